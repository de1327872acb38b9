#!/bin/bash
# usage: bash tools/gpu_variants_gen.sh TAG "S list" "BP list"  -- tools/gen_time.py once per library variant under gpurun_variants/ + the default
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=$1; SL=${2:-"5 9"}; BL=${3:-"1"}
L=hardware-efficient-mua-compression_b200/libmua_b200.so
cp $L /tmp/lib_keep.so
mkdir -p gpurun_out; : > gpurun_out/${TAG}_variants.log
for v in /tmp/lib_keep.so gpurun_variants/lib_*.so; do
  cp $v $L 2>/dev/null
  echo "== $v" >> gpurun_out/${TAG}_variants.log
  for S in $SL; do for BP in $BL; do
    timeout 120 python tools/gen_time.py $S $BP 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in j.items() if k in ('S','BP','encode_ms','decode_ms','bits_per_symbol','parity_ok')})" >> gpurun_out/${TAG}_variants.log 2>&1
  done; done
done
cp /tmp/lib_keep.so $L
cat gpurun_out/${TAG}_variants.log
