#!/bin/bash
# time the general-codebook path (tools/gen_time.py) for S in {4,5,7,9,10} once per prebuilt library variant + the default
L=hardware-efficient-mua-compression_b200/libmua_b200.so
cp $L /tmp/lib_keep.so
for v in gpurun_variants/lib_*.so /tmp/lib_keep.so; do
  cp $v $L 2>/dev/null
  echo "== $v"
  for S in 5 7 9 10; do timeout 100 python tools/gen_time.py $S 1 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in j.items() if k in ('S','encode_ms','decode_ms','bits_per_symbol','parity_ok')})"; done
  timeout 100 python tools/gen_time.py 9 50 2>/dev/null | tail -1 | cut -c1-200
done
cp /tmp/lib_keep.so $L
