#!/bin/bash
# cfg4 bench lines (stages run back to back, cold: the honest short-row numbers) once per prebuilt library variant under gpurun_variants/
# usage: cfg4_variants.sh "S BP" ["S BP" ...]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
L=hardware-efficient-mua-compression_b200/libmua_b200.so
cp $L /tmp/lib_keep.so
for v in gpurun_variants/lib_*.so; do
  cp $v $L
  for cell in "$@"; do
    S=${cell% *}; BP=${cell#* }
    python bench.py --workload cfg4 --alphabet $S --bp $BP --steps 10 --warmup 3 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['stages']
print('$v', 'S=$S BP=$BP', 'step %.4f ms' % d['ms_per_step'], {k: round(st[k],4) for k in ('calibrate_ms','encode_ms','decode_ms','calibrate_frac','encode_frac','decode_frac')})"
  done
done
cp /tmp/lib_keep.so $L
