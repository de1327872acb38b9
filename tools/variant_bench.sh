#!/bin/bash
# run the short bench once per prebuilt library variant under gpurun_variants/ (experiment helper)
L=hardware-efficient-mua-compression_b200/libmua_b200.so
cp $L /tmp/lib_keep.so
for v in gpurun_variants/lib_*.so; do
  cp $v $L
  timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0.3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['stages']['encode_ms'], d['stages']['decode_ms'])"
done
cp /tmp/lib_keep.so $L
