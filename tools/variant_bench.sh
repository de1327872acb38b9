#!/bin/bash
# time encode/decode once per prebuilt library variant under gpurun_variants/ (experiment helper; no parity check)
L=hardware-efficient-mua-compression_b200/libmua_b200.so
cp $L /tmp/lib_keep.so
for v in gpurun_variants/lib_*.so; do
  cp $v $L
  timeout 200 python tools/stage_time.py $v 2>/dev/null
done
cp /tmp/lib_keep.so $L
