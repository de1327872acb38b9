#!/bin/bash
# time calibrate (9 history lengths, 100k channels) for S in {3,5,7,10} x BP {1,10} once per prebuilt library variant
L=hardware-efficient-mua-compression_b200/libmua_b200.so
cp $L /tmp/lib_keep.so
for v in gpurun_variants/lib_*.so; do
  cp $v $L
  echo "== $v"
  for S in 3 5 7 10; do timeout 100 python tools/cal_time.py $S 1 2>/dev/null; done
  timeout 100 python tools/cal_time.py 10 10 2>/dev/null
done
cp /tmp/lib_keep.so $L
