#!/bin/bash
# time the short-row stages once per prebuilt library variant under gpurun_variants/ (experiment helper; parity flag only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
L=hardware-efficient-mua-compression_b200/libmua_b200.so
cp $L /tmp/lib_keep.so
for v in gpurun_variants/lib_*.so; do
  cp $v $L
  echo "== $v"
  ROWS_BPS=${ROWS_BPS:-50,10} timeout 200 python tools/rows_time.py "$@" 2>&1 | tail -4
done
cp /tmp/lib_keep.so $L
