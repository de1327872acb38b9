#!/bin/bash
# usage: bash tools/gpu_ncu_stage.sh TAG [lib.so]  -- ONE ncu --set full capture of the encode and decode kernels of tools/stage_time.py (cfg5 shard), optionally with a variant library
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=$1; V=$2; RX=${3:-"k_encode|k_decode"}
L=hardware-efficient-mua-compression_b200/libmua_b200.so
mkdir -p gpurun_out
if [ -n "$V" ]; then cp $L /tmp/lib_keep.so; cp $V $L; fi
python tools/stage_time.py $TAG > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$RX" --launch-skip 4 --launch-count 2 \
    -o gpurun_out/${TAG}_prof -f python tools/stage_time.py $TAG > gpurun_out/${TAG}_ncu.log 2>&1
if [ -n "$V" ]; then cp /tmp/lib_keep.so $L; fi
tail -1 gpurun_out/${TAG}_plain.log
