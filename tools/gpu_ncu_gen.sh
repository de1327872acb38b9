#!/bin/bash
# usage: bash tools/gpu_ncu_gen.sh TAG S BP  -- plain run of tools/gen_time.py, then ONE ncu --set full capture of its encode and decode kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=$1; S=$2; BP=$3; RX=${4:-"k_encode|k_decode"}
mkdir -p gpurun_out
python tools/gen_time.py $S $BP > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"$RX" --launch-skip 2 --launch-count 2 \
    -o gpurun_out/${TAG}_prof -f python tools/gen_time.py $S $BP > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_plain.log
