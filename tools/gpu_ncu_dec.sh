#!/bin/bash
# usage: bash tools/gpu_ncu_dec.sh TAG S BP [kernel-regex]  -- plain run of tools/gen_time.py, then ONE ncu --set full capture of one launch of the matching kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=$1; S=$2; BP=$3; K=${4:-k_decode}
mkdir -p gpurun_out
python tools/gen_time.py $S $BP > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip 2 --launch-count 1 \
    -o gpurun_out/${TAG}_prof -f python tools/gen_time.py $S $BP > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_plain.log
