#!/bin/bash
# usage: bash tools/gpu_ncu_cal.sh TAG S BP  -- plain run of tools/cal_time.py, then ONE ncu --set full capture of its calibrate kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=$1; S=$2; BP=$3
mkdir -p gpurun_out
python tools/cal_time.py $S $BP > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_calibrate" --launch-skip 2 --launch-count 1 \
    -o gpurun_out/${TAG}_prof -f python tools/cal_time.py $S $BP > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_plain.log
