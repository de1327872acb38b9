#!/bin/bash
# round 2, GPU call D: occupancy probe of the encoders, then encoder variants (S=3 fast, S=5 pair) at three row lengths
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r02d}
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/occ tools/probes/occupancy_probe.cu && /tmp/occ | tee gpurun_out/${TAG}_occ.log
bash tools/gpu_variants_gen.sh $TAG "3 5" "1 10 50"
