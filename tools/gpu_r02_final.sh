#!/bin/bash
# round-2 final evidence in one gpurun call: suite + smoke + headline bench + ncu of the step kernels (tools/evidence.sh), cfg4 bench
# lines, the cfg4 sweep, and one ncu --set full capture of each lane-per-channel kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
R=${1:-r02z}
bash tools/evidence.sh $R
for cell in "3 50" "3 10" "5 50" "9 50" "5 10" "5 1" "9 1"; do
  set -- $cell
  python bench.py --workload cfg4 --alphabet $1 --bp $2 --steps 5 --warmup 3 --no-e2e > gpurun_out/${R}_cfg4_s$1_bp$2.json 2> gpurun_out/${R}_cfg4_s$1_bp$2.err
done
python tools/sweep_bench.py > gpurun_out/${R}_sweep.log 2>&1 && cp gpurun_out/sweep.json gpurun_out/${R}_sweep.json
bash tools/gpu_ncu_gen.sh ${R}_rows3 3 50 "k_encode|k_decode"
bash tools/gpu_ncu_gen.sh ${R}_rows5 5 50 "k_encode|k_decode"
bash tools/gpu_ncu_cal.sh ${R}_cal3 3 50
bash tools/gpu_ncu_cal.sh ${R}_cal9 9 50
ls -la gpurun_out | grep $R | wc -l
