#!/bin/bash
# refresh of the cfg4 evidence after the last calibrate change: cfg4 bench lines, the sweep, ncu of k_calibrate_rows
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
R=${1:-r02y}
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${R}_pytest_gpu.log
for cell in "3 50" "3 10" "5 50" "9 50" "5 10" "9 10" "5 1" "9 1"; do
  S=${cell% *}; BP=${cell#* }
  python bench.py --workload cfg4 --alphabet $S --bp $BP --steps 10 --warmup 3 --no-e2e > gpurun_out/${R}_cfg4_s${S}_bp${BP}.json 2> gpurun_out/${R}_cfg4_s${S}_bp${BP}.err
done
python tools/sweep_bench.py > gpurun_out/${R}_sweep.log 2>&1 && cp gpurun_out/sweep.json gpurun_out/${R}_sweep.json
bash tools/gpu_ncu_cal.sh ${R}_cal3 3 50
bash tools/gpu_ncu_cal.sh ${R}_cal9 9 50
