#!/bin/bash
# last refresh: suite, smoke, headline bench, cfg4 bench lines, sweep, ncu of the pair-table row encoder
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
R=${1:-r02x}
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${R}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
for cell in "3 50" "3 10" "5 50" "9 50" "5 10" "9 10" "5 1" "9 1" "3 1"; do
  S=${cell% *}; BP=${cell#* }
  python bench.py --workload cfg4 --alphabet $S --bp $BP --steps 10 --warmup 3 --no-e2e > gpurun_out/${R}_cfg4_s${S}_bp${BP}.json 2> gpurun_out/${R}_cfg4_s${S}_bp${BP}.err
done
python tools/sweep_bench.py > gpurun_out/${R}_sweep.log 2>&1 && cp gpurun_out/sweep.json gpurun_out/${R}_sweep.json
bash tools/gpu_ncu_gen.sh ${R}_rows5 5 50 "k_encode|k_decode"
