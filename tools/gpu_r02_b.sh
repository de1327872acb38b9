#!/bin/bash
# round 2, GPU call B: whole GPU suite, cfg4 bench lines with the dense pair encoder
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --timeout 1500 ) > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
for S in 5 7 9; do
python bench.py --workload cfg4 --alphabet $S --bp 1 --steps 5 --warmup 3 --no-e2e > gpurun_out/r02b_bench_cfg4_s$S.json 2> gpurun_out/r02b_bench_cfg4_s$S.err
done
tail -3 gpurun_out/r02b_pytest.log
