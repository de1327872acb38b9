#!/bin/bash
# round 2, GPU call A: whole GPU suite (no -x: see every failure), smoke, one short bench line per workload
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r02a_gpu.txt 2>&1
( time python -m pytest tests -m gpu -q -x --timeout 1500 ) > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
python bench.py --workload cfg4 --alphabet 5 --bp 1 --steps 5 --warmup 3 --no-e2e > gpurun_out/r02a_bench_cfg4_s5.json 2> gpurun_out/r02a_bench_cfg4_s5.err
tail -3 gpurun_out/r02a_pytest.log
