#!/bin/bash
# usage: bash tools/gpu_suite.sh TAG [cfg4 alphabets...]   -- whole GPU suite + cfg4 bench lines (device-timed only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=$1; shift
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --timeout 1500 ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
for S in "$@"; do
python bench.py --workload cfg4 --alphabet $S --bp 1 --steps 5 --warmup 3 --no-e2e > gpurun_out/${TAG}_cfg4_s$S.json 2> gpurun_out/${TAG}_cfg4_s$S.err
done
grep -E "passed|failed|FAILED|rc=" gpurun_out/${TAG}_pytest.log | tail -8
