#!/bin/bash
# GPU call: full GPU suite with the default library, then the headline stage times once per library variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r02n}
mkdir -p gpurun_out
( python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -5 ) | tee gpurun_out/${TAG}_pytest.log
echo "== default"; timeout 300 python tools/stage_time.py default 2>&1 | tail -2
bash tools/variant_bench.sh 2>&1 | tee gpurun_out/${TAG}_variants.log
