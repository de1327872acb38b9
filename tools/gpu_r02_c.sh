#!/bin/bash
# round 2, GPU call C: whole GPU suite, cfg4 sweep (all stages, 3 bin periods x 4 alphabets), headline bench (device-timed only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r02c}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x --timeout 1500 ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
python tools/sweep_bench.py > gpurun_out/${TAG}_sweep.log 2>&1 && cp gpurun_out/sweep.json gpurun_out/${TAG}_sweep.json
python tools/show_sweep.py gpurun_out/${TAG}_sweep.json | tail -14
python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0.5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - $TAG <<'PY'
import json,sys
d=json.load(open("gpurun_out/%s_bench.json" % (sys.argv[1] if len(sys.argv)>1 else "r02c")))
print({k: d["stages"][k] for k in ("calibrate_ms","encode_ms","decode_ms","encode_frac","decode_frac")}, d["ms_per_step"], d["lossless"])
PY
