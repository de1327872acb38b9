#!/bin/bash
# One gpurun call that refreshes the round's evidence: GPU parity suite, smoke, bench line, ncu launch list and one
# ncu --set full capture of the step's kernels.  Usage: gpurun --timeout 1500 -- 'bash tools/evidence.sh rNN'
R=${1:-r01}
O=gpurun_out
mkdir -p $O
t0=$(date +%s)
python -m pytest tests -m gpu -x -q > $O/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-t0 )) s" | tee $O/${R}_status.txt
tail -3 $O/${R}_pytest_gpu.log
python __graft_entry__.py smoke > $O/${R}_smoke.log 2>&1; echo "smoke rc=$? $(( $(date +%s)-t0 )) s" | tee -a $O/${R}_status.txt
python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err; echo "bench rc=$? $(( $(date +%s)-t0 )) s" | tee -a $O/${R}_status.txt
cat $O/${R}_bench_n1.json
CMD="python bench.py --steps 5 --warmup 3 --no-e2e --cpu-seconds 0.5"
$CMD > $O/${R}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv $CMD > $O/${R}_ncu_launches.log 2>&1
echo "ncu launches rc=$? $(( $(date +%s)-t0 )) s" | tee -a $O/${R}_status.txt
$CMD > $O/${R}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_encode|k_decode|k_calibrate' -s 9 -c 3 -f -o $O/${R}_prof $CMD > $O/${R}_ncu_full.log 2>&1
echo "ncu full rc=$? $(( $(date +%s)-t0 )) s" | tee -a $O/${R}_status.txt
