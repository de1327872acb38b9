#!/bin/bash
# usage: bash tools/gpu_multi.sh TAG N [extra bench args]  -- bench.py on N GPUs (torchrun), p2p report and NCCL report, device-timed only
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=$1; N=$2; shift; shift
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 10 --warmup 3 --no-e2e --report p2p "$@" > gpurun_out/${TAG}_n${N}_p2p.json 2> gpurun_out/${TAG}_n${N}_p2p.err
echo "p2p rc=$?"
run --steps 10 --warmup 3 --no-e2e --report nccl "$@" > gpurun_out/${TAG}_n${N}_nccl.json 2> gpurun_out/${TAG}_n${N}_nccl.err
echo "nccl rc=$?"
tail -c 600 gpurun_out/${TAG}_n${N}_p2p.err
