#!/usr/bin/env bash
# multi-GPU bench lines (run under gpurun --gpus N): bash tools/multi_bench.sh N tag [configs...]
set -u
N=$1; T=$2; shift 2
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
for C in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --config $C \
     > gpurun_out/${T}_bench_${C}_n${N}.json 2> gpurun_out/${T}_bench_${C}_n${N}.err; echo "$C n=$N rc=$?"
  cut -c1-400 gpurun_out/${T}_bench_${C}_n${N}.json
  tail -3 gpurun_out/${T}_bench_${C}_n${N}.err
done
