#!/bin/bash
# usage: mg2.sh N
N=$1
export RR_BENCH_RANK_TIMINGS=1
for ex in p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --exchange $ex --verify > gpurun_out/mg_${N}_${ex}.json 2> gpurun_out/mg_${N}_${ex}.err
  echo "== $ex rc=$?"; grep -v "^W\|^\*\*\*" gpurun_out/mg_${N}_${ex}.err | tail -12; cat gpurun_out/mg_${N}_${ex}.json | cut -c1-1500
done
