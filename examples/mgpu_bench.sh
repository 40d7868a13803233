#!/bin/bash
# usage: mgpu_bench.sh "N1 N2 ..." "p2p nccl"   — bench.py under torchrun for each N and exchange, results in gpurun_out/
NS=${1:-2}
EXS=${2:-"p2p nccl"}
export RR_BENCH_RANK_TIMINGS=1
for N in $NS; do for ex in $EXS; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --exchange $ex --verify > gpurun_out/mg_${N}_${ex}.json 2> gpurun_out/mg_${N}_${ex}.err
  echo "== N=$N $ex rc=$?"; grep -v "^W\|^\*\*\*\|^rank\|OMP_NUM\|^$" gpurun_out/mg_${N}_${ex}.err | tail -5
  tail -1 gpurun_out/mg_${N}_${ex}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['parallelism'][:40], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'verify', d.get('verify'))"
done; done
