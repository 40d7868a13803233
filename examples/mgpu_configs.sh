#!/bin/bash
# usage: mgpu_configs.sh N "c4 c5"  — the multi-GPU configurations of BASELINE.json through the peer-memory exchange, with --verify
N=${1:-4}
for wl in ${2:-"c4 c5"}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --workload $wl --exchange p2p --verify > gpurun_out/mg_${wl}_${N}.json 2> gpurun_out/mg_${wl}_${N}.err
  echo "== $wl N=$N rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/mg_${wl}_${N}.err | tail -6 | cut -c1-400
  tail -1 gpurun_out/mg_${wl}_${N}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['workload'], d['config']['parallelism'][:30], 'ms', d['ms_per_step'], 'Mtri/s', d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('host_frame_pixels_differing_from_device_composite'), 'verify', d.get('verify'))"
done
