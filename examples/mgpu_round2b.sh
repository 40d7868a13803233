#!/bin/bash
# usage: mgpu_round2b.sh "c3:8 c3:4 c5:8 c4:8"  — bench.py (peer-memory exchange, composite verified against the single-GPU frame)
# for each workload:N, one JSON line per run in gpurun_out/r2b_mgpu/
mkdir -p gpurun_out/r2b_mgpu
export RR_BENCH_RANK_TIMINGS=1
for wn in $1; do
  w=${wn%%:*}; N=${wn##*:}
  out=gpurun_out/r2b_mgpu/bench_${w}_${N}gpu_p2p
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-opencl-reference > $out.json 2> $out.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 30 --warmup 5 --exchange p2p > $out.json 2> $out.err
  fi
  echo "== $w N=$N rc=$?"; grep "^rank" $out.err | cut -c1-260 | head -8
  tail -1 $out.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['workload'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'verify', d.get('verify',{}).get('pixels_differing_from_single_gpu_frame'), 'nvlink', d.get('nvlink',{}).get('face_push_bytes_per_frame'), 'stages', d.get('stages_ms'))"
done
