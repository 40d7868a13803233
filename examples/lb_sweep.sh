#!/bin/bash
# Occupancy sweep of the three heavy kernels: builds librr_b200 variants with different __launch_bounds__ minimum-CTA knobs
# (-DRR_LB_SHADOW_SETUP / -DRR_LB_SHADE / -DRR_LB_SETUP_MAIN, rr_kernels.cuh) and benches each through RR_LIB.
#   build here (no GPU needed):   bash examples/lb_sweep.sh build "8,-,4 8,8,4 8,-,5 10,-,4"
#   run on the GPU box:           bash examples/lb_sweep.sh run
# A triple is shadow_setup,shade,setup_main; "-" keeps the compiler's own choice (k_shade) / the default.
set -e
cd "$(dirname "$0")/.."
mkdir -p ab_libs
if [ "$1" = "build" ]; then
  for v in $2; do
    IFS=, read s k m <<< "$v"
    D=""
    [ "$s" != "-" ] && D="$D -DRR_LB_SHADOW_SETUP=$s"
    [ "$k" != "-" ] && D="$D -DRR_LB_SHADE=$k"
    [ "$m" != "-" ] && D="$D -DRR_LB_SETUP_MAIN=$m"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -std=c++17 -Xcompiler -fPIC -shared $D \
         -o "ab_libs/librr_${s}_${k}_${m}.so" openclrenderer_b200/csrc/rr_api.cu &
  done
  wait
  ls -la ab_libs
else
  for lib in "" ab_libs/*.so; do
    RR_LIB=${lib:+$PWD/$lib} python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-opencl-reference 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('${lib:-default}', d['ms_per_step'], d['stages_ms'])"
  done
fi
