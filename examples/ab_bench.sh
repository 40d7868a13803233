#!/bin/bash
# A/B on one box: the same bench line for a list of configurations "name[,ENV=value...][,lib=ab_libs/x.so]".
#   bash examples/ab_bench.sh 200 default nopersist,RR_SHADOW_PERSIST=0,RR_SETUP_PERSIST=0 base,lib=ab_libs/librr_base.so
# Without configurations: the in-tree product, then every ab_libs/*.so.
cd "$(dirname "$0")/.."
steps=${1:-200}; shift
cfgs=("$@")
if [ ${#cfgs[@]} -eq 0 ]; then cfgs=(default); for l in ab_libs/*.so; do [ -f "$l" ] && cfgs+=("$(basename $l .so),lib=$l"); done; fi
for cfg in "${cfgs[@]}"; do
  IFS=, read -ra parts <<< "$cfg"
  name=${parts[0]}; envs=()
  for p in "${parts[@]:1}"; do
    case "$p" in lib=*) envs+=("RR_LIB=$PWD/${p#lib=}");; *) envs+=("$p");; esac
  done
  env "${envs[@]}" python bench.py --steps $steps --warmup 10 --no-cpu-baseline --no-opencl-reference $AB_FLAGS 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d.get('stages_ms'))"
done
