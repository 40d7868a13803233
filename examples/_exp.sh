(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "e2e or alternative") 2>&1 | tail -3
for g in 4 6; do RR_TILE_GRID=$g python examples/tile_probe.py 2>&1 | tail -1; done
python bench.py --no-cpu-baseline --no-opencl-reference 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], json.dumps(d['e2e'])[:700])"
