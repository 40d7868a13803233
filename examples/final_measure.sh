#!/bin/bash
# End-of-round measurement on one B200: the default bench line, the launch list of the same command, and an ncu --set full capture of
# every kernel of a steady-state frame (exported to CSV on the box; the reports themselves are too large to bring back).
#   bash examples/final_measure.sh <tag>     -> gpurun_out/<tag>/
tag=${1:-final}; out=gpurun_out/$tag; mkdir -p $out
python bench.py > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-opencl-reference > $out/launches_run.log 2>&1
ncu --set full --clock-control none -k regex:'^(void )?k_' -s 64 -c 10 -o $out/frame python examples/rank_frame.py --frames 5 > $out/ncu_frame.log 2>&1
ncu -i $out/frame.ncu-rep --page raw --csv > $out/frame_raw.csv 2>/dev/null; ls -la $out/frame.ncu-rep; rm -f $out/frame.ncu-rep
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; tail -c 400 $out/bench_reference.json; echo
