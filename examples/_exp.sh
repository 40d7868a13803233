mkdir -p gpurun_out/r3k
cd gpurun_out/r3k
for k in k_shade k_shadow_setup k_setup_main; do
  ncu --set full --clock-control none --import-source on -k regex:"${k}\$|${k}<" -s 2 -c 1 -o prof_$k python ../../examples/rank_frame.py --frames 3 > ncu_$k.log 2>&1
  ncu -i prof_$k.ncu-rep --page source --csv --print-source cuda,sass > source_$k.csv 2> /dev/null
  ncu -i prof_$k.ncu-rep --page raw --csv > raw_$k.csv 2> /dev/null
  ls -la prof_$k.ncu-rep source_$k.csv
  rm -f prof_$k.ncu-rep
done
