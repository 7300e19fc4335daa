#!/bin/bash
# source-page captures (SASS rows with stall samples) of the two hot kernels, one launch each
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --repeats 1 --workers 1 --no-cpu-baseline --eval-reps 2 --sharded-fits 0 --bursts 2"
$CMD > $O/prof_src_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
for k in ${KERNELS:-k_vals3 k_gram3}; do
  ncu --set full --clock-control none --import-source on -k regex:"^$k\$|$k<" -s 8 -c 1 -f -o /tmp/src_$k $CMD > $O/prof_src_$k.log 2>&1
  ncu -i /tmp/src_$k.ncu-rep --page source --csv > $O/src_$k.csv 2>/dev/null
  ncu -i /tmp/src_$k.ncu-rep --page raw --csv > $O/src_raw_$k.csv 2>/dev/null
  ls -la $O/src_$k.csv
done
cp fitburst_b200/libfitburst_b200.so $O/lib_profiled.so
