#!/bin/bash
# per-variant kernel times from an ncu launch list (measurement experiments; results may be wrong)
cp fitburst_b200/libfitburst_b200.so /tmp/lib_orig.so
for f in build/variants/lib_*.so; do
  tag=$(basename $f .so)
  cp $f fitburst_b200/libfitburst_b200.so; touch fitburst_b200/libfitburst_b200.so
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_gram3$|^k_vals3$" -c 12 --csv --log-file gpurun_out/sw_$tag.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --eval-reps 2 > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/sw_$tag.csv 2>/dev/null | grep -E "k_gram3|k_vals3" | sed "s/^/$tag /" | cut -c1-150
done
cp /tmp/lib_orig.so fitburst_b200/libfitburst_b200.so
