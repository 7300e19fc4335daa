#!/bin/bash
# Times pre-built library variants (build/variants/lib_<tag>.so) on the GPU box.
cp fitburst_b200/libfitburst_b200.so /tmp/lib_orig.so
for f in build/variants/lib_*.so; do
  tag=$(basename $f .so)
  cp $f fitburst_b200/libfitburst_b200.so; touch fitburst_b200/libfitburst_b200.so
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --eval-reps 30 > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/sweep_$tag.json')); print('$tag', 'fits/s %.1f' % d['value'], 'eval ms %.4f' % d['roofline']['ms_per_launch'], d['fit']['nfev'])"
done
cp /tmp/lib_orig.so fitburst_b200/libfitburst_b200.so
