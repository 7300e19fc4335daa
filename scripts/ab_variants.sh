#!/bin/bash
# evaluation time of library build variants (scripts/build_variant.sh): scripts/ab_variants.sh lib1.so lib2.so ...
O=gpurun_out
CMD="python bench.py --steps 8 --warmup 4 --repeats 1 --workers 1 --no-cpu-baseline --eval-reps 40 --sharded-fits 0 --bursts 2"
for lib in default "$@"; do
  tag=$(basename $lib .so)
  if [ "$lib" = "default" ]; then timeout 300 $CMD > $O/abv_$tag.json 2> $O/abv_$tag.err; else FITBURST_B200_LIB=$PWD/$lib timeout 300 $CMD > $O/abv_$tag.json 2> $O/abv_$tag.err; fi
  python -c "import json;d=json.load(open('$O/abv_$tag.json'));print('$tag eval_ms',round(d['roofline']['ms_per_launch'],4),'latency',round(d['single_fit_latency_ms'],3),'ms/step',round(d['ms_per_step'],3))" || tail -3 $O/abv_$tag.err
done
