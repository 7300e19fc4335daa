#!/bin/bash
# A/B of one environment switch on the evaluation time: scripts/ab_env.sh NAME (runs NAME=1 / default and NAME=0)
O=gpurun_out
NAME=$1
CMD="python bench.py --steps 8 --warmup 4 --repeats 1 --workers 1 --no-cpu-baseline --eval-reps 40 --sharded-fits 0 --bursts 2"
for mode in default 0; do
  if [ "$mode" = "default" ]; then timeout 300 $CMD > $O/ab_env_$mode.json 2> $O/ab_env_$mode.err; else env $NAME=0 timeout 300 $CMD > $O/ab_env_$mode.json 2> $O/ab_env_$mode.err; fi
  python -c "import json;d=json.load(open('$O/ab_env_$mode.json'));print('$NAME=$mode eval_ms',round(d['roofline']['ms_per_launch'],4),'latency',round(d['single_fit_latency_ms'],3),'ms/step',round(d['ms_per_step'],3))"
done
