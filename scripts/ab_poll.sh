#!/bin/bash
# spread of the resident batch time under the three host-wait modes (7 repeats each)
O=gpurun_out
for mode in 2 1 0; do
  FITBURST_B200_POLL=$mode python bench.py --steps 24 --warmup 8 --repeats 7 --no-cpu-baseline --sharded-fits 0 > $O/ab_poll_$mode.json 2> $O/ab_poll_$mode.err
  python -c "import json;d=json.load(open('$O/ab_poll_$mode.json'));print('POLL=$mode value',round(d['value'],1),'repeats',[round(x,2) for x in d['ms_per_step_repeats']],'e2e',round(d['e2e']['value'],1),[round(x,2) for x in d['e2e']['ms_per_step_repeats']],'latency',round(d['single_fit_latency_ms'],3))"
done
