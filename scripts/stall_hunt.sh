#!/bin/bash
O=gpurun_out
run() { label=$1; shift; env "$@" python bench.py --steps 24 --warmup 8 --repeats 9 --no-cpu-baseline --sharded-fits 0 > $O/stall_$label.json 2> $O/stall_$label.err; python -c "import json;d=json.load(open('$O/stall_$label.json'));print('$label',[round(x,2) for x in d['ms_per_step_repeats']],[round(x,2) for x in d['e2e']['ms_per_step_repeats']])"; }
run sampler_on A=1
run sampler_off BENCH_NO_SAMPLER=1
run sampler_off2 BENCH_NO_SAMPLER=1
run sampler_on2 A=1
