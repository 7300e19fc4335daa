#!/bin/bash
# GPU round trip used during development: parity tests, bench line, launch list, ncu capture.
# usage (under gpurun): bash scripts/gpu_check.sh <tag> [ncu-regex]
tag=${1:-dev}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -4 gpurun_out/pytest_gpu_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print('fits/s', d['value'], 'e2e', d['e2e']['value'], 'eval ms', d['roofline']['ms_per_launch'], d['fit'], d['cpu_baseline'])"
# profiler passes: one fit at a time (ncu serialises kernels anyway), only after the same command ran clean
NCU_CMD="python bench.py --steps 2 --warmup 1 --repeats 1 --workers 1 --no-cpu-baseline --eval-reps 2"
$NCU_CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $NCU_CMD > gpurun_out/ncu.log 2>&1
if [ -n "$2" ]; then
ncu --set full --clock-control none --import-source on -k regex:"$2" -s 10 -c 6 -o gpurun_out/prof_$tag $NCU_CMD > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
fi
