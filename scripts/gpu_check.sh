#!/bin/bash
# GPU round trip used during development: parity tests, bench line, launch list.
# usage (under gpurun): bash scripts/gpu_check.sh <tag> [ncu-regex]
tag=${1:-dev}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -4 gpurun_out/pytest_gpu_$tag.log
python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print('fits/s', d['value'], 'e2e', d['e2e']['value'], 'eval ms', d['roofline']['ms_per_launch'], d['fit'])"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eval-reps 2 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eval-reps 2 > gpurun_out/ncu.log 2>&1
if [ -n "$2" ]; then
ncu --set full --clock-control none --import-source on -k regex:"$2" -s 10 -c 2 -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 1 --no-cpu-baseline --eval-reps 2 > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
fi
