#!/bin/bash
# evaluation time against the window length (per-channel overhead vs per-sample work of the fast kernels)
O=gpurun_out
for nt in 256 512 1024 2048; do
  timeout 300 python bench.py --ntime $nt --steps 4 --warmup 3 --repeats 1 --workers 1 --no-cpu-baseline --eval-reps 40 --sharded-fits 0 --bursts 2 > $O/scan_nt_$nt.json 2> $O/scan_nt_$nt.err
  python -c "import json;d=json.load(open('$O/scan_nt_$nt.json'));print('ntime $nt eval_ms',round(d['roofline']['ms_per_launch'],4))" || tail -3 $O/scan_nt_$nt.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/scan_nt_2048_launches.csv python bench.py --ntime 2048 --steps 2 --warmup 1 --repeats 1 --workers 1 --no-cpu-baseline --eval-reps 2 --sharded-fits 0 --bursts 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/scan_nt_512_launches.csv python bench.py --ntime 512 --steps 2 --warmup 1 --repeats 1 --workers 1 --no-cpu-baseline --eval-reps 2 --sharded-fits 0 --bursts 2 > /dev/null 2>&1
