#!/bin/bash
# Round-2 profile captures (run on the GPU box; every capture only after the same command exited 0 plainly).
# The .ncu-rep files are converted to CSV on the box (raw page; source page for the two hot kernels):
# gpurun_out/ is limited to 64 MiB.
CMD="python bench.py --steps 2 --warmup 1 --repeats 1 --workers 1 --no-cpu-baseline --eval-reps 2 --sharded-fits 0 --bursts 2"
O=gpurun_out
$CMD > $O/r2_prof_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/r2_end_launches.csv $CMD > $O/r2_prof_ncu1.log 2>&1
echo "launch list EXIT $?"
ncu --set full --clock-control none --import-source on -k regex:'k_vals3|k_gram3' -s 12 -c 6 -f -o /tmp/r2_hot $CMD > $O/r2_prof_ncu2.log 2>&1
echo "hot EXIT $?"
ncu -i /tmp/r2_hot.ncu-rep --page raw --csv > $O/r2_end_hot_raw.csv
ncu --set full --clock-control none -k regex:'k_tables_group|k_trf_chol|k_hess3|k_gram3_finish' -s 12 -c 8 -f -o /tmp/r2_small $CMD > $O/r2_prof_ncu3.log 2>&1
echo "small EXIT $?"
ncu -i /tmp/r2_small.ncu-rep --page raw --csv > $O/r2_end_small_raw.csv
[ -n "$SKIP_GEN4" ] && { ls -la $O/r2_*csv; exit 0; }
FITBURST_B200_EVAL=4 $CMD > $O/r2_prof_plain4.log 2>&1 || { echo "plain gen4 run failed"; exit 1; }
FITBURST_B200_EVAL=4 ncu --set full --clock-control none --import-source on -k regex:'k_rise3|k_gram4' -s 6 -c 4 -f -o /tmp/r2_gen4 $CMD > $O/r2_prof_ncu4.log 2>&1
echo "gen4 EXIT $?"
ncu -i /tmp/r2_gen4.ncu-rep --page raw --csv > $O/r2_gen4_raw.csv
ls -la /tmp/*.ncu-rep $O/r2_*csv
du -sh $O
