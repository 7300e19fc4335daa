#!/bin/bash
# config-5 shard (scintillation, 8 components, 262144 samples) on ONE GPU: launch list of a short
# fit, then one --set full capture of the evaluation kernel (source page -> CSV).
O=gpurun_out
NF=${NF:-128}
CMD="python scripts/run_configs_multi.py 5 --nfreq5 $NF"
$CMD > $O/c5_plain.jsonl 2> $O/c5_plain.err || { echo "plain run failed"; tail -5 $O/c5_plain.err; exit 1; }
cat $O/c5_plain.jsonl | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/c5_launches.csv $CMD > $O/c5_ncu_list.log 2>&1
K=${KERNEL:-k_eval}
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 3 -c 1 -f -o /tmp/c5_eval $CMD > $O/c5_ncu_full.log 2>&1
ncu -i /tmp/c5_eval.ncu-rep --page source --csv > $O/c5_src.csv 2>/dev/null
ncu -i /tmp/c5_eval.ncu-rep --page raw --csv > $O/c5_raw.csv 2>/dev/null
ls -la $O/c5_src.csv $O/c5_raw.csv
cp fitburst_b200/libfitburst_b200.so $O/lib_profiled.so
