#!/bin/bash
# A/B of the row fetch in k_gram3: cp.async.bulk + mbarrier (default) against per-lane cp.async (FITBURST_B200_BULK=0)
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --repeats 1 --workers 1 --no-cpu-baseline --eval-reps 20 --sharded-fits 0 --bursts 2"
for mode in 1 0; do
  FITBURST_B200_BULK=$mode timeout 300 $CMD > $O/ab_bulk_$mode.json 2> $O/ab_bulk_$mode.err || echo "plain run $mode failed"
  python -c "import json;d=json.load(open('$O/ab_bulk_$mode.json'));print('BULK=$mode eval_ms',round(d['roofline']['ms_per_launch'],4),'latency',round(d['single_fit_latency_ms'],3))"
  FITBURST_B200_BULK=$mode timeout 600 ncu --set full --clock-control none -k regex:'k_gram3' -s 6 -c 2 -f -o /tmp/ab_bulk_$mode $CMD > $O/ab_bulk_ncu_$mode.log 2>&1
  ncu -i /tmp/ab_bulk_$mode.ncu-rep --page raw --csv > $O/ab_bulk_raw_$mode.csv
done
