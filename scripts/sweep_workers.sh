for w in 4 5 6 8; do
  python bench.py --workers $w --no-cpu-baseline --sharded-fits 0 --repeats 3 > gpurun_out/w_$w.json 2> gpurun_out/w_$w.err
  python -c "import json;d=json.load(open('gpurun_out/w_$w.json'));print($w,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['ms_per_step_repeats'],d['e2e']['ms_per_step_repeats'])"
done
