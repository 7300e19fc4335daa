#!/bin/bash
# are the occasional 0.2-0.8 s stalls of a batch garbage-collector passes? (gc.DEBUG_STATS on stderr)
O=gpurun_out
python - <<'PY' > $O/gc_probe.json 2> $O/gc_probe.err
import gc, sys, runpy
gc.set_debug(gc.DEBUG_STATS)
sys.argv = ["bench.py", "--steps", "24", "--warmup", "8", "--repeats", "7", "--no-cpu-baseline", "--sharded-fits", "0"]
runpy.run_path("bench.py", run_name="__main__")
PY
python -c "import json;d=json.load(open('$O/gc_probe.json'));print('repeats',[round(x,2) for x in d['ms_per_step_repeats']])"
grep -E "gc: done|gc: collecting generation" $O/gc_probe.err | awk '/collecting generation 2/{g=1} /done/{ if (g) print; g=0 }' | sort -t, -k2 | tail -5
grep -c "collecting generation 2" $O/gc_probe.err
grep -E "gc: done.* [0-9]+\.[0-9]+s elapsed" $O/gc_probe.err | awk '{for(i=1;i<=NF;i++) if ($i ~ /s$/ && $(i+1)=="elapsed") print $i}' | sort -n | tail -5
