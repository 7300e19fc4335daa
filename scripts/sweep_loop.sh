#!/bin/bash
# A/B of the fit-loop drivers at 4 fits in flight (bench.py value / latency), one line per variant.
run() {
  label=$1; shift
  env "$@" python bench.py --steps 24 --warmup 8 --no-cpu-baseline --sharded-fits 0 > gpurun_out/sweep_$label.json 2> gpurun_out/sweep_$label.err
  python - "$label" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/sweep_{sys.argv[1]}.json"))
print(sys.argv[1], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "latency", round(d["single_fit_latency_ms"], 3), "e2e", round(d["e2e"]["value"], 1), flush=True)
PY
}
run host_loop FITBURST_B200_DEVICE_LOOP=0
run ahead2 FITBURST_B200_AHEAD=2
run ahead4 FITBURST_B200_AHEAD=4
run ahead6_first6 FITBURST_B200_AHEAD=6 FITBURST_B200_FIRST=6
run poll_ahead2 FITBURST_B200_POLL=1 FITBURST_B200_AHEAD=2
run poll_ahead4 FITBURST_B200_POLL=1 FITBURST_B200_AHEAD=4
