#!/bin/bash
# usage: tools/sweep.sh "<label>|<bench args>" ...   -> gpurun_out/sweep_<label>.json + one summary line each
for spec in "$@"; do
  label="${spec%%|*}"; args="${spec#*|}"
  python bench.py --steps 5 --warmup 3 --kernel-breakdown --no-cpu-baseline $args > gpurun_out/sweep_$label.json 2> gpurun_out/sweep_$label.err || tail -3 gpurun_out/sweep_$label.err
  python - "$label" <<'PY'
import json, sys
lab = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/sweep_{lab}.json"))
    kb = {k: round(v * 1e3, 1) for k, v in d["kernel_breakdown_ms"].items()}
    print(lab, "Mvs/s", round(d["value"] / 1e6, 1), "ms/frame", round(d["ms_per_step"], 2), "step_frac", round(d["roofline"]["step_frac"], 3),
          "tiles", d["config"]["tiles_in_pass"], "us:", kb, "e2e", round(d["e2e"]["value"] / 1e6, 1))
except Exception as e:
    print(lab, "FAILED", e)
PY
done
