#!/bin/bash
# 1-GPU check of the round's state: whole GPU test-suite (timed), the bench line
out=gpurun_out; mkdir -p $out; tag=r02o
timeout 1700 python -m pytest tests -m gpu -x -q --durations=8 > $out/pytest_gpu_${tag}.log 2>&1; tail -14 $out/pytest_gpu_${tag}.log
python bench.py --steps 20 --warmup 3 --kernel-breakdown > $out/bench_${tag}.json 2> $out/bench_${tag}.err || tail -5 $out/bench_${tag}.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02o.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("ms/step %.3f value %.4g e2e %.4g (%.3f ms) frac %.3f step_frac %.3f launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], r["frac"], r["step_frac"], d["gpu_launches"] // d["steps"]))
print("bodies", d.get("bodies"))
print("cpu", d.get("cpu_baseline"))
print(d.get("kernel_breakdown_ms"), r.get("traffic_source"))
PY
