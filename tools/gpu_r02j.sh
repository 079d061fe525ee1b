#!/bin/bash
# CTA width / tile size variants on the 1 M block
out=gpurun_out; mkdir -p $out; tag=r02j
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-bodies"
$B --block-threads 160 > $out/bench_${tag}_bt160.json 2>> $out/bench_${tag}.err
$B --block-threads 192 > $out/bench_${tag}_bt192.json 2>> $out/bench_${tag}.err
$B --block-threads 192 --tile-cap 2197 > $out/bench_${tag}_bt192_c2197.json 2>> $out/bench_${tag}.err
$B --block-threads 256 --tile-cap 2197 > $out/bench_${tag}_bt256_c2197.json 2>> $out/bench_${tag}.err
$B --block-threads 160 --tile-cap 2197 > $out/bench_${tag}_bt160_c2197.json 2>> $out/bench_${tag}.err
$B --block-threads 128 --tile-cap 2197 > $out/bench_${tag}_bt128_c2197.json 2>> $out/bench_${tag}.err
SB_MERGE_PCT=115 $B --block-threads 160 > $out/bench_${tag}_bt160_m115.json 2>> $out/bench_${tag}.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r02j*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "step_frac %.3f" % (r.get("step_frac", 0)), "rounds", d["config"]["rounds_per_sweep"], d["config"]["tiles_in_pass"])
    except Exception as e:
        print(f, "ERR", e)
PY
