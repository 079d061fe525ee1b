#!/bin/bash
# per-CTA tile descriptors + owner bits; conflict-free timing experiment; dist workload at N=1 (for the 2-GPU comparison)
out=gpurun_out; mkdir -p $out; tag=r02i
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_${tag}.log 2>&1; tail -3 $out/pytest_gpu_${tag}.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-bodies --kernel-breakdown"
$B > $out/bench_${tag}.json 2>> $out/bench_${tag}.err
SB_DEBUG_NOCONFLICT=1 $B > $out/bench_${tag}_noconflict.json 2>> $out/bench_${tag}.err
$B --block-threads 256 > $out/bench_${tag}_bt256.json 2>> $out/bench_${tag}.err
$B --block-threads 64 > $out/bench_${tag}_bt64.json 2>> $out/bench_${tag}.err
python tools/ab_plan.py --bodies 1184 --variants "" > $out/ab_${tag}_bodies.log 2>&1
python tools/ab_plan.py --n 160 --frames 5 --variants ";SB_DEBUG_NOCONFLICT=1" > $out/ab_${tag}_4M.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r02i*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "frac %.3f step_frac %.3f" % (r.get("frac", 0), r.get("step_frac", 0)), "rounds", d["config"]["rounds_per_sweep"])
        print("   ", d.get("kernel_breakdown_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 3 $out/ab_r02i_*.log
