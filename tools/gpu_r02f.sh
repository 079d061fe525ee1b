#!/bin/bash
# record pipeline A/B: double buffer loaded at round start (+ L2 bulk prefetch) at round width 1 and 2
out=gpurun_out; mkdir -p $out; tag=r02f
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/pytest_gpu_${tag}.log 2>&1; tail -3 $out/pytest_gpu_${tag}.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-bodies --kernel-breakdown"
for w in 1 2; do for l2 in 1 0; do
  SB_ROUND_WIDTH=$w SB_L2_PREFETCH=$l2 $B > $out/bench_${tag}_w${w}_l2${l2}.json 2>> $out/bench_${tag}.err
  SB_ROUND_WIDTH=$w SB_L2_PREFETCH=$l2 python tools/ab_plan.py --bodies 1184 --variants "" > $out/ab_${tag}_bodies_w${w}_l2${l2}.log 2>&1
done; done
SB_ROUND_WIDTH=1 python tools/ab_plan.py --n 160 --frames 5 --variants "" > $out/ab_${tag}_4M_w1.log 2>&1
SB_ROUND_WIDTH=2 python tools/ab_plan.py --n 160 --frames 5 --variants "" > $out/ab_${tag}_4M_w2.log 2>&1
SB_ROUND_WIDTH=1 $B --fast-math > $out/bench_${tag}_w1_fast.json 2>> $out/bench_${tag}.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r02f*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "frac %.3f step_frac %.3f" % (r.get("frac", 0), r.get("step_frac", 0)), "rounds", d["config"]["rounds_per_sweep"])
        print("   ", d.get("kernel_breakdown_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 3 $out/ab_r02f_*.log
