#!/bin/bash
# 1-GPU session: GPU test-suite, the bench line, A/B of the round-2 switches (bi-tets, fused launches, snake order, PDL)
out=gpurun_out; mkdir -p $out; tag=${1:-r02e}
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_${tag}.log 2>&1; tail -4 $out/pytest_gpu_${tag}.log
python bench.py --steps 20 --warmup 3 --kernel-breakdown > $out/bench_${tag}.json 2> $out/bench_${tag}.err || tail -5 $out/bench_${tag}.err
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-bodies --kernel-breakdown"
SB_BITETS=0 $B > $out/bench_${tag}_nobitets.json 2>> $out/bench_${tag}.err
$B --no-fuse > $out/bench_${tag}_nofuse.json 2>> $out/bench_${tag}.err
$B --no-snake > $out/bench_${tag}_nosnake.json 2>> $out/bench_${tag}.err
$B --no-pdl > $out/bench_${tag}_nopdl.json 2>> $out/bench_${tag}.err
$B --fast-math > $out/bench_${tag}_fast.json 2>> $out/bench_${tag}.err
SB_BITETS=0 $B --no-fuse --no-snake > $out/bench_${tag}_r02a_like.json 2>> $out/bench_${tag}.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r02e*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "e2e", (d.get("e2e") or {}).get("ms_per_step"),
              "frac %.3f step_frac %.3f" % (r.get("frac", 0), r.get("step_frac", 0)), "launches", d["gpu_launches"] // d["steps"],
              "rounds", d["config"]["rounds_per_sweep"], "bodies", (d.get("bodies") or {}).get("value"), (d.get("bodies") or {}).get("roofline_frac"))
        print("   ", d.get("kernel_breakdown_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
