#!/bin/bash
# A/B of library builds of several commits (build/ab/<sha>: git archive + in-tree build) on one box
out=$PWD/gpurun_out; mkdir -p $out
for d in build/ab/*/ .; do
  tag=$(basename $d); [ "$d" = "." ] && tag=HEAD
  ( cd $d
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-bodies --kernel-breakdown > $out/ab_${tag}_block.json 2> $out/ab_${tag}.err || tail -3 $out/ab_${tag}.err
    python tools/ab_plan.py --bodies 1184 --variants "" > $out/ab_${tag}_bodies.log 2>&1
    python tools/ab_plan.py --n 160 --frames 5 --variants "" > $out/ab_${tag}_4M.log 2>&1 )
  python - $out/ab_${tag}_block.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "launches", d["gpu_launches"] // d["steps"], "rounds", d["config"]["rounds_per_sweep"], d.get("kernel_breakdown_ms"))
PY
  cat $out/ab_${tag}_bodies.log $out/ab_${tag}_4M.log
done
