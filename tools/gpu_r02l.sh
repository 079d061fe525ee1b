#!/bin/bash
# push-to-next-toucher (virtual ranks on one GPU) + CTA width variants at 4 M vertices and on the body batch
out=gpurun_out; mkdir -p $out; tag=r02l
timeout 1200 python -m pytest tests/test_dist.py tests/test_gpu_parity.py -m gpu -x -q > $out/pytest_gpu_${tag}.log 2>&1; tail -3 $out/pytest_gpu_${tag}.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-bodies"
SB_MERGE_PCT=115 $B --block-threads 160 > $out/bench_${tag}_bt160_m115.json 2>> $out/bench_${tag}.err
SB_MERGE_PCT=130 $B --block-threads 160 > $out/bench_${tag}_bt160_m130.json 2>> $out/bench_${tag}.err
SB_MERGE_PCT=170 $B --block-threads 192 > $out/bench_${tag}_bt192_m170.json 2>> $out/bench_${tag}.err
SB_MERGE_PCT=115 $B --block-threads 256 --tile-cap 2197 > $out/bench_${tag}_bt256_c2197_m115.json 2>> $out/bench_${tag}.err
for bt in 128 160; do
python - $bt <<'PY' > $out/ab_${tag}_4M_bt$bt.log 2>&1
import sys, json
from softbodyunity_b200 import SoftBody, meshgen
bt = int(sys.argv[1])
pos, tets, tris = meshgen.block(160, spacing=0.01, origin=(0.0, 0.002, 0.0))
sb = SoftBody(pos, tets, tris, block_threads=bt)
sb.step(frames=3); sb.synchronize()
ms = min(sb.time_frames(5) / 5 for _ in range(3))
i = sb.info()
print(json.dumps(dict(bt=bt, n_verts=len(pos), ms_per_frame=ms, gvs=len(pos) * 10 / ms / 1e6, rounds=sum(i["rounds_in_pass"]), tiles=i["tiles_in_pass"][:4])))
PY
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r02l*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "step_frac %.3f" % (r.get("step_frac", 0)), "rounds", d["config"]["rounds_per_sweep"], d["config"]["tiles_in_pass"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 2 $out/ab_r02l_*.log
