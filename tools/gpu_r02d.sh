#!/bin/bash
# 2-GPU session: real-GPU bit-identity tests, strong scaling of a 2 M mesh (1 vs 2 GPUs), the driver's N=2 bench line (8 M mesh)
out=gpurun_out; mkdir -p $out; tag=${1:-r02g}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > $out/pytest_multigpu_${tag}.log 2>&1; tail -3 $out/pytest_multigpu_${tag}.log
python bench.py --workload dist --size 126 --steps 10 --no-bodies --no-cpu-baseline > $out/bench_${tag}_2M_n1.json 2> $out/bench_${tag}.err || tail -3 $out/bench_${tag}.err
$TR --master-port 29600 bench.py --gpus 2 --workload dist --size 126 --steps 10 --no-bodies > $out/bench_${tag}_2M_n2.json 2>> $out/bench_${tag}.err || tail -3 $out/bench_${tag}.err
$TR --master-port 29601 bench.py --gpus 2 --workload dist --size 126 --steps 10 --no-bodies --slabs > $out/bench_${tag}_2M_n2_slabs.json 2>> $out/bench_${tag}.err || tail -3 $out/bench_${tag}.err
python - $tag <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/bench_%s_*.json" % sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "e2e", (d.get("e2e") or {}).get("ms_per_step"), "chk", (d.get("state_checksum") or {}).get("x4_words_hi_lo"),
              "tiles0", (d.get("distribution") or {}).get("tiles_rank0"), "bodies", (d.get("bodies") or {}).get("value"), "build_s", d.get("plan_build_seconds"))
    except Exception as e:
        print(f, "ERR", e)
PY
