#!/bin/bash
# strong scaling of the 8 M-vertex mesh: the real-GPU bit-identity test for N GPUs, then bench.py as the driver launches it; logs under gpurun_out/
n=${1:-8}; tag=${2:-r02}; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "[$n-" > $out/pytest_multigpu_${tag}_n$n.log 2>&1; tail -2 $out/pytest_multigpu_${tag}_n$n.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus $n --steps 10 --warmup 3 \
  > $out/bench_${tag}_8M_n$n.json 2> $out/bench_${tag}_8M_n$n.err || tail -5 $out/bench_${tag}_8M_n$n.err
python - $out/bench_${tag}_8M_n$n.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "e2e", (d.get("e2e") or {}).get("ms_per_step"), "chk", (d.get("state_checksum") or {}).get("x4_words_hi_lo"),
      "tiles0", (d.get("distribution") or {}).get("tiles_rank0"), "bodies", (d.get("bodies") or {}).get("value"), "build_s", d.get("plan_build_seconds"))
PY
