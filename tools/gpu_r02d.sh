#!/bin/bash
# 2-GPU session: real-GPU bit-identity tests, the driver's N=2 bench line, strong scaling on a 2 M mesh with A/B switches
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multigpu.py tests/test_partition.py -m gpu -x -q 2>&1 | tail -6
python bench.py --workload dist --n 126 --steps 10 --no-bodies --no-cpu-baseline > $out/bench_r02d_2M_n1.json 2> $out/bench_r02d.err || tail -3 $out/bench_r02d.err
for v in "" "--slabs" "--no-fuse" "--no-pdl"; do
  $TR --master-port 29600 bench.py --gpus 2 --workload dist --n 126 --steps 10 --no-bodies $v > $out/bench_r02d_2M_n2$(echo $v | tr -d " ").json 2>> $out/bench_r02d.err || tail -3 $out/bench_r02d.err
done
$TR --master-port 29601 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_r02d_8M_n2.json 2>> $out/bench_r02d.err || tail -3 $out/bench_r02d.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r02d_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "e2e", (d.get("e2e") or {}).get("ms_per_step"), "chk", (d.get("state_checksum") or {}).get("x4_words_hi_lo"),
              "tiles0", d["config"].get("tiles_rank0"), "bodies", (d.get("bodies") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
