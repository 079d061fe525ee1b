#!/bin/bash
# the round's final artefacts in a few minutes: bench line, ncu launch list of the same command, ncu --set full of 12 tile launches
set -u
tag=${1:-r02z}; out=gpurun_out; mkdir -p $out
timeout 150 python bench.py --steps 20 --warmup 3 --kernel-breakdown --no-bodies > $out/bench_${tag}.json 2> $out/bench_${tag}.err || tail -3 $out/bench_${tag}.err
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 700 --csv --log-file $out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bodies > $out/ncu_launch_${tag}.log 2>&1
python tools/launch_summary.py $out/launches_${tag}.csv $out/${tag}_launches.md > /dev/null
PS="--substeps 2 --iterations 10 --frames 2 --info-out $out/${tag}_plan_info.json"
timeout 150 ncu --set full --clock-control none -k regex:k_tile_rounds -s 61 -c 12 -f -o $out/prof_${tag} \
    python tools/profile_step.py $PS > $out/ncu_full_${tag}.log 2>&1
tail -1 $out/ncu_full_${tag}.log
python tools/ncu_summary.py $out/prof_${tag}.ncu-rep $out/${tag}_ncu_tile_rounds.md > /dev/null
python tools/traffic_json.py $out/prof_${tag}.ncu-rep $out/${tag}_traffic.json "${tag}, 12 consecutive tile launches" $out/${tag}_plan_info.json > /dev/null
ls -la $out/prof_${tag}.ncu-rep; rm -f $out/prof_${tag}.ncu-rep
python - $out/bench_${tag}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g ms/step %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items() if k in ("frac", "step_frac", "launch_ms", "share_of_step")})
PY
head -12 $out/${tag}_launches.md
