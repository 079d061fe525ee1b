#!/bin/bash
# Round artefacts for profiles/: bench lines, ncu launch list of the same command, one ncu --set full of the tile launches.
# usage (on the GPU box): tools/profile_round.sh <round-tag>        (then copy gpurun_out/* summaries into profiles/)
set -u
tag=${1:-r02z}
out=gpurun_out; mkdir -p $out
python bench.py --steps 20 --warmup 3 --kernel-breakdown > $out/bench_${tag}.json 2> $out/bench_${tag}.err || { tail -5 $out/bench_${tag}.err; exit 1; }
python bench.py --steps 20 --warmup 3 --kernel-breakdown --fast-math --no-cpu-baseline --no-bodies > $out/bench_${tag}_fast.json 2>> $out/bench_${tag}.err
python bench.py --impl reference --steps 5 --warmup 3 > $out/bench_${tag}_reference.json 2>> $out/bench_${tag}.err
# A/B on the same box: CTA width / rim merge of the planner, and the launch structure
python tools/ab_plan.py --variants "SB_MERGE_PCT=100;SB_MERGE_RIMS=0;" > $out/ab_plan_${tag}.log 2>&1
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-bodies"
$B --block-threads 128 > $out/bench_${tag}_bt128.json 2>> $out/bench_${tag}.err
$B --no-fuse --no-snake > $out/bench_${tag}_nofuse_nosnake.json 2>> $out/bench_${tag}.err
$B --workload sphere > $out/bench_${tag}_sphere100k.json 2>> $out/bench_${tag}.err
# launch list of a short run of the same command (cold-cache, serialised: compare shares)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bodies > $out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 700 --csv --log-file $out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bodies > $out/ncu_launch_${tag}.log 2>&1
python tools/launch_summary.py $out/launches_${tag}.csv $out/${tag}_launches.md > /dev/null
# one ncu --set full capture of ALL tile launches of one frame of 2 substeps x 10 iterations (61 launches: single, double
# and substep-boundary ones in the proportions of the bench), second frame
PS="--substeps 2 --iterations 10 --frames 2 --info-out $out/${tag}_plan_info.json"
python tools/profile_step.py $PS > $out/plain2_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tile_rounds -s 61 -c 61 -f -o $out/prof_${tag} \
    python tools/profile_step.py $PS > $out/ncu_full_${tag}.log 2>&1
tail -2 $out/ncu_full_${tag}.log
python tools/ncu_summary.py $out/prof_${tag}.ncu-rep $out/${tag}_ncu_tile_rounds.md > /dev/null
python tools/traffic_json.py $out/prof_${tag}.ncu-rep $out/${tag}_traffic.json "${tag}" $out/${tag}_plan_info.json
for f in bench_${tag}.json bench_${tag}_fast.json bench_${tag}_reference.json bench_${tag}_bt128.json bench_${tag}_nofuse_nosnake.json bench_${tag}_sphere100k.json; do python - $out/$f <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.4g" % d["value"], d.get("unit"), "ms/step %.3f" % d["ms_per_step"], "e2e %.4g" % (d.get("e2e") or {}).get("value", 0),
      "roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in (d.get("roofline") or {}).items() if k in ("frac", "achieved", "step_frac", "launch_ms", "share_of_step")},
      "cpu", (d.get("cpu_baseline") or {}).get("value"), "bodies", (d.get("bodies") or {}).get("value"))
PY
done
cat $out/ab_plan_${tag}.log
