#!/bin/bash
# one ncu --set full capture of a few tile launches of the 1 M block (second frame), summary + report back in gpurun_out/
tag=${1:-r02h}; out=gpurun_out; mkdir -p $out
PS="--substeps 2 --iterations 10 --frames 2 --info-out $out/${tag}_plan_info.json"
python tools/profile_step.py $PS > $out/plain2_${tag}.log 2>&1 || { tail -5 $out/plain2_${tag}.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_tile_rounds -s 61 -c 8 -f -o $out/prof_${tag} \
    python tools/profile_step.py $PS > $out/ncu_full_${tag}.log 2>&1
tail -2 $out/ncu_full_${tag}.log
python tools/ncu_summary.py $out/prof_${tag}.ncu-rep $out/${tag}_ncu_tile_rounds.md
ls -la $out/prof_${tag}.ncu-rep
