#!/bin/bash
# final 1-GPU session of the round: whole GPU test-suite, then the profile set
out=gpurun_out; mkdir -p $out; tag=${1:-r02z}
timeout 1700 python -m pytest tests -m gpu -x -q --durations=6 > $out/pytest_gpu_${tag}.log 2>&1; tail -12 $out/pytest_gpu_${tag}.log
bash tools/profile_round.sh $tag
