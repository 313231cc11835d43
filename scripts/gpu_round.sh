#!/bin/bash
# One GPU-box session: GPU tests, launch list of the step loop, full ncu capture of the pair-force kernel, bench.
# Usage (under gpurun): bash scripts/gpu_round.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > $out/${tag}_pytest_gpu.log
B="python bench.py --steps 1 --warmup 1 --md-steps 100 --melt-steps 1000 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 400 --csv \
    --log-file $out/${tag}_launches.csv $B > $out/${tag}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force -s 1200 -c 2 \
    -f -o $out/${tag}_force $B > $out/${tag}_force_ncu.log 2>&1
(timeout 600 python bench.py 2>&1 | tail -3) > $out/${tag}_bench.log
tail -5 $out/${tag}_pytest_gpu.log; tail -2 $out/${tag}_launches.log; tail -2 $out/${tag}_force_ncu.log; cat $out/${tag}_bench.log
