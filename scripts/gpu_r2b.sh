#!/bin/bash
# dev round: GPU tests, shell bench, ncu capture of the shell kernel in the step loop
tag=${1:-r2b}
tests=${2:-tests}
out=gpurun_out
mkdir -p $out
(timeout 1200 python -m pytest $tests -m gpu -q -x 2>&1 | tail -15) > $out/${tag}_pytest_gpu.log
(timeout 300 python scripts/force_bench.py --shells 2>&1 | tail -12) > $out/${tag}_shells.log
B="python bench.py --steps 1 --warmup 1 --md-steps 300 --melt-steps 600 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_shell -s 1500 -c 1 \
    -f -o $out/${tag}_force $B > $out/${tag}_force_ncu.log 2>&1
(timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1) > $out/${tag}_bench.json
tail -5 $out/${tag}_pytest_gpu.log; cat $out/${tag}_shells.log; cut -c1-400 $out/${tag}_bench.json
