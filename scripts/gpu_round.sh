#!/bin/bash
# One GPU-box session: GPU tests, launch list of the steady-state NVE step loop, full ncu captures of the pair-force,
# neighbour-build and integrator kernels, the reference's kernels timed next to the product's, bench.  Usage (under gpurun): bash scripts/gpu_round.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
(timeout 1000 python -m pytest tests -m gpu -q 2>&1 | tail -6) > $out/${tag}_pytest_gpu.log
B="python bench.py --steps 1 --warmup 1 --md-steps 300 --melt-steps 600 --no-cpu-baseline"
# launches: melt 600 Andersen steps (4 launches per step + rebuilds, ~4200 in all), then 2 x 300 NVE steps (2 launches per
# step + 14 per rebuild); list ~900 launches of the NVE part
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4300 -c 900 --csv \
    --log-file $out/${tag}_launches.csv $B > $out/${tag}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_row2 -s 1500 -c 1 \
    -f -o $out/${tag}_force $B > $out/${tag}_force_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nlist_build -s 40 -c 1 \
    -f -o $out/${tag}_nlist $B > $out/${tag}_nlist_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_verlet -s 900 -c 1 \
    -f -o $out/${tag}_verlet $B > $out/${tag}_verlet_ncu.log 2>&1
(timeout 200 python scripts/ref_kernel_times.py 2> $out/${tag}_ref_kernel_times.err) > $out/${tag}_ref_kernel_times.json
(timeout 200 python scripts/e2e_breakdown.py 2>&1 | tail -8) > $out/${tag}_e2e_breakdown.txt
(timeout 600 python bench.py 2>&1 | tail -1) > $out/${tag}_bench.json
(timeout 600 python bench.py --impl reference 2>&1 | tail -1) > $out/${tag}_bench_reference.json
tail -3 $out/${tag}_pytest_gpu.log; tail -1 $out/${tag}_launches.log | cut -c1-200; cut -c1-300 $out/${tag}_bench.json; cut -c1-600 $out/${tag}_bench_reference.json
