#!/bin/bash
# One GPU-box session of round 2: GPU tests, launch list of the steady-state NVE step loop (config 2), full ncu captures of
# the shell pair-force kernel, the list-build kernels and the integrator, of the config-3 sampling kernels, the default
# bench line with all legs, the reference arm, and configs 3 / 5 as bench lines.
# Usage (under gpurun): bash scripts/gpu_round2.sh <tag>
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6) > $out/${tag}_pytest_gpu.log
B="python bench.py --steps 1 --warmup 1 --md-steps 300 --melt-steps 600 --no-cpu-baseline --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4300 -c 900 --csv \
    --log-file $out/${tag}_launches.csv $B > $out/${tag}_launches.log 2>&1
for k in k_force_shell:1500 k_nlist_build:40 k_shell_rows:40 k_verlet:900; do
    name=${k%%:*}; skip=${k##*:}
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 \
        -f -o $out/${tag}_$name $B > $out/${tag}_${name}_ncu.log 2>&1
done
B3="python bench.py --config cfg3 --steps 1 --warmup 1 --md-steps 200 --melt-steps 200 --no-cpu-baseline --no-extra"
for k in k_density_modes:1; do
    name=${k%%:*}; skip=${k##*:}
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 \
        -f -o $out/${tag}_$name $B3 > $out/${tag}_${name}_ncu.log 2>&1
done
(timeout 300 python scripts/dev/cfg3_sampling_cost.py 2>&1 | tail -20) > $out/${tag}_cfg3_sampling_cost.txt
(timeout 900 python bench.py 2>&1 | tail -1) > $out/${tag}_bench.json
(timeout 600 python bench.py --impl reference 2>&1 | tail -1) > $out/${tag}_bench_reference.json
(timeout 600 python bench.py --config cfg3 --no-cpu-baseline --no-extra 2>&1 | tail -1) > $out/${tag}_bench_cfg3.json
(timeout 600 python bench.py --config cfg5 --no-cpu-baseline --no-extra 2>&1 | tail -1) > $out/${tag}_bench_cfg5.json
tail -3 $out/${tag}_pytest_gpu.log; cut -c1-300 $out/${tag}_bench.json; cut -c1-300 $out/${tag}_bench_cfg3.json; cat $out/${tag}_cfg3_sampling_cost.txt
