#!/bin/bash
# Development aid (under gpurun): A/B timing of library variants built by scripts/build_variants.sh on ONE box — box-to-box
# differences (~1 %) exceed the effects being measured, runs on the same box repeat to 0.05 %.
# usage: bash scripts/ab_variants.sh name1 name2 ...
for v in "$@"; do
  HALMD_B200_LIB=scripts/_build/$v.so timeout 200 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e9,4), round(d['ms_per_step'],3), round(d['roofline']['avg_launch_us'],2))"
done
