#!/bin/bash
# Strong scaling of the 64 M-particle system on N GPUs of one box, plus the real-NCCL parity test.
# Usage (under gpurun --gpus N): bash scripts/gpu_strong.sh <N> <tag>
n=${1:-2}; tag=${2:-r2s}
out=gpurun_out
mkdir -p $out
(timeout 600 python -m pytest tests -m gpu -q -k "nccl or blocking_scheme_pairs" 2>&1 | tail -40) > $out/${tag}_pytest_n${n}.log
(timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --n-total 64000000 --steps 2 --warmup 1 --md-steps 100 --melt-steps 300 --no-extra 2> $out/${tag}_strong_n${n}.err | tail -1) > $out/${tag}_strong_n${n}.json
cat $out/${tag}_pytest_n${n}.log; cut -c1-400 $out/${tag}_strong_n${n}.json; tail -5 $out/${tag}_strong_n${n}.err
