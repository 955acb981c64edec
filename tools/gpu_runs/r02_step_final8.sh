#!/usr/bin/env bash
# 8-GPU box, frozen code: the driver's N = 8 line and the from-vacuum run
set -u
mkdir -p gpurun_out
TAG=${1:-final8}
run() { n=$1; shift; name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > gpurun_out/r02${TAG}_bench_n${n}_${name}.json 2> gpurun_out/r02${TAG}_bench_n${n}_${name}.err; }
run 8 default --steps 5 --warmup 3 --no-cpu-baseline
run 8 from_vacuum --steps 5 --warmup 3 --no-cpu-baseline --from-vacuum --no-ten-mode --no-parity
