#!/usr/bin/env bash
# 8-GPU box: the driver's N = 8 line (parity + 10-mode block), overlap off for comparison, from vacuum
set -u
mkdir -p gpurun_out
TAG=${1:-l}
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:3}" > gpurun_out/r02${TAG}_bench_n8_$2.json 2> gpurun_out/r02${TAG}_bench_n8_$2.err; }
run 29508 default --steps 5 --warmup 3 --no-cpu-baseline
run 29509 ov0 --steps 5 --warmup 3 --no-cpu-baseline --exchange-overlap 0 --no-ten-mode --no-parity
run 29510 from_vacuum --steps 5 --warmup 3 --no-cpu-baseline --from-vacuum --no-ten-mode --no-parity
nvidia-smi topo -m > gpurun_out/r02${TAG}_topo.txt 2>&1
