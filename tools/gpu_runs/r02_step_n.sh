#!/usr/bin/env bash
# 8-GPU box, final code: the driver's lines at N = 8 / 4 / 2 / 1 and the from-vacuum runs
set -u
mkdir -p gpurun_out
TAG=${1:-n}
run() { n=$1; shift; name=$1; shift; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > gpurun_out/r02${TAG}_bench_n${n}_${name}.json 2> gpurun_out/r02${TAG}_bench_n${n}_${name}.err; }
run 8 default --steps 5 --warmup 3 --no-cpu-baseline
run 8 ov0 --steps 5 --warmup 3 --no-cpu-baseline --exchange-overlap 0 --no-ten-mode --no-parity
run 4 default --steps 5 --warmup 3 --no-cpu-baseline
run 2 default --steps 5 --warmup 3 --no-cpu-baseline
run 8 from_vacuum --steps 5 --warmup 3 --no-cpu-baseline --from-vacuum --no-ten-mode --no-parity
run 8 from_vacuum_10modes --modes 10 --steps 3 --warmup 2 --no-cpu-baseline --from-vacuum --no-ten-mode --no-parity
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_n1_c2.json 2> gpurun_out/r02${TAG}_bench_n1_c2.err
