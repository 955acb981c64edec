#!/usr/bin/env bash
# 4-GPU box: sharded tests, N=2 (overlap on / off), N=4, and the single-GPU line
set -u
mkdir -p gpurun_out
TAG=${1:-j}
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_c2.json 2> gpurun_out/r02${TAG}_bench_c2.err
timeout 400 python -m pytest tests/test_sharding.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r02${TAG}_pytest_sharding.log
for ov in 8 0; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29502 bench.py --gpus 2 --steps 5 --warmup 3 --exchange-overlap $ov --no-cpu-baseline \
      > gpurun_out/r02${TAG}_bench_n2_ov${ov}.json 2> gpurun_out/r02${TAG}_bench_n2_ov${ov}.err
done
for ov in 8 0; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
      --master-port 29504 bench.py --gpus 4 --steps 5 --warmup 3 --exchange-overlap $ov --no-cpu-baseline \
      > gpurun_out/r02${TAG}_bench_n4_ov${ov}.json 2> gpurun_out/r02${TAG}_bench_n4_ov${ov}.err
done
