#!/usr/bin/env bash
# 2-GPU validation of the sharded path (run as ONE gpurun --gpus 2 call from the repo root).
set -u
mkdir -p gpurun_out
TAG=${1:-a}
timeout 300 python -m pytest tests/test_sharding.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r02${TAG}_pytest_sharding_gpu_n2.log
for ex in auto push; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29502 bench.py --gpus 2 --steps 5 --warmup 3 --exchange $ex --no-cpu-baseline \
      > gpurun_out/r02${TAG}_bench_n2_${ex}.json 2> gpurun_out/r02${TAG}_bench_n2_${ex}.err
done
