#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${1:-e}
timeout 120 tools/probes/tma_probe > gpurun_out/r02${TAG}_tma_probe.json 2>&1
for ex in auto push; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29502 bench.py --gpus 2 --steps 5 --warmup 3 --exchange $ex --no-cpu-baseline \
      > gpurun_out/r02${TAG}_bench_n2_${ex}.json 2> gpurun_out/r02${TAG}_bench_n2_${ex}.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29511 tools/xchg_probe.py 9 > gpurun_out/r02${TAG}_xchg_probe.json 2> gpurun_out/r02${TAG}_xchg_probe.err
timeout 200 python -m pytest tests/test_sharding.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r02${TAG}_pytest_sharding.log
