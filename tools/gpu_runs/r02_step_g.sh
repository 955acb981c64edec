#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${1:-g}
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r02${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_c2.json 2> gpurun_out/r02${TAG}_bench_c2.err
for ex in auto; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29502 bench.py --gpus 2 --steps 5 --warmup 3 --exchange $ex --no-cpu-baseline \
      > gpurun_out/r02${TAG}_bench_n2_${ex}.json 2> gpurun_out/r02${TAG}_bench_n2_${ex}.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c3 > gpurun_out/r02${TAG}_bench_c3.json 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c4 > gpurun_out/r02${TAG}_bench_c4.json 2>&1
