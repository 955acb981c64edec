#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${1:-c}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29511 tools/xchg_probe.py 9 > gpurun_out/r02${TAG}_xchg_probe.json 2> gpurun_out/r02${TAG}_xchg_probe.err
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r02${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_c2.json 2> gpurun_out/r02${TAG}_bench_c2.err
B200_INNER_LEGACY=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_c2_legacy_inner.json 2>&1
