#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${1:-h}
timeout 120 tools/probes/fp64_probe > gpurun_out/r02${TAG}_fp64_probe.json 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r02${TAG}_pytest_kernels.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_c2.json 2> gpurun_out/r02${TAG}_bench_c2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c3 > gpurun_out/r02${TAG}_bench_c3.json 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c4 > gpurun_out/r02${TAG}_bench_c4.json 2>&1
timeout 600 python bench.py --steps 3 --warmup 1 --impl reference > gpurun_out/r02${TAG}_bench_ref.json 2> gpurun_out/r02${TAG}_bench_ref.err
