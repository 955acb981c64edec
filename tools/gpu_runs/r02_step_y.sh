#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${1:-y}
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r02${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_c2.json 2> gpurun_out/r02${TAG}_bench_c2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --from-vacuum > gpurun_out/r02${TAG}_bench_c2_from_vacuum.json 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c3 > gpurun_out/r02${TAG}_bench_c3.json 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c4 --from-vacuum > gpurun_out/r02${TAG}_bench_c4_from_vacuum.json 2>&1
