#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${1:-k}
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r02${TAG}_pytest_kernels.log
timeout 120 python tools/inner_probe.py > gpurun_out/r02${TAG}_inner_probe_default.json 2>&1
B200_INNER_ROWS=flat timeout 120 python tools/inner_probe.py > gpurun_out/r02${TAG}_inner_probe_rowsflat.json 2>&1
B200_INNER_TEAMS=2 timeout 120 python tools/inner_probe.py > gpurun_out/r02${TAG}_inner_probe_teams2.json 2>&1
B200_INNER_TEAMS=1 timeout 120 python tools/inner_probe.py > gpurun_out/r02${TAG}_inner_probe_teams1.json 2>&1
B200_INNER_LEGACY=1 timeout 120 python tools/inner_probe.py > gpurun_out/r02${TAG}_inner_probe_legacy.json 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_c2.json 2> gpurun_out/r02${TAG}_bench_c2.err
