#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${1:-s}
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -12 > gpurun_out/r02${TAG}_pytest_kernels.log
timeout 120 python tools/inner_probe.py > gpurun_out/r02${TAG}_inner_probe.json 2>&1
B200_INNER_TMAP=plain timeout 120 python tools/inner_probe.py > gpurun_out/r02${TAG}_inner_probe_plain.json 2>&1
B200_INNER_TMAP=0 timeout 120 python tools/inner_probe.py > gpurun_out/r02${TAG}_inner_probe_bulk.json 2>&1
