#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${1:-v}
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r02${TAG}_pytest_gpu.log
timeout 300 python tools/bench_reductions.py > gpurun_out/r02${TAG}_bench_reductions.json 2> gpurun_out/r02${TAG}_bench_reductions.err
