#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${1:-i}
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_backend.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r02${TAG}_pytest_kernels.log
for mode in wide narrow flat; do
  B200_INNER_ONE=$mode timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_c2_one_${mode}.json 2>&1
done
B200_INNER_LEGACY=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_c2_legacy.json 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply_inner_tma -s 5 -c 2 \
    -o gpurun_out/r02${TAG}_prof_inner_one python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r02${TAG}_ncu_inner.log 2>&1
