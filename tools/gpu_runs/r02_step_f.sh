#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${1:-f}
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r02${TAG}_pytest_kernels.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_c2.json 2> gpurun_out/r02${TAG}_bench_c2.err
for ex in auto push; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29502 bench.py --gpus 2 --steps 5 --warmup 3 --exchange $ex --no-cpu-baseline \
      > gpurun_out/r02${TAG}_bench_n2_${ex}.json 2> gpurun_out/r02${TAG}_bench_n2_${ex}.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply_inner_tma -s 8 -c 2 \
    -o gpurun_out/r02${TAG}_prof_inner_tma python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r02${TAG}_ncu_inner.log 2>&1
