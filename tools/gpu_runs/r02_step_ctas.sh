#!/usr/bin/env bash
# 4-GPU box: link CTAs of the exchange kernel, 32 against 48
set -u
mkdir -p gpurun_out
for c in 32 48; do
  B200_EXCHANGE_CTAS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
      --master-port 29504 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline --no-parity \
      > gpurun_out/r02ctas_bench_n4_$c.json 2> gpurun_out/r02ctas_bench_n4_$c.err
done
