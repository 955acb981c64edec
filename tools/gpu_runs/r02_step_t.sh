#!/usr/bin/env bash
# 2-GPU box: final validation of the final code
set -u
mkdir -p gpurun_out
TAG=${1:-t}
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r02${TAG}_pytest_gpu.log
timeout 120 python tools/inner_probe.py > gpurun_out/r02${TAG}_inner_probe.json 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_c2.json 2> gpurun_out/r02${TAG}_bench_c2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c4 > gpurun_out/r02${TAG}_bench_c4.json 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02${TAG}_bench_n2.json 2> gpurun_out/r02${TAG}_bench_n2.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02${TAG}_smoke.log 2>&1
