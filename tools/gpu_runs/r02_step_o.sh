#!/usr/bin/env bash
# 4-GPU box, final code: N = 4 / 2 / 1 lines, launch list, full GPU test suite
set -u
mkdir -p gpurun_out
TAG=${1:-o}
run() { n=$1; shift; name=$1; shift; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > gpurun_out/r02${TAG}_bench_n${n}_${name}.json 2> gpurun_out/r02${TAG}_bench_n${n}_${name}.err; }
run 4 default --steps 5 --warmup 3 --no-cpu-baseline
run 2 default --steps 5 --warmup 3 --no-cpu-baseline
run 4 from_vacuum --steps 5 --warmup 3 --no-cpu-baseline --from-vacuum --no-parity
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r02${TAG}_bench_n1_c2.json 2> gpurun_out/r02${TAG}_bench_n1_c2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c3 > gpurun_out/r02${TAG}_bench_c3.json 2>&1
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r02${TAG}_pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02${TAG}_launches_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02${TAG}_launches_c2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02${TAG}_smoke.log 2>&1
