#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
TAG=${1:-m}
timeout 120 python tools/inner_probe.py > gpurun_out/r02${TAG}_inner_probe.json 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_apply_inner_tma -s 3 -c 1 \
    -o gpurun_out/r02${TAG}_prof_rows python tools/inner_probe.py > gpurun_out/r02${TAG}_ncu_rows.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_apply_inner_tma -s 33 -c 1 \
    -o gpurun_out/r02${TAG}_prof_blocks python tools/inner_probe.py > gpurun_out/r02${TAG}_ncu_blocks.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c3 > gpurun_out/r02${TAG}_bench_c3.json 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c4 > gpurun_out/r02${TAG}_bench_c4.json 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --workload c1 > gpurun_out/r02${TAG}_bench_c1.json 2>&1
