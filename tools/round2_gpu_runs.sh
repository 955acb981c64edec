#!/usr/bin/env bash
# First GPU calls of round 2: everything that was written after round 1's GPU budget ended.
# Run each block as ONE gpurun call from the repo root; results land in gpurun_out/.
#
#   gpurun --timeout 1500 -- 'bash tools/round2_gpu_runs.sh single'
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/round2_gpu_runs.sh multi'
set -u
mkdir -p gpurun_out
case "${1:-single}" in
single)
  # 1. parity: the whole GPU suite (new: state observables, torch autograd, lazy vacuum)
  timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu.log
  # 2. default bench (steady state) and the same circuit run from vacuum with lazy factors
  timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
  timeout 300 python bench.py --steps 10 --warmup 3 --from-vacuum --no-cpu-baseline \
      > gpurun_out/r02_bench_c2_from_vacuum.json 2> gpurun_out/r02_bench_c2_from_vacuum.err
  timeout 300 python bench.py --steps 10 --warmup 3 --workload c4 --from-vacuum --no-cpu-baseline \
      > gpurun_out/r02_bench_c4_from_vacuum.json 2>&1
  timeout 300 python bench.py --steps 10 --warmup 3 --workload c3 --from-vacuum --no-cpu-baseline \
      > gpurun_out/r02_bench_c3_from_vacuum.json 2>&1
  # 3. the differentiable path: one QNN training run
  timeout 300 python examples/qnn_torch.py --modes 2 --layers 4 --cutoff 10 --steps 40 > gpurun_out/r02_qnn_torch.log 2>&1
  timeout 300 python tools/bench_autodiff.py > gpurun_out/r02_bench_autodiff_c4.json 2> gpurun_out/r02_bench_autodiff_c4.err
  ;;
multi)
  for n in 2 4 8; do
    for ex in auto push; do
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
          --master-port $((29500 + n)) bench.py --gpus $n --steps 5 --warmup 3 --exchange $ex --no-cpu-baseline \
          > gpurun_out/r02_bench_n${n}_${ex}.json 2> gpurun_out/r02_bench_n${n}_${ex}.err
    done
  done
  # whole program runs from vacuum with the sharded lazy vacuum (replicated prefix, one exchange)
  for m in 9 10; do
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29520 bench.py --gpus 8 --modes $m --steps 3 --warmup 3 --from-vacuum --no-cpu-baseline \
        > gpurun_out/r02_bench_n8_${m}modes_from_vacuum.json 2> gpurun_out/r02_bench_n8_${m}modes_from_vacuum.err
  done
  timeout 200 python -m pytest tests/test_sharding.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_pytest_sharding_gpu.log
  ;;
esac
