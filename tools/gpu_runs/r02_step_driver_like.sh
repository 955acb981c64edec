#!/usr/bin/env bash
# what the driver runs at round end on a 1-GPU box, on the frozen tree
set -u
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r02dl_pytest_gpu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02dl_smoke.log 2>&1
( time python bench.py --gpus 1 --steps 20 --warmup 3 ) > gpurun_out/r02dl_bench.json 2> gpurun_out/r02dl_bench.err
