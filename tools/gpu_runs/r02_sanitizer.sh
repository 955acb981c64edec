#!/usr/bin/env bash
# compute-sanitizer over the kernel tests (SURVEY section 5 plan): the in-place gate kernels rely on every task
# touching a disjoint set of amplitudes, the TMA-staged kernel on its mbarrier hand-offs
set -u
mkdir -p gpurun_out
TAG=${1:-q}
SEL="inner_axis or diag_multi or apply_gate1_every_axis or apply_gate2_batched or gen_gate2 or gram1 or outer_axis"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --launch-timeout 120 \
      python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "$SEL" > gpurun_out/r02${TAG}_sanitizer_${tool}.log 2>&1
  echo "exit code $?" >> gpurun_out/r02${TAG}_sanitizer_${tool}.log
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02${TAG}_sanitizer_smoke_memcheck.log 2>&1
echo "exit code $?" >> gpurun_out/r02${TAG}_sanitizer_smoke_memcheck.log
