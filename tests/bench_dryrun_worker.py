"""Worker of tests/test_bench_dryrun.py::test_sharded_arm: bench.py's multi-rank arm under torchrun
with gloo, the numpy double of the C ABI and wall-clock stand-ins for CUDA events (numbers are
meaningless; the control flow and the JSON line are what is checked)."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


def main():
    from fake_lib import FakeLib
    from strawberryfields_b200 import circuit, lib
    import bench as B

    lib._lib = FakeLib()
    circuit._TEST_HOST_MODE = True
    B.DEVICE = "cpu"
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.Event = _Event
    torch.Tensor.pin_memory = lambda self: self
    world = int(os.environ["WORLD_SIZE"])
    B.b200_arm(argparse.Namespace(gpus=world, steps=2, warmup=1, impl="b200", modes=int(sys.argv[1]),
                                  cutoff=int(sys.argv[2]), no_cpu_baseline=True, workload="c2", batch=2,
                                  exchange=sys.argv[3], fuse="fold", no_parity=False, no_ten_mode=False,
                                  parity_probs=12, exchange_overlap=8,
                                  from_vacuum=len(sys.argv) > 4 and sys.argv[4] == "from_vacuum"))


if __name__ == "__main__":
    main()
