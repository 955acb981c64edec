#!/usr/bin/env python3
"""Time one training step (forward + backward) of a CV-QNN layer stack on the differentiable path
(strawberryfields_b200.autodiff) -- BASELINE config 4's shape by default: 6 modes, cutoff 10, batch 64,
one layer with per-entry weights.  Prints one JSON line.  Not part of bench.py's contract: this is
the measurement for SURVEY 8(f)3.

    python tools/bench_autodiff.py [--modes 6] [--cutoff 10] [--batch 64] [--layers 1] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from strawberryfields_b200 import TorchCircuit, lib  # noqa: E402
from strawberryfields_b200.autodiff import qnn_init_weights, qnn_layer, qnn_layer_size  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", type=int, default=6)
    ap.add_argument("--cutoff", type=int, default=10)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--layers", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()

    handle = lib.load(build_if_missing=False)
    gen = torch.Generator().manual_seed(42)
    w = torch.stack([qnn_init_weights(args.modes, args.layers, active_sd=0.05, generator=gen)
                     for _ in range(args.batch)], dim=-1).cuda().requires_grad_(True)   # [layers, size, batch]

    def step():
        prog = TorchCircuit(args.modes, args.cutoff, batch_size=args.batch)
        for k in range(args.layers):
            qnn_layer(prog, w[k])
        ket = prog.ket()
        loss = (1 - ket.reshape(args.batch, -1)[:, 1].abs() ** 2).mean()
        w.grad = None
        loss.backward()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    handle.b200_reset_launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    gates = args.layers * (2 * (args.modes * (args.modes - 1) // 2 + max(1, args.modes - 1)) + 3 * args.modes)
    elements = args.batch * args.cutoff ** args.modes
    print(json.dumps({
        "what": "forward + backward of %d QNN layer(s), %d modes, cutoff %d, batch %d, %d weights per entry"
                % (args.layers, args.modes, args.cutoff, args.batch, args.layers * qnn_layer_size(args.modes)),
        "ms_per_training_step": dt * 1e3, "gates": gates, "stored_elements": elements,
        "forward_equivalent_updates_per_s": gates * elements / dt,
        "kernel_launches_per_step": int(handle.b200_launch_count()) // args.steps,
        "loss": float(loss.detach()), "peak_memory_GB": torch.cuda.max_memory_allocated() / 1e9,
    }))


if __name__ == "__main__":
    main()
