#!/usr/bin/env python3
"""State learning with a CV quantum neural network on the b200fock kernels.

The torch counterpart of the reference's ``examples/quantum_neural_network.py`` (TensorFlow
backend): the same layer structure and weight layout, the circuit runs on the CUDA Fock kernels
and is differentiated by ``strawberryfields_b200.autodiff`` (adjoint method).

    python examples/qnn_torch.py [--modes 1] [--layers 8] [--cutoff 6] [--steps 200]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from strawberryfields_b200 import TorchCircuit  # noqa: E402
from strawberryfields_b200.autodiff import qnn_init_weights, qnn_layer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", type=int, default=1)
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--cutoff", type=int, default=6)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--lr", type=float, default=0.01)
    args = ap.parse_args()

    gen = torch.Generator().manual_seed(137)
    weights = qnn_init_weights(args.modes, args.layers, generator=gen).requires_grad_(True)
    opt = torch.optim.Adam([weights], lr=args.lr)
    target = (1,) + (0,) * (args.modes - 1)  # single photon in mode 0

    for step in range(args.steps):
        prog = TorchCircuit(args.modes, args.cutoff)
        for k in range(args.layers):
            qnn_layer(prog, weights[k])
        ket = prog.ket()
        fidelity = ket[target].abs() ** 2
        loss = 1 - fidelity
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % 20 == 0 or step == args.steps - 1:
            trace = (ket.detach().abs() ** 2).sum().item()
            print("step %4d  fidelity %.6f  trace %.6f" % (step, fidelity.item(), trace))


if __name__ == "__main__":
    main()
