"""Rates of the read-out kernels on the BASELINE config-2 state (8 modes, cutoff 10, 1e8
amplitudes): norm, all_fock_probs, single- and two-mode marginals (the MeasureFock distribution), reduced
density matrix of one mode, one fock_prob, and the pure -> mixed outer product of a 4-mode ket.  HBM-bound
reductions: algorithmic bytes = 16 B read per amplitude (+ what is written).  One JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from strawberryfields_b200 import B200FockBackend
    from strawberryfields_b200 import workloads as W

    n, D = 8, 10
    be = B200FockBackend()
    be.begin_circuit(n, cutoff_dim=D, lazy_vacuum=False)
    W.run_calls(be, W.config2_circuit(n, seed=42))
    c = be.circuit
    c._flush()
    N = D ** n
    peak = 6546.6
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def timed(fn, nbytes, reps=10):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        return {"ms": round(ms, 4), "GBps": round(nbytes / (ms * 1e-3) / 1e9, 1), "frac_of_measured_peak": round(nbytes / (ms * 1e-3) / 1e9 / peak, 3)}

    out = {"state": "config 2: 8 modes, cutoff 10, 1e8 amplitudes (1.6 GB)", "peak_GBps": peak}
    out["norm (b200_norm2)"] = timed(lambda: c._norm_device(), 16 * N)
    out["all_fock_probs (b200_abs2)"] = timed(lambda: c.fock_probs_device(), 24 * N)
    out["marginal of 1 mode (gather_reduce)"] = timed(lambda: c.marginal_probs_device([3]), 16 * N)
    out["marginal of 2 modes = MeasureFock distribution"] = timed(lambda: c.marginal_probs_device([0, 7]), 16 * N)
    out["marginal of 4 modes"] = timed(lambda: c.marginal_probs_device([1, 2, 5, 6]), 16 * N)
    out["reduced_dm of 1 mode (D x D outputs, 1e7-term sums)"] = timed(lambda: c.reduced_dm_device([4]), 16 * N * D, reps=3)
    st = be.state()
    out["fock_prob (one amplitude, D2H)"] = timed(lambda: st.fock_prob([0, 1, 0, 2, 0, 0, 1, 0]), 16)
    b4 = B200FockBackend()
    b4.begin_circuit(4, cutoff_dim=D, lazy_vacuum=False)
    W.run_calls(b4, W.config2_circuit(4, seed=1))
    b4.circuit._flush()

    def to_mixed():
        snap = b4.circuit.snapshot()
        snap._shared = False
        snap._to_mixed()

    out["pure -> mixed outer product (4 modes: 1e4 -> 1e8 entries written)"] = timed(to_mixed, 16 * D ** 8, reps=5)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
