"""Full-size oracle fixtures for BASELINE configs 2, 3 and 5 (run once in the build container; minutes
to hours of CPU, ~10 GB of RAM).  TEST INFRASTRUCTURE -- only tests/ read the files written here.

    python -m oracle.make_golden_fullsize c2        # 8 modes, cutoff 10: 1e8 amplitudes
    python -m oracle.make_golden_fullsize c3        # 4 modes mixed, cutoff 10: 1e8-element rho + MeasureFock
    python -m oracle.make_golden_fullsize c5        # 9 modes, cutoff 10: 1e9 amplitudes (config 5)

The circuits are ``strawberryfields_b200.workloads.config{2,3}_circuit(seed=42)`` -- exactly what
``bench.py`` times -- run by the oracle (``oracle/fock_oracle.py``, vector style, itself pinned to the
unmodified reference by tests/test_oracle_golden.py at reduced sizes).  Stored, per config, in
``tests/golden/ref_config{2,3,5}_full.npz`` (small):

* ``idx`` / ``amp``: 10 000 sampled Fock indices (low photon numbers, fixed seed) and the amplitudes
  <n|psi> (config 3: the diagonal entries <n|rho|n> and sampled off-diagonal entries);
* ``marg``: every single-mode photon-number marginal; ``trace`` (norm^2 / trace);
* config 3: the seed-7 ``MeasureFock`` outcome on all modes ("outcome compared exactly") and the
  outcomes of eight more seeds, the probabilities of the 64 most likely outcomes, and the post-measurement
  trace check.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def sample_indices(n_axes, D, count, seed):
    rng = np.random.RandomState(seed)
    idx = rng.choice(4, size=(count, n_axes), p=[0.6, 0.25, 0.1, 0.05]).astype(np.int64)
    idx[0] = 0
    idx[1: 1 + min(n_axes, count - 1)] = np.eye(n_axes, dtype=np.int64)[: count - 1]
    return np.minimum(idx, D - 1)


def pure_config(name, n_modes, D=10):
    from oracle.fock_oracle import OracleBackend
    from strawberryfields_b200 import workloads as W

    calls = W.config2_circuit(n_modes, seed=42)
    ob = OracleBackend()
    t0 = time.time()
    ob.begin_circuit(n_modes, cutoff_dim=D)
    for i, c in enumerate(calls):
        getattr(ob, c[0])(*c[1:])
        if i % 10 == 0:
            print("%s: gate %d / %d, %.0f s" % (name, i, len(calls), time.time() - t0), flush=True)
    psi = ob.state().data
    idx = sample_indices(n_modes, D, 10000, seed=2026)
    amp = psi[tuple(idx.T)]
    p = np.abs(psi) ** 2
    marg = np.stack([p.sum(axis=tuple(a for a in range(n_modes) if a != m)) for m in range(n_modes)])
    np.savez_compressed(os.path.join(GOLDEN, "ref_%s_full.npz" % name), idx=idx, amp=amp, marg=marg,
                        trace=np.array(p.sum()), n_modes=np.array(n_modes), cutoff=np.array(D),
                        gates=np.array(len(calls)))
    print("%s done in %.0f s: norm^2 = %.15f" % (name, time.time() - t0, p.sum()))


def mixed_config3(D=10, n_modes=4):
    from oracle.fock_oracle import OracleBackend
    from strawberryfields_b200 import workloads as W

    calls = W.config3_circuit(n_modes, seed=42)
    ob = OracleBackend()
    t0 = time.time()
    ob.begin_circuit(n_modes, cutoff_dim=D, pure=False)
    for i, c in enumerate(calls):
        getattr(ob, c[0])(*c[1:])
        print("c3: call %d / %d, %.0f s" % (i, len(calls), time.time() - t0), flush=True)
    st = ob.state()
    rho = st.data                                    # [D]*8, (ket_0, bra_0, ...)
    probs = np.array(st.all_fock_probs()).reshape([D] * n_modes)
    idx = sample_indices(2 * n_modes, D, 10000, seed=2027)   # entries of rho, mostly off-diagonal
    amp = rho[tuple(idx.T)]
    didx = sample_indices(n_modes, D, 2000, seed=2028)
    dprob = probs[tuple(didx.T)]
    marg = np.stack([probs.sum(axis=tuple(a for a in range(n_modes) if a != m)) for m in range(n_modes)])
    top = np.argsort(probs.reshape(-1))[::-1][:64]
    # further seeds, on copies of the circuit (seed 7 lands on the most likely outcome, the vacuum)
    import copy

    extra_seeds = np.array([1, 4, 24, 37, 45, 56, 118, 194])  # first uniform > 0.93 for all but seed 1: non-vacuum outcomes
    extra_out = []
    for sd in extra_seeds:
        ob2 = copy.deepcopy(ob)
        np.random.seed(int(sd))
        extra_out.append(np.asarray(ob2.measure_fock(list(range(n_modes)))).reshape(-1))
        del ob2
    np.random.seed(7)
    outcome = ob.measure_fock(list(range(n_modes)))
    post = ob.state()
    np.savez_compressed(os.path.join(GOLDEN, "ref_config3_full.npz"), idx=idx, amp=amp, diag_idx=didx,
                        diag_prob=dprob, marg=marg, trace=np.array(st.trace()), top_idx=top,
                        top_prob=probs.reshape(-1)[top], measure_seed=np.array(7), outcome=np.asarray(outcome),
                        extra_seeds=extra_seeds, extra_outcomes=np.stack(extra_out),
                        post_trace=np.array(post.trace()),
                        post_prob_of_outcome=np.array(post.fock_prob([0] * n_modes)),
                        n_modes=np.array(n_modes), cutoff=np.array(D), gates=np.array(len(calls)))
    print("c3 done in %.0f s: trace = %.15f, outcome = %s" % (time.time() - t0, st.trace(), outcome))


if __name__ == "__main__":
    what = sys.argv[1:] or ["c2", "c3"]
    if "c2" in what:
        pure_config("config2", 8)
    if "c3" in what:
        mixed_config3()
    if "c5" in what:
        pure_config("config5", 9)
