"""Backend call scripts shared by the golden-fixture generator and the parity tests.

A script is ``(name, n_modes, cutoff, pure, [(method, args, kwargs), ...])``.  The
methods are the ``BaseFock`` backend API (``/root/reference/strawberryfields/
backends/base.py:155-621``) -- exactly what ``LocalEngine._run_program`` calls
one command at a time (``engine.py:422-457``).  The pseudo-method ``"seed"``
calls ``np.random.seed`` (the reference samples measurements from numpy's
global stream, ``circuit.py:684``).

The same script is run on (a) the unmodified reference ``FockBackend`` in the
build container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``), (b) the
numpy oracle, (c) the CUDA backend.
"""
from __future__ import annotations

import numpy as np


def run_script(backend, script, record_states=False):
    """Run a script; returns (list of non-None return values, final state object)."""
    name, n, D, pure, calls = script
    backend.begin_circuit(n, cutoff_dim=D, pure=pure)
    returns = []
    for method, args, kwargs in calls:
        if method == "seed":
            np.random.seed(args[0])
            continue
        r = getattr(backend, method)(*args, **kwargs)
        if r is not None:
            returns.append(np.asarray(r))
    return returns, backend.state()


def C(method, *args, **kwargs):
    return (method, args, kwargs)


def _rng(seed):
    return np.random.RandomState(seed)


# ---------------------------------------------------------------------------
def boson_sampling(D):
    """BASELINE config 1: examples/boson_sampling.py:5-41 (fixed literals)."""
    calls = [
        C("prepare_fock_state", 1, 0),
        C("prepare_fock_state", 1, 1),
        C("prepare_vacuum_state", 2),
        C("prepare_fock_state", 1, 3),
        C("rotation", 0.5719, 0),
        C("rotation", -1.9782, 1),
        C("rotation", 2.0603, 2),
        C("rotation", 0.0644, 3),
        C("beamsplitter", 0.7804, 0.8578, 0, 1),
        C("beamsplitter", 0.06406, 0.5165, 2, 3),
        C("beamsplitter", 0.473, 0.1176, 1, 2),
        C("beamsplitter", 0.563, 0.1517, 0, 1),
        C("beamsplitter", 0.1323, 0.9946, 2, 3),
        C("beamsplitter", 0.311, 0.3231, 1, 2),
        C("beamsplitter", 0.4348, 0.0798, 0, 1),
        C("beamsplitter", 0.4368, 0.6157, 2, 3),
    ]
    return (f"boson_sampling_d{D}", 4, D, True, calls)


def gbs(D):
    """tests/integration/test_algorithms.py:94-136 (Gaussian boson sampling probs)."""
    calls = [C("squeeze", 1.0, 0.0, m) for m in range(4)]
    calls += [
        C("rotation", 0.5719, 0),
        C("rotation", -1.9782, 1),
        C("rotation", 2.0603, 2),
        C("rotation", 0.0644, 3),
        C("beamsplitter", 0.7804, 0.8578, 0, 1),
        C("beamsplitter", 0.06406, 0.5165, 2, 3),
        C("beamsplitter", 0.473, 0.1176, 1, 2),
        C("beamsplitter", 0.563, 0.1517, 0, 1),
        C("beamsplitter", 0.1323, 0.9946, 2, 3),
        C("beamsplitter", 0.311, 0.3231, 1, 2),
        C("beamsplitter", 0.4348, 0.0798, 0, 1),
        C("beamsplitter", 0.4368, 0.6157, 2, 3),
    ]
    return (f"gbs_d{D}", 4, D, True, calls)


def hamiltonian_simulation(pure):
    """tests/integration/test_algorithms.py:190-225, cutoff 4."""
    J, U, k, t = 1, 1.5, 20, 1.086
    theta = -J * t / k
    r = -U * t / (2 * k)
    calls = [C("prepare_fock_state", 2, 0)]
    for _ in range(k):
        calls += [
            C("beamsplitter", theta, np.pi / 2, 0, 1),
            C("kerr_interaction", r, 0),
            C("rotation", -r, 0),
            C("kerr_interaction", r, 1),
            C("rotation", -r, 1),
        ]
    return (f"hamiltonian_sim_{'pure' if pure else 'mixed'}", 2, 4, pure, calls)


def every_gate(n, D, pure, seed=11, prep=True):
    """Every gate of the path on every mode / ordered mode pair.  Pure-state
    pairs (m, 0) with m > 0 are skipped: the reference applies them to the
    wrong modes; they are covered on the mixed path."""
    rs = _rng(seed)
    calls = []
    for m in range(n):
        if prep:  # NB: a single-mode preparation makes the reference state mixed
            calls.append(C("prepare_coherent_state", float(rs.uniform(0.1, 0.5)), float(rs.uniform(0, 6)), m))
    for m in range(n):
        calls.append(C("displacement", float(rs.uniform(0.1, 0.4)), float(rs.uniform(0, 6)), m))
        calls.append(C("squeeze", float(rs.uniform(0.05, 0.3)), float(rs.uniform(0, 6)), m))
        calls.append(C("rotation", float(rs.uniform(-3, 3)), m))
        calls.append(C("kerr_interaction", float(rs.uniform(-1, 1)), m))
        calls.append(C("cubic_phase", float(rs.uniform(-0.05, 0.05)), m))
    for a in range(n):
        for b in range(n):
            if a == b or (pure and b == 0):
                continue
            calls.append(C("beamsplitter", float(rs.uniform(0, 1.5)), float(rs.uniform(0, 6)), a, b))
            calls.append(C("mzgate", float(rs.uniform(0, 6)), float(rs.uniform(0, 6)), a, b))
            calls.append(C("two_mode_squeeze", float(rs.uniform(0.05, 0.25)), float(rs.uniform(0, 6)), a, b))
            calls.append(C("cross_kerr_interaction", float(rs.uniform(-1, 1)), a, b))
    tag = ("pure" if pure else "mixed") + ("" if prep else "_noprep")
    return (f"every_gate_n{n}_d{D}_{tag}", n, D, pure, calls)


def loss_and_measure(n, D, seed=7):
    """BASELINE config 3 at reduced size: mixed state, Sgate/BSgate layers,
    LossChannel(0.9) on every mode, then MeasureFock on all modes with
    np.random.seed(7)."""
    rs = _rng(3)
    calls = []
    pairs = [(i, i + 1) for i in range(0, n - 1, 2)] + [(i, i + 1) for i in range(1, n - 1, 2)]
    for _ in range(2):
        for m in range(n):
            calls.append(C("squeeze", float(rs.uniform(0.1, 0.4)), float(rs.uniform(0, 6)), m))
        for a, b in pairs:
            calls.append(C("beamsplitter", float(rs.uniform(0, np.pi / 2)), float(rs.uniform(0, 6)), a, b))
    for m in range(n):
        calls.append(C("loss", 0.9, m))
    calls.append(C("seed", seed))
    calls.append(C("measure_fock", list(range(n))))
    return (f"loss_measure_n{n}_d{D}", n, D, False, calls)


def loss_edge_cases(D):
    """loss with T in {0, 0.37, 1} on pure and prepared states (ops.py:471-490)."""
    calls = [
        C("prepare_coherent_state", 0.7, 0.3, 0),
        C("prepare_fock_state", 2, 1),
        C("beamsplitter", 0.5, 0.2, 0, 1),
        C("loss", 0.37, 0),
        C("loss", 1.0, 1),
        C("displacement", 0.2, 0.1, 1),
        C("loss", 0.0, 1),
    ]
    return (f"loss_edges_d{D}", 2, D, True, calls)


def preparations(D):
    """prepare_* on mode subsets, unordered multimode kets/dms, add/del mode
    (circuit.py:373-535)."""
    rs = _rng(5)
    ket2 = rs.randn(D, D) + 1j * rs.randn(D, D)
    ket2 /= np.linalg.norm(ket2)
    v = rs.randn(D) + 1j * rs.randn(D)
    v /= np.linalg.norm(v)
    A = rs.randn(D * D, D * D) + 1j * rs.randn(D * D, D * D)
    rho2 = A @ A.conj().T
    rho2 /= np.trace(rho2)
    calls = [
        C("prepare_squeezed_state", 0.3, 0.4, 0),
        C("prepare_displaced_squeezed_state", 0.3, 0.2, 0.25, 0.9, 1),
        C("beamsplitter", 0.6, 0.3, 0, 1),
        C("prepare_thermal_state", 0.4, 2),
        C("prepare_ket_state", ket2, [2, 0]),
        C("mzgate", 0.4, 1.1, 1, 2),
        C("prepare_ket_state", v, 1),
        C("prepare_dm_state", np.ascontiguousarray(rho2.reshape([D] * 4).transpose(0, 2, 1, 3)), [1, 2]),
        C("add_mode", 1),
        C("prepare_fock_state", 1, 3),
        C("beamsplitter", 0.3, 0.0, 2, 3),
        C("del_mode", 1),
        C("rotation", 0.4, 3),
    ]
    return (f"preparations_d{D}", 3, D, True, calls)


def measure_postselect(D, pure):
    calls = [
        C("prepare_coherent_state", 0.6, 0.2, 0),
        C("two_mode_squeeze", 0.4, 0.3, 1, 2),
        C("beamsplitter", 0.7, 0.1, 0, 1),
        C("measure_fock", [1], select=[1]),
        C("displacement", 0.2, 0.0, 1),
        C("seed", 123),
        C("measure_fock", [2, 0]),
        C("squeeze", 0.2, 0.5, 2),
        C("seed", 5),
        C("measure_fock", [1]),
    ]
    return (f"measure_select_d{D}_{'pure' if pure else 'mixed'}", 3, D, pure, calls)


def homodyne_gkp(D, pure):
    """GKP preparation and homodyne measurements (circuit.py:713-812): a sampled outcome on an entangled
    mode, a post-selected one, and a second sampled one; ``num_bins`` keeps the host grid small."""
    calls = [
        C("prepare_gkp", [0.6, 0.4], 0.35, 1e-3, mode=0),
        C("squeeze", 0.3, 0.2, 1),
        C("beamsplitter", 0.6, 0.3, 0, 1),
        C("two_mode_squeeze", 0.2, 0.1, 1, 2),
        C("seed", 21),
        C("measure_homodyne", 0.4, 1, num_bins=5000),
        C("displacement", 0.2, 0.3, 2),
        C("measure_homodyne", 0.0, 0, select=0.37),
        C("rotation", 0.5, 2),
        C("seed", 4),
        C("measure_homodyne", 1.1, 2, num_bins=5000),
        C("beamsplitter", 0.3, 0.1, 2, 0),
    ]
    return (f"homodyne_gkp_d{D}_{'pure' if pure else 'mixed'}", 3, D, pure, calls)


def all_scripts():
    return [
        boson_sampling(5),
        boson_sampling(7),
        gbs(6),
        hamiltonian_simulation(True),
        hamiltonian_simulation(False),
        every_gate(3, 5, True),
        every_gate(3, 4, False),
        every_gate(3, 4, True, seed=12),
        every_gate(3, 6, True, seed=13, prep=False),
        every_gate(4, 5, True, seed=14, prep=False),
        loss_and_measure(3, 5),
        loss_and_measure(2, 10),
        loss_edge_cases(6),
        preparations(4),
        measure_postselect(6, True),
        measure_postselect(5, False),
        homodyne_gkp(8, True),
        homodyne_gkp(6, False),
    ]


# ---------------------------------------------------------------------------
# Observables of the state object (reference backends/states.py:657-985)
def observable_scripts():
    """A pure 3-mode state, a mixed one (single-mode preparations mix the reference state,
    SURVEY F7) and a lossy 2-mode state at the BASELINE cutoff."""
    lossy = loss_and_measure(2, 10)
    lossy = ("lossy_n2_d10", lossy[1], lossy[2], lossy[3],
             [c for c in lossy[4] if c[0] not in ("measure_fock", "seed")])
    return [every_gate(3, 6, True, seed=13, prep=False), every_gate(3, 4, False), lossy]


def observable_cases(script):
    """(key, method, args) evaluated on the final state of ``script``."""
    n = script[1]
    xs = np.linspace(-2.5, 2.5, 7)
    ps = np.linspace(-2.0, 2.0, 5)
    cases = [("fidelity_vacuum", "fidelity_vacuum", ())]
    alphas = [0.3 * np.exp(0.7j * (m + 1)) for m in range(n)]
    cases.append(("fidelity_coherent", "fidelity_coherent", (alphas,)))
    cases.append(("fidelity_coherent_0", "fidelity_coherent", ([0.0] * n,)))
    for m in range(n):
        cases.append((f"mean_photon_{m}", "mean_photon", (m,)))
        cases.append((f"quad_x_{m}", "quad_expectation", (m, 0.0)))
        cases.append((f"quad_phi_{m}", "quad_expectation", (m, 0.4 + m)))
        cases.append((f"wigner_{m}", "wigner", (m, xs, ps)))
        cases.append((f"number_{m}", "number_expectation", ([m],)))
        cases.append((f"parity_{m}", "parity_expectation", ([m],)))
    cases.append(("number_01", "number_expectation", ([0, 1],)))
    cases.append(("number_10", "number_expectation", ([1, 0],)))
    cases.append(("parity_all", "parity_expectation", (list(range(n)),)))
    if n > 2:
        cases.append(("number_02", "number_expectation", ([0, 2],)))
        cases.append(("parity_21", "parity_expectation", ([2, 1],)))
    # polynomials of quadratures: all modes, a mode subset (only modes 0 and n-1 appear), linear only
    rs = _rng(21)
    A = rs.randn(2 * n, 2 * n)
    A = 0.1 * (A + A.T)
    d = 0.3 * rs.randn(2 * n)
    cases.append(("polyquad_full", "poly_quad_expectation", (A, d, 0.3, 0.0)))
    cases.append(("polyquad_full_phi", "poly_quad_expectation", (A, d, 0.0, 0.37)))
    keep = [0, n - 1, n, 2 * n - 1]
    As, ds = np.zeros_like(A), np.zeros_like(d)
    As[np.ix_(keep, keep)] = A[np.ix_(keep, keep)]
    ds[keep] = d[keep]
    cases.append(("polyquad_subset", "poly_quad_expectation", (As, ds, -0.2, 0.9)))
    cases.append(("polyquad_linear", "poly_quad_expectation", (None, d, 0.0, 0.0)))
    return cases
