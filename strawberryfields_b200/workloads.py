"""Synthetic circuits of the BASELINE configs, as backend call lists.

A circuit is a list of ``(method, *args)`` tuples -- the calls
``LocalEngine._run_program`` would make on a ``BaseFock`` backend
(``/root/reference/strawberryfields/engine.py:422-457``) after the ``fock`` compiler has
decomposed the program.  ``Interferometer(U)`` is emitted in exactly the gate order the
reference front end produces for the default rectangular (Clements) mesh
(``strawberryfields/ops.py:2655-2718`` + ``decompositions.rectangular``): a sweep of
``Rgate, BSgate(theta, 0)`` pairs along the even anti-diagonals, one ``Rgate`` per mode,
then ``BSgate, Rgate`` pairs along the odd anti-diagonals.  Angles are drawn at random
(synthetic data); tests/golden/interferometer_n*.json hold reference-compiled instances
of the same structure.
"""
from __future__ import annotations

import numpy as np


def rectangular_mesh_pairs(N):
    """(first, second) sweeps of mode pairs of the N-mode rectangular mesh."""
    first = [(i - j, i - j + 1) for i in range(0, N - 1, 2) for j in range(i + 1)]
    second = [(N + j - i - 2, N + j - i - 1) for i in range(1, N - 1, 2) for j in range(i + 1)]
    return first, list(reversed(second))


def interferometer_calls(N, rng):
    first, second = rectangular_mesh_pairs(N)
    calls = []
    for a, b in first:
        calls.append(("rotation", float(rng.uniform(-np.pi, np.pi)), a))
        calls.append(("beamsplitter", float(rng.uniform(0, np.pi / 2)), 0.0, a, b))
    for m in range(N):
        calls.append(("rotation", float(rng.uniform(0, 2 * np.pi)), m))
    for a, b in second:
        calls.append(("beamsplitter", float(-rng.uniform(0, np.pi / 2)), 0.0, a, b))
        calls.append(("rotation", float(rng.uniform(-np.pi, np.pi)), a))
    return calls


def interferometer_unitary(N, calls):
    """N x N mode transformation of a passive call list (Rgate / BSgate(theta, phi))."""
    U = np.eye(N, dtype=complex)
    for c in calls:
        if c[0] == "rotation":
            G = np.eye(N, dtype=complex)
            G[c[2], c[2]] = np.exp(1j * c[1])
        elif c[0] == "beamsplitter":
            theta, phi, a, b = c[1:]
            G = np.eye(N, dtype=complex)
            ct, st = np.cos(theta), np.sin(theta)
            G[a, a] = ct
            G[a, b] = -np.exp(-1j * phi) * st
            G[b, a] = np.exp(1j * phi) * st
            G[b, b] = ct
        else:
            raise ValueError("not a passive gate: %s" % c[0])
        U = G @ U
    return U


def config1_circuit():
    """BASELINE config 1: ``examples/boson_sampling.py:5-41`` (fixed literals): Fock inputs
    |1,1,0,1>, four Rgates, eight BSgates on 4 modes; the result is ``all_fock_probs()``."""
    calls = [("prepare_fock_state", 1, 0), ("prepare_fock_state", 1, 1), ("prepare_vacuum_state", 2),
             ("prepare_fock_state", 1, 3)]
    for m, t in enumerate((0.5719, -1.9782, 2.0603, 0.0644)):
        calls.append(("rotation", t, m))
    for t, p, a, b in ((0.7804, 0.8578, 0, 1), (0.06406, 0.5165, 2, 3), (0.473, 0.1176, 1, 2),
                       (0.563, 0.1517, 0, 1), (0.1323, 0.9946, 2, 3), (0.311, 0.3231, 1, 2),
                       (0.4348, 0.0798, 0, 1), (0.4368, 0.6157, 2, 3)):
        calls.append(("beamsplitter", t, p, a, b))
    return calls


def config2_circuit(N=8, seed=42):
    """BASELINE config 2 / 5: Sgate + Dgate on every mode, then a random N-mode
    interferometer.  N=8: 8 S + 8 D + 36 R + 28 BS = 80 gates."""
    rng = np.random.RandomState(seed)
    calls = []
    r = rng.uniform(0, 0.3, N)
    pr = rng.uniform(0, 2 * np.pi, N)
    a = rng.uniform(0, 0.5, N)
    pa = rng.uniform(0, 2 * np.pi, N)
    for i in range(N):
        calls.append(("squeeze", float(r[i]), float(pr[i]), i))
        calls.append(("displacement", float(a[i]), float(pa[i]), i))
    calls += interferometer_calls(N, rng)
    return calls


def config3_circuit(N=4, seed=42):
    """BASELINE config 3: mixed state, two layers of Sgate on all modes + BSgate on
    (0,1),(2,3),(1,2), LossChannel(0.9) on every mode, then MeasureFock on all modes."""
    rng = np.random.RandomState(seed)
    pairs = [(i, i + 1) for i in range(0, N - 1, 2)] + [(i, i + 1) for i in range(1, N - 1, 2)]
    calls = []
    for _ in range(2):
        for m in range(N):
            calls.append(("squeeze", float(rng.uniform(0.05, 0.3)), float(rng.uniform(0, 2 * np.pi)), m))
        for a, b in pairs:
            calls.append(("beamsplitter", float(rng.uniform(0, np.pi / 2)), float(rng.uniform(0, 2 * np.pi)), a, b))
    for m in range(N):
        calls.append(("loss", 0.9, m))
    return calls


def config4_circuit(N=6, batch=64, seed=42):
    """BASELINE config 4: one CV quantum-neural-network layer
    (``examples/quantum_neural_network.py:14-85`` shape): interferometer, Sgate,
    interferometer, Dgate, Kgate, with per-batch-element weights (arrays of length
    ``batch``).  The QNN interferometer is BS(theta, phi) on the rectangular layers
    followed by N-1 rotations (15 BS + 5 R for N = 6)."""
    rng = np.random.RandomState(seed)

    def interferometer():
        calls = []
        for layer in range(N):
            for k, (a, b) in enumerate(zip(range(N - 1), range(1, N))):
                if (layer + k) % 2 != 1:
                    calls.append(("beamsplitter", rng.normal(0, 0.1, batch), rng.normal(0, 0.1, batch), a, b))
        for m in range(N - 1):
            calls.append(("rotation", rng.normal(0, 0.1, batch), m))
        return calls

    calls = interferometer()
    for m in range(N):
        calls.append(("squeeze", rng.normal(0, 0.05, batch), 0.0, m))
    calls += interferometer()
    for m in range(N):
        calls.append(("displacement", np.abs(rng.normal(0, 0.05, batch)), rng.uniform(0, 2 * np.pi, batch), m))
        calls.append(("kerr_interaction", rng.normal(0, 0.05, batch), m))
    return calls


def run_calls(backend, calls):
    for c in calls:
        getattr(backend, c[0])(*c[1:])


def count_updates(calls, elements):
    """amp-gate updates of a call list on a state of ``elements`` stored entries
    (one update = one stored complex128 entry passing through one gate, fused or not)."""
    return len(calls) * elements
