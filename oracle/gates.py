"""ORACLE (test infrastructure, not product code): Fock-basis gate tensors on the CPU.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
reference arm may import this module.  The product path (``strawberryfields_b200``)
never does.

What is restated here
---------------------
The reference builds its D, S, BS, MZ and S2 gate tensors by calling the
third-party package ``thewalrus`` (pinned ``thewalrus==0.22.0`` in
``/root/reference/requirements.txt:12``; call sites
``strawberryfields/backends/fockbackend/ops.py:233,252,266,326,340``).  The source
of ``thewalrus.fock_gradients`` is NOT vendored in the reference tree, so the
published recursions (Quesada et al., "Fast optimization of parametrized quantum
optical circuits", Quantum 4, 366 (2020)) are restated here
in plain numpy loops.  They give the exact infinite-dimensional matrix elements
truncated to ``D`` -- not ``expm`` of truncated generators.

Pinning: ``tests/test_oracle_gates.py`` checks every tensor against
``scipy.linalg.expm`` of the generators documented in the reference front end
(``strawberryfields/ops.py:1497-1514`` D, ``1612-1626`` S, ``1866-1885`` BS,
``2037-2051`` S2, ``1956-1973`` MZ) built in a much larger Fock space and cropped,
and against closed forms the reference's own tests use
(``tests/backend/test_displacement_operation.py:74-94``,
``tests/backend/test_twomode_squeezing_operation.py:32-46``).

The remaining (non-thewalrus) constructors follow
``strawberryfields/backends/fockbackend/ops.py`` line by line in meaning:
``phase`` 309-314, ``kerr`` 274-281, ``cross_kerr`` 284-293, ``cubic_phase``
296-306, ``loss_kraus`` 471-490, state vectors 383-461.
"""
from __future__ import annotations

import functools
from math import factorial

import numpy as np
from scipy.linalg import expm

C128 = np.complex128


def _sqrt_table(D):
    return np.sqrt(np.arange(D + 1, dtype=np.float64))


# ----------------------------------------------------------------------------
# single-mode gates ([out, in])
# ----------------------------------------------------------------------------
@functools.lru_cache(maxsize=512)
def displacement(r, phi, D):
    """D(alpha), alpha = r e^{i phi}.  thewalrus.fock_gradients.displacement
    (called at fockbackend/ops.py:233)."""
    sq = _sqrt_table(D)
    alpha = r * np.exp(1j * phi)
    nac = -np.conj(alpha)
    T = np.zeros((D, D), dtype=C128)
    T[0, 0] = np.exp(-0.5 * r * r)
    for m in range(1, D):
        T[m, 0] = alpha / sq[m] * T[m - 1, 0]
    for m in range(D):
        for n in range(1, D):
            below = T[m - 1, n - 1] if m > 0 else 0.0
            T[m, n] = nac / sq[n] * T[m, n - 1] + sq[m] / sq[n] * below
    T.setflags(write=False)
    return T


@functools.lru_cache(maxsize=512)
def squeezing(r, theta, D):
    """S(z), z = r e^{i theta}.  thewalrus.fock_gradients.squeezing
    (called at fockbackend/ops.py:252)."""
    sq = _sqrt_table(D)
    t = np.exp(1j * theta) * np.tanh(r)
    s = 1.0 / np.cosh(r)
    T = np.zeros((D, D), dtype=C128)
    T[0, 0] = np.sqrt(s)
    for m in range(2, D, 2):
        T[m, 0] = -t * sq[m - 1] / sq[m] * T[m - 2, 0]
    for m in range(D):
        for n in range(1, D):
            if (m + n) % 2:
                continue
            left = T[m, n - 2] if n >= 2 else 0.0
            diag = T[m - 1, n - 1] if m > 0 else 0.0
            T[m, n] = sq[n - 1] / sq[n] * np.conj(t) * left + sq[m] / sq[n] * s * diag
    T.setflags(write=False)
    return T


# ----------------------------------------------------------------------------
# two-mode gates, thewalrus index order [out1, out2, in1, in2]
# ----------------------------------------------------------------------------
def _passive_two_mode(u00, u01, u10, u11, D):
    """Fock tensor of the passive two-mode unitary whose 2x2 mode transformation
    is [[u00, u01], [u10, u11]]; shared body of beamsplitter and mzgate
    (rank-3 loop uses (u00, u10), rank-4 loop (u01, u11))."""
    sq = _sqrt_table(D)
    Z = np.zeros((D + 1,) * 4, dtype=C128)  # index -1 lands on the zero pad
    Z[0, 0, 0, 0] = 1.0
    for m in range(D):
        for n in range(D - m):
            p = m + n
            if 0 < p < D:
                Z[m, n, p, 0] = (
                    u00 * sq[m] / sq[p] * Z[m - 1, n, p - 1, 0]
                    + u10 * sq[n] / sq[p] * Z[m, n - 1, p - 1, 0]
                )
    for m in range(D):
        for n in range(D):
            for p in range(D):
                q = m + n - p
                if 0 < q < D:
                    Z[m, n, p, q] = (
                        u01 * sq[m] / sq[q] * Z[m - 1, n, p, q - 1]
                        + u11 * sq[n] / sq[q] * Z[m, n - 1, p, q - 1]
                    )
    return np.ascontiguousarray(Z[:D, :D, :D, :D])


@functools.lru_cache(maxsize=512)
def beamsplitter_tw(theta, phi, D):
    """thewalrus.fock_gradients.beamsplitter (called at fockbackend/ops.py:326)."""
    c = np.cos(theta)
    s = np.sin(theta) * np.exp(1j * phi)
    Z = _passive_two_mode(c, -np.conj(s), s, c, D)
    Z.setflags(write=False)
    return Z


@functools.lru_cache(maxsize=512)
def mzgate_tw(phi_in, phi_ex, D):
    """thewalrus.fock_gradients.mzgate (called at fockbackend/ops.py:340)."""
    v = np.exp(1j * phi_in)
    u = np.exp(1j * phi_ex)
    v1 = (v - 1) * u / 2
    v2 = 1j * (v + 1) / 2
    v3 = 1j * (v + 1) * u / 2
    v4 = (1 - v) / 2
    Z = _passive_two_mode(v1, v2, v3, v4, D)
    Z.setflags(write=False)
    return Z


@functools.lru_cache(maxsize=512)
def two_mode_squeezing_tw(r, theta, D):
    """thewalrus.fock_gradients.two_mode_squeezing (called at fockbackend/ops.py:266)."""
    sq = _sqrt_table(D)
    s = 1.0 / np.cosh(r)
    e = np.exp(1j * theta) * np.tanh(r)
    Z = np.zeros((D + 1,) * 4, dtype=C128)
    Z[0, 0, 0, 0] = s
    for n in range(1, D):
        Z[n, n, 0, 0] = e * Z[n - 1, n - 1, 0, 0]
    for m in range(D):
        for n in range(m):
            p = m - n
            if 0 < p < D:
                Z[m, n, p, 0] = s * sq[m] / sq[p] * Z[m - 1, n, p - 1, 0]
    for m in range(D):
        for n in range(D):
            for p in range(D):
                q = p - (m - n)
                if 0 < q < D:
                    Z[m, n, p, q] = (
                        s * sq[n] / sq[q] * Z[m, n - 1, p, q - 1]
                        - np.conj(e) * sq[p] / sq[q] * Z[m, n, p - 1, q - 1]
                    )
    Z = np.ascontiguousarray(Z[:D, :D, :D, :D])
    Z.setflags(write=False)
    return Z


# SF convention [out1, in1, out2, in2]  (fockbackend/ops.py:269,329,343)
def beamsplitter(theta, phi, D):
    return beamsplitter_tw(theta, phi, D).transpose(0, 2, 1, 3)


def mzgate(phi_in, phi_ex, D):
    return mzgate_tw(phi_in, phi_ex, D).transpose(0, 2, 1, 3)


def two_mode_squeeze(r, theta, D):
    return two_mode_squeezing_tw(r, theta, D).transpose(0, 2, 1, 3)


# ----------------------------------------------------------------------------
# diagonal / host-built gates  (fockbackend/ops.py:274-314)
# ----------------------------------------------------------------------------
def phase(theta, D):
    return np.diag(np.exp(1j * theta * np.arange(D))).astype(C128)


def kerr(kappa, D):
    n = np.arange(D)
    return np.diag(np.exp(1j * kappa * n**2)).astype(C128)


def cross_kerr(kappa, D):
    """[out1, in1, out2, in2] tensor of exp(i kappa n1 n2)."""
    T = np.zeros((D,) * 4, dtype=C128)
    for a in range(D):
        for b in range(D):
            T[a, a, b, b] = np.exp(1j * kappa * a * b)
    return T


def annihilation(D):
    A = np.zeros((D, D), dtype=C128)
    for i in range(1, D):
        A[i - 1, i] = np.sqrt(i)
    return A


def cubic_phase(gamma, hbar, D):
    """expm(i gamma x^3 / (3 hbar)) with the TRUNCATED x (fockbackend/ops.py:296-306)."""
    a = annihilation(D)
    x = (a + a.conj().T) * np.sqrt(hbar / 2)
    return expm(1j * gamma / (3 * hbar) * (x @ x @ x))


def loss_kraus(T, D):
    """Kraus operators of the loss channel (fockbackend/ops.py:471-490)."""
    if T == 0:
        out = []
        for i in range(D):
            P = np.zeros((D, D), dtype=C128)
            P[0, i] = 1.0
            out.append(P)
        return out
    a = annihilation(D)
    damp = np.diag([T ** (i / 2) for i in range(D)]).astype(C128)
    ops = []
    for n in range(D):
        ops.append(
            ((1 - T) / T) ** (n / 2) * (np.linalg.matrix_power(a, n) / np.sqrt(factorial(n))) @ damp
        )
    return ops


# ----------------------------------------------------------------------------
# single-mode state vectors (fockbackend/ops.py:383-461)
# ----------------------------------------------------------------------------
def fock_state(n, D):
    v = np.zeros(D, dtype=C128)
    v[n] = 1.0
    return v


def coherent_state(r, phi, D):
    alpha = r * np.exp(1j * phi)
    return np.exp(-abs(alpha) ** 2 / 2) * np.array(
        [alpha**n / np.sqrt(factorial(n)) for n in range(D)], dtype=C128
    )


def squeezed_state(r, theta, D):
    v = np.zeros(D, dtype=C128)
    for n in range(0, D, 2):
        k = n // 2
        v[n] = (np.sqrt(factorial(2 * k)) / (2**k * factorial(k))) * (
            -np.exp(1j * theta) * np.tanh(r)
        ) ** k
    return np.sqrt(1 / np.cosh(r)) * v


def displaced_squeezed_state(r_d, phi_d, r_s, phi_s, D):
    from numpy.polynomial.hermite import hermval

    if np.allclose(r_s, 0.0):
        return coherent_state(r_d, phi_d, D)
    if np.allclose(r_d, 0.0):
        return squeezed_state(r_s, phi_s, D)
    ph = np.exp(1j * phi_s)
    ch, sh, th = np.cosh(r_s), np.sinh(r_s), np.tanh(r_s)
    alpha = r_d * np.exp(1j * phi_d)
    gamma = alpha * ch + np.conj(alpha) * ph * sh
    harg = gamma / np.sqrt(ph * np.sinh(2 * r_s) + 1e-10)
    N = np.exp(-0.5 * np.abs(alpha) ** 2 - 0.5 * np.conj(alpha) ** 2 * ph * th)
    coeff = np.array(
        [(0.5 * ph * th) ** (n / 2) / np.sqrt(factorial(n) * ch) for n in range(D)]
    )
    vec = np.array([hermval(harg, row) for row in np.diag(coeff)])
    return N * vec


def square_gkp_state(theta, phi, epsilon, ampl_cutoff, D):
    """fockbackend/ops.py:518-596: cos(theta/2)|0>_gkp + e^{-i phi} sin(theta/2)|1>_gkp, each basis
    state a comb of displaced squeezed states (teeth t = -z_max..z_max at sqrt(pi/2)(2t+k)/cosh(eps),
    weights exp(-pi/2 tanh(eps) (k+2t)^2), squeezing r = -log(tanh eps)/2), normalised at the end."""
    z_max = int(np.ceil(np.sqrt(-0.25 / np.pi * np.log(ampl_cutoff) / np.tanh(epsilon))))
    r = -0.5 * np.log(np.tanh(epsilon))
    basis = []
    for k in (0, 1):
        ket = np.zeros(D, dtype=C128)
        for t in range(-z_max, z_max + 1):
            ket = ket + np.exp(-0.5 * np.pi * np.tanh(epsilon) * (k + 2 * t) ** 2) * displaced_squeezed_state(
                np.sqrt(0.5 * np.pi) * (2 * t + k) / np.cosh(epsilon), 0, r, 0, D)
        basis.append(ket)
    ket = np.cos(theta / 2) * basis[0] + np.sin(theta / 2) * np.exp(-1j * phi) * basis[1]
    return ket / np.linalg.norm(ket)


def thermal_state(nbar, D):
    if nbar == 0:
        v = fock_state(0, D)
        return np.outer(v, v.conj())
    return np.diag([nbar**n / (nbar + 1) ** (n + 1) for n in range(D)]).astype(C128)
