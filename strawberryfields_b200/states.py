"""State object returned by ``B200FockBackend.state()``.

Mirrors the Fock part of the reference's ``BaseFockState``
(``/root/reference/strawberryfields/backends/states.py:477-985``): ``ket``, ``dm``,
``trace``, ``all_fock_probs``, ``fock_prob``, ``reduced_dm``, ``mean_photon``,
``fidelity*``.  The state is a device snapshot; ``all_fock_probs``, ``reduced_dm``,
``trace`` and ``fock_prob`` are reductions computed on the GPU, and the full tensor is
copied to the host only when ``ket()``/``dm()``/``data`` is asked for.  When Strawberry
Fields itself is importable the class also derives from its ``BaseFockState`` so every
other observable of the reference (Wigner function, quadrature expectations ...) keeps
working on the host copy.
"""
from __future__ import annotations

import numpy as np

try:  # optional: only present when the reference package is installed next to us
    from strawberryfields.backends.states import BaseFockState as _SFBase  # type: ignore
except Exception:  # pragma: no cover - the GPU box has no strawberryfields
    _SFBase = None

_LETTERS = "abcdefghijklmnopqrstuvwxyz"


class _StandaloneBase:
    EQ_TOLERANCE = 1e-10

    def __init__(self, num_modes, mode_names=None):
        self._modes = num_modes
        self._hbar = 2
        self._data = None
        self._pure = None
        self._mode_names = mode_names or ["q[{}]".format(i) for i in range(num_modes)]

    @property
    def num_modes(self):
        return self._modes

    @property
    def is_pure(self):
        return self._pure

    @property
    def hbar(self):
        return self._hbar

    @property
    def mode_names(self):
        return dict(enumerate(self._mode_names))


class _FockStateMixin:
    """Device-backed overrides shared by the standalone and the SF-derived class."""

    def _init_device(self, circuit_view, pure, cutoff, batched):
        self._view = circuit_view  # frozen DeviceCircuit holding the snapshot
        self._pure = pure
        self._cutoff = cutoff
        self._batched = batched
        self._host = None
        self._basis = "fock"

    # -- data access -----------------------------------------------------------------
    @property
    def data(self):
        if self._host is None:
            self._host = self._view.host_state()
        return self._host

    @property
    def cutoff_dim(self):
        return self._cutoff

    @property
    def batched(self):
        return self._batched

    def ket(self, **kwargs):
        return self.data if self._pure else None

    def dm(self, **kwargs):
        if not self._pure:
            return self.data
        n = self._modes
        psi = self.data
        if self._batched:
            return np.stack([_outer_interleaved(p, n) for p in psi])
        return _outer_interleaved(psi, n)

    # -- device reductions -------------------------------------------------------------
    def trace(self, **kwargs):
        v = self._view._norm_device().cpu().numpy()
        return v if self._batched else float(v[0])

    def all_fock_probs(self, **kwargs):
        p = self._view.fock_probs_device().cpu().numpy()
        shape = [self._cutoff] * self._modes
        return p.reshape([p.shape[0]] + shape) if self._batched else p.reshape(shape)

    def fock_prob(self, n, **kwargs):
        if len(n) != self._modes:
            raise ValueError("List length should be equal to number of modes")
        if max(n) >= self._cutoff:
            raise ValueError("Can't get distribution beyond truncation level")
        vals = self._view.element(n)
        res = np.abs(vals) ** 2 if self._pure else vals.real
        return res if self._batched else res[0]

    def reduced_dm(self, modes, **kwargs):
        if isinstance(modes, int):
            modes = [modes]
        modes = list(modes)
        if modes == list(range(self._modes)):
            return self.dm()
        if modes != sorted(modes):
            raise ValueError("The specified modes cannot be duplicated.")
        if len(modes) > self._modes:
            raise ValueError(
                "The number of specified modes cannot be larger than the number of subsystems."
            )
        r = self._view.reduced_dm_device(modes).cpu().numpy()
        shape = [self._cutoff] * (2 * len(modes))
        return r.reshape([r.shape[0]] + shape) if self._batched else r.reshape(shape)

    def mean_photon(self, mode, **kwargs):
        n = np.arange(self._cutoff)
        probs = np.diagonal(self.reduced_dm(mode), axis1=-2, axis2=-1)
        mean = np.sum(n * probs, axis=-1).real
        var = np.sum(n ** 2 * probs, axis=-1).real - mean ** 2
        return mean, var

    def fidelity(self, other_state, mode, **kwargs):
        rho = self.reduced_dm([mode])
        other = np.asarray(other_state)
        if other.ndim == 1:
            return np.einsum("i,...ij,j->...", other.conj(), rho, other).real
        return np.einsum("ij,...ji->...", other, rho).real  # valid when either state is pure

    def fidelity_vacuum(self, **kwargs):
        return self.fock_prob([0] * self._modes)

    def fidelity_coherent(self, alpha_list, **kwargs):
        """|<alpha|psi>|^2 / <alpha|rho|alpha> (states.py:687-728), contracted on the device."""
        if not hasattr(alpha_list, "__len__"):
            alpha_list = [alpha_list]
        if len(alpha_list) != self._modes:
            raise ValueError("The number of alpha values must match the number of modes.")
        ns = np.arange(self._cutoff)
        sqrt_fact = np.sqrt(np.cumprod(np.concatenate([[1.0], ns[1:]])))
        vecs = [np.exp(-0.5 * np.abs(a) ** 2) * np.asarray(a, dtype=complex) ** ns / sqrt_fact for a in alpha_list]
        ov = self._view.product_overlap_device(vecs).cpu().numpy()
        res = np.abs(ov) ** 2 if self._pure else ov.real
        return res if self._batched else res[0]

    def diagonal_expectation(self, modes, values):
        """<prod_k f(n_k)> for an operator diagonal in the number basis (states.py:920-945): the
        joint photon-number distribution of ``modes`` is reduced on the device (D^k doubles come
        back), never the state."""
        modes = list(modes)
        if len(modes) != len(set(modes)):
            raise ValueError("There can be no duplicates in the modes specified.")
        values = np.asarray(values)
        ps = self._view.marginal_probs_device(sorted(modes)).cpu().numpy()
        ps = ps.reshape([ps.shape[0]] + [self._cutoff] * len(modes))
        for _ in modes:
            ps = np.tensordot(ps, values, axes=([1], [0]))
        return ps if self._batched else float(ps[0])

    def sample_fock(self, shots, modes=None):
        """``shots`` photon-number samples of ``modes`` (default: all) WITHOUT collapsing the state:
        int64 array (shots, len(modes)).  The reference's ``measure_fock`` refuses ``shots != 1``
        (fockbackend/backend.py:290-295); here the joint distribution is reduced once on the device
        (D^k doubles come back) and numpy's global stream draws all shots from it, with the same
        1e-8 clipping and renormalisation as a single measurement (circuit.py:678-686)."""
        if self._batched:
            raise NotImplementedError("sample_fock is not available for batched states")
        modes = list(range(self._modes)) if modes is None else list(modes)
        if len(modes) != len(set(modes)) or any(not 0 <= m < self._modes for m in modes):
            raise ValueError("The specified modes are not valid.")
        D, k = self._cutoff, len(modes)
        if D ** k > 1 << 24:
            raise NotImplementedError("sample_fock: the joint distribution of %d modes is too large" % k)
        order = sorted(modes)
        dist = self._view.marginal_probs_device(order).cpu().numpy().reshape(-1)
        dist = dist * ~np.isclose(dist, 0.0)
        idx = np.random.choice(len(dist), size=int(shots), p=dist / dist.sum())
        digits = np.stack(np.unravel_index(idx, [D] * k), axis=1)          # columns in ascending-mode order
        return digits[:, [order.index(m) for m in modes]].astype(np.int64)

    def number_expectation(self, modes):
        values = np.arange(self._cutoff)
        mean = self.diagonal_expectation(modes, values)
        var = self.diagonal_expectation(modes, values ** 2) - mean ** 2
        return mean, var

    def parity_expectation(self, modes):
        return self.diagonal_expectation(modes, (-1) ** np.arange(self._cutoff))

    def quad_expectation(self, mode, phi=0, **kwargs):
        """Mean and variance of x_phi on one mode (states.py:787-804) from its device-reduced
        density matrix; the ladder operators are built 5 levels above the cutoff and truncated, as
        the reference does, so x_phi^2 has the reference's boundary terms."""
        D = self._cutoff
        a = np.diag(np.sqrt(np.arange(1, D + 5)), 1)
        x = np.sqrt(self._hbar / 2) * (a + a.T)
        p = -1j * np.sqrt(self._hbar / 2) * (a - a.T)
        xphi = np.cos(phi) * x + np.sin(phi) * p
        xphisq = (xphi @ xphi)[:D, :D]
        xphi = xphi[:D, :D]
        rho = self.reduced_dm([mode])
        mean = np.einsum("ij,...ji->...", xphi, rho).real
        var = np.einsum("ij,...ji->...", xphisq, rho).real - mean ** 2
        return mean, var

    def poly_quad_expectation(self, A, d=None, k=0, phi=0, **kwargs):
        """Mean and variance of P = r^T A r + r^T d + k with r = (x_1..x_N, p_1..p_N)
        (states.py:806-918).  Only the modes P acts on are kept: their reduced density matrix comes
        from the device, the operator is built on the host one level above the cutoff (as the
        reference does), squared there and truncated."""
        N, D = self._modes, self._cutoff
        if A is None:
            A = np.zeros([2 * N, 2 * N])
        A = np.asarray(A)
        if A.shape != (2 * N, 2 * N):
            raise ValueError("Matrix of quadratic coefficients A must be of size 2Nx2N.")
        if not np.allclose(A.T, A):
            raise ValueError("Matrix of quadratic coefficients A must be symmetric.")
        d = np.zeros([2 * N]) if d is None else np.asarray(d)
        if d.shape != (2 * N,):
            raise ValueError("Vector of linear coefficients d must be of length 2N.")
        if self._batched:
            raise NotImplementedError("poly_quad_expectation is not available for batched states")
        modes = sorted({int(i) % N for i in A.nonzero()[0]} | {int(i) % N for i in d.nonzero()[0]})
        if not modes:
            return k, 0.0
        m, dim = len(modes), D + 1
        a = np.diag(np.sqrt(np.arange(1, dim)), 1)
        x0 = np.sqrt(self._hbar / 2) * (a + a.T)
        p0 = -1j * np.sqrt(self._hbar / 2) * (a - a.T)
        x1 = np.cos(phi) * x0 + np.sin(phi) * p0   # x_phi, as in quad_expectation
        p1 = -np.sin(phi) * x0 + np.cos(phi) * p0  # its conjugate quadrature

        def on_mode(op, j):
            out = np.ones((1, 1), dtype=complex)
            for q in range(m):
                out = np.kron(out, op if q == j else np.eye(dim))
            return out

        R = [on_mode(x1, j) for j in range(m)] + [on_mode(p1, j) for j in range(m)]
        rows = modes + [N + q for q in modes]
        Ar, dr = A[np.ix_(rows, rows)], d[rows]
        P = k * np.eye(dim ** m, dtype=complex)
        for i in range(2 * m):
            if dr[i] != 0:
                P = P + dr[i] * R[i]
            for j in range(2 * m):
                if Ar[i, j] != 0:
                    P = P + Ar[i, j] * (R[i] @ R[j])
        Psq = P @ P

        def cut(M):
            sl = tuple([slice(0, D)] * (2 * m))
            return M.reshape([dim] * (2 * m))[sl].reshape(D ** m, D ** m)

        rho = self.reduced_dm(modes)
        rho = rho.transpose(list(range(0, 2 * m, 2)) + list(range(1, 2 * m, 2))).reshape(D ** m, D ** m)
        mean = np.trace(cut(P) @ rho).real
        var = np.trace(cut(Psq) @ rho).real - mean ** 2
        return mean, var

    def wigner(self, mode, xvec, pvec):
        """Discretised Wigner function of one mode, [len(pvec), len(xvec)] like the reference's
        (states.py:730-785), from the device-reduced density matrix:
        W = sum_mn rho[m, n] W_mn with the |m><n| transforms generated by the ladder recurrences
        W_0n = 2A W_0,n-1 / sqrt(n),  W_mn = (2A* W_m-1,n - sqrt(n) W_m-1,n-1) / sqrt(m)."""
        if self._batched:
            raise NotImplementedError("wigner is not available for batched states")
        D = self._cutoff
        rho = self.reduced_dm([mode])
        Q, P = np.meshgrid(xvec, pvec)
        A = (Q + 1j * P) / (2 * np.sqrt(self._hbar / 2))
        prev = [np.exp(-2.0 * np.abs(A) ** 2) / np.pi + 0j]
        for n in range(1, D):
            prev.append(2.0 * A * prev[n - 1] / np.sqrt(n))
        W = sum((rho[0, n] * prev[n]).real * (1 if n == 0 else 2) for n in range(D))
        for m in range(1, D):  # prev[n] = W_{m-1,n} for n >= m-1; the lower triangle follows by symmetry
            row = [None] * D
            for n in range(m, D):
                row[n] = (2 * np.conj(A) * prev[n] - np.sqrt(n) * prev[n - 1]) / np.sqrt(m)
                W = W + (rho[m, n] * row[n]).real * (1 if n == m else 2)
            prev = row
        return W / self._hbar

    def __repr__(self):
        return "<B200FockState: num_modes={}, cutoff={}, pure={}, hbar={}>".format(
            self._modes, self._cutoff, self._pure, self._hbar)


def _outer_interleaved(psi, n):
    rho = np.multiply.outer(psi, psi.conj())
    return np.ascontiguousarray(rho.transpose([x for i in range(n) for x in (i, n + i)]))


if _SFBase is not None:

    class B200FockState(_FockStateMixin, _SFBase):
        def __init__(self, circuit_view, num_modes, pure, cutoff_dim, mode_names=None, batched=False):
            _SFBase.__init__(self, None, num_modes, pure, cutoff_dim, mode_names)
            self._init_device(circuit_view, pure, cutoff_dim, batched)

else:

    class B200FockState(_FockStateMixin, _StandaloneBase):
        def __init__(self, circuit_view, num_modes, pure, cutoff_dim, mode_names=None, batched=False):
            _StandaloneBase.__init__(self, num_modes, mode_names)
            self._init_device(circuit_view, pure, cutoff_dim, batched)
