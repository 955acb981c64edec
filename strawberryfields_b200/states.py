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
