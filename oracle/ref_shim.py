"""ORACLE helper (test infrastructure): run the UNMODIFIED reference fock backend.

``/root/reference`` is a pure-Python package whose import needs third-party
modules that are not installed and cannot be fetched (``thewalrus``,
``blackbird``, ``xir``, ``xcc``).  :func:`install` puts inert
stand-ins for them into ``sys.modules`` and binds the five
``thewalrus.fock_gradients`` functions the fock backend actually calls
(``strawberryfields/backends/fockbackend/ops.py:30-36``) to the restated
recursions in :mod:`oracle.gates`.  After that ``import strawberryfields`` works
and ``sf.Engine("fock")`` / ``FockBackend`` / ``Circuit`` run the reference's own
code, unmodified, from where it lies.

In the build container the package is imported from ``/root/reference``; on the GPU box from
``oracle/_ref``, the copy installed by ``oracle/build_ref.py``.  It is used by
``oracle/make_golden.py`` to write the fixtures in ``tests/golden``, by the container-only
differential tests, and by the reference arm of ``bench.py`` (``--impl reference``), which times the
reference's own fock backend on the host cores.
"""
from __future__ import annotations

import os
import sys
import types
from unittest import mock

# where the unmodified reference package lies: the read-only tree of the build container, else the copy
# installed by oracle/build_ref.py (oracle/_ref: git-ignored, travels to the GPU box with the snapshot)
_INSTALLED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get(
    "SF_REFERENCE_ROOT",
    "/root/reference" if os.path.isdir("/root/reference/strawberryfields") else _INSTALLED)

_STUBS = [
    "thewalrus",
    "thewalrus.fock_gradients",
    "thewalrus.symplectic",
    "thewalrus.quantum",
    "thewalrus.quantum.fock_tensors",
    "thewalrus.samples",
    "thewalrus.csamples",
    "thewalrus.random",
    "thewalrus._hafnian",
    "thewalrus._torontonian",
    "thewalrus._hermite_multidimensional",
    "blackbird",
    "blackbird.utils",
    "blackbird.error",
    "blackbird.listener",
    "blackbird.program",
    "xir",
    "xcc",
]


class _Stub(types.ModuleType):
    __version__ = "0.0.0-stub"
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "strawberryfields"))


def install():
    """Make ``import strawberryfields`` resolve to the reference tree. Idempotent."""
    if "strawberryfields" in sys.modules and getattr(
        sys.modules["strawberryfields"], "_b200_oracle_shim", False
    ):
        return sys.modules["strawberryfields"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    for name in _STUBS:
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[name])

    class BlackbirdSyntaxError(Exception):
        pass

    class TemplateError(Exception):
        pass

    sys.modules["blackbird.error"].BlackbirdSyntaxError = BlackbirdSyntaxError
    sys.modules["blackbird.utils"].TemplateError = TemplateError

    from oracle import gates

    fg = sys.modules["thewalrus.fock_gradients"]
    fg.displacement = lambda r, phi, cutoff: gates.displacement(float(r), float(phi), int(cutoff)).copy()
    fg.squeezing = lambda r, theta, cutoff: gates.squeezing(float(r), float(theta), int(cutoff)).copy()
    fg.beamsplitter = lambda theta, phi, cutoff: gates.beamsplitter_tw(
        float(theta), float(phi), int(cutoff)
    ).copy()
    fg.mzgate = lambda phi_in, phi_ex, cutoff: gates.mzgate_tw(
        float(phi_in), float(phi_ex), int(cutoff)
    ).copy()
    fg.two_mode_squeezing = lambda r, theta, cutoff: gates.two_mode_squeezing_tw(
        float(r), float(theta), int(cutoff)
    ).copy()

    # thewalrus.symplectic: the four index / rotation helpers the reference's state classes and its own tests
    # import (backends/states.py:29-30, decompositions.py:25, tests/backend/test_states_polyquad.py:22-23).
    # Their published definitions (thewalrus 0.22 docs): R(theta) = [[cos, -sin], [sin, cos]];
    # Omega = [[0, I], [-I, 0]]; xpxp <-> xxpp reorders (x_1, p_1, .., x_N, p_N) <-> (x_1..x_N, p_1..p_N).
    import numpy as np

    sy = sys.modules["thewalrus.symplectic"]

    def rotation(theta):
        c, s = np.cos(theta), np.sin(theta)
        return np.array([[c, -s], [s, c]])

    def sympmat(N, dtype=np.float64):
        eye, zero = np.identity(N, dtype=dtype), np.zeros((N, N), dtype=dtype)
        return np.block([[zero, eye], [-eye, zero]])

    def _reorder(S, ind):
        S = np.asarray(S)
        if S.shape[0] % 2:
            raise ValueError("The input array is not even-dimensional")
        if S.ndim == 2 and S.shape[0] != S.shape[1]:
            raise ValueError("The input matrix is not square")
        return S[ind] if S.ndim == 1 else S[:, ind][ind]

    def xpxp_to_xxpp(S):
        n = np.asarray(S).shape[0]
        return _reorder(S, np.concatenate([np.arange(0, n, 2), np.arange(1, n, 2)]))

    def xxpp_to_xpxp(S):
        n = np.asarray(S).shape[0]
        return _reorder(S, np.arange(n).reshape(2, -1).T.flatten())

    sy.rotation, sy.sympmat, sy.xpxp_to_xxpp, sy.xxpp_to_xpxp = rotation, sympmat, xpxp_to_xxpp, xxpp_to_xpxp

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import strawberryfields as sf  # noqa: E402

    sf._b200_oracle_shim = True
    return sf
