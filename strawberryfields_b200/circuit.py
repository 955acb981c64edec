"""Device-resident Fock-basis circuit: the host-side driver of ``libb200fock.so``.

Mirrors the interface of the reference's ``Circuit``
(``/root/reference/strawberryfields/backends/fockbackend/circuit.py:35-812``): same method
names, argument meaning and error behaviour, but the state lives in HBM and every
update is a CUDA kernel launched through the C ABI (``include/b200fock.h``).  PyTorch
is used only to own device memory and the stream.

State layout (same as the reference, ``backend.py:50-56``): C-order complex128,
``[B?][D]*n`` for pure states and ``[B?][D]*2n`` with axes ``(ket_0, bra_0, ket_1, ...)``
for mixed states; ``B`` is an optional leading batch axis (TF-backend semantics,
``tfbackend/backend.py:66-112``).

Lazy gate queue.  The engine delivers gates one call at a time
(``engine.py:422-457``); to let one HBM pass do the work of several gates the circuit
keeps, per mode, one pending single-mode operator:

* consecutive single-mode gates on a mode are pre-multiplied (D x D products on the
  device);
* diagonal gates (``Rgate``, ``Kgate``) are folded into the neighbouring dense or
  two-mode gate table and cost no pass at all;
* whatever is still pending when the state is observed is flushed -- all leftover
  diagonals in ONE multi-axis pass.
"""
from __future__ import annotations

import copy
import ctypes as C
from math import factorial

import numpy as np
import torch

from . import lib as L
from . import scheduler as S

C128 = np.complex128
_HBAR = 2  # circuit.py:61

# Set (together with a stand-in library handle) ONLY by the CPU test-suite, which drives the host
# logic against a numpy double of the C ABI (tests/fake_lib.py).  The product never sets it.
_TEST_HOST_MODE = False


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class DeviceParams:
    """Gate parameters that already live on the device: a contiguous float64 tensor of shape
    ``(2,)`` or ``(2, nbatch)`` (row 0 = first parameter, row 1 = second).  Pass it as the
    first parameter of a gate call (the second is ignored); the gate table is generated on
    the device straight from it, so no parameter ever returns to the host."""

    def __init__(self, tensor):
        if tensor.dtype != torch.float64 or not tensor.is_contiguous() or tensor.shape[0] != 2:
            raise ValueError("DeviceParams needs a contiguous float64 tensor of shape (2,) or (2, nbatch)")
        self.tensor = tensor
        self.nbatch = 1 if tensor.dim() == 1 else int(tensor.shape[1])


class TableCache:
    """Parameter-keyed cache of device gate tables -- what the reference's ten ``functools.lru_cache``
    decorators do for its host matrices (``fockbackend/ops.py:208-343``), extended to the tables the
    fold stage derives from them (products of same-mode gates, diagonals folded into a dense or a
    two-mode table).  Every table carries a *recipe*: a hashable description of how it was built
    (generator kind + host parameters, or an operation on other recipes).  A repeated circuit therefore
    launches no generator, compose or fold kernel at all: each gate pass finds its table by recipe.
    Tables built from device-resident parameters (``DeviceParams``) have no recipe and are never cached.
    Cached tables are shared and must never be written in place."""

    def __init__(self, max_bytes=256 << 20):
        from collections import OrderedDict

        self._d = OrderedDict()
        self._bytes = 0
        self.max_bytes = max_bytes
        self.hits = 0
        self.misses = 0
        self.enabled = True

    def get(self, recipe, build):
        """the table of ``recipe`` (built with ``build()`` on a miss); ``recipe is None``: always build"""
        if recipe is None or not self.enabled:
            t = build()
            t._recipe = recipe
            return t
        t = self._d.get(recipe)
        if t is not None:
            self._d.move_to_end(recipe)
            self.hits += 1
            return t
        self.misses += 1
        t = build()
        t._recipe = recipe
        self._d[recipe] = t
        self._bytes += t.numel() * t.element_size()
        while self._bytes > self.max_bytes and len(self._d) > 1:
            _, old = self._d.popitem(last=False)
            self._bytes -= old.numel() * old.element_size()
        return t

    def clear(self):
        self._d.clear()
        self._bytes = 0


TABLES = TableCache()


def _recipe(t):
    return getattr(t, "_recipe", None) if t is not None else 0


def _derived(op, *parts):
    """recipe of a table derived from other tables (None as soon as one input has none)"""
    rs = tuple(p if isinstance(p, (int, str)) else _recipe(p) for p in parts)
    return None if any(r is None for r in rs) else (op,) + rs


class DeviceCircuit:
    """GPU mirror of ``fockbackend.circuit.Circuit``."""

    def __init__(self, num, trunc, pure=True, batch_size=None, device=None, strict_purity=False,
                 fuse=True, lazy_vacuum=False):
        if num < 0:
            raise ValueError("Number of modes must be non-negative -- got {}".format(num))
        if trunc <= 0:
            raise ValueError("Truncation must be positive -- got {}".format(trunc))
        if trunc > L.MAX_CUTOFF:
            raise ValueError("b200fock supports cutoff_dim <= {}".format(L.MAX_CUTOFF))
        L.load()  # fail loudly when the CUDA library is missing
        if _TEST_HOST_MODE:
            self.device = torch.device("cpu")
        else:
            if not torch.cuda.is_available():
                raise L.B200Error("b200fock needs a CUDA device (there is no CPU fallback)")
            self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._num_modes = num
        self._hbar = _HBAR
        self._batched = batch_size is not None
        self._B = int(batch_size) if batch_size is not None else 1
        if self._B < 1:
            raise ValueError("batch_size must be a positive integer")
        if self._B > L.MAX_BATCH:
            # the kernels map the batch axis to gridDim.z
            raise ValueError("b200fock supports batch_size <= {}".format(L.MAX_BATCH))
        self._strict = bool(strict_purity)
        # True / "fold": fold diagonal and same-mode gates, one streaming pass per remaining gate (default);
        # "tile": additionally group gates into multi-gate tile passes (scheduler.py); False: one pass per gate
        self._fuse = fuse if fuse in ("tile", "fold") else ("fold" if fuse else False)
        # lazy vacuum (pure states, fuse="fold"): a mode that no two-mode gate has touched yet is a
        # product factor (its pending single-mode operator applied to |0>); the device tensor only
        # spans the modes that have been entangled so far and grows by one axis when a gate needs it
        self._lazy_opt = bool(lazy_vacuum)
        self._scratch = None
        self._part = None
        self._norm_part = torch.zeros(4096, dtype=torch.float64, device=self.device)
        self.reset(pure=pure, cutoff_dim=trunc)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        if self.device.type != "cuda":
            return None
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _axes(self):
        return self._num_modes if self._pure else 2 * self._num_modes

    def _size(self):
        """elements per batch entry of the device tensor (all axes, unless lazy-vacuum modes are
        still outside it)"""
        return self._trunc ** len(self._phys)

    def _stride(self, axis):
        """element stride of (virtual) tensor axis ``axis`` in the current physical layout"""
        return self._trunc ** (len(self._phys) - 1 - self._pos[axis])

    def _canon_stride(self, axis):
        return self._trunc ** (self._axes() - 1 - axis)

    def _set_identity_layout(self):
        """physical position p holds virtual axis _phys[p]; _pos is the inverse map.  Tile passes
        permute axes (scheduler.py); everything that reads the state goes through _stride()."""
        self._phys = list(range(self._axes()))
        self._pos = list(range(self._axes()))
        self._inactive = set()

    def _is_canonical(self):
        return len(self._phys) == self._axes() and all(p == v for p, v in enumerate(self._phys))

    def _canonicalize(self):
        """Restore the canonical axis order (one out-of-place strided copy)."""
        if self._is_canonical():
            return
        D, B = self._trunc, self._B
        per = self._size()
        out = self._get_scratch(self._buf.numel())
        oa = [(B, per, 0, per)] if B > 1 else []
        for v in range(self._axes()):
            oa.append((D, self._stride(v), 0, self._canon_stride(v)))
        self._gather(self._buf, None, out, oa)
        if self._shared:
            self._buf, self._scratch, self._shared = out, None, False
        else:
            self._buf, self._scratch = out, self._buf
        self._set_identity_layout()

    def _new(self, n):
        return torch.empty(int(n), dtype=torch.complex128, device=self.device)

    def _get_scratch(self, n):
        if self._scratch is None or self._scratch.numel() < n:
            self._scratch = None
            self._scratch = self._new(n)
        return self._scratch[:n] if self._scratch.numel() != n else self._scratch

    def _get_part(self, n_out):
        need = int(n_out) * 64
        if self._part is None or self._part.numel() < need:
            self._part = self._new(need)
        return self._part

    def _own(self):
        """Copy-on-write: a state object returned by ``snapshot()`` shares the buffer until the
        circuit is modified again."""
        if self._shared:
            self._buf = self._buf.clone()
            self._shared = False

    def snapshot(self):
        """Read-only view of the current state for a state object (no copy now)."""
        self._flush()
        snap = object.__new__(DeviceCircuit)
        snap.__dict__.update(self.__dict__)
        snap._pending = {}
        snap._opq = []
        snap._defer_log = None
        snap._untouched = set()
        snap._inactive = set()
        snap._scratch = None
        snap._part = None
        snap._gram_part = None
        snap._norm_part = torch.zeros(4096, dtype=torch.float64, device=self.device)
        snap._shared = True
        self._shared = True
        return snap

    @classmethod
    def from_buffer(cls, like, buf, num_modes, pure):
        """Wrap an existing device tensor (e.g. a reduced density matrix) as a circuit."""
        obj = object.__new__(cls)
        obj.__dict__.update(like.__dict__)
        obj._buf = buf
        obj._num_modes = num_modes
        obj._pure = pure
        obj._pending = {}
        obj._untouched = set()
        obj._scratch = None
        obj._part = None
        obj._norm_part = torch.zeros(4096, dtype=torch.float64, device=like.device)
        obj._shared = False
        obj._opq = []
        obj._defer_log = None
        obj._pregen = {}
        obj._set_identity_layout()
        return obj

    def permute_modes(self, perm):
        """New mode p <- old mode perm[p] (both tensor axes of a mode move together)."""
        self._flush()
        self._canonicalize()
        D, B, n = self._trunc, self._B, self._num_modes
        per = self._size()
        out = self._new(self._buf.numel())
        oa = [(B, per, 0, per)] if B > 1 else []
        for p in range(n):
            for t, ax in enumerate(self._mode_axes(p)):
                oa.append((D, self._stride(self._mode_axes(int(perm[p]))[t]), 0, self._stride(ax)))
        self._gather(self._buf, None, out, oa)
        self._buf, self._shared, self._scratch = out, False, None

    def _vacuum(self, buf, per):
        L.call("b200_fill_zero", _ptr(buf), buf.numel(), self._stream())
        if self._B == 1:
            L.call("b200_set_element", _ptr(buf), 0, 1.0, 0.0, self._stream())
        else:
            buf[::per] = 1.0  # one strided fill for the whole batch (torch: plumbing, not a hot path)

    # ------------------------------------------------------------------ reset (circuit.py:89-116)
    def reset(self, pure=None, cutoff_dim=None, num_subsystems=None):
        if pure is not None:
            if not isinstance(pure, bool):
                raise ValueError("Argument 'pure' must be either True or False")
            self._pure = pure
        if num_subsystems is not None:
            if not isinstance(num_subsystems, int):
                raise ValueError("Argument 'num_subsystems' must be a positive integer")
            self._num_modes = num_subsystems
        if cutoff_dim is not None:
            if not isinstance(cutoff_dim, int) or cutoff_dim < 1:
                raise ValueError("Argument 'cutoff_dim' must be a positive integer")
            if cutoff_dim > L.MAX_CUTOFF:
                raise ValueError("b200fock supports cutoff_dim <= {}".format(L.MAX_CUTOFF))
            self._trunc = cutoff_dim
        self._scratch = None
        self._buf = None
        self._shared = False
        self._pending = {}
        self._opq = []
        self._untouched = set(range(self._num_modes))
        self._defer_log = None
        self._inner_hint = None
        if self._lazy_opt and self._fuse == "fold" and self._num_modes > 0:
            # nothing is entangled yet: the device tensor has no axes (one amplitude, 1, per batch entry)
            self._phys = []
            self._pos = [None] * self._axes()
            self._inactive = set(range(self._num_modes))
            self._buf = torch.ones(self._B, dtype=torch.complex128, device=self.device)
            # gate calls are only recorded until the state is first observed: the replay then knows the
            # whole program and can choose the tensor layout (_replay)
            self._defer_log = []
            return
        self._set_identity_layout()
        per = self._size()
        self._buf = self._new(self._B * per)
        self._vacuum(self._buf, per)

    # ------------------------------------------------------------------ gate tables
    def _params(self, *ps):
        """scalars or length-B arrays -> (nbatch, scalars, device params or None, cache key or None)"""
        if isinstance(ps[0], DeviceParams):
            if ps[0].nbatch not in (1, self._B):
                raise ValueError("DeviceParams batch does not match the circuit's batch_size")
            return ps[0].nbatch, [0.0, 0.0], ps[0].tensor, None
        arrs = [np.asarray(p, dtype=np.float64) for p in ps]
        if all(a.ndim == 0 for a in arrs):
            sc = [float(a) for a in arrs] + [0.0] * (2 - len(arrs))
            return 1, sc, None, tuple(sc)
        if not self._batched:
            raise ValueError("array-valued gate parameters need a batched circuit (batch_size=...)")
        full = np.zeros((2, self._B), dtype=np.float64)
        for i, a in enumerate(arrs):
            if a.ndim == 0:
                full[i, :] = float(a)
            elif a.shape == (self._B,):
                full[i, :] = a
            else:
                raise ValueError("gate parameter must be a scalar or have shape (batch_size,)")
        return self._B, [0.0, 0.0], full, full.tobytes()

    def _key(self, *parts):
        return parts + (self._trunc, str(self.device))

    def _dev_params(self, dev):
        """per-entry host parameters [2, B] -> device (only on a cache miss)"""
        return torch.from_numpy(dev).to(self.device) if isinstance(dev, np.ndarray) else dev

    def _gen1(self, kind, p0, p1):
        ahead = self._pregenerated(kind, p0)
        if ahead is not None:
            return ahead
        nb, sc, dev, key = self._params(p0, p1)
        D = self._trunc

        def build():
            out = self._new(nb * D * D)
            L.call("b200_gen_gate1", kind, D, nb, sc[0], sc[1], _ptr(self._dev_params(dev)), _ptr(out), self._stream())
            return out.view(nb, D, D)

        return TABLES.get(None if key is None else self._key("g1", kind, key), build)

    def _gen_diag(self, kind, p0):
        ahead = self._pregenerated(kind, p0)
        if ahead is not None:
            return ahead
        nb, sc, dev, key = self._params(p0)
        D = self._trunc
        per = D * D if kind == L.DIAG_CROSS_KERR else D

        def build():
            out = self._new(nb * per)
            L.call("b200_gen_diag", kind, D, nb, sc[0], _ptr(self._dev_params(dev)), _ptr(out), self._stream())
            return out.view(nb, per)

        return TABLES.get(None if key is None else self._key("gd", kind, key), build)

    def _gen2(self, kind, p0, p1=0.0):
        ahead = self._pregenerated(kind, p0)
        if ahead is not None:
            return ahead
        nb, sc, dev, key = self._params(p0, p1)
        D = self._trunc
        if D > L.MAX_PAIR_CUTOFF:
            # the block-packed table of a two-axis operator (D^2 + (D-1)D(2D-1)/3 entries) is staged whole in
            # shared memory by the gate kernels: 227 KB hold it up to cutoff 27
            raise ValueError("b200fock supports two-mode gates and the loss channel up to cutoff_dim = {} "
                             "(single-mode and diagonal gates up to {}); got {}".format(
                                 L.MAX_PAIR_CUTOFF, L.MAX_CUTOFF, D))
        P = L.packed_size(D)

        def build():
            out = self._new(nb * P)
            L.call("b200_gen_gate2", kind, D, nb, sc[0], sc[1], _ptr(self._dev_params(dev)), _ptr(out), self._stream())
            return out.view(nb, P)

        return TABLES.get(None if key is None else self._key("g2", kind, key), build)

    def _upload_matrix(self, mat):
        mat = np.ascontiguousarray(np.asarray(mat, dtype=C128))
        D = self._trunc
        if mat.shape != (D, D):
            raise ValueError("single-mode operator must have shape (cutoff, cutoff)")
        return TABLES.get(self._key("mat", mat.tobytes()),
                          lambda: torch.from_numpy(mat).to(self.device).view(1, D, D))

    @staticmethod
    def _expand(t, nb):
        return t if t.shape[0] == nb else t.expand(nb, *t.shape[1:]).contiguous()

    # ------------------------------------------------------------------ raw kernels
    # A launch may be restricted to a *view*: one index range of the OUTERMOST axis, i.e. a contiguous part
    # of the buffer (offset, elements).  No gate acts on a sharded axis, so the parts of a shard are
    # independent states as far as the gate kernels are concerned -- the sharded exchange uses this to run
    # the first gates on the parts that have arrived while the rest is still in flight (sharding.py).
    def _vptr(self):
        view = self.__dict__.get("_view")
        return _ptr(self._buf) if view is None else C.c_void_p(self._buf.data_ptr() + 16 * view[0])

    def _vsize(self):
        view = self.__dict__.get("_view")
        return self._size() if view is None else view[1]

    def _pass(self, tag, name, *args):
        """Launch one full pass over the state.  With ``self.profile`` set to a list, the
        launch is bracketed by CUDA events on the launching stream and
        ``(tag, algorithmic bytes, start, end)`` is appended (bench.py's roofline leg)."""
        prof = self.__dict__.get("profile")
        if prof is None:
            L.call(name, *args)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.call(name, *args)
        e1.record()
        prof.append((tag, 32 * self._B * self._vsize(), e0, e1))

    def _check_launch(self, table, per_entry, *strides):
        """Host-side bounds of a gate launch: the buffer holds B states of _size() elements, every gate
        axis lies inside one state, the table holds one entry (or B) of ``per_entry`` coefficients."""
        size, D = self._vsize(), self._trunc
        if self._buf.numel() < self._B * self._size():
            raise L.B200Error("state buffer holds %d elements, the launch covers %d" % (self._buf.numel(), self._B * size))
        if any(s < 1 or s * D > size for s in strides):
            raise L.B200Error("gate axis stride %r does not fit a state of %d elements" % (strides, size))
        if table.shape[0] not in (1, self._B) or table.numel() < table.shape[0] * per_entry:
            raise L.B200Error("gate table of shape %r for a batch of %d" % (tuple(table.shape), self._B))

    def _k_gate1(self, U, axis, conj):
        self._own()
        self._check_launch(U, self._trunc ** 2, self._stride(axis))
        D, B = self._trunc, self._B
        naxes = len(self._phys)
        nb = U.shape[0]
        pos = self._pos[axis]
        inner = self._stride(axis)
        self._pass("gate1/axis%d" % (naxes - 1 - pos), "b200_apply_gate1", self._vptr(),
                   self._vsize() // (D * inner), D, inner, _ptr(U),
                   int(conj), B, self._size(), D * D if nb > 1 else 0, self._stream())

    def _k_gate2(self, G, rule, ax1, ax2, conj):
        self._own()
        self._check_launch(G, L.packed_size(self._trunc), self._stride(ax1), self._stride(ax2))
        D, B = self._trunc, self._B
        nb = G.shape[0]
        naxes = len(self._phys)
        # the C side routes pair gates with an axis of stride 1 to the staged kernel (k_apply_inner)
        staged = min(self._stride(ax1), self._stride(ax2)) == 1 and 2 <= D <= L.MAX_FAST_CUTOFF
        self._pass("%s/rule%d/axes%d,%d" % ("inner2" if staged else "gate2", rule, naxes - 1 - self._pos[ax1],
                                            naxes - 1 - self._pos[ax2]),
                   "b200_apply_gate2",
                   self._vptr(), self._vsize(), D, self._stride(ax1), self._stride(ax2),
               rule, _ptr(G), int(conj), B, self._size(), G.shape[1] if nb > 1 else 0, self._stream())

    def _k_diag_pair(self, tab, ax1, ax2, conj):
        self._own()
        self._check_launch(tab, self._trunc ** 2, self._stride(ax1), self._stride(ax2))
        D, B = self._trunc, self._B
        nb = tab.shape[0]
        self._pass("diag2", "b200_apply_diag", self._vptr(), self._vsize(), D, self._stride(ax1), self._stride(ax2),
               _ptr(tab), int(conj), B, self._size(), tab.shape[1] if nb > 1 else 0, self._stream())

    def _k_diag_multi(self, items):
        """items: list of (table [nb, D], mode) -- one pass for all of them."""
        self._own()
        D, B = self._trunc, self._B
        axes = []  # (stride, conj, table)
        for tab, mode in items:
            if self._pure:
                axes.append((self._stride(mode), 0, tab))
            else:
                axes.append((self._stride(2 * mode), 0, tab))
                axes.append((self._stride(2 * mode + 1), 1, tab))
        for s, _, tab in axes:
            self._check_launch(tab, D, s)
        for i in range(0, len(axes), L.MAX_AXES):
            chunk = axes[i:i + L.MAX_AXES]
            nb = max(t.shape[0] for _, _, t in chunk)
            tabs = TABLES.get(_derived("stack", nb, *[t for _, _, t in chunk]),
                              lambda: torch.stack([self._expand(t, nb) for _, _, t in chunk], dim=1).contiguous())
            k = len(chunk)
            strides = (C.c_int64 * k)(*[s for s, _, _ in chunk])
            conjs = (C.c_int * k)(*[c for _, c, _ in chunk])
            self._pass("diag_multi/%d" % k, "b200_apply_diag_multi", self._vptr(), self._vsize(), D, k, strides, conjs, _ptr(tabs),
                   B, self._size(), k * D if nb > 1 else 0, self._stream())

    # ------------------------------------------------------------------ immediate application
    def _apply_dense_now(self, U, mode):
        if self._pure:
            self._k_gate1(U, mode, 0)
        else:
            self._k_gate1(U, 2 * mode, 0)
            self._k_gate1(U, 2 * mode + 1, 1)

    def _apply_pair_now(self, G, rule, m1, m2):
        if self._pure:
            self._k_gate2(G, rule, m1, m2, 0)
        else:
            self._k_gate2(G, rule, 2 * m1, 2 * m2, 0)
            self._k_gate2(G, rule, 2 * m1 + 1, 2 * m2 + 1, 1)

    # ------------------------------------------------------------------ tile queue
    def _tile_mode(self):
        """Tile passes need >= 3 tensor axes, a cutoff the tile kernel is instantiated for, and
        fuse=True (fuse="fold" keeps the one-pass-per-gate path with diagonal folding only)."""
        return (self._fuse == "tile" and 2 <= self._trunc <= L.MAX_FAST_CUTOFF and self._axes() >= 3)

    def _emit_dense(self, U, mode):
        """A dense single-mode operator leaves the fold stage: queue it (tile mode) or apply it."""
        if not self._tile_mode():
            self._apply_dense_now(U, mode)
            return
        sz = self._trunc ** 2
        if self._pure:
            self._opq.append(S.Op(S.KIND_SINGLE, (mode,), U, 0, sz))
        else:
            self._opq.append(S.Op(S.KIND_SINGLE, (2 * mode,), U, 0, sz))
            self._opq.append(S.Op(S.KIND_SINGLE, (2 * mode + 1,), U, 1, sz))
        self._queue_limit()

    def _emit_diags(self, items):
        if not self._tile_mode():
            self._k_diag_multi(items)
            return
        for d, mode in items:
            if self._pure:
                self._opq.append(S.Op(S.KIND_DIAG, (mode,), d, 0, self._trunc, 0.2))
            else:
                self._opq.append(S.Op(S.KIND_DIAG, (2 * mode,), d, 0, self._trunc, 0.2))
                self._opq.append(S.Op(S.KIND_DIAG, (2 * mode + 1,), d, 1, self._trunc, 0.2))
        self._queue_limit()

    def _emit_pair(self, G, rule, m1, m2):
        if not self._tile_mode():
            self._apply_pair_now(G, rule, m1, m2)
            return
        sz = L.packed_size(self._trunc)
        if self._pure:
            self._opq.append(S.Op(rule, (m1, m2), G, 0, sz))
        else:
            self._opq.append(S.Op(rule, (2 * m1, 2 * m2), G, 0, sz))
            self._opq.append(S.Op(rule, (2 * m1 + 1, 2 * m2 + 1), G, 1, sz))
        self._queue_limit()

    def _queue_limit(self):
        if len(self._opq) >= 96:  # bounds table memory and scheduling time on very long programs
            self._run_queue()

    def _run_queue(self):
        """Schedule the queued operators into tile passes (scheduler.py) and launch them."""
        if not self._opq:
            return
        ops, self._opq = self._opq, []
        self._own()
        D = self._trunc
        tile_bytes = L.tile_smem_bytes(D, 0)
        budget = (L.SMEM_LIMIT - tile_bytes) // 16  # one persistent CTA per SM owns all shared memory
        passes, phys = S.plan(ops, self._phys, L.TILE_MAX_OPS, budget)
        for p in passes:
            self._launch_tile_pass(p)
        self._phys = list(phys)
        pos = [0] * len(phys)
        for i, v in enumerate(phys):
            pos[v] = i
        self._pos = pos

    def _launch_tile_pass(self, p):
        D, B, A = self._trunc, self._B, self._axes()
        nb = max([op.table.shape[0] for op, _ in p.ops] + [1])
        tops = (L.TileOp * max(len(p.ops), 1))()
        parts, off = [], 0
        for i, (op, tax) in enumerate(p.ops):
            tops[i].kind = op.kind
            tops[i].axis1 = tax[0]
            tops[i].axis2 = tax[1] if len(tax) > 1 else 0
            tops[i].conj = int(op.conj)
            tops[i].coef_offset = off
            parts.append(self._expand(op.table.reshape(op.table.shape[0], -1), nb))
            off += op.coef_size
        coef = torch.cat(parts, dim=1).contiguous() if parts else None
        perm = (C.c_int * 3)(*p.out_perm)
        stride0 = D ** (A - 1 - p.positions[0])
        stride1 = D ** (A - 1 - p.positions[1])
        self._pass("tile/%dops" % len(p.ops), "b200_apply_tile_pass", _ptr(self._buf), self._size(), D, stride0,
                   stride1, tops, len(p.ops), perm, _ptr(coef), off, B, self._size(), off if nb > 1 else 0,
                   self._stream())

    # ------------------------------------------------------------------ deferred program (lazy vacuum)
    _PAIR_GATES = ("beamsplitter", "mzgate", "two_mode_squeeze", "cross_kerr_interaction")

    def _defer(self, name, *args):
        """Lazy vacuum: while the state of a fresh circuit has not been observed, a gate call is only
        recorded (the engine delivers gates one call at a time, engine.py:427-444; the reference applies
        each immediately).  Returns True when the call was recorded."""
        log = self.__dict__.get("_defer_log")
        if log is None:
            return False
        # host arrays are copied; ``DeviceParams`` tensors are kept by reference and read at the replay (a copy
        # per gate would be a launch per gate): do not overwrite them before the state has been observed
        log.append((name, tuple(a.copy() if isinstance(a, np.ndarray) else a for a in args)))
        return True

    def _replay(self):
        """Run the recorded program.  Knowing all of it, the mode that takes part in the FEWEST two-mode
        gates is made the innermost tensor axis (it is activated first): gates on the innermost axis are the
        slow passes (staged through shared memory), every other axis streams at the HBM rate."""
        log = self.__dict__.get("_defer_log")
        if log is None:
            return
        self._defer_log = None
        count, first = {}, {}
        for i, (name, args) in enumerate(log):
            if name in self._PAIR_GATES:
                for m in args[-2:]:
                    count[m] = count.get(m, 0) + 1
                    first.setdefault(m, i)
        if count:
            self._inner_hint = min(count, key=lambda m: (count[m], first[m]))
        self._pregenerate(log)
        try:
            for name, args in log:
                getattr(self, name)(*args)
        finally:
            self._pregen = {}

    # gate method -> (generator entry point, kind, table entries as a function of the cutoff)
    _GENERATORS = {
        "displacement": ("b200_gen_gate1", L.GATE_DISPLACEMENT, lambda D: D * D),
        "squeeze": ("b200_gen_gate1", L.GATE_SQUEEZE, lambda D: D * D),
        "phase_shift": ("b200_gen_diag", L.DIAG_ROTATION, lambda D: D),
        "kerr_interaction": ("b200_gen_diag", L.DIAG_KERR, lambda D: D),
        "cross_kerr_interaction": ("b200_gen_diag", L.DIAG_CROSS_KERR, lambda D: D * D),
        "beamsplitter": ("b200_gen_gate2", L.GATE_BEAMSPLITTER, L.packed_size),
        "mzgate": ("b200_gen_gate2", L.GATE_MZ, L.packed_size),
        "two_mode_squeeze": ("b200_gen_gate2", L.GATE_S2, L.packed_size),
    }

    def _pregenerate(self, log):
        """Gate tables of device-resident parameters (``DeviceParams``: no recipe, so the table cache cannot
        serve them) are generated in ONE launch per gate kind for the whole recorded program -- the generators
        take a batch of parameter pairs -- instead of one launch per gate while the tensors are still small and
        the host is the bottleneck."""
        self._pregen = {}
        D = self._trunc
        groups = {}
        for name, args in log:
            if name in self._GENERATORS and args and isinstance(args[0], DeviceParams) and args[0].nbatch == 1:
                groups.setdefault(name, []).append(args[0])
        for name, dps in groups.items():
            if len(dps) < 2 or (D > L.MAX_PAIR_CUTOFF and self._GENERATORS[name][0] == "b200_gen_gate2"):
                continue
            entry, kind, size = self._GENERATORS[name]
            per, K = size(D), len(dps)
            params = torch.stack([dp.tensor for dp in dps], dim=1).contiguous()   # [2, K]
            out = self._new(K * per)
            if entry == "b200_gen_diag":
                L.call(entry, kind, D, K, 0.0, _ptr(params), _ptr(out), self._stream())
            else:
                L.call(entry, kind, D, K, 0.0, 0.0, _ptr(params), _ptr(out), self._stream())
            shape = (1, D, D) if entry == "b200_gen_gate1" else (1, per)
            for k, dp in enumerate(dps):
                self._pregen.setdefault((id(dp), kind), []).append(out[k * per:(k + 1) * per].view(shape))

    def _pregenerated(self, kind, p0):
        """the table generated ahead for this DeviceParams object, if any (each is handed out once)"""
        if isinstance(p0, DeviceParams):
            tabs = self.__dict__.get("_pregen", {}).get((id(p0), kind))
            if tabs:
                t = tabs.pop(0)
                t._recipe = None
                return t
        return None

    # ------------------------------------------------------------------ lazy queue
    def _touch(self, *modes):
        for m in modes:
            self._untouched.discard(m)

    def _queue_dense(self, U, mode):
        self._touch(mode)
        if not self._fuse:
            self._apply_dense_now(U, mode)
            return
        D = self._trunc
        pend = self._pending.get(mode)
        if pend is None:
            self._pending[mode] = ("dense", U)
        elif pend[0] == "diag":
            d = pend[1]
            nb = max(U.shape[0], d.shape[0])

            def build():
                out = self._expand(U, nb).clone()  # cached tables are shared: never folded in place
                L.call("b200_fold_diag_gate1", D, nb, _ptr(out), _ptr(self._expand(d, nb)), None, self._stream())
                return out

            self._pending[mode] = ("dense", TABLES.get(_derived("fold1", U, d, 0), build))
        else:
            P = pend[1]
            nb = max(U.shape[0], P.shape[0])

            def build():
                out = self._new(nb * D * D).view(nb, D, D)
                L.call("b200_compose_gate1", D, nb, _ptr(self._expand(U, nb)), _ptr(self._expand(P, nb)),
                       _ptr(out), self._stream())
                return out

            self._pending[mode] = ("dense", TABLES.get(_derived("compose", U, P), build))

    def _queue_diag(self, d, mode):
        self._touch(mode)
        if not self._fuse:
            self._k_diag_multi([(d, mode)])
            return
        D = self._trunc
        pend = self._pending.get(mode)
        if pend is None:
            self._pending[mode] = ("diag", d)
        elif pend[0] == "diag":
            nb = max(d.shape[0], pend[1].shape[0])

            def build():
                out = self._new(nb * D).view(nb, D)
                L.call("b200_mul_tables", nb * D, _ptr(self._expand(d, nb)), _ptr(self._expand(pend[1], nb)), 0,
                       _ptr(out), self._stream())
                return out

            self._pending[mode] = ("diag", TABLES.get(_derived("mul", d, pend[1]), build))
        else:
            P = pend[1]
            nb = max(d.shape[0], P.shape[0])

            def build():
                out = self._expand(P, nb).clone()
                L.call("b200_fold_diag_gate1", D, nb, _ptr(out), None, _ptr(self._expand(d, nb)), self._stream())
                return out

            self._pending[mode] = ("dense", TABLES.get(_derived("fold1", P, 0, d), build))

    def _activate(self, modes):
        """Lazy vacuum: bring the still-factored modes among ``modes`` into the device tensor.  Such
        a mode is |0> with (at most) one pending single-mode operator P, i.e. the factor P|0> = column
        0 of P; the tensor grows by one (outermost) axis with one outer-product launch that WRITES the
        new tensor once -- instead of a read + write pass over the full-size state per such gate.
        Mixed states: the factor is the rank-one matrix P|0><0|P^dagger on the (ket, bra) axes."""
        D, B = self._trunc, self._B
        hint = self.__dict__.get("_inner_hint")
        if not self._phys and hint is not None and hint in self._inactive:
            modes = [hint] + [m for m in modes if m != hint]  # first activated = innermost axis
        for m in modes:
            if m not in self._inactive:
                continue
            pend = self._pending.pop(m, None)
            if pend is None or pend[0] == "diag":
                v = torch.zeros(1 if pend is None else pend[1].shape[0], D, dtype=torch.complex128,
                                device=self.device)
                v[:, 0] = 1.0 if pend is None else pend[1][:, 0]
            else:
                v = pend[1][:, :, 0].contiguous()
            nb = v.shape[0]
            old_per = self._size()
            if self._pure:
                f, new_per = v, old_per * D
            else:
                # conj_physical: the kernel reads raw memory, torch's lazy conjugate bit would be lost
                f = (v.reshape(nb, D, 1) * torch.conj_physical(v).reshape(nb, 1, D)).contiguous()
                new_per = old_per * D * D
            out = self._new(B * new_per)
            # out[b][j][i] = f[b][j] * tensor[b][i]: one coalesced read, nf coalesced writes (b200_outer_axis)
            nf = f[0].numel()
            if self._buf.numel() < B * old_per or f.numel() < nb * nf:
                raise L.B200Error("outer_axis operands are smaller than the launch")
            L.call("b200_outer_axis", _ptr(self._buf), _ptr(f.contiguous()), _ptr(out), old_per, nf, B, old_per,
                   nf if nb > 1 else 0, self._stream())
            self._buf, self._shared, self._scratch = out, False, None
            self._inactive.discard(m)
            self._phys = list(self._mode_axes(m)) + self._phys
            for p, ax in enumerate(self._phys):
                self._pos[ax] = p

    def _flush(self, modes=None):
        """Move the pending single-mode operators of ``modes`` (default: all) out of the fold
        stage.  A full flush (``modes is None``) also runs every queued tile pass, after which
        the device buffer holds the up-to-date state."""
        self._replay()
        if self._inactive:
            # descending: every activation adds the outermost axis, so a fresh vacuum ends up canonical
            self._activate(sorted(self._inactive if modes is None else modes, reverse=True))
        if self._pending:
            keys = sorted(self._pending) if modes is None else [m for m in modes if m in self._pending]
            diags = []
            for m in keys:
                kind, tab = self._pending.pop(m)
                if kind == "dense":
                    self._emit_dense(tab, m)
                else:
                    diags.append((tab, m))
            if diags:
                self._emit_diags(diags)
        if modes is None:
            self._run_queue()

    def _pair_gate(self, G, rule, m1, m2):
        """Two-mode gate: pending diagonals on its modes are folded into the table,
        pending dense operators are flushed first."""
        self._touch(m1, m2)
        if self._inactive:
            self._activate(sorted((m1, m2), reverse=True))
        D = self._trunc
        pre = [None, None]
        for i, m in enumerate((m1, m2)):
            pend = self._pending.get(m)
            if pend is None:
                continue
            if pend[0] == "diag":
                pre[i] = self._pending.pop(m)[1]
            else:
                self._flush([m])
        if pre[0] is not None or pre[1] is not None:
            nb = max([G.shape[0]] + [p.shape[0] for p in pre if p is not None])
            G0 = G

            def build():
                out = self._expand(G0, nb).clone()
                p1 = self._expand(pre[0], nb) if pre[0] is not None else None
                p2 = self._expand(pre[1], nb) if pre[1] is not None else None
                L.call("b200_fold_diag_gate2", rule, D, nb, _ptr(out), _ptr(p1), _ptr(p2), None, None, self._stream())
                return out

            G = TABLES.get(_derived("fold2", rule, G0, pre[0], pre[1]), build)
        self._emit_pair(G, rule, m1, m2)

    # ------------------------------------------------------------------ named gates (circuit.py:537-598)
    def phase_shift(self, theta, mode):
        if self._defer('phase_shift', theta, mode):
            return
        self._queue_diag(self._gen_diag(L.DIAG_ROTATION, theta), mode)

    def kerr_interaction(self, kappa, mode):
        if self._defer('kerr_interaction', kappa, mode):
            return
        self._queue_diag(self._gen_diag(L.DIAG_KERR, kappa), mode)

    def displacement(self, r, phi, mode):
        if self._defer('displacement', r, phi, mode):
            return
        self._queue_dense(self._gen1(L.GATE_DISPLACEMENT, r, phi), mode)

    def squeeze(self, r, theta, mode):
        if self._defer('squeeze', r, theta, mode):
            return
        self._queue_dense(self._gen1(L.GATE_SQUEEZE, r, theta), mode)

    def cubic_phase_shift(self, gamma, mode):
        """expm of the truncated x^3 (fockbackend/ops.py:296-306) is a host-built D x D
        matrix; it is applied with the generic dense kernel."""
        from scipy.linalg import expm

        if self._defer("cubic_phase_shift", gamma, mode):
            return
        D = self._trunc
        a = np.diag(np.sqrt(np.arange(1, D)), 1).astype(C128)
        x = (a + a.conj().T) * np.sqrt(self._hbar / 2)
        self._queue_dense(self._upload_matrix(expm(1j * float(gamma) / (3 * self._hbar) * (x @ x @ x))), mode)

    def apply_matrix(self, mat, mode):
        """Arbitrary single-mode operator (host matrix [out, in])."""
        if self._defer("apply_matrix", np.array(mat, dtype=C128), mode):
            return
        self._queue_dense(self._upload_matrix(mat), mode)

    def beamsplitter(self, theta, phi, mode1, mode2):
        if self._defer('beamsplitter', theta, phi, mode1, mode2):
            return
        self._pair_gate(self._gen2(L.GATE_BEAMSPLITTER, theta, phi), L.RULE_SUM, mode1, mode2)

    def mzgate(self, phi_in, phi_ex, mode1, mode2):
        if self._defer('mzgate', phi_in, phi_ex, mode1, mode2):
            return
        self._pair_gate(self._gen2(L.GATE_MZ, phi_in, phi_ex), L.RULE_SUM, mode1, mode2)

    def two_mode_squeeze(self, r, theta, mode1, mode2):
        if self._defer('two_mode_squeeze', r, theta, mode1, mode2):
            return
        self._pair_gate(self._gen2(L.GATE_S2, r, theta), L.RULE_DIFF, mode1, mode2)

    def cross_kerr_interaction(self, kappa, mode1, mode2):
        # diagonal in both modes: commutes with pending diagonals, not with pending dense gates
        if self._defer("cross_kerr_interaction", kappa, mode1, mode2):
            return
        self._touch(mode1, mode2)
        if self._inactive:
            self._activate(sorted((mode1, mode2), reverse=True))
        self._flush([m for m in (mode1, mode2) if self._pending.get(m, ("", 0))[0] == "dense"])
        self._run_queue()  # a two-axis diagonal is not a tile operator: apply it to the current state
        tab = self._gen_diag(L.DIAG_CROSS_KERR, kappa)
        if self._pure:
            self._k_diag_pair(tab, mode1, mode2, 0)
        else:
            self._k_diag_pair(tab, 2 * mode1, 2 * mode2, 0)
            self._k_diag_pair(tab, 2 * mode1 + 1, 2 * mode2 + 1, 1)

    # ------------------------------------------------------------------ channels (circuit.py:65-87, 617-621)
    def loss(self, T, mode):
        if self._defer("loss", T, mode):
            return
        self._flush([mode])  # operators pending on other modes commute with this channel
        self._touch(mode)
        self._to_mixed()
        G = self._gen2(L.CHANNEL_LOSS, T)
        if self._tile_mode():
            self._opq.append(S.Op(S.KIND_DIFF, (2 * mode, 2 * mode + 1), G, 0, L.packed_size(self._trunc)))
            self._queue_limit()
        else:
            self._k_gate2(G, L.RULE_DIFF, 2 * mode, 2 * mode + 1, 0)

    # ------------------------------------------------------------------ strided gather wrapper
    def _gather(self, A, B, Cout, out_axes, red_axes=(), flags=0, base=(0, 0, 0)):
        """out_axes: [(ext, sa, sb, sc)], red_axes: [(ext, ta, tb)] (C order, last fastest)."""
        def merge(axes, width):
            axes = [a for a in axes if a[0] != 1]
            out = []
            for a in axes:
                if out:
                    p = out[-1]
                    if all(p[i] == a[i] * a[0] for i in range(1, width)) and p[0] * a[0] < 2 ** 31:
                        out[-1] = (p[0] * a[0],) + tuple(a[1:])
                        continue
                out.append(tuple(a))
            return out

        oa, ra = merge(list(out_axes), 4), merge(list(red_axes), 3)
        if len(oa) > L.MAX_AXES or len(ra) > L.MAX_AXES:
            raise L.B200Error("tensor rank exceeds the gather kernel's limit")
        # every index the launch can form must lie inside its tensor: a wrong stride fails here, loudly,
        # instead of as an illegal address on the device (and is caught by the CPU suite too)
        for name, t, b0, col, rcol in (("A", A, base[0], 1, 1), ("B", B, base[1], 2, 2), ("C", Cout, base[2], 3, None)):
            if t is None:
                continue
            pairs = [(a[0], a[col]) for a in oa] + ([(r[0], r[rcol]) for r in ra] if rcol is not None else [])
            lo = b0 + sum((e - 1) * s for e, s in pairs if s < 0)
            hi = b0 + sum((e - 1) * s for e, s in pairs if s > 0)
            if lo < 0 or hi >= t.numel():
                raise L.B200Error("gather descriptor addresses [%d, %d] of operand %s with %d elements"
                                  % (lo, hi, name, t.numel()))
        d = L.GatherDesc()
        d.n_out_axes, d.n_red_axes = len(oa), len(ra)
        for j, (e, sa, sb, sc) in enumerate(oa):
            d.out_ext[j], d.out_sa[j], d.out_sb[j], d.out_sc[j] = e, sa, sb, sc
        for j, (e, ta, tb) in enumerate(ra):
            d.red_ext[j], d.red_ta[j], d.red_tb[j] = e, ta, tb
        d.base_a, d.base_b, d.base_c = base
        n_out = 1
        for a in oa:
            n_out *= a[0]
        part = self._get_part(n_out) if (ra and n_out < 2 ** 16) else None
        L.call("b200_gather_reduce", C.byref(d), _ptr(A), _ptr(B), _ptr(Cout), flags, _ptr(part), self._stream())

    def _mode_axes(self, mode):
        """state axes of a mode: (ket,) or (ket, bra)"""
        return (mode,) if self._pure else (2 * mode, 2 * mode + 1)

    # ------------------------------------------------------------------ pure -> mixed (ops.py:110-120)
    def _to_mixed(self):
        if not self._pure:
            return
        self._flush()
        n, D, B = self._num_modes, self._trunc, self._B
        per_p, per_m = D ** n, D ** (2 * n)
        new = self._new(B * per_m)
        axes = [(B, per_p, per_p, per_m)] if B > 1 else []
        for i in range(n):
            st = self._stride(i)
            axes.append((D, st, 0, D ** (2 * n - 1 - 2 * i)))
            axes.append((D, 0, st, D ** (2 * n - 2 - 2 * i)))
        self._gather(self._buf, self._buf, new, axes, flags=L.FLAG_CONJ_B)
        self._buf = new
        self._shared = False
        self._scratch = None
        self._pure = False
        self._set_identity_layout()

    # ------------------------------------------------------------------ norm (circuit.py:367-371)
    def _norm_device(self, materialize=True):
        """squared norm (pure) or trace (mixed) per batch entry, as a device float64 tensor [B].
        ``materialize=False``: of the device tensor as it is -- right after a measurement the modes outside
        it are plain |0> factors of norm one (lazy vacuum)."""
        if materialize:
            self._flush()
        B, per = self._B, self._size()
        out = torch.zeros(B, dtype=torch.float64, device=self.device)
        if self._pure and B == 1:
            L.call("b200_norm2", _ptr(self._buf), per, _ptr(out), _ptr(self._norm_part), self._stream())
        elif self._pure:
            # all B squared norms in one reduction: out[b] = sum_r psi[b, r] * conj(psi[b, r])
            self._gather(self._buf, self._buf, out, [(B, per, per, 1)], [(per, 1, 1)],
                         flags=L.FLAG_CONJ_B | L.FLAG_REAL_OUT)
        else:
            n, D = self._num_modes, self._trunc
            red = [(D, self._stride(2 * i) + self._stride(2 * i + 1), 0) for i in range(n)
                   if i not in self._inactive]
            oa = [(B, per, 0, 1)] if B > 1 else []
            self._gather(self._buf, None, out, oa, red, flags=L.FLAG_REAL_OUT)
        return out

    def norm(self):
        v = self._norm_device().cpu().numpy()
        v = np.sqrt(v) if self._pure else v
        return v if self._batched else v[0]

    # ------------------------------------------------------------------ modes (circuit.py:373-391)
    def alloc(self, n=1):
        self._flush()
        self._canonicalize()
        D, B = self._trunc, self._B
        k = self._axes()
        add = n if self._pure else 2 * n
        old_per, new_per = D ** k, D ** (k + add)
        new = self._new(B * new_per)
        L.call("b200_fill_zero", _ptr(new), new.numel(), self._stream())
        axes = [(B, old_per, 0, new_per)] if B > 1 else []
        axes.append((old_per, 1, 0, D ** add))
        self._gather(self._buf, None, new, axes)
        self._buf = new
        self._shared = False
        self._scratch = None
        for m in range(self._num_modes, self._num_modes + n):
            self._untouched.add(m)
        self._num_modes += n
        self._set_identity_layout()

    def dealloc(self, modes):
        self._to_mixed()
        self._flush()
        keep = [m for m in range(self._num_modes) if m not in modes]
        self._buf = self._partial_trace_keep(keep)
        self._shared = False
        self._scratch = None
        self._num_modes = len(keep)
        self._set_identity_layout()
        self._untouched = set()
        self._pending = {}

    def _partial_trace_keep(self, keep):
        """mixed state -> mixed state on the (sorted) modes ``keep`` (ops.py:144-157)."""
        n, D, B = self._num_modes, self._trunc, self._B
        k = len(keep)
        per, new_per = D ** (2 * n), D ** (2 * k)
        out = self._new(B * new_per)
        oa = [(B, per, 0, new_per)] if B > 1 else []
        for j, m in enumerate(keep):
            oa.append((D, self._stride(2 * m), 0, D ** (2 * k - 1 - 2 * j)))
            oa.append((D, self._stride(2 * m + 1), 0, D ** (2 * k - 2 - 2 * j)))
        red = [(D, self._stride(2 * m) + self._stride(2 * m + 1), 0) for m in range(n) if m not in keep]
        self._gather(self._buf, None, out, oa, red)
        return out

    # ------------------------------------------------------------------ preparation (circuit.py:393-535)
    def prepare_multimode(self, state, modes, input_state_is_pure=None):
        """``input_state_is_pure`` only matters on a batched circuit, where a ``[B, D^k]`` array of kets and a
        ``[D^k, D^k]`` density matrix can have the same shape (``tfbackend/circuit.py:343-396``)."""
        if isinstance(modes, int):
            modes = [modes]
        modes = list(modes)
        if self._batched:
            if self._prepare_batched(state, modes, input_state_is_pure):
                return
            self._prepare_batched_general(state, modes, input_state_is_pure)
            return
        self._replay()
        D, n, k = self._trunc, self._num_modes, len(modes)
        if (k == 1 and modes[0] in self._inactive and (not self._strict or not self._pure) and n > 1
                and np.shape(state) == (D,)):
            # lazy vacuum: the mode is still a product factor, which simply becomes the new ket (stored
            # as the pending operator whose column 0 it is); the state stays pure
            tab = np.zeros((D, D), dtype=C128)
            tab[:, 0] = np.asarray(state, dtype=C128)
            self._pending[modes[0]] = ("dense", self._upload_matrix(tab))
            self._touch(modes[0])
            return
        self._flush()
        self._canonicalize()
        pure_shape, mixed_shape = (D,) * k, (D,) * (2 * k)
        state = np.asarray(state)
        if state.shape == (D ** k,):
            state = state.reshape(pure_shape)
        elif state.shape == (D ** k, D ** k):
            state = state.reshape(mixed_shape)
        if state.shape != pure_shape and state.shape != mixed_shape:
            raise ValueError("Incorrect shape for state preparation")
        if len(modes) != len(set(modes)):
            raise ValueError("The specified modes cannot appear multiple times.")
        is_ket = state.shape == pure_shape
        host = np.ascontiguousarray(state.astype(C128))

        if n == k:
            # circuit.py:441-444 fast path (+ the mode permutation of 461-473)
            self._pure = bool(is_ket)
            self._set_identity_layout()
            src = torch.from_numpy(host.reshape(-1)).to(self.device)
            if modes == list(range(n)):
                self._buf = src
            else:
                self._buf = self._new(src.numel())
                naxes = self._axes()
                oa = []
                for f in range(n):
                    j = modes.index(f)
                    for t, ax in enumerate(self._mode_axes(f)):
                        src_axis = j if self._pure else 2 * j + t
                        oa.append((D, D ** (naxes - 1 - src_axis), 0, self._stride(ax)))
                self._gather(src, None, self._buf, oa)
            self._shared = False
            self._scratch = None
            self._untouched = set()
            return

        if (self._pure and is_ket and not self._strict
                and all(m in self._untouched for m in modes)):
            # Modes still in the untouched product vacuum: psi = psi_rest (x) |0..0>, so replacing them
            # keeps the state pure (the reference would switch to a D^2n tensor here).
            src = torch.from_numpy(host.reshape(-1)).to(self.device)
            out = self._new(self._buf.numel())
            oa = []
            for f in range(n):
                if f in modes:
                    j = modes.index(f)
                    oa.append((D, 0, D ** (k - 1 - j), self._stride(f)))
                else:
                    oa.append((D, self._stride(f), 0, self._stride(f)))
            self._gather(self._buf, src, out, oa)
            self._buf = out
            self._shared = False
            self._scratch = None
            self._touch(*modes)
            return

        # general case: rho <- Tr_modes(rho) (x) new, written straight into its final axis order
        self._to_mixed()
        if is_ket:
            host = np.multiply.outer(host, host.conj()).transpose(
                [x for i in range(k) for x in (i, k + i)]).copy() if k else host
        src = torch.from_numpy(np.ascontiguousarray(host).reshape(-1)).to(self.device)
        keep = [m for m in range(n) if m not in modes]
        red = self._partial_trace_keep(keep)
        kk = len(keep)
        out = self._new(D ** (2 * n))
        oa = []
        for f in range(n):
            for t in (0, 1):
                sc = D ** (2 * n - 1 - (2 * f + t))
                if f in modes:
                    j = modes.index(f)
                    oa.append((D, 0, D ** (2 * k - 1 - (2 * j + t)), sc))
                else:
                    j = keep.index(f)
                    oa.append((D, D ** (2 * kk - 1 - (2 * j + t)), 0, sc))
        self._gather(red, src, out, oa)
        self._buf = out
        self._shared = False
        self._scratch = None
        self._touch(*modes)

    def _prepare_batched(self, state, modes, input_state_is_pure=None):
        """Batched circuits (TF-backend semantics, ``tfbackend/circuit.py:343-396``: one state for every batch
        entry, or an array with a leading batch axis).  Fast path: single-mode KETS on modes that are still the
        untouched vacuum -- the input encodings of a batched program.  |v> = (|v><0|) |0>, so the preparation is
        queued as a rank-one single-mode operator (per entry when the kets differ) and costs what a gate costs.
        Returns False when the call is not of that kind (``_prepare_batched_general`` takes it)."""
        D, B = self._trunc, self._B
        st = np.asarray(state, dtype=C128)
        if len(modes) != 1 or st.shape not in ((D,), (B, D)) or (st.ndim == 2 and B == D and input_state_is_pure is False):
            return False    # (the last case: a D x D density matrix on a circuit whose batch size equals the cutoff)
        m = modes[0]
        log = self.__dict__.get("_defer_log")
        if log and any(name != "_queue_dense" or args[1] == m for name, args in log):
            self._replay()  # gates were recorded before this preparation: apply them first
        if m not in self._untouched or m in self._pending:
            return False
        kets = st.reshape(-1, D)
        tab = np.zeros((kets.shape[0], D, D), dtype=C128)
        tab[:, :, 0] = kets
        table = TABLES.get(self._key("ketprep", tab.tobytes()),
                           lambda: torch.from_numpy(tab).to(self.device))
        if self._defer("_queue_dense", table, m):
            return True
        self._queue_dense(table, m)
        return True

    def _prepare_batched_general(self, state, modes, input_state_is_pure=None):
        """Everything else on a batched circuit (``tfbackend/circuit.py:343-396`` + ``_replace_and_update``):
        kets or density matrices on any modes, one for all entries or one per entry; the three cases of the
        unbatched method with the batch as one more (outermost) gather axis."""
        self._replay()
        self._flush()
        self._canonicalize()
        D, n, k, B = self._trunc, self._num_modes, len(modes), self._B
        if len(modes) != len(set(modes)):
            raise ValueError("The specified modes cannot appear multiple times.")
        st = np.asarray(state)
        shapes = {"ket": [(D,) * k, (D ** k,)], "dm": [(D,) * (2 * k), (D ** k, D ** k)]}
        kind = per_entry = None
        # (kind, one per batch entry?) in the order the flag resolves equal shapes (B = D^k)
        if input_state_is_pure:
            cands = [("ket", True), ("ket", False), ("dm", False), ("dm", True)]
        elif input_state_is_pure is None:
            cands = [("ket", False), ("dm", False), ("ket", True), ("dm", True)]
        else:
            cands = [("dm", False), ("dm", True), ("ket", False), ("ket", True)]
        for kd, pe in cands:
            if any(st.shape == ((B,) + sh if pe else sh) for sh in shapes[kd]):
                kind, per_entry = kd, pe
                break
        if kind is None:
            raise ValueError("Incorrect shape for state preparation")
        is_ket = kind == "ket"
        inner = (D,) * (k if is_ket else 2 * k)
        host = np.ascontiguousarray(st.astype(C128)).reshape(((B,) if per_entry else ()) + inner)
        sb = (lambda per: per) if per_entry else (lambda per: 0)    # stride of the batch axis in the new state
        per = self._size()

        def finish(out, touched=True):
            self._buf, self._shared, self._scratch = out, False, None
            if touched:
                self._touch(*modes)

        if n == k:      # the whole register is replaced (circuit.py:441-444)
            self._pure = bool(is_ket)
            self._set_identity_layout()
            naxes = self._axes()
            new_per = D ** naxes
            src = torch.from_numpy(np.ascontiguousarray(host).reshape(-1)).to(self.device)
            out = self._new(B * new_per)
            oa = [(B, sb(new_per), 0, new_per)]
            for f in range(n):
                j = modes.index(f)
                for t, ax in enumerate(self._mode_axes(f)):
                    src_axis = j if self._pure else 2 * j + t
                    oa.append((D, D ** (naxes - 1 - src_axis), 0, self._stride(ax)))
            self._gather(src, None, out, oa)
            finish(out, touched=False)
            self._untouched = set()
            return

        if self._pure and is_ket and not self._strict and all(m in self._untouched for m in modes):
            # still the product vacuum on these modes: the state stays pure
            src = torch.from_numpy(host.reshape(-1)).to(self.device)
            out = self._new(self._buf.numel())
            oa = [(B, per, sb(D ** k), per)]
            for f in range(n):
                if f in modes:
                    oa.append((D, 0, D ** (k - 1 - modes.index(f)), self._stride(f)))
                else:
                    oa.append((D, self._stride(f), 0, self._stride(f)))
            self._gather(self._buf, src, out, oa)
            finish(out)
            return

        # general case: rho_b <- Tr_modes(rho_b) (x) new_b
        self._to_mixed()
        if is_ket:
            mix = [x for i in range(k) for x in (i, k + i)]
            if per_entry:
                host = np.stack([np.multiply.outer(h, h.conj()).transpose(mix) for h in host])
            else:
                host = np.multiply.outer(host, host.conj()).transpose(mix)
        src = torch.from_numpy(np.ascontiguousarray(host).reshape(-1)).to(self.device)
        keep = [m for m in range(n) if m not in modes]
        red = self._partial_trace_keep(keep)
        kk = len(keep)
        out = self._new(B * D ** (2 * n))
        oa = [(B, D ** (2 * kk), sb(D ** (2 * k)), D ** (2 * n))]
        for f in range(n):
            for t in (0, 1):
                sc = D ** (2 * n - 1 - (2 * f + t))
                if f in modes:
                    j = modes.index(f)
                    oa.append((D, 0, D ** (2 * k - 1 - (2 * j + t)), sc))
                else:
                    j = keep.index(f)
                    oa.append((D, D ** (2 * kk - 1 - (2 * j + t)), 0, sc))
        self._gather(red, src, out, oa)
        finish(out)

    def prepare(self, state, mode):
        self.prepare_multimode(state, [mode] if isinstance(mode, int) else mode)

    def _prepare_ket(self, ket, mode):
        if self._batched:
            self.prepare_multimode(ket, [mode], input_state_is_pure=True)
        elif self._pure or (mode in self._inactive and self._num_modes > 1):  # lazy vacuum: the factor is the ket
            self.prepare(ket, mode)
        else:
            self.prepare(np.outer(ket, ket.conj()), mode)

    # single-mode kets: fockbackend/ops.py:383-461 (tiny host vectors)
    def _kets(self, fn, *params):
        """single-mode ket(s) from scalar parameters, or one per batch entry from length-B arrays"""
        arrs = [np.asarray(p, dtype=np.float64) for p in params]
        if all(a.ndim == 0 for a in arrs):
            return fn(*[float(a) for a in arrs], self._trunc)
        if not self._batched:
            raise ValueError("array-valued preparation parameters need a batched circuit (batch_size=...)")
        full = [np.broadcast_to(a, (self._B,)) for a in arrs]
        return np.stack([fn(*[float(a[b]) for a in full], self._trunc) for b in range(self._B)])

    def prepare_mode_fock(self, n, mode):
        def fock(k, D):
            v = np.zeros(D, dtype=C128)
            v[int(k)] = 1.0
            return v

        self._prepare_ket(self._kets(fock, n), mode)

    def prepare_mode_coherent(self, r, phi, mode):
        self._prepare_ket(self._kets(_coherent, r, phi), mode)

    def prepare_mode_squeezed(self, r, theta, mode):
        self._prepare_ket(self._kets(_squeezed, r, theta), mode)

    def prepare_mode_displaced_squeezed(self, r_d, phi_d, r_s, phi_s, mode):
        self._prepare_ket(self._kets(_displaced_squeezed, r_d, phi_d, r_s, phi_s), mode)

    def prepare_mode_thermal(self, nbar, mode):
        D = self._trunc
        if nbar == 0:
            st = np.zeros((D, D), dtype=C128)
            st[0, 0] = 1.0
        else:
            st = np.diag([nbar ** n / (nbar + 1) ** (n + 1) for n in range(D)]).astype(C128)
        self.prepare(st, mode)

    # ------------------------------------------------------------------ queries
    def is_vacuum(self, tol):
        """circuit.py:600-609"""
        self._flush()
        per = self._size()
        v = self._buf[::per].cpu().numpy() if self._B > 1 else self._buf[:1].cpu().numpy()
        fid = np.abs(v) ** 2 if self._pure else v
        res = np.abs(fid - 1) <= tol
        return res if self._batched else bool(res[0])

    def element(self, n):
        """<n|psi> (pure) or <n|rho|n> (mixed) for the Fock indices ``n``, one value per batch
        entry (states.py:644-655): a D2H read of B numbers, whatever the physical axis order."""
        self._flush()
        per = self._size()
        if self._pure:
            idx = sum(int(x) * self._stride(i) for i, x in enumerate(n))
        else:
            idx = sum(int(x) * (self._stride(2 * i) + self._stride(2 * i + 1)) for i, x in enumerate(n))
        return self._buf[idx::per].cpu().numpy()

    def get_state(self):
        """(device tensor snapshot, pure) -- the reference returns its live array
        (circuit.py:611-615); later gates mutate device memory, so this is a copy."""
        self._flush()
        return self._buf.clone(), self._pure

    def host_state(self):
        self._flush()
        self._canonicalize()
        arr = self._buf.cpu().numpy()
        shape = ([self._B] if self._batched else []) + [self._trunc] * self._axes()
        return arr.reshape(shape)

    # ------------------------------------------------------------------ reductions
    def fock_probs_device(self):
        """all_fock_probs (states.py:580-608) as a device float64 tensor [B, D^n]."""
        self._flush()
        n, D, B = self._num_modes, self._trunc, self._B
        out = torch.empty(B * D ** n, dtype=torch.float64, device=self.device)
        if self._pure and self._is_canonical():
            L.call("b200_abs2", _ptr(self._buf), _ptr(out), self._buf.numel(), self._stream())
        elif self._pure:
            per = self._size()
            oa = [(B, per, per, per)] if B > 1 else []
            for i in range(n):
                oa.append((D, self._stride(i), self._stride(i), D ** (n - 1 - i)))
            self._gather(self._buf, self._buf, out, oa, flags=L.FLAG_CONJ_B | L.FLAG_REAL_OUT)
        else:
            oa = [(B, self._size(), 0, D ** n)] if B > 1 else []
            for i in range(n):
                oa.append((D, self._stride(2 * i) + self._stride(2 * i + 1), 0, D ** (n - 1 - i)))
            self._gather(self._buf, None, out, oa, flags=L.FLAG_REAL_OUT)
        return out.view(B, -1)

    def marginal_probs_device(self, keep):
        """Photon-number distribution of the (sorted) modes ``keep``: float64 [B, D^k].
        Never forms D^(2n) for pure states (the reference does, circuit.py:632-635,675-677)."""
        self._flush()
        n, D, B = self._num_modes, self._trunc, self._B
        k = len(keep)
        if self._pure and k == 1 and D <= L.MAX_FAST_CUTOFF and n > 1:
            return self._gram1(keep[0], diag_only=True)
        out = torch.empty(B * D ** k, dtype=torch.float64, device=self.device)
        per = self._size()
        if self._pure:
            oa = [(B, per, per, D ** k)] if B > 1 else []
            for j, m in enumerate(keep):
                oa.append((D, self._stride(m), self._stride(m), D ** (k - 1 - j)))
            red = [(D, self._stride(m), self._stride(m)) for m in range(n) if m not in keep]
            self._gather(self._buf, self._buf, out, oa, red, flags=L.FLAG_CONJ_B | L.FLAG_REAL_OUT)
        else:
            oa = [(B, per, 0, D ** k)] if B > 1 else []
            for j, m in enumerate(keep):
                oa.append((D, self._stride(2 * m) + self._stride(2 * m + 1), 0, D ** (k - 1 - j)))
            red = [(D, self._stride(2 * m) + self._stride(2 * m + 1), 0) for m in range(n) if m not in keep]
            self._gather(self._buf, None, out, oa, red, flags=L.FLAG_REAL_OUT)
        return out.view(B, -1)

    def reduced_dm_device(self, keep):
        """Reduced density matrix of the (sorted) modes ``keep`` -> complex128 [B, D^2k]
        with interleaved (ket, bra) axes (backend.py:219-253, states.py:613-642)."""
        self._flush()
        if not self._pure:
            return self._partial_trace_keep(keep).view(self._B, -1)
        n, D, B = self._num_modes, self._trunc, self._B
        k = len(keep)
        if k == 1 and D <= L.GRAM_MAX_CUTOFF and n > 1:
            return self._gram1(keep[0], diag_only=False)
        per, new_per = D ** n, D ** (2 * k)
        out = self._new(B * new_per)
        oa = [(B, per, per, new_per)] if B > 1 else []
        for j, m in enumerate(keep):
            oa.append((D, self._stride(m), 0, D ** (2 * k - 1 - 2 * j)))
            oa.append((D, 0, self._stride(m), D ** (2 * k - 2 - 2 * j)))
        red = [(D, self._stride(m), self._stride(m)) for m in range(n) if m not in keep]
        self._gather(self._buf, self._buf, out, oa, red, flags=L.FLAG_CONJ_B)
        return out.view(B, -1)

    def _gram1(self, mode, diag_only):
        """Reduced density matrix (complex128 [B, D*D]) or photon-number marginal (float64 [B, D]) of ONE mode of
        a ket, in one read of the state (``b200_gram1``): what homodyne measurement, ``mean_photon``,
        ``quad_expectation`` and ``wigner`` reduce to."""
        D, B = self._trunc, self._B
        inner = self._stride(mode)
        per = self._size()
        out = (torch.empty(B * D, dtype=torch.float64, device=self.device) if diag_only
               else self._new(B * D * D))
        need = int(L.load().b200_gram1_part_doubles(D, B))
        part = self.__dict__.get("_gram_part")
        if part is None or part.numel() < need:
            part = self._gram_part = torch.empty(need, dtype=torch.float64, device=self.device)
        L.call("b200_gram1", _ptr(self._buf), per // (D * inner), D, inner, 1 if diag_only else 0, _ptr(out), _ptr(part),
               B, per, self._stream())
        return out.view(B, -1)

    def product_overlap_device(self, vectors):
        """<v_0 x .. x v_{n-1}|psi> (pure) or <v|rho|v> (mixed) for one [D] vector per mode, as a device
        complex128 tensor [B] -- the contraction behind ``fidelity_coherent`` (states.py:687-728),
        done one mode at a time on the resident state instead of against a D^n product vector built
        on the host: the first launch reads the state once, every later one a D-th of the previous."""
        self._flush()
        n, D, B = self._num_modes, self._trunc, self._B
        cur = self._buf
        per = self._size()
        # (stride, lo, extent) of every tensor axis still present in ``cur``, per mode: one axis for kets,
        # (ket, bra) for density matrices; lo / extent = the index range held here (whole axes unless sharded)
        axes = {m: [(self._stride(ax),) + self._axis_range(ax) for ax in self._mode_axes(m)] for m in range(n)}
        for m in sorted(range(n), key=lambda q: axes[q][0][0]):  # innermost first: contiguous reads
            v = np.asarray(vectors[m], dtype=C128).reshape(D)
            (s0, lo0, e0) = axes[m][0]
            if self._pure:
                w, red, flags, base_b = v, [(e0, s0, 1)], L.FLAG_CONJ_B, lo0
            else:
                (s1, lo1, e1) = axes[m][1]
                w, red, flags = np.multiply.outer(v.conj(), v), [(e0, s0, D), (e1, s1, 1)], 0
                base_b = lo0 * D + lo1
            w = torch.from_numpy(np.ascontiguousarray(w)).to(self.device)
            del axes[m]
            rest = sorted(axes, key=lambda q: -axes[q][0][0])  # outermost first in the new tensor
            new_per = 1
            for q in rest:
                for _, _, e in axes[q]:
                    new_per *= e
            out = self._new(B * new_per)
            oa = [(B, per, 0, new_per)] if B > 1 else []
            new_axes, acc = {}, new_per
            for q in rest:
                new_axes[q] = []
                for st, lo, e in axes[q]:
                    acc //= e
                    oa.append((e, st, 0, acc))
                    new_axes[q].append((acc, lo, e))
            self._gather(cur, w, out, oa, red, flags=flags, base=(0, base_b, 0))
            cur, per, axes = out, new_per, new_axes
        return self._sum_over_ranks(cur)

    def _axis_range(self, axis):
        """(first index, extent) of tensor axis ``axis`` held by this process: the whole axis here"""
        return 0, self._trunc

    def _sum_over_ranks(self, t):
        return t

    # ------------------------------------------------------------------ Fock measurement (circuit.py:623-711)
    def _project_reset(self, modes, values):
        """|0..0><x| on ``modes`` (ops.py:179-198), out of place."""
        D = self._trunc
        base_a = 0
        for m, v in zip(modes, values):
            for ax in self._mode_axes(m):
                base_a += int(v) * self._stride(ax)
        if self._lazy_opt and self._fuse == "fold" and not self._batched:
            # lazy vacuum: the measured modes are |0> and unentangled again, so they simply leave the device
            # tensor (which shrinks by D per axis) and become product factors until a gate needs them
            gone = {ax for m in modes for ax in self._mode_axes(m)}
            keep = [ax for ax in self._phys if ax not in gone]
            out = self._new(D ** len(keep))
            oa = [(D, self._stride(ax), 0, D ** (len(keep) - 1 - j)) for j, ax in enumerate(keep)]
            self._gather(self._buf, None, out, oa, base=(base_a, 0, 0))
            self._buf, self._shared, self._scratch = out, False, None
            self._phys = keep
            self._pos = [None] * self._axes()
            for p, ax in enumerate(keep):
                self._pos[ax] = p
            self._inactive.update(modes)
            return
        out = self._get_scratch(self._buf.numel())
        L.call("b200_fill_zero", _ptr(out), out.numel(), self._stream())
        oa = []
        for m in range(self._num_modes):
            if m in modes:
                continue
            for ax in self._mode_axes(m):
                oa.append((D, self._stride(ax), 0, self._stride(ax)))
        self._gather(self._buf, None, out, oa, base=(base_a, 0, 0))
        if self._shared:  # the old buffer belongs to a state object now
            self._buf, self._scratch, self._shared = out, None, False
        else:
            self._buf, self._scratch = out, self._buf

    def _renormalise(self):
        self._own()
        nrm = self._norm_device(materialize=False)  # only ever called right after a flush + projection
        if float(nrm[0].item()) == 0:
            raise ZeroDivisionError("Measurement has zero probability.")
        L.call("b200_scale", _ptr(self._buf), self._buf.numel(), 1.0, 0.0, _ptr(nrm), 1 if self._pure else 0,
               self._stream())

    @staticmethod
    def _sample_index(dist):
        """One categorical draw from numpy's global stream, with the outcome ``circuit.py:682-686`` gets from the
        same stream: the reference tests ``sum(dist) != 1`` with the builtin ``sum`` over numpy scalars -- a plain
        left-to-right double sum, which is what ``np.cumsum(dist)[-1]`` computes bit for bit, 17x faster (2.7 ms ->
        0.16 ms for the 1e4 outcomes of a four-mode measurement, host time the GPU spends idle) -- and draws with
        ``np.random.choice(list(range(n)), p=..)``, which returns the same index as ``choice(n, p=..)``.
        ``tests/test_sampling.py`` compares the two on random distributions and seeds."""
        dist = np.asarray(dist, dtype=np.float64)
        total = np.cumsum(dist)[-1]
        return int(np.random.choice(len(dist), p=dist / total if total != 1 else dist))

    def _measure_fock_batched(self, modes, select):
        """Batched ``measure_fock`` (TF-backend semantics, ``tfbackend/circuit.py:612-760``): every batch entry
        is measured on its own -- ``select`` is one list for all entries or an array ``[B, len(modes)]`` --
        and the result has shape ``[B, len(modes)]``.  The marginals of all entries come from ONE device
        reduction; the draws are numpy's, entry by entry in batch order (one uniform each, as in the
        unbatched path); projection and renormalisation are launched per entry (the outcomes differ)."""
        D, B, k = self._trunc, self._B, len(modes)
        self._flush()
        self._canonicalize()
        if select is not None:
            sel = np.asarray(select)
            if np.any(sel == None):  # noqa: E711
                raise NotImplementedError("Post-selection lists must only contain numerical values.")
            if sel.shape == (k,):
                sel = np.vstack([sel] * B)
            if sel.shape != (B, k):
                raise ValueError("The shape of 'select' is incompatible with 'modes'.")
            outcomes = sel.astype(np.int64)
        else:
            keep = sorted(modes)
            dist_all = self.marginal_probs_device(keep).cpu().numpy()
            order = np.argsort(modes)
            outcomes = np.zeros((B, k), dtype=np.int64)
            for b in range(B):
                dist = dist_all[b] * ~np.isclose(dist_all[b], 0.0)
                i = self._sample_index(dist)
                digits = [i // D ** (k - 1 - j) % D for j in range(k)]
                for j in range(k):
                    outcomes[b, order[j]] = int(digits[j])
        # |0..0><outcome| on every entry: out of place into one zeroed buffer, one strided copy per entry
        per = self._size()
        out = self._get_scratch(self._buf.numel())
        L.call("b200_fill_zero", _ptr(out), out.numel(), self._stream())
        oa = [(D, self._stride(ax), 0, self._stride(ax)) for m in range(self._num_modes) if m not in modes
              for ax in self._mode_axes(m)]
        for b in range(B):
            base_a = b * per + sum(int(v) * self._stride(ax) for m, v in zip(modes, outcomes[b]) for ax in self._mode_axes(m))
            self._gather(self._buf, None, out, oa, base=(base_a, 0, b * per))
        if self._shared:
            self._buf, self._scratch, self._shared = out, None, False
        else:
            self._buf, self._scratch = out, self._buf
        nrm = self._norm_device(materialize=False)
        if bool((nrm == 0).any().item()):
            raise ZeroDivisionError("Measurement has zero probability.")
        for b in range(B):
            L.call("b200_scale", C.c_void_p(self._buf.data_ptr() + 16 * b * per), per, 1.0, 0.0,
                   C.c_void_p(nrm.data_ptr() + 8 * b), 1 if self._pure else 0, self._stream())
        self._touch(*modes)
        return outcomes

    def measure_fock(self, modes, select=None):
        if self._batched:
            return self._measure_fock_batched(list(modes), select)
        if select is not None and np.any(np.array(select) == None):  # noqa: E711
            raise NotImplementedError("Post-selection lists must only contain numerical values.")
        self._flush()
        D = self._trunc

        if select is not None:
            if len(select) != len(modes):
                raise ValueError(
                    "When performing post-selection, the number of "
                    "selected values (including None) must match the number of measured modes"
                )
            if not all(isinstance(s, int) or s is None for s in select):
                raise TypeError("The post-select list elements either be integers or None")
            measure = [i for i, s in zip(modes, select) if s is None]
            selected = [i for i, s in zip(modes, select) if s is not None]
            select_values = [s for s in select if s is not None]
            # NB: like the reference, the distribution below would be taken from the state BEFORE the
            # projection (circuit.py:632-635); with all-numeric `select` nothing is left to sample.
            self._project_reset(selected, select_values)
            self._renormalise()
            self._touch(*selected)
        else:
            measure = list(modes)

        if len(measure) > 0:
            keep = sorted(measure)
            dist = self.marginal_probs_device(keep)[0].cpu().numpy()
            # steps 3-6 of SURVEY Appendix B, with numpy itself so the draw is bit-identical
            dist = dist * ~np.isclose(dist, 0.0)
            i = self._sample_index(dist)
            digits =[i // D ** (len(measure) - 1 - m) % D for m in range(len(measure))]
            permutation = np.argsort(measure)
            outcome = [0] * len(measure)
            for j in range(len(measure)):
                outcome[permutation[j]] = int(digits[j])
            self._project_reset(measure, outcome)
            self._renormalise()
            self._touch(*measure)

        if select is not None:
            outcome = copy.copy(select)
        return np.array([outcome])


    # ------------------------------------------------------------------ homodyne (circuit.py:713-801)
    def measure_homodyne(self, phi, mode, select=None, **kwargs):
        """Homodyne measurement.  The single-mode marginal (a D x D reduced density matrix
        computed on the device) is sampled on the host grid exactly as the reference does; the
        conditional state is obtained by applying |0><x_phi| with the dense gate kernel."""
        import numbers

        if self._batched:
            return self._measure_homodyne_batched(phi, mode, select, **kwargs)
        self._flush()
        D = self._trunc
        w = 1 / self._hbar  # m omega / hbar
        if select is not None:
            if not isinstance(select, numbers.Number):
                raise TypeError("Selected measurement result must be of numeric type.")
            sample = float(select)
        else:
            rho = self.reduced_dm_device([mode])[0].cpu().numpy().reshape(D, D)
            ph = np.exp(-1j * phi * np.arange(D))
            rho = ph[:, None] * rho * ph.conj()[None, :]  # rotate to the measurement basis
            q_mag = kwargs.get("max", 10)
            num_bins = kwargs.get("num_bins", 100000)
            q = np.linspace(-q_mag, q_mag, num_bins)
            x = np.sqrt(w) * q
            H = [np.ones_like(x), 2 * x]
            for i in range(2, D):
                H.append(2 * x * H[i - 1] - 2 * (i - 1) * H[i - 2])
            scale = np.array([1 / np.sqrt(2.0 ** n * factorial(n)) for n in range(D)])
            Hn = np.array(H[:D]) * scale[:, None]  # normalised Hermite functions without the Gaussian
            pdf = np.einsum("nm,nq,mq->q", rho, Hn, Hn)
            pdf = pdf * (w / np.pi) ** 0.5 * np.exp(-w * q ** 2) * (q[1] - q[0])
            probs = pdf.real
            probs /= np.sum(probs)
            probs[np.abs(probs) < 1e-10] = 0
            hist = np.random.multinomial(1, probs)
            sample = self._agree_on(q[int(np.flatnonzero(hist)[0])])  # = list(hist).index(1), without the 1e5-element list

        inf_sq = np.array([(-0.5) ** (n // 2) * np.sqrt(factorial(n)) / factorial(n // 2) if n % 2 == 0 else 0.0
                           for n in range(D)], dtype=C128)
        alpha = sample * np.sqrt(w / 2)
        disp = self._gen1(L.GATE_DISPLACEMENT, float(np.abs(alpha)), float(np.angle(alpha)))[0].cpu().numpy()
        eig = (np.exp(1j * phi * np.arange(D))[:, None] * disp) @ inf_sq
        self._touch(mode)
        if self._lazy_opt and self._fuse == "fold" and type(self) is DeviceCircuit:
            # lazy vacuum: |0><x_phi| leaves the mode in |0>, unentangled -- contract the mode with <x_phi|
            # (one read of the state, a D-times smaller write) and let it leave the device tensor
            axes = self._mode_axes(mode)
            keep = [ax for ax in self._phys if ax not in axes]
            out = self._new(D ** len(keep))
            oa = [(D, self._stride(ax), 0, D ** (len(keep) - 1 - j)) for j, ax in enumerate(keep)]
            if self._pure:
                w, red, flags = eig, [(D, self._stride(axes[0]), 1)], L.FLAG_CONJ_B
            else:
                w = np.multiply.outer(eig.conj(), eig)
                red, flags = [(D, self._stride(axes[0]), D), (D, self._stride(axes[1]), 1)], 0
            w = torch.from_numpy(np.ascontiguousarray(w)).to(self.device)
            self._gather(self._buf, w, out, oa, red, flags=flags)
            self._buf, self._shared, self._scratch = out, False, None
            self._phys = keep
            self._pos = [None] * self._axes()
            for p, ax in enumerate(keep):
                self._pos[ax] = p
            self._inactive.add(mode)
            nrm = self._norm_device(materialize=False)
            L.call("b200_scale", _ptr(self._buf), self._buf.numel(), 1.0, 0.0, _ptr(nrm), 1 if self._pure else 0,
                   self._stream())
            return np.array([[sample]])
        proj = np.zeros((D, D), dtype=C128)
        proj[0, :] = eig.conj()
        self._apply_dense_now(self._upload_matrix(proj), mode)
        self._own()
        nrm = self._norm_device()
        L.call("b200_scale", _ptr(self._buf), self._buf.numel(), 1.0, 0.0, _ptr(nrm), 1 if self._pure else 0,
               self._stream())
        return np.array([[sample]])


    def _measure_homodyne_batched(self, phi, mode, select=None, **kwargs):
        """Batched homodyne (TF-backend semantics, ``tfbackend/circuit.py:812-941``): every entry is measured on
        its own.  One device reduction gives the B reduced density matrices; the draws are numpy's, entry by
        entry in batch order, on the reference Fock backend's grid (``circuit.py:713-801``); the projectors
        |0><x_phi| differ per entry, so they are applied as ONE per-entry dense table.  ``select``: a number
        for all entries or an array ``[B]``.  Returns ``[B, 1]``."""
        self._flush()
        D, B = self._trunc, self._B
        w = 1 / self._hbar
        if select is not None:
            sel = np.asarray(select)
            if not np.issubdtype(sel.dtype, np.number) or sel.dtype.kind == "c":
                raise TypeError("Selected measurement result must be of numeric type.")
            if sel.shape not in ((), (B,)):
                raise ValueError("'select' must be a number or have shape (batch_size,)")
            samples = np.broadcast_to(sel.astype(np.float64), (B,)).copy()
        else:
            rho_all = self.reduced_dm_device([mode]).cpu().numpy().reshape(B, D, D)
            ph = np.exp(-1j * phi * np.arange(D))
            q_mag = kwargs.get("max", 10)
            num_bins = kwargs.get("num_bins", 100000)
            q = np.linspace(-q_mag, q_mag, num_bins)
            x = np.sqrt(w) * q
            H = [np.ones_like(x), 2 * x]
            for i in range(2, D):
                H.append(2 * x * H[i - 1] - 2 * (i - 1) * H[i - 2])
            scale = np.array([1 / np.sqrt(2.0 ** n * factorial(n)) for n in range(D)])
            Hn = np.array(H[:D]) * scale[:, None]
            env = (w / np.pi) ** 0.5 * np.exp(-w * q ** 2) * (q[1] - q[0])
            samples = np.zeros(B)
            for b in range(B):
                rho = ph[:, None] * rho_all[b] * ph.conj()[None, :]
                probs = (np.einsum("nm,nq,mq->q", rho, Hn, Hn) * env).real
                probs /= np.sum(probs)
                probs[np.abs(probs) < 1e-10] = 0
                hist = np.random.multinomial(1, probs)
                samples[b] = q[int(np.flatnonzero(hist)[0])]
        inf_sq = np.array([(-0.5) ** (n // 2) * np.sqrt(factorial(n)) / factorial(n // 2) if n % 2 == 0 else 0.0
                           for n in range(D)], dtype=C128)
        alpha = samples * np.sqrt(w / 2)
        disp = self._gen1(L.GATE_DISPLACEMENT, np.abs(alpha), np.angle(alpha + 0j)).cpu().numpy().reshape(B, D, D)
        eig = np.einsum("n,bnm,m->bn", np.exp(1j * phi * np.arange(D)), disp, inf_sq)
        proj = np.zeros((B, D, D), dtype=C128)
        proj[:, 0, :] = eig.conj()
        self._touch(mode)
        self._apply_dense_now(torch.from_numpy(proj).to(self.device), mode)
        self._own()
        nrm = self._norm_device()
        if bool((nrm == 0).any().item()):
            raise ZeroDivisionError("Measurement has zero probability.")
        per = self._size()
        for b in range(B):
            L.call("b200_scale", C.c_void_p(self._buf.data_ptr() + 16 * b * per), per, 1.0, 0.0,
                   C.c_void_p(nrm.data_ptr() + 8 * b), 1 if self._pure else 0, self._stream())
        return samples.reshape(B, 1)

    def _agree_on(self, value):
        """A sampled outcome every process must share (sharded circuits broadcast rank 0's draw)."""
        return value

    def prepare_gkp(self, theta, phi, epsilon, ampl_cutoff, mode):
        """Finite-energy square-lattice GKP qubit state (circuit.py:803-812): a host-built ket."""
        self._prepare_ket(_square_gkp_state(theta, phi, epsilon, ampl_cutoff, self._trunc), mode)


# ---------------------------------------------------------------------- host-built single-mode kets
def _coherent(r, phi, D):
    alpha = r * np.exp(1j * phi)
    return np.exp(-abs(alpha) ** 2 / 2) * np.array(
        [alpha ** n / np.sqrt(factorial(n)) for n in range(D)], dtype=C128)


def _squeezed(r, theta, D):
    v = np.zeros(D, dtype=C128)
    for n in range(0, D, 2):
        m = n // 2
        v[n] = (np.sqrt(factorial(2 * m)) / (2 ** m * factorial(m))) * (-np.exp(1j * theta) * np.tanh(r)) ** m
    return np.sqrt(1 / np.cosh(r)) * v


def _square_gkp_state(theta, phi, epsilon, ampl_cutoff, D):
    """cos(theta/2)|0>_gkp + e^{-i phi} sin(theta/2)|1>_gkp with Fock-damped (epsilon) basis states:
    each basis state is a Gaussian-weighted comb of displaced squeezed states
    (fockbackend/ops.py:518-596)."""
    def basis(k):
        z_max = int(np.ceil(np.sqrt(-0.25 / np.pi * np.log(ampl_cutoff) / np.tanh(epsilon))))
        r = -0.5 * np.log(np.tanh(epsilon))
        ket = np.zeros(D, dtype=C128)
        for t in range(-z_max, z_max + 1):
            weight = np.exp(-0.5 * np.pi * np.tanh(epsilon) * (k + 2 * t) ** 2)
            alpha = np.sqrt(0.5 * np.pi) * (2 * t + k) / np.cosh(epsilon)
            ket = ket + weight * _displaced_squeezed(alpha, 0, r, 0, D)
        return ket

    ket = np.cos(theta / 2) * basis(0) + np.sin(theta / 2) * np.exp(-1j * phi) * basis(1)
    return ket / np.linalg.norm(ket)


def _displaced_squeezed(r_d, phi_d, r_s, phi_s, D):
    """fockbackend/ops.py:419-446 (Hermite-polynomial closed form)."""
    from numpy.polynomial.hermite import hermval

    if np.allclose(r_s, 0.0):
        return _coherent(r_d, phi_d, D)
    if np.allclose(r_d, 0.0):
        return _squeezed(r_s, phi_s, D)
    ph = np.exp(1j * phi_s)
    ch, sh, th = np.cosh(r_s), np.sinh(r_s), np.tanh(r_s)
    alpha = r_d * np.exp(1j * phi_d)
    gamma = alpha * ch + np.conj(alpha) * ph * sh
    harg = gamma / np.sqrt(ph * np.sinh(2 * r_s) + 1e-10)
    N = np.exp(-0.5 * np.abs(alpha) ** 2 - 0.5 * np.conj(alpha) ** 2 * ph * th)
    coeff = np.array([(0.5 * ph * th) ** (n / 2) / np.sqrt(factorial(n) * ch) for n in range(D)])
    return N * np.array([hermval(harg, row) for row in np.diag(coeff)])
