"""Differentiable (and batched) Fock circuits on the CUDA kernels -- SURVEY 8(f)3.

The reference gets gradients by running its TensorFlow backend: every gate is an einsum on a
``tf`` tensor and the gate tensors carry custom gradients from ``thewalrus.fock_gradients.grad_*``
(``/root/reference/strawberryfields/backends/tfbackend/ops.py:319-455``).  Here the same circuit
runs on the b200fock kernels and is differentiated by the adjoint method inside ONE
``torch.autograd.Function``:

forward    psi_k = U_k psi_{k-1}, one kernel pass per gate; psi_{k-1} is kept (a checkpoint);
backward   lambda_K = dL/d(psi_K) as torch hands it over; for k = K .. 1
               dL/dtheta = Re <lambda_k | (dU_k/dtheta) psi_{k-1}>   (one pass + one reduction)
               lambda_{k-1} = U_k^H lambda_k                          (one pass, adjoint table)

No gate is "uncomputed": truncated Fock gates are not unitary, the adjoint recursion above is
exact for any linear map.  dU/dtheta is exact for the truncated tables too: a parameter either
enters through a phase conjugation (d/dphi = i (m - n) U elementwise) or through the generator,
dU/dr = G U in the infinite space, whose first D rows need U at cutoff D + 1 (D + 2 for the
squeezer) -- the table generators are simply asked for the larger cutoff.

Gates: displacement, squeeze, rotation, kerr_interaction, cross_kerr_interaction, beamsplitter,
mzgate, two_mode_squeeze on pure states (vacuum input), scalar or per-batch-entry parameters.
"""
from __future__ import annotations

import numpy as np
import torch

from . import lib as L
from .circuit import DeviceCircuit, _ptr

C128 = torch.complex128

# name -> (class, kind, rule, number of parameters, extra cutoff needed by the generator derivative)
_GATES = {
    "displacement": ("dense", L.GATE_DISPLACEMENT, L.RULE_SINGLE, 2, 1),
    "squeeze": ("dense", L.GATE_SQUEEZE, L.RULE_SINGLE, 2, 2),
    "rotation": ("diag", L.DIAG_ROTATION, L.RULE_SINGLE, 1, 0),
    "kerr_interaction": ("diag", L.DIAG_KERR, L.RULE_SINGLE, 1, 0),
    "cross_kerr_interaction": ("diag2", L.DIAG_CROSS_KERR, L.RULE_SINGLE, 1, 0),
    "beamsplitter": ("pair", L.GATE_BEAMSPLITTER, L.RULE_SUM, 2, 1),
    "mzgate": ("pair", L.GATE_MZ, L.RULE_SUM, 2, 1),
    "two_mode_squeeze": ("pair", L.GATE_S2, L.RULE_DIFF, 2, 1),
}

_INDEX_CACHE = {}


def pair_index(rule, D, device):
    """Index maps of the block-packed two-axis layout (include/b200fock.h): ``flat`` = position of
    every packed entry in the dense [o1][i1][o2][i2] tensor; ``adj`` = permutation of the packed
    entries that transposes every block (out <-> in)."""
    key = (rule, D, str(device))
    if key not in _INDEX_CACHE:
        flat, adj, off = [], [], 0
        for b in range(2 * D - 1):
            lo = max(0, b - D + 1)
            c = min(b, 2 * D - 2 - b) + 1
            for r in range(c):
                for s in range(c):
                    o1, i1 = lo + r, lo + s
                    if rule == L.RULE_SUM:
                        o2, i2 = b - o1, b - i1
                    else:
                        o2, i2 = o1 - (b - D + 1), i1 - (b - D + 1)
                    flat.append(((o1 * D + i1) * D + o2) * D + i2)
                    adj.append(off + s * c + r)
            off += c * c
        assert off == L.packed_size(D)
        _INDEX_CACHE[key] = (torch.tensor(flat, dtype=torch.int64, device=device),
                             torch.tensor(adj, dtype=torch.int64, device=device))
    return _INDEX_CACHE[key]


def _sqrt_n(n, device):
    return torch.sqrt(torch.arange(n, dtype=torch.float64, device=device))


def _shift(T, axis, k):
    """S[.., m, ..] = T[.., m + k, ..] along ``axis`` (zero where m + k falls outside)."""
    n = T.shape[axis]
    out = torch.zeros_like(T)
    if abs(k) >= n:
        return out
    src = [slice(None)] * T.dim()
    dst = [slice(None)] * T.dim()
    if k > 0:
        src[axis], dst[axis] = slice(k, n), slice(0, n - k)
    else:
        src[axis], dst[axis] = slice(0, n + k), slice(-k, n)
    out[tuple(dst)] = T[tuple(src)]
    return out


def _bc(v, dim, axis):
    """reshape a 1-D vector so that it broadcasts along ``axis`` of a ``dim``-dimensional tensor"""
    shape = [1] * dim
    shape[axis] = -1
    return v.reshape(shape)


def _lower(T, axis):   # (a T)[m] = sqrt(m + 1) T[m + 1]
    s = _sqrt_n(T.shape[axis] + 1, T.device)[1:]
    return _bc(s, T.dim(), axis) * _shift(T, axis, 1)


def _raise(T, axis):   # (a^dagger T)[m] = sqrt(m) T[m - 1]
    s = _sqrt_n(T.shape[axis], T.device)
    return _bc(s, T.dim(), axis) * _shift(T, axis, -1)


def _number(T, axis):
    return _bc(torch.arange(T.shape[axis], dtype=torch.float64, device=T.device), T.dim(), axis)


def dense_derivatives(name, Ue, p, D):
    """d(table)/d(parameter) for a single-mode gate.  ``Ue``: [nb, De, De] (out, in) at the extended
    cutoff, ``p``: [2, nb] parameters.  Returns the two [nb, D, D] derivative tables."""
    ph = torch.exp(1j * p[1].to(C128)).reshape(-1, 1, 1)
    m = _number(Ue, 1) - _number(Ue, 2)                       # out - in
    if name == "displacement":   # D = exp(r (e^{i phi} a^dag - e^{-i phi} a))
        d0 = ph * _raise(Ue, 1) - ph.conj() * _lower(Ue, 1)
        d1 = 1j * m * Ue
    else:                        # S = exp(r (e^{-i theta} a^2 - e^{i theta} a^dag^2) / 2)
        d0 = 0.5 * (ph.conj() * _lower(_lower(Ue, 1), 1) - ph * _raise(_raise(Ue, 1), 1))
        d1 = 0.5j * m * Ue
    return [d[:, :D, :D].contiguous() for d in (d0, d1)]


def pair_derivatives(name, Ue, p, D):
    """d(table)/d(parameter) for a two-mode gate.  ``Ue``: dense [nb, De, De, De, De] as
    [o1][i1][o2][i2] at the extended cutoff.  Returns two dense [nb, D, D, D, D] tensors."""
    ph = torch.exp(1j * p[1].to(C128)).reshape(-1, 1, 1, 1, 1)
    o1, i1, o2, i2 = (_number(Ue, a) for a in (1, 2, 3, 4))
    if name == "beamsplitter":   # B = exp(theta (e^{i phi} a1 a2^dag - e^{-i phi} a1^dag a2))
        d0 = ph * _lower(_raise(Ue, 3), 1) - ph.conj() * _raise(_lower(Ue, 3), 1)
        d1 = 1j * (i1 - o1) * Ue
    elif name == "two_mode_squeeze":  # S2 = exp(r (e^{i theta} a1^dag a2^dag - e^{-i theta} a1 a2))
        d0 = ph * _raise(_raise(Ue, 3), 1) - ph.conj() * _lower(_lower(Ue, 3), 1)
        d1 = 0.5j * ((o1 + o2) - (i1 + i2)) * Ue
    else:  # MZ = B0 R1(phi_in) B0 R1(phi_ex), B0 = B(pi/4, pi/2): d/dphi_in = i b^dag b MZ, b = (a1 - i a2)/sqrt2
        d0 = 0.5j * ((o1 + o2) * Ue - 1j * _raise(_lower(Ue, 3), 1) + 1j * _lower(_raise(Ue, 3), 1))
        d1 = 1j * i1 * Ue
    return [d[:, :D, :D, :D, :D].contiguous() for d in (d0, d1)]


class TorchCircuit:
    """Record a parametrised gate list, then ``ket()`` runs it on the GPU and returns a torch tensor
    that is differentiable with respect to every parameter given as a ``requires_grad`` tensor.

    Parameters are python floats, 0-dim tensors, or (batched circuits) tensors of shape
    ``(batch_size,)``.  Method names and argument order follow the backend API
    (``/root/reference/strawberryfields/backends/base.py:155-621``)."""

    def __init__(self, num_modes, cutoff_dim, batch_size=None, device=None, checkpoint_every=None,
                 checkpoint_bytes=16 << 30):
        self.num_modes, self.cutoff, self.batch_size = int(num_modes), int(cutoff_dim), batch_size
        # Stored states: one per differentiable gate (checkpoint_every=1), or every c-th with the ones in
        # between recomputed during the backward pass (one more forward pass per gate, c + K/c states).
        # None: 1 while K states fit in ``checkpoint_bytes``, else ceil(sqrt(K)).
        self._every, self._ckpt_bytes = checkpoint_every, int(checkpoint_bytes)
        self._work = DeviceCircuit(num_modes, cutoff_dim, pure=True, batch_size=batch_size, device=device,
                                   fuse=False)
        self.device = self._work.device
        self._B = self._work._B
        self._tape = []      # (name, modes, [parameter slots])
        self._tensors = []   # differentiable / tensor-valued parameters, in slot order

    # ------------------------------------------------------------------ recording
    def _slot(self, v):
        if isinstance(v, torch.Tensor):
            if v.dim() > 1 or (v.dim() == 1 and v.shape[0] != self._B):
                raise ValueError("gate parameter must be a scalar or have shape (batch_size,)")
            self._tensors.append(v)
            return ("t", len(self._tensors) - 1)
        arr = np.asarray(v, dtype=np.float64)
        if arr.ndim > 1 or (arr.ndim == 1 and arr.shape[0] != self._B):
            raise ValueError("gate parameter must be a scalar or have shape (batch_size,)")
        return ("c", arr)

    def _record(self, name, params, modes):
        for m in modes:
            if not 0 <= m < self.num_modes:
                raise ValueError("The specified modes are not valid.")
        if len(set(modes)) != len(modes):
            raise ValueError("The specified modes are not valid.")
        self._tape.append((name, tuple(modes), [self._slot(p) for p in params]))

    def displacement(self, r, phi, mode):
        self._record("displacement", (r, phi), (mode,))

    def squeeze(self, r, phi, mode):
        self._record("squeeze", (r, phi), (mode,))

    def rotation(self, theta, mode):
        self._record("rotation", (theta,), (mode,))

    def kerr_interaction(self, kappa, mode):
        self._record("kerr_interaction", (kappa,), (mode,))

    def cross_kerr_interaction(self, kappa, mode1, mode2):
        self._record("cross_kerr_interaction", (kappa,), (mode1, mode2))

    def beamsplitter(self, theta, phi, mode1, mode2):
        self._record("beamsplitter", (theta, phi), (mode1, mode2))

    def mzgate(self, phi_in, phi_ex, mode1, mode2):
        self._record("mzgate", (phi_in, phi_ex), (mode1, mode2))

    def two_mode_squeeze(self, r, phi, mode1, mode2):
        self._record("two_mode_squeeze", (r, phi), (mode1, mode2))

    # ------------------------------------------------------------------ execution
    def ket(self):
        """The final ket, ``[D]*n`` (``[B] + [D]*n`` for a batched circuit), complex128 on the device."""
        out = _CircuitFn.apply(self, *self._tensors)
        shape = [self.cutoff] * self.num_modes
        return out.reshape(([self._B] if self.batch_size is not None else []) + shape)

    def _checkpoint_stride(self, K):
        if self._every is not None:
            return max(1, int(self._every))
        state_bytes = 16 * self._B * self.cutoff ** self.num_modes
        if K <= 1 or K * state_bytes <= self._ckpt_bytes:
            return 1
        return int(np.ceil(np.sqrt(K)))

    # values of a gate's parameters as a [2, nb] float64 device tensor
    def _values(self, slots, tensors):
        cols = []
        for kind, v in slots:
            t = tensors[v].detach() if kind == "t" else torch.from_numpy(np.ascontiguousarray(v))
            cols.append(t.to(device=self.device, dtype=torch.float64).reshape(-1))
        nb = max(c.numel() for c in cols)
        cols = [c.expand(nb) if c.numel() == 1 else c for c in cols]
        while len(cols) < 2:
            cols.append(torch.zeros(nb, dtype=torch.float64, device=self.device))
        return torch.stack(cols).contiguous()

    def _table(self, name, p, D):
        """Gate table at cutoff ``D`` for parameters ``p`` ([2, nb])."""
        cls, kind, rule, _, _ = _GATES[name]
        w, nb = self._work, p.shape[1]
        if cls == "dense":
            out = w._new(nb * D * D)
            L.call("b200_gen_gate1", kind, D, nb, 0.0, 0.0, _ptr(p), _ptr(out), w._stream())
            return out.view(nb, D, D)
        if cls in ("diag", "diag2"):
            per = D * D if cls == "diag2" else D
            out = w._new(nb * per)
            L.call("b200_gen_diag", kind, D, nb, 0.0, _ptr(p), _ptr(out), w._stream())
            return out.view(nb, per)
        out = w._new(nb * L.packed_size(D))
        L.call("b200_gen_gate2", kind, D, nb, 0.0, 0.0, _ptr(p), _ptr(out), w._stream())
        return out.view(nb, -1)

    def _apply(self, buf, name, modes, table, adjoint=False):
        """One kernel pass: buf <- table (or its adjoint) on ``modes``, in place."""
        cls, _, rule, _, _ = _GATES[name]
        w = self._work
        w._buf, w._shared = buf, False
        if cls == "dense":
            # conj_physical: the kernels read raw memory, torch's lazy conjugate bit would be lost
            U = torch.conj_physical(table).transpose(1, 2).contiguous() if adjoint else table
            w._k_gate1(U, modes[0], 0)
        elif cls == "diag":
            w._k_diag_multi([(torch.conj_physical(table) if adjoint else table, modes[0])])
        elif cls == "diag2":
            w._k_diag_pair(table, modes[0], modes[1], 1 if adjoint else 0)
        else:
            if adjoint:
                table = torch.conj_physical(table[:, pair_index(rule, self.cutoff, self.device)[1]]).contiguous()
            w._k_gate2(table, rule, modes[0], modes[1], 0)

    def _derivative_tables(self, name, p):
        """[d table / d p0, d table / d p1] in the layout ``_apply`` takes."""
        cls, kind, rule, npar, ext = _GATES[name]
        D = self.cutoff
        if cls in ("diag", "diag2"):
            tab = self._table(name, p, D)
            n = torch.arange(D, dtype=torch.float64, device=self.device)
            if name == "rotation":
                g = n
            elif name == "kerr_interaction":
                g = n * n
            else:
                g = (n[:, None] * n[None, :]).reshape(-1)
            return [(1j * g) * tab]
        De = D + ext
        if De > L.MAX_CUTOFF:
            raise ValueError("gradients need the gate table at cutoff %d > %d" % (De, L.MAX_CUTOFF))
        Te = self._table(name, p, De)
        if cls == "dense":
            return dense_derivatives(name, Te, p, D)
        nb = p.shape[1]
        flat_e, _ = pair_index(rule, De, self.device)
        dense = torch.zeros(nb, De ** 4, dtype=C128, device=self.device)
        dense[:, flat_e] = Te
        ds = pair_derivatives(name, dense.view(nb, De, De, De, De), p, D)
        flat, _ = pair_index(rule, D, self.device)
        return [d.reshape(nb, -1)[:, flat].contiguous() for d in ds]

    def _overlap_real(self, lam, t):
        """Re <lam | t> per batch entry (generic reduce kernel: sum lam * conj(t))."""
        w = self._work
        per = w._size()
        out = torch.zeros(self._B, dtype=torch.float64, device=self.device)
        oa = [(self._B, per, per, 1)] if self._B > 1 else []
        w._gather(lam, t, out, oa, [(per, 1, 1)], flags=L.FLAG_CONJ_B | L.FLAG_REAL_OUT)
        return out


class _CircuitFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prog, *tensors):
        w = prog._work
        w.reset()
        buf = w._buf
        tape = list(prog._tape)
        needs = [t.requires_grad for t in tensors]
        needed = [any(kind == "t" and needs[v] for kind, v in slots) for _, _, slots in tape]
        first = needed.index(True) if True in needed else len(tape)  # nothing before it is kept
        every = prog._checkpoint_stride(len(tape) - first)
        checkpoints, values = {}, []
        for k, (name, modes, slots) in enumerate(tape):
            p = prog._values(slots, tensors)
            if k >= first and (k - first) % every == 0:
                checkpoints[k] = buf.clone()       # the state BEFORE gate k
            prog._apply(buf, name, modes, prog._table(name, p, prog.cutoff))
            values.append(p)
        ctx.prog, ctx.checkpoints, ctx.values, ctx.tape = prog, checkpoints, values, tape
        ctx.first, ctx.every, ctx.needs = first, every, needs
        ctx.meta = [(t.shape, t.dtype, t.device) for t in tensors]
        return buf.clone().view(prog._B, -1)

    @staticmethod
    def backward(ctx, grad_ket):
        prog, tape, first = ctx.prog, ctx.tape, ctx.first
        if ctx.checkpoints is None:
            raise RuntimeError("TorchCircuit: the checkpoints were consumed by a previous backward pass")
        checkpoints, ctx.checkpoints = ctx.checkpoints, None
        lam = grad_ket.resolve_conj().to(dtype=C128, device=prog.device).contiguous().clone().reshape(-1)
        grads = [None] * len(ctx.meta)
        for start in reversed(range(first, len(tape), ctx.every)):
            end = min(start + ctx.every, len(tape))
            # states before the gates of this segment: the stored one, the others recomputed from it
            states = [checkpoints.pop(start)]
            for k in range(start, end - 1):
                nxt = states[-1].clone()
                prog._apply(nxt, tape[k][0], tape[k][1], prog._table(tape[k][0], ctx.values[k], prog.cutoff))
                states.append(nxt)
            for k in range(end - 1, start - 1, -1):
                (name, modes, slots), p = tape[k], ctx.values[k]
                psi = states.pop()
                wanted = [j for j, (kind, v) in enumerate(slots) if kind == "t" and ctx.needs[v]]
                if wanted:
                    dtabs = prog._derivative_tables(name, p)
                    for j in wanted:
                        t = psi.clone() if j != wanted[-1] else psi  # psi is dead after its last use
                        prog._apply(t, name, modes, dtabs[j])
                        g = prog._overlap_real(lam, t)                    # [B]
                        shape, dtype, device = ctx.meta[slots[j][1]]
                        g = g.sum() if len(shape) == 0 else g
                        g = g.reshape(shape).to(dtype=dtype, device=device)
                        idx = slots[j][1]
                        grads[idx] = g if grads[idx] is None else grads[idx] + g
                if k > first:  # nothing before the first differentiable gate needs lambda
                    prog._apply(lam, name, modes, prog._table(name, p, prog.cutoff), adjoint=True)
        return (None,) + tuple(grads)


# ---------------------------------------------------------------------------------------------
# Differentiable read-outs of the ket a TorchCircuit returns (plain torch reductions: the loss side of a
# training loop, cf. state.ket() / trace() / mean_photon() in the reference's TF examples).  ``ket`` is
# [D]*n or [B] + [D]*n; ``batched`` says which.
def _as_batch(ket, batched):
    return ket if batched else ket.unsqueeze(0)


def fock_probs(ket, batched=False):
    """|<n|psi>|^2, same shape as ``ket``"""
    return ket.real ** 2 + ket.imag ** 2


def trace(ket, batched=False):
    """squared norm (the probability kept inside the cutoff), one value per batch entry"""
    p = fock_probs(_as_batch(ket, batched))
    out = p.reshape(p.shape[0], -1).sum(dim=1)
    return out if batched else out[0]


def mean_photon(ket, mode, batched=False):
    """(mean, variance) of the photon number of ``mode``"""
    p = fock_probs(_as_batch(ket, batched))
    n_modes = p.dim() - 1
    marg = p.sum(dim=[1 + m for m in range(n_modes) if m != mode])
    n = torch.arange(marg.shape[1], dtype=marg.dtype, device=marg.device)
    mean = (marg * n).sum(dim=1)
    var = (marg * n ** 2).sum(dim=1) - mean ** 2
    return (mean, var) if batched else (mean[0], var[0])


def fidelity(ket, target, batched=False):
    """|<target|psi>|^2 for a target ket of the unbatched shape"""
    k = _as_batch(ket, batched)
    t = torch.as_tensor(target, dtype=k.dtype, device=k.device).reshape(-1)
    ov = (k.reshape(k.shape[0], -1) * t.conj()).sum(dim=1)
    out = ov.real ** 2 + ov.imag ** 2
    return out if batched else out[0]


# ---------------------------------------------------------------------------------------------
# CV quantum neural network layers -- the parameter layout of
# /root/reference/examples/quantum_neural_network.py:14-85 (BASELINE config 4)
def qnn_interferometer_size(N):
    return N * (N - 1) + max(1, N - 1)


def qnn_layer_size(N):
    return 2 * qnn_interferometer_size(N) + 4 * N


def qnn_interferometer(prog, params, modes):
    """Rectangular beamsplitter array of depth N on ``modes`` followed by N - 1 rotations;
    ``params`` = N(N-1)/2 angles, N(N-1)/2 phases, max(1, N-1) rotation angles."""
    N = len(modes)
    half = N * (N - 1) // 2
    theta, phi, rphi = params[:half], params[half:2 * half], params[2 * half:]
    if N == 1:
        prog.rotation(rphi[0], modes[0])
        return
    n = 0
    for l in range(N):
        for k in range(N - 1):
            if (l + k) % 2 != 1:
                prog.beamsplitter(theta[n], phi[n], modes[k], modes[k + 1])
                n += 1
    for i in range(max(1, N - 1)):
        prog.rotation(rphi[i], modes[i])


def qnn_layer(prog, params, modes=None):
    """One layer: interferometer, squeezers, interferometer, displacements, Kerr gates.  ``params``
    is a 1-D tensor (or a [size, batch] tensor for per-entry weights) of ``qnn_layer_size(N)`` rows."""
    modes = list(range(prog.num_modes)) if modes is None else list(modes)
    N = len(modes)
    M = qnn_interferometer_size(N)
    if len(params) != qnn_layer_size(N):
        raise ValueError("a %d-mode layer takes %d parameters" % (N, qnn_layer_size(N)))
    qnn_interferometer(prog, params[:M], modes)
    for i in range(N):
        prog.squeeze(params[M + i], 0.0, modes[i])
    qnn_interferometer(prog, params[M + N:2 * M + N], modes)
    for i in range(N):
        prog.displacement(params[2 * M + N + i], params[2 * M + 2 * N + i], modes[i])
        prog.kerr_interaction(params[2 * M + 3 * N + i], modes[i])


def qnn_init_weights(N, layers, active_sd=0.0001, passive_sd=0.1, generator=None, device=None):
    """[layers, qnn_layer_size(N)] float64 weights: N(0, passive_sd) for angles and phases,
    N(0, active_sd) for squeezing, displacement and Kerr magnitudes."""
    M = qnn_interferometer_size(N)
    sd = torch.tensor([passive_sd] * M + [active_sd] * N + [passive_sd] * M + [active_sd] * N
                      + [passive_sd] * N + [active_sd] * N, dtype=torch.float64)
    w = torch.randn(layers, qnn_layer_size(N), dtype=torch.float64, generator=generator) * sd
    return w.to(device) if device is not None else w
