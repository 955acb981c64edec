"""States sharded over several GPUs (one process per GPU, ``torch.distributed``).

The reference has no distributed path; this is the scale-out row of
the hot-path scope table (section 8e): a D^n ket -- or a D^2n density matrix with its
(ket_0, bra_0, ket_1, ..) axes -- that does not fit, or should not sit, on one GPU is cut
along its LEADING tensor axes.  Everything below speaks of tensor axes; for a ket axis m is
mode m, for a density matrix a gate is one queue entry on the ket axes and one (conjugated)
on the bra axes, and reductions walk the ket = bra diagonal across the ranks (``_walk``).

Layout.  The world size P is factored as p_0 * p_1 * ... * p_{g-1} with every p_k a divisor
of the cutoff D (D = 10: P = 2, 4, 8 -> (2), (2,2), (2,2,2)).  Physical axis k < g is split
as i_k = b_k * (D / p_k) + j_k; the digits (b_0 .. b_{g-1}) are the rank, so every rank
holds the tensor [D/p_0, .., D/p_{g-1}, D, .., D].  Axes g .. n-1 are whole on every rank:
a gate on modes stored there is a purely local launch of the single-GPU kernels.

Exchange.  A gate on a mode stored on a sharded axis first swaps ALL g sharded axes with g
local ones in one all-to-all (each rank keeps 1/P of its shard and sends (P-1)/P of it --
cheaper per mode than g pairwise half-shard swaps).  With the ranks' buffers mapped into each
other (CUDA IPC over NVLink / NVSwitch) it is a device-side barrier kernel plus one launch of
``b200_exchange_copy`` per part of the new shard -- bulk-copy-engine pulls straight from the
peers' HBM, no staging, no host synchronisation -- overlapped with the first gates behind it
(``_exchange_p2p``); without peer mapping: pack (strided gather) -> ``all_to_all_single``
(NCCL) -> unpack (``_exchange_nccl``).  Which logical mode lives on which physical axis is
tracked on the host; nothing is swapped back.

Scheduling.  Gates are queued (as in the single-GPU lazy queue) and, at a flush, executed
in dependency order preferring gates whose modes are local; when every runnable gate needs
a sharded mode an exchange brings the sharded modes in.  Which local modes it evicts -- and,
while the state is still the vacuum, which modes start out sharded -- is decided by
``exchange_plan.plan``: a bounded search seeded with the online farthest-next-use (Belady)
rule, with costs that grow when an exchange would move short contiguous runs.  The 9- and
10-mode interferometer circuits need 2 exchanges on 2 or 4 ranks and 3 on 8 ranks.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import lib as L
from . import exchange_plan as X
from . import scheduler as S
from .circuit import DeviceCircuit, _ptr


_PLANS = {}  # (queue structure, layout, g, fresh) -> exchange plan: a repeated circuit is planned once
_PLAN_WINDOW = 256  # queue entries planned (and searched) at a time
_KIND_DIAG2 = 4  # two-axis diagonal (cross-Kerr): a queue kind next to scheduler.KIND_SINGLE/SUM/DIFF/DIAG


def factor_world(P, D):
    """P = p_0 * ... * p_{g-1}, each p_k a divisor of D (largest divisors first)."""
    ps, rem = [], P
    while rem > 1:
        cand = [d for d in range(2, D + 1) if D % d == 0 and rem % d == 0]
        if not cand:
            raise ValueError("world size %d cannot be factored into divisors of the cutoff %d" % (P, D))
        ps.append(max(cand))
        rem //= ps[-1]
    return ps


class ShardedCircuit(DeviceCircuit):
    """``DeviceCircuit`` whose pure state is sharded over the ranks of a process group."""

    def __init__(self, num, trunc, group=None, **opts):
        if not dist.is_available() or not dist.is_initialized():
            raise L.B200Error("ShardedCircuit needs an initialised torch.distributed process group")
        self._pg = group
        self._rank = dist.get_rank(group)
        self._world = dist.get_world_size(group)
        self._set_factors(trunc)
        pure = opts.pop("pure", True)
        if self._g > (num if pure else 2 * num) - 1:
            raise ValueError("%d modes are too few to shard over %d ranks at cutoff %d" % (num, self._world, trunc))
        self.exchanges = 0
        self.exchange_bytes = 0
        # "p2p": ONE bulk-copy launch per exchange (or per part of the new shard when gates overlap it) pulls
        # from the peers' shards straight over NVLink -- no pack / unpack, no staging buffers, a device-side
        # barrier instead of host synchronisation; "push": the same launch with the DESTINATION remote (posted
        # stores into the peers' new shards; GPU-green, 711 against 673 GB/s on long runs but no overlap with
        # the gates behind it); "nccl": pack -> all_to_all_single -> unpack; "auto": p2p when every rank can
        # map every peer's buffers, else nccl
        self._xmode = opts.pop("exchange", "auto")
        if self._xmode not in ("auto", "p2p", "push", "nccl"):
            raise ValueError("exchange must be 'auto', 'p2p', 'push' or 'nccl'")
        self._p2p = False
        self._bufs = None
        # gates that follow an exchange run part by part while the remaining parts are still in flight
        # (peer-memory pull only): how many gates are executed that way, 0 = off.  Four passes over a shard
        # take about as long as its exchange (32 B x shard at ~6 TB/s against 16 B x shard at ~0.65 TB/s);
        # more would only trade full-size launches for small ones (measured on 8 GPUs: 8 gates save 1.9 ms
        # of 43.9 ms per circuit)
        self._overlap_ops = int(opts.pop("exchange_overlap", 4))
        self._xstream = None
        # CTAs that drive NVLink in b200_exchange_copy (0: the library default, 32).  One peer is saturated by 32
        # (tools/xchg_probe.py) and on 4 GPUs 32 and 48 give the same circuit time (77.3 / 77.6 ms); with 7 peers
        # interleaved 48 measured 618-647 GB/s against 539-561 (and 516 against 550 ms for the 10-mode circuit).
        import os

        env = os.environ.get("B200_EXCHANGE_CTAS")
        self.exchange_ctas = int(env) if env else (48 if self._world >= 8 else 0)
        opts.pop("batch_size", None)
        opts["fuse"] = "fold"
        # lazy vacuum, sharded flavour: the part of a fresh program that fits one GPU runs REPLICATED on
        # every rank as a small lazy-vacuum circuit (no communication); see _build_from_replicated
        self._lazy_shard = bool(opts.get("lazy_vacuum", False))
        opts["lazy_vacuum"] = False  # the sharded tensor itself always spans every mode
        super().__init__(num, trunc, pure=pure, **opts)

    def _set_factors(self, D):
        self._ps = factor_world(self._world, D)
        self._g = len(self._ps)
        self._digits = self._digits_of(self._rank)  # b_0 .. b_{g-1}, b_0 most significant

    def _digits_of(self, rank):
        digits, r = [], rank
        for p in reversed(self._ps):
            digits.append(r % p)
            r //= p
        return list(reversed(digits))

    # ------------------------------------------------------------------ geometry
    def _ext(self, pos):
        return self._trunc // self._ps[pos] if pos < self._g else self._trunc

    def _size(self):
        n = 1
        for pos in range(self._axes()):
            n *= self._ext(pos)
        return n

    def _local_stride(self, pos):
        s = 1
        for q in range(pos + 1, self._axes()):
            s *= self._ext(q)
        return s

    def _stride(self, axis):
        return self._local_stride(self._pos[axis])

    def _range(self, axis):
        """global index range [lo, hi) of tensor axis ``axis`` held by this rank"""
        pos = self._pos[axis]
        if pos >= self._g:
            return 0, self._trunc
        sub = self._trunc // self._ps[pos]
        return self._digits[pos] * sub, (self._digits[pos] + 1) * sub

    def _axis_range(self, axis):
        lo, hi = self._range(axis)
        return lo, hi - lo

    def _sum_over_ranks(self, t):
        dist.all_reduce(torch.view_as_real(t), group=self._pg)
        return t

    def _canonicalize(self):  # the sharded layout is never canonical; readers go through _stride()
        return

    _FLAG_PAD = 64  # complex128 entries behind every state buffer; buffer 0's hold the barrier counters

    def _alloc(self):
        """(Re)allocate the two ping-pong state buffers for the current geometry and map them into the
        peers when possible.  Collective.  A buffer that a state object returned earlier still shares
        (copy-on-write snapshots of the NCCL mode) is never reused: the reference's reset allocates a new
        array too."""
        size = self._size()
        if self._bufs is None or self._bufs[0].numel() != size or (getattr(self, "_shared", False) and not self._p2p):
            self._bufs = None
            self._buf = None
            self._backing = None
            self._p2p = False
            self._backing = [self._new(size + self._FLAG_PAD)]
            self._bufs = [self._backing[0][:size]]
            self._p2p = self._setup_p2p(size)
        self._cur = 0
        self._buf = self._bufs[0]
        self._scratch = self._send = self._recv = None
        self._shared = False

    def _to_mixed(self):
        """psi -> |psi><psi| with interleaved (ket, bra) axes (ops.py:110-120), sharded on the leading
        axes of the new 2n-axis tensor.  The ket is small next to the density matrix (D^n against
        D^2n / P per rank), so every rank gathers all of it and writes its own shard of the outer
        product; no exchange of the big tensor is needed."""
        if not self._pure:
            return
        self._flush()
        n, D = self._num_modes, self._trunc
        parts = [torch.empty_like(self._buf) for _ in range(self._world)]
        dist.all_gather(parts, self._buf, group=self._pg)
        full = torch.stack(parts).reshape(-1)          # [rank digits][local layout]
        # strides of mode m inside ``full``: position pos of the OLD layout; sharded axes split in (digit, rest)
        old_ls = [self._local_stride(p) for p in range(n)]
        old_pos, old_size = list(self._pos), self._size()
        rank_stride = []                               # stride of digit k in ``full``
        acc = old_size
        for p in reversed(self._ps):
            rank_stride.append(acc)
            acc *= p
        rank_stride.reverse()

        def psi_axes(m, lo, hi):
            """(ext, stride) pairs + base offset that walk psi along mode m over the global range [lo, hi)"""
            pos = old_pos[m]
            if pos >= self._g:
                return [(hi - lo, old_ls[pos])], lo * old_ls[pos]
            sub = D // self._ps[pos]
            if lo % sub == 0 and (hi - lo) % sub == 0:  # whole old digits: digit-major, then the remainder
                return [((hi - lo) // sub, rank_stride[pos]), (sub, old_ls[pos])], (lo // sub) * rank_stride[pos]
            if lo // sub == (hi - 1) // sub:            # inside one old digit
                return [(hi - lo, old_ls[pos])], (lo // sub) * rank_stride[pos] + (lo % sub) * old_ls[pos]
            raise NotImplementedError("pure -> mixed on a sharded state with these rank factors %r" % (self._ps,))

        self._pure = False
        self._phys = list(range(2 * n))
        self._pos = list(range(2 * n))
        self._bufs = None
        self._alloc()
        oa, base_a, base_b = [], 0, 0
        for ax in range(2 * n):
            lo, hi = self._range(ax)
            pairs, off = psi_axes(ax // 2, lo, hi)
            st = self._stride(ax)
            ext_left = hi - lo
            for e, s in pairs:                         # outer pair first: its output stride spans the inner one
                ext_left //= e
                if ax % 2 == 0:
                    oa.append((e, s, 0, st * ext_left))
                else:
                    oa.append((e, 0, s, st * ext_left))
            if ax % 2 == 0:
                base_a += off
            else:
                base_b += off
        self._gather(full, full, self._buf, oa, flags=L.FLAG_CONJ_B, base=(base_a, base_b, 0))
        self._fresh = False

    def reset(self, pure=None, cutoff_dim=None, num_subsystems=None):
        # same argument checks as DeviceCircuit.reset (circuit.py:89-116 of the reference)
        if pure is not None and not isinstance(pure, bool):
            raise ValueError("Argument 'pure' must be either True or False")
        if num_subsystems is not None and not isinstance(num_subsystems, int):
            raise ValueError("Argument 'num_subsystems' must be a positive integer")
        if cutoff_dim is not None:
            if not isinstance(cutoff_dim, int) or cutoff_dim < 1:
                raise ValueError("Argument 'cutoff_dim' must be a positive integer")
            if cutoff_dim > L.MAX_CUTOFF:
                raise ValueError("b200fock supports cutoff_dim <= {}".format(L.MAX_CUTOFF))
        if num_subsystems is not None:
            self._num_modes = num_subsystems
        if cutoff_dim is not None:
            if cutoff_dim != getattr(self, "_trunc", cutoff_dim):
                self._set_factors(cutoff_dim)
            self._trunc = cutoff_dim
        if pure is not None:
            self._pure = bool(pure)
        if self._g > self._axes() - 1:
            raise ValueError("%d modes are too few to shard over %d ranks at cutoff %d"
                             % (self._num_modes, self._world, self._trunc))
        self._phys = list(range(self._axes()))
        self._pos = list(range(self._axes()))
        self._alloc()
        L.call("b200_fill_zero", _ptr(self._buf), self._buf.numel(), self._stream())
        if self._rank == 0:
            L.call("b200_set_element", _ptr(self._buf), 0, 1.0, 0.0, self._stream())
        self._pending = {}
        self._opq = []
        self._untouched = set(range(self._num_modes))
        self._inactive = set()
        self._fresh = True  # still |0..0>: the first flush may choose which modes start out sharded

    # ------------------------------------------------------------------ queue: everything is deferred
    def _tile_mode(self):
        return False

    # Queue entries name TENSOR AXES (mode m for kets; 2m = ket, 2m+1 = bra for density matrices): that
    # is what the exchange planner and the kernels work on.  A mixed-state gate is two entries, U on the
    # ket axes and conj(U) on the bra axes, which the planner may schedule around different exchanges.
    def _emit_dense(self, U, mode):
        sz = self._trunc ** 2
        if self._pure:
            self._opq.append(S.Op(S.KIND_SINGLE, (mode,), U, 0, sz))
        else:
            self._opq.append(S.Op(S.KIND_SINGLE, (2 * mode,), U, 0, sz))
            self._opq.append(S.Op(S.KIND_SINGLE, (2 * mode + 1,), U, 1, sz))

    def _emit_diags(self, items):
        for d, mode in items:
            axes = (mode,) if self._pure else (2 * mode, 2 * mode + 1)
            op = S.Op(S.KIND_DIAG, axes, d, 0, self._trunc, 0.2)
            op.mode = mode
            self._opq.append(op)

    def _emit_pair(self, G, rule, m1, m2):
        sz = L.packed_size(self._trunc)
        if self._pure:
            self._opq.append(S.Op(rule, (m1, m2), G, 0, sz))
        else:
            self._opq.append(S.Op(rule, (2 * m1, 2 * m2), G, 0, sz))
            self._opq.append(S.Op(rule, (2 * m1 + 1, 2 * m2 + 1), G, 1, sz))

    def cross_kerr_interaction(self, kappa, mode1, mode2):
        """exp(i kappa n1 n2): a two-axis diagonal, queued like any other two-axis operator."""
        self._touch(mode1, mode2)
        self._flush([m for m in (mode1, mode2) if m in self._pending])  # keep program order on both modes
        tab = self._gen_diag(L.DIAG_CROSS_KERR, kappa)
        for conj, axes in enumerate(zip(self._mode_axes(mode1), self._mode_axes(mode2))):
            op = S.Op(_KIND_DIAG2, axes, tab, conj, self._trunc ** 2)
            op.kappa = kappa
            self._opq.append(op)

    def loss(self, T, mode):
        """circuit.py:617-621 as ONE queued pair operator on the (ket, bra) axes of the mode."""
        self._flush([mode])
        self._touch(mode)
        self._to_mixed()
        G = self._gen2(L.CHANNEL_LOSS, T)
        self._opq.append(S.Op(S.KIND_DIFF, (2 * mode, 2 * mode + 1), G, 0, L.packed_size(self._trunc)))

    def _exec(self, op):
        if op.kind == S.KIND_SINGLE:
            self._k_gate1(op.table, op.axes[0], op.conj)
        elif op.kind == S.KIND_DIAG:
            self._k_diag_multi([(op.table, op.mode)])   # both axes of the mode for a density matrix
        elif op.kind == _KIND_DIAG2:
            self._k_diag_pair(op.table, op.axes[0], op.axes[1], op.conj)
        else:
            self._k_gate2(op.table, op.kind, op.axes[0], op.axes[1], op.conj)

    def _run_queue(self):
        """Execute the queue in dependency order, local gates first, exchanging when stuck."""
        if not self._opq:
            return
        queue, self._opq = self._opq, []
        self._own()
        # very long programs are planned window by window: planning time stays bounded, the layout one
        # window ends in is where the next one starts
        for start in range(0, len(queue), _PLAN_WINDOW):
            self._run_ops(queue[start:start + _PLAN_WINDOW])

    def _run_ops(self, ops):
        # The plan (exchange_plan.py) is a pure function of the queue and the layout, so every rank
        # derives the same one.  It never evicts the innermost axis if that can be avoided: swapping it
        # would cut the exchange into 16*D/p-byte (80 B) runs -- measured 110 GB/s instead of 590 GB/s
        # over NVLink -- and while the state is still |0..0> it may also pick the layout.
        replicated = []
        if self._fresh and self._lazy_shard and self._pure:
            # lazy vacuum: gates that act while at most k_max modes are entangled run replicated
            k_max = self._num_modes
            while self._trunc ** k_max * self._world > self._trunc ** self._num_modes:
                k_max -= 1
            replicated, rest = X.replicated_prefix([op.axes for op in ops], k_max)
            all_ops, ops = ops, [ops[i] for i in rest]
        key = (tuple(op.axes for op in ops), tuple(self._phys), self._g, self._fresh)
        if key not in _PLANS:
            if len(_PLANS) >= 64:
                _PLANS.clear()
            try:
                _PLANS[key] = X.plan(key[0], self._phys, self._g, free_layout=self._fresh)
            except ValueError as exc:
                raise L.B200Error("%s (%d ranks)" % (exc, self._world)) from exc
        phys0, steps = _PLANS[key]
        if phys0 != self._phys:
            self._set_layout(phys0)
        if replicated:
            self._build_from_replicated([all_ops[i] for i in replicated])
        self._fresh = False
        done_early = 0
        for k, step in enumerate(steps):
            if step[0] == "run":
                for i in step[1][done_early:]:
                    self._exec(ops[i])
                done_early = 0
            else:
                # the first gates of the next run step can start on the parts of the new shard that have arrived
                nxt = steps[k + 1][1] if k + 1 < len(steps) and steps[k + 1][0] == "run" else []
                follow = [ops[i] for i in nxt[:self._overlap_ops]]
                done_early = self._exchange(sorted(self._pos[m] for m in step[1]), follow)

    def _build_from_replicated(self, ops):
        """Lazy vacuum on a sharded state.  ``ops`` (a dependency-closed prefix of a fresh program that
        never entangles more modes than fit one GPU) runs on every rank as a small lazy-vacuum
        ``DeviceCircuit`` -- redundantly, but on D^2 .. D^k_max tensors and without any exchange.  Then
        this rank's shard is written once: its slice of the small tensor, times the factors of the modes
        that are still untouched, straight into the layout the exchange planner chose for the rest."""
        n, D, g = self._num_modes, self._trunc, self._g
        rep = DeviceCircuit(n, D, pure=True, device=self.device, fuse="fold", lazy_vacuum=True)
        rep._defer_log = None  # the operators below are already in execution order: apply them as they come
        if self.__dict__.get("profile") is not None:
            rep.profile = self.profile
        for op in ops:
            if op.kind == S.KIND_SINGLE:
                rep._queue_dense(op.table, op.axes[0])
            elif op.kind == S.KIND_DIAG:
                rep._queue_diag(op.table, op.axes[0])
            elif op.kind == _KIND_DIAG2:
                rep.cross_kerr_interaction(op.kappa, op.axes[0], op.axes[1])
            else:
                rep._pair_gate(op.table, op.kind, op.axes[0], op.axes[1])
        active = list(rep._phys)
        rep._flush(active)  # pending operators of entangled modes; factored modes keep theirs as factors

        def extent(m):
            return self._ext(self._pos[m])

        def offset(m):  # first index of this rank's range on mode m
            pos = self._pos[m]
            return self._digits[pos] * (D // self._ps[pos]) if pos < g else 0

        def contiguous(modes):
            st, acc = {}, 1
            for m in reversed(modes):
                st[m] = acc
                acc *= extent(m)
            return st, acc

        inactive = [m for m in self._phys if m not in active]
        cur_modes = [m for m in self._phys if m in active]      # the shard's axis order
        cur, cs = rep._buf, {}
        if cur_modes:
            cs, size = contiguous(cur_modes)
            cur = self._buf if not inactive else self._new(size)
            oa = [(extent(m), rep._stride(m), 0, cs[m]) for m in cur_modes]
            self._gather(rep._buf, None, cur, oa, base=(sum(offset(m) * rep._stride(m) for m in cur_modes), 0, 0))
        for t, m in enumerate(inactive):
            pend = rep._pending.get(m)
            v = torch.zeros(D, dtype=torch.complex128, device=self.device)
            if pend is None:
                v[0] = 1.0
            elif pend[0] == "diag":
                v[0] = pend[1][0, 0]
            else:
                v = pend[1][0, :, 0].contiguous()
            new_modes = [x for x in self._phys if x == m or x in cur_modes]
            ns, size = contiguous(new_modes)
            out = self._buf if t == len(inactive) - 1 else self._new(size)
            oa = [(extent(x), 0, 1, ns[x]) if x == m else (extent(x), cs[x], 0, ns[x]) for x in new_modes]
            self._gather(cur, v, out, oa, base=(0, offset(m), 0))
            cur, cur_modes, cs = out, new_modes, ns

    def _set_layout(self, phys):
        self._phys = list(phys)
        pos = [0] * len(phys)
        for p, v in enumerate(phys):
            pos[v] = p
        self._pos = pos

    # ------------------------------------------------------------------ peer mapping (NVLink P2P)
    def _setup_p2p(self, size):
        """Map both state buffers of every rank into this process (CUDA IPC through torch's storage
        sharing) and enable peer access, so that a kernel here can read (or write) a peer's shard
        directly.  Every buffer is followed by a small pad; buffer 0's pad holds the counters of the
        device-side barrier (``b200_peer_barrier``).  Collective; returns True only if it worked on
        every rank."""
        if self._xmode == "nccl" or self._world == 1:
            return False
        if self.device.type != "cuda":
            # CPU test double (gloo tests, exchange="p2p" / "push" only): the "peer memory" is POSIX shared
            # memory, so the exchange logic below -- bases, strides, ping-pong -- runs without a GPU
            return self._xmode in ("p2p", "push") and self._setup_shared_host(size)
        ok = True
        peers = None
        total = size + self._FLAG_PAD
        # Every collective below is reached by every rank whatever fails locally (a rank that bailed
        # out early would leave the others waiting in all_gather_object).
        shared = None
        try:
            self._backing.append(self._new(total))
            self._bufs.append(self._backing[1][:size])
            L.call("b200_fill_zero", C.c_void_p(self._backing[0].data_ptr() + 16 * size), self._FLAG_PAD, self._stream())
            shared = [(b.untyped_storage()._share_cuda_(), b.storage_offset()) for b in self._backing]
        except Exception as exc:
            ok = False
            self._p2p_error = repr(exc)
        everyone = [None] * self._world
        dist.all_gather_object(everyone, (self.device.index, shared), group=self._pg)
        try:
            if any(sh is None for _, sh in everyone):
                raise RuntimeError("a peer could not export its buffers")
            peers = [[None] * self._world for _ in range(2)]
            for r, (dev, sh) in enumerate(everyone):
                for i in range(2):
                    if r == self._rank:
                        peers[i][r] = self._backing[i]
                        continue
                    L.call("b200_enable_peer_access", int(dev))
                    handle, offset = sh[i]
                    # open the IPC handle in THIS device's context (cudaIpcMemLazyEnablePeerAccess maps
                    # the exporter's memory for peer loads from here), not in the exporter's
                    handle = (self.device.index,) + tuple(handle[1:])
                    st = torch.UntypedStorage._new_shared_cuda(*handle)
                    typed = torch.storage.TypedStorage(wrap_storage=st, dtype=torch.complex128, _internal=True)
                    peers[i][r] = torch._utils._rebuild_tensor(typed, offset, (total,), (1,))
        except Exception as exc:  # no IPC / no peer path on this box: fall back to the NCCL exchange
            ok = False
            self._p2p_error = repr(exc)
        self._sync()  # the barrier counters are zero before any peer can bump them
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self._pg)
        ok = bool(flag.item())
        if ok:
            self._peers = peers
            self._epoch = 0
            pf = L.PeerFlags()
            pf.n_ranks, pf.rank = self._world, self._rank
            for r in range(self._world):
                pf.flags[r] = peers[0][r].data_ptr() + 16 * size
            self._pflags = pf
        else:
            if self._xmode in ("p2p", "push"):
                raise L.B200Error("peer-memory exchange requested but not available: %s"
                                  % getattr(self, "_p2p_error", "a peer failed"))
            self._bufs = self._bufs[:1]
            self._backing = self._backing[:1]
            self._peers = None
        return ok

    def _setup_shared_host(self, size):
        total = size + self._FLAG_PAD
        self._backing = [self._new(total).share_memory_() for _ in range(2)]
        self._bufs = [b[:size] for b in self._backing]
        shared = [(b.untyped_storage()._share_filename_cpu_(), b.storage_offset()) for b in self._backing]
        everyone = [None] * self._world
        dist.all_gather_object(everyone, shared, group=self._pg)
        self._peers = [[None] * self._world for _ in range(2)]
        for r, sh in enumerate(everyone):
            for i in range(2):
                if r == self._rank:
                    self._peers[i][r] = self._backing[i]
                    continue
                handle, offset = sh[i]
                st = torch.UntypedStorage._new_shared_filename_cpu(*handle)
                self._peers[i][r] = torch.empty(0, dtype=torch.complex128).set_(st, offset, (total,), (1,))
        self._epoch = 0
        return True

    def _sync(self):
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def _peer_barrier(self):
        """Every rank's stream has reached this point: a stream-ordered barrier kernel over peer-mapped
        counters -- the host neither synchronises nor blocks.  (CPU double: the process group.)"""
        self._epoch += 1
        if self.device.type != "cuda":
            dist.barrier(group=self._pg)
            return
        L.call("b200_peer_barrier", C.byref(self._pflags), self._epoch, 60.0, self._stream())

    def _exchange_p2p(self, T, follow=(), new_phys=None):
        """Exchange without staging buffers and without host stalls: after a device-side barrier the launches
        of ``b200_exchange_copy`` move this rank's whole share -- for every source rank a strided block of
        the old layout goes where it belongs in the new shard (the other ping-pong buffer), sources
        interleaved so that every NVLink peer is busy at once.  "p2p" pulls (the peers' old shards are the
        sources), "push" posts stores into the peers' new shards.

        Overlap (pull only).  The new shard is cut along its outermost axis into its ``D / p_0`` index ranges
        ("parts": contiguous, and independent of each other for every gate, because no gate acts on a sharded
        axis).  Part c is pulled by its own launch on a side stream; as soon as it has landed, the gates of
        ``follow`` run on it (launch views, ``DeviceCircuit._vptr``) while parts c+1.. are still crossing
        NVLink: the link drives 48 CTAs, the other 100 SMs compute.  Returns ``len(follow)`` when it did."""
        g, n, D = self._g, self._axes(), self._trunc
        size = self._size()
        ls = [self._local_stride(p) for p in range(n)]
        sub = [D // p for p in self._ps]
        prof = self.__dict__.get("profile")
        push = self._xmode == "push"
        cuda = self.device.type == "cuda"
        parts = sub[0] if (follow and not push and sub[0] > 1) else 1
        if prof is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        # pull: every rank's current shard is final and nobody still reads this rank's other buffer (it was
        # the source of the previous exchange).  push: nobody still uses the buffer the peers will write.
        self._peer_barrier()
        dst = self._bufs[1 - self._cur]

        def block(digits):
            return sum(digits[k] * sub[k] * ls[T[k]] for k in range(g))

        def descriptor(part):
            """block copy of the whole exchange (part None) or of one index of the new outermost axis"""
            axes = []  # (extent, source stride, destination stride), outermost first
            for pos in range(n):
                if pos < g:                       # new sharded remainder m_k <- source axis T[k]
                    axes.append((sub[pos], ls[T[pos]], ls[pos]))
                elif pos in T:                    # new local axis T[k] <- source's remainder j_k on axis k
                    k = T.index(pos)
                    axes.append((sub[k], ls[k], ls[pos]))
                else:
                    axes.append((D, ls[pos], ls[pos]))
            off_s = off_d = 0
            if part is not None:
                off_s, off_d = part * axes[0][1], part * axes[0][2]
                axes[0] = (1, 0, 0)
            merged = []
            for a in axes:
                if a[0] == 1:
                    continue
                if merged and merged[-1][1] == a[1] * a[0] and merged[-1][2] == a[2] * a[0]:
                    merged[-1] = (merged[-1][0] * a[0], a[1], a[2])
                else:
                    merged.append(a)
            run = 1
            if merged and merged[-1][1] == 1 and merged[-1][2] == 1:
                run = merged.pop()[0]
            if len(merged) > L.XCHG_MAX_AXES or self._world > L.XCHG_MAX_PEERS:
                raise L.B200Error("exchange geometry exceeds the copy kernel's limits")
            d = L.XchgDesc()
            d.n_axes, d.n_src, d.first_src, d.run = len(merged), self._world, self._rank, run
            for j, (e, a, b) in enumerate(merged):
                d.ext[j], d.ss[j], d.ds[j] = e, a, b
            span_s = sum((e - 1) * a for e, a, _ in merged) + run - 1
            span_d = sum((e - 1) * b for e, _, b in merged) + run - 1
            for s in range(self._world):
                sd = self._digits_of(s)
                if push:   # my block for rank s goes straight into ITS new shard
                    d.src[s], d.dst[s] = self._buf.data_ptr(), self._peers[1 - self._cur][s].data_ptr()
                    d.src_base[s], d.dst_base[s] = block(sd) + off_s, block(self._digits) + off_d
                else:      # the receiver's digit selects the block, the sender's digit lands on axis T[k]
                    d.src[s], d.dst[s] = self._peers[self._cur][s].data_ptr(), dst.data_ptr()
                    d.src_base[s], d.dst_base[s] = block(self._digits) + off_s, block(sd) + off_d
                if d.src_base[s] + span_s >= size or d.dst_base[s] + span_d >= size:
                    raise L.B200Error("exchange block of rank %d leaves the shard" % s)
            return d

        ctas = int(self.__dict__.get("exchange_ctas", 0))
        if parts == 1:
            # this rank's own block stays in local HBM: plain copy by the CTAs the link does not need
            L.call("b200_exchange_copy", C.byref(descriptor(None)), self._rank, ctas, self._stream())
            if push:
                self._peer_barrier()  # every peer's stores into my new shard have landed
            done = 0
        else:
            main = torch.cuda.current_stream(self.device) if cuda else None
            if cuda:
                if self._xstream is None:
                    self._xstream = torch.cuda.Stream(self.device)
                ready = torch.cuda.Event()
                ready.record(main)              # behind the barrier: the peers' shards are final
                self._xstream.wait_event(ready)
            arrived = []
            for c in range(parts):
                d = descriptor(c)
                if cuda:
                    with torch.cuda.stream(self._xstream):
                        if prof is not None and c == 0:
                            x0 = torch.cuda.Event(enable_timing=True)
                            x0.record(self._xstream)
                        L.call("b200_exchange_copy", C.byref(d), self._rank, ctas, self._stream())
                        ev = torch.cuda.Event(enable_timing=prof is not None and c == parts - 1)
                        ev.record(self._xstream)
                    arrived.append(ev)
                else:
                    L.call("b200_exchange_copy", C.byref(d), self._rank, ctas, self._stream())
            if prof is not None and cuda:
                # the transfer alone (first part issued .. last part landed), measured on the side stream while
                # the main stream computes
                prof.append(("exchange", 16 * size * (self._world - 1) // self._world, x0, arrived[-1]))
            # the gates of `follow` on part c of the NEW shard, in the NEW layout
            old_layout, old_cur, old_buf = list(self._phys), self._cur, self._buf
            self._cur ^= 1
            self._buf = dst
            self._set_layout(new_phys)
            part_elems = size // parts
            try:
                for c in range(parts):
                    if cuda:
                        main.wait_event(arrived[c])
                    self._view = (c * part_elems, part_elems)
                    for op in follow:
                        self._exec(op)
            finally:
                self._view = None
                self._set_layout(old_layout)   # _exchange() installs the new layout and buffer for good
                self._cur, self._buf = old_cur, old_buf
            done = len(follow)
        if prof is not None:
            ev1.record()
            nbytes = 16 * size * (self._world - 1) // self._world
            tag = "exchange/p2p_push" if push else ("exchange/p2p_pull" if parts == 1 else "exchange/p2p_pull+gates")
            prof.append((tag, nbytes, ev0, ev1))
            if parts == 1:
                prof.append(("exchange", nbytes, ev0, ev1))
        self._cur ^= 1
        self._buf = dst
        return done

    # ------------------------------------------------------------------ the exchange
    def _contig(self, exts):
        st, acc = [], 1
        for e in reversed(exts):
            st.append(acc)
            acc *= e
        return list(reversed(st))

    def _exchange(self, T, follow=()):
        """Swap sharded axis k with local axis T[k] for every k.  ``follow``: the gates that come next; the
        peer-memory pull exchange may run them part by part as the parts of the new shard arrive.  Returns
        how many of them it has executed."""
        g, n = self._g, self._axes()
        assert len(T) == g and all(t >= g for t in T)
        self._own()  # the unpack writes in place: never into a buffer a state object still shares
        phys = list(self._phys)
        for k in range(g):
            phys[k], phys[T[k]] = phys[T[k]], phys[k]
        done = 0
        if self._p2p:
            done = self._exchange_p2p(T, follow, phys)
        else:
            self._exchange_nccl(T)
        self._set_layout(phys)
        self._fresh = False
        self.exchanges += 1
        self.exchange_bytes += 16 * self._size() * (self._world - 1) // self._world
        return done

    def _exchange_nccl(self, T):
        """pack (strided gather) -> all_to_all_single -> unpack (strided gather)."""
        g, n, D = self._g, self._axes(), self._trunc
        size = self._size()
        if self._send is None:
            self._send, self._recv = self._new(size), self._new(size)
        prof = self.__dict__.get("profile")
        if prof is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        ls = [self._local_stride(p) for p in range(n)]
        sub = [D // p for p in self._ps]
        # ---- pack: send[c_0..c_{g-1}][j_0..j_{g-1}][local axes, T[k] restricted to m_k] ----
        exts, src = [], []
        for k in range(g):
            exts.append(self._ps[k])
            src.append(sub[k] * ls[T[k]])
        for k in range(g):
            exts.append(sub[k])
            src.append(ls[k])
        for pos in range(g, n):
            if pos in T:
                exts.append(sub[T.index(pos)])
            else:
                exts.append(D)
            src.append(ls[pos])
        dstc = self._contig(exts)
        self._gather(self._buf, None, self._send, [(e, s, 0, c) for e, s, c in zip(exts, src, dstc)])
        if prof is not None:
            eva = torch.cuda.Event(enable_timing=True)
            eva.record()
        # ---- all-to-all: block q of `send` goes to rank q ----
        dist.all_to_all_single(self._recv.view(torch.float64), self._send.view(torch.float64), group=self._pg)
        if prof is not None:
            evb = torch.cuda.Event(enable_timing=True)
            evb.record()
            prof.append(("exchange/pack", 32 * size, ev0, eva))
            prof.append(("exchange/all_to_all", 16 * size * (self._world - 1) // self._world, eva, evb))
        # ---- unpack: recv[b_0..b_{g-1}][j_0..][.. m_k at T[k] ..] -> new local tensor ----
        rs = dict()  # name -> stride inside recv
        names = [("b", k) for k in range(g)] + [("j", k) for k in range(g)] + [("l", pos) for pos in range(g, n)]
        for nm, st in zip(names, dstc):
            rs[nm] = st
        oa = []
        for pos in range(n):
            if pos < g:
                oa.append((sub[pos], rs[("l", T[pos])], 0, ls[pos]))       # m_k becomes the sharded remainder
            elif pos in T:
                k = T.index(pos)
                oa.append((self._ps[k], rs[("b", k)], 0, sub[k] * ls[pos]))  # sender's rank digit
                oa.append((sub[k], rs[("j", k)], 0, ls[pos]))                # sender's local remainder
            else:
                oa.append((D, rs[("l", pos)], 0, ls[pos]))
        self._gather(self._recv, None, self._buf, oa)
        if prof is not None:
            ev1.record()
            prof.append(("exchange/unpack", 32 * size, evb, ev1))
            prof.append(("exchange", 16 * size * (self._world - 1) // self._world, ev0, ev1))

    # ------------------------------------------------------------------ observation
    def _walk(self, mode):
        """How this rank walks the photon number of ``mode`` through its shard: the numbers
        lo .. lo + ext - 1 sit at local offsets base + j * stride.  Kets: the rank's range of the mode's
        axis.  Density matrices: the diagonal ket = bra, i.e. the intersection of the two axes' ranges
        (ext = 0 when this rank holds no diagonal entry of the mode).  Returns (lo, ext, stride, base)."""
        if self._pure:
            lo, hi = self._range(mode)
            return lo, hi - lo, self._stride(mode), 0
        (kl, kh), (bl, bh) = self._range(2 * mode), self._range(2 * mode + 1)
        lo, hi = max(kl, bl), min(kh, bh)
        if hi <= lo:
            return 0, 0, 0, 0
        sk, sb = self._stride(2 * mode), self._stride(2 * mode + 1)
        return lo, hi - lo, sk + sb, (lo - kl) * sk + (lo - bl) * sb

    def _norm_device(self, materialize=True):
        """squared norm (kets) or trace (density matrices) of the whole state, on every rank"""
        self._flush()
        out = torch.zeros(1, dtype=torch.float64, device=self.device)
        if self._pure:
            L.call("b200_norm2", _ptr(self._buf), self._size(), _ptr(out), _ptr(self._norm_part), self._stream())
        else:
            walks = [self._walk(m) for m in range(self._num_modes)]
            if all(w[1] > 0 for w in walks):
                self._gather(self._buf, None, out, [], [(w[1], w[2], 0) for w in walks], flags=L.FLAG_REAL_OUT,
                             base=(sum(w[3] for w in walks), 0, 0))
        dist.all_reduce(out, group=self._pg)
        return out

    def norm(self):
        v = float(self._norm_device().cpu().numpy()[0])
        return float(np.sqrt(v)) if self._pure else v

    def element(self, n):
        """<n|psi> (kets) or <n|rho|n> (density matrices): read on the owning rank, shared with all."""
        self._flush()
        idx, mine = 0, True
        for mode, x in enumerate(n):
            lo, ext, stride, base = self._walk(mode)
            if not lo <= int(x) < lo + ext:
                mine = False
                break
            idx += base + (int(x) - lo) * stride
        val = torch.zeros(2, dtype=torch.float64, device=self.device)
        if mine:
            val = torch.view_as_real(self._buf[idx:idx + 1]).reshape(2).clone()
        dist.all_reduce(val, group=self._pg)
        v = val.cpu().numpy()
        return np.array([v[0] + 1j * v[1]])

    def host_state(self):
        """The whole ket / density matrix on every rank's host (small states / tests only)."""
        self._flush()
        g, n, D = self._g, self._axes(), self._trunc
        parts = [torch.empty_like(self._buf) for _ in range(self._world)]
        dist.all_gather(parts, self._buf, group=self._pg)
        full = torch.stack(parts).cpu().numpy()
        local_ext = [self._ext(p) for p in range(n)]
        full = full.reshape(self._ps + local_ext)  # [b_0..b_{g-1}, j_0..j_{g-1}, local...]
        order = []
        for k in range(g):
            order += [k, g + k]
        order += list(range(2 * g, g + n))
        full = full.transpose(order).reshape([D] * n)  # physical axis order
        return np.ascontiguousarray(full.transpose([self._pos[m] for m in range(n)]))

    def is_vacuum(self, tol):
        v = self.element([0] * self._num_modes)[0]
        fid = np.abs(v) ** 2 if self._pure else v.real
        return bool(np.abs(fid - 1) <= tol)

    def snapshot(self):
        self._flush()
        snap = object.__new__(ShardedCircuit)
        snap.__dict__.update(self.__dict__)
        snap._pending, snap._opq, snap._untouched = {}, [], set()
        snap._scratch = snap._send = snap._recv = snap._part = None
        snap._norm_part = torch.zeros(4096, dtype=torch.float64, device=self.device)
        if self._p2p:
            # the live buffers are mapped into the peers and reused in place: the state object gets
            # its own copy instead of a copy-on-write share
            snap._buf = self._buf.clone()
            snap._bufs, snap._peers, snap._p2p = None, None, False
            snap._shared = False
        else:
            snap._shared = True
            self._shared = True
        return snap

    # ------------------------------------------------------------------ reductions / measurement
    def marginal_probs_device(self, keep):
        """Photon-number distribution of the (sorted) modes ``keep``, float64 [1, D^k], the same on
        every rank: each rank reduces its shard into its own block of the table, then one all-reduce of
        D^k doubles."""
        self._flush()
        D, k = self._trunc, len(keep)
        out = torch.zeros(D ** k, dtype=torch.float64, device=self.device)
        walks = {m: self._walk(m) for m in range(self._num_modes)}
        if all(w[1] > 0 for w in walks.values()):  # else: no diagonal entry of rho lives on this rank
            sb = 1 if self._pure else 0             # kets: |psi|^2 = psi * conj(psi), same strides for B
            oa, base_c = [], 0
            for j, m in enumerate(keep):
                lo, ext, stride, _ = walks[m]
                oa.append((ext, stride, stride * sb, D ** (k - 1 - j)))
                base_c += lo * D ** (k - 1 - j)
            red = [(w[1], w[2], w[2] * sb) for m, w in walks.items() if m not in keep]
            base_a = sum(w[3] for w in walks.values())
            if self._pure:
                self._gather(self._buf, self._buf, out, oa, red, flags=L.FLAG_CONJ_B | L.FLAG_REAL_OUT,
                             base=(base_a, base_a, base_c))
            else:
                self._gather(self._buf, None, out, oa, red, flags=L.FLAG_REAL_OUT, base=(base_a, 0, base_c))
        dist.all_reduce(out, group=self._pg)
        return out.view(1, -1)

    def fock_probs_device(self):
        """all_fock_probs on every rank (small states only: D^n doubles per rank)."""
        return self.marginal_probs_device(list(range(self._num_modes)))

    def _make_local(self, axes):
        """Bring every tensor axis of ``axes`` onto a whole (unsharded) position."""
        g, n = self._g, self._axes()
        if all(self._pos[a] >= g for a in axes):
            return
        cand = [p for p in range(g, n) if self._phys[p] not in axes]
        if len(cand) < g:
            raise NotImplementedError("too many measured modes for a %d-rank sharded state" % self._world)
        cand.sort(key=lambda p: (p == n - 1, p))  # keep the innermost axis resident if possible
        self._exchange(sorted(cand[:g]))

    def _sample_index(self, dist_):
        """Every rank holds the same distribution and draws from its own numpy stream (so a seeded
        program advances the stream on every rank as the reference does); rank 0's draw is
        authoritative."""
        i = DeviceCircuit._sample_index(dist_)
        t = torch.tensor([int(i)], dtype=torch.int64, device=self.device)
        dist.broadcast(t, src=dist.get_global_rank(self._pg, 0) if self._pg is not None else 0, group=self._pg)
        return int(t.item())

    def _project_reset(self, modes, values):
        """|0..0><x| on whole-axis modes: rank-local, out of place (into the other ping-pong buffer
        when the buffers are mapped into the peers)."""
        n = self._axes()
        axes_of = {m: self._mode_axes(m) for m in modes}
        assert all(self._pos[a] >= self._g for ax in axes_of.values() for a in ax)
        self._fresh = False
        if self._p2p:
            self._peer_barrier()  # a slower peer may still be pulling from the other ping-pong buffer
            out = self._bufs[1 - self._cur]
        else:
            out = self._get_scratch(self._buf.numel())
        L.call("b200_fill_zero", _ptr(out), out.numel(), self._stream())
        base_a = sum(int(v) * self._stride(a) for m, v in zip(modes, values) for a in axes_of[m])
        mpos = {self._pos[a] for ax in axes_of.values() for a in ax}
        oa = [(self._ext(p), self._local_stride(p), 0, self._local_stride(p)) for p in range(n) if p not in mpos]
        self._gather(self._buf, None, out, oa, base=(base_a, 0, 0))
        if self._p2p:
            self._cur ^= 1
            self._buf = out
        elif self._shared:
            self._buf, self._scratch, self._shared = out, None, False
        else:
            self._buf, self._scratch = out, self._buf
            self._bufs = [self._buf]

    def measure_fock(self, modes, select=None):
        self._flush()
        self._make_local([a for m in modes for a in self._mode_axes(m)])
        return DeviceCircuit.measure_fock(self, modes, select)

    def reduced_dm_device(self, keep):
        """Reduced density matrix of the (sorted) modes ``keep`` -> complex128 [1, D^2k] with interleaved
        (ket, bra) axes, the same on every rank (states.py:613-642).  The kept modes' axes are made local
        first; every rank then contracts its part of the traced modes and D^2k numbers are all-reduced."""
        self._flush()
        self._make_local([a for m in keep for a in self._mode_axes(m)])
        D, k, n = self._trunc, len(keep), self._num_modes
        out = torch.zeros(D ** (2 * k), dtype=torch.complex128, device=self.device)
        rest = [m for m in range(n) if m not in keep]
        oa = []
        for j, m in enumerate(keep):
            axes = self._mode_axes(m)
            wk, wb = D ** (2 * k - 1 - 2 * j), D ** (2 * k - 2 - 2 * j)
            if self._pure:
                oa += [(D, self._stride(axes[0]), 0, wk), (D, 0, self._stride(axes[0]), wb)]
            else:
                oa += [(D, self._stride(axes[0]), 0, wk), (D, self._stride(axes[1]), 0, wb)]
        if self._pure:
            red = [(self._ext(self._pos[m]), self._stride(m), self._stride(m)) for m in rest]
            self._gather(self._buf, self._buf, out, oa, red, flags=L.FLAG_CONJ_B)
        else:
            walks = [self._walk(m) for m in rest]
            if all(w[1] > 0 for w in walks):
                self._gather(self._buf, None, out, oa, [(w[1], w[2], 0) for w in walks],
                             base=(sum(w[3] for w in walks), 0, 0))
        ri = torch.view_as_real(out)
        dist.all_reduce(ri, group=self._pg)
        return out.view(1, -1)

    def prepare_multimode(self, state, modes, input_state_is_pure=None):
        """Single-mode kets on modes that are still the untouched vacuum (the inputs of a boson-sampling
        program): |v> = (|v><0|) |0>, so the preparation is queued as a rank-one single-mode operator and
        costs what a gate costs.  Anything else needs a partial trace of the sharded state (D^2(n-k) entries: it
        would not fit where sharding is needed) and is refused."""
        modes = [modes] if isinstance(modes, int) else list(modes)
        D = self._trunc
        if len(modes) == 1 and modes[0] in self._untouched and self._pure and np.shape(state) == (D,):
            tab = np.zeros((D, D), dtype=np.complex128)
            tab[:, 0] = np.asarray(state, dtype=np.complex128)
            self._queue_dense(self._upload_matrix(tab), modes[0])
            return
        self._unsupported()

    # homodyne (DeviceCircuit.measure_homodyne): the D x D marginal comes from reduced_dm_device above, the
    # sample is rank 0's, and the projector |0><x_phi| is one more queued single-mode operator
    def _agree_on(self, value):
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device)
        dist.broadcast(t, src=dist.get_global_rank(self._pg, 0) if self._pg is not None else 0, group=self._pg)
        return float(t.item())

    def _apply_dense_now(self, U, mode):
        self._emit_dense(U, mode)
        self._run_queue()

    # ------------------------------------------------------------------ per-rank checkpoints
    def save_shard(self, directory):
        """Write this rank's shard and the layout it is stored in to ``directory`` (one ``.npy`` + one
        ``.json`` per rank): the checkpoint of a state that does not fit one host.  Collective."""
        import json
        import os

        self._flush()
        os.makedirs(directory, exist_ok=True)
        self._sync()
        np.save(os.path.join(directory, "shard_%04d.npy" % self._rank), self._buf.cpu().numpy())
        with open(os.path.join(directory, "shard_%04d.json" % self._rank), "w") as f:
            json.dump({"version": 1, "world": self._world, "rank": self._rank, "num_modes": self._num_modes,
                       "cutoff_dim": self._trunc, "pure": bool(self._pure), "factors": self._ps,
                       "phys": [int(x) for x in self._phys], "elements": int(self._size())}, f)
        dist.barrier(group=self._pg)

    def load_shard(self, directory):
        """Restore a state written by ``save_shard`` with the same world size, cutoff and mode count."""
        import json
        import os

        with open(os.path.join(directory, "shard_%04d.json" % self._rank)) as f:
            meta = json.load(f)
        if (meta["world"], meta["num_modes"], meta["cutoff_dim"], meta["factors"]) != \
                (self._world, self._num_modes, self._trunc, self._ps):
            raise ValueError("checkpoint was written by %d ranks for %d modes at cutoff %d; this circuit differs"
                             % (meta["world"], meta["num_modes"], meta["cutoff_dim"]))
        self._flush()
        if bool(meta["pure"]) != self._pure:
            self._pure = bool(meta["pure"])
            self._bufs = None
        self._set_layout(meta["phys"])
        self._alloc()
        data = np.load(os.path.join(directory, "shard_%04d.npy" % self._rank))
        if data.size != self._size():
            raise ValueError("shard of rank %d has %d elements, expected %d" % (self._rank, data.size, self._size()))
        self._buf.copy_(torch.from_numpy(data).to(self.device))
        self._pending, self._opq, self._untouched, self._fresh = {}, [], set(), False
        dist.barrier(group=self._pg)

    # ------------------------------------------------------------------ modes (circuit.py:373-391)
    def alloc(self, n=1):
        """``add_mode``: tensor n vacuum modes onto the state (circuit.py:373-382).  The new axes are whole
        (local) axes at the innermost positions, so every rank just copies its shard into the |0..0> slice of
        a D^n (D^2n for density matrices) times larger one; no communication.  Collective."""
        self._flush()
        D = self._trunc
        add = n if self._pure else 2 * n
        old, old_size, old_axes = self._buf, self._size(), self._axes()
        self._num_modes += n
        self._phys = list(self._phys) + list(range(old_axes, old_axes + add))
        self._set_layout(self._phys)
        self._bufs = None
        self._alloc()
        L.call("b200_fill_zero", _ptr(self._buf), self._buf.numel(), self._stream())
        self._gather(old, None, self._buf, [(old_size, 1, 0, D ** add)])
        for m in range(self._num_modes - n, self._num_modes):
            self._untouched.add(m)
        self._fresh = False

    def _unsupported(self, *a, **k):
        raise NotImplementedError("this operation is not available on a sharded b200fock circuit yet")

    dealloc = _unsupported   # del_mode needs a partial trace of the sharded state
