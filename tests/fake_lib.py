"""TEST DOUBLE of the C ABI in include/b200fock.h, in numpy, on host memory.

Only the CPU test-suite uses it (``-m "not gpu"``): it lets the host-side logic of
``strawberryfields_b200`` (mode/axis geometry, the lazy gate queue, gather descriptors,
measurement bookkeeping) run without a GPU, so that a wrong stride is caught here and
not on the B200 box.  It is NOT a fallback: the product never imports it, and
``DeviceCircuit`` refuses to run without CUDA unless a test monkeypatches both the
library handle and ``circuit._TEST_HOST_MODE``.

Every function takes the ctypes-converted arguments ``lib.call`` passes and follows the
semantics documented in the header; gate tables come from the oracle's restated
recursions, so the GPU generators are pinned independently (tests/test_gpu_*.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from oracle import gates as og

C128 = np.complex128
RULE_SUM, RULE_DIFF = 1, 2


def _addr(p):
    if p is None:
        return None
    if isinstance(p, C.c_void_p):
        return p.value
    return int(p)


def _c(p, n):
    """complex128 view of n entries at device(host) address p"""
    a = _addr(p)
    return np.ctypeslib.as_array((C.c_double * (2 * int(n))).from_address(a)).view(C128)


def _d(p, n):
    a = _addr(p)
    return np.ctypeslib.as_array((C.c_double * int(n)).from_address(a))


def blk_size(b, D):
    return min(b, 2 * D - 2 - b) + 1


def blk_lo(b, D):
    return max(0, b - D + 1)


def blk_off(b, D):
    return sum(blk_size(x, D) ** 2 for x in range(b))


def packed_size(D):
    return blk_off(2 * D - 1, D)


def member(rule, D, b, m):
    lo = blk_lo(b, D)
    k = lo + m
    l = (b - lo - m) if rule == RULE_SUM else (lo + m - (b - (D - 1)))
    return k, l


def pack(T, rule, D):
    """dense [o1,i1,o2,i2] -> block-packed"""
    out = np.zeros(packed_size(D), dtype=C128)
    for b in range(2 * D - 1):
        c = blk_size(b, D)
        for a in range(c):
            ko, lo_ = member(rule, D, b, a)
            for j in range(c):
                ki, li = member(rule, D, b, j)
                out[blk_off(b, D) + a * c + j] = T[ko, ki, lo_, li]
    return out


def unpack(P, rule, D):
    T = np.zeros((D,) * 4, dtype=C128)
    for b in range(2 * D - 1):
        c = blk_size(b, D)
        for a in range(c):
            ko, lo_ = member(rule, D, b, a)
            for j in range(c):
                ki, li = member(rule, D, b, j)
                T[ko, ki, lo_, li] = P[blk_off(b, D) + a * c + j]
    return T


def _batch_params(nb, p0, p1, params):
    if _addr(params) is None:
        return [(p0, p1)] * nb
    arr = _d(params, 2 * nb).reshape(2, nb)
    return [(arr[0, b], arr[1, b]) for b in range(nb)]


class FakeLib:
    def __init__(self):
        self.launches = 0
        self.err = b""

    # -- library
    def b200_version(self):
        return 100

    def b200_last_error(self):
        return self.err

    def b200_packed_size(self, D):
        return packed_size(D)

    def b200_launch_count(self):
        return self.launches

    def b200_reset_launch_count(self):
        self.launches = 0

    # -- generators
    def b200_gen_gate1(self, kind, D, nb, p0, p1, params, out, stream):
        o = _c(out, nb * D * D).reshape(nb, D, D)
        for b, (a, c) in enumerate(_batch_params(nb, p0, p1, params)):
            o[b] = og.displacement(float(a), float(c), D) if kind == 1 else og.squeezing(float(a), float(c), D)
        self.launches += 1
        return 0

    def b200_gen_diag(self, kind, D, nb, p0, params, out, stream):
        per = D * D if kind == 12 else D
        o = _c(out, nb * per).reshape(nb, per)
        ps = [p0] * nb if _addr(params) is None else list(_d(params, nb))
        n = np.arange(D)
        for b, p in enumerate(ps):
            if kind == 10:
                o[b] = np.exp(1j * p * n)
            elif kind == 11:
                o[b] = np.exp(1j * p * n ** 2)
            else:
                o[b] = np.exp(1j * p * np.multiply.outer(n, n)).reshape(-1)
        self.launches += 1
        return 0

    def b200_gen_gate2(self, kind, D, nb, p0, p1, params, out, stream):
        P = packed_size(D)
        o = _c(out, nb * P).reshape(nb, P)
        for b, (a, c) in enumerate(_batch_params(nb, p0, p1, params)):
            a, c = float(a), float(c)
            if kind == 20:
                o[b] = pack(og.beamsplitter(a, c, D), RULE_SUM, D)
            elif kind == 21:
                o[b] = pack(og.mzgate(a, c, D), RULE_SUM, D)
            elif kind == 22:
                o[b] = pack(og.two_mode_squeeze(a, c, D), RULE_DIFF, D)
            elif kind == 30:
                # superoperator on (ket, bra): T[a, a', d, d'] = sum_l E_l[a, a'] conj(E_l[d, d'])
                S = np.zeros((D,) * 4, dtype=C128)
                for E in og.loss_kraus(a, D):
                    S += np.einsum("ab,cd->abcd", E, E.conj())
                o[b] = pack(S, RULE_DIFF, D)
            else:
                return -1
        self.launches += 1
        return 0

    def b200_compose_gate1(self, D, nb, A, B, Cc, stream):
        a = _c(A, nb * D * D).reshape(nb, D, D)
        b = _c(B, nb * D * D).reshape(nb, D, D)
        _c(Cc, nb * D * D).reshape(nb, D, D)[:] = a @ b
        self.launches += 1
        return 0

    def b200_fold_diag_gate1(self, D, nb, U, pre, post, stream):
        u = _c(U, nb * D * D).reshape(nb, D, D)
        if _addr(pre) is not None:
            u *= _c(pre, nb * D).reshape(nb, 1, D)
        if _addr(post) is not None:
            u *= _c(post, nb * D).reshape(nb, D, 1)
        self.launches += 1
        return 0

    def b200_fold_diag_gate2(self, rule, D, nb, G, pre1, pre2, post1, post2, stream):
        P = packed_size(D)
        g = _c(G, nb * P).reshape(nb, P)
        for b in range(nb):
            T = unpack(g[b], rule, D)  # [o1,i1,o2,i2]
            if _addr(pre1) is not None:
                T = T * _c(pre1, nb * D).reshape(nb, D)[b][None, :, None, None]
            if _addr(pre2) is not None:
                T = T * _c(pre2, nb * D).reshape(nb, D)[b][None, None, None, :]
            if _addr(post1) is not None:
                T = T * _c(post1, nb * D).reshape(nb, D)[b][:, None, None, None]
            if _addr(post2) is not None:
                T = T * _c(post2, nb * D).reshape(nb, D)[b][None, None, :, None]
            g[b] = pack(T, rule, D)
        self.launches += 1
        return 0

    def b200_mul_tables(self, n, a, b, conj_b, out, stream):
        bb = _c(b, n)
        _c(out, n)[:] = _c(a, n) * (bb.conj() if conj_b else bb)
        self.launches += 1
        return 0

    def b200_unpack_gate2(self, rule, D, packed, dense, stream):
        _c(dense, D ** 4)[:] = unpack(_c(packed, packed_size(D)), rule, D).reshape(-1)
        self.launches += 1
        return 0

    # -- application
    def b200_apply_gate1(self, state, outer, D, inner, U, conj, nb, sbs, gbs, stream):
        for b in range(nb):
            st = _c(_addr(state) + 16 * b * sbs, outer * D * inner).reshape(outer, D, inner)
            u = _c(_addr(U) + 16 * b * gbs, D * D).reshape(D, D)
            if conj:
                u = u.conj()
            st[:] = np.einsum("ab,obi->oai", u, st)
        self.launches += 1
        return 0

    def b200_apply_gate2(self, state, total, D, s1, s2, rule, packed, conj, nb, sbs, gbs, stream):
        hi, lo = max(s1, s2), min(s1, s2)
        if hi % (D * lo) or total % (D * hi):
            return -1
        mid = hi // (D * lo)
        outer = total // (D * hi)
        for b in range(nb):
            st = _c(_addr(state) + 16 * b * sbs, total).reshape(outer, D, mid, D, lo)
            T = unpack(_c(_addr(packed) + 16 * b * gbs, packed_size(D)), rule, D)
            if conj:
                T = T.conj()
            if s1 > s2:  # axis 1 = first gate index
                st[:] = np.einsum("akbl,okmli->oambi", T, st)
            else:
                st[:] = np.einsum("akbl,olmki->obmai", T, st)
        self.launches += 1
        return 0

    def b200_apply_diag(self, state, total, D, s1, s2, tab, conj, nb, sbs, tbs, stream):
        e = np.arange(total)
        d1 = (e // s1) % D
        idx = d1 if not s2 else d1 * D + (e // s2) % D
        ntab = D * D if s2 else D
        for b in range(nb):
            t = _c(_addr(tab) + 16 * b * tbs, ntab)
            if conj:
                t = t.conj()
            _c(_addr(state) + 16 * b * sbs, total)[:] *= t[idx]
        self.launches += 1
        return 0

    def b200_apply_diag_multi(self, state, total, D, naxes, strides, conjs, tabs, nb, sbs, tbs, stream):
        e = np.arange(total)
        for b in range(nb):
            t = _c(_addr(tabs) + 16 * b * tbs, naxes * D).reshape(naxes, D)
            f = np.ones(total, dtype=C128)
            for k in range(naxes):
                row = t[k].conj() if conjs[k] else t[k]
                f *= row[(e // strides[k]) % D]
            _c(_addr(state) + 16 * b * sbs, total)[:] *= f
        self.launches += 1
        return 0

    # -- tile pass
    def b200_tile_groups(self, D):
        return 1 if D >= 14 else max(1, 32 // D)

    def b200_tile_smem_bytes(self, D, coef_count):
        s1 = D | 1
        s0 = (D * s1) | 1
        return (2 * self.b200_tile_groups(D) * ((D * s0) | 1) + coef_count) * 16

    def b200_apply_tile_pass(self, state, total, D, stride0, stride1, ops, nops, out_perm, coef, coef_count,
                             nb, sbs, cbs, stream):
        if D < 2 or D > 16 or stride1 % D or stride0 % (D * stride1) or total % (D * stride0):
            return -1
        if self.b200_tile_smem_bytes(D, coef_count) > 227 * 1024:
            return -2
        lo, mid, hi = stride1 // D, stride0 // (D * stride1), total // (D * stride0)
        pos = [1, 3, 5]
        for b in range(nb):
            st = _c(_addr(state) + 16 * b * sbs, total)
            cur = st.reshape(hi, D, mid, D, lo, D).copy()
            cf = _c(_addr(coef) + 16 * b * cbs, coef_count) if coef_count else None
            for o in range(nops):
                op = ops[o]
                off = op.coef_offset
                if op.kind == 3:
                    t = cf[off:off + D]
                    t = t.conj() if op.conj else t
                    shape = [1] * 6
                    shape[pos[op.axis1]] = D
                    cur = cur * t.reshape(shape)
                elif op.kind == 0:
                    U = cf[off:off + D * D].reshape(D, D)
                    U = U.conj() if op.conj else U
                    cur = np.moveaxis(np.tensordot(U, cur, axes=(1, pos[op.axis1])), 0, pos[op.axis1])
                else:
                    T = unpack(cf[off:off + packed_size(D)], op.kind, D)
                    T = T.conj() if op.conj else T
                    a1, a2 = pos[op.axis1], pos[op.axis2]
                    cur = np.moveaxis(np.tensordot(T, cur, axes=([1, 3], [a1, a2])), [0, 1], [a1, a2])
            order = list(range(6))
            for k in range(3):
                order[pos[out_perm[k]]] = pos[k]
            st[:] = np.ascontiguousarray(cur.transpose(order)).reshape(-1)
        self.launches += 1
        return 0

    # -- gather
    def b200_gather_reduce(self, desc, A, B, Cc, flags, part, stream):
        d = desc._obj
        no, nr = d.n_out_axes, d.n_red_axes
        oext = [d.out_ext[j] for j in range(no)]
        rext = [d.red_ext[j] for j in range(nr)]

        def offsets(exts, strides):
            off = np.zeros(1, dtype=np.int64)
            for e, s in zip(exts, strides):
                off = (off[:, None] + (np.arange(e, dtype=np.int64) * s)[None, :]).reshape(-1)
            return off

        oa = offsets(oext, [d.out_sa[j] for j in range(no)]) + d.base_a
        ob = offsets(oext, [d.out_sb[j] for j in range(no)]) + d.base_b
        oc = offsets(oext, [d.out_sc[j] for j in range(no)]) + d.base_c
        ta = offsets(rext, [d.red_ta[j] for j in range(nr)])
        tb = offsets(rext, [d.red_tb[j] for j in range(nr)])
        ia = oa[:, None] + ta[None, :]
        ib = ob[:, None] + tb[None, :]
        if flags & 4:
            a = _d(A, int(ia.max()) + 1)[ia].astype(C128)
        else:
            a = _c(A, int(ia.max()) + 1)[ia]
        if _addr(B) is not None:
            bb = _c(B, int(ib.max()) + 1)[ib]
            a = a * (bb.conj() if flags & 1 else bb)
        res = a.sum(axis=1)
        if flags & 2:
            _d(Cc, int(oc.max()) + 1)[oc] = res.real
        else:
            _c(Cc, int(oc.max()) + 1)[oc] = res
        self.launches += 1
        return 0

    # -- elementwise
    def b200_fill_zero(self, p, n, stream):
        if n:
            _c(p, n)[:] = 0
        self.launches += 1
        return 0

    def b200_set_element(self, p, idx, re, im, stream):
        _c(p, idx + 1)[idx] = re + 1j * im
        self.launches += 1
        return 0

    def b200_abs2(self, psi, out, n, stream):
        v = _c(psi, n)
        _d(out, n)[:] = v.real ** 2 + v.imag ** 2
        self.launches += 1
        return 0

    def b200_norm2(self, psi, n, out, part, stream):
        v = _c(psi, n)
        _d(out, 1)[0] = np.vdot(v, v).real
        self.launches += 2
        return 0

    def b200_scale(self, p, n, re, im, divisor, sqrt_div, stream):
        f = re + 1j * im
        if _addr(divisor) is not None:
            dv = _d(divisor, 1)[0]
            f = f / (np.sqrt(dv) if sqrt_div else dv)
        _c(p, n)[:] *= f
        self.launches += 1
        return 0

    # -- multi-GPU exchange (the "peers" of the CPU double are POSIX shared-memory tensors)
    def b200_exchange_copy(self, desc, local_src, bulk_ctas, stream):
        d = desc._obj
        exts = [d.ext[j] for j in range(d.n_axes)] + [int(d.run)]
        ss = [d.ss[j] for j in range(d.n_axes)] + [1]
        ds = [d.ds[j] for j in range(d.n_axes)] + [1]
        so = np.zeros(1, dtype=np.int64)
        do = np.zeros(1, dtype=np.int64)
        for e, a, b in zip(exts, ss, ds):
            so = (so[:, None] + (np.arange(e, dtype=np.int64) * a)[None, :]).reshape(-1)
            do = (do[:, None] + (np.arange(e, dtype=np.int64) * b)[None, :]).reshape(-1)
        assert len(np.unique(do)) == len(do), "exchange destinations overlap"
        for k in range(d.n_src):
            s = (k + d.first_src) % d.n_src
            si, di = so + d.src_base[s], do + d.dst_base[s]
            _c(d.dst[s], int(di.max()) + 1)[di] = _c(d.src[s], int(si.max()) + 1)[si]
        self.launches += 1
        return 0

    def b200_peer_barrier(self, pf, epoch, timeout_s, stream):
        raise AssertionError("the CPU double synchronises with the process group, not with device flags")

    # -- single-mode Gram matrix / marginal
    def b200_gram1_part_doubles(self, D, nbatch):
        return nbatch * 444 * D * (D + 1)

    def b200_gram1(self, psi, outer, D, inner, diag_only, out, part, nb, sbs, stream):
        assert D <= (16 if diag_only else 12)
        for b in range(nb):
            v = _c(_addr(psi) + 16 * b * sbs, outer * D * inner).reshape(outer, D, inner)
            rho = np.einsum("oai,obi->ab", v, v.conj())
            if diag_only:
                _d(_addr(out) + 8 * b * D, D)[:] = np.real(np.diag(rho))
            else:
                _c(_addr(out) + 16 * b * D * D, D * D)[:] = rho.reshape(-1)
        self.launches += 2
        return 0

    def b200_outer_axis(self, inp, f, out, n_in, nf, nb, in_bs, f_bs, stream):
        for b in range(nb):
            v = _c(_addr(inp) + 16 * b * in_bs, n_in)
            fac = _c(_addr(f) + 16 * b * f_bs, nf)
            _c(_addr(out) + 16 * b * n_in * nf, n_in * nf)[:] = np.multiply.outer(fac, v).reshape(-1)
        self.launches += 1
        return 0
