"""Kernel-level parity through the C ABI (libb200fock.so) against the oracle, on the GPU.

Tolerance 1e-12 absolute (complex128), as stated in BASELINE.json's north_star.  Covers
every cutoff the register-blocked kernels are instantiated for, the local-memory path
above B200_MAX_FAST_CUTOFF, every axis position (first / middle / last mode), both
selection rules, ragged slice counts and the batched layouts."""
import ctypes as C

import numpy as np
import pytest

from oracle import gates as og

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def lib():
    from strawberryfields_b200 import lib as L

    L.load()
    return L


@pytest.fixture(scope="module")
def torch_mod():
    import torch

    return torch


def _dev(torch, arr):
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def _p(t):
    return C.c_void_p(t.data_ptr())


def _rand(rs, *shape):
    return rs.randn(*shape) + 1j * rs.randn(*shape)


# ---------------------------------------------------------------------------- gate tables
@pytest.mark.parametrize("D", [1, 2, 3, 5, 7, 10, 13, 16, 20])
def test_gen_gate1(lib, torch_mod, D):
    torch = torch_mod
    for kind, fn, (a, b) in ((lib.GATE_DISPLACEMENT, og.displacement, (0.7, 1.3)),
                             (lib.GATE_SQUEEZE, og.squeezing, (0.45, -0.8))):
        out = torch.empty(D * D, dtype=torch.complex128, device="cuda")
        lib.call("b200_gen_gate1", kind, D, 1, a, b, None, _p(out), None)
        assert np.abs(out.cpu().numpy().reshape(D, D) - fn(a, b, D)).max() < TOL


def test_gen_gate1_batched(lib, torch_mod):
    torch = torch_mod
    D, B = 9, 5
    rs = np.random.RandomState(1)
    params = np.stack([rs.uniform(0, 0.6, B), rs.uniform(-3, 3, B)])
    pd = _dev(torch, params)
    out = torch.empty(B * D * D, dtype=torch.complex128, device="cuda")
    lib.call("b200_gen_gate1", lib.GATE_SQUEEZE, D, B, 0.0, 0.0, _p(pd), _p(out), None)
    got = out.cpu().numpy().reshape(B, D, D)
    for b in range(B):
        assert np.abs(got[b] - og.squeezing(params[0, b], params[1, b], D)).max() < TOL


@pytest.mark.parametrize("D", [1, 2, 3, 4, 7, 10, 12, 16])
def test_gen_gate2_and_unpack(lib, torch_mod, D):
    torch = torch_mod
    P = lib.packed_size(D)
    cases = [
        (lib.GATE_BEAMSPLITTER, lib.RULE_SUM, og.beamsplitter, (0.6, 0.9)),
        (lib.GATE_BEAMSPLITTER, lib.RULE_SUM, og.beamsplitter, (-1.2, 0.0)),
        (lib.GATE_MZ, lib.RULE_SUM, og.mzgate, (0.3, 1.2)),
        (lib.GATE_S2, lib.RULE_DIFF, og.two_mode_squeeze, (0.35, 0.4)),
    ]
    for kind, rule, fn, (a, b) in cases:
        packed = torch.empty(P, dtype=torch.complex128, device="cuda")
        dense = torch.empty(D ** 4, dtype=torch.complex128, device="cuda")
        lib.call("b200_gen_gate2", kind, D, 1, a, b, None, _p(packed), None)
        lib.call("b200_unpack_gate2", rule, D, _p(packed), _p(dense), None)
        assert np.abs(dense.cpu().numpy().reshape((D,) * 4) - fn(a, b, D)).max() < TOL


@pytest.mark.parametrize("T", [0.0, 0.37, 0.9, 1.0])
def test_gen_loss_superoperator(lib, torch_mod, T):
    torch = torch_mod
    D = 8
    P = lib.packed_size(D)
    packed = torch.empty(P, dtype=torch.complex128, device="cuda")
    dense = torch.empty(D ** 4, dtype=torch.complex128, device="cuda")
    lib.call("b200_gen_gate2", lib.CHANNEL_LOSS, D, 1, T, 0.0, None, _p(packed), None)
    lib.call("b200_unpack_gate2", lib.RULE_DIFF, D, _p(packed), _p(dense), None)
    S = np.zeros((D,) * 4, dtype=complex)
    for E in og.loss_kraus(T, D):  # fockbackend/ops.py:471-490
        S += np.einsum("ab,cd->abcd", E, E.conj())
    assert np.abs(dense.cpu().numpy().reshape((D,) * 4) - S).max() < TOL


def test_gen_diag(lib, torch_mod):
    torch = torch_mod
    D = 11
    n = np.arange(D)
    for kind, p, want in ((lib.DIAG_ROTATION, 0.77, np.exp(1j * 0.77 * n)),
                          (lib.DIAG_KERR, -0.31, np.exp(-1j * 0.31 * n ** 2)),
                          (lib.DIAG_CROSS_KERR, 0.2, np.exp(1j * 0.2 * np.multiply.outer(n, n)).reshape(-1))):
        out = torch.empty(want.size, dtype=torch.complex128, device="cuda")
        lib.call("b200_gen_diag", kind, D, 1, p, None, _p(out), None)
        assert np.abs(out.cpu().numpy() - want).max() < TOL


# ---------------------------------------------------------------------------- gate application
@pytest.mark.parametrize("D", [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 21])
def test_apply_gate1_every_axis(lib, torch_mod, D):
    torch = torch_mod
    rs = np.random.RandomState(D)
    n = 4 if D <= 8 else 3
    psi = _rand(rs, *([D] * n))
    U = _rand(rs, D, D)
    for axis in range(n):
        for conj in (0, 1):
            st = _dev(torch, psi.reshape(-1))
            lib.call("b200_apply_gate1", _p(st), D ** axis, D, D ** (n - 1 - axis), _p(_dev(torch, U)), conj,
                     1, D ** n, 0, None)
            want = np.moveaxis(np.tensordot(U.conj() if conj else U, psi, axes=(1, axis)), 0, axis)
            assert np.abs(st.cpu().numpy().reshape(psi.shape) - want).max() < 1e-11 * max(1, D)


def test_apply_gate1_ragged_and_batched(lib, torch_mod):
    torch = torch_mod
    rs = np.random.RandomState(5)
    D, B = 7, 3
    outer, inner = 13, 5  # neither a multiple of the warp size
    psi = _rand(rs, B, outer, D, inner)
    U = _rand(rs, B, D, D)
    st = _dev(torch, psi.reshape(-1))
    lib.call("b200_apply_gate1", _p(st), outer, D, inner, _p(_dev(torch, U)), 0, B, outer * D * inner, D * D, None)
    want = np.einsum("bxy,boyi->boxi", U, psi)
    assert np.abs(st.cpu().numpy().reshape(psi.shape) - want).max() < TOL * 10


@pytest.mark.parametrize("D", [2, 3, 5, 7, 10, 12, 16, 18])
@pytest.mark.parametrize("rule", ["sum", "diff"])
def test_apply_gate2_every_ordered_pair(lib, torch_mod, D, rule):
    torch = torch_mod
    rs = np.random.RandomState(100 + D)
    n = 4 if D <= 7 else 3
    psi = _rand(rs, *([D] * n))
    if rule == "sum":
        T, kind, rl = og.beamsplitter(0.7, 0.4, D), lib.GATE_BEAMSPLITTER, lib.RULE_SUM
        args = (0.7, 0.4)
    else:
        T, kind, rl = og.two_mode_squeeze(0.3, 1.1, D), lib.GATE_S2, lib.RULE_DIFF
        args = (0.3, 1.1)
    packed = torch.empty(lib.packed_size(D), dtype=torch.complex128, device="cuda")
    lib.call("b200_gen_gate2", kind, D, 1, args[0], args[1], None, _p(packed), None)
    for a in range(n):
        for b in range(n):
            if a == b:
                continue
            for conj in (0, 1):
                st = _dev(torch, psi.reshape(-1))
                lib.call("b200_apply_gate2", _p(st), D ** n, D, D ** (n - 1 - a), D ** (n - 1 - b), rl,
                         _p(packed), conj, 1, D ** n, 0, None)
                Tc = T.conj() if conj else T
                want = np.moveaxis(np.tensordot(Tc, psi, axes=([1, 3], [a, b])), [0, 1], [a, b])
                assert np.abs(st.cpu().numpy().reshape(psi.shape) - want).max() < 1e-11


def test_apply_gate2_batched_gate_tables(lib, torch_mod):
    torch = torch_mod
    rs = np.random.RandomState(9)
    D, B, n = 6, 4, 3
    psi = _rand(rs, B, *([D] * n))
    params = np.stack([rs.uniform(0, 1.5, B), rs.uniform(0, 6, B)])
    P = lib.packed_size(D)
    packed = torch.empty(B * P, dtype=torch.complex128, device="cuda")
    lib.call("b200_gen_gate2", lib.GATE_BEAMSPLITTER, D, B, 0.0, 0.0, _p(_dev(torch, params)), _p(packed), None)
    st = _dev(torch, psi.reshape(-1))
    lib.call("b200_apply_gate2", _p(st), D ** n, D, D ** 0, D ** 2, lib.RULE_SUM, _p(packed), 0, B, D ** n, P, None)
    got = st.cpu().numpy().reshape(psi.shape)
    for b in range(B):
        T = og.beamsplitter(params[0, b], params[1, b], D)
        want = np.moveaxis(np.tensordot(T, psi[b], axes=([1, 3], [2, 0])), [0, 1], [2, 0])
        assert np.abs(got[b] - want).max() < TOL * 10


def test_apply_diag_and_multi(lib, torch_mod):
    torch = torch_mod
    rs = np.random.RandomState(3)
    D, n = 6, 4
    psi = _rand(rs, *([D] * n))
    t1 = np.exp(1j * rs.uniform(0, 6, D))
    t2 = np.exp(1j * rs.uniform(0, 6, (D, D)))
    st = _dev(torch, psi.reshape(-1))
    lib.call("b200_apply_diag", _p(st), D ** n, D, D ** 2, 0, _p(_dev(torch, t1)), 0, 1, D ** n, 0, None)
    assert np.abs(st.cpu().numpy().reshape(psi.shape) - psi * t1[None, :, None, None]).max() < TOL
    st = _dev(torch, psi.reshape(-1))
    lib.call("b200_apply_diag", _p(st), D ** n, D, D ** 0, D ** 3, _p(_dev(torch, t2.reshape(-1))), 1, 1, D ** n, 0, None)
    want = psi * t2.conj().T[:, None, None, :]
    assert np.abs(st.cpu().numpy().reshape(psi.shape) - want).max() < TOL
    tabs = np.exp(1j * rs.uniform(0, 6, (3, D)))
    st = _dev(torch, psi.reshape(-1))
    strides = (C.c_int64 * 3)(D ** 3, D ** 1, D ** 0)
    conjs = (C.c_int * 3)(0, 1, 0)
    lib.call("b200_apply_diag_multi", _p(st), D ** n, D, 3, strides, conjs, _p(_dev(torch, tabs.reshape(-1))), 1,
             D ** n, 0, None)
    want = psi * tabs[0][:, None, None, None] * tabs[1].conj()[None, None, :, None] * tabs[2][None, None, None, :]
    assert np.abs(st.cpu().numpy().reshape(psi.shape) - want).max() < TOL


# ---------------------------------------------------------------------------- reductions
def test_norm_abs2_scale(lib, torch_mod):
    torch = torch_mod
    rs = np.random.RandomState(4)
    n = 1_234_567
    psi = _rand(rs, n)
    st = _dev(torch, psi)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    part = torch.zeros(4096, dtype=torch.float64, device="cuda")
    lib.call("b200_norm2", _p(st), n, _p(out), _p(part), None)
    assert abs(out.item() - np.vdot(psi, psi).real) < 1e-9 * n ** 0.5
    pr = torch.empty(n, dtype=torch.float64, device="cuda")
    lib.call("b200_abs2", _p(st), _p(pr), n, None)
    assert np.abs(pr.cpu().numpy() - np.abs(psi) ** 2).max() < TOL
    lib.call("b200_scale", _p(st), n, 1.0, 0.0, _p(out), 1, None)
    assert np.abs(st.cpu().numpy() - psi / np.linalg.norm(psi)).max() < TOL


def test_gather_reduce_split_reduction(lib, torch_mod):
    """single-mode marginal of a large pure state: few outputs, long reduction (split path)"""
    torch = torch_mod
    rs = np.random.RandomState(8)
    D, n = 10, 5
    psi = _rand(rs, *([D] * n))
    st = _dev(torch, psi.reshape(-1))
    d = lib.GatherDesc()
    d.n_out_axes, d.n_red_axes = 1, 2
    d.out_ext[0], d.out_sa[0], d.out_sb[0], d.out_sc[0] = D, D ** 2, D ** 2, 1
    d.red_ext[0], d.red_ta[0], d.red_tb[0] = D ** 2, D ** 3, D ** 3
    d.red_ext[1], d.red_ta[1], d.red_tb[1] = D ** 2, 1, 1
    out = torch.zeros(D, dtype=torch.float64, device="cuda")
    part = torch.zeros(D * 64, dtype=torch.complex128, device="cuda")
    lib.call("b200_gather_reduce", C.byref(d), _p(st), _p(st), _p(out), lib.FLAG_CONJ_B | lib.FLAG_REAL_OUT,
             _p(part), None)
    want = (np.abs(psi) ** 2).sum(axis=(0, 1, 3, 4))
    assert np.abs(out.cpu().numpy() - want).max() < 1e-9


def test_bad_arguments_fail_loudly(lib, torch_mod):
    torch = torch_mod
    st = torch.zeros(16, dtype=torch.complex128, device="cuda")
    with pytest.raises(lib.B200Error):
        lib.call("b200_apply_gate1", _p(st), 1, 200, 1, _p(st), 0, 1, 16, 0, None)
    with pytest.raises(lib.B200Error):
        lib.call("b200_apply_gate2", _p(st), 16, 2, 3, 1, lib.RULE_SUM, _p(st), 0, 1, 16, 0, None)


# ---------------------------------------------------------------------------- tile pass
@pytest.mark.parametrize("D", [2, 3, 4, 5, 7, 8, 10, 11, 12, 13, 16])
def test_tile_pass_matches_numpy_double(lib, torch_mod, D):
    """b200_apply_tile_pass (cp.async-staged tiles, op list, output permutation) against the
    numpy double of the ABI, for every cutoff the kernel is instantiated for."""
    from fake_lib import FakeLib, pack

    torch = torch_mod
    rs = np.random.RandomState(200 + D)
    n = 5 if D <= 5 else 4
    total = D ** n
    psi = _rand(rs, total)
    P = lib.packed_size(D)
    U = _rand(rs, D, D)
    d = np.exp(1j * rs.uniform(0, 6, D))
    bs = pack(og.beamsplitter(0.7, 0.3, D), lib.RULE_SUM, D)
    s2 = pack(og.two_mode_squeeze(0.3, 0.9, D), lib.RULE_DIFF, D)
    coef = np.concatenate([U.reshape(-1), d, bs, s2, U.reshape(-1)])
    offs = [0, D * D, D * D + D, D * D + D + P, D * D + D + 2 * P]
    fake = FakeLib()
    cases = [(0, 1), (0, n - 2), (n - 3, n - 2)]
    perms = [(0, 1, 2), (2, 0, 1), (1, 2, 0), (0, 2, 1)]
    for ci, (p0, p1) in enumerate(cases):
        ops = (lib.TileOp * 6)()
        spec = [(lib.RULE_SINGLE, 0, 0, 0, offs[0]), (lib.TILE_DIAG, 2, 0, 1, offs[1]),
                (lib.RULE_SUM, 1, 2, 0, offs[2]), (lib.RULE_DIFF, 2, 0, 1, offs[3]),
                (lib.RULE_SINGLE, 2, 0, 1, offs[4]), (lib.RULE_SUM, 0, 1, 0, offs[2])]
        for i, (k, a1, a2, cj, off) in enumerate(spec):
            ops[i].kind, ops[i].axis1, ops[i].axis2, ops[i].conj, ops[i].coef_offset = k, a1, a2, cj, off
        # above cutoff 13 two packed tables no longer fit next to the tiles in 227 KB of shared memory
        nops, ncoef = (6, coef.size) if D <= 13 else (3, D * D + D + P)
        perm = (C.c_int * 3)(*perms[ci % len(perms)])
        s0, s1 = D ** (n - 1 - p0), D ** (n - 1 - p1)
        st = _dev(torch, psi)
        cf = _dev(torch, coef)
        lib.call("b200_apply_tile_pass", _p(st), total, D, s0, s1, ops, nops, perm, _p(cf), ncoef, 1, total, 0,
                 None)
        want = psi.copy()
        cfh = coef.copy()
        rc = fake.b200_apply_tile_pass(C.c_void_p(want.ctypes.data), total, D, s0, s1, ops, nops, perm,
                                       C.c_void_p(cfh.ctypes.data), ncoef, 1, total, 0, None)
        assert rc == 0
        assert np.abs(st.cpu().numpy() - want).max() < 1e-11


def test_tile_pass_batched_and_ragged(lib, torch_mod):
    from fake_lib import FakeLib, pack

    torch = torch_mod
    rs = np.random.RandomState(77)
    D, n, B = 10, 4, 3  # 10 tiles per batch entry: not a multiple of the 3 tiles a CTA stages
    total = D ** n
    psi = _rand(rs, B * total)
    P = lib.packed_size(D)
    coef = np.concatenate([np.concatenate([pack(og.beamsplitter(0.2 + 0.3 * b, 0.1 * b, D), lib.RULE_SUM, D),
                                           _rand(rs, D * D)]) for b in range(B)])
    per = P + D * D
    ops = (lib.TileOp * 2)()
    ops[0].kind, ops[0].axis1, ops[0].axis2, ops[0].conj, ops[0].coef_offset = lib.RULE_SUM, 0, 2, 0, 0
    ops[1].kind, ops[1].axis1, ops[1].axis2, ops[1].conj, ops[1].coef_offset = lib.RULE_SINGLE, 1, 0, 0, P
    perm = (C.c_int * 3)(1, 0, 2)
    st, cf = _dev(torch, psi), _dev(torch, coef)
    lib.call("b200_apply_tile_pass", _p(st), total, D, D ** 3, D ** 1, ops, 2, perm, _p(cf), per, B, total, per, None)
    want, cfh = psi.copy(), coef.copy()
    FakeLib().b200_apply_tile_pass(C.c_void_p(want.ctypes.data), total, D, D ** 3, D ** 1, ops, 2, perm,
                                   C.c_void_p(cfh.ctypes.data), per, B, total, per, None)
    assert np.abs(st.cpu().numpy() - want).max() < 1e-11


# ---------------------------------------------------------------------------- innermost-axis geometries
# The TMA-staged persistent kernel (csrc/inner.cu) has one tile shape per geometry: adjacent axes (mid = 1),
# a short mid (< 32 slices per row run: whole outer blocks), rows (mid >= 32), one-mode gates on the last
# axis; partial tiles at the end of a state and of a row run; per-entry gate tables on a batch.
@pytest.mark.parametrize("D,n", [(10, 5), (6, 6), (4, 7), (3, 7), (7, 4), (12, 4), (16, 3)])
def test_inner_axis_pair_gates_all_partners(lib, torch_mod, D, n):
    torch = torch_mod
    rs = np.random.RandomState(17 * D + n)
    psi = _rand(rs, *([D] * n))
    for rule in ("sum", "diff"):
        if rule == "sum":
            T, kind, rl, args = og.beamsplitter(0.6, 0.3, D), lib.GATE_BEAMSPLITTER, lib.RULE_SUM, (0.6, 0.3)
        else:
            T, kind, rl, args = og.two_mode_squeeze(0.25, 0.9, D), lib.GATE_S2, lib.RULE_DIFF, (0.25, 0.9)
        packed = torch.empty(lib.packed_size(D), dtype=torch.complex128, device="cuda")
        lib.call("b200_gen_gate2", kind, D, 1, args[0], args[1], None, _p(packed), None)
        for a in range(n - 1):
            for (x, y) in ((a, n - 1), (n - 1, a)):
                st = _dev(torch, psi.reshape(-1))
                lib.call("b200_apply_gate2", _p(st), D ** n, D, D ** (n - 1 - x), D ** (n - 1 - y), rl,
                         _p(packed), 0, 1, D ** n, 0, None)
                want = np.moveaxis(np.tensordot(T, psi, axes=([1, 3], [x, y])), [0, 1], [x, y])
                assert np.abs(st.cpu().numpy().reshape(psi.shape) - want).max() < 1e-11, (rule, x, y)


@pytest.mark.parametrize("D,slices", [(10, 1000), (10, 333), (6, 4099), (5, 77), (16, 40), (2, 100001)])
def test_inner_axis_one_mode_gate_ragged(lib, torch_mod, D, slices):
    """one-mode gate on the innermost axis; slice counts that are not multiples of a tile"""
    torch = torch_mod
    rs = np.random.RandomState(D + slices)
    psi = _rand(rs, slices, D)
    U = _rand(rs, D, D)
    st = _dev(torch, psi.reshape(-1))
    lib.call("b200_apply_gate1", _p(st), slices, D, 1, _p(_dev(torch, U)), 1, 1, slices * D, 0, None)
    want = psi @ U.conj().T
    assert np.abs(st.cpu().numpy().reshape(psi.shape) - want).max() < 1e-11 * D


def test_inner_axis_batched_tables(lib, torch_mod):
    """per-entry gate tables on a batch: the persistent CTAs reload the table when their tile range crosses
    into the next batch entry"""
    torch = torch_mod
    rs = np.random.RandomState(23)
    D, B, n = 10, 5, 4
    psi = _rand(rs, B, *([D] * n))
    params = np.stack([rs.uniform(0, 1.5, B), rs.uniform(0, 6, B)])
    P = lib.packed_size(D)
    packed = torch.empty(B * P, dtype=torch.complex128, device="cuda")
    lib.call("b200_gen_gate2", lib.GATE_BEAMSPLITTER, D, B, 0.0, 0.0, _p(_dev(torch, params)), _p(packed), None)
    for (x, y) in ((2, 3), (3, 0), (1, 3)):
        st = _dev(torch, psi.reshape(-1))
        lib.call("b200_apply_gate2", _p(st), D ** n, D, D ** (n - 1 - x), D ** (n - 1 - y), lib.RULE_SUM, _p(packed),
                 0, B, D ** n, P, None)
        got = st.cpu().numpy().reshape(psi.shape)
        for b in range(B):
            T = og.beamsplitter(params[0, b], params[1, b], D)
            want = np.moveaxis(np.tensordot(T, psi[b], axes=([1, 3], [x, y])), [0, 1], [x, y])
            assert np.abs(got[b] - want).max() < 1e-11, (x, y, b)
    U = _rand(rs, B, D, D)
    st = _dev(torch, psi.reshape(-1))
    lib.call("b200_apply_gate1", _p(st), D ** (n - 1), D, 1, _p(_dev(torch, U)), 0, B, D ** n, D * D, None)
    want = np.einsum("bxy,b...y->b...x", U, psi)
    assert np.abs(st.cpu().numpy().reshape(psi.shape) - want).max() < 1e-10


def test_diag_multi_geometries(lib, torch_mod):
    """b200_apply_diag_multi: super-row decomposition (rows of D^j elements, D rows per super-row) for axis
    subsets that include / exclude the axis of stride L, conjugated entries, small states, a batch"""
    torch = torch_mod
    rs = np.random.RandomState(31)
    for D, n, axes in ((10, 5, [0, 1, 4]), (10, 5, [1, 2]), (10, 5, [3]), (10, 4, [0, 1, 2, 3]), (7, 2, [0, 1]),
                       (6, 6, [0, 2, 5]), (3, 1, [0]), (16, 3, [0, 2])):
        psi = _rand(rs, *([D] * n))
        tabs = np.exp(1j * rs.uniform(0, 6, (len(axes), D)))
        conj = [int(rs.rand() < 0.5) for _ in axes]
        want = psi.copy()
        for t, ax, cj in zip(tabs, axes, conj):
            shape = [1] * n
            shape[ax] = D
            want = want * (t.conj() if cj else t).reshape(shape)
        st = _dev(torch, psi.reshape(-1))
        k = len(axes)
        strides = (C.c_int64 * k)(*[D ** (n - 1 - a) for a in axes])
        conjs = (C.c_int * k)(*conj)
        lib.call("b200_apply_diag_multi", _p(st), D ** n, D, k, strides, conjs, _p(_dev(torch, tabs.reshape(-1))), 1,
                 D ** n, 0, None)
        assert np.abs(st.cpu().numpy().reshape(psi.shape) - want).max() < TOL * 10, (D, n, axes)
    D, n, B = 5, 4, 3
    psi = _rand(rs, B, *([D] * n))
    tabs = np.exp(1j * rs.uniform(0, 6, (B, 2, D)))
    st = _dev(torch, psi.reshape(-1))
    lib.call("b200_apply_diag_multi", _p(st), D ** n, D, 2, (C.c_int64 * 2)(D ** 3, D), (C.c_int * 2)(0, 1),
             _p(_dev(torch, tabs.reshape(-1))), B, D ** n, 2 * D, None)
    want = psi * tabs[:, 0].reshape(B, D, 1, 1, 1) * tabs[:, 1].conj().reshape(B, 1, 1, D, 1)
    assert np.abs(st.cpu().numpy().reshape(psi.shape) - want).max() < TOL * 10


@pytest.mark.parametrize("D,n", [(2, 5), (5, 4), (10, 4), (12, 3), (16, 3), (7, 2)])
def test_gram1_every_axis(lib, torch_mod, D, n):
    """b200_gram1: reduced density matrix / marginal of one mode of a ket in one read of the state, every axis
    position, full matrix (cutoffs <= 12) and diagonal (<= 16), and a batch"""
    torch = torch_mod
    rs = np.random.RandomState(D * 10 + n)
    B = 3
    psi = _rand(rs, B, *([D] * n))
    st = _dev(torch, psi.reshape(-1))
    part = torch.empty(int(lib.load().b200_gram1_part_doubles(D, B)), dtype=torch.float64, device="cuda")
    letters = "abcdefgh"[:n]
    for axis in range(n):
        inner = D ** (n - 1 - axis)
        sub_in = "z" + letters
        sub_cj = "z" + letters.replace(letters[axis], "X")
        want = np.einsum("%s,%s->z%sX" % (sub_in, sub_cj, letters[axis]), psi, psi.conj())
        probs = torch.empty(B * D, dtype=torch.float64, device="cuda")
        lib.call("b200_gram1", _p(st), D ** axis, D, inner, 1, _p(probs), _p(part), B, D ** n, None)
        assert np.abs(probs.cpu().numpy().reshape(B, D) - np.real(np.einsum("zaa->za", want))).max() < 1e-10
        if D <= 12:
            rho = torch.empty(B * D * D, dtype=torch.complex128, device="cuda")
            lib.call("b200_gram1", _p(st), D ** axis, D, inner, 0, _p(rho), _p(part), B, D ** n, None)
            assert np.abs(rho.cpu().numpy().reshape(B, D, D) - want).max() < 1e-10
    if D > 12:
        with pytest.raises(lib.B200Error):
            rho = torch.empty(B * D * D, dtype=torch.complex128, device="cuda")
            lib.call("b200_gram1", _p(st), 1, D, D ** (n - 1), 0, _p(rho), _p(part), B, D ** n, None)


def test_outer_axis(lib, torch_mod):
    """b200_outer_axis: out[b][j][i] = f[b][j] * in[b][i] (odd and even lengths, shared and per-entry factors)"""
    torch = torch_mod
    rs = np.random.RandomState(41)
    for n_in, nf, B, shared in ((1, 10, 1, True), (7, 5, 3, False), (1000, 100, 2, True), (4097, 10, 1, True),
                                (10 ** 5, 10, 2, False)):
        x = _rand(rs, B, n_in)
        f = _rand(rs, 1 if shared else B, nf)
        out = torch.empty(B * nf * n_in, dtype=torch.complex128, device="cuda")
        xd, fd = _dev(torch, x.reshape(-1)), _dev(torch, f.reshape(-1))   # keep both alive across the launch
        lib.call("b200_outer_axis", _p(xd), _p(fd), _p(out), n_in, nf, B, n_in, 0 if shared else nf, None)
        want = np.einsum("bj,bi->bji", np.broadcast_to(f, (B, nf)), x)
        assert np.abs(out.cpu().numpy().reshape(B, nf, n_in) - want).max() < 1e-14
