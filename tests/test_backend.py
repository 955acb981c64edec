"""B200FockBackend against the reference-run golden fixtures and the oracle.

Every test runs twice:
  * ``gpu``  -- the real thing: CUDA kernels through libb200fock.so (marked ``gpu``);
  * ``host`` -- the same host-side logic (mode/axis geometry, lazy gate queue, gather
    descriptors, measurement bookkeeping) against the numpy double of the C ABI
    (tests/fake_lib.py), so a wrong stride is caught without a GPU.
Tolerance: 1e-12 absolute on complex128 amplitudes / probabilities (BASELINE north_star);
measurement outcomes must be identical."""
import json
import os

import numpy as np
import pytest

import scripts
from fake_lib import FakeLib

TOL = 1e-12


@pytest.fixture(params=["host", pytest.param("gpu", marks=pytest.mark.gpu)])
def host_backend(request, monkeypatch):
    from strawberryfields_b200 import circuit, lib
    from strawberryfields_b200.backend import B200FockBackend

    if request.param == "host":
        monkeypatch.setattr(lib, "_lib", FakeLib())
        monkeypatch.setattr(circuit, "_TEST_HOST_MODE", True)

    def make(**opts):
        be = B200FockBackend()
        be._test_opts = opts
        orig = be.begin_circuit

        def begin(n, **kw):
            kw.setdefault("lazy_vacuum", False)  # the eager path; tests/test_vacuum_lazy.py runs the default
            kw.update(opts)
            return orig(n, **kw)

        be.begin_circuit = begin
        return be

    return make


@pytest.mark.parametrize("fuse", [True, "tile", False])
@pytest.mark.parametrize("script", scripts.all_scripts(), ids=lambda s: s[0])
def test_script_matches_reference_fixture(script, fuse, host_backend, golden_dir, request):
    if script[0] == "boson_sampling_d7" and request.node.callspec.params["host_backend"] == "host":
        pytest.skip("large for the numpy double; run on the GPU")
    if script[0].startswith("homodyne_gkp") and request.node.callspec.params["host_backend"] == "gpu":
        # written after round 1's GPU budget ended: their first GPU run is in tests/test_vacuum_lazy.py, the
        # last file of the suite, so that a surprise there cannot hide the results of the validated tests
        pytest.skip("GPU variant runs in tests/test_vacuum_lazy.py::test_homodyne_scripts_eager")
    ref = np.load(os.path.join(golden_dir, f"ref_{script[0]}.npz"))
    rets, st = scripts.run_script(host_backend(strict_purity=True, fuse=fuse), script)
    assert bool(ref["pure"]) == st.is_pure
    if "data" in ref:
        assert st.data.shape == ref["data"].shape
        assert np.abs(st.data - ref["data"]).max() < TOL
    else:
        assert np.abs(st.all_fock_probs() - ref["probs"]).max() < TOL
    for i, r in enumerate(rets):
        assert np.array_equal(r, ref[f"ret{i}"])


def test_pure_fast_path_gives_same_probabilities(host_backend, golden_dir):
    """Default (non-strict) mode keeps the state pure when untouched vacuum modes are
    prepared; probabilities equal the reference's mixed-state result."""
    script = scripts.boson_sampling(5)
    ref = np.load(os.path.join(golden_dir, f"ref_{script[0]}.npz"))
    _, st = scripts.run_script(host_backend(), script)
    assert st.is_pure
    assert np.abs(st.all_fock_probs() - ref["probs"]).max() < TOL


@pytest.mark.parametrize("N,D", [(4, 6), (5, 5)])
def test_compiled_interferometer(N, D, host_backend, golden_dir):
    with open(os.path.join(golden_dir, f"interferometer_n{N}.json")) as f:
        gl = json.load(f)["gates"]
    be = host_backend()
    be.begin_circuit(N, cutoff_dim=D)
    for g in gl:
        getattr(be, g[0])(*g[1:])
    ref = np.load(os.path.join(golden_dir, f"ref_interferometer_n{N}_d{D}.npz"))["data"]
    st = be.state()
    assert np.abs(st.ket() - ref).max() < TOL
    # the lazy queue must have saved passes: 2N dense + all R gates folded away
    assert np.abs(st.trace() - np.vdot(ref, ref).real) < TOL


def test_state_subsets_and_reductions(host_backend):
    from oracle.fock_oracle import OracleBackend

    sc = scripts.every_gate(3, 4, True, seed=3, prep=False)
    _, st = scripts.run_script(host_backend(), sc)
    ob = OracleBackend()
    _, ost = scripts.run_script(ob, sc)
    assert np.abs(st.all_fock_probs() - ost.all_fock_probs()).max() < TOL
    for modes in ([0], [1], [2], [0, 2], [1, 2]):
        assert np.abs(st.reduced_dm(modes) - ost.reduced_dm(modes)).max() < TOL
    assert abs(st.fock_prob([1, 0, 2]) - ost.fock_prob([1, 0, 2])) < TOL
    assert abs(st.trace() - ost.trace()) < TOL
    m, v = st.mean_photon(1)
    om, ov = ost.mean_photon(1)
    assert abs(m - om) < 1e-10 and abs(v - ov) < 1e-10


def test_backend_state_with_mode_order(host_backend):
    from oracle.fock_oracle import OracleBackend

    sc = scripts.every_gate(3, 4, False, seed=4)
    be = host_backend()
    scripts.run_script(be, sc)
    ob = OracleBackend()
    scripts.run_script(ob, sc)
    for modes in ([2, 0], [1], [2, 1, 0], [0, 1]):
        a, b = be.state(modes), ob.state(modes)
        assert a.num_modes == len(modes)
        assert np.abs(a.dm() - b.dm()).max() < TOL


def test_batched_gates(host_backend):
    """Batch = leading axis, per-element parameters (tfbackend semantics):
    equals B independent runs of the oracle."""
    from oracle.fock_oracle import OracleBackend

    B, n, D = 3, 3, 5
    rs = np.random.RandomState(0)
    r = rs.uniform(0.1, 0.3, B)
    ph = rs.uniform(0, 6, B)
    th = rs.uniform(0, 1.5, B)
    be = host_backend(batch_size=B)
    be.begin_circuit(n, cutoff_dim=D)
    be.squeeze(r, ph, 0)
    be.displacement(0.2, ph, 1)
    be.rotation(th, 1)
    be.beamsplitter(th, 0.3, 0, 1)
    be.kerr_interaction(0.1, 2)
    be.mzgate(ph, th, 1, 2)
    be.two_mode_squeeze(r, 0.2, 2, 0)
    st = be.state()
    kets = st.ket()
    assert kets.shape == (B,) + (D,) * n
    for b in range(B):
        ob = OracleBackend()
        ob.begin_circuit(n, cutoff_dim=D)
        ob.squeeze(r[b], ph[b], 0)
        ob.displacement(0.2, ph[b], 1)
        ob.rotation(th[b], 1)
        ob.beamsplitter(th[b], 0.3, 0, 1)
        ob.kerr_interaction(0.1, 2)
        ob.mzgate(ph[b], th[b], 1, 2)
        ob.two_mode_squeeze(r[b], 0.2, 2, 0)
        assert np.abs(kets[b] - ob.state().data).max() < TOL
    assert np.abs(st.all_fock_probs()[1] - np.abs(kets[1]) ** 2).max() < TOL


def test_error_behaviour(host_backend):
    be = host_backend()
    with pytest.raises(ValueError, match="cutoff_dim"):
        be.begin_circuit(2)
    be.begin_circuit(2, cutoff_dim=4)
    with pytest.raises(ValueError, match="not valid"):
        be.rotation(0.1, 5)
    with pytest.raises(NotImplementedError):
        be.measure_fock([0], shots=2)
    with pytest.raises(NotImplementedError):
        be.measure_fock([0, 1], select=[None, 1])
    with pytest.raises(ValueError):
        be.measure_fock([0, 1], select=[1])
    be.prepare_fock_state(1, 0)
    with pytest.raises(ZeroDivisionError):
        be.measure_fock([0], select=[2])


def test_copy_on_write_snapshot(host_backend):
    be = host_backend()
    be.begin_circuit(2, cutoff_dim=4)
    be.displacement(0.3, 0.1, 0)
    st = be.state()
    before = st.ket().copy()
    be.squeeze(0.2, 0.0, 0)
    be.beamsplitter(0.4, 0.1, 0, 1)
    st2 = be.state()
    assert np.abs(st.all_fock_probs() - np.abs(before) ** 2).max() < TOL  # snapshot unaffected
    assert np.abs(st2.ket() - before).max() > 1e-3


def test_reset_changes_cutoff(host_backend):
    # tests/integration/test_engine_integration.py:74-91
    be = host_backend()
    be.begin_circuit(2, cutoff_dim=4)
    be.displacement(0.3, 0.1, 0)
    be.reset(cutoff_dim=6)
    assert be.get_cutoff_dim() == 6
    assert be.is_vacuum(0.0)
    assert be.state().ket().shape == (6, 6)


def test_tile_mode_fuses_gates_into_fewer_launches(monkeypatch, golden_dir):
    """With fuse=True the 35-gate compiled interferometer program runs as a handful of
    b200_apply_tile_pass launches and no per-gate pass; same ket as the reference."""
    from strawberryfields_b200 import circuit, lib
    from strawberryfields_b200.backend import B200FockBackend

    log = []

    class Spy(FakeLib):
        def __getattribute__(self, name):
            attr = FakeLib.__getattribute__(self, name)
            if name.startswith("b200_apply"):
                log.append(name)
            return attr

    monkeypatch.setattr(lib, "_lib", Spy())
    monkeypatch.setattr(circuit, "_TEST_HOST_MODE", True)
    with open(os.path.join(golden_dir, "interferometer_n5.json")) as f:
        gl = json.load(f)["gates"]
    be = B200FockBackend()
    be.begin_circuit(5, cutoff_dim=5, fuse="tile")
    for g in gl:
        getattr(be, g[0])(*g[1:])
    ket = be.state().ket()
    ref = np.load(os.path.join(golden_dir, "ref_interferometer_n5_d5.npz"))["data"]
    assert np.abs(ket - ref).max() < TOL
    assert set(log) == {"b200_apply_tile_pass"}
    assert len(log) < len(gl) / 2


def test_wrong_strides_fail_on_the_host(monkeypatch):
    """Every gather descriptor and gate launch is bounds-checked before it reaches the library: a
    wrong stride, base offset or table raises B200Error instead of addressing memory outside its
    tensor (an illegal address on the device)."""
    import torch

    from strawberryfields_b200 import circuit, lib

    monkeypatch.setattr(lib, "_lib", FakeLib())
    monkeypatch.setattr(circuit, "_TEST_HOST_MODE", True)
    c = circuit.DeviceCircuit(3, 4)
    n = c._buf.numel()
    out = torch.zeros(n, dtype=torch.complex128)
    c._gather(c._buf, None, out, [(n, 1, 0, 1)])                                  # fine
    with pytest.raises(lib.B200Error, match="operand A"):
        c._gather(c._buf, None, out, [(n, 2, 0, 1)])                              # reads past the state
    with pytest.raises(lib.B200Error, match="operand C"):
        c._gather(c._buf, None, out, [(n, 1, 0, 1)], base=(0, 0, 1))              # writes past the output
    with pytest.raises(lib.B200Error, match="operand A"):
        c._gather(c._buf, None, out[:1], [], [(n, 1, 0)], base=(-1, 0, 0))        # negative offset
    with pytest.raises(lib.B200Error, match="operand B"):
        c._gather(c._buf, out[:4], out, [(n, 1, 1, 1)])                           # second operand too short
    U = c._gen1(lib.GATE_DISPLACEMENT, 0.1, 0.2)
    c._k_gate1(U, 1, 0)                                                           # fine
    with pytest.raises(lib.B200Error, match="gate table"):
        c._k_gate1(U.reshape(-1)[:8].view(1, 2, 4), 1, 0)                         # truncated table
    c._buf = c._buf[: n // 2]
    with pytest.raises(lib.B200Error, match="state buffer"):
        c._k_gate1(U, 1, 0)                                                       # buffer smaller than the state


# ---------------------------------------------------------------------------- batched preparation / measurement
# TF-backend batch semantics (tfbackend/circuit.py:343 prepare_multimode, :612 measure_fock)
@pytest.mark.parametrize("lazy", [False, True])
def test_batched_preparations_and_measure_fock(host_backend, lazy):
    from oracle.fock_oracle import OracleBackend

    D, B, n = 5, 4, 3
    rs = np.random.RandomState(8)
    r = rs.uniform(0.1, 0.5, B)
    th = rs.uniform(0, 1.2, B)
    be = host_backend(lazy_vacuum=lazy)
    be.begin_circuit(n, cutoff_dim=D, batch_size=B)

    def program(b_, rr, tt):
        b_.prepare_coherent_state(rr, 0.3, 0)         # per-entry parameters
        b_.prepare_fock_state(1, 1)                   # the same ket for every entry
        b_.prepare_squeezed_state(0.2, 0.1, 2)
        b_.beamsplitter(tt, 0.2, 0, 1)
        b_.beamsplitter(0.4, 0.0, 1, 2)
        b_.kerr_interaction(0.1, 0)

    program(be, r, th)
    kets = be.state().ket()
    obs = []
    for b in range(B):
        ob = OracleBackend()
        ob.begin_circuit(n, cutoff_dim=D)
        # the oracle (like the reference) would mix the state on single-mode preparations: prepare the product ket
        from strawberryfields_b200.circuit import _coherent, _squeezed

        f1 = np.zeros(D, dtype=complex)
        f1[1] = 1
        ob.prepare_ket_state(np.multiply.outer(np.multiply.outer(_coherent(r[b], 0.3, D), f1), _squeezed(0.2, 0.1, D)),
                             [0, 1, 2])
        ob.beamsplitter(th[b], 0.2, 0, 1)
        ob.beamsplitter(0.4, 0.0, 1, 2)
        ob.kerr_interaction(0.1, 0)
        assert np.abs(kets[b] - ob.state().data).max() < TOL
        obs.append(ob)
    # seeded measurement: one uniform per entry, in batch order -- the same draws as B sequential oracle runs
    np.random.seed(3)
    got = be.measure_fock([2, 0])
    assert got.shape == (B, 2)
    np.random.seed(3)
    post = be.state().ket()
    for b in range(B):
        want = obs[b].measure_fock([2, 0])
        assert np.array_equal(got[b], np.asarray(want).reshape(-1))
        assert np.abs(post[b] - obs[b].state().data).max() < TOL
    # post-selection: one list for all entries, or one per entry
    be.reset()
    program(be, r, th)
    sel = np.array([[0, 1], [1, 0], [0, 0], [1, 1]])
    out = be.measure_fock([0, 2], select=sel)
    assert np.array_equal(out, sel)
    post = be.state().ket()
    for b in range(B):
        ob = OracleBackend()
        ob.begin_circuit(n, cutoff_dim=D)
        f1 = np.zeros(D, dtype=complex)
        f1[1] = 1
        ob.prepare_ket_state(np.multiply.outer(np.multiply.outer(_coherent(r[b], 0.3, D), f1), _squeezed(0.2, 0.1, D)),
                             [0, 1, 2])
        ob.beamsplitter(th[b], 0.2, 0, 1)
        ob.beamsplitter(0.4, 0.0, 1, 2)
        ob.kerr_interaction(0.1, 0)
        ob.measure_fock([0, 2], select=[int(x) for x in sel[b]])
        assert np.abs(post[b] - ob.state().data).max() < TOL
    with pytest.raises(ValueError, match="shape of 'select'"):
        be.measure_fock([0, 2], select=[1, 0, 0])


def batched_general_preparations(make_backend):
    """Preparations a batched circuit cannot queue as a rank-one operator (TF-backend ``prepare_multimode`` +
    ``_replace_and_update``, tfbackend/circuit.py:343-396): multi-mode kets, density matrices, modes that have
    been touched, one state for all entries or one per entry -- every batch entry against the oracle run
    alone.  (Shared with tests/test_vacuum_lazy.py, which holds the GPU variant: last file of the suite.)"""
    from oracle.fock_oracle import OracleBackend

    D, B, n = 4, 3, 3
    rs = np.random.RandomState(21)

    th = rs.uniform(0.2, 1.0, B)
    k2 = rs.randn(B, D, D) + 1j * rs.randn(B, D, D)
    k2 /= np.sqrt((np.abs(k2) ** 2).sum(axis=(1, 2), keepdims=True))          # per-entry two-mode kets
    k1 = rs.randn(D) + 1j * rs.randn(D)
    k1 /= np.linalg.norm(k1)                                                  # one single-mode ket for all
    a = rs.randn(B, D, D) + 1j * rs.randn(B, D, D)
    dm = np.einsum("bij,bkj->bik", a, a.conj())
    dm /= np.trace(dm, axis1=1, axis2=2)[:, None, None]                       # per-entry single-mode density matrices
    full = rs.randn(D, D, D) + 1j * rs.randn(D, D, D)
    full /= np.sqrt((np.abs(full) ** 2).sum())                                # whole register, one for all

    def steps(b_, e):
        """e = batch entry for the oracle (None: the batched backend gets the whole arrays)"""
        pick = (lambda x: x) if e is None else (lambda x: x[e])
        b_.prepare_ket_state(pick(k2), [2, 0])                 # untouched modes, unordered: stays pure
        b_.beamsplitter(pick(th) if e is None else float(th[e]), 0.3, 0, 1)
        yield "pure two-mode kets"
        b_.prepare_ket_state(k1, [1])                          # touched mode: the state becomes mixed
        b_.squeeze(0.2, 0.1, 1)
        yield "ket on a touched mode"
        b_.prepare_dm_state(pick(dm), [0])                     # per-entry density matrices
        b_.beamsplitter(0.5, 0.0, 0, 2)
        yield "per-entry density matrix"
        b_.prepare_dm_state(np.outer(k1, k1.conj()), [2])      # one density matrix for all
        yield "broadcast density matrix"
        b_.prepare_ket_state(full, [1, 2, 0])                  # whole register, permuted: pure again
        b_.kerr_interaction(0.2, 2)
        yield "whole register"

    be = make_backend()
    be.begin_circuit(n, cutoff_dim=D, batch_size=B)
    obs = []
    for e in range(B):
        ob = OracleBackend()
        ob.begin_circuit(n, cutoff_dim=D)
        obs.append((ob, steps(ob, e)))
    for label in steps(be, None):
        st = be.state()
        got = st.dm()
        for e, (ob, it) in enumerate(obs):
            assert next(it) == label
            assert np.abs(got[e] - ob.state().dm()).max() < TOL, (label, e)
        if label in ("pure two-mode kets", "whole register"):
            assert st.is_pure
    with pytest.raises(ValueError, match="Incorrect shape"):
        be.prepare_ket_state(np.ones((B + 2, D)), [0])
    with pytest.raises(ValueError, match="multiple times"):
        be.prepare_ket_state(np.ones((D, D)) / D, [1, 1])


def test_batched_general_preparations(monkeypatch):
    from strawberryfields_b200 import circuit, lib
    from strawberryfields_b200.backend import B200FockBackend

    monkeypatch.setattr(lib, "_lib", FakeLib())
    monkeypatch.setattr(circuit, "_TEST_HOST_MODE", True)
    for lazy in (False, True):
        def make():
            be = B200FockBackend()
            orig = be.begin_circuit
            be.begin_circuit = lambda n_, **kw: orig(n_, lazy_vacuum=lazy, **kw)
            return be

        batched_general_preparations(make)


def batched_homodyne(make_backend):
    """Batched ``measure_homodyne`` (tfbackend/circuit.py:812-941 semantics on the Fock backend's sampling grid):
    the B draws in batch order are the draws of B oracle runs made one after the other."""
    from oracle.fock_oracle import OracleBackend

    D, B, n = 6, 3, 2
    r = np.array([0.2, 0.5, 0.35])
    phi = 0.4

    def program(b_, rr):
        b_.prepare_coherent_state(rr, 0.3, 0)
        b_.squeeze(0.3, 0.2, 1)
        b_.beamsplitter(0.6, 0.1, 0, 1)

    for pure in (True, False):
        be = make_backend()
        be.begin_circuit(n, cutoff_dim=D, batch_size=B, pure=pure)
        program(be, r)
        np.random.seed(17)
        got = be.measure_homodyne(phi, 0)
        assert got.shape == (B, 1)
        post = be.state().dm()
        np.random.seed(17)
        for e in range(B):
            ob = OracleBackend()
            ob.begin_circuit(n, cutoff_dim=D, pure=pure)
            program(ob, float(r[e]))
            want = ob.measure_homodyne(phi, 0)
            assert abs(float(np.asarray(want).reshape(-1)[0]) - got[e, 0]) < 1e-12
            assert np.abs(post[e] - ob.state().dm()).max() < TOL
        # post-selection: one value per entry
        be.reset(pure=pure)
        program(be, r)
        sel = np.array([0.3, -0.2, 1.1])
        assert np.array_equal(be.measure_homodyne(phi, 1, select=sel), sel.reshape(B, 1))
        post = be.state().dm()
        for e in range(B):
            ob = OracleBackend()
            ob.begin_circuit(n, cutoff_dim=D, pure=pure)
            program(ob, float(r[e]))
            ob.measure_homodyne(phi, 1, select=float(sel[e]))
            assert np.abs(post[e] - ob.state().dm()).max() < TOL
    with pytest.raises(ValueError, match="batch_size"):
        be.measure_homodyne(phi, 0, select=[0.1, 0.2])
    with pytest.raises(TypeError, match="numeric"):
        be.measure_homodyne(phi, 0, select="x")


def test_batched_homodyne(monkeypatch):
    from strawberryfields_b200 import circuit, lib
    from strawberryfields_b200.backend import B200FockBackend

    monkeypatch.setattr(lib, "_lib", FakeLib())
    monkeypatch.setattr(circuit, "_TEST_HOST_MODE", True)
    for lazy in (False, True):
        def make():
            be = B200FockBackend()
            orig = be.begin_circuit
            be.begin_circuit = lambda n_, **kw: orig(n_, lazy_vacuum=lazy, **kw)
            return be

        batched_homodyne(make)
