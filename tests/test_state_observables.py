"""Observables of the state object: ``fidelity_coherent``, ``mean_photon``,
``quad_expectation``, ``poly_quad_expectation``, ``wigner``, ``number_expectation``,
``parity_expectation`` computed from device reductions of the resident state.

* tests/golden/ref_observables.json holds what the UNMODIFIED reference ``BaseFockState``
  returned (written by oracle/make_golden_observables.py in the build container);
* the oracle's restatements are pinned against those values;
* the b200fock state object (numpy double of the C ABI here, CUDA kernels with ``-m gpu``) is
  compared with both.  Tolerance 1e-12 absolute (values are O(1))."""
import json
import os

import numpy as np
import pytest

import scripts
from fake_lib import FakeLib
from oracle.fock_oracle import OracleBackend

TOL = 1e-12


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "ref_observables.json")) as f:
        return json.load(f)


@pytest.fixture(params=["host", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    from strawberryfields_b200 import circuit, lib
    from strawberryfields_b200.backend import B200FockBackend

    if request.param == "host":
        monkeypatch.setattr(lib, "_lib", FakeLib())
        monkeypatch.setattr(circuit, "_TEST_HOST_MODE", True)
    return B200FockBackend


def _values(st, script, skip=()):
    return {key: np.asarray(getattr(st, method)(*args), dtype=np.float64)
            for key, method, args in scripts.observable_cases(script)
            if hasattr(st, method) and method not in skip}


@pytest.mark.parametrize("script", scripts.observable_scripts(), ids=lambda s: s[0])
def test_oracle_matches_reference(script, golden):
    _, st = scripts.run_script(OracleBackend(), script)
    got = _values(st, script)
    want = golden[script[0]]
    assert set(got) == {k for k in want if not k.startswith("polyquad")}  # polyquad: fixture-pinned only
    for key in got:
        assert np.abs(got[key] - np.asarray(want[key])).max() < TOL, key


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("script", scripts.observable_scripts(), ids=lambda s: s[0])
def test_b200_matches_reference_and_oracle(script, strict, backend, golden):
    be = backend()
    orig = be.begin_circuit
    be.begin_circuit = lambda n, **kw: orig(n, strict_purity=strict, **kw)
    _, st = scripts.run_script(be, script)
    got = _values(st, script)
    _, ost = scripts.run_script(OracleBackend(), script)
    want_o = _values(ost, script)
    want_r = golden[script[0]]
    assert set(got) == set(want_r)
    for key in want_r:
        scale = max(1.0, np.abs(np.asarray(want_r[key])).max())  # polyquad variances are O(100)
        assert np.abs(got[key] - np.asarray(want_r[key])).max() < TOL * scale, key
        if key in want_o:
            assert np.abs(got[key] - want_o[key]).max() < TOL, key


def test_after_tile_permutation(backend):
    """The reductions go through the logical -> physical axis map: same values when tile
    passes have permuted the axes (fuse="tile")."""
    script = scripts.every_gate(4, 5, True, seed=14, prep=False)
    vals = []
    for fuse in ("fold", "tile"):
        be = backend()
        be.begin_circuit(4, cutoff_dim=5, fuse=fuse)
        for method, args, kwargs in script[4]:
            getattr(be, method)(*args, **kwargs)
        st = be.state()
        vals.append(_values(st, script, skip=("poly_quad_expectation",)))  # host-only algebra, slow at 4 modes
    for key in vals[0]:
        assert np.abs(vals[0][key] - vals[1][key]).max() < TOL, key


def test_batched_observables(backend):
    """Batched circuits: one value per batch entry, equal to the oracle run entry by entry."""
    B, n, D = 3, 2, 6
    r = np.array([0.1, 0.25, 0.4])
    be = backend()
    be.begin_circuit(n, cutoff_dim=D, batch_size=B)
    be.displacement(r, np.array([0.3, 0.6, 0.9]), 0)
    be.squeeze(0.2, 0.1, 1)
    be.beamsplitter(np.array([0.4, 0.5, 0.6]), 0.3, 0, 1)
    st = be.state()
    alphas = [0.2 + 0.1j, -0.1j]
    got = {
        "fc": st.fidelity_coherent(alphas),
        "num": np.stack(st.number_expectation([0, 1])),
        "par": st.parity_expectation([1]),
        "quad": np.stack(st.quad_expectation(0, 0.3)),
        "mp": np.stack(st.mean_photon(1)),
    }
    for b in range(B):
        ob = OracleBackend()
        ob.begin_circuit(n, cutoff_dim=D)
        ob.displacement(float(r[b]), [0.3, 0.6, 0.9][b], 0)
        ob.squeeze(0.2, 0.1, 1)
        ob.beamsplitter([0.4, 0.5, 0.6][b], 0.3, 0, 1)
        ost = ob.state()
        assert abs(got["fc"][b] - ost.fidelity_coherent(alphas)) < TOL
        assert np.abs(got["num"][:, b] - np.array(ost.number_expectation([0, 1]))).max() < TOL
        assert abs(got["par"][b] - ost.parity_expectation([1])) < TOL
        assert np.abs(got["quad"][:, b] - np.array(ost.quad_expectation(0, 0.3))).max() < TOL
        assert np.abs(got["mp"][:, b] - np.array(ost.mean_photon(1))).max() < TOL


def test_sample_fock(backend):
    """Multi-shot sampling without collapse: outcomes follow the joint distribution, columns follow
    the requested mode order, the state is unchanged, and one shot equals what measure_fock draws."""
    n, D = 3, 5
    be, ob = backend(), OracleBackend()
    for b in (be, ob):
        b.begin_circuit(n, cutoff_dim=D)
        b.squeeze(0.5, 0.2, 0)
        b.displacement(0.6, 0.1, 2)
        b.beamsplitter(0.7, 0.3, 0, 1)
    st = be.state()
    probs = ob.state().all_fock_probs()
    np.random.seed(3)
    s = st.sample_fock(4000, modes=[2, 0])
    assert s.shape == (4000, 2) and s.dtype == np.int64
    marg = probs.sum(axis=1)                                # [n0, n2]
    freq = np.zeros((D, D))
    np.add.at(freq, (s[:, 1], s[:, 0]), 1.0 / len(s))      # column 0 is mode 2
    assert np.abs(freq - marg).max() < 0.03
    assert np.abs(be.state().ket() - ob.state().data).max() < TOL   # no collapse
    np.random.seed(9)
    one = st.sample_fock(1, modes=[1, 2])
    np.random.seed(9)
    assert np.array_equal(one, ob.measure_fock([1, 2]))
    with pytest.raises(ValueError, match="not valid"):
        st.sample_fock(2, modes=[0, 0])


def test_argument_errors(backend):
    be = backend()
    be.begin_circuit(2, cutoff_dim=4)
    st = be.state()
    with pytest.raises(ValueError, match="must match the number of modes"):
        st.fidelity_coherent([0.1])
    with pytest.raises(ValueError, match="no duplicates"):
        st.number_expectation([0, 0])
