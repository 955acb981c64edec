"""Pin oracle/fock_oracle.py against (a) the reference's own golden vectors and
(b) fixtures written by the UNMODIFIED reference backend (oracle/make_golden.py)."""
import glob
import json
import os

import numpy as np
import pytest

import scripts
from oracle.fock_oracle import OracleBackend

TOL = 1e-12


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, f"ref_{name}.npz"))


@pytest.mark.parametrize("style", ["vector", "reference"])
@pytest.mark.parametrize("script", scripts.all_scripts(), ids=lambda s: s[0])
def test_script_matches_reference_fixture(script, style, golden_dir):
    if style == "reference" and script[0] in ("boson_sampling_d7", "every_gate_n3_d5_pure"):
        pytest.skip("slow in the loop style; covered by the vector style")
    ref = _load(golden_dir, script[0])
    rets, st = scripts.run_script(OracleBackend(style=style), script)
    assert bool(ref["pure"]) == st.is_pure
    if "data" in ref:
        assert st.data.shape == ref["data"].shape
        assert np.abs(st.data - ref["data"]).max() < TOL
    else:
        assert np.abs(st.all_fock_probs() - ref["probs"]).max() < TOL
    for i, r in enumerate(rets):
        assert np.array_equal(r, ref[f"ret{i}"])


def test_reference_golden_boson_sampling(golden_dir):
    # /root/reference/tests/integration/test_algorithms.py:177 (cutoff 6 there; the
    # two listed probabilities do not depend on the cutoff for cutoff >= 4)
    _, st = scripts.run_script(OracleBackend(), scripts.boson_sampling(6))
    got = [st.fock_prob([1, 1, 0, 1]), st.fock_prob([2, 0, 0, 1])]
    assert np.allclose(got, [0.174689160486, 0.106441927246], atol=1e-11, rtol=0)


def test_reference_golden_gbs():
    # /root/reference/tests/integration/test_algorithms.py:120-136 (tol 1e-3 there:
    # the listed values are the infinite-cutoff ones)
    _, st = scripts.run_script(OracleBackend(), scripts.gbs(6))
    states = [[0, 0, 0, 0], [1, 1, 0, 0], [0, 1, 0, 1], [1, 1, 1, 1], [2, 0, 0, 0]]
    want = [0.176378447614135, 0.0685595637122246, 0.002056097258977398, 0.00834294639986785, 0.01031294525345511]
    assert np.allclose([st.fock_prob(s) for s in states], want, atol=1e-3, rtol=0)


@pytest.mark.parametrize("pure", [True, False])
def test_reference_golden_hamiltonian_simulation(pure):
    # /root/reference/tests/integration/test_algorithms.py:190-225
    _, st = scripts.run_script(OracleBackend(), scripts.hamiltonian_simulation(pure))
    got = [st.fock_prob(s) for s in ([0, 2], [1, 1], [2, 0])]
    assert np.allclose(got, [0.52240124572, 0.235652876857, 0.241945877423], atol=1e-10, rtol=0)


@pytest.mark.parametrize("N,D", [(4, 6), (5, 5)])
def test_compiled_interferometer_matches_reference(N, D, golden_dir):
    with open(os.path.join(golden_dir, f"interferometer_n{N}.json")) as f:
        gl = json.load(f)["gates"]
    be = OracleBackend()
    be.begin_circuit(N, cutoff_dim=D)
    for g in gl:
        getattr(be, g[0])(*g[1:])
    ref = np.load(os.path.join(golden_dir, f"ref_interferometer_n{N}_d{D}.npz"))["data"]
    assert np.abs(be.state().data - ref).max() < TOL


def test_measure_fock_deterministic_on_fock_input():
    # /root/reference/tests/backend/test_fock_measurement.py:107-139
    be = OracleBackend()
    be.begin_circuit(3, cutoff_dim=5)
    for m, n in enumerate([2, 0, 3]):
        be.prepare_fock_state(n, m)
    out = be.measure_fock([0, 1, 2])
    assert out.tolist() == [[2, 0, 3]]
    assert be.is_vacuum(1e-12)


def test_loss_coherent_amplitude():
    # /root/reference/tests/backend/test_loss_channel.py:188-207: alpha -> sqrt(T) alpha
    from oracle import gates

    D, T, r, phi = 14, 0.6, 0.5, 0.3
    be = OracleBackend()
    be.begin_circuit(1, cutoff_dim=D)
    be.prepare_coherent_state(r, phi, 0)
    be.loss(T, 0)
    v = gates.coherent_state(np.sqrt(T) * r, phi, D)
    assert np.abs(be.state().dm() - np.outer(v, v.conj())).max() < 1e-9
