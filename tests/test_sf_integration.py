"""Drop-in check with the real Strawberry Fields front end (build container only, needs
/root/reference): sf.Engine("b200fock") vs sf.Engine("fock") on the same Program objects."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.reference
def test_engine_drop_in():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "sf_integration_worker.py")],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    out = json.loads(line[len("RESULT "):])
    assert out["registered"] and out["is_basefock"] and out["state_is_basefockstate"]
    assert out["boson_probs_err"] < 1e-12
    # /root/reference/tests/integration/test_algorithms.py:177
    assert abs(out["golden"][0] - 0.174689160486) < 1e-11 and abs(out["golden"][1] - 0.106441927246) < 1e-11
    assert out["samples_equal"]
    assert out["dm_err"] < 1e-12
    assert out["mean_photon_err"] < 1e-10
    assert out["wigner_shape"] == [5, 5]


@pytest.mark.reference
def test_reference_suite_subset_runs_on_b200fock():
    """Three files of the reference's own backend test-suite (beamsplitter on every mode pair, loss
    channel, Fock measurement) with FockBackend swapped for B200FockBackend (tests/b200_ref_plugin.py).
    The full run is recorded in profiles/r01_reference_suite.md."""
    files = ["test_loss_channel.py", "test_fock_measurement.py", "test_modes.py"]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "tests")]))
    res = subprocess.run([sys.executable, "-m", "pytest", "-p", "b200_ref_plugin", "-p", "no:cacheprovider", "-m", "fock",
                          "-q"] + [os.path.join("/root/reference/tests/backend", f) for f in files],
                         capture_output=True, text=True, timeout=900, cwd="/tmp", env=env)
    tail = res.stdout.strip().splitlines()[-1]
    assert res.returncode == 0 and " passed" in tail and "failed" not in tail, res.stdout[-2000:]
