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
