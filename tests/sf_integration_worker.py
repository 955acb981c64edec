"""Run by tests/test_sf_integration.py in a fresh process (build container only): the UNMODIFIED
reference front end / compiler / engine drive `b200fock` through sf.Engine, next to the
reference's own fock backend.  Device calls go to the numpy double of the C ABI (no GPU here)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402

sf = ref_shim.install()  # must precede the plugin import so that it derives from the real BaseFock
import strawberryfields_b200 as plugin  # noqa: E402
from strawberryfields_b200 import circuit, lib  # noqa: E402
from fake_lib import FakeLib  # noqa: E402
from strawberryfields import ops  # noqa: E402
from strawberryfields.backends import BaseFock  # noqa: E402
from strawberryfields.backends.states import BaseFockState  # noqa: E402

lib._lib = FakeLib()
circuit._TEST_HOST_MODE = True
plugin.register()

out = {"registered": "b200fock" in sf.backends.local_backends,
       "is_basefock": isinstance(plugin.B200FockBackend(), BaseFock)}


def boson_sampling():
    prog = sf.Program(4)
    with prog.context as q:
        ops.Fock(1) | q[0]
        ops.Fock(1) | q[1]
        ops.Vac | q[2]
        ops.Fock(1) | q[3]
        ops.Rgate(0.5719) | q[0]
        ops.Rgate(-1.9782) | q[1]
        ops.Rgate(2.0603) | q[2]
        ops.Rgate(0.0644) | q[3]
        ops.BSgate(0.7804, 0.8578) | (q[0], q[1])
        ops.BSgate(0.06406, 0.5165) | (q[2], q[3])
        ops.BSgate(0.473, 0.1176) | (q[1], q[2])
        ops.BSgate(0.563, 0.1517) | (q[0], q[1])
        ops.BSgate(0.1323, 0.9946) | (q[2], q[3])
        ops.BSgate(0.311, 0.3231) | (q[1], q[2])
        ops.BSgate(0.4348, 0.0798) | (q[0], q[1])
        ops.BSgate(0.4368, 0.6157) | (q[2], q[3])
    return prog


def mixed_program():
    prog = sf.Program(3)
    U = sf.utils.random_interferometer(3)
    with prog.context as q:
        ops.Sgate(0.3, 0.2) | q[0]
        ops.Dgate(0.2, 0.4) | q[1]
        ops.Interferometer(U) | q
        ops.Kgate(0.1) | q[2]
        ops.LossChannel(0.8) | q[1]
        ops.MZgate(0.3, 0.7) | (q[0], q[2])
        ops.S2gate(0.1, 0.3) | (q[1], q[2])
        ops.MeasureFock() | q[0]
    return prog


# BASELINE config 1 through both engines
res = {}
for name in ("fock", "b200fock"):
    eng = sf.Engine(name, backend_options={"cutoff_dim": 5})
    res[name] = eng.run(boson_sampling()).state
out["state_is_basefockstate"] = isinstance(res["b200fock"], BaseFockState)
out["boson_probs_err"] = float(np.abs(res["fock"].all_fock_probs() - res["b200fock"].all_fock_probs()).max())
out["golden"] = [float(res["b200fock"].fock_prob([1, 1, 0, 1])), float(res["b200fock"].fock_prob([2, 0, 0, 1]))]

# decompositions, loss, measurement with the global numpy stream
np.random.seed(3)
prog = mixed_program()
samples, dms = {}, {}
for name in ("fock", "b200fock"):
    np.random.seed(11)
    eng = sf.Engine(name, backend_options={"cutoff_dim": 5})
    r = eng.run(prog)
    samples[name] = np.asarray(r.samples).tolist()
    dms[name] = r.state.dm()
out["samples_equal"] = samples["fock"] == samples["b200fock"]
out["dm_err"] = float(np.abs(dms["fock"] - dms["b200fock"]).max())
# observables inherited from the reference's BaseFockState work on the device-backed state
st = res["b200fock"]
out["mean_photon_err"] = float(abs(st.mean_photon(0)[0] - res["fock"].mean_photon(0)[0]))
out["wigner_shape"] = list(np.asarray(st.wigner(0, np.linspace(-2, 2, 5), np.linspace(-2, 2, 5))).shape)
print("RESULT " + json.dumps(out))
