"""ORACLE helper (build container only): write ``tests/golden/ref_observables.json`` -- the
values the UNMODIFIED reference ``BaseFockState`` (``/root/reference/strawberryfields/backends/
states.py:657-985``) returns for the observables of SURVEY 8(f)2 on the final states of a few
``tests/scripts.py`` scripts.

    python -m oracle.make_golden_observables        # needs /root/reference (see ref_shim.py)

``observable_cases`` (the list of calls and their arguments) lives in ``tests/scripts.py`` so the
tests evaluate exactly the same calls on the oracle and on the b200fock state objects.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from oracle import ref_shim

    ref_shim.install()
    from strawberryfields.backends.fockbackend import FockBackend
    import scripts

    out = {}
    for sc in scripts.observable_scripts():
        _, st = scripts.run_script(FockBackend(), sc)
        vals = {}
        for key, method, args in scripts.observable_cases(sc):
            v = getattr(st, method)(*args)
            vals[key] = np.asarray(v, dtype=np.float64).tolist()
        out[sc[0]] = vals
        print("wrote", sc[0], len(vals), "observables")
    with open(os.path.join(GOLD, "ref_observables.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
