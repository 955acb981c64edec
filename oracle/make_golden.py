"""ORACLE helper (build container only): write ``tests/golden/*.npz`` by running
the UNMODIFIED reference ``FockBackend`` (``/root/reference/strawberryfields/
backends/fockbackend/backend.py``) on the scripts in ``tests/scripts.py``, and
the reference front end + ``fock`` compiler on the BASELINE circuit shapes.

    python -m oracle.make_golden            # needs /root/reference (see ref_shim.py)

The gate tensors inside the reference come from ``oracle/gates.py`` (thewalrus is
absent), so these fixtures pin everything except those five
recursions, which ``tests/test_oracle_gates.py`` pins against ``expm``.
Each fixture stores the final state tensor, its purity flag, and every value a
backend call returned (measurement outcomes).
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


def _compiled_gate_list(sf, prog):
    """Compile with the reference 'fock' compiler (program.py:633-768) and
    translate each Command into the backend call ``Operation._apply`` makes
    (ops.py: Sgate 1636, Dgate 1523, Rgate 1440, BSgate 1946, Kgate 2160 ...)."""
    table = {
        "Sgate": "squeeze",
        "Dgate": "displacement",
        "Rgate": "rotation",
        "BSgate": "beamsplitter",
        "MZgate": "mzgate",
        "S2gate": "two_mode_squeeze",
        "Kgate": "kerr_interaction",
        "Vgate": "cubic_phase",
        "CKgate": "cross_kerr_interaction",
    }
    out = []
    for cmd in prog.compile(compiler="fock").circuit:
        name = cmd.op.__class__.__name__
        params = [float(x) for x in cmd.op.p]
        if params and params[0] == 0:  # Gate.apply skips identity gates (ops.py:494-509)
            continue
        out.append([table[name]] + params + [r.ind for r in cmd.reg])
    return out


def interferometer_circuit(sf, N, seed=42):
    """BASELINE config 2/5 generator: Sgate+Dgate on every mode,
    then Interferometer(random_interferometer(N)), rectangular mesh."""
    from strawberryfields import ops

    np.random.seed(seed)
    U = sf.utils.random_interferometer(N)
    r = np.random.uniform(0, 0.3, N)
    pr = np.random.uniform(0, 2 * np.pi, N)
    a = np.random.uniform(0, 0.5, N)
    pa = np.random.uniform(0, 2 * np.pi, N)
    prog = sf.Program(N)
    with prog.context as q:
        for i in range(N):
            ops.Sgate(r[i], pr[i]) | q[i]
            ops.Dgate(a[i], pa[i]) | q[i]
        ops.Interferometer(U) | q
    return _compiled_gate_list(sf, prog), U


def main():
    from oracle import ref_shim

    sf = ref_shim.install()
    from strawberryfields.backends.fockbackend import FockBackend
    import scripts

    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[1:]  # e.g. `python -m oracle.make_golden homodyne`: just the scripts whose name matches
    for sc in scripts.all_scripts():
        name = sc[0]
        if only and not any(o in name for o in only):
            continue
        be = FockBackend()
        rets, st = scripts.run_script(be, sc)
        data = np.ascontiguousarray(st.data)
        out = {"data": data, "pure": np.array(st.is_pure), "n_modes": np.array(st.num_modes)}
        if data.size > 200_000:  # keep fixtures small: store probabilities only
            out = {
                "probs": np.ascontiguousarray(st.all_fock_probs()),
                "pure": np.array(st.is_pure),
                "n_modes": np.array(st.num_modes),
            }
        for i, r in enumerate(rets):
            out[f"ret{i}"] = r
        np.savez_compressed(os.path.join(GOLD, f"ref_{name}.npz"), **out)
        print("wrote", name, {k: getattr(v, "shape", None) for k, v in out.items()})

    if only:
        return
    # compiled BASELINE-shaped circuits: gate list from the reference front end,
    # final ket from the reference backend through sf.Engine
    for N, D in ((4, 6), (5, 5)):
        gl, U = interferometer_circuit(sf, N)
        be = FockBackend()
        be.begin_circuit(N, cutoff_dim=D)
        for g in gl:
            getattr(be, g[0])(*g[1:])
        ket = np.ascontiguousarray(be.state().data)
        with open(os.path.join(GOLD, f"interferometer_n{N}.json"), "w") as f:
            json.dump({"N": N, "gates": gl, "U_re": U.real.tolist(), "U_im": U.imag.tolist()}, f)
        np.savez_compressed(os.path.join(GOLD, f"ref_interferometer_n{N}_d{D}.npz"), data=ket)
        print("wrote interferometer", N, D, len(gl))
    # full-size gate lists only (no reference state: 33 min .. 6 h on the CPU)
    for N in (8, 9, 10):
        gl, U = interferometer_circuit(sf, N)
        with open(os.path.join(GOLD, f"interferometer_n{N}.json"), "w") as f:
            json.dump({"N": N, "gates": gl, "U_re": U.real.tolist(), "U_im": U.imag.tolist()}, f)
        print("wrote gate list", N, len(gl))


if __name__ == "__main__":
    main()
