"""Program I/O: Blackbird / XIR loaders and writers and the state checkpoint
(``strawberryfields_b200/io.py``; reference: ``strawberryfields/io/__init__.py:67,145,169``).
The loaders are checked on the reference's own example script, on round trips, against the oracle,
and -- where /root/reference exists -- against the reference front end compiling the same program."""
import io as _io
import os

import numpy as np
import pytest

from fake_lib import FakeLib
from strawberryfields_b200 import io as bio
from strawberryfields_b200 import workloads as W

X8 = """\
name example_job_X8
version 1.0
target X8_01 (shots = 20)

complex array U[4, 4] =
    -0.13879438-0.47517904j, -0.29303954-0.47264099j, -0.43951987+0.12977568j, -0.03496718-0.48418713j
    0.06065372-0.11292765j, 0.54733962+0.1215551j, -0.50721513+0.56195975j, -0.15923161+0.26606674j
    0.42212573-0.53182417j, -0.2642572+0.50625182j, 0.19448705+0.28321781j, 0.30281396-0.05582391j
    0.43097587-0.30288974j, 0.07419772-0.21155126j, 0.28335618-0.13633175j, -0.75113453+0.09580304j

# Initial states are two-mode squeezed states
S2gate(1.0, 0.0) | [0, 4]
S2gate(1.0, 0.0) | [1, 5]
S2gate(1.0, 0.0) | [3, 7]

Interferometer(U) | [0, 1, 2, 3]
BSgate(0.543, 0.123) | [2, 0]
Rgate(0.453) | 1
MZgate(0.65, -0.54) | [2, 3]

Interferometer(U) | [4, 5, 6, 7]
BSgate(0.543, 0.123) | [6, 4]
Rgate(0.453) | 5
MZgate(0.65, -0.54) | [6, 7]

MeasureFock() | [0, 1, 2, 3, 4, 5, 6, 7]
"""


@pytest.fixture
def host(monkeypatch):
    from strawberryfields_b200 import circuit, lib

    monkeypatch.setattr(lib, "_lib", FakeLib())
    monkeypatch.setattr(circuit, "_TEST_HOST_MODE", True)


def test_blackbird_header_arrays_and_operations():
    prog = bio.loads(X8)
    assert prog.name == "example_job_X8" and prog.version == "1.0"
    assert prog.target == {"name": "X8_01", "options": {"shots": 20}} and prog.run_options == {"shots": 20}
    assert prog.num_subsystems == 8 and len(prog.operations) == 12
    op = prog.operations[3]
    assert op["op"] == "Interferometer" and op["modes"] == [0, 1, 2, 3] and op["args"][0].shape == (4, 4)
    assert op["args"][0][1, 2] == -0.50721513 + 0.56195975j
    assert prog.operations[5] == {"op": "Rgate", "args": [0.453], "kwargs": {}, "modes": [1]}
    assert prog.operations[-1]["op"] == "MeasureFock" and prog.operations[-1]["args"] == []


def test_reference_example_file_loads_if_present():
    path = "/root/reference/examples/example_job_X8.xbb"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    prog = bio.load(path)
    assert [o["op"] for o in prog.operations] == [o["op"] for o in bio.loads(X8).operations]


def test_expressions_variables_and_keywords():
    prog = bio.loads("""name t
version 1.0
float alpha = 0.3423
int n = 2
Coherent(alpha, sqrt(pi)) | 0
Sgate(-alpha*2, phi=pi/4) | 1
Fock(n) | 2
MeasureHomodyne(phi=0.43, select=0.32) | 2   # trailing comment
MeasureX | 0
""")
    ops = prog.operations
    assert ops[0]["args"] == [0.3423, np.sqrt(np.pi)]
    assert ops[1]["args"] == [-0.6846] and ops[1]["kwargs"] == {"phi": np.pi / 4}
    assert ops[2]["args"] == [2]
    assert ops[3]["kwargs"] == {"phi": 0.43, "select": 0.32} and ops[4] == {"op": "MeasureX", "args": [], "kwargs": {}, "modes": [0]}
    calls = prog.calls()
    assert calls[0] == ("prepare_coherent_state", 0.3423, float(np.sqrt(np.pi)), 0)
    assert calls[-2] == ("measure_homodyne", 0.43, 2, {"select": 0.32}) and calls[-1] == ("measure_homodyne", 0.0, 0, {"select": None})


def test_for_loops_are_unrolled():
    prog = bio.loads("""name loops
version 1.0
float array phase[1, 4] =
    0.1, 0.2, 0.3, 0.4
for int m in 0:3
    MZgate(phase[0, m], 0.5) | [m, m + 1]
    Rgate(0.1 * m) | m

for int k in [3, 1]
    MeasureFock() | k
for int j in 0:6:2
    Sgate(0.1) | j
""")
    ops = prog.operations
    assert [o["op"] for o in ops] == ["MZgate", "Rgate"] * 3 + ["MeasureFock"] * 2 + ["Sgate"] * 3
    assert [o["modes"] for o in ops[:6]] == [[0, 1], [0], [1, 2], [1], [2, 3], [2]]
    assert ops[2]["args"] == [0.2, 0.5] and ops[3]["args"] == [0.1]
    assert [o["modes"] for o in ops[6:8]] == [[3], [1]] and [o["modes"] for o in ops[8:]] == [[0], [2], [4]]
    with pytest.raises(bio.ProgramSyntaxError):
        bio.loads("name x\nversion 1.0\nfor int m in 0:3\nSgate(0.1) | m\n")           # empty block
    with pytest.raises(bio.ProgramSyntaxError):
        bio.loads("name x\nversion 1.0\nfor int m in 0:3\n    float y = 2\n")           # not an operation
    with pytest.raises(bio.ProgramSyntaxError):
        bio.loads("name x\nversion 1.0\nfor int m in 0:2\n    Sgate(0.1) | m - 1\n")    # negative mode


@pytest.mark.parametrize("bad,exc", [("Sgate(0.3 | 0", bio.ProgramSyntaxError), ("Sgate(foo) | 0", NameError),
                                     ("type tdm (temporal_modes=2)\nSgate(0.1, 0.0) | 0", NotImplementedError),
                                     ("type tdm 3", bio.ProgramSyntaxError),
                                     ("include \"lib.xbb\"", NotImplementedError),
                                     ("Sgate(__import__('os')) | 0", bio.ProgramSyntaxError)])
def test_bad_scripts_are_refused(bad, exc):
    with pytest.raises(exc):
        bio.loads("name x\nversion 1.0\n" + bad + "\n")


def test_invalid_ir_names():
    with pytest.raises(ValueError, match="not recognized as a valid IR option"):
        bio.loads("", ir="qasm")
    with pytest.raises(ValueError, match="not recognized as a valid IR option"):
        bio.dumps(bio.CircuitProgram(), ir="qasm")


def test_xir_script():
    prog = bio.loads("""options:
    _name_: test_program;
    cutoff_dim: 5;
    shots: 2;
end;
use xstd;
Vacuum | [1];
Squeezed(0.12, 0.0) | [2];
Sgate(1, 0.0) | [0];      // a comment
S2gate(0.543, -0.12) | [0, 3];
Interferometer([[(0.6+0j), (0.8+0j)], [(-0.8+0j), (0.6+0j)]]) | [0, 1];
MeasureHomodyne(phi: 0.43, select: 0.32) | [2];
""", ir="xir")
    assert prog.name == "test_program" and prog.backend_options == {"cutoff_dim": 5} and prog.run_options == {"shots": 2}
    assert [o["op"] for o in prog.operations] == ["Vacuum", "Squeezed", "Sgate", "S2gate", "Interferometer", "MeasureHomodyne"]
    assert prog.operations[3] == {"op": "S2gate", "args": [0.543, -0.12], "kwargs": {}, "modes": [0, 3]}
    assert prog.operations[4]["args"][0].shape == (2, 2) and prog.operations[5]["kwargs"] == {"phi": 0.43, "select": 0.32}


@pytest.mark.parametrize("ir", ["blackbird", "xir"])
def test_save_load_round_trip(ir, tmp_path):
    prog = bio.loads(X8)
    path = tmp_path / "prog"
    bio.save(str(path), prog, ir=ir)
    written = str(path) + (".xbb" if ir == "blackbird" else ".xir")
    assert os.path.exists(written)            # the extension is appended (io/__init__.py:118-121)
    back = bio.load(written, ir=ir)
    assert back.name == prog.name and len(back.operations) == len(prog.operations)
    for a, b in zip(prog.operations, back.operations):
        assert a["op"] == b["op"] and a["modes"] == b["modes"] and a["kwargs"] == b["kwargs"]
        for x, y in zip(a["args"], b["args"]):
            assert np.allclose(x, y, atol=0, rtol=0)
    buf = _io.StringIO()
    bio.save(buf, prog, ir=ir)
    assert bio.loads(buf.getvalue(), ir=ir).num_subsystems == 8


@pytest.mark.parametrize("N", [2, 3, 4, 5, 8])
def test_clements_mesh_reproduces_the_unitary(N):
    rng = np.random.RandomState(N)
    A = rng.randn(N, N) + 1j * rng.randn(N, N)
    U, _ = np.linalg.qr(A)
    calls = bio.clements_rectangular(U)
    assert np.abs(W.interferometer_unitary(N, calls) - U).max() < 1e-13
    assert sum(c[0] == "beamsplitter" for c in calls) == N * (N - 1) // 2
    assert all(abs(c[-1] - c[-2]) == 1 for c in calls if c[0] == "beamsplitter")   # neighbouring modes only
    # the structure the reference emits: (R, BS) pairs, one R per mode, (BS, R) pairs
    kinds = [c[0][0] for c in calls]
    first = N * (N - 1) // 2 - sum(1 for i in range(1, N - 1, 2) for _ in range(i + 1))
    assert kinds == ["r", "b"] * first + ["r"] * N + ["b", "r"] * (N * (N - 1) // 2 - first)
    with pytest.raises(ValueError, match="not unitary"):
        bio.clements_rectangular(A)


def test_reference_compiled_mesh_instances_agree(golden_dir):
    """tests/golden/interferometer_n*.json hold gate lists the reference front end compiled for given
    unitaries: our mesh realises the same unitary with the same gate structure."""
    import json

    for N in (4, 5):
        path = os.path.join(golden_dir, "interferometer_n%d.json" % N)
        ref = json.load(open(path))
        if "U_re" not in ref:
            pytest.skip("fixture holds no unitary")
        U = np.array(ref["U_re"]) + 1j * np.array(ref["U_im"])
        calls = bio.clements_rectangular(U)
        assert np.abs(W.interferometer_unitary(N, calls) - U).max() < 1e-12


def test_program_runs_and_matches_the_oracle(host):
    """a loaded script on the plugin == the same calls on the oracle (decompositions included)"""
    from oracle.fock_oracle import OracleBackend
    from strawberryfields_b200.backend import B200FockBackend

    rng = np.random.RandomState(4)
    U, _ = np.linalg.qr(rng.randn(3, 3) + 1j * rng.randn(3, 3))
    rows = "\n".join("    " + ", ".join("%r%s%rj" % (float(z.real), "+" if z.imag >= 0 else "-", abs(float(z.imag))) for z in row) for row in U)
    script = """name mixed_bag
version 1.0
complex array U[3, 3] =
%s
Squeezed(0.3, 0.2) | 0
Coherent(0.4, 1.0) | 1
Interferometer(U) | [2, 0, 1]
Xgate(0.3) | 0
Zgate(-0.2) | 1
Pgate(0.4) | 2
CXgate(0.3) | [0, 1]
CZgate(-0.2) | [1, 2]
Fouriergate() | 0
Kgate(0.1) | 1
CKgate(0.2) | [0, 2]
Vgate(0.05) | 2
LossChannel(0.9) | 1
""" % rows
    prog = bio.loads(script)
    be = B200FockBackend()
    prog.run(be, cutoff_dim=5)
    ob = OracleBackend()
    ob.begin_circuit(3, cutoff_dim=5)
    for c in prog.calls():
        getattr(ob, c[0])(*c[1:])
    assert np.abs(be.state().dm() - ob.state().dm()).max() < 1e-12


def test_measurement_samples_are_returned(host):
    from strawberryfields_b200.backend import B200FockBackend

    prog = bio.loads("name m\nversion 1.0\nFock(2) | 0\nFock(1) | 1\nBSgate(0.0, 0.0) | [0, 1]\nMeasureFock() | [1, 0]\n")
    samples = prog.run(B200FockBackend(), cutoff_dim=4)
    assert samples == {1: [1], 0: [2]}
    with pytest.raises(ValueError, match="cutoff_dim"):
        prog.run(B200FockBackend())


@pytest.mark.parametrize("pure", [True, False])
def test_state_checkpoint_round_trip(host, pure, tmp_path):
    from strawberryfields_b200.backend import B200FockBackend

    be = B200FockBackend()
    be.begin_circuit(3, cutoff_dim=4, pure=pure)
    W.run_calls(be, W.config2_circuit(3, seed=2))
    st = be.state()
    path = str(tmp_path / "ckpt.npz")
    bio.save_state(path, st)
    b2 = B200FockBackend()
    b2.begin_circuit(3, cutoff_dim=4)
    data, was_pure = bio.load_state(path, b2)
    assert was_pure == pure and np.array_equal(data, st.data)
    assert b2.state().is_pure == pure and np.abs(b2.state().data - st.data).max() < 1e-15
    b3 = B200FockBackend()
    b3.begin_circuit(2, cutoff_dim=4)
    with pytest.raises(ValueError, match="checkpoint holds"):
        bio.load_state(path, b3)


@pytest.mark.reference
def test_lowering_equals_the_reference_front_end():
    """Differential: the reference front end (blackbird stubbed away, so the program is built with its own
    ops) compiles the same operations with its ``fock`` compiler; the backend call lists must produce the same
    state on the oracle."""
    from oracle import ref_shim
    from oracle.fock_oracle import OracleBackend

    sf = ref_shim.install()
    from strawberryfields import ops

    prog = bio.loads(X8.replace("S2gate(1.0", "S2gate(0.3").replace("| [0, 4]", "| [0, 2]").replace("| [1, 5]", "| [1, 3]")
                     .replace("S2gate(0.3, 0.0) | [3, 7]\n", "").split("Interferometer(U) | [4, 5, 6, 7]")[0])
    # (q[2], q[0]) would hit the reference's pure-state axis bug: keep the pair ascending
    assert prog.operations[3]["modes"] == [2, 0]
    prog.operations[3]["modes"] = [0, 2]
    # the script prints U to 8 digits; both sides get the same exactly unitary matrix
    prog.operations[2]["args"][0] = np.linalg.qr(prog.operations[2]["args"][0])[0]
    sfp = sf.Program(4)
    with sfp.context as q:
        for op in prog.operations:
            getattr(ops, op["op"])(*op["args"], **op["kwargs"]) | [q[i] for i in op["modes"]]
    eng = sf.Engine("fock", backend_options={"cutoff_dim": 5})
    want = eng.run(sfp).state
    ob = OracleBackend()
    ob.begin_circuit(4, cutoff_dim=5)
    for c in prog.calls():
        getattr(ob, c[0])(*c[1:])
    got = ob.state()
    assert np.abs(got.data - want.data).max() < 1e-12


# ------------------------------------------------------------------ templates and measured parameters
TEMPLATE = """name tmpl
version 1.0
target fock (cutoff_dim=5)
float a = 0.3
Sgate({r}, 0.1) | 0
Dgate(a*{alpha}**2 + 0.1, phi=sqrt({r})) | 1
BSgate({theta}, pi/2 - {theta}) | [0, 1]
"""


def test_template_parameters_stay_symbolic_until_bound():
    prog = bio.loads(TEMPLATE)
    assert prog.is_template and prog.free_parameters == ["alpha", "r", "theta"] and not prog.has_feed_forward
    assert isinstance(prog.operations[0]["args"][0], bio.Parameter) and prog.operations[0]["args"][1] == 0.1
    # written back as Blackbird and loaded again: the same script
    assert bio.loads(prog.serialize()).serialize() == prog.serialize()
    assert "{alpha}" in prog.serialize() and "sqrt({r})" in prog.serialize()
    # XIR has no {name} syntax: free parameters are bare names there (io.to_xir writes them so, xir_io.py:287-303)
    xprog = bio.loads(prog.serialize("xir"), ir="xir")
    assert "Sgate(r, 0.1) | [0];" in prog.serialize("xir") and xprog.free_parameters == prog.free_parameters
    with pytest.raises(ValueError, match="alpha"):
        prog.calls(args={"r": 0.2, "theta": 0.1})
    vals = {"r": 0.25, "alpha": 0.5, "theta": 0.4}
    calls = prog.calls(args=vals)
    assert calls[0] == ("squeeze", 0.25, 0.1, 0)
    assert calls[1][0] == "displacement" and abs(calls[1][1] - (0.3 * 0.25 + 0.1)) < 1e-15 and calls[1][2] == 0.5
    assert calls[2] == ("beamsplitter", 0.4, np.pi / 2 - 0.4, 0, 1)
    # partial binding keeps the rest symbolic and rewrites the text
    part = prog.bind(r=0.25)
    assert part.free_parameters == ["alpha", "theta"] and part.operations[0]["args"][0] == 0.25
    assert part.operations[1]["kwargs"]["phi"] == 0.5 and "{r}" not in part.serialize()
    assert part.bind(alpha=0.5, theta=0.4).calls() == calls
    assert xprog.calls(args=vals) == calls
    with pytest.raises(NotImplementedError, match="measured parameters"):
        bio.loads("MeasureFock() | 0\nRgate(q0) | 1").serialize("xir")
    with pytest.raises(ValueError, match="no free parameter"):
        prog.bind(gamma=1.0)


def test_parameter_arithmetic_with_numpy_scalars():
    x = bio.Parameter.free("x")
    e = np.float64(2.0) * x - (1 + 2j) / x ** 2 + (-x)
    assert e.free_names == ["x"] and e.measured_modes == []
    assert abs(e.evaluate(lambda kind, key: 0.5) - (1.0 - (1 + 2j) / 0.25 - 0.5)) < 1e-15
    assert e.substitute({"x": 0.5}) == e.evaluate(lambda kind, key: 0.5)
    m = bio.Parameter.measured(3) * x
    assert m.measured_modes == [3] and m.substitute({"x": 2}).text == "q3 * 2"
    assert bio.loads("Rgate(%s) | 0" % e.text).operations[0]["args"][0].evaluate(lambda k, n: 0.5) == e.evaluate(lambda k, n: 0.5)


def test_template_runs_with_args(host):
    from oracle.fock_oracle import OracleBackend
    from strawberryfields_b200.backend import B200FockBackend

    prog = bio.loads(TEMPLATE)
    vals = {"r": 0.25, "alpha": 0.5, "theta": 0.4}
    be = B200FockBackend()
    with pytest.raises(ValueError, match="alpha, r, theta"):
        prog.run(be)
    prog.run(be, args=vals)          # cutoff from the target line
    ob = OracleBackend()
    ob.begin_circuit(2, cutoff_dim=5)
    ob.squeeze(0.25, 0.1, 0)
    ob.displacement(0.3 * 0.25 + 0.1, 0.5, 1)
    ob.beamsplitter(0.4, np.pi / 2 - 0.4, 0, 1)
    assert np.abs(be.state().ket() - ob.state().ket()).max() < 1e-12


FEED_FORWARD = """name ff
version 1.0
Fock(2) | 0
Squeezed(0.6, 0.0) | 1
BSgate(0.7, 0.3) | [0, 1]
MeasureFock() | 0
Rgate(q0*pi/3) | 1
Dgate(0.1*q0 + 0.05, 0.0) | 2
MeasureFock() | 1
Rgate(q0 - q1) | 2
"""


@pytest.mark.parametrize("seed", [3, 11, 12])
def test_measured_parameters_are_fed_forward(host, seed):
    """``q<m>`` takes the latest outcome of mode m when its operation is reached (engine.py:427-444)."""
    from oracle.fock_oracle import OracleBackend
    from strawberryfields_b200.backend import B200FockBackend

    prog = bio.loads(FEED_FORWARD)
    assert prog.has_feed_forward and not prog.is_template
    with pytest.raises(ValueError, match="measured parameters"):
        prog.calls()
    be = B200FockBackend()
    np.random.seed(seed)
    samples = prog.run(be, cutoff_dim=6)
    ob = OracleBackend()
    ob.begin_circuit(3, cutoff_dim=6)
    np.random.seed(seed)
    ob.prepare_fock_state(2, 0)
    ob.prepare_squeezed_state(0.6, 0.0, 1)
    ob.beamsplitter(0.7, 0.3, 0, 1)
    n0 = int(np.asarray(ob.measure_fock([0])).reshape(-1)[0])
    ob.rotation(n0 * np.pi / 3, 1)
    ob.displacement(0.1 * n0 + 0.05, 0.0, 2)
    n1 = int(np.asarray(ob.measure_fock([1])).reshape(-1)[0])
    ob.rotation(n0 - n1, 2)
    assert samples == {0: [n0], 1: [n1]}
    assert np.abs(be.state().dm() - ob.state().dm()).max() < 1e-12


def test_measured_parameter_before_its_measurement(host):
    from strawberryfields_b200.backend import B200FockBackend

    with pytest.raises(ValueError, match="before mode 1 has been measured"):
        bio.loads("Rgate(q1) | 0\nMeasureFock() | 1").run(B200FockBackend(), cutoff_dim=3)


@pytest.mark.reference
def test_symbolic_programs_run_like_the_reference_engine(host):
    """``to_sf`` hands the reference FreeParameter / MeasuredParameter atoms (par_convert, parameters.py:249-275);
    its engine binds ``args`` and feeds outcomes forward -- the result must equal ``CircuitProgram.run``."""
    from oracle import ref_shim
    from strawberryfields_b200.backend import B200FockBackend

    sf = ref_shim.install()
    script = TEMPLATE + "MeasureFock() | 0\nRgate(q0*{theta} + sin(q0)) | 1\n"
    prog = bio.loads(script)
    vals = {"r": 0.6, "alpha": 0.5, "theta": 0.4}
    sfp = prog.to_sf()
    assert sorted(sfp.free_params) == prog.free_parameters
    eng = sf.Engine("fock", backend_options={"cutoff_dim": 5})
    for seed in (1, 5, 6):
        np.random.seed(seed)
        res = eng.run(sfp, args=vals)
        be = B200FockBackend()
        np.random.seed(seed)
        samples = prog.run(be, args=vals)
        assert samples[0] == [int(np.asarray(res.samples).reshape(-1)[0])]
        assert np.abs(be.state().dm() - res.state.dm()).max() < 1e-12
        eng.reset()


@pytest.mark.reference
def test_catstate_gkp_smzgate_lower_like_the_reference():
    """the remaining Fock-compiler primitives / decompositions a script can name (compilers/fock.py:23-71)"""
    from oracle import ref_shim
    from oracle.fock_oracle import OracleBackend

    sf = ref_shim.install()
    prog = bio.loads("name extras\nversion 1.0\ntarget fock (cutoff_dim=8)\n"
                     "Catstate(0.8, 0.3, 1) | 0\nGKP([0.5, 0.2], 0.35) | 1\nsMZgate(0.3, 0.9) | [0, 1]\n"
                     "sMZgate(pi/2, 0.2) | [1, 0]\n")
    # (q[1], q[0]) would hit the reference's pure-state axis bug: compare on ascending pairs only
    prog.operations[3]["modes"] = [0, 1]
    from strawberryfields import ops

    with pytest.raises(NameError, match="sMZgate"):      # not in ops.__all__: the reference loader refuses it too
        prog.to_sf()
    sfp = sf.Program(2)
    with sfp.context as q:
        for op in prog.operations:
            getattr(ops, op["op"])(*op["args"], **op["kwargs"]) | [q[i] for i in op["modes"]]
    want = sf.Engine("fock", backend_options={"cutoff_dim": 8}).run(sfp).state
    ob = OracleBackend()
    ob.begin_circuit(2, cutoff_dim=8)
    calls = prog.calls()
    assert [c[0] for c in calls].count("rotation") == 3      # Rgate(pi/2 - pi/2) is skipped
    for c in calls:
        getattr(ob, c[0])(*c[1:-1], **c[-1]) if isinstance(c[-1], dict) else getattr(ob, c[0])(*c[1:])
    assert np.abs(ob.state().dm() - want.dm()).max() < 1e-12
    with pytest.raises(ValueError, match="cutoff"):
        bio.loads("Catstate(0.8) | 0").calls()


# ------------------------------------------------------------------ the reference's own script tests as golden vectors
# (tests/frontend/io/test_io_blackbird.py:73-92,386-580 and test_io_xir.py:68-79,470-600: the scripts and what the
# reference asserts about the converted programs; here checked on the operation list of the loader)
REF_U = np.array([
    [0.219546940711 - 0.256534554457j, 0.611076853957 + 0.524178937791j, -0.102700187435 + 0.474478834685j, -0.027250232925 + 0.03729094623j],
    [0.451281863394 + 0.602582912475j, 0.456952590016 + 0.01230749109j, 0.131625867435 - 0.450417744715j, 0.035283194078 - 0.053244267184j],
    [0.038710094355 + 0.492715562066j, -0.019212744068 - 0.321842852355j, -0.240776471286 + 0.524432833034j, -0.458388143039 + 0.329633367819j],
    [-0.156619083736 + 0.224568570065j, 0.109992223305 - 0.163750223027j, -0.421179844245 + 0.183644837982j, 0.818769184612 + 0.068015658737j]])

REF_BLACKBIRD = """\
name test_program
version 1.0

complex array A0[4, 4] =
    0.219546940711-0.256534554457j, 0.611076853957+0.524178937791j, -0.102700187435+0.474478834685j, -0.027250232925+0.03729094623j
    0.451281863394+0.602582912475j, 0.456952590016+0.01230749109j, 0.131625867435-0.450417744715j, 0.035283194078-0.053244267184j
    0.038710094355+0.492715562066j, -0.019212744068-0.321842852355j, -0.240776471286+0.524432833034j, -0.458388143039+0.329633367819j
    -0.156619083736+0.224568570065j, 0.109992223305-0.163750223027j, -0.421179844245+0.183644837982j, 0.818769184612+0.068015658737j

Vacuum() | 1
Squeezed(0.12, 0.0) | 2
Sgate(1, 0.0) | 0
Dgate(0.735934779718964, 0.7469555733762603) | 1
S2gate(0.543, -0.12) | [0, 3]
Interferometer(A0) | [0, 1, 2, 3]
MeasureHomodyne(0) | 0
MeasureHomodyne(0.43, select=0.32) | 2
MeasureHomodyne(0.43, select=0.32) | 2
"""

REF_XIR = """\
Vacuum | [1];
Squeezed(0.12, 0.0) | [2];
Sgate(1, 0.0) | [0];
Dgate(0.735934779718964, 0.7469555733762603) | [1];
S2gate(0.543, -0.12) | [0, 3];
Interferometer([[(0.219546940711-0.256534554457j), (0.611076853957+0.524178937791j), (-0.102700187435+0.474478834685j), (-0.027250232925+0.03729094623j)], [(0.451281863394+0.602582912475j), (0.456952590016+0.01230749109j), (0.131625867435-0.450417744715j), (0.035283194078-0.053244267184j)], [(0.038710094355+0.492715562066j), (-0.019212744068-0.321842852355j), (-0.240776471286+0.524432833034j), (-0.458388143039+0.329633367819j)], [(-0.156619083736+0.224568570065j), (0.109992223305-0.163750223027j), (-0.421179844245+0.183644837982j), (0.818769184612+0.068015658737j)]]) | [0, 1, 2, 3];
MeasureHomodyne(phi: 0) | [0];
MeasureHomodyne(phi: 0.43, select: 0.32) | [2];
MeasureHomodyne(phi: 0.43, select: 0.32) | [2];\
"""


@pytest.mark.parametrize("ir,script", [("blackbird", REF_BLACKBIRD), ("xir", REF_XIR)], ids=["blackbird", "xir"])
def test_reference_not_compiled_program_scripts(ir, script):
    """the program of the reference's ``prog`` fixture (test_io_xir.py:46-66) as its two serialisations"""
    prog = bio.loads(script, ir=ir)
    assert [op["op"] for op in prog.operations] == ["Vacuum", "Squeezed", "Sgate", "Dgate", "S2gate", "Interferometer",
                                                    "MeasureHomodyne", "MeasureHomodyne", "MeasureHomodyne"]
    assert [op["modes"] for op in prog.operations] == [[1], [2], [0], [1], [0, 3], [0, 1, 2, 3], [0], [2], [2]]
    assert prog.operations[1]["args"] == [0.12, 0.0] and prog.operations[2]["args"] == [1, 0.0]
    assert np.allclose(prog.operations[3]["args"], [abs(0.54 + 0.5j), np.angle(0.54 + 0.5j)], rtol=0, atol=1e-15)
    assert prog.operations[4]["args"] == [0.543, -0.12]
    assert np.array_equal(np.asarray(prog.operations[5]["args"][0]), REF_U)
    last = prog.operations[8]
    assert (last["args"] + [last["kwargs"].get("phi")])[0] == 0.43 and last["kwargs"]["select"] == 0.32
    assert prog.num_subsystems == 4
    # both serialisations lower to the same backend calls
    other = bio.loads(REF_XIR if ir == "blackbird" else REF_BLACKBIRD, ir="xir" if ir == "blackbird" else "blackbird")
    for a, b in zip(prog.calls(), other.calls()):
        assert a[0] == b[0] and all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))


def test_reference_blackbird_conversion_cases():
    indent = lambda text: "\n".join("        " + line for line in text.splitlines())  # noqa: E731  (the reference's are indented)
    with pytest.raises(ValueError, match="contains no quantum operations"):
        bio.loads(indent("name test_program\nversion 1.0\n"))
    prog = bio.loads(indent("name test_program\nversion 1.0\nVac | 0\n"))
    assert prog.name == "test_program" and prog.operations == [{"op": "Vac", "args": [], "kwargs": {}, "modes": [0]}]
    assert prog.calls() == [("prepare_vacuum_state", 0)]
    prog = bio.loads(indent("name test_program\nversion 1.0\n\nDgate(r=0.54, phi=0) | 0\n"))
    assert prog.operations[0]["kwargs"] == {"r": 0.54, "phi": 0} and prog.calls() == [("displacement", 0.54, 0.0, 0)]
    prog = bio.loads(indent("name test_program\nversion 1.0\n\nBSgate(theta=0.54, phi=pi) | [0, 2]\n"))
    assert prog.calls() == [("beamsplitter", 0.54, np.pi, 0, 2)]
    # test_gate_measured_par
    prog = bio.loads(indent("name test_program\nversion 1.0\n\nMeasureX | 0\nDgate(q0) | 1\nRgate(2*q0) | 2\n"))
    p1, p2 = prog.operations[1]["args"][0], prog.operations[2]["args"][0]
    assert p1.node == ("measured", 0) and p2.measured_modes == [0] and p2.evaluate(lambda k, m: 0.25) == 0.5
    # test_gate_free_par
    prog = bio.loads(indent("name test_program\nversion 1.0\n\n"
                            "Dgate(1-{ALPHA}, 0) | 0     # keyword arg, compound expr\n"
                            "Rgate(theta={foo_bar1}) | 0  # keyword arg, atomic\n"
                            "Dgate({ALPHA}**2, 0) | 0        # positional arg, compound expr\n"
                            "Rgate({foo_bar2}) | 0        # positional arg, atomic\n"))
    assert set(prog.free_parameters) == {"foo_bar1", "foo_bar2", "ALPHA"} and len(prog.operations) == 4
    assert prog.operations[1]["kwargs"]["theta"].node == ("free", "foo_bar1")
    assert prog.operations[3]["args"][0].node == ("free", "foo_bar2")
    assert prog.calls(args={"ALPHA": 0.5, "foo_bar1": 0.1, "foo_bar2": 0.2}) == [
        ("displacement", 0.5, 0.0, 0), ("rotation", 0.1, 0), ("displacement", 0.25, 0.0, 0), ("rotation", 0.2, 0)]


def test_reference_xir_gate_definitions():
    """test_io_xir.py:561-600: user-defined gates are expanded in place, wires default to the body's labels"""
    script = """
        gate Aubergine(x, y)[w]:
            Squeezed(x, y) | [w];
        end;

        gate Banana(a, b, c, d, x, y):
            Aubergine(x, y) | [0];
            Aubergine(x, y) | [1];
            Rgate(a) | [0];
            BSgate(b, c) | [0, 1];
            Rgate(d) | [1];
        end;

        Vacuum | [1];
        Banana(0.5, 0.4, 0.0, 0.5, 1.0, 0.0) | [3, 0];
    """
    prog = bio.loads(script, ir="xir")
    assert [op["op"] for op in prog.operations] == ["Vacuum", "Squeezed", "Squeezed", "Rgate", "BSgate", "Rgate"]
    assert [op["args"] for op in prog.operations] == [[], [1.0, 0.0], [1.0, 0.0], [0.5], [0.4, 0.0], [0.5]]
    assert [op["modes"] for op in prog.operations] == [[1], [3], [0], [3], [3, 0], [0]]
    with pytest.raises(bio.ProgramSyntaxError, match="acts on 2 wire"):
        bio.loads(script.replace("| [3, 0];", "| [3];"), ir="xir")
    with pytest.raises(bio.ProgramSyntaxError, match="takes parameters"):
        bio.loads(script.replace("Banana(0.5, 0.4, 0.0, 0.5, 1.0, 0.0)", "Banana(0.5)"), ir="xir")
    with pytest.raises(ValueError, match="XIR program is empty"):
        bio.loads("options:\n  cutoff_dim: 5;\nend;\n", ir="xir")
    with pytest.raises(NotImplementedError, match="tdm"):
        bio.loads("options:\n  _type_: tdm;\n  N: [2, 3];\nend;\nSgate(0.1, 0.0) | [2];", ir="xir")


def test_serialisers_reproduce_the_reference_text(tmp_path):
    """what ``sf.save`` must write for the reference's ``prog`` fixture (test_io_blackbird.py:736-776,
    test_io_xir.py:658-700), file-name handling included"""
    assert bio.loads(REF_BLACKBIRD).serialize() == REF_BLACKBIRD
    assert bio.loads(REF_XIR, ir="xir").serialize("xir") == REF_XIR
    prog = bio.loads(REF_BLACKBIRD)
    bio.save(tmp_path / "test.xbb", prog)                       # path object
    assert (tmp_path / "test.xbb").read_text() == REF_BLACKBIRD
    bio.save(str(tmp_path / "test.txt"), prog, ir="xir")        # extension appended
    xprog = bio.load(str(tmp_path / "test.txt.xir"), ir="xir")
    assert [op["op"] for op in xprog.operations] == [op["op"] for op in prog.operations]
    with open(tmp_path / "obj.xbb", "w") as f:                  # file object
        bio.save(f, prog)
    with open(tmp_path / "obj.xbb") as f:
        assert bio.load(f).serialize() == REF_BLACKBIRD
    for fn in (bio.load, lambda x: bio.save(x, prog)):
        with pytest.raises(ValueError, match="must be a string, path"):
            fn(1)


def test_batched_state_checkpoint_round_trip(host, tmp_path):
    from strawberryfields_b200.backend import B200FockBackend

    B, D, n = 3, 4, 2
    for pure in (True, False):
        be = B200FockBackend()
        be.begin_circuit(n, cutoff_dim=D, batch_size=B, pure=pure)
        be.prepare_coherent_state(np.array([0.1, 0.3, 0.5]), 0.2, 0)
        be.squeeze(0.2, 0.0, 1)
        be.beamsplitter(np.array([0.3, 0.6, 0.9]), 0.1, 0, 1)
        st = be.state()
        path = str(tmp_path / ("b%d.npz" % pure))
        bio.save_state(path, st)
        b2 = B200FockBackend()
        b2.begin_circuit(n, cutoff_dim=D, batch_size=B)
        data, was_pure = bio.load_state(path, b2)
        assert was_pure == st.is_pure and data.shape[0] == B
        assert b2.state().is_pure == st.is_pure and np.abs(b2.state().data - st.data).max() < 1e-15


@pytest.mark.reference
def test_from_sf_writes_what_the_reference_saves():
    """``from_sf`` == ``io.to_blackbird`` (blackbird_io.py:164-232) on the reference's own cases
    (test_io_blackbird.py:43-70,98-350): the ``prog`` fixture must serialise to the text ``sf.save`` writes."""
    from oracle import ref_shim

    sf = ref_shim.install()
    from strawberryfields import ops
    from strawberryfields.parameters import par_funcs as pf

    prog = sf.Program(4, name="test_program")
    with prog.context as q:
        ops.Vac | q[1]
        ops.Squeezed(0.12) | q[2]
        ops.Sgate(1) | q[0]
        ops.Dgate(np.abs(0.54 + 0.5j), np.angle(0.54 + 0.5j)) | q[1]
        ops.S2gate(0.543, -0.12) | (q[0], q[3])
        ops.Interferometer(REF_U) | q
        ops.MeasureX | q[0]
        ops.MeasureHomodyne(0.43, select=0.32) | q[2]
        ops.MeasureHomodyne(phi=0.43, select=0.32) | q[2]
    cp = bio.from_sf(prog)
    assert cp.serialize() == REF_BLACKBIRD
    assert cp.target["name"] is None and cp.version == "1.0"
    # measurement keyword arguments
    prog = sf.Program(1)
    with prog.context as q:
        ops.MeasureFock(select=2) | q[0]
    assert bio.from_sf(prog).operations[0] == {"op": "MeasureFock", "modes": [0], "args": [], "kwargs": {"select": [2]}}
    prog = sf.Program(1)
    with prog.context as q:
        ops.MeasureFock(dark_counts=2) | q[0]
    assert bio.from_sf(prog).operations[0]["kwargs"] == {"dark_counts": [2]}
    # symbolic parameters (test_measured_par_str / test_free_par_str)
    prog = sf.Program(2)
    with prog.context as q:
        ops.Sgate(0.43) | q[0]
        ops.MeasureX | q[0]
        ops.Zgate(2 * pf.sin(q[0].par)) | q[1]
    last = bio.from_sf(prog).operations[-1]
    assert last["op"] == "Zgate" and last["modes"] == [1] and last["args"][0].measured_modes == [0]
    assert abs(last["args"][0].evaluate(lambda k, m: 0.3) - 2 * np.sin(0.3)) < 1e-15
    prog = sf.Program(2)
    r, alpha = prog.params("r", "alpha")
    with prog.context as q:
        ops.Sgate(r) | q[0]
        ops.Zgate(3 * pf.log(-alpha)) | q[1]
    cp = bio.from_sf(prog)
    assert cp.operations[0]["args"][0].node == ("free", "r") and cp.operations[0]["args"][1] == 0.0
    assert cp.free_parameters == ["alpha", "r"]
    assert abs(cp.operations[1]["args"][0].evaluate(lambda k, n: -0.7) - 3 * np.log(0.7)) < 1e-15
    # .. and the text loads back to the same expression
    again = bio.loads(cp.serialize())
    assert abs(again.operations[1]["args"][0].evaluate(lambda k, n: -0.7) - 3 * np.log(0.7)) < 1e-15
    # a compiled program carries its target and options
    prog = sf.Program(2, name="c")
    with prog.context as q:
        ops.Pgate(0.43) | q[0]
        ops.BSgate(0.3, 0.1) | (q[0], q[1])
    comp = bio.from_sf(prog.compile(compiler="fock", shots=7))
    assert comp.target == {"name": "fock", "options": {"shots": 7}}
    assert [op["op"] for op in comp.operations] == ["Sgate", "Rgate", "BSgate"]


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("ir", ["blackbird", "xir"])
def test_random_programs_round_trip(seed, ir):
    """write -> read -> write is a fixed point and keeps every number bit for bit (repr round trip), for random
    programs over the whole operation set incl. arrays, keyword arguments, free parameters and expressions"""
    rs = np.random.RandomState(100 + seed)
    n = int(rs.randint(2, 6))
    free = [bio.Parameter.free(nm) for nm in ("alpha", "beta_1", "G")]

    def number():
        kind = rs.randint(5)
        if kind == 0:
            return int(rs.randint(-3, 4))
        if kind == 1:
            return float(rs.randn() * 10.0 ** rs.randint(-8, 8))
        if kind == 2:
            p = free[rs.randint(3)]
            return [p, 2 * p - 0.5, p ** 2 / 3, -p, bio.Parameter.call("sin", [p]) * 1.5][rs.randint(5)]
        return float(rs.uniform(-1, 1))

    ops1 = ["Sgate", "Dgate", "Rgate", "Kgate", "Vgate", "Xgate", "Zgate", "Pgate", "LossChannel", "Coherent", "Squeezed"]
    ops2 = ["BSgate", "S2gate", "MZgate", "CKgate", "CXgate", "CZgate"]
    prog = bio.CircuitProgram(name="fuzz_%d" % seed, target={"name": "fock", "options": {"cutoff_dim": 5, "shots": 3}}
                              if ir == "blackbird" else None, options={"cutoff_dim": 5} if ir == "xir" else None)
    for _ in range(int(rs.randint(5, 25))):
        r = rs.rand()
        if r < 0.5:
            nargs = int(rs.randint(1, 3))
            prog.operations.append({"op": ops1[rs.randint(len(ops1))], "args": [number() for _ in range(nargs)], "kwargs": {},
                                    "modes": [int(rs.randint(n))]})
        elif r < 0.85:
            a, b = rs.choice(n, 2, replace=False)
            kw = {"phi": number()} if rs.rand() < 0.3 else {}
            prog.operations.append({"op": ops2[rs.randint(len(ops2))], "args": [number()], "kwargs": kw, "modes": [int(a), int(b)]})
        elif r < 0.93:
            k = int(rs.randint(2, n + 1))
            U = np.linalg.qr(rs.randn(k, k) + 1j * rs.randn(k, k))[0]
            prog.operations.append({"op": "Interferometer", "args": [U], "kwargs": {}, "modes": [int(x) for x in rs.permutation(n)[:k]]})
        else:
            prog.operations.append({"op": "MeasureFock", "args": [], "kwargs": {"select": [int(rs.randint(3))]} if rs.rand() < 0.5 else {},
                                    "modes": [int(rs.randint(n))]})
    text = bio.dumps(prog, ir)
    again = bio.loads(text, ir)
    assert bio.dumps(again, ir) == text
    assert len(again.operations) == len(prog.operations) and again.free_parameters == prog.free_parameters
    look = lambda kind, key: {"alpha": 0.3, "beta_1": -1.2, "G": 2.5}[key]  # noqa: E731
    for a, b in zip(prog.operations, again.operations):
        assert a["op"] == b["op"] and a["modes"] == b["modes"] and sorted(a["kwargs"]) == sorted(b["kwargs"])
        for x, y in zip(a["args"] + [a["kwargs"][k] for k in sorted(a["kwargs"])],
                        b["args"] + [b["kwargs"][k] for k in sorted(b["kwargs"])]):
            x, y = bio._resolve(x, look), bio._resolve(y, look)
            if isinstance(x, np.ndarray):
                assert np.array_equal(x, np.asarray(y))
            elif isinstance(x, float):
                assert np.isclose(x, y, rtol=1e-15, atol=0)     # expressions re-associate constants at most
            else:
                assert x == y
