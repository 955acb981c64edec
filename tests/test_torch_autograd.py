"""Differentiable circuits (strawberryfields_b200/autodiff.py).

* forward: the ket equals the oracle's (1e-12);
* derivative tables: d(table)/d(parameter) against central finite differences of the oracle's gate
  tensors (the restated thewalrus recursions), every gate, both parameters;
* backward: gradients of real losses against central finite differences of the same circuit
  (step 1e-6, agreement 1e-7 -- the truncation error of the difference, not of the gradient),
  scalar and per-batch-entry parameters, a parameter used twice, constants mixed in.
``host``: numpy double of the C ABI; ``gpu``: the CUDA kernels."""
import numpy as np
import pytest
import torch

from fake_lib import FakeLib
from oracle import gates as og
from oracle.fock_oracle import OracleBackend

TOL = 1e-12


@pytest.fixture(params=["host", pytest.param("gpu", marks=pytest.mark.gpu)])
def make(request, monkeypatch):
    from strawberryfields_b200 import circuit, lib
    from strawberryfields_b200.autodiff import TorchCircuit

    if request.param == "host":
        monkeypatch.setattr(lib, "_lib", FakeLib())
        monkeypatch.setattr(circuit, "_TEST_HOST_MODE", True)
    return TorchCircuit


GATES = [
    ("displacement", (0.31, 0.7), 1), ("squeeze", (0.22, -0.4), 1), ("rotation", (0.9,), 1),
    ("kerr_interaction", (0.35,), 1), ("cross_kerr_interaction", (0.6,), 2),
    ("beamsplitter", (0.45, 1.1), 2), ("mzgate", (0.8, -0.3), 2), ("two_mode_squeeze", (0.17, 0.6), 2),
]


def oracle_table(name, params, D):
    """dense gate tensor: [out, in] or [o1, i1, o2, i2] (the layout fockbackend/ops.py returns)"""
    n = np.arange(D)
    if name == "displacement":
        return og.displacement(params[0], params[1], D)
    if name == "squeeze":
        return og.squeezing(params[0], params[1], D)
    if name == "rotation":
        return np.diag(np.exp(1j * params[0] * n))
    if name == "kerr_interaction":
        return np.diag(np.exp(1j * params[0] * n ** 2))
    if name == "cross_kerr_interaction":
        T = np.zeros([D] * 4, dtype=complex)
        for a in range(D):
            for b in range(D):
                T[a, a, b, b] = np.exp(1j * params[0] * a * b)
        return T
    fn = {"beamsplitter": og.beamsplitter, "mzgate": og.mzgate, "two_mode_squeeze": og.two_mode_squeeze}[name]
    return fn(params[0], params[1], D)  # already [out1, in1, out2, in2]


@pytest.mark.parametrize("name,params,nmodes", GATES, ids=[g[0] for g in GATES])
def test_derivative_tables(name, params, nmodes, make):
    from strawberryfields_b200 import lib as L
    from strawberryfields_b200.autodiff import _GATES, pair_index

    D = 6
    prog = make(2, D)
    cls, _, rule, npar, _ = _GATES[name]
    p = torch.zeros(2, 1, dtype=torch.float64, device=prog.device)
    for j, v in enumerate(params):
        p[j, 0] = v
    dtabs = prog._derivative_tables(name, p)
    assert len(dtabs) >= npar
    h = 1e-6
    for j in range(npar):
        hi, lo = list(params), list(params)
        hi[j] += h
        lo[j] -= h
        want = (oracle_table(name, hi, D) - oracle_table(name, lo, D)) / (2 * h)
        got = dtabs[j][0].cpu().numpy()
        if cls == "dense":
            got = got.reshape(D, D)
        elif cls == "diag":
            got = np.diag(got)
        elif cls == "diag2":
            full = np.zeros([D] * 4, dtype=complex)
            for a in range(D):
                for b in range(D):
                    full[a, a, b, b] = got.reshape(D, D)[a, b]
            got = full
        else:
            flat, _ = pair_index(rule, D, prog.device)
            full = np.zeros(D ** 4, dtype=complex)
            full[flat.cpu().numpy()] = got
            got = full.reshape([D] * 4)
        assert np.abs(got - want).max() < 1e-8, (name, j)


def build(prog, th):
    """3 modes, every differentiable gate, parameter th[0] used twice, one constant parameter."""
    prog.squeeze(th[0], th[1], 0)
    prog.displacement(th[2], th[3], 1)
    prog.displacement(0.2, th[4], 2)
    prog.beamsplitter(th[5], th[6], 0, 1)
    prog.rotation(th[7], 1)
    prog.kerr_interaction(th[8], 2)
    prog.two_mode_squeeze(th[9], th[10], 2, 1)
    prog.mzgate(th[11], th[12], 2, 0)
    prog.cross_kerr_interaction(th[13], 0, 2)
    prog.beamsplitter(th[0], 0.3, 1, 2)
    prog.squeeze(th[14], 0.0, 2)
    return prog


def oracle_ket(th, n, D):
    ob = OracleBackend()
    ob.begin_circuit(n, cutoff_dim=D)
    build(ob, [float(t) for t in th])
    return ob.state().data


def loss_of(ket):
    """a real loss that sees amplitudes and phases: weighted probabilities + overlap with a fixed vector"""
    n = ket.numel()
    w = torch.linspace(0.0, 1.0, n, dtype=torch.float64, device=ket.device)
    v = torch.exp(1j * torch.arange(n, dtype=torch.float64, device=ket.device) * 0.37) / np.sqrt(n)
    flat = ket.reshape(-1)
    return (w * (flat.abs() ** 2)).sum() + (v.conj() * flat).sum().real + ((v * flat).sum().imag) ** 2


def test_forward_and_gradient_match_finite_differences(make):
    n, D = 3, 5
    rng = np.random.RandomState(3)
    th0 = rng.uniform(0.05, 0.4, 15)
    th = [torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in th0]
    ket = build(make(n, D), th).ket()
    assert np.abs(ket.detach().cpu().numpy() - oracle_ket(th0, n, D)).max() < TOL
    loss_of(ket).backward()
    got = np.array([t.grad.item() for t in th])
    h = 1e-6
    for j in range(len(th0)):
        vals = []
        for s in (+1, -1):
            x = th0.copy()
            x[j] += s * h
            with torch.no_grad():
                vals.append(loss_of(build(make(n, D), [torch.tensor(v, dtype=torch.float64) for v in x]).ket()).item())
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(got[j] - fd) < 1e-7, (j, got[j], fd)


def test_batched_parameters(make):
    """per-entry parameters ([B]) get per-entry gradients; a scalar shared by the batch gets the sum"""
    n, D, B = 2, 5, 3
    r = torch.tensor([0.1, 0.2, 0.3], dtype=torch.float64, requires_grad=True)
    theta = torch.tensor(0.4, dtype=torch.float64, requires_grad=True)

    def run(rv, tv, batch):
        prog = make(n, D, batch_size=batch)
        prog.displacement(rv, 0.5, 0)
        prog.squeeze(0.15, tv, 1)
        prog.beamsplitter(tv, 0.2, 0, 1)
        prog.kerr_interaction(rv, 1)
        return prog.ket()

    ket = run(r, theta, B)
    assert ket.shape == (B, D, D)
    loss = sum(loss_of(ket[b]) * (b + 1) for b in range(B))
    loss.backward()
    h = 1e-6
    for b in range(B):
        fd_r, fd_t = [], []
        for s in (+1, -1):
            with torch.no_grad():
                k1 = run(torch.tensor(r[b].item() + s * h, dtype=torch.float64), theta.detach(), None)
                fd_r.append(loss_of(k1).item() * (b + 1))
        assert abs(r.grad[b].item() - (fd_r[0] - fd_r[1]) / (2 * h)) < 1e-7
    tot = []
    for s in (+1, -1):
        with torch.no_grad():
            k = run(r.detach(), torch.tensor(theta.item() + s * h, dtype=torch.float64), B)
            tot.append(sum(loss_of(k[b]).item() * (b + 1) for b in range(B)))
    assert abs(theta.grad.item() - (tot[0] - tot[1]) / (2 * h)) < 1e-7


def test_training_step_reduces_loss(make):
    """A few gradient steps on a 2-mode circuit raise the fidelity with a target Fock state -- the
    shape of examples/quantum_neural_network.py (one layer: BS, R, S, D, K)."""
    n, D = 2, 6
    rng = np.random.RandomState(0)
    w = torch.tensor(rng.normal(0, 0.1, 9), dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([w], lr=0.05)

    def fidelity():
        prog = make(n, D)
        prog.beamsplitter(w[0], w[1], 0, 1)
        prog.rotation(w[2], 0)
        for m in range(n):
            prog.squeeze(w[3 + m], 0.0, m)
            prog.displacement(w[5 + m], 0.0, m)
            prog.kerr_interaction(w[7 + m], m)
        return prog.ket()[1, 0].abs() ** 2

    first = fidelity().item()
    for _ in range(15):
        opt.zero_grad()
        loss = 1 - fidelity()
        loss.backward()
        opt.step()
    assert fidelity().item() > first + 0.05


def test_qnn_layers_batched_weights(make):
    """BASELINE config 4's shape at reduced size: QNN layers (examples/quantum_neural_network.py:14-85)
    with per-entry weights; forward equals the oracle entry by entry, gradients equal finite differences."""
    from strawberryfields_b200.autodiff import qnn_init_weights, qnn_layer, qnn_layer_size

    N, D, B, layers = 3, 4, 2, 2
    gen = torch.Generator().manual_seed(1)
    w0 = torch.stack([qnn_init_weights(N, layers, active_sd=0.05, generator=gen) for _ in range(B)], dim=-1)
    assert w0.shape == (layers, qnn_layer_size(N), B)

    def run(w):
        prog = make(N, D, batch_size=B)
        for k in range(layers):
            qnn_layer(prog, w[k])
        return prog.ket()

    w = w0.clone().requires_grad_(True)
    ket = run(w)
    for b in range(B):
        ob = OracleBackend()
        ob.begin_circuit(N, cutoff_dim=D)
        for k in range(layers):
            qnn_layer(ob, [float(x) for x in w0[k, :, b]], modes=list(range(N)))
        assert np.abs(ket[b].detach().cpu().numpy() - ob.state().data).max() < TOL
    loss = sum(loss_of(ket[b]) * (b + 1) for b in range(B))
    loss.backward()
    h = 1e-6
    rng = np.random.RandomState(0)
    for _ in range(6):
        k, i, b = rng.randint(layers), rng.randint(qnn_layer_size(N)), rng.randint(B)
        vals = []
        for s in (+1, -1):
            x = w0.clone()
            x[k, i, b] += s * h
            with torch.no_grad():
                kk = run(x)
                vals.append(sum(loss_of(kk[c]).item() * (c + 1) for c in range(B)))
        assert abs(w.grad[k, i, b].item() - (vals[0] - vals[1]) / (2 * h)) < 1e-7, (k, i, b)


@pytest.mark.parametrize("every", [2, 4, 100, None])
def test_checkpoint_thinning_gives_the_same_gradients(every, make):
    """Keeping every c-th state and recomputing the rest must not change a single gradient."""
    n, D = 3, 4
    rng = np.random.RandomState(8)
    th0 = rng.uniform(0.05, 0.4, 15)
    grads = []
    for ev, budget in ((1, 16 << 30), (every, 1 if every is None else 16 << 30)):  # None + tiny budget: sqrt(K)
        th = [torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in th0]
        prog = build(make(n, D, checkpoint_every=ev, checkpoint_bytes=budget), th)
        loss_of(prog.ket()).backward()
        grads.append(np.array([t.grad.item() for t in th]))
    assert np.abs(grads[0] - grads[1]).max() < 1e-13


def test_readouts_match_the_oracle_and_differentiate(make):
    """trace / mean_photon / fidelity / fock_probs of the returned ket: oracle values, and a gradient that
    flows through them (finite differences)."""
    from strawberryfields_b200 import autodiff as A

    n, D = 2, 7
    r = torch.tensor(0.4, dtype=torch.float64, requires_grad=True)

    def run(rv):
        prog = make(n, D)
        prog.displacement(rv, 0.3, 0)
        prog.squeeze(0.25, 0.1, 1)
        prog.beamsplitter(0.5, 0.2, 0, 1)
        return prog.ket()

    ket = run(r)
    ob = OracleBackend()
    ob.begin_circuit(n, cutoff_dim=D)
    ob.displacement(0.4, 0.3, 0)
    ob.squeeze(0.25, 0.1, 1)
    ob.beamsplitter(0.5, 0.2, 0, 1)
    ost = ob.state()
    assert abs(A.trace(ket).item() - ost.trace()) < TOL
    got = A.mean_photon(ket, 1)
    assert np.abs(np.array([got[0].item(), got[1].item()]) - np.array(ost.mean_photon(1))).max() < TOL
    assert np.abs(A.fock_probs(ket).detach().cpu().numpy() - ost.all_fock_probs()).max() < TOL
    target = np.zeros((D, D), dtype=complex)
    target[1, 0] = 1.0
    assert abs(A.fidelity(ket, target).item() - ost.fock_prob([1, 0])) < TOL
    loss = A.mean_photon(ket, 0)[0] + 0.5 * A.trace(ket)
    loss.backward()
    h = 1e-6
    with torch.no_grad():
        vals = [(A.mean_photon(run(0.4 + s * h), 0)[0] + 0.5 * A.trace(run(0.4 + s * h))).item() for s in (1, -1)]
    assert abs(r.grad.item() - (vals[0] - vals[1]) / (2 * h)) < 1e-7


def test_constant_prefix_and_second_backward(make):
    """Gates before the first differentiable one keep no checkpoint; a second backward pass over the
    same graph is refused (the checkpoints are consumed)."""
    n, D = 2, 5
    t = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)

    def run(tv):
        prog = make(n, D)
        prog.squeeze(0.2, 0.4, 0)
        prog.beamsplitter(0.5, 0.1, 0, 1)
        prog.displacement(tv, 0.2, 1)
        prog.beamsplitter(0.3, tv, 1, 0)
        return prog.ket()

    loss = loss_of(run(t))
    loss.backward(retain_graph=True)
    h = 1e-6
    with torch.no_grad():
        fd = (loss_of(run(0.3 + h)).item() - loss_of(run(0.3 - h)).item()) / (2 * h)
    assert abs(t.grad.item() - fd) < 1e-7
    with pytest.raises(RuntimeError, match="consumed"):
        loss.backward()


def test_argument_errors(make):
    prog = make(2, 4)
    with pytest.raises(ValueError, match="modes are not valid"):
        prog.beamsplitter(0.1, 0.2, 0, 0)
    with pytest.raises(ValueError, match="modes are not valid"):
        prog.rotation(0.1, 2)
    with pytest.raises(ValueError, match="scalar or have shape"):
        prog.rotation(torch.zeros(3, dtype=torch.float64), 0)
