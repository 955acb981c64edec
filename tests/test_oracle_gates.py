"""Pin oracle/gates.py (restated thewalrus.fock_gradients)
against expm of the generators documented in the reference front end
(strawberryfields/ops.py:1497-1514, 1612-1626, 1866-1885, 1956-1973, 2037-2051)
and against closed forms the reference's own backend tests use."""
import numpy as np
import pytest
from scipy.linalg import expm

from oracle import gates as g

TOL = 1e-12
D = 8


def _a(N):
    return np.diag(np.sqrt(np.arange(1, N)), 1).astype(complex)


@pytest.mark.parametrize("r,phi", [(0.4, 0.7), (0.0, 0.3), (1.1, -2.0)])
def test_displacement_vs_expm(r, phi):
    N = 80
    a = _a(N)
    al = r * np.exp(1j * phi)
    U = expm(al * a.conj().T - np.conj(al) * a)[:D, :D]
    assert np.abs(U - g.displacement(r, phi, D)).max() < TOL


@pytest.mark.parametrize("r,theta", [(0.3, 1.1), (0.0, 0.0), (0.6, -0.4)])
def test_squeezing_vs_expm(r, theta):
    N = 120
    a = _a(N)
    ad = a.conj().T
    z = r * np.exp(1j * theta)
    U = expm((np.conj(z) * a @ a - z * ad @ ad) / 2)[:D, :D]
    assert np.abs(U - g.squeezing(r, theta, D)).max() < TOL


def _two(N):
    a = _a(N)
    eye = np.eye(N)
    return np.kron(a, eye), np.kron(eye, a)


@pytest.mark.parametrize("theta,phi", [(0.6, 0.9), (np.pi / 4, np.pi / 2), (-0.3, 0.0)])
def test_beamsplitter_vs_expm(theta, phi):
    N = 2 * D  # BS conserves photon number: exact once N > 2(D-1)
    a1, a2 = _two(N)
    G = theta * (np.exp(1j * phi) * a1 @ a2.conj().T - np.exp(-1j * phi) * a1.conj().T @ a2)
    U = expm(G).reshape(N, N, N, N)[:D, :D, :D, :D]
    assert np.abs(U - g.beamsplitter_tw(theta, phi, D)).max() < TOL


def test_mzgate_vs_decomposition():
    # ops.py:1956-1973: MZ = BS(pi/4, pi/2) (R(phi_in) x I) BS(pi/4, pi/2) (R(phi_ex) x I)
    N = 2 * D
    a1, a2 = _two(N)
    BS = expm((np.pi / 4) * (1j * a1 @ a2.conj().T + 1j * a1.conj().T @ a2))
    eye = np.eye(N)

    def R(t):
        return np.kron(np.diag(np.exp(1j * t * np.arange(N))), eye)

    pin, pex = 0.3, 1.2
    MZ = (BS @ R(pin) @ BS @ R(pex)).reshape(N, N, N, N)[:D, :D, :D, :D]
    assert np.abs(MZ - g.mzgate_tw(pin, pex, D)).max() < TOL


@pytest.mark.parametrize("r,theta", [(0.25, 0.4), (0.5, -1.0)])
def test_two_mode_squeezing_vs_expm(r, theta):
    N = 48
    a1, a2 = _two(N)
    z = r * np.exp(1j * theta)
    U = expm(z * a1.conj().T @ a2.conj().T - np.conj(z) * a1 @ a2).reshape(N, N, N, N)[:D, :D, :D, :D]
    assert np.abs(U - g.two_mode_squeezing_tw(r, theta, D)).max() < 1e-11


def test_tmsv_closed_form():
    # tests/backend/test_twomode_squeezing_operation.py:32-46
    r, theta = 0.4, 0.3
    Z = g.two_mode_squeezing_tw(r, theta, D)
    k = np.arange(D)
    expect = (np.exp(1j * theta) * np.tanh(r)) ** k / np.cosh(r)
    assert np.abs(Z[k, k, 0, 0] - expect).max() < TOL


def test_coherent_column():
    # tests/backend/test_displacement_operation.py:74-94: D(alpha)|0> is the coherent state
    r, phi = 0.5, 0.3
    assert np.abs(g.displacement(r, phi, D)[:, 0] - g.coherent_state(r, phi, D)).max() < TOL


def test_loss_kraus_completeness_on_low_photon_block():
    T = 0.37
    S = sum(E.conj().T @ E for E in g.loss_kraus(T, D))
    assert np.abs(S - np.eye(D)).max() < TOL
