"""``DeviceCircuit._sample_index`` must return what the reference's sampling lines return from the same numpy
stream (``fockbackend/circuit.py:682-686``; SURVEY Appendix B: "identical outcomes given identical uniform draws"),
although it does not sum with the builtin ``sum``."""
import numpy as np
import pytest

from strawberryfields_b200.circuit import DeviceCircuit


def reference_draw(dist):
    # circuit.py:682-686, verbatim in behaviour: builtin sum over the numpy array, choice over a list of indices
    if sum(dist) != 1:
        return np.random.choice(list(range(len(dist))), p=dist / sum(dist))
    return np.random.choice(list(range(len(dist))), p=dist)


@pytest.mark.parametrize("seed", range(8))
def test_same_draws_as_the_reference_lines(seed):
    rs = np.random.RandomState(seed)
    for trial in range(40):
        n = int(rs.randint(1, 4000))
        dist = rs.rand(n) ** int(rs.randint(1, 7))
        dist /= dist.sum()
        if rs.rand() < 0.5:
            dist[rs.randint(n, size=max(1, n // 3))] = 0.0       # measured distributions are sparse
        dist = dist * ~np.isclose(dist, 0.0)                     # step 4 of Appendix B
        assert np.cumsum(dist)[-1] == sum(dist)                  # the same plain left-to-right double sum
        s = int(rs.randint(2 ** 31))
        np.random.seed(s)
        want = [reference_draw(dist) for _ in range(5)]
        after_ref = np.random.random()
        np.random.seed(s)
        got = [DeviceCircuit._sample_index(dist) for _ in range(5)]
        assert got == want and np.random.random() == after_ref   # same outcomes, same stream position


def test_exactly_normalised_and_single_outcome():
    for dist in (np.array([0.5, 0.25, 0.25]), np.array([1.0]), np.array([0.0, 1.0, 0.0])):
        np.random.seed(5)
        want = [reference_draw(dist) for _ in range(10)]
        np.random.seed(5)
        assert [DeviceCircuit._sample_index(dist) for _ in range(10)] == want
