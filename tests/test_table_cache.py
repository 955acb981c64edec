"""Parameter-keyed gate-table cache (reference: the lru_cache decorators of
``fockbackend/ops.py:208-343``): a repeated circuit launches no generator / compose / fold kernel,
cached tables are never modified by later folds, and results are unchanged."""
import collections

import numpy as np
import pytest

from fake_lib import FakeLib


class CountingLib(FakeLib):
    def __init__(self):
        super().__init__()
        object.__setattr__(self, "calls", collections.Counter())

    def __getattribute__(self, name):
        if name.startswith("b200_"):
            object.__getattribute__(self, "calls")[name] += 1
        return object.__getattribute__(self, name)


@pytest.fixture
def counting(monkeypatch):
    from strawberryfields_b200 import circuit, lib

    fake = CountingLib()
    monkeypatch.setattr(lib, "_lib", fake)
    monkeypatch.setattr(circuit, "_TEST_HOST_MODE", True)
    circuit.TABLES.clear()
    yield fake
    circuit.TABLES.clear()


TABLE_KERNELS = ("b200_gen_gate1", "b200_gen_gate2", "b200_gen_diag", "b200_compose_gate1", "b200_fold_diag_gate1",
                 "b200_fold_diag_gate2", "b200_mul_tables")


@pytest.mark.parametrize("pure", [True, False])
def test_repeated_circuit_hits_the_cache(counting, pure):
    from oracle.fock_oracle import OracleBackend
    from strawberryfields_b200 import workloads as W
    from strawberryfields_b200.backend import B200FockBackend

    calls = W.config2_circuit(4, seed=3) + [("kerr_interaction", 0.2, 1), ("two_mode_squeeze", 0.1, 0.4, 0, 3),
                                            ("rotation", 0.3, 0), ("rotation", 0.5, 0), ("displacement", 0.2, 0.1, 0)]
    be = B200FockBackend()
    be.begin_circuit(4, cutoff_dim=5, pure=pure)
    W.run_calls(be, calls)
    first = be.state().data.copy()
    n_first = sum(counting.calls[k] for k in TABLE_KERNELS)
    assert n_first > 0
    counting.calls.clear()
    be.reset(pure=pure)
    W.run_calls(be, calls)
    second = be.state().data.copy()
    assert sum(counting.calls[k] for k in TABLE_KERNELS) == 0   # every table found by its recipe
    assert counting.calls["b200_apply_gate2"] > 0                # the passes themselves still run
    assert np.array_equal(first, second)
    ob = OracleBackend()
    ob.begin_circuit(4, cutoff_dim=5, pure=pure)
    W.run_calls(ob, calls)
    assert np.abs(second - ob.state().data).max() < 1e-12


def test_cached_tables_are_not_modified_by_folds(counting):
    """R -> D folds the rotation INTO a copy of the displacement table; the cached displacement table
    must stay what the generator wrote."""
    from strawberryfields_b200 import circuit
    from strawberryfields_b200.backend import B200FockBackend

    be = B200FockBackend()
    be.begin_circuit(2, cutoff_dim=6)
    be.displacement(0.3, 0.2, 0)
    plain = be.state().ket().copy()
    be.reset()
    be.rotation(0.7, 0)        # pending diagonal
    be.displacement(0.3, 0.2, 0)  # folded: D * R
    be.rotation(-0.4, 0)       # folded on the other side
    be.beamsplitter(0.4, 0.1, 0, 1)
    be.state()
    be.reset()
    be.displacement(0.3, 0.2, 0)
    assert np.array_equal(be.state().ket(), plain)
    assert circuit.TABLES.hits > 0


def test_device_params_are_never_cached(counting):
    import torch

    from strawberryfields_b200 import DeviceParams
    from strawberryfields_b200.backend import B200FockBackend

    be = B200FockBackend()
    be.begin_circuit(2, cutoff_dim=5)
    for val in (0.2, 0.5):
        be.reset()
        counting.calls.clear()
        be.displacement(DeviceParams(torch.tensor([val, 0.1], dtype=torch.float64)), None, 0)
        k = be.state().ket()
        assert counting.calls["b200_gen_gate1"] == 1
        assert abs(k[0, 0] - np.exp(-val ** 2 / 2)) < 1e-12


def test_cache_is_bounded():
    from strawberryfields_b200.circuit import TableCache
    import torch

    c = TableCache(max_bytes=10 * 16 * 4)
    for i in range(10):
        c.get(("k", i), lambda: torch.zeros(10, dtype=torch.complex128))
    assert len(c._d) <= 4
    assert ("k", 9) in c._d


def test_limits_are_checked_up_front(counting):
    """cutoffs 28..64 are accepted for single-mode work but a two-mode gate says clearly
    that it is not supported (instead of a shared-memory opt-in failure mid-circuit); batch sizes beyond the
    grid limit are refused at construction."""
    from strawberryfields_b200.backend import B200FockBackend

    be = B200FockBackend()
    be.begin_circuit(2, cutoff_dim=30, lazy_vacuum=False)
    be.rotation(0.3, 0)
    with pytest.raises(ValueError, match="two-mode gates"):
        be.beamsplitter(0.3, 0.1, 0, 1)
    with pytest.raises(ValueError, match="two-mode gates"):
        be.loss(0.5, 0)
    with pytest.raises(ValueError, match="batch_size"):
        B200FockBackend().begin_circuit(1, cutoff_dim=3, batch_size=70000)


def test_batched_vacuum_and_norm_use_one_launch_each(counting):
    from strawberryfields_b200.backend import B200FockBackend

    be = B200FockBackend()
    be.begin_circuit(2, cutoff_dim=4, batch_size=6, lazy_vacuum=False)
    assert counting.calls["b200_set_element"] == 0           # no per-entry launches
    be.displacement(np.linspace(0.1, 0.6, 6), 0.2, 0)
    counting.calls.clear()
    tr = be.state().trace()
    assert counting.calls["b200_norm2"] == 0 and counting.calls["b200_gather_reduce"] >= 1
    assert np.allclose(tr, 1.0, atol=1e-3) and np.shape(tr) == (6,)


def test_device_params_tables_are_generated_in_batches_at_replay(counting):
    """Deferred program + DeviceParams (no recipe: the cache cannot serve them): one generator launch per gate
    KIND for the whole recorded program instead of one per gate; same state as with host parameters."""
    import torch

    from strawberryfields_b200 import DeviceParams
    from strawberryfields_b200 import workloads as W
    from strawberryfields_b200.backend import B200FockBackend

    n, D = 4, 5
    calls = W.config2_circuit(n, seed=9) + [("kerr_interaction", 0.2, 1), ("two_mode_squeeze", 0.1, 0.4, 0, 3)]
    ref = B200FockBackend()
    ref.begin_circuit(n, cutoff_dim=D)
    W.run_calls(ref, calls)
    want = ref.state().ket().copy()

    be = B200FockBackend()
    be.begin_circuit(n, cutoff_dim=D)          # lazy vacuum: gate calls are recorded
    two = ("squeeze", "displacement", "beamsplitter", "mzgate", "two_mode_squeeze")
    counting.calls.clear()
    for c in calls:
        vals = [float(x) for x in c[1:] if not isinstance(x, int)] + [0.0]
        modes = [x for x in c[1:] if isinstance(x, int)]
        dp = DeviceParams(torch.tensor(vals[:2], dtype=torch.float64))
        getattr(be, c[0])(*([dp] + ([None] if c[0] in two else []) + modes))
    assert sum(counting.calls[k] for k in ("b200_gen_gate1", "b200_gen_gate2", "b200_gen_diag")) == 0  # recorded only
    got = be.state().ket()
    assert np.abs(got - want).max() < 1e-12
    # S, D -> 2 launches of gen_gate1; R (+ the single K) -> gen_diag; BS (+ the single S2) -> gen_gate2
    assert counting.calls["b200_gen_gate1"] == 2
    assert counting.calls["b200_gen_diag"] == 2      # one batch of rotations + the lone Kerr gate
    assert counting.calls["b200_gen_gate2"] == 2     # one batch of beamsplitters + the lone S2 gate
