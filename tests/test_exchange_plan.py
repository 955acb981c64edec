"""Exchange planner of the sharded path (strawberryfields_b200/exchange_plan.py): pure host
logic.  Every plan is replayed by ``check`` (each gate once, only on local modes, program order
kept per mode); the search is never worse than the online Belady rule, and on the BASELINE
config-5 circuits (C2 generator with N = 9 / 10) it reaches the exchange counts
DESIGN.md quotes."""
import numpy as np
import pytest

from strawberryfields_b200 import exchange_plan as X
from strawberryfields_b200 import workloads as W


def queue_axes(n, seed=42):
    """Modes of the gates the lazy queue emits for the C2/C5 circuit: per-mode S.D products fold
    into one dense gate, rotations fold into the neighbouring beamsplitter, leftovers are diagonals."""
    ops, pending = [], {}
    for c in W.config2_circuit(n, seed=seed):
        if c[0] == "beamsplitter":
            a, b = c[-2:]
            for m in (a, b):
                if pending.pop(m, None) == "dense":
                    ops.append((m,))
            ops.append((a, b))
        else:
            m = c[-1]
            kind = "diag" if c[0] == "rotation" else "dense"
            pending[m] = "dense" if "dense" in (kind, pending.get(m)) else "diag"
    ops += [(m,) for m in sorted(pending)]
    return ops


@pytest.mark.parametrize("n,g,want", [(9, 1, 2), (9, 2, 2), (9, 3, 3), (10, 1, 2), (10, 2, 2), (10, 3, 3), (8, 3, 4)])
def test_config5_exchange_counts(n, g, want):
    ops = queue_axes(n)
    phys0, steps = X.plan(ops, list(range(n)), g, free_layout=True)
    X.check(ops, phys0, g, steps)
    assert X.exchanges(steps) == want
    cost, online = X.greedy(ops, list(range(n)), g)
    assert X.exchanges(steps) <= X.exchanges(online)
    if n >= 9:
        # no exchange evicts one of the two innermost positions: the block copies move runs of >= 100
        # amplitudes (1.6 KB), which is what the bulk-copy exchange kernel needs to stay near the link rate
        assert max(max(T) for T in evicted_positions(phys0, g, steps)) <= n - 3
    # the same holds for the steady state of a REPEATED circuit (the layout one run ends in is where the
    # next one starts; bench.py's timed region), at no more exchanges
    phys = phys0
    for _ in range(3):
        phys1, again = X.plan(ops, phys, g, free_layout=False)
        assert phys1 == phys
        phys = X.check(ops, phys, g, again)
    assert X.exchanges(again) <= want + (1 if (n, g) == (10, 2) else 0)
    if n >= 9 and (n, g) != (10, 2):
        assert max(max(T) for T in evicted_positions(phys1, g, again)) <= n - 3


def evicted_positions(phys0, g, steps):
    phys, out = list(phys0), []
    for s in steps:
        if s[0] == "exchange":
            pos = {m: p for p, m in enumerate(phys)}
            T = sorted(pos[m] for m in s[1])
            out.append(T)
            phys = X._swap(phys, g, T)
    return out


def test_fixed_layout_is_respected():
    ops = queue_axes(9)
    phys = [3, 1, 2, 0, 4, 5, 6, 7, 8]
    phys0, steps = X.plan(ops, phys, 2, free_layout=False)
    assert phys0 == phys
    X.check(ops, phys0, 2, steps)


def test_innermost_axis_is_kept_when_possible():
    ops = queue_axes(9)
    for g in (1, 2, 3):
        phys0, steps = X.plan(ops, list(range(9)), g, free_layout=True)
        inner = phys0[-1]
        assert all(inner not in s[1] for s in steps if s[0] == "exchange")


@pytest.mark.parametrize("seed", range(12))
def test_random_queues(seed):
    rng = np.random.RandomState(seed)
    n = int(rng.randint(5, 12))
    g = int(rng.randint(1, min(3, (n - 2) // 2) + 1))
    ops = []
    for _ in range(int(rng.randint(1, 50))):
        if rng.rand() < 0.3:
            ops.append((int(rng.randint(n)),))
        else:
            a, b = rng.choice(n, 2, replace=False)
            ops.append((int(a), int(b)))
    phys = [int(x) for x in rng.permutation(n)]
    cost, online = X.greedy(ops, phys, g)
    X.check(ops, phys, g, online)
    for free in (False, True):
        phys0, steps = X.plan(ops, phys, g, free_layout=free)
        final = X.check(ops, phys0, g, steps)
        assert sorted(final) == list(range(n))
        assert X.exchanges(steps) <= X.exchanges(online)
        if not free:
            assert phys0 == phys


@pytest.mark.parametrize("n,k_max,want_rep,max_exchanges", [(9, 8, 25, 1), (10, 9, 26, 1), (8, 5, 12, 4)])
def test_replicated_prefix(n, k_max, want_rep, max_exchanges):
    """Lazy vacuum on sharded states: the prefix never entangles more than k_max modes, keeps the
    program order per mode, and together with the rest covers every gate once."""
    ops = queue_axes(n)
    rep, rest = X.replicated_prefix(ops, k_max)
    assert sorted(rep + rest) == list(range(len(ops))) and len(rep) == want_rep
    entangled = set()
    for i in rep:
        if len(ops[i]) > 1:
            entangled.update(ops[i])
    assert len(entangled) <= k_max
    for i in rep:  # nothing in the prefix comes after a postponed gate on the same mode
        assert not any(j < i and set(ops[j]) & set(ops[i]) for j in rest)
    # the rest is an ordinary queue for the exchange planner
    phys0, steps = X.plan([ops[i] for i in rest], list(range(n)), 3, free_layout=True)
    X.check([ops[i] for i in rest], phys0, 3, steps)
    assert X.exchanges(steps) <= max_exchanges  # 9 / 10 modes on 8 ranks: ONE exchange after the prefix


def test_no_sharding_and_empty_queue():
    assert X.plan([], [0, 1, 2], 1) == ([0, 1, 2], [])
    phys0, steps = X.plan([(0, 1), (1, 2)], [0, 1, 2], 0)
    assert steps == [("run", [0, 1])]


def test_too_few_whole_axes():
    # 3 modes, 2 sharded: a two-mode gate can never have both modes local
    with pytest.raises(ValueError):
        X.plan([(0, 1)], [0, 1, 2], 2)


def test_budget_bounds_the_work():
    ops = queue_axes(10)
    phys0, steps = X.plan(ops, list(range(10)), 3, free_layout=True, budget=40)
    X.check(ops, phys0, 3, steps)  # a starved search still returns a valid (online) plan
