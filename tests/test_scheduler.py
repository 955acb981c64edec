"""Pass scheduler (strawberryfields_b200/scheduler.py): pure logic, checked by replaying its
plan on a symbolic state -- every op must run exactly once, in an order consistent with the
program order on each axis, on axes that really are inside the tile at that moment."""
import numpy as np
import pytest

from strawberryfields_b200 import scheduler as S
from strawberryfields_b200 import workloads as W


def _ops_from_calls(calls):
    ops = []
    for c in calls:
        modes = tuple(x for x in c[1:] if isinstance(x, int))
        if c[0] == "beamsplitter":
            ops.append(S.Op(S.KIND_SUM, modes, coef_size=670))
        elif c[0] == "rotation":
            ops.append(S.Op(S.KIND_DIAG, modes, coef_size=10, weight=0.2))
        else:
            ops.append(S.Op(S.KIND_SINGLE, modes, coef_size=100))
    return ops


@pytest.mark.parametrize("N", [3, 4, 6, 8, 10])
def test_plan_is_a_valid_reordering(N):
    ops = _ops_from_calls(W.config2_circuit(N, seed=1))
    passes, phys = S.plan(ops, list(range(N)), 16, 4000)
    cur = list(range(N))
    seen = []
    for p in passes:
        assert cur[-1] == p.vaxes[2] and cur[p.positions[0]] == p.vaxes[0] and cur[p.positions[1]] == p.vaxes[1]
        assert p.positions[0] < p.positions[1] < N - 1
        assert len(p.ops) <= 16 and sum(o.coef_size for o, _ in p.ops) <= 4000
        for op, tax in p.ops:
            assert tuple(p.vaxes[t] for t in tax) == op.axes
            seen.append(op)
        where = (p.positions[0], p.positions[1], N - 1)
        assert sorted(p.out_perm) == [0, 1, 2]
        for k in range(3):
            cur[where[p.out_perm[k]]] = p.vaxes[k]
    assert cur == phys and sorted(phys) == list(range(N))
    assert len(seen) == len(ops) and {id(o) for o in seen} == {id(o) for o in ops}
    # per-axis program order is preserved
    rank = {id(o): i for i, o in enumerate(ops)}
    last = {}
    for o in seen:
        for a in o.axes:
            assert last.get(a, -1) < rank[id(o)]
            last[a] = rank[id(o)]


def test_fusion_reduces_passes():
    ops = _ops_from_calls(W.config2_circuit(8, seed=42))
    passes, _ = S.plan(ops, list(range(8)), 16, 4000)
    assert len(passes) <= 24  # 80 gates; one pass per gate would be 80


def test_respects_budget():
    ops = [S.Op(S.KIND_SUM, (0, 1), coef_size=670) for _ in range(5)]
    passes, _ = S.plan(ops, [0, 1, 2], 16, 1400)
    assert [len(p.ops) for p in passes] == [2, 2, 1]


def test_needs_three_axes():
    with pytest.raises(ValueError):
        S.plan([S.Op(S.KIND_SINGLE, (0,))], [0, 1], 16, 1000)
