"""Pass scheduler of the lazy gate queue: which gates share one HBM round trip.

The engine hands the backend one gate at a time (``/root/reference/strawberryfields/
engine.py:422-457``) and the reference applies each with a full pass over the state.  Here
gates are queued and, when the state is observed, grouped into *tile passes*
(``b200_apply_tile_pass``): a pass stages D x D x D tiles spanned by two arbitrary tensor
axes plus the innermost axis and applies every queued operator that lives on those three
axes.  Because a pass may write its tile back with the three axes permuted, the scheduler
also chooses which axis is left in the innermost position -- the only constraint linking
consecutive passes -- and keeps the resulting logical -> physical axis permutation.

Pure logic, no device access: unit-tested on the CPU (tests/test_scheduler.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

KIND_SINGLE, KIND_SUM, KIND_DIFF, KIND_DIAG = 0, 1, 2, 3


@dataclass
class Op:
    kind: int
    axes: Tuple[int, ...]          # virtual axes (1 for SINGLE / DIAG, 2 for SUM / DIFF)
    table: object = None           # device tensor [nbatch, size]
    conj: int = 0
    coef_size: int = 0             # entries of `table` per batch element
    weight: float = 1.0            # what the op is worth when ranking candidate tiles


@dataclass
class TilePassPlan:
    positions: Tuple[int, int]                  # physical positions of tile axes 0 and 1 (2 = innermost)
    vaxes: Tuple[int, int, int]                 # virtual axes on tile axes 0, 1, 2 before the pass
    ops: List[Tuple[Op, Tuple[int, ...]]] = field(default_factory=list)   # (op, tile axes it acts on)
    out_perm: Tuple[int, int, int] = (0, 1, 2)  # tile axis k is written to tile position out_perm[k]


def _executable(ops: Sequence[Op], order: Sequence[int], triple, max_ops, coef_budget):
    """Indices (in program order) of the queued ops that can run inside a tile on `triple`:
    an op runs if all its axes are in the tile and no earlier, skipped op shares an axis."""
    blocked = set()
    out = []
    coef = 0
    for i in order:
        op = ops[i]
        ok = True
        for a in op.axes:
            if a not in triple or a in blocked:
                ok = False
                break
        if ok and (len(out) >= max_ops or coef + op.coef_size > coef_budget):
            ok = False
        if ok:
            out.append(i)
            coef += op.coef_size
        else:
            blocked.update(op.axes)
            if len(blocked) >= 3 and all(t in blocked for t in triple):
                break
    return out


def _ready_axes(ops: Sequence[Op], order: Sequence[int], limit=12):
    """Axis sets of the first ops of every dependency chain (the DAG frontier)."""
    blocked = set()
    ready = []
    for i in order:
        ax = ops[i].axes
        if not any(a in blocked for a in ax):
            ready.append(ax)
            if len(ready) >= limit:
                break
        blocked.update(ax)
    return ready


def _candidates(ready, inn, all_axes):
    """Tiles {inn, x, y} worth scoring: built from the frontier ops' axes."""
    pool = []
    for ax in ready:
        for a in ax:
            if a != inn and a not in pool:
                pool.append(a)
    if len(pool) < 2:
        for a in all_axes:
            if a != inn and a not in pool:
                pool.append(a)
            if len(pool) >= 2:
                break
    cands = set()
    for ax in ready:
        s = set(ax) - {inn}
        if len(s) == 2:
            cands.add(tuple(sorted(s)))
        else:
            for b in pool:
                t = s | {b}
                if len(t) == 2:
                    cands.add(tuple(sorted(t)))
    if not cands and len(pool) >= 2:
        cands.add(tuple(sorted(pool[:2])))
    return [(x, y) for x, y in cands]


def _best(ops, order, inn, all_axes, max_ops, coef_budget):
    ready = _ready_axes(ops, order)
    best, best_score, best_exec = None, -1.0, []
    for x, y in _candidates(ready, inn, all_axes):
        ex = _executable(ops, order, (inn, x, y), max_ops, coef_budget)
        score = sum(ops[i].weight for i in ex)
        if score > best_score:
            best, best_score, best_exec = (x, y), score, ex
    return best, best_score, best_exec


def plan(ops: Sequence[Op], phys: List[int], max_ops: int, coef_budget: int) -> Tuple[List[TilePassPlan], List[int]]:
    """Group `ops` (program order) into tile passes.

    phys[pos] = virtual axis stored at physical position pos (last = innermost).
    Returns (passes, new phys)."""
    phys = list(phys)
    A = len(phys)
    if A < 3:
        raise ValueError("tile passes need at least three tensor axes")
    all_axes = list(range(A))
    order = list(range(len(ops)))
    passes: List[TilePassPlan] = []
    guard = 0
    while order:
        guard += 1
        if guard > 4 * len(ops) + 8:
            raise RuntimeError("scheduler failed to make progress")
        inn = phys[-1]
        pair, score, ex = _best(ops, order, inn, all_axes, max_ops, coef_budget)
        if not ex:
            # the frontier op does not fit the budget together with anything: run it alone, unbudgeted
            first = ops[order[0]]
            s = [a for a in first.axes if a != inn]
            for a in all_axes:
                if len(s) >= 2:
                    break
                if a != inn and a not in s:
                    s.append(a)
            pair = tuple(sorted(s[:2]))
            ex = [order[0]]
        x, y = pair
        px, py = phys.index(x), phys.index(y)
        if px > py:
            x, y, px, py = y, x, py, px
        tile_vaxes = (x, y, inn)
        done = set(ex)
        rest = [i for i in order if i not in done]
        # which of the three axes to leave innermost: one-step lookahead
        choice, choice_score = inn, -1.0
        if rest:
            for z in (inn, x, y):
                _, sc, _ = _best(ops, rest, z, all_axes, max_ops, coef_budget)
                if sc > choice_score + 1e-9:
                    choice, choice_score = z, sc
        # tile axis k -> tile position out_perm[k]; the two non-innermost axes keep their order
        others = [k for k in range(3) if tile_vaxes[k] != choice]
        out_perm = [0, 0, 0]
        out_perm[tile_vaxes.index(choice)] = 2
        out_perm[others[0]], out_perm[others[1]] = 0, 1
        p = TilePassPlan(positions=(px, py), vaxes=tile_vaxes, out_perm=tuple(out_perm))
        for i in ex:
            p.ops.append((ops[i], tuple(tile_vaxes.index(a) for a in ops[i].axes)))
        passes.append(p)
        pos_of_tile = (px, py, A - 1)
        for k in range(3):
            phys[pos_of_tile[out_perm[k]]] = tile_vaxes[k]
        order = rest
    return passes, phys
