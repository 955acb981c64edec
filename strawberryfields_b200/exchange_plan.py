"""Exchange planning for sharded states (pure host logic, no torch, no device).

A sharded state keeps ``g`` of its ``n`` tensor axes split over the ranks; a gate can only run
when all its modes sit on whole (local) axes.  One exchange swaps ALL ``g`` sharded axes with
``g`` local ones (``ShardedCircuit._exchange``), moving (P-1)/P of every shard over NVLink, so
the number of exchanges is the communication cost of a flush.  This module decides, for a
queue of gates,

* the order the gates run in (any order that keeps gates sharing a mode in program order),
* which local modes every exchange evicts,
* and, for a state that is still the vacuum, which modes start out sharded and which mode
  lives on the innermost axis (the layout of |0..0> is free).

``greedy`` is the online rule (evict the modes whose next use is farthest away -- Belady);
``plan`` runs a bounded depth-first search seeded with the greedy answer and never returns
anything worse.  On the 9- and 10-mode interferometer circuits of BASELINE config 5 over 8
ranks the search needs 3 exchanges where the greedy rule needs 5 and 4.  The work is bounded
by a budget of queue sweeps (a few milliseconds of Python) and ``ShardedCircuit`` caches the
plan per queue structure.

The reference has no distributed path; the gate-order freedom used here
is the one its ``pu.optimize_circuit`` / DAG utilities express
(``/root/reference/strawberryfields/program_utils.py:329``): gates on disjoint modes commute.
"""
from __future__ import annotations

from itertools import combinations

# Costs are in quarter exchanges.  What an exchange costs depends on the INNERMOST position it evicts: the
# block copy moves contiguous runs of (elements behind that position) * 16 bytes.  Evicting the innermost
# axis itself leaves 16*D/p-byte (80 B) runs -- measured ~3x slower over NVLink; the axis next to it leaves
# 16*D-byte (160 B) runs -- measured 382 GB/s against 600-700 GB/s for runs of 1.6 KB and more (round 2:
# bulk-copy exchange kernel, profiles/r02_xchg_probe.json).
EXCHANGE_COST = 4
INNER_PENALTY = 8     # evicting position n-1
NEXT_PENALTY = 2      # evicting position n-2
THIRD_PENALTY = 1     # evicting position n-3: 16*D*D-byte (1.6 KB) runs, measured 600 against 675-710 GB/s
DEFAULT_BUDGET = 2500  # queue sweeps one plan() may spend (~5 us each for a 60-gate queue)
_FAR = 1 << 30


def _mask(modes):
    m = 0
    for a in modes:
        m |= 1 << a
    return m


def sweep(masks, order, sharded):
    """Run everything runnable: gates (in queue order) whose modes are local and not behind an
    earlier blocked gate.  ``masks[i]`` / ``sharded`` are mode bit masks.  Returns (run, rest)."""
    blocked, run, rest = sharded, [], []
    for i in order:
        m = masks[i]
        if m & blocked:
            blocked |= m
            rest.append(i)
        else:
            run.append(i)
    return run, rest


def _swap(phys, g, T):
    phys = list(phys)
    for k in range(g):
        phys[k], phys[T[k]] = phys[T[k]], phys[k]
    return phys


def _belady(op_axes, order, phys, g, keep_runs_long=True):
    """Positions to evict by the online rule; ``order`` is the blocked remainder of the queue."""
    n = len(phys)
    next_use = {}
    for r, i in enumerate(order):
        for a in op_axes[i]:
            next_use.setdefault(a, r)
    # the earliest blocked gate must be runnable after the exchange (progress guarantee): the whole
    # axes that hold ITS modes are never evicted; the innermost axis only when nothing else is left
    need = set(op_axes[order[0]])
    cand = [p for p in range(g, n) if phys[p] not in need]
    if len(cand) < g:
        raise ValueError("a %d-mode state has too few whole axes to exchange %d sharded ones" % (n, g))
    # keep the runs of the block copy long when there is a choice (positions are not final while the layout
    # of a fresh state is still being chosen: then only the innermost axis is avoided)
    for avoid in ((n - 1, n - 2, n - 3), (n - 1, n - 2), (n - 1,)) if keep_runs_long else ((n - 1,),):
        if len([p for p in cand if p not in avoid]) >= g:
            cand = [p for p in cand if p not in avoid]
            break
    far = sorted(((next_use.get(phys[p], _FAR), p) for p in cand), reverse=True)
    return sorted(p for _, p in far[:g])


def exchange_cost(T, n, penalty=True):
    """cost of one exchange that evicts the (sorted) positions ``T`` of an ``n``-axis layout"""
    if not penalty:
        return EXCHANGE_COST
    last = T[-1]
    return EXCHANGE_COST + (INNER_PENALTY if last == n - 1 else NEXT_PENALTY if last == n - 2 else
                            THIRD_PENALTY if last == n - 3 else 0)


def greedy(op_axes, phys, g, left=None, keep_runs_long=True):
    """The online plan.  Returns (cost, steps): steps are ``("run", [i, ..])`` /
    ``("exchange", (evicted modes))``; cost = sum of ``exchange_cost`` (quarter exchanges, with the
    penalties for short runs)."""
    phys = list(phys)
    n = len(phys)
    masks = [_mask(a) for a in op_axes]
    order = list(range(len(op_axes)))
    steps, cost = [], 0
    while True:
        run, order = sweep(masks, order, _mask(phys[:g]))
        if left is not None:
            left[0] -= 1
        if run:
            steps.append(("run", run))
        if not order:
            return cost, steps
        T = _belady(op_axes, order, phys, g, keep_runs_long)
        cost += exchange_cost(T, n, keep_runs_long)
        steps.append(("exchange", tuple(phys[p] for p in T)))
        phys = _swap(phys, g, T)


def search(op_axes, phys, g, left, penalty=True, bound=None):
    """Cheapest plan found by a depth-first search from layout ``phys`` (cost = sum of
    ``exchange_cost``; ``penalty=False`` counts exchanges only); ``left[0]`` is the
    remaining sweep budget, shared with the caller.  Returns (cost, steps), or None when nothing
    cheaper than ``bound`` was found."""
    n = len(phys)
    masks = [_mask(a) for a in op_axes]
    best = [bound if bound is not None else _FAR, None]
    seen = {}

    def rec(order, phys, cost, trail):
        run, rest = sweep(masks, order, _mask(phys[:g]))
        left[0] -= 1
        if run:
            trail = trail + [("run", run)]
        if not rest:
            if cost < best[0]:
                best[0], best[1] = cost, trail
            return
        if cost + EXCHANGE_COST >= best[0] or left[0] <= 0:
            return
        key = (tuple(rest), _mask(phys[:g]), phys[n - 1], phys[n - 2], phys[n - 3])
        if seen.get(key, _FAR) <= cost:
            return
        seen[key] = cost
        opts = []
        for T in combinations(range(g, n), g):
            c = exchange_cost(T, n, penalty)
            if cost + c >= best[0]:
                continue
            if left[0] <= 0:
                break
            left[0] -= 1
            new = _swap(phys, g, T)
            r2 = sweep(masks, rest, _mask(new[:g]))[1]
            if len(r2) < len(rest):  # an exchange that unblocks nothing is never useful
                opts.append((c, len(r2), T, new))
        opts.sort()
        for c, _, T, new in opts:
            if cost + c < best[0]:
                rec(rest, new, cost + c, trail + [("exchange", tuple(phys[p] for p in T))])

    rec(list(range(len(op_axes))), list(phys), 0, [])
    return None if best[1] is None else (best[0], best[1])


def replicated_prefix(op_axes, k_max):
    """Lazy vacuum on a sharded state: split a fresh program into the gates that can run while at
    most ``k_max`` modes are entangled (a tensor small enough to be replicated on every rank) and
    the rest.  Only two-mode gates entangle; single-mode gates on untouched modes just update a
    product factor.  A gate goes to the rest when it would exceed ``k_max`` or shares a mode with a
    gate already there (program order per mode).  Returns (replicated, rest) index lists."""
    active, blocked, rep, rest = set(), set(), [], []
    for i, ax in enumerate(op_axes):
        if not blocked.isdisjoint(ax):
            blocked.update(ax)
            rest.append(i)
            continue
        grown = active | set(ax) if len(ax) > 1 else active
        if len(grown) <= k_max:
            active = grown
            rep.append(i)
        else:
            blocked.update(ax)
            rest.append(i)
    return rep, rest


def plan_cost(phys0, g, steps):
    """cost of a plan, with the short-run penalties, when it starts from layout ``phys0``"""
    phys, n, cost = list(phys0), len(phys0), 0
    for s in steps:
        if s[0] == "exchange":
            pos = {m: p for p, m in enumerate(phys)}
            T = sorted(pos[m] for m in s[1])
            cost += exchange_cost(T, n)
            phys = _swap(phys, g, T)
    return cost


def exchanges(steps):
    return sum(1 for s in steps if s[0] == "exchange")


def plan(op_axes, phys, g, free_layout=False, budget=DEFAULT_BUDGET):
    """Plan a flush.  ``op_axes``: modes of every queued gate, in program order; ``phys``: mode on
    every tensor axis (the first ``g`` are sharded); ``free_layout``: the state is still |0..0>, so
    the planner may also pick the layout.  Returns ``(phys0, steps)``: the layout to start from
    (== ``phys`` unless ``free_layout``) and the steps, exchanges given as the tuple of evicted
    MODES (the ``g`` sharded modes all come in)."""
    phys = list(phys)
    n = len(phys)
    op_axes = [tuple(a) for a in op_axes]
    if g == 0 or not op_axes:
        return phys, ([("run", list(range(len(op_axes))))] if op_axes else [])
    left = [budget if not free_layout else budget // 4]
    cost, base = greedy(op_axes, phys, g)
    if cost > 0:
        found = search(op_axes, phys, g, left, bound=cost)
        if found:
            base = found[1]
    if not free_layout or exchanges(base) == 0:
        return phys, base

    # ---- vacuum: choose the initially sharded modes, then the two innermost modes; the given sharded set
    # is kept unless another one saves at least one exchange ----
    def layout(sh):
        return list(sh) + [m for m in phys if m not in sh]

    left = [budget - budget // 4]
    starts = []
    for sh in combinations(sorted(phys), g):
        if left[0] <= budget // 3:  # keep at least a third of the budget for the searches
            break
        try:
            starts.append((greedy(op_axes, layout(sh), g, left, keep_runs_long=False)[0], sh))
        except ValueError:
            continue
    starts.sort()
    best = (exchanges(base) * EXCHANGE_COST, base, tuple(phys[:g]))
    for _, sh in starts[:4]:
        if left[0] <= 0:
            break
        # positions are not final yet: search without the short-run penalties, fix the layout afterwards
        found = search(op_axes, layout(sh), g, left, penalty=False, bound=best[0])
        if found is not None:
            best = (found[0], found[1], sh)
    _, steps, sh = best
    uses = {m: 0 for m in phys}
    for a in op_axes:
        for m in a:
            uses[m] += 1
    # A plan names the modes every exchange evicts, so it is valid for ANY order of the whole axes; the
    # order decides which positions the exchanges touch.  Innermost axis: a mode that is never exchanged,
    # the least-used one (the streaming kernels are slowest on the last axis); the axis next to it: a mode
    # that keeps the exchanges away from the two innermost positions, where the runs of the block copy
    # are short; everything else keeps ascending order.
    rest = [m for m in phys if m not in sh]
    if len(rest) < 3:
        return list(sh) + rest, steps
    together = {}   # two-mode gates per unordered mode pair
    for a in op_axes:
        if len(a) == 2:
            together[frozenset(a)] = together.get(frozenset(a), 0) + 1
    choice = None
    for last in rest:
        # gates of the innermost mode with a partner that is NOT its neighbour in memory take the slow tile
        # shapes of the innermost-axis kernel (rows: 3.8-4.8 TB/s against 5.7-6.1 for adjacent axes)
        pairs_of_last = sum(c for k, c in together.items() if last in k)
        for nxt in rest:
            if nxt == last:
                continue
            lay = list(sh) + [m for m in rest if m not in (last, nxt)] + [nxt, last]
            apart = pairs_of_last - together.get(frozenset((last, nxt)), 0)
            key = (plan_cost(lay, g, steps), uses[last] + apart, uses[nxt], last, nxt)
            if choice is None or key < choice[0]:
                choice = (key, lay)
    return choice[1], steps


def check(op_axes, phys0, g, steps):
    """Replay a plan on the host and verify it: every gate runs exactly once, on local modes, and
    gates sharing a mode keep their program order.  Returns the final layout."""
    phys = list(phys0)
    done, last_on_mode = set(), {}
    for s in steps:
        if s[0] == "run":
            sharded = set(phys[:g])
            for i in s[1]:
                assert i not in done, "gate %d runs twice" % i
                for m in op_axes[i]:
                    assert m not in sharded, "gate %d runs on a sharded mode" % i
                    assert last_on_mode.get(m, -1) < i, "gate %d overtakes a later gate on mode %d" % (i, m)
                    last_on_mode[m] = i
                done.add(i)
        else:
            pos = {m: p for p, m in enumerate(phys)}
            T = sorted(pos[m] for m in s[1])
            assert len(T) == g and T[0] >= g, "exchange must evict %d local modes" % g
            phys = _swap(phys, g, T)
    assert done == set(range(len(op_axes))), "gates left over"
    return phys
