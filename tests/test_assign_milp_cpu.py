"""Exact assignment stand-in (amod_b200/assign_milp.py: ilp_assign.py:10-98 on scipy/HiGHS; Gurobi is absent, so
parity with the reference's solver is unpinned): model feasibility, optimality against brute force, and never
worse than the reference's greedy assigner on real pools."""
import enum
import itertools
import types

import numpy as np
import pytest

from conftest import load_golden
from amod_b200 import marshal
from amod_b200.assign_milp import milp_assignment
import oracle
from test_marshal_cpu import _objects


class OrderStatus(enum.Enum):
    PENDING = 1
    PICKING = 2


def _check(pairs, sel, vehs, considered, reqs, ensure):
    chosen = [pairs[i] for i in sel]
    assert sorted(p[0].id for p in chosen) == [v.id for v in vehs]            # every vehicle exactly one pair
    served = [r.id for p in chosen for r in p[1]]
    assert len(served) == len(set(served))                                     # a request at most once
    if ensure:
        for rid in considered:
            if reqs[rid].status == OrderStatus.PICKING:
                assert rid in served
    return sum(p[4] for p in chosen)


def _brute(pairs, vehs, considered, reqs, ensure):
    by_veh = [[i for i, p in enumerate(pairs) if p[0].id == v.id] for v in vehs]
    best = None
    for combo in itertools.product(*by_veh):
        served = [r.id for i in combo for r in pairs[i][1]]
        if len(served) != len(set(served)):
            continue
        if ensure and any(reqs[rid].status == OrderStatus.PICKING and rid not in served for rid in considered):
            continue
        val = sum(pairs[i][4] for i in combo)
        best = val if best is None else max(best, val)
    return best


@pytest.mark.parametrize("seed", range(6))
def test_matches_brute_force_on_small_pools(seed):
    rng = np.random.default_rng(seed)
    n_veh, n_req = 4, 6
    reqs = [types.SimpleNamespace(id=i, status=OrderStatus.PICKING if rng.random() < 0.25 else OrderStatus.PENDING)
            for i in range(n_req)]
    vehs = [types.SimpleNamespace(id=v) for v in range(n_veh)]
    pairs = []
    picking_owner = {}
    for v in vehs:
        pairs.append([v, [], [], 0.0, 0.0])
    for r in reqs:               # a picking request is already in its vehicle's plan: make that pair exist
        if r.status == OrderStatus.PICKING:
            picking_owner[r.id] = int(rng.integers(n_veh))
    for v in vehs:
        mine = [reqs[r] for r, o in picking_owner.items() if o == v.id]
        if mine:
            pairs.append([v, sorted(mine, key=lambda r: r.id), [], 0.0, float(len(mine))])
        for _ in range(5):
            k = int(rng.integers(1, 4))
            trip = sorted(rng.choice(n_req, size=k, replace=False).tolist())
            pairs.append([v, [reqs[i] for i in trip], [], 0.0, float(k) + float(rng.integers(0, 3)) * 0.25])
    order = rng.permutation(len(pairs))
    pairs = [pairs[i] for i in order]                    # the assigner re-sorts by vehicle id itself
    considered = list(range(n_req))
    for ensure in (True, False):
        work = list(pairs)
        sel = milp_assignment(work, considered, reqs, vehs, ensure)
        assert [p[0].id for p in work] == sorted(p[0].id for p in work)       # stably re-sorted in place (ilp_assign.py:23)
        want = _brute(work, vehs, considered, reqs, ensure)
        if want is None:
            assert sel == []                                                   # infeasible: nothing selected, like the reference
            continue
        got = _check(work, sel, vehs, considered, reqs, ensure)
        assert got == pytest.approx(want, abs=1e-9)


def test_not_worse_than_greedy_on_reference_pools():
    tt, cases = load_golden("rand_real_osp_cap4")
    for ep, gold in cases[:3]:
        reqs_by_id, objs, vehs = _objects(ep)
        for r in objs:
            r.status = OrderStatus.PENDING
        recs = marshal.pool_to_records(gold, ep, vehs, objs)
        for p in recs:
            p[4] = float(len(p[1]))                                            # the reference's scorer (scheduling.py:160-161)
        considered = [int(ep.req_id[i]) for i in ep.search]
        sel = milp_assignment(recs, considered, reqs_by_id, vehs, True)
        got = _check(recs, sel, vehs, considered, reqs_by_id, False)
        _order, gsel = oracle.greedy_assign(gold, ep.n_veh, ep.n_req)
        greedy = float(np.diff(gold.ent_trip_off)[_order[gsel]].sum())
        assert got >= greedy


def test_array_path_equals_record_path():
    """install(exact_assignment=True) hands the solver a LazyPoolRecords: the model is built from the pool arrays
    and must be the record path's model (same rows, same order -> same solution)."""
    tt, cases = load_golden("rand_real_osp_cap4")
    for ep, gold in cases[:3]:
        reqs_by_id, objs, vehs = _objects(ep)
        for k, r in enumerate(objs):
            r.status = OrderStatus.PENDING
        considered = [int(ep.req_id[i]) for i in ep.search]
        recs = marshal.pool_to_records(gold, ep, vehs, objs)
        for p in recs:
            p[4] = float(len(p[1]))
        eager = milp_assignment(recs, considered, reqs_by_id, vehs, True)
        lazy = marshal.LazyPoolRecords(gold, ep, vehs, objs)
        lazy.scored = True
        sel = milp_assignment(lazy, considered, reqs_by_id, vehs, True)
        assert sel == eager
        assert not lazy._cache                                           # nothing was materialised to build the model
        for i in sel[:5]:
            a, b = lazy[i], recs[i]
            assert a[0] is b[0] and a[1] == b[1] and a[2] == b[2] and a[4] == b[4]
