"""Exact assignment over a candidate pool without Gurobi: `ilp_assignment` (ilp_assign.py:10-98) restated on
scipy.optimize.milp (HiGHS).  HOST-side consumer of the pool, not part of the GPU hot path:
with Gurobi absent the reference can only run its greedy assigner; this lets the exact formulation run on the
pool the CUDA path builds.

Same model as the reference:
    maximise   sum_i score_i * x_i                                   (ilp_assign.py:39-43)
    s.t.       sum_{i of vehicle v} x_i            = 1   for every vehicle         (:62-66)
               sum_{i containing r_j} x_i + y_j    = 1   for every considered request (:69-74)
               y_j = 0 for requests that are being picked up (optional)             (:77-80)
               x, y binary
and the same side effects: the pair list is stably re-sorted by vehicle id in place (:23) and the result is the
ascending list of selected positions in that re-sorted list (:86-88).

PARITY UNPINNED: Gurobi is not installed here, so no golden vectors exist for the exact solver; when several
optimal solutions exist HiGHS and Gurobi may return different ones.  The tests check feasibility, optimality
against brute force on small pools, and that the objective is never below the greedy assigner's.
"""
from __future__ import annotations

import numpy as np


def _solve(pair_veh, score, tr_rows, tr_cols, V, R, P, picking_j, time_limit_sec):
    """rows 0..V-1: one pair per vehicle; rows V..V+R-1: request j in at most one selected pair (+ y_j)."""
    from scipy.optimize import Bounds, LinearConstraint, milp
    from scipy.sparse import csr_matrix
    rows = np.concatenate([pair_veh, V + tr_rows, V + np.arange(R, dtype=np.int64)])
    cols = np.concatenate([np.arange(P, dtype=np.int64), tr_cols, P + np.arange(R, dtype=np.int64)])
    A = csr_matrix((np.ones(rows.size), (rows, cols)), shape=(V + R, P + R))
    ub = np.ones(P + R)
    ub[P + np.asarray(picking_j, dtype=np.int64)] = 0.0
    c = np.concatenate([-np.asarray(score, dtype=np.float64), np.zeros(R)])
    options = {"disp": False}
    if time_limit_sec is not None:
        options["time_limit"] = float(time_limit_sec)
    res = milp(c, constraints=LinearConstraint(A, lb=np.ones(V + R), ub=np.ones(V + R)), integrality=np.ones(P + R),
               bounds=Bounds(np.zeros(P + R), ub), options=options)
    if res.x is None:          # infeasible model: the reference prints the solver error and selects nothing (:90-93)
        return []
    return np.flatnonzero(res.x[:P] > 0.5).tolist()


def _picking(considered_rids, reqs, ensure):
    if not ensure:
        return []
    out = []
    for j, rid in enumerate(considered_rids):
        st = reqs[rid].status
        if getattr(st, "name", st) == "PICKING":
            out.append(j)
    return out


def _from_pool_arrays(lazy, considered_rids, reqs, vehs, ensure, time_limit_sec):
    """Same model straight from the SoA pool behind a LazyPoolRecords (install(..., exact_assignment=True)): no
    Python record is built for the 10^4-10^6 pairs; the caller then materialises only the selected ones."""
    p = lazy._pool
    P, R, V = p.n_entries, len(considered_rids), len(vehs)
    cur = lazy._perm if lazy._perm is not None else np.arange(P, dtype=np.int64)
    veh_ids = np.fromiter((v.id for v in vehs), dtype=np.int64, count=V)
    order = np.argsort(veh_ids[p.ent_veh[cur]], kind="stable")          # ilp_assign.py:23 as a permutation
    lazy.apply_order(order)
    perm = np.asarray(lazy._perm, dtype=np.int64)
    toff = p.ent_trip_off.astype(np.int64)
    tlen = (toff[1:] - toff[:-1])[perm]
    score = tlen.astype(np.float64) if lazy.scored else np.zeros(P)
    pos = np.repeat(np.arange(P, dtype=np.int64), tlen)
    first = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(tlen, out=first[1:])
    src = np.repeat(toff[:-1][perm] - first[:-1], tlen) + np.arange(int(first[-1]), dtype=np.int64)
    rid = lazy._req_id[p.trip_req[src]]
    cons = np.asarray(considered_rids, dtype=np.int64)
    by = np.argsort(cons, kind="stable")
    at = np.searchsorted(cons[by], rid)
    at = np.minimum(at, max(R - 1, 0))
    known = (cons[by][at] == rid) if R else np.zeros(rid.size, dtype=bool)   # requests outside considered_rids are skipped (:52-58)
    return _solve(veh_ids[p.ent_veh[perm]], score, by[at[known]], pos[known], V, R, P, _picking(considered_rids, reqs, ensure),
                  time_limit_sec)


def milp_assignment(veh_trip_pairs, considered_rids, reqs, vehs, ensure_assigning_orders_that_are_picking: bool = True,
                    time_limit_sec: float | None = None) -> list[int]:
    """Drop-in for `ilp_assign.ilp_assignment`: rebind `dispatch_osp.ilp_assignment` / `dispatch_sba.ilp_assignment`
    (the callers look the name up in their own module, dispatch_osp.py:62, dispatch_sba.py:24)."""
    from .marshal import LazyPoolRecords
    if len(veh_trip_pairs) == 0:
        return []
    if isinstance(veh_trip_pairs, LazyPoolRecords) and veh_trip_pairs._full is None and not veh_trip_pairs._cache:
        return _from_pool_arrays(veh_trip_pairs, considered_rids, reqs, vehs, ensure_assigning_orders_that_are_picking, time_limit_sec)
    veh_trip_pairs.sort(key=lambda e: e[0].id)            # ilp_assign.py:23 (stable)
    P, R, V = len(veh_trip_pairs), len(considered_rids), len(vehs)
    rid_to_j = {rid: j for j, rid in enumerate(considered_rids)}
    pair_veh = np.fromiter((p[0].id for p in veh_trip_pairs), dtype=np.int64, count=P)
    score = np.fromiter((p[4] for p in veh_trip_pairs), dtype=np.float64, count=P)
    tr_rows, tr_cols = [], []
    for i, p in enumerate(veh_trip_pairs):
        for req in p[1]:
            j = rid_to_j.get(req.id)
            if j is not None:                              # requests outside considered_rids are skipped (:52-58)
                tr_rows.append(j); tr_cols.append(i)
    return _solve(pair_veh, score, np.asarray(tr_rows, dtype=np.int64), np.asarray(tr_cols, dtype=np.int64), V, R, P,
                  _picking(considered_rids, reqs, ensure_assigning_orders_that_are_picking), time_limit_sec)
