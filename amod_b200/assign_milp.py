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


def milp_assignment(veh_trip_pairs, considered_rids, reqs, vehs, ensure_assigning_orders_that_are_picking: bool = True,
                    time_limit_sec: float | None = None) -> list[int]:
    """Drop-in for `ilp_assign.ilp_assignment`: rebind `dispatch_osp.ilp_assignment` / `dispatch_sba.ilp_assignment`
    (the callers look the name up in their own module, dispatch_osp.py:62, dispatch_sba.py:24)."""
    from scipy.optimize import Bounds, LinearConstraint, milp
    from scipy.sparse import csr_matrix

    if len(veh_trip_pairs) == 0:
        return []
    veh_trip_pairs.sort(key=lambda e: e[0].id)            # ilp_assign.py:23 (stable)
    P, R, V = len(veh_trip_pairs), len(considered_rids), len(vehs)
    rid_to_j = {rid: j for j, rid in enumerate(considered_rids)}
    pair_veh = np.fromiter((p[0].id for p in veh_trip_pairs), dtype=np.int64, count=P)
    score = np.fromiter((p[4] for p in veh_trip_pairs), dtype=np.float64, count=P)
    rows, cols = [pair_veh], [np.arange(P, dtype=np.int64)]          # vehicle rows 0..V-1
    tr_rows, tr_cols = [], []
    for i, p in enumerate(veh_trip_pairs):
        for req in p[1]:
            j = rid_to_j.get(req.id)
            if j is not None:                              # requests outside considered_rids are skipped (:52-58)
                tr_rows.append(V + j); tr_cols.append(i)
    rows.append(np.asarray(tr_rows, dtype=np.int64)); cols.append(np.asarray(tr_cols, dtype=np.int64))
    rows.append(V + np.arange(R, dtype=np.int64)); cols.append(P + np.arange(R, dtype=np.int64))   # + y_j
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    A = csr_matrix((np.ones(rows.size), (rows, cols)), shape=(V + R, P + R))
    ub = np.ones(P + R)
    if ensure_assigning_orders_that_are_picking:
        for j, rid in enumerate(considered_rids):
            st = reqs[rid].status
            if getattr(st, "name", st) == "PICKING":
                ub[P + j] = 0.0
    c = np.concatenate([-score, np.zeros(R)])
    options = {"disp": False}
    if time_limit_sec is not None:
        options["time_limit"] = float(time_limit_sec)
    res = milp(c, constraints=LinearConstraint(A, lb=np.ones(V + R), ub=np.ones(V + R)), integrality=np.ones(P + R),
               bounds=Bounds(np.zeros(P + R), ub), options=options)
    if res.x is None:          # infeasible model: the reference prints the solver error and selects nothing (:90-93)
        return []
    return np.flatnonzero(res.x[:P] > 0.5).tolist()
