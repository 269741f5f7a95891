"""Install the GPU pool builder behind the reference's seams by rebinding module globals
(SURVEY.md §8b "How to install").  Callers resolve the names in their own module at call time, so
the dispatcher entry points, the assigner and the simulator run unchanged.

    import src.dispatcher.dispatch_osp as dispatch_osp, src.dispatcher.dispatch_sba as dispatch_sba
    import src.simulator.platform as platform, src.simulator.route_functions as rf
    from amod_b200.install import install
    install(rf.mean_travel_time_table, dispatch_osp, dispatch_sba, platform)
"""
from __future__ import annotations

import numpy as np

from . import pool as _pool


def install(mean_travel_time_table, dispatch_osp=None, dispatch_sba=None, platform=None, device: int = 0,
            table_dtype=np.float32, use_device_delays: bool = True, greedy_on_gpu: bool = False,
            exact_assignment: bool = False) -> _pool.PoolBuilder:
    """`table_dtype=np.float32` is the north-star layout (67 MB, L2-resident) and is exact when the
    table's values are fp32-representable; pass np.float64 for bit-exact results on arbitrary tables."""
    tt = np.ascontiguousarray(mean_travel_time_table, dtype=table_dtype)
    if table_dtype == np.float32 and not np.array_equal(tt.astype(np.float64), np.asarray(mean_travel_time_table, dtype=np.float64)):
        raise ValueError("table is not exactly representable in fp32: install with table_dtype=np.float64")
    b = _pool.PoolBuilder(tt, device=device)
    if dispatch_osp is not None:      # B1, dispatch_osp.py:48-51 looks the name up in its own module
        dispatch_osp.compute_candidate_veh_trip_pairs = b.compute_candidate_veh_trip_pairs
        if use_device_delays or greedy_on_gpu or exact_assignment:
            dispatch_osp.score_vt_pairs_with_num_of_orders_and_sche_cost = _pool.score_vt_pairs_with_num_of_orders_and_sche_cost
    if dispatch_sba is not None:      # B2, dispatch_sba.py:15
        dispatch_sba.compute_candidate_veh_req_pairs = b.compute_candidate_veh_req_pairs
        if use_device_delays or greedy_on_gpu or exact_assignment:
            dispatch_sba.score_vt_pairs_with_num_of_orders_and_sche_cost = _pool.score_vt_pairs_with_num_of_orders_and_sche_cost
    if greedy_on_gpu:                 # next row
        b.lazy_records = True         # only the selected pairs are ever materialised as Python records: the README-sanctioned greedy assigner (ilp_assign.py:101-130) on the device pool
        ga = lambda pairs, rids=None, reqs=None, vehs=None, ensure=True: b.greedy_assignment(pairs)
        for mod in (dispatch_osp, dispatch_sba):
            if mod is not None:
                mod.ilp_assignment = ga           # dispatch_osp.py:62, dispatch_sba.py:23 look the name up in their own module
                mod.greedy_assignment = b.greedy_assignment
    if exact_assignment:              # the reference's ILP (ilp_assign.py:10-98) on HiGHS where Gurobi is absent; host-side consumer
        if greedy_on_gpu:
            raise ValueError("choose one assigner: greedy_on_gpu or exact_assignment")
        from .assign_milp import milp_assignment
        b.lazy_records = True         # the model is built from the pool arrays; only the selected pairs become Python records
        for mod in (dispatch_osp, dispatch_sba):
            if mod is not None:
                mod.ilp_assignment = milp_assignment
    if platform is not None:          # B3 was imported by name into the platform module (platform.py:9-11)
        platform.reposition_idle_vehicles_to_nearest_pending_orders = b.reposition_idle_vehicles_to_nearest_pending_orders
    return b
