"""SURVEY.md 8f-3, second half: the fleet's state advance, the request life cycle and the apply step on the device
(include/amod_b200.h amod_fleet_*; kernels amod_b200/csrc/fleet.cuh).

`DeviceFleetSim` is the host handle of a fleet whose whole epoch state lives in HBM: vehicles (vehicle.py:41-70 incl. the
route as legs and steps), requests (request.py:25-41 incl. status / Tp / Td).  It mirrors, call for call, what
Platform.run_cycle does (platform.py:121-166):

    sim.advance(T, in_main_horizon)            # upd_vehs_and_reqs_stat_to_time (platform.py:213-233)
    sim.add_requests(first, onid, dnid, ...)   # gen_reqs_to_time (platform.py:252-271); Req.__init__ fields from the host
    sim.dispatch(mode, cap, T, n_new)          # pool build -> scorer -> greedy_assignment -> upd_schedule_for_vehicles_...
    sim.rebalance(T)                           # reposition_idle_vehicles_to_nearest_pending_orders (rebalancing_npo.py:8-62)

and `epoch()` runs the four in order.  Nothing but the arrivals goes up; the host reads what it wants with `array()` /
`export()`.  There is no CPU fallback: every method goes through the C-ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .marshal import MODE_OSP, MODE_SBA

ST_IDLE, ST_WORKING, ST_REBALANCING = 1, 2, 3                                   # types.py:18-21
RQ_PENDING, RQ_PICKING, RQ_ONBOARD, RQ_COMPLETE, RQ_WALKAWAY = 1, 2, 3, 4, 5    # types.py:11-16
STATS = ("Ds", "Ts", "Ds_empty", "Ts_empty", "Dr", "Tr", "Lt", "Ld")            # vehicle.py:57-64


class DeviceFleetSim:
    def __init__(self, builder, graph, start_nid, max_stops: int = 24, max_route_steps: int = 4096, request_capacity: int = 1 << 20):
        """builder: PoolBuilder over graph.tt.  graph: tt / dist / pred (1-based predecessors) / lng / lat per node.
        start_nid: the vehicles' start nodes (platform.py:50-55)."""
        self.b = builder
        pred = np.ascontiguousarray(graph.pred, dtype=np.int32)
        dist = np.ascontiguousarray(graph.dist, dtype=np.float64 if builder.tt_dtype == capi.TT_F64 else np.float32)
        if pred.shape != (builder.n_nodes, builder.n_nodes) or dist.shape != pred.shape:
            raise ValueError("the path and distance tables must have the travel-time table's shape")
        builder._check(builder.lib.amod_routes_upload(builder.h, pred.ctypes.data, dist.ctypes.data))
        start = np.ascontiguousarray(start_nid, dtype=np.int32)
        lng = np.ascontiguousarray(graph.lng, dtype=np.float64)
        lat = np.ascontiguousarray(graph.lat, dtype=np.float64)
        self.n_veh, self.max_stops, self.max_legs, self.max_steps = int(start.size), int(max_stops), int(max_stops) + 1, int(max_route_steps)
        builder._check(builder.lib.amod_fleet_create(builder.h, self.n_veh, self.max_stops, self.max_steps, int(request_capacity),
                                                     start.ctypes.data, lng.ctypes.data, lat.ctypes.data))
        self.n_req = 0
        # the host keeps the immutable request fields it uploaded (to name stops in export(); not a fallback for anything)
        self._req_cols = [np.zeros(1024, np.int32), np.zeros(1024, np.int32), np.zeros(1024, np.float64), np.zeros(1024, np.float64)]
        self.last_stats: dict = {}
        self.last_n_search = 0

    # ---- platform.py:252-271 ---------------------------------------------------------------------------------------
    def add_requests(self, first, onid, dnid, tr, ts, clp, cld):
        cols = [np.ascontiguousarray(onid, np.int32), np.ascontiguousarray(dnid, np.int32), np.ascontiguousarray(clp, np.float64),
                np.ascontiguousarray(cld, np.float64), np.ascontiguousarray(tr, np.float64), np.ascontiguousarray(ts, np.float64)]
        n = int(cols[0].size)
        b = self.b
        b._check(b.lib.amod_fleet_add_requests(b.h, int(first), n, *[c.ctypes.data for c in cols]))
        top = int(first) + n
        if top > self._req_cols[0].size:
            self._req_cols = [np.concatenate([a, np.zeros(max(top, 2 * a.size) - a.size, a.dtype)]) for a in self._req_cols]
        for a, c in zip(self._req_cols, cols[:4]):
            a[int(first):top] = c
        self.n_req = top

    _onid = property(lambda self: self._req_cols[0][:self.n_req])
    _dnid = property(lambda self: self._req_cols[1][:self.n_req])
    _clp = property(lambda self: self._req_cols[2][:self.n_req])
    _cld = property(lambda self: self._req_cols[3][:self.n_req])

    # ---- platform.py:213-233 ---------------------------------------------------------------------------------------
    def advance(self, T, update_stats: bool) -> dict:
        b = self.b
        n = C.c_int32(0)
        b._check(b.lib.amod_fleet_advance(b.h, int(T), int(bool(update_stats)), C.byref(n)))
        k = n.value
        ev = dict(veh=np.zeros(k, np.int32), seq=np.zeros(k, np.int32), rid=np.zeros(k, np.int32), pod=np.zeros(k, np.int32), time=np.zeros(k, np.float64))
        if k:
            b._check(b.lib.amod_fleet_events(b.h, *[ev[c].ctypes.data for c in ("veh", "seq", "rid", "pod", "time")], k))
            order = np.lexsort((ev["seq"], ev["veh"]))
            ev = {c: a[order] for c, a in ev.items()}
        return ev

    def set_status(self, rids, status: int):
        r = np.ascontiguousarray(rids, dtype=np.int32)
        self.b._check(self.b.lib.amod_fleet_set_status(self.b.h, r.ctypes.data, int(r.size), int(status)))

    # ---- vehicle.py:231-288 for a batch of vehicles ----------------------------------------------------------------
    def apply_codes(self, veh, off, code, reb_node, reb_ddl):
        veh = np.ascontiguousarray(veh, np.int32); off = np.ascontiguousarray(off, np.int32); code = np.ascontiguousarray(code, np.int32)
        reb_node = np.ascontiguousarray(reb_node, np.int32); reb_ddl = np.ascontiguousarray(reb_ddl, np.float64)
        b = self.b
        b._check(b.lib.amod_fleet_apply(b.h, int(veh.size), veh.ctypes.data, off.ctypes.data, code.ctypes.data, reb_node.ctypes.data, reb_ddl.ctypes.data))

    def apply(self, veh, off, rid, pod, tnid, ddl):
        """Schedules as the reference's (rid, pod, tnid, ddl) stops (types.py:66-69)."""
        veh = np.asarray(veh, np.int32); off = np.asarray(off, np.int64)
        rid = np.asarray(rid, np.int64); pod = np.asarray(pod, np.int64)
        code = np.where(rid < 0, -1, 2 * rid + (pod == -1)).astype(np.int32)
        reb_node = np.zeros(veh.size, np.int32); reb_ddl = np.zeros(veh.size, np.float64)
        is_reb = np.flatnonzero(rid < 0)
        if is_reb.size:
            call = np.searchsorted(off, is_reb, side="right") - 1
            reb_node[call] = np.asarray(tnid)[is_reb]; reb_ddl[call] = np.asarray(ddl)[is_reb]
        self.apply_codes(veh, off, code, reb_node, reb_ddl)

    def apply_selected(self, entries=None, mark_picking: bool = True, reset_considered: bool = False):
        b = self.b
        if entries is None:
            b._check(b.lib.amod_fleet_apply_selected(b.h, None, 0, int(mark_picking), int(reset_considered)))
        else:
            e = np.ascontiguousarray(entries, np.int32)
            b._check(b.lib.amod_fleet_apply_selected(b.h, e.ctypes.data, int(e.size), int(mark_picking), int(reset_considered)))

    # ---- the epoch (platform.py:121-166) ---------------------------------------------------------------------------
    def build(self, mode: int, cap: int, T, n_new: int) -> dict:
        b = self.b
        st = capi.AmodStats()
        ns = C.c_int32(0)
        rc = b.lib.amod_fleet_build(b.h, int(mode), int(cap), int(T), int(n_new), C.byref(st), C.byref(ns))
        b.last_stats = st.as_dict()
        b._check(rc)
        self.last_stats, self.last_n_search = b.last_stats, ns.value
        return b.last_stats

    def dispatch(self, mode: int, cap: int, T, n_new: int):
        """assign_orders_through_osp / _sba with the greedy assigner (dispatch_osp.py:12-92, dispatch_sba.py:8-41): the pool, the
        selection and the new routes stay on the device."""
        self.build(mode, cap, T, n_new)
        _order, sel = self.b.greedy_assign()
        self.apply_selected(None, mark_picking=True, reset_considered=(mode == MODE_OSP))
        return int(sel.size)

    def rebalance(self, T) -> int:
        """rebalancing_npo.py:8-62: idle vehicles x pending requests, nearest first."""
        status = self.array("status"); nid = self.array("nid"); rs = self.array("req_status")
        idle = np.flatnonzero(status == ST_IDLE)
        pend = np.flatnonzero(rs == RQ_PENDING)
        if not idle.size or not pend.size:
            return 0
        oi, op, od = self.b.npo_match(nid[idle], self._onid[pend])
        if not oi.size:
            return 0
        n = oi.size
        self.apply_codes(idle[oi], np.arange(n + 1, dtype=np.int32), np.full(n, -1, np.int32), self._onid[pend[op]],
                         (np.float64(T) + od) + 240.0)              # rebalancing_npo.py:33-37
        return int(n)

    def epoch(self, T, update_stats: bool, mode: int, cap: int, arrivals=None, rebalance: bool = True) -> dict:
        ev = self.advance(T, update_stats)
        n_new = 0
        if arrivals is not None and len(arrivals[1]):
            self.add_requests(*arrivals)
            n_new = len(arrivals[1])
        n_sel = self.dispatch(mode, cap, T, n_new)
        n_reb = self.rebalance(T) if rebalance else 0
        return dict(events=ev, n_selected=n_sel, n_rebalanced=n_reb)

    # ---- host mirror -----------------------------------------------------------------------------------------------
    def array(self, name: str) -> np.ndarray:
        idx, dt, shape = capi.FLEET_ARRAYS[name]
        V, S, L, R = self.n_veh, self.max_stops, self.max_legs, self.n_req
        dims = {"V": (V,), "8V": (8, V), "VS": (V, S), "VL": (V, L), "R": (R,)}[shape]
        out = np.zeros(dims, dtype=np.dtype(dt))
        self.b._check(self.b.lib.amod_fleet_array(self.b.h, idx, out.ctypes.data, out.nbytes))
        return out

    def export(self) -> dict:
        """Every field the reference's Veh / Req objects hold, as arrays (schedules and legs as CSR over vehicles, stops as the
        reference's (rid, pod, tnid, ddl))."""
        out = {k: self.array(k) for k in ("nid", "lng", "lat", "t_to_nid", "d_to_nid", "tnid", "load", "status", "state_time", "t", "d",
                                          "has_step", "step_t", "step_d", "step_u", "step_v", "sche_len")}
        st = self.array("stats")
        for i, s in enumerate(STATS):
            out[s] = st[i]
        V, S = self.n_veh, self.max_stops
        code = self.array("sche_code"); rn = self.array("reb_node"); rd = self.array("reb_ddl")
        mask = np.arange(S)[None, :] < out["sche_len"][:, None]
        c = code[mask].astype(np.int64)
        veh = np.repeat(np.arange(V), out["sche_len"])
        r = np.maximum(c, 0) >> 1
        is_reb, drop = c < 0, (c & 1) == 1
        if self.n_req:
            node = np.where(drop, self._dnid[r], self._onid[r]); ddl = np.where(drop, self._cld[r], self._clp[r])
        else:
            node = np.zeros(c.size, np.int32); ddl = np.zeros(c.size, np.float64)
        out["sche_rid"] = np.where(is_reb, -1, r).astype(np.int32)
        out["sche_pod"] = np.where(is_reb, 0, np.where(drop, -1, 1)).astype(np.int32)
        out["sche_tnid"] = np.where(is_reb, rn[veh], node).astype(np.int32)
        out["sche_ddl"] = np.where(is_reb, rd[veh], ddl).astype(np.float64)
        out["sche_onboard"] = self.array("sche_onboard")[mask]
        head, n = self.array("leg_head"), self.array("leg_n")
        L = self.max_legs
        lm = (np.arange(L)[None, :] >= head[:, None]) & (np.arange(L)[None, :] < n[:, None])
        out["n_legs"] = (n - head).astype(np.int32)
        for k in ("rid", "pod", "tnid", "ddl", "t", "d"):
            out["leg_" + k] = self.array("leg_" + k)[lm]
        out["leg_n_steps"] = (self.array("leg_se") - self.array("leg_sb"))[lm]
        out["req_status"] = self.array("req_status"); out["req_tp"] = self.array("req_tp"); out["req_td"] = self.array("req_td")
        return out
