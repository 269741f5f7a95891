"""batched route construction on the GPU behind the reference's own interface.

The reference builds a vehicle's route leg by leg (vehicle.py:290-302 add_leg -> route_functions.py:36-70
build_route_from_origin_to_dest: a Python walk of the predecessor table per leg); with greedy + OSP every vehicle's
route is rebuilt every epoch, which is the largest cost of a simulated run at small scale.  Here all legs
of the schedules selected in one epoch are built by ONE amod_routes_build call; `RouteBuilder.
build_route_from_origin_to_dest` then serves the reference's unchanged `Veh.build_route` / `add_leg` from that batch with
the reference's own return value: (duration, distance, [(t, d, [u, v], [u_geo, v_geo]), ..., (0.0, 0.0, [tnid, tnid], ...)]).

    rb = RouteBuilder(builder, rf.shortest_path_table, rf.travel_distance_table, node_geo)
    install_routes(rb, vehicle_module, [dispatch_osp, dispatch_sba])
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class Routes:
    """amod_routes arrays: leg g owns steps [leg_off[g], leg_off[g+1])."""

    def __init__(self, leg_off, step_u, step_v, step_t, step_d, leg_t, leg_d):
        self.leg_off, self.step_u, self.step_v, self.step_t, self.step_d, self.leg_t, self.leg_d = leg_off, step_u, step_v, step_t, step_d, leg_t, leg_d

    @property
    def n_legs(self):
        return self.leg_t.size


class RouteBuilder:
    def __init__(self, builder, path_table: np.ndarray, dist_table: np.ndarray, node_geo: np.ndarray | None = None):
        """`builder` is the PoolBuilder whose context holds the travel-time table; the distance table is uploaded in the same
        element type.  node_geo[n-1] = (lng, lat) of node n (route_functions.py:74-76)."""
        self.b = builder
        n = builder.n_nodes
        pred = np.ascontiguousarray(path_table, dtype=np.int32)
        el = np.float64 if builder.tt_dtype == capi.TT_F64 else np.float32
        dist = np.ascontiguousarray(dist_table, dtype=el)
        if pred.shape != (n, n) or dist.shape != (n, n):
            raise ValueError("path / distance tables must match the travel-time table")
        if el == np.float32 and not np.array_equal(dist.astype(np.float64), np.asarray(dist_table, dtype=np.float64)):
            raise ValueError("distance table is not exactly representable in fp32: build the PoolBuilder on float64 tables")
        builder._check(builder.lib.amod_routes_upload(builder.h, pred.ctypes.data, dist.ctypes.data))
        self.node_geo = None if node_geo is None else np.asarray(node_geo, dtype=np.float64)
        self._cache: dict = {}
        self.n_batches = 0

    # ---- array level -----------------------------------------------------------------------------------------------
    def build(self, leg_o, leg_d) -> Routes:
        leg_o = np.ascontiguousarray(leg_o, dtype=np.int32)
        leg_d = np.ascontiguousarray(leg_d, dtype=np.int32)
        L = leg_o.size
        b = self.b
        cap = max(64 * L, 1024)
        while True:
            off = np.zeros(L + 1, np.int64)
            su, sv = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
            st, sd = np.zeros(cap, np.float64), np.zeros(cap, np.float64)
            lt, ld = np.zeros(max(L, 1), np.float64), np.zeros(max(L, 1), np.float64)
            r = capi.AmodRoutes()
            r.cap_steps = cap
            r.leg_off, r.step_u, r.step_v, r.step_t, r.step_d = off.ctypes.data, su.ctypes.data, sv.ctypes.data, st.ctypes.data, sd.ctypes.data
            r.leg_t, r.leg_d = lt.ctypes.data, ld.ctypes.data
            rc = b.lib.amod_routes_build(b.h, leg_o.ctypes.data, leg_d.ctypes.data, L, capi.LOC_HOST, C.byref(r))
            if rc == capi.E_CAPACITY:
                cap = int(r.n_steps) + 1024
                continue
            b._check(rc)
            n = int(r.n_steps)
            self.n_batches += 1
            return Routes(off, su[:n], sv[:n], st[:n], sd[:n], lt[:L], ld[:L])

    # ---- the reference's interface ---------------------------------------------------------------------------------
    def _geo(self, nid: int):
        g = self.node_geo[nid - 1]
        return [float(g[0]), float(g[1])]

    def _to_reference(self, r: Routes, g: int):
        a, z = int(r.leg_off[g]), int(r.leg_off[g + 1])
        u, v = r.step_u[a:z].tolist(), r.step_v[a:z].tolist()
        t, d = r.step_t[a:z], r.step_d[a:z]
        steps = [(t[k], d[k], [u[k], v[k]], [self._geo(u[k]), self._geo(v[k])]) for k in range(z - a - 1)]
        tn = v[-1]
        tg = self._geo(tn)
        steps.append((0.0, 0.0, [tn, tn], [tg, list(tg)]))
        return r.leg_t[g], r.leg_d[g], steps

    def prefetch(self, pairs):
        """Build every (origin, destination) leg of `pairs` not yet cached in ONE GPU batch."""
        todo = sorted({(int(o), int(d)) for o, d in pairs} - self._cache.keys())
        if not todo:
            return
        r = self.build([p[0] for p in todo], [p[1] for p in todo])
        for g, key in enumerate(todo):
            self._cache[key] = (r, g)

    def clear(self):
        self._cache.clear()

    def build_route_from_origin_to_dest(self, onid: int, dnid: int):
        """Drop-in for route_functions.build_route_from_origin_to_dest (:36-70)."""
        key = (int(onid), int(dnid))
        if key not in self._cache:
            self.prefetch([key])
        r, g = self._cache[key]
        return self._to_reference(r, g)


def legs_of_selected_pairs(veh_trip_pairs, selected_indices):
    """(origin, destination) of every leg Veh.build_route will ask for when the selected schedules are applied
    (scheduling.py:110-119 -> vehicle.py:231-302): the chain starts at the end of the unfinished step if the vehicle is between
    two nodes (vehicle.py:247-266), else at its node; a rebalancing stop is dropped when a trip is assigned (:234-241)."""
    pairs = []
    for idx in selected_indices:
        veh, _trip, sche = veh_trip_pairs[idx][0], veh_trip_pairs[idx][1], veh_trip_pairs[idx][2]
        stops = list(sche)
        if getattr(getattr(veh, "status", None), "name", "") == "REBALANCING" and len(stops) > 1:
            for k, s in enumerate(stops):
                if s[0] == -1:
                    stops.pop(k)
                    break
        cur = veh.step_to_nid.nid_pair[1] if getattr(veh, "step_to_nid", None) else veh.nid
        for (_rid, _pod, tnid, _ddl) in stops:
            pairs.append((cur, tnid))
            cur = tnid
    return pairs


def install_routes(rb: RouteBuilder, vehicle_module, dispatch_modules=()):
    """vehicle.py looks build_route_from_origin_to_dest up in its own module (star import, vehicle.py:290-291): rebind it
    there; wrap the dispatchers' apply step so that all legs of an epoch travel to the GPU in one batch."""
    vehicle_module.build_route_from_origin_to_dest = rb.build_route_from_origin_to_dest
    for mod in dispatch_modules:
        if mod is None:
            continue
        orig = mod.upd_schedule_for_vehicles_in_selected_vt_pairs

        def upd(veh_trip_pairs, selected_veh_trip_pair_indices, _orig=orig):
            rb.clear()
            rb.prefetch(legs_of_selected_pairs(veh_trip_pairs, selected_veh_trip_pair_indices))
            return _orig(veh_trip_pairs, selected_veh_trip_pair_indices)

        mod.upd_schedule_for_vehicles_in_selected_vt_pairs = upd
    return rb
