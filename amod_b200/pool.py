"""Host-side mirror of the reference's three seam functions on top of the
C-ABI (include/amod_b200.h).  Same names, argument meaning and return records as

  B1  dispatch_osp.compute_candidate_veh_trip_pairs   (dispatch_osp.py:94-129)
  B2  dispatch_sba.compute_candidate_veh_req_pairs    (dispatch_sba.py:44-74)
  B3  rebalancing_npo.reposition_idle_vehicles_to_nearest_pending_orders (rebalancing_npo.py:8-62)

so a maintainer can rebind the module globals (amod_b200/install.py).  All compute is CUDA;
there is no CPU path here — a missing library or device raises RuntimeError.
"""
from __future__ import annotations

import ctypes as C

import operator

import numpy as np

from . import capi, marshal
from .marshal import MODE_OSP, MODE_OSP_NR, MODE_SBA, Epoch, Pool


class AmodError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"amod_b200 error {code}: {msg}")
        self.code = code


class PoolBuilder:
    """One GPU context holding the travel-time table (route_functions.py:14-15) L2-resident."""

    def __init__(self, tt: np.ndarray, device: int = 0, lib_path: str | None = None):
        self.lib = capi.load_library(lib_path)
        tt = np.ascontiguousarray(tt)
        if tt.dtype not in (np.float32, np.float64) or tt.ndim != 2 or tt.shape[0] != tt.shape[1]:
            raise ValueError("travel-time table must be a square float32/float64 array")
        self.n_nodes = tt.shape[0]
        self.tt_dtype = capi.TT_F64 if tt.dtype == np.float64 else capi.TT_F32
        self.device = device
        h = C.c_void_p()
        rc = self.lib.amod_create(C.byref(h), device, self.n_nodes, tt.ctypes.data, self.tt_dtype)
        if rc != capi.OK:
            raise AmodError(rc, self.lib.amod_last_error(None).decode())
        self.h = h
        self.last_stats: dict = {}
        self.lazy_records = False      # seams return LazyPoolRecords (see install(greedy_on_gpu=True))

    def _release_host_bufs(self):
        buf = getattr(self, "_host_buf", None)
        if buf is not None and getattr(self, "h", None):
            self.lib.amod_host_unregister(self.h, buf.ctypes.data)
        self._host_buf = None

    def close(self):
        if getattr(self, "h", None):
            self._release_host_bufs()
            self.lib.amod_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc == capi.E_LEVELS:
            # same exception type the reference raises at dispatch_osp.py:226
            raise IndexError(self.lib.amod_last_error(self.h).decode())
        if rc != capi.OK:
            raise AmodError(rc, self.lib.amod_last_error(self.h).decode())

    def set_graph(self, enable: bool):
        self._check(self.lib.amod_set_graph(self.h, int(bool(enable))))

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.amod_set_stream(self.h, C.c_void_p(cuda_stream)))

    # ---- array level -----------------------------------------------------------------------
    def build(self, ep, loc: int = capi.LOC_HOST, ptr_of=None) -> dict:
        """Build the pool for `ep` (Epoch of numpy arrays, or device tensors with `ptr_of`);
        the pool stays on the device.  Returns the stats dict."""
        s = capi.epoch_struct(ep, ptr_of)
        st = capi.AmodStats()
        fn = self.lib.amod_sba_build if ep.mode == MODE_SBA else self.lib.amod_osp_build
        rc = fn(self.h, C.byref(s), loc, C.byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return self.last_stats

    def pool_sizes(self):
        v = capi.AmodPool()
        self._check(self.lib.amod_pool_view(self.h, C.byref(v)))
        return int(v.n_entries), int(v.n_trip_reqs), int(v.n_stops)

    def pool_view(self) -> capi.AmodPool:
        v = capi.AmodPool()
        self._check(self.lib.amod_pool_view(self.h, C.byref(v)))
        return v

    def fetch(self, reuse: bool = False) -> Pool:
        """Copy the device pool into numpy arrays.  With reuse=True the arrays are views of grow-only
        host buffers owned by the builder (no page faults per epoch); they are overwritten by the next
        fetch, which is fine for the seams (records are materialised immediately)."""
        P, NT, NS = self.pool_sizes()
        n = {"e": P, "e1": P + 1, "t": NT, "s": NS}
        if reuse:
            arrs = packed_pool_views(self, n)
        else:
            arrs = {name: np.empty(max(n[key], 1), dtype=dt) for name, dt, key in capi.POOL_FIELDS}
        o = capi.AmodPool()
        o.cap_entries, o.cap_trip_reqs, o.cap_stops = P, NT, NS
        for name, _dt, _key in capi.POOL_FIELDS:
            setattr(o, name, arrs[name].ctypes.data)
        self._check(self.lib.amod_pool_fetch(self.h, C.byref(o), capi.LOC_PINNED if reuse else capi.LOC_HOST))
        return Pool(stats=dict(self.last_stats), **{name: arrs[name][:n[key]] for name, _dt, key in capi.POOL_FIELDS})

    def build_pool(self, ep: Epoch, reuse: bool = False) -> Pool:
        self.build(ep)
        return self.fetch(reuse)

    def npo_match(self, idle_nid, pend_onid):
        idle_nid = np.ascontiguousarray(idle_nid, dtype=np.int32)
        pend_onid = np.ascontiguousarray(pend_onid, dtype=np.int32)
        m = min(idle_nid.size, pend_onid.size)
        oi, op, od = np.zeros(max(m, 1), np.int32), np.zeros(max(m, 1), np.int32), np.zeros(max(m, 1), np.float64)
        n_out, n_rounds = C.c_int32(0), C.c_int32(0)
        self._check(self.lib.amod_npo_match(self.h, idle_nid.ctypes.data, idle_nid.size, pend_onid.ctypes.data,
                                            pend_onid.size, capi.LOC_HOST, oi.ctypes.data, op.ctypes.data,
                                            od.ctypes.data, C.byref(n_out), C.byref(n_rounds)))
        self.last_npo_rounds = n_rounds.value
        k = n_out.value
        return oi[:k], op[:k], od[:k]

    def greedy_assign(self):
        """greedy_assignment (ilp_assign.py:101-130) over the pool of the last build, scored len(trip) as the
        reference's scorer leaves it.  Returns (order, selected): order[i] = pool entry at position i of the
        stably re-sorted list, selected = positions in that list."""
        P, _nt, _ns = self.pool_sizes()
        order, sel = np.empty(max(P, 1), np.int32), np.empty(max(P, 1), np.int32)
        n_sel, n_rounds = C.c_int32(0), C.c_int32(0)
        self._check(self.lib.amod_greedy_assign(self.h, order.ctypes.data, sel.ctypes.data, C.byref(n_sel), capi.LOC_HOST,
                                                C.byref(n_rounds)))
        self.last_greedy_rounds = n_rounds.value
        return order[:P], sel[:n_sel.value]

    def greedy_assignment(self, veh_trip_pairs):
        """Drop-in for ilp_assign.greedy_assignment when `veh_trip_pairs` is the record list of this builder's
        last build, scored by the reference's scorer (pair[4] = len(trip)): sorts the list in place exactly as
        the reference does (ilp_assign.py:109) and returns the selected indices."""
        if getattr(veh_trip_pairs, "builder_token", None) != getattr(self, "_pool_token", object()):
            raise RuntimeError("greedy_assignment: the pairs are not this builder's last pool")
        lazy = isinstance(veh_trip_pairs, marshal.LazyPoolRecords)
        if lazy:
            if not veh_trip_pairs.scored:
                raise RuntimeError("greedy_assignment: score the pairs first (score_vt_pairs_with_num_of_orders_and_sche_cost)")
        elif any(pair[4] != len(pair[1]) for pair in veh_trip_pairs):
            raise RuntimeError("greedy_assignment: only the reference's default scoring (len(trip)) is supported on the GPU")
        order, sel = self.greedy_assign()
        if lazy:
            veh_trip_pairs.apply_order(order)
        else:
            veh_trip_pairs[:] = [veh_trip_pairs[i] for i in order.tolist()]
        return sel.tolist()

    # ---- the reference's seams ---------------------------------------------------------------
    def compute_candidate_veh_trip_pairs(self, new_received_rids, considered_rids, reqs, vehs, system_time_sec,
                                         cutoff_time_for_a_size_k_trip_search_per_veh_sec, enable_reoptimization,
                                         enable_fast_compute, capacity: int | None = None):
        """B1.  The wall-clock cut-off argument is accepted and ignored: the GPU search is
        exhaustive, i.e. the reference with the cut-off disabled (SURVEY.md hard part 4)."""
        if enable_fast_compute:
            raise NotImplementedError("enable_fast_compute is disabled in the reference (dispatch_osp.py:29)")
        cap = capacity if capacity is not None else _capacity_of(vehs)
        mode = MODE_OSP if enable_reoptimization else MODE_OSP_NR
        ep, req_objs = marshal.marshal_epoch(reqs, vehs, considered_rids, system_time_sec, cap, mode, req_table=self._req_table())
        pool = self.build_pool(ep, reuse=True)
        return self._tag(self._records(pool, ep, vehs, req_objs))

    def _req_table(self):
        t = getattr(self, "_reqs_seen", None)
        if t is None:
            t = self._reqs_seen = marshal.ReqTable()
        return t

    def _records(self, pool, ep, vehs, req_objs):
        if getattr(self, "lazy_records", False):
            return marshal.LazyPoolRecords(pool, ep, vehs, req_objs)
        return marshal.pool_to_records(pool, ep, vehs, req_objs)

    def _tag(self, records):
        self._pool_token = object()
        records.builder_token = self._pool_token
        return records

    def compute_candidate_veh_req_pairs(self, new_received_rids, reqs, vehs, system_time_sec,
                                        capacity: int | None = None):
        """B2."""
        cap = capacity if capacity is not None else _capacity_of(vehs)
        ep, req_objs = marshal.marshal_epoch(reqs, vehs, new_received_rids, system_time_sec, cap, MODE_SBA, req_table=self._req_table())
        pool = self.build_pool(ep, reuse=True)
        return self._tag(self._records(pool, ep, vehs, req_objs))

    def reposition_idle_vehicles_to_nearest_pending_orders(self, reqs, vehs, system_time_sec, num_of_new_reqs,
                                                           value_func, detour_sec: int = 240):
        """B3: candidates, sort and greedy selection on the GPU; veh.build_route stays host
        (rebalancing_npo.py:62)."""
        pending = _with_status(reqs, 1)                      # OrderStatus.PENDING (:13-16), in id order
        idle = _with_status(vehs, marshal.ST_IDLE)           # (:29-31), in fleet order
        if not pending or not idle:
            return
        oi, op, od = self.npo_match(np.fromiter(map(_get_nid, idle), dtype=np.int32, count=len(idle)),
                                    np.fromiter(map(_get_onid, pending), dtype=np.int32, count=len(pending)))
        tt_type = np.float64
        for vi, ri, dt in zip(oi.tolist(), op.tolist(), od):
            req = pending[ri]
            idle[vi].build_route([(-1, 0, req.onid, system_time_sec + tt_type(dt) + detour_sec)])   # (:37,:62)


# order and 64 B alignment of the segments = what amod_pool_fetch / amod_gpool_fetch recognise as "back to back"
_FETCH_ORDER = ("ent_veh", "ent_kind", "ent_nfeas", "ent_cost", "ent_delay", "ent_trip_off", "ent_sche_off", "trip_req", "sche_code")


def packed_pool_views(owner, n: dict) -> dict:
    """Views of ONE grow-only page-locked buffer owned by `owner` (a PoolBuilder or PeerPool), laid out so that the fetch is a
    single DMA: segment g starts where segment g-1 ends, rounded up to 64 B.  Overwritten by the next fetch."""
    builder = getattr(owner, "b", owner)
    spec = {name: (dt, key) for name, dt, key in capi.POOL_FIELDS}
    sizes = [(name, np.dtype(spec[name][0]), int(n[spec[name][1]])) for name in _FETCH_ORDER]
    total = sum((cnt * dt.itemsize + 63) & ~63 for _nm, dt, cnt in sizes) + 64
    buf = getattr(owner, "_host_buf", None)
    if buf is None or buf.size < total:
        if buf is not None:
            builder.lib.amod_host_unregister(builder.h, buf.ctypes.data)
        buf = np.zeros(int(1.5 * total) + 4096, dtype=np.uint8)      # touched once, then page-locked
        builder._check(builder.lib.amod_host_register(builder.h, buf.ctypes.data, buf.nbytes))
        owner._host_buf = buf
    base = (-buf.ctypes.data) % 64          # 64 B aligned start
    out, at = {}, base
    for name, dt, cnt in sizes:
        out[name] = buf[at:at + max(cnt, 1) * dt.itemsize].view(dt) if cnt else np.zeros(1, dtype=dt)
        at += (cnt * dt.itemsize + 63) & ~63
    return out


def _status_value(s) -> int:
    return int(getattr(s, "value", s))


_get_status_value = operator.attrgetter("status.value")
_get_nid = operator.attrgetter("nid")
_get_onid = operator.attrgetter("onid")


def _with_status(objs, value: int) -> list:
    """Objects whose status equals `value`, in order.  The reference scans every request of the day each epoch
    (rebalancing_npo.py:13-16); the scan stays, at C speed."""
    objs = objs if isinstance(objs, list) else list(objs)
    try:
        st = np.fromiter(map(_get_status_value, objs), dtype=np.int64, count=len(objs))
    except AttributeError:      # plain ints instead of the reference's enums
        st = np.fromiter((_status_value(o.status) for o in objs), dtype=np.int64, count=len(objs))
    return [objs[i] for i in np.flatnonzero(st == value).tolist()]


def _capacity_of(vehs) -> int:
    # VEH_CAPACITY[0] (scheduling.py:76) equals every vehicle's K (platform.py:55)
    return int(vehs[0].K)


def score_vt_pairs_with_num_of_orders_and_sche_cost(veh_trip_pairs, reqs, system_time_sec):
    """Drop-in for scheduling.py:140-161 when the pool came from PoolBuilder: pair[3] := delay
    (computed on the GPU in the pool pass), pair[4] := len(trip).  Falls back to nothing: a
    pool without device delays is an error."""
    if isinstance(veh_trip_pairs, marshal.LazyPoolRecords):
        veh_trip_pairs.scored = True       # records are built with pair[3] = delay, pair[4] = len(trip)
        return
    delays = getattr(veh_trip_pairs, "delays", None)
    if delays is None or len(delays) != len(veh_trip_pairs):
        raise RuntimeError("pool was not produced by amod_b200 (no device delays attached)")
    for pair, d in zip(veh_trip_pairs, delays):
        pair[3] = d
        pair[4] = len(pair[1])
