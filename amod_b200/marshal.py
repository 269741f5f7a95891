"""Host-side marshalling between the reference's Python records and the SoA epoch
arrays that cross the C-ABI (include/amod_b200.h, struct amod_epoch / amod_pool).

Record layouts mirrored here:
  stop       (rid, pod, tnid, ddl)                    types.py:66-69, scheduling.py:56-57
  Veh fields id, nid, t_to_nid, load, status, sche, onboard_rids   vehicle.py:41-70
  Req fields id, onid, dnid, Clp, Cld, Tr, Ts, status              request.py:25-41
  pool entry [veh, trip, sche, cost, score]           dispatch_osp.py:105

Stop code (int32): 2*req_index + (0 pick-up | 1 drop-off), or -1 for the rebalancing
stop (rid=-1, pod=0) whose node/deadline are per-vehicle (rebalancing_npo.py:37).
"""
from __future__ import annotations

import gc
import itertools
import operator

import numpy as np

MODE_OSP = 0      # enable_reoptimization=True   dispatch_osp.py:301-322
MODE_OSP_NR = 1   # enable_reoptimization=False  dispatch_osp.py:301-303
MODE_SBA = 2      # dispatch_sba.py:44-74

KIND_BASIC, KIND_TRIP, KIND_CURRENT, KIND_SBA_PAIR, KIND_SBA_EMPTY = 0, 1, 2, 3, 4

ST_IDLE, ST_WORKING, ST_REBALANCING = 1, 2, 3  # types.py:18-21


def _status_value(s) -> int:
    return int(getattr(s, "value", s))


_get_sche = operator.attrgetter("sche")
_get_onboard = operator.attrgetter("onboard_rids")
_get_veh3 = operator.attrgetter("nid", "t_to_nid", "load")
_get_status_value = operator.attrgetter("status.value")
_chain = itertools.chain.from_iterable


class Epoch:
    """SoA view of one epoch's seam inputs.  All arrays are C-contiguous numpy."""

    FIELDS = ("veh_nid", "veh_t_to_nid", "veh_load", "veh_status", "veh_sche_off", "sche_code",
              "sche_onboard", "veh_reb_node", "veh_reb_ddl", "req_id", "req_onid", "req_dnid",
              "req_clp", "req_cld", "req_tr", "req_ts", "search")

    def __init__(self, **kw):
        self.T = int(kw.pop("T"))
        self.cap = int(kw.pop("cap"))
        self.mode = int(kw.pop("mode"))
        for f in self.FIELDS:
            setattr(self, f, kw.pop(f))
        assert not kw, kw
        self.canon()

    def canon(self):
        dt = dict(veh_nid=np.int32, veh_t_to_nid=np.float64, veh_load=np.int32, veh_status=np.int32,
                  veh_sche_off=np.int32, sche_code=np.int32, sche_onboard=np.uint8, veh_reb_node=np.int32,
                  veh_reb_ddl=np.float64, req_id=np.int32, req_onid=np.int32, req_dnid=np.int32,
                  req_clp=np.float64, req_cld=np.float64, req_tr=np.float64, req_ts=np.float64,
                  search=np.int32)
        for f, d in dt.items():
            setattr(self, f, np.ascontiguousarray(getattr(self, f), dtype=d))

    @property
    def n_veh(self):
        return self.veh_nid.size

    @property
    def n_req(self):
        return self.req_onid.size

    @property
    def n_search(self):
        return self.search.size

    def to_npz_dict(self, prefix=""):
        d = {prefix + f: getattr(self, f) for f in self.FIELDS}
        d[prefix + "scalars"] = np.array([self.T, self.cap, self.mode], dtype=np.int64)
        return d

    @classmethod
    def from_npz_dict(cls, d, prefix=""):
        T, cap, mode = (int(x) for x in d[prefix + "scalars"])
        return cls(T=T, cap=cap, mode=mode, **{f: d[prefix + f] for f in cls.FIELDS})

    def take(self, sel: np.ndarray) -> "Epoch":
        """The epoch restricted to vehicles `sel` (in that order).  Requests are kept as they are."""
        sel = np.asarray(sel, dtype=np.int64)
        lens = np.diff(self.veh_sche_off)[sel]
        off = np.zeros(sel.size + 1, dtype=np.int32)
        np.cumsum(lens, out=off[1:])
        # stop indices of the selected vehicles, vehicle by vehicle (vectorised ragged gather)
        starts = np.repeat(self.veh_sche_off[:-1][sel].astype(np.int64) - off[:-1], lens)
        idx = starts + np.arange(int(off[-1]), dtype=np.int64)
        kw = {f: getattr(self, f) for f in self.FIELDS}
        for f in ("veh_nid", "veh_t_to_nid", "veh_load", "veh_status", "veh_reb_node", "veh_reb_ddl"):
            kw[f] = kw[f][sel]
        kw["veh_sche_off"] = off
        kw["sche_code"] = self.sche_code[idx]
        kw["sche_onboard"] = self.sche_onboard[idx]
        return Epoch(T=self.T, cap=self.cap, mode=self.mode, **kw)

    def shard(self, rank: int, world: int) -> tuple["Epoch", np.ndarray]:
        """Vehicles rank, rank+world, ... (interleaved: vehicle work is heavy-tailed and
        spatially correlated with fleet order, platform.py:51-55).  Requests are replicated."""
        sel = np.arange(rank, self.n_veh, world, dtype=np.int64)
        return self.take(sel), sel


class ReqTable:
    """Per-simulation table of the immutable request fields, filled the first time a request id is seen, so that
    an epoch only reads the ~10^2 new request objects instead of all ~10^4 live ones.  Bound to one `reqs`
    container (the platform's list, platform.py:60-65); a different container starts a new table."""

    FIELDS = (("onid", np.int32), ("dnid", np.int32), ("Clp", np.float64), ("Cld", np.float64), ("Tr", np.float64), ("Ts", np.float64))

    def __init__(self):
        self._container = None
        self._known = np.zeros(0, dtype=bool)
        self._cols = [np.zeros(0, dtype=dt) for _f, dt in self.FIELDS]

    def gather(self, reqs, rids: np.ndarray, req_objs: list):
        if reqs is not self._container:
            self.__init__()
            self._container = reqs
        top = int(rids[-1]) + 1 if rids.size else 0
        if top > self._known.size:
            n = max(top, 2 * self._known.size, 1024)
            self._known = np.concatenate([self._known, np.zeros(n - self._known.size, dtype=bool)])
            self._cols = [np.concatenate([c, np.zeros(n - c.size, dtype=c.dtype)]) for c in self._cols]
        new = np.flatnonzero(~self._known[rids])
        if new.size:
            objs = [req_objs[k] for k in new.tolist()]
            at = rids[new]
            for c, (f, dt) in zip(self._cols, self.FIELDS):
                c[at] = np.fromiter((getattr(r, f) for r in objs), dtype=dt, count=len(objs))
            self._known[at] = True
        return tuple(c[rids] for c in self._cols)


def marshal_epoch(reqs, vehs, search_rids, T: int, cap: int, mode: int, check: bool = True, req_table: ReqTable | None = None):
    """Veh/Req objects -> Epoch.  `search_rids` is considered_rids (OSP, dispatch_osp.py:33-39)
    or new_received_rids (SBA, dispatch_sba.py:56).  Returns (epoch, req_objs) where
    req_objs[i] is the Req behind request index i.  One pass over the objects, the rest is numpy."""
    nv = len(vehs)
    # attribute reads at C speed (operator.attrgetter / map / chain): this loop over ~10^3-10^4 Python objects is
    # what the seam costs per epoch once the kernels take < 1 ms
    sches = list(map(_get_sche, vehs))
    lens = np.fromiter(map(len, sches), dtype=np.int64, count=nv)
    off = np.zeros(nv + 1, dtype=np.int32)
    np.cumsum(lens, out=off[1:])
    ns = int(off[-1])
    flat = list(_chain(sches))
    stops = np.fromiter(_chain(flat), dtype=np.float64, count=4 * ns).reshape(ns, 4) if ns else np.zeros((0, 4))
    s_rid = stops[:, 0].astype(np.int64)
    s_pod = stops[:, 1].astype(np.int64)
    s_node = stops[:, 2].astype(np.int64)
    s_ddl = stops[:, 3]
    s_veh = np.repeat(np.arange(nv, dtype=np.int64), lens)
    is_reb = s_rid < 0
    search_arr = np.asarray(search_rids, dtype=np.int64).reshape(-1)
    onbs = list(map(_get_onboard, vehs))
    onb_len = np.fromiter(map(len, onbs), dtype=np.int64, count=nv)
    onb_rid = np.fromiter(_chain(onbs), dtype=np.int64, count=int(onb_len.sum()))
    onb_veh = np.repeat(np.arange(nv, dtype=np.int64), onb_len)
    vals = np.concatenate([search_arr, s_rid[~is_reb]])
    top = int(max(vals.max() if vals.size else -1, onb_rid.max() if onb_rid.size else -1)) + 1
    dense = 0 < top <= (1 << 22) and (vals.size == 0 or int(vals.min()) >= 0)    # request ids index the platform's list: direct tables
    if dense:
        present = np.zeros(top, dtype=bool)
        present[vals] = True
        rids = np.flatnonzero(present)
    else:
        rids = np.unique(vals)
    req_objs = [reqs[r] for r in rids.tolist()]
    nr = len(req_objs)
    if req_table is not None:       # a request's (onid, dnid, Clp, Cld, Tr, Ts) never change (request.py:25-41): read each object once
        req_onid, req_dnid, req_clp, req_cld, req_tr, req_ts = req_table.gather(reqs, rids, req_objs)
    else:
        req_onid = np.fromiter((r.onid for r in req_objs), dtype=np.int32, count=nr)
        req_dnid = np.fromiter((r.dnid for r in req_objs), dtype=np.int32, count=nr)
        req_clp = np.fromiter((r.Clp for r in req_objs), dtype=np.float64, count=nr)
        req_cld = np.fromiter((r.Cld for r in req_objs), dtype=np.float64, count=nr)
        req_tr = np.fromiter((r.Tr for r in req_objs), dtype=np.float64, count=nr)
        req_ts = np.fromiter((r.Ts for r in req_objs), dtype=np.float64, count=nr)
    s_rid0 = np.where(is_reb, rids[0] if nr else 0, s_rid)
    if dense:
        index_of = np.full(top, -1, dtype=np.int64)
        index_of[rids] = np.arange(nr, dtype=np.int64)
        search = index_of[search_arr].astype(np.int32)
        idx = index_of[s_rid0] if ns else np.zeros(0, dtype=np.int64)
        # onboard flag: (vehicle, rid) pairs of veh.onboard_rids; a request is on board of at most one vehicle
        veh_of = np.full(top, -1, dtype=np.int64)
        veh_of[onb_rid] = onb_veh
        onboard = ((veh_of[s_rid0] == s_veh) & ~is_reb).astype(np.uint8) if ns else np.zeros(0, dtype=np.uint8)
    else:
        search = np.searchsorted(rids, search_arr).astype(np.int32)
        idx = np.searchsorted(rids, s_rid0) if ns else np.zeros(0, dtype=np.int64)
        big = int(max(rids[-1] if nr else 0, onb_rid.max() if onb_rid.size else 0)) + 2
        onboard = (np.isin(s_veh * big + s_rid, onb_veh * big + onb_rid) & ~is_reb).astype(np.uint8) if ns else np.zeros(0, dtype=np.uint8)
    code = np.where(is_reb, -1, 2 * idx + (s_pod == -1)).astype(np.int32)
    reb_node = np.zeros(nv, dtype=np.int32)
    reb_ddl = np.zeros(nv, dtype=np.float64)
    reb_node[s_veh[is_reb]] = s_node[is_reb]
    reb_ddl[s_veh[is_reb]] = s_ddl[is_reb]
    if check and ns:
        bad_reb = is_reb & ~((s_rid == -1) & (s_pod == 0))
        if bad_reb.any():
            k = int(np.flatnonzero(bad_reb)[0])
            raise ValueError(f"vehicle {int(s_veh[k])}: unexpected stop {flat[k]}")
        req = ~is_reb
        pick = s_pod == 1
        exp_node = np.where(pick, req_onid[idx], req_dnid[idx])
        exp_ddl = np.where(pick, req_clp[idx], req_cld[idx])
        bad = req & (~np.isin(s_pod, (1, -1)) | (s_node != exp_node) | (s_ddl != exp_ddl))
        if bad.any():
            k = int(np.flatnonzero(bad)[0])
            raise ValueError(f"vehicle {int(s_veh[k])}: stop {flat[k]} does not match its request")
    vcols = np.fromiter(_chain(map(_get_veh3, vehs)), dtype=np.float64, count=3 * nv).reshape(nv, 3)   # (nid, t_to_nid, load): exact in float64
    veh_nid = vcols[:, 0].astype(np.int32)
    veh_t = np.ascontiguousarray(vcols[:, 1])
    veh_load = vcols[:, 2].astype(np.int32)
    try:
        veh_status = np.fromiter(map(_get_status_value, vehs), dtype=np.int32, count=nv)
    except AttributeError:          # plain ints instead of the reference's enum
        veh_status = np.fromiter((_status_value(v.status) for v in vehs), dtype=np.int32, count=nv)
    ep = Epoch(T=T, cap=cap, mode=mode, veh_nid=veh_nid, veh_t_to_nid=veh_t, veh_load=veh_load,
               veh_status=veh_status, veh_sche_off=off, sche_code=code, sche_onboard=onboard,
               veh_reb_node=reb_node, veh_reb_ddl=reb_ddl, req_id=rids.astype(np.int32), req_onid=req_onid,
               req_dnid=req_dnid, req_clp=req_clp, req_cld=req_cld, req_tr=req_tr, req_ts=req_ts,
               search=search)
    return ep, req_objs


class Pool:
    """SoA pool as returned over the C-ABI (struct amod_pool)."""

    FIELDS = ("ent_veh", "ent_kind", "ent_trip_off", "ent_sche_off", "ent_cost", "ent_delay",
              "ent_nfeas", "trip_req", "sche_code")

    def __init__(self, **kw):
        self.stats = dict(kw.pop("stats", {}))
        for f in self.FIELDS:
            setattr(self, f, kw.pop(f))
        assert not kw, kw

    @property
    def n_entries(self):
        return self.ent_veh.size

    def to_npz_dict(self, prefix=""):
        return {prefix + f: getattr(self, f) for f in self.FIELDS}

    @classmethod
    def from_npz_dict(cls, d, prefix=""):
        return cls(**{f: d[prefix + f] for f in cls.FIELDS})

    def take(self, idx: np.ndarray) -> "Pool":
        """Entries `idx` in that order (re-packs the two CSR payloads)."""
        idx = np.asarray(idx, dtype=np.int64)

        def regather(off, data):
            lens = (off[1:] - off[:-1])[idx]
            new_off = np.zeros(idx.size + 1, dtype=np.int64)
            np.cumsum(lens, out=new_off[1:])
            src = np.repeat(off[:-1][idx] - new_off[:-1], lens) + np.arange(new_off[-1])
            return new_off, data[src]

        toff, treq = regather(self.ent_trip_off.astype(np.int64), self.trip_req)
        soff, scode = regather(self.ent_sche_off.astype(np.int64), self.sche_code)
        return Pool(ent_veh=self.ent_veh[idx], ent_kind=self.ent_kind[idx], ent_trip_off=toff,
                    ent_sche_off=soff, ent_cost=self.ent_cost[idx], ent_delay=self.ent_delay[idx],
                    ent_nfeas=self.ent_nfeas[idx], trip_req=treq, sche_code=scode, stats=self.stats)


def concat_pools(pools: list["Pool"]) -> "Pool":
    def cat_off(name):
        outs, base = [np.zeros(1, dtype=np.int64)], 0
        for p in pools:
            o = getattr(p, name).astype(np.int64)
            outs.append(o[1:] + base)
            base += int(o[-1])
        return np.concatenate(outs)

    kw = {f: np.concatenate([getattr(p, f) for p in pools]) for f in Pool.FIELDS
          if f not in ("ent_trip_off", "ent_sche_off")}
    kw["ent_trip_off"] = cat_off("ent_trip_off")
    kw["ent_sche_off"] = cat_off("ent_sche_off")
    stats = {}
    for p in pools:
        for k, v in p.stats.items():
            if isinstance(v, (list, tuple)):      # per-level counters
                stats[k] = [a + b for a, b in zip(stats.get(k, [0] * len(v)), v)]
            else:
                stats[k] = stats.get(k, 0) + v
    return Pool(stats=stats, **kw)


def records_to_pool(records, vehs, index_of_rid: dict, mode: int) -> Pool:
    """Reference pool records -> Pool arrays (used to store golden vectors).  `cost` is
    record[3] as returned by the seam, i.e. before the scorer overwrites it
    (scheduling.py:145); delay/nfeas are filled by the caller if known."""
    veh_index = {id(v): i for i, v in enumerate(vehs)}
    n = len(records)
    ent_veh = np.empty(n, dtype=np.int32)
    ent_kind = np.zeros(n, dtype=np.int32)
    ent_cost = np.empty(n, dtype=np.float64)
    toff = np.zeros(n + 1, dtype=np.int64)
    soff = np.zeros(n + 1, dtype=np.int64)
    treq, scode = [], []
    for e, (veh, trip, sche, cost, _score) in enumerate(records):
        ent_veh[e] = veh_index[id(veh)]
        ent_cost[e] = cost
        treq.extend(index_of_rid[r.id] for r in trip)
        for (rid, pod, _tnid, _ddl) in sche:
            scode.append(-1 if rid < 0 else 2 * index_of_rid[rid] + (1 if pod == -1 else 0))
        toff[e + 1] = len(treq)
        soff[e + 1] = len(scode)
    return Pool(ent_veh=ent_veh, ent_kind=ent_kind, ent_trip_off=toff, ent_sche_off=soff,
                ent_cost=ent_cost, ent_delay=np.zeros(n), ent_nfeas=np.zeros(n, dtype=np.int32),
                trip_req=np.asarray(treq, dtype=np.int32), sche_code=np.asarray(scode, dtype=np.int32))


class PoolRecords(list):
    """The list the seams return: plain 5-element lists [veh, trip, sche, cost, score]
    (consumers index and overwrite pair[3], pair[4]: scheduling.py:145,161), plus the
    device-computed delays so the scorer pass can be skipped."""
    delays: np.ndarray | None = None
    stats: dict | None = None
    builder_token: object | None = None


def pool_to_records(pool: Pool, ep: Epoch, vehs, req_objs) -> PoolRecords:
    """Pool arrays -> the reference's record list with live Veh/Req references and a
    fresh mutable list of 4-tuples per schedule (vehicle.py:234-245 pops from it).
    Vectorised: one object-array gather turns all stop codes into tuples, then plain slicing."""
    n_req = len(req_objs)
    n = pool.n_entries
    out = PoolRecords()
    # ~10^5 container objects are born here and none of them is garbage: without this the cyclic collector runs
    # a generation pass every few hundred allocations and takes 3/4 of the time
    gc_was_on = gc.isenabled()
    gc.disable()
    try:
        _fill_records(out, pool, ep, vehs, req_objs, n_req, n)
    finally:
        if gc_was_on:
            gc.enable()
    out.delays = pool.ent_delay.copy()
    out.stats = pool.stats
    return out


def _fill_records(out, pool, ep, vehs, req_objs, n_req, n):
    if n:
        # stop tuples: [pick-up of request i | drop-off of request i | rebalancing stop of vehicle v]
        lut = np.empty(2 * n_req + len(vehs), dtype=object)
        for i, r in enumerate(req_objs):
            lut[2 * i] = (r.id, 1, r.onid, r.Clp)
            lut[2 * i + 1] = (r.id, -1, r.dnid, r.Cld)
        soff = pool.ent_sche_off.astype(np.int64)
        scode = pool.sche_code.astype(np.int64)
        reb = np.flatnonzero(scode < 0)
        if reb.size:
            stop_ent = np.searchsorted(soff, reb, side="right") - 1        # entry of each rebalancing stop
            veh_of = pool.ent_veh[stop_ent]
            for v in np.unique(veh_of).tolist():
                lut[2 * n_req + v] = next(s for s in vehs[v].sche if s[0] == -1)
            scode = scode.copy()
            scode[reb] = 2 * n_req + veh_of
        stops = lut[scode].tolist()
        req_lut = np.empty(max(n_req, 1), dtype=object)
        for i, r in enumerate(req_objs):
            req_lut[i] = r
        trips = req_lut[pool.trip_req].tolist() if pool.trip_req.size else []
        veh_lut = np.empty(len(vehs), dtype=object)
        for i, v in enumerate(vehs):
            veh_lut[i] = v
        ent_vehs = veh_lut[pool.ent_veh].tolist()
        so = soff.tolist()
        to = pool.ent_trip_off.tolist()
        cost = pool.ent_cost.copy()      # numpy float64 scalars like the reference; the arrays may be views of reused buffers
        out.extend([ent_vehs[e], trips[to[e]:to[e + 1]], stops[so[e]:so[e + 1]], cost[e], 0.0] for e in range(n))


class LazyPoolRecords:
    """Sequence view of the pool that materialises a record only when it is indexed.  For the pipeline
    scorer(GPU delays) -> greedy assignment on the GPU -> upd_schedule_for_vehicles_in_selected_vt_pairs
    (scheduling.py:110-119) only the <= n_veh selected pairs are ever touched, so the 10^4-10^6 other
    records are never built.  Any other consumer (iteration, sort) still works: it materialises everything."""

    def __init__(self, pool: Pool, ep: Epoch, vehs, req_objs):
        self._pool = Pool(stats=pool.stats, **{f: np.array(getattr(pool, f)) for f in Pool.FIELDS})   # own copy
        self._vehs, self._reqs = vehs, req_objs
        self._req_id = np.asarray(ep.req_id, dtype=np.int64)     # request index -> request id (array-level consumers)
        self._perm = None              # position -> pool entry after an in-place "sort"
        self._cache = {}
        self._full = None
        self.delays = self._pool.ent_delay
        self.stats = pool.stats
        self.builder_token = None
        self._scored = False
        self.user_sorted = False       # list.sort() by a consumer: the GPU assigner's permutation no longer applies

    @property
    def scored(self) -> bool:
        return self._scored

    @scored.setter
    def scored(self, value: bool):
        """The scorer (scheduling.py:140-161) overwrites pair[3] with the delay and pair[4] with len(trip): records
        built before the flag flips carry cost / 0.0 and are rewritten in place (identity of the lists is kept)."""
        value = bool(value)
        if value and not self._scored:
            p = self._pool
            for e, rec in self._cache.items():      # every record ever built lives in the cache, keyed by pool entry
                rec[3] = p.ent_delay[e]
                rec[4] = len(rec[1])
        self._scored = value

    def __len__(self):
        return self._pool.n_entries

    def _entry(self, e: int):
        rec = self._cache.get(e)
        if rec is None:
            p = self._pool
            veh = self._vehs[int(p.ent_veh[e])]
            sche = []
            for c in p.sche_code[p.ent_sche_off[e]:p.ent_sche_off[e + 1]].tolist():
                if c < 0:
                    sche.append(next(s for s in veh.sche if s[0] == -1))
                else:
                    r = self._reqs[c >> 1]
                    sche.append((r.id, -1, r.dnid, r.Cld) if c & 1 else (r.id, 1, r.onid, r.Clp))
            trip = [self._reqs[i] for i in p.trip_req[p.ent_trip_off[e]:p.ent_trip_off[e + 1]].tolist()]
            if self.scored:      # what the reference's scorer leaves behind (scheduling.py:145,161)
                rec = [veh, trip, sche, p.ent_delay[e], len(trip)]
            else:
                rec = [veh, trip, sche, p.ent_cost[e], 0.0]
            self._cache[e] = rec
        return rec

    def __getitem__(self, i):
        if self._full is not None:
            return self._full[i]
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return self._entry(int(self._perm[i]) if self._perm is not None else i)

    def apply_order(self, order: np.ndarray):
        """The in-place stable sort the reference's assigner performs (ilp_assign.py:109), as a permutation:
        order[i] = pool entry (in the builder's order) that stands at position i of the sorted list.  If something
        already iterated the list (records materialised), the materialised list is permuted in place as well, so
        positions mean "index in the sorted list" either way."""
        if self.user_sorted:
            raise RuntimeError("the pair list was re-sorted by the caller: positions no longer refer to the builder's pool order")
        order = np.asarray(order)
        if self._full is not None:
            self._full[:] = [self._cache[e] for e in order.tolist()]
        self._perm = order

    def materialise(self) -> list:
        if self._full is None:
            self._full = [self[i] for i in range(len(self))]
        return self._full

    def __iter__(self):
        return iter(self.materialise())

    def sort(self, key=None, reverse=False):
        self.materialise().sort(key=key, reverse=reverse)
        self.user_sorted = True


def pool_of_vehicles(pool: Pool, step: int, lo: int = 0) -> Pool:
    """The entries of vehicles lo, lo+step, ... in pool order (vehicles are independent, dispatch_osp.py:107-111, so a
    vehicle sample of the pool can be checked against the oracle's build of just those vehicles)."""
    if step <= 1 and lo == 0:
        return pool
    v = pool.ent_veh.astype(np.int64)
    return pool.take(np.flatnonzero((v >= lo) & ((v - lo) % step == 0)))


def pool_digest(pool: Pool) -> str:
    """Order-sensitive fingerprint of a pool: entry count, vehicle / kind / CSR offsets / trip requests / stop
    codes, and the float64 BITS of costs and delays.  Two pools share a digest iff they are bit-identical in
    every field the parity tests compare (pools_equal with rtol=0, check_kind=True)."""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    h.update(np.int64(pool.n_entries).tobytes())
    for f, dt in (("ent_veh", "<i4"), ("ent_kind", "<i4"), ("ent_trip_off", "<i8"), ("ent_sche_off", "<i8"), ("trip_req", "<i4"),
                  ("sche_code", "<i4"), ("ent_nfeas", "<i4"), ("ent_cost", "<f8"), ("ent_delay", "<f8")):
        a = np.ascontiguousarray(getattr(pool, f), dtype=np.dtype(dt))
        h.update(f.encode())
        h.update(np.int64(a.size).tobytes())
        h.update(a.tobytes())
    return h.hexdigest()


def pools_equal(a: Pool, b: Pool, rtol: float = 0.0, check_delay: bool = True, check_nfeas: bool = True,
                check_kind: bool = False) -> tuple[bool, str]:
    """Bit-exact structure; cost/delay within rtol (0.0 = bit-exact)."""
    if a.n_entries != b.n_entries:
        return False, f"n_entries {a.n_entries} != {b.n_entries}"
    for f in ("ent_veh", "ent_trip_off", "ent_sche_off", "trip_req", "sche_code"):
        x, y = getattr(a, f).astype(np.int64), getattr(b, f).astype(np.int64)
        if x.shape != y.shape or not np.array_equal(x, y):
            bad = np.flatnonzero(x[:min(x.size, y.size)] != y[:min(x.size, y.size)])
            return False, f"{f} differs (first at {bad[:1]}; sizes {x.size},{y.size})"
    if check_kind and not np.array_equal(a.ent_kind, b.ent_kind):
        return False, "ent_kind differs"
    if check_nfeas and not np.array_equal(a.ent_nfeas, b.ent_nfeas):
        bad = np.flatnonzero(a.ent_nfeas != b.ent_nfeas)
        return False, f"ent_nfeas differs at {bad[:5]}"
    names = ("ent_cost", "ent_delay") if check_delay else ("ent_cost",)
    for f in names:
        x, y = getattr(a, f), getattr(b, f)
        ok = (x == y) if rtol == 0.0 else (np.abs(x - y) <= rtol * np.maximum(np.abs(x), np.abs(y)))
        if not ok.all():
            bad = np.flatnonzero(~ok)
            return False, f"{f} differs at {bad[:5]}: {x[bad[:5]]} vs {y[bad[:5]]}"
    return True, ""
