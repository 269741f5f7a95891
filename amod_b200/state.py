"""the epoch state kept in HBM across epochs (include/amod_b200.h amod_state_*).

`DeviceFleet` mirrors, on the host, what the device holds: the request table (indexed by request id; only the requests that
arrived since the last epoch are uploaded) and one row per vehicle (only the rows that changed since the last epoch are
uploaded: a vehicle standing idle costs nothing).  `sync()` takes one epoch's arrays in request-ID space (what
`marshal.marshal_epoch` produces, with its per-epoch request indices mapped back to ids) and uploads the delta;
`build()` runs the pool build from the resident state.  The pool names requests by id.

What is NOT here: the vehicles' state advance itself (vehicle.py:73-162 move_to_time) still runs in the reference's
simulator on the host, so every moving vehicle's row changes every epoch; the delta is the requests plus those rows.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .marshal import MODE_SBA, Epoch, Pool


class DeviceFleet:
    def __init__(self, builder, n_veh: int, max_stops: int, request_capacity: int = 1 << 20):
        self.b = builder
        self.n_veh, self.max_stops = int(n_veh), int(max_stops)
        builder._check(builder.lib.amod_state_create(builder.h, self.n_veh, self.max_stops, int(request_capacity)))
        self.n_req = 0
        z = lambda dt, *shape: np.zeros(shape, dtype=dt)
        V, S = self.n_veh, self.max_stops
        self.m = dict(nid=z(np.int32, V), t=z(np.float64, V), load=z(np.int32, V), status=z(np.int32, V), reb_node=z(np.int32, V),
                      reb_ddl=z(np.float64, V), len=z(np.int32, V), code=z(np.int32, V, S), onboard=z(np.uint8, V, S))
        self.last_upload_bytes = 0
        self.last_changed = 0

    # ---- epoch arrays in request-id space ---------------------------------------------------------------------------------
    @staticmethod
    def rows_of(ep: Epoch, max_stops: int):
        """Fixed-stride vehicle rows of an Epoch, stop codes over request IDS."""
        V = ep.n_veh
        lens = np.diff(ep.veh_sche_off).astype(np.int32)
        if lens.size and int(lens.max()) > max_stops:
            raise ValueError(f"a schedule has {int(lens.max())} stops; the resident state holds {max_stops}")
        code = np.zeros((V, max_stops), dtype=np.int32)
        onb = np.zeros((V, max_stops), dtype=np.uint8)
        if ep.sche_code.size:
            rid = ep.req_id.astype(np.int64)
            c = ep.sche_code.astype(np.int64)
            cid = np.where(c < 0, -1, 2 * rid[np.maximum(c, 0) >> 1] + (c & 1)).astype(np.int32)
            row = np.repeat(np.arange(V), lens)
            col = np.arange(c.size) - np.repeat(ep.veh_sche_off[:-1].astype(np.int64), lens)
            code[row, col] = cid
            onb[row, col] = ep.sche_onboard
        return dict(nid=ep.veh_nid, t=ep.veh_t_to_nid, load=ep.veh_load, status=ep.veh_status, reb_node=ep.veh_reb_node,
                    reb_ddl=ep.veh_reb_ddl, len=lens, code=code, onboard=onb)

    def add_requests(self, rid, onid, dnid, clp, cld, tr, ts):
        """Requests with ids >= the table's size, in id order (request.py:25-41 fields; they never change afterwards)."""
        rid = np.asarray(rid, dtype=np.int64)
        new = rid >= self.n_req
        if not new.any():
            return 0
        r = rid[new]
        order = np.argsort(r)
        r = r[order]
        first, top = int(r[0]), int(r[-1]) + 1
        if first != self.n_req or not np.array_equal(r, np.arange(first, top)):
            # ids missing in between (requests this epoch never saw): the table rows stay zero until they arrive
            full = lambda a, dt: _scatter(np.asarray(a)[new][order], r - self.n_req, top - self.n_req, dt)
            first = self.n_req
        else:
            full = lambda a, dt: np.ascontiguousarray(np.asarray(a)[new][order], dtype=dt)
        cols = [full(onid, np.int32), full(dnid, np.int32), full(clp, np.float64), full(cld, np.float64), full(tr, np.float64), full(ts, np.float64)]
        cols[0][cols[0] == 0] = 1; cols[1][cols[1] == 0] = 1            # (never-seen ids: any valid node; no stop can name them)
        b = self.b
        b._check(b.lib.amod_state_add_requests(b.h, first, top - first, *[c.ctypes.data for c in cols]))
        self.n_req = top
        return sum(c.nbytes for c in cols)

    def sync(self, ep: Epoch) -> int:
        """Upload what changed since the last epoch; returns the bytes that travelled."""
        nbytes = self.add_requests(ep.req_id, ep.req_onid, ep.req_dnid, ep.req_clp, ep.req_cld, ep.req_tr, ep.req_ts)
        rows = self.rows_of(ep, self.max_stops)
        m = self.m
        changed = np.zeros(self.n_veh, dtype=bool)
        for k in ("nid", "t", "load", "status", "reb_node", "reb_ddl", "len"):
            changed |= rows[k] != m[k]
        changed |= (rows["code"] != m["code"]).any(axis=1) | (rows["onboard"] != m["onboard"]).any(axis=1)
        idx = np.flatnonzero(changed).astype(np.int32)
        self.last_changed = int(idx.size)
        if idx.size:
            lens = rows["len"][idx]
            off = np.zeros(idx.size + 1, dtype=np.int32)
            np.cumsum(lens, out=off[1:])
            mask = np.arange(self.max_stops)[None, :] < lens[:, None]
            code = np.ascontiguousarray(rows["code"][idx][mask], dtype=np.int32)
            onb = np.ascontiguousarray(rows["onboard"][idx][mask], dtype=np.uint8)
            cols = [np.ascontiguousarray(rows[k][idx]) for k in ("nid", "t", "load", "status", "reb_node", "reb_ddl")]
            b = self.b
            b._check(b.lib.amod_state_update_vehicles(b.h, idx.size, idx.ctypes.data, *[c.ctypes.data for c in cols], off.ctypes.data,
                                                      code.ctypes.data, onb.ctypes.data))
            nbytes += idx.nbytes + off.nbytes + code.nbytes + onb.nbytes + sum(c.nbytes for c in cols)
            for k in m:
                m[k][idx] = rows[k][idx]
        self.last_upload_bytes = nbytes
        return nbytes

    def build(self, mode: int, cap: int, T: int, search_rids) -> dict:
        b = self.b
        s = np.ascontiguousarray(search_rids, dtype=np.int32)
        st = capi.AmodStats()
        rc = b.lib.amod_state_build(b.h, mode, cap, T, s.ctypes.data, s.size, C.byref(st))
        b.last_stats = st.as_dict()
        b._check(rc)
        return b.last_stats

    def download(self) -> dict:
        V, S = self.n_veh, self.max_stops
        z = lambda dt, *shape: np.zeros(shape, dtype=dt)
        d = dict(nid=z(np.int32, V), t=z(np.float64, V), load=z(np.int32, V), status=z(np.int32, V), reb_node=z(np.int32, V),
                 reb_ddl=z(np.float64, V), len=z(np.int32, V), code=z(np.int32, V, S), onboard=z(np.uint8, V, S))
        b = self.b
        b._check(b.lib.amod_state_download(b.h, *[d[k].ctypes.data for k in ("nid", "t", "load", "status", "reb_node", "reb_ddl", "len", "code", "onboard")]))
        return d


def _scatter(vals, at, n, dt):
    out = np.zeros(n, dtype=dt)
    out[at] = vals
    return out


def pool_in_request_ids(pool: Pool, ep: Epoch) -> Pool:
    """A pool built from an Epoch (per-epoch request indices) rewritten over request ids, for comparison with a pool built
    from the resident state."""
    rid = ep.req_id.astype(np.int64)
    c = pool.sche_code.astype(np.int64)
    code = np.where(c < 0, -1, 2 * rid[np.maximum(c, 0) >> 1] + (c & 1)).astype(np.int32)
    kw = {f: getattr(pool, f) for f in Pool.FIELDS}
    kw["sche_code"] = code
    kw["trip_req"] = rid[pool.trip_req].astype(np.int32) if pool.trip_req.size else pool.trip_req
    return Pool(stats=pool.stats, **kw)
