"""Synthetic Manhattan-sized inputs for the OSP hot path.

Nothing here reads the reference.  The road graph is a jittered directed grid whose
edge times are multiples of 1/8 s, so every shortest travel time is exact in fp32
and in fp64 and the table is bit-reproducible on any host (the reference keeps the
table as a float64 ndarray, route_functions.py:14-15; we hand it the exact upcast).
"""
from __future__ import annotations

import numpy as np

# Map box of the reference (config.py:71-79).
MAP_WIDTH_KM = 10.71
MAP_HEIGHT_KM = 20.85
OLNG, OLAT, DLNG, DLAT = -74.0300, 40.6950, -73.9030, 40.8825
N_NODES_MANHATTAN = 4091  # value_function.py:14


class RoadGraph:
    """All-pairs tables in the reference's file formats (route_functions.py:8-17)."""

    def __init__(self, tt: np.ndarray, dist: np.ndarray | None, pred: np.ndarray | None,
                 lng: np.ndarray, lat: np.ndarray, stations: np.ndarray):
        self.tt = tt            # float32 [N,N]  row = origin, col = destination, seconds
        self.dist = dist        # float32 [N,N]  metres (or None)
        self.pred = pred        # int32   [N,N]  1-based predecessor of d on path o->d, 0 at o==d (or None)
        self.lng = lng
        self.lat = lat
        self.stations = stations  # 1-based node ids

    @property
    def n_nodes(self) -> int:
        return self.tt.shape[0]


def make_road_graph(n_nodes: int = N_NODES_MANHATTAN, seed: int = 0, nx: int | None = None,
                    with_paths: bool = False, n_stations: int = 101, edge_times: tuple | None = None) -> RoadGraph:
    """Directed 4-neighbour grid, per-direction jittered edge times quantised to 1/8 s.  `edge_times`: draw every edge's time
    from these exact values instead (e.g. (10, 15, 30): paths then end exactly on epoch boundaries, the tie cases of
    vehicle.py:104-127)."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra

    if nx is None:
        nx = max(2, int(round(np.sqrt(n_nodes * MAP_WIDTH_KM / MAP_HEIGHT_KM))))
    ny = (n_nodes + nx - 1) // nx
    rng = np.random.default_rng(seed)
    idx = np.arange(n_nodes)
    gx, gy = idx % nx, idx // nx
    dx_m = MAP_WIDTH_KM * 1000.0 / max(nx - 1, 1)
    dy_m = MAP_HEIGHT_KM * 1000.0 / max(ny - 1, 1)
    src, dst, length = [], [], []
    right = idx[(gx + 1 < nx) & (idx + 1 < n_nodes)]
    up = idx[idx + nx < n_nodes]
    for a, b, l in ((right, right + 1, dx_m), (up, up + nx, dy_m)):
        src += [a, b]
        dst += [b, a]
        length += [np.full(a.size, l), np.full(a.size, l)]
    src, dst, length = np.concatenate(src), np.concatenate(dst), np.concatenate(length)
    speed = 5.5 * rng.uniform(0.7, 1.3, size=src.size)         # m/s, per directed edge
    t_edge = np.maximum(np.round(length / speed * 8.0), 1.0) / 8.0
    if edge_times is not None:
        t_edge = rng.choice(np.asarray(edge_times, dtype=np.float64), size=src.size)
    w = csr_matrix((t_edge, (src, dst)), shape=(n_nodes, n_nodes))
    if with_paths:
        tt64, pred0 = dijkstra(w, directed=True, return_predecessors=True)
        pred = np.where(pred0 < 0, 0, pred0 + 1).astype(np.int32)
        # distance along the chosen time-shortest path: accumulate edge lengths by
        # walking the predecessor tree in order of increasing travel time.
        elen = csr_matrix((np.round(length), (src, dst)), shape=(n_nodes, n_nodes)).toarray()
        dist = np.zeros((n_nodes, n_nodes), dtype=np.float64)
        order = np.argsort(tt64, axis=1, kind="stable")
        rows = np.arange(n_nodes)
        for r in range(1, n_nodes):
            d = order[:, r]
            p = pred0[rows, d]
            dist[rows, d] = dist[rows, p] + elen[p, d]
        dist = dist.astype(np.float32)
    else:
        tt64 = dijkstra(w, directed=True)
        pred, dist = None, None
    assert np.isfinite(tt64).all()
    tt = tt64.astype(np.float32)
    assert (tt.astype(np.float64) == tt64).all()
    lng = OLNG + (DLNG - OLNG) * gx / max(nx - 1, 1)
    lat = OLAT + (DLAT - OLAT) * gy / max(ny - 1, 1)
    stations = (np.linspace(0, n_nodes - 1, n_stations).astype(np.int64) + 1).astype(np.int32)
    return RoadGraph(tt, dist, pred, lng, lat, stations)


def make_request_stream(graph: RoadGraph, n_requests: int, t0_sec: int, t1_sec: int, seed: int = 1,
                        mean_trip_sec: float = 600.0, core_frac: float = 0.6):
    """Requests (onid, dnid, t) sorted by time, uniform rate on [t0, t1).

    Origins are concentrated in a central band; each destination sits at an
    exponentially distributed travel time from its origin (uniform OD
    is far too sparse to be Manhattan-like)."""
    rng = np.random.default_rng(seed)
    n = graph.n_nodes
    t = np.sort(rng.integers(t0_sec, t1_sec, size=n_requests)).astype(np.int64)
    lat = graph.lat
    mid = 0.5 * (lat.min() + lat.max())
    span = lat.max() - lat.min() + 1e-9
    wgt = np.exp(-0.5 * ((lat - mid) / (0.22 * span)) ** 2)
    wgt = core_frac * wgt / wgt.sum() + (1.0 - core_frac) / n
    onid = rng.choice(n, size=n_requests, p=wgt / wgt.sum())
    target = np.maximum(rng.exponential(mean_trip_sec, size=n_requests), 90.0)
    dnid = np.empty(n_requests, dtype=np.int64)
    # pick, among 64 random candidates, the node whose travel time is closest to target
    cand = rng.integers(0, n, size=(n_requests, 64))
    tt_c = graph.tt[onid[:, None], cand]
    tt_c = np.where(cand == onid[:, None], np.inf, tt_c)
    best = np.argmin(np.abs(tt_c - target[:, None]), axis=1)
    dnid[:] = cand[np.arange(n_requests), best]
    return (onid + 1).astype(np.int32), (dnid + 1).astype(np.int32), t


def diurnal_times(n_requests: int, seed: int = 1) -> np.ndarray:
    """Request times over a whole day (seconds from 0:00) with an evening peak."""
    rng = np.random.default_rng(seed)
    h = np.arange(24)
    rate = 0.35 + 0.65 * np.exp(-0.5 * ((h - 19) / 3.5) ** 2) + 0.45 * np.exp(-0.5 * ((h - 9) / 2.5) ** 2)
    rate[:5] *= 0.35
    p = rate / rate.sum()
    hours = rng.choice(24, size=n_requests, p=p)
    return np.sort(hours * 3600 + rng.integers(0, 3600, size=n_requests)).astype(np.int64)


def random_table(n_nodes: int, seed: int, kind: str = "ties", dtype=np.float64) -> np.ndarray:
    """Small stress tables (NOT metric): 'ties' = multiples of 30 s with many equal entries and
    zeros, 'real' = arbitrary float64 values so that summation order shows in the last bit."""
    rng = np.random.default_rng(seed)
    if kind == "ties":
        tt = rng.integers(0, 9, size=(n_nodes, n_nodes)).astype(np.float64) * 30.0
    else:
        tt = rng.uniform(1.0, 240.0, size=(n_nodes, n_nodes))
        if dtype == np.float32:
            tt = tt.astype(np.float32).astype(np.float64)
    np.fill_diagonal(tt, 0.0)
    return np.ascontiguousarray(tt.astype(dtype))


def random_epoch(tt: np.ndarray, n_veh: int, n_pending: int, cap: int, mode: int, T: int = 3600, seed: int = 0,
                 p_idle: float = 0.3, p_reb: float = 0.1, max_picking: int = 2, wait_sec: int = 420,
                 near: bool = False, frac_new: float = 0.5):
    """A self-consistent random epoch state (vehicle.py:41-70 / request.py:25-41 invariants):
    working vehicles carry `load` onboard requests (drop-offs only in sche) and some picking
    requests (pick-up before drop-off); picking requests belong to exactly one vehicle."""
    from .marshal import Epoch, MODE_OSP, ST_IDLE, ST_REBALANCING, ST_WORKING
    rng = np.random.default_rng(seed)
    n = tt.shape[0]
    tt64 = tt.astype(np.float64)
    status = rng.choice([ST_IDLE, ST_REBALANCING, ST_WORKING], size=n_veh,
                        p=[p_idle, p_reb, 1.0 - p_idle - p_reb])
    veh_nid = rng.integers(1, n + 1, size=n_veh)
    veh_t = np.where(rng.random(n_veh) < 0.5, 0.0, rng.uniform(0.0, 40.0, size=n_veh))
    on, dn, tr, onboard_flag, picking_flag = [], [], [], [], []
    veh_sche, veh_onb, veh_load = [], [], []

    def new_req(onboard, picking, origin=None):
        o = int(rng.integers(1, n + 1)) if origin is None else origin
        d = int(rng.integers(1, n + 1))
        if rng.random() < 0.15 and len(on):     # shared nodes -> exact cost ties
            k = int(rng.integers(0, len(on)))
            o, d = (on[k], d) if rng.random() < 0.5 else (o, dn[k])
        on.append(o); dn.append(d)
        tr.append(int(T - rng.integers(0, 600 if onboard else 140)))
        onboard_flag.append(onboard); picking_flag.append(picking)
        return len(on) - 1

    for v in range(n_veh):
        sche, onb = [], []
        if status[v] == ST_WORKING:
            load = int(rng.integers(0, cap + 1))
            npick = int(rng.integers(0 if load else 1, max_picking + 1))
            items = []
            for _ in range(load):
                items.append((new_req(True, False), 1))
            for _ in range(npick):
                r = new_req(False, True)
                items.append((r, 0)); items.append((r, 1))
            # random order with pick-up before drop-off and capacity respected when possible
            order = list(rng.permutation(len(items)))
            seq = [items[i] for i in order]
            for r in {r for r, _ in seq}:
                pos = [i for i, it in enumerate(seq) if it[0] == r]
                if len(pos) == 2 and seq[pos[0]][1] == 1:
                    seq[pos[0]], seq[pos[1]] = seq[pos[1]], seq[pos[0]]
            sche = [2 * r + d for r, d in seq]
            onb = [1 if onboard_flag[r] else 0 for r, _ in seq]
        elif status[v] == ST_REBALANCING:
            load = 0
            sche, onb = [-1], [0]
        else:
            load = 0
        veh_sche.append(sche); veh_onb.append(onb); veh_load.append(load)
    first_pending = len(on)
    for _ in range(n_pending):
        origin = None
        if near:
            v = int(rng.integers(0, n_veh))
            origin = int(veh_nid[v])
        new_req(False, False, origin)
    on = np.asarray(on, dtype=np.int32); dn = np.asarray(dn, dtype=np.int32)
    tr = np.asarray(tr, dtype=np.float64)
    ts = tt64[on - 1, dn - 1]
    clp = tr + wait_sec
    cld = tr + ts + 2 * wait_sec
    onboard_flag = np.asarray(onboard_flag, dtype=bool)
    cld = np.where(onboard_flag, cld + rng.uniform(0, 900, size=cld.size), cld)
    # shuffle request ids so that index order is not creation order
    nreq = on.size
    perm = rng.permutation(nreq)           # new index of old request i is inv[i]
    inv = np.empty(nreq, dtype=np.int64); inv[perm] = np.arange(nreq)
    remap = lambda c: -1 if c < 0 else 2 * int(inv[c >> 1]) + (c & 1)
    off = np.zeros(n_veh + 1, dtype=np.int32)
    np.cumsum([len(s) for s in veh_sche], out=off[1:])
    code = np.asarray([remap(c) for s in veh_sche for c in s], dtype=np.int32)
    onb = np.asarray([b for s in veh_onb for b in s], dtype=np.uint8)
    reb_node = np.where(status == ST_REBALANCING, rng.integers(1, n + 1, size=n_veh), 0)
    reb_ddl = np.where(status == ST_REBALANCING, T + rng.uniform(100, 700, size=n_veh), 0.0)
    pending = np.zeros(nreq, dtype=bool); pending[first_pending:] = True
    picking = np.asarray(picking_flag, dtype=bool)
    if mode == MODE_OSP:
        considered_old = np.flatnonzero(pending | picking)
    else:
        considered_old = np.flatnonzero(pending)
        considered_old = considered_old[rng.random(considered_old.size) < frac_new]
    search = np.sort(inv[considered_old]).astype(np.int32)
    gid = np.cumsum(rng.integers(1, 4, size=nreq)).astype(np.int32)   # ascending, sparse global ids
    return Epoch(T=T, cap=cap, mode=mode, veh_nid=veh_nid, veh_t_to_nid=veh_t, veh_load=np.asarray(veh_load),
                 veh_status=status, veh_sche_off=off, sche_code=code, sche_onboard=onb, veh_reb_node=reb_node,
                 veh_reb_ddl=reb_ddl, req_id=gid, req_onid=on[perm], req_dnid=dn[perm], req_clp=clp[perm],
                 req_cld=cld[perm], req_tr=tr[perm], req_ts=ts[perm], search=search)


def tile_fleet(ep, n: int):
    """Fleet of n x V vehicles on the same requests: the snapshot's fleet n times, every further copy in a
    different (seeded) vehicle order, so that an interleaved rank slice is a sample of the whole fleet and not
    the same V/n vehicles n times over.  (bench.py's weak-scaling and size-sweep workloads.)"""
    if n == 1:
        return ep
    from .marshal import Epoch
    parts = [ep] + [ep.take(np.random.default_rng(1000 + c).permutation(ep.n_veh)) for c in range(1, n)]
    kw = {f: getattr(ep, f) for f in Epoch.FIELDS}
    for f in ("veh_nid", "veh_t_to_nid", "veh_load", "veh_status", "veh_reb_node", "veh_reb_ddl", "sche_code", "sche_onboard"):
        kw[f] = np.concatenate([getattr(p, f) for p in parts])
    lens = np.concatenate([np.diff(p.veh_sche_off) for p in parts])
    off = np.zeros(lens.size + 1, dtype=np.int32)
    np.cumsum(lens, out=off[1:])
    kw["veh_sche_off"] = off
    return Epoch(T=ep.T, cap=ep.cap, mode=ep.mode, **kw)
