"""Multi-GPU: vehicles shard across ranks (each vehicle's trip table depends only on that
vehicle, the shared request list and the read-only table: dispatch_osp.py:107-111), every rank
builds its slice of the pool, and one all-gather assembles the vehicle-major pool for the
unchanged assigner.  One process per GPU, torch.distributed for the plumbing."""
from __future__ import annotations

import numpy as np

from .marshal import Pool, concat_pools


def merge_vehicle_major(parts: list[tuple[Pool, np.ndarray]], n_veh: int) -> Pool:
    """parts[r] = (pool of rank r, global vehicle index of each of its local vehicles).
    Rank pools are vehicle-major over their own vehicles; the merged pool is vehicle-major over
    the whole fleet (the reference's order, dispatch_osp.py:107)."""
    pools = []
    for pool, sel in parts:
        p = Pool(stats=pool.stats, **{f: getattr(pool, f) for f in Pool.FIELDS})
        p.ent_veh = np.asarray(sel, dtype=np.int32)[pool.ent_veh] if pool.n_entries else pool.ent_veh
        pools.append(p)
    cat = concat_pools(pools)
    order = np.argsort(cat.ent_veh, kind="stable")
    return cat.take(order)


def merge_sba(parts: list[tuple[Pool, np.ndarray]], n_veh: int, search: np.ndarray | None = None) -> Pool:
    """SBA order is request-major then one empty pair per vehicle (dispatch_sba.py:56-69): pairs sort by
    (position of the request in new_received_rids, vehicle), the empty pairs by vehicle.  `search` is the epoch's
    search array (request indices in the reference's loop order, dispatch_sba.py:56); without it the request
    index itself is taken as the position (true when the simulator hands ascending rids)."""
    pools = []
    for pool, sel in parts:
        p = Pool(stats=pool.stats, **{f: getattr(pool, f) for f in Pool.FIELDS})
        p.ent_veh = np.asarray(sel, dtype=np.int32)[pool.ent_veh] if pool.n_entries else pool.ent_veh
        pools.append(p)
    cat = concat_pools(pools)
    is_empty = (cat.ent_trip_off[1:] - cat.ent_trip_off[:-1]) == 0
    req = cat.trip_req[np.minimum(cat.ent_trip_off[:-1], max(cat.trip_req.size - 1, 0))] if cat.trip_req.size else np.zeros(cat.n_entries, dtype=np.int64)
    if search is not None and req.size:
        search = np.asarray(search, dtype=np.int64)
        pos_of = np.full(int(max(search.max(initial=-1), req.max(initial=-1))) + 1, np.iinfo(np.int32).max, dtype=np.int64)
        pos_of[search[::-1]] = np.arange(search.size - 1, -1, -1)      # first position of a request in the loop
        req = pos_of[req]
    req = np.where(is_empty, np.iinfo(np.int64).max, req)
    order = np.lexsort((cat.ent_veh, req))
    return cat.take(order)


def balanced_placement(cost: np.ndarray, world: int, root: int | None = None, root_share: float = 1.0):
    """Placement of every vehicle on a rank by its expected work (e.g. the size of its share of the previous epochs'
    pools): longest-processing-time-first bin packing, ties towards the emptier rank.  A rank's vehicles keep their
    global order (the assembled pool is vehicle-major, dispatch_osp.py:107).  `root_share` < 1 gives the gathering rank
    that fraction of an even share of the work: it also merges the fleet-wide pool after its own build, so a lighter
    slice shortens the epoch.  Returns (veh_rank, veh_local, sels) where sels[r] are the global indices of rank r's
    vehicles in local order."""
    cost = np.asarray(cost, dtype=np.float64)
    n = cost.size
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world)
    if root is not None and world > 1 and root_share < 1.0:
        load[root] = (1.0 - root_share) * float(cost.sum()) / world        # a head start the packing fills last
    count = np.zeros(world, dtype=np.int64)
    veh_rank = np.zeros(n, dtype=np.int32)
    cap = (n + world - 1) // world + 1               # a staging slot holds at most this many vehicles
    for v in order.tolist():
        open_ranks = np.flatnonzero(count < cap)
        r = int(open_ranks[np.argmin(load[open_ranks])])
        veh_rank[v] = r
        load[r] += cost[v]
        count[r] += 1
    veh_local = np.zeros(n, dtype=np.int32)
    sels = []
    for r in range(world):
        sel = np.flatnonzero(veh_rank == r)
        veh_local[sel] = np.arange(sel.size, dtype=np.int32)
        sels.append(sel.astype(np.int64))
    return veh_rank, veh_local, sels


def allgather_pool(local: Pool, sel: np.ndarray, n_veh: int, mode_sba: bool = False, device=None, search=None) -> Pool:
    """All-gather the rank-local pools (torch.distributed; NCCL over NVLink when `device` is a
    CUDA device, gloo on CPU) and return the full vehicle-major pool on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    sizes = torch.tensor([local.n_entries, local.trip_req.size, local.sche_code.size, len(sel)], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    mx = all_sizes.max(axis=0)

    def gather(arr: np.ndarray, n_max: int, col: int, extra: int = 0):
        n_max = int(n_max) + extra
        buf = torch.zeros(max(n_max, 1), dtype=torch.from_numpy(np.zeros(1, arr.dtype)).dtype, device=dev)
        if arr.size:
            buf[:arr.size] = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
        outs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf)
        return [o.cpu().numpy()[:int(all_sizes[r, col]) + extra] for r, o in enumerate(outs)]

    fields = {}
    for f, col, extra in (("ent_veh", 0, 0), ("ent_kind", 0, 0), ("ent_cost", 0, 0), ("ent_delay", 0, 0), ("ent_nfeas", 0, 0),
                          ("ent_trip_off", 0, 1), ("ent_sche_off", 0, 1), ("trip_req", 1, 0), ("sche_code", 2, 0)):
        fields[f] = gather(getattr(local, f), mx[col], col, extra)
    sels = gather(np.asarray(sel, dtype=np.int64), mx[3], 3)
    parts = [(Pool(**{f: fields[f][r] for f in Pool.FIELDS}), sels[r]) for r in range(world)]
    if mode_sba:
        return merge_sba(parts, n_veh, search)
    return merge_vehicle_major(parts, n_veh)


def merge_vehicle_major_torch(rank_pools: list[dict], world: int, totals: dict | None = None) -> dict:
    """Device-side version of merge_vehicle_major for interleaved sharding (vehicle v lives on rank
    v % world as local vehicle v // world).  rank_pools[r] maps Pool field name -> torch tensor
    (already trimmed to the rank's sizes) on one device; returns the merged pool tensors in the
    reference's vehicle-major order (dispatch_osp.py:107).  Pure torch indexing: plumbing."""
    import torch

    veh = torch.cat([p["ent_veh"].to(torch.int64) * world + r for r, p in enumerate(rank_pools)])
    order = torch.argsort(veh, stable=True)
    out = {"ent_veh": veh[order].to(torch.int32)}
    for f in ("ent_kind", "ent_cost", "ent_delay", "ent_nfeas"):
        out[f] = torch.cat([p[f] for p in rank_pools])[order]
    for off_name, data_name in (("ent_trip_off", "trip_req"), ("ent_sche_off", "sche_code")):
        lens, starts, base = [], [], 0
        for p in rank_pools:
            o = p[off_name]
            lens.append(o[1:] - o[:-1])
            starts.append(o[:-1] + base)
            base += int(p[data_name].numel())
        lens = torch.cat(lens)[order]
        starts = torch.cat(starts)[order]
        new_off = torch.zeros(lens.numel() + 1, dtype=torch.int64, device=lens.device)
        torch.cumsum(lens, 0, out=new_off[1:])
        data = torch.cat([p[data_name] for p in rank_pools])
        total = int(totals[data_name]) if totals is not None else int(data.numel())   # known on the host: no device sync
        idx = torch.repeat_interleave(starts - new_off[:-1], lens, output_size=total) + torch.arange(total, device=lens.device)
        out[off_name] = new_off
        out[data_name] = data[idx]
    return out


class PeerPool:
    """Vehicle-major pool of the whole fleet assembled on every rank by peer stores over NVLink
    (include/amod_b200.h, amod_gpool_*): the only collectives are the all-gather of per-vehicle
    counts and a barrier.  OSP modes; vehicles interleaved v % world."""

    def __init__(self, builder, world: int, rank: int, n_veh_total: int, cap_entries: int, cap_trip_reqs: int,
                 cap_stops: int, device):
        import ctypes as C

        import torch
        import torch.distributed as dist
        self.b, self.world, self.rank, self.n_veh_total, self.device = builder, world, rank, n_veh_total, device
        self.vmax = (n_veh_total + world - 1) // world
        # every rank must use the SAME capacities: the buffer layout is computed from them on each side
        caps = [None] * world
        dist.all_gather_object(caps, (int(cap_entries), int(cap_trip_reqs), int(cap_stops)))
        cap_entries, cap_trip_reqs, cap_stops = (max(c[i] for c in caps) for i in range(3))
        handle = C.create_string_buffer(64)
        builder._check(builder.lib.amod_gpool_create(builder.h, cap_entries, cap_trip_reqs, cap_stops, n_veh_total, world, handle))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw))
        blob = C.create_string_buffer(b"".join(handles), 64 * world)
        builder._check(builder.lib.amod_gpool_open_peers(builder.h, world, rank, blob))
        self.counts = torch.zeros(self.vmax * 3, dtype=torch.int32, device=device)
        self.counts_all = torch.zeros(world * self.vmax * 3, dtype=torch.int32, device=device)
        self.nccl = dist.get_backend() == "nccl"
        # one GPU per rank -> counts, barriers and payload all go through peer memory (no host collective)
        self.fused = self.nccl
        dist.barrier()

    def attach(self, root: int = 0, veh_rank=None, veh_local=None):
        """From now on every builder.build(...) of this rank's slice also assembles: the slice is pushed (to `root` only,
        or to every rank with root < 0) and the merging ranks build the fleet-wide pool inside the build's own captured
        graph and single host synchronisation (amod_gpool_attach).  veh_rank / veh_local: placement of every global
        vehicle (see balanced_placement); None = interleaved."""
        b = self.b
        if veh_rank is None:
            rc = b.lib.amod_gpool_attach(b.h, self.n_veh_total, root, None, None)
        else:
            self._vr = np.ascontiguousarray(veh_rank, dtype=np.int32)
            self._vl = np.ascontiguousarray(veh_local, dtype=np.int32)
            assert self._vr.size == self.n_veh_total == self._vl.size
            rc = b.lib.amod_gpool_attach(b.h, self.n_veh_total, root, self._vr.ctypes.data, self._vl.ctypes.data)
        b._check(rc)
        self.attached, self.root = True, root

    def detach(self):
        self.b._check(self.b.lib.amod_gpool_detach(self.b.h))
        self.attached = False

    def assemble(self):
        """Call after builder.build(...) of this rank's slice.  Returns nothing; the pool is in this
        rank's global pool (view()/fetch()).  A no-op while attached (the build already assembled)."""
        import torch
        import torch.distributed as dist
        b = self.b
        if getattr(self, "attached", False):
            return
        if self.fused:
            b._check(b.lib.amod_gpool_assemble(b.h, self.n_veh_total))
            return
        self.counts.zero_()
        b._check(b.lib.amod_pool_veh_counts(b.h, self.counts.data_ptr()))
        if self.nccl:
            dist.all_gather_into_tensor(self.counts_all, self.counts)
        else:   # CPU collectives (tests on one GPU): stage through the host
            torch.cuda.current_stream().synchronize()
            parts = [torch.zeros(self.vmax * 3, dtype=torch.int32) for _ in range(self.world)]
            dist.all_gather(parts, self.counts.cpu())
            self.counts_all.copy_(torch.cat(parts))
            torch.cuda.current_stream().synchronize()
        b._check(b.lib.amod_gpool_scatter(b.h, self.counts_all.data_ptr(), self.vmax, self.n_veh_total))
        dist.barrier()        # every rank's peer stores have landed before anyone reads

    def view(self):
        from . import capi
        import ctypes as C
        v = capi.AmodPool()
        self.b._check(self.b.lib.amod_gpool_view(self.b.h, C.byref(v)))
        return v

    def fetch(self, reuse: bool = False) -> Pool:
        """This rank's assembled pool as numpy arrays.  reuse=True: views of grow-only page-locked buffers owned by
        this object (straight DMA, overwritten by the next fetch), like PoolBuilder.fetch(reuse=True)."""
        import ctypes as C

        from . import capi
        v = self.view()
        n = {"e": int(v.n_entries), "e1": int(v.n_entries) + 1, "t": int(v.n_trip_reqs), "s": int(v.n_stops)}
        if reuse:
            from .pool import packed_pool_views
            arrs = packed_pool_views(self, n)
        else:
            arrs = {name: np.empty(max(n[key], 1), dtype=dt) for name, dt, key in capi.POOL_FIELDS}
        o = capi.AmodPool()
        o.cap_entries, o.cap_trip_reqs, o.cap_stops = v.n_entries, v.n_trip_reqs, v.n_stops
        for name, _dt, _key in capi.POOL_FIELDS:
            setattr(o, name, arrs[name].ctypes.data)
        self.b._check(self.b.lib.amod_gpool_fetch(self.b.h, C.byref(o)))
        return Pool(**{name: arrs[name][:n[key]] for name, _dt, key in capi.POOL_FIELDS})

    def _release_host_bufs(self):
        buf = getattr(self, "_host_buf", None)
        if buf is not None and getattr(self.b, "h", None):
            self.b.lib.amod_host_unregister(self.b.h, buf.ctypes.data)
        self._host_buf = None
