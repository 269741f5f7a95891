"""Host logic: marshalling Veh/Req objects <-> SoA arrays <-> reference-style records."""
import collections
import types

import numpy as np

from conftest import load_golden
from amod_b200 import marshal
import oracle


class _Status:
    def __init__(self, v):
        self.value = v


def _objects(ep):
    """Minimal duck-typed Veh/Req objects from an Epoch (what the seams read)."""
    reqs = {}
    objs = []
    for i in range(ep.n_req):
        r = types.SimpleNamespace(id=int(ep.req_id[i]), onid=int(ep.req_onid[i]), dnid=int(ep.req_dnid[i]),
                                  Clp=float(ep.req_clp[i]), Cld=float(ep.req_cld[i]), Tr=float(ep.req_tr[i]),
                                  Ts=float(ep.req_ts[i]), status=_Status(1))
        reqs[r.id] = r
        objs.append(r)
    vehs = []
    for v in range(ep.n_veh):
        sche, onboard = [], []
        for k in range(ep.veh_sche_off[v], ep.veh_sche_off[v + 1]):
            c = int(ep.sche_code[k])
            if c < 0:
                sche.append((-1, 0, int(ep.veh_reb_node[v]), float(ep.veh_reb_ddl[v])))
            else:
                r = objs[c >> 1]
                sche.append((r.id, -1, r.dnid, r.Cld) if c & 1 else (r.id, 1, r.onid, r.Clp))
                if ep.sche_onboard[k]:
                    onboard.append(r.id)
        vehs.append(types.SimpleNamespace(id=v, nid=int(ep.veh_nid[v]), t_to_nid=float(ep.veh_t_to_nid[v]),
                                          load=int(ep.veh_load[v]), status=_Status(int(ep.veh_status[v])), sche=sche,
                                          onboard_rids=onboard, K=ep.cap))
    return reqs, objs, vehs


def test_marshal_roundtrip_reproduces_epoch():
    tt, cases = load_golden("rand_real_osp_cap4")
    for ep, _gold in cases[:3]:
        reqs, objs, vehs = _objects(ep)
        search_rids = [int(ep.req_id[i]) for i in ep.search]
        ep2, req_objs = marshal.marshal_epoch(reqs, vehs, search_rids, ep.T, ep.cap, ep.mode)
        # request set = searched + referenced by a schedule; indices may be a subset of the original
        remap = {int(ep.req_id[i]): i for i in range(ep.n_req)}
        assert [remap[int(r)] for r in ep2.req_id[ep2.search]] == ep.search.tolist()
        a = oracle.osp_build(ep, tt)
        b = oracle.osp_build(ep2, tt)
        back = np.array([remap[int(r)] for r in ep2.req_id])
        assert np.array_equal(a.ent_veh, b.ent_veh) and np.array_equal(a.ent_cost, b.ent_cost)
        assert np.array_equal(a.trip_req, back[b.trip_req])
        bc = np.where(b.sche_code < 0, -1, 2 * back[np.maximum(b.sche_code, 0) >> 1] + (b.sche_code & 1))
        assert np.array_equal(a.sche_code, bc)


def test_pool_to_records_layout():
    tt, cases = load_golden("rand_ties_osp_cap4")
    ep, gold = cases[0]
    reqs, objs, vehs = _objects(ep)
    recs = marshal.pool_to_records(gold, ep, vehs, objs)
    assert len(recs) == gold.n_entries
    for e, rec in enumerate(recs):
        veh, trip, sche, cost, score = rec
        assert isinstance(rec, list) and isinstance(sche, list) and all(isinstance(s, tuple) and len(s) == 4 for s in sche)
        assert veh is vehs[gold.ent_veh[e]]
        assert [r.id for r in trip] == sorted(r.id for r in trip)
        assert cost == gold.ent_cost[e] and score == 0.0
    # records_to_pool inverts it
    index_of = {int(r): i for i, r in enumerate(ep.req_id)}
    back = marshal.records_to_pool(recs, vehs, index_of, ep.mode)
    ok, why = marshal.pools_equal(back, gold, check_delay=False, check_nfeas=False)
    assert ok, why
    # a consumer may pop the rebalancing stop in place (vehicle.py:234-241): lists must be fresh
    recs2 = marshal.pool_to_records(gold, ep, vehs, objs)
    assert all(a[2] is not b[2] for a, b in zip(recs, recs2))


def test_shard_and_merge_is_vehicle_major():
    from amod_b200 import dist
    tt, cases = load_golden("rand_real_osp_cap4")
    ep, gold = cases[1]
    world = 3
    parts = []
    for rank in range(world):
        sub, sel = ep.shard(rank, world)
        p = oracle.osp_build(sub, tt)
        parts.append((p, sel))
    merged = dist.merge_vehicle_major(parts, ep.n_veh)
    ok, why = marshal.pools_equal(merged, gold, check_kind=True)
    assert ok, why


def test_torch_merge_matches_numpy_merge():
    import torch
    from amod_b200 import dist
    tt, cases = load_golden("rand_real_osp_cap4")
    ep, gold = cases[2]
    world = 3
    rank_pools = []
    for rank in range(world):
        sub, sel = ep.shard(rank, world)
        p = oracle.osp_build(sub, tt)
        rank_pools.append({f: torch.from_numpy(np.ascontiguousarray(getattr(p, f))) for f in marshal.Pool.FIELDS})
    m = dist.merge_vehicle_major_torch(rank_pools, world)
    got = marshal.Pool(**{f: m[f].numpy() for f in marshal.Pool.FIELDS})
    ok, why = marshal.pools_equal(got, gold, check_kind=True)
    assert ok, why


def test_take_reorders_vehicles_with_their_schedules():
    """Epoch.take (used by shard and by bench.py's weak-scaling fleet): any vehicle order, schedules follow, and
    building the permuted epoch gives the original pool with the vehicles renamed."""
    tt, cases = load_golden("rand_real_osp_cap4")
    ep = cases[0][0]
    perm = np.random.default_rng(3).permutation(ep.n_veh)
    pe = ep.take(perm)
    assert pe.n_veh == ep.n_veh and pe.sche_code.size == ep.sche_code.size
    for k in (0, ep.n_veh // 2, ep.n_veh - 1):
        v = perm[k]
        assert np.array_equal(pe.sche_code[pe.veh_sche_off[k]:pe.veh_sche_off[k + 1]],
                              ep.sche_code[ep.veh_sche_off[v]:ep.veh_sche_off[v + 1]])
        assert pe.veh_nid[k] == ep.veh_nid[v] and pe.veh_load[k] == ep.veh_load[v]
    a, b = oracle.osp_build(ep, tt), oracle.osp_build(pe, tt)
    assert a.n_entries == b.n_entries
    # per-vehicle entry counts travel with the vehicles
    ca = np.bincount(a.ent_veh, minlength=ep.n_veh)
    cb = np.bincount(b.ent_veh, minlength=ep.n_veh)
    assert np.array_equal(cb, ca[perm])


def test_request_table_and_index_paths_agree():
    """marshal_epoch with the per-simulation request table, without it, and with request ids too sparse for the
    direct index tables must produce the same epoch; a different `reqs` container starts a fresh table."""
    tt, cases = load_golden("rand_real_osp_cap4")
    ep = cases[0][0]
    reqs, objs, vehs = _objects(ep)
    search = [int(ep.req_id[i]) for i in ep.search]
    plain, _ = marshal.marshal_epoch(reqs, vehs, search, ep.T, ep.cap, ep.mode)
    table = marshal.ReqTable()
    for _ in range(2):                                   # second call: every request already known
        cached, _ = marshal.marshal_epoch(reqs, vehs, search, ep.T, ep.cap, ep.mode, req_table=table)
        for f in marshal.Epoch.FIELDS:
            assert np.array_equal(getattr(plain, f), getattr(cached, f)) and getattr(plain, f).dtype == getattr(cached, f).dtype, f
    # same ids, different request data, different container: the table must not serve the old values
    reqs2 = {rid: types.SimpleNamespace(**{**r.__dict__, "Clp": r.Clp + 7.0, "Cld": r.Cld + 9.0}) for rid, r in reqs.items()}
    vehs2 = [types.SimpleNamespace(**{**v.__dict__, "sche": [(s[0], s[1], s[2], s[3] + (0.0 if s[0] < 0 else (7.0 if s[1] == 1 else 9.0)))
                                                           for s in v.sche]}) for v in vehs]
    shifted, _ = marshal.marshal_epoch(reqs2, vehs2, search, ep.T, ep.cap, ep.mode, req_table=table)
    assert np.array_equal(shifted.req_clp, plain.req_clp + 7.0) and np.array_equal(shifted.req_cld, plain.req_cld + 9.0)
    # ids beyond the direct-table limit take the sort-based path
    big = 1 << 23
    reqs3 = {rid + big: types.SimpleNamespace(**{**r.__dict__, "id": rid + big}) for rid, r in reqs.items()}
    vehs3 = [types.SimpleNamespace(**{**v.__dict__, "sche": [(s[0] + big if s[0] >= 0 else s[0], s[1], s[2], s[3]) for s in v.sche],
                                      "onboard_rids": [x + big for x in v.onboard_rids]}) for v in vehs]
    sparse, _ = marshal.marshal_epoch(reqs3, vehs3, [r + big for r in search], ep.T, ep.cap, ep.mode)
    for f in marshal.Epoch.FIELDS:
        want = getattr(plain, f) + (big if f == "req_id" else 0)
        assert np.array_equal(want, getattr(sparse, f)), f


def test_lazy_records_keep_sorted_positions_after_iteration():
    """anything that iterates the lazy pair list between scoring and the GPU assigner materialises the
    records; apply_order must then permute the materialised list too, records built before scoring must carry
    delay / len(trip) afterwards, and a consumer's own sort() makes the assigner's permutation refuse to apply."""
    import pytest
    tt, cases = load_golden("sim_osp_cap4")
    ep, gold = cases[2]
    reqs, objs, vehs = _objects(ep)
    order, sel = oracle.greedy_assign(gold, ep.n_veh, ep.n_req)

    def picked(lazy):
        return [(lazy[i][0].id, [r.id for r in lazy[i][1]], lazy[i][2], float(lazy[i][3]), lazy[i][4]) for i in sel.tolist()]

    clean = marshal.LazyPoolRecords(gold, ep, vehs, objs)
    clean.scored = True
    clean.apply_order(order)
    want = picked(clean)

    touched = marshal.LazyPoolRecords(gold, ep, vehs, objs)
    first = touched[0]                      # built before scoring: cost / 0.0
    assert first[3] == gold.ent_cost[0] and first[4] == 0.0
    n_seen = sum(1 for _pair in touched)    # a debug hook iterating the candidate list
    assert n_seen == gold.n_entries
    touched.scored = True
    assert first[3] == gold.ent_delay[0] and first[4] == len(first[1])
    touched.apply_order(order)
    assert picked(touched) == want
    assert touched[int(sel[0])] is touched.materialise()[int(sel[0])]

    resorted = marshal.LazyPoolRecords(gold, ep, vehs, objs)
    resorted.scored = True
    resorted.sort(key=lambda p: p[0].id, reverse=True)
    with pytest.raises(RuntimeError):
        resorted.apply_order(order)


def test_state_rows_and_request_id_space():
    """amod_b200/state.py host logic: fixed-stride vehicle rows over request IDS, and a pool rewritten over ids."""
    from amod_b200 import state as astate
    tt, cases = load_golden("sim_osp_cap4")
    ep, gold = cases[2]
    rows = astate.DeviceFleet.rows_of(ep, 16)
    lens = np.diff(ep.veh_sche_off)
    assert np.array_equal(rows["len"], lens)
    for v in range(ep.n_veh):
        codes = ep.sche_code[ep.veh_sche_off[v]:ep.veh_sche_off[v + 1]]
        want = [(-1 if c < 0 else 2 * int(ep.req_id[c >> 1]) + (c & 1)) for c in codes.tolist()]
        assert rows["code"][v, :lens[v]].tolist() == want and not rows["code"][v, lens[v]:].any()
    p = astate.pool_in_request_ids(gold, ep)
    assert np.array_equal(p.trip_req, ep.req_id[gold.trip_req]) and np.array_equal(p.ent_cost, gold.ent_cost)
    back = np.where(p.sche_code < 0, -1, 2 * np.searchsorted(ep.req_id, np.maximum(p.sche_code, 0) >> 1) + (p.sche_code & 1))
    assert np.array_equal(back, gold.sche_code)
