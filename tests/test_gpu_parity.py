"""GPU parity (run on the B200 box): the CUDA pool, through the C-ABI, against
 (a) the committed golden vectors produced by the unmodified reference, and
 (b) the C oracle on larger seeded inputs.
Bar: bit-exact entry set, order, trips, chosen schedules AND float64 costs/delays (rtol 0):
the kernels accumulate in the reference's order, so no tolerance is needed (north_star allows 1e-6)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_families, load_golden, table_from_spec
from amod_b200 import marshal, synth
from amod_b200.pool import PoolBuilder

pytestmark = pytest.mark.gpu

_builders = {}


def builder_for(tt):
    key = (tt.ctypes.data, tt.shape, tt.dtype.str)
    if key not in _builders:
        _builders[key] = (PoolBuilder(tt), tt)
    return _builders[key][0]


@pytest.mark.parametrize("family", golden_families(("rand_", "sim_")))
def test_cuda_pool_matches_reference_golden(family):
    import oracle
    tt, cases = load_golden(family)
    b = builder_for(tt)
    for i, (ep, gold) in enumerate(cases):
        got = b.build_pool(ep)
        ok, why = marshal.pools_equal(got, gold, rtol=0.0, check_kind=True)
        assert ok, f"{family} case {i}: {why}"
        ref = (oracle.sba_build if ep.mode == marshal.MODE_SBA else oracle.osp_build)(ep, tt)
        for key in ("n_evals", "n_lookups", "n_tasks", "n_sches_stored", "n_screens"):
            assert got.stats[key] == ref.stats[key], (family, i, key, got.stats[key], ref.stats[key])


def test_cuda_level_overflow_raises_indexerror_like_reference():
    tt, cases = load_golden("levels")
    b = builder_for(tt)
    (ep_ok, gold_ok), (ep_bad, _none) = cases
    ok, why = marshal.pools_equal(b.build_pool(ep_ok), gold_ok, check_kind=True)
    assert ok, why
    with pytest.raises(IndexError):
        b.build_pool(ep_bad)


def test_cuda_f32_and_f64_tables_agree_on_fp32_representable_data():
    tt, cases = load_golden("rand_real32_osp_cap2")
    assert tt.dtype == np.float32
    b32, b64 = PoolBuilder(tt), PoolBuilder(tt.astype(np.float64))
    for ep, gold in cases[:2]:
        for b in (b32, b64):
            ok, why = marshal.pools_equal(b.build_pool(ep), gold, check_kind=True)
            assert ok, why


@pytest.mark.parametrize("mode,cap,n_veh,n_pending,wait", [
    (marshal.MODE_OSP, 4, 300, 120, 260), (marshal.MODE_OSP, 8, 120, 60, 200), (marshal.MODE_OSP_NR, 4, 300, 150, 300),
    (marshal.MODE_SBA, 4, 800, 500, 420), (marshal.MODE_OSP, 1, 200, 100, 300)])
def test_cuda_pool_matches_oracle_on_manhattan_grid(mode, cap, n_veh, n_pending, wait):
    """Larger seeded epochs on the 4,091-node synthetic Manhattan table (fp32)."""
    import oracle
    tt = _grid()
    b = builder_for(tt)
    ep = synth.random_epoch(tt, n_veh=n_veh, n_pending=n_pending, cap=cap, mode=mode, seed=77 + cap, wait_sec=wait,
                            near=True, T=5000)
    ref = (oracle.sba_build if mode == marshal.MODE_SBA else oracle.osp_build)(ep, tt, n_threads=oracle.max_threads())
    assert ref.error == 0
    got = b.build_pool(ep)
    ok, why = marshal.pools_equal(got, ref, rtol=0.0, check_kind=True)
    assert ok, why
    for key in ("n_evals", "n_lookups", "n_tasks", "n_sches_stored"):
        assert got.stats[key] == ref.stats[key], (key, got.stats[key], ref.stats[key])
    assert got.stats["n_evals"] > 1000


_grid_tt = None


def _grid():
    global _grid_tt
    if _grid_tt is None:
        _grid_tt = synth.make_road_graph(seed=0).tt
    return _grid_tt


def test_cuda_empty_and_degenerate_inputs():
    tt = synth.random_table(12, 3, "ties")
    b = builder_for(tt)
    import oracle
    for mode in (marshal.MODE_OSP, marshal.MODE_OSP_NR, marshal.MODE_SBA):
        # no requests at all
        ep = synth.random_epoch(tt, n_veh=9, n_pending=0, cap=3, mode=mode, seed=5, max_picking=0 if mode != marshal.MODE_OSP else 1)
        ref = (oracle.sba_build if mode == marshal.MODE_SBA else oracle.osp_build)(ep, tt)
        ok, why = marshal.pools_equal(b.build_pool(ep), ref, check_kind=True)
        assert ok, (mode, why)
    # no vehicles
    ep = synth.random_epoch(tt, n_veh=4, n_pending=5, cap=3, mode=marshal.MODE_OSP, seed=6)
    sub, _sel = ep.shard(7, 8)          # rank 7 of 8 gets no vehicle
    assert sub.n_veh == 0
    assert b.build_pool(sub).n_entries == 0


@pytest.mark.parametrize("family", golden_families(("npo_",)))
def test_cuda_npo_matches_reference(family):
    z = np.load(os.path.join(GOLDEN, family + ".npz"))
    tt = table_from_spec(json.loads(str(z["spec"])))
    b = builder_for(tt)
    for i in range(int(z["n_cases"])):
        idle, pend, T = z[f"c{i}_idle_nid"], z[f"c{i}_pend_onid"], int(z[f"c{i}_T"])
        oi, op, od = b.npo_match(idle, pend)
        order = np.argsort(oi)
        assert np.array_equal(oi[order], z[f"c{i}_m_idle"])
        assert np.array_equal(pend[op[order]], z[f"c{i}_m_node"])
        assert np.array_equal(T + od[order] + 240, z[f"c{i}_m_ddl"])


@pytest.mark.parametrize("persistent", [True, False], ids=["one-launch", "launch-per-round"])
def test_cuda_npo_matches_oracle_with_ties(persistent, monkeypatch):
    """Both forms of the prefix rounds (all rounds in one cooperative launch / two launches per round), on tables full of exact
    ties, with stations (many idle vehicles on one node), more vehicles than requests and the reverse."""
    import oracle
    if not persistent:
        monkeypatch.setenv("AMOD_NPO_NO_PERSIST", "1")
    rng = np.random.default_rng(4)
    for kind, n_nodes in (("ties", 30), ("ties", 200), ("real", 200)):
        tt = synth.random_table(n_nodes, 9, kind)
        b = PoolBuilder(tt)
        sizes = ((40, 25), (25, 40), (1, 7), (300, 200), (900, 1100), (2500, 600)) + (((60, 17000),) if kind == "real" else ())   # (> 16,384 requests: ordered on the host)
        for n_idle, n_pend in sizes:
            stations = rng.integers(1, n_nodes + 1, size=5)
            idle = np.where(rng.random(n_idle) < 0.6, rng.choice(stations, n_idle), rng.integers(1, n_nodes + 1, size=n_idle)).astype(np.int32)
            pend = rng.integers(1, n_nodes + 1, size=n_pend).astype(np.int32)
            a = oracle.npo_match(tt, idle, pend)
            g = b.npo_match(idle, pend)
            for x, y in zip(a, g):
                assert np.array_equal(x, y), (kind, n_nodes, n_idle, n_pend)
        b.close()


def test_seam_functions_return_reference_style_records():
    """B1/B2 through duck-typed Veh/Req objects: live references, fresh mutable schedules."""
    from test_marshal_cpu import _objects
    import oracle
    tt, cases = load_golden("sim_osp_cap4")
    b = builder_for(tt)
    ep, gold = cases[1]
    reqs, objs, vehs = _objects(ep)
    search_rids = [int(ep.req_id[i]) for i in ep.search]
    recs = b.compute_candidate_veh_trip_pairs(search_rids, search_rids, reqs, vehs, ep.T, 0.1, True, False)
    assert len(recs) == gold.n_entries
    for e, (veh, trip, sche, cost, score) in enumerate(recs):
        assert veh is vehs[gold.ent_veh[e]] and cost == gold.ent_cost[e]
    assert np.array_equal(recs.delays, gold.ent_delay)


def test_cuda_sba_and_npo_at_peak_epoch_scale():
    """BASELINE.json configs[4]: SBA matrix + NPO match at 20k vehicles x 5k pending requests."""
    import oracle
    tt = _grid()
    b = builder_for(tt)
    ep = synth.random_epoch(tt, n_veh=20000, n_pending=5000, cap=4, mode=marshal.MODE_SBA, seed=5, wait_sec=420, near=True,
                            frac_new=1.0, T=5000)
    assert ep.n_veh * ep.n_search == 100_000_000
    ref = oracle.sba_build(ep, tt, n_threads=oracle.max_threads())
    got = b.build_pool(ep)
    ok, why = marshal.pools_equal(got, ref, rtol=0.0, check_kind=True)
    assert ok, why
    assert got.stats["n_evals"] == ref.stats["n_evals"] and got.stats["n_screens"] == 100_000_000
    # NPO: idle vehicles x pending requests of the same epoch (rebalancing_npo.py:26-59)
    idle = ep.veh_nid[ep.veh_status == marshal.ST_IDLE][:6000]
    pend = ep.req_onid[ep.search][:1500]
    a = oracle.npo_match(tt, idle, pend)
    g = b.npo_match(idle, pend)
    for x, y in zip(a, g):
        assert np.array_equal(x, y)


def test_cuda_long_schedules_take_two_warp_steps(mode=marshal.MODE_SBA):
    """Base schedules with more than 31 stops (a row wider than a warp): SBA inserts into the vehicle's
    full current schedule (dispatch_sba.py:62)."""
    import oracle
    tt = synth.random_table(40, 21, "real", np.float32)
    b = builder_for(tt)
    ep = synth.random_epoch(tt, n_veh=60, n_pending=40, cap=10, mode=mode, seed=9, wait_sec=3000, near=True, max_picking=13,
                            frac_new=0.9, T=9000)
    assert np.diff(ep.veh_sche_off).max() >= 32
    ref = (oracle.sba_build if mode == marshal.MODE_SBA else oracle.osp_build)(ep, tt)
    assert ref.error == 0
    got = b.build_pool(ep)
    ok, why = marshal.pools_equal(got, ref, rtol=0.0, check_kind=True)
    assert ok, why
    assert got.stats["n_evals"] == ref.stats["n_evals"] and got.stats["n_lookups"] == ref.stats["n_lookups"]


@pytest.mark.parametrize("persistent", [True, False], ids=["one-launch", "launch-per-round"])
def test_cuda_greedy_assignment_matches_reference_and_oracle(persistent, monkeypatch):
    """Next row: greedy_assignment over the device-resident pool; both forms of the rounds (all rounds in one
    cooperative launch / two launches per round with a host check every four rounds)."""
    if not persistent:
        monkeypatch.setenv("AMOD_GA_NO_PERSIST", "1")
    import oracle
    z = np.load(os.path.join(GOLDEN, "greedy.npz"))
    for g in range(int(z["n"])):
        tt, cases = load_golden(str(z[f"g{g}_family"]))
        ep, _pool = cases[int(z[f"g{g}_case"])]
        b = builder_for(tt)
        b.build(ep)
        order, sel = b.greedy_assign()
        assert np.array_equal(order, z[f"g{g}_order"]), g
        assert np.array_equal(sel, z[f"g{g}_sel"]), g
    # larger: cfg2-like random epoch against the oracle
    tt = _grid()
    b = builder_for(tt)
    ep = synth.random_epoch(tt, n_veh=600, n_pending=300, cap=4, mode=marshal.MODE_OSP, seed=11, wait_sec=300, near=True, T=5000)
    pool = b.build_pool(ep)
    order, sel = b.greedy_assign()
    o2, s2 = oracle.greedy_assign(pool, ep.n_veh, ep.n_req)
    assert np.array_equal(order, o2) and np.array_equal(sel, s2)
    assert sel.size == ep.n_veh          # every vehicle keeps exactly one pair (its empty pair at worst)


@pytest.mark.parametrize("tag", ["cfg2", "cfg3"])
def test_cuda_pool_matches_oracle_on_full_size_simulator_snapshots(tag):
    """BASELINE.json configs[1] / configs[2] at full size: 2,000 vehicles, capacity 4 / 8, epoch states recorded
    inside the reference simulator (bench_data/make_snapshots.py).  Bit-exact pool, stats equal."""
    import glob
    import oracle
    from conftest import ROOT
    files = sorted(glob.glob(os.path.join(ROOT, "bench_data", f"{tag}_epoch*.npz")))
    assert files, f"no {tag} snapshots"
    tt = _grid()
    b = builder_for(tt)
    for f in files:
        ep = marshal.Epoch.from_npz_dict(np.load(f), "ep_")
        assert ep.n_veh == 2000 and ep.cap == (4 if tag == "cfg2" else 8)
        ref = oracle.osp_build(ep, tt, n_threads=oracle.max_threads())
        assert ref.error == 0
        got = b.build_pool(ep)
        ok, why = marshal.pools_equal(got, ref, rtol=0.0, check_kind=True)
        assert ok, f"{os.path.basename(f)}: {why}"
        for key in ("n_evals", "n_lookups", "n_tasks", "n_sches_stored"):
            assert got.stats[key] == ref.stats[key], (f, key)
        order, sel = b.greedy_assign()
        o2, s2 = oracle.greedy_assign(got, ep.n_veh, ep.n_req)
        assert np.array_equal(order, o2) and np.array_equal(sel, s2)


def test_lazy_records_with_gpu_greedy_select_the_same_pairs():
    """install(greedy_on_gpu=True) wiring: lazy records, device delays as scores' companion, greedy on the GPU."""
    from test_marshal_cpu import _objects
    from amod_b200 import pool as pool_mod
    import oracle
    tt, cases = load_golden("sim_osp_cap4")
    b = builder_for(tt)
    ep, gold = cases[2]
    reqs, objs, vehs = _objects(ep)
    rids = [int(ep.req_id[i]) for i in ep.search]
    # eager reference-style path
    b.lazy_records = False
    eager = b.compute_candidate_veh_trip_pairs(rids, rids, reqs, vehs, ep.T, 0.1, True, False)
    pool_mod.score_vt_pairs_with_num_of_orders_and_sche_cost(eager, reqs, ep.T)
    sel_e = b.greedy_assignment(eager)
    picked_e = [(eager[i][0].id, [r.id for r in eager[i][1]], eager[i][2], eager[i][3], eager[i][4]) for i in sel_e]
    # lazy path
    b.lazy_records = True
    try:
        lazy = b.compute_candidate_veh_trip_pairs(rids, rids, reqs, vehs, ep.T, 0.1, True, False)
        pool_mod.score_vt_pairs_with_num_of_orders_and_sche_cost(lazy, reqs, ep.T)
        sel_l = b.greedy_assignment(lazy)
        picked_l = [(lazy[i][0].id, [r.id for r in lazy[i][1]], lazy[i][2], lazy[i][3], lazy[i][4]) for i in sel_l]
        assert len(lazy._cache) <= ep.n_veh           # only the selected pairs were ever built
    finally:
        b.lazy_records = False
    assert sel_e == sel_l and picked_e == picked_l
    order, sel = oracle.greedy_assign(gold, ep.n_veh, ep.n_req)
    assert sel.tolist() == sel_e


def test_block_per_trip_classification_path(monkeypatch):
    """Force the heavy-trip list order classification (one block per trip) onto ordinary trips."""
    import oracle
    monkeypatch.setenv("AMOD_HEAVY_MIN", "40")
    tt, cases = load_golden("rand_real_osp_cap4")
    b = PoolBuilder(tt)          # a fresh context reads the override
    for ep, gold in cases[:3]:
        assert gold.ent_nfeas.max() >= 40
        ok, why = marshal.pools_equal(b.build_pool(ep), gold, rtol=0.0, check_kind=True)
        assert ok, why
    b.close()


def test_pinned_fetch_and_repeated_builds_are_identical():
    """fetch(reuse=True) lands in page-locked builder-owned buffers by DMA (AMOD_LOC_PINNED); the result is the
    staged fetch's.  The same epoch built many times in a row (graph replay, both graph branches racing freely)
    gives the same pool every time."""
    import glob
    from conftest import ROOT
    files = sorted(glob.glob(os.path.join(ROOT, "bench_data", "cfg2_epoch*.npz")))
    tt = _grid()
    b = PoolBuilder(tt)
    eps = [marshal.Epoch.from_npz_dict(np.load(f), "ep_") for f in files[:3]]
    firsts = []
    for ep in eps:
        b.build(ep)
        staged = b.fetch(reuse=False)
        pinned = b.fetch(reuse=True)
        ok, why = marshal.pools_equal(pinned, staged, rtol=0.0, check_kind=True)
        assert ok, why
        firsts.append(staged)
    for rep in range(12):
        for ep, first in zip(eps, firsts):
            again = b.build_pool(ep, reuse=True)
            ok, why = marshal.pools_equal(again, first, rtol=0.0, check_kind=True)
            assert ok, f"repeat {rep}: {why}"
            assert again.stats["n_evals"] == first.stats["n_evals"]
    b.close()


def test_candidate_wide_rank_path(monkeypatch):
    """k_gen_emit ranks a row's (j, q) pairs by shuffles when there are at most 32 of them and through memory otherwise.
    Force every row onto the memory path and compare with the reference goldens."""
    monkeypatch.setenv("AMOD_GEN_WIDE", "1")
    for family in ("rand_real_osp_cap4", "rand_ties_osp_cap8"):
        if family not in golden_families(("rand_",)):
            continue
        tt, cases = load_golden(family)
        b = PoolBuilder(tt)          # a fresh context reads the override
        for ep, gold in cases[:3]:
            ok, why = marshal.pools_equal(b.build_pool(ep), gold, rtol=0.0, check_kind=True)
            assert ok, f"{family}: {why}"
        b.close()


def test_trip_lookup_with_colliding_fingerprints(monkeypatch):
    """The trip lookup table keeps a 32-bit fingerprint beside each trip index so that a miss ends at the first probe; a hit is
    still verified against the vehicle and the whole request tuple.  With the fingerprints cut to two bits most probes meet a
    matching fingerprint of ANOTHER trip: the verification must reject it and the probe must go on."""
    monkeypatch.setenv("AMOD_HASH_FP_BITS", "2")
    for family in ("rand_real_osp_cap4", "rand_ties_osp_cap8", "sim_osp_cap4"):
        if family not in golden_families(("rand_", "sim_")):
            continue
        tt, cases = load_golden(family)
        b = PoolBuilder(tt)          # a fresh context reads the override
        for ep, gold in cases[:3]:
            ok, why = marshal.pools_equal(b.build_pool(ep), gold, rtol=0.0, check_kind=True)
            assert ok, f"{family}: {why}"
        b.close()


def test_cuda_pool_at_cfg4_fleet_scale_and_sharded():
    """BASELINE.json configs[3] scale: 8,000 vehicles (bench.py's weak-scaling fleet for N = 4), capacity 4.  The
    whole-fleet pool is bit-exact against the oracle, and the four interleaved rank slices, built one after the
    other on this GPU and merged vehicle-major on the host, give the same pool."""
    import glob
    import oracle
    from conftest import ROOT
    import bench
    from amod_b200 import dist as adist
    f = sorted(glob.glob(os.path.join(ROOT, "bench_data", "cfg2_epoch*.npz")))[2]
    ep = bench.tile_fleet(marshal.Epoch.from_npz_dict(np.load(f), "ep_"), 4)
    assert ep.n_veh == 8000
    tt = _grid()
    b = builder_for(tt)
    ref = oracle.osp_build(ep, tt, n_threads=oracle.max_threads())
    got = b.build_pool(ep)
    ok, why = marshal.pools_equal(got, ref, rtol=0.0, check_kind=True)
    assert ok, why
    assert got.stats["n_evals"] == ref.stats["n_evals"]
    parts = []
    for rank in range(4):
        sub, sel = ep.shard(rank, 4)
        parts.append((b.build_pool(sub), sel))
    merged = adist.merge_vehicle_major(parts, ep.n_veh)
    ok, why = marshal.pools_equal(merged, ref, rtol=0.0, check_kind=True)
    assert ok, why


def test_cuda_cfg5_peak_epoch_snapshot_sba_and_npo_at_named_size():
    """BASELINE.json configs[4] on the committed peak-epoch snapshot (bench_data/cfg5_*: 20,000 vehicles in mixed states from a
    demand-x4 run of the reference simulator, a 5,000-request burst): SBA pool and both NPO matches (the epoch's idle vehicles,
    and all 20,000 x 5,000 = 10^8 pairs) against the oracle fingerprints of bench_data/digests.json."""
    import glob
    import hashlib
    from conftest import ROOT
    f = sorted(glob.glob(os.path.join(ROOT, "bench_data", "cfg5_epoch*.npz")))[0]
    dg = json.load(open(os.path.join(ROOT, "bench_data", "digests.json")))[os.path.basename(f) + "@x1"]
    z = np.load(f)
    ep = marshal.Epoch.from_npz_dict(z, "ep_")
    assert ep.n_veh == 20000 and ep.n_search == 5000 and ep.mode == marshal.MODE_SBA
    b = builder_for(_grid())
    got = b.build_pool(ep)
    assert got.n_entries == dg["n_entries"] and got.stats["n_evals"] == dg["n_evals"] and got.stats["n_screens"] == 100_000_000
    assert marshal.pool_digest(got) == dg["digest"]

    def npo_digest(oi, op, od):
        h = hashlib.blake2b(digest_size=16)
        for a, dt in ((oi, "<i4"), (op, "<i4"), (od, "<f8")):
            h.update(np.ascontiguousarray(a, dtype=np.dtype(dt)).tobytes())
        return h.hexdigest()

    assert npo_digest(*b.npo_match(z["npo_idle_nid"], z["npo_pend_onid"])) == dg["npo_digest"]
    oi, op, od = b.npo_match(ep.veh_nid, z["npo_pend_onid"])          # 20,000 x 5,000
    assert oi.size == 5000 and npo_digest(oi, op, od) == dg["npo_full_digest"]


def test_cuda_dense_demand_x4_vehicle_sample():
    """The dense reading of BASELINE.json configs[3] (8,000 vehicles, demand x4: 8,129 considered requests at epoch 62 of the
    reference simulator, bench_data/cfg4dense_*): every 80th vehicle of the fleet, 44 M schedule evaluations and 353 k pool
    entries for 100 vehicles, against the oracle fingerprint of exactly those vehicles (bench_data/digests.json)."""
    import glob
    from conftest import ROOT
    f = sorted(glob.glob(os.path.join(ROOT, "bench_data", "cfg4dense_epoch*.npz")))[0]
    dg = json.load(open(os.path.join(ROOT, "bench_data", "digests.json")))[os.path.basename(f) + "@x1"]
    ep = marshal.Epoch.from_npz_dict(np.load(f), "ep_")
    step = dg["sample_step"]
    sub = ep.take(np.arange(0, ep.n_veh, step))
    b = builder_for(_grid())
    got = b.build_pool(sub)
    got.ent_veh = (got.ent_veh.astype(np.int64) * step).astype(np.int32)      # back to fleet indices, as the oracle reports them
    assert got.n_entries == dg["n_entries"] and got.stats["n_evals"] == dg["n_evals"] and got.stats["n_lookups"] == dg["n_lookups"]
    assert marshal.pool_digest(got) == dg["digest"]
