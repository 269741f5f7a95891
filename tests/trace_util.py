"""Replay of recorded horizons (tests/golden/trace_*.npz, made by tests/golden/make_trace.py) through the product's
seams: PoolBuilder.compute_candidate_veh_{req,trip}_pairs -> scorer -> greedy_assignment, and the NPO match.  The builder
is the CUDA PoolBuilder in the -m gpu tests and an oracle-backed double of its three device calls in the CPU tests."""
import hashlib
import json
import os

import numpy as np

from conftest import GOLDEN
from amod_b200 import marshal, pool as pool_mod
from test_marshal_cpu import _objects


def digest_arrays(*arrs) -> str:
    h = hashlib.blake2b(digest_size=16)
    for a, dt in arrs:
        a = np.ascontiguousarray(a, dtype=np.dtype(dt))
        h.update(np.int64(a.size).tobytes())
        h.update(a.tobytes())
    return h.hexdigest()


def oracle_backed_builder(tt):
    """PoolBuilder whose three device calls (pool build, greedy assignment, NPO match) are served by the oracle: every
    other line of the seam path (marshalling, lazy records, scorer, host wrappers) is the product's own."""
    import oracle

    class OracleBackedBuilder(pool_mod.PoolBuilder):
        def __init__(self):
            self.last_stats = {}
            self.lazy_records = True

        def build_pool(self, ep, reuse=False):
            self._last = ((oracle.sba_build if ep.mode == marshal.MODE_SBA else oracle.osp_build)(ep, tt), ep)
            return self._last[0]

        def greedy_assign(self):
            pool, ep = self._last
            return oracle.greedy_assign(pool, ep.n_veh, ep.n_req)

        def npo_match(self, idle_nid, pend_onid):
            return oracle.npo_match(tt, np.asarray(idle_nid, dtype=np.int32), np.asarray(pend_onid, dtype=np.int32))

        def close(self):
            pass

    return OracleBackedBuilder()


def replay_cfg1(b, z, every: int = 1) -> int:
    """Every recorded epoch of the cfg1 horizon (SBA + greedy + NPO, 100 vehicles): the pool the seam returns must equal
    the pool the UNMODIFIED reference returned (bit-exact, delays included), the greedy selection must be the
    reference's, the NPO matches the reference's.  Returns the number of epochs checked."""
    n = 0
    b.lazy_records = True
    for i in range(0, int(z["n_epochs"]), every):
        if f"e{i}_ep_veh_nid" not in z.files:
            continue
        ep = marshal.Epoch.from_npz_dict(z, f"e{i}_ep_")
        gold = marshal.Pool.from_npz_dict(z, f"e{i}_pool_")
        reqs, objs, vehs = _objects(ep)
        rids = [int(ep.req_id[k]) for k in ep.search]
        recs = b.compute_candidate_veh_req_pairs(rids, reqs, vehs, ep.T)
        ok, why = marshal.pools_equal(recs._pool, gold, rtol=0.0, check_kind=True)
        assert ok, f"cfg1 epoch {i}: {why}"
        pool_mod.score_vt_pairs_with_num_of_orders_and_sche_cost(recs, reqs, ep.T)
        sel = b.greedy_assignment(recs)
        assert sel == z[f"e{i}_sel"].tolist(), f"cfg1 epoch {i}: greedy selection differs"
        assert [recs[s][0].id for s in sel] == z[f"e{i}_sel_veh"].tolist(), i
        assert [r.id for s in sel for r in recs[s][1]] == z[f"e{i}_sel_rid"].tolist(), i
        if f"e{i}_npo_idle_nid" in z.files and z[f"e{i}_npo_idle_nid"].size and z[f"e{i}_npo_pend_onid"].size:
            idle, pend, T = z[f"e{i}_npo_idle_nid"], z[f"e{i}_npo_pend_onid"], int(z[f"e{i}_T"])
            oi, op, od = b.npo_match(idle, pend)
            order = np.argsort(oi)
            assert np.array_equal(oi[order], z[f"e{i}_npo_m_idle"]), i
            assert np.array_equal(pend[op[order]], z[f"e{i}_npo_m_node"]), i
            assert np.array_equal(T + od[order] + 240, z[f"e{i}_npo_m_ddl"]), i
        n += 1
    return n


def replay_cfg2day(b, z, only=None) -> int:
    """The sampled epochs of the whole synthetic day (OSP + greedy, 2,000 vehicles): pool fingerprint, greedy order and
    selection, NPO matches, all against the values recorded when the day was simulated."""
    n = 0
    b.lazy_records = True
    for e in z["epochs"].tolist():
        if only is not None and e not in only:
            continue
        exp = json.loads(str(z[f"e{e}_expect"]))
        ep = marshal.Epoch.from_npz_dict(z, f"e{e}_ep_")
        reqs, objs, vehs = _objects(ep)
        rids = [int(ep.req_id[k]) for k in ep.search]
        if ep.mode == marshal.MODE_SBA:
            recs = b.compute_candidate_veh_req_pairs(rids, reqs, vehs, ep.T)
        else:
            recs = b.compute_candidate_veh_trip_pairs(rids, rids, reqs, vehs, ep.T, 0.1, ep.mode == marshal.MODE_OSP, False)
        assert len(recs) == exp["n_entries"], (e, len(recs), exp["n_entries"])
        assert marshal.pool_digest(recs._pool) == exp["pool_digest"], f"cfg2 day epoch {e}: pool differs"
        pool_mod.score_vt_pairs_with_num_of_orders_and_sche_cost(recs, reqs, ep.T)
        order, sel = b.greedy_assign()
        assert digest_arrays((order, "<i4")) == exp["order_digest"], e
        assert np.array_equal(sel, z[f"e{e}_sel"]), e
        if f"e{e}_npo_idle_nid" in z.files:
            idle, pend = z[f"e{e}_npo_idle_nid"], z[f"e{e}_npo_pend_onid"]
            if idle.size and pend.size:
                oi, op, od = b.npo_match(idle, pend)
            else:
                oi, op, od = np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0)
            assert digest_arrays((oi, "<i4"), (op, "<i4"), (od, "<f8")) == str(z[f"e{e}_npo_digest"]), e
        n += 1
    return n


def load_trace(name):
    path = os.path.join(GOLDEN, name + ".npz")
    return np.load(path) if os.path.exists(path) else None
