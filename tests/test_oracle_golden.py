"""Pins the C oracle (oracle/amod_oracle.c) to the reference: every golden vector in
tests/golden/ was produced by the unmodified reference (tests/golden/make_golden.py).
Bit-exact on structure, order, chosen schedules, costs and delays."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_families, load_golden, table_from_spec
from amod_b200 import marshal
import oracle


@pytest.mark.parametrize("family", golden_families(("rand_", "sim_")))
def test_oracle_matches_reference_pools(family):
    tt, cases = load_golden(family)
    assert cases
    for i, (ep, gold) in enumerate(cases):
        build = oracle.sba_build if ep.mode == marshal.MODE_SBA else oracle.osp_build
        got = build(ep, tt, n_threads=2)
        assert got.error == 0
        ok, why = marshal.pools_equal(got, gold, rtol=0.0, check_kind=True)
        assert ok, f"{family} case {i}: {why}"


def test_oracle_level_overflow_mirrors_indexerror():
    tt, cases = load_golden("levels")
    (ep_ok, gold_ok), (ep_bad, gold_bad) = cases
    got = oracle.osp_build(ep_ok, tt)
    assert got.error == 0
    ok, why = marshal.pools_equal(got, gold_ok, check_kind=True)
    assert ok, why
    assert gold_bad is None                      # the reference raised IndexError (dispatch_osp.py:226)
    assert oracle.osp_build(ep_bad, tt).error == -4


def test_oracle_test_constraints_unit_vectors():
    z = np.load(os.path.join(GOLDEN, "unit.npz"))
    tt = table_from_spec(json.loads(str(z["spec"])))
    eps = [marshal.Epoch.from_npz_dict(z, f"c{i}_ep_") for i in range(int(z["n_cases"]))]
    seen = set()
    for (case, v, q, i, j, viol), cost in zip(z["rows"].tolist(), z["costs"].tolist()):
        ep = eps[case]
        base = ep.sche_code[ep.veh_sche_off[v]:ep.veh_sche_off[v + 1]].tolist()
        base.insert(i, 2 * q)
        base.insert(j, 2 * q + 1)
        got_viol, got_cost = oracle.test_constraints(ep, tt, v, base, q, i, j)
        assert got_viol == viol
        assert got_cost == cost or (np.isinf(cost) and np.isinf(got_cost))
        seen.add(viol)
    assert seen == {-1, 0, 1, 2, 3, 4}


@pytest.mark.parametrize("family", [f for f in golden_families(("npo_",))])
def test_oracle_npo_matches_reference(family):
    z = np.load(os.path.join(GOLDEN, family + ".npz"))
    tt = table_from_spec(json.loads(str(z["spec"])))
    n = int(z["n_cases"])
    assert n > 0
    for i in range(n):
        idle, pend, T = z[f"c{i}_idle_nid"], z[f"c{i}_pend_onid"], int(z[f"c{i}_T"])
        oi, op, od = oracle.npo_match(tt, idle, pend)
        assert oi.size == min(idle.size, pend.size)
        order = np.argsort(oi)
        assert np.array_equal(oi[order], z[f"c{i}_m_idle"])
        assert np.array_equal(pend[op[order]], z[f"c{i}_m_node"])
        assert np.array_equal(T + od[order] + 240, z[f"c{i}_m_ddl"])


def test_oracle_thread_count_does_not_change_result():
    tt, cases = load_golden("rand_real_osp_cap4")
    ep = cases[0][0]
    a = oracle.osp_build(ep, tt, n_threads=1)
    b = oracle.osp_build(ep, tt, n_threads=4)
    ok, why = marshal.pools_equal(a, b, check_kind=True)
    assert ok, why
    assert a.stats["n_evals"] == b.stats["n_evals"] and a.stats["n_lookups"] == b.stats["n_lookups"]


def test_oracle_greedy_assignment_matches_reference():
    z = np.load(os.path.join(GOLDEN, "greedy.npz"))
    assert int(z["n"]) > 0
    for g in range(int(z["n"])):
        _tt, cases = load_golden(str(z[f"g{g}_family"]))
        ep, pool = cases[int(z[f"g{g}_case"])]
        order, sel = oracle.greedy_assign(pool, ep.n_veh, ep.n_req)
        assert np.array_equal(order, z[f"g{g}_order"])
        assert np.array_equal(sel, z[f"g{g}_sel"])


def test_bench_digests_are_the_oracles():
    """bench.py states parity against bench_data/digests.json without executing the oracle: the stored fingerprints of the
    plain (x1) cfg2 / cfg3 snapshots must be what the oracle produces today."""
    import glob
    import json
    from conftest import ROOT
    from amod_b200 import marshal, synth
    import oracle
    dg = json.load(open(os.path.join(ROOT, "bench_data", "digests.json")))
    tt = synth.make_road_graph(seed=0).tt
    files = sorted(glob.glob(os.path.join(ROOT, "bench_data", "cfg2_epoch*.npz")))[:3] + sorted(glob.glob(os.path.join(ROOT, "bench_data", "cfg3_epoch*.npz")))[:1]
    for f in files:
        ep = marshal.Epoch.from_npz_dict(np.load(f), "ep_")
        pool = oracle.osp_build(ep, tt, n_threads=oracle.max_threads())
        rec = dg[os.path.basename(f) + "@x1"]
        assert marshal.pool_digest(pool) == rec["digest"] and pool.n_entries == rec["n_entries"] and pool.stats["n_evals"] == rec["n_evals"], f
