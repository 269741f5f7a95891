"""Golden vectors for the 'next' row greedy_assignment (ilp_assign.py:101-130): the unmodified reference
function applied to pools of the existing golden families (scored as the reference's scorer leaves them:
pair[4] = len(trip), scheduling.py:160-161).  Needs /root/reference.

    python tests/golden/make_golden_greedy.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))

from amod_b200 import synth  # noqa: E402
import ref_harness as rh  # noqa: E402
from conftest import load_golden  # noqa: E402


def main():
    g = synth.make_road_graph(16, seed=1, with_paths=True)
    ref = rh.setup_reference("/tmp/amod_ref_greedy", g, np.array([1, 2], dtype=np.int32), np.array([2, 3], dtype=np.int32),
                             np.array([70000, 80000]), fleet_size=4, capacity=4)
    out = {}
    n = 0
    for family in ("rand_real_osp_cap4", "rand_ties_osp_cap4", "sim_osp_cap4", "rand_ties_sba_cap4", "rand_real_nr_cap4", "sim_sba_cap6"):
        _tt, cases = load_golden(family)
        for ci, (ep, pool) in enumerate(cases[:3]):
            vehs = [types.SimpleNamespace(id=v) for v in range(ep.n_veh)]
            reqs = [types.SimpleNamespace(id=int(r)) for r in ep.req_id]
            recs = []
            for e in range(pool.n_entries):
                trip = [reqs[i] for i in pool.trip_req[pool.ent_trip_off[e]:pool.ent_trip_off[e + 1]]]
                recs.append([vehs[pool.ent_veh[e]], trip, None, 0.0, len(trip)])
            orig = {id(r): e for e, r in enumerate(recs)}
            sel = ref.ilp_assign.greedy_assignment(recs)      # sorts recs in place, returns positions in the sorted list
            out[f"g{n}_family"] = np.array(family); out[f"g{n}_case"] = np.array(ci)
            out[f"g{n}_order"] = np.array([orig[id(r)] for r in recs], dtype=np.int32)
            out[f"g{n}_sel"] = np.array(sel, dtype=np.int32)
            n += 1
    out["n"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "greedy.npz"), **out)
    print("wrote greedy.npz with", n, "cases")


if __name__ == "__main__":
    main()
