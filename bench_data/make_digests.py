"""Oracle fingerprints of every bench snapshot (and of its tiled fleets), so that bench.py can state parity in
its JSON line at every N WITHOUT executing the oracle: bench_data/digests.json.

    python bench_data/make_digests.py

For each bench_data/<tag>_epoch*.npz and each fleet multiple n in TILES[tag] the C oracle (oracle/amod_oracle.c,
pinned to the reference by tests/test_oracle_golden.py) builds the pool of synth.tile_fleet(snapshot, n); the entry
"<file>@x<n>" records marshal.pool_digest(pool) plus the counters the kernels must reproduce.  OSP snapshots are
built with oracle.osp_build, SBA snapshots (cfg5) with oracle.sba_build; cfg5 also records the NPO match digest.
"""
from __future__ import annotations

import glob
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from amod_b200 import marshal, synth  # noqa: E402
import oracle  # noqa: E402

TILES = {"cfg2": (1, 2, 4, 8, 16), "cfg3": (1, 2, 4, 8), "cfg4dense": (1,), "cfg5": (1,)}
# cfg4dense (8,000 vehicles, demand x4, SURVEY.md 8d's reading of configs[3]) is ~3.5 G schedule evaluations per epoch AND holds
# feasible trips of size cap+4, where the reference raises IndexError (dispatch_osp.py:8,226): it cannot be built exhaustively,
# neither by the reference nor here (profiles/r02_cfg4_demand_x4_full_fleet_raises_indexerror.log).  Its every-80th-vehicle
# sample (100 vehicles x 8,129 requests, 44 M evaluations, no size-8 trip) is kept as a dense parity case: vehicles are
# independent (dispatch_osp.py:107-111), so the oracle builds just those vehicles (tests/test_gpu_parity.py).
SAMPLE = {"cfg4dense": 80}


def npo_digest(oi, op, od) -> str:
    h = hashlib.blake2b(digest_size=16)
    for a, dt in ((oi, "<i4"), (op, "<i4"), (od, "<f8")):
        h.update(np.ascontiguousarray(a, dtype=np.dtype(dt)).tobytes())
    return h.hexdigest()


def main():
    tt = synth.make_road_graph(seed=0).tt
    only = set(sys.argv[1:])
    try:
        previous = json.load(open(os.path.join(HERE, "digests.json")))
    except Exception:
        previous = {}
    out = {}
    nth = oracle.max_threads()
    for f in sorted(glob.glob(os.path.join(HERE, "*_epoch*.npz"))):
        name = os.path.basename(f)
        tag = name.split("_")[0]
        if tag not in TILES:
            continue
        if only and tag not in only:
            out.update({k: v for k, v in previous.items() if k.startswith(name)})
            continue
        z = np.load(f)
        ep = marshal.Epoch.from_npz_dict(z, "ep_")
        for n in TILES.get(tag, (1,)):
            full = synth.tile_fleet(ep, n)
            step = SAMPLE.get(tag, 1)
            if full.mode == marshal.MODE_SBA:
                pool = oracle.sba_build(full, tt, n_threads=nth)
            else:
                pool = oracle.osp_build(full, tt, n_threads=nth, veh_lo=0, veh_step=step)
            assert pool.error == 0
            rec = {"digest": marshal.pool_digest(pool), "sample_step": step, "n_entries": int(pool.n_entries), "n_evals": int(pool.stats["n_evals"]),
                   "n_lookups": int(pool.stats["n_lookups"]), "n_veh": int(full.n_veh), "n_search": int(full.n_search)}
            if "npo_idle_nid" in z.files and n == 1:
                oi, op, od = oracle.npo_match(tt, z["npo_idle_nid"], z["npo_pend_onid"])
                rec["npo_digest"] = npo_digest(oi, op, od)
                rec["npo_matches"] = int(oi.size)
                # the named peak size (BASELINE.json configs[4]): EVERY vehicle's node as an idle vehicle x the 5,000 origins = 10^8 pairs
                oi, op, od = oracle.npo_match(tt, ep.veh_nid, z["npo_pend_onid"])
                rec["npo_full_digest"] = npo_digest(oi, op, od)
                rec["npo_full_pairs"] = int(ep.n_veh) * int(z["npo_pend_onid"].size)
            out[f"{name}@x{n}"] = rec
            print(name, n, rec, flush=True)
    json.dump(out, open(os.path.join(HERE, "digests.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
