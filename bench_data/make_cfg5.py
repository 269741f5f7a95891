"""BASELINE.json configs[4]: one peak-epoch snapshot for the SBA vehicle-request matrix + NPO nearest-idle-vehicle
match at 20,000 vehicles x 5,000 pending requests ("mixed idle/working states drawn from a cfg4-style
run, tiled").

    ADVANCE=sba SNAP=25 python bench_data/make_snapshots.py 8000 4 108000 cfg5src     # fleet state (reference simulator)
    python bench_data/make_cfg5.py                                                     # -> bench_data/cfg5_epoch0025.npz

Fleet: the 8,000-vehicle fleet of a demand-x4 run of the reference simulator at epoch 25 (warm-up: idle, working and
rebalancing vehicles mixed), tiled to 20,000 vehicles (synth.tile_fleet).  Requests: a burst of 5,000 new requests in
the epoch's 30 s (origins/destinations from the same generator as the day's stream), Clp = Tr + 420, Cld = Tr + Ts + 840
(request.py:33-34): the SBA search list (new_received_rids, dispatch_sba.py:56).  NPO: the fleet's idle vehicles x the
5,000 pending origins (rebalancing_npo.py:26-38).  n_evals / n_lookups are filled by bench_data/make_digests.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from amod_b200 import marshal, synth  # noqa: E402


def main():
    graph = synth.make_road_graph(seed=0)
    z = np.load(os.path.join(HERE, "cfg5src_epoch0025.npz"))
    src = marshal.Epoch.from_npz_dict(z, "ep_")
    fleet = synth.tile_fleet(src, 3).take(np.arange(20000))
    T = src.T
    n_new = 5000
    on, dn, tm = synth.make_request_stream(graph, n_new, T - 30, T, seed=5, mean_trip_sec=600.0)
    ts = graph.tt[on - 1, dn - 1].astype(np.float64)
    tr = tm.astype(np.float64)
    base = src.n_req
    kw = {f: getattr(fleet, f) for f in marshal.Epoch.FIELDS}
    kw["req_id"] = np.concatenate([src.req_id, src.req_id.max() + 1 + np.arange(n_new)]).astype(np.int32)
    kw["req_onid"] = np.concatenate([src.req_onid, on]).astype(np.int32)
    kw["req_dnid"] = np.concatenate([src.req_dnid, dn]).astype(np.int32)
    kw["req_tr"] = np.concatenate([src.req_tr, tr])
    kw["req_ts"] = np.concatenate([src.req_ts, ts])
    kw["req_clp"] = np.concatenate([src.req_clp, tr + 420.0])
    kw["req_cld"] = np.concatenate([src.req_cld, tr + ts + 840.0])
    kw["search"] = (base + np.arange(n_new)).astype(np.int32)
    ep = marshal.Epoch(T=T, cap=src.cap, mode=marshal.MODE_SBA, **kw)
    idle = ep.veh_nid[ep.veh_status == marshal.ST_IDLE].astype(np.int32)
    pend = ep.req_onid[ep.search].astype(np.int32)
    print(f"cfg5: V={ep.n_veh} new={ep.n_search} idle={idle.size} status counts {np.bincount(ep.veh_status).tolist()}")
    np.savez_compressed(os.path.join(HERE, "cfg5_epoch0025.npz"), **ep.to_npz_dict("ep_"), n_evals=0, n_lookups=0, n_entries=0,
                        npo_idle_nid=idle, npo_pend_onid=pend)


if __name__ == "__main__":
    main()
