"""Host-side cost of the Python seam per epoch (no GPU needed): Veh/Req objects -> SoA epoch (marshal_epoch),
pool arrays -> reference-style records (eager / lazy).  The pool itself comes from the oracle here; on a GPU box
the C-ABI call takes its place.  Numbers quoted in BASELINE.md §4.   python scripts/prof_seam_cpu.py [cfg2|cfg3]"""
import glob
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from amod_b200 import marshal, synth  # noqa: E402
import oracle  # noqa: E402  (test infrastructure: stands in for the CUDA library on a CPU-only host)


class _Status:
    def __init__(self, v):
        self.value = v


def objects_of(ep):
    """Duck-typed Veh/Req objects carrying what the seams read (vehicle.py:41-70, request.py:25-41)."""
    objs = [types.SimpleNamespace(id=int(ep.req_id[i]), onid=int(ep.req_onid[i]), dnid=int(ep.req_dnid[i]), Clp=float(ep.req_clp[i]),
                                  Cld=float(ep.req_cld[i]), Tr=float(ep.req_tr[i]), Ts=float(ep.req_ts[i]), status=_Status(1))
            for i in range(ep.n_req)]
    reqs = {r.id: r for r in objs}
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
        vehs.append(types.SimpleNamespace(id=v, nid=int(ep.veh_nid[v]), t_to_nid=float(ep.veh_t_to_nid[v]), load=int(ep.veh_load[v]),
                                          status=_Status(int(ep.veh_status[v])), sche=sche, onboard_rids=onboard, K=ep.cap))
    return reqs, vehs


def median_ms(fn, n=15):
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t) * 1e3)
    return float(np.median(ts))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    f = sorted(glob.glob(os.path.join(ROOT, "bench_data", f"{tag}_epoch*.npz")))[2]
    ep = marshal.Epoch.from_npz_dict(np.load(f), "ep_")
    tt = synth.make_road_graph(seed=0).tt
    reqs, vehs = objects_of(ep)
    search = [int(ep.req_id[i]) for i in ep.search]
    table = marshal.ReqTable()
    ep2, req_objs = marshal.marshal_epoch(reqs, vehs, search, ep.T, ep.cap, ep.mode, req_table=table)
    pool = oracle.osp_build(ep2, tt, n_threads=oracle.max_threads())
    print(f"{tag}: {ep2.n_veh} vehicles, {ep2.n_req} live requests, {len(search)} considered, pool {pool.n_entries} entries")
    print(f"  marshal_epoch (request table warm) {median_ms(lambda: marshal.marshal_epoch(reqs, vehs, search, ep.T, ep.cap, ep.mode, req_table=table)):7.2f} ms")
    print(f"  marshal_epoch (no request table)   {median_ms(lambda: marshal.marshal_epoch(reqs, vehs, search, ep.T, ep.cap, ep.mode)):7.2f} ms")
    print(f"  pool_to_records (eager list)       {median_ms(lambda: marshal.pool_to_records(pool, ep2, vehs, req_objs), 7):7.2f} ms")
    print(f"  LazyPoolRecords (lazy view)        {median_ms(lambda: marshal.LazyPoolRecords(pool, ep2, vehs, req_objs)):7.2f} ms")


if __name__ == "__main__":
    main()
