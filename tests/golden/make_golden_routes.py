"""tests/golden/routes.npz — route_functions.build_route_from_origin_to_dest (route_functions.py:36-70) of the UNMODIFIED
reference on seeded (origin, destination) pairs of two synthetic road graphs (needs /root/reference)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from amod_b200 import synth  # noqa: E402
import ref_harness as rh  # noqa: E402


def main():
    spec = dict(kind="grid", n_nodes=700, seed=8)
    g = synth.make_road_graph(spec["n_nodes"], seed=spec["seed"], with_paths=True)
    ref = rh.setup_reference("/tmp/amod_ref_routes", g, np.array([1, 2], dtype=np.int32), np.array([2, 3], dtype=np.int32),
                             np.array([70000, 80000]), fleet_size=4, capacity=4)
    rf = ref.route_functions
    rng = np.random.default_rng(3)
    o = rng.integers(1, g.n_nodes + 1, size=400)
    d = rng.integers(1, g.n_nodes + 1, size=400)
    d[:12] = o[:12]                                   # origin == destination: only the closing step
    off, U, V, T, D, LT, LD = [0], [], [], [], [], [], []
    for a, b in zip(o.tolist(), d.tolist()):
        duration, distance, steps = rf.build_route_from_origin_to_dest(a, b)
        for (t, dd, (u, v), _geo) in steps:
            U.append(u); V.append(v); T.append(float(t)); D.append(float(dd))
        off.append(len(U)); LT.append(float(duration)); LD.append(float(distance))
    geo = np.array([rf.get_node_geo(n) for n in range(1, g.n_nodes + 1)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "routes.npz"), spec=np.array(json.dumps(spec)), leg_o=o.astype(np.int32), leg_d=d.astype(np.int32),
                        leg_off=np.asarray(off, dtype=np.int64), step_u=np.asarray(U, dtype=np.int32), step_v=np.asarray(V, dtype=np.int32),
                        step_t=np.asarray(T), step_d=np.asarray(D), leg_t=np.asarray(LT), leg_d_sum=np.asarray(LD), node_geo=geo)
    print(f"{len(LT)} legs, {len(U)} steps")


if __name__ == "__main__":
    main()
