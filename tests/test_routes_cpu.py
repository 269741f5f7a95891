"""SURVEY.md 8f-2 (route construction): the oracle's restatement of route_functions.build_route_from_origin_to_dest against
the reference's own outputs (tests/golden/routes.npz), and the host-side adapter that serves the reference's
Veh.build_route from a batch (amod_b200/routes.py) with the oracle standing in for the device call."""
import json
import os
import types

import numpy as np

from conftest import GOLDEN
from amod_b200 import routes as aroutes, synth
import oracle


def _golden():
    z = np.load(os.path.join(GOLDEN, "routes.npz"))
    spec = json.loads(str(z["spec"]))
    g = synth.make_road_graph(spec["n_nodes"], seed=spec["seed"], with_paths=True)
    return z, g


def check_routes_against_golden(r, z):
    for name, gold in (("leg_off", "leg_off"), ("step_u", "step_u"), ("step_v", "step_v"), ("step_t", "step_t"), ("step_d", "step_d"),
                       ("leg_t", "leg_t"), ("leg_d", "leg_d_sum")):
        got = r[name] if isinstance(r, dict) else getattr(r, name)
        assert np.array_equal(np.asarray(got), z[gold]), name


def test_oracle_routes_match_reference_golden():
    z, g = _golden()
    off, u, v, t, d, lt, ld = oracle.routes_build(g.pred, g.tt, g.dist, z["leg_o"], z["leg_d"])
    check_routes_against_golden(dict(leg_off=off, step_u=u, step_v=v, step_t=t, step_d=d, leg_t=lt, leg_d=ld), z)
    assert (np.diff(off)[:12] == 1).all()            # origin == destination: only the closing step


class OracleRouteBuilder(aroutes.RouteBuilder):
    """RouteBuilder whose device call is served by the oracle (CPU tests)."""

    def __init__(self, g, node_geo):
        self.g, self.node_geo, self._cache, self.n_batches = g, np.asarray(node_geo, dtype=np.float64), {}, 0

    def build(self, leg_o, leg_d):
        self.n_batches += 1
        return aroutes.Routes(*oracle.routes_build(self.g.pred, self.g.tt, self.g.dist, leg_o, leg_d))


def test_adapter_returns_the_reference_record():
    z, g = _golden()
    rb = OracleRouteBuilder(g, z["node_geo"])
    rb.prefetch(zip(z["leg_o"][:50].tolist(), z["leg_d"][:50].tolist()))
    assert rb.n_batches == 1
    for k in range(50):
        dur, dist, steps = rb.build_route_from_origin_to_dest(int(z["leg_o"][k]), int(z["leg_d"][k]))
        a, b = int(z["leg_off"][k]), int(z["leg_off"][k + 1])
        assert dur == z["leg_t"][k] and dist == z["leg_d_sum"][k] and len(steps) == b - a
        for s, (t, d, uv, geo) in enumerate(steps):
            assert t == z["step_t"][a + s] and d == z["step_d"][a + s] and uv == [int(z["step_u"][a + s]), int(z["step_v"][a + s])]
            assert geo == [z["node_geo"][uv[0] - 1].tolist(), z["node_geo"][uv[1] - 1].tolist()]
        assert steps[-1][0] == 0.0 and steps[-1][2][0] == steps[-1][2][1] == int(z["leg_d"][k])
    assert rb.n_batches == 1                           # everything came from the one batch


def test_legs_of_selected_pairs_follow_build_route():
    """vehicle.py:231-302: the chain starts after the unfinished step, a rebalancing stop is dropped when a trip is assigned."""
    st = types.SimpleNamespace
    v1 = st(status=st(name="IDLE"), nid=5, step_to_nid=None)
    v2 = st(status=st(name="REBALANCING"), nid=9, step_to_nid=st(nid_pair=[9, 11]))
    pairs = [[v1, [], [(3, 1, 20, 0.0), (3, -1, 30, 0.0)], 0.0, 0.0], [v2, [], [(-1, 0, 40, 0.0), (4, 1, 50, 0.0), (4, -1, 60, 0.0)], 0.0, 0.0]]
    assert aroutes.legs_of_selected_pairs(pairs, [0, 1]) == [(5, 20), (20, 30), (11, 50), (50, 60)]
