"""SURVEY.md 8f-2 on the GPU: amod_routes_build through the C-ABI against the reference's own outputs (routes.npz) and, on
the full-size 4,091-node graph, against the oracle for a whole epoch's worth of legs."""
import numpy as np
import pytest

from amod_b200 import routes as aroutes, synth
from amod_b200.pool import PoolBuilder
from test_routes_cpu import _golden, check_routes_against_golden

pytestmark = pytest.mark.gpu


def test_cuda_routes_match_reference_golden():
    z, g = _golden()
    b = PoolBuilder(g.tt)
    rb = aroutes.RouteBuilder(b, g.pred, g.dist, z["node_geo"])
    check_routes_against_golden(rb.build(z["leg_o"], z["leg_d"]), z)
    dur, dist, steps = rb.build_route_from_origin_to_dest(int(z["leg_o"][20]), int(z["leg_d"][20]))
    assert dur == z["leg_t"][20] and dist == z["leg_d_sum"][20] and len(steps) == int(z["leg_off"][21] - z["leg_off"][20])
    b.close()


def test_cuda_routes_full_size_epoch_against_oracle():
    import oracle
    g = synth.make_road_graph(seed=0, with_paths=True)
    b = PoolBuilder(g.tt)
    rb = aroutes.RouteBuilder(b, g.pred, g.dist)
    rng = np.random.default_rng(11)
    o = rng.integers(1, g.n_nodes + 1, size=10000).astype(np.int32)     # 2,000 vehicles x ~5 legs
    d = rng.integers(1, g.n_nodes + 1, size=10000).astype(np.int32)
    got = rb.build(o, d)
    off, u, v, t, dd, lt, ld = oracle.routes_build(g.pred, g.tt, g.dist, o[:1500], d[:1500])
    n = int(off[-1])
    assert np.array_equal(got.leg_off[:1501], off) and np.array_equal(got.step_u[:n], u) and np.array_equal(got.step_v[:n], v)
    assert np.array_equal(got.step_t[:n], t) and np.array_equal(got.step_d[:n], dd)
    assert np.array_equal(got.leg_t[:1500], lt) and np.array_equal(got.leg_d[:1500], ld)
    # size-independent property at full size: a leg's duration / distance equal the tables' entries (the reference asserts
    # |sum - table| <= 0.005, route_functions.py:64-68; here the 1/8 s grid makes it exact)
    assert np.array_equal(got.leg_t, g.tt[o - 1, d - 1].astype(np.float64))
    assert np.allclose(got.leg_d, g.dist[o - 1, d - 1].astype(np.float64), atol=0.005)
    b.close()
