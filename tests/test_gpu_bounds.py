"""Memory safety without compute-sanitizer (closed on this GPU pool): the -DAMOD_BOUNDS build of the same sources guards
every store into a level, pool, task or record buffer (common.cuh BOK) and the host turns a violated guard into an error.
The golden families, a cap-8 epoch and a full-size snapshot run under it from a FRESH context (all buffers start at their
small default capacities, so every regrow / rerun path is exercised) and must be bit-exact with no guard hit."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT, golden_families, load_golden
from amod_b200 import marshal, synth
from amod_b200.pool import PoolBuilder

pytestmark = pytest.mark.gpu
LIB = os.path.join(ROOT, "amod_b200", "csrc", "libamod_b200_bounds.so")


@pytest.mark.parametrize("family", golden_families(("rand_", "sim_")))
def test_golden_families_under_guarded_stores(family):
    tt, cases = load_golden(family)
    b = PoolBuilder(tt, lib_path=LIB)
    assert b"AMOD_BOUNDS" in b.lib.amod_version()
    for i, (ep, gold) in enumerate(cases):
        ok, why = marshal.pools_equal(b.build_pool(ep), gold, rtol=0.0, check_kind=True)
        assert ok, f"{family} case {i}: {why}"
    b.close()


def test_full_size_snapshots_under_guarded_stores():
    import oracle
    tt = synth.make_road_graph(seed=0).tt
    b = PoolBuilder(tt, lib_path=LIB)
    for tag in ("cfg2", "cfg3"):
        f = sorted(glob.glob(os.path.join(ROOT, "bench_data", f"{tag}_epoch*.npz")))[0]
        ep = marshal.Epoch.from_npz_dict(np.load(f), "ep_")
        ref = oracle.osp_build(ep, tt, n_threads=oracle.max_threads())
        got = b.build_pool(ep)
        ok, why = marshal.pools_equal(got, ref, rtol=0.0, check_kind=True)
        assert ok, f"{tag}: {why}"
        assert got.stats["n_attempts"] >= 1
    b.close()
