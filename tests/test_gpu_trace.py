"""North-star target run with CUDA in the loop: recorded horizons (tests/golden/make_trace.py) replayed through the
product's seams on the GPU — PoolBuilder.compute_candidate_* (marshal -> amod_*_build -> amod_pool_fetch), the scorer
on device delays, amod_greedy_assign, amod_npo_match — epoch after epoch on one context (captured graph reused, buffers
regrown, modes alternating).  Bit-exact pools and selections are required at every epoch.
  cfg1     all 258 epochs of BASELINE.json configs[0] against what the UNMODIFIED reference returned
  cfg2day  the sampled epochs of a whole synthetic day of configs[1] (SBA warm-up, OSP main horizon, both peaks)"""
import json

import pytest

from conftest import table_from_spec
from amod_b200.pool import PoolBuilder
import trace_util as tu

pytestmark = pytest.mark.gpu


def test_cfg1_whole_horizon_on_cuda():
    z = tu.load_trace("trace_cfg1")
    assert z is not None
    b = PoolBuilder(table_from_spec(json.loads(str(z["spec"]))))
    n = tu.replay_cfg1(b, z)
    assert n >= 250
    b.close()


def test_cfg2_whole_day_samples_on_cuda():
    z = tu.load_trace("trace_cfg2day")
    if z is None or "spec" not in z.files:
        pytest.skip("tests/golden/trace_cfg2day.npz not recorded (or still being written)")
    b = PoolBuilder(table_from_spec(json.loads(str(z["spec"]))))
    n = tu.replay_cfg2day(b, z)
    assert n == len(z["epochs"])
    b.close()
