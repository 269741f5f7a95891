"""The recorded horizons replayed on the CPU: the oracle stands in for the three device calls, every other line of the
seam path is the product's.  (The same replay with CUDA in the loop is tests/test_gpu_trace.py.)"""
import pytest

from conftest import table_from_spec
import json
import trace_util as tu


def test_cfg1_horizon_replays_through_the_seams_on_cpu():
    z = tu.load_trace("trace_cfg1")
    assert z is not None, "tests/golden/trace_cfg1.npz missing (tests/golden/make_trace.py cfg1)"
    tt = table_from_spec(json.loads(str(z["spec"])))
    n = tu.replay_cfg1(tu.oracle_backed_builder(tt), z, every=3)
    assert n >= 80


def test_cfg2_day_samples_replay_on_cpu():
    z = tu.load_trace("trace_cfg2day")
    if z is None or "spec" not in z.files:
        pytest.skip("tests/golden/trace_cfg2day.npz not recorded (or still being written)")
    tt = table_from_spec(json.loads(str(z["spec"])))
    epochs = z["epochs"].tolist()
    n = tu.replay_cfg2day(tu.oracle_backed_builder(tt), z, only=set(epochs[::6]))
    assert n >= 1
