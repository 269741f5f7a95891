"""the epoch state resident in HBM, updated by deltas (amod_state_*, amod_b200/state.py), replayed over the
recorded horizons: after every delta upload the device rows equal what marshal_epoch produced for that epoch, and the pool
built from the resident state equals the pool built from the epoch's host arrays (request indices mapped to ids)."""
import json

import numpy as np
import pytest

from conftest import table_from_spec
from amod_b200 import marshal, state as astate
from amod_b200.pool import PoolBuilder
import trace_util as tu

pytestmark = pytest.mark.gpu


def _check_epoch(b, fleet, ep, gold=None):
    fleet.sync(ep)
    rows, dev = astate.DeviceFleet.rows_of(ep, fleet.max_stops), fleet.download()
    for k in rows:
        assert np.array_equal(np.asarray(rows[k]), dev[k]), k
    fleet.build(ep.mode, ep.cap, ep.T, ep.req_id[ep.search])
    got = b.fetch()
    want = astate.pool_in_request_ids(gold if gold is not None else b.build_pool(ep), ep)
    ok, why = marshal.pools_equal(got, want, rtol=0.0, check_kind=True)
    assert ok, why


def test_resident_state_follows_the_cfg1_horizon():
    z = tu.load_trace("trace_cfg1")
    b = PoolBuilder(table_from_spec(json.loads(str(z["spec"]))))
    fleet = astate.DeviceFleet(b, 100, max_stops=16)
    full = delta = 0
    n = 0
    for i in range(int(z["n_epochs"])):
        if f"e{i}_ep_veh_nid" not in z.files:
            continue
        ep = marshal.Epoch.from_npz_dict(z, f"e{i}_ep_")
        _check_epoch(b, fleet, ep, marshal.Pool.from_npz_dict(z, f"e{i}_pool_"))
        delta += fleet.last_upload_bytes
        full += sum(getattr(ep, f).nbytes for f in marshal.Epoch.FIELDS)
        n += 1
    assert n >= 250 and delta < full          # the deltas are smaller than re-uploading every epoch
    b.close()


def test_resident_state_on_the_whole_day_samples():
    z = tu.load_trace("trace_cfg2day")
    if z is None or "spec" not in z.files:
        pytest.skip("tests/golden/trace_cfg2day.npz not recorded")
    b = PoolBuilder(table_from_spec(json.loads(str(z["spec"]))))
    fleet = astate.DeviceFleet(b, 2000, max_stops=16)
    for e in z["epochs"].tolist()[::3]:
        ep = marshal.Epoch.from_npz_dict(z, f"e{e}_ep_")
        _check_epoch(b, fleet, ep)
        exp = json.loads(str(z[f"e{e}_expect"]))
        assert b.last_stats["n_entries"] == exp["n_entries"]
    b.close()
