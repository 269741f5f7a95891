"""The fleet oracle (oracle/fleet_oracle.c: move_to_time, walk-away scan, build_route restated) against the fleet state
recorded from the unmodified reference simulator, every epoch of both horizons, bit for bit."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

import fleet_oracle
import fleet_util


@pytest.mark.parametrize("which", ["sba", "osp", "ties"])
def test_fleet_oracle_reproduces_the_reference_horizon(which):
    """`ties`: a grid whose edges take exactly 10 / 15 / 30 s, so steps and legs end exactly on the 30 s epoch boundaries
    (step.t == dT: pct == 1 and a zero-length remainder; leg.t == dT: the step loop finishes the leg, vehicle.py:104-127)."""
    tr = fleet_util.FleetTrace(which)
    graph = tr.graph()
    fleet = fleet_oracle.FleetOracle(graph, tr.z["start_nid"])
    n = fleet_util.replay_open_loop(tr, fleet)
    assert n == tr.E
    x = fleet.export()
    fleet_util._same("final Tp", x["req_tp"], tr.z["req_tp"], n)
    fleet_util._same("final Td", x["req_td"], tr.z["req_td"], n)
    fleet_util._same("final status", x["req_status"], tr.z["req_status"], n)
