"""The fleet oracle (oracle/fleet_oracle.c: move_to_time, walk-away scan, build_route restated) against the fleet state
recorded from the unmodified reference simulator, every epoch of both horizons, bit for bit."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

from amod_b200 import synth
import fleet_oracle
import fleet_util


@pytest.fixture(scope="module")
def graph():
    return synth.make_road_graph(seed=0, with_paths=True)


@pytest.mark.parametrize("which", ["sba", "osp"])
def test_fleet_oracle_reproduces_the_reference_horizon(which, graph):
    tr = fleet_util.FleetTrace(which)
    fleet = fleet_oracle.FleetOracle(graph, tr.z["start_nid"])
    n = fleet_util.replay_open_loop(tr, fleet)
    assert n == tr.E
    x = fleet.export()
    fleet_util._same("final Tp", x["req_tp"], tr.z["req_tp"], n)
    fleet_util._same("final Td", x["req_td"], tr.z["req_td"], n)
    fleet_util._same("final status", x["req_status"], tr.z["req_status"], n)
