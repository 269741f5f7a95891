"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference (Leot6/AMoD, Python) in this
container to pin the oracle and to generate golden vectors.

The reference has no tests and needs data files and three absent third-party modules,
so this harness (never shipped, never imported by the product path):
  1. copies /root/reference/src into a scratch directory under /tmp (the reference tree
     is read-only and its config derives data paths from its own location, config.py:11),
  2. writes synthetic pickles in the formats route_functions.py:8-17 / platform.py:60-65
     unpickle, using the reference's own `Pos` / `RawRequest` classes (types.py:90-102),
  3. injects empty stand-ins for matplotlib / gurobipy and for the (disabled, config.py:59)
     value-function module so that torch is not imported,
  4. sets COLLECT_DATA=False in the scratch copy's config (value_function.py:153 hard-codes
     a 1500-vehicle loop), and leaves everything else as shipped.
Nothing from /root/reference is copied into the repository.
"""
from __future__ import annotations

import os
import pickle
import re
import shutil
import sys
import types

REFERENCE_ROOT = os.environ.get("AMOD_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "dispatcher"))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def setup_reference(workdir: str, graph, req_onid, req_dnid, req_time_sec, *, fleet_size: int, capacity: int,
                    dispatcher: str = "OSP", rebalancer: str = "NPO", sim_start: str = "2016-05-25 18:30:00",
                    warmup_min: int = 30, main_min: int = 60, winddown_min: int = 39, day: str = "25"):
    """Create the scratch copy + data files and import the reference from it.
    Must be called at most once per process (the reference freezes config scalars at import)."""
    assert reference_available(), "reference tree not present"
    if os.path.exists(workdir):
        shutil.rmtree(workdir)
    shutil.copytree(os.path.join(REFERENCE_ROOT, "src"), os.path.join(workdir, "src"))
    cfg_path = os.path.join(workdir, "src", "simulator", "config.py")
    cfg = open(cfg_path).read()

    def sub(pattern, repl):
        nonlocal cfg
        cfg, n = re.subn(pattern, repl, cfg, count=1, flags=re.M)
        assert n == 1, pattern

    sub(r"^COLLECT_DATA = True", "COLLECT_DATA = False")
    sub(r"^FLEET_SIZE = \[\d+\]", f"FLEET_SIZE = [{fleet_size}]")
    sub(r"^VEH_CAPACITY = \[\d+\]", f"VEH_CAPACITY = [{capacity}]")
    sub(r'^DISPATCHER = "[^"]+"', f'DISPATCHER = "{dispatcher}"')
    sub(r'^REBALANCER = "[^"]+"', f'REBALANCER = "{rebalancer}"')
    sub(r'^SIMULATION_START_TIME = "[^"]+"', f'SIMULATION_START_TIME = "{sim_start}"')
    sub(r"^WARMUP_DURATION_MIN = \d+", f"WARMUP_DURATION_MIN = {warmup_min}")
    sub(r"^SIMULATION_DURATION_MIN = \d+", f"SIMULATION_DURATION_MIN = {main_min}")
    sub(r"^WINDDOWN_DURATION_MIN = \d+", f"WINDDOWN_DURATION_MIN = {winddown_min}")
    sub(r'^SIMULATION_DAYs = \["26", "25"\]', f'SIMULATION_DAYs = ["{day}"]')
    open(cfg_path, "w").write(cfg)

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image", "matplotlib.animation"):
        _stub(name)
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    _stub("gurobipy", GRB=types.SimpleNamespace(BINARY=0, MAXIMIZE=0, Attr=types.SimpleNamespace(X="X")),
          GurobiError=type("GurobiError", (Exception,), {}), Model=None)

    class ValueFunction:  # stand-in: ENABLE_VALUE_FUNCTION is False (config.py:59)
        pass

    sys.path.insert(0, workdir)
    _stub("src.value_function", __path__=[])
    _stub("src.value_function.value_function", ValueFunction=ValueFunction)

    from src.simulator import types as ref_types  # no dependencies

    map_dir = os.path.join(workdir, "datalog-gitignore", "map-data")
    taxi_dir = os.path.join(workdir, "datalog-gitignore", "taxi-data")
    os.makedirs(map_dir)
    os.makedirs(taxi_dir)

    def mkpos(nid):
        p = ref_types.Pos()
        p.node_id, p.lng, p.lat = int(nid), float(graph.lng[nid - 1]), float(graph.lat[nid - 1])
        return p

    import numpy as np
    dump = lambda obj, path: pickle.dump(obj, open(path, "wb"), protocol=4)
    dump([mkpos(i + 1) for i in range(graph.n_nodes)], os.path.join(map_dir, "nodes.pickle"))
    dump([mkpos(int(s)) for s in graph.stations], os.path.join(map_dir, "stations-101.pickle"))
    dump(graph.tt.astype(np.float64), os.path.join(map_dir, "mean-table.pickle"))
    dump(graph.dist.astype(np.float64), os.path.join(map_dir, "dist-table.pickle"))
    dump(graph.pred.astype(np.int64), os.path.join(map_dir, "path-table.pickle"))
    raws = []
    for o, d, t in zip(req_onid.tolist(), req_dnid.tolist(), req_time_sec.tolist()):
        r = ref_types.RawRequest()
        r.origin_node_id, r.destination_node_id, r.request_time_sec = int(o), int(d), int(t)
        raws.append(r)
    dump(raws, os.path.join(taxi_dir, f"manhattan-taxi-201605{day}-peak.pickle"))

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        import src.dispatcher.dispatch_osp as dispatch_osp
        import src.dispatcher.dispatch_sba as dispatch_sba
        import src.dispatcher.ilp_assign as ilp_assign
        import src.dispatcher.scheduling as scheduling
        import src.rebalancer.rebalancing_npo as rebalancing_npo
        import src.simulator.platform as platform
        import src.simulator.route_functions as route_functions
    # README.md:29 sanctions the greedy assigner when Gurobi is absent.
    dispatch_osp.ilp_assignment = lambda pairs, rids, reqs, vehs, ensure=True: ilp_assign.greedy_assignment(pairs)
    dispatch_sba.ilp_assignment = lambda pairs, rids, reqs, vehs, ensure=True: ilp_assign.greedy_assignment(pairs)
    return types.SimpleNamespace(dispatch_osp=dispatch_osp, dispatch_sba=dispatch_sba, ilp_assign=ilp_assign,
                                 scheduling=scheduling, rebalancing_npo=rebalancing_npo, platform=platform,
                                 route_functions=route_functions, types=ref_types, ValueFunction=ValueFunction,
                                 workdir=workdir, day=day)


def make_platform(ref):
    import contextlib
    import datetime
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return ref.platform.Platform(f"201605{ref.day}-peak", ref.ValueFunction(), datetime.datetime.now())
