"""Replay of the fleet-state traces recorded from the reference simulator (tests/golden/make_trace_fleet.py) through
anything that looks like a fleet (oracle/fleet_oracle.py: FleetOracle, amod_b200/fleet.py: DeviceFleetSim), comparing
every vehicle and request field bit for bit."""
from __future__ import annotations

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

A_EXACT = ("nid", "lng", "lat", "t_to_nid", "d_to_nid", "tnid", "load", "status", "state_time", "t", "d",
           "Ds", "Ts", "Ds_empty", "Ts_empty", "Dr", "Tr", "Lt", "Ld", "has_step", "step_t", "step_d", "step_u", "step_v")


_graphs = {}


class FleetTrace:
    def graph(self):
        """The road graph the trace was recorded on (tests/golden/make_trace_fleet.py: gspec)."""
        from amod_b200 import synth
        sp = self.spec
        key = (sp["n_nodes"], sp["seed"], tuple(sp.get("edge_times") or ()))
        if key not in _graphs:
            _graphs[key] = synth.make_road_graph(sp["n_nodes"], seed=sp["seed"], with_paths=True, edge_times=sp.get("edge_times"))
        return _graphs[key]

    def __init__(self, which: str):
        z = np.load(os.path.join(GOLDEN, f"trace_fleet_{which}.npz"))
        self.z = {k: z[k] for k in z.files}
        self.spec = json.loads(str(self.z["spec"]))
        self.E, self.V = self.z["A_nid"].shape
        self.T = self.z["T"]
        self.n_after = self.z["n_reqs_after"]

    def in_main(self, T) -> bool:          # utility_functions.py:40-43
        a = self.spec["warmup_min"] * 60
        return a < T <= a + self.spec["main_min"] * 60

    def new_requests(self, e):
        lo = int(self.n_after[e - 1]) if e else 0
        hi = int(self.n_after[e])
        return lo, hi

    def rows(self, prefix, e, per_vehicle=True):
        """CSR rows of epoch e: (offsets relative to the epoch, columns)."""
        off = self.z[prefix + "off"]
        if per_vehicle:
            o = off[e * self.V:(e + 1) * self.V + 1]
        else:
            o = off[e:e + 2]
        return o - o[0], int(o[0]), int(o[-1])

    def calls(self, e):
        off = self.z["C_off"]
        a, b = int(off[e]), int(off[e + 1])
        so = self.z["C_stop_off"][a:b + 1]
        cols = {k: self.z["C_stop_" + k][int(so[0]):int(so[-1])] for k in ("rid", "pod", "tnid", "ddl")}
        return self.z["C_veh"][a:b], self.z["C_phase"][a:b], (so - so[0]).astype(np.int32), cols


def _same(name, got, want, e):
    got = np.asarray(got); want = np.asarray(want)
    assert got.shape == want.shape, f"epoch {e}: {name} shape {got.shape} != {want.shape}"
    if got.dtype.kind == "f" or want.dtype.kind == "f":
        bad = np.flatnonzero(got.astype(np.float64).view(np.int64) != want.astype(np.float64).view(np.int64))
        # (+0.0 and -0.0 compare equal in the reference's own asserts; bits are still required to match here)
    else:
        bad = np.flatnonzero(got.astype(np.int64) != want.astype(np.int64))
    assert bad.size == 0, f"epoch {e}: {name} differs at {bad[:8].tolist()}: got {got[bad[:4]].tolist()} want {want[bad[:4]].tolist()}"


def check_after_advance(tr: FleetTrace, e: int, x: dict, ev: dict):
    z = tr.z
    for c in A_EXACT:
        _same("A." + c, x[c], z["A_" + c][e], e)
    o, a, b = tr.rows("A_sche_", e)
    _same("A.sche_len", x["sche_len"], np.diff(o), e)
    for k in ("rid", "pod", "tnid", "ddl"):
        _same("A.sche_" + k, x["sche_" + k], z["A_sche_" + k][a:b], e)
    _same("A.n_legs", x["n_legs"], z["A_n_legs"][e], e)
    _o, a, b = tr.rows("A_ev_", e, per_vehicle=False)
    order = np.lexsort((np.arange(ev["veh"].size), ev["veh"]))          # the reference emits events vehicle by vehicle, in leg order
    for k in ("veh", "rid", "pod", "time"):
        _same("A.ev_" + k, ev[k][order], z["A_ev_" + k][a:b], e)
    _o, a, b = tr.rows("A_rstat_", e, per_vehicle=False)
    _same("A.req_status", x["req_status"], z["A_rstat_status"][a:b], e)


def check_after_epoch(tr: FleetTrace, e: int, x: dict):
    z = tr.z
    for c in ("t", "d", "tnid", "status", "n_legs"):
        _same("D." + c, x[c], z["D_" + c][e], e)
    o, a, b = tr.rows("D_sche_", e)
    _same("D.sche_len", x["sche_len"], np.diff(o), e)
    for k in ("rid", "pod", "tnid", "ddl"):
        _same("D.sche_" + k, x["sche_" + k], z["D_sche_" + k][a:b], e)
    o, a, b = tr.rows("D_leg_", e)
    for k in ("rid", "pod", "tnid", "ddl", "t", "d", "n_steps"):
        _same("D.leg_" + k, x["leg_" + k], z["D_leg_" + k][a:b], e)
    _o, a, b = tr.rows("D_rstat_", e, per_vehicle=False)
    _same("D.req_status", x["req_status"], z["D_rstat_status"][a:b], e)


def replay_open_loop(tr: FleetTrace, fleet, epochs=None, check_every: int = 1):
    """Drive `fleet` with the recorded arrivals, status changes and build_route calls; compare after every advance and
    at the end of every epoch."""
    z = tr.z
    E = tr.E if epochs is None else min(epochs, tr.E)
    for e in range(E):
        T = int(tr.T[e])
        ev = fleet.advance(T, tr.in_main(T))
        if e % check_every == 0 or e == E - 1:
            check_after_advance(tr, e, fleet.export(), ev)
        lo, hi = tr.new_requests(e)
        if hi > lo:
            fleet.add_requests(lo, z["req_onid"][lo:hi], z["req_dnid"][lo:hi], z["req_tr"][lo:hi], z["req_ts"][lo:hi],
                               z["req_clp"][lo:hi], z["req_cld"][lo:hi])
        # what the dispatcher did to the request statuses (scheduling.py:116-117, dispatch_osp.py:77-78), as recorded
        _o, a, b = tr.rows("D_rstat_", e, per_vehicle=False)
        want = z["D_rstat_status"][a:b].astype(np.int32)
        _o, a, b = tr.rows("A_rstat_", e, per_vehicle=False)
        have = np.concatenate([z["A_rstat_status"][a:b].astype(np.int32), np.full(hi - lo, 1, dtype=np.int32)])
        for s in (1, 2):
            ch = np.flatnonzero((want != have) & (want == s))
            if ch.size:
                fleet.set_status(ch, s)
        veh, phase, off, cols = tr.calls(e)
        for ph in (0, 1):
            m = np.flatnonzero(phase == ph)
            if not m.size:
                continue
            lens = np.diff(off)[m]
            o2 = np.zeros(m.size + 1, dtype=np.int32); np.cumsum(lens, out=o2[1:])
            take = np.concatenate([np.arange(off[i], off[i + 1]) for i in m]) if lens.sum() else np.zeros(0, dtype=np.int64)
            fleet.apply(veh[m], o2, cols["rid"][take], cols["pod"][take], cols["tnid"][take], cols["ddl"][take])
        if e % check_every == 0 or e == E - 1:
            check_after_epoch(tr, e, fleet.export())
    return E
