"""TEST INFRASTRUCTURE ONLY — ctypes wrapper around oracle/fleet_oracle.c (CPU restatement of the reference's vehicle
state advance, walk-away scan and build_route: vehicle.py:73-310, platform.py:213-233).  Imported only by tests/ and
__graft_entry__ (build / smoke); the product package amod_b200 never imports it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libfleet_oracle.so")
SRC = os.path.join(_HERE, "fleet_oracle.c")

VEH_COLS = ("nid", "lng", "lat", "t_to_nid", "d_to_nid", "tnid", "load", "status", "state_time", "t", "d",
            "Ds", "Ts", "Ds_empty", "Ts_empty", "Dr", "Tr", "Lt", "Ld", "has_step", "step_t", "step_d", "step_u", "step_v")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.ofleet_create.restype = C.c_void_p
        _lib.ofleet_create.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
        _lib.ofleet_destroy.argtypes = [C.c_void_p]
        _lib.ofleet_add_requests.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ofleet_set_status.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib.ofleet_advance.argtypes = [C.c_void_p, C.c_double, C.c_int]
        _lib.ofleet_events.argtypes = [C.c_void_p] * 5
        _lib.ofleet_apply.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        _lib.ofleet_error.argtypes = [C.c_void_p]
        _lib.ofleet_n_req.argtypes = [C.c_void_p]
        _lib.ofleet_export.argtypes = [C.c_void_p] * 4
        _lib.ofleet_export_sche.argtypes = [C.c_void_p] * 5
        _lib.ofleet_export_legs.argtypes = [C.c_void_p] * 8
        _lib.ofleet_export_requests.argtypes = [C.c_void_p] * 4
    return _lib


def _p(a):
    return a.ctypes.data


class FleetOracle:
    """graph: synth.RoadGraph (tt, dist, pred 1-based predecessors, lng, lat); start_nid: the vehicles' stations
    (platform.py:50-55)."""

    def __init__(self, graph, start_nid):
        L = lib()
        self.tt = np.ascontiguousarray(graph.tt)
        self.dist = np.ascontiguousarray(graph.dist, dtype=self.tt.dtype)
        self.pred = np.ascontiguousarray(graph.pred, dtype=np.int32)
        self.lng = np.ascontiguousarray(graph.lng, dtype=np.float64)
        self.lat = np.ascontiguousarray(graph.lat, dtype=np.float64)
        start = np.ascontiguousarray(start_nid, dtype=np.int32)
        self.n_veh = int(start.size)
        self.h = L.ofleet_create(self.n_veh, int(self.tt.shape[0]), 1 if self.tt.dtype == np.float64 else 0, _p(self.tt), _p(self.dist),
                                 _p(self.pred), _p(self.lng), _p(self.lat), _p(start))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ofleet_destroy(self.h)
            self.h = None

    def add_requests(self, first, onid, dnid, tr, ts, clp, cld):
        assert int(first) == lib().ofleet_n_req(self.h), "requests arrive in id order (platform.py:252-271)"
        tr, clp, ts = (np.ascontiguousarray(a, dtype=np.float64) for a in (tr, clp, ts))
        lib().ofleet_add_requests(self.h, int(tr.size), _p(tr), _p(clp), _p(ts))

    def set_status(self, rids, status: int):
        r = np.ascontiguousarray(rids, dtype=np.int32)
        lib().ofleet_set_status(self.h, _p(r), int(r.size), int(status))

    def advance(self, T, update_stats: bool):
        n = lib().ofleet_advance(self.h, float(T), int(bool(update_stats)))
        ev = dict(veh=np.zeros(n, np.int32), rid=np.zeros(n, np.int32), pod=np.zeros(n, np.int32), time=np.zeros(n, np.float64))
        lib().ofleet_events(self.h, _p(ev["veh"]), _p(ev["rid"]), _p(ev["pod"]), _p(ev["time"]))
        self._check()
        return ev

    def apply(self, veh, off, rid, pod, tnid, ddl):
        veh = np.ascontiguousarray(veh, dtype=np.int32); off = np.ascontiguousarray(off, dtype=np.int32)
        rid = np.ascontiguousarray(rid, dtype=np.int32); pod = np.ascontiguousarray(pod, dtype=np.int32)
        tnid = np.ascontiguousarray(tnid, dtype=np.int32); ddl = np.ascontiguousarray(ddl, dtype=np.float64)
        lib().ofleet_apply(self.h, int(veh.size), _p(veh), _p(off), _p(rid), _p(pod), _p(tnid), _p(ddl))
        self._check()

    def _check(self):
        e = lib().ofleet_error(self.h)
        assert e == 0, f"a reference assertion failed in the fleet oracle (mask {e})"

    def export(self) -> dict:
        V = self.n_veh
        cols = np.zeros((V, 24), dtype=np.float64)
        sl = np.zeros(V, dtype=np.int32); nl = np.zeros(V, dtype=np.int32)
        lib().ofleet_export(self.h, _p(cols), _p(sl), _p(nl))
        out = {c: cols[:, i].copy() for i, c in enumerate(VEH_COLS)}
        ns, ng = int(sl.sum()), int(nl.sum())
        s = dict(rid=np.zeros(ns, np.int32), pod=np.zeros(ns, np.int32), tnid=np.zeros(ns, np.int32), ddl=np.zeros(ns, np.float64))
        lib().ofleet_export_sche(self.h, _p(s["rid"]), _p(s["pod"]), _p(s["tnid"]), _p(s["ddl"]))
        g = dict(rid=np.zeros(ng, np.int32), pod=np.zeros(ng, np.int32), tnid=np.zeros(ng, np.int32), ddl=np.zeros(ng, np.float64),
                 t=np.zeros(ng, np.float64), d=np.zeros(ng, np.float64), n_steps=np.zeros(ng, np.int32))
        lib().ofleet_export_legs(self.h, *[_p(g[k]) for k in ("rid", "pod", "tnid", "ddl", "t", "d", "n_steps")])
        out["sche_len"] = sl; out["n_legs"] = nl
        out.update({"sche_" + k: v for k, v in s.items()})
        out.update({"leg_" + k: v for k, v in g.items()})
        nr = lib().ofleet_n_req(self.h)
        rs = np.zeros(nr, np.int32); tp = np.zeros(nr, np.float64); td = np.zeros(nr, np.float64)
        lib().ofleet_export_requests(self.h, _p(rs), _p(tp), _p(td))
        out["req_status"] = rs; out["req_tp"] = tp; out["req_td"] = td
        return out
