"""TEST INFRASTRUCTURE ONLY — ctypes wrapper around oracle/amod_oracle.c (the CPU
restatement of the reference's OSP/SBA/NPO path).  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; the
product package amod_b200 never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from amod_b200.capi import AmodPool, AmodStats, POOL_FIELDS, epoch_struct  # noqa: E402  (struct layouts only)
from amod_b200.marshal import Pool  # noqa: E402

LIB = os.path.join(_HERE, "liboracle.so")
SRC = os.path.join(_HERE, "amod_oracle.c")


def build(force: bool = False) -> str:
    hdr = os.path.join(os.path.dirname(_HERE), "include", "amod_b200.h")     # struct layouts are shared with the product ABI
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


class _Result(C.Structure):
    _fields_ = [("pool", AmodPool), ("stats", AmodStats), ("error", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_osp_build.restype = C.POINTER(_Result)
        _lib.oracle_osp_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.oracle_sba_build.restype = C.POINTER(_Result)
        _lib.oracle_sba_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.oracle_free.argtypes = [C.POINTER(_Result)]
        _lib.oracle_free.restype = None
        _lib.oracle_npo_match.restype = C.c_int
        _lib.oracle_npo_match.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_test_constraints.restype = C.c_int
        _lib.oracle_test_constraints.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                 C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lib.oracle_max_threads.restype = C.c_int
        _lib.oracle_greedy_assign.restype = C.c_int
        _lib.oracle_greedy_assign.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _tt_args(tt: np.ndarray):
    assert tt.ndim == 2 and tt.shape[0] == tt.shape[1] and tt.flags.c_contiguous
    assert tt.dtype in (np.float32, np.float64)
    return tt.ctypes.data, (1 if tt.dtype == np.float64 else 0), tt.shape[0]


def _take(res_p) -> Pool:
    r = res_p.contents
    p = r.pool
    n = {"e": p.n_entries, "e1": p.n_entries + 1, "t": p.n_trip_reqs, "s": p.n_stops}
    kw = {}
    for name, dt, key in POOL_FIELDS:
        cnt = int(n[key])
        addr = getattr(p, name)
        kw[name] = np.ctypeslib.as_array(C.cast(addr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(cnt,)).copy() \
            if cnt else np.zeros(0, dtype=dt)
    stats = r.stats.as_dict()
    err = int(r.error)
    lib().oracle_free(res_p)
    pool = Pool(stats=stats, **kw)
    pool.error = err
    return pool


def osp_build(ep, tt: np.ndarray, n_threads: int = 1, veh_lo: int = 0, veh_hi: int | None = None,
              veh_step: int = 1) -> Pool:
    """dispatch_osp.compute_candidate_veh_trip_pairs restated (cut-off disabled)."""
    s = epoch_struct(ep)
    a, d, n = _tt_args(tt)
    hi = ep.n_veh if veh_hi is None else veh_hi
    return _take(lib().oracle_osp_build(C.addressof(s), a, d, n, n_threads, veh_lo, hi, veh_step))


def sba_build(ep, tt: np.ndarray, n_threads: int = 1) -> Pool:
    s = epoch_struct(ep)
    a, d, n = _tt_args(tt)
    return _take(lib().oracle_sba_build(C.addressof(s), a, d, n, n_threads))


def npo_match(tt: np.ndarray, idle_nid: np.ndarray, pend_onid: np.ndarray):
    a, d, n = _tt_args(tt)
    idle_nid = np.ascontiguousarray(idle_nid, dtype=np.int32)
    pend_onid = np.ascontiguousarray(pend_onid, dtype=np.int32)
    m = min(idle_nid.size, pend_onid.size)
    oi, op, od = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros(m, np.float64)
    k = lib().oracle_npo_match(a, d, n, idle_nid.ctypes.data, idle_nid.size, pend_onid.ctypes.data, pend_onid.size,
                               oi.ctypes.data, op.ctypes.data, od.ctypes.data)
    return oi[:k], op[:k], od[:k]


def test_constraints(ep, tt: np.ndarray, v: int, sche_codes, new_req: int, idx_p: int, idx_d: int):
    s = epoch_struct(ep)
    a, d, n = _tt_args(tt)
    sc = np.ascontiguousarray(sche_codes, dtype=np.int32)
    cost = C.c_double(0.0)
    viol = lib().oracle_test_constraints(C.addressof(s), a, d, n, v, sc.ctypes.data, sc.size, new_req, idx_p, idx_d,
                                         C.byref(cost))
    return viol, cost.value


def greedy_assign(pool, n_veh: int, n_req: int, score=None):
    """ilp_assign.greedy_assignment restated: returns (order, selected) where order[i] is the pool entry at
    position i of the stably re-sorted list and `selected` are positions in that list."""
    score = np.ascontiguousarray(np.diff(pool.ent_trip_off) if score is None else score, dtype=np.float64)
    keep = {name: np.ascontiguousarray(getattr(pool, name), dtype=dt) for name, dt, _k in POOL_FIELDS}
    s = AmodPool()
    s.n_entries, s.n_trip_reqs, s.n_stops = pool.n_entries, keep["trip_req"].size, keep["sche_code"].size
    for name in keep:
        setattr(s, name, keep[name].ctypes.data)
    order = np.zeros(max(pool.n_entries, 1), np.int32)
    sel = np.zeros(max(pool.n_entries, 1), np.int32)
    m = lib().oracle_greedy_assign(C.addressof(s), score.ctypes.data, n_veh, n_req, order.ctypes.data, sel.ctypes.data)
    return order[:pool.n_entries], sel[:m]


def routes_build(pred: np.ndarray, tt: np.ndarray, dist: np.ndarray, leg_o, leg_d):
    """route_functions.build_route_from_origin_to_dest restated, for a batch of legs; same array layout as
    amod_routes_build: (leg_off, step_u, step_v, step_t, step_d, leg_t, leg_d)."""
    L = lib()
    if not hasattr(L, "_routes_ready"):
        L.oracle_route_build.restype = C.c_int64
        L.oracle_route_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L._routes_ready = True
    pred = np.ascontiguousarray(pred, dtype=np.int32)
    assert tt.dtype == dist.dtype and tt.flags.c_contiguous and dist.flags.c_contiguous
    n = tt.shape[0]
    dtype = 1 if tt.dtype == np.float64 else 0
    leg_o = np.asarray(leg_o, dtype=np.int32); leg_d = np.asarray(leg_d, dtype=np.int32)
    cap = 4 * n
    su, sv, st, sd = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap), np.zeros(cap)
    off = [0]
    U, V, Tt, D, LT, LD = [], [], [], [], [], []
    for o, d in zip(leg_o.tolist(), leg_d.tolist()):
        lt, ld = C.c_double(0.0), C.c_double(0.0)
        k = L.oracle_route_build(pred.ctypes.data, tt.ctypes.data, dist.ctypes.data, dtype, n, o, d, su.ctypes.data, sv.ctypes.data,
                                 st.ctypes.data, sd.ctypes.data, cap, C.byref(lt), C.byref(ld))
        assert k > 0, (o, d, k)
        U.append(su[:k].copy()); V.append(sv[:k].copy()); Tt.append(st[:k].copy()); D.append(sd[:k].copy())
        LT.append(lt.value); LD.append(ld.value)
        off.append(off[-1] + k)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return (np.asarray(off, dtype=np.int64), cat(U, np.int32), cat(V, np.int32), cat(Tt, np.float64), cat(D, np.float64),
            np.asarray(LT, dtype=np.float64), np.asarray(LD, dtype=np.float64))
