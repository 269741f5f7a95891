"""ctypes binding of include/amod_b200.h (libamod_b200.so).

This is the whole Python↔CUDA boundary: plain pointers and sizes.  There is no CPU
fallback — if the shared library is missing or no CUDA device is present the calls
raise (SURVEY.md §8b "Errors").
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AMOD_B200_LIB") or os.path.join(_HERE, "csrc", "libamod_b200.so")   # (override: tuning variants of the same source)

OK, E_BADARG, E_CUDA, E_CAPACITY, E_LEVELS, E_NOMEM, E_LIMIT = 0, -1, -2, -3, -4, -5, -6
LOC_HOST, LOC_DEVICE, LOC_PINNED = 0, 1, 2
TT_F32, TT_F64 = 0, 1

_i32p, _f64p, _u8p, _i64p = C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int64)


class AmodEpoch(C.Structure):
    _fields_ = [("n_veh", C.c_int32), ("n_req", C.c_int32), ("n_search", C.c_int32), ("cap", C.c_int32),
                ("mode", C.c_int32), ("reserved", C.c_int32), ("T", C.c_int64),
                ("veh_nid", C.c_void_p), ("veh_t_to_nid", C.c_void_p), ("veh_load", C.c_void_p),
                ("veh_status", C.c_void_p), ("veh_sche_off", C.c_void_p), ("sche_code", C.c_void_p),
                ("sche_onboard", C.c_void_p), ("veh_reb_node", C.c_void_p), ("veh_reb_ddl", C.c_void_p),
                ("req_onid", C.c_void_p), ("req_dnid", C.c_void_p), ("req_clp", C.c_void_p),
                ("req_cld", C.c_void_p), ("req_tr", C.c_void_p), ("req_ts", C.c_void_p), ("search", C.c_void_p)]


EPOCH_PTR_FIELDS = ("veh_nid", "veh_t_to_nid", "veh_load", "veh_status", "veh_sche_off", "sche_code",
                    "sche_onboard", "veh_reb_node", "veh_reb_ddl", "req_onid", "req_dnid", "req_clp",
                    "req_cld", "req_tr", "req_ts", "search")


class AmodPool(C.Structure):
    _fields_ = [("cap_entries", C.c_int64), ("cap_trip_reqs", C.c_int64), ("cap_stops", C.c_int64),
                ("n_entries", C.c_int64), ("n_trip_reqs", C.c_int64), ("n_stops", C.c_int64),
                ("ent_veh", C.c_void_p), ("ent_kind", C.c_void_p), ("ent_trip_off", C.c_void_p),
                ("ent_sche_off", C.c_void_p), ("ent_cost", C.c_void_p), ("ent_delay", C.c_void_p),
                ("ent_nfeas", C.c_void_p), ("trip_req", C.c_void_p), ("sche_code", C.c_void_p)]


POOL_FIELDS = (("ent_veh", np.int32, "e"), ("ent_kind", np.int32, "e"), ("ent_trip_off", np.int64, "e1"),
               ("ent_sche_off", np.int64, "e1"), ("ent_cost", np.float64, "e"), ("ent_delay", np.float64, "e"),
               ("ent_nfeas", np.int32, "e"), ("trip_req", np.int32, "t"), ("sche_code", np.int32, "s"))


class AmodRoutes(C.Structure):
    _fields_ = [("cap_steps", C.c_int64), ("n_steps", C.c_int64), ("leg_off", C.c_void_p), ("step_u", C.c_void_p),
                ("step_v", C.c_void_p), ("step_t", C.c_void_p), ("step_d", C.c_void_p), ("leg_t", C.c_void_p), ("leg_d", C.c_void_p)]


class AmodStats(C.Structure):
    _fields_ = [("n_screens", C.c_int64), ("n_evals", C.c_int64), ("n_lookups", C.c_int64),
                ("n_entries", C.c_int64), ("n_sches_stored", C.c_int64), ("n_tasks", C.c_int64),
                ("n_levels", C.c_int32), ("n_kernel_launches", C.c_int32), ("ms_build", C.c_double),
                ("ms_eval", C.c_double), ("n_eval_launches", C.c_int32), ("n_attempts", C.c_int32),
                ("n_chunks", C.c_int64), ("n_items", C.c_int64), ("n_lookups_eval", C.c_int64),
                ("evals_level", C.c_int64 * 16), ("lookups_level", C.c_int64 * 16), ("ms_screen", C.c_double)]

    def as_dict(self):
        return {f: (list(getattr(self, f)) if f.endswith("_level") else getattr(self, f)) for f, _ in self._fields_}


def epoch_struct(ep, ptr_of=None) -> AmodEpoch:
    """Fill an AmodEpoch from an amod_b200.marshal.Epoch (numpy arrays -> host pointers), or
    from any object whose fields `ptr_of` can turn into addresses (e.g. torch CUDA tensors)."""
    s = AmodEpoch()
    s.n_veh, s.n_req, s.n_search, s.cap = ep.n_veh, ep.n_req, ep.n_search, ep.cap
    s.mode, s.T = ep.mode, ep.T
    if ptr_of is None:
        ptr_of = lambda a: a.ctypes.data
    for f in EPOCH_PTR_FIELDS:
        setattr(s, f, ptr_of(getattr(ep, f)))
    return s


EXPORTS = ("amod_create", "amod_destroy", "amod_last_error", "amod_version", "amod_set_stream",
           "amod_osp_build", "amod_sba_build", "amod_pool_fetch", "amod_host_register", "amod_host_unregister",
           "amod_pool_view", "amod_npo_match",
           "amod_gather_bench", "amod_gpool_create", "amod_gpool_open_peers", "amod_pool_veh_counts", "amod_gpool_scatter",
           "amod_gpool_view", "amod_gpool_fetch", "amod_set_graph", "amod_gpool_assemble", "amod_greedy_assign",
           "amod_gpool_attach", "amod_gpool_detach", "amod_routes_upload", "amod_routes_build",
           "amod_state_create", "amod_state_add_requests", "amod_state_update_vehicles", "amod_state_build", "amod_state_download",
           "amod_fleet_create", "amod_fleet_add_requests", "amod_fleet_advance", "amod_fleet_events", "amod_fleet_set_status",
           "amod_fleet_apply", "amod_fleet_apply_selected", "amod_fleet_build", "amod_fleet_array")

# amod_fleet_array selectors (include/amod_b200.h enum AMOD_FLEET_*): name -> (index, dtype, shape key)
FLEET_ARRAYS = {name: (i, dt, shape) for i, (name, dt, shape) in enumerate((
    ("nid", "i4", "V"), ("lng", "f8", "V"), ("lat", "f8", "V"), ("t_to_nid", "f8", "V"), ("d_to_nid", "f8", "V"), ("tnid", "i4", "V"),
    ("load", "i4", "V"), ("status", "i4", "V"), ("state_time", "f8", "V"), ("t", "f8", "V"), ("d", "f8", "V"), ("stats", "f8", "8V"),
    ("has_step", "u1", "V"), ("step_t", "f8", "V"), ("step_d", "f8", "V"), ("step_u", "i4", "V"), ("step_v", "i4", "V"),
    ("sche_len", "i4", "V"), ("sche_code", "i4", "VS"), ("sche_onboard", "u1", "VS"), ("reb_node", "i4", "V"), ("reb_ddl", "f8", "V"),
    ("leg_head", "i4", "V"), ("leg_n", "i4", "V"), ("leg_rid", "i4", "VL"), ("leg_pod", "i4", "VL"), ("leg_tnid", "i4", "VL"),
    ("leg_ddl", "f8", "VL"), ("leg_t", "f8", "VL"), ("leg_d", "f8", "VL"), ("leg_sb", "i4", "VL"), ("leg_se", "i4", "VL"),
    ("req_status", "i4", "R"), ("req_tp", "f8", "R"), ("req_td", "f8", "R")))}

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """Load libamod_b200.so; raises (never falls back) if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} not found: build it with `python __graft_entry__.py` "
                           f"(there is no CPU fallback for the OSP pool builder)")
    lib = C.CDLL(p)
    lib.amod_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.amod_create.restype = C.c_int
    lib.amod_destroy.argtypes = [C.c_void_p]
    lib.amod_destroy.restype = None
    lib.amod_last_error.argtypes = [C.c_void_p]
    lib.amod_last_error.restype = C.c_char_p
    lib.amod_version.argtypes = []
    lib.amod_version.restype = C.c_char_p
    lib.amod_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.amod_set_stream.restype = C.c_int
    for name in ("amod_osp_build", "amod_sba_build"):
        fn = getattr(lib, name)
        fn.argtypes = [C.c_void_p, C.POINTER(AmodEpoch), C.c_int, C.POINTER(AmodStats)]
        fn.restype = C.c_int
    lib.amod_pool_fetch.argtypes = [C.c_void_p, C.POINTER(AmodPool), C.c_int]
    lib.amod_pool_fetch.restype = C.c_int
    lib.amod_host_register.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.amod_host_register.restype = C.c_int
    lib.amod_host_unregister.argtypes = [C.c_void_p, C.c_void_p]
    lib.amod_host_unregister.restype = C.c_int
    lib.amod_pool_view.argtypes = [C.c_void_p, C.POINTER(AmodPool)]
    lib.amod_pool_view.restype = C.c_int
    lib.amod_npo_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.amod_npo_match.restype = C.c_int
    lib.amod_gather_bench.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                      C.POINTER(C.c_double)]
    lib.amod_gather_bench.restype = C.c_int
    lib.amod_gpool_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    lib.amod_gpool_create.restype = C.c_int
    lib.amod_gpool_open_peers.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.amod_gpool_open_peers.restype = C.c_int
    lib.amod_pool_veh_counts.argtypes = [C.c_void_p, C.c_void_p]
    lib.amod_pool_veh_counts.restype = C.c_int
    lib.amod_gpool_scatter.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.amod_gpool_scatter.restype = C.c_int
    lib.amod_gpool_view.argtypes = [C.c_void_p, C.POINTER(AmodPool)]
    lib.amod_gpool_view.restype = C.c_int
    lib.amod_gpool_fetch.argtypes = [C.c_void_p, C.POINTER(AmodPool)]
    lib.amod_gpool_fetch.restype = C.c_int
    lib.amod_greedy_assign.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32)]
    lib.amod_greedy_assign.restype = C.c_int
    lib.amod_gpool_assemble.argtypes = [C.c_void_p, C.c_int32]
    lib.amod_gpool_assemble.restype = C.c_int
    lib.amod_gpool_attach.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.amod_gpool_attach.restype = C.c_int
    lib.amod_gpool_detach.argtypes = [C.c_void_p]
    lib.amod_gpool_detach.restype = C.c_int
    lib.amod_routes_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.amod_routes_upload.restype = C.c_int
    lib.amod_routes_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int, C.POINTER(AmodRoutes)]
    lib.amod_routes_build.restype = C.c_int
    lib.amod_state_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.amod_state_create.restype = C.c_int
    lib.amod_state_add_requests.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 6
    lib.amod_state_add_requests.restype = C.c_int
    lib.amod_state_update_vehicles.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 10
    lib.amod_state_update_vehicles.restype = C.c_int
    lib.amod_state_build.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.POINTER(AmodStats)]
    lib.amod_state_build.restype = C.c_int
    lib.amod_state_download.argtypes = [C.c_void_p] + [C.c_void_p] * 9
    lib.amod_state_download.restype = C.c_int
    lib.amod_set_graph.argtypes = [C.c_void_p, C.c_int]
    lib.amod_set_graph.restype = C.c_int
    lib.amod_fleet_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.amod_fleet_add_requests.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 6
    lib.amod_fleet_advance.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int32)]
    lib.amod_fleet_events.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int32]
    lib.amod_fleet_set_status.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.amod_fleet_apply.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 5
    lib.amod_fleet_apply_selected.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.amod_fleet_build.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.POINTER(AmodStats), C.POINTER(C.c_int32)]
    lib.amod_fleet_array.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
    for name in EXPORTS:
        if name.startswith("amod_fleet_"):
            getattr(lib, name).restype = C.c_int
    if path is None:
        _lib = lib
    return lib
