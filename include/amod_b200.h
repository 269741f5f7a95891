/* amod_b200.h — C-ABI of the B200-native Optimal-Schedule-Pool builder.
 *
 * The reference (Leot6/AMoD) is pure Python and has no FFI; the seams this library
 * stands behind are three module-level functions:
 *   B1  src/dispatcher/dispatch_osp.py:94-129   compute_candidate_veh_trip_pairs
 *   B2  src/dispatcher/dispatch_sba.py:44-74    compute_candidate_veh_req_pairs
 *   B3  src/rebalancer/rebalancing_npo.py:8-62  reposition_idle_vehicles_to_nearest_pending_orders
 * A maintainer binds these entry points with ctypes; the Python
 * mirror in amod_b200/pool.py does exactly that.  Plain pointers and sizes only.
 *
 * Conventions: every function returns 0 (AMOD_OK) or a negative error code;
 * amod_last_error() gives the text.  Node ids are 1-based as in the reference
 * (route_functions.py:23).  There is no CPU fallback: without a CUDA device
 * amod_create fails with AMOD_E_CUDA.
 */
#ifndef AMOD_B200_H
#define AMOD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMOD_OK 0
#define AMOD_E_BADARG (-1)   /* malformed input (sizes, node ids, codes out of range)            */
#define AMOD_E_CUDA (-2)     /* CUDA runtime error or no device                                   */
#define AMOD_E_CAPACITY (-3) /* caller buffer too small; required sizes are written to n_*        */
#define AMOD_E_LEVELS (-4)   /* a feasible trip of size cap+4 exists: the reference raises
                                IndexError there (dispatch_osp.py:8,226); so do we               */
#define AMOD_E_NOMEM (-5)
#define AMOD_E_LIMIT (-6)    /* input exceeds a compiled limit (stops per schedule, requests)     */

#define AMOD_LOC_HOST 0      /* pointers in the struct are host memory   */
#define AMOD_LOC_DEVICE 1    /* pointers are device memory of ctx's GPU.  PRECONDITION: the arrays are well formed (node ids in
                              * 1..n_nodes, stop codes in -1..2*n_req-1, monotone CSR offsets): host epochs are validated by
                              * the library (AMOD_E_BADARG), device epochs are not read by the host and are trusted; the
                              * resident state (amod_state_*) validates what it uploads */
#define AMOD_LOC_PINNED 2    /* host memory that is page-locked (amod_host_register, cudaHostAlloc, a pinned torch
                              * tensor): amod_pool_fetch copies into it by DMA, without the staging copy */

#define AMOD_TT_F32 0        /* travel-time table element type */
#define AMOD_TT_F64 1

/* dispatch mode = which reference seam/branch is being replaced */
#define AMOD_MODE_OSP 0      /* dispatch_osp.py, enable_reoptimization=True  */
#define AMOD_MODE_OSP_NR 1   /* dispatch_osp.py, enable_reoptimization=False */
#define AMOD_MODE_SBA 2      /* dispatch_sba.py                              */

/* pool entry kinds */
#define AMOD_KIND_BASIC 0     /* dispatch_osp.py:113,148  [veh, [], basic_sches[0], cost]     */
#define AMOD_KIND_TRIP 1      /* dispatch_osp.py:115-117  one feasible trip                   */
#define AMOD_KIND_CURRENT 2   /* dispatch_osp.py:120-125  the current working schedule        */
#define AMOD_KIND_SBA_PAIR 3  /* dispatch_sba.py:65                                           */
#define AMOD_KIND_SBA_EMPTY 4 /* dispatch_sba.py:68-69                                        */

#define AMOD_MAX_STOPS 40     /* stops in one schedule: 3*cap+8 (cap<=10)     */
#define AMOD_MAX_CAP 10
#define AMOD_MAX_TRIP 14      /* cap+4 */

typedef struct amod_ctx amod_ctx;

/* One epoch's inputs, structure-of-arrays (vehicle.py:41-70, request.py:25-41).
 * A stop code is 2*req_index + (0 pick-up | 1 drop-off); -1 is the vehicle's
 * rebalancing stop (rid=-1,pod=0; rebalancing_npo.py:37) whose node/deadline are
 * veh_reb_node/veh_reb_ddl.  A request stop's node/deadline are (onid,Clp) or
 * (dnid,Cld) of its request (scheduling.py:56-57). */
typedef struct amod_epoch {
    int32_t n_veh, n_req, n_search, cap;
    int32_t mode;                 /* AMOD_MODE_*                                            */
    int32_t reserved;
    int64_t T;                    /* system_time_sec (epoch end)                            */
    const int32_t* veh_nid;       /* [n_veh] current node                                   */
    const double*  veh_t_to_nid;  /* [n_veh]                                                */
    const int32_t* veh_load;      /* [n_veh]                                                */
    const int32_t* veh_status;    /* [n_veh] 1 idle, 2 working, 3 rebalancing (types.py:18) */
    const int32_t* veh_sche_off;  /* [n_veh+1] CSR offsets into sche_code/sche_onboard      */
    const int32_t* sche_code;     /* [n_stops] veh.sche as stop codes                       */
    const uint8_t* sche_onboard;  /* [n_stops] 1 if the stop's rid is in veh.onboard_rids   */
    const int32_t* veh_reb_node;  /* [n_veh]                                                */
    const double*  veh_reb_ddl;   /* [n_veh]                                                */
    const int32_t* req_onid;      /* [n_req]                                                */
    const int32_t* req_dnid;      /* [n_req]                                                */
    const double*  req_clp;       /* [n_req] latest pick-up  (request.py:33)                */
    const double*  req_cld;       /* [n_req] latest drop-off (request.py:34)                */
    const double*  req_tr;        /* [n_req] request time                                   */
    const double*  req_ts;        /* [n_req] shortest travel time                           */
    const int32_t* search;        /* [n_search] request indices: considered_rids (OSP) or
                                     new_received_rids (SBA), in the reference's order      */
} amod_epoch;

/* The candidate pool in the reference's order (dispatch_osp.py:107-125 vehicle-major;
 * dispatch_sba.py:56-69 request-major then one empty pair per vehicle). */
typedef struct amod_pool {
    int64_t cap_entries, cap_trip_reqs, cap_stops; /* in: capacities of the buffers below  */
    int64_t n_entries, n_trip_reqs, n_stops;       /* out                                  */
    int32_t* ent_veh;       /* [n_entries] vehicle index                                   */
    int32_t* ent_kind;      /* [n_entries] AMOD_KIND_*                                     */
    int64_t* ent_trip_off;  /* [n_entries+1] CSR into trip_req                             */
    int64_t* ent_sche_off;  /* [n_entries+1] CSR into sche_code                            */
    double*  ent_cost;      /* [n_entries] schedule completion time (scheduling.py:98,107) */
    double*  ent_delay;     /* [n_entries] compute_sche_delay (scheduling.py:126-137)      */
    int32_t* ent_nfeas;     /* [n_entries] number of feasible schedules of the trip        */
    int32_t* trip_req;      /* request indices, ascending (dispatch_osp.py:124,187)        */
    int32_t* sche_code;     /* stop codes of the entry's schedule                          */
} amod_pool;

typedef struct amod_stats {
    int64_t n_screens;      /* reachability screens (dispatch_osp.py:160, dispatch_sba.py:59)        */
    int64_t n_evals;        /* test_constraints_get_cost executions AS THE REFERENCE WOULD PERFORM
                               THEM (scheduling.py:28, dispatch_osp.py:318), i.e. honouring its
                               early exits (scheduling.py:41-44)                                     */
    int64_t n_lookups;      /* travel-time lookups the reference would perform       */
    int64_t n_entries;
    int64_t n_sches_stored; /* feasible schedules materialised over all levels                       */
    int64_t n_tasks;        /* compute_schedule calls                                                */
    int32_t n_levels;
    int32_t n_kernel_launches;
    double  ms_build;       /* device time of the whole build (CUDA events)                          */
    double  ms_eval;        /* device time inside the insertion-search kernels                       */
    int32_t n_eval_launches;
    int32_t n_attempts;     /* 1 unless a device buffer had to be regrown and the epoch rerun                */
    int64_t n_chunks;       /* warp work units of the insertion search                                       */
    int64_t n_items;        /* (base schedule, pick-up index) lanes inside them (<= 32 per chunk)            */
    int64_t n_lookups_eval; /* the part of n_lookups performed inside insertion-search evaluations           */
    int64_t evals_level[16];   /* insertion-search evaluations per trip level (index = trip size; one count + one fill
                                  launch of the search kernel each): the per-launch work behind the roofline figures */
    int64_t lookups_level[16]; /* reference lookups inside those evaluations, per trip level                         */
    double  ms_screen;         /* device time of the reachability screen (plain stream launches only, like ms_eval)  */
} amod_stats;

/* Upload the travel-time table (row = origin, col = destination; route_functions.py:22-25). */
int amod_create(amod_ctx** out, int device, int n_nodes, const void* tt_host, int tt_dtype);
void amod_destroy(amod_ctx* ctx);
const char* amod_last_error(amod_ctx* ctx);
const char* amod_version(void);

/* Replay the level loop as one captured CUDA graph (default 1).  0 = plain stream launches, which also
 * brackets every insertion-search kernel with CUDA events (amod_stats.ms_eval) for profiling. */
int amod_set_graph(amod_ctx* ctx, int enable);

/* Use an existing CUDA stream (cudaStream_t) for all work of this context; 0 = own stream. */
int amod_set_stream(amod_ctx* ctx, void* cuda_stream);

/* B1: build the OSP pool for one epoch into the context's device-resident pool.
 * ep->mode is AMOD_MODE_OSP or AMOD_MODE_OSP_NR.  `loc` says where ep's arrays live. */
int amod_osp_build(amod_ctx* ctx, const amod_epoch* ep, int loc, amod_stats* stats);

/* B2: SBA vehicle-request feasibility/cost matrix (ep->mode = AMOD_MODE_SBA). */
int amod_sba_build(amod_ctx* ctx, const amod_epoch* ep, int loc, amod_stats* stats);

/* Copy the last built pool into caller-owned buffers (host or device).  Returns
 * AMOD_E_CAPACITY with the required n_* filled if a buffer is too small. */
int amod_pool_fetch(amod_ctx* ctx, amod_pool* out, int loc);

/* Page-lock / release a caller-owned host buffer so that it can be passed as AMOD_LOC_PINNED
 * (thin wrappers over cudaHostRegister / cudaHostUnregister: the host side needs no CUDA binding). */
int amod_host_register(amod_ctx* ctx, void* p, size_t bytes);
int amod_host_unregister(amod_ctx* ctx, void* p);

/* Device pointers of the last built pool (valid until the next build on ctx). */
int amod_pool_view(amod_ctx* ctx, amod_pool* view);

/* ---- multi-GPU pool assembly over NVLink peer memory (one process per GPU) ----------
 * Vehicles are interleaved over ranks (global vehicle v lives on rank v % world as local vehicle
 * v / world).  Every rank owns one "global pool" buffer; after its local build each rank writes its
 * vehicles' entries straight into EVERY rank's global pool at their final vehicle-major positions
 * with peer stores (cudaIpc-mapped memory), so the only collectives left are the all-gather of the
 * per-vehicle entry counts (a few KB) and a barrier.  OSP modes only (SBA order is request-major). */
#define AMOD_IPC_HANDLE_BYTES 64
int amod_gpool_create(amod_ctx* ctx, int64_t cap_entries, int64_t cap_trip_reqs, int64_t cap_stops,
                      int32_t max_veh_total, int32_t world, void* ipc_handle_out /* AMOD_IPC_HANDLE_BYTES */);
int amod_gpool_open_peers(amod_ctx* ctx, int32_t world, int32_t rank, const void* handles /* world x 64 B */);
/* (entries, trip requests, stops) of every local vehicle of the last built pool -> counts_dev[n_veh*3] */
int amod_pool_veh_counts(amod_ctx* ctx, int32_t* counts_dev);
/* counts_all_dev[world][veh_per_rank_max][3] = the all-gathered counts.  Writes this rank's entries
 * into all peers' global pools.  The caller must run a barrier before any rank reads its pool. */
int amod_gpool_scatter(amod_ctx* ctx, const int32_t* counts_all_dev, int32_t veh_per_rank_max, int32_t n_veh_total);
/* The same assembly with NO host-side collective.  Each rank pushes its whole local pool (contiguous arrays,
 * 16 B peer stores) into its staging slot on every rank and raises an epoch flag there (st.release.sys); it
 * then waits for all slots in its own memory (ld.acquire.sys, bounded spin), prefix-sums the per-vehicle
 * counts in global vehicle order and merges the staged slices into its own vehicle-major pool locally.
 * Requires one GPU per rank (kernels of different ranks wait on each other).  On return this rank's pool is
 * complete; a rank starts pushing the next epoch only after every peer has finished reading the previous one. */
int amod_gpool_assemble(amod_ctx* ctx, int32_t n_veh_total);
/* Assembly INSIDE the build: after this call every amod_osp_build on ctx also pushes this rank's slice and, on the
 * merging ranks, assembles the fleet-wide pool, as part of the same captured graph and under the build's single host
 * synchronisation (amod_gpool_assemble must not be called while attached).
 *   root >= 0  gather: only rank `root` (the process that hosts the unchanged assigner) receives the slices and merges;
 *              a rank sends its slice once and is done.   root < 0: every rank receives and merges (all-gather).
 *   veh_rank / veh_local [n_veh_total] (host): placement of every global vehicle, e.g. a load-balanced sharding by the
 *              previous epoch's per-vehicle work; NULL = interleaved (v % world, v / world).  Vehicles of one rank must
 *              keep their global order.  May be called again at any time to change the placement; changing `root`
 *              requires that all ranks have met at a host-side barrier since their last build.
 * A rank whose attempt overflowed a buffer or must be searched deeper reruns its epoch and only then pushes; the
 * others wait for it (bounded: AMOD_GP_TIMEOUT_S seconds, default 30; a time-out is AMOD_E_CUDA on the waiting rank and
 * the global pool must then be rebuilt on every rank). */
int amod_gpool_attach(amod_ctx* ctx, int32_t n_veh_total, int32_t root, const int32_t* veh_rank, const int32_t* veh_local);
int amod_gpool_detach(amod_ctx* ctx);
/* device pointers + sizes of this rank's global pool (valid after the barrier) */
int amod_gpool_view(amod_ctx* ctx, amod_pool* view);
/* copy this rank's global pool into caller-owned HOST buffers (AMOD_E_CAPACITY reports the sizes needed) */
int amod_gpool_fetch(amod_ctx* ctx, amod_pool* out_host);

/* SURVEY.md §8f rank 1 (the consumer of the pool): greedy_assignment (ilp_assign.py:101-130) over the last
 * built pool, scored as the reference's scorer leaves it (score = len(trip), scheduling.py:160-161).
 * order_out[i] = pool entry at position i of the stably re-sorted list (n_entries values); sel_out = the
 * selected positions IN THAT SORTED LIST, ascending (what the reference returns); buffers host or device. */
int amod_greedy_assign(amod_ctx* ctx, int32_t* order_out, int32_t* sel_out, int32_t* n_sel, int loc, int32_t* n_rounds);

/* B3: NPO matching (rebalancing_npo.py:26-59).  idle_nid[i] = node of the i-th IDLE vehicle
 * in fleet order, pend_onid[j] = origin of the j-th PENDING request in id order.  Writes
 * min(n_idle,n_pend) matches in the reference's selection order. */
int amod_npo_match(amod_ctx* ctx, const int32_t* idle_nid, int32_t n_idle, const int32_t* pend_onid,
                   int32_t n_pend, int loc, int32_t* out_idle_idx, int32_t* out_pend_idx, double* out_dt,
                   int32_t* n_out, int32_t* n_rounds);

/* ---- batched route construction ------------------------------------------------------------------
 * route_functions.build_route_from_origin_to_dest (route_functions.py:36-70) for a batch of legs, e.g. all legs of the
 * schedules selected in one epoch (scheduling.py:110-119 -> vehicle.py:231-302 add_leg).
 * amod_routes_upload: shortest_path_table (int32 [n*n], P[o-1][d-1] = 1-based predecessor of d on the best path from o,
 * <= 0 at d == o; route_functions.py:38-43) and travel_distance_table (same element type as the travel-time table).
 * amod_routes_build: leg g goes from leg_o[g] to leg_d[g].  Its steps are rows [leg_off[g], leg_off[g+1]) of the step
 * arrays: (t, d, [u, v]) per consecutive node pair of the path, then the closing (0, 0, [tnid, tnid]) step (:59-61).
 * leg_t / leg_d are the reference's left-to-right float64 sums (:56-57).  Buffers are caller-owned (host or device);
 * AMOD_E_CAPACITY reports the n_steps needed. */
typedef struct amod_routes {
    int64_t cap_steps;       /* in: capacity of the step arrays */
    int64_t n_steps;         /* out */
    int64_t* leg_off;        /* [n_legs+1] */
    int32_t* step_u;         /* [n_steps] */
    int32_t* step_v;
    double*  step_t;
    double*  step_d;
    double*  leg_t;          /* [n_legs] */
    double*  leg_d;
} amod_routes;
int amod_routes_upload(amod_ctx* ctx, const int32_t* path_table_host, const void* dist_table_host);
int amod_routes_build(amod_ctx* ctx, const int32_t* leg_o, const int32_t* leg_d, int32_t n_legs, int loc, amod_routes* out);

/* ---- epoch state resident in HBM, updated by deltas ----------------------------------------------------
 * The request table is indexed by request id (request.py:25-41: a request's fields never change; the platform numbers them
 * 0, 1, 2, ... platform.py:262) and only grows; vehicles (vehicle.py:41-70) are rows of fixed-stride device arrays.  An epoch
 * uploads the requests that arrived and the vehicles whose row changed, then builds from the resident state; stop codes and the
 * pool's trip_req then hold request IDS (code = 2*rid + drop-off, -1 = rebalancing stop).
 * amod_state_update_vehicles: row i of the arrays describes vehicle veh_idx[i]; its schedule is stops [sche_off[i], sche_off[i+1])
 * of sche_code / sche_onboard (stop codes over request ids).  All pointers are host memory. */
int amod_state_create(amod_ctx* ctx, int32_t n_veh, int32_t max_stops_per_veh, int32_t request_capacity);
int amod_state_add_requests(amod_ctx* ctx, int32_t first_rid, int32_t n, const int32_t* onid, const int32_t* dnid, const double* clp,
                            const double* cld, const double* tr, const double* ts);
int amod_state_update_vehicles(amod_ctx* ctx, int32_t n, const int32_t* veh_idx, const int32_t* nid, const double* t_to_nid,
                               const int32_t* load, const int32_t* status, const int32_t* reb_node, const double* reb_ddl,
                               const int32_t* sche_off, const int32_t* sche_code, const uint8_t* sche_onboard);
/* build from the resident state: mode / cap / T as in amod_epoch, search = request ids in the reference's order (host). */
int amod_state_build(amod_ctx* ctx, int32_t mode, int32_t cap, int64_t T, const int32_t* search_rids, int32_t n_search, amod_stats* stats);
/* download the resident state as an amod_epoch-shaped set of host arrays (tests: parity with the host marshalling) */
int amod_state_download(amod_ctx* ctx, int32_t* nid, double* t, int32_t* load, int32_t* status, int32_t* reb_node, double* reb_ddl,
                        int32_t* sche_len, int32_t* sche_code /*[n_veh*max_stops]*/, uint8_t* sche_onboard);

/* ---- SURVEY.md 8f-3 (second half): the vehicles' state advance, the request life cycle and the apply step ON THE DEVICE ----
 * replaces, for a simulator that keeps its epoch state in HBM:
 *   Veh.move_to_time + update_service_status_after_veh_finish_a_leg / pop_leg / pop_step / cut_step   src/simulator/vehicle.py:73-226
 *   Platform.upd_vehs_and_reqs_stat_to_time (pick / drop info, walk-away scan)                          src/simulator/platform.py:213-233
 *   upd_schedule_for_vehicles_in_selected_vt_pairs + Veh.build_route / add_leg / clear_route            src/dispatcher/scheduling.py:110-119,
 *                                                                                                       vehicle.py:231-310
 *   the considered-request list of assign_orders_through_osp                                            src/dispatcher/dispatch_osp.py:33-37
 * Needs amod_routes_upload first.  The fleet owns a resident state (amod_state_*): its vehicle rows are what amod_fleet_build
 * feeds to the unchanged pool pipeline, so one epoch is
 *   amod_fleet_advance(T) -> amod_fleet_add_requests(arrivals) -> amod_fleet_build -> amod_greedy_assign ->
 *   amod_fleet_apply_selected -> amod_npo_match -> amod_fleet_apply(rebalancing stops)
 * with only the arrivals going up and the `done` events coming down.  Request status values are the reference's
 * (types.py:11-16: 1 PENDING 2 PICKING 3 ONBOARD 4 COMPLETE 5 WALKAWAY), vehicle status types.py:18-21.  All pointers are host
 * memory.  Results are bit-identical to the reference's float64 arithmetic (tests/golden/trace_fleet_*.npz). */
int amod_fleet_create(amod_ctx* ctx, int32_t n_veh, int32_t max_stops_per_veh, int32_t max_route_steps_per_veh, int32_t request_capacity,
                      const int32_t* start_nid /*[n_veh] platform.py:50-55*/, const double* node_lng, const double* node_lat /*[n_nodes]*/);
/* Req.__init__ (request.py:25-41) for the requests first_rid .. first_rid+n-1 (ids are consecutive, platform.py:262) */
int amod_fleet_add_requests(amod_ctx* ctx, int32_t first_rid, int32_t n, const int32_t* onid, const int32_t* dnid, const double* clp,
                            const double* cld, const double* tr, const double* ts);
/* platform.py:213-233 to time T; update_stats = verify_the_current_epoch_is_in_the_main_study_horizon(T).  *n_events = number of
 * (rid, pod, time) tuples the vehicles' move_to_time returned; amod_fleet_events copies them out (any order; (veh, seq) sorts
 * them into the reference's). */
int amod_fleet_advance(amod_ctx* ctx, int64_t T, int32_t update_stats, int32_t* n_events);
int amod_fleet_events(amod_ctx* ctx, int32_t* veh, int32_t* seq, int32_t* rid, int32_t* pod, double* time, int32_t cap);
int amod_fleet_set_status(amod_ctx* ctx, const int32_t* rids, int32_t n, int32_t status);
/* Veh.build_route for a batch of vehicles (each at most once): stop codes 2*rid + drop-off, -1 = rebalancing stop (reb_node[i], reb_ddl[i]) */
int amod_fleet_apply(amod_ctx* ctx, int32_t n, const int32_t* veh, const int32_t* sche_off /*[n+1]*/, const int32_t* sche_code,
                     const int32_t* reb_node /*[n]*/, const double* reb_ddl /*[n]*/);
/* the apply step on the device pool: entries == NULL applies the selection of the last amod_greedy_assign without it leaving the
 * device; mark_picking: the trips' requests become PICKING (scheduling.py:116-117); reset_considered: PICKING -> PENDING first
 * (dispatch_osp.py:77-78) */
int amod_fleet_apply_selected(amod_ctx* ctx, const int32_t* entries, int32_t n, int32_t mark_picking, int32_t reset_considered);
/* pool build from the fleet state; OSP searches the considered requests (found on the device), SBA / OSP-NR the last n_new */
int amod_fleet_build(amod_ctx* ctx, int32_t mode, int32_t cap, int64_t T, int32_t n_new, amod_stats* stats, int32_t* n_search);
/* one array of the fleet state -> host (mirror / tests); bytes must be the array's exact size */
enum { AMOD_FLEET_NID = 0 /*i32[V]*/, AMOD_FLEET_LNG, AMOD_FLEET_LAT, AMOD_FLEET_T_TO_NID, AMOD_FLEET_D_TO_NID /*f64[V]*/, AMOD_FLEET_TNID,
       AMOD_FLEET_LOAD, AMOD_FLEET_STATUS /*i32[V]*/, AMOD_FLEET_STATE_TIME, AMOD_FLEET_T, AMOD_FLEET_D /*f64[V]*/,
       AMOD_FLEET_STATS /*f64[8][V]: Ds Ts Ds_empty Ts_empty Dr Tr Lt Ld*/, AMOD_FLEET_HAS_STEP /*u8[V]*/, AMOD_FLEET_STEP_T, AMOD_FLEET_STEP_D /*f64[V]*/,
       AMOD_FLEET_STEP_U, AMOD_FLEET_STEP_V, AMOD_FLEET_SCHE_LEN /*i32[V]*/, AMOD_FLEET_SCHE_CODE /*i32[V][max_stops]*/,
       AMOD_FLEET_SCHE_ONBOARD /*u8[V][max_stops]*/, AMOD_FLEET_REB_NODE /*i32[V]*/, AMOD_FLEET_REB_DDL /*f64[V]*/, AMOD_FLEET_LEG_HEAD,
       AMOD_FLEET_LEG_N /*i32[V]*/, AMOD_FLEET_LEG_RID, AMOD_FLEET_LEG_POD, AMOD_FLEET_LEG_TNID /*i32[V][max_stops+1]*/, AMOD_FLEET_LEG_DDL,
       AMOD_FLEET_LEG_T, AMOD_FLEET_LEG_D /*f64[V][max_stops+1]*/, AMOD_FLEET_LEG_SB, AMOD_FLEET_LEG_SE /*i32[V][max_stops+1]*/,
       AMOD_FLEET_REQ_STATUS /*i32[n_req]*/, AMOD_FLEET_REQ_TP, AMOD_FLEET_REQ_TD /*f64[n_req]*/ };
int amod_fleet_array(amod_ctx* ctx, int32_t which, void* out_host, int64_t bytes);

/* Raw travel-time gather out[i] = tt[o[i]][d[i]] (route_functions.py:22-25): used by the
 * L2-gather roofline microbenchmark.  Device pointers.  Returns device ms in *ms. */
int amod_gather_bench(amod_ctx* ctx, const int32_t* o_dev, const int32_t* d_dev, float* out_dev,
                      int64_t n, int iters, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* AMOD_B200_H */
