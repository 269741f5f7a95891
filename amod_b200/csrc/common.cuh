// common.cuh — shared device structs and helpers for the OSP pool builder (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/amod_b200.h"

#define MAXS AMOD_MAX_STOPS
#define WARPS_PER_BLOCK 8
#define FULL 0xffffffffu

typedef long long i64;
typedef int32_t code_t;  // stop code: 2*req + (0 pick | 1 drop), -1 = rebalancing stop

#define MAX_LEVELS (AMOD_MAX_TRIP + 1)

// Device view of one epoch (include/amod_b200.h: struct amod_epoch).
struct DEpoch {
    int n_veh, n_req, n_search, cap, mode;
    double T;
    const int* veh_nid; const double* veh_t; const int* veh_load; const int* veh_status;
    const int* veh_sche_off; const int* sche_code; const uint8_t* sche_onboard;
    const int* reb_node; const double* reb_ddl;
    const int* onid; const int* dnid; const double* clp; const double* cld; const double* tr; const double* ts;
    const int* search;
};

// Travel-time table mean_travel_time_table[o-1][d-1] (route_functions.py:22-25), kept L2-resident.
template <typename TT>
struct Table {
    const TT* p;
    int n;
    __device__ __forceinline__ double operator()(int o, int d) const {
        return (double)__ldg(p + (size_t)(o - 1) * (size_t)n + (size_t)(d - 1));
    }
    __device__ __forceinline__ TT raw(int o, int d) const { return __ldg(p + (size_t)(o - 1) * (size_t)n + (size_t)(d - 1)); }
};

// One stop of a stored schedule: the stop code plus the node and deadline it stands for (scheduling.py:56-57), so
// that the insertion search stages a base schedule with one 16 B load per stop and no dependent look-ups.
struct __align__(16) Stop { int code; int node; double ddl; };

// One trip-size level of the per-vehicle feasible-trip table (FEASIBLE_TRIP_TABLE[veh][k-1],
// dispatch_osp.py:8) flattened over all vehicles.  Level 0 holds the basic schedules.
// Feasible schedules of trip t are rows [s0[t], s0[t]+nfeas[t]) of the level's schedule store
// (fixed `stride` stops per row), kept in the reference's LIST order (running-minimum records newest-first,
// then the rest in encounter order; scheduling.py:31-36): row s0[t] is the best schedule and the next level
// iterates the rows exactly as the reference iterates feasible_sches.  The insertion search reports its
// results in ENCOUNTER order (the reference's loop order); `sch_cost` holds the costs in that order and
// `final[s0 + e]` is the row the e-th encountered schedule of the trip is stored in.
struct Level {
    int k;
    int stride;          // stops per schedule row (>= longest schedule of this level)
    int cap_trips;
    i64 cap_sch;
    int* trip_veh;
    int* trip_req;       // [n_trips * k] ascending request indices
    i64* s0;             // first schedule row of the trip
    int* nfeas;
    double* cost;        // min cost (scheduling.py:33)
    Stop* sch;           // [n_sch * stride] list order
    double* sch_cost;    // [n_sch] cost of each feasible schedule, encounter order
    int* final;          // [n_sch] encounter position -> row
    int* veh_start;      // [n_veh+1] trips of vehicle v are [veh_start[v], veh_start[v+1])  (OSP modes)
};

struct Stats {            // device counters
    unsigned long long n_screens, n_evals, n_lookups, n_sches, n_tasks, flags, n_chunks, n_items, n_lookups_eval;
    unsigned long long lvl_evals[MAX_LEVELS + 1], lvl_lookups[MAX_LEVELS + 1];   // insertion-search work per trip level (one k_eval launch pair each)
};

// Device-resident control block: every size the level-synchronous pipeline produces lives here, so
// the host never has to synchronise between phases.  A kernel that would exceed a buffer capacity
// records the size it needs, raises `overflow` and clamps; the host regrows and reruns the epoch.
#define OVF_TASKS 1ull
#define OVF_CHUNKS 2ull
#define OVF_TRIPS 4ull
#define OVF_SCH 8ull
#define OVF_POOL 16ull
#define OVF_ROWS 32ull
#define OVF_ITEMS 64ull
#define OVF_TUPLES 128ull
struct Control {
    int n_tasks[2];
    int n_chunks[2];      // rows / chunks of the task buffer with the same index: level k+1 is prepared while
    int n_rows[2];        // level k's schedules are still being written (two branches of the epoch graph)
    int n_perm_items;     // warp work units of the basic-schedule permutation search (capacity > 5)
    int n_trips[MAX_LEVELS];
    i64 n_sch[MAX_LEVELS];
    i64 n_ent, n_ptr, n_pst;
    unsigned long long overflow;
    i64 need_rows;
    i64 need_perm_items;
    i64 need_tuples;      // feasible-insertion records per launch region (x regions) the search would have needed
    i64 need_tasks, need_chunks, need_trips[MAX_LEVELS], need_sch[MAX_LEVELS], need_ent, need_ptr, need_pst;
    int max_level_nonempty;
    int max_tlen;         // longest trip of a pool entry (= the highest score the greedy assigner will see, assign.cuh)
    int n_heavy[MAX_LEVELS];  // trips with very many feasible schedules, classified by a whole block each
    int pad2;
    Stats stats;
};

// One base schedule of one task = one "row" of insertion-search work (l+1 lanes).  32 B, one load.
struct __align__(16) RowDesc {
    int v, q, s, l;               // vehicle, inserted request, list position of the base schedule, its length
    i64 s0;                       // first schedule row of the base trip in the previous level
    int task, pad;
};

// Programmatic dependent launch (sm_90+): every kernel of the epoch pipeline is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its blocks may be scheduled while the kernel before it in the
// stream is still draining.  First statement of every such kernel: wait until the predecessor grid has completed and its
// writes are visible, then allow the successor's launch to begin.  (A no-op for launches without the attribute.)
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// -DAMOD_BOUNDS (debug build, libamod_b200_bounds.so): every store into a level, pool, task or record buffer is guarded by
// BOK(index < capacity); a violated guard skips the store and records where (file tag * 100000 + line), and the host turns
// that into an error after the epoch.  compute-sanitizer is closed on this GPU pool, so this build is what the GPU tests
// run the golden families under.  In the release build BOK(x) is `true` and costs nothing.
#ifdef AMOD_BOUNDS
__device__ unsigned long long g_bounds_hit;
__device__ __forceinline__ bool bounds_ok(bool ok, int where) {
    if (!ok) atomicCAS(&g_bounds_hit, 0ull, (unsigned long long)where);
    return ok;
}
#define BOK(cond) bounds_ok((cond), BFILE * 100000 + __LINE__)
#else
#define BOK(cond) true
#endif

__device__ __forceinline__ int pod_of(int code) { return code < 0 ? 0 : ((code & 1) ? -1 : 1); }

__device__ __forceinline__ void stop_info(const DEpoch& ep, int v, int code, int& node, double& ddl, int& pod) {
    if (code < 0) { node = ep.reb_node[v]; ddl = ep.reb_ddl[v]; pod = 0; return; }
    int r = code >> 1;
    if (code & 1) { node = ep.dnid[r]; ddl = ep.cld[r]; pod = -1; }
    else          { node = ep.onid[r]; ddl = ep.clp[r]; pod = 1; }
}

__device__ __forceinline__ Stop make_stop(const DEpoch& ep, int v, int code) {
    Stop s; int pod;
    s.code = code;
    stop_info(ep, v, code, s.node, s.ddl, pod);
    return s;
}

__device__ __forceinline__ double warp_min(double x) {
    for (int o = 16; o; o >>= 1) x = fmin(x, __shfl_xor_sync(FULL, x, o));
    return x;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long x) {
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
__device__ __forceinline__ int warp_sum(int x) {
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
// exclusive prefix over lanes
__device__ __forceinline__ int warp_excl_sum(int x, int lane) {
    int y = x;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, y, o); if (lane >= o) y += t; }
    return y - x;
}
__device__ __forceinline__ double warp_excl_min(double x, int lane) {  // +inf for lane 0
    double y = x;
    for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(FULL, y, o); if (lane >= o) y = fmin(y, t); }
    double e = __shfl_up_sync(FULL, y, 1);
    return lane == 0 ? __longlong_as_double(0x7ff0000000000000LL) : e;
}
#define DINF __longlong_as_double(0x7ff0000000000000LL)

// Kernels that run many rounds inside one (cooperative) launch: what the rounds change is read with ld.global.cg (L2), because a
// line cached in an SM's L1 in an earlier round would be stale, and the phases of a round are separated by a device-wide barrier
// (arrive with a release, spin on an acquire load; bounded, so a fault cannot hang the device; all blocks must be resident).
__device__ __forceinline__ int ldcg_i(const int* p) { return __ldcg(p); }
__device__ __forceinline__ int ldcg_b(const uint8_t* p) { return (int)__ldcg((const unsigned char*)p); }
__device__ __forceinline__ bool grid_barrier(unsigned* bar, unsigned want, int* err) {
    __shared__ int s_ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v = 0;
        long long spins = 0;
        do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < want && ++spins < (1ll << 26) && !ldcg_i(err));
        if (v < want) atomicExch(err, 1);
        s_ok = v >= want;
        __threadfence();
    }
    __syncthreads();
    return s_ok != 0;
}

