// amod_b200.cu — host side of the C-ABI (include/amod_b200.h): level-synchronous construction of
// the Optimal Schedule Pool on one B200.  See DESIGN.md for the data layout and kernel list.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "assign.cuh"
#include "eval.cuh"
#include "gpool.cuh"
#include "levels.cuh"
#include "npo.cuh"
#include "place.cuh"
#include "pool.cuh"
#include "routes.cuh"
#include "state.cuh"
#include "fleet.cuh"
#include "sbam.cuh"
#include "scan.cuh"

#ifndef GRID_ELEM_PER_SM
#define GRID_ELEM_PER_SM 2
#endif
#ifndef GRID_WARP_PER_SM
#define GRID_WARP_PER_SM 6      // swept 2..16 on the cfg2 snapshots (profiles/r01_grid_sweep.txt)
#endif
#ifndef GRID_EVAL_PER_SM
#define GRID_EVAL_PER_SM EVAL_BLOCKS_PER_SM
#endif

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return ctx->fail(AMOD_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,   \
                                                 cudaGetErrorString(e_));                                \
    } while (0)
#define CKR(expr)                    \
    do {                             \
        int r_ = (expr);             \
        if (r_ != AMOD_OK) return r_; \
    } while (0)

static thread_local unsigned long long g_reallocs = 0;   // bumped whenever a device buffer moves (graphs must be re-captured)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        g_reallocs++;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = std::max(bytes + bytes / 4, (size_t)4096);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() { return (T*)p; }
};

struct LevelBufs { DevBuf trip_veh, trip_req, s0, nfeas, cost, sch, sch_cost, final, veh_start; };

struct amod_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int n_nodes = 0, tt_dtype = 0;
    void* d_tt = nullptr;
    size_t tt_bytes = 0;
    int n_sm = 148;
    std::string err;
    // staging of host epochs
    DevBuf ep_dev, fetch_dev; void* ep_pinned = nullptr; size_t ep_pinned_cap = 0;
    void* fetch_pinned = nullptr; size_t fetch_pinned_cap = 0;
    // device control block + pinned host mirror
    DevBuf ctl; Control* h_ctl = nullptr;
    DevBuf d_epoch; DEpoch* h_epoch = nullptr;          // the epoch header (sizes + array pointers) lives on the device
    // one captured CUDA graph of the whole level loop per dispatch mode; replayed while buffers keep their addresses
    struct GraphSlot { cudaGraphExec_t exec = nullptr; unsigned long long gen = 0; int cap = 0, base_len = 0, max_level = 0, launches = 0;
                       size_t n_ev = 0; void* stream = nullptr; };
    GraphSlot graphs[3];
    unsigned long long buf_gen = 0;
    bool use_graph = true;
    int gen_pairs_occ[4] = {0, 0, 0, 0};   // resident blocks per SM of k_gen_pairs<1 / 4 / 8 / AMOD_MAX_TRIP>
    unsigned hash_fp_mask = 0xffffffffu;   // AMOD_HASH_FP_BITS=n keeps n bits of the slots' fingerprints (tests: forces the full-tuple comparison to reject)
    int gen_wide = 0;              // AMOD_GEN_WIDE=1: k_gen_emit ranks every row's pairs through memory, the path of rows with > 32 pairs (tests)
    int heavy_min = -1;            // AMOD_HEAVY_MIN overrides the block-per-trip classification threshold (tests)
    int level_hint[3] = {0, 0, 0};   // deepest non-empty trip level seen so far (+1 is searched before looking further)
    // capacities (grown from the sizes an epoch needed; persisted across epochs)
    i64 cap_tasks = 1 << 18, cap_chunks = 1 << 18, cap_rows = 1 << 19, cap_items = 1 << 16, cap_ent = 1 << 18, cap_ptr = 1 << 19, cap_pst = 1 << 21;
    i64 cap_tuples = 1 << 20;       // feasible-insertion records of one search launch (all block regions together)
    i64 cap_trips[MAX_LEVELS], cap_sch[MAX_LEVELS];
    // workspace
    DevBuf len0, S0, s0pos, perm_base, perm_nitems, perm_off, perm_cnt, perm_pos, flag8, pos, major_off, partial, scan_sites, task_veh[2], task_base[2], task_q[2], task_nchunk[2], task_chunk0, trip_pos,
        row_desc, task_nfeas, chunk_cnt, lane_cnt, chunk_pos, tuples, tup_cnt, tup_pos, sq_onid, sq_clp, hash, row_cnt, row_off, gen_pairs, heavy_list, gen_partial, task_first;
    LevelBufs lv[MAX_LEVELS];
    // pool
    DevBuf veh_cnt, veh_off, ent_veh, ent_kind, ent_nfeas, ent_tlen, ent_slen, ent_src, ent_trip, ent_cost, ent_delay,
        trip_off, sche_off, trip_req, sche_code;
    i64 pool_entries = 0, pool_trip_reqs = 0, pool_stops = 0;
    bool pool_valid = false;
    int n_attempts = 0;
    // multi-GPU global pool (peer-mapped)
    char* gp_base = nullptr; GPoolLayout gp_layout; GPoolPeers gp_peers; bool gp_open = false; int gp_max_veh = 0;
    GSlotLayout gp_slot; i64 gp_o_stage = 0; int gp_world = 0;
    DevBuf gp_off, gp_tot, gp_place, gp_glob; void* gp_peer_ptr[GPOOL_MAX_WORLD] = {nullptr};
    DevBuf sm_pos, sm_epos, sm_egv, sm_dest, sm_rs, sm_base, sm_slen;      // SBA merge workspace (sbam.cuh)
    GpState* gp_h_state = nullptr;           // pinned mirror of the device-resident assembly state
    bool gp_attached = false;                // every build also assembles (inside the build's single synchronisation)
    int gp_root = -1;                        // >= 0: only this rank receives and merges the slices (gather); < 0: every rank does
    int gp_n_total = 0;
    // greedy assignment
    DevBuf ga_pos, ga_rank, ga_order, ga_state, ga_bv, ga_br, ga_tv, ga_tr, ga_sel, ga_cnt;
    // npo
    DevBuf npo_in, npo_vm, npo_rm, npo_rb, npo_vb, npo_rdt, npo_mdt, npo_cnt, npo_grp, npo_rg, npo_D, npo_best, npo_tie, npo_scr, npo_out, ga_sync;
    // routes (8f-2)
    int* d_pred = nullptr; void* d_dist = nullptr;
    DevBuf rt_in, rt_len, rt_off, rt_u, rt_v, rt_t, rt_d, rt_lt, rt_ld, rt_err;
    // resident epoch state (8f-3)
    FleetState fs; bool fs_valid = false; int fs_req_cap = 0, fs_req_n = 0, fs_max_len = 0;
    DevBuf fs_nid, fs_t, fs_load, fs_status, fs_rnode, fs_rddl, fs_len, fs_code, fs_onb, fs_off64, fs_off32, fs_ccode, fs_conb, fs_search, fs_stage, fs_err;
    DevBuf rq_onid, rq_dnid, rq_clp, rq_cld, rq_tr, rq_ts;
    void* fs_pinned = nullptr; size_t fs_pinned_cap = 0;
    // fleet simulation state (8f-3: advance + apply on the device, fleet.cuh)
    FleetSim fl; bool fl_valid = false; std::vector<DevBuf> fl_bufs; DevBuf fl_call, fl_pos, fl_stage, fl_rids;
    int fl_n_events = 0; int ga_n_sel = -1;     // ga_n_sel: selected pairs of the last greedy assignment (on the device in ga_sel / ga_order)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> ev_pool;
    cudaStream_t side = nullptr;             // second branch of the epoch graph (next level's candidates)
    std::vector<cudaEvent_t> sync_pool;      // fork / join edges between the two branches
    int n_launches = 0;
    unsigned hash_slots = 0;                 // slots per trip lookup table (two tables in `hash`)
    size_t sq_n = 0;                         // searched requests of the epoch (sizes the search cache even when the fleet is empty)
    cudaAccessPolicyWindow l2_window; bool l2_window_valid = false, l2_on = false;
    int base_len_hint[3] = {0, 0, 0};        // longest current schedule seen so far per mode (SBA / OSP_NR): keeps graph + strides stable

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf;
        return code;
    }
};

static thread_local std::string g_err;

template <typename F>
static int dispatch_tt(amod_ctx* ctx, F&& f) {
    if (ctx->tt_dtype == AMOD_TT_F64) return f(Table<double>{(const double*)ctx->d_tt, ctx->n_nodes});
    return f(Table<float>{(const float*)ctx->d_tt, ctx->n_nodes});
}

static inline int nblk(i64 n, int per) { return (int)((n + per - 1) / per); }

// Every kernel of the epoch pipeline is launched with programmatic stream serialization (PDL): its blocks may be
// scheduled while the previous kernel of the stream drains; the kernel itself waits for that grid's completion before
// touching memory (pdl_enter(), common.cuh).  AMOD_NO_PDL=1 launches them plainly (A/B: profiles/r02_pdl_ab.txt).
static bool g_pdl = getenv("AMOD_NO_PDL") == nullptr;
template <typename... KArgs, typename... Args>
static inline void launch_ks(cudaStream_t st, int grid, int block, size_t smem, void (*kern)(KArgs...), Args&&... args) {
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.stream = st; cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline void launch_k(cudaStream_t st, int grid, int block, void (*kern)(KArgs...), Args&&... args) {
    launch_ks(st, grid, block, 0, kern, static_cast<Args&&>(args)...);
}

// exclusive scan (length and result sizes live on the device; no synchronisation)
template <typename In, typename Fin>
static void scan_dev(amod_ctx* ctx, In in, i64* out, Fin fin) {
    const int g = ctx->n_sm;
    i64* partial = ctx->partial.as<i64>();
    launch_k(ctx->stream, g, SCAN_THREADS, k_scan_a<In>, in, partial);
    launch_k(ctx->stream, g, SCAN_THREADS, k_scan_b<In, Fin>, in, partial, out, fin);
    ctx->n_launches += 2;
}

// phase (b) only: the producer kernel left the block sums in `partial` (zeroed before the epoch; see scan_seg_len)
template <typename InA, typename FinA, typename InB, typename FinB>
static void scan_fin2(amod_ctx* ctx, InA a, i64* outa, FinA fa, InB b, i64* outb, FinB fb, i64* partial) {
    const int g = (ctx->n_sm / 2) * 2;
    launch_k(ctx->stream, g, SCAN_THREADS, k_scan2_b<InA, FinA, InB, FinB>, a, fa, outa, b, fb, outb, partial);
    ctx->n_launches += 1;
}

template <typename In, typename Fin>
static void scan_fin(amod_ctx* ctx, In in, i64* out, Fin fin, i64* partial, cudaStream_t st = nullptr) {
    launch_k(st ? st : ctx->stream, ctx->n_sm, SCAN_THREADS, k_scan_b<In, Fin>, in, partial, out, fin);
    ctx->n_launches += 1;
}

// Keep the travel-time table resident in L2 (126 MB on B200; 67 MB as fp32): persisting
// access-policy window on the streams the kernels run on AND on every kernel node of the captured
// epoch graph (graph kernel nodes carry their own launch attributes).  AMOD_NO_L2_POLICY=1 switches
// it off (A/B evidence: profiles/r02_l2_policy_ab.txt).
static void apply_l2_policy(amod_ctx* ctx, cudaStream_t stream) {
    if (!ctx->l2_window_valid) {
        ctx->l2_on = false;
        if (getenv("AMOD_NO_L2_POLICY")) return;
        int max_win = 0, max_persist = 0;
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
        if (max_win <= 0 || max_persist <= 0 || !ctx->d_tt) return;
        size_t win = std::min(ctx->tt_bytes, (size_t)max_win);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min((size_t)max_persist, win));
        memset(&ctx->l2_window, 0, sizeof ctx->l2_window);
        ctx->l2_window.base_ptr = ctx->d_tt;
        ctx->l2_window.num_bytes = win;
        ctx->l2_window.hitRatio = (float)std::min(1.0, (double)max_persist / (double)win);
        ctx->l2_window.hitProp = cudaAccessPropertyPersisting;
        ctx->l2_window.missProp = cudaAccessPropertyStreaming;
        ctx->l2_window_valid = true;
        ctx->l2_on = true;
    }
    if (!ctx->l2_on) return;
    cudaStreamAttrValue av; memset(&av, 0, sizeof av);
    av.accessPolicyWindow = ctx->l2_window;
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av);
    cudaGetLastError();
}

// the same window on every kernel node of a captured graph (before instantiation)
static void apply_l2_policy_to_graph(amod_ctx* ctx, cudaGraph_t graph) {
    if (!ctx->l2_on) return;
    size_t n = 0;
    if (cudaGraphGetNodes(graph, nullptr, &n) != cudaSuccess || n == 0) { cudaGetLastError(); return; }
    std::vector<cudaGraphNode_t> nodes(n);
    if (cudaGraphGetNodes(graph, nodes.data(), &n) != cudaSuccess) { cudaGetLastError(); return; }
    cudaKernelNodeAttrValue av; memset(&av, 0, sizeof av);
    av.accessPolicyWindow = ctx->l2_window;
    for (size_t i = 0; i < n; i++) {
        cudaGraphNodeType ty;
        if (cudaGraphNodeGetType(nodes[i], &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
        cudaGraphKernelNodeSetAttribute(nodes[i], cudaKernelNodeAttributeAccessPolicyWindow, &av);
    }
    cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
#ifdef AMOD_BOUNDS
extern "C" const char* amod_version(void) { return "amod_b200 0.2 (sm_100a, AMOD_BOUNDS debug build)"; }
#else
extern "C" const char* amod_version(void) { return "amod_b200 0.2 (sm_100a)"; }
#endif

extern "C" const char* amod_last_error(amod_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C" int amod_create(amod_ctx** out, int device, int n_nodes, const void* tt_host, int tt_dtype) {
    if (!out || !tt_host || n_nodes <= 0 || (tt_dtype != AMOD_TT_F32 && tt_dtype != AMOD_TT_F64)) {
        g_err = "amod_create: bad argument";
        return AMOD_E_BADARG;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device >= ndev) {
        g_err = std::string("amod_create: no usable CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        return AMOD_E_CUDA;
    }
    amod_ctx* ctx = new amod_ctx();
    ctx->device = device; ctx->n_nodes = n_nodes; ctx->tt_dtype = tt_dtype;
    auto bail = [&](const char* what, cudaError_t ce) {
        g_err = std::string("amod_create: ") + what + ": " + cudaGetErrorString(ce);
        delete ctx;
        return AMOD_E_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    ctx->own_stream = true;
    cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device);
    size_t bytes = (size_t)n_nodes * n_nodes * (tt_dtype == AMOD_TT_F64 ? 8 : 4);
    if ((e = cudaMalloc(&ctx->d_tt, bytes)) != cudaSuccess) return bail("cudaMalloc(table)", e);
    if ((e = cudaMemcpy(ctx->d_tt, tt_host, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return bail("upload table", e);
    ctx->tt_bytes = bytes;
    for (int k = 0; k < MAX_LEVELS; k++) { ctx->cap_trips[k] = 1 << 16; ctx->cap_sch[k] = 1 << 18; }
    if ((e = cudaMallocHost((void**)&ctx->h_ctl, sizeof(Control))) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void**)&ctx->h_epoch, sizeof(DEpoch))) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = ctx->d_epoch.ensure(sizeof(DEpoch))) != cudaSuccess) return bail("cudaMalloc", e);
    ctx->use_graph = getenv("AMOD_NO_GRAPH") == nullptr;
    if (const char* hm = getenv("AMOD_HEAVY_MIN")) ctx->heavy_min = atoi(hm);
    if (const char* gi = getenv("AMOD_GEN_WIDE")) ctx->gen_wide = atoi(gi) != 0;
    if (const char* fb = getenv("AMOD_HASH_FP_BITS")) { const int n = std::max(0, std::min(32, atoi(fb))); ctx->hash_fp_mask = n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }
    apply_l2_policy(ctx, ctx->stream);
    apply_l2_policy(ctx, ctx->side);         // k_basic<fill>, candidate generation: the side branch looks travel times up too
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1);
    *out = ctx;
    return AMOD_OK;
}

extern "C" void amod_destroy(amod_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    DevBuf* all[] = {&ctx->ep_dev, &ctx->fetch_dev, &ctx->ctl, &ctx->len0, &ctx->S0, &ctx->s0pos, &ctx->perm_base, &ctx->perm_nitems, &ctx->perm_off, &ctx->perm_cnt, &ctx->perm_pos, &ctx->flag8, &ctx->pos, &ctx->major_off,
                     &ctx->partial, &ctx->scan_sites,
                     &ctx->task_veh[0], &ctx->task_veh[1], &ctx->task_base[0], &ctx->task_base[1], &ctx->task_q[0], &ctx->task_q[1],
                     &ctx->task_nchunk[0], &ctx->task_nchunk[1], &ctx->task_chunk0, &ctx->trip_pos, &ctx->row_desc, &ctx->task_nfeas, &ctx->chunk_cnt,
                     &ctx->lane_cnt, &ctx->chunk_pos, &ctx->tuples, &ctx->tup_cnt, &ctx->tup_pos, &ctx->sq_onid, &ctx->sq_clp, &ctx->hash, &ctx->row_cnt, &ctx->row_off, &ctx->gen_pairs, &ctx->heavy_list, &ctx->gen_partial, &ctx->task_first,
                     &ctx->veh_cnt, &ctx->veh_off, &ctx->ent_veh, &ctx->ent_kind, &ctx->ent_nfeas, &ctx->ent_tlen, &ctx->ent_slen,
                     &ctx->ent_src, &ctx->ent_trip, &ctx->ent_cost, &ctx->ent_delay, &ctx->trip_off, &ctx->sche_off, &ctx->trip_req,
                     &ctx->sche_code, &ctx->ga_pos, &ctx->ga_rank, &ctx->ga_order, &ctx->ga_state, &ctx->ga_bv, &ctx->ga_br, &ctx->ga_tv, &ctx->ga_tr,
                     &ctx->ga_sel, &ctx->ga_cnt, &ctx->npo_in, &ctx->npo_vm, &ctx->npo_rm, &ctx->npo_rb, &ctx->npo_vb, &ctx->npo_rdt,
                     &ctx->npo_mdt, &ctx->npo_cnt, &ctx->npo_grp, &ctx->npo_rg, &ctx->npo_D, &ctx->npo_best, &ctx->npo_tie, &ctx->npo_scr, &ctx->npo_out, &ctx->ga_sync};
    for (DevBuf* b : all) b->release();
    for (auto& l : ctx->lv) { l.trip_veh.release(); l.trip_req.release(); l.s0.release(); l.nfeas.release(); l.cost.release();
                              l.sch.release(); l.sch_cost.release(); l.final.release(); l.veh_start.release(); }
    if (ctx->h_ctl) cudaFreeHost(ctx->h_ctl);
    if (ctx->h_epoch) cudaFreeHost(ctx->h_epoch);
    if (ctx->fetch_pinned) cudaFreeHost(ctx->fetch_pinned);
    for (int r = 0; r < GPOOL_MAX_WORLD; r++) if (ctx->gp_peer_ptr[r]) cudaIpcCloseMemHandle(ctx->gp_peer_ptr[r]);
    if (ctx->gp_base) cudaFree(ctx->gp_base);
    ctx->gp_off.release(); ctx->gp_tot.release(); ctx->gp_place.release(); ctx->gp_glob.release();
    { DevBuf* mb[] = {&ctx->sm_pos, &ctx->sm_epos, &ctx->sm_egv, &ctx->sm_dest, &ctx->sm_rs, &ctx->sm_base, &ctx->sm_slen}; for (DevBuf* b : mb) b->release(); }
    if (ctx->gp_h_state) cudaFreeHost(ctx->gp_h_state);
    ctx->d_epoch.release();
    for (auto& g : ctx->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (ctx->ep_pinned) cudaFreeHost(ctx->ep_pinned);
    if (ctx->d_tt) cudaFree(ctx->d_tt);
    { DevBuf* sb[] = {&ctx->fs_nid, &ctx->fs_t, &ctx->fs_load, &ctx->fs_status, &ctx->fs_rnode, &ctx->fs_rddl, &ctx->fs_len, &ctx->fs_code, &ctx->fs_onb,
                      &ctx->fs_off64, &ctx->fs_off32, &ctx->fs_ccode, &ctx->fs_conb, &ctx->fs_search, &ctx->fs_stage, &ctx->fs_err, &ctx->rq_onid, &ctx->rq_dnid,
                      &ctx->rq_clp, &ctx->rq_cld, &ctx->rq_tr, &ctx->rq_ts};
      for (DevBuf* b : sb) b->release(); }
    if (ctx->fs_pinned) cudaFreeHost(ctx->fs_pinned);
    if (ctx->d_pred) cudaFree(ctx->d_pred);
    if (ctx->d_dist) cudaFree(ctx->d_dist);
    { DevBuf* rb[] = {&ctx->rt_in, &ctx->rt_len, &ctx->rt_off, &ctx->rt_u, &ctx->rt_v, &ctx->rt_t, &ctx->rt_d, &ctx->rt_lt, &ctx->rt_ld, &ctx->rt_err};
      for (DevBuf* b : rb) b->release(); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    for (auto e : ctx->sync_pool) cudaEventDestroy(e);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int amod_set_graph(amod_ctx* ctx, int enable) {
    if (!ctx) return AMOD_E_BADARG;
    ctx->use_graph = enable != 0;
    return AMOD_OK;
}

extern "C" int amod_set_stream(amod_ctx* ctx, void* cuda_stream) {
    if (!ctx) return AMOD_E_BADARG;
    if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; apply_l2_policy(ctx, ctx->stream); }
    else { CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; apply_l2_policy(ctx, ctx->stream); }
    return AMOD_OK;
}

// ------------------------------------------------------------------------------------------------
// epoch upload / validation
// ------------------------------------------------------------------------------------------------
static int validate_host_epoch(amod_ctx* ctx, const amod_epoch* ep) {
    const int N = ctx->n_nodes;
    for (int v = 0; v < ep->n_veh; v++) {
        if (ep->veh_nid[v] < 1 || ep->veh_nid[v] > N) return ctx->fail(AMOD_E_BADARG, "vehicle %d: node id %d out of range", v, ep->veh_nid[v]);
        if (ep->veh_sche_off[v + 1] < ep->veh_sche_off[v]) return ctx->fail(AMOD_E_BADARG, "veh_sche_off not monotone at %d", v);
        if (ep->veh_sche_off[v + 1] - ep->veh_sche_off[v] > AMOD_MAX_STOPS - 2)
            return ctx->fail(AMOD_E_LIMIT, "vehicle %d: schedule longer than %d stops", v, AMOD_MAX_STOPS - 2);
        if (ep->veh_load[v] < 0) return ctx->fail(AMOD_E_BADARG, "vehicle %d: negative load", v);
    }
    for (int r = 0; r < ep->n_req; r++)
        if (ep->req_onid[r] < 1 || ep->req_onid[r] > N || ep->req_dnid[r] < 1 || ep->req_dnid[r] > N)
            return ctx->fail(AMOD_E_BADARG, "request %d: node id out of range", r);
    const int ns = ep->n_veh ? ep->veh_sche_off[ep->n_veh] : 0;
    for (int v = 0, k = 0; v < ep->n_veh; v++)
        for (k = ep->veh_sche_off[v]; k < ep->veh_sche_off[v + 1]; k++) {
            int c = ep->sche_code[k];
            if (c < -1 || (c >> 1) >= ep->n_req) return ctx->fail(AMOD_E_BADARG, "stop code %d out of range", c);
            if (c == -1 && (ep->veh_reb_node[v] < 1 || ep->veh_reb_node[v] > N))
                return ctx->fail(AMOD_E_BADARG, "vehicle %d: rebalancing node out of range", v);
        }
    (void)ns;
    for (int s = 0; s < ep->n_search; s++)
        if (ep->search[s] < 0 || ep->search[s] >= ep->n_req) return ctx->fail(AMOD_E_BADARG, "search[%d] out of range", s);
    return AMOD_OK;
}

static int upload_epoch(amod_ctx* ctx, const amod_epoch* ep, int loc, DEpoch* d) {
    if (ep->n_veh < 0 || ep->n_req < 0 || ep->n_search < 0) return ctx->fail(AMOD_E_BADARG, "negative size");
    if (ep->cap < 1 || ep->cap > AMOD_MAX_CAP) return ctx->fail(AMOD_E_LIMIT, "capacity %d outside 1..%d", ep->cap, AMOD_MAX_CAP);
    if (ep->n_req > (1 << 29)) return ctx->fail(AMOD_E_LIMIT, "too many live requests in one epoch");
    d->n_veh = ep->n_veh; d->n_req = ep->n_req; d->n_search = ep->n_search; d->cap = ep->cap; d->mode = ep->mode;
    d->T = (double)ep->T;
    if (loc == AMOD_LOC_DEVICE) {
        d->veh_nid = ep->veh_nid; d->veh_t = ep->veh_t_to_nid; d->veh_load = ep->veh_load; d->veh_status = ep->veh_status;
        d->veh_sche_off = ep->veh_sche_off; d->sche_code = ep->sche_code; d->sche_onboard = ep->sche_onboard;
        d->reb_node = ep->veh_reb_node; d->reb_ddl = ep->veh_reb_ddl; d->onid = ep->req_onid; d->dnid = ep->req_dnid;
        d->clp = ep->req_clp; d->cld = ep->req_cld; d->tr = ep->req_tr; d->ts = ep->req_ts; d->search = ep->search;
        *ctx->h_epoch = *d;
        CK(cudaMemcpyAsync(ctx->d_epoch.p, ctx->h_epoch, sizeof(DEpoch), cudaMemcpyHostToDevice, ctx->stream));
        return AMOD_OK;
    }
    CKR(validate_host_epoch(ctx, ep));
    const size_t V = ep->n_veh, R = ep->n_req, Q = ep->n_search;
    const size_t NS = V ? (size_t)ep->veh_sche_off[V] : 0;
    struct Part { const void* src; size_t bytes; const void** dst; };
    Part parts[] = {
        {ep->veh_t_to_nid, V * 8, (const void**)&d->veh_t}, {ep->veh_reb_ddl, V * 8, (const void**)&d->reb_ddl},
        {ep->req_clp, R * 8, (const void**)&d->clp}, {ep->req_cld, R * 8, (const void**)&d->cld},
        {ep->req_tr, R * 8, (const void**)&d->tr}, {ep->req_ts, R * 8, (const void**)&d->ts},
        {ep->veh_nid, V * 4, (const void**)&d->veh_nid}, {ep->veh_load, V * 4, (const void**)&d->veh_load},
        {ep->veh_status, V * 4, (const void**)&d->veh_status}, {ep->veh_sche_off, (V + 1) * 4, (const void**)&d->veh_sche_off},
        {ep->veh_reb_node, V * 4, (const void**)&d->reb_node}, {ep->req_onid, R * 4, (const void**)&d->onid},
        {ep->req_dnid, R * 4, (const void**)&d->dnid}, {ep->search, Q * 4, (const void**)&d->search},
        {ep->sche_code, NS * 4, (const void**)&d->sche_code}, {ep->sche_onboard, NS, (const void**)&d->sche_onboard},
    };
    size_t total = 0;
    for (auto& p : parts) total += (p.bytes + 15) & ~(size_t)15;
    total = std::max(total, (size_t)16);
    if (total > ctx->ep_pinned_cap) {
        if (ctx->ep_pinned) cudaFreeHost(ctx->ep_pinned);
        ctx->ep_pinned = nullptr; ctx->ep_pinned_cap = 0;
        CK(cudaMallocHost(&ctx->ep_pinned, total * 2));
        ctx->ep_pinned_cap = total * 2;
    }
    CK(ctx->ep_dev.ensure(total));
    size_t off = 0;
    for (auto& p : parts) {
        if (p.bytes) memcpy((char*)ctx->ep_pinned + off, p.src, p.bytes);
        *p.dst = (char*)ctx->ep_dev.p + off;
        off += (p.bytes + 15) & ~(size_t)15;
    }
    CK(cudaMemcpyAsync(ctx->ep_dev.p, ctx->ep_pinned, total, cudaMemcpyHostToDevice, ctx->stream));
    *ctx->h_epoch = *d;
    CK(cudaMemcpyAsync(ctx->d_epoch.p, ctx->h_epoch, sizeof(DEpoch), cudaMemcpyHostToDevice, ctx->stream));
    return AMOD_OK;
}

// ------------------------------------------------------------------------------------------------
// the build
// ------------------------------------------------------------------------------------------------
// Block sums left by producer kernels, one zeroed array per scan of the epoch (levels, basic schedules, screen, pool)
#define SCAN_SITES (MAX_LEVELS + 4)
static size_t scan_sites_bytes(const amod_ctx* ctx) { return (size_t)SCAN_SITES * 8 * (size_t)(ctx->n_sm + 2); }
static unsigned long long* scan_site(amod_ctx* ctx, int i) { return ctx->scan_sites.as<unsigned long long>() + (size_t)i * (size_t)(ctx->n_sm + 2); }

static cudaEvent_t get_sync_event(amod_ctx* ctx, size_t i) {
    while (ctx->sync_pool.size() <= i) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); ctx->sync_pool.push_back(e); }
    return ctx->sync_pool[i];
}

#define EV_SCREEN 100     // ev_pool slots: [0, 64) insertion-search launches, 64.. peer-assembly profiling, 100/101 the screen
static cudaEvent_t get_event(amod_ctx* ctx, size_t i) {
    while (ctx->ev_pool.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); ctx->ev_pool.push_back(e); }
    return ctx->ev_pool[i];
}

static int ensure_buffers(amod_ctx* ctx, int V, i64 n_screen, int max_level, const int* stride, LevelSet* ls, PoolBufs* pb) {
    const size_t T = (size_t)ctx->cap_tasks, C = (size_t)ctx->cap_chunks;
    CK(ctx->ctl.ensure(sizeof(Control)));
    CK(ctx->len0.ensure(4 * (size_t)(V + 1))); CK(ctx->S0.ensure(4 * (size_t)(V + 1))); CK(ctx->s0pos.ensure(8 * (size_t)(V + 2)));
    CK(ctx->perm_base.ensure(4 * (size_t)(V + 1) * AMOD_MAX_CAP)); CK(ctx->perm_nitems.ensure(4 * (size_t)(V + 1)));
    CK(ctx->perm_off.ensure(8 * (size_t)(V + 2))); CK(ctx->perm_cnt.ensure(4 * (size_t)ctx->cap_items)); CK(ctx->perm_pos.ensure(8 * ((size_t)ctx->cap_items + 2)));
    {   // screen: survivor bitmask (one word per 32 minors of each major) + per-major packed counts / offsets
        const size_t Qs = V > 0 ? (size_t)(n_screen / V) : 0;
        const size_t words = std::max((size_t)V * ((Qs + 31) / 32), Qs * (((size_t)V + 31) / 32)) + 1;
        const size_t majors = std::max((size_t)V, Qs) + 2;
        CK(ctx->flag8.ensure(4 * words)); CK(ctx->pos.ensure(8 * majors)); CK(ctx->major_off.ensure(8 * majors));
    }
    CK(ctx->partial.ensure(2 * 8 * (size_t)(ctx->n_sm + 1)));
    CK(ctx->scan_sites.ensure(scan_sites_bytes(ctx)));
    {   // searched requests' origin / latest pick-up by search position (levels.cuh SearchCache)
        const size_t Qs = V > 0 ? (size_t)(n_screen / V) : 0;
        CK(ctx->sq_onid.ensure(4 * (std::max(Qs, ctx->sq_n) + 1))); CK(ctx->sq_clp.ensure(8 * (std::max(Qs, ctx->sq_n) + 1)));
    }
    for (int w = 0; w < 2; w++) {
        CK(ctx->task_veh[w].ensure(4 * T)); CK(ctx->task_base[w].ensure(4 * T)); CK(ctx->task_q[w].ensure(4 * T));
        CK(ctx->task_nchunk[w].ensure(4 * T));
    }
    CK(ctx->task_chunk0.ensure(8 * (T + 2))); CK(ctx->trip_pos.ensure(8 * (T + 2)));
    CK(ctx->row_desc.ensure(2 * sizeof(RowDesc) * (size_t)ctx->cap_rows)); CK(ctx->task_nfeas.ensure(4 * T)); CK(ctx->chunk_cnt.ensure(4 * C)); CK(ctx->lane_cnt.ensure(64 * C)); CK(ctx->chunk_pos.ensure(8 * (C + 2)));
    CK(ctx->gen_pairs.ensure(sizeof(int2) * T));      // (j, q) pairs of a level's candidates = the next level's tasks
    {   // one record region per block of the search launch
        const size_t n_regions = (size_t)ctx->n_sm * GRID_EVAL_PER_SM;
        const size_t region = ((size_t)ctx->cap_tuples + n_regions - 1) / n_regions;
        CK(ctx->tuples.ensure(sizeof(Tuple) * region * n_regions)); CK(ctx->tup_pos.ensure(4 * region * n_regions)); CK(ctx->tup_cnt.ensure(4 * n_regions));
    }
    size_t max_trips = 1;
    memset(ls, 0, sizeof *ls);
    for (int k = 0; k <= max_level; k++) {
        LevelBufs& b = ctx->lv[k];
        const size_t nt = (size_t)std::max(ctx->cap_trips[k], (i64)(k == 0 ? V + 1 : 1));
        ctx->cap_trips[k] = (i64)nt;
        const size_t nsch = (size_t)std::max(ctx->cap_sch[k], (i64)1);
        max_trips = std::max(max_trips, nt);
        CK(b.trip_veh.ensure(4 * nt)); CK(b.trip_req.ensure(4 * nt * (size_t)std::max(k, 1))); CK(b.s0.ensure(8 * nt));
        CK(b.nfeas.ensure(4 * nt)); CK(b.cost.ensure(8 * nt));
        CK(b.sch.ensure(sizeof(Stop) * nsch * (size_t)stride[k])); CK(b.sch_cost.ensure(8 * nsch)); CK(b.final.ensure(4 * nsch));
        CK(b.veh_start.ensure(4 * ((size_t)V + 2)));
        Level& lv = ls->lv[k];
        lv.k = k; lv.stride = stride[k]; lv.cap_trips = (int)nt; lv.cap_sch = (i64)nsch;
        lv.trip_veh = b.trip_veh.as<int>(); lv.trip_req = b.trip_req.as<int>(); lv.s0 = b.s0.as<i64>(); lv.nfeas = b.nfeas.as<int>();
        lv.cost = b.cost.as<double>(); lv.sch = b.sch.as<Stop>(); lv.sch_cost = b.sch_cost.as<double>(); lv.final = b.final.as<int>();
        lv.veh_start = b.veh_start.as<int>();
    }
    ls->n = max_level + 1;
    CK(ctx->row_cnt.ensure(4 * max_trips)); CK(ctx->row_off.ensure(8 * (max_trips + 2))); CK(ctx->heavy_list.ensure(4 * max_trips));
    unsigned hsz = 64; while (hsz < 2u * (unsigned)max_trips + 2u) hsz <<= 1;
    CK(ctx->hash.ensure(2 * 8 * (size_t)hsz));      // two tables: level k+1's is cleared while level k's is read
    ctx->hash_slots = hsz;
    CK(ctx->gen_partial.ensure(8 * ((size_t)ctx->n_sm * GRID_WARP_PER_SM + 1)));      // segment sums + the pair-buffer cursor
    CK(ctx->task_first.ensure(8 * (max_trips + 2)));
    CK(ctx->veh_cnt.ensure(4 * (size_t)(V + 1))); CK(ctx->veh_off.ensure(8 * (size_t)(V + 2)));
    const size_t E = (size_t)ctx->cap_ent, PT = (size_t)ctx->cap_ptr, PS = (size_t)ctx->cap_pst;
    CK(ctx->ent_veh.ensure(4 * E)); CK(ctx->ent_kind.ensure(4 * E)); CK(ctx->ent_nfeas.ensure(4 * E)); CK(ctx->ent_tlen.ensure(4 * E));
    CK(ctx->ent_slen.ensure(4 * E)); CK(ctx->ent_src.ensure(4 * E)); CK(ctx->ent_trip.ensure(4 * E)); CK(ctx->ent_cost.ensure(8 * E));
    CK(ctx->ent_delay.ensure(8 * E)); CK(ctx->trip_off.ensure(8 * (E + 2))); CK(ctx->sche_off.ensure(8 * (E + 2)));
    CK(ctx->trip_req.ensure(4 * PT)); CK(ctx->sche_code.ensure(4 * PS));
    pb->cap_ent = (i64)E; pb->cap_ptr = (i64)PT; pb->cap_pst = (i64)PS;
    pb->ent_veh = ctx->ent_veh.as<int>(); pb->ent_kind = ctx->ent_kind.as<int>(); pb->ent_nfeas = ctx->ent_nfeas.as<int>();
    pb->ent_tlen = ctx->ent_tlen.as<int>(); pb->ent_slen = ctx->ent_slen.as<int>(); pb->ent_src = ctx->ent_src.as<int>();
    pb->ent_trip = ctx->ent_trip.as<int>(); pb->ent_cost = ctx->ent_cost.as<double>(); pb->ent_delay = ctx->ent_delay.as<double>();
    pb->trip_off = ctx->trip_off.as<i64>(); pb->sche_off = ctx->sche_off.as<i64>(); pb->trip_req = ctx->trip_req.as<int>();
    pb->sche_code = ctx->sche_code.as<int>();
    return AMOD_OK;
}

static void enqueue_gpool(amod_ctx* ctx, int run_levels, int max_level, bool sba);
static int check_gpool_state(amod_ctx* ctx);

// One attempt: enqueue the whole level-synchronous pipeline on the stream without any host
// synchronisation; every size is produced and consumed on the device (Control block).
template <typename TT>
static int enqueue_build(amod_ctx* ctx, int mode, int cap, Table<TT> tt, int max_level, const LevelSet& ls, const PoolBufs& pb,
                         size_t* n_ev_out) {
    cudaStream_t st = ctx->stream;
    const DEpoch* d = ctx->d_epoch.as<DEpoch>();      // sizes and array pointers are read on the device
    const int sba = mode == AMOD_MODE_SBA;
    Control* ctl = ctx->ctl.as<Control>();
    Stats* dstats = &ctl->stats;
    const int G = ctx->n_sm * GRID_ELEM_PER_SM;   // elementwise kernels: grid-stride on a fixed grid
    const int GE = ctx->n_sm * GRID_WARP_PER_SM;   // warp-per-item kernels
    const int GEV = ctx->n_sm * GRID_EVAL_PER_SM;  // the insertion search: exactly the resident blocks, one wave
    size_t n_ev = 0;
    int* len0 = ctx->len0.as<int>();
    const SearchCache sc{ctx->sq_onid.as<int>(), ctx->sq_clp.as<double>()};

    cudaStream_t sd = ctx->side;
    const int half = ctx->n_sm / 2;            // blocks per scan of the pair scan (k_scan2_b)
    size_t n_sync = 0;
    auto edge = [&](cudaStream_t from, cudaStream_t to) -> cudaError_t {
        cudaEvent_t e = get_sync_event(ctx, n_sync++);
        cudaError_t ce = cudaEventRecord(e, from);
        return ce != cudaSuccess ? ce : cudaStreamWaitEvent(to, e, 0);
    };
    // ---- level 0: basic schedules -----------------------------------------------------------------
    if (cap <= 5) {      // at most 120 orderings per vehicle: one warp per vehicle is enough
        launch_k(st, GE, WARPS_PER_BLOCK * 32, k_basic<TT, 0>, d, tt, ls.lv[0], len0, ctx->S0.as<int>(), nullptr, ctl, dstats,
                                                            scan_site(ctx, MAX_LEVELS), ctx->n_sm, sc);
        // the level-0 store is written beside the reachability screen, which only needs the counts (S0)
        CK(edge(st, sd));
        scan_fin(ctx, InVehInt{ctx->S0.as<int>(), d}, ctx->s0pos.as<i64>(),
                 FinCountI64{&ctl->n_sch[0], ls.lv[0].cap_sch, &ctl->need_sch[0], &ctl->overflow, OVF_SCH}, (i64*)scan_site(ctx, MAX_LEVELS), sd);
        launch_k(sd, GE, WARPS_PER_BLOCK * 32, k_basic<TT, 1>, d, tt, ls.lv[0], len0, ctx->S0.as<int>(), ctx->s0pos.as<i64>(), ctl, dstats,
                                                            nullptr, ctx->n_sm, sc);
        ctx->n_launches += 2;
    } else {             // up to cap! orderings: 32 permutation ranks per warp work item, spread over the GPU
        launch_k(st, G, 256, k_perm_prep, d, ls.lv[0], len0, ctx->perm_base.as<int>(), ctx->perm_nitems.as<int>(), dstats, sc);
        scan_dev(ctx, InVehInt{ctx->perm_nitems.as<int>(), d}, ctx->perm_off.as<i64>(), FinPermItems{ctl, ctx->cap_items});
        launch_k(st, GE, WARPS_PER_BLOCK * 32, k_perm_items<TT, 0>, d, tt, ls.lv[0], ctl, len0, ctx->perm_base.as<int>(), ctx->perm_off.as<i64>(),
                                                                ctx->perm_cnt.as<int>(), nullptr, dstats);
        scan_dev(ctx, InArray<int>{ctx->perm_cnt.as<int>(), &ctl->n_perm_items, 0}, ctx->perm_pos.as<i64>(),
                 FinBasicRows{ctl, d, ls.lv[0].cap_sch});
        launch_k(st, GE, WARPS_PER_BLOCK * 32, k_perm_items<TT, 1>, d, tt, ls.lv[0], ctl, len0, ctx->perm_base.as<int>(), ctx->perm_off.as<i64>(),
                                                                ctx->perm_cnt.as<int>(), ctx->perm_pos.as<i64>(), dstats);
        launch_k(st, G, 256, k_perm_rows, d, ls.lv[0], ctl, len0, ctx->S0.as<int>(), ctx->perm_off.as<i64>(), ctx->perm_pos.as<i64>());
        CK(edge(st, sd));     // (the side branch joins the graph here; it has nothing to do before the levels)
        ctx->n_launches += 4;
    }
    // ---- level-1 tasks: reachability screen ---------------------------------------------------------
    {
        const int G1 = std::max(1, 32 / (ls.lv[0].stride + 1));
        CK(cudaEventRecord(get_event(ctx, EV_SCREEN), st));
        launch_k(st, ctx->n_sm * 16, SCREEN_WARPS * 32, k_screen_count<TT>, d, tt, ctx->S0.as<int>(), ctx->flag8.as<unsigned>(), ctx->pos.as<i64>(),
                                                                         scan_site(ctx, MAX_LEVELS + 1), ctx->n_sm, sc);
        scan_fin(ctx, InMajorPacked{ctx->pos.as<i64>(), d}, ctx->major_off.as<i64>(),
                 FinTasksRows{ctl, 0, ctx->cap_tasks, ctx->cap_rows, ctx->cap_chunks, G1}, (i64*)scan_site(ctx, MAX_LEVELS + 1));
        launch_k(st, GE, WARPS_PER_BLOCK * 32, k_screen_fill, d, ctl, ctx->S0.as<int>(), ctx->flag8.as<unsigned>(), ctx->major_off.as<i64>(),
                                                          ctx->task_veh[0].as<int>(), ctx->task_base[0].as<int>(), ctx->task_q[0].as<int>(),
                                                          ctx->task_nchunk[0].as<int>(), ctx->task_chunk0.as<i64>());
        CK(cudaEventRecord(get_event(ctx, EV_SCREEN + 1), st));
        ctx->n_launches += 2;
    }
    // ---- levels 1..max_level ----------------------------------------------------------------------------
    // Two branches per level once the schedule offsets are known (fork after the scans):
    //   main: k_place_cost -> k_classify -> k_place_rows     (orders the level's schedules, writes them)
    //   side: k_compact_trips -> k_gen_pairs -> k_gen_emit
    //         (names and indexes the level's trips, prepares level k+1's tasks and rows; touches no schedule)
    // They meet again before k_eval<count> of level k+1.  Buffers both branches touch at once are doubled
    // (task lists, row descriptors, Control::n_rows / n_chunks, scan partials).
    auto hash_size_of = [&](int k) -> unsigned {   // trips of level k are looked up by the candidates of level k+1
        if (k < 2 || k >= max_level) return 0u;
        unsigned sz = 64; while (sz < 2u * (unsigned)ls.lv[k].cap_trips + 2u) sz <<= 1;
        return sz;
    };
    auto hash_of = [&](int k) -> TripHash {        // two tables by level parity: k+1's is cleared while k's is read
        TripHash th; th.slots = ctx->hash.as<unsigned long long>() + (size_t)(k & 1) * ctx->hash_slots;
        th.mask = hash_size_of(k) ? hash_size_of(k) - 1 : 0;
        th.fp_mask = ctx->hash_fp_mask;
        return th;
    };
    RowDesc* rows_buf[2] = {ctx->row_desc.as<RowDesc>(), ctx->row_desc.as<RowDesc>() + ctx->cap_rows};
    i64* row0 = ctx->task_chunk0.as<i64>();
    i64* cpos = ctx->chunk_pos.as<i64>();
    int* tnf = ctx->task_nfeas.as<int>();
    CK(edge(sd, st));     // level 0 is stored
    launch_k(st, GE, 256, k_expand_rows, ctl, 0, ls.lv[0], len0, ctx->task_veh[0].as<int>(), ctx->task_base[0].as<int>(), ctx->task_q[0].as<int>(),
                                      ctx->task_nchunk[0].as<int>(), row0, rows_buf[0], tnf);
    ctx->n_launches++;
    const int n_regions = GEV;
    const int region_cap = (int)(((size_t)ctx->cap_tuples + n_regions - 1) / n_regions);
    Tuple* tup = ctx->tuples.as<Tuple>();
    int* tup_cnt = ctx->tup_cnt.as<int>();
    unsigned* tup_pos = ctx->tup_pos.as<unsigned>();
    for (int k = 1; k <= max_level; k++) {
        const int cur = (k - 1) & 1, nxt = k & 1;
        int* tv = ctx->task_veh[cur].as<int>(); int* tb = ctx->task_base[cur].as<int>(); int* tq = ctx->task_q[cur].as<int>();
        const Level& prev = ls.lv[k - 1];
        const Level& lv = ls.lv[k];
        const int Gk = std::max(1, 32 / (prev.stride + 1));      // rows per chunk: Gk * (longest schedule + 1) <= 32 lanes
        RowDesc* rows = rows_buf[cur];
        CK(cudaEventRecord(get_event(ctx, n_ev++), st));
        launch_k(st, GEV, EVAL_WARPS * 32, k_eval<TT>, d, tt, prev, lv, ctl, cur, Gk, rows, ctx->chunk_cnt.as<int>(),
                                                       ctx->lane_cnt.as<uint8_t>(), tnf, dstats, scan_site(ctx, k), half,
                                                       tup, tup_cnt, region_cap);
        CK(cudaEventRecord(get_event(ctx, n_ev++), st));
        scan_fin2(ctx, InArray<int>{ctx->chunk_cnt.as<int>(), &ctl->n_chunks[cur], 0}, cpos,
                  FinCountI64{&ctl->n_sch[k], lv.cap_sch, &ctl->need_sch[k], &ctl->overflow, OVF_SCH},
                  InTaskFeasible{ctl, cur, tnf}, ctx->trip_pos.as<i64>(),
                  FinTrips{&ctl->n_trips[k], (i64)lv.cap_trips, &ctl->need_trips[k], &ctl->overflow}, (i64*)scan_site(ctx, k));
        CK(edge(st, sd));                                         // fork
        const TripHash th = hash_of(k);
        auto kct = KM_IDX(k) == 0 ? k_compact_trips<1> : KM_IDX(k) == 1 ? k_compact_trips<4> : KM_IDX(k) == 2 ? k_compact_trips<8> : k_compact_trips<AMOD_MAX_TRIP>;
        launch_k(sd, G, 256, kct, ctl, cur, tv, tb, tq, tnf, ctx->trip_pos.as<i64>(), prev, lv, th, dstats, d, sba,
                                          ctx->major_off.as<i64>(), ctx->task_first.as<i64>(), ctx->gen_partial.as<unsigned long long>(), GE + 1);
        cudaEvent_t named = get_sync_event(ctx, n_sync++);        // the level's trips have their rows (s0, nfeas)
        CK(cudaEventRecord(named, sd));
        // main: costs to their encounter positions -> list order per trip -> schedules straight into their final rows
        launch_k(st, n_regions, PLACE_THREADS, k_place_cost, ctl, tup, tup_cnt, region_cap, ctx->lane_cnt.as<uint8_t>(), cpos, lv, tup_pos);
        CK(cudaStreamWaitEvent(st, named, 0));
        const int heavy_min = ctx->heavy_min >= 0 ? ctx->heavy_min : (cap > 5 ? 2048 : 0);   // large capacities: trips with 10^3..10^5 schedules
        launch_k(st, GE, WARPS_PER_BLOCK * 32, k_classify, ctl, lv, heavy_min, ctx->heavy_list.as<int>());
        if (heavy_min) { launch_k(st, ctx->n_sm * 2, HEAVY_THREADS, k_classify_heavy, ctl, lv, ctx->heavy_list.as<int>()); ctx->n_launches++; }
        launch_k(st, n_regions, PLACE_THREADS, k_place_rows, d, ctl, tup, tup_cnt, region_cap, tup_pos, rows, Gk, prev, lv);
        ctx->n_launches += 5;       // k_eval, k_compact_trips, k_place_cost, k_classify, k_place_rows (the scans count themselves)
        if (k < max_level) {
            // ---- candidates of size k+1 (dispatch_osp.py:183-216), their tasks and rows ----------------------
            const TripHash thn = hash_of(k + 1);
            unsigned long long* gpart = ctx->gen_partial.as<unsigned long long>();
            // one wave: the grid is what fits (the rows are dealt cyclically; a second wave would only start when the first has drained)
            auto kgp = KM_IDX(k) == 0 ? k_gen_pairs<1> : KM_IDX(k) == 1 ? k_gen_pairs<4> : KM_IDX(k) == 2 ? k_gen_pairs<8> : k_gen_pairs<AMOD_MAX_TRIP>;
            int& occ = ctx->gen_pairs_occ[KM_IDX(k)];
            if (occ == 0) { if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kgp, WARPS_PER_BLOCK * 32, 0) != cudaSuccess || occ < 1) occ = 1; }
            launch_k(sd, ctx->n_sm * std::min(occ, GRID_WARP_PER_SM), WARPS_PER_BLOCK * 32, kgp, ctl, ls.lv[1], lv, th, ctx->row_cnt.as<int>(), ctx->row_off.as<int>(),
                                                            ctx->gen_pairs.as<int2>(), ctx->cap_tasks, gpart, GE, gpart + GE, thn.slots, thn.mask ? thn.mask + 1 : 0u);
            GenOut go;
            go.task_veh = ctx->task_veh[nxt].as<int>(); go.task_base = ctx->task_base[nxt].as<int>(); go.task_q = ctx->task_q[nxt].as<int>();
            go.task_row0 = row0; go.rows = rows_buf[nxt]; go.task_nfeas = tnf; go.task_first = ctx->task_first.as<i64>();
            go.cap_tasks = ctx->cap_tasks; go.cap_rows = ctx->cap_rows; go.cap_chunks = ctx->cap_chunks;
            go.G_next = std::max(1, 32 / (lv.stride + 1));
            launch_k(sd, GE, WARPS_PER_BLOCK * 32, k_gen_emit, ctl, nxt, lv, len0, ctx->row_cnt.as<int>(), ctx->row_off.as<int>(),
                                                           ctx->gen_pairs.as<int2>(), gpart, go, ctx->gen_wide);
            ctx->n_launches += 2;
        }
        CK(edge(sd, st));                                         // join
    }
    // ---- pool assembly --------------------------------------------------------------------------------
    {
        LevelSet lsp = ls;
        lsp.n = max_level + 1;          // only the levels searched this epoch hold current data
        launch_k(st, G, 256, k_pool_meta, d, lsp, len0, ctx->veh_off.as<i64>(), ctl, pb, scan_site(ctx, MAX_LEVELS + 2), half);
        scan_fin2(ctx, InEntLen{pb.ent_tlen, ctl}, pb.trip_off, FinCountI64{&ctl->n_ptr, pb.cap_ptr, &ctl->need_ptr, &ctl->overflow, OVF_POOL},
                  InEntLen{pb.ent_slen, ctl}, pb.sche_off, FinCountI64{&ctl->n_pst, pb.cap_pst, &ctl->need_pst, &ctl->overflow, OVF_POOL},
                  (i64*)scan_site(ctx, MAX_LEVELS + 2));
        launch_k(st, G * 2, 128, k_pool_payload<TT>, d, tt, lsp, ctl, pb);
        ctx->n_launches += 2;
    }
    // ---- multi-GPU: this rank's slice travels (and, on the merging ranks, the fleet-wide pool is assembled) in the same graph
    if (ctx->gp_attached) enqueue_gpool(ctx, max_level, sba ? max_level : cap + 4, sba != 0);
    *n_ev_out = n_ev;
    return AMOD_OK;
}

template <typename TT>
static int build_impl(amod_ctx* ctx, const amod_epoch* ep, int loc, amod_stats* out_stats, Table<TT> tt) {
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    ctx->pool_valid = false; ctx->ga_n_sel = -1;
    DEpoch d;
    CK(cudaEventRecord(ctx->ev0, st));
    CKR(upload_epoch(ctx, ep, loc, &d));
    const int V = d.n_veh, mode = d.mode;
    const int max_level = mode == AMOD_MODE_SBA ? 1 : d.cap + 4;
    if (max_level > AMOD_MAX_TRIP) return ctx->fail(AMOD_E_LIMIT, "trip size limit");
    // row stride of each level's schedule store: longest possible base schedule + 2 stops per level
    int base_len = ep->reserved > 0 ? ep->reserved : 0;
    if (mode == AMOD_MODE_OSP) base_len = std::max(d.cap, 1);
    else if (base_len == 0) {
        if (loc == AMOD_LOC_HOST) { for (int v = 0; v < V; v++) base_len = std::max(base_len, ep->veh_sche_off[v + 1] - ep->veh_sche_off[v]); }
        else if (V > 0) {
            std::vector<int> off((size_t)V + 1);
            CK(cudaMemcpyAsync(off.data(), ep->veh_sche_off, 4 * ((size_t)V + 1), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            for (int v = 0; v < V; v++) base_len = std::max(base_len, off[v + 1] - off[v]);
        }
    }
    if (mode != AMOD_MODE_OSP) {     // monotone per mode: the captured graph and the level strides stay put once the fleet's longest schedule was seen
        base_len = std::max(base_len, ctx->base_len_hint[mode]);
        ctx->base_len_hint[mode] = base_len;
    }
    int stride[MAX_LEVELS];
    for (int k = 0; k < MAX_LEVELS; k++) stride[k] = std::max(1, std::min(MAXS, base_len + 2 * k));

    if (ctx->gp_attached && mode == AMOD_MODE_SBA && (ctx->gp_root < 0 || ctx->gp_root == ctx->gp_peers.rank)) {
        // workspace of the request-major merge on this (merging) rank
        const unsigned long long r0 = g_reallocs;
        const size_t W = (size_t)ctx->gp_peers.world, CE = (size_t)ctx->gp_slot.cap_ent, Q = (size_t)d.n_search;
        CK(ctx->sm_pos.ensure(4 * ((size_t)d.n_req + 1))); CK(ctx->sm_epos.ensure(4 * W * CE)); CK(ctx->sm_egv.ensure(4 * W * CE));
        CK(ctx->sm_dest.ensure(4 * W * CE)); CK(ctx->sm_rs.ensure(4 * W * (Q + 2))); CK(ctx->sm_base.ensure(8 * (Q + 2)));
        CK(ctx->sm_slen.ensure(4 * ((size_t)ctx->gp_layout.cap_ent + 1))); CK(ctx->partial.ensure(2 * 8 * (size_t)(ctx->n_sm + 1)));
        if (g_reallocs != r0) ctx->buf_gen++;
    }
    LevelSet ls; PoolBufs pb;
    size_t n_ev = 0;
    Control* h = ctx->h_ctl;
    int attempt = 0;
    // Trip levels above the deepest feasible one are empty: search one level past what the previous epoch
    // reached and look further only if that level still produced trips (then the epoch is rerun deeper).
    int run_levels = mode == AMOD_MODE_SBA ? 1 : std::min(max_level, std::max(2, ctx->level_hint[mode] + 1));
    for (;; attempt++) {
        if (attempt > 48) return ctx->fail(AMOD_E_NOMEM, "pool build did not converge after %d regrow attempts", attempt);
        const unsigned long long r0 = g_reallocs;
        ctx->sq_n = (size_t)d.n_search;
        CKR(ensure_buffers(ctx, V, (i64)V * d.n_search, max_level, stride, &ls, &pb));
        if (g_reallocs != r0) ctx->buf_gen++;
        memset(h, 0, sizeof(Control));
        h->n_trips[0] = V;
        CK(cudaMemcpyAsync(ctx->ctl.p, h, sizeof(Control), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(ctx->scan_sites.p, 0, scan_sites_bytes(ctx), st));     // block sums the producer kernels add into
        if (ctx->use_graph) {
            // the whole level loop is one CUDA graph (two branches, ~60-90 kernel nodes) whose sizes all live on the device,
            // so a captured graph stays valid until a buffer moves or the dispatch mode / capacity changes
            amod_ctx::GraphSlot& gs = ctx->graphs[mode];
            if (!gs.exec || gs.gen != ctx->buf_gen || gs.cap != d.cap || gs.base_len != base_len || gs.max_level != run_levels ||
                gs.stream != (void*)st) {
                if (gs.exec) { cudaGraphExecDestroy(gs.exec); gs.exec = nullptr; }
                cudaGraph_t graph = nullptr;
                CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
                ctx->n_launches = 0;
                int rc = enqueue_build(ctx, mode, d.cap, tt, run_levels, ls, pb, &n_ev);
                cudaError_t ce = cudaStreamEndCapture(st, &graph);
                if (rc != AMOD_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
                CK(ce);
                apply_l2_policy_to_graph(ctx, graph);
                ce = cudaGraphInstantiate(&gs.exec, graph, 0);
                cudaGraphDestroy(graph);
                CK(ce);
                gs.gen = ctx->buf_gen; gs.cap = d.cap; gs.base_len = base_len; gs.max_level = run_levels; gs.launches = ctx->n_launches;
                gs.n_ev = n_ev; gs.stream = (void*)st;
            }
            n_ev = gs.n_ev;
            ctx->n_launches = gs.launches;
            CK(cudaGraphLaunch(gs.exec, st));
        } else {
            ctx->n_launches = 0;
            CKR(enqueue_build(ctx, mode, d.cap, tt, run_levels, ls, pb, &n_ev));
        }
        CK(cudaEventRecord(ctx->ev1, st));
        CK(cudaMemcpyAsync(h, ctx->ctl.p, sizeof(Control), cudaMemcpyDeviceToHost, st));
        const bool with_gp = ctx->gp_attached;
        if (with_gp) CK(cudaMemcpyAsync(ctx->gp_h_state, ctx->gp_tot.p, sizeof(GpState), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
#ifdef AMOD_BOUNDS
        {   // debug build: a guarded store was out of range (common.cuh BOK): file tag * 100000 + line
            unsigned long long hit = 0;
            CK(cudaMemcpyFromSymbol(&hit, g_bounds_hit, sizeof hit));
            if (hit) {
                unsigned long long zero = 0;
                cudaMemcpyToSymbol(g_bounds_hit, &zero, sizeof zero);
                return ctx->fail(AMOD_E_CUDA, "AMOD_BOUNDS: guarded store out of range at site %llu (file tag * 100000 + line)", hit);
            }
        }
#endif
        if (!h->overflow) {
            if (run_levels < max_level && h->n_trips[run_levels] > 0) { run_levels = std::min(max_level, run_levels + 2); continue; }
            break;
        }
        // regrow exactly what overflowed (with headroom) and rerun the epoch
        auto grow = [](i64& cap, i64 need) { if (need > cap) cap = need + need / 2 + 1024; };
        grow(ctx->cap_tasks, h->need_tasks); grow(ctx->cap_chunks, h->need_chunks); grow(ctx->cap_rows, h->need_rows); grow(ctx->cap_items, h->need_perm_items);
        for (int k = 0; k <= max_level; k++) { grow(ctx->cap_trips[k], h->need_trips[k]); grow(ctx->cap_sch[k], h->need_sch[k]); }
        grow(ctx->cap_ent, h->need_ent); grow(ctx->cap_ptr, h->need_ptr); grow(ctx->cap_pst, h->need_pst);
        grow(ctx->cap_tuples, h->need_tuples);
        for (int k = 0; k <= max_level; k++) if (ctx->cap_sch[k] > 0x7ffffff0) return ctx->fail(AMOD_E_LIMIT, "more than 2^31 feasible schedules in one level");
        if (ctx->cap_tasks > 0x7ffffff0 || ctx->cap_chunks > 0x7ffffff0 || ctx->cap_rows > 0x7ffffff0)
            return ctx->fail(AMOD_E_LIMIT, "more than 2^31 candidates or base-schedule rows in one level");
    }
    ctx->n_attempts = attempt + 1;
    {
        int deepest = 0;
        for (int k = 1; k <= run_levels; k++) if (h->n_trips[k] > 0) deepest = k;
        ctx->level_hint[mode] = std::max(ctx->level_hint[mode], deepest);   // monotone: a captured graph is reused, never thrashed
    }
    if (h->stats.flags & 1ull) return ctx->fail(AMOD_E_LIMIT, "a vehicle schedule exceeds the compiled stop/capacity limits");
    const i64 P = h->n_ent, NT = h->n_ptr, NS = h->n_pst;
    ctx->pool_entries = P; ctx->pool_trip_reqs = NT; ctx->pool_stops = NS;
    ctx->pool_valid = true;
    int levels_done = 0;
    i64 total_tasks = 0;
    for (int k = 1; k <= run_levels; k++) if (h->n_trips[k] > 0) levels_done = k;
    if (out_stats) {
        memset(out_stats, 0, sizeof *out_stats);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        out_stats->ms_build = ms;
        double me = 0.0;
        for (size_t i = 0; i + 1 < n_ev; i += 2) {
            float x = 0.f;
            if (cudaEventElapsedTime(&x, ctx->ev_pool[i], ctx->ev_pool[i + 1]) == cudaSuccess) me += x; else cudaGetLastError();
        }
        out_stats->ms_eval = me;
        if (!ctx->use_graph && ctx->ev_pool.size() > EV_SCREEN + 1) {
            float x = 0.f;
            if (cudaEventElapsedTime(&x, ctx->ev_pool[EV_SCREEN], ctx->ev_pool[EV_SCREEN + 1]) == cudaSuccess) out_stats->ms_screen = x; else cudaGetLastError();
        }
        out_stats->n_eval_launches = (int)(n_ev / 2);
        const i64 n_screen = (i64)V * d.n_search;
        out_stats->n_screens = n_screen;
        out_stats->n_evals = (int64_t)h->stats.n_evals;
        // lookups as the reference would do them: evaluations + one per screen + the cost walk of the
        // basic/current pairs + the delay walk of every pool entry
        out_stats->n_lookups = (int64_t)h->stats.n_lookups + n_screen + NS;
        out_stats->n_entries = P;
        out_stats->n_tasks = (int64_t)h->stats.n_tasks;
        out_stats->n_levels = levels_done;
        out_stats->n_kernel_launches = ctx->n_launches;
        out_stats->n_sches_stored = (int64_t)h->stats.n_sches;
        out_stats->n_attempts = ctx->n_attempts;
        out_stats->n_chunks = (int64_t)h->stats.n_chunks;
        out_stats->n_items = (int64_t)h->stats.n_items;
        out_stats->n_lookups_eval = (int64_t)h->stats.n_lookups_eval;
        for (int k = 0; k < 16 && k <= MAX_LEVELS; k++) {
            out_stats->evals_level[k] = (int64_t)h->stats.lvl_evals[k];
            out_stats->lookups_level[k] = (int64_t)h->stats.lvl_lookups[k];
        }
    }
    (void)total_tasks;
    if (ctx->gp_attached) CKR(check_gpool_state(ctx));
    if (mode != AMOD_MODE_SBA && run_levels == max_level && h->n_trips[max_level] > 0)
        return ctx->fail(AMOD_E_LEVELS, "a feasible trip of size cap+4=%d exists; the reference raises IndexError here "
                                        "(dispatch_osp.py:8,226)", max_level);
    return AMOD_OK;
}

extern "C" int amod_osp_build(amod_ctx* ctx, const amod_epoch* ep, int loc, amod_stats* stats) {
    if (!ctx || !ep) return AMOD_E_BADARG;
    if (ep->mode != AMOD_MODE_OSP && ep->mode != AMOD_MODE_OSP_NR) return ctx->fail(AMOD_E_BADARG, "amod_osp_build: mode must be OSP or OSP_NR");
    return dispatch_tt(ctx, [&](auto tt) { return build_impl(ctx, ep, loc, stats, tt); });
}

extern "C" int amod_sba_build(amod_ctx* ctx, const amod_epoch* ep, int loc, amod_stats* stats) {
    if (!ctx || !ep) return AMOD_E_BADARG;
    if (ep->mode != AMOD_MODE_SBA) return ctx->fail(AMOD_E_BADARG, "amod_sba_build: mode must be SBA");
    return dispatch_tt(ctx, [&](auto tt) { return build_impl(ctx, ep, loc, stats, tt); });
}

extern "C" int amod_pool_view(amod_ctx* ctx, amod_pool* v) {
    if (!ctx || !v) return AMOD_E_BADARG;
    if (!ctx->pool_valid) return ctx->fail(AMOD_E_BADARG, "no pool has been built on this context");
    v->n_entries = v->cap_entries = ctx->pool_entries;
    v->n_trip_reqs = v->cap_trip_reqs = ctx->pool_trip_reqs;
    v->n_stops = v->cap_stops = ctx->pool_stops;
    v->ent_veh = ctx->ent_veh.as<int32_t>(); v->ent_kind = ctx->ent_kind.as<int32_t>();
    v->ent_trip_off = (int64_t*)ctx->trip_off.p; v->ent_sche_off = (int64_t*)ctx->sche_off.p;
    v->ent_cost = ctx->ent_cost.as<double>(); v->ent_delay = ctx->ent_delay.as<double>(); v->ent_nfeas = ctx->ent_nfeas.as<int32_t>();
    v->trip_req = ctx->trip_req.as<int32_t>(); v->sche_code = ctx->sche_code.as<int32_t>();
    return AMOD_OK;
}

// Page-locked destinations that are laid out back to back (64 B aligned segments, the order of the parts): the live ranges
// of the nine pool arrays are packed by one kernel and travel as ONE DMA instead of nine (each copy has a fixed latency of
// several microseconds; PoolBuilder.fetch(reuse=True) allocates its buffers that way).
struct PackSegs { const char* src[9]; i64 bytes[9]; i64 at[9]; };
__global__ void k_pack_pool(PackSegs p, char* __restrict__ dst) {
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x, nth = (i64)gridDim.x * blockDim.x;
    for (int g = 0; g < 9; g++) {
        const int4* s = (const int4*)p.src[g];
        int4* d = (int4*)(dst + p.at[g]);
        const i64 n16 = (p.bytes[g] + 15) >> 4;          // (sources are cudaMalloc'd with slack: reading the last partial 16 B is in bounds)
        for (i64 x = tid; x < n16; x += nth) d[x] = s[x];
    }
}
struct FetchPart { void* dst; const void* src; size_t bytes; };
static bool fetch_packed(amod_ctx* ctx, const FetchPart* parts, cudaError_t* ce) {
    size_t at[9], total = 0;
    for (int g = 0; g < 9; g++) { at[g] = total; total += (parts[g].bytes + 63) & ~(size_t)63; }
    for (int g = 0; g < 9; g++)
        if ((char*)parts[g].dst != (char*)parts[0].dst + at[g] || ((uintptr_t)parts[g].src & 15)) return false;
    *ce = ctx->fetch_dev.ensure(total + 64);
    if (*ce != cudaSuccess) return true;
    PackSegs p;
    for (int g = 0; g < 9; g++) { p.src[g] = (const char*)parts[g].src; p.bytes[g] = (i64)parts[g].bytes; p.at[g] = (i64)at[g]; }
    k_pack_pool<<<ctx->n_sm * 2, 256, 0, ctx->stream>>>(p, ctx->fetch_dev.as<char>());
    *ce = cudaMemcpyAsync(parts[0].dst, ctx->fetch_dev.p, total, cudaMemcpyDeviceToHost, ctx->stream);
    if (*ce == cudaSuccess) *ce = cudaStreamSynchronize(ctx->stream);
    return true;
}

extern "C" int amod_pool_fetch(amod_ctx* ctx, amod_pool* o, int loc) {
    if (!ctx || !o) return AMOD_E_BADARG;
    if (!ctx->pool_valid) return ctx->fail(AMOD_E_BADARG, "no pool has been built on this context");
    const i64 P = ctx->pool_entries, NT = ctx->pool_trip_reqs, NS = ctx->pool_stops;
    o->n_entries = P; o->n_trip_reqs = NT; o->n_stops = NS;
    if (o->cap_entries < P || o->cap_trip_reqs < NT || o->cap_stops < NS)
        return ctx->fail(AMOD_E_CAPACITY, "pool buffers too small: need %lld entries, %lld trip requests, %lld stops", P, NT, NS);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    struct Part { void* dst; const void* src; size_t bytes; };
    Part parts[] = {
        {o->ent_veh, ctx->ent_veh.p, 4 * (size_t)P}, {o->ent_kind, ctx->ent_kind.p, 4 * (size_t)P},
        {o->ent_nfeas, ctx->ent_nfeas.p, 4 * (size_t)P}, {o->ent_cost, ctx->ent_cost.p, 8 * (size_t)P},
        {o->ent_delay, ctx->ent_delay.p, 8 * (size_t)P}, {o->ent_trip_off, ctx->trip_off.p, 8 * (size_t)(P + 1)},
        {o->ent_sche_off, ctx->sche_off.p, 8 * (size_t)(P + 1)}, {o->trip_req, ctx->trip_req.p, 4 * (size_t)NT},
        {o->sche_code, ctx->sche_code.p, 4 * (size_t)NS},
    };
    if (loc == AMOD_LOC_DEVICE) {
        for (auto& p : parts) if (p.bytes) CK(cudaMemcpyAsync(p.dst, p.src, p.bytes, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        return AMOD_OK;
    }
    if (loc == AMOD_LOC_PINNED) {      // page-locked host buffers: straight DMA (one copy when they are laid out back to back)
        {
            FetchPart fp[9];
            for (int g = 0; g < 9; g++) { fp[g].dst = parts[g].dst; fp[g].src = parts[g].src; fp[g].bytes = parts[g].bytes; }
            cudaError_t ce = cudaSuccess;
            if (P > 0 && fetch_packed(ctx, fp, &ce)) { CK(ce); return AMOD_OK; }
        }
        for (auto& p : parts) if (p.bytes) CK(cudaMemcpyAsync(p.dst, p.src, p.bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return AMOD_OK;
    }
    // host buffers are pageable: DMA everything into one pinned staging area, then memcpy
    size_t total = 0;
    for (auto& p : parts) total += (p.bytes + 63) & ~(size_t)63;
    if (total > ctx->fetch_pinned_cap) {
        if (ctx->fetch_pinned) cudaFreeHost(ctx->fetch_pinned);
        ctx->fetch_pinned = nullptr; ctx->fetch_pinned_cap = 0;
        CK(cudaMallocHost(&ctx->fetch_pinned, total + total / 2));
        ctx->fetch_pinned_cap = total + total / 2;
    }
    size_t off = 0;
    for (auto& p : parts) {
        if (p.bytes) CK(cudaMemcpyAsync((char*)ctx->fetch_pinned + off, p.src, p.bytes, cudaMemcpyDeviceToHost, st));
        off += (p.bytes + 63) & ~(size_t)63;
    }
    CK(cudaStreamSynchronize(st));
    off = 0;
    for (auto& p : parts) {
        if (p.bytes) memcpy(p.dst, (char*)ctx->fetch_pinned + off, p.bytes);
        off += (p.bytes + 63) & ~(size_t)63;
    }
    return AMOD_OK;
}

extern "C" int amod_host_register(amod_ctx* ctx, void* p, size_t bytes) {
    if (!ctx || !p || !bytes) return AMOD_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    return AMOD_OK;
}

extern "C" int amod_host_unregister(amod_ctx* ctx, void* p) {
    if (!ctx || !p) return AMOD_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostUnregister(p));
    return AMOD_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU: global pool in peer-mapped memory
// ------------------------------------------------------------------------------------------------
static GPoolLayout gpool_layout(i64 ce, i64 ct, i64 cs, i64 max_veh = 0) {
    GPoolLayout L; memset(&L, 0, sizeof L);
    L.cap_ent = ce; L.cap_ptr = ct; L.cap_pst = cs;
    i64 o = 0;
    auto take = [&](i64 bytes) { i64 at = o; o += (bytes + 255) / 256 * 256; return at; };
    L.o_flags = take(8 * 2 * GPOOL_MAX_WORLD); L.o_counts = 0;
    L.o_hdr = take(64); L.o_veh = take(4 * ce); L.o_kind = take(4 * ce); L.o_nfeas = take(4 * ce); L.o_cost = take(8 * ce);
    L.o_delay = take(8 * ce); L.o_toff = take(8 * (ce + 1)); L.o_soff = take(8 * (ce + 1)); L.o_treq = take(4 * ct); L.o_scode = take(4 * cs);
    L.o_counts = take(4 * 3 * (max_veh + GPOOL_MAX_WORLD));     // [world][ceil(max_veh/world)][3] fits for any world
    L.bytes = o;
    return L;
}

static GSlotLayout gslot_layout(i64 cv, i64 ce, i64 ct, i64 cs) {
    GSlotLayout S; memset(&S, 0, sizeof S);
    S.cap_veh = cv; S.cap_ent = ce; S.cap_ptr = ct; S.cap_pst = cs;
    i64 o = 0;
    auto take = [&](i64 bytes) { i64 at = o; o += (bytes + 255) / 256 * 256; return at; };
    S.o_hdr = take(64); S.o_vehoff = take(8 * (cv + 1)); S.o_kind = take(4 * ce); S.o_nfeas = take(4 * ce); S.o_cost = take(8 * ce);
    S.o_delay = take(8 * ce); S.o_toff = take(8 * (ce + 1)); S.o_soff = take(8 * (ce + 1)); S.o_treq = take(4 * ct); S.o_scode = take(4 * cs);
    S.o_cnt = take(4 * 3 * cv);
    S.o_veh = take(4 * ce);
    S.bytes = o;
    return S;
}

extern "C" int amod_gpool_create(amod_ctx* ctx, int64_t cap_entries, int64_t cap_trip_reqs, int64_t cap_stops, int32_t max_veh_total,
                                 int32_t world, void* ipc_handle_out) {
    if (!ctx || !ipc_handle_out || cap_entries <= 0 || cap_trip_reqs <= 0 || cap_stops <= 0 || max_veh_total <= 0 || world < 1 ||
        world > GPOOL_MAX_WORLD) return AMOD_E_BADARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == AMOD_IPC_HANDLE_BYTES, "ipc handle size");
    CK(cudaSetDevice(ctx->device));
    if (ctx->gp_base) return ctx->fail(AMOD_E_BADARG, "global pool already created on this context");
    ctx->gp_layout = gpool_layout(cap_entries, cap_trip_reqs, cap_stops, max_veh_total);
    // staging: one slot per rank, each able to hold twice an even share of the global pool
    ctx->gp_world = world;
    ctx->gp_slot = gslot_layout((max_veh_total + world - 1) / world + 1, 2 * (cap_entries / world) + 1024, 2 * (cap_trip_reqs / world) + 1024,
                                2 * (cap_stops / world) + 1024);
    ctx->gp_o_stage = ctx->gp_layout.bytes;
    CK(cudaMalloc((void**)&ctx->gp_base, (size_t)(ctx->gp_layout.bytes + (i64)world * ctx->gp_slot.bytes)));
    CK(cudaMemset(ctx->gp_base, 0, (size_t)ctx->gp_layout.o_veh));      // flags + header start at zero
    ctx->gp_max_veh = max_veh_total;
    CK(ctx->gp_off.ensure(8 * 3 * ((size_t)max_veh_total + 2)));
    CK(ctx->gp_tot.ensure(sizeof(GpState)));
    CK(cudaMemset(ctx->gp_tot.p, 0, sizeof(GpState)));
    CK(cudaMallocHost((void**)&ctx->gp_h_state, sizeof(GpState)));
    {   // bounded spins: AMOD_GP_TIMEOUT_S seconds (default 30) of the SM clock
        GpState init; memset(&init, 0, sizeof init);
        const char* ts = getenv("AMOD_GP_TIMEOUT_S");
        init.timeout_clk = (long long)((ts ? std::max(1.0, atof(ts)) : 30.0) * 1.9e9);
        CK(cudaMemcpy(ctx->gp_tot.p, &init, sizeof init, cudaMemcpyHostToDevice));
    }
    {   // default placement: vehicles interleaved over the ranks (global v = local v / world of rank v % world)
        std::vector<int> place(2 * (size_t)max_veh_total);
        for (int v = 0; v < max_veh_total; v++) { place[v] = v % world; place[(size_t)max_veh_total + v] = v / world; }
        CK(ctx->gp_place.ensure(4 * place.size()));
        CK(cudaMemcpy(ctx->gp_place.p, place.data(), 4 * place.size(), cudaMemcpyHostToDevice));
        const size_t cv = (size_t)ctx->gp_slot.cap_veh;
        std::vector<int> glob((size_t)world * cv, 0);           // inverse: global vehicle of (rank, local vehicle)
        for (int v = 0; v < max_veh_total; v++) glob[(size_t)(v % world) * cv + (size_t)(v / world)] = v;
        CK(ctx->gp_glob.ensure(4 * glob.size()));
        CK(cudaMemcpy(ctx->gp_glob.p, glob.data(), 4 * glob.size(), cudaMemcpyHostToDevice));
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->gp_base));
    memcpy(ipc_handle_out, &h, sizeof h);
    return AMOD_OK;
}

extern "C" int amod_gpool_open_peers(amod_ctx* ctx, int32_t world, int32_t rank, const void* handles) {
    if (!ctx || !handles || world < 1 || world > GPOOL_MAX_WORLD || rank < 0 || rank >= world) return AMOD_E_BADARG;
    if (!ctx->gp_base) return ctx->fail(AMOD_E_BADARG, "amod_gpool_create first");
    CK(cudaSetDevice(ctx->device));
    memset(&ctx->gp_peers, 0, sizeof ctx->gp_peers);
    ctx->gp_peers.world = world; ctx->gp_peers.rank = rank;
    for (int r = 0; r < world; r++) {
        if (r == rank) { ctx->gp_peers.base[r] = ctx->gp_base; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * AMOD_IPC_HANDLE_BYTES, sizeof h);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->gp_peer_ptr[r] = p;
        ctx->gp_peers.base[r] = (char*)p;
    }
    ctx->gp_open = true;
    return AMOD_OK;
}

static PoolBufs pool_bufs_of(amod_ctx* ctx) {
    PoolBufs pb; memset(&pb, 0, sizeof pb);
    pb.cap_ent = ctx->cap_ent; pb.cap_ptr = ctx->cap_ptr; pb.cap_pst = ctx->cap_pst;
    pb.ent_veh = ctx->ent_veh.as<int>(); pb.ent_kind = ctx->ent_kind.as<int>(); pb.ent_nfeas = ctx->ent_nfeas.as<int>();
    pb.ent_tlen = ctx->ent_tlen.as<int>(); pb.ent_slen = ctx->ent_slen.as<int>(); pb.ent_src = ctx->ent_src.as<int>();
    pb.ent_trip = ctx->ent_trip.as<int>(); pb.ent_cost = ctx->ent_cost.as<double>(); pb.ent_delay = ctx->ent_delay.as<double>();
    pb.trip_off = ctx->trip_off.as<i64>(); pb.sche_off = ctx->sche_off.as<i64>(); pb.trip_req = ctx->trip_req.as<int>();
    pb.sche_code = ctx->sche_code.as<int>();
    return pb;
}

extern "C" int amod_pool_veh_counts(amod_ctx* ctx, int32_t* counts_dev) {
    if (!ctx || !counts_dev) return AMOD_E_BADARG;
    if (!ctx->pool_valid) return ctx->fail(AMOD_E_BADARG, "no pool has been built on this context");
    if (ctx->h_epoch->mode == AMOD_MODE_SBA) return ctx->fail(AMOD_E_BADARG, "per-vehicle counts need a vehicle-major (OSP) pool");
    CK(cudaSetDevice(ctx->device));
    k_veh_counts<<<ctx->n_sm * 2, 256, 0, ctx->stream>>>(ctx->d_epoch.as<DEpoch>(), ctx->veh_off.as<i64>(), ctx->ctl.as<Control>(),
                                                       pool_bufs_of(ctx), counts_dev);
    CK(cudaGetLastError());
    return AMOD_OK;
}

extern "C" int amod_gpool_scatter(amod_ctx* ctx, const int32_t* counts_all_dev, int32_t veh_per_rank_max, int32_t n_veh_total) {
    if (!ctx || !counts_all_dev) return AMOD_E_BADARG;
    if (!ctx->gp_open) return ctx->fail(AMOD_E_BADARG, "amod_gpool_open_peers first");
    if (!ctx->pool_valid) return ctx->fail(AMOD_E_BADARG, "no pool has been built on this context");
    if (n_veh_total > ctx->gp_max_veh) return ctx->fail(AMOD_E_CAPACITY, "global pool was created for at most %d vehicles", ctx->gp_max_veh);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int world = ctx->gp_peers.world;
    i64* off = ctx->gp_off.as<i64>();
    const size_t stride = (size_t)ctx->gp_max_veh + 2;
    i64* tot = ctx->gp_tot.as<i64>();       // GpState: tot[0..2], err at [3]
    CK(cudaMemsetAsync(tot + 3, 0, 8, st));
    for (int c = 0; c < 3; c++)
        scan_dev(ctx, InGlobalVeh{counts_all_dev, world, veh_per_rank_max, c, n_veh_total}, off + c * stride, FinStore{tot + c});
    k_gpool_scatter<<<ctx->n_sm * 4, 256, 0, st>>>(ctx->d_epoch.as<DEpoch>(), ctx->veh_off.as<i64>(), pool_bufs_of(ctx), ctx->gp_layout,
                                                 ctx->gp_peers, off, off + stride, off + 2 * stride, tot, n_veh_total,
                                                 (unsigned long long*)(tot + 3));
    i64 h[4];
    CK(cudaMemcpyAsync(h, tot, sizeof h, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (h[3]) return ctx->fail(AMOD_E_CAPACITY, "global pool too small: need %lld entries, %lld trip requests, %lld stops", h[0], h[1], h[2]);
    return AMOD_OK;
}

// the three assembly kernels on the build stream: push my slice (+ flag) | wait for all slices, global offsets | merge
// (+ "read" flag).  Ranks that do not merge (gather mode, rank != root) only push.  Capturable: no per-epoch argument.
static void enqueue_gpool(amod_ctx* ctx, int run_levels, int max_level, bool sba = false) {
    cudaStream_t st = ctx->stream;
    const GPoolLayout& L = ctx->gp_layout;
    const GSlotLayout& S = ctx->gp_slot;
    const int world = ctx->gp_peers.world, root = ctx->gp_root, n_total = ctx->gp_n_total;
    i64* off = ctx->gp_off.as<i64>();
    const size_t stride = (size_t)ctx->gp_max_veh + 2;
    GpState* gst = ctx->gp_tot.as<GpState>();
    const int* vrank = ctx->gp_place.as<int>();
    const int* vlocal = vrank + ctx->gp_max_veh;
    GPush P; memset(&P, 0, sizeof P);
    const void* srcs[11] = {nullptr, ctx->veh_off.p, ctx->ent_kind.p, ctx->ent_nfeas.p, ctx->ent_cost.p, ctx->ent_delay.p, ctx->trip_off.p,
                            ctx->sche_off.p, ctx->trip_req.p, ctx->sche_code.p, ctx->ent_veh.p};
    const i64 dsts[11] = {S.o_hdr, S.o_vehoff, S.o_kind, S.o_nfeas, S.o_cost, S.o_delay, S.o_toff, S.o_soff, S.o_treq, S.o_scode, S.o_veh};
    for (int g = 0; g < 11; g++) { P.seg[g].src = (const char*)srcs[g]; P.seg[g].dst_off = dsts[g]; P.seg[g].bytes = 0; }
    P.n = 11;
    k_gp_push_bulk<<<ctx->n_sm * 2, 256, 0, st>>>(P, L, S, ctx->gp_o_stage, ctx->gp_peers, ctx->d_epoch.as<DEpoch>(), ctx->ctl.as<Control>(),
                                                gst, ctx->gp_base, root, run_levels, max_level);
    ctx->n_launches++;
    if (root >= 0 && root != ctx->gp_peers.rank) return;
    const char* stage = ctx->gp_base + ctx->gp_o_stage;
    if (sba) {      // request-major merge (sbam.cuh): positions by counting, one prefix sum for the schedule offsets
        SbaMergeBufs m;
        m.pos_of = ctx->sm_pos.as<int>(); m.epos = ctx->sm_epos.as<int>(); m.egv = ctx->sm_egv.as<int>(); m.dest = ctx->sm_dest.as<int>();
        m.rs = ctx->sm_rs.as<int>(); m.base = ctx->sm_base.as<i64>(); m.slen = ctx->sm_slen.as<int>(); m.glob = ctx->gp_glob.as<int>();
        m.cap_ent_slot = S.cap_ent; m.cap_veh = (int)S.cap_veh;
        const DEpoch* d = ctx->d_epoch.as<DEpoch>();
        Control* ctl = ctx->ctl.as<Control>();
        const int G = ctx->n_sm * 4;
        k_sbam_pos<<<G, 256, 0, st>>>(d, m);
        k_sbam_pos2<<<G, 256, 0, st>>>(d, m);
        k_sbam_keys<<<G, 256, 0, st>>>(stage, S, world, d, m, gst, ctx->gp_base, L, ctl);
        k_sbam_reqstart<<<G, 256, 0, st>>>(stage, S, world, d, m, ctl);
        k_sbam_base<<<ctx->n_sm, 256, 0, st>>>(world, d, m, gst, L, n_total, ctl);
        k_sbam_dest<<<G * 2, 256, 0, st>>>(stage, S, ctx->gp_base, L, world, d, m, gst, ctl);
        scan_dev(ctx, InSbaSlen{m.slen, gst}, (i64*)(ctx->gp_base + L.o_soff), FinSbaStops{gst, L.cap_pst});
        k_sbam_codes<<<G * 2, 256, 0, st>>>(stage, S, ctx->gp_base, L, world, m, gst, ctx->gp_peers, ctl);
        ctx->n_launches += 7;
        return;
    }
    k_gp_offsets<<<1, 1024, 0, st>>>(stage, S, world, n_total, off, off + stride, off + 2 * stride, gst, ctx->gp_base, L, vrank, vlocal,
                                     ctx->ctl.as<Control>(), run_levels, max_level);
    k_gp_merge<<<ctx->n_sm * 8, 256, 0, st>>>(stage, S, ctx->gp_base, L, world, off, off + stride, off + 2 * stride, gst, n_total,
                                              ctx->gp_peers, vrank, vlocal, ctx->ctl.as<Control>(), run_levels, max_level);
    ctx->n_launches += 2;
}

static int check_gpool_state(amod_ctx* ctx) {
    GpState* h = ctx->gp_h_state;
    if (h->err) cudaMemsetAsync(&ctx->gp_tot.as<GpState>()->err, 0, 8, ctx->stream);      // reported once
    if (h->err & 2) return ctx->fail(AMOD_E_CUDA, "peer assembly timed out waiting for another rank (the PeerPool must be rebuilt on every rank)");
    if (h->err & 1) return ctx->fail(AMOD_E_CAPACITY, "global pool or staging slot too small (need %lld entries, %lld trip requests, %lld stops)",
                                     (long long)h->tot[0], (long long)h->tot[1], (long long)h->tot[2]);
    return AMOD_OK;
}

extern "C" int amod_gpool_attach(amod_ctx* ctx, int32_t n_veh_total, int32_t root, const int32_t* veh_rank, const int32_t* veh_local) {
    if (!ctx) return AMOD_E_BADARG;
    if (!ctx->gp_open) return ctx->fail(AMOD_E_BADARG, "amod_gpool_open_peers first");
    if (n_veh_total < 0 || n_veh_total > ctx->gp_max_veh) return ctx->fail(AMOD_E_CAPACITY, "global pool was created for at most %d vehicles", ctx->gp_max_veh);
    if (root >= ctx->gp_peers.world) return ctx->fail(AMOD_E_BADARG, "root outside the world");
    CK(cudaSetDevice(ctx->device));
    if ((veh_rank == nullptr) != (veh_local == nullptr)) return ctx->fail(AMOD_E_BADARG, "veh_rank and veh_local come together");
    std::vector<int> place(2 * (size_t)ctx->gp_max_veh, 0);
    const int world = ctx->gp_peers.world;
    for (int v = 0; v < n_veh_total; v++) {
        const int r = veh_rank ? veh_rank[v] : v % world, u = veh_local ? veh_local[v] : v / world;
        if (r < 0 || r >= world || u < 0 || u >= ctx->gp_slot.cap_veh) return ctx->fail(AMOD_E_BADARG, "vehicle %d: placement (%d, %d) out of range", v, r, u);
        place[v] = r; place[(size_t)ctx->gp_max_veh + v] = u;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(ctx->gp_place.p, place.data(), 4 * place.size(), cudaMemcpyHostToDevice));
    {
        const size_t cv = (size_t)ctx->gp_slot.cap_veh;
        std::vector<int> glob((size_t)world * cv, 0);
        for (int v = 0; v < n_veh_total; v++) glob[(size_t)place[v] * cv + (size_t)place[(size_t)ctx->gp_max_veh + v]] = v;
        CK(cudaMemcpy(ctx->gp_glob.p, glob.data(), 4 * glob.size(), cudaMemcpyHostToDevice));
    }
    if (!ctx->gp_attached || ctx->gp_root != root || ctx->gp_n_total != n_veh_total) ctx->buf_gen++;      // the epoch graph changes
    if (!ctx->gp_attached || ctx->gp_root != root) {
        // the readers change: the ranks must have met at a host-side barrier since the last assembly; the next push does not
        // wait for "read" flags of epochs that were read under the previous root / mode
        GpState* gst = ctx->gp_tot.as<GpState>();
        CK(cudaMemcpy(&gst->read_base, &gst->epoch, 8, cudaMemcpyDeviceToDevice));
    }
    ctx->gp_attached = true; ctx->gp_root = root; ctx->gp_n_total = n_veh_total;
    return AMOD_OK;
}

extern "C" int amod_gpool_detach(amod_ctx* ctx) {
    if (!ctx) return AMOD_E_BADARG;
    if (ctx->gp_attached) {
        ctx->buf_gen++;
        GpState* gst = ctx->gp_tot.as<GpState>();
        CK(cudaSetDevice(ctx->device));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaMemcpy(&gst->read_base, &gst->epoch, 8, cudaMemcpyDeviceToDevice));
    }
    ctx->gp_attached = false; ctx->gp_root = -1;
    return AMOD_OK;
}

extern "C" int amod_gpool_assemble(amod_ctx* ctx, int32_t n_veh_total) {
    if (!ctx) return AMOD_E_BADARG;
    if (!ctx->gp_open) return ctx->fail(AMOD_E_BADARG, "amod_gpool_open_peers first");
    if (!ctx->pool_valid) return ctx->fail(AMOD_E_BADARG, "no pool has been built on this context");
    if (ctx->h_epoch->mode == AMOD_MODE_SBA) return ctx->fail(AMOD_E_BADARG, "peer assembly needs a vehicle-major (OSP) pool");
    if (n_veh_total > ctx->gp_max_veh) return ctx->fail(AMOD_E_CAPACITY, "global pool was created for at most %d vehicles", ctx->gp_max_veh);
    if (ctx->gp_attached) return ctx->fail(AMOD_E_BADARG, "the pool is assembled by every build while attached (amod_gpool_attach)");
    if (ctx->gp_peers.world != ctx->gp_world) return ctx->fail(AMOD_E_BADARG, "world size differs from amod_gpool_create");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->gp_n_total = n_veh_total;
    enqueue_gpool(ctx, 0, 0, false);   // (the build is complete: nothing to rerun)
    CK(cudaMemcpyAsync(ctx->gp_h_state, ctx->gp_tot.p, sizeof(GpState), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return check_gpool_state(ctx);
}

extern "C" int amod_gpool_view(amod_ctx* ctx, amod_pool* v) {
    if (!ctx || !v) return AMOD_E_BADARG;
    if (!ctx->gp_base) return ctx->fail(AMOD_E_BADARG, "amod_gpool_create first");
    CK(cudaSetDevice(ctx->device));
    i64 h[3];
    CK(cudaMemcpyAsync(h, ctx->gp_base + ctx->gp_layout.o_hdr, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const GPoolLayout& L = ctx->gp_layout;
    char* B = ctx->gp_base;
    v->cap_entries = L.cap_ent; v->cap_trip_reqs = L.cap_ptr; v->cap_stops = L.cap_pst;
    v->n_entries = h[0]; v->n_trip_reqs = h[1]; v->n_stops = h[2];
    v->ent_veh = (int32_t*)(B + L.o_veh); v->ent_kind = (int32_t*)(B + L.o_kind); v->ent_nfeas = (int32_t*)(B + L.o_nfeas);
    v->ent_cost = (double*)(B + L.o_cost); v->ent_delay = (double*)(B + L.o_delay); v->ent_trip_off = (int64_t*)(B + L.o_toff);
    v->ent_sche_off = (int64_t*)(B + L.o_soff); v->trip_req = (int32_t*)(B + L.o_treq); v->sche_code = (int32_t*)(B + L.o_scode);
    return AMOD_OK;
}

extern "C" int amod_gpool_fetch(amod_ctx* ctx, amod_pool* o) {
    if (!ctx || !o) return AMOD_E_BADARG;
    amod_pool v;
    CKR(amod_gpool_view(ctx, &v));
    o->n_entries = v.n_entries; o->n_trip_reqs = v.n_trip_reqs; o->n_stops = v.n_stops;
    if (o->cap_entries < v.n_entries || o->cap_trip_reqs < v.n_trip_reqs || o->cap_stops < v.n_stops)
        return ctx->fail(AMOD_E_CAPACITY, "pool buffers too small: need %lld entries, %lld trip requests, %lld stops",
                         (long long)v.n_entries, (long long)v.n_trip_reqs, (long long)v.n_stops);
    const size_t P = (size_t)v.n_entries, NT = (size_t)v.n_trip_reqs, NS = (size_t)v.n_stops;
    cudaStream_t st = ctx->stream;
    if (P) {      // one DMA when the (page-locked) destinations are laid out back to back, in the order of amod_pool_fetch
        FetchPart fp[9] = {{o->ent_veh, v.ent_veh, 4 * P}, {o->ent_kind, v.ent_kind, 4 * P}, {o->ent_nfeas, v.ent_nfeas, 4 * P},
                           {o->ent_cost, v.ent_cost, 8 * P}, {o->ent_delay, v.ent_delay, 8 * P}, {o->ent_trip_off, v.ent_trip_off, 8 * (P + 1)},
                           {o->ent_sche_off, v.ent_sche_off, 8 * (P + 1)}, {o->trip_req, v.trip_req, 4 * NT}, {o->sche_code, v.sche_code, 4 * NS}};
        cudaPointerAttributes pa;
        const bool pinned = cudaPointerGetAttributes(&pa, o->ent_veh) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
        cudaError_t ce = cudaSuccess;
        if (pinned && fetch_packed(ctx, fp, &ce)) { CK(ce); return AMOD_OK; }
    }
    if (P) {
        CK(cudaMemcpyAsync(o->ent_veh, v.ent_veh, 4 * P, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(o->ent_kind, v.ent_kind, 4 * P, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(o->ent_nfeas, v.ent_nfeas, 4 * P, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(o->ent_cost, v.ent_cost, 8 * P, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(o->ent_delay, v.ent_delay, 8 * P, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaMemcpyAsync(o->ent_trip_off, v.ent_trip_off, 8 * (P + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(o->ent_sche_off, v.ent_sche_off, 8 * (P + 1), cudaMemcpyDeviceToHost, st));
    if (NT) CK(cudaMemcpyAsync(o->trip_req, v.trip_req, 4 * NT, cudaMemcpyDeviceToHost, st));
    if (NS) CK(cudaMemcpyAsync(o->sche_code, v.sche_code, 4 * NS, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return AMOD_OK;
}

// ------------------------------------------------------------------------------------------------
// greedy assignment over the device-resident pool (next row)
// ------------------------------------------------------------------------------------------------
extern "C" int amod_greedy_assign(amod_ctx* ctx, int32_t* order_out, int32_t* sel_out, int32_t* n_sel, int loc, int32_t* n_rounds) {
    if (!ctx || !order_out || !sel_out || !n_sel) return AMOD_E_BADARG;
    if (!ctx->pool_valid) return ctx->fail(AMOD_E_BADARG, "no pool has been built on this context");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const i64 P = ctx->pool_entries;
    const int V = ctx->h_epoch->n_veh, R = ctx->h_epoch->n_req;
    *n_sel = 0;
    if (n_rounds) *n_rounds = 0;
    if (P == 0) return AMOD_OK;
    if (P > 0x7ffffff0) return ctx->fail(AMOD_E_LIMIT, "pool too large for 32-bit ranks");
    Control* ctl = ctx->ctl.as<Control>();
    const PoolBufs pb = pool_bufs_of(ctx);
    const size_t Pz = (size_t)P;
    const int nb = std::max(1, ctx->h_ctl->max_tlen + 1);      // score = len(trip): buckets 0 .. the pool's longest trip
    if (nb > GA_BUCKETS) return ctx->fail(AMOD_E_LIMIT, "a pair score is outside 0..%d", GA_BUCKETS - 1);
    CK(ctx->ga_pos.ensure(8 * ((size_t)nb * Pz + 2))); CK(ctx->ga_rank.ensure(4 * Pz)); CK(ctx->ga_order.ensure(4 * Pz));
    CK(ctx->ga_state.ensure(Pz)); CK(ctx->ga_bv.ensure(8 * (size_t)(V + 1))); CK(ctx->ga_br.ensure(8 * (size_t)(R + 1)));
    CK(ctx->ga_tv.ensure((size_t)V + 1)); CK(ctx->ga_tr.ensure((size_t)R + 1)); CK(ctx->ga_sel.ensure(4 * Pz)); CK(ctx->ga_cnt.ensure(32));
    CK(ctx->partial.ensure(2 * 8 * (size_t)(ctx->n_sm + 1)));
    const int G = ctx->n_sm * 2;
    i64* pos = ctx->ga_pos.as<i64>();
    unsigned long long* err = (unsigned long long*)(ctx->ga_cnt.as<char>() + 8);
    i64* tot = (i64*)(ctx->ga_cnt.as<char>() + 16);
    int* n_left = ctx->ga_cnt.as<int>();
    CK(cudaMemsetAsync(ctx->ga_cnt.p, 0, 32, st));
    // 1. rank = position after the stable sort by -score (counting sort through one prefix sum)
    scan_dev(ctx, InScoreBucket{pb.ent_tlen, ctl, nb}, pos, FinNone{});
    k_ga_init<<<G, 256, 0, st>>>(ctl, pb.ent_tlen, pos, ctx->ga_rank.as<int>(), ctx->ga_order.as<int>(), ctx->ga_state.as<uint8_t>(),
                                ctx->ga_bv.as<int>(), V, ctx->ga_br.as<int>(), R, ctx->ga_tv.as<uint8_t>(), ctx->ga_tr.as<uint8_t>(), err, nb);
    // 2. rounds: keep, then drop + post the survivors' ranks into the other set of minima (two launches per round)
    int rounds = 0;
    int* bv[2] = {ctx->ga_bv.as<int>(), ctx->ga_bv.as<int>() + V};
    int* br[2] = {ctx->ga_br.as<int>(), ctx->ga_br.as<int>() + R};
    k_ga_min<<<G, 256, 0, st>>>(ctl, pb, ctx->ga_rank.as<int>(), ctx->ga_state.as<uint8_t>(), bv[0], br[0]);
    bool done = false;
    {   // every round in one cooperative launch (AMOD_GA_NO_PERSIST=1: two launches per round, a host check every four rounds)
        static int occ = -1;
        if (occ < 0) { if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ga_rounds, 256, 0) != cudaSuccess) occ = 0; cudaGetLastError(); }
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        if (coop && occ >= 2 && !getenv("AMOD_GA_NO_PERSIST")) {
            CK(ctx->ga_sync.ensure(32));
            CK(cudaMemsetAsync(ctx->ga_sync.p, 0, 32, st));      // left[2] | barrier | error | rounds
            int* d_left = ctx->ga_sync.as<int>(); unsigned* d_bar = (unsigned*)(d_left + 2); int* d_err = d_left + 3; int* d_rounds = d_left + 4;
            const Control* cctl = ctl; PoolBufs pbb = pb;
            const int* rk = ctx->ga_rank.as<int>(); uint8_t* stt = ctx->ga_state.as<uint8_t>();
            int* bv0 = bv[0]; int* br0 = br[0]; int* bv1 = bv[1]; int* br1 = br[1];
            int nv = V, nr = R, max_rounds = 4 * (V + 8);
            uint8_t* tv = ctx->ga_tv.as<uint8_t>(); uint8_t* tr = ctx->ga_tr.as<uint8_t>();
            void* args[] = {&cctl, &pbb, &rk, &stt, &bv0, &br0, &bv1, &br1, &nv, &nr, &tv, &tr, &d_left, &d_bar, &d_err, &max_rounds, &d_rounds};
            if (cudaLaunchCooperativeKernel((void*)k_ga_rounds, dim3((unsigned)G), dim3(256), args, 0, st) == cudaSuccess) {
                int h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                CK(cudaMemcpyAsync(h, ctx->ga_sync.p, 32, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                if (h[3]) return ctx->fail(AMOD_E_CUDA, "greedy assignment: the device-wide barrier timed out");
                rounds = h[4];
                if (rounds >= max_rounds) return ctx->fail(AMOD_E_CUDA, "greedy assignment did not converge");
                done = true;
            } else cudaGetLastError();
        }
    }
    while (!done) {
        for (int rep = 0; rep < 4; rep++) {       // four rounds per host check
            const int cur = rounds & 1;
            if (rep == 3) CK(cudaMemsetAsync(n_left, 0, 4, st));
            k_ga_keep<<<G, 256, 0, st>>>(ctl, pb, ctx->ga_rank.as<int>(), ctx->ga_state.as<uint8_t>(), bv[cur], br[cur],
                                        ctx->ga_tv.as<uint8_t>(), ctx->ga_tr.as<uint8_t>());
            k_ga_drop<<<G, 256, 0, st>>>(ctl, pb, ctx->ga_rank.as<int>(), ctx->ga_state.as<uint8_t>(), ctx->ga_tv.as<uint8_t>(), ctx->ga_tr.as<uint8_t>(),
                                        bv[cur], V, br[cur], R, bv[cur ^ 1], br[cur ^ 1], n_left);
            rounds++;
        }
        int left = 0;
        CK(cudaMemcpyAsync(&left, n_left, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (left == 0) break;
        if (rounds > 4 * (V + 8)) return ctx->fail(AMOD_E_CUDA, "greedy assignment did not converge");
    }
    // 3. selected positions in the sorted list, ascending
    scan_dev(ctx, InKeptByRank{ctx->ga_order.as<int>(), ctx->ga_state.as<uint8_t>(), ctl}, pos, FinStore{tot});
    k_ga_emit<<<G, 256, 0, st>>>(ctl, ctx->ga_order.as<int>(), ctx->ga_state.as<uint8_t>(), pos, ctx->ga_sel.as<int>());
    i64 h[2] = {0, 0};
    CK(cudaMemcpyAsync(&h[0], tot, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h[1], err, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h[1]) return ctx->fail(AMOD_E_LIMIT, "a pair score is outside 0..%d", GA_BUCKETS - 1);
    const cudaMemcpyKind kind = loc == AMOD_LOC_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    CK(cudaMemcpyAsync(order_out, ctx->ga_order.p, 4 * Pz, kind, st));
    if (h[0]) CK(cudaMemcpyAsync(sel_out, ctx->ga_sel.p, 4 * (size_t)h[0], kind, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    *n_sel = (int32_t)h[0];
    ctx->ga_n_sel = (int)h[0];
    if (n_rounds) *n_rounds = rounds;
    return AMOD_OK;
}

// ------------------------------------------------------------------------------------------------
// NPO
// ------------------------------------------------------------------------------------------------
template <typename TT>
static int npo_impl(amod_ctx* ctx, const int32_t* idle_nid, int n_idle, const int32_t* pend_onid, int n_pend, int loc,
                    int32_t* out_idle, int32_t* out_pend, double* out_dt, int32_t* n_out, int32_t* n_rounds, Table<TT> tt) {
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    *n_out = 0;
    if (n_rounds) *n_rounds = 0;
    const int target = std::min(n_idle, n_pend);
    if (target == 0) return AMOD_OK;
    const int* d_idle = idle_nid; const int* d_pend = pend_onid;
    if (loc == AMOD_LOC_HOST) {
        for (int i = 0; i < n_idle; i++) if (idle_nid[i] < 1 || idle_nid[i] > ctx->n_nodes) return ctx->fail(AMOD_E_BADARG, "idle node out of range");
        for (int i = 0; i < n_pend; i++) if (pend_onid[i] < 1 || pend_onid[i] > ctx->n_nodes) return ctx->fail(AMOD_E_BADARG, "pending node out of range");
        CK(ctx->npo_in.ensure(4 * (size_t)(n_idle + n_pend)));
        CK(cudaMemcpyAsync(ctx->npo_in.p, idle_nid, 4 * (size_t)n_idle, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->npo_in.as<int>() + n_idle, pend_onid, 4 * (size_t)n_pend, cudaMemcpyHostToDevice, st));
        d_idle = ctx->npo_in.as<int>(); d_pend = ctx->npo_in.as<int>() + n_idle;
    }
    // groups of idle vehicles standing at the same node (npo.cuh): node, vehicles in fleet order, head pointer
    std::vector<int> h_idle;
    const int* hi = idle_nid;
    if (loc == AMOD_LOC_DEVICE) {
        h_idle.resize(n_idle);
        CK(cudaMemcpyAsync(h_idle.data(), idle_nid, 4 * (size_t)n_idle, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        hi = h_idle.data();
        for (int i = 0; i < n_idle; i++) if (hi[i] < 1 || hi[i] > ctx->n_nodes) return ctx->fail(AMOD_E_BADARG, "idle node out of range");
    }
    // counting sort by node (stable: vehicles of a group stay in fleet order)
    std::vector<int> by_node(n_idle), g_node, g_start;
    {
        std::vector<int> cnt((size_t)ctx->n_nodes + 2, 0);
        for (int i = 0; i < n_idle; i++) cnt[hi[i] + 1]++;
        for (int nd = 1; nd <= ctx->n_nodes; nd++)
            if (cnt[nd + 1]) { g_node.push_back(nd); g_start.push_back(cnt[nd]); cnt[nd + 1] += cnt[nd]; } else cnt[nd + 1] = cnt[nd];
        std::vector<int> pos((size_t)ctx->n_nodes + 2, 0);
        for (size_t gi = 0; gi < g_node.size(); gi++) pos[g_node[gi]] = g_start[gi];
        for (int i = 0; i < n_idle; i++) by_node[pos[hi[i]]++] = i;
    }
    g_start.push_back(n_idle);
    const int G = (int)g_node.size();
    // packed upload: node[G] | start[G+1] | veh[n_idle] | head[G] (zero) | hrank[G] (set on the device)
    std::vector<int> packed((size_t)G + (size_t)G + 1 + (size_t)n_idle + 2 * (size_t)G, 0);
    memcpy(packed.data(), g_node.data(), 4 * (size_t)G);
    memcpy(packed.data() + G, g_start.data(), 4 * ((size_t)G + 1));
    memcpy(packed.data() + 2 * (size_t)G + 1, by_node.data(), 4 * (size_t)n_idle);
    CK(ctx->npo_grp.ensure(4 * packed.size()));
    CK(cudaMemcpyAsync(ctx->npo_grp.p, packed.data(), 4 * packed.size(), cudaMemcpyHostToDevice, st));
    NpoGroups grp;
    grp.node = ctx->npo_grp.as<int>(); grp.start = grp.node + G; grp.veh = grp.start + G + 1;
    grp.head = ctx->npo_grp.as<int>() + 2 * (size_t)G + 1 + (size_t)n_idle; grp.hrank = grp.head + G; grp.n_groups = G;
    CK(ctx->npo_vm.ensure(4 * (size_t)n_idle)); CK(ctx->npo_rm.ensure(4 * (size_t)n_pend)); CK(ctx->npo_mdt.ensure(8 * (size_t)n_pend));
    CK(ctx->npo_cnt.ensure(32));
    if (2.0 * sizeof(TT) * (double)G * (double)n_pend > 32e9)      // the group x request travel times are held twice (both orientations)
        return ctx->fail(AMOD_E_LIMIT, "npo: %d vehicle nodes x %d requests need more than 32 GB of travel times", G, n_pend);
    CK(ctx->npo_D.ensure(2 * sizeof(TT) * (size_t)G * (size_t)n_pend)); CK(ctx->npo_best.ensure(4 * (size_t)n_pend)); CK(ctx->npo_tie.ensure((size_t)n_pend));
    CK(ctx->npo_scr.ensure(sizeof(NpoPick) * (size_t)n_pend));
    CK(cudaMemsetAsync(ctx->npo_vm.p, 0xff, 4 * (size_t)n_idle, st));
    CK(cudaMemsetAsync(ctx->npo_rm.p, 0xff, 4 * (size_t)n_pend, st));
    CK(cudaMemsetAsync(ctx->npo_cnt.p, 0, 32, st));      // matched | - | cursor (8 B) | barrier | rounds | error
    int* d_cnt = ctx->npo_cnt.as<int>();
    unsigned long long* d_cursor = (unsigned long long*)(ctx->npo_cnt.as<char>() + 8);
    TT* D = ctx->npo_D.as<TT>();
    TT* DT = D + (size_t)G * (size_t)n_pend;
    k_npo_matrix<TT><<<ctx->n_sm * 8, 256, 0, st>>>(tt, grp, d_pend, n_pend, D, DT);
    int matched = 0, rounds = 0;
    // every round in one cooperative launch when all the groups' blocks are resident together (AMOD_NPO_NO_PERSIST=1: never)
    {
        static int occ_blocks = -1;
        if (occ_blocks < 0) { if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_blocks, k_npo_rounds<TT>, NPO_THREADS, 0) != cudaSuccess) occ_blocks = 0; cudaGetLastError(); }
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        if (coop && G <= occ_blocks * ctx->n_sm && !getenv("AMOD_NPO_NO_PERSIST")) {
            unsigned* d_bar = (unsigned*)(ctx->npo_cnt.as<char>() + 16);
            int* d_rounds = (int*)(ctx->npo_cnt.as<char>() + 20);
            int* d_err = (int*)(ctx->npo_cnt.as<char>() + 24);
            const TT* cD = D; const TT* cDT = DT;
            int np = n_pend, tg = target, max_rounds = target + 8;
            int* vm = ctx->npo_vm.as<int>(); int* rmm = ctx->npo_rm.as<int>(); double* md = ctx->npo_mdt.as<double>();
            int* bgp = ctx->npo_best.as<int>(); uint8_t* tp = ctx->npo_tie.as<uint8_t>(); NpoPick* scr = ctx->npo_scr.as<NpoPick>();
            void* args[] = {&cD, &cDT, &grp, &np, &vm, &rmm, &md, &bgp, &tp, &scr, &d_cursor, &d_cnt, &tg, &d_bar, &max_rounds, &d_rounds, &d_err};
            if (cudaLaunchCooperativeKernel((void*)k_npo_rounds<TT>, dim3((unsigned)(occ_blocks * ctx->n_sm)), dim3(NPO_THREADS), args, 0, st) == cudaSuccess) {
                int h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                CK(cudaMemcpyAsync(h, ctx->npo_cnt.p, 32, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                if (h[6]) return ctx->fail(AMOD_E_CUDA, "npo: the device-wide barrier timed out");
                matched = h[0]; rounds = h[5];
            } else cudaGetLastError();       // (not launchable after all: the per-round launches below do the work)
        }
    }
    int per_sync = 2;             // rounds between two looks at the match count (a round after the last match returns at once):
    while (matched < target) {    // 2, 4, 8, 8, ... - a handful of idle vehicles is matched in one or two rounds
        for (int i = 0; i < per_sync; i++) {
            k_npo_best<TT><<<(n_pend + 7) / 8, 256, 0, st>>>(DT, grp, n_pend, ctx->npo_rm.as<int>(), ctx->npo_best.as<int>(), ctx->npo_tie.as<uint8_t>(),
                                                                d_cnt, target, d_cursor);
            k_npo_accept<TT><<<G, NPO_THREADS, 0, st>>>(D, grp, n_pend, ctx->npo_vm.as<int>(), ctx->npo_rm.as<int>(), ctx->npo_mdt.as<double>(),
                                                        ctx->npo_best.as<int>(), ctx->npo_tie.as<uint8_t>(), ctx->npo_scr.as<NpoPick>(), d_cursor, d_cnt);
        }
        rounds += per_sync;
        per_sync = std::min(8, per_sync * 2);
        int prev = matched;
        CK(cudaMemcpyAsync(&matched, ctx->npo_cnt.p, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (matched == prev) return ctx->fail(AMOD_E_CUDA, "npo: no progress in round %d", rounds);
    }
    // selection order of the reference: ascending (dt, request rank, vehicle rank)
    if (n_pend <= NPO_ORDER_MAX) {      // ordered on the device: three short copies come back
        const size_t M = (size_t)matched;
        CK(ctx->npo_out.ensure(16 * (size_t)target + 64));
        int* o_idle = ctx->npo_out.as<int>(); int* o_pend = o_idle + target; double* o_dt = (double*)(ctx->npo_out.as<char>() + 8 * (((size_t)target + 7) & ~(size_t)7));
        k_npo_order<<<(n_pend * 32 + 255) / 256, 256, 0, st>>>(ctx->npo_rm.as<int>(), ctx->npo_mdt.as<double>(), n_pend, o_idle, o_pend, o_dt);
        CK(cudaMemcpyAsync(out_idle, o_idle, 4 * M, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out_pend, o_pend, 4 * M, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out_dt, o_dt, 8 * M, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        *n_out = matched;
        if (n_rounds) *n_rounds = rounds;
        return AMOD_OK;
    }
    std::vector<int> rm(n_pend);
    std::vector<double> mdt(n_pend);
    CK(cudaMemcpyAsync(rm.data(), ctx->npo_rm.p, 4 * (size_t)n_pend, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(mdt.data(), ctx->npo_mdt.p, 8 * (size_t)n_pend, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    struct Sel { double dt; int r; };
    std::vector<Sel> order;
    order.reserve((size_t)target);
    for (int r = 0; r < n_pend; r++) if (rm[r] >= 0) order.push_back(Sel{mdt[r], r});
    std::sort(order.begin(), order.end(), [](const Sel& a, const Sel& b) { return a.dt < b.dt || (a.dt == b.dt && a.r < b.r); });
    for (size_t i = 0; i < order.size(); i++) { out_pend[i] = order[i].r; out_idle[i] = rm[order[i].r]; out_dt[i] = order[i].dt; }
    *n_out = (int)order.size();
    if (n_rounds) *n_rounds = rounds;
    return AMOD_OK;
}

extern "C" int amod_npo_match(amod_ctx* ctx, const int32_t* idle_nid, int32_t n_idle, const int32_t* pend_onid, int32_t n_pend,
                              int loc, int32_t* out_idle_idx, int32_t* out_pend_idx, double* out_dt, int32_t* n_out, int32_t* n_rounds) {
    if (!ctx || !n_out || n_idle < 0 || n_pend < 0) return AMOD_E_BADARG;
    return dispatch_tt(ctx, [&](auto tt) {
        return npo_impl(ctx, idle_nid, n_idle, pend_onid, n_pend, loc, out_idle_idx, out_pend_idx, out_dt, n_out, n_rounds, tt);
    });
}

// ------------------------------------------------------------------------------------------------
// batched route construction
// ------------------------------------------------------------------------------------------------
extern "C" int amod_routes_upload(amod_ctx* ctx, const int32_t* path_table_host, const void* dist_table_host) {
    if (!ctx || !path_table_host || !dist_table_host) return AMOD_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    const size_t n2 = (size_t)ctx->n_nodes * ctx->n_nodes;
    const size_t el = ctx->tt_dtype == AMOD_TT_F64 ? 8 : 4;
    if (!ctx->d_pred) CK(cudaMalloc((void**)&ctx->d_pred, 4 * n2));
    if (!ctx->d_dist) CK(cudaMalloc(&ctx->d_dist, el * n2));
    CK(cudaMemcpy(ctx->d_pred, path_table_host, 4 * n2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_dist, dist_table_host, el * n2, cudaMemcpyHostToDevice));
    return AMOD_OK;
}

template <typename TT>
static int routes_impl(amod_ctx* ctx, const int32_t* leg_o, const int32_t* leg_d, int n_legs, int loc, amod_routes* out, Table<TT> tt) {
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    out->n_steps = 0;
    if (n_legs == 0) return AMOD_OK;
    const size_t L = (size_t)n_legs;
    const int* d_o = leg_o; const int* d_d = leg_d;
    if (loc != AMOD_LOC_DEVICE) {
        CK(ctx->rt_in.ensure(8 * L));
        CK(cudaMemcpyAsync(ctx->rt_in.p, leg_o, 4 * L, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->rt_in.as<int>() + L, leg_d, 4 * L, cudaMemcpyHostToDevice, st));
        d_o = ctx->rt_in.as<int>(); d_d = ctx->rt_in.as<int>() + L;
    }
    CK(ctx->rt_len.ensure(4 * L)); CK(ctx->rt_off.ensure(8 * (L + 2))); CK(ctx->rt_lt.ensure(8 * L)); CK(ctx->rt_ld.ensure(8 * L));
    CK(ctx->rt_err.ensure(16)); CK(ctx->partial.ensure(2 * 8 * (size_t)(ctx->n_sm + 1)));
    CK(cudaMemsetAsync(ctx->rt_err.p, 0, 16, st));
    RouteTables rt; rt.pred = ctx->d_pred; rt.dist = ctx->d_dist; rt.n = ctx->n_nodes;
    unsigned long long* err = ctx->rt_err.as<unsigned long long>();
    i64* tot = (i64*)(err + 1);
    k_route_count<<<ctx->n_sm * 4, 256, 0, st>>>(rt, d_o, d_d, n_legs, ctx->rt_len.as<int>(), err);
    scan_dev(ctx, InLegLen{ctx->rt_len.as<int>(), (i64)n_legs}, ctx->rt_off.as<i64>(), FinStoreI64{tot});
    unsigned long long h[2];
    CK(cudaMemcpyAsync(h, err, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h[0]) return ctx->fail(AMOD_E_BADARG, "a leg's node is out of range or the path table does not lead back to its origin");
    const i64 n_steps = (i64)h[1];
    out->n_steps = n_steps;
    if (out->cap_steps < n_steps) return ctx->fail(AMOD_E_CAPACITY, "step buffers too small: need %lld steps", (long long)n_steps);
    const size_t S = (size_t)n_steps;
    const bool dev = loc == AMOD_LOC_DEVICE;
    int* su = out->step_u; int* sv = out->step_v; double* stt = out->step_t; double* sdd = out->step_d;
    double* lt = out->leg_t; double* ld = out->leg_d; i64* lo = (i64*)out->leg_off;
    if (!dev) {
        CK(ctx->rt_u.ensure(4 * S)); CK(ctx->rt_v.ensure(4 * S)); CK(ctx->rt_t.ensure(8 * S)); CK(ctx->rt_d.ensure(8 * S));
        su = ctx->rt_u.as<int>(); sv = ctx->rt_v.as<int>(); stt = ctx->rt_t.as<double>(); sdd = ctx->rt_d.as<double>();
        lt = ctx->rt_lt.as<double>(); ld = ctx->rt_ld.as<double>();
    }
    k_route_fill<TT><<<ctx->n_sm * 8, 256, 0, st>>>(rt, tt, d_o, d_d, n_legs, ctx->rt_off.as<i64>(), n_steps, su, sv, stt, sdd, lt, ld);
    const cudaMemcpyKind kind = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    CK(cudaMemcpyAsync(lo, ctx->rt_off.p, 8 * (L + 1), kind, st));
    if (!dev) {
        CK(cudaMemcpyAsync(out->step_u, su, 4 * S, kind, st)); CK(cudaMemcpyAsync(out->step_v, sv, 4 * S, kind, st));
        CK(cudaMemcpyAsync(out->step_t, stt, 8 * S, kind, st)); CK(cudaMemcpyAsync(out->step_d, sdd, 8 * S, kind, st));
        CK(cudaMemcpyAsync(out->leg_t, lt, 8 * L, kind, st)); CK(cudaMemcpyAsync(out->leg_d, ld, 8 * L, kind, st));
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return AMOD_OK;
}

extern "C" int amod_routes_build(amod_ctx* ctx, const int32_t* leg_o, const int32_t* leg_d, int32_t n_legs, int loc, amod_routes* out) {
    if (!ctx || !out || n_legs < 0 || (n_legs && (!leg_o || !leg_d))) return AMOD_E_BADARG;
    if (!ctx->d_pred) return ctx->fail(AMOD_E_BADARG, "amod_routes_upload first");
    return dispatch_tt(ctx, [&](auto tt) { return routes_impl(ctx, leg_o, leg_d, n_legs, loc, out, tt); });
}

// ------------------------------------------------------------------------------------------------
// epoch state resident in HBM, updated by deltas
// ------------------------------------------------------------------------------------------------
extern "C" int amod_state_create(amod_ctx* ctx, int32_t n_veh, int32_t max_stops, int32_t req_cap) {
    if (!ctx || n_veh < 0 || max_stops < 1 || max_stops > AMOD_MAX_STOPS || req_cap < 1) return AMOD_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    const size_t V = (size_t)std::max(n_veh, 1), S = (size_t)max_stops, R = (size_t)req_cap;
    CK(ctx->fs_nid.ensure(4 * V)); CK(ctx->fs_t.ensure(8 * V)); CK(ctx->fs_load.ensure(4 * V)); CK(ctx->fs_status.ensure(4 * V));
    CK(ctx->fs_rnode.ensure(4 * V)); CK(ctx->fs_rddl.ensure(8 * V)); CK(ctx->fs_len.ensure(4 * V)); CK(ctx->fs_code.ensure(4 * V * S)); CK(ctx->fs_onb.ensure(V * S));
    CK(ctx->fs_off64.ensure(8 * (V + 2))); CK(ctx->fs_off32.ensure(4 * (V + 2))); CK(ctx->fs_ccode.ensure(4 * V * S)); CK(ctx->fs_conb.ensure(V * S));
    CK(ctx->fs_err.ensure(8));
    CK(ctx->rq_onid.ensure(4 * R)); CK(ctx->rq_dnid.ensure(4 * R)); CK(ctx->rq_clp.ensure(8 * R)); CK(ctx->rq_cld.ensure(8 * R));
    CK(ctx->rq_tr.ensure(8 * R)); CK(ctx->rq_ts.ensure(8 * R));
    DevBuf* zero[] = {&ctx->fs_nid, &ctx->fs_t, &ctx->fs_load, &ctx->fs_status, &ctx->fs_rnode, &ctx->fs_rddl, &ctx->fs_len, &ctx->fs_code, &ctx->fs_onb};
    for (DevBuf* b : zero) CK(cudaMemset(b->p, 0, b->cap));
    FleetState& fs = ctx->fs;
    fs.nid = ctx->fs_nid.as<int>(); fs.t = ctx->fs_t.as<double>(); fs.load = ctx->fs_load.as<int>(); fs.status = ctx->fs_status.as<int>();
    fs.reb_node = ctx->fs_rnode.as<int>(); fs.reb_ddl = ctx->fs_rddl.as<double>(); fs.len = ctx->fs_len.as<int>(); fs.code = ctx->fs_code.as<int>();
    fs.onboard = ctx->fs_onb.as<uint8_t>(); fs.n_veh = n_veh; fs.max_stops = max_stops;
    ctx->fs_valid = true; ctx->fs_req_cap = req_cap; ctx->fs_req_n = 0; ctx->fs_max_len = 0;
    ctx->buf_gen++;
    return AMOD_OK;
}

extern "C" int amod_state_add_requests(amod_ctx* ctx, int32_t first, int32_t n, const int32_t* onid, const int32_t* dnid, const double* clp,
                                       const double* cld, const double* tr, const double* ts) {
    if (!ctx || n < 0 || first < 0) return AMOD_E_BADARG;
    if (!ctx->fs_valid) return ctx->fail(AMOD_E_BADARG, "amod_state_create first");
    if (n == 0) return AMOD_OK;
    if ((i64)first + n > ctx->fs_req_cap) return ctx->fail(AMOD_E_CAPACITY, "request table holds %d requests", ctx->fs_req_cap);
    for (int i = 0; i < n; i++)
        if (onid[i] < 1 || onid[i] > ctx->n_nodes || dnid[i] < 1 || dnid[i] > ctx->n_nodes) return ctx->fail(AMOD_E_BADARG, "request %d: node id out of range", first + i);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t N = (size_t)n, F = (size_t)first;
    CK(cudaMemcpyAsync(ctx->rq_onid.as<int>() + F, onid, 4 * N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->rq_dnid.as<int>() + F, dnid, 4 * N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->rq_clp.as<double>() + F, clp, 8 * N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->rq_cld.as<double>() + F, cld, 8 * N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->rq_tr.as<double>() + F, tr, 8 * N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->rq_ts.as<double>() + F, ts, 8 * N, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));      // (the host arrays may be reused by the caller)
    ctx->fs_req_n = std::max(ctx->fs_req_n, first + n);
    return AMOD_OK;
}

extern "C" int amod_state_update_vehicles(amod_ctx* ctx, int32_t n, const int32_t* veh_idx, const int32_t* nid, const double* t,
                                          const int32_t* load, const int32_t* status, const int32_t* reb_node, const double* reb_ddl,
                                          const int32_t* sche_off, const int32_t* sche_code, const uint8_t* sche_onboard) {
    if (!ctx || n < 0) return AMOD_E_BADARG;
    if (!ctx->fs_valid) return ctx->fail(AMOD_E_BADARG, "amod_state_create first");
    if (n == 0) return AMOD_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t N = (size_t)n, NS = (size_t)sche_off[n];
    for (int i = 0; i < n; i++) {
        if (nid[i] < 1 || nid[i] > ctx->n_nodes) return ctx->fail(AMOD_E_BADARG, "vehicle %d: node id out of range", veh_idx[i]);
        if (sche_off[i + 1] - sche_off[i] > ctx->fs.max_stops) return ctx->fail(AMOD_E_LIMIT, "vehicle %d: schedule longer than the state's %d stops", veh_idx[i], ctx->fs.max_stops);
        ctx->fs_max_len = std::max(ctx->fs_max_len, sche_off[i + 1] - sche_off[i]);      // monotone: the SBA / OSP-NR graph and strides stay put
        for (int k = sche_off[i]; k < sche_off[i + 1]; k++) {
            const int c = sche_code[k];
            if (c < -1 || (c >> 1) >= ctx->fs_req_n) return ctx->fail(AMOD_E_BADARG, "vehicle %d: stop code %d names an unknown request", veh_idx[i], c);
            if (c == -1 && (reb_node[i] < 1 || reb_node[i] > ctx->n_nodes)) return ctx->fail(AMOD_E_BADARG, "vehicle %d: rebalancing node out of range", veh_idx[i]);
        }
    }
    // one packed staging buffer, one copy
    struct Part { const void* src; size_t bytes; size_t at; };
    Part parts[] = {{t, 8 * N, 0}, {reb_ddl, 8 * N, 0}, {veh_idx, 4 * N, 0}, {nid, 4 * N, 0}, {load, 4 * N, 0}, {status, 4 * N, 0}, {reb_node, 4 * N, 0},
                    {sche_off, 4 * (N + 1), 0}, {sche_code, 4 * NS, 0}, {sche_onboard, NS, 0}};
    size_t total = 0;
    for (auto& p : parts) { p.at = total; total += (p.bytes + 15) & ~(size_t)15; }
    if (total > ctx->fs_pinned_cap) {
        if (ctx->fs_pinned) cudaFreeHost(ctx->fs_pinned);
        ctx->fs_pinned = nullptr; ctx->fs_pinned_cap = 0;
        CK(cudaMallocHost(&ctx->fs_pinned, total * 2));
        ctx->fs_pinned_cap = total * 2;
    }
    CK(ctx->fs_stage.ensure(total));
    for (auto& p : parts) if (p.bytes) memcpy((char*)ctx->fs_pinned + p.at, p.src, p.bytes);
    CK(cudaMemcpyAsync(ctx->fs_stage.p, ctx->fs_pinned, total, cudaMemcpyHostToDevice, st));
    const char* B = ctx->fs_stage.as<char>();
    FleetDelta d;
    d.t = (const double*)(B + parts[0].at); d.reb_ddl = (const double*)(B + parts[1].at); d.idx = (const int*)(B + parts[2].at);
    d.nid = (const int*)(B + parts[3].at); d.load = (const int*)(B + parts[4].at); d.status = (const int*)(B + parts[5].at);
    d.reb_node = (const int*)(B + parts[6].at); d.off = (const int*)(B + parts[7].at); d.code = (const int*)(B + parts[8].at);
    d.onboard = (const uint8_t*)(B + parts[9].at); d.n = n;
    CK(cudaMemsetAsync(ctx->fs_err.p, 0, 8, st));
    k_state_scatter<<<std::max(1, std::min(ctx->n_sm * 4, (n + 255) / 256)), 256, 0, st>>>(ctx->fs, d, ctx->fs_err.as<unsigned long long>());
    unsigned long long err = 0;
    CK(cudaMemcpyAsync(&err, ctx->fs_err.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (err) return ctx->fail(AMOD_E_BADARG, "a vehicle index or schedule length is out of range for the resident state");
    return AMOD_OK;
}

// the pool build over the resident rows; the searched request ids are already in ctx->fs_search (device)
static int state_build_dev(amod_ctx* ctx, int32_t mode, int32_t cap, int64_t T, int32_t n_search, amod_stats* stats) {
    cudaStream_t st = ctx->stream;
    const FleetState& fs = ctx->fs;
    CK(ctx->partial.ensure(2 * 8 * (size_t)(ctx->n_sm + 1)));
    scan_dev(ctx, InStateLen{fs.len, (i64)fs.n_veh}, ctx->fs_off64.as<i64>(), FinNone{});
    k_state_csr<<<std::max(1, std::min(ctx->n_sm * 4, (fs.n_veh + 256) / 256)), 256, 0, st>>>(fs, ctx->fs_off64.as<i64>(), ctx->fs_off32.as<int>(),
                                                                                       ctx->fs_ccode.as<int>(), ctx->fs_conb.as<uint8_t>());
    amod_epoch ep; memset(&ep, 0, sizeof ep);
    ep.n_veh = fs.n_veh; ep.n_req = ctx->fs_req_n; ep.n_search = n_search; ep.cap = cap; ep.mode = mode; ep.T = T;
    ep.reserved = mode == AMOD_MODE_OSP ? 0 : std::max(ctx->fs_max_len, 1);     // longest schedule ever resident (monotone)
    ep.veh_nid = fs.nid; ep.veh_t_to_nid = fs.t; ep.veh_load = fs.load; ep.veh_status = fs.status;
    ep.veh_sche_off = ctx->fs_off32.as<int>(); ep.sche_code = ctx->fs_ccode.as<int>(); ep.sche_onboard = ctx->fs_conb.as<uint8_t>();
    ep.veh_reb_node = fs.reb_node; ep.veh_reb_ddl = fs.reb_ddl;
    ep.req_onid = ctx->rq_onid.as<int>(); ep.req_dnid = ctx->rq_dnid.as<int>(); ep.req_clp = ctx->rq_clp.as<double>(); ep.req_cld = ctx->rq_cld.as<double>();
    ep.req_tr = ctx->rq_tr.as<double>(); ep.req_ts = ctx->rq_ts.as<double>(); ep.search = ctx->fs_search.as<int>();
    return dispatch_tt(ctx, [&](auto tt) { return build_impl(ctx, &ep, AMOD_LOC_DEVICE, stats, tt); });
}

extern "C" int amod_state_build(amod_ctx* ctx, int32_t mode, int32_t cap, int64_t T, const int32_t* search_rids, int32_t n_search, amod_stats* stats) {
    if (!ctx || n_search < 0 || (n_search && !search_rids)) return AMOD_E_BADARG;
    if (!ctx->fs_valid) return ctx->fail(AMOD_E_BADARG, "amod_state_create first");
    for (int i = 0; i < n_search; i++)
        if (search_rids[i] < 0 || search_rids[i] >= ctx->fs_req_n) return ctx->fail(AMOD_E_BADARG, "search[%d] names an unknown request", i);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(ctx->fs_search.ensure(4 * (size_t)std::max(n_search, 1)));
    if (n_search) CK(cudaMemcpyAsync(ctx->fs_search.p, search_rids, 4 * (size_t)n_search, cudaMemcpyHostToDevice, st));
    return state_build_dev(ctx, mode, cap, T, n_search, stats);
}

extern "C" int amod_state_download(amod_ctx* ctx, int32_t* nid, double* t, int32_t* load, int32_t* status, int32_t* reb_node, double* reb_ddl,
                                   int32_t* sche_len, int32_t* sche_code, uint8_t* sche_onboard) {
    if (!ctx) return AMOD_E_BADARG;
    if (!ctx->fs_valid) return ctx->fail(AMOD_E_BADARG, "amod_state_create first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t V = (size_t)ctx->fs.n_veh, S = (size_t)ctx->fs.max_stops;
    if (V == 0) return AMOD_OK;
    CK(cudaMemcpyAsync(nid, ctx->fs.nid, 4 * V, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(t, ctx->fs.t, 8 * V, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(load, ctx->fs.load, 4 * V, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(status, ctx->fs.status, 4 * V, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(reb_node, ctx->fs.reb_node, 4 * V, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(reb_ddl, ctx->fs.reb_ddl, 8 * V, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sche_len, ctx->fs.len, 4 * V, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sche_code, ctx->fs.code, 4 * V * S, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sche_onboard, ctx->fs.onboard, V * S, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return AMOD_OK;
}

#include "fleet_api.cuh"

extern "C" int amod_gather_bench(amod_ctx* ctx, const int32_t* o_dev, const int32_t* d_dev, float* out_dev, int64_t n, int iters,
                                 double* ms) {
    if (!ctx || !ms) return AMOD_E_BADARG;
    return dispatch_tt(ctx, [&](auto tt) -> int {
        cudaStream_t st = ctx->stream;
        CK(cudaSetDevice(ctx->device));
        using T = decltype(tt);
        auto launch = [&]() {
            if constexpr (std::is_same<T, Table<double>>::value) k_gather_bench<double><<<ctx->n_sm * 16, 256, 0, st>>>(tt, o_dev, d_dev, out_dev, n);
            else k_gather_bench<float><<<ctx->n_sm * 16, 256, 0, st>>>(tt, o_dev, d_dev, out_dev, n);
        };
        launch();
        CK(cudaEventRecord(ctx->ev0, st));
        for (int i = 0; i < iters; i++) launch();
        CK(cudaEventRecord(ctx->ev1, st));
        CK(cudaStreamSynchronize(st));
        float f = 0.f;
        CK(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
        *ms = f / std::max(iters, 1);
        return AMOD_OK;
    });
}
