// gpool.cuh — multi-GPU assembly of the vehicle-major pool with peer stores over NVLink.
// Each rank scatters its own vehicles' entries into every rank's global pool (cudaIpc-mapped), at the
// positions given by a prefix sum over the all-gathered per-vehicle counts (dispatch_osp.py:107 order).
#pragma once
#include "common.cuh"
#include "pool.cuh"

#define GPOOL_MAX_WORLD 16

struct GPoolLayout {         // one contiguous allocation per rank; identical layout on every rank
    i64 cap_ent, cap_ptr, cap_pst;
    i64 o_flags, o_counts;   // peer-written: 2*GPOOL_MAX_WORLD epoch flags; all ranks' per-vehicle counts [world][vmax][3]
    i64 o_hdr, o_veh, o_kind, o_nfeas, o_cost, o_delay, o_toff, o_soff, o_treq, o_scode, bytes;
};
struct GPoolPeers { char* base[GPOOL_MAX_WORLD]; int world, rank; };

// per local vehicle: entries, trip requests, stops of the local (vehicle-major) pool
__global__ void k_veh_counts(const DEpoch* __restrict__ epp, const i64* __restrict__ veh_off, const Control* __restrict__ ctl,
                             PoolBufs pb, int* __restrict__ out) {
    const int n = epp->n_veh;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        const i64 e0 = veh_off[v], e1 = veh_off[v + 1];
        out[3 * v + 0] = (int)(e1 - e0);
        out[3 * v + 1] = (int)(pb.trip_off[e1] - pb.trip_off[e0]);
        out[3 * v + 2] = (int)(pb.sche_off[e1] - pb.sche_off[e0]);
    }
}

struct InGlobalVeh {         // counts in global vehicle order: vehicle v = local v / world of rank v % world
    const int* all; int world, vmax, comp, n_total;
    __device__ __forceinline__ i64 size() const { return n_total; }
    __device__ __forceinline__ i64 operator()(i64 v) const { return all[((v % world) * (i64)vmax + v / world) * 3 + comp]; }
};
struct FinStore { i64* dst; __device__ __forceinline__ void operator()(i64 tot) const { *dst = tot; } };

// one warp per local vehicle: copy its entries + payload to every rank's global pool
__global__ void __launch_bounds__(256)
k_gpool_scatter(const DEpoch* __restrict__ epp, const i64* __restrict__ veh_off, PoolBufs pb, GPoolLayout L, GPoolPeers peers,
                const i64* __restrict__ goff_e, const i64* __restrict__ goff_t, const i64* __restrict__ goff_s,
                const i64* __restrict__ totals /*[3]*/, int n_veh_total, unsigned long long* err) {
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const int n_local = epp->n_veh;
    const bool fits = totals[0] <= L.cap_ent && totals[1] <= L.cap_ptr && totals[2] <= L.cap_pst;
    if (!fits) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(err, 1ull); return; }
    if (blockIdx.x == 0 && threadIdx.x < 3) {    // header of OUR pool: global sizes (every rank computes the same totals)
        ((i64*)(peers.base[peers.rank] + L.o_hdr))[threadIdx.x] = totals[threadIdx.x];
    }
    for (int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < n_local; u += nwarps) {
        const int v = u * peers.world + peers.rank;
        const i64 e0 = veh_off[u], e1 = veh_off[u + 1];
        const i64 E0 = goff_e[v], T0 = goff_t[v], S0 = goff_s[v];
        const i64 t0 = pb.trip_off[e0], s0 = pb.sche_off[e0];
        const i64 nt = pb.trip_off[e1] - t0, ns = pb.sche_off[e1] - s0;
        const bool last = v == n_veh_total - 1;
        for (int d = 0; d < peers.world; d++) {
            char* B = peers.base[d];
            int* g_veh = (int*)(B + L.o_veh); int* g_kind = (int*)(B + L.o_kind); int* g_nfeas = (int*)(B + L.o_nfeas);
            double* g_cost = (double*)(B + L.o_cost); double* g_delay = (double*)(B + L.o_delay);
            i64* g_toff = (i64*)(B + L.o_toff); i64* g_soff = (i64*)(B + L.o_soff);
            int* g_treq = (int*)(B + L.o_treq); int* g_scode = (int*)(B + L.o_scode);
            for (i64 e = e0 + lane; e < e1; e += 32) {
                const i64 E = E0 + (e - e0);
                g_veh[E] = v; g_kind[E] = pb.ent_kind[e]; g_nfeas[E] = pb.ent_nfeas[e];
                g_cost[E] = pb.ent_cost[e]; g_delay[E] = pb.ent_delay[e];
                g_toff[E] = T0 + (pb.trip_off[e] - t0); g_soff[E] = S0 + (pb.sche_off[e] - s0);
            }
            for (i64 x = lane; x < nt; x += 32) g_treq[T0 + x] = pb.trip_req[t0 + x];
            for (i64 x = lane; x < ns; x += 32) g_scode[S0 + x] = pb.sche_code[s0 + x];
            if (last && lane == 0) { g_toff[totals[0]] = totals[1]; g_soff[totals[0]] = totals[2]; }   // CSR end sentinels
        }
    }
    __threadfence_system();      // peer stores are performed before the kernel (and the flag that follows it) completes
}


// ---- assembly without host collectives: flags in peer memory ---------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Device-resident state of the peer assembly (one per context): the epoch counter lives on the device so that the
// assembly kernels can be part of the captured epoch graph (no per-epoch kernel argument).
struct GpState {
    i64 tot[3];                     // entries, trip requests, stops of the assembled pool
    unsigned long long err;         // 1 = capacity, 2 = a rank never arrived (time-out)
    unsigned ticket[2];
    unsigned long long epoch;       // assemblies completed (pushes issued) by this rank
    long long timeout_clk;          // bounded spins give up after this many clock64 ticks
    unsigned long long read_base;   // epochs up to this one were read under another root / mode: their "read" flags are not waited for
};

// Wait until the flags `which` of the ranks in `mask` in MY pool have reached `epoch` (bounded spin); call with a whole
// block, threads < world poll one flag each.  which 0 = "slice pushed", 1 = "slices read".
__device__ __forceinline__ void gp_wait_flags(const char* __restrict__ my_base, const GPoolLayout& L, int world, int which,
                                              unsigned long long epoch, GpState* st, unsigned mask) {
    if ((int)threadIdx.x < world && ((mask >> threadIdx.x) & 1u)) {
        const unsigned long long* f = (const unsigned long long*)(my_base + L.o_flags) + which * GPOOL_MAX_WORLD + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < epoch) {
            if (clock64() - t0 > st->timeout_clk) { atomicOr(&st->err, 2ull); break; }   // a rank never arrived
            __nanosleep(200);
        }
        __threadfence_system();
    }
    __syncthreads();
}

// this rank's build must be complete and final for its slice to travel: an overflowed attempt or one that has to be
// searched deeper is rerun by the host (amod_b200.cu: build_impl), and only the rerun pushes
__device__ __forceinline__ bool gp_build_final(const Control* __restrict__ ctl, int run_levels, int max_level) {
    return !ctl->overflow && !(run_levels < max_level && ctl->n_trips[run_levels] > 0);
}


// ---- bulk variant: whole rank slices travel as coalesced peer stores, the interleave is done locally --------
// Every rank has one staging slot per peer.  A rank pushes its local pool arrays (contiguous, 16 B stores) into
// its slot on every rank, raises a flag, waits for all slots in its own memory, prefix-sums the per-vehicle
// counts in global vehicle order and merges the staged slices into its own vehicle-major pool.  Peer traffic is
// bulk; the small scattered copies are local.
struct GSlotLayout {        // one staging slot (offsets inside the slot); identical on every rank
    i64 cap_veh, cap_ent, cap_ptr, cap_pst;
    i64 o_hdr, o_vehoff, o_kind, o_nfeas, o_cost, o_delay, o_toff, o_soff, o_treq, o_scode, bytes;
    i64 o_cnt;            // int32 [3][cap_veh]: entries / trip requests / stops of each local vehicle (so readers need no dependent loads)
    i64 o_veh;            // int32 [cap_ent]: local vehicle of every entry (the SBA merge needs it; OSP slices are vehicle-major)
};
struct GSeg { const char* src; i64 dst_off; i64 bytes; };
struct GPush { GSeg seg[11]; int n; };

// root < 0: every rank receives every slice (all-gather); root >= 0: only `root` does (gather: the rank that hosts the
// assigner), so a rank sends its slice once instead of `world` times and only the root merges.
__global__ void __launch_bounds__(256)
k_gp_push_bulk(GPush P, GPoolLayout L, GSlotLayout S, i64 o_stage, GPoolPeers peers, const DEpoch* __restrict__ epp,
               const Control* __restrict__ ctl, GpState* st, const char* __restrict__ my_base, int root, int run_levels, int max_level) {
    if (!gp_build_final(ctl, run_levels, max_level)) return;
    const unsigned long long epoch = st->epoch + 1;        // (every block reads it before the last block to finish advances it)
    const int d0 = root < 0 ? 0 : root, d1 = root < 0 ? peers.world : root + 1;
    // the readers of my slot must have finished with the previous epoch before I overwrite it
    if (epoch - 1 > st->read_base) gp_wait_flags(my_base, L, peers.world, 1, epoch - 1, st, root < 0 ? 0xffffffffu : (1u << root));
    const i64 n_veh = epp->n_veh, n_ent = ctl->n_ent, n_ptr = ctl->n_ptr, n_pst = ctl->n_pst;
    const bool fits = n_veh <= S.cap_veh && n_ent <= S.cap_ent && n_ptr <= S.cap_ptr && n_pst <= S.cap_pst;
    if (!fits && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&st->err, 1ull);
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x, nth = (i64)gridDim.x * blockDim.x;
    if (!fits && tid == 0)      // tell every reader that this slot holds nothing usable
        for (int d = d0; d < d1; d++) ((i64*)(peers.base[d] + o_stage + (i64)peers.rank * S.bytes + S.o_hdr))[0] = -1;
    if (fits) {
        // live sizes of the segments (the host passes capacities; trim to what this epoch produced)
        const i64 live[11] = {32, 8 * (n_veh + 1), 4 * n_ent, 4 * n_ent, 8 * n_ent, 8 * n_ent, 8 * (n_ent + 1), 8 * (n_ent + 1),
                              4 * n_ptr, 4 * n_pst, 4 * n_ent};
        for (int d = d0; d < d1; d++) {
            char* slot = peers.base[d] + o_stage + (i64)peers.rank * S.bytes;
            if (tid == 0) { i64* h = (i64*)(slot + S.o_hdr); h[0] = n_veh; h[1] = n_ent; h[2] = n_ptr; h[3] = n_pst; h[4] = epp->mode; }
            for (int g = 1; g < P.n; g++) {
                const i64 bytes = live[g];
                const int4* src = (const int4*)P.seg[g].src;
                int4* dst = (int4*)(slot + P.seg[g].dst_off);
                const i64 n16 = bytes >> 4;
                for (i64 x = tid; x < n16; x += nth) dst[x] = src[x];
                const i64 done = n16 << 4;
                for (i64 x = done + tid * 4; x < bytes; x += nth * 4) *(int*)((char*)dst + x) = *(const int*)((const char*)src + x);
            }
        }
        // per-vehicle (entries, trip requests, stops), so that the readers' prefix sums are plain array reads
        const i64* vo = (const i64*)P.seg[1].src; const i64* toff = (const i64*)P.seg[6].src; const i64* soff = (const i64*)P.seg[7].src;
        for (i64 u = tid; u < n_veh; u += nth) {
            const i64 e0 = vo[u], e1 = vo[u + 1];
            const int ce = (int)(e1 - e0), ct = (int)(toff[e1] - toff[e0]), cs = (int)(soff[e1] - soff[e0]);
            for (int d = d0; d < d1; d++) {
                int* cnt = (int*)(peers.base[d] + o_stage + (i64)peers.rank * S.bytes + S.o_cnt);
                cnt[u] = ce; cnt[S.cap_veh + u] = ct; cnt[2 * S.cap_veh + u] = cs;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&st->ticket[0], 1u) == gridDim.x - 1) {
        st->ticket[0] = 0;
        st->epoch = epoch;
        __threadfence_system();
        for (int d = d0; d < d1; d++)
            st_release_sys((unsigned long long*)(peers.base[d] + L.o_flags) + peers.rank, epoch);
    }
}

struct InStagedVeh {        // counts of global vehicle v read from the staged slices: vehicle v is local vehicle vlocal[v] of rank vrank[v]
    const char* stage; GSlotLayout S; const int* vrank; const int* vlocal; int comp, n_total;
    __device__ __forceinline__ i64 size() const { return n_total; }
    __device__ __forceinline__ i64 operator()(i64 v) const {
        const char* slot = stage + (i64)vrank[v] * S.bytes;
        const i64 u = vlocal[v];
        if (((const i64*)(slot + S.o_hdr))[0] <= u) return 0;        // slot unusable (its rank overflowed) or short
        return ((const int*)(slot + S.o_cnt))[comp * S.cap_veh + u];
    }
};

// Staged slices -> my vehicle-major pool (all local memory).  One warp per 32 consecutive GLOBAL entries: a lane
// finds its entry's vehicle by binary search in the global entry offsets, copies the fixed fields and rebases the
// two CSR offsets; the trip requests and stop codes of the 32 entries are one contiguous destination range, copied
// by the whole warp (each element finds its entry among the lanes by a shuffle search).  No per-vehicle loops: a
// vehicle with thousands of entries costs what its entries cost.
__global__ void __launch_bounds__(256)
k_gp_merge(const char* __restrict__ stage, GSlotLayout S, char* __restrict__ B, GPoolLayout L, int world,
           const i64* __restrict__ goff_e, const i64* __restrict__ goff_t, const i64* __restrict__ goff_s,
           GpState* st, int n_veh_total, GPoolPeers peers, const int* __restrict__ vrank, const int* __restrict__ vlocal,
           const Control* __restrict__ ctl, int run_levels, int max_level) {
    if (!gp_build_final(ctl, run_levels, max_level)) return;
    const i64* totals = st->tot;
    const unsigned long long epoch = st->epoch;
    const int lane = threadIdx.x & 31;
    const i64 nwarps = (i64)gridDim.x * (blockDim.x >> 5);
    bool usable = totals[0] <= L.cap_ent && totals[1] <= L.cap_ptr && totals[2] <= L.cap_pst && !(st->err & 2ull);
    for (int r = 0; r < world; r++) usable &= ((const i64*)(stage + (i64)r * S.bytes + S.o_hdr))[0] >= 0;
    if (!usable && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&st->err, 1ull);
    int* g_veh = (int*)(B + L.o_veh); int* g_kind = (int*)(B + L.o_kind); int* g_nfeas = (int*)(B + L.o_nfeas);
    double* g_cost = (double*)(B + L.o_cost); double* g_delay = (double*)(B + L.o_delay);
    i64* g_toff = (i64*)(B + L.o_toff); i64* g_soff = (i64*)(B + L.o_soff);
    int* g_treq = (int*)(B + L.o_treq); int* g_scode = (int*)(B + L.o_scode);
    const i64 n_ent = usable ? totals[0] : 0;
    if (blockIdx.x == 0 && threadIdx.x < 3) ((i64*)(B + L.o_hdr))[threadIdx.x] = totals[threadIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0) { g_toff[n_ent] = totals[1]; g_soff[n_ent] = totals[2]; }
    for (i64 E0 = ((i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; E0 < n_ent; E0 += nwarps * 32) {
        const i64 E = E0 + lane;
        const bool valid = E < n_ent;
        i64 dt = 0, ds = 0;                 // destination of the entry's trip requests / stop codes
        int lt = 0, lsc = 0;                // their lengths
        const int* pt = nullptr; const int* ps = nullptr;
        if (valid) {
            int a = 0, b = n_veh_total;     // last vehicle whose first entry is <= E (vehicles without entries share their successor's offset)
            while (b - a > 1) { const int m = (a + b) >> 1; if (goff_e[m] <= E) a = m; else b = m; }
            const int v = a;
            const char* slot = stage + (i64)vrank[v] * S.bytes;
            const i64 u = vlocal[v];
            const i64* s_toff = (const i64*)(slot + S.o_toff); const i64* s_soff = (const i64*)(slot + S.o_soff);
            const i64 e0 = ((const i64*)(slot + S.o_vehoff))[u];
            const i64 e = e0 + (E - goff_e[v]);
            const i64 t0 = s_toff[e0], s0 = s_soff[e0], te = s_toff[e], se = s_soff[e];
            lt = (int)(s_toff[e + 1] - te); lsc = (int)(s_soff[e + 1] - se);
            dt = goff_t[v] + (te - t0); ds = goff_s[v] + (se - s0);
            pt = (const int*)(slot + S.o_treq) + te; ps = (const int*)(slot + S.o_scode) + se;
            g_veh[E] = v; g_kind[E] = ((const int*)(slot + S.o_kind))[e]; g_nfeas[E] = ((const int*)(slot + S.o_nfeas))[e];
            g_cost[E] = ((const double*)(slot + S.o_cost))[e]; g_delay[E] = ((const double*)(slot + S.o_delay))[e];
            g_toff[E] = dt; g_soff[E] = ds;
        }
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            const i64 d0 = pass ? ds : dt;
            const int len = pass ? lsc : lt;
            const int* src = pass ? ps : pt;
            int* dst = pass ? g_scode : g_treq;
            const i64 first = __shfl_sync(FULL, d0, 0);
            const int n_valid = __popc(__ballot_sync(FULL, valid));
            const i64 end = __shfl_sync(FULL, d0 + len, n_valid - 1);
            const int rel = valid ? (int)(d0 - first) : 0x7fffffff;          // monotone over the lanes
            const int n = (int)(end - first);
            for (int x0 = 0; x0 < n; x0 += 32) {
                const int x = x0 + lane;
                int own = 0;                                                  // last lane whose range starts at or before x
#pragma unroll
                for (int h = 16; h; h >>= 1) { const int r = __shfl_sync(FULL, rel, own + h); if (r <= x) own += h; }
                const int* sp = (const int*)__shfl_sync(FULL, (unsigned long long)src, own);
                const int r0 = __shfl_sync(FULL, rel, own);
                if (x < n) dst[first + x] = sp[x - r0];
            }
        }
    }
    // "I have read everyone's slot of this epoch": the last block to finish raises my flag on every rank
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&st->ticket[1], 1u) == gridDim.x - 1) {
            st->ticket[1] = 0;
            __threadfence_system();
            for (int d = 0; d < peers.world; d++)
                st_release_sys((unsigned long long*)(peers.base[d] + L.o_flags) + GPOOL_MAX_WORLD + peers.rank, epoch);
        }
    }
}

// all three prefix sums (entries, trip requests, stops) over the global vehicle order in ONE single-block launch
// (a fleet has at most a few 10^4 vehicles): thread = contiguous range of vehicles; the counts are plain arrays
// every rank pushed with its slice
__global__ void __launch_bounds__(1024)
k_gp_offsets(const char* __restrict__ stage, GSlotLayout S, int world, int n_total, i64* __restrict__ goff_e, i64* __restrict__ goff_t,
             i64* __restrict__ goff_s, GpState* st, const char* __restrict__ my_base, GPoolLayout L, const int* __restrict__ vrank,
             const int* __restrict__ vlocal, const Control* __restrict__ ctl, int run_levels, int max_level) {
    __shared__ i64 sm[3][32];
    if (!gp_build_final(ctl, run_levels, max_level)) return;
    gp_wait_flags(my_base, L, world, 0, st->epoch, st, 0xffffffffu);       // every rank's slice of this epoch has landed in my memory
    // pass 1 (coalesced over vehicles, loads independent of each other): the three counts of every global vehicle, parked in the
    // offset arrays themselves
    for (int v = threadIdx.x; v < n_total; v += 1024) {
        const char* slot = stage + (i64)vrank[v] * S.bytes;
        const i64 u = vlocal[v];
        const bool ok = ((const i64*)(slot + S.o_hdr))[0] > u;             // (a slot is unusable when its rank overflowed)
        const int* cnt = (const int*)(slot + S.o_cnt);
        goff_e[v] = ok ? cnt[u] : 0; goff_t[v] = ok ? cnt[S.cap_veh + u] : 0; goff_s[v] = ok ? cnt[2 * S.cap_veh + u] : 0;
    }
    __syncthreads();
    // pass 2: every thread owns a contiguous range of vehicles; block-wide exclusive scan of the three running sums
    const int per = (n_total + 1023) / 1024;
    const int v0 = min(n_total, (int)threadIdx.x * per), v1 = min(n_total, v0 + per);
    i64 a[3] = {0, 0, 0};
    for (int v = v0; v < v1; v++) { a[0] += goff_e[v]; a[1] += goff_t[v]; a[2] += goff_s[v]; }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    i64 ex[3];
    for (int c = 0; c < 3; c++) {
        i64 y = a[c];
        for (int o = 1; o < 32; o <<= 1) { i64 t = __shfl_up_sync(FULL, y, o); if (lane >= o) y += t; }
        if (lane == 31) sm[c][w] = y;
        ex[c] = y - a[c];
    }
    __syncthreads();
    if (w == 0)
        for (int c = 0; c < 3; c++) {
            i64 z = sm[c][lane], y = z;
            for (int o = 1; o < 32; o <<= 1) { i64 t = __shfl_up_sync(FULL, y, o); if (lane >= o) y += t; }
            sm[c][lane] = y - z;
            if (lane == 31) st->tot[c] = y;
        }
    __syncthreads();
    i64 run[3] = {ex[0] + sm[0][w], ex[1] + sm[1][w], ex[2] + sm[2][w]};
    for (int v = v0; v < v1; v++) {
        const i64 ce = goff_e[v], ct = goff_t[v], cs = goff_s[v];
        goff_e[v] = run[0]; goff_t[v] = run[1]; goff_s[v] = run[2];
        run[0] += ce; run[1] += ct; run[2] += cs;
    }
    __syncthreads();
    if (threadIdx.x == 1023) { goff_e[n_total] = st->tot[0]; goff_t[n_total] = st->tot[1]; goff_s[n_total] = st->tot[2]; }
}
