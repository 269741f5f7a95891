// sbam.cuh — multi-GPU SBA: the request-major merge of the ranks' slices on the gathering rank.
//
// Every rank builds the SBA pairs of ITS vehicles for all new requests (dispatch_sba.py:56-65) and pushes its slice like an
// OSP slice (gpool.cuh).  The reference's order is request-major in new_received_rids order, vehicles in fleet order within
// a request, then one empty pair per vehicle in fleet order (:68-69).  A rank's vehicles keep their fleet order, so every
// slice is already sorted by (request position, global vehicle); the position of an entry in the merged pool is
//     (pairs of earlier requests, all ranks) + (pairs of the same request with a smaller global vehicle, all ranks)
// and for an empty pair (all pairs) + its global vehicle: counting, no sort.
#pragma once
#undef BFILE
#define BFILE 7
#include "common.cuh"
#include "gpool.cuh"

struct SbaMergeBufs {
    int* pos_of;        // [n_req] search position of a request index (n_search = not searched)
    int* epos;          // [world * cap_ent] request position of every staged entry (n_search for the empty pairs)
    int* egv;           // [world * cap_ent] global vehicle of every staged entry
    int* dest;          // [world * cap_ent] position in the merged pool
    int* rs;            // [world * (n_search + 2)] first entry of slice r with request position >= s
    i64* base;          // [n_search + 2] first merged entry with request position >= s
    int* slen;          // [cap_ent_global] schedule length of every merged entry
    const int* glob;    // [world * cap_veh] global vehicle of (rank, local vehicle)
    i64 cap_ent_slot;
    int cap_veh;
};

__global__ void k_sbam_pos(const DEpoch* __restrict__ epp, SbaMergeBufs m) {
    const DEpoch& ep = *epp;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ep.n_req; i += stride) m.pos_of[i] = ep.n_search;
}
__global__ void k_sbam_pos2(const DEpoch* __restrict__ epp, SbaMergeBufs m) {
    const DEpoch& ep = *epp;
    const int stride = gridDim.x * blockDim.x;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < ep.n_search; s += stride) atomicMin(&m.pos_of[ep.search[s]], s);
}

// entries of slice r (header: n_veh, n_ent, n_ptr, n_pst, mode)
__device__ __forceinline__ i64 slot_n_ent(const char* stage, const GSlotLayout& S, int r) { return ((const i64*)(stage + (i64)r * S.bytes + S.o_hdr))[1]; }

__global__ void k_sbam_keys(const char* __restrict__ stage, GSlotLayout S, int world, const DEpoch* __restrict__ epp, SbaMergeBufs m,
                            GpState* st, const char* __restrict__ my_base, GPoolLayout L, const Control* __restrict__ ctl) {
    if (ctl->overflow) return;
    gp_wait_flags(my_base, L, world, 0, st->epoch, st, 0xffffffffu);       // every rank's slice of this epoch has landed
    const int Q = epp->n_search;
    for (int r = 0; r < world; r++) {
        const char* slot = stage + (i64)r * S.bytes;
        const i64 n = min(slot_n_ent(stage, S, r), S.cap_ent);
        const int* kind = (const int*)(slot + S.o_kind); const int* veh = (const int*)(slot + S.o_veh);
        const i64* toff = (const i64*)(slot + S.o_toff); const int* treq = (const int*)(slot + S.o_treq);
        for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (i64)gridDim.x * blockDim.x) {
            const bool pair = kind[e] == AMOD_KIND_SBA_PAIR;
            m.epos[(i64)r * m.cap_ent_slot + e] = pair ? m.pos_of[treq[toff[e]]] : Q;
            m.egv[(i64)r * m.cap_ent_slot + e] = m.glob[(i64)r * m.cap_veh + veh[e]];
        }
    }
}

__device__ __forceinline__ int lower_bound_i(const int* __restrict__ a, int lo, int hi, int key) {   // first index in [lo, hi) with a[i] >= key
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void k_sbam_reqstart(const char* __restrict__ stage, GSlotLayout S, int world, const DEpoch* __restrict__ epp, SbaMergeBufs m,
                                const Control* __restrict__ ctl) {
    if (ctl->overflow) return;
    const int Q = epp->n_search;
    const int n_items = world * (Q + 2);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
        const int r = i / (Q + 2), s = i - r * (Q + 2);
        const int n = (int)min(slot_n_ent(stage, S, r), S.cap_ent);
        m.rs[i] = s <= Q ? lower_bound_i(m.epos + (i64)r * m.cap_ent_slot, 0, n, s) : n;
    }
}

__global__ void k_sbam_base(int world, const DEpoch* __restrict__ epp, SbaMergeBufs m, GpState* st, GPoolLayout L, int n_veh_total,
                            const Control* __restrict__ ctl) {
    if (ctl->overflow) return;
    const int Q = epp->n_search;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s <= Q + 1; s += gridDim.x * blockDim.x) {
        i64 b = 0;
        for (int r = 0; r < world; r++) b += m.rs[r * (Q + 2) + s];
        m.base[s] = b;
        if (s == Q) {     // all pairs; the merged pool = pairs + one empty pair per vehicle of the fleet
            st->tot[0] = b + n_veh_total; st->tot[1] = b;
            if (b + n_veh_total > L.cap_ent || b > L.cap_ptr) atomicOr(&st->err, 1ull);
        }
    }
}

__global__ void __launch_bounds__(256)
k_sbam_dest(const char* __restrict__ stage, GSlotLayout S, char* __restrict__ B, GPoolLayout L, int world, const DEpoch* __restrict__ epp,
            SbaMergeBufs m, GpState* st, const Control* __restrict__ ctl) {
    if (ctl->overflow || (st->err & 3ull)) return;
    const int Q = epp->n_search;
    int* g_veh = (int*)(B + L.o_veh); int* g_kind = (int*)(B + L.o_kind); int* g_nfeas = (int*)(B + L.o_nfeas);
    double* g_cost = (double*)(B + L.o_cost); double* g_delay = (double*)(B + L.o_delay);
    i64* g_toff = (i64*)(B + L.o_toff); int* g_treq = (int*)(B + L.o_treq);
    const i64 n_pairs = m.base[Q];
    for (int r = 0; r < world; r++) {
        const char* slot = stage + (i64)r * S.bytes;
        const i64 n = min(slot_n_ent(stage, S, r), S.cap_ent);
        const i64* toff = (const i64*)(slot + S.o_toff); const i64* soff = (const i64*)(slot + S.o_soff);
        const int* treq = (const int*)(slot + S.o_treq);
        for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (i64)gridDim.x * blockDim.x) {
            const int s = m.epos[(i64)r * m.cap_ent_slot + e], gv = m.egv[(i64)r * m.cap_ent_slot + e];
            i64 d;
            if (s < Q) {
                d = m.base[s];
                for (int r2 = 0; r2 < world; r2++) {
                    const int a = m.rs[r2 * (Q + 2) + s], b = m.rs[r2 * (Q + 2) + s + 1];
                    d += r2 == r ? (int)e - a : lower_bound_i(m.egv + (i64)r2 * m.cap_ent_slot, a, b, gv) - a;
                }
                g_treq[d] = treq[toff[e]];
            } else d = n_pairs + gv;
            m.dest[(i64)r * m.cap_ent_slot + e] = (int)d;
            g_veh[d] = gv; g_kind[d] = ((const int*)(slot + S.o_kind))[e]; g_nfeas[d] = ((const int*)(slot + S.o_nfeas))[e];
            g_cost[d] = ((const double*)(slot + S.o_cost))[e]; g_delay[d] = ((const double*)(slot + S.o_delay))[e];
            g_toff[d] = d < n_pairs ? d : n_pairs;
            m.slen[d] = (int)(soff[e + 1] - soff[e]);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) g_toff[st->tot[0]] = n_pairs;
}

struct InSbaSlen {
    const int* slen; const GpState* st;
    __device__ __forceinline__ i64 size() const { return (st->err & 3ull) ? 0 : st->tot[0]; }
    __device__ __forceinline__ i64 operator()(i64 i) const { return (i64)slen[i]; }
};
struct FinSbaStops {
    GpState* st; i64 cap;
    __device__ __forceinline__ void operator()(i64 tot) const { st->tot[2] = tot; if (tot > cap) atomicOr(&st->err, 1ull); }
};

__global__ void __launch_bounds__(256)
k_sbam_codes(const char* __restrict__ stage, GSlotLayout S, char* __restrict__ B, GPoolLayout L, int world, SbaMergeBufs m, GpState* st,
             GPoolPeers peers, const Control* __restrict__ ctl) {
    const bool run = !ctl->overflow && !(st->err & 3ull);
    const i64* g_soff = (const i64*)(B + L.o_soff); int* g_scode = (int*)(B + L.o_scode);
    if (run) {
        if (blockIdx.x == 0 && threadIdx.x < 3) ((i64*)(B + L.o_hdr))[threadIdx.x] = st->tot[threadIdx.x];
        for (int r = 0; r < world; r++) {
            const char* slot = stage + (i64)r * S.bytes;
            const i64 n = min(slot_n_ent(stage, S, r), S.cap_ent);
            const i64* soff = (const i64*)(slot + S.o_soff); const int* scode = (const int*)(slot + S.o_scode);
            for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (i64)gridDim.x * blockDim.x) {
                const i64 d = m.dest[(i64)r * m.cap_ent_slot + e];
                const i64 a = soff[e], len = soff[e + 1] - a, o = g_soff[d];
                for (i64 k = 0; k < len; k++) g_scode[o + k] = scode[a + k];
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
                st_release_sys((unsigned long long*)(peers.base[d] + L.o_flags) + GPOOL_MAX_WORLD + peers.rank, st->epoch);
        }
    }
}
