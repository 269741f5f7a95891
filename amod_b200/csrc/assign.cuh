// assign.cuh — SURVEY.md §8f rank 1: greedy_assignment (ilp_assign.py:101-130) on the device-resident pool.
// The reference stably sorts the pairs by -score (score = len(trip) after its scorer, scheduling.py:160-161)
// and keeps a pair iff its vehicle and all its requests are still free.  With the total order
// (rank in the sorted list) this equals: a pair is kept iff no kept pair of smaller rank shares a resource.
// Rounds: every undecided pair that is the smallest-rank undecided pair on its vehicle and on each of its
// requests is kept (all of them at once); pairs touching a taken resource are dropped; repeat.  A round is two launches
// (k_ga_keep, k_ga_drop): the survivors post their ranks for the next round while they are being counted.
#pragma once
#include "common.cuh"
#include "pool.cuh"

#define GA_BUCKETS 32
#define GA_UNDECIDED 0
#define GA_KEPT 1
#define GA_DROPPED 2

// stable counting sort by descending score: scan over [bucket][entry] flags; nb = highest score of the pool + 1 buckets
// (Control::max_tlen, left by k_pool_meta), at most GA_BUCKETS
struct InScoreBucket {
    const int* tlen; const Control* ctl; int nb;
    __device__ __forceinline__ i64 size() const { return (i64)nb * ctl->n_ent; }
    __device__ __forceinline__ i64 operator()(i64 i) const {
        const i64 P = ctl->n_ent;
        const int b = (int)(i / P);
        return tlen[i - (i64)b * P] == nb - 1 - b;
    }
};

__global__ void k_ga_init(const Control* __restrict__ ctl, const int* __restrict__ tlen, const i64* __restrict__ pos, int* __restrict__ rank,
                          int* __restrict__ order, uint8_t* __restrict__ state, int* best_v, int n_veh, int* best_r, int n_req,
                          uint8_t* taken_v, uint8_t* taken_r, unsigned long long* err, int nb) {
    const i64 P = ctl->n_ent;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < P; e += stride) {
        const int s = tlen[e];
        if (s < 0 || s >= nb) { atomicOr(err, 1ull); continue; }
        const int r = (int)pos[(i64)(nb - 1 - s) * P + e];
        rank[e] = r; order[r] = (int)e; state[e] = GA_UNDECIDED;
    }
    // (both sets of per-resource minima: best_v / best_r hold 2 x n entries)
    for (i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x; x < n_veh; x += stride) { best_v[x] = 0x7fffffff; best_v[n_veh + x] = 0x7fffffff; taken_v[x] = 0; }
    for (i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x; x < n_req; x += stride) { best_r[x] = 0x7fffffff; best_r[n_req + x] = 0x7fffffff; taken_r[x] = 0; }
}

__global__ void k_ga_min(const Control* __restrict__ ctl, PoolBufs pb, const int* __restrict__ rank, const uint8_t* __restrict__ state,
                         int* best_v, int* best_r) {
    const i64 P = ctl->n_ent;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < P; e += stride) {
        if (state[e] != GA_UNDECIDED) continue;
        const int r = rank[e];
        atomicMin(&best_v[pb.ent_veh[e]], r);
        for (i64 t = pb.trip_off[e]; t < pb.trip_off[e + 1]; t++) atomicMin(&best_r[pb.trip_req[t]], r);
    }
}

__global__ void k_ga_keep(const Control* __restrict__ ctl, PoolBufs pb, const int* __restrict__ rank, uint8_t* __restrict__ state,
                          const int* __restrict__ best_v, const int* __restrict__ best_r, uint8_t* taken_v, uint8_t* taken_r) {
    const i64 P = ctl->n_ent;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < P; e += stride) {
        if (state[e] != GA_UNDECIDED) continue;
        const int r = rank[e];
        bool top = best_v[pb.ent_veh[e]] == r;
        for (i64 t = pb.trip_off[e]; t < pb.trip_off[e + 1] && top; t++) top = best_r[pb.trip_req[t]] == r;
        if (!top) continue;
        state[e] = GA_KEPT;
        taken_v[pb.ent_veh[e]] = 1;
        for (i64 t = pb.trip_off[e]; t < pb.trip_off[e + 1]; t++) taken_r[pb.trip_req[t]] = 1;
    }
}

// drop pairs touching a taken resource and count what is left; the survivors already post their rank to the NEXT round's
// per-resource minima (two sets of minima alternate), and this round's set is reset for the round after: two launches per round
__global__ void k_ga_drop(const Control* __restrict__ ctl, PoolBufs pb, const int* __restrict__ rank, uint8_t* __restrict__ state,
                          const uint8_t* __restrict__ taken_v, const uint8_t* __restrict__ taken_r, int* best_v, int n_veh, int* best_r, int n_req,
                          int* next_v, int* next_r, int* n_left) {
    const i64 P = ctl->n_ent;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    int left = 0;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < P; e += stride) {
        if (state[e] != GA_UNDECIDED) continue;
        bool hit = taken_v[pb.ent_veh[e]];
        for (i64 t = pb.trip_off[e]; t < pb.trip_off[e + 1] && !hit; t++) hit = taken_r[pb.trip_req[t]];
        if (hit) { state[e] = GA_DROPPED; continue; }
        left++;
        const int r = rank[e];
        atomicMin(&next_v[pb.ent_veh[e]], r);
        for (i64 t = pb.trip_off[e]; t < pb.trip_off[e + 1]; t++) atomicMin(&next_r[pb.trip_req[t]], r);
    }
    for (i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x; x < n_veh; x += stride) best_v[x] = 0x7fffffff;
    for (i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x; x < n_req; x += stride) best_r[x] = 0x7fffffff;
    left = warp_sum(left);
    if ((threadIdx.x & 31) == 0 && left) atomicAdd(n_left, left);
}

// All rounds in one cooperative launch (every block resident): the keep and drop phases of a round as above, separated by a
// device-wide barrier instead of a kernel boundary, the count of undecided pairs read on the device - one host synchronisation for the
// whole assignment.  An entry is handled by the same thread in every phase; what other threads change (the per-resource minima,
// the taken flags, the counts) is read from L2.
__global__ void __launch_bounds__(256)
k_ga_rounds(const Control* __restrict__ ctl, PoolBufs pb, const int* __restrict__ rank, uint8_t* __restrict__ state,
            int* bv0, int* br0, int* bv1, int* br1, int n_veh, int n_req, uint8_t* taken_v, uint8_t* taken_r,
            int* n_left /* [2] */, unsigned* bar, int* err, int max_rounds, int* rounds_out) {
    const i64 P = ctl->n_ent;
    const i64 stride = (i64)gridDim.x * blockDim.x, t0 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned phase = 0;
    int round = 0;
    while (round < max_rounds) {
        const int cur = round & 1;
        const int* best_v = cur ? bv1 : bv0; const int* best_r = cur ? br1 : br0;
        int* next_v = cur ? bv0 : bv1; int* next_r = cur ? br0 : br1;
        for (i64 e = t0; e < P; e += stride) {        // keep
            if (state[e] != GA_UNDECIDED) continue;
            const int r = rank[e], v = pb.ent_veh[e];
            const i64 a = pb.trip_off[e], b = pb.trip_off[e + 1];
            bool top = ldcg_i(best_v + v) == r;
            for (i64 t = a; t < b && top; t++) top = ldcg_i(best_r + pb.trip_req[t]) == r;
            if (!top) continue;
            state[e] = GA_KEPT;
            taken_v[v] = 1;
            for (i64 t = a; t < b; t++) taken_r[pb.trip_req[t]] = 1;
        }
        if (!grid_barrier(bar, ++phase * gridDim.x, err)) break;
        if (t0 == 0) n_left[cur ^ 1] = 0;             // the other round's count: every block has read it (they all passed this barrier)
        int left = 0;
        for (i64 e = t0; e < P; e += stride) {        // drop, count, post the survivors' ranks for the next round
            if (state[e] != GA_UNDECIDED) continue;
            const int v = pb.ent_veh[e];
            const i64 a = pb.trip_off[e], b = pb.trip_off[e + 1];
            bool hit = ldcg_b(taken_v + v) != 0;
            for (i64 t = a; t < b && !hit; t++) hit = ldcg_b(taken_r + pb.trip_req[t]) != 0;
            if (hit) { state[e] = GA_DROPPED; continue; }
            left++;
            const int r = rank[e];
            atomicMin(&next_v[v], r);
            for (i64 t = a; t < b; t++) atomicMin(&next_r[pb.trip_req[t]], r);
        }
        for (i64 x = t0; x < n_veh; x += stride) ((int*)best_v)[x] = 0x7fffffff;
        for (i64 x = t0; x < n_req; x += stride) ((int*)best_r)[x] = 0x7fffffff;
        left = warp_sum(left);
        if ((threadIdx.x & 31) == 0 && left) atomicAdd(&n_left[cur], left);
        round++;
        if (!grid_barrier(bar, ++phase * gridDim.x, err)) break;
        if (ldcg_i(n_left + cur) == 0) break;
    }
    if (t0 == 0) *rounds_out = round;
}

struct InKeptByRank {       // scan input over sorted positions: 1 if the pair at that position was kept
    const int* order; const uint8_t* state; const Control* ctl;
    __device__ __forceinline__ i64 size() const { return ctl->n_ent; }
    __device__ __forceinline__ i64 operator()(i64 i) const { return state[order[i]] == GA_KEPT; }
};

__global__ void k_ga_emit(const Control* __restrict__ ctl, const int* __restrict__ order, const uint8_t* __restrict__ state,
                          const i64* __restrict__ pos, int* __restrict__ sel) {
    const i64 P = ctl->n_ent;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += stride)
        if (state[order[i]] == GA_KEPT) sel[pos[i]] = (int)i;       // position in the sorted list (ilp_assign.py:122)
}
