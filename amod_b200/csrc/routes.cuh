// routes.cuh — SURVEY.md 8f-2: batched route construction.  route_functions.build_route_from_origin_to_dest
// (route_functions.py:36-70) for all legs of the schedules selected in one epoch: the path of a leg is recovered by
// walking the predecessor table backwards from the destination (:38-43), its steps carry the travel time and distance of
// each consecutive node pair (:49-57) plus the closing zero-length step (:59-61), and the leg's duration / distance are the
// left-to-right float64 sums the reference forms (:56-57), so they are bit-identical.
#pragma once
#undef BFILE
#define BFILE 5
#include "common.cuh"

struct RouteTables {
    const int* pred;        // [n*n] 1-based predecessor of d on the best path o -> d; <= 0 at d == o  (shortest_path_table)
    const void* dist;       // [n*n] travel_distance_table, element type = tt's
    int n;
};

// nodes on the path o -> d (both included); 0 = the table does not lead back to o within n hops
__device__ __forceinline__ int route_len(const RouteTables& rt, int o, int d) {
    const int* row = rt.pred + (size_t)(o - 1) * (size_t)rt.n;
    int len = 1;
    for (int x = row[d - 1]; x > 0; x = row[x - 1]) { if (++len > rt.n) return 0; }
    return len;
}

__global__ void k_route_count(RouteTables rt, const int* __restrict__ leg_o, const int* __restrict__ leg_d, int n_legs,
                              int* __restrict__ leg_len, unsigned long long* err) {
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n_legs; g += gridDim.x * blockDim.x) {
        const int o = leg_o[g], d = leg_d[g];
        int len = 0;
        if (o >= 1 && o <= rt.n && d >= 1 && d <= rt.n) len = route_len(rt, o, d);
        if (len == 0) { atomicOr(err, 1ull); len = 1; }
        leg_len[g] = len;
    }
}

// one warp per leg: lane 0 walks the predecessor chain backwards writing the nodes at their forward positions, the lanes
// then look the step times / distances up in parallel, and lane 0 forms the ordered sums
template <typename TT>
__global__ void __launch_bounds__(256)
k_route_fill(RouteTables rt, Table<TT> tt, const int* __restrict__ leg_o, const int* __restrict__ leg_d, int n_legs,
             const i64* __restrict__ off, i64 cap_steps, int* __restrict__ step_u, int* __restrict__ step_v, double* __restrict__ step_t,
             double* __restrict__ step_d, double* __restrict__ leg_t, double* __restrict__ leg_dist) {
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    if (off[n_legs] > cap_steps) return;
    const TT* dist = (const TT*)rt.dist;
    for (int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < n_legs; g += nwarps) {
        const int o = leg_o[g], d = leg_d[g];
        const i64 a = off[g];
        const int len = (int)(off[g + 1] - a);
        if (lane == 0) {
            const int* row = rt.pred + (size_t)(o - 1) * (size_t)rt.n;
            int x = d;
            for (int p = len - 1; p >= 0; p--) {          // path.reverse() (:43): node p of the forward path
                step_v[a + p] = x;                        // (node p, parked in step_v until the lanes pair the nodes up)
                x = p > 0 ? row[x - 1] : x;
            }
        }
        __syncwarp();
        // step k (k < len-1) = (node k, node k+1); nodes were stored as step_v[a + k] = node k
        for (int k = lane; k < len; k += 32) {
            const int u = step_v[a + k];
            step_u[a + k] = u;
        }
        __syncwarp();
        for (int k = lane; k < len; k += 32) {
            const int u = step_u[a + k];
            const int v = k + 1 < len ? step_u[a + k + 1] : u;
            double t = 0.0, dd = 0.0;
            if (k + 1 < len) { t = tt(u, v); dd = (double)__ldg(dist + (size_t)(u - 1) * (size_t)rt.n + (size_t)(v - 1)); }
            step_t[a + k] = t; step_d[a + k] = dd;
        }
        __syncwarp();
        for (int k = lane; k < len; k += 32) step_v[a + k] = k + 1 < len ? step_u[a + k + 1] : step_u[a + k];
        __syncwarp();
        if (lane == 0) {
            double st = 0.0, sd = 0.0;                    // duration += t; distance += d (:56-57), in path order
            for (int k = 0; k + 1 < len; k++) { st += step_t[a + k]; sd += step_d[a + k]; }
            leg_t[g] = st; leg_dist[g] = sd;
        }
        __syncwarp();
    }
}

struct InLegLen {
    const int* p; i64 n;
    __device__ __forceinline__ i64 size() const { return n; }
    __device__ __forceinline__ i64 operator()(i64 i) const { return (i64)p[i]; }
};
struct FinStoreI64 { i64* dst; __device__ __forceinline__ void operator()(i64 tot) const { *dst = tot; } };
