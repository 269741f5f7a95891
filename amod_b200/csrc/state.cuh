// state.cuh — SURVEY.md 8f-3: the epoch state kept in HBM across epochs and updated by deltas.
//
// Requests are immutable once known (request.py:25-41): the request TABLE is indexed by request id and only grows.
// Vehicles (vehicle.py:41-70) are rows of fixed-stride arrays; an epoch uploads only the rows that changed since the last
// one (k_state_scatter) and the build reads the resident arrays: a prefix sum over the schedule lengths and one gather turn the
// fixed-stride schedules into the CSR layout the pipeline consumes (k_state_csr).  Stop codes reference request ids directly
// (2*rid + drop-off), so the pool that comes back names requests by id.
#pragma once
#undef BFILE
#define BFILE 6
#include "common.cuh"

struct FleetState {            // device arrays, capacity n_veh / n_veh * max_stops
    int* nid; double* t; int* load; int* status; int* reb_node; double* reb_ddl;
    int* len; int* code; uint8_t* onboard;
    int n_veh, max_stops;
};

// rows of changed vehicles: packed host->device staging (one memcpy), scattered into the resident arrays
struct FleetDelta {
    const int* idx; const int* nid; const double* t; const int* load; const int* status; const int* reb_node; const double* reb_ddl;
    const int* off;            // [n+1] into code / onboard
    const int* code; const uint8_t* onboard;
    int n;
};

__global__ void k_state_scatter(FleetState s, FleetDelta d, unsigned long long* err) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += gridDim.x * blockDim.x) {
        const int v = d.idx[i];
        const int a = d.off[i], L = d.off[i + 1] - a;
        if (v < 0 || v >= s.n_veh || L < 0 || L > s.max_stops) { atomicOr(err, 1ull); continue; }
        s.nid[v] = d.nid[i]; s.t[v] = d.t[i]; s.load[v] = d.load[i]; s.status[v] = d.status[i];
        s.reb_node[v] = d.reb_node[i]; s.reb_ddl[v] = d.reb_ddl[i]; s.len[v] = L;
        for (int m = 0; m < s.max_stops; m++) {          // (the tail of the row is cleared: the row is exactly the vehicle's schedule)
            s.code[(i64)v * s.max_stops + m] = m < L ? d.code[a + m] : 0;
            s.onboard[(i64)v * s.max_stops + m] = m < L ? d.onboard[a + m] : (uint8_t)0;
        }
    }
}

struct InStateLen {
    const int* len; i64 n;
    __device__ __forceinline__ i64 size() const { return n; }
    __device__ __forceinline__ i64 operator()(i64 v) const { return (i64)len[v]; }
};

// fixed-stride schedules -> CSR (veh_sche_off int32, sche_code, sche_onboard)
__global__ void k_state_csr(FleetState s, const i64* __restrict__ off64, int* __restrict__ off32, int* __restrict__ code, uint8_t* __restrict__ onboard) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v <= s.n_veh; v += gridDim.x * blockDim.x) {
        off32[v] = (int)off64[v];
        if (v == s.n_veh) break;
        const i64 a = off64[v];
        const int L = s.len[v];
        for (int m = 0; m < L; m++) { code[a + m] = s.code[(i64)v * s.max_stops + m]; onboard[a + m] = s.onboard[(i64)v * s.max_stops + m]; }
    }
}
