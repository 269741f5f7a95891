// Microbenchmark: what does a nested IF conditional graph node cost on the critical path of a chain of tiny kernels?
//   A: chain of 8 "levels" x 5 kernels, plain graph
//   B: the same with levels 4..8 wrapped in nested IF nodes whose condition is true for `live` levels
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cond_nodes cond_nodes.cu && ./cond_nodes
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_work(int* p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] += 1; }
__global__ void k_set(cudaGraphConditionalHandle h, const int* live, int level) { if (threadIdx.x == 0 && blockIdx.x == 0) cudaGraphSetConditional(h, level <= *live ? 1u : 0u); }
__global__ void k_work_set(int* p, int n, cudaGraphConditionalHandle h, const int* live, int level) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] += 1;
    if (i == 0) cudaGraphSetConditional(h, level <= *live ? 1u : 0u);
}

static const int LEVELS = 8, PER = 5, FIRST_COND = 4;
static int* d_p; static int* d_live;

static void level_body(cudaStream_t st, int grid) { for (int i = 0; i < PER; i++) k_work<<<grid, 256, 0, st>>>(d_p, grid * 256); }

// capture levels [k, LEVELS] into the stream's current capture; levels >= FIRST_COND go into nested IF bodies
static int capture_from(cudaStream_t st, int k, int grid, std::vector<cudaGraph_t>& todo_graph, std::vector<int>& todo_level) {
    for (; k <= LEVELS; k++) {
        if (k >= FIRST_COND) {
            cudaStreamCaptureStatus status; cudaGraph_t g; const cudaGraphNode_t* deps; size_t nd;
            CK(cudaStreamGetCaptureInfo(st, &status, nullptr, &g, &deps, &nd));
            cudaGraphConditionalHandle h;
            CK(cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault));
            // the kernel before the IF node sets the condition
            k_set<<<1, 32, 0, st>>>(h, d_live, k);
            CK(cudaStreamGetCaptureInfo(st, &status, nullptr, &g, &deps, &nd));
            cudaGraphNodeParams cp = {};
            cp.type = cudaGraphNodeTypeConditional;
            cp.conditional.handle = h; cp.conditional.type = cudaGraphCondTypeIf; cp.conditional.size = 1;
            cudaGraphNode_t node;
            CK(cudaGraphAddNode(&node, g, deps, nd, &cp));
            CK(cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies));
            todo_graph.push_back(cp.conditional.phGraph_out[0]); todo_level.push_back(k);
            return 0;       // the rest of the levels live inside the body
        }
        level_body(st, grid);
    }
    return 0;
}

int main() {
    const int grid = 148;
    CK(cudaMalloc(&d_p, grid * 256 * 4)); CK(cudaMemset(d_p, 0, grid * 256 * 4)); CK(cudaMalloc(&d_live, 4));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // A: plain chains of 3..8 levels
    for (int nl = 3; nl <= LEVELS; nl++) {
        cudaGraph_t g; cudaGraphExec_t ex;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        for (int k = 1; k <= nl; k++) level_body(st, grid);
        CK(cudaStreamEndCapture(st, &g)); CK(cudaGraphInstantiate(&ex, g, 0));
        for (int i = 0; i < 20; i++) CK(cudaGraphLaunch(ex, st));
        CK(cudaEventRecord(e0, st));
        for (int i = 0; i < 200; i++) CK(cudaGraphLaunch(ex, st));
        CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("plain  levels=%d  %.2f us per graph\n", nl, 1e3 * ms / 200);
    }
    // B: nested IF from level FIRST_COND
    {
        cudaGraph_t g; cudaGraphExec_t ex;
        std::vector<cudaGraph_t> tg; std::vector<int> tl;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        if (capture_from(st, 1, grid, tg, tl)) return 1;
        k_work<<<grid, 256, 0, st>>>(d_p, grid * 256);      // the work after the loop (depends on the IF node)
        CK(cudaStreamEndCapture(st, &g));
        for (size_t i = 0; i < tg.size(); i++) {           // bodies (nested ones are appended while we go)
            CK(cudaStreamBeginCaptureToGraph(st, tg[i], nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
            level_body(st, grid);
            if (capture_from(st, tl[i] + 1, grid, tg, tl)) return 1;
            CK(cudaStreamEndCapture(st, nullptr));
        }
        CK(cudaGraphInstantiate(&ex, g, 0));
        for (int live = 3; live <= LEVELS; live++) {
            CK(cudaMemcpy(d_live, &live, 4, cudaMemcpyHostToDevice));
            for (int i = 0; i < 20; i++) CK(cudaGraphLaunch(ex, st));
            CK(cudaEventRecord(e0, st));
            for (int i = 0; i < 200; i++) CK(cudaGraphLaunch(ex, st));
            CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("nested live=%d  %.2f us per graph (graph holds %d levels, IF from level %d, +1 set kernel per IF, +1 tail kernel)\n", live, 1e3 * ms / 200, LEVELS, FIRST_COND);
        }
    }
    return 0;
}
