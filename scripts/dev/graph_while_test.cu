// Can a conditional WHILE node's body hold COOPERATIVE kernel launches (grid.sync) and be ended from the device?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/gw scripts/dev/graph_while_test.cu && /tmp/gw
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;
__global__ void work(int* counter, double* acc) {
    cg::grid_group grid = cg::this_grid();
    if (threadIdx.x == 0) atomicAdd(acc + 1, 1.0);
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) acc[0] += acc[1];   // sees every block's add
    grid.sync();
}
__global__ void post(int* counter, double* acc, cudaGraphConditionalHandle h, int rounds) {
    cg::grid_group grid = cg::this_grid();
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int c = atomicAdd(counter, 1) + 1;
        cudaGraphSetConditional(h, c < rounds ? 1u : 0u);
    }
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle h;
    CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
    p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node; CK(cudaGraphAddNode(&node, g, nullptr, 0, &p));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    int* d; double* acc; CK(cudaMalloc(&d, 4)); CK(cudaMalloc(&acc, 16)); CK(cudaMemset(d, 0, 4)); CK(cudaMemset(acc, 0, 16));
    int rounds = 7;
    cudaKernelNodeParams kp = {};
    void* a1[] = {&d, &acc};
    kp.func = (void*)work; kp.gridDim = dim3(sms); kp.blockDim = dim3(128); kp.kernelParams = a1;
    cudaGraphNode_t k1, k2; CK(cudaGraphAddKernelNode(&k1, body, nullptr, 0, &kp));
    void* a2[] = {&d, &acc, &h, &rounds};
    kp.func = (void*)post; kp.kernelParams = a2;
    CK(cudaGraphAddKernelNode(&k2, body, &k1, 1, &kp));
    cudaLaunchAttributeValue v = {}; v.cooperative = 1;
    CK(cudaGraphKernelNodeSetAttribute(k1, cudaLaunchAttributeCooperative, &v));
    CK(cudaGraphKernelNodeSetAttribute(k2, cudaLaunchAttributeCooperative, &v));
    cudaGraphExec_t e; CK(cudaGraphInstantiate(&e, g, 0));
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemsetAsync(d, 0, 4, st));
        CK(cudaGraphLaunch(e, st)); CK(cudaStreamSynchronize(st));
        int hc; double ha[2]; CK(cudaMemcpy(&hc, d, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ha, acc, 16, cudaMemcpyDeviceToHost));
        printf("rep %d: rounds run %d (want %d), blocks counted %.0f, acc %.0f\n", rep, hc, rounds, ha[1], ha[0]);
    }
    return 0;
}
