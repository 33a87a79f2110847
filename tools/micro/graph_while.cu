#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* c) { atomicAdd(c, 1); }
__global__ void setc(cudaGraphConditionalHandle h, const int* c) { cudaGraphSetConditional(h, *c < 5 ? 1u : 0u); }
int main() {
    int* c; cudaMalloc(&c, 4); cudaMemset(c, 0, 4);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t n; printf("add %d\n", (int)cudaGraphAddNode(&n, g, nullptr, 0, &p));
    cudaGraph_t b = p.conditional.phGraph_out[0];
    cudaStream_t s; cudaStreamCreate(&s);
    printf("begin %d\n", (int)cudaStreamBeginCaptureToGraph(s, b, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    body<<<1, 1, 0, s>>>(c); setc<<<1, 1, 0, s>>>(h, c);
    printf("end %d\n", (int)cudaStreamEndCapture(s, nullptr));
    cudaGraphExec_t e; printf("inst %d\n", (int)cudaGraphInstantiate(&e, g, 0));
    cudaGraphLaunch(e, 0); cudaDeviceSynchronize();
    int hc; cudaMemcpy(&hc, c, 4, cudaMemcpyDeviceToHost); printf("count %d\n", hc);
}
