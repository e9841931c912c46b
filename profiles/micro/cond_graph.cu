#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* c, cudaGraphConditionalHandle h) {
  int v = atomicAdd(c, 1);
  cudaGraphSetConditional(h, v < 9 ? 1u : 0u);
}
int main() {
  int* c; cudaMalloc(&c, 4); cudaMemset(c, 0, 4);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h;
  cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t n;
  cudaError_t e = cudaGraphAddNode(&n, g, nullptr, 0, &p);
  printf("add node: %s\n", cudaGetErrorString(e));
  cudaGraph_t bg = p.conditional.phGraph_out[0];
  cudaStream_t s; cudaStreamCreate(&s);
  e = cudaStreamBeginCaptureToGraph(s, bg, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
  printf("begin capture: %s\n", cudaGetErrorString(e));
  body<<<1,1,0,s>>>(c, h);
  e = cudaStreamEndCapture(s, nullptr);
  printf("end capture: %s\n", cudaGetErrorString(e));
  cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0);
  printf("instantiate: %s\n", cudaGetErrorString(e));
  cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
  int hc; cudaMemcpy(&hc, c, 4, cudaMemcpyDeviceToHost);
  printf("count = %d (expect 10)\n", hc);
  return 0;
}
