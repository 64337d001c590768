// Floor of a dependent kernel chain inside a CUDA graph: n empty kernels (one 32-thread CTA each), with and without
// programmatic dependent launch; microseconds per replay.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_graph_chain tools/probe_graph_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void empty_kernel(int* p) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p && threadIdx.x == 0) p[blockIdx.x] = 1;
}
static float run(int n, bool pdl, int blocks) {
  cudaStream_t s;
  cudaStreamCreate(&s);
  int* d;
  cudaMalloc(&d, 4096);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(32);
    cfg.stream = s;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    a[0].val.programmaticStreamSerializationAllowed = 1;
    if (pdl) { cfg.attrs = a; cfg.numAttrs = 1; }
    cudaLaunchKernelEx(&cfg, empty_kernel, d);
  }
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  for (int i = 0; i < 50; ++i) cudaGraphLaunch(ge, s);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, s);
  for (int i = 0; i < 1000; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;  // us per replay (1000 replays, ms total)
}
int main() {
  for (int n : {1, 2, 4, 7, 8})
    for (int blocks : {1, 8})
      printf("chain of %d empty kernels x %d CTAs: %.2f us/replay with PDL, %.2f without  [%s]\n", n, blocks, run(n, true, blocks),
             run(n, false, blocks), cudaGetErrorString(cudaGetLastError()));
  return 0;
}
