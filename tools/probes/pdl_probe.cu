// Probe: how much does programmatic dependent launch save on a chain of short dependent kernels (stream vs graph)?
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
__global__ void k_chain(float* buf, int iters, int pdl, int pre_iters) {
  // prologue that does not depend on the previous kernel
  float acc = threadIdx.x;
  for (int i = 0; i < pre_iters; ++i) acc = acc * 1.0001f + 0.5f;
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  float v = buf[blockIdx.x * blockDim.x + threadIdx.x];
  for (int i = 0; i < iters; ++i) v = v * 1.0001f + acc * 1e-9f;
  buf[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
static void launch(cudaStream_t st, float* buf, int grid, int iters, int pdl, int pre) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
  cudaLaunchAttribute a[1];
  a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  a[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = a; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k_chain, buf, iters, pdl, pre);
}
int main() {
  float* buf; cudaMalloc(&buf, 148 * 256 * 4 * 8); cudaMemset(buf, 0, 148 * 256 * 4 * 8);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int N = 2000;
  for (int grid : {32, 148, 592}) for (int iters : {200, 2000}) for (int pre : {0, 2000}) for (int pdl = 0; pdl < 2; ++pdl) {
    // stream
    for (int i = 0; i < 50; ++i) launch(st, buf, grid, iters, pdl, pre);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int i = 0; i < N; ++i) launch(st, buf, grid, iters, pdl, pre);
    cudaEventRecord(e1, st); cudaStreamSynchronize(st);
    float ms_s; cudaEventElapsedTime(&ms_s, e0, e1);
    // graph
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < N; ++i) launch(st, buf, grid, iters, pdl, pre);
    cudaStreamEndCapture(st, &g);
    cudaError_t er = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
    float ms_g; cudaEventElapsedTime(&ms_g, e0, e1);
    printf("grid=%3d iters=%4d pre=%4d pdl=%d : stream %.2f us/kernel  graph %.2f us/kernel  (%s)\n", grid, iters, pre, pdl,
           1e3 * ms_s / N, 1e3 * ms_g / N, cudaGetErrorString(er));
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  }
  return 0;
}
