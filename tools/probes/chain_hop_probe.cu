// Probe: cost of one hop of a dependent-launch chain inside a CUDA graph, (a) griddepcontrol.wait (what the engine does),
// (b) the same programmatic launch but the consumer polls a completion counter the producer CTAs bump after their stores
// (release / acquire at gpu scope) and never executes griddepcontrol.wait.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/chain_hop_probe tools/probes/chain_hop_probe.cu && /tmp/chain_hop_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// mode 0: pdl wait; mode 1: flag polling. in/out: [grid * 128] floats; flags: one counter per launch
__global__ void __launch_bounds__(128) hop_kernel(const float* __restrict__ in, float* __restrict__ out, unsigned* flags, int idx,
                                                  int prev_grid, int mode, int work) {
  pdl_launch();
  if (mode == 0) {
    pdl_wait();
  } else if (idx > 0) {
    if (threadIdx.x == 0) {
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + idx - 1) : "memory");
      } while (v < (unsigned)prev_grid);
    }
    __syncthreads();
  }
  const int i = blockIdx.x * 128 + threadIdx.x;
  float v = in[(i * 7 + 3) % (prev_grid * 128)];
  for (int k = 0; k < work; ++k) v = v * 1.0001f + 0.5f;
  out[i] = v;
  if (mode == 1) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(flags + idx) : "memory");
    }
  }
}

int main() {
  const int hops = 2000, grids[] = {16, 72, 144};
  float *a, *b;
  unsigned* flags;
  cudaMalloc(&a, 148 * 128 * 4);
  cudaMalloc(&b, 148 * 128 * 4);
  cudaMalloc(&flags, hops * 4);
  cudaMemset(a, 0, 148 * 128 * 4);
  cudaStream_t st;
  cudaStreamCreate(&st);
  for (int grid : grids)
    for (int work : {0, 2000})
      for (int mode = 0; mode < 2; ++mode) {
        cudaGraph_t g;
        cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        cudaMemsetAsync(flags, 0, hops * 4, st);
        for (int h = 0; h < hops; ++h) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(grid);
          cfg.blockDim = dim3(128);
          cfg.stream = st;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          at[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at;
          cfg.numAttrs = 1;
          cudaLaunchKernelEx(&cfg, hop_kernel, (const float*)((h & 1) ? b : a), (h & 1) ? a : b, flags, h, grid, mode, work);
        }
        cudaStreamEndCapture(st, &g);
        cudaGraphInstantiate(&ge, g, 0);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int w = 0; w < 2; ++w) cudaGraphLaunch(ge, st);
        cudaStreamSynchronize(st);
        cudaEventRecord(e0, st);
        for (int w = 0; w < 5; ++w) cudaGraphLaunch(ge, st);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("grid %3d work %4d mode %s: %.3f us per hop (%s)\n", grid, work, mode ? "flag-poll" : "pdl-wait ", ms * 1e3 / (5 * hops),
               cudaGetErrorString(cudaGetLastError()));
        cudaGraphExecDestroy(ge);
        cudaGraphDestroy(g);
      }
  return 0;
}
