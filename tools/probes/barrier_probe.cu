// Micro-benchmark of grid-barrier variants for the persistent expert-stream kernel (148 co-resident CTAs, one
// barrier between every two phases). Each variant runs ITERS back-to-back barriers; prints ns per barrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/barrier_probe tools/probes/barrier_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int ITERS = 2000;
constexpr int THREADS = 320;

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_relaxed(unsigned* p) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void red_release(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned atom_acq_rel(unsigned* p) {
  unsigned v;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// variant, stride (words between consecutive barrier counters), pollers per CTA (1 or 2 threads in different warps)
template <int V>
__global__ void __launch_bounds__(THREADS, 1) barrier_kernel(unsigned* ctr, int stride, int pollers, float* sink, long long* out) {
  const unsigned G = gridDim.x;
  const int tid = threadIdx.x;
  float acc = 0.f;
  long long t0 = 0;
  if (blockIdx.x == 0 && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int it = 0; it < ITERS; ++it) {
    // a little per-thread "work" whose global store must be published by the barrier
    if (tid >= 64) sink[(size_t)blockIdx.x * THREADS + tid] = acc + it;
    if (tid >= 64) asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 worker warps (tid >= 64)
    unsigned* c = ctr + (size_t)it * stride;
    if constexpr (V == 0) {          // fence + relaxed red; relaxed poll; fence  (the kernel's current form)
      if (tid == 64) { fence_gpu(); red_relaxed(c); }
      if (tid == 64 || (pollers > 1 && tid == 0)) { while (ld_relaxed(c) < G) {} fence_gpu(); }
    } else if constexpr (V == 1) {   // red.release; ld.acquire poll (no fences)
      if (tid == 64) red_release(c);
      if (tid == 64 || (pollers > 1 && tid == 0)) { while (ld_acquire(c) < G) {} }
    } else if constexpr (V == 2) {   // fence + relaxed red; relaxed poll with nanosleep back-off; fence
      if (tid == 64) { fence_gpu(); red_relaxed(c); }
      if (tid == 64 || (pollers > 1 && tid == 0)) { while (ld_relaxed(c) < G) { __nanosleep(64); } fence_gpu(); }
    } else if constexpr (V == 3) {   // last arriver publishes a go-flag on another line; everyone polls the flag
      unsigned* go = c + 64;
      if (tid == 64) { if (atom_acq_rel(c) == G - 1) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(go), "r"(1u) : "memory"); } }
      if (tid == 64 || (pollers > 1 && tid == 0)) { while (ld_acquire(go) == 0) {} }
    } else if constexpr (V == 4) {   // two-level: 4 group counters (one per 37 CTAs) -> root counter
      unsigned* grp = c + 32 * (1 + (blockIdx.x & 3));
      const unsigned gsize = (G + 3 - (blockIdx.x & 3)) / 4;
      if (tid == 64) { if (atom_acq_rel(grp) == gsize - 1) red_release(c); }
      if (tid == 64 || (pollers > 1 && tid == 0)) { while (ld_acquire(c) < 4) {} }
    } else if constexpr (V == 5) {   // flag per CTA, polled by one warp (5 words per lane)
      if (tid == 64) { fence_gpu(); asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(ctr + blockIdx.x), "r"((unsigned)it + 1u) : "memory"); }
      if (tid >= 64 && tid < 96) {
        const int lane = tid & 31;
        while (true) {
          bool ok = true;
          for (int i = 0; i < 5; ++i) if (lane + 32 * i < (int)G) ok = ok && ld_relaxed(ctr + lane + 32 * i) >= (unsigned)it + 1u;
          if (__all_sync(0xffffffffu, ok)) break;
        }
        fence_gpu();
      }
    } else if constexpr (V == 6) {   // V0 without the fences (NOT a correct barrier: lower bound of the counter mechanics)
      if (tid == 64) red_relaxed(c);
      if (tid == 64 || (pollers > 1 && tid == 0)) { while (ld_relaxed(c) < G) {} }
    }
    if (tid >= 64) asm volatile("bar.sync 1, 256;" ::: "memory");
    acc += 1.f;
  }
  if (blockIdx.x == 0 && tid == 0) {
    long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    out[0] = t1 - t0;
  }
}

template <int V>
void run(const char* name, int grid, int stride, int pollers, unsigned* ctr, size_t ctr_bytes, float* sink, long long* out) {
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(ctr, 0, ctr_bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    barrier_kernel<V><<<grid, THREADS, 200 * 1024>>>(ctr, stride, pollers, sink, out);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  printf("%-64s stride %4d pollers %d : %7.1f ns / barrier\n", name, stride, pollers, best * 1e6f / ITERS);
  fflush(stdout);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int grid = prop.multiProcessorCount;
  printf("%s, %d SMs, %d barriers per run, %d threads per CTA\n", prop.name, grid, ITERS, THREADS);
  const size_t words = (size_t)ITERS * 256 + 1024;
  unsigned* ctr;
  float* sink;
  long long* out;
  cudaMalloc(&ctr, words * 4);
  cudaMalloc(&sink, (size_t)grid * THREADS * 4);
  cudaMalloc(&out, 64);
#define SETATTR(V) cudaFuncSetAttribute(barrier_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
  SETATTR(0); SETATTR(1); SETATTR(2); SETATTR(3); SETATTR(4); SETATTR(5); SETATTR(6);
  for (int pollers = 1; pollers <= 2; ++pollers) {
    for (int stride : {1, 32, 256}) {
      run<0>("V0 fence+red.relaxed / ld.relaxed poll / fence", grid, stride, pollers, ctr, words * 4, sink, out);
      run<1>("V1 red.release / ld.acquire poll", grid, stride, pollers, ctr, words * 4, sink, out);
    }
    run<2>("V2 V0 + nanosleep(64) back-off", grid, 32, pollers, ctr, words * 4, sink, out);
    run<3>("V3 atom.acq_rel, last arriver sets go-flag / ld.acquire poll", grid, 256, pollers, ctr, words * 4, sink, out);
    run<4>("V4 two-level (4 group counters -> root)", grid, 256, pollers, ctr, words * 4, sink, out);
    run<6>("V6 V0 without fences (lower bound, not a barrier)", grid, 32, pollers, ctr, words * 4, sink, out);
  }
  run<5>("V5 per-CTA flags polled by one warp", grid, 1, 1, ctr, words * 4, sink, out);
  return 0;
}
