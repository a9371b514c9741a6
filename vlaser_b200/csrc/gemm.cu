// tcgen05 / TMEM / TMA GEMM for sm_100a. See gemm.cuh for the contract.
//
// CTA = 192 threads, persistent over output tiles (static round-robin), one CTA per SM:
//   warp 0      TMA producer (one lane): A tile [128 x 64] + W tile [BLOCK_N x 64] per stage
//   warp 1      TMEM allocator + MMA issuer (one lane): 4 x tcgen05.mma (K=16) per stage
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns per warp, fused epilogue, global stores
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM accumulator double buffer full/empty
// (MMA <-> epilogue) so the epilogue of tile i overlaps the mainloop of tile i+1.
// The live row count M may come from device memory (var-len packed prefix), which keeps every
// launch configuration static and the whole policy call CUDA-graph capturable.
#include "gemm.cuh"

#include <mutex>

namespace vlaser {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int GEMM_THREADS = 192;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : 6;
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct TileSched {
  int m_tiles, n_tiles, base, splits, kb_per_split, k_blocks, total;
};

template <int BLOCK_N, int EPI>
__device__ __forceinline__ TileSched make_sched(const GemmDims& d, int M) {
  TileSched s;
  s.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  s.n_tiles = d.N / BLOCK_N;
  s.k_blocks = d.K / BLOCK_K;
  s.base = s.m_tiles * s.n_tiles;
  int splits = 1;
  if (EPI == EPI_RESID_F32 && d.max_splits > 1 && s.base > 0) {
    splits = (int)gridDim.x / s.base;
    if (splits > d.max_splits) splits = d.max_splits;
    if (splits > s.k_blocks / 2) splits = s.k_blocks / 2;  // at least two k-blocks per split
    if (splits < 1) splits = 1;
  }
  s.kb_per_split = (s.k_blocks + splits - 1) / splits;
  s.splits = (s.k_blocks + s.kb_per_split - 1) / s.kb_per_split;
  s.total = s.base * s.splits;
  return s;
}

// ------------------------------------------------------------------------------------------
// Fused epilogue on one 32-column chunk `v` of one row (plus a partner chunk `w` for the
// epilogues that combine two chunks: SwiGLU gate/up, RoPE x[i] / x[i+64]).
//   col0  : global column of v[0];  for SWIGLU w starts at col0+32, for ROPE at col0+64.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
  uint4* p = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
    p[i] = u;
  }
}

__device__ __forceinline__ void add_vec32(float (&v)[32], const float* __restrict__ src) {
  const float4* p = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 b = __ldg(p + i);
    v[4 * i + 0] += b.x;
    v[4 * i + 1] += b.y;
    v[4 * i + 2] += b.z;
    v[4 * i + 3] += b.w;
  }
}

template <int EPI>
__device__ __forceinline__ void epi_chunk(const EpiParams& ep, int N, int row, bool row_ok, int col0, float (&v)[32],
                                          float (&w)[32], bool first_split, bool atomic) {
  if constexpr (EPI == EPI_STORE_BF16 || EPI == EPI_GELU_BF16 || EPI == EPI_SILU_BF16) {
    if (ep.bias) add_vec32(v, ep.bias + col0);
    if constexpr (EPI == EPI_GELU_BF16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    }
    if constexpr (EPI == EPI_SILU_BF16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = silu(v[j]);
    }
    if (row_ok) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldo + col0, v);
  } else if constexpr (EPI == EPI_RESID_F32) {
    if (ep.bias && first_split) add_vec32(v, ep.bias + col0);
    if (ep.scale) {
      const float4* sp = reinterpret_cast<const float4*>(ep.scale + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 s = __ldg(sp + i);
        v[4 * i + 0] *= s.x;
        v[4 * i + 1] *= s.y;
        v[4 * i + 2] *= s.z;
        v[4 * i + 3] *= s.w;
      }
    }
    if (row_ok) {
      float* o = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col0;
      if (atomic) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * i), "f"(v[4 * i]),
                       "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                       : "memory");
      } else {
        float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 x = o4[i];
          x.x += v[4 * i + 0];
          x.y += v[4 * i + 1];
          x.z += v[4 * i + 2];
          x.w += v[4 * i + 3];
          o4[i] = x;
        }
      }
    }
  } else if constexpr (EPI == EPI_SWIGLU_BF16) {
    // v = gate chunk (cols col0..col0+31), w = up chunk (col0+32..); output column col0/2.
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = silu(v[j]) * w[j];
    if (row_ok) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldo + (col0 >> 1), v);
  } else if constexpr (EPI == EPI_QKV_ROPE) {
    // one head = 128 columns; v = x[i0..i0+31], w = x[i0+64..i0+95] of that head, i0 in {0,32}.
    const int head = col0 >> 7;
    const int i0 = col0 & 127;
    if (ep.bias) {
      add_vec32(v, ep.bias + col0);
      add_vec32(w, ep.bias + col0 + 64);
    }
    const bool is_q = head < ep.n_q_heads;
    const bool is_k = !is_q && head < ep.n_q_heads + ep.n_kv_heads;
    if (row_ok) {
      if (is_q || is_k) {
        const int pos = ep.row_pos[row];
        const float4* c4 = reinterpret_cast<const float4*>(ep.cos_tab + (size_t)pos * 64 + i0);
        const float4* s4 = reinterpret_cast<const float4*>(ep.sin_tab + (size_t)pos * 64 + i0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 c = __ldg(c4 + i), s = __ldg(s4 + i);
          const float cc[4] = {c.x, c.y, c.z, c.w};
          const float ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = v[4 * i + e], b = w[4 * i + e];
            v[4 * i + e] = a * cc[e] - b * ss[e];   // x*cos + rot(x)*sin, rot(x)[i] = -x[i+64]
            w[4 * i + e] = b * cc[e] + a * ss[e];   //                     rot(x)[i+64] = x[i]
          }
        }
      }
      __nv_bfloat16* dst;
      if (is_q) {
        dst = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldo + head * 128 + i0;
      } else {
        const int kvh = is_k ? head - ep.n_q_heads : head - ep.n_q_heads - ep.n_kv_heads;
        __nv_bfloat16* base = is_k ? ep.k_cache : ep.v_cache;
        dst = base + (size_t)kvh * ep.kv_head_stride + (size_t)ep.row_kvrow[row] * 128 + i0;
      }
      store_bf16x32(dst, v);
      store_bf16x32(dst + 64, w);
    }
  } else if constexpr (EPI == EPI_PATCH_EMBED) {
    if (ep.bias) add_vec32(v, ep.bias + col0);
    const int img = row / ep.patches, p = row - img * ep.patches;
    if (row_ok) {
      add_vec32(v, ep.pos_emb + (size_t)(1 + p) * N + col0);
      float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) +
                                             (size_t)(img * (ep.patches + 1) + 1 + p) * ep.ldo + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  } else if constexpr (EPI == EPI_STORE_F32) {
    if (ep.bias) add_vec32(v, ep.bias + col0);
    if (row_ok) {
      const int dr = ep.row_map ? ep.row_map[row] : row;
      if (dr >= 0) {
        float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (size_t)dr * ep.ldo + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------
template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmDims d,
                    EpiParams ep) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int M = d.M_dev ? min(*d.M_dev, d.M) : d.M;
  const TileSched sc = make_sched<BLOCK_N, EPI>(d, M);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < sc.total; t += gridDim.x) {
        const int split = t / sc.base, rem = t - split * sc.base;
        const int n_idx = rem / sc.m_tiles, m_idx = rem - n_idx * sc.m_tiles;
        const int kb0 = split * sc.kb_per_split;
        const int kb1 = min(sc.k_blocks, kb0 + sc.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, m_idx * BLOCK_M);
          tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, n_idx * BLOCK_N);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < sc.total; t += gridDim.x) {
        const int split = t / sc.base;
        const int kb0 = split * sc.kb_per_split;
        const int kb1 = min(sc.k_blocks, kb0 + sc.kb_per_split);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t da = make_umma_desc_sw128(sa);
          const uint64_t db = make_umma_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128B swizzle atom: +2 in the (addr>>4) field
            umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);  // accumulator ready for the epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ================================ epilogue warps ================================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < sc.total; t += gridDim.x) {
      const int split = t / sc.base, rem = t - split * sc.base;
      const int n_idx = rem / sc.m_tiles, m_idx = rem - n_idx * sc.m_tiles;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int row = m_idx * BLOCK_M + quarter * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BLOCK_N;
      const int n0 = n_idx * BLOCK_N;
      const bool first_split = split == 0;
      const bool atomic = sc.splits > 1;
      if constexpr (EPI == EPI_SWIGLU_BF16) {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 64; ++c) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr + c * 64, r0);
          tmem_ld_32x32(taddr + c * 64 + 32, r1);
          tmem_ld_wait();
          float v[32], w[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = __uint_as_float(r0[j]);
            w[j] = __uint_as_float(r1[j]);
          }
          epi_chunk<EPI>(ep, d.N, row, row_ok, n0 + c * 64, v, w, first_split, atomic);
        }
      } else if constexpr (EPI == EPI_QKV_ROPE) {
#pragma unroll 1
        for (int hh = 0; hh < BLOCK_N / 128; ++hh) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(taddr + hh * 128 + c * 32, r0);
            tmem_ld_32x32(taddr + hh * 128 + c * 32 + 64, r1);
            tmem_ld_wait();
            float v[32], w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = __uint_as_float(r0[j]);
              w[j] = __uint_as_float(r1[j]);
            }
            epi_chunk<EPI>(ep, d.N, row, row_ok, n0 + hh * 128 + c * 32, v, w, first_split, atomic);
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t r0[32];
          tmem_ld_32x32(taddr + c * 32, r0);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
          epi_chunk<EPI>(ep, d.N, row, row_ok, n0 + c * 32, v, v, first_split, atomic);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// CUDA-core cross-check kernel (tests / bring-up only): same epilogue code, plain pointers.
// One thread = one row x one 32-column chunk (plus its partner chunk where the epilogue needs one).
// ------------------------------------------------------------------------------------------
template <int EPI>
__global__ void gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ W, GemmDims d,
                                 EpiParams ep) {
  const int M = d.M_dev ? min(*d.M_dev, d.M) : d.M;
  const int row = blockIdx.y * blockDim.y + threadIdx.y;
  const int unit = blockIdx.x * blockDim.x + threadIdx.x;
  int col0, colw;
  if (EPI == EPI_SWIGLU_BF16) {
    col0 = unit * 64;
    colw = col0 + 32;
  } else if (EPI == EPI_QKV_ROPE) {
    col0 = (unit >> 1) * 128 + (unit & 1) * 32;
    colw = col0 + 64;
  } else {
    col0 = unit * 32;
    colw = col0;
  }
  if (col0 >= d.N) return;
  const bool row_ok = row < M;
  float v[32], w[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = w[j] = 0.f;
  if (row_ok) {
    const __nv_bfloat16* a = A + (size_t)row * d.K;
    for (int k = 0; k < d.K; ++k) {
      const float av = __bfloat162float(a[k]);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += av * __bfloat162float(W[(size_t)(col0 + j) * d.K + k]);
      if (EPI == EPI_SWIGLU_BF16 || EPI == EPI_QKV_ROPE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) w[j] += av * __bfloat162float(W[(size_t)(colw + j) * d.K + k]);
      }
    }
  }
  epi_chunk<EPI>(ep, d.N, row, row_ok, col0, v, w, true, false);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

template <int BLOCK_N, int EPI>
int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDims& d, const EpiParams& ep, int grid,
               cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  static bool attr_set = false;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, EPI>;
  if (!attr_set) {
    VL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    attr_set = true;
  }
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, d, ep);
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <int EPI>
int launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDims& d, const EpiParams& ep, int block_n,
              int grid, cudaStream_t stream) {
  if (block_n == 128) return launch_one<128, EPI>(tmA, tmB, d, ep, grid, stream);
  if constexpr (EPI != EPI_QKV_ROPE) {
    if (block_n == 256) return launch_one<256, EPI>(tmA, tmB, d, ep, grid, stream);
  }
  set_last_error("gemm_launch: unsupported block_n");
  return -2;
}

}  // namespace

int gemm_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  VL_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  VL_CHECK(cols % 8 == 0, "TMA needs a 16-byte aligned row pitch");
  VL_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA needs a 16-byte aligned base");
  VL_CHECK(box_rows <= 256, "TMA box rows <= 256");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

int gemm_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDims& d, const EpiParams& ep, int block_n,
                int epi, cudaStream_t stream) {
  VL_CHECK(d.K % BLOCK_K == 0 && d.K > 0, "gemm: K must be a positive multiple of 64");
  VL_CHECK(block_n == 128 || block_n == 256, "gemm: block_n must be 128 or 256");
  VL_CHECK(d.N % block_n == 0, "gemm: N must be a multiple of block_n");
  VL_CHECK(ep.out != nullptr, "gemm: null output");
  if (d.M <= 0) return 0;
  const int m_tiles = (d.M + BLOCK_M - 1) / BLOCK_M;
  long long tiles = (long long)m_tiles * (d.N / block_n);
  if (epi == EPI_RESID_F32 && d.max_splits > 1) tiles *= d.max_splits;
  const int grid = (int)std::min<long long>(tiles, gemm_num_sms());
  switch (epi) {
    case EPI_STORE_BF16: return launch_bn<EPI_STORE_BF16>(tmA, tmB, d, ep, block_n, grid, stream);
    case EPI_GELU_BF16: return launch_bn<EPI_GELU_BF16>(tmA, tmB, d, ep, block_n, grid, stream);
    case EPI_SILU_BF16: return launch_bn<EPI_SILU_BF16>(tmA, tmB, d, ep, block_n, grid, stream);
    case EPI_RESID_F32: return launch_bn<EPI_RESID_F32>(tmA, tmB, d, ep, block_n, grid, stream);
    case EPI_SWIGLU_BF16: return launch_bn<EPI_SWIGLU_BF16>(tmA, tmB, d, ep, block_n, grid, stream);
    case EPI_QKV_ROPE: return launch_bn<EPI_QKV_ROPE>(tmA, tmB, d, ep, block_n, grid, stream);
    case EPI_PATCH_EMBED: return launch_bn<EPI_PATCH_EMBED>(tmA, tmB, d, ep, block_n, grid, stream);
    case EPI_STORE_F32: return launch_bn<EPI_STORE_F32>(tmA, tmB, d, ep, block_n, grid, stream);
    default: set_last_error("gemm_launch: unknown epilogue"); return -2;
  }
}

template <int EPI>
static int launch_simt_one(const __nv_bfloat16* A, const __nv_bfloat16* W, const GemmDims& d, const EpiParams& ep,
                           cudaStream_t stream) {
  int units = d.N / 32;
  if (EPI == EPI_SWIGLU_BF16) units = d.N / 64;
  if (EPI == EPI_QKV_ROPE) units = d.N / 64;
  dim3 block(32, 4);
  dim3 grid((units + 31) / 32, (d.M + 3) / 4);
  gemm_simt_kernel<EPI><<<grid, block, 0, stream>>>(A, W, d, ep);
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int gemm_launch_simt(const __nv_bfloat16* A, const __nv_bfloat16* W, const GemmDims& d, const EpiParams& ep, int epi,
                     cudaStream_t stream) {
  VL_CHECK(d.N % 128 == 0, "gemm_simt: N must be a multiple of 128");
  if (d.M <= 0) return 0;
  switch (epi) {
    case EPI_STORE_BF16: return launch_simt_one<EPI_STORE_BF16>(A, W, d, ep, stream);
    case EPI_GELU_BF16: return launch_simt_one<EPI_GELU_BF16>(A, W, d, ep, stream);
    case EPI_SILU_BF16: return launch_simt_one<EPI_SILU_BF16>(A, W, d, ep, stream);
    case EPI_RESID_F32: return launch_simt_one<EPI_RESID_F32>(A, W, d, ep, stream);
    case EPI_SWIGLU_BF16: return launch_simt_one<EPI_SWIGLU_BF16>(A, W, d, ep, stream);
    case EPI_QKV_ROPE: return launch_simt_one<EPI_QKV_ROPE>(A, W, d, ep, stream);
    case EPI_PATCH_EMBED: return launch_simt_one<EPI_PATCH_EMBED>(A, W, d, ep, stream);
    case EPI_STORE_F32: return launch_simt_one<EPI_STORE_F32>(A, W, d, ep, stream);
    default: set_last_error("gemm_launch_simt: unknown epilogue"); return -2;
  }
}

}  // namespace vlaser
