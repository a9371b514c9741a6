// Fused GEMM epilogues shared by the tcgen05 GEMM kernels (gemm.cu) and the persistent expert-stream kernel
// (expert.cu): one 32-column accumulator chunk of one row at a time, read from TMEM with tcgen05.ld.
#pragma once
#include "gemm.cuh"

namespace vlaser {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;

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

// a_next[row, col0..col0+31] = bf16(v * gamma_next), returns sum v^2 (fused RMSNorm producer side)
__device__ __forceinline__ float norm_out_chunk(const EpiParams& ep, int row, int col0, const float (&v)[32]) {
  float ss = 0.f;
  float y[32];
  const float4* g4 = reinterpret_cast<const float4*>(ep.gamma_next + col0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 g = __ldg(g4 + i);
    ss += (v[4 * i] * v[4 * i] + v[4 * i + 1] * v[4 * i + 1]) + (v[4 * i + 2] * v[4 * i + 2] + v[4 * i + 3] * v[4 * i + 3]);
    y[4 * i + 0] = v[4 * i + 0] * g.x;
    y[4 * i + 1] = v[4 * i + 1] * g.y;
    y[4 * i + 2] = v[4 * i + 2] * g.z;
    y[4 * i + 3] = v[4 * i + 3] * g.w;
  }
  store_bf16x32(ep.a_next + (size_t)row * ep.lda_next + col0, y);
  return ss;
}

// ---- EPI_QKV_ROPE: one head = 128 columns; a unit is x[i0..i0+31] (chunk v) and x[i0+64..i0+95] (chunk w) of that
// head, i0 in {0, 32}. The RoPE factors / KV-cache row of a unit depend on the row only, not on the accumulator:
// rope_prefetch may run while the mainloop is still in flight (the tcgen05 kernels call it before waiting on the
// accumulator), rope_apply finishes bias + rotation + the q / KV-cache stores.
struct RopeOperands {
  float cs[32], sn[32];
  int kvrow;
};
__device__ __forceinline__ void rope_prefetch(const EpiParams& ep, int row, bool row_ok, int col0, RopeOperands& ro) {
  const int head = col0 >> 7, i0 = col0 & 127;
  const bool is_q = head < ep.n_q_heads;
  const bool is_k = !is_q && head < ep.n_q_heads + ep.n_kv_heads;
  ro.kvrow = 0;
  if (!row_ok) {
#pragma unroll
    for (int j = 0; j < 32; ++j) ro.cs[j] = 1.f, ro.sn[j] = 0.f;
    return;
  }
  if (!is_q) ro.kvrow = __ldg(ep.row_kvrow + row);
  if (is_q || is_k) {
    const int pos = __ldg(ep.row_pos + row);
    const float4* c4 = reinterpret_cast<const float4*>(ep.cos_tab + (size_t)pos * 64 + i0);
    const float4* s4 = reinterpret_cast<const float4*>(ep.sin_tab + (size_t)pos * 64 + i0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 c = __ldg(c4 + i), s = __ldg(s4 + i);
      ro.cs[4 * i] = c.x, ro.cs[4 * i + 1] = c.y, ro.cs[4 * i + 2] = c.z, ro.cs[4 * i + 3] = c.w;
      ro.sn[4 * i] = s.x, ro.sn[4 * i + 1] = s.y, ro.sn[4 * i + 2] = s.z, ro.sn[4 * i + 3] = s.w;
    }
  }
}
// the cos / sin factors of a row whose position id is already known (see the tcgen05 kernel: issued beside the TMEM loads)
__device__ __forceinline__ void rope_factors(const EpiParams& ep, int pos, bool row_ok, int col0, RopeOperands& ro) {
  const int head = col0 >> 7, i0 = col0 & 127;
  if (row_ok && head < ep.n_q_heads + ep.n_kv_heads) {
    const float4* c4 = reinterpret_cast<const float4*>(ep.cos_tab + (size_t)pos * 64 + i0);
    const float4* s4 = reinterpret_cast<const float4*>(ep.sin_tab + (size_t)pos * 64 + i0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 c = __ldg(c4 + i), s = __ldg(s4 + i);
      ro.cs[4 * i] = c.x, ro.cs[4 * i + 1] = c.y, ro.cs[4 * i + 2] = c.z, ro.cs[4 * i + 3] = c.w;
      ro.sn[4 * i] = s.x, ro.sn[4 * i + 1] = s.y, ro.sn[4 * i + 2] = s.z, ro.sn[4 * i + 3] = s.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) ro.cs[j] = 1.f, ro.sn[j] = 0.f;
  }
}
// bias + rotation of one unit in registers
// (bias_s: optional shared-memory copy of bias[col0..col0+31] | bias[col0+64..col0+95], fetched ahead of the accumulator)
__device__ __forceinline__ void rope_rotate(const EpiParams& ep, int col0, float (&v)[32], float (&w)[32],
                                            const RopeOperands& ro, const float* bias_s = nullptr) {
  const int head = col0 >> 7;
  if (ep.bias) {
    if (bias_s) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(bias_s + 4 * i);
        const float4 b = *reinterpret_cast<const float4*>(bias_s + 32 + 4 * i);
        v[4 * i] += a.x, v[4 * i + 1] += a.y, v[4 * i + 2] += a.z, v[4 * i + 3] += a.w;
        w[4 * i] += b.x, w[4 * i + 1] += b.y, w[4 * i + 2] += b.z, w[4 * i + 3] += b.w;
      }
    } else {
      add_vec32(v, ep.bias + col0);
      add_vec32(w, ep.bias + col0 + 64);
    }
  }
  if (head < ep.n_q_heads + ep.n_kv_heads) {  // q or k head
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float a = v[j], b = w[j];
      v[j] = a * ro.cs[j] - b * ro.sn[j];   // x*cos + rot(x)*sin, rot(x)[i] = -x[i+64]
      w[j] = b * ro.cs[j] + a * ro.sn[j];   //                     rot(x)[i+64] = x[i]
    }
  }
}
// where x[i0] of this row's unit goes: the q buffer, or the row's slot in the K / V cache
__device__ __forceinline__ __nv_bfloat16* rope_dst(const EpiParams& ep, int row, int col0, const RopeOperands& ro) {
  const int head = col0 >> 7, i0 = col0 & 127;
  if (head < ep.n_q_heads) return reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldo + head * 128 + i0;
  const bool is_k = head < ep.n_q_heads + ep.n_kv_heads;
  const int kvh = is_k ? head - ep.n_q_heads : head - ep.n_q_heads - ep.n_kv_heads;
  return (is_k ? ep.k_cache : ep.v_cache) + (size_t)kvh * ep.kv_head_stride + (size_t)ro.kvrow * 128 + i0;
}
__device__ __forceinline__ void rope_apply(const EpiParams& ep, int row, bool row_ok, int col0, float (&v)[32],
                                           float (&w)[32], const RopeOperands& ro) {
  rope_rotate(ep, col0, v, w, ro);
  if (!row_ok) return;
  __nv_bfloat16* dst = rope_dst(ep, row, col0, ro);
  store_bf16x32(dst, v);
  store_bf16x32(dst + 64, w);
}

// `rstd`: per-row scale of the accumulator (fused RMSNorm consumer side, 1 when unused);
// `ss`: per-thread running sum of squares of the produced row (producer side).
template <int EPI>
__device__ __forceinline__ void epi_chunk(const EpiParams& ep, int N, int row, bool row_ok, int col0, float (&v)[32],
                                          float (&w)[32], bool first_split, bool atomic, float rstd, float& ss,
                                          float* part_tile = nullptr) {
  if (ep.row_ss) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= rstd;
    if constexpr (EPI == EPI_SWIGLU_BF16 || EPI == EPI_QKV_ROPE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) w[j] *= rstd;
    }
  }
  if constexpr (EPI == EPI_STORE_BF16 || EPI == EPI_GELU_BF16 || EPI == EPI_SILU_BF16) {
    if (ep.bias) add_vec32(v, ep.bias + col0);
    if constexpr (EPI == EPI_GELU_BF16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
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
      if (part_tile) {
        // deterministic split-K: this split's partial goes to the workspace, the consumer sums in split order
        float4* p4 = reinterpret_cast<float4*>(part_tile + (size_t)row * ep.ldo + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) __stcg(p4 + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
      } else if (atomic) {
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
          v[4 * i + 0] = x.x;
          v[4 * i + 1] = x.y;
          v[4 * i + 2] = x.z;
          v[4 * i + 3] = x.w;
        }
        if (ep.a_next) ss += norm_out_chunk(ep, row, col0, v);
      }
    }
  } else if constexpr (EPI == EPI_SWIGLU_BF16) {
    // v = gate chunk (cols col0..col0+31), w = up chunk (col0+32..); output column col0/2.
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = silu(v[j]) * w[j];
    if (row_ok) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldo + (col0 >> 1), v);
  } else if constexpr (EPI == EPI_QKV_ROPE) {
    RopeOperands ro;
    rope_prefetch(ep, row, row_ok, col0, ro);
    rope_apply(ep, row, row_ok, col0, v, w, ro);
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
        if (ep.a_next) ss += norm_out_chunk(ep, dr, col0, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// One epilogue warp's share of an accumulator tile: TMEM lane quarter `taddr` (32 rows, one per lane), column
// half `half` of the BN-wide tile. Units of 32 (plain), 64 (SwiGLU gate|up) or 2x32 (RoPE x[i], x[i+64]) columns.
// ------------------------------------------------------------------------------------------
template <int BN, int EPI, int PARTS = 2>
__device__ __forceinline__ void epilogue_tile(const EpiParams& ep, int N, uint32_t taddr, int half, int row, bool row_ok,
                                              int n0, bool first_split, bool atomic, float rstd, float& ss,
                                              float* part_tile = nullptr) {
  if constexpr (EPI == EPI_SWIGLU_BF16) {
    constexpr int U = BN / 64 / PARTS;
#pragma unroll 1
    for (int c = half * U; c < (half + 1) * U; ++c) {
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
      epi_chunk<EPI>(ep, N, row, row_ok, n0 + c * 64, v, w, first_split, atomic, rstd, ss);
    }
  } else if constexpr (EPI == EPI_QKV_ROPE) {
    constexpr int U = BN / 64 / PARTS;  // unit u: head u>>1, half-head chunk u&1
#pragma unroll 1
    for (int u = half * U; u < (half + 1) * U; ++u) {
      const int off = (u >> 1) * 128 + (u & 1) * 32;
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(taddr + off, r0);
      tmem_ld_32x32(taddr + off + 64, r1);
      tmem_ld_wait();
      float v[32], w[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = __uint_as_float(r0[j]);
        w[j] = __uint_as_float(r1[j]);
      }
      epi_chunk<EPI>(ep, N, row, row_ok, n0 + off, v, w, first_split, atomic, rstd, ss);
    }
  } else {
    constexpr int U = BN / 32 / PARTS;
#pragma unroll 1
    for (int c = half * U; c < (half + 1) * U; ++c) {
      uint32_t r0[32];
      tmem_ld_32x32(taddr + c * 32, r0);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
      epi_chunk<EPI>(ep, N, row, row_ok, n0 + c * 32, v, v, first_split, atomic, rstd, ss, part_tile);
    }
  }
}

// ---- staging of one accumulator row chunk into a 32-row x 128-byte, 128B-swizzled smem tile (one row per lane) that
// leaves the SM as ONE TMA transaction (bulk tensor store, or reduce-add performed at L2)
__device__ __forceinline__ void stage_row_f32(uint8_t* buf, int lane, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
// 32 bf16 outputs = chunks [j0, j0+4) of the 128-byte row
__device__ __forceinline__ void stage_row_bf16(uint8_t* buf, int lane, int j0, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 u;
    u.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
    u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
    u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
    u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(buf + lane * 128 + (((j0 + j) ^ (lane & 7)) << 4)) = u;
  }
}
// two 32-column chunks with ONE wait: a tcgen05.ld round trip is ~0.3 us (in-situ stamps, profiles/r02j), and a one-shot
// epilogue is a chain of them
__device__ __forceinline__ void load_acc32x2(uint32_t taddr0, uint32_t taddr1, float (&v0)[32], float (&v1)[32]) {
  uint32_t r0[32], r1[32];
  tmem_ld_32x32(taddr0, r0);
  tmem_ld_32x32(taddr1, r1);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    v0[j] = __uint_as_float(r0[j]);
    v1[j] = __uint_as_float(r1[j]);
  }
}
__device__ __forceinline__ void load_acc32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}


}  // namespace
}  // namespace vlaser
