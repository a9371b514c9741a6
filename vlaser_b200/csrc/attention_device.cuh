// Device-side body of the fused attention tile (see attention.cuh), shared by the stand-alone attention kernel
// (attention.cu) and the persistent expert-stream kernel (expert.cu). One call = one (sample, kv head, 64-row
// query block) work item processed by a group of ATT_THREADS threads; `sync` is the barrier of that group.
#pragma once
#include "attention.cuh"

namespace vlaser {
namespace {

constexpr int ATT_THREADS = 128;
constexpr int ATT_BM = 64;   // rows per CTA
constexpr int ATT_BN = 64;   // kv rows per block

template <int HD>
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  // 16-byte chunk index within a row of HD bf16, XOR-swizzled with the low 3 row bits
  return static_cast<uint32_t>(row * (HD * 2) + ((chunk ^ (row & 7)) << 4));
}

template <int HD>
constexpr int attention_tile_smem() { return (ATT_BM + 4 * ATT_BN) * HD * 2; }

// `dep_wait` is called exactly once by every thread that does not return early: after it the outputs of the
// previous grid (Q and the new K/V rows) may be read.
// CAUSAL is a template parameter so that the policy call's instantiation is the code it was before the text path existed.
template <int HD, bool CAUSAL = false, typename Sync, typename DepWait>
__device__ __forceinline__ void attention_tile(const AttnParams& p, int b, int kvh, int qb, uint8_t* smem, int tid,
                                               Sync sync, DepWait dep_wait) {
  constexpr int CH = HD / 8;  // 16-byte chunks per row
  uint8_t* Qs = smem;                                  // [64][HD]
  uint8_t* Ks = Qs + ATT_BM * HD * 2;                  // [2][64][HD]
  uint8_t* Vs = Ks + 2 * ATT_BN * HD * 2;              // [2][64][HD]

  const int warp = tid >> 5, lane = tid & 31;
  const bool has_extra = p.q2 != nullptr;
  const int q_len = p.q_len ? p.q_len[b] : p.uni_q_len;
  const int q_start = p.q_len ? p.q_start[b] : b * p.uni_q_len;
  const int kv_len = p.kv_len ? p.kv_len[b] : p.uni_kv_len;
  const int kv_start = p.kv_len ? p.kv_start[b] : b * p.uni_kv_len;
  const int group = p.group;
  const int rows_total = group * (q_len + (has_extra ? 1 : 0));
  const int r0 = qb * ATT_BM;
  if (r0 >= rows_total || kv_len < p.min_kv_len) return;  // (exits without dep_wait: nothing was read)
  const bool cta_has_extra = has_extra && (r0 + ATT_BM > group * q_len);
  // causal (text prefill of the chat VLM): query token t sees keys [0, kv_len - q_len + t]; the CTA stops at the last key
  // its last token sees
  const int causal_off = kv_len - q_len;
  int kv_hi = kv_len + (cta_has_extra ? 1 : 0);
  if constexpr (CAUSAL) kv_hi = min(kv_len, causal_off + min(q_len - 1, (r0 + ATT_BM - 1) / group) + 1);
  const int n_blocks = (kv_hi + ATT_BN - 1) / ATT_BN;

  const __nv_bfloat16* kbase = p.k + (size_t)kvh * p.kv_head_stride + (size_t)kv_start * p.kv_row_stride;
  const __nv_bfloat16* vbase = p.v + (size_t)kvh * p.kv_head_stride + (size_t)kv_start * p.kv_row_stride;

  auto load_kv = [&](int blk, int buf) {
    uint8_t* kd = Ks + buf * ATT_BN * HD * 2;
    uint8_t* vd = Vs + buf * ATT_BN * HD * 2;
    for (int i = tid; i < ATT_BN * CH; i += ATT_THREADS) {
      const int row = i / CH, c = i - row * CH;
      const int kr = blk * ATT_BN + row;
      const bool ok = kr < kv_hi;
      const size_t off = (size_t)(ok ? kr : 0) * p.kv_row_stride + c * 8;
      cp_async_16(kd + swz<HD>(row, c), kbase + off, ok);
      cp_async_16(vd + swz<HD>(row, c), vbase + off, ok);
    }
  };
  // block 0 lies entirely in the stable prefix: fetch it before waiting on the previous grid
  const bool pre = p.stable_prefix && !has_extra && (kv_len - q_len >= ATT_BN);
  if (pre) {
    load_kv(0, 0);
    cp_async_commit();
  }
  dep_wait();
  // ---- Q tile -> smem (zero-fill rows past rows_total)
  for (int i = tid; i < ATT_BM * CH; i += ATT_THREADS) {
    const int row = i / CH, c = i - row * CH;
    const int r = r0 + row;
    const bool ok = r < rows_total;
    const int t = ok ? r / group : 0, hh = ok ? r - t * group : 0;
    const __nv_bfloat16* src;
    if (has_extra && t == q_len)
      src = p.q2 + (size_t)b * p.q2_row_stride + (size_t)(kvh * group + hh) * HD + c * 8;
    else
      src = p.q + (size_t)(q_start + t) * p.q_row_stride + (size_t)(kvh * group + hh) * HD + c * 8;
    cp_async_16(Qs + swz<HD>(row, c), src, ok);
  }
  if (!pre) load_kv(0, 0);
  cp_async_commit();

  // ---- per-thread row bookkeeping: rows rlo = warp*16 + lane/4 and rlo + 8
  const int rlo = warp * 16 + (lane >> 2);
  int lim[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = r0 + rlo + h * 8;
    if (r >= rows_total) lim[h] = 0;
    else if (CAUSAL) lim[h] = min(kv_len, causal_off + r / group + 1);
    else lim[h] = (has_extra && (r / group) == q_len) ? kv_len + 1 : kv_len;
  }
  const bool warp_active = (r0 + warp * 16) < rows_total;
  const float sl2 = p.scale * 1.4426950408889634f;

  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[HD / 16][4];

  const uint32_t qs_u = smem_u32(Qs);
  for (int jb = 0; jb < n_blocks; ++jb) {
    const int buf = jb & 1;
    if (jb + 1 < n_blocks) {
      load_kv(jb + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    sync();
    if (warp_active) {
      if (jb == 0) {
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const int mi = lane >> 3;
          const int row = warp * 16 + (mi & 1) * 8 + (lane & 7);
          ldmatrix_x4(qf[kk], qs_u + swz<HD>(row, kk * 2 + (mi >> 1)));
        }
      }
      const uint32_t ks_u = smem_u32(Ks + buf * ATT_BN * HD * 2);
      const uint32_t vs_u = smem_u32(Vs + buf * ATT_BN * HD * 2);
      // ---- S = Q K^T  (16 x 64 per warp)
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          uint32_t kf[4];
          const int mi = lane >> 3;
          const int row = jp * 16 + (mi >> 1) * 8 + (lane & 7);
          ldmatrix_x4(kf, ks_u + swz<HD>(row, kk * 2 + (mi & 1)));
          mma_bf16_16816(s[2 * jp], qf[kk], kf[0], kf[1]);
          mma_bf16_16816(s[2 * jp + 1], qf[kk], kf[2], kf[3]);
        }
      }
      // ---- mask + online softmax (exp2 domain)
      const int colbase = jb * ATT_BN + (lane & 3) * 2;
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = colbase + j * 8 + (e & 1);
          const int h = e >> 1;
          const float val = (col < lim[h]) ? s[j][e] * sl2 : -INFINITY;
          s[j][e] = val;
          mx[h] = fmaxf(mx[h], val);
        }
      }
      float alpha[2], moff[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
        const float m_new = fmaxf(m_run[h], mx[h]);
        moff[h] = (m_new == -INFINITY) ? 0.f : m_new;
        alpha[h] = exp2f(m_run[h] - moff[h]);  // m_run = -inf -> 0
        m_run[h] = m_new;
      }
      float rs[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = exp2f(s[j][e] - moff[e >> 1]);
          s[j][e] = pv;
          rs[e >> 1] += pv;
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        rs[h] += __shfl_xor_sync(0xffffffffu, rs[h], 1);
        rs[h] += __shfl_xor_sync(0xffffffffu, rs[h], 2);
        l_run[h] = l_run[h] * alpha[h] + rs[h];
      }
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        o[i][0] *= alpha[0];
        o[i][1] *= alpha[0];
        o[i][2] *= alpha[1];
        o[i][3] *= alpha[1];
      }
      // ---- O += P V
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t pf[4];
        pf[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
        pf[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
        pf[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
        pf[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
        for (int dp = 0; dp < HD / 16; ++dp) {
          uint32_t vf[4];
          const int mi = lane >> 3;
          const int row = ks * 16 + (mi & 1) * 8 + (lane & 7);
          ldmatrix_x4_trans(vf, vs_u + swz<HD>(row, dp * 2 + (mi >> 1)));
          mma_bf16_16816(o[2 * dp], pf, vf[0], vf[1]);
          mma_bf16_16816(o[2 * dp + 1], pf, vf[2], vf[3]);
        }
      }
    }
    sync();
  }

  // ---- normalise and store
  if (!warp_active) return;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = r0 + rlo + h * 8;
    if (r >= rows_total) continue;
    const int t = r / group, hh = r - t * group;
    const float inv = l_run[h] > 0.f ? 1.f / l_run[h] : 0.f;
    __nv_bfloat16* dst;
    if (has_extra && t == q_len)
      dst = p.out2 + (size_t)b * p.out2_row_stride + (size_t)(kvh * group + hh) * HD;
    else
      dst = p.out + (size_t)(q_start + t) * p.out_row_stride + (size_t)(kvh * group + hh) * HD;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      const uint32_t u = pack_bf16(o[i][2 * h] * inv, o[i][2 * h + 1] * inv);
      *reinterpret_cast<uint32_t*>(dst + i * 8 + (lane & 3) * 2) = u;
    }
  }
}

}  // namespace
}  // namespace vlaser
