// Weight-streaming expert pass for a SINGLE observation (the B = 1 latency case: M = 1 proprio row in the prefill,
// M = H <= 8 action rows in a denoise step). See expert.cuh (expert_gemv_pass_launch).
//
// With <= 8 rows the tensor cores have nothing to do: a 128-row MMA tile is > 94 % padding and the pass is bound by
// streaming the 1.3 GB of expert weights once (0.2 ms at HBM peak) plus the latency of the dependent steps. One
// persistent launch per pass; 148 CTAs x 16 warps stay resident. A layer is 5 phases separated by a grid barrier
// (release-red / acquire-poll, see expert.cu):
//   0  RMSNorm (every CTA redundantly: M x he floats from L2) + q|k|v projection + bias + RoPE + KV-cache write
//   1  attention of the M rows over the cached prefix + their own K/V, one CTA per query head
//   2  o-projection, split-K partial sums red.add'ed into the fp32 residual stream
//   3  RMSNorm (redundant) + gate/up projection + SwiGLU
//   4  down projection, split-K, red.add into the residual stream
// A work unit is one warp x 4 weight rows x one K chunk of 768 elements: every lane issues its 12 independent 16-byte
// weight loads (coalesced 512-byte warp requests, 6 KB per warp, 96 KB per SM in flight) before consuming any of them;
// the activations of the chunk (bf16, staged once per phase in shared memory) are expanded once per 4 weight rows.
// The loads of a warp's FIRST unit of the next phase are issued BEFORE the grid barrier (weights do not depend on it),
// so the barrier and the activation staging hide the HBM latency of the next phase. Weights are read exactly once.
// One thread per CTA also issues bulk HBM -> L2 prefetches (cp.async.bulk.prefetch.L2, through the TMA unit) of the
// matrices one to two phases ahead and of the next layer's norm gains / bias / K/V prefix rows, so that HBM keeps
// streaming while barriers, staging and attention run; the per-layer pointer table lives in shared memory; the
// attention CTAs hold the K/V rows that date from the prefill in registers across the barrier (attn_load_rows).
// Measured per-phase times, the ncu capture and the variants that were tried and not kept: DESIGN.md section 4d,
// profiles/r01k_b1_latency_phases.txt, profiles/r01m_ncu_full_expert_gemv.txt.
#include "expert.cuh"

namespace vlaser {
namespace {

constexpr int GV_WARPS = 16;
constexpr int GV_THREADS = 32 * GV_WARPS;
constexpr int GV_PHASES = 5;
constexpr int GV_NR = 4;              // weight rows per unit
constexpr int GV_NV = 3;              // 16-byte loads per lane and row
constexpr int GV_CH = GV_NV * 256;    // K elements per chunk
constexpr int GV_JS = GV_THREADS / 64;  // kv splits of the PV product

__device__ __forceinline__ void gv_arrive(unsigned int* ctr) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
// (a CTA of the grid is missing -- the launcher checks that the grid is resident, so this is insurance: give up after
// ~2 s, flag the call as failed for the host and run on without waiting; no __trap, which would kill the CUDA context of
// the evaluator that called us)
__device__ __forceinline__ void gv_wait(const unsigned int* ctr, unsigned int target, int* err_flag) {
  unsigned int v;
  long long t0 = 0;
  while (true) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= target) break;
    if (err_flag && *reinterpret_cast<volatile int*>(err_flag) == EXPERT_ERR_BARRIER) break;  // another barrier already gave up
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 4000000000LL) {
      if (err_flag) atomicExch(err_flag, EXPERT_ERR_BARRIER);
      break;
    }
  }
}
// weights: read once, do not pollute L1. volatile: the prefetch before the grid barrier must stay before it
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float round_bf16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x), f[1] = bf16_hi(u.x), f[2] = bf16_lo(u.y), f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z), f[5] = bf16_hi(u.z), f[6] = bf16_lo(u.w), f[7] = bf16_hi(u.w);
}

// 4 weight rows x the K range [k0, k1) (row-major [N, K] bf16 matrices, row pointers at column 0)
struct Unit {
  const __nv_bfloat16* w[GV_NR];
  int k0, k1;
};

struct Dims {
  int he, mlp, qdim, heads, G, b, warp;
};

// the per-layer pointers, copied to shared memory once per launch (the table entry in global memory would be a
// dependent DRAM miss at the head of every phase: the 1.3 GB of weights streamed per pass flush it out of L2)
constexpr int GV_MAX_LAYERS = 48;
struct LayerPtrs {
  const __nv_bfloat16 *w_qkv, *w_o, *w_gu, *w_down;
  const float *ln_in, *ln_post, *bqkv;
};

__device__ __forceinline__ int unit_count(const Dims& d, int kind) {
  if (kind == 0) return d.heads * 32;
  if (kind == 3) return d.mlp / 2;
  const int K = kind == 2 ? d.qdim : d.mlp;
  return (d.he / GV_NR) * ((K + GV_CH - 1) / GV_CH);
}
// first unit of this warp: spread over the CTAs, except the down projection (CTA-contiguous units share one K chunk,
// so a CTA stages only that slice of the SwiGLU output)
__device__ __forceinline__ int first_unit(const Dims& d, int kind) {
  return kind == 4 ? d.b * GV_WARPS + d.warp : d.warp * d.G + d.b;
}
__device__ __forceinline__ bool make_unit(const Dims& d, const LayerPtrs* L, int kind, int u, Unit& U) {
  if (u >= unit_count(d, kind)) return false;
  if (kind == 0) {
    // head h, rotation pairs (i, i + 64) and (i + 32, i + 96)
    const __nv_bfloat16* W = L->w_qkv + (size_t)((u >> 5) * 128 + (u & 31)) * d.he;
    U.w[0] = W, U.w[1] = W + (size_t)64 * d.he, U.w[2] = W + (size_t)32 * d.he, U.w[3] = W + (size_t)96 * d.he;
    U.k0 = 0, U.k1 = d.he;
  } else if (kind == 3) {
    // output columns 2u, 2u + 1 of the interleaved matrix (32-row groups: gate rows [64 g, 64 g + 32), up rows
    // [64 g + 32, 64 g + 64))
    const int col = 2 * u;
    const __nv_bfloat16* W = L->w_gu + (size_t)((col >> 5) * 64 + (col & 31)) * d.he;
    U.w[0] = W, U.w[1] = W + d.he, U.w[2] = W + (size_t)32 * d.he, U.w[3] = W + (size_t)33 * d.he;
    U.k0 = 0, U.k1 = d.he;
  } else {
    const int K = kind == 2 ? d.qdim : d.mlp;
    const int cgs = d.he / GV_NR, ch = u / cgs, cg = u - ch * cgs;
    const __nv_bfloat16* W = (kind == 2 ? L->w_o : L->w_down) + (size_t)(cg * GV_NR) * K;
#pragma unroll
    for (int i = 0; i < GV_NR; ++i) U.w[i] = W + (size_t)i * K;
    U.k0 = ch * GV_CH, U.k1 = min(K, U.k0 + GV_CH);
  }
  return true;
}

__device__ __forceinline__ void issue_loads(const Unit& U, int k, int lane, uint4 (&wr)[GV_NV][GV_NR]) {  // = WRegs
#pragma unroll
  for (int v = 0; v < GV_NV; ++v) {
    const int kk = k + v * 256 + lane * 8;
#pragma unroll
    for (int i = 0; i < GV_NR; ++i) wr[v][i] = kk < U.k1 ? ldg_stream(U.w[i] + kk) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// acc[i][r] += sum over this lane's elements of chunk k of w_i * a_r; a_s: bf16 [R][pitch] in shared memory whose
// column 0 is K index abase
template <int R>
__device__ __forceinline__ void consume(const uint4 (&wr)[GV_NV][GV_NR], const __nv_bfloat16* a_s, int pitch, int abase,
                                        int k, int k1, int lane, float (&acc)[GV_NR][R]) {
#pragma unroll
  for (int v = 0; v < GV_NV; ++v) {
    const int kk = k + v * 256 + lane * 8;
    if (kk < k1) {
      float wf[GV_NR][8];
#pragma unroll
      for (int i = 0; i < GV_NR; ++i) unpack8(wr[v][i], wf[i]);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float af[8];
        unpack8(*reinterpret_cast<const uint4*>(a_s + (size_t)r * pitch + (kk - abase)), af);
#pragma unroll
        for (int i = 0; i < GV_NR; ++i) {
          float s = acc[i][r];
#pragma unroll
          for (int e = 0; e < 8; ++e) s = fmaf(wf[i][e], af[e], s);
          acc[i][r] = s;
        }
      }
    }
  }
}

// warp-reduces acc[i][r] over the lanes and hands activation row r's value of weight row i to lane r (r < R <= 32)
template <int R>
__device__ __forceinline__ void reduce_to_lane(float (&acc)[GV_NR][R], int lane, float (&mine)[GV_NR]) {
#pragma unroll
  for (int i = 0; i < GV_NR; ++i) {
    mine[i] = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float t = warp_sum(acc[i][r]);
      if (lane == r) mine[i] = t;
    }
  }
}

// RMSNorm of the M live rows of x into shared memory as bf16 (HF Qwen2RMSNorm: fp32 statistics; the normalised row is
// the bf16 GEMM operand of the per-kernel schedule as well). All 16 warps: warp w owns part w / R of the columns of row
// w % R, x is read ONCE (registers), the partial sums of squares meet in shared memory. he <= 512 * (16 / R).
constexpr float GV_FIX = 1073741824.f;        // 2^30: fixed-point scale of ExpertPass::xfix
constexpr float GV_FIX_INV = 1.f / 1073741824.f;

template <int R>
__device__ __forceinline__ void norm_rows_to_smem(const float* __restrict__ x, const long long* __restrict__ xfix,
                                                  const float* __restrict__ gamma, int M,
                                                  int he, float eps, __nv_bfloat16* a_s, float* part_ss, int warp, int lane) {
  constexpr int PARTS = GV_WARPS / R;
  constexpr int NI = R;  // float2 slots per lane: he / PARTS / 64 <= R for he <= 1024
  const int r = warp % R, part = warp / R, plen = he / PARTS;
  const float* xr = x + (size_t)r * he + part * plen;
  const long long* fr = xfix ? xfix + (size_t)r * he + part * plen : nullptr;
  const float* gr = gamma + part * plen;
  float2 xv[NI], gv[NI];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = lane * 2 + 64 * i;
    xv[i] = make_float2(0.f, 0.f), gv[i] = xv[i];
    if (c < plen && r < M) {
      xv[i] = __ldcg(reinterpret_cast<const float2*>(xr + c));
      if (fr) {
        const longlong2 f = __ldcg(reinterpret_cast<const longlong2*>(fr + c));
        xv[i].x += (float)f.x * GV_FIX_INV;
        xv[i].y += (float)f.y * GV_FIX_INV;
      }
      gv[i] = __ldg(reinterpret_cast<const float2*>(gr + c));
    }
  }
#pragma unroll
  for (int i = 0; i < NI; ++i) q += xv[i].x * xv[i].x + xv[i].y * xv[i].y;
  q = warp_sum(q);
  if (lane == 0) part_ss[warp] = q;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int pp = 0; pp < PARTS; ++pp) tot += part_ss[pp * R + r];
  const float rstd = rsqrtf(tot / (float)he + eps);
  __nv_bfloat16* dst = a_s + (size_t)r * he + part * plen;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = lane * 2 + 64 * i;
    if (c < plen) *reinterpret_cast<uint32_t*>(dst + c) = pack_bf16(xv[i].x * rstd * gv[i].x, xv[i].y * rstd * gv[i].y);
  }
}

// bf16 columns [k_lo, k_hi) of the M live rows of src [M, width] (global, written by other CTAs in the previous phase)
// -> shared memory [R][pitch], zero rows beyond M
template <int R>
__device__ __forceinline__ void bf16_rows_to_smem(const __nv_bfloat16* __restrict__ src, int M, int width, int k_lo,
                                                  int k_hi, int pitch, __nv_bfloat16* dst, int tid) {
  const int groups = (k_hi - k_lo) / 8;
  for (int i = tid; i < R * groups; i += GV_THREADS) {
    const int r = i / groups, c = (i - r * groups) * 8;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (r < M) u = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)r * width + k_lo + c));
    *reinterpret_cast<uint4*>(dst + (size_t)r * pitch + c) = u;
  }
}

// ---- attention of the q_len rows of ONE query head over kv_len cached rows, by the whole CTA. The K/V rows of the
// prefix (and, in a denoise step, of the proprio token) date from the prefill: the CTA fetches them into registers
// BEFORE the grid barrier that publishes this layer's q and its M fresh K/V rows (attn_prefetch), so the phase itself
// costs one L2 round trip (q + the fresh rows) instead of three dependent ones.
constexpr int GV_KIT = 2;    // score sweeps held in registers (64 kv rows each)
constexpr int GV_VT = 16;    // PV steps held in registers (GV_JS kv rows each)
static_assert(GV_KIT * (GV_THREADS / 8) == GV_VT * GV_JS, "register-resident kv ranges of the two products differ");
constexpr int GV_KV_REG = GV_VT * GV_JS;  // kv rows covered by the register-resident part (128)

// The register-resident K/V rows live in the weight-load registers `wr` (dead during the attention phase of a CTA that
// runs attention): K sweeps in wr[it][0..1], the 16 V words in wr[2][0..3].
static_assert(GV_NV >= 3 && GV_NR >= 4 && GV_KIT == 2 && GV_VT == 16, "K/V register overlay does not fit the weight registers");
typedef uint4 WRegs[GV_NV][GV_NR];
__device__ __forceinline__ uint32_t& v_slot(WRegs& wr, int t) {
  uint4& q = wr[2][t >> 2];
  return (t & 3) == 0 ? q.x : (t & 3) == 1 ? q.y : (t & 3) == 2 ? q.z : q.w;
}
struct Seg {
  int q_len, kv_len, kv0, qs0;
};

// score product: 8 lanes per kv row, lane l8 owns head dims {32 c + 4 l8 .. + 3 : c = 0..3} (the 8 lanes of a row read
// 128 contiguous bytes of q from shared memory per step: no bank conflicts)
__device__ __forceinline__ void load_k_row(const __nv_bfloat16* K, int j, int l8, uint4& a, uint4& b) {
  const uint2* kr = reinterpret_cast<const uint2*>(K + (size_t)j * 128 + l8 * 4);
  const uint2 c0 = __ldcg(kr), c1 = __ldcg(kr + 8), c2 = __ldcg(kr + 16), c3 = __ldcg(kr + 24);
  a = make_uint4(c0.x, c0.y, c1.x, c1.y);
  b = make_uint4(c2.x, c2.y, c3.x, c3.y);
}
__device__ __forceinline__ void attn_load_rows(const __nv_bfloat16* K, const __nv_bfloat16* V, int lo, int hi, int tid,
                                               WRegs& wr, bool clear) {
  const int l8 = tid & 7, jrow = tid >> 3;
#pragma unroll
  for (int it = 0; it < GV_KIT; ++it) {
    const int j = jrow + (GV_THREADS / 8) * it;
    if (j >= lo && j < hi) {
      load_k_row(K, j, l8, wr[it][0], wr[it][1]);
    } else if (clear) {
      wr[it][0] = wr[it][1] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  const int d2 = tid & 63, js = tid >> 6;
#pragma unroll
  for (int t = 0; t < GV_VT; ++t) {
    const int j = js + GV_JS * t;
    if (j >= lo && j < hi) v_slot(wr, t) = __ldcg(reinterpret_cast<const uint32_t*>(V + (size_t)j * 128) + d2);
    else if (clear) v_slot(wr, t) = 0u;
  }
}

// qrow: this lane's first float4 of the q row (q + 4 l8); step c is 32 floats further
__device__ __forceinline__ float score_part(const uint4& u0, const uint4& u1, const float* qrow) {
  float kf[16];
  {
    float t[8];
    unpack8(u0, t);
#pragma unroll
    for (int e = 0; e < 8; ++e) kf[e] = t[e];
    unpack8(u1, t);
#pragma unroll
    for (int e = 0; e < 8; ++e) kf[8 + e] = t[e];
  }
  float a = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 qv = *reinterpret_cast<const float4*>(qrow + 32 * c);
    a += kf[4 * c] * qv.x + kf[4 * c + 1] * qv.y + kf[4 * c + 2] * qv.z + kf[4 * c + 3] * qv.w;
  }
  return a;
}

__device__ __forceinline__ void gv_stamp(long long* st, int slot);
template <int R>
__device__ __forceinline__ void attention_head(const ExpertPass& p, const Seg& sg, int M, int h, const __nv_bfloat16* K,
                                               const __nv_bfloat16* V, WRegs& wr, float* f_s, int kvpad, int tid,
                                               int warp, int lane, long long* st) {
  const int qdim = p.n_q_heads * 128;
  const int q_len = sg.q_len, kv_len = sg.kv_len;
  float* qs = f_s;                        // [R][128]
  float* sc = qs + R * 128;               // [R][kvpad] scores, then bf16-rounded probabilities
  float* red = sc + (size_t)R * kvpad;    // [GV_JS][R][128] partial PV sums
  float* lsum = red + GV_JS * R * 128;    // [R]
  // one L2 round trip: q and the M fresh K/V rows of this layer
  constexpr int QX = (R * 128 + GV_THREADS - 1) / GV_THREADS;
  float qv[QX];
#pragma unroll
  for (int x = 0; x < QX; ++x) {
    const int i = tid + x * GV_THREADS, r = i >> 7, dd = i & 127;
    qv[x] = (i < R * 128 && r < q_len) ? __bfloat162float(__ldcg(p.q + (size_t)(sg.qs0 + r) * qdim + h * 128 + dd)) : 0.f;
  }
  attn_load_rows(K, V, kv_len - M, kv_len, tid, wr, false);
#pragma unroll
  for (int x = 0; x < QX; ++x)
    if (tid + x * GV_THREADS < R * 128) qs[tid + x * GV_THREADS] = qv[x];
  __syncthreads();
  gv_stamp(st, 5);
  // scores: 8 lanes per kv row (16 head dims each), 64 rows per sweep; sc is laid out [kv row][R]
  {
    const int l8 = tid & 7, jrow = tid >> 3;
    const float sl2 = p.scale * 1.4426950408889634f;
    {
      // the two register-resident sweeps share every q fragment read from shared memory
      float k0[16], k1[16];
      {
        float t[8];
        unpack8(wr[0][0], t);
#pragma unroll
        for (int e = 0; e < 8; ++e) k0[e] = t[e];
        unpack8(wr[0][1], t);
#pragma unroll
        for (int e = 0; e < 8; ++e) k0[8 + e] = t[e];
        unpack8(wr[1][0], t);
#pragma unroll
        for (int e = 0; e < 8; ++e) k1[e] = t[e];
        unpack8(wr[1][1], t);
#pragma unroll
        for (int e = 0; e < 8; ++e) k1[8 + e] = t[e];
      }
      const int j0 = jrow, j1 = jrow + GV_THREADS / 8;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 qv = *reinterpret_cast<const float4*>(qs + r * 128 + l8 * 4 + 32 * c);
          a0 += k0[4 * c] * qv.x + k0[4 * c + 1] * qv.y + k0[4 * c + 2] * qv.z + k0[4 * c + 3] * qv.w;
          a1 += k1[4 * c] * qv.x + k1[4 * c + 1] * qv.y + k1[4 * c + 2] * qv.z + k1[4 * c + 3] * qv.w;
        }
        a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
        a0 += __shfl_xor_sync(0xffffffffu, a0, 2);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
        a0 += __shfl_xor_sync(0xffffffffu, a0, 4);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 4);
        if (l8 == 0 && j0 < kv_len) sc[(size_t)j0 * R + r] = a0 * sl2;
        if (l8 == 0 && j1 < kv_len) sc[(size_t)j1 * R + r] = a1 * sl2;
      }
    }
    for (int it = GV_KIT; it * (GV_THREADS / 8) < kv_len; ++it) {  // kv rows beyond the register-resident 128
      const int j = jrow + (GV_THREADS / 8) * it;
      const bool valid = j < kv_len;
      uint4 u0 = make_uint4(0u, 0u, 0u, 0u), u1 = u0;
      if (valid) load_k_row(K, j, l8, u0, u1);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float a = score_part(u0, u1, qs + r * 128 + l8 * 4);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (valid && l8 == 0) sc[(size_t)j * R + r] = a * sl2;
      }
    }
  }
  __syncthreads();
  gv_stamp(st, 6);
  // softmax (exp2 domain), warp r owns row r; the probabilities are rounded to bf16 as the PV operand of the
  // tensor-core path is, the row sum stays fp32
  for (int r = warp; r < R; r += GV_WARPS) {
    float* row = sc + r;
    float mx = -INFINITY;
    for (int j = lane; j < kv_len; j += 32) mx = fmaxf(mx, row[(size_t)j * R]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < kv_len; j += 32) {
      const float pv = exp2f(row[(size_t)j * R] - mx);
      sum += pv;
      row[(size_t)j * R] = round_bf16(pv);
    }
    sum = warp_sum(sum);
    if (lane == 0) lsum[r] = sum;
  }
  __syncthreads();
  gv_stamp(st, 7);
  // PV: thread = (pair of head dims, kv split)
  {
    const int d2 = tid & 63, js = tid >> 6;
    float o[R][2];
#pragma unroll
    for (int r = 0; r < R; ++r) o[r][0] = o[r][1] = 0.f;
    auto pv_step = [&](int j, uint32_t vw) {
      const float v0 = bf16_lo(vw), v1 = bf16_hi(vw);
      float pj[R];
      if constexpr (R % 4 == 0) {
#pragma unroll
        for (int r4 = 0; r4 < R / 4; ++r4) {
          const float4 t = *reinterpret_cast<const float4*>(sc + (size_t)j * R + 4 * r4);
          pj[4 * r4] = t.x, pj[4 * r4 + 1] = t.y, pj[4 * r4 + 2] = t.z, pj[4 * r4 + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) pj[r] = sc[(size_t)j * R + r];
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        o[r][0] = fmaf(pj[r], v0, o[r][0]);
        o[r][1] = fmaf(pj[r], v1, o[r][1]);
      }
    };
#pragma unroll
    for (int t = 0; t < GV_VT; ++t) {
      const int j = js + GV_JS * t;
      if (j < kv_len) pv_step(j, v_slot(wr, t));
    }
    const uint32_t* Vw = reinterpret_cast<const uint32_t*>(V) + d2;
#pragma unroll 4
    for (int j = GV_KV_REG + js; j < kv_len; j += GV_JS) pv_step(j, __ldcg(Vw + (size_t)j * 64));
#pragma unroll
    for (int r = 0; r < R; ++r) *reinterpret_cast<float2*>(red + ((size_t)js * R + r) * 128 + 2 * d2) = make_float2(o[r][0], o[r][1]);
  }
  __syncthreads();
  gv_stamp(st, 8);
  for (int i = tid; i < R * 64; i += GV_THREADS) {
    const int r = i >> 6, d2 = i & 63;
    if (r < q_len) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int js = 0; js < GV_JS; ++js) {
        const float2 t = *reinterpret_cast<const float2*>(red + ((size_t)js * R + r) * 128 + 2 * d2);
        a0 += t.x, a1 += t.y;
      }
      const float inv = lsum[r] > 0.f ? 1.f / lsum[r] : 0.f;
      *reinterpret_cast<uint32_t*>(p.ao + (size_t)(sg.qs0 + r) * qdim + h * 128 + 2 * d2) = pack_bf16(a0 * inv, a1 * inv);
    }
  }
}

// profiling aid: fine-grained stamps of CTA 0 / thread 0 in layer 1 go to the unused tail of the phase-timing row
__device__ __forceinline__ void gv_stamp(long long* st, int slot) {
  if (st) {
    long long now;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
    st[slot] = now;
  }
}

// fire-and-forget HBM -> L2 prefetch of slice idx of n of a byte range (bulk prefetches of <= 32 KB through the TMA
// unit: a handful of instructions by ONE thread, nothing queues in the SM's load/store pipe). bytes % 16 == 0.
__device__ __forceinline__ void l2_prefetch(const void* base, size_t bytes, int idx, int n) {
  const size_t slice = ((bytes + n - 1) / n + 127) & ~(size_t)127;
  const size_t lo = (size_t)idx * slice, hi = min(bytes, lo + slice);
  const char* c = reinterpret_cast<const char*>(base);
  for (size_t off = lo; off < hi; off += 32768) {
    const unsigned int sz = (unsigned int)min((size_t)32768, hi - off);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(c + off), "r"(sz) : "memory");
  }
}

template <int R>
__global__ void __launch_bounds__(GV_THREADS, 1) expert_gemv_pass_kernel(const ExpertPass p) {
  extern __shared__ __align__(16) unsigned char gv_smem[];
  __nv_bfloat16* a_s = reinterpret_cast<__nv_bfloat16*>(gv_smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Dims d;
  d.he = p.he, d.mlp = p.mlp, d.qdim = p.n_q_heads * 128, d.heads = p.n_q_heads + 2 * p.n_kv_heads;
  d.G = (int)gridDim.x, d.b = (int)blockIdx.x, d.warp = warp;
  const int nw = d.G * GV_WARPS;
  const int M = p.M;
  const int n_phases = (p.n_layers - 1) * GV_PHASES + (p.last_kv_only ? 1 : GV_PHASES);
  const int kvpad = (p.max_kv_len + 3) & ~3;
  const bool skip_attn = p.flags & 2, skip_units = p.flags & 4, no_l2pf = p.flags & 16;  // profiling aids
  const bool attn_cta = d.b < p.n_q_heads;
  const bool pf_thread = tid == 32 && !no_l2pf;  // the CTA's prefetch issuer (not thread 0: that one arrives and polls)
  const size_t qkv_bytes = (size_t)d.heads * 128 * d.he * 2, o_bytes = (size_t)d.he * d.qdim * 2;
  const size_t gu_bytes = (size_t)2 * d.mlp * d.he * 2, down_bytes = (size_t)d.he * d.mlp * 2;

  // everything up to the dependent-launch wait touches constants only: the layer table, the first unit's weight
  // loads and the L2 prefetch of layer 0's small matrices and parameters
  __shared__ LayerPtrs lt[GV_MAX_LAYERS];
  __shared__ float part_ss[GV_WARPS];
  if (tid < p.n_layers) {
    const ExpertLayerDev* e = p.layers + tid;
    LayerPtrs t;
    t.w_qkv = reinterpret_cast<const __nv_bfloat16*>(e->w_qkv), t.w_o = reinterpret_cast<const __nv_bfloat16*>(e->w_o);
    t.w_gu = reinterpret_cast<const __nv_bfloat16*>(e->w_gu), t.w_down = reinterpret_cast<const __nv_bfloat16*>(e->w_down);
    t.ln_in = e->ln_in, t.ln_post = e->ln_post, t.bqkv = e->bqkv;
    lt[tid] = t;
  }
  __syncthreads();
  WRegs wr;
  Unit U;
  bool have = !skip_units && make_unit(d, lt, 0, first_unit(d, 0), U);
  if (have) issue_loads(U, U.k0, lane, wr);
  if (pf_thread) {
    l2_prefetch(lt[0].w_qkv, qkv_bytes, d.b, d.G);
    l2_prefetch(lt[0].w_o, o_bytes, d.b, d.G);
    if (d.b == d.G - 1) {
      l2_prefetch(lt[0].ln_in, (size_t)d.he * 4, 0, 1);
      l2_prefetch(lt[0].ln_post, (size_t)d.he * 4, 0, 1);
      l2_prefetch(lt[0].bqkv, (size_t)d.heads * 128 * 4, 0, 1);
    }
  }
  pdl_launch_dependents();
  pdl_wait();  // the previous grid's x / K / V / plan are complete
  if (p.timing && blockIdx.x == 0 && tid == 0) {  // profiling aid (option "expert_timing"): pass start
    long long now;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
    p.timing[7 * p.n_layers - 1] = now;
  }
  const int my_pos = lane < M ? p.row_pos[lane] : 0, my_kvrow = lane < M ? p.row_kvrow[lane] : 0;
  Seg sg;
  sg.q_len = p.q_len[0], sg.kv_len = p.kv_len[0], sg.kv0 = p.kv_start[0], sg.qs0 = p.q_start[0];
  const int kvh = attn_cta ? d.b / (p.n_q_heads / p.n_kv_heads) : 0;

  for (int layer = 0; layer < p.n_layers; ++layer) {
    const LayerPtrs* L = lt + layer;
    __nv_bfloat16* kc = p.k_cache + (size_t)layer * p.kv_layer_stride;
    __nv_bfloat16* vc = p.v_cache + (size_t)layer * p.kv_layer_stride;
    const __nv_bfloat16* Kh = kc + (size_t)kvh * p.kv_head_stride + (size_t)sg.kv0 * 128;
    const __nv_bfloat16* Vh = vc + (size_t)kvh * p.kv_head_stride + (size_t)sg.kv0 * 128;
#pragma unroll 1
    for (int j = 0; j < GV_PHASES; ++j) {
      const int ph = layer * GV_PHASES + j;
      if (ph >= n_phases) break;
      if (ph > 0) {
        if (tid == 0) {
          gv_wait(p.barrier + (ph - 1), (unsigned int)d.G, p.err_flag);
          if (p.timing && blockIdx.x == 0) {  // when did the whole grid finish phase ph - 1
            long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            p.timing[ph - 1] = now;
          }
        }
        __syncthreads();
      }
      long long* st = (p.timing && blockIdx.x == 0 && tid == 0 && layer == 1) ? p.timing + GV_PHASES * p.n_layers : nullptr;
      if (j == 0) gv_stamp(st, 0);
      if (j == 1) gv_stamp(st, 4);
      if (j == 3) gv_stamp(st, 10);
      if (j == 1) {
        if (attn_cta && !skip_attn)
          attention_head<R>(p, sg, M, d.b, Kh, Vh, wr, reinterpret_cast<float*>(gv_smem), kvpad, tid, warp, lane, st);
        gv_stamp(st, 9);
      } else {
        const int total = unit_count(d, j);
        int pitch = d.he, abase = 0;
        if (j == 0 || j == 3) {
          norm_rows_to_smem<R>(p.x, p.xfix, j == 0 ? L->ln_in : L->ln_post, M, d.he, p.eps, a_s, part_ss, warp, lane);
        } else if (j == 2) {
          pitch = d.qdim;
          bf16_rows_to_smem<R>(p.ao, M, d.qdim, 0, d.qdim, pitch, a_s, tid);
        } else {
          pitch = 2 * GV_CH;  // 16 consecutive units span at most two K chunks
        }
        for (int base = 0; base < total; base += nw) {
          if (j == 4) {
            // this round's units of the CTA: [base + b * W, + W) -> chunk range -> K slice of the SwiGLU output
            const int cgs = d.he / GV_NR;
            const int u_lo = base + d.b * GV_WARPS, u_hi = min(u_lo + GV_WARPS, total) - 1;
            if (base > 0) __syncthreads();
            if (u_lo <= u_hi) {
              abase = (u_lo / cgs) * GV_CH;
              const int k_hi = min(d.mlp, (u_hi / cgs + 1) * GV_CH);
              bf16_rows_to_smem<R>(p.act, M, d.mlp, abase, k_hi, pitch, a_s, tid);
            }
          }
          if (base == 0 || j == 4) __syncthreads();
          if (base == 0 && j == 0) gv_stamp(st, 1);
          if (base == 0 && j == 3) gv_stamp(st, 11);
          const int u = base + first_unit(d, j);
          const bool pre = base == 0 && have;  // this unit's loads went out before the grid barrier
          const bool ok = pre ? true : (!skip_units && make_unit(d, L, j, u, U));
          if (ok) {
            // q|k|v epilogue operands (bias, RoPE factors of this lane's row) fetched ahead of the dot products
            float eb[4] = {0.f, 0.f, 0.f, 0.f}, ecs[2] = {1.f, 1.f}, esn[2] = {0.f, 0.f};
            if (j == 0 && lane < M) {
              const float* bq = L->bqkv + (u >> 5) * 128 + (u & 31);
              eb[0] = __ldg(bq), eb[1] = __ldg(bq + 64), eb[2] = __ldg(bq + 32), eb[3] = __ldg(bq + 96);
              const size_t ci = (size_t)my_pos * 64 + (u & 31);
              ecs[0] = __ldg(p.cos_tab + ci), esn[0] = __ldg(p.sin_tab + ci);
              ecs[1] = __ldg(p.cos_tab + ci + 32), esn[1] = __ldg(p.sin_tab + ci + 32);
            }
            float acc[GV_NR][R];
#pragma unroll
            for (int i = 0; i < GV_NR; ++i)
#pragma unroll
              for (int r = 0; r < R; ++r) acc[i][r] = 0.f;
            for (int k = U.k0; k < U.k1; k += GV_CH) {
              if (!pre || k != U.k0) issue_loads(U, k, lane, wr);
              consume<R>(wr, a_s, pitch, abase, k, U.k1, lane, acc);
            }
            float mine[GV_NR];
            reduce_to_lane<R>(acc, lane, mine);
            if (lane < M) {
              const int r = lane;
              if (j == 0) {
                const int head = u >> 5, i0 = u & 31;
                const bool is_q = head < p.n_q_heads, is_k = !is_q && head < p.n_q_heads + p.n_kv_heads;
                __nv_bfloat16* dst;
                if (is_q) dst = p.q + (size_t)r * d.qdim + head * 128;
                else {
                  const int kh = is_k ? head - p.n_q_heads : head - p.n_q_heads - p.n_kv_heads;
                  dst = (is_k ? kc : vc) + (size_t)kh * p.kv_head_stride + (size_t)my_kvrow * 128;
                }
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  const int i = i0 + 32 * hf;
                  float v0 = mine[2 * hf] + eb[2 * hf];
                  float v1 = mine[2 * hf + 1] + eb[2 * hf + 1];
                  if (is_q || is_k) {
                    const float cs = ecs[hf], sn = esn[hf];
                    const float x0 = v0, x1 = v1;
                    v0 = x0 * cs - x1 * sn;  // x*cos + rotate_half(x)*sin
                    v1 = x1 * cs + x0 * sn;
                  }
                  dst[i] = __float2bfloat16_rn(v0);
                  dst[i + 64] = __float2bfloat16_rn(v1);
                }
              } else if (j == 3) {
                *reinterpret_cast<uint32_t*>(p.act + (size_t)r * d.mlp + 2 * u) =
                    pack_bf16(silu(mine[0]) * mine[2], silu(mine[1]) * mine[3]);
              } else {
                const int cgs = d.he / GV_NR, cg = u % cgs;
                if (p.xfix) {
                  // integer adds commute: the sum does not depend on the order in which the K chunks arrive
#pragma unroll
                  for (int i = 0; i < GV_NR; ++i)
                    atomicAdd(reinterpret_cast<unsigned long long*>(p.xfix + (size_t)r * d.he + cg * GV_NR + i),
                              (unsigned long long)__float2ll_rn(mine[i] * GV_FIX));
                } else {
#pragma unroll
                  for (int i = 0; i < GV_NR; ++i) atomicAdd(p.x + (size_t)r * d.he + cg * GV_NR + i, mine[i]);
                }
              }
            }
          }
        }
      }
      have = false;
      if (j == 0) gv_stamp(st, 2);
      if (j == 3) gv_stamp(st, 12);
      if (ph + 1 < n_phases) {
        __syncthreads();  // every thread's writes happen-before the CTA's single release
        if (j == 0) gv_stamp(st, 3);
        if (j == 3) gv_stamp(st, 13);
        if (tid == 0) gv_arrive(p.barrier + ph);
        __syncthreads();  // the arrival is in the load/store pipe ahead of everything below
        // ---- while the grid barrier settles: everything of the coming phases that does not depend on it
        const int nj = j == GV_PHASES - 1 ? 0 : j + 1;
        if (nj != 1) {
          // first unit of the next weight phase into registers
          // (profiling flag 32: only the even warps load across the barrier, so that the warps of a CTA alternate between
          // loading and computing instead of all doing the same thing at the same time)
          have = !skip_units && (!(p.flags & 32) || (warp & 1) == 0) && make_unit(d, j == GV_PHASES - 1 ? L + 1 : L, nj, first_unit(d, nj), U);
          if (have) issue_loads(U, U.k0, lane, wr);
        } else if (attn_cta && !skip_attn) {
          attn_load_rows(Kh, Vh, 0, sg.kv_len - M, tid, wr, true);  // K/V rows that date from the prefill
        }
        // HBM -> L2 one to two phases ahead, so that HBM streams while barriers, staging and attention run:
        //   after q|k|v: gate/up of this layer (by the CTAs that do not run attention)
        //   after attention: the next layer's K/V prefix rows of this CTA's kv head
        //   after o: down
        //   after gate/up: q|k|v, o and the norm gains / bias of the next layer
        if (pf_thread) {
          if (j == 0 && !attn_cta) l2_prefetch(L->w_gu, gu_bytes, d.b - p.n_q_heads, d.G - p.n_q_heads);
          else if (j == 1 && attn_cta && layer + 1 < p.n_layers) {
            const size_t rows16 = ((size_t)(sg.kv_len - M) * 256) & ~(size_t)15;
            l2_prefetch(Kh + p.kv_layer_stride, rows16, 0, 1);
            l2_prefetch(Vh + p.kv_layer_stride, rows16, 0, 1);
          } else if (j == 2) l2_prefetch(L->w_down, down_bytes, d.b, d.G);
          else if (j == 3 && layer + 1 < p.n_layers) {
            l2_prefetch((L + 1)->w_qkv, qkv_bytes, d.b, d.G);
            l2_prefetch((L + 1)->w_o, o_bytes, d.b, d.G);
            if (d.b == d.G - 1) {
              l2_prefetch((L + 1)->ln_in, (size_t)d.he * 4, 0, 1);
              l2_prefetch((L + 1)->ln_post, (size_t)d.he * 4, 0, 1);
              l2_prefetch((L + 1)->bqkv, (size_t)d.heads * 128 * 4, 0, 1);
            }
          }
        }
      }
    }
  }
}

template <int R>
int launch_r(const ExpertPass& p, int grid, size_t smem, cudaStream_t stream) {
  static unsigned long long attr_devs = 0;  // devices this kernel's attributes were set on
  if (vl_first_on_device(attr_devs)) {
    VL_CUDA_CHECK(cudaFuncSetAttribute(expert_gemv_pass_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  // the kernel synchronises its CTAs through a spin barrier: every CTA of the grid has to be resident at once
  int per_sm = 0;
  VL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, expert_gemv_pass_kernel<R>, GV_THREADS, smem));
  VL_CHECK(per_sm >= 1 && per_sm * gemm_num_sms() >= grid,
           "expert gemv pass: the persistent grid does not fit on the device (use option expert_gemv = 0)");
  VL_CUDA_CHECK(launch_kernel(expert_gemv_pass_kernel<R>, dim3(grid), dim3(GV_THREADS), smem, stream, p));
  return 0;
}

}  // namespace

int expert_gemv_pass_launch(const ExpertPass& p, cudaStream_t stream) {
  VL_CHECK(p.layers && p.n_layers > 0 && p.M > 0, "expert gemv pass: empty problem");
  VL_CHECK(p.batch == 1 && p.M <= 8, "expert gemv pass: one observation, at most 8 rows");
  VL_CHECK(p.he % 128 == 0 && p.he <= 1024 && p.mlp % 64 == 0, "expert gemv pass: bad widths");
  VL_CHECK(p.barrier != nullptr && p.max_kv_len > 0, "expert gemv pass: null barrier counter / no kv bound");
  VL_CHECK(p.n_layers <= GV_MAX_LAYERS, "expert gemv pass: more layers than the shared-memory layer table holds");
  const int R = p.M <= 1 ? 1 : p.M <= 4 ? 4 : 8;
  const int qdim = p.n_q_heads * 128;
  const int kvpad = (p.max_kv_len + 3) & ~3;
  const size_t act_bytes = (size_t)R * std::max(std::max(p.he, qdim), 2 * GV_CH) * sizeof(__nv_bfloat16);
  const size_t attn_bytes = ((size_t)R * 128 + (size_t)R * kvpad + (size_t)GV_JS * R * 128 + R) * sizeof(float);
  const size_t smem = std::max(act_bytes, attn_bytes);
  VL_CHECK(smem <= 200 * 1024, "expert gemv pass: activations do not fit in shared memory");
  const int grid = gemm_num_sms();  // one CTA per SM: all resident (the grid barrier needs that)
  VL_CHECK(grid >= p.n_q_heads, "expert gemv pass: fewer CTAs than query heads");
  if (R == 1) return launch_r<1>(p, grid, smem, stream);
  if (R == 4) return launch_r<4>(p, grid, smem, stream);
  return launch_r<8>(p, grid, smem, stream);
}

}  // namespace vlaser
