// Flash-style fused attention (see attention.cuh). CTA = 4 warps x 16 query rows = 64 "rows",
// where a row is a (token, q-head-within-group) pair so that the `group` query heads sharing one
// KV head are packed into the same tile (GQA packing: K/V tiles are loaded once per kv head).
// K/V stream through a 2-stage cp.async ring in 64-row blocks; QK^T and PV run on mma.sync
// m16n8k16 bf16 with fp32 accumulation, softmax statistics in fp32 (exp2 domain).
// Attention is <3% of the FLOPs of this path (sequences are <= ~900 tokens); the tensor-core
// heavy lifting is in gemm.cu.
#include "attention.cuh"
#include "attention_device.cuh"
#include "gemm.cuh"  // gemm_num_sms()

namespace vlaser {

namespace {

template <int HD, bool CAUSAL = false>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(AttnParams p) {
  pdl_launch_dependents();
  extern __shared__ __align__(128) uint8_t smem[];
  long long* tr = (blockIdx.x + blockIdx.y + blockIdx.z == 0 && threadIdx.x == 0) ? p.trace : nullptr;
  trace_stamp(tr, 0);
  attention_tile<HD, CAUSAL>(p, blockIdx.z, blockIdx.y, blockIdx.x, smem, threadIdx.x, [] { __syncthreads(); },
                     [tr] { pdl_wait(); trace_stamp(tr, 1); });
  trace_stamp(tr, 6);
}

// ------------------------------------------------------------------------------------------
// InternViT self-attention, one CTA per (image, head): N = P+1 tokens (257 at 224 px), head_dim 64, no mask
// (modeling_intern_vit.py:210-227). K and V of the head are staged in shared memory ONCE (the generic kernel
// re-reads them for every 64-row query block), 8 warps sweep the 16-row query tiles, the ragged tail of N
// (257 = 4 x 64 + 1 keys, 16 x 16 + 1 query rows) is handled with partial MMA tiles for the keys and a
// cooperative CUDA-core pass for the left-over query rows instead of whole padded blocks.
// ------------------------------------------------------------------------------------------
constexpr int VA_THREADS = 256;
constexpr int VA_MAX_N = 320;

__global__ void __launch_bounds__(VA_THREADS, 2)
vit_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int N, int heads,
                     float scale) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int HD = 64, CH = 8;
  extern __shared__ __align__(128) uint8_t smem[];
  const int NR = (N + 15) & ~15;          // rows staged (zero-filled past N)
  uint8_t* Qs = smem;                      // [NR][64] bf16, 16-byte chunks XOR-swizzled with the row
  uint8_t* Ks = Qs + NR * 128;
  uint8_t* Vs = Ks + NR * 128;
  float* ps = reinterpret_cast<float*>(Vs + NR * 128);   // [NR] probabilities of the CUDA-core pass
  float* red = ps + NR;                                   // [4][64] + [16] scratch

  const int h = blockIdx.x, img = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ld = 3 * heads * HD;
  const __nv_bfloat16* base = qkv + (size_t)img * N * ld + h * HD;
  const int koff = heads * HD, voff = 2 * heads * HD;

  // ---- stage: group 0 = Q (all rows) + K/V rows [0,64); group 1 = the remaining K/V rows
  const int first = min(64, NR);
  for (int i = tid; i < NR * CH; i += VA_THREADS) {
    const int row = i >> 3, c = i & 7;
    const bool ok = row < N;
    cp_async_16(Qs + swz<HD>(row, c), base + (size_t)(ok ? row : 0) * ld + c * 8, ok);
  }
  for (int i = tid; i < first * CH; i += VA_THREADS) {
    const int row = i >> 3, c = i & 7;
    const bool ok = row < N;
    const __nv_bfloat16* src = base + (size_t)(ok ? row : 0) * ld + c * 8;
    cp_async_16(Ks + swz<HD>(row, c), src + koff, ok);
    cp_async_16(Vs + swz<HD>(row, c), src + voff, ok);
  }
  cp_async_commit();
  for (int i = tid + first * CH; i < NR * CH; i += VA_THREADS) {
    const int row = i >> 3, c = i & 7;
    const bool ok = row < N;
    const __nv_bfloat16* src = base + (size_t)(ok ? row : 0) * ld + c * 8;
    cp_async_16(Ks + swz<HD>(row, c), src + koff, ok);
    cp_async_16(Vs + swz<HD>(row, c), src + voff, ok);
  }
  cp_async_commit();

  const float sl2 = scale * 1.4426950408889634f;
  const uint32_t qs_u = smem_u32(Qs), ks_u = smem_u32(Ks), vs_u = smem_u32(Vs);
  const int n_tiles = N >> 4;              // full 16-row query tiles (MMA path)
  const int n_kb = (N + 63) >> 6;          // 64-key blocks, the last one possibly ragged
  bool staged_all = false;

  // few images (B = 1, 2): gridDim.z CTAs share the query tiles of one (image, head), so that a 16-head single image
  // is not left to 16 SMs with two sweeps each
  for (int tile = warp + 8 * blockIdx.z; tile < n_tiles || !staged_all; tile += 8 * gridDim.z) {
    const bool active = tile < n_tiles;   // every warp walks the staging barriers of the first sweep
    uint32_t qf[HD / 16][4];
    float o[HD / 8][4];
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

    for (int jb = 0; jb < n_kb; ++jb) {
      if (!staged_all && jb < 2) {
        if (jb == 0) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncthreads();
        if (jb == 1 || n_kb == 1) staged_all = true;
      }
      if (!active) continue;
      if (jb == 0) {
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const int mi = lane >> 3;
          const int row = tile * 16 + (mi & 1) * 8 + (lane & 7);
          ldmatrix_x4(qf[kk], qs_u + swz<HD>(row, kk * 2 + (mi >> 1)));
        }
      }
      const int keys = min(64, N - jb * 64);       // live keys of this block
      const int nt16 = (keys + 15) >> 4;            // 16-key groups to compute (zero rows pad the last)
      // ---- S = Q K^T
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        if (jp < nt16) {
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            uint32_t kf[4];
            const int mi = lane >> 3;
            const int row = jb * 64 + jp * 16 + (mi >> 1) * 8 + (lane & 7);
            ldmatrix_x4(kf, ks_u + swz<HD>(row, kk * 2 + (mi & 1)));
            mma_bf16_16816(s[2 * jp], qf[kk], kf[0], kf[1]);
            mma_bf16_16816(s[2 * jp + 1], qf[kk], kf[2], kf[3]);
          }
        }
      }
      // ---- online softmax (exp2 domain); only a ragged block needs the key mask
      const bool ragged = keys < 64;
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float val = s[j][e] * sl2;
          if (ragged && (j * 8 + (lane & 3) * 2 + (e & 1)) >= keys) val = -INFINITY;
          s[j][e] = val;
          mx[e >> 1] = fmaxf(mx[e >> 1], val);
        }
      }
      float alpha[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
        const float m_new = fmaxf(m_run[hh], mx[hh]);   // finite: every block holds at least one live key
        alpha[hh] = ex2_approx(m_run[hh] - m_new);
        m_run[hh] = m_new;
      }
      float rs[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = ex2_approx(s[j][e] - m_run[e >> 1]);
          s[j][e] = pv;
          rs[e >> 1] += pv;
        }
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        rs[hh] += __shfl_xor_sync(0xffffffffu, rs[hh], 1);
        rs[hh] += __shfl_xor_sync(0xffffffffu, rs[hh], 2);
        l_run[hh] = l_run[hh] * alpha[hh] + rs[hh];
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
        if (ks < nt16) {
          uint32_t pf[4];
          pf[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
          pf[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
          pf[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
          pf[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
          for (int dp = 0; dp < HD / 16; ++dp) {
            uint32_t vf[4];
            const int mi = lane >> 3;
            const int row = jb * 64 + ks * 16 + (mi & 1) * 8 + (lane & 7);
            ldmatrix_x4_trans(vf, vs_u + swz<HD>(row, dp * 2 + (mi >> 1)));
            mma_bf16_16816(o[2 * dp], pf, vf[0], vf[1]);
            mma_bf16_16816(o[2 * dp + 1], pf, vf[2], vf[3]);
          }
        }
      }
    }
    if (!active) continue;
    // ---- normalise, transpose through this tile's (now dead) Q rows, store 16-byte chunks coalesced
    const int r_lo = tile * 16 + (lane >> 2);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int row = r_lo + hh * 8;
      const float inv = 1.f / l_run[hh];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i)
        *reinterpret_cast<uint32_t*>(Qs + swz<HD>(row, i) + (lane & 3) * 4) =
            pack_bf16(o[i][2 * hh] * inv, o[i][2 * hh + 1] * inv);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = tile * 16 + it * 4 + (lane >> 3), c = lane & 7;
      const uint4 u = *reinterpret_cast<const uint4*>(Qs + swz<HD>(row, c));
      *reinterpret_cast<uint4*>(out + ((size_t)img * N + row) * (heads * HD) + h * HD + c * 8) = u;
    }
  }

  // ---- left-over query rows (N % 16 of them; the single CLS/patch row at N = 257): CUDA-core pass, all threads
  for (int r = n_tiles * 16; r < N && blockIdx.z == gridDim.z - 1; ++r) {
    __syncthreads();
    // scores: thread t owns keys t, t + 256
    float sc[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int j = tid + rep * VA_THREADS;
      if (j < N) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const uint4 qv = *reinterpret_cast<const uint4*>(Qs + swz<HD>(r, c));
          const uint4 kv = *reinterpret_cast<const uint4*>(Ks + swz<HD>(j, c));
          const uint32_t qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) acc += bf16_lo(qa[e]) * bf16_lo(ka[e]) + bf16_hi(qa[e]) * bf16_hi(ka[e]);
        }
        sc[rep] = acc * sl2;
      }
    }
    float mx = warp_max(fmaxf(sc[0], sc[1]));
    if (lane == 0) red[256 + warp] = mx;
    __syncthreads();
    mx = red[256];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[256 + w]);
    float sum = 0.f;
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int j = tid + rep * VA_THREADS;
      if (j < N) {
        const float p = ex2_approx(sc[rep] - mx);
        ps[j] = p;
        sum += p;
      }
    }
    sum = warp_sum(sum);
    if (lane == 0) red[264 + warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[264 + w];
    // output: thread (part, d) accumulates keys j = part, part + 4, ...
    const int d = tid & 63, part = tid >> 6;
    float acc = 0.f;
    for (int j = part; j < N; j += 4)
      acc += ps[j] * __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(Vs + swz<HD>(j, d >> 3) + (d & 7) * 2));
    red[part * 64 + d] = acc;
    __syncthreads();
    if (tid < 64) {
      const float v = (red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid]) / sum;
      out[((size_t)img * N + r) * (heads * HD) + h * HD + tid] = __float2bfloat16(v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Decode-step attention (see DecodeAttnParams): split-KV, CUDA cores -- a step has `group` query rows per kv head, far
// too few for an MMA tile, and is bound by streaming the K / V window once.
// ------------------------------------------------------------------------------------------
constexpr int DA_THREADS = 128;
constexpr int DA_MAX_GROUP = 8;
constexpr int DA_MAX_CHUNK = 256;

__global__ void __launch_bounds__(DA_THREADS) decode_attention_split_kernel(DecodeAttnParams p) {
  extern __shared__ __align__(128) uint8_t da_smem[];         // K chunk, V chunk: [chunk][256 B], 16-byte chunks XOR-swizzled
  __shared__ __align__(16) float qs[DA_MAX_GROUP][128];       // queries, scaled into the exp2 domain
  __shared__ float ps[DA_MAX_GROUP][DA_MAX_CHUNK];            // scores, then probabilities
  __shared__ float red[DA_MAX_GROUP][4];
  __shared__ float oacc[DA_MAX_GROUP][128];
  pdl_launch_dependents();
  pdl_wait();
  const int sp = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kv_len = p.kv_len[b];
  const int k0 = sp * p.chunk;
  if (k0 >= kv_len) return;
  const int n = min(p.chunk, kv_len - k0);
  const int G = p.group;
  const float sl2 = p.scale * 1.4426950408889634f;
  uint8_t* Ks = da_smem;
  uint8_t* Vs = da_smem + (size_t)p.chunk * 256;
  // ---- the whole K / V chunk in flight at once (a CTA is bound by HBM latency, not by arithmetic)
  const size_t kvbase = (size_t)kvh * p.kv_head_stride + (size_t)(p.kv_start[b] + k0) * p.kv_row_stride;
  for (int i = tid; i < n * 16; i += DA_THREADS) {
    const int row = i >> 4, c = i & 15;
    cp_async_16(Ks + swz<128>(row, c), p.k + kvbase + (size_t)row * p.kv_row_stride + c * 8, true);
  }
  cp_async_commit();
  for (int i = tid; i < n * 16; i += DA_THREADS) {
    const int row = i >> 4, c = i & 15;
    cp_async_16(Vs + swz<128>(row, c), p.v + kvbase + (size_t)row * p.kv_row_stride + c * 8, true);
  }
  cp_async_commit();
  const __nv_bfloat16* qrow = p.q + (size_t)p.q_start[b] * p.q_row_stride + (size_t)kvh * G * 128;
  for (int i = tid; i < G * 128; i += DA_THREADS) qs[i >> 7][i & 127] = __bfloat162float(qrow[i]) * sl2;
  cp_async_wait<1>();  // K landed
  __syncthreads();
  // ---- scores: one key per thread and round
  for (int j = tid; j < p.chunk; j += DA_THREADS) {
    float acc[DA_MAX_GROUP];
#pragma unroll
    for (int g = 0; g < DA_MAX_GROUP; ++g) acc[g] = 0.f;
    if (j < n) {
#pragma unroll 4
      for (int c = 0; c < 16; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(Ks + swz<128>(j, c));
        const float kf[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
        for (int g = 0; g < DA_MAX_GROUP; ++g) {
          if (g < G) {
            const float4 q0 = *reinterpret_cast<const float4*>(&qs[g][c * 8]);
            const float4 q1 = *reinterpret_cast<const float4*>(&qs[g][c * 8 + 4]);
            acc[g] += (q0.x * kf[0] + q0.y * kf[1]) + (q0.z * kf[2] + q0.w * kf[3]) + (q1.x * kf[4] + q1.y * kf[5]) +
                      (q1.z * kf[6] + q1.w * kf[7]);
          }
        }
      }
    }
#pragma unroll
    for (int g = 0; g < DA_MAX_GROUP; ++g)
      if (g < G) ps[g][j] = j < n ? acc[g] : -INFINITY;
  }
  cp_async_wait<0>();  // V landed
  __syncthreads();
  // ---- chunk maximum and probabilities: warp w owns rows w, w + 4
  for (int g = warp; g < G; g += 4) {
    float mx = -INFINITY;
    for (int j = lane; j < p.chunk; j += 32) mx = fmaxf(mx, ps[g][j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < p.chunk; j += 32) {
      const float pj = exp2f(ps[g][j] - mx);  // -inf -> 0
      ps[g][j] = pj;
      sum += pj;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[g][0] = mx, red[g][1] = sum;
  }
  __syncthreads();
  // ---- P V: thread = (two dims, key parity)
  const int d2 = (tid & 63) * 2, par = tid >> 6;
  float o[DA_MAX_GROUP][2];
#pragma unroll
  for (int g = 0; g < DA_MAX_GROUP; ++g) o[g][0] = o[g][1] = 0.f;
#pragma unroll 4
  for (int j = par; j < n; j += 2) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(Vs + swz<128>(j, d2 >> 3) + (d2 & 7) * 2);
    const float v0 = bf16_lo(u), v1 = bf16_hi(u);
#pragma unroll
    for (int g = 0; g < DA_MAX_GROUP; ++g)
      if (g < G) {
        const float pj = ps[g][j];
        o[g][0] += pj * v0;
        o[g][1] += pj * v1;
      }
  }
  if (par == 1) {
#pragma unroll
    for (int g = 0; g < DA_MAX_GROUP; ++g)
      if (g < G) oacc[g][d2] = o[g][0], oacc[g][d2 + 1] = o[g][1];
  }
  __syncthreads();
  const int splits = gridDim.x;
  const size_t wbase = ((size_t)(b * p.n_kv_heads + kvh) * splits + sp) * G;
  if (par == 0) {
#pragma unroll
    for (int g = 0; g < DA_MAX_GROUP; ++g)
      if (g < G)
        *reinterpret_cast<float2*>(p.ws_o + (wbase + g) * 128 + d2) = make_float2(o[g][0] + oacc[g][d2], o[g][1] + oacc[g][d2 + 1]);
  }
  if (tid < G) *reinterpret_cast<float2*>(p.ws_ml + (wbase + tid) * 2) = make_float2(red[tid][0], red[tid][1]);
}

__global__ void __launch_bounds__(128) decode_attention_merge_kernel(DecodeAttnParams p, int splits) {
  pdl_launch_dependents();
  pdl_wait();
  const int head = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  const int kvh = head / p.group, g = head - kvh * p.group;
  const int ns = (p.kv_len[b] + p.chunk - 1) / p.chunk;
  const size_t wbase = (size_t)(b * p.n_kv_heads + kvh) * splits;
  float M = -INFINITY;
  for (int s = 0; s < ns; ++s) M = fmaxf(M, p.ws_ml[((wbase + s) * p.group + g) * 2]);
  float L = 0.f, o = 0.f;
  for (int s = 0; s < ns; ++s) {  // fixed order: bit-identical from call to call
    const float2 ml = *reinterpret_cast<const float2*>(p.ws_ml + ((wbase + s) * p.group + g) * 2);
    const float w = exp2f(ml.x - M);
    L += w * ml.y;
    o += w * p.ws_o[((wbase + s) * p.group + g) * 128 + d];
  }
  p.out[(size_t)p.q_start[b] * p.out_row_stride + (size_t)head * 128 + d] = __float2bfloat16(L > 0.f ? o / L : 0.f);
}

}  // namespace

int decode_attention_launch(const DecodeAttnParams& p, cudaStream_t stream) {
  VL_CHECK(p.batch > 0 && p.batch <= 65535 && p.n_kv_heads > 0 && p.group >= 1 && p.group <= DA_MAX_GROUP, "decode attention: bad shape");
  VL_CHECK(p.chunk >= 64 && p.chunk <= DA_MAX_CHUNK && p.chunk % 64 == 0, "decode attention: chunk must be 64 .. 256, a multiple of 64");
  VL_CHECK(p.max_kv_len > 0 && p.ws_o && p.ws_ml, "decode attention: no workspace");
  const int splits = (p.max_kv_len + p.chunk - 1) / p.chunk;
  static unsigned long long attr_devs = 0;
  if (vl_first_on_device(attr_devs))
    VL_CUDA_CHECK(cudaFuncSetAttribute(decode_attention_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DA_MAX_CHUNK * 512));
  VL_CUDA_CHECK(launch_kernel(decode_attention_split_kernel, dim3(splits, p.n_kv_heads, p.batch), dim3(DA_THREADS),
                              (size_t)p.chunk * 512, stream, p));
  VL_CUDA_CHECK(launch_kernel(decode_attention_merge_kernel, dim3(p.n_kv_heads * p.group, p.batch), dim3(128), 0, stream, p, splits));
  return 0;
}

namespace {

// ------------------------------------------------------------------------------------------
// Few query rows per (sample, kv head) -- the denoise steps: H action tokens x `group` heads = 24 rows of the 64-row tile of
// the generic kernel, i.e. two of its four warps idle while the other two walk the key blocks one after the other (the
// legacy mma.sync path costs ~2 us per 64-key block and warp). Here rows <= 32 and 8 warps: warp w owns row tile (w & 1)
// and, in every round of 128 keys, the 32 keys of group (w >> 1), so a 94-key window is multiplied in one round by six warps
// at once; the partial softmax states of a row tile are merged through shared memory in a fixed order (bit-identical from
// call to call).
// ------------------------------------------------------------------------------------------
constexpr int AS_THREADS = 256;
constexpr int AS_ROUND = 128;  // keys per round: 4 key groups x 32

__global__ void __launch_bounds__(AS_THREADS) attention_small_kernel(AttnParams p) {
  constexpr int HD = 128, CH = HD / 8;
  pdl_launch_dependents();
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* Qs = smem;                          // [32][HD]
  uint8_t* Ks = Qs + 32 * HD * 2;              // [128][HD]
  uint8_t* Vs = Ks + AS_ROUND * HD * 2;        // [128][HD]
  long long* tr = (blockIdx.x + blockIdx.y + blockIdx.z == 0 && threadIdx.x == 0) ? p.trace : nullptr;
  trace_stamp(tr, 0);
  const int b = blockIdx.z, kvh = blockIdx.y, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int rt = warp & 1, kg = warp >> 1;
  const int q_len = p.q_len ? p.q_len[b] : p.uni_q_len;
  const int q_start = p.q_len ? p.q_start[b] : b * p.uni_q_len;
  const int kv_len = p.kv_len ? p.kv_len[b] : p.uni_kv_len;
  const int kv_start = p.kv_len ? p.kv_start[b] : b * p.uni_kv_len;
  const int group = p.group;
  const int rows_total = group * q_len;  // <= 32
  if (rows_total <= 0 || kv_len < p.min_kv_len) return;
  const int n_rounds = (kv_len + AS_ROUND - 1) / AS_ROUND;
  const __nv_bfloat16* kbase = p.k + (size_t)kvh * p.kv_head_stride + (size_t)kv_start * p.kv_row_stride;
  const __nv_bfloat16* vbase = p.v + (size_t)kvh * p.kv_head_stride + (size_t)kv_start * p.kv_row_stride;
  // keys [k0 + r0, k0 + r1) of a round into rows [r0, r1) of the buffers
  auto load_rows = [&](int k0, int r0, int r1) {
    for (int i = tid + r0 * CH; i < r1 * CH; i += AS_THREADS) {
      const int row = i / CH, c = i - row * CH;
      const bool ok = k0 + row < kv_len;
      const size_t off = (size_t)(ok ? k0 + row : 0) * p.kv_row_stride + c * 8;
      cp_async_16(Ks + swz<HD>(row, c), kbase + off, ok);
      cp_async_16(Vs + swz<HD>(row, c), vbase + off, ok);
    }
  };
  // the first 64 keys lie in the stable prefix: fetch them before waiting on the previous grid
  const bool pre = p.stable_prefix && (kv_len - q_len >= 64);
  if (pre) load_rows(0, 0, 64);
  pdl_wait();
  trace_stamp(tr, 1);
  for (int i = tid; i < 32 * CH; i += AS_THREADS) {
    const int row = i / CH, c = i - row * CH;
    const bool ok = row < rows_total;
    const int t = ok ? row / group : 0, hh = ok ? row - t * group : 0;
    cp_async_16(Qs + swz<HD>(row, c), p.q + (size_t)(q_start + t) * p.q_row_stride + (size_t)(kvh * group + hh) * HD + c * 8, ok);
  }
  load_rows(0, pre ? 64 : 0, AS_ROUND);
  cp_async_commit();

  const int rlo = rt * 16 + (lane >> 2);
  const bool warp_active = rt * 16 < rows_total;
  const float sl2 = p.scale * 1.4426950408889634f;
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[HD / 16][4];
  const uint32_t qs_u = smem_u32(Qs), ks_u = smem_u32(Ks), vs_u = smem_u32(Vs);
  for (int r2 = 0; r2 < n_rounds; ++r2) {
    if (r2 > 0) {
      __syncthreads();  // the buffers have been consumed
      load_rows(r2 * AS_ROUND, 0, AS_ROUND);
      cp_async_commit();
    }
    cp_async_wait<0>();
    __syncthreads();
    const int key0 = r2 * AS_ROUND + kg * 32;  // this warp's 32 keys: buffer rows kg * 32 ..
    if (warp_active && key0 < kv_len) {
      if (r2 == 0) {
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const int mi = lane >> 3;
          const int row = rt * 16 + (mi & 1) * 8 + (lane & 7);
          ldmatrix_x4(qf[kk], qs_u + swz<HD>(row, kk * 2 + (mi >> 1)));
        }
      }
      float sc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {
          uint32_t kf[4];
          const int mi = lane >> 3;
          const int row = kg * 32 + jp * 16 + (mi >> 1) * 8 + (lane & 7);
          ldmatrix_x4(kf, ks_u + swz<HD>(row, kk * 2 + (mi & 1)));
          mma_bf16_16816(sc[2 * jp], qf[kk], kf[0], kf[1]);
          mma_bf16_16816(sc[2 * jp + 1], qf[kk], kf[2], kf[3]);
        }
      }
      const int colbase = key0 + (lane & 3) * 2;
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = colbase + j * 8 + (e & 1);
          const float val = (col < kv_len) ? sc[j][e] * sl2 : -INFINITY;
          sc[j][e] = val;
          mx[e >> 1] = fmaxf(mx[e >> 1], val);
        }
      }
      float alpha[2], moff[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
        const float m_new = fmaxf(m_run[h], mx[h]);
        moff[h] = (m_new == -INFINITY) ? 0.f : m_new;
        alpha[h] = exp2f(m_run[h] - moff[h]);
        m_run[h] = m_new;
      }
      float rs[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = exp2f(sc[j][e] - moff[e >> 1]);
          sc[j][e] = pv;
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
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t pf[4];
        pf[0] = pack_bf16(sc[2 * ks][0], sc[2 * ks][1]);
        pf[1] = pack_bf16(sc[2 * ks][2], sc[2 * ks][3]);
        pf[2] = pack_bf16(sc[2 * ks + 1][0], sc[2 * ks + 1][1]);
        pf[3] = pack_bf16(sc[2 * ks + 1][2], sc[2 * ks + 1][3]);
#pragma unroll
        for (int dp = 0; dp < HD / 16; ++dp) {
          uint32_t vf[4];
          const int mi = lane >> 3;
          const int row = kg * 32 + ks * 16 + (mi & 1) * 8 + (lane & 7);
          ldmatrix_x4_trans(vf, vs_u + swz<HD>(row, dp * 2 + (mi >> 1)));
          mma_bf16_16816(o[2 * dp], pf, vf[0], vf[1]);
          mma_bf16_16816(o[2 * dp + 1], pf, vf[2], vf[3]);
        }
      }
    }
  }
  // ---- merge the key-group states of a row tile: warps of groups 1..3 hand (m, l, o) to group 0 through the K / V buffers
  __syncthreads();
  float* xch = reinterpret_cast<float*>(Ks);  // [2 row tiles][3 groups][32 lanes][4 + 64]
  if (kg > 0) {
    float* w = xch + (((size_t)rt * 3 + (kg - 1)) * 32 + lane) * 68;
    w[0] = m_run[0], w[1] = m_run[1], w[2] = l_run[0], w[3] = l_run[1];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) *reinterpret_cast<float4*>(w + 4 + 4 * i) = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
  }
  __syncthreads();
  if (kg > 0 || !warp_active) return;
#pragma unroll 1
  for (int g2 = 0; g2 < 3; ++g2) {  // fixed order
    const float* w = xch + (((size_t)rt * 3 + g2) * 32 + lane) * 68;
    float a0[2], a1[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float m1 = w[h], l1 = w[2 + h];
      const float M = fmaxf(m_run[h], m1);
      const float mo = (M == -INFINITY) ? 0.f : M;
      a0[h] = exp2f(m_run[h] - mo);
      a1[h] = exp2f(m1 - mo);
      l_run[h] = l_run[h] * a0[h] + l1 * a1[h];
      m_run[h] = M;
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      const float4 o1 = *reinterpret_cast<const float4*>(w + 4 + 4 * i);
      o[i][0] = o[i][0] * a0[0] + o1.x * a1[0];
      o[i][1] = o[i][1] * a0[0] + o1.y * a1[0];
      o[i][2] = o[i][2] * a0[1] + o1.z * a1[1];
      o[i][3] = o[i][3] * a0[1] + o1.w * a1[1];
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = rlo + h * 8;
    if (r >= rows_total) continue;
    const int t = r / group, hh = r - t * group;
    const float inv = l_run[h] > 0.f ? 1.f / l_run[h] : 0.f;
    __nv_bfloat16* dst = p.out + (size_t)(q_start + t) * p.out_row_stride + (size_t)(kvh * group + hh) * HD;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i)
      *reinterpret_cast<uint32_t*>(dst + i * 8 + (lane & 3) * 2) = pack_bf16(o[i][2 * h] * inv, o[i][2 * h + 1] * inv);
  }
  trace_stamp(tr, 6);
}

template <int HD>
int launch_hd(const AttnParams& p, cudaStream_t stream) {
  constexpr int SMEM = (ATT_BM + 4 * ATT_BN) * HD * 2;
  static unsigned long long attr_devs = 0;  // devices this kernel's attributes were set on
  if (vl_first_on_device(attr_devs)) {
    VL_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  }
  const int rows_max = p.group * p.max_q_len;
  dim3 grid((rows_max + ATT_BM - 1) / ATT_BM, p.n_kv_heads, p.batch);
  if constexpr (HD == 128) {
    if (p.small_rows && rows_max <= 32 && !p.causal && p.q2 == nullptr) {  // the denoise steps: both warp pairs busy
      constexpr int SSMEM = (32 + 2 * AS_ROUND) * HD * 2;
      static unsigned long long attr_devs_s = 0;
      if (vl_first_on_device(attr_devs_s))
        VL_CUDA_CHECK(cudaFuncSetAttribute(attention_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SSMEM));
      VL_CUDA_CHECK(launch_kernel(attention_small_kernel, dim3(1, p.n_kv_heads, p.batch), dim3(AS_THREADS), SSMEM, stream, p));
      return 0;
    }
  }
  if (p.causal) {  // text prefill / decode of the chat VLM
    VL_CHECK(p.q2 == nullptr, "attention: the causal rule has no extra query row");
    static unsigned long long attr_devs_c = 0;
    if (vl_first_on_device(attr_devs_c))
      VL_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    VL_CUDA_CHECK(launch_kernel(attention_kernel<HD, true>, dim3(grid), dim3(ATT_THREADS), SMEM, stream, p));
    return 0;
  }
  VL_CUDA_CHECK(launch_kernel(attention_kernel<HD>, dim3(grid), dim3(ATT_THREADS), SMEM, stream, p));
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace

int vit_attention_launch(const __nv_bfloat16* qkv, __nv_bfloat16* out, int images, int tokens, int heads, float scale,
                         cudaStream_t stream) {
  VL_CHECK(tokens >= 128 && tokens <= VA_MAX_N, "vit_attention: 128 <= tokens <= 320");
  VL_CHECK(images > 0 && images <= 65535 && heads > 0, "vit_attention: bad grid");
  const int NR = (tokens + 15) & ~15;
  const int smem = 3 * NR * 128 + NR * 4 + (4 * 64 + 16) * 4;
  static unsigned long long attr_devs = 0;  // devices this kernel's attributes were set on
  if (vl_first_on_device(attr_devs)) {
    const int max_nr = (VA_MAX_N + 15) & ~15;
    VL_CUDA_CHECK(cudaFuncSetAttribute(vit_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       3 * max_nr * 128 + max_nr * 4 + (4 * 64 + 16) * 4));
  }
  const int qsplit = (heads * images * 2 <= gemm_num_sms() && tokens > 128) ? 2 : 1;
  VL_CUDA_CHECK(launch_kernel(vit_attention_kernel, dim3(heads, images, qsplit), dim3(VA_THREADS), smem, stream, qkv, out, tokens,
                              heads, scale));
  return 0;
}

int attention_launch(const AttnParams& p, int head_dim, cudaStream_t stream) {
  VL_CHECK(p.batch > 0 && p.max_q_len > 0, "attention: empty problem");
  VL_CHECK(p.batch <= 65535 && p.n_kv_heads <= 65535, "attention: grid too large");
  if (head_dim == 64) return launch_hd<64>(p, stream);
  if (head_dim == 128) return launch_hd<128>(p, stream);
  set_last_error("attention: head_dim must be 64 or 128");
  return -2;
}

}  // namespace vlaser
