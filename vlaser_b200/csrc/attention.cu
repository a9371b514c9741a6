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

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(AttnParams p) {
  pdl_launch_dependents();
  extern __shared__ __align__(128) uint8_t smem[];
  long long* tr = (blockIdx.x + blockIdx.y + blockIdx.z == 0 && threadIdx.x == 0) ? p.trace : nullptr;
  trace_stamp(tr, 0);
  attention_tile<HD>(p, blockIdx.z, blockIdx.y, blockIdx.x, smem, threadIdx.x, [] { __syncthreads(); },
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

template <int HD>
int launch_hd(const AttnParams& p, cudaStream_t stream) {
  constexpr int SMEM = (ATT_BM + 4 * ATT_BN) * HD * 2;
  static unsigned long long attr_devs = 0;  // devices this kernel's attributes were set on
  if (vl_first_on_device(attr_devs)) {
    VL_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  }
  const int rows_max = p.group * p.max_q_len;
  dim3 grid((rows_max + ATT_BM - 1) / ATT_BM, p.n_kv_heads, p.batch);
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
