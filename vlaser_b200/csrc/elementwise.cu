// Memory-bound kernels (see elementwise.cuh). All of them are one pass over their operands with
// 16-byte vector accesses; norms keep the row in registers (one warp per row).
#include "elementwise.cuh"

namespace vlaser {

static int g_resid_norm_small = 1;  // measurement switch (option "rn_small"): narrow split-K streams on resid_rmsnorm_small_kernel
void set_resid_norm_small(int on) { g_resid_norm_small = on; }

namespace {

// ------------------------------------------------------------------------------------------
// norms
// ------------------------------------------------------------------------------------------
template <int VPL, bool RMS, bool GATHER>
__global__ void __launch_bounds__(128) norm_kernel(float* __restrict__ x, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                                                   int rows, const int* __restrict__ rows_dev, float eps, int g,
                                                   int C, const L2Prefetch pf, long long* trace,
                                                   const RowBlocksReady ready = RowBlocksReady()) {
  pdl_launch_dependents();
  long long* tr = (blockIdx.x == 0 && threadIdx.x == 0) ? trace : nullptr;
  trace_stamp(tr, 0);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (pf.bytes[i] == 0) continue;
      const unsigned long long slice = ((pf.bytes[i] + gridDim.x - 1) / gridDim.x + 127ull) & ~127ull;
      const unsigned long long lo = blockIdx.x * slice, hi = min(pf.bytes[i], lo + slice);
      for (unsigned long long off = lo; off < hi; off += 32768ull) {
        const unsigned int sz = (unsigned int)min(32768ull, hi - off);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(pf.ptr[i]) + off), "r"(sz) : "memory");
      }
    }
  }
  if (ready.ctr == nullptr) pdl_wait();  // (with `ready` the rows are released block by block, see RowBlocksReady)
  trace_stamp(tr, 1);
  constexpr int W = VPL * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrows = rows_dev ? min(*rows_dev, rows) : rows;
  // (a device-resident row count comes with a capped grid: the warps stride over the live rows instead of launching one
  // warp per row of the capacity, most of which would only read the count and exit)
  for (int row = blockIdx.x * 4 + warp; row < nrows; row += gridDim.x * 4) {
  if (ready.ctr) {
    if (lane == 0) {
      const int* c = ready.ctr + (row >> 8);
      int v;
      long long t0 = 0;
      while (true) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
        if (v >= ready.expect) break;
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 2000000000LL) {  // the producer never reported this block: flag the call, go on
          if (ready.err) atomicExch(ready.err, NORM_ERR_READY_TIMEOUT);
          break;
        }
        __nanosleep(100);
      }
    }
    __syncwarp();
  }
  float4 v[VPL];
  if constexpr (GATHER) {
    // row = (img, r2, c2) of the pixel-shuffled grid; source = 4 ViT tokens (CLS skipped)
    const int hg = g >> 1, T = hg * hg, P = g * g;
    const int img = row / T, t = row - img * T;
    const int r2 = t / hg, c2 = t - r2 * hg;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int e = (i * 32 + lane) * 4;
      const int seg = e / C, within = e - seg * C;
      const int tok = (2 * r2 + (seg >> 1)) * g + 2 * c2 + (seg & 1);
      v[i] = *reinterpret_cast<const float4*>(x + (size_t)(img * (P + 1) + 1 + tok) * C + within);
    }
  } else {
    const float4* src = reinterpret_cast<const float4*>(x + (size_t)row * W);
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = src[i * 32 + lane];
  }
  float mean = 0.f;
  if constexpr (!RMS) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    mean = warp_sum(s) * (1.0f / W);
  }
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / W) + eps);
  uint2* dst = reinterpret_cast<uint2*>(out + (size_t)row * W);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int e4 = i * 32 + lane;
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + e4);
    float4 y;
    y.x = (v[i].x - mean) * rstd * gm.x;
    y.y = (v[i].y - mean) * rstd * gm.y;
    y.z = (v[i].z - mean) * rstd * gm.z;
    y.w = (v[i].w - mean) * rstd * gm.w;
    if constexpr (!RMS) {
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + e4);
      y.x += bt.x;
      y.y += bt.y;
      y.z += bt.z;
      y.w += bt.w;
    }
    uint2 u;
    u.x = pack_bf16(y.x, y.y);
    u.y = pack_bf16(y.z, y.w);
    dst[e4] = u;
  }
  }
  trace_stamp(tr, 6);
}

// ---- residual update + norm for a stream with pending split-K partials. One CTA per row; thread (c, g) owns float4
// column c and the splits g, g + G, g + 2G, ...: all of its loads are in flight at once (one L2 round trip for up to
// 8 G splits; the warp-per-row kernel above would walk them in dependent batches), the G group sums meet in shared memory
// and are added in a fixed order, so the result does not depend on scheduling. out == nullptr: residual update only.
template <bool RMS>
__global__ void __launch_bounds__(1024) resid_norm_kernel(float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                                                          int rows, const int* __restrict__ rows_dev, int W, float eps,
                                                          const SplitKPartials sp, long long* trace) {
  pdl_launch_dependents();
  extern __shared__ float4 rn_smem[];  // [G - 1][C4] group sums, then 32 floats of reduction scratch
  long long* tr = (blockIdx.x == 0 && threadIdx.x == 0) ? trace : nullptr;
  trace_stamp(tr, 0);
  const int C4 = W >> 2, G = blockDim.x / C4;
  const int c = threadIdx.x % C4, g = threadIdx.x / C4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* red = reinterpret_cast<float*>(rn_smem + (size_t)(G - 1) * C4);
  float4 gm = make_float4(0.f, 0.f, 0.f, 0.f), bt = gm;
  if (g == 0 && out) {  // constants: ahead of the dependent-launch wait
    gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
    if (!RMS) bt = __ldg(reinterpret_cast<const float4*>(beta) + c);
  }
  pdl_wait();
  trace_stamp(tr, 1);
  const int row = blockIdx.x;
  const int nrows = rows_dev ? min(*rows_dev, rows) : rows;
  if (row >= nrows) return;
  const int S = sp.splits ? *sp.splits : 0;
  const size_t off = (size_t)row * W + 4 * c;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g == 0) v = *reinterpret_cast<const float4*>(x + off);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s0 = g; s0 < S; s0 += 8 * G) {
    float4 pv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int s = s0 + u * G;
      pv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s < S) pv[u] = __ldcg(reinterpret_cast<const float4*>(sp.part + (size_t)s * sp.stride + off));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc.x += pv[u].x;
      acc.y += pv[u].y;
      acc.z += pv[u].z;
      acc.w += pv[u].w;
    }
  }
  if (g > 0) rn_smem[(size_t)(g - 1) * C4 + c] = acc;
  __syncthreads();
  if (g == 0) {
    v.x += acc.x, v.y += acc.y, v.z += acc.z, v.w += acc.w;
    for (int gg = 1; gg < G; ++gg) {
      const float4 o = rn_smem[(size_t)(gg - 1) * C4 + c];
      v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
    }
    if (S > 0) *reinterpret_cast<float4*>(x + off) = v;
  }
  if (out == nullptr) return;
  auto block_sum = [&](float val) {
    val = warp_sum(val);
    __syncthreads();  // scratch free again
    if (lane == 0) red[warp] = val;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += red[w];
    return t;
  };
  float mean = 0.f;
  if (!RMS) mean = block_sum(g == 0 ? (v.x + v.y) + (v.z + v.w) : 0.f) / (float)W;
  const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, dd = v.w - mean;
  const float q = block_sum(g == 0 ? (a * a + b * b) + (cc * cc + dd * dd) : 0.f);
  if (g == 0) {
    const float rstd = rsqrtf(q / (float)W + eps);
    float4 y = make_float4(a * rstd * gm.x, b * rstd * gm.y, cc * rstd * gm.z, dd * rstd * gm.w);
    if (!RMS) y.x += bt.x, y.y += bt.y, y.z += bt.z, y.w += bt.w;
    uint2 u;
    u.x = pack_bf16(y.x, y.y);
    u.y = pack_bf16(y.z, y.w);
    *reinterpret_cast<uint2*>(out + off) = u;
  }
  trace_stamp(tr, 6);
}

// The same residual update + RMSNorm for narrow streams with few splits (the action expert: 768 columns, <= 12 partial
// tiles): one CTA per row with ONE thread per float4 column -- all 13 loads of a thread in flight at once, no exchange of
// group sums through shared memory, a block reduction over 6 warps instead of 24. The partial sums keep the association of
// resid_norm_kernel (four interleaved split groups, each summed in ascending order, then added to x in group order), so the
// residual stream is bit-identical; only the reduction tree of the sum of squares differs.
__global__ void __launch_bounds__(256) resid_rmsnorm_small_kernel(float* __restrict__ x, const float* __restrict__ gamma,
                                                                  __nv_bfloat16* __restrict__ out, int rows,
                                                                  const int* __restrict__ rows_dev, int W, float eps,
                                                                  const SplitKPartials sp, long long* trace) {
  __shared__ float red[8];
  pdl_launch_dependents();
  long long* tr = (blockIdx.x == 0 && threadIdx.x == 0) ? trace : nullptr;
  trace_stamp(tr, 0);
  const int C4 = W >> 2, c = threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float4 gm = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C4 && out) gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
  pdl_wait();
  trace_stamp(tr, 1);
  const int row = blockIdx.x;
  const int nrows = rows_dev ? min(*rows_dev, rows) : rows;
  if (row >= nrows) return;
  const int S = sp.known_splits >= 0 ? sp.known_splits : (sp.splits ? *sp.splits : 0);  // <= 12 on this path
  const size_t off = (size_t)row * W + 4 * c;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C4) {
    v = *reinterpret_cast<const float4*>(x + off);
    float4 pv[12];
#pragma unroll
    for (int u = 0; u < 12; ++u) {
      pv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u < S) pv[u] = __ldcg(reinterpret_cast<const float4*>(sp.part + (size_t)u * sp.stride + off));
    }
    float4 acc[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 12; ++u) {
      acc[u & 3].x += pv[u].x;
      acc[u & 3].y += pv[u].y;
      acc[u & 3].z += pv[u].z;
      acc[u & 3].w += pv[u].w;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) v.x += acc[g].x, v.y += acc[g].y, v.z += acc[g].z, v.w += acc[g].w;
    if (S > 0) *reinterpret_cast<float4*>(x + off) = v;
  }
  if (out == nullptr) return;
  float q = warp_sum((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
  if (lane == 0) red[warp] = q;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < nwarps; ++w) tot += red[w];
  if (c < C4) {
    const float rstd = rsqrtf(tot / (float)W + eps);
    uint2 u;
    u.x = pack_bf16(v.x * rstd * gm.x, v.y * rstd * gm.y);
    u.y = pack_bf16(v.z * rstd * gm.z, v.w * rstd * gm.w);
    *reinterpret_cast<uint2*>(out + off) = u;
  }
  trace_stamp(tr, 6);
}

template <bool RMS>
int resid_norm_launch(float* x, const float* gamma, const float* beta, __nv_bfloat16* out, int rows, const int* rows_dev,
                      int width, float eps, const SplitKPartials& sp, cudaStream_t stream, long long* trace) {
  VL_CHECK(width % 128 == 0 && width <= 4096, "resid_norm: width must be a multiple of 128, <= 4096");
  if (rows <= 0) return 0;
  if constexpr (RMS) {
    // narrow stream, few splits (sp.max_splits is the launch's cap): one thread per column, see the kernel
    if (g_resid_norm_small && width <= 1024 && sp.max_splits > 0 && sp.max_splits <= 12) {
      const int threads = (width / 4 + 31) / 32 * 32;
      VL_CUDA_CHECK(launch_kernel(resid_rmsnorm_small_kernel, dim3(rows), dim3(threads), 0, stream, x, gamma, out, rows, rows_dev,
                                  width, eps, sp, trace));
      return 0;
    }
  }
  const int C4 = width / 4;
  const int G = std::max(1, std::min(4, 1024 / C4));
  const size_t smem = (size_t)(G - 1) * C4 * sizeof(float4) + 32 * sizeof(float);
  VL_CUDA_CHECK(launch_kernel(resid_norm_kernel<RMS>, dim3(rows), dim3(C4 * G), smem, stream, x, gamma, beta, out, rows,
                              rows_dev, width, eps, sp, trace));
  return 0;
}

template <bool RMS, bool GATHER>
int norm_dispatch(float* x, const float* gamma, const float* beta, __nv_bfloat16* out, int rows,
                  const int* rows_dev, int width, float eps, int g, int C, cudaStream_t stream,
                  const L2Prefetch& pf = L2Prefetch(), long long* trace = nullptr,
                  const SplitKPartials& sp = SplitKPartials(), const RowBlocksReady& ready = RowBlocksReady()) {
  VL_CHECK(!ready.ctr || (!GATHER && !sp.part && !rows_dev), "norm: row-block release needs the plain kernel at a static row count");
  VL_CHECK(!(GATHER && sp.part), "norm: the pixel-shuffle gather variant does not take split-K partials");
  if constexpr (!GATHER) {
    if (sp.part) return resid_norm_launch<RMS>(x, gamma, beta, out, rows, rows_dev, width, eps, sp, stream, trace);
  }
  VL_CHECK(width % 128 == 0 && width >= 128 && width <= 4096, "norm: width must be a multiple of 128, <= 4096");
  if (rows <= 0) return 0;
  // live rows in device memory (the packed prefix): `rows` is the capacity, a few times the live count
  const int grid = rows_dev ? std::min((rows + 3) / 4, 8 * 148) : (rows + 3) / 4;
#define VL_NORM_CASE(V)                                                                                          \
  case V:                                                                                                        \
    VL_CUDA_CHECK(launch_kernel(norm_kernel<V, RMS, GATHER>, dim3(grid), dim3(128), 0, stream, x, gamma, beta, out, rows, rows_dev, eps, g, C, pf, trace, ready)); \
    break;
  switch (width / 128) {
    VL_NORM_CASE(1) VL_NORM_CASE(2) VL_NORM_CASE(3) VL_NORM_CASE(4) VL_NORM_CASE(6) VL_NORM_CASE(8)
    VL_NORM_CASE(12) VL_NORM_CASE(16) VL_NORM_CASE(28) VL_NORM_CASE(32)
    default:
      set_last_error("norm: unsupported width " + std::to_string(width));
      return -2;
  }
#undef VL_NORM_CASE
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// pixel_shuffle (standalone, bit exact)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void pixel_shuffle_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int g, int C) {
  pdl_launch_dependents();
  pdl_wait();
  // out[img, r2, c2, b*2C + a*C + c] = in[img, 2*r2+b, 2*c2+a, c]
  const int hg = g >> 1;
  const long long total = (long long)n * hg * hg * 4 * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int seg = (int)(r & 3);
    r >>= 2;
    const int c2 = (int)(r % hg);
    r /= hg;
    const int r2 = (int)(r % hg);
    const int img = (int)(r / hg);
    const long long src = (((long long)img * g + (2 * r2 + (seg >> 1))) * g + (2 * c2 + (seg & 1))) * C + c;
    out[i] = in[src];
  }
}

// ------------------------------------------------------------------------------------------
// im2col / cls rows
// ------------------------------------------------------------------------------------------
// One thread per (patch row, unit): units 0 .. 3 ps - 1 are the runs (channel, ky) of ps contiguous pixels -- contiguous in
// the image and in the output row (k = ch ps^2 + ky ps + kx) --, the last unit zeroes the K padding. (One thread per output
// ELEMENT with 64-bit div / mod took 73 us for 40 MB of traffic at 64 images, profiles/r02s2zz launch list.)
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ pix, __nv_bfloat16* __restrict__ cols, int n, int px, int ps,
                              int kpad) {
  pdl_launch_dependents();
  pdl_wait();
  const int g = px / ps;
  const int kreal = 3 * ps * ps;
  const int units = 3 * ps + 1;
  const int total = n * g * g * units;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int row = i / units, u = i - row * units;
    __nv_bfloat16* dst = cols + (size_t)row * kpad;
    if (u == 3 * ps) {
      for (int k = kreal; k < kpad; ++k) dst[k] = __float2bfloat16(0.f);
      continue;
    }
    const int img = row / (g * g), p = row - img * (g * g);
    const int pr = p / g, pc = p - pr * g;
    const int ch = u / ps, ky = u - ch * ps;
    const T* src = pix + (((size_t)img * 3 + ch) * px + pr * ps + ky) * px + pc * ps;
    dst += ch * ps * ps + ky * ps;
    if (sizeof(T) == 2 && (ps & 1) == 0 && (px & 1) == 0 && (kpad & 1) == 0) {
      // bf16 in, bf16 out: a bit copy, two pixels per word (every offset above is even)
      const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
      uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
      for (int k = 0; k < ps / 2; ++k) d32[k] = s32[k];
    } else {
      for (int k = 0; k < ps; ++k) dst[k] = __float2bfloat16(static_cast<float>(src[k]));
    }
  }
}

__global__ void cls_rows_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                                int n, int tokens, int width) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  const int img = i / width, c = i - img * width;
  x[(size_t)img * tokens * width + c] = cls[c] + pos[c];
}

// ------------------------------------------------------------------------------------------
// planning
// ------------------------------------------------------------------------------------------
__global__ void plan_lengths_kernel(const long long* __restrict__ ids, const int* __restrict__ valid_in, PlanBuffers pb,
                                    int B, int S, int H, int pad_token, int image_token) {
  pdl_launch_dependents();
  pdl_wait();
  // one warp per sample counts non-pad ids
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  int cnt = 0;
  if (valid_in) {
    cnt = valid_in[warp];
  } else {
    for (int j = lane; j < S; j += 32) cnt += ids[(size_t)warp * S + j] != pad_token;
    cnt = (int)warp_sum((float)cnt);
  }
  cnt = min(max(cnt, 0), S);
  int imgs = 0;
  for (int j = lane; j < cnt; j += 32) imgs += ids[(size_t)warp * S + j] == image_token;
  imgs = (int)warp_sum((float)imgs);
  if (lane == 0) {
    pb.valid_len[warp] = cnt;
    pb.img_cnt[warp] = imgs;
  }
}

__global__ void plan_scan_kernel(PlanBuffers pb, int B, int H, int kv_stride) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0, img = 0;
    for (int b = 0; b < B; ++b) {
      pb.cu[b] = acc;
      // policy call: K/V rows packed like the prefix (+ proprio + H action rows per sample); text generation: every sample
      // owns a fixed window of kv_stride rows that its decode steps fill
      pb.kvoff[b] = kv_stride > 0 ? b * kv_stride : acc + b * (1 + H);
      pb.img_off[b] = img;
      acc += pb.valid_len[b];
      img += pb.img_cnt[b];
    }
    pb.cu[B] = acc;
    *pb.m_tot = acc;
  }
}

__global__ void plan_rows_kernel(const long long* __restrict__ ids, PlanBuffers pb, int B, int S, int H,
                                 int img_tokens, int image_token, int pad_token, int embed_rows, int pos0) {
  pdl_launch_dependents();
  pdl_wait();
  // one warp per sample
  const int b = blockIdx.x, lane = threadIdx.x;
  const int vl = pb.valid_len[b], base = pb.cu[b], kvo = pb.kvoff[b];
  int seen = pb.img_off[b];  // image rows are numbered across the batch in row-major order
  for (int j0 = 0; j0 < S; j0 += 32) {
    const int j = j0 + lane;
    const bool valid = j < vl;
    const long long id = (j < S) ? ids[(size_t)b * S + j] : pad_token;
    const bool is_img = valid && id == image_token;
    const unsigned mask = __ballot_sync(0xffffffffu, is_img);
    if (valid) {
      const int m = base + j;
      pb.row_pos[m] = j + pos0;  // policy call: ids start at 1 (pizero_internvl.py:575-577); chat VLM: at 0 (HF)
      pb.row_kvrow[m] = kvo + j;
      int src;
      if (is_img) {
        const int k = seen + __popc(mask & ((1u << lane) - 1u));
        if (k < B * img_tokens) {
          pb.img_row_map[k] = m;
          src = -2;
        } else {
          src = -1;
        }
      } else {
        src = (id == pad_token) ? -1 : (int)id;
        if (id != pad_token && (id < 0 || id >= embed_rows)) {  // out of the embedding table: zeros + error flag
          src = -1;
          atomicCAS(pb.bad_id, 0, 1 + (int)min(max(id, (long long)0), (long long)0x7ffffff0));
        }
      }
      pb.row_src[m] = src;
    }
    seen += __popc(mask);
  }
  if (lane == 0) {
    pb.prop_pos[b] = 1;
    pb.prop_kvrow[b] = kvo + vl;
    pb.pre_kv_len[b] = vl;
    pb.den_q_start[b] = b * H;
    pb.den_q_len[b] = H;
    pb.den_kv_len[b] = vl + 1 + H;
    pb.prop_q_start[b] = b;
    pb.prop_q_len[b] = 1;
    pb.prop_kv_len[b] = vl + 1;
  }
  for (int i = lane; i < H; i += 32) {
    pb.act_pos[b * H + i] = 2 + i;
    pb.act_kvrow[b * H + i] = kvo + vl + 1 + i;
  }
}

__global__ void assemble_text_kernel(const __nv_bfloat16* __restrict__ embed, const int* __restrict__ row_src,
                                     const int* __restrict__ m_tot, float* __restrict__ x, int rows_cap, int width) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= min(*m_tot, rows_cap)) return;
  const int src = row_src[row];
  if (src == -2) return;
  float* dst = x + (size_t)row * width;
  if (src < 0) {
    for (int c = lane * 4; c < width; c += 128) *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const __nv_bfloat16* s = embed + (size_t)src * width;
  for (int c = lane * 8; c < width; c += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(s + c);
    *reinterpret_cast<float4*>(dst + c) = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
    *reinterpret_cast<float4*>(dst + c + 4) = make_float4(bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w));
  }
}

// ------------------------------------------------------------------------------------------
// small heads
// ------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void small_linear_kernel(const TI* __restrict__ in, const float* __restrict__ W,
                                    const float* __restrict__ bias, TO* __restrict__ out, int rows, int K, int N) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)rows * N) return;
  const int r = (int)(i / N), n = (int)(i - (long long)r * N);
  float acc = bias ? bias[n] : 0.f;
  for (int k = 0; k < K; ++k) acc += W[(size_t)n * K + k] * static_cast<float>(in[(size_t)r * K + k]);
  out[i] = static_cast<TO>(acc);
}

template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = static_cast<TO>(static_cast<float>(in[i]));
}

// ------------------------------------------------------------------------------------------
// final norm + decoder + Euler
// ------------------------------------------------------------------------------------------
// One warp per (row, decoder output): every warp normalises its row itself (the row is a few KB in L2) and then has all
// the loads of ONE decoder-weight row in flight at once -- two dependent load rounds per warp instead of a serial sweep
// over the A decoder rows, and rows x A warps instead of rows (B = 1: 28 warps instead of 4).
template <typename TO>
__global__ void __launch_bounds__(128) euler_step_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ Wdec,
                                                         const float* __restrict__ bdec, float* __restrict__ action,
                                                         TO* __restrict__ out, int rows, int width, int A, int H,
                                                         int horizon, float dt, float eps, int last_step, int do_clip,
                                                         float clip, const long long* __restrict__ xfix) {
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * 4 + warp;
  if (unit >= rows * A) return;
  const int r = unit / A, a = unit - r * A;
  // constants first: they do not depend on the previous launch (width <= 1024: up to 8 float4 per lane)
  const float4* wr = reinterpret_cast<const float4*>(Wdec + (size_t)a * width);
  const float4* gr = reinterpret_cast<const float4*>(gamma);
  float4 wv[8], gv[8], xv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c4 = lane + 32 * i;
    wv[i] = gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 * 4 < width) wv[i] = __ldg(wr + c4), gv[i] = __ldg(gr + c4);
  }
  const float bias = __ldg(bdec + a);
  pdl_wait();
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)r * width);
  float q = 0.f, dot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c4 = lane + 32 * i;
    xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 * 4 < width) {
      xv[i] = xr[c4];
      if (xfix) {
        const longlong2* f2 = reinterpret_cast<const longlong2*>(xfix + (size_t)r * width) + 2 * c4;
        const longlong2 f0 = f2[0], f1 = f2[1];
        xv[i].x += (float)f0.x * (1.f / 1073741824.f);
        xv[i].y += (float)f0.y * (1.f / 1073741824.f);
        xv[i].z += (float)f1.x * (1.f / 1073741824.f);
        xv[i].w += (float)f1.y * (1.f / 1073741824.f);
      }
    }
  }
  const float prev = action[(size_t)r * A + a];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
    dot += (xv[i].x * gv[i].x * wv[i].x + xv[i].y * gv[i].y * wv[i].y) + (xv[i].z * gv[i].z * wv[i].z + xv[i].w * gv[i].w * wv[i].w);
  }
  q = warp_sum(q);
  dot = warp_sum(dot);
  if (lane == 0) {
    const float rstd = rsqrtf(q / width + eps);
    float val = prev + dt * (dot * rstd + bias);
    action[(size_t)r * A + a] = val;
    if (last_step) {
      const int b = r / H, i = r - b * H;
      if (i >= H - horizon) {
        if (do_clip) val = fminf(fmaxf(val, -clip), clip);
        out[((size_t)b * horizon + (i - (H - horizon))) * A + a] = static_cast<TO>(val);
      }
    }
  }
}

// ---- flow-matching training forward (see elementwise.cuh)
__device__ __forceinline__ float io_load(const void* p, int io, long long i) {
  return io == DT_F32 ? reinterpret_cast<const float*>(p)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

__global__ void __launch_bounds__(256) flow_prepare_kernel(const void* __restrict__ x0, const void* __restrict__ x1,
                                                           const void* __restrict__ tin, int io, float sig_min, float emb,
                                                           const float* __restrict__ W2a, const float* __restrict__ b2,
                                                           float* __restrict__ action, float* __restrict__ tbias,
                                                           float* __restrict__ vel, int HA, int he) {
  extern __shared__ float tau[];  // [he]
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, half = he / 2;
  const float t = io_load(tin, io, b);
  for (int i = threadIdx.x; i < HA; i += blockDim.x) {
    const long long k = (long long)b * HA + i;
    action[k] = (1.f - (1.f - sig_min) * t) * io_load(x0, io, k) + t * io_load(x1, io, k);
    vel[k] = 0.f;
  }
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float a = t * expf((float)i * -emb);
    tau[i] = sinf(a);
    tau[half + i] = cosf(a);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int n = warp; n < he; n += nw) {
    const float* w = W2a + (size_t)n * he;
    float acc = 0.f;
    for (int j = lane; j < he; j += 32) acc = fmaf(__ldg(w + j), tau[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) tbias[(size_t)b * he + n] = acc + __ldg(b2 + n);
  }
}

__global__ void __launch_bounds__(256) bias_silu_rows_kernel(const float* __restrict__ z, const float* __restrict__ tbias,
                                                             __nv_bfloat16* __restrict__ h, int rows, int H, int width) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n = (long long)rows * width;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / width), c = (int)(i - (long long)r * width);
    const float v = z[i] + tbias[(size_t)(r / H) * width + c];
    h[i] = __float2bfloat16(silu(v));
  }
}

__global__ void __launch_bounds__(256) flow_loss_kernel(const float* __restrict__ vel, const void* __restrict__ x0,
                                                        const void* __restrict__ x1, int io, float sig_min, long long n,
                                                        float* __restrict__ loss) {
  __shared__ float part[256];
  pdl_launch_dependents();
  pdl_wait();
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += 256) {  // fixed thread -> element map and tree: bit-identical from call to call
    const float d = vel[i] - (io_load(x1, io, i) - (1.f - sig_min) * io_load(x0, io, i));
    acc = fmaf(d, d, acc);
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = part[0] / (float)n;
}

// ---- greedy text generation of the chat VLM (see elementwise.cuh)
__global__ void __launch_bounds__(256) last_rows_kernel(const float* __restrict__ x, const int* __restrict__ cu,
                                                        const int* __restrict__ valid_len, float* __restrict__ out, int width) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int vl = valid_len[b];
  const float* src = x + (size_t)(cu[b] + max(vl, 1) - 1) * width;
  for (int c = threadIdx.x * 4; c < width; c += 1024)
    *reinterpret_cast<float4*>(out + (size_t)b * width + c) =
        vl > 0 ? *reinterpret_cast<const float4*>(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(1024) argmax_rows_kernel(const float* __restrict__ logits, long long ld, int n,
                                                           int* __restrict__ next_tok, int* __restrict__ n_gen, int bump) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  pdl_launch_dependents();
  pdl_wait();
  const float* row = logits + (size_t)blockIdx.x * ld;
  float best = -INFINITY;
  int idx = 0x7fffffff;
  // ld is a multiple of 128 floats: float4 loads, four in flight per thread; ties keep the lowest index, like torch.argmax
  const int n4 = n >> 2;
  const float4* row4 = reinterpret_cast<const float4*>(row);
  for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      v[u] = i < n4 ? __ldcs(row4 + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = 4 * (i0 + u * (int)blockDim.x);
      const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (e[k] > best || (e[k] == best && i + k < idx)) best = e[k], idx = i + k;
    }
  }
  for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) {
    const float v = row[i];
    if (v > best || (v == best && i < idx)) best = v, idx = i;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > best || (ov == best && oi < idx)) best = ov, idx = oi;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) bv[warp] = best, bi[warp] = idx;
  __syncthreads();
  if (warp == 0) {
    best = lane < (int)(blockDim.x >> 5) ? bv[lane] : -INFINITY;
    idx = lane < (int)(blockDim.x >> 5) ? bi[lane] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > best || (ov == best && oi < idx)) best = ov, idx = oi;
    }
    if (lane == 0) {
      next_tok[blockIdx.x] = idx == 0x7fffffff ? 0 : idx;
      if (bump && blockIdx.x == 0) *n_gen += 1;  // (every reader of n_gen in this step ran in an earlier launch)
    }
  }
}

__global__ void __launch_bounds__(256) decode_advance_kernel(TextGenBuffers g, const int* __restrict__ valid_len,
                                                             const int* __restrict__ kvoff,
                                                             const __nv_bfloat16* __restrict__ embed, float* __restrict__ x,
                                                             int width, int window, int gen_cap, long long eos, long long fill) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int k = *g.n_gen;  // tokens emitted so far
  const int vl = valid_len[b];
  const bool done = g.finished[b] != 0;
  long long tok = done ? fill : (long long)g.next_tok[b];
  // the emitted token becomes the input of the next step at position vl + k; a full window ends the sample
  const bool full = vl + k + 1 > window;
  __syncthreads();  // every thread has read finished[b] before thread 0 updates it
  if (threadIdx.x == 0) {
    if (k < gen_cap) g.tokens[(size_t)b * gen_cap + k] = tok;
    if (!done && (tok == eos || full)) g.finished[b] = 1;
    const int pos = min(vl + k, window - 1);
    g.pos[b] = pos;
    g.kvrow[b] = kvoff[b] + pos;
    g.kv_len[b] = pos + 1;
    g.q_start[b] = b;
    g.q_len[b] = 1;
  }
  const __nv_bfloat16* src = embed + (size_t)(done ? 0 : tok) * width;
  for (int c = threadIdx.x * 8; c < width; c += 2048) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + c);
    float* dst = x + (size_t)b * width + c;
    *reinterpret_cast<float4*>(dst) = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
    *reinterpret_cast<float4*>(dst + 4) = make_float4(bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w));
  }
}

}  // namespace

// ==========================================================================================
int layernorm_launch(float* x, const float* gamma, const float* beta, __nv_bfloat16* out, int rows, int width,
                     float eps, ShuffleGather sg, cudaStream_t stream, const SplitKPartials* sp, const RowBlocksReady* ready) {
  if (sg.grid > 0) {
    VL_CHECK(width == 4 * sg.channels && sg.channels % 4 == 0 && sg.grid % 2 == 0, "layernorm gather: bad shape");
    VL_CHECK(!sp || !sp->part, "layernorm gather: reduce the split-K partials first (splitk_reduce_launch)");
    return norm_dispatch<false, true>(x, gamma, beta, out, rows, nullptr, width, eps, sg.grid, sg.channels, stream);
  }
  return norm_dispatch<false, false>(x, gamma, beta, out, rows, nullptr, width, eps, 0, 0, stream, L2Prefetch(), nullptr,
                                     sp ? *sp : SplitKPartials(), ready ? *ready : RowBlocksReady());
}

int rmsnorm_launch(float* x, const float* gamma, __nv_bfloat16* out, int rows, const int* rows_dev, int width,
                   float eps, cudaStream_t stream, const L2Prefetch* pf, long long* trace, const SplitKPartials* sp) {
  return norm_dispatch<true, false>(x, gamma, nullptr, out, rows, rows_dev, width, eps, 0, 0, stream,
                                    pf ? *pf : L2Prefetch(), trace, sp ? *sp : SplitKPartials());
}

int splitk_reduce_launch(float* x, int rows, const int* rows_dev, int width, const SplitKPartials& sp, cudaStream_t stream) {
  VL_CHECK(sp.part && sp.splits, "splitk_reduce: bad arguments");
  return resid_norm_launch<true>(x, nullptr, nullptr, nullptr, rows, rows_dev, width, 0.f, sp, stream, nullptr);
}

int pixel_shuffle_launch(const void* in, void* out, int n, int grid, int channels, int elem_bytes,
                         cudaStream_t stream) {
  VL_CHECK(grid % 2 == 0 && grid > 0, "pixel_shuffle: grid must be even");
  const long long total = (long long)n * grid * grid * channels;
  if (total == 0) return 0;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  if (elem_bytes == 4)
    VL_CUDA_CHECK(launch_kernel(pixel_shuffle_kernel<uint32_t>, dim3(blocks), dim3(256), 0, stream, (const uint32_t*)in, (uint32_t*)out, n, grid, channels));
  else if (elem_bytes == 2)
    VL_CUDA_CHECK(launch_kernel(pixel_shuffle_kernel<uint16_t>, dim3(blocks), dim3(256), 0, stream, (const uint16_t*)in, (uint16_t*)out, n, grid, channels));
  else {
    set_last_error("pixel_shuffle: elem_bytes must be 2 or 4");
    return -2;
  }
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int im2col_launch(const void* pixel, int pixel_dtype, __nv_bfloat16* cols, int n, int px, int ps, int kpad,
                  cudaStream_t stream) {
  VL_CHECK(px % ps == 0 && kpad >= 3 * ps * ps, "im2col: bad shape");
  const long long total = (long long)n * (px / ps) * (px / ps) * (3 * ps + 1);
  if (total == 0) return 0;
  VL_CHECK(total < (1ll << 31), "im2col: too many patch rows for one launch");
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 32);
  if (pixel_dtype == DT_F32)
    VL_CUDA_CHECK(launch_kernel(im2col_kernel<float>, dim3(blocks), dim3(256), 0, stream, (const float*)pixel, cols, n, px, ps, kpad));
  else
    VL_CUDA_CHECK(launch_kernel(im2col_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, stream, (const __nv_bfloat16*)pixel, cols, n, px, ps, kpad));
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int cls_rows_launch(float* x, const float* cls, const float* pos, int n, int tokens, int width, cudaStream_t stream) {
  if (n <= 0) return 0;
  VL_CUDA_CHECK(launch_kernel(cls_rows_kernel, dim3((n * width + 255) / 256), dim3(256), 0, stream, x, cls, pos, n, tokens, width));
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int plan_launch(const long long* input_ids, const int* valid_len_in, PlanBuffers pb, int B, int S, int H,
                int img_tokens_per_sample, int image_token, int pad_token, int embed_rows, cudaStream_t stream,
                int text_kv_stride) {
  VL_CHECK(B > 0, "plan: empty batch");
  VL_CUDA_CHECK(cudaMemsetAsync(pb.img_row_map, 0xFF, sizeof(int) * (size_t)B * img_tokens_per_sample, stream));
  VL_CUDA_CHECK(launch_kernel(plan_lengths_kernel, dim3((B * 32 + 127) / 128), dim3(128), 0, stream, input_ids, valid_len_in, pb, B, S, H, pad_token, image_token));
  VL_CUDA_CHECK(launch_kernel(plan_scan_kernel, dim3(1), dim3(32), 0, stream, pb, B, H, text_kv_stride));
  VL_CUDA_CHECK(launch_kernel(plan_rows_kernel, dim3(B), dim3(32), 0, stream, input_ids, pb, B, S, H, img_tokens_per_sample, image_token, pad_token, embed_rows,
                              text_kv_stride > 0 ? 0 : 1));
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int assemble_text_launch(const __nv_bfloat16* embed, const int* row_src, const int* m_tot, float* x, int rows_cap,
                         int width, cudaStream_t stream) {
  VL_CHECK(width % 8 == 0, "assemble: width % 8");
  if (rows_cap <= 0) return 0;
  VL_CUDA_CHECK(launch_kernel(assemble_text_kernel, dim3((rows_cap + 3) / 4), dim3(128), 0, stream, embed, row_src, m_tot, x, rows_cap, width));
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int small_linear_launch(const void* in, int in_dtype, const float* W, const float* bias, void* out, int out_dtype,
                        int rows, int K, int N, cudaStream_t stream) {
  const long long total = (long long)rows * N;
  if (total <= 0) return 0;
  const int blocks = (int)((total + 255) / 256);
  if (in_dtype == DT_F32 && out_dtype == DT_F32)
    VL_CUDA_CHECK(launch_kernel(small_linear_kernel<float, float>, dim3(blocks), dim3(256), 0, stream, (const float*)in, W, bias, (float*)out, rows, K, N));
  else if (in_dtype == DT_F32 && out_dtype == DT_BF16)
    VL_CUDA_CHECK(launch_kernel(small_linear_kernel<float, __nv_bfloat16>, dim3(blocks), dim3(256), 0, stream, (const float*)in, W, bias, (__nv_bfloat16*)out, rows, K, N));
  else if (in_dtype == DT_BF16 && out_dtype == DT_F32)
    VL_CUDA_CHECK(launch_kernel(small_linear_kernel<__nv_bfloat16, float>, dim3(blocks), dim3(256), 0, stream, (const __nv_bfloat16*)in, W, bias, (float*)out, rows, K, N));
  else
    VL_CUDA_CHECK(launch_kernel(small_linear_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(blocks), dim3(256), 0, stream, (const __nv_bfloat16*)in, W, bias, (__nv_bfloat16*)out, rows, K, N));
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int convert_launch(const void* in, int in_dtype, void* out, int out_dtype, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  if (in_dtype == DT_F32 && out_dtype == DT_F32)
    VL_CUDA_CHECK(launch_kernel(convert_kernel<float, float>, dim3(blocks), dim3(256), 0, stream, (const float*)in, (float*)out, n));
  else if (in_dtype == DT_BF16 && out_dtype == DT_F32)
    VL_CUDA_CHECK(launch_kernel(convert_kernel<__nv_bfloat16, float>, dim3(blocks), dim3(256), 0, stream, (const __nv_bfloat16*)in, (float*)out, n));
  else if (in_dtype == DT_F32 && out_dtype == DT_BF16)
    VL_CUDA_CHECK(launch_kernel(convert_kernel<float, __nv_bfloat16>, dim3(blocks), dim3(256), 0, stream, (const float*)in, (__nv_bfloat16*)out, n));
  else
    VL_CUDA_CHECK(launch_kernel(convert_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(blocks), dim3(256), 0, stream, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, n));
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int euler_step_launch(const float* x, const float* gamma, const float* Wdec, const float* bdec, float* action,
                      void* out, int out_dtype, int rows, int width, int A, int H, int horizon, float dt, float eps,
                      int last_step, int do_clip, float clip, cudaStream_t stream, const long long* xfix) {
  if (rows <= 0) return 0;
  VL_CHECK(width <= 1024, "euler_step: width must be <= 1024");
  VL_CHECK(width % 4 == 0 && width <= 1024, "euler step: width must be a multiple of 4, <= 1024");
  const int grid = (rows * A + 3) / 4;
  if (out_dtype == DT_F32)
    VL_CUDA_CHECK(launch_kernel(euler_step_kernel<float>, dim3(grid), dim3(128), 0, stream, x, gamma, Wdec, bdec, action, (float*)out, rows, width, A, H,
                                                       horizon, dt, eps, last_step, do_clip, clip, xfix));
  else
    VL_CUDA_CHECK(launch_kernel(euler_step_kernel<__nv_bfloat16>, dim3(grid), dim3(128), 0, stream, x, gamma, Wdec, bdec, action, (__nv_bfloat16*)out, rows,
                                                               width, A, H, horizon, dt, eps, last_step, do_clip, clip, xfix));
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int flow_prepare_launch(const void* x0, const void* x1, const void* t, int io, float sig_min, float time_max_period,
                        const float* W2a, const float* b2, float* action, float* tbias, float* vel, int B, int HA, int he,
                        cudaStream_t stream) {
  VL_CHECK(he % 2 == 0 && he <= 8192, "flow_prepare: bad expert width");
  const float emb = (float)(std::log((double)time_max_period) / (he / 2 - 1));
  VL_CUDA_CHECK(launch_kernel(flow_prepare_kernel, dim3(B), dim3(256), (size_t)he * sizeof(float), stream, x0, x1, t, io, sig_min,
                              emb, W2a, b2, action, tbias, vel, HA, he));
  return 0;
}

int bias_silu_rows_launch(const float* z, const float* tbias, __nv_bfloat16* h, int rows, int H, int width,
                          cudaStream_t stream) {
  const long long n = (long long)rows * width;
  const int grid = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  VL_CUDA_CHECK(launch_kernel(bias_silu_rows_kernel, dim3(grid), dim3(256), 0, stream, z, tbias, h, rows, H, width));
  return 0;
}

int flow_loss_launch(const float* vel, const void* x0, const void* x1, int io, float sig_min, long long n, float* loss,
                     cudaStream_t stream) {
  VL_CUDA_CHECK(launch_kernel(flow_loss_kernel, dim3(1), dim3(256), 0, stream, vel, x0, x1, io, sig_min, n, loss));
  return 0;
}

int last_rows_launch(const float* x, const int* cu, const int* valid_len, float* out, int B, int width, cudaStream_t stream) {
  VL_CHECK(width % 4 == 0, "last_rows: width % 4");
  VL_CUDA_CHECK(launch_kernel(last_rows_kernel, dim3(B), dim3(256), 0, stream, x, cu, valid_len, out, width));
  return 0;
}

int argmax_rows_launch(const float* logits, long long ld, int n, int rows, int* next_tok, int* n_gen, int bump,
                       cudaStream_t stream) {
  VL_CUDA_CHECK(launch_kernel(argmax_rows_kernel, dim3(rows), dim3(1024), 0, stream, logits, ld, n, next_tok, n_gen, bump));
  return 0;
}

int decode_advance_launch(const TextGenBuffers& g, const int* valid_len, const int* kvoff, const __nv_bfloat16* embed, float* x,
                          int B, int width, int window, int gen_cap, long long eos, long long fill, cudaStream_t stream) {
  VL_CHECK(width % 8 == 0, "decode_advance: width % 8");
  VL_CUDA_CHECK(launch_kernel(decode_advance_kernel, dim3(B), dim3(256), 0, stream, g, valid_len, kvoff, embed, x, width, window,
                              gen_cap, eos, fill));
  return 0;
}

}  // namespace vlaser
