// Input staging and action post-processing: the two steps either side of `infer_action`.
//
// stage_frames_kernel — camera frames u8 (HWC as the simulator hands them over, or CHW) -> optional centre crop
//   (simpler.py:72-79) -> 8-bit Lanczos4 resize, bit-exact with `cv2.resize(.., INTER_LANCZOS4)` (simpler.py:80-84:
//   taps from OpenCV's interpolateLanczos4 scaled by 2^11 to int16, replicate border, int32 accumulation of the
//   horizontal then the vertical pass, (sum + 2^21) >> 22 saturated to u8) -> `x * (1/255)` -> `(x - mean) / std`
//   (processing.py:27-63) -> bf16 / f32 NCHW pixel_values. HBM-bound integer work: every source byte is read once
//   into a shared-memory window per 32x8 output tile (coalesced rows), the two separable passes run out of shared
//   memory, the output planes are written as contiguous rows.
// postprocess_actions_kernel — normalised action chunks -> environment actions (simpler.py:136-177): bound / gaussian
//   de-normalisation (base.py:21-49, first two ops in float32 exactly as numpy promotes them), euler (sxyz) ->
//   axis-angle (geometry.py:261-434), gripper binarisation (Bridge, :217-221) or the sticky-gripper state machine
//   (EDR, :259-286) with per-environment state carried across calls. One thread per environment (the gripper state is
//   sequential along the chunk); float64 like the reference's numpy.
#include <math.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "../../include/vlaser_b200.h"

namespace vlaser {
namespace {

constexpr int ST_TX = 32, ST_TY = 8;      // output tile
constexpr int ST_THREADS = 256;
constexpr int COEF_BITS = 11;

// ---------------------------------------------------------------------------------------------- tap tables (host)
// OpenCV interpolateLanczos4(float x, float* coeffs), restated
void lanczos4_coeffs(float x, float* coeffs) {
  static const double s45 = 0.70710678118654752440084436210485;
  static const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  const double pi = 3.1415926535897932384626433832795;
  float sum = 0;
  const double y0 = -(x + 3) * pi * 0.25, s0 = sin(y0), c0 = cos(y0);
  for (int i = 0; i < 8; i++) {
    const float y0_ = (x + 3 - i);
    if (fabs(y0_) >= 1e-6f) {
      const double y = -y0_ * pi * 0.25;
      coeffs[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
    } else {
      coeffs[i] = 1e30f;
    }
    sum += coeffs[i];
  }
  sum = 1.f / sum;
  for (int i = 0; i < 8; i++) coeffs[i] *= sum;
}

struct AxisTable {
  int32_t* ofs = nullptr;   // [dst] first tap index in the source axis (may be < 0 / run past the end: taps clamp)
  int16_t* w = nullptr;     // [dst][8]
  int max_span_x = 0;       // largest source extent covered by ST_TX consecutive outputs
  int max_span_y = 0;       // ... by ST_TY consecutive outputs
};

int build_axis_table(int src, int dst, AxisTable* out) {
  std::vector<int32_t> ofs(dst);
  std::vector<int16_t> w((size_t)dst * 8);
  const double scale = (double)src / (double)dst;
  for (int d = 0; d < dst; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    const int s = (int)floorf(f);
    f -= s;
    float c[8];
    lanczos4_coeffs(f, c);
    ofs[d] = s - 3;
    for (int k = 0; k < 8; ++k) {
      const long r = lrintf(c[k] * (float)(1 << COEF_BITS));  // saturate_cast<short>: round half to even
      w[(size_t)d * 8 + k] = (int16_t)(r < -32768 ? -32768 : r > 32767 ? 32767 : r);
    }
  }
  auto span = [&](int tile) {
    int m = 0;
    for (int d0 = 0; d0 < dst; d0 += tile) {
      const int d1 = std::min(dst, d0 + tile) - 1;
      m = std::max(m, ofs[d1] + 8 - ofs[d0]);
    }
    return m;
  };
  out->max_span_x = span(ST_TX);
  out->max_span_y = span(ST_TY);
  VL_CUDA_CHECK(cudaMalloc(&out->ofs, sizeof(int32_t) * dst));
  VL_CUDA_CHECK(cudaMalloc(&out->w, sizeof(int16_t) * dst * 8));
  VL_CUDA_CHECK(cudaMemcpy(out->ofs, ofs.data(), sizeof(int32_t) * dst, cudaMemcpyHostToDevice));
  VL_CUDA_CHECK(cudaMemcpy(out->w, w.data(), sizeof(int16_t) * dst * 8, cudaMemcpyHostToDevice));
  return 0;
}

// tables are a pure function of (source extent, destination extent): cached per device for the process lifetime
std::mutex g_tab_mu;
std::map<std::tuple<int, int, int>, AxisTable> g_tabs;

int get_axis_table(int src, int dst, const AxisTable** out) {
  int dev = 0;
  VL_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_tab_mu);
  const auto key = std::make_tuple(dev, src, dst);
  auto it = g_tabs.find(key);
  if (it == g_tabs.end()) {
    AxisTable t;
    const int rc = build_axis_table(src, dst, &t);
    if (rc) return rc;
    it = g_tabs.emplace(key, t).first;
  }
  *out = &it->second;
  return 0;
}

// ---------------------------------------------------------------------------------------------- staging kernel
struct StageParams {
  const uint8_t* frames;   // [n][H][W][3] (layout 0) or [n][3][H][W] (layout 1)
  void* out;               // [n][3][px][px]
  int n, H, W, layout;
  int cx0, cy0, cw, ch;    // crop window inside the frame (the resize sees a cw x ch image)
  int px, out_bf16, identity;
  const int32_t* xofs;
  const int16_t* xw;
  const int32_t* yofs;
  const int16_t* yw;
  int span_x, span_y;      // shared-memory window extent
  float mean[3], std[3];
};

__device__ __forceinline__ float normalise(int v, float mean, float sd) {
  // processing.py:27-48 in float32, one rounding per op (no FMA contraction)
  return __fdiv_rn(__fsub_rn(__fmul_rn((float)v, 1.0f / 255.0f), mean), sd);
}

template <typename OutT>
__device__ __forceinline__ void store_px(void* out, size_t idx, float v) {
  if constexpr (sizeof(OutT) == 2) reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(out)[idx] = v;
}

template <typename OutT>
__global__ void __launch_bounds__(ST_THREADS) stage_frames_kernel(const StageParams p) {
  extern __shared__ __align__(16) uint8_t st_smem[];
  const int img = blockIdx.z;
  const int ox0 = blockIdx.x * ST_TX, oy0 = blockIdx.y * ST_TY;
  const int tid = threadIdx.x;
  const size_t plane = (size_t)p.px * p.px;
  if (p.identity) {
    // frame already at the target size: cv2.resize returns a copy
    for (int i = tid; i < ST_TX * ST_TY * 3; i += ST_THREADS) {
      const int c = i / (ST_TX * ST_TY), r = i - c * ST_TX * ST_TY;
      const int oy = oy0 + r / ST_TX, ox = ox0 + r % ST_TX;
      if (oy >= p.px || ox >= p.px) continue;
      const int sy = p.cy0 + oy, sx = p.cx0 + ox;
      const uint8_t v = p.layout == 0 ? p.frames[(((size_t)img * p.H + sy) * p.W + sx) * 3 + c]
                                      : p.frames[(((size_t)img * 3 + c) * p.H + sy) * p.W + sx];
      store_px<OutT>(p.out, ((size_t)img * 3 + c) * plane + (size_t)oy * p.px + ox, normalise(v, p.mean[c], p.std[c]));
    }
    return;
  }
  // shared memory: source window u8 [span_y][span_x][3] | horizontal sums i32 [span_y][ST_TX][3]
  uint8_t* win = st_smem;
  const int win_bytes = (p.span_y * p.span_x * 3 + 15) & ~15;
  int32_t* hor = reinterpret_cast<int32_t*>(st_smem + win_bytes);
  const int wx0 = p.xofs[ox0], wy0 = p.yofs[oy0];
  const int ox_last = min(p.px, ox0 + ST_TX) - 1, oy_last = min(p.px, oy0 + ST_TY) - 1;
  const int wcols = p.xofs[ox_last] + 8 - wx0, wrows = p.yofs[oy_last] + 8 - wy0;
  // ---- window load, replicate border (clamp to the crop window), coalesced along the source row
  if (p.layout == 0) {
    const int row_bytes = wcols * 3;
    for (int i = tid; i < wrows * row_bytes; i += ST_THREADS) {
      const int r = i / row_bytes, b = i - r * row_bytes;
      const int col = b / 3, c = b - col * 3;
      const int sy = p.cy0 + min(max(wy0 + r, 0), p.ch - 1), sx = p.cx0 + min(max(wx0 + col, 0), p.cw - 1);
      win[(r * p.span_x + col) * 3 + c] = p.frames[(((size_t)img * p.H + sy) * p.W + sx) * 3 + c];
    }
  } else {
    for (int i = tid; i < 3 * wrows * wcols; i += ST_THREADS) {
      const int c = i / (wrows * wcols), rem = i - c * wrows * wcols;
      const int r = rem / wcols, col = rem - r * wcols;
      const int sy = p.cy0 + min(max(wy0 + r, 0), p.ch - 1), sx = p.cx0 + min(max(wx0 + col, 0), p.cw - 1);
      win[(r * p.span_x + col) * 3 + c] = p.frames[(((size_t)img * 3 + c) * p.H + sy) * p.W + sx];
    }
  }
  __syncthreads();
  // ---- horizontal pass: every window row x ST_TX outputs x 3 channels, 8 taps, int32
  for (int i = tid; i < wrows * ST_TX; i += ST_THREADS) {
    const int r = i / ST_TX, t = i - r * ST_TX;
    const int ox = ox0 + t;
    int a0 = 0, a1 = 0, a2 = 0;
    if (ox < p.px) {
      const int c0 = p.xofs[ox] - wx0;
      const int16_t* w = p.xw + (size_t)ox * 8;
      const uint8_t* src = win + (r * p.span_x + c0) * 3;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int wk = w[k];
        a0 += wk * src[3 * k + 0];
        a1 += wk * src[3 * k + 1];
        a2 += wk * src[3 * k + 2];
      }
    }
    int32_t* h = hor + (r * ST_TX + t) * 3;
    h[0] = a0;
    h[1] = a1;
    h[2] = a2;
  }
  __syncthreads();
  // ---- vertical pass + rounding + normalisation; thread -> (channel, row, x) with x fastest (row-contiguous stores)
  for (int i = tid; i < 3 * ST_TY * ST_TX; i += ST_THREADS) {
    const int c = i / (ST_TY * ST_TX), rem = i - c * ST_TY * ST_TX;
    const int ty = rem / ST_TX, t = rem - ty * ST_TX;
    const int oy = oy0 + ty, ox = ox0 + t;
    if (oy >= p.px || ox >= p.px) continue;
    const int r0 = p.yofs[oy] - wy0;
    const int16_t* w = p.yw + (size_t)oy * 8;
    int acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += (int)w[k] * hor[((r0 + k) * ST_TX + t) * 3 + c];
    int v = (acc + (1 << (2 * COEF_BITS - 1))) >> (2 * COEF_BITS);
    v = min(max(v, 0), 255);
    store_px<OutT>(p.out, ((size_t)img * 3 + c) * plane + (size_t)oy * p.px + ox, normalise(v, p.mean[c], p.std[c]));
  }
}

// ---------------------------------------------------------------------------------------------- post-processing
struct PostParams {
  const void* actions;     // [n_env][T][A] normalised
  int in_bf16;
  int n_env, T, A;
  int norm;                // 0 bound (p01/p99), 1 gaussian (mean/std)
  const double* stat_a;    // [A] p01 | mean
  const double* stat_b;    // [A] p99 | std
  int gripper;             // 0 Bridge binarisation, 1 EDR sticky gripper
  int sticky_repeat;
  double* sticky_state;    // [n_env][3] = {on, repeat count, held value}; read and updated
  double* out;             // [n_env][T][7]
};

__global__ void postprocess_actions_kernel(const PostParams p) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.n_env) return;
  const double eps = 2.220446049250313e-16;
  double on = 0.0, rep = 0.0, held = 0.0;
  if (p.gripper == 1) {
    on = p.sticky_state[e * 3 + 0];
    rep = p.sticky_state[e * 3 + 1];
    held = p.sticky_state[e * 3 + 2];
  }
  for (int t = 0; t < p.T; ++t) {
    const size_t base = ((size_t)e * p.T + t) * p.A;
    double raw[6];
    float last = 0.f;
    for (int i = 0; i < p.A; ++i) {
      const float a = p.in_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.actions)[base + i])
                                : reinterpret_cast<const float*>(p.actions)[base + i];
      if (i == p.A - 1) {
        last = a;  // the gripper channel is not normalised (simpler.py:139)
      } else if (i < 6) {
        if (p.norm == 0) {
          // base.py:21-31 with numpy promotion: (float32 - (-1)) / 2 stay float32, the rest is float64
          const float u = __fdiv_rn(__fsub_rn(a, -1.0f), 2.0f);
          raw[i] = __dadd_rn(__dmul_rn((double)u, p.stat_b[i] - p.stat_a[i]), p.stat_a[i]);  // (no FMA contraction)
        } else {
          raw[i] = __dadd_rn(__dmul_rn((double)a, p.stat_b[i] + 1e-8), p.stat_a[i]);  // base.py:42-49
        }
      }
    }
    // ---- euler (sxyz) -> quaternion -> axis * angle (geometry.py:294-434)
    const double ai = raw[3] / 2.0, aj = raw[4] / 2.0, ak = raw[5] / 2.0;
    const double ci = cos(ai), si = sin(ai), cj = cos(aj), sj = sin(aj), ck = cos(ak), sk = sin(ak);
    const double cc = ci * ck, cs = ci * sk, sc = si * ck, ss = si * sk;
    double q0 = cj * cc + sj * ss, q1 = cj * sc - sj * cs, q2 = cj * ss + sj * cc, q3 = cj * cs - sj * sc;
    double ax = 1.0, ay = 0.0, az = 0.0, theta = 0.0;
    const double Nq = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
    if (!isfinite(Nq)) {
      theta = nan("");
    } else if (Nq >= eps * eps) {
      if (Nq != 1.0) {
        const double s = sqrt(Nq);
        q0 /= s, q1 /= s, q2 /= s, q3 /= s;
      }
      const double len2 = q1 * q1 + q2 * q2 + q3 * q3;
      if (len2 >= (eps * 3) * (eps * 3)) {
        theta = 2.0 * acos(fmax(fmin(q0, 1.0), -1.0));
        const double l = sqrt(len2);
        ax = q1 / l, ay = q2 / l, az = q3 / l;
      }
    }
    // ---- gripper
    double g;
    if (p.gripper == 0) {
      g = 2.0 * ((double)last > 0.5 ? 1.0 : 0.0) - 1.0;       // simpler.py:217-221
    } else {
      const double act = ((double)last * 2.0) - 1.0;           // simpler.py:259-286
      double rel = -act;
      if (fabs(rel) > 0.5 && on == 0.0) {
        on = 1.0;
        held = rel;
      }
      if (on != 0.0) {
        rep += 1.0;
        rel = held;
      }
      if (rep == (double)p.sticky_repeat) {
        on = 0.0;
        rep = 0.0;
        held = 0.0;
      }
      g = rel;
    }
    double* o = p.out + ((size_t)e * p.T + t) * 7;
    o[0] = raw[0];
    o[1] = raw[1];
    o[2] = raw[2];
    o[3] = ax * theta;
    o[4] = ay * theta;
    o[5] = az * theta;
    o[6] = g;
  }
  if (p.gripper == 1) {
    p.sticky_state[e * 3 + 0] = on;
    p.sticky_state[e * 3 + 1] = rep;
    p.sticky_state[e * 3 + 2] = held;
  }
}

// BaseEnvAdapter (de)normalisation (Vlaser_VLA/Simpler/src/agent/env_adapter/base.py:8-49), elementwise over [rows, dim] in
// float64 as numpy computes it: bound (stat_a = p01, stat_b = p99): 2 (x - a) / (b - a + 1e-8) - 1 clipped to [-1, 1], inverse
// (x + 1) / 2 (b - a) + a; gaussian (stat_a = mean, stat_b = std): (x - a) / (b + 1e-8), inverse x (b + 1e-8) + a.
// dtype codes: 0 f32, 1 bf16, 2 f64.
__device__ __forceinline__ double load_any(const void* p, int dtype, long long i) {
  if (dtype == 2) return reinterpret_cast<const double*>(p)[i];
  if (dtype == 1) return (double)__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return (double)reinterpret_cast<const float*>(p)[i];
}
__global__ void normalize_kernel(const void* __restrict__ in, int in_dtype, void* __restrict__ out, int out_dtype, long long n,
                                 int dim, int norm, int inverse, const double* __restrict__ stat_a,
                                 const double* __restrict__ stat_b) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = (int)(i % dim);
  const double x = load_any(in, in_dtype, i), a = stat_a[d], b = stat_b[d];
  double y;
  if (norm == 0) {
    if (!inverse) y = fmin(fmax(2.0 * (x - a) / (b - a + 1e-8) - 1.0, -1.0), 1.0);
    else y = (x + 1.0) / 2.0 * (b - a) + a;
  } else {
    y = inverse ? x * (b + 1e-8) + a : (x - a) / (b + 1e-8);
  }
  if (out_dtype == 2) reinterpret_cast<double*>(out)[i] = y;
  else if (out_dtype == 1) reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16((float)y);
  else reinterpret_cast<float*>(out)[i] = (float)y;
}

}  // namespace
}  // namespace vlaser

using namespace vlaser;

extern "C" {

int32_t vlaser_op_stage_frames(const uint8_t* frames, int32_t n, int32_t src_h, int32_t src_w, int32_t layout,
                               int32_t center_crop, void* pixel_values, int32_t out_dtype, int32_t px, const float* mean3,
                               const float* std3, void* stream) {
  VL_CHECK(frames && pixel_values && n > 0 && src_h > 0 && src_w > 0 && px > 0, "stage_frames: bad argument");
  VL_CHECK(layout == 0 || layout == 1, "stage_frames: layout must be 0 (HWC) or 1 (CHW)");
  VL_CHECK(out_dtype == VLASER_F32 || out_dtype == VLASER_BF16, "stage_frames: out dtype must be f32 or bf16");
  StageParams p = {};
  p.frames = frames;
  p.out = pixel_values;
  p.n = n, p.H = src_h, p.W = src_w, p.layout = layout;
  p.cw = src_w, p.ch = src_h;
  if (center_crop) {  // simpler.py:72-79
    const int c = std::min(src_h, src_w);
    p.cx0 = (src_w - c) / 2, p.cy0 = (src_h - c) / 2, p.cw = c, p.ch = c;
  }
  p.px = px;
  p.out_bf16 = out_dtype == VLASER_BF16;
  p.identity = (p.cw == px && p.ch == px);
  static const float kMean[3] = {0.485f, 0.456f, 0.406f}, kStd[3] = {0.229f, 0.224f, 0.225f};  // processing.py:303-304
  for (int c = 0; c < 3; ++c) p.mean[c] = mean3 ? mean3[c] : kMean[c], p.std[c] = std3 ? std3[c] : kStd[c];
  size_t smem = 0;
  if (!p.identity) {
    const AxisTable *tx = nullptr, *ty = nullptr;
    int rc = get_axis_table(p.cw, px, &tx);
    if (rc) return rc;
    rc = get_axis_table(p.ch, px, &ty);
    if (rc) return rc;
    p.xofs = tx->ofs, p.xw = tx->w, p.yofs = ty->ofs, p.yw = ty->w;
    p.span_x = tx->max_span_x, p.span_y = ty->max_span_y;
    smem = (((size_t)p.span_y * p.span_x * 3 + 15) & ~(size_t)15) + (size_t)p.span_y * ST_TX * 3 * sizeof(int32_t);
    VL_CHECK(smem <= 200 * 1024, "stage_frames: down-scaling factor too large for the shared-memory window");
  }
  const dim3 grid((px + ST_TX - 1) / ST_TX, (px + ST_TY - 1) / ST_TY, n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.out_bf16) {
    if (smem > 48 * 1024)
      VL_CUDA_CHECK(cudaFuncSetAttribute(stage_frames_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stage_frames_kernel<__nv_bfloat16><<<grid, ST_THREADS, smem, st>>>(p);
  } else {
    if (smem > 48 * 1024)
      VL_CUDA_CHECK(cudaFuncSetAttribute(stage_frames_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stage_frames_kernel<float><<<grid, ST_THREADS, smem, st>>>(p);
  }
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int32_t vlaser_op_normalize(const void* in, int32_t in_dtype, void* out, int32_t out_dtype, int64_t rows, int32_t dim,
                            int32_t norm, int32_t inverse, const double* stat_a, const double* stat_b, void* stream) {
  VL_CHECK(in && out && stat_a && stat_b && rows > 0 && dim > 0, "normalize: bad argument");
  VL_CHECK(in_dtype >= 0 && in_dtype <= 2 && out_dtype >= 0 && out_dtype <= 2, "normalize: dtype must be f32 (0), bf16 (1) or f64 (2)");
  VL_CHECK(norm == 0 || norm == 1, "normalize: norm must be 0 (bound) or 1 (gaussian)");
  const long long n = (long long)rows * dim;
  normalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, in_dtype, out, out_dtype, n, dim, norm, inverse, stat_a, stat_b);
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int32_t vlaser_op_postprocess_actions(const void* actions, int32_t dtype, int32_t n_env, int32_t T, int32_t A, int32_t norm,
                                      const double* stat_a, const double* stat_b, int32_t gripper, int32_t sticky_repeat,
                                      double* sticky_state, double* out, void* stream) {
  VL_CHECK(actions && out && stat_a && stat_b && n_env > 0 && T > 0, "postprocess_actions: bad argument");
  VL_CHECK(A == 7, "postprocess_actions: the SimplerEnv action is xyz + euler + gripper (7)");
  VL_CHECK(dtype == VLASER_F32 || dtype == VLASER_BF16, "postprocess_actions: dtype must be f32 or bf16");
  VL_CHECK(norm == 0 || norm == 1, "postprocess_actions: norm must be 0 (bound) or 1 (gaussian)");
  VL_CHECK(gripper == 0 || (gripper == 1 && sticky_state && sticky_repeat > 0), "postprocess_actions: bad gripper mode");
  PostParams p = {};
  p.actions = actions, p.in_bf16 = dtype == VLASER_BF16;
  p.n_env = n_env, p.T = T, p.A = A, p.norm = norm;
  p.stat_a = stat_a, p.stat_b = stat_b;
  p.gripper = gripper, p.sticky_repeat = sticky_repeat, p.sticky_state = sticky_state;
  p.out = out;
  postprocess_actions_kernel<<<(n_env + 63) / 64, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  VL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
