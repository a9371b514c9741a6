// Persistent "expert stream" kernel: ONE launch runs a whole pass of the action-expert transformer
// (joint_model.py:124-260 restricted to one mixture; pizero_internvl.py:884-912 calls it once per Euler step)
// over a small row count M (B proprio tokens in the prefill, B*H action tokens in a denoise step):
//
//   for each layer:  RMSNorm -> q|k|v GEMM (+bias, RoPE, KV-cache write) -> attention over the cached prefix KV
//                    -> o-proj (+residual) -> RMSNorm -> gate/up GEMM + SwiGLU -> down-proj (+residual)
//
// At M <= 512 every one of these is a few microseconds of work; launched as separate kernels the pass is bound
// by launch / prologue / drain latency (7 kernels x 28 layers x 10 steps), not by HBM or the tensor pipe. Here the
// 148 CTAs stay resident, keep their smem ring, mbarriers and TMEM allocation across phases, prefetch the next
// phase's weight tiles while the grid barrier between two phases settles, and hand activations to each other
// through L2.
#pragma once
#include "attention.cuh"
#include "gemm.cuh"

namespace vlaser {

constexpr int EXPERT_MAX_STREAMS = 4;

// Per-layer operands, resident in device memory (array of `layers` entries, 64-byte aligned for the tensor maps).
struct alignas(64) ExpertLayerDev {
  CUtensorMap tm_qkv, tm_o, tm_gu, tm_down;  // W operands, [N, K] bf16, box 128 rows x 64
  const float* ln_in;
  const float* ln_post;
  const float* bqkv;
  const void* w_qkv;   // the same four matrices as plain pointers (row-major [N, K] bf16) for the weight-streaming
  const void* w_o;     // single-observation kernel (expert_gemv.cu)
  const void* w_gu;
  const void* w_down;
  const void* pad_[1];
};
static_assert(sizeof(ExpertLayerDev) % 64 == 0, "tensor maps of the next table entry must stay 64-byte aligned");

constexpr int EXPERT_ERR_BARRIER = 0x7ffffff1;  // value left in ExpertPass::err_flag by a timed-out grid barrier

struct ExpertPass {
  const ExpertLayerDev* layers = nullptr;
  int n_layers = 0;
  int last_kv_only = 0;  // 1: the last layer stops after its K/V are cached (joint_model.py:192-193, 690-691)
  int M = 0;             // live rows
  int he = 0, mlp = 0;   // hidden width / MLP width of the expert
  int n_q_heads = 0, n_kv_heads = 0;  // head_dim is 128
  float eps = 1e-6f;
  float* x = nullptr;                 // [M, he] fp32 residual stream (in/out)
  // expert_gemv.cu, deterministic mode: the split-K partial sums of the o / down projections are NOT added to x with
  // fp32 atomics (their order varies from call to call) but accumulated as 2^-30 fixed point in xfix [M, he] with
  // 64-bit integer atomics, which commute exactly. x stays what it was on entry for the whole pass; the stream a
  // reader sees is x + xfix * 2^-30 (one rounding). Zero on entry.
  long long* xfix = nullptr;
  __nv_bfloat16* xn = nullptr;        // [Mpad, he]   normalised rows (GEMM A operand, tmXn)
  __nv_bfloat16* q = nullptr;         // [M, n_q_heads*128]
  __nv_bfloat16* ao = nullptr;        // [Mpad, n_q_heads*128] attention output (tmAo)
  __nv_bfloat16* act = nullptr;       // [Mpad, mlp] SwiGLU output (tmAct)
  // RoPE + KV cache of this pass' rows
  const int* row_pos = nullptr;
  const int* row_kvrow = nullptr;
  const float* cos_tab = nullptr;
  const float* sin_tab = nullptr;
  __nv_bfloat16* k_cache = nullptr;   // layer 0; [layer][kv_head][rows][128]
  __nv_bfloat16* v_cache = nullptr;
  long long kv_layer_stride = 0;      // elements between layers
  long long kv_head_stride = 0;       // elements between kv heads
  // attention segments (per sample)
  const int* q_start = nullptr;
  const int* q_len = nullptr;
  const int* kv_start = nullptr;
  const int* kv_len = nullptr;
  int max_q_len = 0;
  int max_kv_len = 0;                 // upper bound of kv_len[] (sizes the score rows of expert_gemv.cu)
  int batch = 0;
  float scale = 1.f;
  int flags = 0;                      // profiling aids (wrong results!): 1 = drop the residual red.adds, 2 = skip attention,
                                      // 4 = no GEMM tiles, 8 = no norm rows (2|4|8 = the bare grid-barrier skeleton);
                                      // expert_gemv.cu: 2, 4 as above, 16 = no L2 prefetch of the coming phases' weights,
                                      // 32 = only the even warps load across the grid barrier (correct results; measured slower)
  long long* timing = nullptr;        // optional [7 * n_layers]: globaltimer (ns) at which CTA 0 saw each phase complete
  unsigned int* barrier = nullptr;    // grid-barrier arrival counters of this pass, [n_streams][7 * n_layers], zero on entry
  int* err_flag = nullptr;            // set to EXPERT_ERR_BARRIER when a grid barrier gives up (~2 s: a CTA of the grid never
                                      // arrived); the kernel then runs on without waiting and the host reports the call as failed
  int n_streams = 1;                  // independent sub-grids (see expert.cu); needs batch % S == 0, (M / S) % 32 == 0
};

// tmXn / tmAo / tmAct: A-operand maps (box 128 rows x 64) of p.xn / p.ao / p.act; tmX: output map of the fp32 residual
// stream p.x (make_tmap_out, box 32 rows x 32 columns) through which the split-K partial sums are reduce-added.
int expert_pass_launch(const ExpertPass& p, const CUtensorMap& tmXn, const CUtensorMap& tmAo, const CUtensorMap& tmAct,
                       const CUtensorMap& tmX, cudaStream_t stream);

// Same pass for ONE observation with at most 8 rows (B = 1: the proprio row, the H action rows): no tensor cores, the
// expert weights are streamed once with coalesced 16-byte loads, 5 phases per layer (expert_gemv.cu). Uses p.x, p.q,
// p.ao, p.act, the RoPE / KV-cache fields, the per-sample segments of sample 0 and p.barrier ([5 * n_layers] words).
int expert_gemv_pass_launch(const ExpertPass& p, cudaStream_t stream);

}  // namespace vlaser
