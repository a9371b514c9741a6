// bf16 GEMM  C[M,N] = A[M,K] * W[N,K]^T  (nn.Linear layout, both operands K-major) on the
// sm_100a 5th-gen tensor cores: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared-memory ring ->
// tcgen05.mma (single issuing thread, fp32 accumulators in TMEM, double buffered) ->
// tcgen05.ld epilogue fused with the op that follows the projection in the reference
// (bias / GELU / SiLU / SwiGLU / layer-scale + residual / RoPE + KV-cache write / patch-embed).
#pragma once
#include "common.cuh"

namespace vlaser {

enum Epi : int {
  EPI_STORE_BF16 = 0,   // out_bf16[m,n] = acc + bias[n]
  EPI_GELU_BF16 = 1,    // out_bf16[m,n] = gelu_erf(acc + bias[n])                (InternMLP fc1, mlp1[1..2])
  EPI_SILU_BF16 = 2,    // out_bf16[m,n] = silu(acc + bias[n])                    (ActionEncoder linear_2)
  EPI_RESID_F32 = 3,    // out_f32[m,n] += scale[n] * (acc + bias[n])             (o_proj/down_proj/proj/fc2 + residual)
  EPI_SWIGLU_BF16 = 4,  // out_bf16[m,j] = silu(gate_j) * up_j, W rows interleaved in 32-row gate/up groups
  EPI_QKV_ROPE = 5,     // fused q|k|v projection: +bias, rotate-half RoPE on q,k; q -> out, k,v -> KV cache rows
  EPI_PATCH_EMBED = 6,  // out_f32[img*(P+1)+1+p, n] = acc + bias[n] + pos[1+p, n]  (InternVisionEmbeddings)
  EPI_STORE_F32 = 7,    // out_f32[row_map[m], n] = acc + bias[n]
  EPI_COUNT = 8
};

struct EpiParams {
  void* out = nullptr;        // primary output (bf16 or f32, see Epi)
  int ldo = 0;                // leading dimension of out, in elements
  const float* bias = nullptr;
  const float* scale = nullptr;
  const int* row_map = nullptr;
  // EPI_QKV_ROPE
  const int* row_pos = nullptr;     // RoPE position id per row
  const int* row_kvrow = nullptr;   // KV-cache row per row
  const float* cos_tab = nullptr;   // [pos][64]
  const float* sin_tab = nullptr;
  __nv_bfloat16* k_cache = nullptr; // [kv_head][rows][128]
  __nv_bfloat16* v_cache = nullptr;
  long long kv_head_stride = 0;     // elements between kv heads
  int n_q_heads = 0;
  int n_kv_heads = 0;
  // EPI_PATCH_EMBED
  const float* pos_emb = nullptr;   // [(P+1), N]
  int patches = 0;
  // ---- RMSNorm fused across the GEMM chain (expert streams, HF Qwen2RMSNorm):
  //   y = (x * rstd * gamma) W^T  ==  rstd[m] * ((x * gamma) W^T),  rstd[m] = rsqrt(ss[m]/width + eps)
  // producer side (EPI_RESID_F32 / EPI_STORE_F32): besides the fp32 residual row x_new the epilogue writes the NEXT
  // GEMM's A operand a_next[m,n] = bf16(x_new[m,n] * gamma_next[n]) and accumulates ss_next[m] += sum_n x_new[m,n]^2.
  // With split-K (fp32 atomics into x) the CTA that retires the last split of a tile (tile_counters) re-reads the
  // finished tile and does that work.
  const float* gamma_next = nullptr;   // [N]
  __nv_bfloat16* a_next = nullptr;     // [M, lda_next]
  int lda_next = 0;
  float* ss_next = nullptr;            // [M], zeroed by the caller
  int* tile_counters = nullptr;        // [m_tiles*n_tiles] zero-initialised, self-resetting (split-K only)
  // consumer side (any epilogue): acc is multiplied by rstd[m] before bias / activation.
  // ss_slots > 0 (CTA-pair kernel, large M: the LLM prefill): the sum of squares of a row is kept as ss_slots partial sums,
  // one per (256-column tile, column half) of the producing GEMM, written with plain stores by the producer's transposed
  // read-modify-write epilogue (epilogue_resid_rmw_pair) and added in slot order by the consumer: no atomics, bit-
  // reproducible. ss_next / row_ss then point at [M][ss_slots].
  int ss_slots = 0;
  const float* row_ss = nullptr;       // [M] sum of squares of the un-normalised row
  float inv_width = 0.f;               // 1 / hidden width of the normalised row
  float eps = 0.f;
  // ---- deterministic split-K (EPI_RESID_F32). With `part` set the K splits of a tile do not meet through fp32 atomics /
  // L2 reduce-adds in `out` (whose summation order changes from call to call): split s stores its partial tile -- scaled,
  // split 0 biased -- to part[s * part_stride + row * ldo + col], and the consumer of the residual stream (the norm
  // kernels of elementwise.cu, splitk_reduce) adds the partials to `out` in split order. part_splits receives the
  // split count the launch actually used (0: one split, `out` was updated directly).
  float* part = nullptr;
  long long part_stride = 0;           // floats between consecutive splits, a multiple of 256 * ldo covering the rows
  int* part_splits = nullptr;
  // ---- fused tail (pair kernel, 256 x 128 tiles, EPI_RESID_F32 with `part`: the o / down projections of the action expert
  // at one 256-row tile). The launch is one dependent hop of a latency-bound chain and so is the norm launch behind it;
  // with tail_ctr set every CTA, once its partial tiles have landed, passes an in-kernel grid barrier (all CTAs of the
  // launch are resident: one per SM, at most gemm_num_sms() of them) and then finishes rows blockIdx.x, + gridDim.x, ...:
  // out[row] += partials in split order (the same association as resid_norm_kernel), then RMSNorm(out[row]) * tail_gamma
  // -> tail_out (bf16) -- the work of the norm launch that would follow, without its hop.
  // ---- row-block completion counters (pair kernel, EPI_RESID_F32 through the TMA reduce-add epilogue, one split): every
  // epilogue warp, once the reduce-adds of a tile have landed, bumps done_ctr[m_idx] (release, gpu scope); a 256-row block
  // is complete at (N / 256) * 2 * epilogue-warps arrivals. A norm launch behind this GEMM -- its programmatic dependent --
  // then works under the GEMM, row block by row block, instead of behind it (norm_kernel `ready`).
  int* done_ctr = nullptr;             // [ceil(M / 256)], zero on entry
  unsigned int* tail_ctr = nullptr;    // arrival counter of this launch, zero on entry
  const float* tail_gamma = nullptr;   // [N] gain of the following RMSNorm
  __nv_bfloat16* tail_out = nullptr;   // [M, ldo]
  float tail_eps = 0.f;
  int* tail_err = nullptr;             // receives GEMM_ERR_TAIL_BARRIER if the barrier gives up (~1 s)
};
constexpr int GEMM_ERR_TAIL_BARRIER = 0x7ffffff2;

// K splits of the pair kernel: as many as the idle pairs allow, at least two k-blocks each, evened out
__host__ __device__ inline int pair_split_count(int base, int n_pairs, int k_blocks, int max_splits) {
  int splits = 1;
  if (max_splits > 1 && base > 0) {
    splits = n_pairs / base;
    if (splits > max_splits) splits = max_splits;
    if (splits > k_blocks / 2) splits = k_blocks / 2;
    if (splits < 1) splits = 1;
  }
  const int kb_per_split = (k_blocks + splits - 1) / splits;
  return (k_blocks + kb_per_split - 1) / kb_per_split;
}

struct GemmDims {
  int M = 0, N = 0, K = 0;       // K % 64 == 0, N % block_n == 0
  const int* M_dev = nullptr;    // optional: live row count in device memory (M is then the capacity)
  int max_splits = 1;            // split-K cap (EPI_RESID_F32 only)
  int stages = 0;                // smem ring depth; 0 = default for the tile shape
  int pair = 0;                  // 1: CTA-pair kernel (cta_group::2, 256x256 tile per pair; W map with 128-row boxes)
  int pair_bn = 256;             // pair kernel tile width: 256, or 128 (residual epilogue only; W map with 64-row boxes): half the
                                 // output bytes per CTA and twice the CTAs for the one-row-tile split-K projections
  int half_sm = 0;               // pair kernel: 1 = the half-SM variant (static M, one tile per CTA pair; see gemm.cu)
  int gelu_warps = 16;           // pair kernel, exact-GELU epilogue: epilogue warps (16, or 8 with one more ring stage)
  int n_fast = 0;                // pair kernel tile order: 0 = m fastest, 1 = n fastest
  int tma_out = 0;               // set by the launcher: epilogue leaves through the output tensor map
  long long* trace = nullptr;    // optional 8-slot timeline record of this launch (common.cuh trace_stamp)
};

// Host API -------------------------------------------------------------------------------
// Encodes a 2-D bf16 row-major [rows, cols] tensor map with a [box_rows x 64] box, 128B swizzle.
int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
// Launches the tcgen05 kernel (block_n in {128,256}); `simt` launches the CUDA-core cross-check kernel
// (debug / tests only) that shares the epilogue code but reads A and W through plain pointers.
// Output tensor map for the pair kernel's TMA epilogue: row-major [rows, cols] bf16 (elem_bytes 2) or f32 (4), box =
// 32 rows x 128 bytes, 128B swizzle. For EPI_SWIGLU_BF16 it describes the [M, N/2] result.
int make_tmap_out(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, int elem_bytes);
// `tmC` (optional, pair kernel + bf16/residual epilogues only): output tensor map; the epilogue then stores through TMA.
// `tmP` (pair kernel, split-K with ep.part): make_tmap_out map of the partial workspace viewed as
// [max_splits * part_stride / ldo, ldo] f32.
int gemm_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDims& d, const EpiParams& ep, int block_n,
                int epi, cudaStream_t stream, const CUtensorMap* tmC = nullptr, const CUtensorMap* tmP = nullptr);
int gemm_launch_simt(const __nv_bfloat16* A, const __nv_bfloat16* W, const GemmDims& d, const EpiParams& ep, int epi,
                     cudaStream_t stream);
int gemm_num_sms();
// Split count a CTA-pair launch (d.pair, EPI_RESID_F32, TMA epilogue) with a STATIC row count d.M will use and report through
// EpiParams::part_splits (0: one split, `out` updated directly) -- the arithmetic of the kernel, for callers that want
// the count at launch time.
int gemm_pair_split_count(const GemmDims& d, bool with_partials);

}  // namespace vlaser
