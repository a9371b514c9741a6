// Memory-bound kernels of the hot path (HBM roofline): norms, pixel-shuffle gather, im2col,
// token planning / embedding assembly, the action encoder head and the fused
// final-norm + action decoder + Euler update.
#pragma once
#include "common.cuh"

namespace vlaser {

enum DType : int { DT_F32 = 0, DT_BF16 = 1 };

// ---- norms: fp32 residual stream in, bf16 GEMM operand out. width % 128 == 0, width <= 4096.
// LayerNorm (nn.LayerNorm, two-pass statistics) and HF Qwen2RMSNorm.
struct ShuffleGather {   // pixel_shuffle(0.5) gather fused into the projector's LayerNorm
  int grid = 0;          // g (patch grid side); 0 = no gather
  int channels = 0;      // C (ViT hidden)
};
// Deterministic split-K (see EpiParams::part in gemm.cuh): the partial tiles a split-K residual projection left in its
// workspace. A norm launch that is handed them first finishes the residual update, x[row] += part[0][row] + part[1][row]
// + ... in split order (one writer per row, so the sum does not depend on scheduling), stores the new x and then
// normalises it. `splits` is the device word the GEMM wrote (0 = nothing to add).
struct SplitKPartials {
  const float* part = nullptr;
  long long stride = 0;        // floats between consecutive splits
  const int* splits = nullptr;
  int max_splits = 0;          // host-side cap of *splits for this launch (0: unknown)
  int known_splits = -1;       // >= 0: the split count itself, derived on the host (static row count: gemm_pair_split_count);
                               // the norm launch then issues its partial loads without first reading *splits (one L2 round
                               // trip less on the dependent chain of the action expert)
};
void set_resid_norm_small(int on);  // measurement switch, see elementwise.cu
// `ready` (optional, plain LayerNorm of a static row count only): row-block completion counters of the residual GEMM this
// launch is the programmatic dependent of (EpiParams::done_ctr). The kernel then does NOT wait for the GEMM to finish: the
// warp of a row waits until ready[row / 256] has reached `ready_expect` and normalises the row while the GEMM is still
// working on later row blocks.
struct RowBlocksReady {
  const int* ctr = nullptr;
  int expect = 0;
  int* err = nullptr;  // receives NORM_ERR_READY_TIMEOUT if a counter never arrives (~1 s)
};
constexpr int NORM_ERR_READY_TIMEOUT = 0x7ffffff3;
int layernorm_launch(float* x, const float* gamma, const float* beta, __nv_bfloat16* out, int rows, int width,
                     float eps, ShuffleGather sg, cudaStream_t stream, const SplitKPartials* sp = nullptr,
                     const RowBlocksReady* ready = nullptr);
// x[row] += sum of the partials (no norm): for consumers of the residual stream that are not a plain norm launch
int splitk_reduce_launch(float* x, int rows, const int* rows_dev, int width, const SplitKPartials& sp, cudaStream_t stream);
// Optional HBM -> L2 prefetch riding on a norm launch: before its dependent-launch wait, thread 0 of every CTA issues
// bulk L2 prefetches (cp.async.bulk.prefetch.L2, <= 32 KB each) for its slice of up to three constant byte ranges —
// the weights of the GEMMs a few launches further down the chain, which would otherwise start on cold HBM lines
// (the 1.3 GB of expert weights streamed per pass flush the 126 MB L2 between passes).
struct L2Prefetch {
  const void* ptr[3] = {nullptr, nullptr, nullptr};
  unsigned long long bytes[3] = {0, 0, 0};  // multiples of 16
};
int rmsnorm_launch(float* x, const float* gamma, __nv_bfloat16* out, int rows, const int* rows_dev, int width,
                   float eps, cudaStream_t stream, const L2Prefetch* pf = nullptr, long long* trace = nullptr,
                   const SplitKPartials* sp = nullptr);

// ---- pixel_shuffle (pizero_internvl.py:706-716), pure re-index, bit exact. elem_bytes in {2,4}.
int pixel_shuffle_launch(const void* in, void* out, int n, int grid, int channels, int elem_bytes, cudaStream_t stream);

// ---- patch embedding support (modeling_intern_vit.py:162-174)
// pixel [n,3,px,px] (f32|bf16) -> cols bf16 [n*P, kpad], column = ch*ps*ps + ky*ps + kx, zero padded.
int im2col_launch(const void* pixel, int pixel_dtype, __nv_bfloat16* cols, int n, int px, int ps, int kpad,
                  cudaStream_t stream);
// x[img*(P+1), :] = cls + pos[0]
int cls_rows_launch(float* x, const float* cls, const float* pos, int n, int tokens, int width, cudaStream_t stream);

// ---- prefix planning (valid lengths, packed offsets, row tables) — replaces the dense mask /
// position-id tensors of build_causal_mask_and_position_ids (pizero_internvl.py:517-587)
struct PlanBuffers {
  int* valid_len;     // [B]
  int* cu;            // [B+1] packed-row offsets
  int* kvoff;         // [B]   KV-cache row offsets: cu[b] + b*(1+H)
  int* m_tot;         // [1]
  int* row_pos;       // [B*S] RoPE position per packed row (j+1)
  int* row_kvrow;     // [B*S]
  int* row_src;       // [B*S] >=0 embed row, -1 zeros, -2 image row (written by the projector epilogue)
  int* img_row_map;   // [B*n_img*T] projector output row -> packed row (or -1); rows are numbered in batch order, so a
                      // sample owns as many image tiles as its ids hold image tokens (dynamic tiles), like the
                      // reference's masked scatter (pizero_internvl.py:757-791)
  int* img_cnt;       // [B] image tokens among the valid ids of a sample
  int* img_off;       // [B] exclusive scan of img_cnt
  int* prop_pos;      // [B] == 1
  int* prop_kvrow;    // [B]
  int* act_pos;       // [B*H] 2..H+1
  int* act_kvrow;     // [B*H]
  int* pre_kv_len;    // [B]  == valid_len
  int* den_q_start;   // [B]
  int* den_q_len;     // [B]
  int* den_kv_len;    // [B]  valid_len + 1 + H
  int* prop_q_start;  // [B]  == b   (proprio pass of the expert stream: one query row per sample)
  int* prop_q_len;    // [B]  == 1
  int* prop_kv_len;   // [B]  valid_len + 1 (prefix + the proprio token itself)
  int* bad_id;        // [1]  0, or 1 + the first token id seen outside [0, embed_rows) (nn.Embedding would raise IndexError;
                      //      here the row embeds as zeros and the call reports the error, engine.cu)
};
// text_kv_stride > 0: the chat-VLM layout -- sample b owns K/V rows [b * stride, (b + 1) * stride) and RoPE positions start
// at 0 (HF Qwen2) instead of 1
int plan_launch(const long long* input_ids, const int* valid_len_in, PlanBuffers pb, int B, int S, int H,
                int img_tokens_per_sample, int image_token, int pad_token, int embed_rows, cudaStream_t stream,
                int text_kv_stride = 0);
// text rows of the packed prefix: x[m,:] = embed[row_src[m],:] (bf16 -> fp32) or zeros
int assemble_text_launch(const __nv_bfloat16* embed, const int* row_src, const int* m_tot, float* x, int rows_cap,
                         int width, cudaStream_t stream);

// ---- small dense heads on CUDA cores (K = action/proprio dim <= 32)
// out[r,n] = bias[n] + sum_k W[n,k] * in[r,k]   (in f32|bf16; out fp32 or bf16)
int small_linear_launch(const void* in, int in_dtype, const float* W, const float* bias, void* out, int out_dtype,
                        int rows, int K, int N, cudaStream_t stream);
// convert/copy a [n] vector between dtypes (noise ingest)
int convert_launch(const void* in, int in_dtype, void* out, int out_dtype, long long n, cudaStream_t stream);

// ---- final RMSNorm + action decoder + Euler step (pizero_internvl.py:911-912, 927-934)
// action[r,:] += dt * (Wdec * RMS(x[r,:], gamma) + bdec); on the last step also writes
// out[b, i-(H-horizon), :] = clamp(action) for the last `horizon` tokens.
int euler_step_launch(const float* x, const float* gamma, const float* Wdec, const float* bdec, float* action,
                      void* out, int out_dtype, int rows, int width, int A, int H, int horizon, float dt, float eps,
                      int last_step, int do_clip, float clip, cudaStream_t stream, const long long* xfix = nullptr);
// (xfix: optional [rows, width] 2^-30 fixed-point increments to add to x, see ExpertPass::xfix)

// ---- flow-matching training forward (pizero_internvl.py:1064-1197, forward only): the pieces around the one expert pass
// Per sample b with time t[b] (io dtype): action[b] = psi_t = (1 - (1 - sig_min) t) x0 + t x1 (:1053-1062, fp32),
// tbias[b, n] = b2[n] + sum_j W2a[n, j] * cat(sin(t w), cos(t w))[j]  -- the time half of ActionEncoder.linear_2
// (modules.py:9-22, 39-53) for an arbitrary per-sample time --, vel[b] = 0. x0 / x1: io dtype [B, HA].
int flow_prepare_launch(const void* x0, const void* x1, const void* t, int io, float sig_min, float time_max_period,
                        const float* W2a, const float* b2, float* action, float* tbias, float* vel, int B, int HA, int he,
                        cudaStream_t stream);
// h[r, :] = bf16(silu(z[r, :] + tbias[r / H, :]))  (the rest of linear_2 + swish with a per-sample bias)
int bias_silu_rows_launch(const float* z, const float* tbias, __nv_bfloat16* h, int rows, int H, int width,
                          cudaStream_t stream);
// loss = mean((vel - (x1 - (1 - sig_min) x0))^2) over n elements (:1190-1193), one CTA, fixed summation order
int flow_loss_launch(const float* vel, const void* x0, const void* x1, int io, float sig_min, long long n, float* loss,
                     cudaStream_t stream);

// ---- greedy text generation of the chat VLM (InternVLChatModel.generate -> HF generate(do_sample=False),
// Vlaser_VLM/internvl_chat/internvl/model/internvl_chat/modeling_internvl_chat.py:401-439; the loop of
// pizero_internvl.py:1412-1447 around PiZero.infer_text :1005-1046)
struct TextGenBuffers {
  int* next_tok;       // [B] argmax of the latest logits
  int* n_gen;          // [1] tokens emitted so far
  int* finished;       // [B] the sample has emitted eos (or filled its window)
  long long* tokens;   // [B, gen_cap]
  int *pos, *kvrow, *kv_len, *q_start, *q_len;  // [B] plan of the coming decode step
};
// out[b, :] = x[cu[b] + valid_len[b] - 1, :]: the residual row of the last prompt token of every sample
int last_rows_launch(const float* x, const int* cu, const int* valid_len, float* out, int B, int width, cudaStream_t stream);
// next_tok[r] = argmax(logits[r, :n]) (first maximum); bump: *n_gen += 1 afterwards
int argmax_rows_launch(const float* logits, long long ld, int n, int rows, int* next_tok, int* n_gen, int bump,
                       cudaStream_t stream);
// Emits token k = *n_gen of every sample (next_tok, or `fill` once the sample is finished; eos / a full window finish it),
// writes the plan of the decode step that consumes it (position valid_len + k, its K/V row, kv_len) and gathers its
// embedding row into x[b, :] (fp32).
int decode_advance_launch(const TextGenBuffers& g, const int* valid_len, const int* kvoff, const __nv_bfloat16* embed, float* x,
                          int B, int width, int window, int gen_cap, long long eos, long long fill, cudaStream_t stream);

}  // namespace vlaser
