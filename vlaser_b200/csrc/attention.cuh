// Fused (flash-style, online-softmax) attention used at the three sites of the hot path:
//   * InternViT self-attention    (head_dim 64, no mask, N = P+1 tokens per image)       -- _naive_attn
//   * prefix prefill              (head_dim 128, GQA, bidirectional over the valid prefix; the
//                                  proprio token is an extra query row that also sees its own key)
//   * denoise step                (head_dim 128, GQA, H action queries over cached prefix + proprio
//                                  + H fresh action keys)
// No mask tensor is read: the block-attention rule of build_causal_mask_and_position_ids
// (pizero_internvl.py:517-587) is encoded as per-sample (q_len, kv_len) plus the extra-row rule.
#pragma once
#include "common.cuh"

namespace vlaser {

struct AttnParams {
  const __nv_bfloat16* q = nullptr;   // row r, head h at q + r*q_row_stride + h*HD
  long long q_row_stride = 0;
  const __nv_bfloat16* q2 = nullptr;  // optional extra query row per sample (proprio): q2 + b*q2_row_stride + h*HD
  long long q2_row_stride = 0;
  const __nv_bfloat16* k = nullptr;   // kv row r, kv head g at k + g*kv_head_stride + r*kv_row_stride
  const __nv_bfloat16* v = nullptr;
  long long kv_row_stride = 0;
  long long kv_head_stride = 0;
  __nv_bfloat16* out = nullptr;       // same addressing as q
  long long out_row_stride = 0;
  __nv_bfloat16* out2 = nullptr;      // same addressing as q2
  long long out2_row_stride = 0;
  // per-sample segments; when q_len == nullptr the uniform values are used (start = b * uni_*)
  const int* q_start = nullptr;
  const int* q_len = nullptr;
  const int* kv_start = nullptr;
  const int* kv_len = nullptr;
  int uni_q_len = 0;
  int uni_kv_len = 0;
  int max_q_len = 0;    // upper bound on q_len (+1 if q2) for grid sizing
  int group = 1;        // query heads per kv head
  int n_kv_heads = 1;
  int batch = 0;
  float scale = 1.f;
  // 1: the K/V rows [0, kv_len - q_len) of every sample (cached prefix + proprio row of a denoise step) were written
  // by kernels that completed long before the previous launch, so the first K/V block may be fetched BEFORE the
  // programmatic-dependent-launch wait and overlaps the tail of the q|k|v projection that produces the new rows
  int stable_prefix = 0;
  int min_kv_len = 0;   // samples with fewer kv rows are skipped (they were handled by the tcgen05 prefill kernel)
  int causal = 0;       // 1: query token t of a sample sees keys [0, kv_len - q_len + t] (HF causal mask; no q2 row)
  int small_rows = 0;   // 1: when group * max_q_len <= 32 (the denoise steps) use the kernel that splits the key blocks between
                        // its two warp pairs (attention_small_kernel)
  long long* trace = nullptr;  // optional timeline record of this launch (common.cuh trace_stamp)
};

int attention_launch(const AttnParams& p, int head_dim, cudaStream_t stream);

// Decode-step attention of the chat VLM (one new query token per sample over its whole K/V window, head_dim 128, GQA):
// the generic kernel gives such a step batch * n_kv_heads CTAs that each walk the window serially (91 us per layer at 3.3 k
// keys). Here the window is cut into chunks of `chunk` keys: one CTA per (chunk, kv head, sample) leaves the un-normalised
// output and the (max, sum) of its chunk in a workspace, a second launch merges the chunks of a row (split-KV softmax).
struct DecodeAttnParams {
  const __nv_bfloat16* q = nullptr;   // query row of sample b: q + q_start[b] * q_row_stride, head h at + h * 128
  long long q_row_stride = 0;
  const __nv_bfloat16* k = nullptr;   // as AttnParams
  const __nv_bfloat16* v = nullptr;
  long long kv_row_stride = 0, kv_head_stride = 0;
  __nv_bfloat16* out = nullptr;       // addressed like q
  long long out_row_stride = 0;
  const int* q_start = nullptr;       // [batch]
  const int* kv_start = nullptr;      // [batch]
  const int* kv_len = nullptr;        // [batch]
  int group = 1, n_kv_heads = 1, batch = 0;
  int max_kv_len = 0;                 // upper bound of kv_len (grid sizing)
  int chunk = 256;                    // keys per CTA, a multiple of 64, <= 256
  float scale = 1.f;
  float* ws_o = nullptr;              // [batch, n_kv_heads, splits, group, 128] f32, splits = ceil(max_kv_len / chunk)
  float* ws_ml = nullptr;             // [batch, n_kv_heads, splits, group, 2] f32
};
int decode_attention_launch(const DecodeAttnParams& p, cudaStream_t stream);

// InternViT self-attention fast path (head_dim 64, no mask, 128 <= tokens <= 320): qkv [images*tokens, 3*heads*64]
// packed as the fused qkv projection writes it, out [images*tokens, heads*64]. One CTA per (image, head).
int vit_attention_launch(const __nv_bfloat16* qkv, __nv_bfloat16* out, int images, int tokens, int heads, float scale,
                         cudaStream_t stream);

// The same op on the tcgen05 tensor cores for tokens == 257 (224 px): attention_tc.cu. tmQKV: make_tmap_bf16 map of qkv
// [images*257, 3*heads*64] with 128-row boxes; tmOut: the same kind of map of out [images*257 (or more rows), heads*64].
int vit_attention_tc_launch(const CUtensorMap& tmQKV, const CUtensorMap& tmOut, const __nv_bfloat16* qkv,
                            __nv_bfloat16* out, int images, int tokens, int heads, float scale, cudaStream_t stream,
                            long long* trace = nullptr);

// Prefix-prefill attention (head_dim 128, GQA-packed query tiles, L <= 256 valid tokens per sample read from device memory)
// on the tcgen05 tensor cores: attention_tc.cu. tmQ: make_tmap_q3d map of the q buffer; tmK / tmV: make_tmap_bf16 maps of
// the whole K / V cache viewed as [layers * kv_heads * kv_rows, 128] with 32-row boxes; kv_layer_row0 = layer * kv_heads *
// kv_rows. Samples with L > 256 are skipped (AttnParams::min_kv_len of the generic kernel takes them).
int prefill_attention_tc_launch(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const int* q_start,
                                const int* q_len, const int* kv_start, __nv_bfloat16* out, long long out_row_stride,
                                long long kv_layer_row0, long long kv_rows, int batch, int n_kv_heads, int group, float scale,
                                cudaStream_t stream, long long* trace = nullptr);
// Long prefixes (samples with min_len <= L tokens; any length) on the tensor cores, bidirectional or causal: K / V stream
// in blocks of 128 keys, exact two-pass softmax (attention_tc.cu). Same maps as the kernel above.
int prefill_attention_tc_long_launch(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const int* q_start,
                                     const int* q_len, const int* kv_start, __nv_bfloat16* out, long long out_row_stride,
                                     long long kv_layer_row0, long long kv_rows, int batch, int n_kv_heads, int group, int max_tokens,
                                     int min_len, int causal, float scale, cudaStream_t stream);
int make_tmap_q3d(CUtensorMap* out, const void* ptr, uint64_t rows, int n_heads, int group);

}  // namespace vlaser
