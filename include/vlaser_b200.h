/* vlaser_b200 — C ABI of the B200-native Vlaser-VLA `infer_action` hot path.
 *
 * Drop-in boundary: the reference exposes this path as a Python class,
 * `PiZeroInference.forward -> PiZero.infer_action`
 * (Vlaser_VLA/Simpler/src/model/vla/pizero_internvl.py:1286-1307, 798-936), constructed in
 * `EvalAgent.__init__` (Vlaser_VLA/Simpler/src/agent/eval.py:33-56) and fed a checkpoint by
 * `load_state_dict` (eval.py:196-212). The entry points below are exactly what a binding for that
 * class needs: create from the cfg keys `PiZero.__init__` reads (:156-337), set weights by
 * state-dict key, finalize (pack + build schedules), infer, destroy. `vlaser_b200/policy.py` is the
 * ctypes binding that mirrors the reference class; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions: plain pointers and sizes only; 0 = OK, negative = error (text via
 * vlaser_last_error()); one handle per GPU (the current CUDA device at vlaser_create); calls on one
 * handle are serialised by the caller; device entry points are asynchronous w.r.t. the host
 * (stream-ordered on `stream`, a cudaStream_t passed as void*), borrow their pointers for the call
 * only, and do not synchronise -- with one exception: the FIRST vlaser_infer_action call for a new
 * (batch size, io dtype, valid_len present) combination waits for the staged inputs, runs one
 * un-captured pass and captures the CUDA graph that later calls replay (csrc/engine.cu,
 * engine_execute), so it blocks the host for about one call; changing a schedule option drops the
 * captured graphs. Results are bit-identical from call to call for the same inputs (split-K partial
 * sums are combined in a fixed order, option "det_splitk"). All GEMMs run in bf16 with fp32
 * accumulation whatever dtype the I/O tensors have. There is no CPU fallback: without a CUDA device
 * every compute entry point fails with an error.
 */
#ifndef VLASER_B200_H
#define VLASER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLASER_ABI_VERSION 1

/* dtype codes for floating-point I/O tensors and weights */
#define VLASER_F32 0
#define VLASER_BF16 1

typedef struct vlaser_handle vlaser_t;

/* Mirrors the hydra cfg keys read by PiZero.__init__ (pizero_internvl.py:156-337) plus the
 * InternVL3 / InternViT / Qwen2.5 dims the reference takes from the HF checkpoint config
 * (get_internvl3, :32-152). */
typedef struct vlaser_cfg {
  /* InternViT (S/ivl/modeling_intern_vit.py) */
  int32_t image_size, patch_size, vit_hidden, vit_layers, vit_heads, vit_mlp;
  float vit_eps;  /* layer_norm_eps, 1e-6 */
  float proj_eps; /* mlp1[0] nn.LayerNorm default eps, 1e-5 */
  /* LLM ("vlm" mixture) */
  int32_t hidden, layers, n_heads, n_kv_heads, head_dim, mlp;
  float rms_eps, rope_theta;
  /* action expert ("proprio"/"action" mixtures) */
  int32_t hidden_e, mlp_e;
  /* tokens */
  int32_t embed_rows, image_token_index, pad_token_id, max_image_text_tokens;
  /* action parameterisation */
  int32_t cond_steps, horizon_steps, action_dim, proprio_dim, num_inference_steps;
  int32_t clip_enabled;          /* final_action_clip_value is not None */
  float final_action_clip_value;
  float time_max_period;
  /* capacity */
  int32_t max_batch;             /* largest B any call will use on this handle */
} vlaser_cfg;

int32_t vlaser_abi_version(void);
const char* vlaser_last_error(void);

/* PiZeroInference(cfg)  — pizero_internvl.py:156-337 */
int32_t vlaser_create(const vlaser_cfg* cfg, vlaser_t** out);
/* load_state_dict, one tensor at a time — eval.py:196-212. `key` is the reference state-dict key
 * (aliases `_orig_mod.`, `joint_model.mixtures.proprio.`, `internvl_model.*` accepted); `data` may
 * be a host or device pointer; dtype VLASER_F32 or VLASER_BF16. Unknown keys return 1 (ignored). */
int32_t vlaser_set_weight(vlaser_t* h, const char* key, const void* data, const int64_t* shape, int32_t ndim,
                          int32_t dtype);
/* .to(dtype).to(device).eval() + schedule construction — eval.py:43-56. Fails if a weight is missing. */
int32_t vlaser_finalize(vlaser_t* h);
/* forward == infer_action — pizero_internvl.py:1287-1307, 798-936. All pointers are DEVICE pointers.
 *   input_ids    i64 [B, max_image_text_tokens]
 *   pixel_values io_dtype [B*cond_steps, 3, image_size, image_size], already (x/255-mean)/std
 *   valid_len    i32 [B] or NULL: number of valid prefix tokens per sample == number of zeros in
 *                image_text_proprio_mask[b,0,0,:S]; NULL derives it as count(input_ids != pad_token_id)
 *   proprios     io_dtype [B, 1, proprio_dim]
 *   noise        io_dtype [B, H, action_dim], H = horizon_steps+cond_steps-1: the draw of
 *                torch.randn at :879-881, made by the caller so results are reproducible
 *   actions_out  io_dtype [B, horizon_steps, action_dim], normalised actions */
int32_t vlaser_infer_action(vlaser_t* h, const int64_t* input_ids, const void* pixel_values, const int32_t* valid_len,
                            const void* proprios, const void* noise, void* actions_out, int32_t B, int32_t io_dtype,
                            void* stream);
/* Same call with HOST buffers (what an evaluator holds before eval.py:130 / after :139): stages the
 * inputs through pinned memory, runs on the handle's own stream, copies the actions back and
 * synchronises. */
int32_t vlaser_infer_action_host(vlaser_t* h, const int64_t* input_ids, const void* pixel_values,
                                 const int32_t* valid_len, const void* proprios, const void* noise, void* actions_out,
                                 int32_t B, int32_t io_dtype);
/* Flow-matching loss of the training step, forward only — PiZero.forward (pizero_internvl.py:1064-1197) with
 * is_vlm=False: psi_t = (1-(1-sig_min)t) x0 + t x1 (:1053-1062), ONE pass of the action expert at the per-sample time
 * t over the prefix / proprio K/V (no block of the training mask attends to a later block, so this equals the reference's
 * joint forward without KV cache), v_psi = action_decoder(.), loss = mean((v_psi - (x1 - (1-sig_min) x0))^2).
 * Device pointers:
 *   input_ids / pixel_values / valid_len / proprios   as vlaser_infer_action
 *   actions  io_dtype [B, H, action_dim]  x1, the normalised target chunk
 *   noise    io_dtype [B, H, action_dim]  x0, the draw of torch.randn_like at :1094, made by the caller
 *   t        io_dtype [B]                 flow time per sample
 *   loss_out f32 [1];  velocity_out f32 [B, H, action_dim] or NULL (v_psi, for parity checks)
 * No gradients: the optimiser step stays with the reference trainer (out of scope). */
int32_t vlaser_flow_matching_loss(vlaser_t* h, const int64_t* input_ids, const void* pixel_values, const int32_t* valid_len,
                                  const void* proprios, const void* actions, const void* noise, const void* t,
                                  float flow_sig_min, float* loss_out, float* velocity_out, int32_t B, int32_t io_dtype,
                                  void* stream);
/* Greedy text generation of the chat VLM — InternVLChatModel.generate
 * (Vlaser_VLM/internvl_chat/internvl/model/internvl_chat/modeling_internvl_chat.py:401-439: extract_feature, image rows
 * scattered into the token embeddings, then HF `language_model.generate(inputs_embeds, do_sample=False, use_cache=True)`),
 * which is also the loop of pizero_internvl.py:1412-1447 around PiZero.infer_text (:1005-1046). The LLM mixture alone:
 * causal attention, RoPE positions 0.. (HF), all layers complete, final RMSNorm, output head, argmax; the prompt K/V stay
 * in the handle's cache and every decode step (one CUDA graph, replayed) appends one row per sample.
 *   input_ids      i64 [B, max_image_text_tokens], right padded with pad_token_id; image rows = image_token_index
 *   pixel_values   io_dtype [B*cond_steps, 3, image_size, image_size] image tiles in batch order (a sample owns as many as
 *                  its ids hold runs of image tokens), or NULL for a text-only prompt
 *   valid_len      i32 [B] or NULL (count of non-pad ids)
 *   max_new_tokens 1 .. max_image_text_tokens + 1 + H: a sample's K/V window holds prompt + generated tokens; a sample
 *                  that fills it stops like one that emitted eos
 *   tokens_out     i64 [B, max_new_tokens]; positions after a sample's eos hold fill_token_id
 *   first_logits_out f32 [B, embed_rows] or NULL: logits of the last prompt position (== infer_text(...)["logits"][:, -1])
 * Needs the LLM's final norm (`joint_model.mixtures.vlm.norm.weight` / `...language_model.model.norm.weight`) and takes
 * `lm_head.weight` when the checkpoint has one (else the tied embedding table, Qwen2.5-1.5B). Device pointers;
 * stream-ordered like vlaser_infer_action (the first call per batch size captures two graphs). */
int32_t vlaser_generate(vlaser_t* h, const int64_t* input_ids, const void* pixel_values, const int32_t* valid_len,
                        int32_t max_new_tokens, int64_t eos_token_id, int64_t fill_token_id, int64_t* tokens_out,
                        float* first_logits_out, int32_t B, int32_t io_dtype, void* stream);
void vlaser_destroy(vlaser_t* h);

/* Options: "pdl" (process-wide; 0 off, 1 programmatic dependent launch inside the denoise steps only, 2 (default) on every launch),
 * "pair_gemm" (default 1: large-M GEMMs on the cta_group::2 kernel), "fuse_norm" (default 0; 1 fuses the expert streams' RMSNorm across the GEMM chain), "exp_stages" (smem ring depth
 * of the expert GEMMs, 0 = default),
 * "expert_mega" (default 0; 1 runs each proprio / denoise pass of the action expert as ONE persistent kernel, csrc/expert.cu),
 * "expert_streams" (default 2: independent sub-grids of that kernel, each owning a slice of the batch),
 * "prop_stream" (default 1: the proprio expert pass of the prefill is launched on a second stream, forked after the
 * proprio encoder and joined before the first denoise step; it only reads the prefix K/V of each layer),
 * "small_m_split" (default 1: split-K over the idle SMs for the ViT / LLM residual projections of small batches),
 * "det_splitk" (default 1: split-K partial sums go to a workspace and are added to the residual stream in split order by the
 * next norm launch, so that the same inputs give bit-identical actions; 0 = fp32 atomics / L2 reduce-adds, whose summation
 * order varies from call to call),
 * "l2_prefetch" (default 0, measured no gain: the RMSNorm launches of a denoise step issue bulk L2 prefetches of the expert weights the
 * next launches of the chain will read),
 * "expert_gemv" (default 1: at B = 1 the expert passes run on the weight-streaming single-observation kernel, csrc/expert_gemv.cu),
 * "rn_small" (process-wide, default 1: the residual update + RMSNorm of narrow split-K streams -- the action expert -- with one
 * thread per column instead of four split groups of threads; bit-identical residual stream),
 * "attn_small" (default 1: the denoise-step attention, group * H <= 32 query rows per (sample, kv head), on the kernel that splits
 * the key blocks between its two warp pairs),
 * "prefill_tc_long" (default 1: prefixes of more than 256 tokens and the causal text prefill on the two-pass tcgen05 attention
 * kernel instead of the generic mma.sync one),
 * "decode_chunk" (default 128: keys per CTA of the split-KV decode-step attention of vlaser_generate; 0 = the generic kernel),
 * "fuse_tail" (default 0, measured no gain: the o / down projections of the action expert finish the residual update + the
 * following RMSNorm themselves behind an in-kernel grid barrier instead of leaving them to a norm launch),
 * "use_graph" (default 1: replay the whole call as one CUDA graph per batch size),
 * "stop_after" (debug: 0 all, 1 ViT, 2 prefix embedding, 3 prefill, 4 first denoise step),
 * "simt_gemm" (debug/tests: route GEMMs through the CUDA-core cross-check kernel),
 * "profile_gemm" / "profile_mtot" (see vlaser_profile_read). */
int32_t vlaser_set_option(vlaser_t* h, const char* key, int64_t value);
/* Introspection for parity tests: copies the named internal device buffer to host memory.
 * Returns the buffer size in bytes (>= 0) or a negative error; copies min(size, max_bytes). */
int64_t vlaser_debug_copy(vlaser_t* h, const char* name, void* host_dst, int64_t max_bytes);
/* Number of kernels launched by the last vlaser_infer_action* call on this handle. */
int64_t vlaser_last_launch_count(vlaser_t* h);
/* In-situ GEMM timing for the roofline figure: with options "use_graph"=0, "profile_gemm"=1 and
 * "profile_mtot"=<sum of valid lengths> every tcgen05 GEMM launch of the following calls is bracketed by
 * CUDA events on its launch stream; this returns (and resets) the summed device time, the summed
 * algorithmic FLOPs (2*M*N*K with the live M) and the number of GEMM launches. */
int32_t vlaser_profile_read(vlaser_t* h, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches);
/* Same counters split by regime: out12 = {ms, flops, bytes, launches} of the GEMMs at M <= 512 (action-expert / action-
 * head projections, weight-streaming bound), then of the CTA-pair kernel at M > 512 (ViT + prefill, tensor bound), then of the persistent
 * expert-stream kernel (one launch per pass; flops = its projections, bytes = the expert weights streamed once).
 * bytes of the GEMM classes = both operands once + the output once (fp32 residual epilogues: read + write). */
int32_t vlaser_profile_read_classes(vlaser_t* h, double* out12);
/* Algorithmic FLOPs of one call at batch B with the given valid lengths (SURVEY.md §8d formulae). */
double vlaser_flops_per_call(vlaser_t* h, const int32_t* valid_len_host, int32_t B);

/* ---- single-kernel entry points (used by the parity tests and micro-benchmarks) ---------- */
typedef struct vlaser_gemm_args {
  const void* A;        /* bf16 [M,K] */
  const void* W;        /* bf16 [N,K] (nn.Linear layout) */
  int32_t M, N, K;
  const int32_t* M_dev; /* optional live row count on device */
  int32_t epi;          /* see vlaser::Epi in csrc/gemm.cuh */
  int32_t block_n;      /* 128 | 256 */
  int32_t max_splits;
  int32_t simt;         /* 1 = CUDA-core cross-check kernel */
  void* out;
  int32_t ldo;
  const float* bias;
  const float* scale;
  const int32_t* row_map;
  const int32_t* row_pos;
  const int32_t* row_kvrow;
  const float* cos_tab;
  const float* sin_tab;
  void* k_cache;
  void* v_cache;
  int64_t kv_head_stride;
  int32_t n_q_heads, n_kv_heads;
  const float* pos_emb;
  int32_t patches;
  /* fused RMSNorm chain (see csrc/gemm.cuh): producer outputs and consumer row scale */
  const float* gamma_next;
  void* a_next;          /* bf16 [M, lda_next] */
  int32_t lda_next;
  float* ss_next;        /* f32 [M], caller-zeroed */
  int32_t* tile_counters;/* i32, zero-initialised (split-K only) */
  const float* row_ss;   /* f32 [M] */
  float inv_width, eps;
  int32_t stages;        /* smem ring depth, 0 = default */
  int32_t pair;          /* 1 = CTA-pair kernel (cta_group::2, 256x256 tile per pair); block_n is ignored */
  int32_t tma_out;       /* pair kernel: epilogue through TMA store / reduce-add (bf16 and residual epilogues) */
  /* deterministic split-K (residual epilogue, max_splits > 1): split s leaves its partial tile at
   * partial[s * partial_stride + m * ldo + n] instead of meeting the other splits through fp32 atomics in `out`;
   * *partial_splits receives the split count used (0 = one split, `out` updated directly). The caller finishes with
   * vlaser_op_splitk_reduce. partial_stride: a multiple of 256 * ldo covering M rows. */
  float* partial;
  int64_t partial_stride;
  int32_t* partial_splits;
  /* > 0 (pair kernel): ss_next / row_ss are [M, ss_slots] partial sums of squares, one slot per (256-column tile, column
   * half) of the producer, plain stores, added in slot order by the consumer (csrc/gemm.cuh EpiParams::ss_slots) */
  int32_t ss_slots;
} vlaser_gemm_args;
int32_t vlaser_op_gemm(const vlaser_gemm_args* a, void* stream);

typedef struct vlaser_attn_args {
  const void* q; int64_t q_row_stride;
  const void* q2; int64_t q2_row_stride;
  const void* k; const void* v; int64_t kv_row_stride; int64_t kv_head_stride;
  void* out; int64_t out_row_stride;
  void* out2; int64_t out2_row_stride;
  const int32_t* q_start; const int32_t* q_len; const int32_t* kv_start; const int32_t* kv_len;
  int32_t uni_q_len, uni_kv_len, max_q_len, group, n_kv_heads, batch, head_dim;
  float scale;
  int32_t causal;  /* 1: query token t of a sample sees keys [0, kv_len - q_len + t] (text prefill of the chat VLM) */
  int32_t small_rows; /* 1: with group * max_q_len <= 32 (head_dim 128) use the kernel that splits the key blocks between its warp pairs */
} vlaser_attn_args;
int32_t vlaser_op_attention(const vlaser_attn_args* a, void* stream);
/* Decode-step attention of the chat VLM (one query row per sample over its K/V window, head_dim 128, GQA; HF Qwen2
 * eager_attention_forward with q_len = 1): split-KV -- one CTA per (chunk of keys, kv head, sample), then a merge launch.
 * q / out bf16 [rows, n_heads*128] (sample b at row q_start[b]); caches bf16 [n_kv_heads, kv_rows, 128] (sample b from row
 * kv_start[b], kv_len[b] keys); chunk 64..256 (multiple of 64); ws_o f32 [batch, n_kv_heads, ceil(max_kv_len/chunk),
 * group, 128], ws_ml f32 [.., group, 2]. */
int32_t vlaser_op_decode_attention(const void* q, const void* k_cache, const void* v_cache, void* out, const int32_t* kv_start,
                                   const int32_t* kv_len, int32_t batch, int32_t n_heads, int32_t n_kv_heads, int64_t kv_rows,
                                   int32_t max_kv_len, int32_t chunk, float scale, float* ws_o, float* ws_ml, const int32_t* q_start,
                                   void* stream);
/* InternViT self-attention fast path (_naive_attn, modeling_intern_vit.py:210-227): qkv bf16 [images*tokens, 3*heads*64]
 * as written by the fused qkv projection, out bf16 [images*tokens, heads*64]; 128 <= tokens <= 320. */
int32_t vlaser_op_vit_attention(const void* qkv, void* out, int32_t images, int32_t tokens, int32_t heads, float scale,
                                void* stream);

/* The same op on the tcgen05 tensor cores (csrc/attention_tc.cu): tokens == 257 (224 px images, single-pass kernel) or
 * 1 + a multiple of 256 up to 1025 (448 px tiles: exact two-pass softmax, K resident in shared memory). */
int32_t vlaser_op_vit_attention_tc(const void* qkv, void* out, int32_t images, int32_t tokens, int32_t heads, float scale,
                                   void* stream);

/* Prefix-prefill attention on the tcgen05 tensor cores (csrc/attention_tc.cu; joint_model.py:631-656 restricted to the valid
 * prefix): q bf16 [q_rows, n_heads*128] (packed rows, sample b at q_start[b] .. + q_len[b]), k / v cache bf16
 * [n_kv_heads, kv_rows, 128] (sample b at kv_start[b]), out like q. Bidirectional over the q_len[b] <= 256 tokens of a
 * sample; longer samples are left untouched (the generic kernel takes them). */
int32_t vlaser_op_prefill_attention_tc(const void* q, const void* k_cache, const void* v_cache, void* out, const int32_t* q_start,
                                       const int32_t* q_len, const int32_t* kv_start, int32_t q_rows, int32_t n_heads,
                                       int32_t n_kv_heads, int64_t kv_rows, int32_t batch, float scale, void* stream);

/* The same attention for samples of min_len <= q_len[b] tokens of any length, bidirectional (causal = 0) or under the HF
 * causal mask (causal = 1: token t sees keys [0, t]): K / V stream in blocks of 128 keys, exact two-pass softmax, O
 * accumulated in TMEM (csrc/attention_tc.cu, prefill_attention_tc_long_kernel). max_tokens: upper bound of q_len. */
int32_t vlaser_op_prefill_attention_tc_long(const void* q, const void* k_cache, const void* v_cache, void* out,
                                            const int32_t* q_start, const int32_t* q_len, const int32_t* kv_start, int32_t q_rows,
                                            int32_t n_heads, int32_t n_kv_heads, int64_t kv_rows, int32_t batch, int32_t max_tokens,
                                            int32_t min_len, int32_t causal, float scale, void* stream);

int32_t vlaser_op_layernorm(const float* x, const float* gamma, const float* beta, void* out_bf16, int32_t rows,
                            int32_t width, float eps, int32_t shuffle_grid, int32_t shuffle_channels, void* stream);
int32_t vlaser_op_rmsnorm(const float* x, const float* gamma, void* out_bf16, int32_t rows, int32_t width, float eps,
                          void* stream);
/* x[m, :] += partial[0][m, :] + partial[1][m, :] + ... in split order (one writer per row: bit-identical from call to
 * call), then, when gamma and out_bf16 are given, out = RMSNorm(x) * gamma as vlaser_op_rmsnorm. */
int32_t vlaser_op_splitk_reduce(float* x, int32_t rows, int32_t width, const float* part, int64_t part_stride,
                                const int32_t* splits_dev, const float* gamma, void* out_bf16, float eps, void* stream);
int32_t vlaser_op_pixel_shuffle(const void* in, void* out, int32_t n, int32_t grid, int32_t channels,
                                int32_t elem_bytes, void* stream);
int32_t vlaser_op_im2col(const void* pixel, int32_t pixel_dtype, void* cols_bf16, int32_t n, int32_t px, int32_t ps,
                         int32_t kpad, void* stream);
int32_t vlaser_op_euler_step(const float* x, const float* gamma, const float* Wdec, const float* bdec, float* action,
                             void* out, int32_t out_dtype, int32_t rows, int32_t width, int32_t A, int32_t H,
                             int32_t horizon, float dt, float eps, int32_t last_step, int32_t do_clip, float clip,
                             void* stream);

/* ---- either side of infer_action -------------------------------------------------
 * Input staging — replaces, for a whole batch on the device, `cv2.resize(image, image_size, INTER_LANCZOS4)` +
 * the optional centre crop (Vlaser_VLA/Simpler/src/agent/env_adapter/simpler.py:70-84) and `process_images`
 * (src/model/vla/processing.py:27-63, mean/std :303-304).
 *   frames        u8 device pointer, [n, src_h, src_w, 3] (layout 0, as the simulator returns them) or
 *                 [n, 3, src_h, src_w] (layout 1)
 *   center_crop   1 = the CENTER_CROP branch (crop to min(h, w) around the centre before the resize)
 *   pixel_values  out_dtype [n, 3, px, px]: exactly what vlaser_infer_action takes
 *   mean3/std3    HOST pointers to 3 floats, or NULL for the ImageNet constants of processing.py
 * The 8-bit Lanczos4 resize is bit-exact with OpenCV's; the normalisation is bit-exact with the reference's float32
 * arithmetic followed by the cast to out_dtype. */
int32_t vlaser_op_stage_frames(const uint8_t* frames, int32_t n, int32_t src_h, int32_t src_w, int32_t layout,
                               int32_t center_crop, void* pixel_values, int32_t out_dtype, int32_t px, const float* mean3,
                               const float* std3, void* stream);
/* Action post-processing — `SimplerAdapter.postprocess` for n_env environments at once (simpler.py:136-177;
 * de-normalisation base.py:21-49; euler2axangle src/utils/geometry.py:261-434; gripper rules simpler.py:217-221
 * (gripper 0, Bridge) and :259-286 (gripper 1, EDR sticky gripper)). All pointers are DEVICE pointers.
 *   actions       dtype [n_env, T, 7] normalised action chunks (the output of vlaser_infer_action)
 *   norm          0 = "bound": stat_a = p01, stat_b = p99; 1 = "gaussian": stat_a = mean, stat_b = std (f64 [7])
 *   sticky_state  f64 [n_env, 3] = {sticky_action_is_on, gripper_action_repeat, sticky_gripper_action}, zero after
 *                 reset(); read and updated (gripper 1 only; may be NULL otherwise)
 *   out           f64 [n_env, T, 7] = (xyz, axis * angle, gripper), float64 like the reference's numpy */
int32_t vlaser_op_postprocess_actions(const void* actions, int32_t dtype, int32_t n_env, int32_t T, int32_t A, int32_t norm,
                                      const double* stat_a, const double* stat_b, int32_t gripper, int32_t sticky_repeat,
                                      double* sticky_state, double* out, void* stream);

/* Proprio normalisation / joint de-normalisation — `BaseEnvAdapter.{normalize,denormalize}_{bound,gaussian}`
 * (Vlaser_VLA/Simpler/src/agent/env_adapter/base.py:8-49) as `SimplerAdapter.preprocess` applies them to the proprio vector
 * (simpler.py:107-121), elementwise over a DEVICE array [rows, dim], float64 arithmetic like numpy. dtype codes: 0 f32,
 * 1 bf16, 2 f64. norm 0 = "bound" (stat_a = p01, stat_b = p99; the forward direction clips to [-1, 1]), 1 = "gaussian"
 * (stat_a = mean, stat_b = std); inverse = 1 de-normalises (RoboTwin: the 14 joint targets of an action chunk). */
int32_t vlaser_op_normalize(const void* in, int32_t in_dtype, void* out, int32_t out_dtype, int64_t rows, int32_t dim,
                            int32_t norm, int32_t inverse, const double* stat_a, const double* stat_b, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VLASER_B200_H */
