// Engine behind the C ABI (include/vlaser_b200.h): weight packing, HBM-resident buffers, the
// launch schedule of PiZero.infer_action (pizero_internvl.py:798-936) and its CUDA-graph replay.
//
// Data layout in HBM (all sized for cfg.max_batch at create):
//   weights      bf16 [N,K] row-major (nn.Linear layout == K-major UMMA operand), q|k|v fused,
//                gate/up interleaved in 32-row groups; norm gains, biases, layer-scales, RoPE
//                tables, time-embedding biases in fp32
//   residuals    fp32  ViT [imgs*(P+1), d], prefix [sum valid_len, h] PACKED (no padding rows),
//                proprio [B, he], action [B*H, he]
//   GEMM inputs  bf16 activations written by the norm kernels / epilogues
//   KV cache     bf16 [layer][K|V][kv_head][rows][128], sample b owns rows kvoff[b].. :
//                valid prefix tokens, then the proprio token, then the H action tokens
// The valid lengths live on the device (plan kernels), every launch configuration depends on B
// only, so one CUDA graph per batch size replays the whole call.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/vlaser_b200.h"
#include "attention.cuh"
#include "elementwise.cuh"
#include "expert.cuh"
#include "gemm.cuh"

namespace vlaser {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
// PDL policy: 0 = off, 1 = only inside the latency-bound expert chain of the denoise steps, 2 = every kernel.
static int g_pdl_mode = 2;  // 0 off, 1 inside the denoise steps only, 2 (default) every launch: 39.8 vs 40.6 ms per call at B=64 (profiles/r01e_schedule_ab.txt)
static thread_local bool g_pdl_scope = false;  // set by the schedule while it enqueues a latency-bound section (per host
                                               // thread: handles driven from different threads do not see each other's)
bool pdl_enabled() { return g_pdl_mode == 2 || (g_pdl_mode == 1 && g_pdl_scope); }
void set_pdl_enabled(bool on) { g_pdl_scope = on; }

namespace {

using bf16 = __nv_bfloat16;

// ------------------------------------------------------------------------------------------
// weight ingest
// ------------------------------------------------------------------------------------------
enum PackMode : int { PK_PLAIN = 0, PK_GATE = 1, PK_UP = 2 };

struct Slot {
  void* dst = nullptr;
  int dst_dtype = DT_F32;
  long long rows = 1, cols = 1;       // expected source view [rows, cols]
  long long src_col0 = 0, ncols = 0;  // copied column range (ncols==0 -> all)
  long long dst_ld = 0, dst_row_off = 0;
  int mode = PK_PLAIN;
  bool filled = false;
  bool optional = false;  // finalize does not insist on it (text head of the chat VLM)
};

template <typename TS, typename TD>
__global__ void pack_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long rows, long long src_ld,
                            long long src_col0, long long ncols, long long dst_ld, long long dst_row_off, int mode) {
  const long long total = rows * ncols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ncols, c = i - r * ncols;
    long long dr = r;
    if (mode == PK_GATE) dr = (r >> 5) * 64 + (r & 31);
    if (mode == PK_UP) dr = (r >> 5) * 64 + 32 + (r & 31);
    dst[(dst_row_off + dr) * dst_ld + c] = static_cast<TD>(static_cast<float>(src[r * src_ld + src_col0 + c]));
  }
}

struct Lin {            // one packed projection weight
  bf16* w = nullptr;
  int N = 0, K = 0;
  CUtensorMap tm128, tm256, tm64;
  bool has256 = false;
};
struct Act {            // one bf16 activation buffer usable as GEMM A operand (tm) and as TMA-epilogue output (otm)
  bf16* p = nullptr;
  int rows = 0, K = 0;
  CUtensorMap tm, otm;
};

struct VitLayer {
  float *n1w, *n1b, *n2w, *n2b, *bqkv, *bproj, *ls1, *bfc1, *bfc2, *ls2;
  Lin qkv, proj, fc1, fc2;
};
struct LlmLayer {
  float *ln_in, *ln_post, *bqkv;
  Lin qkv, o, gu, down;
};

struct Engine {
  vlaser_cfg c{};
  int device = 0;
  int P = 0, T = 0, G = 0, H = 0, S = 0, Bc = 0, kpad = 0, n_img = 0;
  int mlp_p = 0, mlp_e_p = 0;  // MLP widths padded to a multiple of 128 (8B expert: 5600 -> 5632); pad rows / columns are zero
  int qdim = 0, kvdim = 0, qkvN = 0;
  long long kv_rows = 0;
  bool finalized = false;
  std::vector<void*> allocs;
  std::multimap<std::string, Slot> slots;

  // weights
  Lin patch;
  float *patch_b = nullptr, *cls = nullptr, *pos = nullptr;
  std::vector<VitLayer> vit;
  float *pj_lnw = nullptr, *pj_lnb = nullptr, *pj_b1 = nullptr, *pj_b2 = nullptr;
  Lin pj1, pj2;
  bf16* embed = nullptr;
  std::vector<LlmLayer> llm, exp;
  float* final_norm = nullptr;
  float *ae_w1 = nullptr, *ae_b1 = nullptr, *ae_w2a = nullptr, *ae_b2 = nullptr, *ae_b3 = nullptr;
  Lin ae_w2b, ae_w3;
  float *pe_w = nullptr, *pe_b = nullptr, *dec_w = nullptr, *dec_b = nullptr;
  float *cos_tab = nullptr, *sin_tab = nullptr, *time_bias = nullptr;  // [steps][he]

  // activations
  Act cols, vn, vattn, vmlp, ps, pa, xn, ao, act, xpn, aop, actp, xan, aoa, acta, e1, h2;
  float *vx = nullptr, *xv = nullptr, *xp = nullptr, *xa = nullptr, *action = nullptr;
  CUtensorMap vx_otm, xv_otm, vqkv_otm, xp_otm, xa_otm;
  CUtensorMap q3d_tm, kc_tm, vc_tm;  // tcgen05 prefill attention: 3-D map of the LLM q buffer, 32-row maps of the K / V caches
  CUtensorMap vqkv_tm;  // load map (128-row boxes) of the ViT q|k|v buffer for the tcgen05 attention kernel  // output maps of the fp32 residual streams / ViT qkv buffer
  bf16 *vqkv = nullptr, *q = nullptr, *qp = nullptr, *qa = nullptr;
  bf16 *kc = nullptr, *vc = nullptr;  // [layers][kv_heads][kv_rows][128]
  // fused RMSNorm chain of the expert streams: per pass (prefill proprio, then each denoise step) one row-sum-of-
  // squares slot per norm site [(steps+1)][2*layers+1][ss_rows], zeroed once per call; split-K tile counters
  float* ss = nullptr;
  int ss_rows = 0;
  // LLM prefill with RMSNorm fused across the GEMM chain (option fuse_norm_llm): row sums of squares of the prefix stream
  // as partial sums per (256-column tile, column half) of the producing projection, [2 sites][rows][hidden / 128]
  float* llm_ss = nullptr;
  int llm_ss_slots = 0;
  // ViT LayerNorm launches under the residual GEMM in front of them (option ln_overlap): row-block completion counters,
  // [2 * vit_layers][ceil(rows / 256)], zeroed at the start of a call
  int* ln_ctr = nullptr;
  int ln_ctr_blocks = 0;
  size_t ss_floats = 0;
  int* tile_ctr = nullptr;
  // persistent expert-stream kernel (expert.cu): per-layer operand table and one grid-barrier counter per pass
  ExpertLayerDev* xp_layers = nullptr;
  unsigned int* xp_bar = nullptr;
  size_t xp_bar_words() const { return (size_t)EXPERT_MAX_STREAMS * 7 * c.layers; }  // barrier counters of one pass
  int xp_streams = 2;              // sub-grids of the expert-stream kernel (option "expert_streams"), when the rows allow
  long long* xfix = nullptr;       // [(steps+1)][8][he] fixed-point residual increments of the single-observation kernel
  size_t xfix_words() const { return (size_t)8 * c.hidden_e; }
  long long* xp_timing = nullptr;  // [(steps+1)][7*layers] phase-completion timestamps, filled when xp_time is set
  int xp_time = 0, xp_flags = 0;
  PlanBuffers pb{};
  // static I/O staging (graph operands)
  long long* in_ids = nullptr;
  void *in_pix = nullptr, *in_prop = nullptr, *in_noise = nullptr, *out_act = nullptr;
  int* in_valid = nullptr;
  // flow-matching training forward (vlaser_flow_matching_loss): target actions x1, per-sample times, per-sample time bias
  // of the action encoder, fp32 pre-activation of its linear_2, velocity, loss
  void *in_x1 = nullptr, *in_t = nullptr;
  float *flow_tbias = nullptr, *flow_z = nullptr, *flow_vel = nullptr, *flow_loss = nullptr;
  float flow_sig_min = 0.001f, flow_sig_graph = -1.f;
  int mode = 0;                // what engine_execute enqueues: 0 infer_action, 1 flow-matching loss (forward only),
                               // 2 causal text prefill + first token, 3 one greedy decode step
  // ---- greedy text generation of the chat VLM (vlaser_generate): final norm of the LLM and its output head (optional
  // weights; the head falls back to the tied embedding table), the residual rows of the tokens being decoded, logits
  float* llm_norm = nullptr;
  bool llm_norm_set = false;
  Lin lm_head;                 // N = embed_rows padded to 128; the tensor maps clip at embed_rows (zero rows behind)
  bool lm_head_own = false;
  float* x_dec = nullptr;      // [Bc padded to 256, hidden] fp32
  CUtensorMap xdec_otm;
  float* logits = nullptr;     // [Bc, lm_head.N] fp32
  TextGenBuffers gen{};
  int gen_cap = 0;             // = the K/V window of a sample, S + 1 + H
  long long gen_eos = -1, gen_fill = 0;
  int text_no_image = 0;       // the prompt of this vlaser_generate call holds no image tile
  float *dec_ws_o = nullptr, *dec_ws_ml = nullptr;  // split-KV workspace of the decode-step attention
  int dec_chunk = 128;         // keys per CTA of the decode-step attention (option "decode_chunk"; 0: the generic kernel)
  void* pinned = nullptr;
  size_t pinned_bytes = 0;

  // options / state
  int use_graph = 1, stop_after = 0, simt = 0;
  int exp_stages = 0;          // smem ring depth of the expert-stream GEMMs (0 = kernel default)
  int pair_min_tiles = 36;     // fewest pair tiles (x splits) for which the pair kernel is preferred. 36 puts the o / gate-up /
                               // down projections of the action expert (36 / 70 / 72 pair tiles at 256 rows) on it: one
                               // 256-row tile reads every weight once (gpurun r02e: 42.5 ms per call at 74, 41.9 at 70 / 36)
  int tma_epi = 1;             // 1: pair-kernel epilogues leave through TMA stores / reduce-adds
  int pair_gemm = 1;           // 1: large-M GEMMs run on the CTA-pair (cta_group::2) kernel
  int mega = 0;                // 1: the expert streams (proprio pass of the prefill, every denoise step) run as ONE
                               // persistent kernel per pass (expert.cu) instead of 7 kernels per layer. Measured on
                               // B200 at B=64 (profiles/r01e_schedule_ab.txt): 43.2 ms per call against 40.3 ms for the
                               // per-kernel schedule with PDL, so the per-kernel schedule is the default
  int prop_stream = 1;         // 1: the proprio expert pass of the prefill runs on a second stream beside the LLM layers
                               // (B=64, gpurun r01ab: 38.87 / 38.74 ms with, 39.31 / 39.36 ms without)
  int small_m_split = 1;       // 1: split-K for the ViT / LLM residual projections when their tiles cover < half of the SMs
  int l2pf = 0;                // 1: the norm launches of a denoise step prefetch the coming expert weights into L2. Measured at
                               // B=64 (gpurun r01l, two runs each): 39.8 / 40.7 ms with, 39.5 / 39.8 ms without -> off
  int gemv = 1;                // 1: at B = 1 the expert passes run on the weight-streaming kernel (expert_gemv.cu)
  // deterministic split-K (option "det_splitk", default 1): partial-tile workspaces of the two launch streams, see
  // EpiParams::part (gemm.cuh) and SplitKPartials (elementwise.cuh)
  struct SplitWs {
    float* part = nullptr;
    size_t floats = 0;
    int* splits = nullptr;
    std::map<int, CUtensorMap> maps;  // workspace viewed as [floats / N, N] f32, by N
  };
  SplitWs ws_main, ws_side;
  int det_splitk = 1;
  // in-situ timeline (option "trace"): one 8-slot %globaltimer record per launch of the call, see common.cuh
  long long* trace = nullptr;
  int trace_on = 0;
  size_t trace_idx = 0;
  std::string trace_names;
  static constexpr size_t TRACE_MAX = 8192;
  int prefill_tc = 1;          // 1: prefix-prefill attention on the tcgen05 kernel (samples of <= 256 valid tokens; attention_tc.cu)
  int attn_small = 1;          // 1: the denoise-step attention (group * H <= 32 query rows) on the kernel that splits the key blocks
                               // between its two warp pairs (attention_small_kernel)
  int prefill_tc_long = 1;     // 1: samples of more than 256 tokens (and the causal text prefill) on the two-pass tcgen05 kernel
  int gelu_warps = 8;          // epilogue warps of the pair kernel's exact-GELU variant (8 with a 6-stage ring | 16 with 5 stages).
                               // With the packed-FP32 GELU (gelu_erf_fast2) 8 warps win: ViT 11.50 vs 11.62 ms sustained
                               // (profiles/r02s3n); with the scalar form the two were equal
  int pair_bn128 = 1;          // 1: 256 x 128 pair tiles for the one-row-tile split-K projections (see Run::gemm)
  int den_chains = 1;          // sub-batch chains of the denoise steps (1 | 2), see Run::denoise_all
  int half_sm = 0;             // 1: one-tile-per-CTA pair launches at a static M <= 256 on the half-SM variant of the pair kernel
                               // (two CTAs per SM, so the next launch of the denoise chain is resident and through its prologue
                               // while the current one runs). Measured (profiles/r02s3f): the late-entry stragglers of the o /
                               // gate-up / down launches disappear, but a 3-stage ring does not cover the HBM latency of the
                               // gate/up weights (mainloop 4.3 -> 8.4 us), 4 epilogue warps issue the partial-tile TMA stores in
                               // twice the time, and the denoise attention capped at 128 registers to make room loses 1.8 us:
                               // 38.9 vs 37.7 ms per call. Off.
  int fuse_norm_llm = 0;       // 1: LLM prefill at large M (all four projections on the CTA-pair kernel): the o / down projections
                               // write the next GEMM's operand bf16(x * gamma) and the row sums of squares themselves, the q|k|v /
                               // gate-up epilogues scale their accumulator rows by rstd -- the 2 RMSNorm launches per layer go.
                               // Measured at B = 64 (profiles/r02s3j, r02s3o: both schedules traced in one run): the 55 norm
                               // launches (21.7 us per layer) disappear, but every GEMM around them gets slower -- o + 9.4 us and
                               // down + 5.6 us (the read-modify-write epilogue of the last tile is 12 us against 3-5 for the L2-side
                               // reduce-add), gate/up + 10 us and q|k|v + 3.7 us (row statistics + row scale in epilogues that
                               // already pace their tile loops) -- 36.6 us per layer: 25.4 vs 24.7 ms for ViT + prefill sustained,
                               // 37.9 vs 37.1 ms per call. Off.
  int ln_overlap = 1;          // 1: the LayerNorm launches of the ViT run UNDER the residual GEMM (proj / fc2) in front of them: the
                               // GEMM's epilogue warps report finished 256-row blocks (EpiParams::done_ctr), the norm launch --
                               // the GEMM's programmatic dependent -- does not wait for the GEMM but for the block of each row
  int known_splits = 1;        // 1: norm launches behind a static-M split-K projection get the split count from the host
  int fuse_tail = 0;           // 1: the o / down projections of the action expert at one 256-row tile finish the residual update +
                               // the following RMSNorm themselves behind an in-kernel grid barrier (EpiParams::tail_ctr): two of
                               // the seven dependent launches of an expert layer disappear. Bit-identical actions; measured at
                               // B = 64 (profiles/r02s2k_ab_fuse_tail.txt, three alternations): 39.88 / 39.36 / 39.69 ms with,
                               // 39.56 / 39.99 / 39.57 ms without -- the in-kernel barrier + the two-rows-per-CTA tail cost what
                               // the dependent-launch hop + the row-per-CTA norm launch cost. Off.
  unsigned int* tail_ctrs = nullptr;  // one arrival counter per fused launch of a call, zeroed at the start of the call
  int tail_used = 0;
  static constexpr int TAIL_CTRS = 4096;
  int xs_o = 12, xs_down = 24; // split-K caps of the action expert's o / down projections (options "exp_splits_o/down")
  int vit_tc = 1;              // 1: InternViT attention of 257-token images on the tcgen05 kernel (attention_tc.cu)
  int ablate = 0;              // timing aid: bit mask of denoise-step launches to skip (1 norm_in, 2 qkv, 4 attention, 8 o,
                               // 16 norm_post, 32 gate/up, 64 down); the actions are then garbage
  int fuse_norm = 0;           // 1: RMSNorm of the expert streams fused across the GEMM chain (see exp_pre/exp_post)
  int profile_gemm = 0;        // bracket every GEMM launch with CUDA events (non-graph runs only)
  long long profile_mtot = 0;  // live packed-row count used for the FLOP accounting of var-len GEMMs
  std::vector<cudaEvent_t> prof_events;
  std::vector<double> prof_flops, prof_bytes;
  std::vector<int> prof_class;  // 0 = small-M GEMMs (M <= 512: expert / action-head projections, weight streaming),
                                // 1 = CTA-pair kernel at large M (ViT + prefill, tensor bound), 2 = expert-stream kernel
  size_t prof_used = 0;
  long long launches = 0, last_launches = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t side_stream = nullptr;      // the proprio pass of the prefill when option "prop_stream" is set
  std::vector<cudaEvent_t> side_events;    // [layers] K/V of LLM layer l written, + fork + join
  struct GraphKey {
    int B, io, stop, simt, has_valid, mode, no_img;
    bool operator<(const GraphKey& o) const {
      return std::tie(B, io, stop, simt, has_valid, mode, no_img) <
             std::tie(o.B, o.io, o.stop, o.simt, o.has_valid, o.mode, o.no_img);
    }
  };
  struct GraphVal {
    cudaGraphExec_t exec;
    long long launches;
  };
  std::map<GraphKey, GraphVal> graphs;

  ~Engine() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.second.exec);
    for (auto ev : prof_events) cudaEventDestroy(ev);
    for (void* p : allocs) cudaFree(p);
    if (pinned) cudaFreeHost(pinned);
    if (own_stream) cudaStreamDestroy(own_stream);
    if (side_stream) cudaStreamDestroy(side_stream);
    for (cudaEvent_t ev : side_events) cudaEventDestroy(ev);
  }

  template <typename Tp>
  int alloc(Tp** out, size_t count) {
    void* p = nullptr;
    const size_t bytes = std::max<size_t>(count * sizeof(Tp), 256);
    VL_CUDA_CHECK(cudaMalloc(&p, bytes));
    VL_CUDA_CHECK(cudaMemset(p, 0, bytes));
    allocs.push_back(p);
    *out = reinterpret_cast<Tp*>(p);
    return 0;
  }

  // ---- slot helpers
  void add_slot(const std::string& key, void* dst, int dtype, long long rows, long long cols, long long dst_ld = 0,
                long long dst_row_off = 0, int mode = PK_PLAIN, long long src_col0 = 0, long long ncols = 0) {
    Slot s;
    s.dst = dst;
    s.dst_dtype = dtype;
    s.rows = rows;
    s.cols = cols;
    s.src_col0 = src_col0;
    s.ncols = ncols ? ncols : cols;
    s.dst_ld = dst_ld ? dst_ld : s.ncols;
    s.dst_row_off = dst_row_off;
    s.mode = mode;
    slots.emplace(key, s);
  }
  int add_vec(const std::string& key, float** dst, long long n) {
    if (alloc(dst, (size_t)n)) return -1;
    add_slot(key, *dst, DT_F32, 1, n);
    return 0;
  }
  int add_lin(const std::string& key, Lin* l, int N, int K) {
    l->N = N;
    l->K = K;
    if (alloc(&l->w, (size_t)N * K)) return -1;
    if (!key.empty()) add_slot(key, l->w, DT_BF16, N, K);
    return 0;
  }
  int make_lin_maps(Lin* l) {
    if (make_tmap_bf16(&l->tm128, l->w, l->N, l->K, 128)) return -1;
    if (make_tmap_bf16(&l->tm64, l->w, l->N, l->K, 64)) return -1;
    l->has256 = (l->N % 256 == 0);
    if (l->has256 && make_tmap_bf16(&l->tm256, l->w, l->N, l->K, 256)) return -1;
    return 0;
  }
  int add_act(Act* a, int rows, int K) {
    a->rows = rows;
    a->K = K;
    // pad the row count to a multiple of 256 (one CTA-pair tile) so that TMA boxes never start out of bounds
    const int rows_alloc = (rows + 255) / 256 * 256;
    if (alloc(&a->p, (size_t)rows_alloc * K)) return -1;
    if (make_tmap_out(&a->otm, a->p, rows_alloc, K, 2)) return -1;
    return make_tmap_bf16(&a->tm, a->p, rows_alloc, K, 128);
  }
};

#define VL_TRY(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

// ------------------------------------------------------------------------------------------
// construction
// ------------------------------------------------------------------------------------------
int engine_build(Engine& e) {
  const vlaser_cfg& c = e.c;
  VL_CHECK(c.head_dim == 128, "head_dim must be 128 (joint_model.py:437)");
  VL_CHECK(c.vit_hidden == c.vit_heads * 64, "InternViT head dim must be 64");
  VL_CHECK(c.image_size % c.patch_size == 0, "image_size % patch_size");
  VL_CHECK(c.n_heads % c.n_kv_heads == 0, "n_heads % n_kv_heads");
  VL_CHECK(c.max_batch > 0, "max_batch > 0");
  VL_CHECK(c.mlp % 32 == 0 && c.mlp_e % 32 == 0, "mlp sizes must be multiples of 32");
  VL_CHECK(c.hidden % 128 == 0 && c.hidden_e % 128 == 0 && c.vit_hidden % 128 == 0 && c.vit_mlp % 128 == 0,
           "hidden sizes must be multiples of 128");
  VL_CHECK(c.action_dim <= 32 && c.proprio_dim <= 64, "action/proprio dims too large");
  e.G = c.image_size / c.patch_size;
  VL_CHECK(e.G % 2 == 0, "patch grid must be even for pixel_shuffle(0.5)");
  e.P = e.G * e.G;
  e.T = e.P / 4;
  e.n_img = c.cond_steps;
  e.H = c.horizon_steps + c.cond_steps - 1;
  e.S = c.max_image_text_tokens;
  e.Bc = c.max_batch;
  e.kpad = (3 * c.patch_size * c.patch_size + 63) / 64 * 64;
  e.mlp_p = (c.mlp + 127) / 128 * 128;
  e.mlp_e_p = (c.mlp_e + 127) / 128 * 128;
  e.qdim = c.n_heads * 128;
  e.kvdim = c.n_kv_heads * 128;
  e.qkvN = e.qdim + 2 * e.kvdim;
  e.kv_rows = (long long)e.Bc * (e.S + 1 + e.H);
  VL_CHECK(e.n_img * e.T <= e.S, "image tokens exceed max_image_text_tokens");
  VL_CUDA_CHECK(cudaGetDevice(&e.device));
  // The side stream carries latency-bound launches (proprio pass of the prefill) next to the large GEMMs of the main
  // stream: lowest priority, so that the block scheduler hands free SMs to the main stream's CTAs first (its launches
  // otherwise queue behind side-stream CTAs that hold an SM each while they wait on a dependent load).
  int prio_lo = 0, prio_hi = 0;
  VL_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  VL_CUDA_CHECK(cudaStreamCreateWithPriority(&e.own_stream, cudaStreamNonBlocking, prio_hi));
  VL_CUDA_CHECK(cudaStreamCreateWithPriority(&e.side_stream, cudaStreamNonBlocking, prio_lo));
  e.side_events.resize((size_t)c.layers + 1);
  for (cudaEvent_t& ev : e.side_events) VL_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));

  const int d = c.vit_hidden, f = c.vit_mlp, h = c.hidden, he = c.hidden_e;
  const std::string v = "vision_tower.vision_model.";
  // --- ViT
  VL_TRY(e.add_lin("", &e.patch, d, e.kpad));
  e.add_slot(v + "embeddings.patch_embedding.weight", e.patch.w, DT_BF16, d, 3 * c.patch_size * c.patch_size, e.kpad);
  VL_TRY(e.add_vec(v + "embeddings.patch_embedding.bias", &e.patch_b, d));
  VL_TRY(e.add_vec(v + "embeddings.class_embedding", &e.cls, d));
  VL_TRY(e.add_vec(v + "embeddings.position_embedding", &e.pos, (long long)(e.P + 1) * d));
  e.vit.resize(c.vit_layers);
  for (int i = 0; i < c.vit_layers; ++i) {
    VitLayer& L = e.vit[i];
    const std::string p = v + "encoder.layers." + std::to_string(i) + ".";
    VL_TRY(e.add_vec(p + "norm1.weight", &L.n1w, d));
    VL_TRY(e.add_vec(p + "norm1.bias", &L.n1b, d));
    VL_TRY(e.add_lin(p + "attn.qkv.weight", &L.qkv, 3 * d, d));
    VL_TRY(e.add_vec(p + "attn.qkv.bias", &L.bqkv, 3 * d));
    VL_TRY(e.add_lin(p + "attn.proj.weight", &L.proj, d, d));
    VL_TRY(e.add_vec(p + "attn.proj.bias", &L.bproj, d));
    VL_TRY(e.add_vec(p + "ls1", &L.ls1, d));
    VL_TRY(e.add_vec(p + "norm2.weight", &L.n2w, d));
    VL_TRY(e.add_vec(p + "norm2.bias", &L.n2b, d));
    VL_TRY(e.add_lin(p + "mlp.fc1.weight", &L.fc1, f, d));
    VL_TRY(e.add_vec(p + "mlp.fc1.bias", &L.bfc1, f));
    VL_TRY(e.add_lin(p + "mlp.fc2.weight", &L.fc2, d, f));
    VL_TRY(e.add_vec(p + "mlp.fc2.bias", &L.bfc2, d));
    VL_TRY(e.add_vec(p + "ls2", &L.ls2, d));
  }
  // --- projector
  VL_TRY(e.add_vec("multi_modal_projector.0.weight", &e.pj_lnw, 4 * d));
  VL_TRY(e.add_vec("multi_modal_projector.0.bias", &e.pj_lnb, 4 * d));
  VL_TRY(e.add_lin("multi_modal_projector.1.weight", &e.pj1, h, 4 * d));
  VL_TRY(e.add_vec("multi_modal_projector.1.bias", &e.pj_b1, h));
  VL_TRY(e.add_lin("multi_modal_projector.3.weight", &e.pj2, h, h));
  VL_TRY(e.add_vec("multi_modal_projector.3.bias", &e.pj_b2, h));
  VL_TRY(e.alloc(&e.embed, (size_t)c.embed_rows * h));
  e.add_slot("embed_tokens.weight", e.embed, DT_BF16, c.embed_rows, h);
  // --- LLM + expert
  // gate/up rows and down columns past `mlp` (up to the padded width mp) stay zero: silu(0) * 0 = 0 feeds zero
  // columns into a zero-weighted slice of the down projection
  auto add_llm = [&](std::vector<LlmLayer>& Ls, const std::string& prefix, int hid, int mlp, int mp) -> int {
    Ls.resize(c.layers);
    for (int i = 0; i < c.layers; ++i) {
      LlmLayer& L = Ls[i];
      const std::string p = prefix + "layers." + std::to_string(i) + ".";
      VL_TRY(e.add_vec(p + "input_layernorm.weight", &L.ln_in, hid));
      VL_TRY(e.add_vec(p + "post_attention_layernorm.weight", &L.ln_post, hid));
      VL_TRY(e.add_lin("", &L.qkv, e.qkvN, hid));
      e.add_slot(p + "self_attn.q_proj.weight", L.qkv.w, DT_BF16, e.qdim, hid, hid, 0);
      e.add_slot(p + "self_attn.k_proj.weight", L.qkv.w, DT_BF16, e.kvdim, hid, hid, e.qdim);
      e.add_slot(p + "self_attn.v_proj.weight", L.qkv.w, DT_BF16, e.kvdim, hid, hid, e.qdim + e.kvdim);
      VL_TRY(e.alloc(&L.bqkv, (size_t)e.qkvN));
      e.add_slot(p + "self_attn.q_proj.bias", L.bqkv, DT_F32, 1, e.qdim);
      e.add_slot(p + "self_attn.k_proj.bias", L.bqkv + e.qdim, DT_F32, 1, e.kvdim);
      e.add_slot(p + "self_attn.v_proj.bias", L.bqkv + e.qdim + e.kvdim, DT_F32, 1, e.kvdim);
      VL_TRY(e.add_lin(p + "self_attn.o_proj.weight", &L.o, hid, e.qdim));
      VL_TRY(e.add_lin("", &L.gu, 2 * mp, hid));
      e.add_slot(p + "mlp.gate_proj.weight", L.gu.w, DT_BF16, mlp, hid, hid, 0, PK_GATE);
      e.add_slot(p + "mlp.up_proj.weight", L.gu.w, DT_BF16, mlp, hid, hid, 0, PK_UP);
      VL_TRY(e.add_lin("", &L.down, hid, mp));
      e.add_slot(p + "mlp.down_proj.weight", L.down.w, DT_BF16, hid, mlp, mp);
    }
    return 0;
  };
  VL_TRY(add_llm(e.llm, "joint_model.mixtures.vlm.", h, c.mlp, e.mlp_p));
  // text head of the chat VLM (optional: the policy call never reads them, pizero_internvl.py:192-193 skip rule)
  VL_TRY(e.add_vec("joint_model.mixtures.vlm.norm.weight", &e.llm_norm, h));
  e.slots.find("joint_model.mixtures.vlm.norm.weight")->second.optional = true;
  VL_TRY(add_llm(e.exp, "joint_model.mixtures.action.", he, c.mlp_e, e.mlp_e_p));
  VL_TRY(e.add_vec("joint_model.mixtures.action.norm.weight", &e.final_norm, he));
  // --- heads
  VL_TRY(e.add_vec("action_encoder.linear_1.weight", &e.ae_w1, (long long)he * c.action_dim));
  VL_TRY(e.add_vec("action_encoder.linear_1.bias", &e.ae_b1, he));
  VL_TRY(e.alloc(&e.ae_w2a, (size_t)he * he));
  VL_TRY(e.add_lin("", &e.ae_w2b, he, he));
  e.add_slot("action_encoder.linear_2.weight", e.ae_w2a, DT_F32, he, 2 * he, he, 0, PK_PLAIN, 0, he);
  e.add_slot("action_encoder.linear_2.weight", e.ae_w2b.w, DT_BF16, he, 2 * he, he, 0, PK_PLAIN, he, he);
  VL_TRY(e.add_vec("action_encoder.linear_2.bias", &e.ae_b2, he));
  VL_TRY(e.add_lin("action_encoder.linear_3.weight", &e.ae_w3, he, he));
  VL_TRY(e.add_vec("action_encoder.linear_3.bias", &e.ae_b3, he));
  VL_TRY(e.add_vec("proprio_encoder.weight", &e.pe_w, (long long)he * c.proprio_dim));
  VL_TRY(e.add_vec("proprio_encoder.bias", &e.pe_b, he));
  VL_TRY(e.add_vec("action_decoder.weight", &e.dec_w, (long long)c.action_dim * he));
  VL_TRY(e.add_vec("action_decoder.bias", &e.dec_b, c.action_dim));

  // --- activations
  const int Ni = e.Bc * e.n_img;
  const int vrows = Ni * (e.P + 1);
  const int prow = e.Bc * e.S;
  const int arow = e.Bc * e.H;
  VL_TRY(e.add_act(&e.cols, Ni * e.P, e.kpad));
  VL_TRY(e.alloc(&e.vx, (size_t)vrows * d));
  VL_TRY(make_tmap_out(&e.vx_otm, e.vx, vrows, d, 4));
  VL_TRY(e.add_act(&e.vn, vrows, d));
  VL_TRY(e.alloc(&e.vqkv, (size_t)vrows * 3 * d));
  VL_TRY(make_tmap_out(&e.vqkv_otm, e.vqkv, vrows, 3 * d, 2));
  if (vrows >= 128) VL_TRY(make_tmap_bf16(&e.vqkv_tm, e.vqkv, vrows, 3 * d, 128));
  VL_TRY(e.add_act(&e.vattn, vrows, d));
  VL_TRY(e.add_act(&e.vmlp, vrows, f));
  VL_TRY(e.add_act(&e.ps, Ni * e.T, 4 * d));
  VL_TRY(e.add_act(&e.pa, Ni * e.T, h));
  VL_TRY(e.alloc(&e.xv, (size_t)prow * h));
  VL_TRY(make_tmap_out(&e.xv_otm, e.xv, prow, h, 4));
  VL_TRY(e.add_act(&e.xn, prow, h));
  VL_TRY(e.alloc(&e.q, (size_t)prow * e.qdim));
  VL_TRY(e.add_act(&e.ao, prow, e.qdim));
  VL_TRY(e.add_act(&e.act, prow, e.mlp_p));
  VL_TRY(e.alloc(&e.xp, (size_t)e.Bc * he));
  VL_TRY(make_tmap_out(&e.xp_otm, e.xp, e.Bc, he, 4));
  VL_TRY(e.add_act(&e.xpn, e.Bc, he));
  VL_TRY(e.alloc(&e.qp, (size_t)e.Bc * e.qdim));
  VL_TRY(e.add_act(&e.aop, e.Bc, e.qdim));
  VL_TRY(e.add_act(&e.actp, e.Bc, e.mlp_e_p));
  VL_TRY(e.alloc(&e.xa, (size_t)arow * he));
  VL_TRY(make_tmap_out(&e.xa_otm, e.xa, arow, he, 4));
  VL_TRY(e.add_act(&e.xan, arow, he));
  VL_TRY(e.alloc(&e.qa, (size_t)arow * e.qdim));
  VL_TRY(e.add_act(&e.aoa, arow, e.qdim));
  VL_TRY(e.add_act(&e.acta, arow, e.mlp_e_p));
  VL_TRY(e.add_act(&e.e1, arow, he));
  VL_TRY(e.add_act(&e.h2, arow, he));
  VL_TRY(e.alloc(&e.action, (size_t)arow * c.action_dim));
  e.ss_rows = std::max(arow, e.Bc);
  e.ss_floats = (size_t)(c.num_inference_steps + 1) * (2 * c.layers + 1) * e.ss_rows;
  VL_TRY(e.alloc(&e.ss, e.ss_floats));
  e.ln_ctr_blocks = (e.Bc * e.n_img * (e.P + 1) + 255) / 256;
  VL_TRY(e.alloc(&e.ln_ctr, (size_t)2 * c.vit_layers * e.ln_ctr_blocks));
  e.llm_ss_slots = c.hidden / 128;
  VL_TRY(e.alloc(&e.llm_ss, (size_t)2 * e.Bc * e.S * e.llm_ss_slots));
  VL_TRY(e.alloc(&e.tile_ctr, 4096));
  VL_TRY(e.alloc(&e.trace, Engine::TRACE_MAX * 16));
  {
    // split-K partial workspaces: [splits <= 32][rows padded to 256][width]. The expert streams always split (o / down
    // projections); the ViT / LLM residual projections only while their tiles cover less than half of the SMs
    // (option "small_m_split"), i.e. up to rows_trig(N) rows.
    auto r256 = [](long long r) { return (r + 255) / 256 * 256; };
    auto rows_trig = [](int N) { return 128LL * std::max(1, 74 / (N / 128)); };
    size_t need = (size_t)32 * r256(arow) * he;
    need = std::max<size_t>(need, (size_t)32 * r256(std::min<long long>(vrows, rows_trig(d))) * d);
    need = std::max<size_t>(need, (size_t)32 * r256(std::min<long long>(prow, rows_trig(h))) * h);
    e.ws_main.floats = need;
    VL_TRY(e.alloc(&e.ws_main.part, need));
    VL_TRY(e.alloc(&e.ws_main.splits, 1));
    e.ws_side.floats = (size_t)32 * r256(e.Bc) * he;
    VL_TRY(e.alloc(&e.ws_side.part, e.ws_side.floats));
    VL_TRY(e.alloc(&e.ws_side.splits, 1));
  }
  VL_TRY(e.alloc(&e.xp_layers, (size_t)c.layers));
  VL_TRY(e.alloc(&e.xp_bar, (size_t)(c.num_inference_steps + 1) * e.xp_bar_words()));
  VL_TRY(e.alloc(&e.xp_timing, (size_t)(c.num_inference_steps + 1) * 7 * c.layers));
  VL_TRY(e.alloc(&e.xfix, (size_t)(c.num_inference_steps + 1) * e.xfix_words()));
  const size_t kv_elems = (size_t)c.layers * c.n_kv_heads * e.kv_rows * 128;
  VL_TRY(e.alloc(&e.kc, kv_elems));
  VL_TRY(e.alloc(&e.vc, kv_elems));
  if (c.n_heads / c.n_kv_heads <= 8) {
    VL_TRY(make_tmap_q3d(&e.q3d_tm, e.q, (uint64_t)prow, c.n_heads, c.n_heads / c.n_kv_heads));
    VL_TRY(make_tmap_bf16(&e.kc_tm, e.kc, (uint64_t)c.layers * c.n_kv_heads * e.kv_rows, 128, 32));
    VL_TRY(make_tmap_bf16(&e.vc_tm, e.vc, (uint64_t)c.layers * c.n_kv_heads * e.kv_rows, 128, 32));
  }
  // --- plan buffers
  PlanBuffers& pb = e.pb;
  VL_TRY(e.alloc(&pb.valid_len, e.Bc));
  VL_TRY(e.alloc(&pb.cu, e.Bc + 1));
  VL_TRY(e.alloc(&pb.kvoff, e.Bc));
  VL_TRY(e.alloc(&pb.m_tot, 1));
  VL_TRY(e.alloc(&pb.row_pos, prow));
  VL_TRY(e.alloc(&pb.row_kvrow, prow));
  VL_TRY(e.alloc(&pb.row_src, prow));
  VL_TRY(e.alloc(&pb.img_row_map, e.Bc * e.n_img * e.T));
  VL_TRY(e.alloc(&pb.img_cnt, e.Bc));
  VL_TRY(e.alloc(&pb.img_off, e.Bc));
  VL_TRY(e.alloc(&pb.prop_pos, e.Bc));
  VL_TRY(e.alloc(&pb.prop_kvrow, e.Bc));
  VL_TRY(e.alloc(&pb.act_pos, arow));
  VL_TRY(e.alloc(&pb.act_kvrow, arow));
  VL_TRY(e.alloc(&pb.pre_kv_len, e.Bc));
  VL_TRY(e.alloc(&pb.den_q_start, e.Bc));
  VL_TRY(e.alloc(&pb.den_q_len, e.Bc));
  VL_TRY(e.alloc(&pb.den_kv_len, e.Bc));
  VL_TRY(e.alloc(&pb.prop_q_start, e.Bc));
  VL_TRY(e.alloc(&pb.prop_q_len, e.Bc));
  VL_TRY(e.alloc(&pb.prop_kv_len, e.Bc));
  VL_TRY(e.alloc(&pb.bad_id, 1));
  // --- I/O staging
  VL_TRY(e.alloc(&e.in_ids, (size_t)e.Bc * e.S));
  VL_TRY(e.alloc(&e.in_valid, e.Bc));
  const size_t pix_elems = (size_t)Ni * 3 * c.image_size * c.image_size;
  VL_TRY(e.alloc((float**)&e.in_pix, pix_elems));
  VL_TRY(e.alloc((float**)&e.in_prop, (size_t)e.Bc * c.proprio_dim));
  VL_TRY(e.alloc((float**)&e.in_noise, (size_t)arow * c.action_dim));
  VL_TRY(e.alloc((float**)&e.out_act, (size_t)e.Bc * c.horizon_steps * c.action_dim));
  VL_TRY(e.alloc((float**)&e.in_x1, (size_t)arow * c.action_dim));
  VL_TRY(e.alloc((float**)&e.in_t, e.Bc));
  VL_TRY(e.alloc(&e.flow_tbias, (size_t)e.Bc * c.hidden_e));
  VL_TRY(e.alloc(&e.flow_z, (size_t)arow * c.hidden_e));
  VL_TRY(e.alloc(&e.flow_vel, (size_t)arow * c.action_dim));
  VL_TRY(e.alloc(&e.flow_loss, 4));
  VL_TRY(e.alloc(&e.tail_ctrs, Engine::TAIL_CTRS));
  {
    const int bpad = (e.Bc + 255) / 256 * 256;
    e.gen_cap = e.S + 1 + e.H;
    e.lm_head.N = (c.embed_rows + 127) / 128 * 128;
    e.lm_head.K = h;
    VL_TRY(e.alloc(&e.x_dec, (size_t)bpad * h));
    VL_TRY(make_tmap_out(&e.xdec_otm, e.x_dec, bpad, h, 4));
    VL_TRY(e.alloc(&e.gen.next_tok, e.Bc));
    VL_TRY(e.alloc(&e.gen.n_gen, 4));
    VL_TRY(e.alloc(&e.gen.finished, e.Bc));
    VL_TRY(e.alloc(&e.gen.tokens, (size_t)e.Bc * e.gen_cap));
    VL_TRY(e.alloc(&e.gen.pos, e.Bc));
    VL_TRY(e.alloc(&e.gen.kvrow, e.Bc));
    VL_TRY(e.alloc(&e.gen.kv_len, e.Bc));
    VL_TRY(e.alloc(&e.gen.q_start, e.Bc));
    VL_TRY(e.alloc(&e.gen.q_len, e.Bc));
    const size_t max_splits = (size_t)(e.gen_cap + 63) / 64;
    VL_TRY(e.alloc(&e.dec_ws_o, (size_t)e.Bc * c.n_heads * max_splits * 128));
    VL_TRY(e.alloc(&e.dec_ws_ml, (size_t)e.Bc * c.n_heads * max_splits * 2));
  }
  e.pinned_bytes = (size_t)e.Bc * e.S * 8 + pix_elems * 4 + (size_t)e.Bc * 4 + (size_t)e.Bc * c.proprio_dim * 4 +
                   (size_t)arow * c.action_dim * 4 + (size_t)e.Bc * c.horizon_steps * c.action_dim * 4 + 4096;
  VL_CUDA_CHECK(cudaMallocHost(&e.pinned, e.pinned_bytes));
  memset(e.pinned, 0, e.pinned_bytes);
  return 0;
}

int engine_set_weight(Engine& e, std::string key, const void* data, const int64_t* shape, int ndim, int dtype) {
  VL_CHECK(!e.finalized, "set_weight after finalize");
  VL_CHECK(dtype == DT_F32 || dtype == DT_BF16, "weight dtype must be f32 or bf16");
  static const std::pair<const char*, const char*> aliases[] = {
      {"_orig_mod.", ""},
      {"joint_model.mixtures.proprio.", "joint_model.mixtures.action."},
      {"internvl_model.vision_model.", "vision_tower.vision_model."},
      {"internvl_model.mlp1.", "multi_modal_projector."},
      {"internvl_model.language_model.model.embed_tokens.", "embed_tokens."},
      {"internvl_model.language_model.model.layers.", "joint_model.mixtures.vlm.layers."},
      {"internvl_model.action_expert.model.layers.", "joint_model.mixtures.action.layers."},
      {"internvl_model.action_expert.model.norm.", "joint_model.mixtures.action.norm."},
      {"internvl_model.language_model.model.norm.", "joint_model.mixtures.vlm.norm."},
      {"internvl_model.language_model.lm_head.", "lm_head."},
      {"language_model.lm_head.", "lm_head."},
  };
  for (auto& a : aliases) {
    const size_t n = strlen(a.first);
    if (key.compare(0, n, a.first) == 0) key = std::string(a.second) + key.substr(n);
  }
  if (key == "lm_head.weight" && e.slots.find(key) == e.slots.end()) {
    // output head of the chat VLM: storage only when a checkpoint brings one (otherwise the tied embedding table)
    VL_TRY(e.alloc(&e.lm_head.w, (size_t)e.c.embed_rows * e.c.hidden));
    e.add_slot(key, e.lm_head.w, DT_BF16, e.c.embed_rows, e.c.hidden);
    e.lm_head_own = true;
  }
  auto range = e.slots.equal_range(key);
  if (range.first == range.second) return 1;  // not a tensor of this path
  long long numel = 1;
  for (int i = 0; i < ndim; ++i) numel *= shape[i];
  const size_t esz = dtype == DT_F32 ? 4 : 2;
  // stage host data on the device
  cudaPointerAttributes attr{};
  const void* src = data;
  void* tmp = nullptr;
  cudaError_t pe = cudaPointerGetAttributes(&attr, data);
  if (pe != cudaSuccess) cudaGetLastError();
  if (pe != cudaSuccess || attr.type == cudaMemoryTypeUnregistered || attr.type == cudaMemoryTypeHost) {
    VL_CUDA_CHECK(cudaMalloc(&tmp, (size_t)numel * esz));
    VL_CUDA_CHECK(cudaMemcpy(tmp, data, (size_t)numel * esz, cudaMemcpyHostToDevice));
    src = tmp;
  }
  int rc = 0;
  for (auto it = range.first; it != range.second && rc == 0; ++it) {
    Slot& s = it->second;
    if (numel != s.rows * s.cols) {
      set_last_error("set_weight: " + key + " has " + std::to_string(numel) + " elements, expected " +
                     std::to_string(s.rows * s.cols));
      rc = -2;
      break;
    }
    const long long total = s.rows * s.ncols;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
#define VL_PACK(TS, TD)                                                                                       \
  pack_kernel<TS, TD><<<blocks, 256>>>((const TS*)src, (TD*)s.dst, s.rows, s.cols, s.src_col0, s.ncols,       \
                                       s.dst_ld, s.dst_row_off, s.mode)
    if (dtype == DT_F32 && s.dst_dtype == DT_F32) VL_PACK(float, float);
    else if (dtype == DT_F32 && s.dst_dtype == DT_BF16) VL_PACK(float, bf16);
    else if (dtype == DT_BF16 && s.dst_dtype == DT_F32) VL_PACK(bf16, float);
    else VL_PACK(bf16, bf16);
#undef VL_PACK
    if (cudaGetLastError() != cudaSuccess) {
      set_last_error("set_weight: pack kernel launch failed for " + key);
      rc = -1;
    }
    s.filled = true;
  }
  if (tmp) {
    cudaDeviceSynchronize();
    cudaFree(tmp);
  }
  return rc;
}

int engine_finalize(Engine& e) {
  const vlaser_cfg& c = e.c;
  for (auto& kv : e.slots) {
    if (!kv.second.filled && !kv.second.optional) {
      set_last_error("finalize: missing weight " + kv.first);
      return -3;
    }
  }
  VL_CUDA_CHECK(cudaDeviceSynchronize());
  // tensor maps for every packed weight
  VL_TRY(e.make_lin_maps(&e.patch));
  for (auto& L : e.vit) {
    VL_TRY(e.make_lin_maps(&L.qkv));
    VL_TRY(e.make_lin_maps(&L.proj));
    VL_TRY(e.make_lin_maps(&L.fc1));
    VL_TRY(e.make_lin_maps(&L.fc2));
  }
  VL_TRY(e.make_lin_maps(&e.pj1));
  VL_TRY(e.make_lin_maps(&e.pj2));
  for (auto* Ls : {&e.llm, &e.exp})
    for (auto& L : *Ls) {
      VL_TRY(e.make_lin_maps(&L.qkv));
      VL_TRY(e.make_lin_maps(&L.o));
      VL_TRY(e.make_lin_maps(&L.gu));
      VL_TRY(e.make_lin_maps(&L.down));
    }
  VL_TRY(e.make_lin_maps(&e.ae_w2b));
  VL_TRY(e.make_lin_maps(&e.ae_w3));
  {
    // text head: maps over the true row count -- TMA zero-fills the rows of the last 128-row box behind embed_rows
    if (!e.lm_head_own) e.lm_head.w = e.embed;  // tie_word_embeddings (Qwen2.5-1.5B; pizero_internvl.py:325-334)
    VL_TRY(make_tmap_bf16(&e.lm_head.tm128, e.lm_head.w, c.embed_rows, c.hidden, 128));
    VL_TRY(make_tmap_bf16(&e.lm_head.tm64, e.lm_head.w, c.embed_rows, c.hidden, 64));
    e.lm_head.has256 = false;
    e.llm_norm_set = e.slots.find("joint_model.mixtures.vlm.norm.weight")->second.filled;
    VL_TRY(e.alloc(&e.logits, (size_t)e.Bc * e.lm_head.N));
  }
  {
    std::vector<ExpertLayerDev> tab(c.layers);
    for (int i = 0; i < c.layers; ++i) {
      const LlmLayer& L = e.exp[i];
      memset(&tab[i], 0, sizeof(ExpertLayerDev));
      tab[i].tm_qkv = L.qkv.tm128;
      tab[i].tm_o = L.o.tm128;
      tab[i].tm_gu = L.gu.tm128;
      tab[i].tm_down = L.down.tm128;
      tab[i].ln_in = L.ln_in;
      tab[i].ln_post = L.ln_post;
      tab[i].bqkv = L.bqkv;
      tab[i].w_qkv = L.qkv.w;
      tab[i].w_o = L.o.w;
      tab[i].w_gu = L.gu.w;
      tab[i].w_down = L.down.w;
    }
    VL_CUDA_CHECK(cudaMemcpy(e.xp_layers, tab.data(), tab.size() * sizeof(ExpertLayerDev), cudaMemcpyHostToDevice));
  }
  // RoPE tables (HF Qwen2RotaryEmbedding, default rope): cos/sin(pos * theta^(-2i/128)), i < 64
  const int max_pos = e.S + e.H + 2;  // policy call: ids 1..S, 2..H+1; text generation: 0..S+H (the K/V window)
  std::vector<float> ct((size_t)max_pos * 64), st((size_t)max_pos * 64);
  for (int p = 0; p < max_pos; ++p)
    for (int i = 0; i < 64; ++i) {
      const float inv = 1.0f / std::pow((float)c.rope_theta, (float)(2 * i) / 128.0f);
      const float ang = (float)p * inv;  // fp32 product as in the reference
      ct[(size_t)p * 64 + i] = (float)std::cos((double)ang);
      st[(size_t)p * 64 + i] = (float)std::sin((double)ang);
    }
  VL_TRY(e.alloc(&e.cos_tab, ct.size()));
  VL_TRY(e.alloc(&e.sin_tab, st.size()));
  VL_CUDA_CHECK(cudaMemcpy(e.cos_tab, ct.data(), ct.size() * 4, cudaMemcpyHostToDevice));
  VL_CUDA_CHECK(cudaMemcpy(e.sin_tab, st.data(), st.size() * 4, cudaMemcpyHostToDevice));
  // time-embedding contribution of ActionEncoder.linear_2 for each Euler step:
  //   tb[k] = b2 + W2[:, :he] * cat(sin(t_k w), cos(t_k w))   (modules.py:15-22, 39-53)
  const int he = c.hidden_e, half = he / 2, steps = c.num_inference_steps;
  std::vector<float> w2a((size_t)he * he), b2(he), tb((size_t)steps * he);
  VL_CUDA_CHECK(cudaMemcpy(w2a.data(), e.ae_w2a, w2a.size() * 4, cudaMemcpyDeviceToHost));
  VL_CUDA_CHECK(cudaMemcpy(b2.data(), e.ae_b2, b2.size() * 4, cudaMemcpyDeviceToHost));
  float t = 0.f;
  const float dtf = (float)(1.0 / steps);
  const float emb = (float)(std::log((double)c.time_max_period) / (half - 1));
  std::vector<double> tau(he);
  for (int k = 0; k < steps; ++k) {
    for (int i = 0; i < half; ++i) {
      const float w = std::exp((float)i * -emb);
      const float a = t * w;
      tau[i] = std::sin((double)a);
      tau[half + i] = std::cos((double)a);
    }
    for (int n = 0; n < he; ++n) {
      double acc = b2[n];
      for (int j = 0; j < he; ++j) acc += (double)w2a[(size_t)n * he + j] * tau[j];
      tb[(size_t)k * he + n] = (float)acc;
    }
    t += dtf;
  }
  VL_TRY(e.alloc(&e.time_bias, tb.size()));
  VL_CUDA_CHECK(cudaMemcpy(e.time_bias, tb.data(), tb.size() * 4, cudaMemcpyHostToDevice));
  e.finalized = true;
  return 0;
}

// ------------------------------------------------------------------------------------------
// schedule
// ------------------------------------------------------------------------------------------
struct Run {
  Engine& e;
  cudaStream_t st;
  int B;
  int io;
  int rc = 0;

  long long* tr(const char* name, int layer = -1) {
    if (!e.trace_on || e.trace_idx >= Engine::TRACE_MAX) return nullptr;
    e.trace_names += name;
    if (layer >= 0) e.trace_names += " L" + std::to_string(layer);
    e.trace_names += st == e.side_stream ? " [side]\n" : "\n";
    return e.trace + 16 * e.trace_idx++;
  }
  const char* tr_name = "gemm";
  int tr_layer = -1;
  bool gemm_reports_blocks = false;  // the last gemm() bumps its EpiParams::done_ctr (see Engine::ln_overlap)

  // ---- deterministic split-K bookkeeping: the partials a split-K residual projection left for the next reader of x
  struct Pending {
    float* x = nullptr;
    SplitKPartials sp;
    int rows = 0, width = 0;
    const int* rows_dev = nullptr;
  };
  Pending pend[2];  // main stream, side stream
  int si() const { return st == e.side_stream ? 1 : 0; }
  // the partials pending on x, for a norm launch that will fold them in (nullptr: none)
  const SplitKPartials* take(const float* x, SplitKPartials& tmp) {
    Pending& p = pend[si()];
    if (p.x == nullptr || p.x != x) return nullptr;
    tmp = p.sp;
    p = Pending();
    return &tmp;
  }
  // any other reader of the stream: finish the residual update first
  int flush() {
    Pending& p = pend[si()];
    if (p.x == nullptr) return 0;
    const Pending q = p;
    p = Pending();
    e.launches++;
    return splitk_reduce_launch(q.x, q.rows, q.rows_dev, q.width, q.sp, st);
  }

  // fused tail (EpiParams::tail_ctr): a caller that will launch RMSNorm(x) * gamma -> out right behind a split-K residual
  // projection offers it to gemm(); when the launch qualifies (pair kernel, 256 x 128 tiles, deterministic partials) the
  // projection does that work itself and tail_done tells the caller (and the next block_pre) to skip its norm launch
  struct Tail {
    const float* gamma = nullptr;
    bf16* out = nullptr;
  } tail_offer;
  struct {
    const float* x = nullptr;
    const float* gamma = nullptr;
  } tail_done[2];
  bool take_tail(const float* x, const float* gamma) {
    auto& t = tail_done[si()];
    if (t.x == nullptr || t.x != x || t.gamma != gamma) return false;
    t.x = nullptr;
    return true;
  }

  int gemm(const Act& a, const Lin& w, int M, const int* M_dev, int epi, EpiParams ep, int max_splits = 1,
           int stages = 0, const CUtensorMap* out_tm = nullptr) {
    const Tail offer = tail_offer;
    tail_offer = Tail();
    GemmDims d;
    d.trace = tr(tr_name, tr_layer);
    d.gelu_warps = e.gelu_warps;
    d.M = M;
    d.N = w.N;
    d.K = w.K;
    d.M_dev = M_dev;
    d.max_splits = max_splits;
    d.stages = stages;
    e.launches++;
    if (e.simt) return gemm_launch_simt(a.p, w.w, d, ep, epi, st);
    // Few observations (B = 1, 2, ...): a residual projection of the ViT / LLM has a handful of output tiles (down
    // projection of the LLM at B = 1: 12 tiles for 27.5 MB of weights, 23 us on 12 SMs). Split K over the idle SMs
    // (fp32 red.adds into the residual stream, single-CTA kernel; the split count follows the live tile count in the
    // kernel). Large batches are untouched: their tile count already covers the GPU.
    if (e.small_m_split && epi == EPI_RESID_F32 && max_splits <= 1 && !ep.a_next &&
        (long long)((M + 127) / 128) * (w.N / 128) * 2 <= gemm_num_sms()) {
      max_splits = d.max_splits = 32;
      out_tm = nullptr;
    }
    const CUtensorMap* part_tm = nullptr;
    if (epi == EPI_RESID_F32 && pend[si()].x == ep.out) VL_TRY(flush());
    if (e.det_splitk && epi == EPI_RESID_F32 && max_splits > 1 && !ep.a_next) {
      Engine::SplitWs& ws = si() ? e.ws_side : e.ws_main;
      const long long stride = (long long)((M + 255) / 256 * 256) * w.N;
      const int fit = (int)std::min<long long>(max_splits, (long long)(ws.floats / (size_t)stride));
      if (fit >= 2) {
        max_splits = d.max_splits = fit;
        ep.part = ws.part;
        ep.part_stride = stride;
        ep.part_splits = ws.splits;
        auto it = ws.maps.find(w.N);
        if (it == ws.maps.end()) {
          CUtensorMap m;
          VL_TRY(make_tmap_out(&m, ws.part, ws.floats / (size_t)w.N, (uint64_t)w.N, 4));
          it = ws.maps.emplace(w.N, m).first;
        }
        part_tm = &it->second;
        VL_TRY(flush());
        Pending& p = pend[si()];
        p.x = reinterpret_cast<float*>(ep.out);
        p.sp.part = ws.part;
        p.sp.stride = stride;
        p.sp.splits = ws.splits;
        p.rows = M;
        p.rows_dev = M_dev;
        p.width = w.N;
      } else {
        max_splits = d.max_splits = 1;  // no room for partials: one split per tile (slower, still deterministic)
      }
    }
    int bn = 128;
    // large-M GEMMs: CTA-pair kernel (256x256 tile per pair) when there is at least one tile per pair
    // (with split-K through the TMA reduce-add epilogue the splits multiply the tile count)
    long long pair_tiles = (long long)((M + 255) / 256) * (w.N / 256);
    const bool pair_split = max_splits > 1 && epi == EPI_RESID_F32 && out_tm != nullptr && e.tma_epi && !ep.a_next;
    if (pair_split) pair_tiles *= std::min(max_splits, std::max(1, w.K / 128));
    // (a static M of at most one 128-row tile -- the proprio pass, a sub-batch chain of the denoise steps -- would leave
    // the second CTA of every pair multiplying zero rows: single-CTA kernel)
    const bool half_tile = M_dev == nullptr && M <= 128;
    if (e.pair_gemm && w.N % 256 == 0 && (max_splits <= 1 || pair_split) && pair_tiles >= e.pair_min_tiles && !half_tile) d.pair = 1;
    // one 256-row tile of rows, split-K partials (the o / down projections of the action expert): 256 x 128 tiles -- every
    // CTA then stores 64 KB of partials instead of 128 KB (the epilogue of these launches is bound by the ~55 GB/s an SM
    // stores through TMA, in-situ timeline r02j) and N / 128 x splits pairs cover the GPU with half as many splits
    if (d.pair && pair_split && ep.part && e.pair_bn128 && M_dev == nullptr && M <= 256 && w.N % 128 == 0) {
      const int tiles128 = w.N / 128;
      const int fit = std::max(1, (gemm_num_sms() / 2) / tiles128);
      if (fit >= 2) {
        d.pair_bn = 128;
        max_splits = d.max_splits = std::min(max_splits, fit);
      }
    }
    if (ep.part && pend[si()].x == reinterpret_cast<float*>(ep.out)) {
      pend[si()].sp.max_splits = d.max_splits;  // final cap
      // static row count on the pair kernel's TMA epilogue: the count itself is known here
      if (d.pair && pair_split && M_dev == nullptr && e.tma_epi && out_tm != nullptr && e.known_splits)
        pend[si()].sp.known_splits = gemm_pair_split_count(d, true);
    }
    // tile order: with many more rows than weight rows a wave of pairs should share A tiles (n fastest); measured on
    // B200: ViT GEMMs and the LLM o/down projections gain, the wide-N LLM qkv/gate-up GEMMs prefer m fastest
    d.n_fast = (M_dev == nullptr) ? (M > w.N) : (w.N <= 1536);
    if (!d.pair && epi != EPI_QKV_ROPE && w.has256) {
      const long long tiles256 = (long long)((M + 127) / 128) * (w.N / 256);
      // wide projections of a half tile (gate/up: N = 17920): 128 x 256 tiles read each weight once per CTA column and keep
      // the launch at N / 256 CTAs, so that two sub-batch chains fit on the GPU side by side
      if (tiles256 >= gemm_num_sms() || (half_tile && w.N >= 8192)) bn = 256;
    }
    const bool prof = e.profile_gemm && !e.use_graph;
    if (prof) {
      while (e.prof_events.size() < e.prof_used + 2) {
        cudaEvent_t ev;
        VL_CUDA_CHECK(cudaEventCreate(&ev));
        e.prof_events.push_back(ev);
      }
      const double rows = M_dev ? (double)e.profile_mtot : (double)M;
      e.prof_flops.push_back(2.0 * rows * (double)w.N * (double)w.K);
      // algorithmic bytes: both operands once + the output once (residual epilogues: fp32 read-modify-write)
      const double out_elems = rows * (double)(epi == EPI_SWIGLU_BF16 ? w.N / 2 : w.N);
      const double out_bytes = (epi == EPI_RESID_F32) ? 8.0 * out_elems : ((epi == EPI_STORE_F32 || epi == EPI_PATCH_EMBED) ? 4.0 : 2.0) * out_elems;
      e.prof_bytes.push_back(2.0 * (rows * (double)w.K + (double)w.N * (double)w.K) + out_bytes);
      e.prof_class.push_back((d.pair && rows > 512.0) ? 1 : 0);  // M <= 512: weight-streaming regime whatever the kernel
      VL_CUDA_CHECK(cudaEventRecord(e.prof_events[e.prof_used], st));
    }
    if (e.fuse_tail && offer.gamma && d.pair && d.pair_bn == 128 && ep.part && epi == EPI_RESID_F32 && w.N <= 1024 &&
        e.tail_used < Engine::TAIL_CTRS) {
      // all CTAs of the launch resident (one per SM): 2 * min(tiles, SMs / 2) <= SMs
      ep.tail_ctr = e.tail_ctrs + e.tail_used++;
      ep.tail_gamma = offer.gamma;
      ep.tail_out = offer.out;
      ep.tail_eps = e.c.rms_eps;
      ep.tail_err = e.pb.bad_id;
      pend[si()] = Pending();  // the projection itself finishes the residual update
      tail_done[si()].x = reinterpret_cast<const float*>(ep.out);
      tail_done[si()].gamma = offer.gamma;
    }
    // one tile per CTA pair at a static row count (the o / gate-up / down launches of a denoise step): the half-SM variant,
    // whose CTAs become resident and run their prologue while the previous launch of the chain still holds the SMs
    if (e.half_sm && d.pair && M_dev == nullptr && M <= 256 && out_tm != nullptr && e.tma_epi && !ep.a_next && !ep.row_ss &&
        !ep.tail_ctr) {
      const long long t = (long long)(w.N / d.pair_bn) * (epi == EPI_RESID_F32 ? std::max(1, d.max_splits) : 1);
      if (((epi == EPI_RESID_F32 && d.pair_bn == 128) || (epi == EPI_SWIGLU_BF16 && d.pair_bn == 256)) &&
          t <= gemm_num_sms() / 2)
        d.half_sm = 1;
    }
    // row-block completion counters: only the one-split residual epilogue of the pair kernel through TMA reports them
    if (ep.done_ctr && !(d.pair && d.pair_bn == 256 && !d.half_sm && epi == EPI_RESID_F32 && d.max_splits <= 1 && out_tm != nullptr &&
                         e.tma_epi && !ep.a_next && !ep.tail_ctr && M_dev == nullptr))
      ep.done_ctr = nullptr;
    gemm_reports_blocks = ep.done_ctr != nullptr;
    const CUtensorMap& wtm = (d.pair && d.pair_bn == 128) ? w.tm64 : (bn == 256 ? w.tm256 : w.tm128);
    const int rc = gemm_launch(a.tm, wtm, d, ep, bn, epi, st, e.tma_epi ? out_tm : nullptr, part_tm);
    if (prof) {
      VL_CUDA_CHECK(cudaEventRecord(e.prof_events[e.prof_used + 1], st));
      e.prof_used += 2;
    }
    return rc;
  }

  EpiParams rope_ep(bf16* qout, const float* bias, const int* pos, const int* kvrow, int layer) {
    EpiParams ep;
    ep.out = qout;
    ep.ldo = e.qdim;
    ep.bias = bias;
    ep.row_pos = pos;
    ep.row_kvrow = kvrow;
    ep.cos_tab = e.cos_tab;
    ep.sin_tab = e.sin_tab;
    const size_t off = (size_t)layer * e.c.n_kv_heads * e.kv_rows * 128;
    ep.k_cache = e.kc + off;
    ep.v_cache = e.vc + off;
    ep.kv_head_stride = e.kv_rows * 128;
    ep.n_q_heads = e.c.n_heads;
    ep.n_kv_heads = e.c.n_kv_heads;
    return ep;
  }

  // the norm launch behind a residual GEMM that reports its row blocks: arrivals per 256-row block = (N / 256) column
  // tiles x 2 CTAs x 8 epilogue warps
  RowBlocksReady ln_ready;
  void note_blocks(const int* ctr, int width) {
    ln_ready = RowBlocksReady();
    if (ctr && gemm_reports_blocks) {
      ln_ready.ctr = ctr;
      ln_ready.expect = (width / 256) * 2 * 8;
      ln_ready.err = e.pb.bad_id;
    }
  }
  int vit() {
    const vlaser_cfg& c = e.c;
    const int d = c.vit_hidden, f = c.vit_mlp;
    const int Ni = B * e.n_img, rows = Ni * (e.P + 1);
    VL_TRY(im2col_launch(e.in_pix, io, e.cols.p, Ni, c.image_size, c.patch_size, e.kpad, st));
    VL_TRY(cls_rows_launch(e.vx, e.cls, e.pos, Ni, e.P + 1, d, st));
    e.launches += 2;
    EpiParams ep;
    ep.out = e.vx;
    ep.ldo = d;
    ep.bias = e.patch_b;
    ep.pos_emb = e.pos;
    ep.patches = e.P;
    tr_name = "vit_patch";
    VL_TRY(gemm(e.cols, e.patch, Ni * e.P, nullptr, EPI_PATCH_EMBED, ep));
    for (int i = 0; i < c.vit_layers; ++i) {
      VitLayer& L = e.vit[i];
      tr_layer = i;
      SplitKPartials spt;
      // (i > 0: under the fc2 GEMM of the previous layer when that launch reports its row blocks)
      VL_TRY(layernorm_launch(e.vx, L.n1w, L.n1b, e.vn.p, rows, d, c.vit_eps, ShuffleGather{}, st, take(e.vx, spt),
                              ln_ready.ctr ? &ln_ready : nullptr));
      ln_ready = RowBlocksReady();
      EpiParams e1;
      e1.out = e.vqkv;
      e1.ldo = 3 * d;
      e1.bias = L.bqkv;
      tr_name = "vit_qkv";
      VL_TRY(gemm(e.vn, L.qkv, rows, nullptr, EPI_STORE_BF16, e1, 1, 0, &e.vqkv_otm));
      if (e.vit_tc && !e.simt && (e.P + 1 == 257 || (e.P % 256 == 0 && e.P <= 1024))) {  // 224 px; 448 px tiles (two-pass kernel)
        VL_TRY(vit_attention_tc_launch(e.vqkv_tm, e.vattn.tm, e.vqkv, e.vattn.p, Ni, e.P + 1, c.vit_heads, 0.125f, st,
                                       tr("vit_attention", i)));
      } else if (e.P + 1 >= 128 && e.P + 1 <= 320) {
        VL_TRY(vit_attention_launch(e.vqkv, e.vattn.p, Ni, e.P + 1, c.vit_heads, 0.125f, st));
      } else {
      AttnParams ap;
      ap.q = e.vqkv;
      ap.k = e.vqkv + d;
      ap.v = e.vqkv + 2 * d;
      ap.q_row_stride = ap.kv_row_stride = 3 * d;
      ap.kv_head_stride = 64;
      ap.out = e.vattn.p;
      ap.out_row_stride = d;
      ap.uni_q_len = ap.uni_kv_len = ap.max_q_len = e.P + 1;
      ap.group = 1;
      ap.n_kv_heads = c.vit_heads;
      ap.batch = Ni;
      ap.scale = 0.125f;
      VL_TRY(attention_launch(ap, 64, st));
      }
      EpiParams e2;
      e2.out = e.vx;
      e2.ldo = d;
      e2.bias = L.bproj;
      e2.scale = L.ls1;
      tr_name = "vit_proj";
      if (e.ln_overlap && !e.simt) e2.done_ctr = e.ln_ctr + (size_t)(2 * i) * e.ln_ctr_blocks;
      VL_TRY(gemm(e.vattn, L.proj, rows, nullptr, EPI_RESID_F32, e2, 1, 0, &e.vx_otm));
      note_blocks(e2.done_ctr, d);
      VL_TRY(layernorm_launch(e.vx, L.n2w, L.n2b, e.vn.p, rows, d, c.vit_eps, ShuffleGather{}, st, take(e.vx, spt),
                              ln_ready.ctr ? &ln_ready : nullptr));
      ln_ready = RowBlocksReady();
      EpiParams e3;
      e3.out = e.vmlp.p;
      e3.ldo = f;
      e3.bias = L.bfc1;
      tr_name = "vit_fc1";
      VL_TRY(gemm(e.vn, L.fc1, rows, nullptr, EPI_GELU_BF16, e3, 1, 0, &e.vmlp.otm));
      EpiParams e4;
      e4.out = e.vx;
      e4.ldo = d;
      e4.bias = L.bfc2;
      e4.scale = L.ls2;
      tr_name = "vit_fc2";
      if (e.ln_overlap && !e.simt && i + 1 < c.vit_layers) e4.done_ctr = e.ln_ctr + (size_t)(2 * i + 1) * e.ln_ctr_blocks;
      VL_TRY(gemm(e.vmlp, L.fc2, rows, nullptr, EPI_RESID_F32, e4, 1, 0, &e.vx_otm));
      note_blocks(e4.done_ctr, d);
      e.launches += 3;
    }
    return flush();  // the last fc2 of a small batch may have left split-K partials
  }

  int embed() {
    const vlaser_cfg& c = e.c;
    const int d = c.vit_hidden, h = c.hidden;
    const int Ni = B * e.n_img, prows = Ni * e.T;
    ShuffleGather sg;
    sg.grid = e.G;
    sg.channels = d;
    if (e.text_no_image) {  // text-only prompt of the chat VLM: no tile went through the ViT
      VL_TRY(assemble_text_launch(e.embed, e.pb.row_src, e.pb.m_tot, e.xv, B * e.S, h, st));
      e.launches++;
      return 0;
    }
    VL_TRY(layernorm_launch(e.vx, e.pj_lnw, e.pj_lnb, e.ps.p, prows, 4 * d, c.proj_eps, sg, st));
    EpiParams e1;
    e1.out = e.pa.p;
    e1.ldo = h;
    e1.bias = e.pj_b1;
    tr_name = "proj1";
    tr_layer = -1;
    VL_TRY(gemm(e.ps, e.pj1, prows, nullptr, EPI_GELU_BF16, e1, 1, 0, &e.pa.otm));
    EpiParams e2;
    e2.out = e.xv;
    e2.ldo = h;
    e2.bias = e.pj_b2;
    e2.row_map = e.pb.img_row_map;
    tr_name = "proj2";
    VL_TRY(gemm(e.pa, e.pj2, prows, nullptr, EPI_STORE_F32, e2));
    VL_TRY(assemble_text_launch(e.embed, e.pb.row_src, e.pb.m_tot, e.xv, B * e.S, h, st));
    if (e.mode < 2) VL_TRY(small_linear_launch(e.in_prop, io, e.pe_w, e.pe_b, e.xp, DT_F32, B, c.proprio_dim, c.hidden_e, st));
    e.launches += 3;
    return 0;
  }

  // one transformer block of a stream (residual x fp32, expert or LLM weights)
  int abl = 0;  // timing aid (option "ablate"): launches of the denoise steps to leave out, results are then garbage
  int block_pre(LlmLayer& L, float* x, Act& xn, bf16* qout, int M, const int* M_dev, int hid, const int* pos,
                const int* kvrow, int layer, const L2Prefetch* pf = nullptr) {
    tr_layer = layer;
    EpiParams qe = rope_ep(qout, L.bqkv, pos, kvrow, layer);
    if (llm_fused_in && x == e.xv) {
      // fused chain: the down projection of the previous layer left bf16(x * ln_in) in xn and the row sums of squares in
      // site 1; the q|k|v epilogue scales its accumulator rows by rstd
      llm_consumer(qe, 1, hid);
    } else if (take_tail(x, L.ln_in)) {
      // the down projection of the previous layer already left RMSNorm(x) * ln_in in xn
    } else if (!(abl & 1)) {
      SplitKPartials spt;
      VL_TRY(rmsnorm_launch(x, L.ln_in, xn.p, M, M_dev, hid, e.c.rms_eps, st, pf, tr("norm_in", layer), take(x, spt)));
      e.launches++;
    }
    if (abl & 2) return 0;
    tr_name = "qkv";
    return gemm(xn, L.qkv, M, M_dev, EPI_QKV_ROPE, qe);
  }
  // ---- LLM prefill, RMSNorm fused across the GEMM chain (see Engine::fuse_norm_llm)
  bool llm_fused = false;     // this prefill runs fused (set by prefill())
  bool llm_fused_in = false;  // xn / site 1 hold the operand and statistics of the coming q|k|v projection
  void llm_consumer(EpiParams& ep, int site, int hid) {
    ep.row_ss = e.llm_ss + (size_t)site * e.Bc * e.S * e.llm_ss_slots;
    ep.ss_slots = e.llm_ss_slots;
    ep.inv_width = 1.0f / (float)hid;
    ep.eps = e.c.rms_eps;
  }
  void llm_producer(EpiParams& ep, Act& xn, const float* gamma, int site, int hid) {
    ep.a_next = xn.p;
    ep.lda_next = hid;
    ep.gamma_next = gamma;
    ep.ss_next = e.llm_ss + (size_t)site * e.Bc * e.S * e.llm_ss_slots;
    ep.ss_slots = e.llm_ss_slots;
  }
  int block_post(LlmLayer& L, float* x, Act& xn, Act& ao, Act& act, int M, const int* M_dev, int hid, int mlp,
                 int splits, const CUtensorMap* x_otm = nullptr, const L2Prefetch* pf = nullptr,
                 const float* next_gamma = nullptr) {
    EpiParams eo;
    eo.out = x;
    eo.ldo = hid;
    tr_name = "o";
    const bool fused = llm_fused && M_dev != nullptr && x == e.xv;
    if (fused) {
      EpiParams ef = eo;
      llm_producer(ef, xn, L.ln_post, 0, hid);
      VL_TRY(gemm(ao, L.o, M, M_dev, EPI_RESID_F32, ef, 1, 0, x_otm));
      EpiParams eg;
      eg.out = act.p;
      eg.ldo = mlp;
      llm_consumer(eg, 0, hid);
      tr_name = "gate_up";
      VL_TRY(gemm(xn, L.gu, M, M_dev, EPI_SWIGLU_BF16, eg, 1, 0, &act.otm));
      tr_name = "down";
      EpiParams ed = eo;
      if (next_gamma) llm_producer(ed, xn, next_gamma, 1, hid);
      llm_fused_in = next_gamma != nullptr;
      return gemm(act, L.down, M, M_dev, EPI_RESID_F32, ed, 1, 0, x_otm);
    }
    if (!abl) tail_offer.gamma = L.ln_post, tail_offer.out = xn.p;
    if (!(abl & 8)) VL_TRY(gemm(ao, L.o, M, M_dev, EPI_RESID_F32, eo, splits > 1 ? std::min(splits, e.xs_o) : 1, 0, x_otm));
    if (take_tail(x, L.ln_post)) {
      // the o projection already left RMSNorm(x) * ln_post in xn
    } else if (!(abl & 16)) {
      SplitKPartials spt;
      VL_TRY(rmsnorm_launch(x, L.ln_post, xn.p, M, M_dev, hid, e.c.rms_eps, st, pf, tr("norm_post", tr_layer), take(x, spt)));
      e.launches++;
    }
    EpiParams eg;
    eg.out = act.p;
    eg.ldo = mlp;
    tr_name = "gate_up";
    if (!(abl & 32)) VL_TRY(gemm(xn, L.gu, M, M_dev, EPI_SWIGLU_BF16, eg, 1, 0, &act.otm));
    if (abl & 64) return 0;
    tr_name = "down";
    // (half tile: 6 output tiles x 12 splits = 72 CTAs, again half of the GPU)
    const int xs_down = (M_dev == nullptr && M <= 128) ? std::min(e.xs_down, 12) : e.xs_down;
    if (!abl && next_gamma) tail_offer.gamma = next_gamma, tail_offer.out = xn.p;
    return gemm(act, L.down, M, M_dev, EPI_RESID_F32, eo, splits > 1 ? std::min(splits, xs_down) : 1, 0, x_otm);
  }

  // ---- expert streams (proprio token in the prefill, action tokens in the denoise steps): RMSNorm is fused across
  // the GEMM chain. `xn` always holds bf16(x * gamma) for the NEXT norm site and ss slot `site` its row sums of
  // squares; the consuming GEMM scales its accumulator rows by rsqrt(ss/he + eps).
  float* ss_slot(int pass, int site) { return e.ss + ((size_t)pass * (2 * e.c.layers + 1) + site) * e.ss_rows; }
  void norm_consumer(EpiParams& ep, int pass, int site) {
    ep.row_ss = ss_slot(pass, site);
    ep.inv_width = 1.0f / (float)e.c.hidden_e;
    ep.eps = e.c.rms_eps;
  }
  void norm_producer(EpiParams& ep, Act& xn, const float* gamma, int pass, int site) {
    ep.a_next = xn.p;
    ep.lda_next = e.c.hidden_e;
    ep.gamma_next = gamma;
    ep.ss_next = ss_slot(pass, site);
    ep.tile_counters = e.tile_ctr;
  }
  int exp_pre(int li, Act& xn, bf16* qout, int M, const int* pos, const int* kvrow, int pass) {
    EpiParams ep = rope_ep(qout, e.exp[li].bqkv, pos, kvrow, li);
    norm_consumer(ep, pass, 2 * li);
    return gemm(xn, e.exp[li].qkv, M, nullptr, EPI_QKV_ROPE, ep, 1, e.exp_stages);
  }
  int exp_post(int li, float* x, Act& xn, Act& ao, Act& act, int M, int pass, bool last_needs_norm) {
    const vlaser_cfg& c = e.c;
    LlmLayer& L = e.exp[li];
    EpiParams eo;
    eo.out = x;
    eo.ldo = c.hidden_e;
    norm_producer(eo, xn, L.ln_post, pass, 2 * li + 1);
    VL_TRY(gemm(ao, L.o, M, nullptr, EPI_RESID_F32, eo, 32, e.exp_stages));
    EpiParams eg;
    eg.out = act.p;
    eg.ldo = e.mlp_e_p;
    norm_consumer(eg, pass, 2 * li + 1);
    VL_TRY(gemm(xn, L.gu, M, nullptr, EPI_SWIGLU_BF16, eg, 1, e.exp_stages));
    EpiParams ed;
    ed.out = x;
    ed.ldo = c.hidden_e;
    if (li + 1 < c.layers) norm_producer(ed, xn, e.exp[li + 1].ln_in, pass, 2 * li + 2);
    else if (last_needs_norm) norm_producer(ed, xn, e.final_norm, pass, 2 * li + 2);
    return gemm(act, L.down, M, nullptr, EPI_RESID_F32, ed, 32, e.exp_stages);
  }

  // B = 1 (the latency case): the expert passes run on the weight-streaming single-observation kernel
  // (the flow-matching loss runs its one expert pass on the per-kernel schedule: the persistent kernels index their time
  // bias by Euler step)
  bool gemv() const { return e.gemv && B == 1 && e.H <= 8 && !e.simt && e.mode == 0; }
  bool mega() const { return (e.mega || gemv()) && !e.simt && e.mode == 0; }
  // largest power-of-two sub-grid count <= the option that keeps whole samples, a multiple of 32 rows and at least 64
  // rows per sub-grid
  int expert_streams(int M, int B) const {
    int s = std::min(e.xp_streams, (int)EXPERT_MAX_STREAMS);
    while (s > 1 && !(B % s == 0 && (M / s) % 32 == 0 && M % s == 0 && M / s >= 64)) s >>= 1;
    return std::max(s, 1);
  }

  // One pass of the expert transformer over M rows as a single persistent kernel (expert.cu). pass 0 = the proprio
  // token of the prefill (its K/V only in the last layer), pass k+1 = the action tokens of Euler step k.
  int expert_pass(int pass, int M, float* x, const CUtensorMap& x_otm, Act& xn, bf16* q, Act& ao, Act& act, const int* pos, const int* kvrow,
                  const int* q_start, const int* q_len, const int* kv_len, int max_q_len, bool last_kv_only) {
    const vlaser_cfg& c = e.c;
    ExpertPass p;
    p.layers = e.xp_layers;
    p.n_layers = c.layers;
    p.last_kv_only = last_kv_only ? 1 : 0;
    p.M = M;
    p.he = c.hidden_e;
    p.mlp = e.mlp_e_p;
    p.n_q_heads = c.n_heads;
    p.n_kv_heads = c.n_kv_heads;
    p.eps = c.rms_eps;
    p.x = x;
    p.xn = xn.p;
    p.q = q;
    p.ao = ao.p;
    p.act = act.p;
    p.row_pos = pos;
    p.row_kvrow = kvrow;
    p.cos_tab = e.cos_tab;
    p.sin_tab = e.sin_tab;
    p.k_cache = e.kc;
    p.v_cache = e.vc;
    p.kv_layer_stride = (long long)c.n_kv_heads * e.kv_rows * 128;
    p.kv_head_stride = e.kv_rows * 128;
    p.q_start = q_start;
    p.q_len = q_len;
    p.kv_start = e.pb.kvoff;
    p.kv_len = kv_len;
    p.max_q_len = max_q_len;
    p.max_kv_len = e.S + 1 + e.H;
    p.batch = B;
    p.scale = 0.08838834764831845f;  // 128^-0.5
    p.barrier = e.xp_bar + (size_t)pass * e.xp_bar_words();
    p.err_flag = e.pb.bad_id;  // shares the call's error word (reported by the next call / the host entry)
    p.xfix = (gemv() && e.det_splitk) ? e.xfix + (size_t)pass * e.xfix_words() : nullptr;
    p.n_streams = expert_streams(M, B);
    p.flags = e.xp_flags;
    p.timing = e.xp_time ? e.xp_timing + (size_t)pass * 7 * c.layers : nullptr;
    e.launches++;
    const bool prof = e.profile_gemm && !e.use_graph;
    if (prof) {
      while (e.prof_events.size() < e.prof_used + 2) {
        cudaEvent_t ev;
        VL_CUDA_CHECK(cudaEventCreate(&ev));
        e.prof_events.push_back(ev);
      }
      // GEMM work of the pass: q|k|v, o, gate/up, down per layer (the last layer of a kv-only pass: q|k|v only)
      const double he = c.hidden_e, f = c.mlp_e, rows = M;
      const double w_full = he * e.qkvN + (double)e.qdim * he + 3.0 * he * f, w_kv = he * e.qkvN;
      const double w = (c.layers - (last_kv_only ? 1 : 0)) * w_full + (last_kv_only ? w_kv : 0.0);
      e.prof_flops.push_back(2.0 * rows * w);
      e.prof_bytes.push_back(2.0 * w);  // algorithmic bytes: every expert weight streamed once from HBM
      e.prof_class.push_back(2);
      VL_CUDA_CHECK(cudaEventRecord(e.prof_events[e.prof_used], st));
    }
    const int rc = gemv() ? expert_gemv_pass_launch(p, st) : expert_pass_launch(p, xn.tm, ao.tm, act.tm, x_otm, st);
    if (prof) {
      VL_CUDA_CHECK(cudaEventRecord(e.prof_events[e.prof_used + 1], st));
      e.prof_used += 2;
    }
    return rc;
  }

  // proprio expert block of layer li on the current stream `st` (its own attention launch: one query row per sample over
  // the prefix K/V of the layer and its own K/V row)
  int proprio_layer(int li) {
    const vlaser_cfg& c = e.c;
    VL_TRY(block_pre(e.exp[li], e.xp, e.xpn, e.qp, B, nullptr, c.hidden_e, e.pb.prop_pos, e.pb.prop_kvrow, li));
    if (li == c.layers - 1) return 0;
    AttnParams ap;
    const size_t off = (size_t)li * c.n_kv_heads * e.kv_rows * 128;
    ap.q = e.qp;
    ap.q_row_stride = e.qdim;
    ap.k = e.kc + off;
    ap.v = e.vc + off;
    ap.kv_row_stride = 128;
    ap.kv_head_stride = e.kv_rows * 128;
    ap.out = e.aop.p;
    ap.out_row_stride = e.qdim;
    ap.q_start = e.pb.prop_q_start;
    ap.q_len = e.pb.prop_q_len;
    ap.kv_start = e.pb.kvoff;
    ap.kv_len = e.pb.prop_kv_len;
    ap.max_q_len = 1;
    ap.group = c.n_heads / c.n_kv_heads;
    ap.n_kv_heads = c.n_kv_heads;
    ap.batch = B;
    ap.scale = 0.08838834764831845f;
    ap.trace = tr("attention", li);
    VL_TRY(attention_launch(ap, 128, st));
    e.launches++;
    return block_post(e.exp[li], e.xp, e.xpn, e.aop, e.actp, B, nullptr, c.hidden_e, e.mlp_e_p, 32, &e.xp_otm);
  }

  int prefill() {
    const vlaser_cfg& c = e.c;
    const int cap = B * e.S;
    // option "prop_stream": the proprio token never feeds the prefix stream, it only reads the prefix K/V of each layer.
    // Its 28 x 7 latency-bound launches then run on a second stream (forked here, joined before the first denoise
    // step) and fill the tails of the LLM launches instead of sitting between them.
    // (B = 1: the denoise passes run on the single-observation kernel, the proprio pass still takes the second stream --
    // with one observation the LLM launches leave most SMs idle, so its chain hides completely)
    const bool side = e.prop_stream && !e.mega && !e.fuse_norm;
    cudaStream_t main_st = st;
    if (side) {
      VL_CUDA_CHECK(cudaEventRecord(e.side_events[c.layers], main_st));        // fork: xp (proprio encoder) is ready
      VL_CUDA_CHECK(cudaStreamWaitEvent(e.side_stream, e.side_events[c.layers], 0));
    }
    // RMSNorm fused across the GEMM chain of the prefix stream: only when all four projections run on the CTA-pair kernel
    // through its TMA / transposed epilogues (the o projection has the fewest tiles)
    llm_fused = e.fuse_norm_llm && !e.simt && e.pair_gemm && e.tma_epi && !abl && c.hidden % 256 == 0 &&
                (long long)((cap + 255) / 256) * (c.hidden / 256) >= e.pair_min_tiles;
    llm_fused_in = false;
    for (int li = 0; li < c.layers; ++li) {
      VL_TRY(block_pre(e.llm[li], e.xv, e.xn, e.q, cap, e.pb.m_tot, c.hidden, e.pb.row_pos, e.pb.row_kvrow, li));
      if (side) {
        VL_CUDA_CHECK(cudaEventRecord(e.side_events[li], main_st));            // prefix K/V of layer li written
        VL_CUDA_CHECK(cudaStreamWaitEvent(e.side_stream, e.side_events[li], 0));
        st = e.side_stream;
        const int rc = proprio_layer(li);
        st = main_st;
        if (rc) return rc;
      } else if (mega()) {
        // the proprio token never feeds the prefix stream: its whole expert pass runs after the LLM prefill
      } else if (li == 0 || !e.fuse_norm)  // (fused: the proprio encoder output has no producer epilogue -> plain RMSNorm)
        VL_TRY(block_pre(e.exp[li], e.xp, e.xpn, e.qp, B, nullptr, c.hidden_e, e.pb.prop_pos, e.pb.prop_kvrow, li));
      else
        VL_TRY(exp_pre(li, e.xpn, e.qp, B, e.pb.prop_pos, e.pb.prop_kvrow, 0));
      if (li == c.layers - 1) break;  // last layer: only K/V are needed (joint_model.py:192-193, 690-691)
      AttnParams ap;
      const size_t off = (size_t)li * c.n_kv_heads * e.kv_rows * 128;
      ap.q = e.q;
      ap.q_row_stride = e.qdim;
      ap.q2 = (mega() || side) ? nullptr : e.qp;
      ap.q2_row_stride = e.qdim;
      ap.k = e.kc + off;
      ap.v = e.vc + off;
      ap.kv_row_stride = 128;
      ap.kv_head_stride = e.kv_rows * 128;
      ap.out = e.ao.p;
      ap.out_row_stride = e.qdim;
      ap.out2 = e.aop.p;
      ap.out2_row_stride = e.qdim;
      ap.q_start = e.pb.cu;
      ap.q_len = e.pb.valid_len;
      ap.kv_start = e.pb.kvoff;
      ap.kv_len = e.pb.pre_kv_len;
      ap.max_q_len = e.S + ((mega() || side) ? 0 : 1);
      ap.group = c.n_heads / c.n_kv_heads;
      ap.n_kv_heads = c.n_kv_heads;
      ap.batch = B;
      ap.scale = 0.08838834764831845f;  // 128^-0.5
      const bool tc_ok = e.prefill_tc && ap.q2 == nullptr && !e.simt && ap.group <= 8;
      if (tc_ok) {
        VL_TRY(prefill_attention_tc_launch(e.q3d_tm, e.kc_tm, e.vc_tm, e.pb.cu, e.pb.valid_len, e.pb.kvoff, e.ao.p, e.qdim,
                                           (long long)li * c.n_kv_heads * e.kv_rows, e.kv_rows, B, c.n_kv_heads, ap.group,
                                           ap.scale, st, tr("attention_tc", li)));
        e.launches++;
        ap.min_kv_len = 257;  // longer prefixes: generic kernel (every CTA of a short sample exits at once)
        if (e.S > 256 && e.prefill_tc_long) {  // ... or the two-pass tcgen05 kernel
          VL_TRY(prefill_attention_tc_long_launch(e.q3d_tm, e.kc_tm, e.vc_tm, e.pb.cu, e.pb.valid_len, e.pb.kvoff, e.ao.p, e.qdim,
                                                  (long long)li * c.n_kv_heads * e.kv_rows, e.kv_rows, B, c.n_kv_heads, ap.group,
                                                  e.S, 257, 0, ap.scale, st));
          e.launches++;
        }
      }
      if (!tc_ok || (e.S > 256 && !e.prefill_tc_long)) {
        ap.trace = tr("attention", li);
        VL_TRY(attention_launch(ap, 128, st));
        e.launches++;
      }
      VL_TRY(block_post(e.llm[li], e.xv, e.xn, e.ao, e.act, cap, e.pb.m_tot, c.hidden, e.mlp_p, 1, &e.xv_otm, nullptr,
                        llm_fused ? e.llm[li + 1].ln_in : nullptr));
      if (mega() || side) continue;
      if (e.fuse_norm) VL_TRY(exp_post(li, e.xp, e.xpn, e.aop, e.actp, B, 0, false));
      else VL_TRY(block_post(e.exp[li], e.xp, e.xpn, e.aop, e.actp, B, nullptr, c.hidden_e, e.mlp_e_p, 32, &e.xp_otm));
    }
    if (side) {
      st = e.side_stream;
      const int frc = flush();
      st = main_st;
      if (frc) return frc;
      VL_CUDA_CHECK(cudaEventRecord(e.side_events[c.layers], e.side_stream));  // join
      VL_CUDA_CHECK(cudaStreamWaitEvent(main_st, e.side_events[c.layers], 0));
    }
    VL_TRY(flush());
    if (mega() && !side)
      VL_TRY(expert_pass(0, B, e.xp, e.xp_otm, e.xpn, e.qp, e.aop, e.actp, e.pb.prop_pos, e.pb.prop_kvrow, e.pb.prop_q_start,
                         e.pb.prop_q_len, e.pb.prop_kv_len, 1, true));
    return 0;
  }

  // ---- denoise steps. The flow integration of one observation never reads another one's state, so the batch may be cut
  // into `den_chains` sub-batches whose 10-step chains run side by side on their own streams inside the captured graph
  // (option "den_chains"): every launch of a chain is a few microseconds of dependent latency on a fraction of the SMs,
  // and a second chain fills the SMs and the bubbles the first one leaves. A chain owns a row slice of the action-stream
  // buffers (tensor-map views), its own split-K workspace and its own stream; nothing is shared but the weights.
  struct Chain {
    int id = 0;               // 0: main stream, 1: side stream
    int b0 = 0, nb = 0;       // samples [b0, b0 + nb)
    int r0 = 0, M = 0;        // action rows [r0, r0 + M)
    Act xan, aoa, acta, e1, h2;
    CUtensorMap xa_otm;
  };
  int make_chain(Chain& ch, int id, int b0, int nb) {
    ch.id = id;
    ch.b0 = b0;
    ch.nb = nb;
    ch.r0 = b0 * e.H;
    ch.M = nb * e.H;
    if (b0 == 0 && nb == B) {  // the whole batch: the handle's own maps
      ch.xan = e.xan, ch.aoa = e.aoa, ch.acta = e.acta, ch.e1 = e.e1, ch.h2 = e.h2;
      ch.xa_otm = e.xa_otm;
      return 0;
    }
    auto view = [&](const Act& a, Act& v) -> int {
      v.rows = ch.M;
      v.K = a.K;
      v.p = a.p + (size_t)ch.r0 * a.K;
      // the slice ends where the next chain's rows begin: the maps clip there (loads zero-fill, stores are dropped)
      if (make_tmap_out(&v.otm, v.p, ch.M, a.K, 2)) return -1;
      return make_tmap_bf16(&v.tm, v.p, ch.M, a.K, 128);
    };
    VL_TRY(view(e.xan, ch.xan));
    VL_TRY(view(e.aoa, ch.aoa));
    VL_TRY(view(e.acta, ch.acta));
    VL_TRY(view(e.e1, ch.e1));
    VL_TRY(view(e.h2, ch.h2));
    return make_tmap_out(&ch.xa_otm, e.xa + (size_t)ch.r0 * e.c.hidden_e, ch.M, e.c.hidden_e, 4);
  }

  bool flow = false;  // the pass is the flow-matching training forward (per-sample time bias, velocity out)
  int nch_active = 1;  // sub-batch chains of the denoise steps running side by side (fused tails need the whole GPU)
  int denoise(int step, Chain& ch) {
    const vlaser_cfg& c = e.c;
    const int M = ch.M, he = c.hidden_e;
    float* xa = e.xa + (size_t)ch.r0 * he;
    bf16* qa = e.qa + (size_t)ch.r0 * e.qdim;
    float* action = e.action + (size_t)ch.r0 * c.action_dim;
    const int* act_pos = e.pb.act_pos + ch.r0;
    const int* act_kvrow = e.pb.act_kvrow + ch.r0;
    struct AblScope {
      int& a;
      ~AblScope() { a = 0; }
    } abl_scope{abl};
    abl = e.ablate;
    // action encoder (modules.py:39-53) with the time term pre-reduced into a per-step bias
    VL_TRY(small_linear_launch(action, DT_F32, e.ae_w1, e.ae_b1, ch.e1.p, DT_BF16, M, c.action_dim, he, st));
    e.launches++;
    EpiParams e2;
    e2.out = ch.h2.p;
    e2.ldo = he;
    e2.bias = e.time_bias + (size_t)step * he;
    tr_name = "act_enc2";
    tr_layer = -1;
    if (flow) {
      // per-sample time: linear_2 leaves its fp32 pre-activation, the per-sample time bias and the swish follow
      e2.out = e.flow_z + (size_t)ch.r0 * he;
      e2.bias = nullptr;
      VL_TRY(gemm(ch.e1, e.ae_w2b, M, nullptr, EPI_STORE_F32, e2));
      VL_TRY(bias_silu_rows_launch(e.flow_z + (size_t)ch.r0 * he, e.flow_tbias + (size_t)ch.b0 * he, ch.h2.p, M, e.H, he, st));
      e.launches++;
    } else
    VL_TRY(gemm(ch.e1, e.ae_w2b, M, nullptr, EPI_SILU_BF16, e2));
    EpiParams e3;
    e3.out = xa;
    e3.ldo = he;
    e3.bias = e.ae_b3;
    if (e.fuse_norm && !mega()) norm_producer(e3, ch.xan, e.exp[0].ln_in, step + 1, 0);
    tr_name = "act_enc3";
    VL_TRY(gemm(ch.h2, e.ae_w3, M, nullptr, EPI_STORE_F32, e3));
    if (mega())
      VL_TRY(expert_pass(step + 1, M, e.xa, e.xa_otm, e.xan, e.qa, e.aoa, e.acta, e.pb.act_pos, e.pb.act_kvrow, e.pb.den_q_start,
                         e.pb.den_q_len, e.pb.den_kv_len, e.H, false));
    for (int li = 0; li < (mega() ? 0 : c.layers); ++li) {
      // HBM -> L2 ahead of the chain (option "l2_prefetch"): the ln_in norm pulls this layer's o and gate/up matrices,
      // the ln_post norm its down matrix and the next layer's q|k|v
      auto wbytes = [](const Lin& w) { return (unsigned long long)w.N * w.K * sizeof(bf16); };
      L2Prefetch pf_in, pf_post;
      pf_in.ptr[0] = e.exp[li].o.w, pf_in.bytes[0] = wbytes(e.exp[li].o);
      pf_in.ptr[1] = e.exp[li].gu.w, pf_in.bytes[1] = wbytes(e.exp[li].gu);
      pf_post.ptr[0] = e.exp[li].down.w, pf_post.bytes[0] = wbytes(e.exp[li].down);
      if (li + 1 < c.layers) pf_post.ptr[1] = e.exp[li + 1].qkv.w, pf_post.bytes[1] = wbytes(e.exp[li + 1].qkv);
      if (e.fuse_norm) VL_TRY(exp_pre(li, ch.xan, qa, M, act_pos, act_kvrow, step + 1));
      else VL_TRY(block_pre(e.exp[li], xa, ch.xan, qa, M, nullptr, he, act_pos, act_kvrow, li, e.l2pf ? &pf_in : nullptr));
      AttnParams ap;
      const size_t off = (size_t)li * c.n_kv_heads * e.kv_rows * 128;
      ap.q = e.qa;  // q / out rows are addressed through the absolute per-sample starts below
      ap.q_row_stride = e.qdim;
      ap.k = e.kc + off;
      ap.v = e.vc + off;
      ap.kv_row_stride = 128;
      ap.kv_head_stride = e.kv_rows * 128;
      ap.out = e.aoa.p;
      ap.out_row_stride = e.qdim;
      ap.q_start = e.pb.den_q_start + ch.b0;
      ap.q_len = e.pb.den_q_len + ch.b0;
      ap.kv_start = e.pb.kvoff + ch.b0;
      ap.kv_len = e.pb.den_kv_len + ch.b0;
      ap.stable_prefix = 1;  // prefix + proprio K/V of this layer date from the prefill
      ap.small_rows = e.attn_small;
      ap.max_q_len = e.H;
      ap.group = c.n_heads / c.n_kv_heads;
      ap.n_kv_heads = c.n_kv_heads;
      ap.batch = ch.nb;
      ap.scale = 0.08838834764831845f;
      if (!(abl & 4)) {
        ap.trace = tr("attention", li);
        VL_TRY(attention_launch(ap, 128, st));
        e.launches++;
      }
      if (e.fuse_norm) VL_TRY(exp_post(li, xa, ch.xan, ch.aoa, ch.acta, M, step + 1, false));
      else VL_TRY(block_post(e.exp[li], xa, ch.xan, ch.aoa, ch.acta, M, nullptr, he, e.mlp_e_p, 32, &ch.xa_otm, e.l2pf ? &pf_post : nullptr,
                             (li + 1 < c.layers && nch_active == 1) ? e.exp[li + 1].ln_in : nullptr));
    }
    const int last = step == c.num_inference_steps - 1;
    VL_TRY(flush());  // split-K partials of the last down projection
    const size_t out_off = (size_t)ch.b0 * c.horizon_steps * c.action_dim * (io == DT_F32 ? 4 : 2);
    if (flow) {  // the velocity itself: vel = 0 + 1 * (decoder(norm(x)))
      VL_TRY(euler_step_launch(xa, e.final_norm, e.dec_w, e.dec_b, e.flow_vel + (size_t)ch.r0 * c.action_dim, e.out_act, io, M, he,
                               c.action_dim, e.H, c.horizon_steps, 1.0f, c.rms_eps, 0, 0, 0.f, st, nullptr));
      e.launches++;
      return 0;
    }
    VL_TRY(euler_step_launch(xa, e.final_norm, e.dec_w, e.dec_b, action, reinterpret_cast<uint8_t*>(e.out_act) + out_off, io, M, he,
                             c.action_dim, e.H, c.horizon_steps, (float)(1.0 / c.num_inference_steps), c.rms_eps, last,
                             c.clip_enabled, c.final_action_clip_value, st,
                             (gemv() && e.det_splitk) ? e.xfix + (size_t)(step + 1) * e.xfix_words() : nullptr));
    e.launches++;
    return 0;
  }

  // all denoise steps: one chain, or `den_chains` sub-batch chains forked from / joined into the main stream
  int denoise_all() {
    const vlaser_cfg& c = e.c;
    const int steps = e.stop_after == 4 ? 1 : c.num_inference_steps;
    int nch = 1;
    // two chains when asked for and the halves are whole 128-row tiles of whole samples (tensor-map views, one M tile
    // each); the expert-stream / single-observation / fused-norm schedules keep one chain
    if (e.den_chains >= 2 && !mega() && !e.fuse_norm && !e.simt && B % 2 == 0 && ((B / 2) * e.H) % 128 == 0) nch = 2;
    cudaStream_t main_st = st;
    nch_active = nch;
    std::vector<Chain> chains(nch);
    for (int i = 0; i < nch; ++i) VL_TRY(make_chain(chains[i], i, i * (B / nch), B / nch));
    if (nch > 1) {
      VL_CUDA_CHECK(cudaEventRecord(e.side_events[0], main_st));  // fork: prefill complete, noise converted
      VL_CUDA_CHECK(cudaStreamWaitEvent(e.side_stream, e.side_events[0], 0));
    }
    int rc = 0;
    for (int i = 0; i < nch && rc == 0; ++i) {
      st = i == 0 ? main_st : e.side_stream;
      for (int k = 0; k < steps && rc == 0; ++k) rc = denoise(k, chains[i]);
    }
    st = main_st;
    if (rc) return rc;
    if (nch > 1) {
      VL_CUDA_CHECK(cudaEventRecord(e.side_events[1], e.side_stream));  // join
      VL_CUDA_CHECK(cudaStreamWaitEvent(main_st, e.side_events[1], 0));
    }
    return 0;
  }

  // Flow-matching loss of the training step, forward only (pizero_internvl.py:1064-1197): the joint forward over
  // (prefix, proprio, noisy action chunk) under the block mask equals the prefill followed by ONE pass of the action expert
  // over the cached prefix / proprio K/V -- no block attends to a later one -- at the per-sample time t, then
  // v_psi = decoder(norm(x)) and mean((v_psi - (x1 - (1 - sig_min) x0))^2).
  int flow_loss() {
    const vlaser_cfg& c = e.c;
    const int HA = e.H * c.action_dim;
    VL_TRY(flow_prepare_launch(e.in_noise, e.in_x1, e.in_t, io, e.flow_sig_min, c.time_max_period, e.ae_w2a, e.ae_b2, e.action,
                               e.flow_tbias, e.flow_vel, B, HA, c.hidden_e, st));
    e.launches++;
    Chain ch;
    VL_TRY(make_chain(ch, 0, 0, B));
    flow = true;
    set_pdl_enabled(true);
    const int rc = denoise(0, ch);
    set_pdl_enabled(false);
    flow = false;
    if (rc) return rc;
    VL_TRY(flow_loss_launch(e.flow_vel, e.in_noise, e.in_x1, io, e.flow_sig_min, (long long)B * HA, e.flow_loss, st));
    e.launches++;
    return 0;
  }

  // ---- greedy text generation of the chat VLM (modeling_internvl_chat.py:401-439 -> HF Qwen2 generate, do_sample=False;
  // PiZero.infer_text, pizero_internvl.py:1005-1046): the LLM mixture alone, causal, RoPE positions from 0, every layer
  // complete (no last-layer skip), final RMSNorm, output head, argmax.
  // final norm of the B residual rows in x_dec -> logits -> next_tok
  int text_head(int bump) {
    const vlaser_cfg& c = e.c;
    SplitKPartials spt;
    VL_TRY(rmsnorm_launch(e.x_dec, e.llm_norm, e.xn.p, B, nullptr, c.hidden, c.rms_eps, st, nullptr, nullptr, take(e.x_dec, spt)));
    e.launches++;
    EpiParams ep;
    ep.out = e.logits;
    ep.ldo = e.lm_head.N;
    tr_name = "lm_head";
    tr_layer = -1;
    VL_TRY(gemm(e.xn, e.lm_head, B, nullptr, EPI_STORE_F32, ep));
    VL_TRY(argmax_rows_launch(e.logits, e.lm_head.N, c.embed_rows, B, e.gen.next_tok, e.gen.n_gen, bump, st));
    e.launches++;
    return 0;
  }
  AttnParams text_attn(int li) {
    const vlaser_cfg& c = e.c;
    AttnParams ap;
    const size_t off = (size_t)li * c.n_kv_heads * e.kv_rows * 128;
    ap.q = e.q;
    ap.q_row_stride = e.qdim;
    ap.k = e.kc + off;
    ap.v = e.vc + off;
    ap.kv_row_stride = 128;
    ap.kv_head_stride = e.kv_rows * 128;
    ap.out = e.ao.p;
    ap.out_row_stride = e.qdim;
    ap.kv_start = e.pb.kvoff;
    ap.group = c.n_heads / c.n_kv_heads;
    ap.n_kv_heads = c.n_kv_heads;
    ap.batch = B;
    ap.scale = 0.08838834764831845f;
    ap.causal = 1;
    return ap;
  }
  int text_prefill() {
    const vlaser_cfg& c = e.c;
    const int cap = B * e.S;
    for (int li = 0; li < c.layers; ++li) {
      VL_TRY(block_pre(e.llm[li], e.xv, e.xn, e.q, cap, e.pb.m_tot, c.hidden, e.pb.row_pos, e.pb.row_kvrow, li));
      AttnParams ap = text_attn(li);
      if (e.prefill_tc_long && !e.simt && ap.group <= 8) {
        VL_TRY(prefill_attention_tc_long_launch(e.q3d_tm, e.kc_tm, e.vc_tm, e.pb.cu, e.pb.valid_len, e.pb.kvoff, e.ao.p, e.qdim,
                                                (long long)li * c.n_kv_heads * e.kv_rows, e.kv_rows, B, c.n_kv_heads, ap.group,
                                                e.S, 1, 1, ap.scale, st));
      } else {
        ap.q_start = e.pb.cu;
        ap.q_len = e.pb.valid_len;
        ap.kv_len = e.pb.pre_kv_len;
        ap.max_q_len = e.S;
        VL_TRY(attention_launch(ap, 128, st));
      }
      e.launches++;
      VL_TRY(block_post(e.llm[li], e.xv, e.xn, e.ao, e.act, cap, e.pb.m_tot, c.hidden, e.mlp_p, 1, &e.xv_otm));
    }
    VL_TRY(flush());
    VL_TRY(last_rows_launch(e.xv, e.pb.cu, e.pb.valid_len, e.x_dec, B, c.hidden, st));
    e.launches++;
    return text_head(0);
  }
  // one decode step: emit the pending token, run it through the LLM over its K/V window, leave the next argmax
  int text_decode() {
    const vlaser_cfg& c = e.c;
    VL_TRY(decode_advance_launch(e.gen, e.pb.valid_len, e.pb.kvoff, e.embed, e.x_dec, B, c.hidden, e.gen_cap, e.gen_cap,
                                 e.gen_eos, e.gen_fill, st));
    e.launches++;
    for (int li = 0; li < c.layers; ++li) {
      VL_TRY(block_pre(e.llm[li], e.x_dec, e.xn, e.q, B, nullptr, c.hidden, e.gen.pos, e.gen.kvrow, li));
      AttnParams ap = text_attn(li);
      if (e.dec_chunk > 0 && ap.group <= 8) {
        // one query row per sample over its whole window: split-KV kernel (the generic one walks the window serially)
        DecodeAttnParams dp;
        dp.q = ap.q, dp.q_row_stride = ap.q_row_stride;
        dp.k = ap.k, dp.v = ap.v, dp.kv_row_stride = ap.kv_row_stride, dp.kv_head_stride = ap.kv_head_stride;
        dp.out = ap.out, dp.out_row_stride = ap.out_row_stride;
        dp.q_start = e.gen.q_start, dp.kv_start = e.pb.kvoff, dp.kv_len = e.gen.kv_len;
        dp.group = ap.group, dp.n_kv_heads = ap.n_kv_heads, dp.batch = B;
        dp.max_kv_len = e.gen_cap;
        dp.chunk = e.dec_chunk;
        dp.scale = ap.scale;
        dp.ws_o = e.dec_ws_o, dp.ws_ml = e.dec_ws_ml;
        VL_TRY(decode_attention_launch(dp, st));
        e.launches += 2;
      } else {
        ap.q_start = e.gen.q_start;
        ap.q_len = e.gen.q_len;
        ap.kv_len = e.gen.kv_len;
        ap.max_q_len = 1;
        VL_TRY(attention_launch(ap, 128, st));
        e.launches++;
      }
      VL_TRY(block_post(e.llm[li], e.x_dec, e.xn, e.ao, e.act, B, nullptr, c.hidden, e.mlp_p, 1, &e.xdec_otm));
    }
    return text_head(1);
  }

  int all(bool has_valid) {
    const vlaser_cfg& c = e.c;
    VL_CUDA_CHECK(cudaMemsetAsync(e.tail_ctrs, 0, sizeof(unsigned int) * Engine::TAIL_CTRS, st));  // fused-tail barriers
    if (e.ln_overlap) VL_CUDA_CHECK(cudaMemsetAsync(e.ln_ctr, 0, sizeof(int) * (size_t)2 * c.vit_layers * e.ln_ctr_blocks, st));
    e.tail_used = 0;
    if (e.mode == 3) return text_decode();
    if (e.mode == 2) {
      VL_TRY(plan_launch(e.in_ids, has_valid ? e.in_valid : nullptr, e.pb, B, e.S, e.H, e.n_img * e.T, c.image_token_index,
                         c.pad_token_id, c.embed_rows, st, e.gen_cap));
      VL_CUDA_CHECK(cudaMemsetAsync(e.gen.n_gen, 0, sizeof(int), st));
      VL_CUDA_CHECK(cudaMemsetAsync(e.gen.finished, 0, sizeof(int) * (size_t)B, st));
      e.launches += 5;
      if (!e.text_no_image) VL_TRY(vit());
      VL_TRY(embed());
      return text_prefill();
    }
    VL_TRY(plan_launch(e.in_ids, has_valid ? e.in_valid : nullptr, e.pb, B, e.S, e.H, e.n_img * e.T,
                       c.image_token_index, c.pad_token_id, c.embed_rows, st));
    VL_TRY(convert_launch(e.in_noise, io, e.action, DT_F32, (long long)B * e.H * c.action_dim, st));
    if (e.fuse_norm) VL_CUDA_CHECK(cudaMemsetAsync(e.ss, 0, e.ss_floats * sizeof(float), st));
    VL_CUDA_CHECK(cudaMemsetAsync(e.xp_bar, 0, sizeof(unsigned int) * (size_t)(c.num_inference_steps + 1) * e.xp_bar_words(), st));

    if (gemv() && e.det_splitk)
      VL_CUDA_CHECK(cudaMemsetAsync(e.xfix, 0, sizeof(long long) * (size_t)(c.num_inference_steps + 1) * e.xfix_words(), st));
    e.launches += 5;
    VL_TRY(vit());
    if (e.stop_after == 1) return 0;
    VL_TRY(embed());
    if (e.stop_after == 2) return 0;
    VL_TRY(prefill());
    if (e.stop_after == 3) return 0;
    if (e.mode == 1) return flow_loss();
    set_pdl_enabled(true);
    const int rc = denoise_all();
    set_pdl_enabled(false);
    return rc;
  }
};

int engine_execute(Engine& e, int B, int io, bool has_valid, cudaStream_t st) {
  e.launches = 0;
  if (!e.use_graph) {
    e.trace_idx = 0;
    e.trace_names.clear();
    Run r{e, st, B, io};
    const int rc = r.all(has_valid);
    e.last_launches = e.launches;
    return rc;
  }
  Engine::GraphKey key{B, io, e.stop_after, e.simt, has_valid ? 1 : 0, e.mode, e.mode >= 2 ? e.text_no_image : 0};
  auto it = e.graphs.find(key);
  if (it == e.graphs.end()) {
    // Build on the handle's own stream (the caller's may be the legacy default stream, which cannot
    // be captured). A warm, non-captured pass first sets the per-kernel attributes outside capture.
    VL_CUDA_CHECK(cudaStreamSynchronize(st));  // staged inputs have landed
    cudaStream_t cs = e.own_stream;
    e.trace_idx = 0;
    e.trace_names.clear();
    {
      Run r{e, cs, B, io};
      VL_TRY(r.all(has_valid));
      VL_CUDA_CHECK(cudaStreamSynchronize(cs));
    }
    e.launches = 0;
    e.trace_idx = 0;
    e.trace_names.clear();
    cudaGraph_t graph = nullptr;
    VL_CUDA_CHECK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    Run r{e, cs, B, io};
    const int rc = r.all(has_valid);
    cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    if (rc != 0) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    VL_CUDA_CHECK(ce);
    cudaGraphExec_t exec = nullptr;
    VL_CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
    cudaGraphDestroy(graph);
    it = e.graphs.emplace(key, Engine::GraphVal{exec, e.launches}).first;
    if (e.mode == 3) {  // a decode step changes state: the un-captured pass above WAS this step
      e.last_launches = it->second.launches;
      return 0;
    }
  }
  VL_CUDA_CHECK(cudaGraphLaunch(it->second.exec, st));
  e.last_launches = it->second.launches;
  return 0;
}

size_t io_size(int io) { return io == DT_F32 ? 4 : 2; }

}  // namespace
}  // namespace vlaser

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace vlaser;

struct vlaser_handle {
  Engine e;
};

extern "C" {

int32_t vlaser_abi_version(void) { return VLASER_ABI_VERSION; }
const char* vlaser_last_error(void) { return g_last_error.c_str(); }

int32_t vlaser_create(const vlaser_cfg* cfg, vlaser_t** out) {
  VL_CHECK(cfg && out, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_last_error("vlaser_create: no CUDA device (this library has no CPU fallback)");
    return -10;
  }
  auto* h = new vlaser_handle();
  h->e.c = *cfg;
  const int rc = engine_build(h->e);
  if (rc != 0) {
    delete h;
    return rc;
  }
  *out = h;
  return 0;
}

int32_t vlaser_set_weight(vlaser_t* h, const char* key, const void* data, const int64_t* shape, int32_t ndim,
                          int32_t dtype) {
  VL_CHECK(h && key && data && shape, "null argument");
  return engine_set_weight(h->e, key, data, shape, ndim, dtype);
}

int32_t vlaser_finalize(vlaser_t* h) {
  VL_CHECK(h, "null handle");
  return engine_finalize(h->e);
}

int32_t vlaser_set_option(vlaser_t* h, const char* key, int64_t value) {
  VL_CHECK(h && key, "null argument");
  const std::string k(key);
  if (k == "use_graph") h->e.use_graph = (int)value;
  else if (k == "stop_after") h->e.stop_after = (int)value;
  else if (k == "simt_gemm") h->e.simt = (int)value;
  else if (k == "pdl" || k == "exp_stages" || k == "fuse_norm" || k == "pair_gemm" || k == "tma_epi" ||
           k == "pair_min_tiles" || k == "expert_mega" || k == "expert_streams" || k == "expert_gemv" || k == "l2_prefetch" || k == "small_m_split" || k == "prop_stream" || k == "expert_timing" || k == "expert_flags" || k == "ablate" || k == "trace" || k == "det_splitk" || k == "vit_tc" || k == "exp_splits_o" || k == "exp_splits_down" || k == "den_chains" || k == "pair_bn128" || k == "gelu_warps" || k == "prefill_tc" || k == "fuse_tail" || k == "decode_chunk" || k == "prefill_tc_long" || k == "attn_small" || k == "rn_small" || k == "known_splits" || k == "half_sm" || k == "fuse_norm_llm" || k == "ln_overlap") {
    // schedule-shaping options: captured graphs are stale afterwards
    if (k == "pdl") g_pdl_mode = (int)value;
    else if (k == "fuse_norm") h->e.fuse_norm = (int)value;
    else if (k == "expert_mega") h->e.mega = (int)value;
    else if (k == "expert_streams") h->e.xp_streams = (int)value;
    else if (k == "expert_gemv") h->e.gemv = (int)value;
    else if (k == "l2_prefetch") h->e.l2pf = (int)value;
    else if (k == "small_m_split") h->e.small_m_split = (int)value;
    else if (k == "prop_stream") h->e.prop_stream = (int)value;
    else if (k == "expert_timing") h->e.xp_time = (int)value;
    else if (k == "expert_flags") h->e.xp_flags = (int)value;
    else if (k == "ablate") h->e.ablate = (int)value;
    else if (k == "trace") h->e.trace_on = (int)value;
    else if (k == "det_splitk") h->e.det_splitk = (int)value;
    else if (k == "vit_tc") h->e.vit_tc = (int)value;
    else if (k == "den_chains") h->e.den_chains = (int)value;
    else if (k == "pair_bn128") h->e.pair_bn128 = (int)value;
    else if (k == "gelu_warps") h->e.gelu_warps = (int)value == 8 ? 8 : 16;
    else if (k == "prefill_tc") h->e.prefill_tc = (int)value;
    else if (k == "fuse_tail") h->e.fuse_tail = (int)value;
    else if (k == "prefill_tc_long") h->e.prefill_tc_long = (int)value;
    else if (k == "attn_small") h->e.attn_small = (int)value;
    else if (k == "known_splits") h->e.known_splits = (int)value;
    else if (k == "half_sm") h->e.half_sm = (int)value;
    else if (k == "fuse_norm_llm") h->e.fuse_norm_llm = (int)value;
    else if (k == "ln_overlap") h->e.ln_overlap = (int)value;
    else if (k == "rn_small") set_resid_norm_small((int)value);
    else if (k == "decode_chunk") h->e.dec_chunk = (value <= 0) ? 0 : std::min(256, std::max(64, (int)value / 64 * 64));
    else if (k == "exp_splits_o") h->e.xs_o = std::max(1, (int)value);
    else if (k == "exp_splits_down") h->e.xs_down = std::max(1, (int)value);
    else if (k == "pair_gemm") h->e.pair_gemm = (int)value;
    else if (k == "tma_epi") h->e.tma_epi = (int)value;
    else if (k == "pair_min_tiles") h->e.pair_min_tiles = (int)value;
    else h->e.exp_stages = (int)value;
    cudaDeviceSynchronize();
    for (auto& g : h->e.graphs) cudaGraphExecDestroy(g.second.exec);
    h->e.graphs.clear();
  }
  else if (k == "profile_gemm") {
    h->e.profile_gemm = (int)value;
    h->e.prof_used = 0;
    h->e.prof_flops.clear();
    h->e.prof_bytes.clear();
    h->e.prof_class.clear();
  } else if (k == "profile_mtot") h->e.profile_mtot = value;
  else {
    set_last_error("unknown option " + k);
    return -2;
  }
  return 0;
}

int32_t vlaser_infer_action(vlaser_t* h, const int64_t* input_ids, const void* pixel_values, const int32_t* valid_len,
                            const void* proprios, const void* noise, void* actions_out, int32_t B, int32_t io_dtype,
                            void* stream) {
  VL_CHECK(h, "null handle");
  Engine& e = h->e;
  VL_CHECK(e.finalized, "infer_action before finalize");
  VL_CHECK(B > 0 && B <= e.Bc, "batch exceeds max_batch");
  VL_CHECK(io_dtype == DT_F32 || io_dtype == DT_BF16, "io_dtype must be f32 or bf16");
  VL_CHECK(input_ids && pixel_values && proprios && noise && actions_out, "null tensor");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const vlaser_cfg& c = e.c;
  const size_t es = io_size(io_dtype);
  const size_t pix = (size_t)B * e.n_img * 3 * c.image_size * c.image_size;
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_ids, input_ids, (size_t)B * e.S * 8, cudaMemcpyDeviceToDevice, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_pix, pixel_values, pix * es, cudaMemcpyDeviceToDevice, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_prop, proprios, (size_t)B * c.proprio_dim * es, cudaMemcpyDeviceToDevice, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_noise, noise, (size_t)B * e.H * c.action_dim * es, cudaMemcpyDeviceToDevice, st));
  if (valid_len) VL_CUDA_CHECK(cudaMemcpyAsync(e.in_valid, valid_len, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  // token ids outside the embedding table: the device call cannot synchronise, so the flag of the PREVIOUS call (copied to
  // pinned memory behind it) is reported here, at the latest one call late
  int* host_bad = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(e.pinned) + e.pinned_bytes - 64);
  if (*host_bad != 0) {
    const int bad = *host_bad - 1;
    *host_bad = 0;
    cudaMemsetAsync(e.pb.bad_id, 0, sizeof(int), st);
    if (bad + 1 == NORM_ERR_READY_TIMEOUT) {
      set_last_error("infer_action: a norm launch waited ~1 s for a row block its residual GEMM never reported (option ln_overlap) "
                     "in an earlier call; the actions of that call are invalid");
      return -5;
    }
    if (bad + 1 == EXPERT_ERR_BARRIER) {
      set_last_error("infer_action: a grid barrier of the persistent expert kernel timed out in an earlier call (a CTA of its grid "
                     "was not resident); the actions of that call are invalid");
      return -5;
    }
    set_last_error("infer_action: an earlier call passed token id " + std::to_string(bad) + " outside the embedding table of " +
                   std::to_string(c.embed_rows) + " rows (nn.Embedding would raise IndexError)");
    return -4;
  }
  VL_TRY(engine_execute(e, B, io_dtype, valid_len != nullptr, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(actions_out, e.out_act, (size_t)B * c.horizon_steps * c.action_dim * es,
                                cudaMemcpyDeviceToDevice, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(host_bad, e.pb.bad_id, sizeof(int), cudaMemcpyDeviceToHost, st));
  return 0;
}

int32_t vlaser_flow_matching_loss(vlaser_t* h, const int64_t* input_ids, const void* pixel_values, const int32_t* valid_len,
                                  const void* proprios, const void* actions, const void* noise, const void* t,
                                  float flow_sig_min, float* loss_out, float* velocity_out, int32_t B, int32_t io_dtype,
                                  void* stream) {
  VL_CHECK(h, "null handle");
  Engine& e = h->e;
  VL_CHECK(e.finalized, "flow_matching_loss before finalize");
  VL_CHECK(B > 0 && B <= e.Bc, "batch exceeds max_batch");
  VL_CHECK(io_dtype == DT_F32 || io_dtype == DT_BF16, "io_dtype must be f32 or bf16");
  VL_CHECK(input_ids && pixel_values && proprios && actions && noise && t && loss_out, "null tensor");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const vlaser_cfg& c = e.c;
  const size_t es = io_size(io_dtype);
  const size_t pix = (size_t)B * e.n_img * 3 * c.image_size * c.image_size;
  const size_t na = (size_t)B * e.H * c.action_dim;
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_ids, input_ids, (size_t)B * e.S * 8, cudaMemcpyDeviceToDevice, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_pix, pixel_values, pix * es, cudaMemcpyDeviceToDevice, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_prop, proprios, (size_t)B * c.proprio_dim * es, cudaMemcpyDeviceToDevice, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_noise, noise, na * es, cudaMemcpyDeviceToDevice, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_x1, actions, na * es, cudaMemcpyDeviceToDevice, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_t, t, (size_t)B * es, cudaMemcpyDeviceToDevice, st));
  if (valid_len) VL_CUDA_CHECK(cudaMemcpyAsync(e.in_valid, valid_len, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  e.flow_sig_min = flow_sig_min;
  e.mode = 1;
  // (flow_sig_min is a launch constant of the captured graph: a new value needs a new capture)
  if (e.flow_sig_graph != flow_sig_min) {
    for (auto it = e.graphs.begin(); it != e.graphs.end();)
      if (it->first.mode == 1) {
        cudaStreamSynchronize(e.own_stream);
        cudaGraphExecDestroy(it->second.exec);
        it = e.graphs.erase(it);
      } else {
        ++it;
      }
    e.flow_sig_graph = flow_sig_min;
  }
  const int rc = engine_execute(e, B, io_dtype, valid_len != nullptr, st);
  e.mode = 0;
  if (rc) return rc;
  VL_CUDA_CHECK(cudaMemcpyAsync(loss_out, e.flow_loss, sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (velocity_out) VL_CUDA_CHECK(cudaMemcpyAsync(velocity_out, e.flow_vel, na * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int32_t vlaser_generate(vlaser_t* h, const int64_t* input_ids, const void* pixel_values, const int32_t* valid_len,
                        int32_t max_new_tokens, int64_t eos_token_id, int64_t fill_token_id, int64_t* tokens_out,
                        float* first_logits_out, int32_t B, int32_t io_dtype, void* stream) {
  VL_CHECK(h, "null handle");
  Engine& e = h->e;
  VL_CHECK(e.finalized, "generate before finalize");
  VL_CHECK(e.llm_norm_set, "generate: the final norm of the LLM (joint_model.mixtures.vlm.norm.weight / "
                           "language_model.model.norm.weight) was not loaded");
  VL_CHECK(B > 0 && B <= e.Bc, "batch exceeds max_batch");
  VL_CHECK(io_dtype == DT_F32 || io_dtype == DT_BF16, "io_dtype must be f32 or bf16");
  VL_CHECK(input_ids && tokens_out, "null tensor");
  VL_CHECK(max_new_tokens >= 1 && max_new_tokens <= e.gen_cap, "max_new_tokens must be in [1, max_image_text_tokens + 1 + H]");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const vlaser_cfg& c = e.c;
  const size_t es = io_size(io_dtype);
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_ids, input_ids, (size_t)B * e.S * 8, cudaMemcpyDeviceToDevice, st));
  if (pixel_values)
    VL_CUDA_CHECK(cudaMemcpyAsync(e.in_pix, pixel_values, (size_t)B * e.n_img * 3 * c.image_size * c.image_size * es,
                                  cudaMemcpyDeviceToDevice, st));
  if (valid_len) VL_CUDA_CHECK(cudaMemcpyAsync(e.in_valid, valid_len, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  if (e.gen_eos != eos_token_id || e.gen_fill != fill_token_id) {  // launch constants of the captured decode step
    for (auto it = e.graphs.begin(); it != e.graphs.end();)
      if (it->first.mode == 3) {
        cudaStreamSynchronize(e.own_stream);
        cudaGraphExecDestroy(it->second.exec);
        it = e.graphs.erase(it);
      } else {
        ++it;
      }
    e.gen_eos = eos_token_id;
    e.gen_fill = fill_token_id;
  }
  e.text_no_image = pixel_values ? 0 : 1;
  int rc = 0;
  e.mode = 2;
  rc = engine_execute(e, B, io_dtype, valid_len != nullptr, st);
  if (rc == 0 && first_logits_out)
    rc = cudaMemcpy2DAsync(first_logits_out, (size_t)c.embed_rows * 4, e.logits, (size_t)e.lm_head.N * 4, (size_t)c.embed_rows * 4, B,
                           cudaMemcpyDeviceToDevice, st) == cudaSuccess ? 0 : -1;
  e.mode = 3;
  for (int k = 0; k + 1 < max_new_tokens && rc == 0; ++k) rc = engine_execute(e, B, io_dtype, valid_len != nullptr, st);
  e.mode = 0;
  e.text_no_image = 0;
  if (rc) return rc;
  // the last token: emitted, not run through the model
  VL_TRY(decode_advance_launch(e.gen, e.pb.valid_len, e.pb.kvoff, e.embed, e.x_dec, B, c.hidden, e.gen_cap, e.gen_cap, e.gen_eos,
                               e.gen_fill, st));
  VL_CUDA_CHECK(cudaMemcpy2DAsync(tokens_out, (size_t)max_new_tokens * 8, e.gen.tokens, (size_t)e.gen_cap * 8,
                                  (size_t)max_new_tokens * 8, B, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int32_t vlaser_infer_action_host(vlaser_t* h, const int64_t* input_ids, const void* pixel_values,
                                 const int32_t* valid_len, const void* proprios, const void* noise, void* actions_out,
                                 int32_t B, int32_t io_dtype) {
  VL_CHECK(h, "null handle");
  Engine& e = h->e;
  VL_CHECK(e.finalized, "infer_action before finalize");
  VL_CHECK(B > 0 && B <= e.Bc, "batch exceeds max_batch");
  VL_CHECK(io_dtype == DT_F32 || io_dtype == DT_BF16, "io_dtype must be f32 or bf16");
  VL_CHECK(input_ids && pixel_values && proprios && noise && actions_out, "null tensor");
  cudaStream_t st = e.own_stream;
  const vlaser_cfg& c = e.c;
  const size_t es = io_size(io_dtype);
  const size_t n_ids = (size_t)B * e.S * 8;
  const size_t n_pix = (size_t)B * e.n_img * 3 * c.image_size * c.image_size * es;
  const size_t n_prop = (size_t)B * c.proprio_dim * es;
  const size_t n_noise = (size_t)B * e.H * c.action_dim * es;
  const size_t n_out = (size_t)B * c.horizon_steps * c.action_dim * es;
  // stage through the handle's pinned buffer so the H2D copies are truly asynchronous
  uint8_t* pin = reinterpret_cast<uint8_t*>(e.pinned);
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  uint8_t* p_ids = pin;
  uint8_t* p_pix = p_ids + up(n_ids);
  uint8_t* p_prop = p_pix + up(n_pix);
  uint8_t* p_noise = p_prop + up(n_prop);
  uint8_t* p_valid = p_noise + up(n_noise);
  uint8_t* p_out = p_valid + up((size_t)B * 4);
  VL_CHECK((size_t)(p_out + up(n_out) + 256 - pin) <= e.pinned_bytes, "pinned staging buffer too small");
  // page-locked caller buffers are copied from directly; pageable ones are staged through the pinned buffer
  auto stage = [&](uint8_t* pin_dst, const void* src, size_t n) -> const void* {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost) return src;
    cudaGetLastError();
    memcpy(pin_dst, src, n);
    return pin_dst;
  };
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_ids, stage(p_ids, input_ids, n_ids), n_ids, cudaMemcpyHostToDevice, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_pix, stage(p_pix, pixel_values, n_pix), n_pix, cudaMemcpyHostToDevice, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_prop, stage(p_prop, proprios, n_prop), n_prop, cudaMemcpyHostToDevice, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(e.in_noise, stage(p_noise, noise, n_noise), n_noise, cudaMemcpyHostToDevice, st));
  if (valid_len)
    VL_CUDA_CHECK(cudaMemcpyAsync(e.in_valid, stage(p_valid, valid_len, (size_t)B * 4), (size_t)B * 4,
                                  cudaMemcpyHostToDevice, st));
  VL_TRY(engine_execute(e, B, io_dtype, valid_len != nullptr, st));
  VL_CUDA_CHECK(cudaMemcpyAsync(p_out, e.out_act, n_out, cudaMemcpyDeviceToHost, st));
  int* p_bad = reinterpret_cast<int*>(p_out + up(n_out));
  VL_CUDA_CHECK(cudaMemcpyAsync(p_bad, e.pb.bad_id, sizeof(int), cudaMemcpyDeviceToHost, st));
  VL_CUDA_CHECK(cudaStreamSynchronize(st));
  memcpy(actions_out, p_out, n_out);
  if (*p_bad != 0) {
    const int bad = *p_bad - 1;
    cudaMemsetAsync(e.pb.bad_id, 0, sizeof(int), st);
    if (*p_bad == NORM_ERR_READY_TIMEOUT) {
      set_last_error("infer_action: a norm launch waited ~1 s for a row block its residual GEMM never reported (option ln_overlap); "
                     "the actions are invalid");
      return -5;
    }
    if (*p_bad == EXPERT_ERR_BARRIER) {
      set_last_error("infer_action: a grid barrier of the persistent expert kernel timed out (a CTA of its grid was not resident); "
                     "the actions are invalid");
      return -5;
    }
    set_last_error("infer_action: input_ids holds token id " + std::to_string(bad) + " outside the embedding table of " +
                   std::to_string(c.embed_rows) + " rows (nn.Embedding would raise IndexError)");
    return -4;
  }
  return 0;
}

void vlaser_destroy(vlaser_t* h) {
  if (!h) return;
  cudaDeviceSynchronize();
  delete h;
}

int64_t vlaser_last_launch_count(vlaser_t* h) { return h ? h->e.last_launches : -1; }

int64_t vlaser_debug_copy(vlaser_t* h, const char* name, void* host_dst, int64_t max_bytes) {
  VL_CHECK(h && name, "null argument");
  Engine& e = h->e;
  const vlaser_cfg& c = e.c;
  const std::string n(name);
  const void* src = nullptr;
  size_t bytes = 0;
  const size_t Ni = (size_t)e.Bc * e.n_img;
  const size_t kvb = (size_t)c.layers * c.n_kv_heads * e.kv_rows * 128 * 2;
  if (n == "vit_x") src = e.vx, bytes = Ni * (e.P + 1) * c.vit_hidden * 4;
  else if (n == "prefix_x") src = e.xv, bytes = (size_t)e.Bc * e.S * c.hidden * 4;
  else if (n == "proprio_x") src = e.xp, bytes = (size_t)e.Bc * c.hidden_e * 4;
  else if (n == "action_x") src = e.xa, bytes = (size_t)e.Bc * e.H * c.hidden_e * 4;
  else if (n == "action") src = e.action, bytes = (size_t)e.Bc * e.H * c.action_dim * 4;
  else if (n == "k_cache") src = e.kc, bytes = kvb;
  else if (n == "v_cache") src = e.vc, bytes = kvb;
  else if (n == "valid_len") src = e.pb.valid_len, bytes = (size_t)e.Bc * 4;
  else if (n == "cu") src = e.pb.cu, bytes = (size_t)(e.Bc + 1) * 4;
  else if (n == "kvoff") src = e.pb.kvoff, bytes = (size_t)e.Bc * 4;
  else if (n == "img_row_map") src = e.pb.img_row_map, bytes = Ni * e.T * 4;
  else if (n == "row_src") src = e.pb.row_src, bytes = (size_t)e.Bc * e.S * 4;
  else if (n == "expert_timing") src = e.xp_timing, bytes = (size_t)(c.num_inference_steps + 1) * 7 * c.layers * 8;
  else if (n == "trace") src = e.trace, bytes = Engine::TRACE_MAX * 16 * 8;
  else if (n == "trace_names") {
    // newline-separated launch names in record order (valid after a call that built its graph with option "trace" on)
    if (host_dst && max_bytes > 0) {
      const size_t nb = std::min<size_t>(e.trace_names.size(), (size_t)max_bytes);
      memcpy(host_dst, e.trace_names.data(), nb);
    }
    return (int64_t)e.trace_names.size();
  } else if (n == "kv_rows") {
    const int64_t v = e.kv_rows;
    if (host_dst && max_bytes >= 8) memcpy(host_dst, &v, 8);
    return 8;
  } else {
    set_last_error("debug_copy: unknown buffer " + n);
    return -2;
  }
  if (host_dst && max_bytes > 0) {
    VL_CUDA_CHECK(cudaDeviceSynchronize());
    VL_CUDA_CHECK(cudaMemcpy(host_dst, src, std::min<size_t>(bytes, (size_t)max_bytes), cudaMemcpyDeviceToHost));
  }
  return (int64_t)bytes;
}

int32_t vlaser_profile_read(vlaser_t* h, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches) {
  VL_CHECK(h && gemm_ms && gemm_flops && gemm_launches, "null argument");
  Engine& e = h->e;
  VL_CUDA_CHECK(cudaDeviceSynchronize());
  double ms = 0, fl = 0;
  for (size_t i = 0; i + 1 < e.prof_used; i += 2) {
    float t = 0;
    VL_CUDA_CHECK(cudaEventElapsedTime(&t, e.prof_events[i], e.prof_events[i + 1]));
    ms += t;
    fl += e.prof_flops[i / 2];
  }
  *gemm_ms = ms;
  *gemm_flops = fl;
  *gemm_launches = (int64_t)(e.prof_used / 2);
  e.prof_used = 0;
  e.prof_flops.clear();
  e.prof_bytes.clear();
  e.prof_class.clear();
  return 0;
}

int32_t vlaser_profile_read_classes(vlaser_t* h, double* out12) {
  VL_CHECK(h && out12, "null argument");
  Engine& e = h->e;
  VL_CUDA_CHECK(cudaDeviceSynchronize());
  double* out8 = out12;
  for (int i = 0; i < 12; ++i) out12[i] = 0.0;
  for (size_t i = 0; i + 1 < e.prof_used; i += 2) {
    float t = 0;
    VL_CUDA_CHECK(cudaEventElapsedTime(&t, e.prof_events[i], e.prof_events[i + 1]));
    const int c = std::min(std::max(e.prof_class[i / 2], 0), 2);
    out8[4 * c + 0] += t;
    out8[4 * c + 1] += e.prof_flops[i / 2];
    out8[4 * c + 2] += e.prof_bytes[i / 2];
    out8[4 * c + 3] += 1.0;
  }
  e.prof_used = 0;
  e.prof_flops.clear();
  e.prof_bytes.clear();
  e.prof_class.clear();
  return 0;
}

double vlaser_flops_per_call(vlaser_t* h, const int32_t* valid_len_host, int32_t B) {
  if (!h || !valid_len_host) return -1.0;
  const vlaser_cfg& c = h->e.c;
  const Engine& e = h->e;
  // W(h,f) = h*nh*hd + 2*h*nkv*hd + nh*hd*h + 3*h*f
  auto W = [&](double hid, double f) { return hid * e.qdim + 2 * hid * e.kvdim + e.qdim * hid + 3 * hid * f; };
  const double N = e.P + 1, dv = c.vit_hidden, fv = c.vit_mlp, L = c.layers, hh = c.hidden, he = c.hidden_e;
  const double nhd = e.qdim, H = e.H, A = c.action_dim;
  const double k_patch = 3.0 * c.patch_size * c.patch_size;
  const double F_vit =
      e.n_img * (c.vit_layers * (2 * N * dv * 3 * dv + 2 * N * dv * dv + 4 * N * dv * fv + 4 * N * N * dv) +
                 2 * e.P * k_patch * dv);
  const double F_proj = e.n_img * e.T * (2 * 4 * dv * hh + 2 * hh * hh);
  double total = 0;
  for (int b = 0; b < B; ++b) {
    const double Sx = valid_len_host[b];
    const double F_prefill = L * (2 * Sx * W(hh, c.mlp) + 2 * W(he, c.mlp_e) + 4 * (Sx * Sx + Sx + 1) * nhd) -
                             2 * Sx * (nhd * hh + 3 * hh * c.mlp) - 2 * (nhd * he + 3 * he * c.mlp_e);
    const double F_step = L * (2 * H * W(he, c.mlp_e) + 4 * H * (Sx + 1 + H) * nhd) + 2 * H * (A * he + 3 * he * he) +
                          2 * H * he * A;
    total += F_vit + F_proj + F_prefill + c.num_inference_steps * F_step;
  }
  return total;
}

// ---- single-kernel entry points ---------------------------------------------------------------
int32_t vlaser_op_gemm(const vlaser_gemm_args* a, void* stream) {
  VL_CHECK(a, "null args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GemmDims d;
  d.M = a->M;
  d.N = a->N;
  d.K = a->K;
  d.M_dev = a->M_dev;
  d.max_splits = a->max_splits > 0 ? a->max_splits : 1;
  EpiParams ep;
  ep.out = a->out;
  ep.ldo = a->ldo;
  ep.bias = a->bias;
  ep.scale = a->scale;
  ep.row_map = a->row_map;
  ep.row_pos = a->row_pos;
  ep.row_kvrow = a->row_kvrow;
  ep.cos_tab = a->cos_tab;
  ep.sin_tab = a->sin_tab;
  ep.k_cache = reinterpret_cast<__nv_bfloat16*>(a->k_cache);
  ep.v_cache = reinterpret_cast<__nv_bfloat16*>(a->v_cache);
  ep.kv_head_stride = a->kv_head_stride;
  ep.n_q_heads = a->n_q_heads;
  ep.n_kv_heads = a->n_kv_heads;
  ep.pos_emb = a->pos_emb;
  ep.patches = a->patches;
  ep.gamma_next = a->gamma_next;
  ep.a_next = reinterpret_cast<__nv_bfloat16*>(a->a_next);
  ep.lda_next = a->lda_next;
  ep.ss_next = a->ss_next;
  ep.tile_counters = a->tile_counters;
  ep.row_ss = a->row_ss;
  ep.inv_width = a->inv_width;
  ep.eps = a->eps;
  ep.part = a->partial;
  ep.part_stride = a->partial_stride;
  ep.part_splits = a->partial_splits;
  ep.ss_slots = a->ss_slots;
  d.stages = a->stages;
  d.pair = a->pair & 1;
  d.n_fast = (a->pair >> 1) & 1;
  d.pair_bn = (a->pair & 4) ? 128 : 256;
  d.half_sm = (a->pair >> 3) & 1;
  if (a->simt)
    return gemm_launch_simt(reinterpret_cast<const __nv_bfloat16*>(a->A), reinterpret_cast<const __nv_bfloat16*>(a->W),
                            d, ep, a->epi, st);
  CUtensorMap tmA, tmB;
  // rows of A beyond M are zero-filled by TMA (the map's row extent is the true M)
  VL_TRY(make_tmap_bf16(&tmA, a->A, (uint64_t)a->M, (uint64_t)a->K, 128));
  VL_TRY(make_tmap_bf16(&tmB, a->W, (uint64_t)a->N, (uint64_t)a->K, a->pair ? (uint32_t)(d.pair_bn / 2) : (uint32_t)a->block_n));
  CUtensorMap tmC;
  const bool tma_out = a->pair && a->tma_out;
  if (tma_out) VL_TRY(make_tmap_out(&tmC, a->out, (uint64_t)a->M, (uint64_t)a->ldo, a->epi == EPI_RESID_F32 ? 4 : 2));
  CUtensorMap tmP;
  const bool part_map = a->pair && a->partial != nullptr;
  if (part_map) {
    VL_CHECK(a->ldo > 0 && a->partial_stride > 0, "op_gemm: split-K partials need ldo and partial_stride");
    VL_TRY(make_tmap_out(&tmP, a->partial, (uint64_t)(d.max_splits * (a->partial_stride / a->ldo)), (uint64_t)a->ldo, 4));
  }
  return gemm_launch(tmA, tmB, d, ep, a->pair ? 128 : a->block_n, a->epi, st, tma_out ? &tmC : nullptr,
                     part_map ? &tmP : nullptr);
}

int32_t vlaser_op_attention(const vlaser_attn_args* a, void* stream) {
  VL_CHECK(a, "null args");
  AttnParams p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(a->q);
  p.q_row_stride = a->q_row_stride;
  p.q2 = reinterpret_cast<const __nv_bfloat16*>(a->q2);
  p.q2_row_stride = a->q2_row_stride;
  p.k = reinterpret_cast<const __nv_bfloat16*>(a->k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(a->v);
  p.kv_row_stride = a->kv_row_stride;
  p.kv_head_stride = a->kv_head_stride;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.out_row_stride = a->out_row_stride;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a->out2);
  p.out2_row_stride = a->out2_row_stride;
  p.q_start = a->q_start;
  p.q_len = a->q_len;
  p.kv_start = a->kv_start;
  p.kv_len = a->kv_len;
  p.uni_q_len = a->uni_q_len;
  p.uni_kv_len = a->uni_kv_len;
  p.max_q_len = a->max_q_len;
  p.group = a->group;
  p.n_kv_heads = a->n_kv_heads;
  p.batch = a->batch;
  p.scale = a->scale;
  p.causal = a->causal;
  p.small_rows = a->small_rows;
  VL_CHECK(!(p.causal && p.q2), "attention: the causal rule has no extra query row");
  return attention_launch(p, a->head_dim, reinterpret_cast<cudaStream_t>(stream));
}

int32_t vlaser_op_decode_attention(const void* q, const void* k_cache, const void* v_cache, void* out, const int32_t* kv_start,
                                   const int32_t* kv_len, int32_t batch, int32_t n_heads, int32_t n_kv_heads, int64_t kv_rows,
                                   int32_t max_kv_len, int32_t chunk, float scale, float* ws_o, float* ws_ml, const int32_t* q_start,
                                   void* stream) {
  VL_CHECK(q && k_cache && v_cache && out && kv_start && kv_len && ws_o && ws_ml && q_start, "null argument");
  VL_CHECK(n_kv_heads > 0 && n_heads % n_kv_heads == 0, "decode_attention: bad head counts");
  DecodeAttnParams p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q), p.q_row_stride = (long long)n_heads * 128;
  p.k = reinterpret_cast<const __nv_bfloat16*>(k_cache), p.v = reinterpret_cast<const __nv_bfloat16*>(v_cache);
  p.kv_row_stride = 128, p.kv_head_stride = kv_rows * 128;
  p.out = reinterpret_cast<__nv_bfloat16*>(out), p.out_row_stride = (long long)n_heads * 128;
  p.q_start = q_start, p.kv_start = kv_start, p.kv_len = kv_len;
  p.group = n_heads / n_kv_heads, p.n_kv_heads = n_kv_heads, p.batch = batch;
  p.max_kv_len = max_kv_len, p.chunk = chunk, p.scale = scale, p.ws_o = ws_o, p.ws_ml = ws_ml;
  return decode_attention_launch(p, reinterpret_cast<cudaStream_t>(stream));
}

int32_t vlaser_op_vit_attention(const void* qkv, void* out, int32_t images, int32_t tokens, int32_t heads, float scale,
                                void* stream) {
  return vit_attention_launch(reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), images,
                              tokens, heads, scale, reinterpret_cast<cudaStream_t>(stream));
}

int32_t vlaser_op_vit_attention_tc(const void* qkv, void* out, int32_t images, int32_t tokens, int32_t heads, float scale,
                                   void* stream) {
  VL_CHECK(qkv && out, "null argument");
  CUtensorMap tq, to;
  VL_TRY(make_tmap_bf16(&tq, qkv, (uint64_t)images * tokens, (uint64_t)3 * heads * 64, 128));
  VL_TRY(make_tmap_bf16(&to, out, (uint64_t)images * tokens, (uint64_t)heads * 64, 128));
  return vit_attention_tc_launch(tq, to, reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out),
                                 images, tokens, heads, scale, reinterpret_cast<cudaStream_t>(stream));
}

int32_t vlaser_op_prefill_attention_tc(const void* q, const void* k_cache, const void* v_cache, void* out, const int32_t* q_start,
                                       const int32_t* q_len, const int32_t* kv_start, int32_t q_rows, int32_t n_heads,
                                       int32_t n_kv_heads, int64_t kv_rows, int32_t batch, float scale, void* stream) {
  VL_CHECK(q && k_cache && v_cache && out && q_start && q_len && kv_start, "null argument");
  VL_CHECK(n_kv_heads > 0 && n_heads % n_kv_heads == 0, "prefill_attention_tc: bad head counts");
  CUtensorMap tq, tk, tv;
  VL_TRY(make_tmap_q3d(&tq, q, (uint64_t)q_rows, n_heads, n_heads / n_kv_heads));
  VL_TRY(make_tmap_bf16(&tk, k_cache, (uint64_t)n_kv_heads * kv_rows, 128, 32));
  VL_TRY(make_tmap_bf16(&tv, v_cache, (uint64_t)n_kv_heads * kv_rows, 128, 32));
  return prefill_attention_tc_launch(tq, tk, tv, q_start, q_len, kv_start, reinterpret_cast<__nv_bfloat16*>(out),
                                     (long long)n_heads * 128, 0, kv_rows, batch, n_kv_heads, n_heads / n_kv_heads, scale,
                                     reinterpret_cast<cudaStream_t>(stream));
}

int32_t vlaser_op_prefill_attention_tc_long(const void* q, const void* k_cache, const void* v_cache, void* out,
                                            const int32_t* q_start, const int32_t* q_len, const int32_t* kv_start, int32_t q_rows,
                                            int32_t n_heads, int32_t n_kv_heads, int64_t kv_rows, int32_t batch, int32_t max_tokens,
                                            int32_t min_len, int32_t causal, float scale, void* stream) {
  VL_CHECK(q && k_cache && v_cache && out && q_start && q_len && kv_start, "null argument");
  VL_CHECK(n_kv_heads > 0 && n_heads % n_kv_heads == 0, "prefill_attention_tc_long: bad head counts");
  CUtensorMap tq, tk, tv;
  VL_TRY(make_tmap_q3d(&tq, q, (uint64_t)q_rows, n_heads, n_heads / n_kv_heads));
  VL_TRY(make_tmap_bf16(&tk, k_cache, (uint64_t)n_kv_heads * kv_rows, 128, 32));
  VL_TRY(make_tmap_bf16(&tv, v_cache, (uint64_t)n_kv_heads * kv_rows, 128, 32));
  return prefill_attention_tc_long_launch(tq, tk, tv, q_start, q_len, kv_start, reinterpret_cast<__nv_bfloat16*>(out),
                                          (long long)n_heads * 128, 0, kv_rows, batch, n_kv_heads, n_heads / n_kv_heads, max_tokens,
                                          min_len, causal, scale, reinterpret_cast<cudaStream_t>(stream));
}

int32_t vlaser_op_layernorm(const float* x, const float* gamma, const float* beta, void* out_bf16, int32_t rows,
                            int32_t width, float eps, int32_t shuffle_grid, int32_t shuffle_channels, void* stream) {
  ShuffleGather sg;
  sg.grid = shuffle_grid;
  sg.channels = shuffle_channels;
  return layernorm_launch(const_cast<float*>(x), gamma, beta, reinterpret_cast<__nv_bfloat16*>(out_bf16), rows, width, eps, sg,
                          reinterpret_cast<cudaStream_t>(stream));
}
int32_t vlaser_op_rmsnorm(const float* x, const float* gamma, void* out_bf16, int32_t rows, int32_t width, float eps,
                          void* stream) {
  return rmsnorm_launch(const_cast<float*>(x), gamma, reinterpret_cast<__nv_bfloat16*>(out_bf16), rows, nullptr, width, eps,
                        reinterpret_cast<cudaStream_t>(stream));
}
int32_t vlaser_op_splitk_reduce(float* x, int32_t rows, int32_t width, const float* part, int64_t part_stride,
                                const int32_t* splits_dev, const float* gamma, void* out_bf16, float eps, void* stream) {
  VL_CHECK(x && part && splits_dev, "null argument");
  SplitKPartials sp;
  sp.part = part;
  sp.stride = part_stride;
  sp.splits = splits_dev;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (gamma && out_bf16)
    return rmsnorm_launch(x, gamma, reinterpret_cast<__nv_bfloat16*>(out_bf16), rows, nullptr, width, eps, st, nullptr, nullptr, &sp);
  return splitk_reduce_launch(x, rows, nullptr, width, sp, st);
}
int32_t vlaser_op_pixel_shuffle(const void* in, void* out, int32_t n, int32_t grid, int32_t channels,
                                int32_t elem_bytes, void* stream) {
  return pixel_shuffle_launch(in, out, n, grid, channels, elem_bytes, reinterpret_cast<cudaStream_t>(stream));
}
int32_t vlaser_op_im2col(const void* pixel, int32_t pixel_dtype, void* cols_bf16, int32_t n, int32_t px, int32_t ps,
                         int32_t kpad, void* stream) {
  return im2col_launch(pixel, pixel_dtype, reinterpret_cast<__nv_bfloat16*>(cols_bf16), n, px, ps, kpad,
                       reinterpret_cast<cudaStream_t>(stream));
}
int32_t vlaser_op_euler_step(const float* x, const float* gamma, const float* Wdec, const float* bdec, float* action,
                             void* out, int32_t out_dtype, int32_t rows, int32_t width, int32_t A, int32_t H,
                             int32_t horizon, float dt, float eps, int32_t last_step, int32_t do_clip, float clip,
                             void* stream) {
  return euler_step_launch(x, gamma, Wdec, bdec, action, out, out_dtype, rows, width, A, H, horizon, dt, eps, last_step,
                           do_clip, clip, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
