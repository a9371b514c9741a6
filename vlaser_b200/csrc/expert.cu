// Persistent expert-stream kernel (see expert.cuh).
//
// Grid = one CTA per SM, all co-resident; CTA = 320 threads:
//   warp 0      TMA producer (one lane): per ring stage one W block [128 x 64] and one or two A blocks [128 x 64]
//   warp 1      TMEM owner + tcgen05.mma issuer (one lane)
//   warps 2..9  workers: GEMM epilogues (tcgen05.ld + fused bias/RoPE/KV-write, SwiGLU, residual red.add),
//               RMSNorm rows, attention tiles (two 4-warp groups, mma.sync), grid-barrier arrivals
// A pass is a list of phases (7 per layer); consecutive phases are separated by a grid barrier (one global counter,
// release/acquire at gpu scope). Only consumers of another CTA's output wait on it:
//   * the producer lane before it loads ACTIVATION blocks of a GEMM phase -- the WEIGHT blocks of that phase are
//     issued into the ring before the wait, so HBM latency overlaps the barrier,
//   * the workers before a norm / attention phase.
// The MMA lane and the epilogues are ordered behind those loads by the ring's mbarriers. Split-K partial sums of the
// o/down projections meet in L2 (red.global.add.f32 into the fp32 residual stream), the following norm phase reads
// them back with ld.cg.
// A-operand rows beyond M come from zero-initialised padding of the activation buffers (never written).
//
// Sub-grids: every phase of a pass at M <= 512 is latency-bound (ring round trips, epilogue, barrier), not throughput-
// bound, and samples are independent. With p.n_streams = S > 1 the grid is cut into S independent sub-grids of
// gridDim.x / S CTAs; sub-grid s owns samples [s * batch/S, ...) = rows [s * M/S, ...) (a multiple of 32 rows, so no
// 32-row reduce-add box straddles two owners), has its own barrier counters and runs the whole pass on its rows.
// The sub-grids drift freely against each other, so one sub-grid's GEMM phase fills the SMs and the L2->SM fabric
// while another sits in a barrier or an epilogue. Each streams the weights itself; the second reader hits L2.
#include "expert.cuh"

#include "attention_device.cuh"
#include "gemm_epilogue.cuh"

namespace vlaser {

namespace {

constexpr int XP_WORKER_WARPS = 8;
constexpr int XP_WORKERS = 32 * XP_WORKER_WARPS;
constexpr int XP_THREADS = 64 + XP_WORKERS;
constexpr int XP_STAGES = 4;
constexpr uint32_t XP_SUB = BLOCK_M * BLOCK_K * 2;   // 16 KB: one 128-row x 64-column bf16 operand block
constexpr uint32_t XP_STAGE_BYTES = 3 * XP_SUB;      // A rows [0,128) | W | A rows [128,256)
constexpr uint32_t XP_RING_BYTES = XP_STAGES * XP_STAGE_BYTES;
constexpr uint32_t XP_OUT_BYTES = XP_WORKER_WARPS * 4096;  // one 32-row x 128-byte staging tile per worker warp
constexpr uint32_t XP_SMEM = XP_RING_BYTES + XP_OUT_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr uint32_t XP_ACC_COLS = 256;                // TMEM columns per accumulator stage (two 128-column sub-tiles)
constexpr int XP_PHASES = 7;
constexpr int XP_MAX_STREAMS = EXPERT_MAX_STREAMS;
enum : int { PH_NORM1 = 0, PH_QKV = 1, PH_ATTN = 2, PH_O = 3, PH_NORM2 = 4, PH_GU = 5, PH_DOWN = 6 };

static_assert(2 * attention_tile_smem<128>() <= (int)XP_RING_BYTES, "attention staging must fit in the idle ring");

struct GemmPhase {
  const CUtensorMap* tmA;
  const CUtensorMap* tmB;
  int N, K;
  int MT;          // 128-row sub-tiles per output tile (they share the W block)
  int epi;
  int max_splits;  // split-K cap (residual epilogue only)
  int ahead;       // W blocks may be issued before the grid barrier (the ring is not aliased by the previous phase)
};

__device__ __forceinline__ GemmPhase gemm_phase(const ExpertPass& p, int rows, const CUtensorMap* tmXn,
                                                const CUtensorMap* tmAo, const CUtensorMap* tmAct, int layer, int j) {
  const ExpertLayerDev* L = p.layers + layer;
  const int qdim = p.n_q_heads * 128;
  const int mt = rows > BLOCK_M ? 2 : 1;
  GemmPhase g;
  if (j == PH_QKV) g = GemmPhase{tmXn, &L->tm_qkv, qdim + 2 * p.n_kv_heads * 128, p.he, 1, EPI_QKV_ROPE, 1, 1};
  else if (j == PH_O) g = GemmPhase{tmAo, &L->tm_o, p.he, qdim, mt, EPI_RESID_F32, 32, 0};  // ring aliased by attention
  else if (j == PH_GU) g = GemmPhase{tmXn, &L->tm_gu, 2 * p.mlp, p.he, mt, EPI_SWIGLU_BF16, 1, 1};
  else g = GemmPhase{tmAct, &L->tm_down, p.he, p.mlp, mt, EPI_RESID_F32, 32, 1};
  return g;
}

struct Sched {
  int m_groups, base, kb_per_split, k_blocks, total;
};
__device__ __forceinline__ Sched make_sched(const GemmPhase& g, int M, int ctas, int flags) {
  Sched s;
  s.m_groups = (M + g.MT * BLOCK_M - 1) / (g.MT * BLOCK_M);
  s.k_blocks = g.K / BLOCK_K;
  s.base = s.m_groups * (g.N / 128);
  int splits = 1;
  if (g.max_splits > 1) {
    splits = min(min(ctas / s.base, g.max_splits), s.k_blocks / 2);  // >= two k-blocks per split
    if (splits < 1) splits = 1;
  }
  s.kb_per_split = (s.k_blocks + splits - 1) / splits;
  splits = (s.k_blocks + s.kb_per_split - 1) / s.kb_per_split;
  s.total = (flags & 4) ? 0 : s.base * splits;  // flag 4 (profiling aid): barrier skeleton only
  return s;
}
struct Tile {
  int split, n_idx, m_grp, kb0, kb1;
};
__device__ __forceinline__ Tile tile_of(const Sched& s, int t) {
  Tile c;
  c.split = t / s.base;
  const int rem = t - c.split * s.base;
  c.n_idx = rem / s.m_groups;
  c.m_grp = rem - c.n_idx * s.m_groups;
  c.kb0 = c.split * s.kb_per_split;
  c.kb1 = min(s.k_blocks, c.kb0 + s.kb_per_split);
  return c;
}

// ---- grid barrier: one arrival counter per phase (a CTA without work in a GEMM phase arrives at once and may run
// several phases ahead of a busy one, so a single running sum over phases would not prove that EVERY CTA is done);
// phase ph may consume other CTAs' output once counter[ph - 1] == gridDim.x
// Arrival (one thread per CTA, after a CTA barrier over the workers): a gpu-scope release is cumulative over the
// writes of every thread that synchronised with this one. red.release + ld.acquire polling measured 1.16 us per
// barrier on 148 CTAs against 1.43 us for fence + relaxed red / relaxed poll + fence and 3.4 us for per-CTA flag
// words polled by a warp (tools/probes/barrier_probe.cu, profiles/r01_barrier_probe.txt).
__device__ __forceinline__ void grid_arrive(unsigned int* ctr) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
// (~2 s without the other CTAs: one of them is missing -- the launcher checks residency, so this is insurance. No __trap,
// which would kill the caller's CUDA context: flag the call as failed for the host and run on without waiting.)
__device__ __forceinline__ void grid_wait(const unsigned int* ctr, unsigned int target, int* err_flag) {
  unsigned int v;
  long long t0 = 0;
  while (true) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= target) break;
    if (err_flag && *reinterpret_cast<volatile int*>(err_flag) == EXPERT_ERR_BARRIER) break;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 4000000000LL) {
      if (err_flag) atomicExch(err_flag, EXPERT_ERR_BARRIER);
      break;
    }
  }
}
// mbarrier wait that traps instead of hanging the GPU if a pipeline partner never arrives (bring-up insurance: a
// persistent kernel with a grid barrier must never spin forever)
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(XP_WORKERS) : "memory"); }

__global__ void __launch_bounds__(XP_THREADS, 1)
expert_pass_kernel(const __grid_constant__ CUtensorMap tmXn, const __grid_constant__ CUtensorMap tmAo,
                   const __grid_constant__ CUtensorMap tmAct, const __grid_constant__ CUtensorMap tmX,
                   const ExpertPass p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_stage = smem + XP_RING_BYTES;  // [XP_WORKER_WARPS][4096], 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + XP_OUT_BYTES);
  uint64_t* full_bar = bars;                        // [XP_STAGES]
  uint64_t* empty_bar = bars + XP_STAGES;           // [XP_STAGES]
  uint64_t* tfull_bar = bars + 2 * XP_STAGES;       // [2]
  uint64_t* tempty_bar = bars + 2 * XP_STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * XP_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // sub-grid of this CTA: G CTAs, index cta, rows [row_base, row_base + Mg), samples [b_base, b_base + Bg)
  const int S = p.n_streams > 1 ? p.n_streams : 1;
  const unsigned int G = gridDim.x / S;
  const int sg = (int)(blockIdx.x / G);
  const int cta = (int)(blockIdx.x - sg * G);
  const bool idle = sg >= S;  // gridDim.x % S left-over CTAs
  const int rows_sg = p.M / S, batch_sg = p.batch / S;
  const int row_base = sg * rows_sg, b_base = sg * batch_sg;
  const int Mg = rows_sg, Bg = batch_sg;
  unsigned int* const bar = p.barrier + (size_t)sg * XP_PHASES * p.n_layers;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmXn);
    tma_prefetch_desc(&tmAo);
    tma_prefetch_desc(&tmAct);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < XP_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], XP_WORKER_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * XP_ACC_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (idle) {
    // nothing to do (keeps the TMEM allocation symmetric; falls through to the common exit)
  } else if (warp == 0) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int layer = 0; layer < p.n_layers; ++layer) {
        const bool kv_only = p.last_kv_only && layer == p.n_layers - 1;
        for (int jj = 0; jj < 4; ++jj) {
          const int j = jj == 0 ? PH_QKV : jj == 1 ? PH_O : jj == 2 ? PH_GU : PH_DOWN;
          if (kv_only && j != PH_QKV) break;
          const unsigned int ph = (unsigned int)(layer * XP_PHASES + j);
          const GemmPhase g = gemm_phase(p, Mg, &tmXn, &tmAo, &tmAct, layer, j);
          const Sched sc = make_sched(g, Mg, (int)G, p.flags);
          const uint32_t tx = (uint32_t)(g.MT + 1) * XP_SUB;
          bool open = false;
          int npend = 0;
          int pend_stage[XP_STAGES], pend_kb[XP_STAGES], pend_row[XP_STAGES];
          auto load_a = [&](int st, int kb, int row0) {
            uint8_t* s = smem + st * XP_STAGE_BYTES;
            tma_load_2d(s, g.tmA, &full_bar[st], kb * BLOCK_K, row0);
            if (g.MT == 2) tma_load_2d(s + 2 * XP_SUB, g.tmA, &full_bar[st], kb * BLOCK_K, row0 + BLOCK_M);
          };
          auto open_gate = [&]() {
            // every CTA has published all of its work of the earlier phases (and the previous grid is complete:
            // the workers arrive only after their own griddepcontrol.wait)
            grid_wait(bar + (ph - 1), G, p.err_flag);
            fence_proxy_async_all();  // generic-proxy writes of other SMs -> this thread's async-proxy (TMA) reads
            for (int i = 0; i < npend; ++i) load_a(pend_stage[i], pend_kb[i], pend_row[i]);
            npend = 0;
            open = true;
          };
          // the attention phase stages K/V in the ring: nothing may be loaded (not even by a CTA without o-proj
          // tiles, which would otherwise run ahead into the gate/up weights) before every attention tile is done
          if (j == PH_O) open_gate();
          for (int t = cta; t < sc.total; t += (int)G) {
            const Tile tc = tile_of(sc, t);
            const int row0 = row_base + tc.m_grp * g.MT * BLOCK_M;
            for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
              if (!open && (!g.ahead || npend == XP_STAGES)) open_gate();
              mbar_wait_guard(&empty_bar[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full_bar[stage], tx);
              tma_load_2d(smem + stage * XP_STAGE_BYTES + XP_SUB, g.tmB, &full_bar[stage], kb * BLOCK_K, tc.n_idx * 128);
              if (open) {
                load_a(stage, kb, row0);
              } else {
                pend_stage[npend] = stage;
                pend_kb[npend] = kb;
                pend_row[npend] = row0;
                ++npend;
              }
              if (++stage == XP_STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
          if (!open && npend) open_gate();
        }
      }
    }
  } else if (warp == 1) {
    // ======================================= MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, 128);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int layer = 0; layer < p.n_layers; ++layer) {
        const bool kv_only = p.last_kv_only && layer == p.n_layers - 1;
        for (int jj = 0; jj < 4; ++jj) {
          const int j = jj == 0 ? PH_QKV : jj == 1 ? PH_O : jj == 2 ? PH_GU : PH_DOWN;
          if (kv_only && j != PH_QKV) break;
          const GemmPhase g = gemm_phase(p, Mg, &tmXn, &tmAo, &tmAct, layer, j);
          const Sched sc = make_sched(g, Mg, (int)G, p.flags);
          for (int t = cta; t < sc.total; t += (int)G) {
            const Tile tc = tile_of(sc, t);
            mbar_wait_guard(&tempty_bar[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * XP_ACC_COLS;
            for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
              mbar_wait_guard(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem + stage * XP_STAGE_BYTES);
              const uint64_t db = make_umma_desc_sw128(sa + XP_SUB);
              for (int sub = 0; sub < g.MT; ++sub) {
                const uint64_t da = make_umma_desc_sw128(sa + sub * 2 * XP_SUB);
#pragma unroll
                for (int k = 0; k < BLOCK_K / 16; ++k)
                  umma_bf16(tmem_d + sub * 128, da + 2 * k, db + 2 * k, idesc, (kb > tc.kb0 || k > 0) ? 1u : 0u);
              }
              umma_commit(&empty_bar[stage]);
              if (++stage == XP_STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
            umma_commit(&tfull_bar[as]);
            as ^= 1;
            if (as == 0) aphase ^= 1;
          }
        }
      }
    }
  } else {
    // ======================================= workers =======================================
    const int wtid = threadIdx.x - 64;
    const int wwarp = wtid >> 5;
    const int quarter = warp & 3;      // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which half of a sub-tile's columns it drains
    const int qdim = p.n_q_heads * 128;
    uint8_t* wbuf = out_stage + (warp - 2) * 4096;
    pdl_wait();                        // the previous grid's activations (x, KV cache, plan) are complete
    if (p.timing && blockIdx.x == 0 && wtid == 0) {
      long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      p.timing[7 * p.n_layers - 1] = now;  // pass start (the last phase never records: no barrier follows it)
    }
    int as = 0;
    uint32_t aphase = 0;
    const int n_phases = (p.n_layers - 1) * XP_PHASES + (p.last_kv_only ? 2 : XP_PHASES);
    for (int layer = 0; layer < p.n_layers; ++layer) {
      const ExpertLayerDev* L = p.layers + layer;
      for (int j = 0; j < XP_PHASES; ++j) {
        const int ph = layer * XP_PHASES + j;
        if (ph >= n_phases) break;
        if (j == PH_ATTN) {
          // the per-sample segment table is constant: pull this group's first entry into L1 ahead of the barrier
          const int grp0 = wtid >> 7;
          const int qblocks0 = ((p.n_q_heads / p.n_kv_heads) * p.max_q_len + ATT_BM - 1) / ATT_BM;
          const int it0 = cta + (int)G * grp0;
          if (it0 < Bg * p.n_kv_heads * qblocks0) {
            const int b0 = b_base + it0 / qblocks0 / p.n_kv_heads;
            const int pre = p.q_len[b0] + p.q_start[b0] + p.kv_len[b0] + p.kv_start[b0];
            asm volatile("" ::"r"(pre));
          }
        }
        if (j == PH_NORM1 || j == PH_NORM2 || j == PH_ATTN) {
          if (ph > 0) {
            if (wtid == 0) grid_wait(bar + (ph - 1), G, p.err_flag);
            worker_bar();
          }
        }
        if (j == PH_NORM1 || j == PH_NORM2) {
          // ---- RMSNorm (HF Qwen2RMSNorm): one warp per row, fp32 row from L2, bf16 GEMM operand out
          const float* gamma = j == PH_NORM1 ? L->ln_in : L->ln_post;
          const int nch = p.he >> 7;  // float4 chunks per lane
          const float inv_w = 1.0f / (float)p.he;
          for (int r = row_base + cta + (int)G * wwarp; r < ((p.flags & 8) ? 0 : row_base + Mg); r += (int)G * XP_WORKER_WARPS) {
            const float4* src = reinterpret_cast<const float4*>(p.x + (size_t)r * p.he);
            float4 v[8];
            float qs = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (i < nch) {
                v[i] = __ldcg(src + i * 32 + lane);
                qs += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
              }
            }
            const float rstd = rsqrtf(warp_sum(qs) * inv_w + p.eps);
            uint2* dst = reinterpret_cast<uint2*>(p.xn + (size_t)r * p.he);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (i < nch) {
                const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
                uint2 u;
                u.x = pack_bf16(v[i].x * rstd * gm.x, v[i].y * rstd * gm.y);
                u.y = pack_bf16(v[i].z * rstd * gm.z, v[i].w * rstd * gm.w);
                dst[i * 32 + lane] = u;
              }
            }
          }
        } else if (j == PH_ATTN) {
          // ---- attention over the cached prefix KV + this pass' own rows: two 4-warp groups, one work item each
          AttnParams ap;
          ap.q = p.q;
          ap.q_row_stride = qdim;
          ap.k = p.k_cache + (size_t)layer * p.kv_layer_stride;
          ap.v = p.v_cache + (size_t)layer * p.kv_layer_stride;
          ap.kv_row_stride = 128;
          ap.kv_head_stride = p.kv_head_stride;
          ap.out = p.ao;
          ap.out_row_stride = qdim;
          ap.q_start = p.q_start;
          ap.q_len = p.q_len;
          ap.kv_start = p.kv_start;
          ap.kv_len = p.kv_len;
          ap.max_q_len = p.max_q_len;
          ap.group = p.n_q_heads / p.n_kv_heads;
          ap.n_kv_heads = p.n_kv_heads;
          ap.batch = p.batch;
          ap.scale = p.scale;
          const int grp = wtid >> 7, gtid = wtid & 127;
          uint8_t* asmem = smem + grp * attention_tile_smem<128>();
          const int qblocks = (ap.group * p.max_q_len + ATT_BM - 1) / ATT_BM;
          const int items = Bg * p.n_kv_heads * qblocks;
          for (int it = cta + (int)G * grp; it < ((p.flags & 2) ? 0 : items); it += 2 * (int)G) {
            const int qb = it % qblocks;
            const int rest = it / qblocks;
            const int kvh = rest % p.n_kv_heads;
            const int b = b_base + rest / p.n_kv_heads;
            attention_tile<128>(ap, b, kvh, qb, asmem, gtid,
                                [grp] { asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory"); }, [] {});
          }
          fence_proxy_async();  // this thread's generic-proxy smem accesses precede the TMA writes that reuse the ring
        } else {
          // ---- GEMM epilogue
          const GemmPhase g = gemm_phase(p, Mg, &tmXn, &tmAo, &tmAct, layer, j);
          const Sched sc = make_sched(g, Mg, (int)G, p.flags);
          EpiParams ep;
          if (j == PH_QKV) {
            ep.out = p.q;
            ep.ldo = qdim;
            ep.bias = L->bqkv;
            ep.row_pos = p.row_pos;
            ep.row_kvrow = p.row_kvrow;
            ep.cos_tab = p.cos_tab;
            ep.sin_tab = p.sin_tab;
            ep.k_cache = p.k_cache + (size_t)layer * p.kv_layer_stride;
            ep.v_cache = p.v_cache + (size_t)layer * p.kv_layer_stride;
            ep.kv_head_stride = p.kv_head_stride;
            ep.n_q_heads = p.n_q_heads;
            ep.n_kv_heads = p.n_kv_heads;
          } else if (j == PH_GU) {
            ep.out = p.act;
            ep.ldo = p.mlp;
          } else {
            ep.out = p.x;
            ep.ldo = p.he;
          }
          for (int t = cta; t < sc.total; t += (int)G) {
            const Tile tc = tile_of(sc, t);
            mbar_wait_guard(&tfull_bar[as], aphase);
            tc_fence_after();
            const int n0 = tc.n_idx * 128;
            for (int sub = 0; sub < g.MT; ++sub) {
              const int row = row_base + (tc.m_grp * g.MT + sub) * BLOCK_M + quarter * 32 + lane;
              const bool row_ok = row < row_base + Mg;
              const uint32_t taddr =
                  tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * XP_ACC_COLS + sub * 128;
              float ss = 0.f;
              // residual sums always go through red.add at L2: another CTA may own a different K split of this tile,
              // and plain loads of x could hit a stale L1 line of an earlier layer
              if (j == PH_QKV) epilogue_tile<128, EPI_QKV_ROPE>(ep, g.N, taddr, half, row, row_ok, n0, true, false, 1.f, ss);
              else if (j == PH_GU) epilogue_tile<128, EPI_SWIGLU_BF16>(ep, g.N, taddr, half, row, row_ok, n0, true, false, 1.f, ss);
              else if (!(p.flags & 1) && row - lane < row_base + Mg) {  // (a box past the sub-grid's rows: not ours)
                // split-K partial sums meet in L2: the 32 x 32 fp32 chunk of this warp leaves as one bulk tensor
                // reduce-add (full 128-byte lines; per-lane red.v4 costs 32 L2 transactions per instruction)
#pragma unroll 1
                for (int c = half * 2; c < half * 2 + 2; ++c) {
                  float v[32];
                  load_acc32(taddr + c * 32, v);
                  if (lane == 0) bulk_wait_read<0>();  // the staging tile has been read out by the previous reduce
                  __syncwarp();
                  stage_row_f32(wbuf, lane, v);
                  fence_proxy_async();
                  __syncwarp();
                  if (lane == 0) {
                    tma_reduce_add_2d(&tmX, wbuf, n0 + c * 32, row - lane);
                    bulk_commit();
                  }
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            as ^= 1;
            if (as == 0) aphase ^= 1;
          }
        }
        if (ph + 1 < n_phases) {
          if ((j == PH_O || j == PH_DOWN) && lane == 0) bulk_wait<0>();  // this warp's reduce-adds have been performed
          fence_proxy_async_all();  // this thread's global writes are ordered before async-proxy (TMA) reads of them
          worker_bar();             // every worker's writes happen-before the CTA's single (fenced) arrival
          if (wtid == 0) grid_arrive(bar + ph);
          if (p.timing && blockIdx.x == 0 && wtid == 0) {  // profiling aid: when did the whole grid finish this phase
            grid_wait(bar + ph, G, p.err_flag);
            long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            p.timing[ph] = now;
          }
        }
      }
    }
  }

  if (warp >= 2 && lane == 0) bulk_wait<0>();  // reduce-adds of the last phase have landed before the grid completes
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * XP_ACC_COLS);
}

}  // namespace

int expert_pass_launch(const ExpertPass& p, const CUtensorMap& tmXn, const CUtensorMap& tmAo, const CUtensorMap& tmAct,
                       const CUtensorMap& tmX, cudaStream_t stream) {
  VL_CHECK(p.layers && p.n_layers > 0 && p.M > 0, "expert pass: empty problem");
  VL_CHECK(p.he % 128 == 0 && p.he <= 1024, "expert pass: hidden width must be a multiple of 128, <= 1024");
  VL_CHECK(p.mlp % 64 == 0, "expert pass: MLP width must be a multiple of 64");
  VL_CHECK(p.n_kv_heads > 0 && p.n_q_heads % p.n_kv_heads == 0, "expert pass: bad head counts");
  VL_CHECK(p.barrier != nullptr, "expert pass: null barrier counter");
  const int S = p.n_streams > 1 ? p.n_streams : 1;
  VL_CHECK(S == 1 || (p.batch % S == 0 && p.M % p.batch == 0 && (p.M / S) % 32 == 0 && S <= XP_MAX_STREAMS),
           "expert pass: sub-grids need whole samples and a multiple of 32 rows each");
  static int max_ctas = -1;
  if (max_ctas < 0) {
    VL_CUDA_CHECK(cudaFuncSetAttribute(expert_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XP_SMEM));
    int per_sm = 0;
    VL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, expert_pass_kernel, XP_THREADS, XP_SMEM));
    VL_CHECK(per_sm >= 1, "expert pass: kernel does not fit on an SM");
    max_ctas = per_sm * gemm_num_sms();
  }
  // the grid barrier needs every CTA resident at once: one CTA per SM
  const int grid = std::min(gemm_num_sms(), max_ctas);
  VL_CUDA_CHECK(launch_kernel(expert_pass_kernel, dim3(grid), dim3(XP_THREADS), XP_SMEM, stream, tmXn, tmAo, tmAct, tmX, p));
  return 0;
}

}  // namespace vlaser
