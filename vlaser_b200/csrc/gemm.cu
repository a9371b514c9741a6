// tcgen05 / TMEM / TMA GEMM for sm_100a. See gemm.cuh for the contract.
//
// CTA = 192 threads, persistent over output tiles (static round-robin), one CTA per SM:
//   warp 0      TMA producer (one lane): A tile [128 x 64] + W tile [BLOCK_N x 64] per stage
//   warp 1      TMEM allocator + MMA issuer (one lane): 4 x tcgen05.mma (K=16) per stage
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns per warp, fused epilogue, global stores
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM accumulator double buffer full/empty
// (MMA <-> epilogue) so the epilogue of tile i overlaps the mainloop of tile i+1.
// The live row count M may come from device memory (var-len packed prefix), which keeps every
// launch configuration static and the whole policy call CUDA-graph capturable.
#include "gemm.cuh"
#include "gemm_epilogue.cuh"

#include <mutex>

namespace vlaser {

namespace {

constexpr int EPI_WARPS = 8;                        // two warps per TMEM lane quarter, each drains half the columns
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;   // TMA warp + MMA warp + epilogue warps

constexpr int MAX_STAGES = 8;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int DEFAULT_STAGES = (BLOCK_N == 256) ? 4 : 6;
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages
  // the q|k|v epilogue keeps two 32-row x 128-byte tiles per epilogue warp (the unit's RoPE factors, fetched while the
  // mainloop runs; the first one then transposes the rotated unit for the stores); other epilogues: none
  static constexpr uint32_t epi_stage_bytes(int epi) { return epi == EPI_QKV_ROPE ? EPI_WARPS * 8192 : 0; }
  static constexpr uint32_t smem_bytes(int stages, int epi) {
    return stages * STAGE_BYTES + epi_stage_bytes(epi) + 1024 /*align slack*/ + 256 /*barriers*/;
  }
  // ring depth: the q|k|v variant gives one stage to the transposition tiles (and leaves L1 some room: 198 KB of shared memory)
  static constexpr int default_stages(int epi) { return epi == EPI_QKV_ROPE ? DEFAULT_STAGES - 1 : DEFAULT_STAGES; }
};

struct TileSched {
  int m_tiles, n_tiles, base, splits, kb_per_split, k_blocks, total;
};

template <int BLOCK_N, int EPI>
__device__ __forceinline__ TileSched make_sched(const GemmDims& d, int M) {
  TileSched s;
  s.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  s.n_tiles = d.N / BLOCK_N;
  s.k_blocks = d.K / BLOCK_K;
  s.base = s.m_tiles * s.n_tiles;
  int splits = 1;
  if (EPI == EPI_RESID_F32 && d.max_splits > 1 && s.base > 0) {
    splits = (int)gridDim.x / s.base;
    if (splits > d.max_splits) splits = d.max_splits;
    if (splits > s.k_blocks / 2) splits = s.k_blocks / 2;  // at least two k-blocks per split
    if (splits < 1) splits = 1;
  }
  s.kb_per_split = (s.k_blocks + splits - 1) / splits;
  s.splits = (s.k_blocks + s.kb_per_split - 1) / s.kb_per_split;
  s.total = s.base * s.splits;
  return s;
}


// ------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------
struct TileCoord {
  int split, n_idx, m_idx, kb0, kb1, rem;
};
__device__ __forceinline__ TileCoord tile_coord(const TileSched& sc, int t) {
  TileCoord c;
  c.split = t / sc.base;
  c.rem = t - c.split * sc.base;
  c.n_idx = c.rem / sc.m_tiles;
  c.m_idx = c.rem - c.n_idx * sc.m_tiles;
  c.kb0 = c.split * sc.kb_per_split;
  c.kb1 = min(sc.k_blocks, c.kb0 + sc.kb_per_split);
  return c;
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmDims d,
                    EpiParams ep) {
  using Cfg = GemmCfg<BLOCK_N>;
  const int STAGES = d.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES;  // [EPI_WARPS][4096]
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + Cfg::epi_stage_bytes(EPI));
  uint64_t* full_bar = bars;                          // [MAX_STAGES]
  uint64_t* empty_bar = bars + MAX_STAGES;            // [MAX_STAGES]
  uint64_t* tfull_bar = bars + 2 * MAX_STAGES;        // [2]
  uint64_t* tempty_bar = bars + 2 * MAX_STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  volatile int* fin_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long* tr = (blockIdx.x == 0 && lane == 0) ? d.trace : nullptr;

  if (warp == 0 && lane == 0) {
    trace_stamp(tr, 0);
    trace_max(d.trace, 14);
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // The next kernel of the chain may start its own prologue now (its CTAs find TMEM/smem as ours retire).
  pdl_launch_dependents();

  // Live row count: a device-resident M belongs to this call's plan, so it is read after the PDL wait.
  int M = d.M;
  if (d.M_dev) {
    pdl_wait();
    M = min(*d.M_dev, d.M);
  }
  const TileSched sc = make_sched<BLOCK_N, EPI>(d, M);

  if (warp == 0) {
    // ================================ TMA producer ================================
    // (first ring fill: one lane per stage, see the pair kernel; steady state: lane 0)
    {
      int stage = 0;
      uint32_t phase = 0;
      int t = blockIdx.x;
      int kb = 0;
      TileCoord tc{};
      bool have = t < sc.total;
      if (have) {
        tc = tile_coord(sc, t);
        kb = tc.kb0;
      }
      auto advance = [&]() {
        if (++kb >= tc.kb1) {
          t += gridDim.x;
          have = t < sc.total;
          if (have) {
            tc = tile_coord(sc, t);
            kb = tc.kb0;
          }
        }
      };
      if (!d.M_dev) {
        // Weights do not depend on the previous kernel: fill the ring's B halves before the PDL wait,
        // then add the A halves (activations written by the previous kernel) after it.
        int my_n = 0, my_m = 0, my_kb = 0;
        bool mine = false;
        int npre = 0;
        while (have && npre < STAGES) {
          if (lane == npre) mine = true, my_n = tc.n_idx, my_m = tc.m_idx, my_kb = kb;
          ++npre;
          advance();
        }
        if (mine) {
          mbar_arrive_expect_tx(&full_bar[lane], Cfg::STAGE_BYTES);
          tma_load_2d(smem + lane * Cfg::STAGE_BYTES + Cfg::A_BYTES, &tmB, &full_bar[lane], my_kb * BLOCK_K, my_n * BLOCK_N);
        }
        if (lane == 0) trace_max(d.trace, 15);
        pdl_wait();
        trace_stamp(tr, 1);
        if (mine) tma_load_2d(smem + lane * Cfg::STAGE_BYTES, &tmA, &full_bar[lane], my_kb * BLOCK_K, my_m * BLOCK_M);
        stage = npre;
        if (stage == STAGES) {
          stage = 0;
          phase = 1;
        }
      }
      if (lane == 0) {
        while (have) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, tc.m_idx * BLOCK_M);
          tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, tc.n_idx * BLOCK_N);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
          advance();
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < sc.total; t += gridDim.x) {
        const TileCoord tc = tile_coord(sc, t);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (t == (int)blockIdx.x && kb == tc.kb0) trace_stamp(tr, 2);
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t da = make_umma_desc_sw128(sa);
          const uint64_t db = make_umma_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128B swizzle atom: +2 in the (addr>>4) field
            umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > tc.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);  // accumulator ready for the epilogue
        trace_stamp(tr, 3);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ================================ epilogue warps ================================
    pdl_wait();
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // which half of the tile's columns it drains
    int as = 0;
    uint32_t aphase = 0;
    if (warp != 2) tr = nullptr;
    if constexpr (EPI == EPI_RESID_F32) {
      if (ep.part_splits && blockIdx.x == 0 && threadIdx.x == 64) *ep.part_splits = sc.splits > 1 ? sc.splits : 0;
    }
    for (int t = blockIdx.x; t < sc.total; t += gridDim.x) {
      const TileCoord tc = tile_coord(sc, t);
      const int n_idx = tc.n_idx, m_idx = tc.m_idx;
      const int row = m_idx * BLOCK_M + quarter * 32 + lane;
      const bool row_ok = row < M;
      const int n0 = n_idx * BLOCK_N;
      float rstd = 1.f;
      if (ep.row_ss && row_ok) rstd = rsqrtf(__ldcg(ep.row_ss + row) * ep.inv_width + ep.eps);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BLOCK_N;
      const bool first_split = tc.split == 0;
      const bool atomic = sc.splits > 1;
      float ss = 0.f;
      float* part_tile = nullptr;
      if constexpr (EPI == EPI_RESID_F32) {
        if (atomic && ep.part && !ep.a_next) part_tile = ep.part + (size_t)tc.split * ep.part_stride;
      }
      // q|k|v projection, 128-column tiles: this warp owns ONE unit of the tile (half-head chunk `half`); its RoPE
      // factors and KV-cache row are fetched while the mainloop is still running
      constexpr bool ROPE_PRE = EPI == EPI_QKV_ROPE && BLOCK_N == 128 && EPI_WARPS == 8;
      RopeOperands ro;
      int rope_pos = 0;
      float bias_v = 0.f, bias_w = 0.f;
      if constexpr (ROPE_PRE) {
        // Everything of this unit that does not depend on the accumulator is fetched while the mainloop runs: the row's
        // position id and KV-cache row (registers), the tile's bias and the unit's 2 x 32 RoPE factors per row (shared
        // memory: held in registers across the wait they spill). The factor rows are read in the transposed domain --
        // 8 lanes share a row, 128-byte runs -- instead of 32 scattered 16-byte pieces per instruction behind the
        // accumulator (in-situ stamps: that epilogue took 3.8 us of the launch's 8, profiles/r02s2y).
        uint8_t* cs_tile = epi_stage + (warp - 2) * 8192;
        uint8_t* sn_tile = cs_tile + 4096;
        if (row_ok) {
          rope_pos = __ldg(ep.row_pos + row);
          ro.kvrow = __ldg(ep.row_kvrow + row);
        }
        if (ep.bias) {  // lane j keeps bias[j] of both chunks (cold in L2: a DRAM round trip); broadcast by shuffle below
          bias_v = __ldg(ep.bias + n0 + half * 32 + lane);
          bias_w = __ldg(ep.bias + n0 + half * 32 + 64 + lane);
        }
        const bool rot = (n0 >> 7) < ep.n_q_heads + ep.n_kv_heads;
        if (rot) {
          const int ch = lane & 7;
          float4 c4[8], s4[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + (lane >> 3);
            const int p = __shfl_sync(0xffffffffu, rope_pos, r);
            c4[it] = __ldg(reinterpret_cast<const float4*>(ep.cos_tab + (size_t)p * 64 + half * 32 + 4 * ch));
            s4[it] = __ldg(reinterpret_cast<const float4*>(ep.sin_tab + (size_t)p * 64 + half * 32 + 4 * ch));
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + (lane >> 3);
            const int sw = r * 128 + ((ch ^ (r & 7)) << 4);
            *reinterpret_cast<float4*>(cs_tile + sw) = c4[it];
            *reinterpret_cast<float4*>(sn_tile + sw) = s4[it];
          }
        }
        __syncwarp();
      }
      {
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        if (t == (int)blockIdx.x) trace_stamp(tr, 4);
        if constexpr (ROPE_PRE) {
          uint8_t* buf = epi_stage + (warp - 2) * 8192;  // factor tile, then the transposition tile of the stores
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr + half * 32, r0);
          tmem_ld_32x32(taddr + half * 32 + 64, r1);
          const int col0 = n0 + half * 32;
          if ((n0 >> 7) < ep.n_q_heads + ep.n_kv_heads) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int sw = lane * 128 + ((c ^ (lane & 7)) << 4);
              const float4 cc = *reinterpret_cast<const float4*>(buf + sw);
              const float4 sn = *reinterpret_cast<const float4*>(buf + 4096 + sw);
              ro.cs[4 * c] = cc.x, ro.cs[4 * c + 1] = cc.y, ro.cs[4 * c + 2] = cc.z, ro.cs[4 * c + 3] = cc.w;
              ro.sn[4 * c] = sn.x, ro.sn[4 * c + 1] = sn.y, ro.sn[4 * c + 2] = sn.z, ro.sn[4 * c + 3] = sn.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) ro.cs[j] = 1.f, ro.sn[j] = 0.f;
          }
          tmem_ld_wait();
          float v[32], w[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = __uint_as_float(r0[j]) * rstd + __shfl_sync(0xffffffffu, bias_v, j);
            w[j] = __uint_as_float(r1[j]) * rstd + __shfl_sync(0xffffffffu, bias_w, j);
          }
          // Transpose the unit through this warp's 4 KB swizzled staging tile so that every store instruction writes
          // whole 64-byte runs (4 rows x 2 runs) instead of 32 scattered 16-byte pieces
          if ((n0 >> 7) < ep.n_q_heads + ep.n_kv_heads) {  // q or k head: x*cos + rot(x)*sin (rope_rotate)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float a = v[j], b = w[j];
              v[j] = a * ro.cs[j] - b * ro.sn[j];
              w[j] = b * ro.cs[j] + a * ro.sn[j];
            }
          }
          __syncwarp();  // every lane has read its factors before the tile is overwritten
          stage_row_bf16(buf, lane, 0, v);
          stage_row_bf16(buf, lane, 4, w);
          __nv_bfloat16* my_dst = row_ok ? rope_dst(ep, row, col0, ro) : nullptr;
          const unsigned long long dst_bits = reinterpret_cast<unsigned long long>(my_dst);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + (lane >> 3), ch = lane & 7;
            const unsigned long long db = __shfl_sync(0xffffffffu, dst_bits, r);
            const uint4 u = *reinterpret_cast<const uint4*>(buf + r * 128 + ((ch ^ (r & 7)) << 4));
            if (db) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(db) + (ch < 4 ? ch * 8 : 64 + (ch - 4) * 8)) = u;
          }
          __syncwarp();
        } else {
          epilogue_tile<BLOCK_N, EPI>(ep, d.N, taddr, half, row, row_ok, n0, first_split, atomic, rstd, ss, part_tile);
        }
      }
      // accumulator drained: hand the TMEM stage back before any slow tail work
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;

      if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_STORE_F32) {
        if (ep.a_next) {
          if (!atomic) {
            const int dr = (EPI == EPI_STORE_F32 && ep.row_map && row_ok) ? ep.row_map[row] : row;
            if (row_ok && dr >= 0) atomicAdd(ep.ss_next + dr, ss);
          } else {
            // split-K: the CTA that retires the last split of this tile finishes the row statistics
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            if (threadIdx.x == 64) {
              const int old = atomicAdd(ep.tile_counters + tc.rem, 1);
              const int last = (old == sc.splits - 1);
              if (last) ep.tile_counters[tc.rem] = 0;
              *fin_flag = last;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            const bool last = *fin_flag != 0;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");  // flag consumed before the next tile may rewrite it
            if (last) {
              // coalesced pass over the finished tile: one warp per row, 8 rows in flight per warp
              __threadfence();
              const int m0 = m_idx * BLOCK_M + quarter * 32;
#pragma unroll 1
              for (int cb = 0; cb < BLOCK_N / 128; ++cb) {
                const int col = n0 + cb * 128 + lane * 4;
                const float4 g = __ldg(reinterpret_cast<const float4*>(ep.gamma_next + col));
#pragma unroll 1
                for (int rb = half * 16; rb < half * 16 + 16; rb += 8) {
                  float4 xv[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const int r = m0 + rb + i;
                    xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < M)
                      xv[i] = __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.out) +
                                                                     (size_t)r * ep.ldo + col));
                  }
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const int r = m0 + rb + i;
                    const float q = warp_sum((xv[i].x * xv[i].x + xv[i].y * xv[i].y) +
                                             (xv[i].z * xv[i].z + xv[i].w * xv[i].w));
                    if (r < M) {
                      uint2 u;
                      u.x = pack_bf16(xv[i].x * g.x, xv[i].y * g.y);
                      u.y = pack_bf16(xv[i].z * g.z, xv[i].w * g.w);
                      *reinterpret_cast<uint2*>(ep.a_next + (size_t)r * ep.lda_next + col) = u;
                      if (lane == 0) atomicAdd(ep.ss_next + r, q);
                    }
                  }
                }
              }
            }
          }
        }
      }
    }
    trace_stamp(tr, 5);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (warp == 2) trace_stamp(tr, 6);
  if (threadIdx.x == 0) trace_max(d.trace, 7);
}

// ------------------------------------------------------------------------------------------
// Coalesced epilogue of the pair kernel: accumulator rows are transposed through a 128B-swizzled smem tile
// (conflict-free 16-byte stores, one row per lane) and leave the SM as ONE TMA transaction per 32 x 128-byte
// tile: a bulk tensor store for bf16 outputs, a bulk tensor reduce-add (fp32 add at L2) for the residual
// streams. Row-per-thread global accesses cost 32 L1 wavefronts per instruction and paced the tile loop.
// Rows past the live M inside the last tile are written too (never read; TMA clips at the tensor bounds).
// ------------------------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ constexpr bool epi_has_tma_out() {
  return EPI == EPI_STORE_BF16 || EPI == EPI_GELU_BF16 || EPI == EPI_SILU_BF16 || EPI == EPI_RESID_F32 ||
         EPI == EPI_SWIGLU_BF16;
}

template <int BN, int EPI, int PARTS, int NBUF>
__device__ __forceinline__ void epilogue_tile_tma(const EpiParams& ep, const CUtensorMap* tmC, uint8_t* wbuf,
                                                  uint32_t& par, uint32_t taddr, int half, int lane, int row0, int n0,
                                                  float rstd, bool first_split, const CUtensorMap* tmP = nullptr,
                                                  int part_row0 = -1, uint8_t* ring = nullptr, long long* tr = nullptr) {
  auto acquire = [&]() -> uint8_t* {
    // the staging tile about to be reused (two units ago with NBUF == 2, the previous one otherwise) has been read out
    if (lane == 0) bulk_wait_read<NBUF - 1>();
    __syncwarp();
    uint8_t* b = wbuf + par * 4096;
    if (NBUF == 2) par ^= 1;
    return b;
  };
  auto scale_rows = [&](float (&v)[32]) {
    if (ep.row_ss) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= rstd;
    }
  };
  if constexpr (EPI == EPI_RESID_F32) {
    constexpr int U = BN / 32 / PARTS;
    // `ring` (last tile of this CTA: the operand ring is idle): every chunk gets its own 4 KB staging tile there, so the
    // U stores of the warp are issued back to back behind ONE proxy fence instead of each waiting for the TMA unit to
    // have drained the single staging tile -- the small-M projections, one tile per CTA, are paced by exactly that
    auto finish = [&](float (&v)[32], int c) -> bool {  // scale / bias / layer-scale one chunk; true: staged in the ring
      const int col0 = n0 + c * 32;
      scale_rows(v);
      if (ep.bias && first_split) add_vec32(v, ep.bias + col0);
      if (ep.scale) {
        const float4* sp = reinterpret_cast<const float4*>(ep.scale + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 sc = __ldg(sp + i);
          v[4 * i + 0] *= sc.x;
          v[4 * i + 1] *= sc.y;
          v[4 * i + 2] *= sc.z;
          v[4 * i + 3] *= sc.w;
        }
      }
      if (ring) {
        stage_row_f32(ring + (c - half * U) * 4096, lane, v);
        return true;
      }
      uint8_t* b = acquire();
      stage_row_f32(b, lane, v);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (part_row0 >= 0) tma_store_2d(tmP, b, col0, part_row0);  // deterministic split-K: partial tile to the workspace
        else tma_reduce_add_2d(tmC, b, col0, row0);
        bulk_commit();
      }
      return false;
    };
    static_assert(U % 2 == 0, "residual TMA epilogue: chunks are drained in pairs");
#pragma unroll 1
    for (int c = half * U; c < (half + 1) * U; c += 2) {
      float v0[32], v1[32];
      load_acc32x2(taddr + c * 32, taddr + c * 32 + 32, v0, v1);
      if (c == half * U) trace_stamp(tr, 8);
      finish(v0, c);
      finish(v1, c + 1);
      trace_stamp(tr, 9);
    }
    if (ring) {
      fence_proxy_async();
      __syncwarp();
      trace_stamp(tr, 10);
      if (lane == 0) {
#pragma unroll 1
        for (int c = half * U; c < (half + 1) * U; ++c) {
          const uint8_t* b = ring + (c - half * U) * 4096;
          if (part_row0 >= 0) tma_store_2d(tmP, b, n0 + c * 32, part_row0);
          else tma_reduce_add_2d(tmC, b, n0 + c * 32, row0);
        }
        bulk_commit();
      }
      trace_stamp(tr, 11);
    }
  } else if constexpr (EPI == EPI_SWIGLU_BF16) {
    constexpr int U = BN / 128 / PARTS;  // 128 accumulator columns -> 64 outputs -> one 128-byte row
    static_assert(U >= 1, "SwiGLU TMA epilogue: at least 128 accumulator columns per warp");
#pragma unroll 1
    for (int u = half * U; u < (half + 1) * U; ++u) {
      uint8_t* b = acquire();
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        float g[32], up[32];
        load_acc32x2(taddr + u * 128 + q * 64, taddr + u * 128 + q * 64 + 32, g, up);
        scale_rows(g);
        scale_rows(up);
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] = silu(g[j]) * up[j];
        stage_row_bf16(b, lane, q * 4, g);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, b, (n0 >> 1) + u * 64, row0);
        bulk_commit();
      }
    }
  } else {
    constexpr int U = BN / 64 / PARTS;
#pragma unroll 1
    for (int u = half * U; u < (half + 1) * U; ++u) {
      uint8_t* b = acquire();
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        const int col0 = n0 + u * 64 + q * 32;
        float v[32];
        load_acc32(taddr + u * 64 + q * 32, v);
        scale_rows(v);
        if constexpr (EPI == EPI_GELU_BF16) {
          // bias + exact GELU on pairs (packed FP32 instructions, see gelu_erf_fast2)
          const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ep.bias) b = __ldg(b4 + i);
            gelu_erf_fast2(v[4 * i], v[4 * i + 1], b.x, b.y);
            gelu_erf_fast2(v[4 * i + 2], v[4 * i + 3], b.z, b.w);
          }
        } else {
          if (ep.bias) add_vec32(v, ep.bias + col0);
        }
        if constexpr (EPI == EPI_SILU_BF16) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu(v[j]);
        }
        stage_row_bf16(b, lane, q * 4, v);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, b, n0 + u * 64, row0);
        bulk_commit();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// tcgen05 kernel, CTA pairs (cta_group::2): one 256(M) x 256(N) tile per pair. Each CTA stages its own
// 128 rows of A and HALF of the W tile (128 of the 256 N rows); the pair's tensor cores read both halves, so
// the L2->SM operand traffic per FLOP is 2/3 of the single-CTA 128x256 tile (L2 bandwidth, not the tensor
// pipe, bounds the large-M GEMMs of this path). The leader CTA (rank 0) issues the MMAs; its full barriers
// collect the TMA bytes of both CTAs; tcgen05.commit multicasts to the empty / accumulator-full barriers of both.
// ------------------------------------------------------------------------------------------
constexpr int PAIR_BLOCK_N = 256;
constexpr int PAIR_MAX_STAGES = 6;
constexpr uint32_t PAIR_A_BYTES = BLOCK_M * BLOCK_K * 2;           // this CTA's 128 rows of A
constexpr uint32_t PAIR_B_BYTES = (PAIR_BLOCK_N / 2) * BLOCK_K * 2;  // this CTA's half of the W tile
constexpr uint32_t PAIR_STAGE_BYTES = PAIR_A_BYTES + PAIR_B_BYTES;
// smem budget (227 KB): one 32-row x 128-byte output staging tile per epilogue warp + as many ring stages as fit
// (the q|k|v + RoPE epilogue transposes a unit as fp32 -- 32 rows x 256 bytes per warp -- and gives one ring stage for it)
constexpr int pair_stages(int nw, int epi) { return (nw == 16 || epi == EPI_QKV_ROPE) ? 5 : 6; }
constexpr uint32_t pair_warp_stage_bytes(int epi) { return epi == EPI_QKV_ROPE ? 8192 : 4096; }
constexpr uint32_t pair_out_bytes(int nw, int epi) { return nw * pair_warp_stage_bytes(epi); }
constexpr uint32_t pair_smem_bytes(int nw, int epi) {
  // (no alignment slack: the kernel declares its dynamic shared memory 1024-byte aligned and has no static shared memory.
  // 229,632 bytes at 6 stages + the 1 KB the system reserves per CTA leave 2.8 KB of the SM: the two row-per-CTA norm CTAs
  // an SM gets between the launches of the denoise chain fit beside a CTA of this kernel -- with the slack only one did, the
  // second had to wait for the projection to finish, and with it the launch of everything behind the norm)
  return pair_stages(nw, epi) * PAIR_STAGE_BYTES + pair_out_bytes(nw, epi) + 256;
}
// ---- half-SM variant (template parameter HALF): the launches of the action expert's denoise chain. Every CTA of such a
// launch owns ONE tile, and the launch is one hop of a latency-bound chain of ~2500 dependent launches. A CTA that
// needs a whole SM (225 KB of shared memory, 320 threads x 168 registers) cannot enter until a CTA of the previous launch
// has left its SM, and only then starts its ~2.8 us prologue (barriers, TMEM, cluster sync, weight prefetch): the in-situ
// stamps over all CTAs (profiles/r02s3e) put that prologue on the critical path of the o / gate-up / down launches,
// 2 - 2.8 us each. The half-SM variant fits twice on an SM -- 4 epilogue warps (192 threads), a short ring (the K loop of a
// CTA is 2 - 12 k-blocks), no dedicated staging tiles (the epilogue stages in the idle ring), ONE accumulator stage
// (BN TMEM columns) -- so the next launch's CTAs are resident, initialised and have their weights in flight while the
// current launch still runs.
constexpr int pair_half_stages(int bn) { return bn == 128 ? 4 : 3; }
constexpr uint32_t pair_half_slot_bytes(int bn) { return PAIR_A_BYTES + (bn / 2) * BLOCK_K * 2; }  // 24 KB | 32 KB
constexpr uint32_t pair_half_smem_bytes(int bn) { return pair_half_stages(bn) * pair_half_slot_bytes(bn) + 256; }

// tile order of the persistent loop: m fastest (the pairs of a wave share one W tile) or n fastest (share one A tile)
__device__ __forceinline__ void pair_tile(int t, int m_tiles, int n_tiles, int n_fast, int& m_idx, int& n_idx) {
  if (n_fast) {
    m_idx = t / n_tiles;
    n_idx = t - m_idx * n_tiles;
  } else {
    n_idx = t / m_tiles;
    m_idx = t - n_idx * m_tiles;
  }
}

// Fused tail of a split-K residual projection (see EpiParams::tail_ctr): in-kernel grid barrier, then this CTA's share of
// the residual update + RMSNorm rows. Called by every thread of every CTA after the CTA-wide (cluster-wide) barrier that
// follows the completion of its TMA stores. N <= 1024 columns (one float4 per thread of the first N / 4 threads).
__device__ __forceinline__ void pair_tail_resid_norm(const EpiParams& ep, const GemmDims& d, int M, int S, float* tail_red) {
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ep.tail_ctr) : "memory");
    unsigned int v;
    long long t0 = 0;
    while (true) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ep.tail_ctr) : "memory");
      if (v >= gridDim.x) break;
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 2000000000LL) {  // a CTA of the launch is not resident: give up, flag the call
        if (ep.tail_err) atomicExch(ep.tail_err, GEMM_ERR_TAIL_BARRIER);
        break;
      }
    }
  }
  __syncthreads();
  const int W = d.N, C4 = W >> 2;
  const int c = threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = (C4 + 31) >> 5;  // warps that own columns
  float* x = reinterpret_cast<float*>(ep.out);
  float4 gm = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C4) gm = __ldg(reinterpret_cast<const float4*>(ep.tail_gamma) + c);
  // two rows per round, all of their loads in flight at once (256 rows over 144 CTAs: one round)
  for (int row0 = blockIdx.x; row0 < M; row0 += 2 * gridDim.x) {
    const int rows[2] = {row0, row0 + (int)gridDim.x};
    float4 v[2];
    float q[2] = {0.f, 0.f};
    if (c < C4) {
      float4 pv[2][12];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const bool ok = rows[r] < M;
        const size_t off = (size_t)(ok ? rows[r] : row0) * ep.ldo + 4 * c;
        v[r] = *reinterpret_cast<const float4*>(x + off);
#pragma unroll
        for (int u = 0; u < 12; ++u) {
          pv[r][u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (u < S && ok) pv[r][u] = __ldcg(reinterpret_cast<const float4*>(ep.part + (size_t)u * ep.part_stride + off));
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (rows[r] >= M) continue;
        const size_t off = (size_t)rows[r] * ep.ldo + 4 * c;
        // four interleaved split groups, each summed in ascending split order from 0 and then added to x in group order:
        // the association of resid_norm_kernel (G = 4), so the residual stream is bit-identical with the unfused schedule
        float4 acc[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 12; ++u) {
          acc[u & 3].x += pv[r][u].x;
          acc[u & 3].y += pv[r][u].y;
          acc[u & 3].z += pv[r][u].z;
          acc[u & 3].w += pv[r][u].w;
        }
        for (int s0 = 12; s0 < S; s0 += 4) {  // (more than 12 splits: the rest in dependent batches of one per group)
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (s0 + u < S) {
              const float4 p = __ldcg(reinterpret_cast<const float4*>(ep.part + (size_t)(s0 + u) * ep.part_stride + off));
              acc[u].x += p.x, acc[u].y += p.y, acc[u].z += p.z, acc[u].w += p.w;
            }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) v[r].x += acc[g].x, v[r].y += acc[g].y, v[r].z += acc[g].z, v[r].w += acc[g].w;
        if (S > 0) *reinterpret_cast<float4*>(x + off) = v[r];
        q[r] = (v[r].x * v[r].x + v[r].y * v[r].y) + (v[r].z * v[r].z + v[r].w * v[r].w);
      }
    }
    q[0] = warp_sum(q[0]);
    q[1] = warp_sum(q[1]);
    __syncthreads();  // scratch free again
    if (lane == 0 && warp < 16) tail_red[warp] = q[0], tail_red[16 + warp] = q[1];
    __syncthreads();
    if (c < C4) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (rows[r] >= M) continue;
        float tot = 0.f;
        for (int w = 0; w < nw; ++w) tot += tail_red[16 * r + w];
        const float rstd = rsqrtf(tot / (float)W + ep.tail_eps);
        uint2 u;
        u.x = pack_bf16(v[r].x * rstd * gm.x, v[r].y * rstd * gm.y);
        u.y = pack_bf16(v[r].z * rstd * gm.z, v[r].w * rstd * gm.w);
        *reinterpret_cast<uint2*>(ep.tail_out + (size_t)rows[r] * ep.ldo + 4 * c) = u;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// q|k|v + RoPE epilogue of the pair kernel (256-column tiles, 8 epilogue warps): a warp owns 32 rows x one head (128
// columns) = two units {x[i0..i0+31], x[i0+64..i0+95]}, i0 in {0, 32}. The row-per-lane form (epilogue_tile) issues
// every RoPE-factor load and every q / KV-cache store as 32 scattered 16-byte pieces (32 L1 wavefronts per instruction,
// the factor table missing in the 20 KB of L1 this kernel leaves): 19 us per tile, three times the tile's mainloop, so the
// LLM-prefill q|k|v projection ran at 40 % of the tensor roofline. Here a unit is transposed as fp32 through the warp's
// 8 KB swizzled tile; in the transposed domain 8 lanes share a row: factor loads are 128-byte runs, stores 64-byte runs,
// bias / rotation in fp32 exactly as before. The accumulator stage is handed back once the second unit is staged.
// ------------------------------------------------------------------------------------------
template <class Release>
__device__ __forceinline__ void epilogue_rope_pair(const EpiParams& ep, uint8_t* tile, uint32_t taddr, int lane, int row,
                                                   bool row_ok, int col_head, float rstd, Release&& release) {
  const int head = col_head >> 7;
  const bool is_q = head < ep.n_q_heads;
  const bool is_k = !is_q && head < ep.n_q_heads + ep.n_kv_heads;
  const bool rot = is_q || is_k;
  const int ch = lane & 7, rsub = lane >> 3;
  // per-row operands (this call's plan: read after the dependent-launch wait the caller has passed)
  int pos = 0;
  unsigned long long dst_bits = 0;
  if (row_ok) {
    if (rot) pos = __ldg(ep.row_pos + row);
    __nv_bfloat16* dst;
    if (is_q) {
      dst = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldo + head * 128;
    } else {
      const int kvrow = __ldg(ep.row_kvrow + row);
      const int kvh = is_k ? head - ep.n_q_heads : head - ep.n_q_heads - ep.n_kv_heads;
      dst = (is_k ? ep.k_cache : ep.v_cache) + (size_t)kvh * ep.kv_head_stride + (size_t)kvrow * 128;
    }
    dst_bits = reinterpret_cast<unsigned long long>(dst);
  }
#pragma unroll 1
  for (int u = 0; u < 2; ++u) {
    const int i0 = u * 32;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), bw = bv;
    if (ep.bias) {
      bv = __ldg(reinterpret_cast<const float4*>(ep.bias + col_head + i0 + 4 * ch));
      bw = __ldg(reinterpret_cast<const float4*>(ep.bias + col_head + i0 + 64 + 4 * ch));
    }
    {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(taddr + i0, r0);
      tmem_ld_32x32(taddr + i0 + 64, r1);
      tmem_ld_wait();
      if (u == 1) release();
      uint8_t* rowp = tile + lane * 256;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int sw = (c ^ (lane & 7)) << 4;
        *reinterpret_cast<float4*>(rowp + sw) =
            make_float4(__uint_as_float(r0[4 * c]) * rstd, __uint_as_float(r0[4 * c + 1]) * rstd,
                        __uint_as_float(r0[4 * c + 2]) * rstd, __uint_as_float(r0[4 * c + 3]) * rstd);
        *reinterpret_cast<float4*>(rowp + 128 + sw) =
            make_float4(__uint_as_float(r1[4 * c]) * rstd, __uint_as_float(r1[4 * c + 1]) * rstd,
                        __uint_as_float(r1[4 * c + 2]) * rstd, __uint_as_float(r1[4 * c + 3]) * rstd);
      }
    }
    __syncwarp();
    // transposed domain: lane (rsub, ch) finishes pairs 4 ch .. 4 ch + 3 of rows it * 4 + rsub
#pragma unroll 1
    for (int hb = 0; hb < 2; ++hb) {  // two batches of four rows: the factor loads of a batch are in flight together
      float4 cs[4], sn[4];
      unsigned long long db[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (hb * 4 + i) * 4 + rsub;
        db[i] = __shfl_sync(0xffffffffu, dst_bits, r);
        const int p = __shfl_sync(0xffffffffu, pos, r);
        if (rot) {
          cs[i] = __ldg(reinterpret_cast<const float4*>(ep.cos_tab + (size_t)p * 64 + i0 + 4 * ch));
          sn[i] = __ldg(reinterpret_cast<const float4*>(ep.sin_tab + (size_t)p * 64 + i0 + 4 * ch));
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (hb * 4 + i) * 4 + rsub;
        const int sw = (ch ^ (r & 7)) << 4;
        float4 a = *reinterpret_cast<const float4*>(tile + r * 256 + sw);
        float4 b = *reinterpret_cast<const float4*>(tile + r * 256 + 128 + sw);
        a.x += bv.x, a.y += bv.y, a.z += bv.z, a.w += bv.w;
        b.x += bw.x, b.y += bw.y, b.z += bw.z, b.w += bw.w;
        if (rot) {  // x*cos + rot(x)*sin, rot(x)[i] = -x[i+64], rot(x)[i+64] = x[i]
          const float4 c = cs[i], s = sn[i];
          const float4 a0 = a, b0 = b;
          a.x = a0.x * c.x - b0.x * s.x, a.y = a0.y * c.y - b0.y * s.y;
          a.z = a0.z * c.z - b0.z * s.z, a.w = a0.w * c.w - b0.w * s.w;
          b.x = b0.x * c.x + a0.x * s.x, b.y = b0.y * c.y + a0.y * s.y;
          b.z = b0.z * c.z + a0.z * s.z, b.w = b0.w * c.w + a0.w * s.w;
        }
        if (db[i]) {
          __nv_bfloat16* d0 = reinterpret_cast<__nv_bfloat16*>(db[i]) + i0 + 4 * ch;
          *reinterpret_cast<uint2*>(d0) = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
          *reinterpret_cast<uint2*>(d0 + 64) = make_uint2(pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
        }
      }
    }
    __syncwarp();  // the tile is free for the next unit
  }
}

// ------------------------------------------------------------------------------------------
// Residual epilogue of the pair kernel that also produces the NEXT GEMM's operand (RMSNorm fused across the GEMM chain at
// large M: the o / down projections of the LLM prefill). The TMA reduce-add epilogue never sees x_new = x_old + acc -- the
// add happens at L2 --, so the norm needed a launch of its own that read x_new back (34 MB) and wrote bf16 (17 MB): 55
// launches of ~10 us per call. Here a warp's 32 rows x 128 columns go through its 4 KB tile one 32-column chunk at a time:
// accumulator rows in (row per lane), then in the transposed domain (8 lanes share a row, 128-byte runs) x_old is read
// (the loads are issued before the accumulator chunk is fetched), x_new and bf16(x_new * gamma_next) are written, and the
// squares are summed per row; the row's sum over this warp's 128 columns goes to ITS slot of ss_next[row][ss_slots]
// (plain store; the consumer adds the slots in order: bit-reproducible, unlike an atomicAdd per tile).
// ------------------------------------------------------------------------------------------
template <int BN, int PARTS, class Release>
__device__ __forceinline__ void epilogue_resid_rmw_pair(const EpiParams& ep, uint8_t* tile, uint32_t taddr, int half,
                                                        int lane, int row0, int M, int n0, Release&& release) {
  constexpr int U = BN / 32 / PARTS;
  const int ch = lane & 7, rsub = lane >> 3;
  float* x = reinterpret_cast<float*>(ep.out);
  float ssum[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) ssum[it] = 0.f;
#pragma unroll 1
  for (int c = half * U; c < (half + 1) * U; ++c) {
    const int col = n0 + c * 32 + 4 * ch;
    float4 xo[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = row0 + it * 4 + rsub;
      xo[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < M) xo[it] = __ldcg(reinterpret_cast<const float4*>(x + (size_t)r * ep.ldo + col));
    }
    float4 bs = make_float4(0.f, 0.f, 0.f, 0.f), sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (ep.bias) bs = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
    if (ep.scale) sc = __ldg(reinterpret_cast<const float4*>(ep.scale + col));
    const float4 gm = __ldg(reinterpret_cast<const float4*>(ep.gamma_next + col));
    {
      float v[32];
      load_acc32(taddr + c * 32, v);
      if (c == (half + 1) * U - 1) release();
      stage_row_f32(tile, lane, v);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rl = it * 4 + rsub, r = row0 + rl;
      float4 v = *reinterpret_cast<const float4*>(tile + rl * 128 + ((ch ^ (rl & 7)) << 4));
      v.x = (v.x + bs.x) * sc.x, v.y = (v.y + bs.y) * sc.y, v.z = (v.z + bs.z) * sc.z, v.w = (v.w + bs.w) * sc.w;
      float4 xn = xo[it];
      xn.x += v.x, xn.y += v.y, xn.z += v.z, xn.w += v.w;
      if (r < M) {
        *reinterpret_cast<float4*>(x + (size_t)r * ep.ldo + col) = xn;
        *reinterpret_cast<uint2*>(ep.a_next + (size_t)r * ep.lda_next + col) =
            make_uint2(pack_bf16(xn.x * gm.x, xn.y * gm.y), pack_bf16(xn.z * gm.z, xn.w * gm.w));
        ssum[it] += (xn.x * xn.x + xn.y * xn.y) + (xn.z * xn.z + xn.w * xn.w);
      }
    }
    __syncwarp();
  }
  const int slot = (n0 / BN) * PARTS + half;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    float s = ssum[it];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    const int r = row0 + it * 4 + rsub;
    if (ch == 0 && r < M) ep.ss_next[(size_t)r * ep.ss_slots + slot] = s;
  }
}

template <int EPI>
constexpr int pair_epi_warps() { return EPI == EPI_GELU_BF16 ? 16 : 8; }  // the exact-GELU epilogue is issue-bound at 8

template <int EPI, int NW, int BN = PAIR_BLOCK_N, bool HALF = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * NW, HALF ? 2 : 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP, GemmDims d,
                         EpiParams ep) {
  static_assert(!HALF || (NW == 4 && (EPI == EPI_RESID_F32 || EPI == EPI_SWIGLU_BF16)), "half-SM variant: 4 epilogue warps");
  constexpr int STAGES = HALF ? pair_half_stages(BN) : pair_stages(NW, EPI);
  // (BN = 128: the ring keeps its 32 KB slots, the W half of a slot is simply half full; the half-SM variant packs them)
  constexpr uint32_t SLOT = HALF ? pair_half_slot_bytes(BN) : PAIR_STAGE_BYTES;
  constexpr int ACC = HALF ? 1 : 2;                                           // accumulator stages in TMEM
  constexpr uint32_t STAGE_TX = 2 * (PAIR_A_BYTES + (BN / 2) * BLOCK_K * 2);  // bytes both CTAs land per stage
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;  // 1024-byte aligned by declaration (128B-swizzle atoms); the kernel has no static shared memory
  uint8_t* out_stage = smem + STAGES * SLOT;  // [NW][NBUF][4096], 1024-byte aligned (half-SM variant: none)
  constexpr int PARTS = NW / 4;                           // column parts per TMEM lane quarter
  constexpr int NBUF = 1;                                 // staging tiles per epilogue warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + (HALF ? 0u : pair_out_bytes(NW, EPI)));
  uint64_t* full_bar = bars;                          // [STAGES]   used in the leader only
  uint64_t* empty_bar = bars + MAX_STAGES;            // [STAGES]   one set per CTA
  uint64_t* tfull_bar = bars + 2 * MAX_STAGES;        // [2]        one set per CTA
  uint64_t* tempty_bar = bars + 2 * MAX_STAGES + 2;   // [2]        used in the leader only (8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  long long* tr = (blockIdx.x == 0 && lane == 0) ? d.trace : nullptr;

  if (warp == 0 && lane == 0) {
    trace_stamp(tr, 0);
    trace_max(d.trace, 14);
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (d.tma_out) tma_prefetch_desc(&tmC);
    if (EPI == EPI_RESID_F32 && ep.part) tma_prefetch_desc(&tmP);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * NW);  // epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  // Both CTAs' barriers are initialised before any remote arrive / multicast commit. The TMEM allocation comes AFTER this
  // sync: the producer does not need it, and its first ring fill (the weights) is what a CTA that could only enter when the
  // previous launch of a dependent chain left its SM is waited for (prologue stamps over all CTAs, profiles/r02s3g: barriers
  // 0.3 us, TMEM 1.0, cluster sync 1.5, weights issued 2.9 - 3.8 us after entry with the allocation in front of the sync)
  cluster_sync_all();
  pdl_launch_dependents();
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, ACC * BN);
    tmem_relinquish_pair();
  }
  if (warp != 0) {  // the MMA and epilogue warps meet on the TMEM address; the producer is already filling the ring
    tc_fence_before();
    asm volatile("bar.sync 2, %0;" ::"n"(32 * (NW + 1)) : "memory");
    tc_fence_after();
  }
  const uint32_t tmem_base = warp != 0 ? *tmem_slot : 0u;
  // A device-resident row count belongs to this call's plan: it is read after the dependent-launch wait. With a static
  // M only the threads that touch activations wait (producer: after its weight prefetch; epilogue warps).
  int M = d.M;
  if (d.M_dev) {
    pdl_wait();
    M = min(*d.M_dev, d.M);
  }
  const int m_tiles = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int n_tiles = d.N / BN;
  const int k_blocks = d.K / BLOCK_K;
  const int base = m_tiles * n_tiles;
  // split-K (residual epilogue through TMA reduce-add / partial stores only): partial sums meet in L2 or in the workspace
  const int splits = (EPI == EPI_RESID_F32 && d.tma_out) ? pair_split_count(base, n_pairs, k_blocks, d.max_splits) : 1;
  const int kb_per_split = (k_blocks + splits - 1) / splits;
  const int total = base * splits;
  if constexpr (EPI == EPI_RESID_F32) {
    if (ep.part_splits && blockIdx.x == 0 && threadIdx.x == 64) {
      pdl_wait();  // the previous reader of the word has finished
      *ep.part_splits = (splits > 1 && ep.part) ? splits : 0;
    }
  }

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    // The whole warp walks the (tile, k-block) cursor; the FIRST ring fill is issued by one lane per stage -- a TMA issue
    // costs a thread 0.25 - 0.35 us at the head of a launch (prologue stamps, profiles/r02s3g: six stages took 1.4 - 2.2 us
    // from one thread), and a CTA of the denoise chain that could only enter when the previous launch left its SM is
    // waited for exactly there. The steady state stays with lane 0.
    {
      const uint32_t full0 = mapa_u32(smem_u32(full_bar), 0);  // the leader's full barriers
      int stage = 0;
      uint32_t phase = 0;
      // (tile, k-block) cursor of this pair
      int t = pair, kb = 0, kb1 = 0, row_a = 0, row_b = 0;
      bool have = t < total;
      auto open_tile = [&]() {
        int n_idx, m_idx;
        const int split = t / base;
        pair_tile(t - split * base, m_tiles, n_tiles, d.n_fast, m_idx, n_idx);
        row_a = (m_idx * 2 + (int)rank) * BLOCK_M;
        row_b = n_idx * BN + (int)rank * (BN / 2);
        kb = split * kb_per_split;
        kb1 = min(k_blocks, kb + kb_per_split);
      };
      auto advance = [&]() {
        if (++kb >= kb1) {
          t += n_pairs;
          have = t < total;
          if (have) open_tile();
        }
      };
      if (have) open_tile();
      if (!d.M_dev) {
        // Weights do not depend on the previous kernel: the W halves of the first ring fill go out before the
        // dependent-launch wait, the A halves (activations written by the previous kernel) after it.
        int my_kb = 0, my_row_a = 0, my_row_b = 0;
        bool mine = false;
        int npre = 0;
        while (have && npre < STAGES) {
          if (lane == npre) mine = true, my_kb = kb, my_row_a = row_a, my_row_b = row_b;
          ++npre;
          advance();
        }
        if (mine) {
          if (leader) mbar_arrive_expect_tx(&full_bar[lane], STAGE_TX);
          tma_load_2d_pair(smem + lane * SLOT + PAIR_A_BYTES, &tmB, full0 + lane * 8, my_kb * BLOCK_K, my_row_b);
        }
        if (lane == 0) trace_max(d.trace, 15);
        pdl_wait();
        trace_stamp(tr, 1);
        if (mine) tma_load_2d_pair(smem + lane * SLOT, &tmA, full0 + lane * 8, my_kb * BLOCK_K, my_row_a);
        stage = npre;
        if (stage == STAGES) {
          stage = 0;
          phase = 1;
        }
      } else {
        trace_stamp(tr, 1);
      }
      if (lane < 2) {
        // steady state: lane 0 loads the A half, lane 1 the W half of a stage -- ONE instruction with two active lanes
        // instead of two issues in a row from one thread
        const CUtensorMap* my_map = lane ? &tmB : &tmA;
        const uint32_t my_off = lane ? PAIR_A_BYTES : 0u;
        while (have) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader && lane == 0) mbar_arrive_expect_tx(&full_bar[stage], STAGE_TX);
          tma_load_2d_pair(smem + stage * SLOT + my_off, my_map, full0 + stage * 8, kb * BLOCK_K, lane ? row_b : row_a);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
          advance();
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = pair; t < total; t += n_pairs) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        const int split = t / base;
        const int kb0 = split * kb_per_split, kb1 = min(k_blocks, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (t == pair && kb == kb0) trace_stamp(tr, 2);
          const uint32_t sa = smem_u32(smem + stage * SLOT);
          const uint64_t da = make_umma_desc_sw128(sa);
          const uint64_t db = make_umma_desc_sw128(sa + PAIR_A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            umma_bf16_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit_pair(&empty_bar[stage], 0x3);  // frees this smem slot in both CTAs
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_pair(&tfull_bar[as], 0x3);  // accumulators (both CTAs' TMEM) ready
        trace_stamp(tr, 3);
        if (ACC == 2) as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ================================ epilogue warps (both CTAs) ================================
    if (!d.M_dev) pdl_wait();  // residual rows / activations of the previous kernel
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t tempty0 = mapa_u32(smem_u32(tempty_bar), 0);
    // (half-SM variant: one tile per CTA, so the operand ring is idle whenever the epilogue runs -- it stages there)
    uint8_t* wbuf = HALF ? smem + (warp - 2) * 4096 : out_stage + (warp - 2) * (NBUF * (int)pair_warp_stage_bytes(EPI));
    uint32_t out_par = 0;
    int as = 0;
    uint32_t aphase = 0;
    if (warp != 2) tr = nullptr;
    int pend_m = -1;  // row block whose reduce-adds this warp has issued but not yet reported (EpiParams::done_ctr)
    auto report_done = [&]() {
      if constexpr (EPI == EPI_RESID_F32) {
        if (ep.done_ctr && pend_m >= 0) {
          if (lane == 0) {
            bulk_wait<0>();        // the warp's reduce-adds have been performed ...
            fence_proxy_async();   // ... and are ordered (async proxy -> generic) before the release below
            asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ep.done_ctr + pend_m) : "memory");
          }
          pend_m = -1;
        }
      }
    };
    for (int t = pair; t < total; t += n_pairs) {
      int n_idx, m_idx;
      const int split = t / base;
      pair_tile(t - split * base, m_tiles, n_tiles, d.n_fast, m_idx, n_idx);
      const int row = (m_idx * 2 + (int)rank) * BLOCK_M + quarter * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
      const int n0 = n_idx * BN;
      float rstd = 1.f;
      if (ep.row_ss && row_ok) {
        float sq;
        if (ep.ss_slots > 0) {  // partial sums of the producer's column halves, added in slot order
          sq = 0.f;
          for (int j = 0; j < ep.ss_slots; ++j) sq += __ldcg(ep.row_ss + (size_t)row * ep.ss_slots + j);
        } else {
          sq = __ldcg(ep.row_ss + row);
        }
        rstd = rsqrtf(sq * ep.inv_width + ep.eps);
      }
      float ss = 0.f;
      int part_row0 = -1;
      if (EPI == EPI_RESID_F32 && splits > 1 && ep.part) part_row0 = split * (int)(ep.part_stride / ep.ldo) + (row - lane);
      {
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        if (t == pair) trace_stamp(tr, 4);
        report_done();  // the previous tile's stores have had this tile's mainloop to land
        bool done = false;
        if constexpr (epi_has_tma_out<EPI>()) {
          if (d.tma_out) {
            // last tile of this pair: the operand ring (STAGES x 32 KB) is idle once the accumulator is complete
            uint8_t* ring = nullptr;
            if (EPI == EPI_RESID_F32 && t + n_pairs >= total &&
                (BN / 32 / PARTS) * 4096 * NW <= STAGES * (int)SLOT)
              ring = smem + (warp - 2) * ((BN / 32 / PARTS) * 4096);
            epilogue_tile_tma<BN, EPI, PARTS, NBUF>(ep, &tmC, wbuf, out_par, taddr, half, lane, row - lane, n0, rstd,
                                                    split == 0, &tmP, part_row0, ring, t == pair ? tr : nullptr);
            done = true;
            pend_m = m_idx;
          }
        }
        bool released = false;
        if constexpr (EPI == EPI_QKV_ROPE && BN == 256 && NW == 8) {
          epilogue_rope_pair(ep, wbuf, taddr + half * 128, lane, row, row_ok, n0 + half * 128, rstd, [&]() {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty0 + as * 8);
          });
          done = released = true;
        }
        if constexpr (EPI == EPI_RESID_F32 && BN == 256 && !HALF) {
          if (!done && ep.a_next && ep.ss_slots > 0) {
            epilogue_resid_rmw_pair<BN, PARTS>(ep, wbuf, taddr, half, lane, row - lane, M, n0, [&]() {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(tempty0 + as * 8);
            });
            done = released = true;
          }
        }
        if (!done) epilogue_tile<BN, EPI, PARTS>(ep, d.N, taddr, half, row, row_ok, n0, true, false, rstd, ss);
        if (released) {
          if (t == pair) trace_stamp(tr, 12);
          if (ACC == 2) as ^= 1;
          if (as == 0) aphase ^= 1;
          continue;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (t == pair) trace_stamp(tr, 12);
      if (lane == 0) mbar_arrive_cluster(tempty0 + as * 8);  // the leader's MMA thread owns both accumulators
      if (t == pair) trace_stamp(tr, 13);
      if (ACC == 2) as ^= 1;
      if (as == 0) aphase ^= 1;
      if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_STORE_F32) {
        if (ep.a_next) {
          const int dr = (EPI == EPI_STORE_F32 && ep.row_map && row_ok) ? ep.row_map[row] : row;
          if (row_ok && dr >= 0) atomicAdd(ep.ss_next + dr, ss);
        }
      }
    }
    report_done();  // the last tile of this warp
  }

  if (warp == 2) trace_stamp(tr, 5);
  if (warp >= 2 && lane == 0) bulk_wait<0>();  // this warp's TMA stores have landed
  if constexpr (EPI == EPI_RESID_F32) {
    // (the last tile of an epilogue warp; `pend_m` lives in the epilogue branch above, so the report is made there)
  }
  if constexpr (EPI == EPI_RESID_F32 && BN == 128) {
    if (ep.tail_ctr) fence_proxy_async();  // ... and are ordered before the release below (async proxy -> generic)
  }
  // neither CTA may exit (or free TMEM) while its peer can still signal its barriers or read its smem
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, ACC * BN);
  if (warp == 2) trace_stamp(tr, 6);
  if (threadIdx.x == 0) trace_max(d.trace, 7);
  if constexpr (EPI == EPI_RESID_F32 && BN == 128) {
    if (ep.tail_ctr) pair_tail_resid_norm(ep, d, M, (splits > 1 && ep.part) ? splits : 0, reinterpret_cast<float*>(smem));
  }
}

// ------------------------------------------------------------------------------------------
// CUDA-core cross-check kernel (tests / bring-up only): same epilogue code, plain pointers.
// One thread = one row x one 32-column chunk (plus its partner chunk where the epilogue needs one).
// ------------------------------------------------------------------------------------------
template <int EPI>
__global__ void gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ W, GemmDims d,
                                 EpiParams ep) {
  pdl_launch_dependents();
  pdl_wait();
  const int M = d.M_dev ? min(*d.M_dev, d.M) : d.M;
  const int row = blockIdx.y * blockDim.y + threadIdx.y;
  const int unit = blockIdx.x * blockDim.x + threadIdx.x;
  int col0, colw;
  if (EPI == EPI_SWIGLU_BF16) {
    col0 = unit * 64;
    colw = col0 + 32;
  } else if (EPI == EPI_QKV_ROPE) {
    col0 = (unit >> 1) * 128 + (unit & 1) * 32;
    colw = col0 + 64;
  } else {
    col0 = unit * 32;
    colw = col0;
  }
  if (col0 >= d.N) return;
  const bool row_ok = row < M;
  float v[32], w[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = w[j] = 0.f;
  if (row_ok) {
    const __nv_bfloat16* a = A + (size_t)row * d.K;
    for (int k = 0; k < d.K; ++k) {
      const float av = __bfloat162float(a[k]);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += av * __bfloat162float(W[(size_t)(col0 + j) * d.K + k]);
      if (EPI == EPI_SWIGLU_BF16 || EPI == EPI_QKV_ROPE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) w[j] += av * __bfloat162float(W[(size_t)(colw + j) * d.K + k]);
      }
    }
  }
  float rstd = 1.f, ss = 0.f;
  if (ep.row_ss && row_ok) rstd = rsqrtf(ep.row_ss[row] * ep.inv_width + ep.eps);
  epi_chunk<EPI>(ep, d.N, row, row_ok, col0, v, w, true, false, rstd, ss);
  if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_STORE_F32) {
    if (ep.a_next && row_ok) {
      const int dr = (EPI == EPI_STORE_F32 && ep.row_map) ? ep.row_map[row] : row;
      if (dr >= 0) atomicAdd(ep.ss_next + dr, ss);
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

template <int BLOCK_N, int EPI>
int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmDims d, const EpiParams& ep, int grid,
               cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  static unsigned long long attr_devs = 0;  // devices this kernel's attributes were set on
  auto kern = gemm_tcgen05_kernel<BLOCK_N, EPI>;
  if (vl_first_on_device(attr_devs)) {
    VL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Cfg::smem_bytes(Cfg::default_stages(EPI), EPI)));
  }
  if (d.stages <= 0 || d.stages > Cfg::default_stages(EPI)) d.stages = Cfg::default_stages(EPI);
  if (d.stages < 2) d.stages = 2;
  VL_CUDA_CHECK(launch_kernel(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::smem_bytes(d.stages, EPI), stream, tmA, tmB, d, ep));
  return 0;
}

template <int EPI, int BN, int NW = pair_epi_warps<EPI>(), bool HALF = false>
int launch_pair_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tc, const CUtensorMap& tp, const GemmDims& d,
                   const EpiParams& ep, cudaStream_t stream) {
  static unsigned long long attr_devs = 0;  // devices this kernel's attributes were set on
  auto kern = gemm_tcgen05_pair_kernel<EPI, NW, BN, HALF>;
  constexpr uint32_t smem_bytes = HALF ? pair_half_smem_bytes(BN) : pair_smem_bytes(NW, EPI);
  if (vl_first_on_device(attr_devs)) {
    VL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  }
  long long tiles = (long long)((d.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * (d.N / BN);
  if (EPI == EPI_RESID_F32 && d.max_splits > 1) {
    VL_CHECK(d.tma_out, "gemm: split-K on the pair kernel needs the TMA reduce-add epilogue (output tensor map)");
    tiles *= d.max_splits;
  }
  const int pairs = (int)std::min<long long>(tiles, gemm_num_sms() / 2);
  // __cluster_dims__(2,1,1) is part of the kernel: a plain launch (with the optional PDL attribute) suffices
  VL_CHECK(!ep.done_ctr || (EPI == EPI_RESID_F32 && d.tma_out && d.max_splits <= 1 && !HALF && !ep.tail_ctr),
           "gemm: row-block completion counters need the one-split residual epilogue of the pair kernel through TMA");
  if (HALF) {
    VL_CHECK(d.tma_out && !d.M_dev && tiles <= gemm_num_sms() / 2 && !ep.tail_ctr && !ep.row_ss,
             "gemm: the half-SM pair kernel takes static-M launches of at most one tile per CTA pair through the TMA epilogue");
  }
  VL_CUDA_CHECK(launch_kernel(kern, dim3(2 * pairs), dim3(64 + 32 * NW), smem_bytes, stream, tmA, tmB, tc, tp, d, ep));
  return 0;
}

template <int EPI>
int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmC, GemmDims d, const EpiParams& ep,
                cudaStream_t stream, const CUtensorMap* tmP) {
  d.tma_out = (tmC != nullptr && epi_has_tma_out<EPI>() && !ep.a_next) ? 1 : 0;
  const CUtensorMap& tc = d.tma_out ? *tmC : tmA;
  VL_CHECK(!(EPI == EPI_RESID_F32 && ep.part) || tmP != nullptr, "gemm: pair kernel with split-K partials needs the workspace tensor map");
  const CUtensorMap& tp = tmP ? *tmP : tmA;
  if constexpr (EPI == EPI_RESID_F32) {
    if (d.pair_bn == 128 && d.half_sm) return launch_pair_bn<EPI, 128, 4, true>(tmA, tmB, tc, tp, d, ep, stream);
    if (d.pair_bn == 128) return launch_pair_bn<EPI, 128>(tmA, tmB, tc, tp, d, ep, stream);
  }
  if constexpr (EPI == EPI_SWIGLU_BF16) {
    if (d.pair_bn == 256 && d.half_sm) return launch_pair_bn<EPI, PAIR_BLOCK_N, 4, true>(tmA, tmB, tc, tp, d, ep, stream);
  }
  VL_CHECK(!d.half_sm, "gemm: the half-SM pair kernel exists for the residual (128-column tiles) and SwiGLU epilogues");
  VL_CHECK(d.pair_bn == 256, "gemm: 128-column pair tiles exist for the residual epilogue only");
  if constexpr (EPI == EPI_GELU_BF16) {
    if (d.gelu_warps == 8) return launch_pair_bn<EPI, PAIR_BLOCK_N, 8>(tmA, tmB, tc, tp, d, ep, stream);
  }
  return launch_pair_bn<EPI, PAIR_BLOCK_N>(tmA, tmB, tc, tp, d, ep, stream);
}

template <int EPI>
int launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmC, const GemmDims& d,
              const EpiParams& ep, int block_n, int grid, cudaStream_t stream, const CUtensorMap* tmP) {
  if (d.pair) return launch_pair<EPI>(tmA, tmB, tmC, d, ep, stream, tmP);
  if (block_n == 128) return launch_one<128, EPI>(tmA, tmB, d, ep, grid, stream);
  if constexpr (EPI != EPI_QKV_ROPE) {
    if (block_n == 256) return launch_one<256, EPI>(tmA, tmB, d, ep, grid, stream);
  }
  set_last_error("gemm_launch: unsupported block_n");
  return -2;
}

}  // namespace

int gemm_pair_split_count(const GemmDims& d, bool with_partials) {
  const int bn = d.pair_bn;
  const int base = ((d.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * (d.N / bn);
  long long tiles = base;
  if (d.max_splits > 1) tiles *= d.max_splits;
  const int n_pairs = (int)std::min<long long>(tiles, gemm_num_sms() / 2);  // launch_pair_bn
  const int splits = pair_split_count(base, n_pairs, d.K / BLOCK_K, d.max_splits);
  return (splits > 1 && with_partials) ? splits : 0;
}

int gemm_num_sms() {
  static int sms[64] = {0};  // per device ordinal
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = sms[dev & 63];
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  VL_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  VL_CHECK(cols % 8 == 0, "TMA needs a 16-byte aligned row pitch");
  VL_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA needs a 16-byte aligned base");
  VL_CHECK(box_rows <= 256, "TMA box rows <= 256");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

int make_tmap_out(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, int elem_bytes) {
  PFN_encodeTiled fn = get_encode_fn();
  VL_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  VL_CHECK(elem_bytes == 2 || elem_bytes == 4, "output tensor map: bf16 or f32");
  VL_CHECK((cols * elem_bytes) % 16 == 0 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA alignment");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), 32};  // 128-byte rows, 32 rows = one epilogue warp
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (output) failed with CUresult " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

int gemm_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDims& d, const EpiParams& ep, int block_n,
                int epi, cudaStream_t stream, const CUtensorMap* tmC, const CUtensorMap* tmP) {
  VL_CHECK(!ep.part || (epi == EPI_RESID_F32 && ep.part_stride > 0 && ep.part_stride % ((long long)256 * ep.ldo) == 0 &&
                        ep.part_stride >= (long long)d.M * ep.ldo && !ep.a_next),
           "gemm: split-K partials need the residual epilogue and a stride of whole 256-row blocks covering M");
  VL_CHECK(d.K % BLOCK_K == 0 && d.K > 0, "gemm: K must be a positive multiple of 64");
  VL_CHECK(block_n == 128 || block_n == 256, "gemm: block_n must be 128 or 256");
  VL_CHECK(!d.pair || (block_n == 128 && (d.pair_bn == 256 || d.pair_bn == 128) && d.N % d.pair_bn == 0),
           "gemm: pair mode needs the W map with pair_bn / 2 row boxes, N % pair_bn == 0");
  VL_CHECK(d.N % block_n == 0, "gemm: N must be a multiple of block_n");
  VL_CHECK(ep.out != nullptr, "gemm: null output");
  VL_CHECK(!ep.done_ctr || d.pair, "gemm: row-block completion counters exist in the pair kernel only");
  VL_CHECK(!ep.a_next || (ep.gamma_next && ep.ss_next && ep.lda_next > 0), "gemm: incomplete fused-norm outputs");
  VL_CHECK(!(ep.a_next && epi == EPI_RESID_F32 && d.max_splits > 1) || ep.tile_counters,
           "gemm: split-K with fused-norm outputs needs tile counters");
  if (d.M <= 0) return 0;
  const int m_tiles = (d.M + BLOCK_M - 1) / BLOCK_M;
  long long tiles = (long long)m_tiles * (d.N / block_n);
  if (epi == EPI_RESID_F32 && d.max_splits > 1) tiles *= d.max_splits;
  const int grid = (int)std::min<long long>(tiles, gemm_num_sms());
  switch (epi) {
    case EPI_STORE_BF16: return launch_bn<EPI_STORE_BF16>(tmA, tmB, tmC, d, ep, block_n, grid, stream, tmP);
    case EPI_GELU_BF16: return launch_bn<EPI_GELU_BF16>(tmA, tmB, tmC, d, ep, block_n, grid, stream, tmP);
    case EPI_SILU_BF16: return launch_bn<EPI_SILU_BF16>(tmA, tmB, tmC, d, ep, block_n, grid, stream, tmP);
    case EPI_RESID_F32: return launch_bn<EPI_RESID_F32>(tmA, tmB, tmC, d, ep, block_n, grid, stream, tmP);
    case EPI_SWIGLU_BF16: return launch_bn<EPI_SWIGLU_BF16>(tmA, tmB, tmC, d, ep, block_n, grid, stream, tmP);
    case EPI_QKV_ROPE: return launch_bn<EPI_QKV_ROPE>(tmA, tmB, tmC, d, ep, block_n, grid, stream, tmP);
    case EPI_PATCH_EMBED: return launch_bn<EPI_PATCH_EMBED>(tmA, tmB, tmC, d, ep, block_n, grid, stream, tmP);
    case EPI_STORE_F32: return launch_bn<EPI_STORE_F32>(tmA, tmB, tmC, d, ep, block_n, grid, stream, tmP);
    default: set_last_error("gemm_launch: unknown epilogue"); return -2;
  }
}

template <int EPI>
static int launch_simt_one(const __nv_bfloat16* A, const __nv_bfloat16* W, const GemmDims& d, const EpiParams& ep,
                           cudaStream_t stream) {
  int units = d.N / 32;
  if (EPI == EPI_SWIGLU_BF16) units = d.N / 64;
  if (EPI == EPI_QKV_ROPE) units = d.N / 64;
  dim3 block(32, 4);
  dim3 grid((units + 31) / 32, (d.M + 3) / 4);
  VL_CUDA_CHECK(launch_kernel(gemm_simt_kernel<EPI>, grid, block, 0, stream, A, W, d, ep));
  return 0;
}

int gemm_launch_simt(const __nv_bfloat16* A, const __nv_bfloat16* W, const GemmDims& d, const EpiParams& ep, int epi,
                     cudaStream_t stream) {
  VL_CHECK(d.N % 128 == 0, "gemm_simt: N must be a multiple of 128");
  if (d.M <= 0) return 0;
  switch (epi) {
    case EPI_STORE_BF16: return launch_simt_one<EPI_STORE_BF16>(A, W, d, ep, stream);
    case EPI_GELU_BF16: return launch_simt_one<EPI_GELU_BF16>(A, W, d, ep, stream);
    case EPI_SILU_BF16: return launch_simt_one<EPI_SILU_BF16>(A, W, d, ep, stream);
    case EPI_RESID_F32: return launch_simt_one<EPI_RESID_F32>(A, W, d, ep, stream);
    case EPI_SWIGLU_BF16: return launch_simt_one<EPI_SWIGLU_BF16>(A, W, d, ep, stream);
    case EPI_QKV_ROPE: return launch_simt_one<EPI_QKV_ROPE>(A, W, d, ep, stream);
    case EPI_PATCH_EMBED: return launch_simt_one<EPI_PATCH_EMBED>(A, W, d, ep, stream);
    case EPI_STORE_F32: return launch_simt_one<EPI_STORE_F32>(A, W, d, ep, stream);
    default: set_last_error("gemm_launch_simt: unknown epilogue"); return -2;
  }
}

}  // namespace vlaser
