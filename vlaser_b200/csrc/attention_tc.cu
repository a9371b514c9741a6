// InternViT self-attention on the 5th-generation tensor cores (modeling_intern_vit.py:210-227, _naive_attn; no mask):
// head_dim 64, N = 257 tokens per image (224 px: 256 patches + CLS). See attention.cuh (vit_attention_tc_launch); the
// 1025-token rows of 448 px tiles take vit_attention_tc_long_kernel further down.
//
// One persistent CTA per SM walks (image, head) items. An item is 2 query tiles of 128 rows + 1 left-over query row
// against 256 keys + 1 left-over key. Warp-specialised, everything on chip between the q|k|v projection and the output:
//   warp 0        TMA producer: Q (2 x [128 x 64]), K and V ([256 x 64]) of the item as 128B-swizzled boxes of the packed
//                 qkv buffer, plus the three 128-byte rows of the left-over token (bulk copies); 2-deep smem ring
//   warp 1        TMEM allocator + MMA issuer (one lane): S_t = Q_t K^T (tcgen05.mma, M 128, N 256, 4 k-steps, fp32
//                 accumulator in TMEM) and, once the softmax group has turned S_t into probabilities,
//                 O_t = P_t V (A operand read from TMEM, V as MN-major smem operand, 16 k-steps)
//   warps 2..5    softmax group of query tile 0: one thread per query row (TMEM lane). The 256 scores of the row fit in
//   warps 6..9    TMEM, so softmax is exact and single-pass -- row max, then exp2 / sum / bf16 pack written back over
//                 the scores (tcgen05.st); no online rescaling and no correction step. The left-over key is a rank-1
//                 term on CUDA cores (score from the row's q in smem, p_x * v_x added to O). O leaves through a swizzled
//                 smem tile (the dead Q tile) and ONE TMA store per tile.
//   warps 10..13  the left-over query row on CUDA cores (257 scores, block softmax, PV), beside the tensor-core pipeline
// TMEM (512 columns): tile t owns columns [256 t, 256 t + 256): S fp32 in all of them, then P bf16 in [0, 128) and
// O fp32 in [192, 256).
#include "attention.cuh"
#include "gemm.cuh"  // gemm_num_sms(), make_tmap_bf16

#include <mutex>

namespace vlaser {
namespace {

typedef CUresult (*PFN_encodeTiled_tc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled_tc get_encode_fn_tc() {
  static PFN_encodeTiled_tc fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled_tc>(q);
  });
  return fn;
}

constexpr int VT_THREADS = 448;
constexpr int VT_HD = 64;
constexpr int VT_TILE_BYTES = 128 * 128;             // 128 rows x 64 bf16
constexpr int VT_BUF_BYTES = 6 * VT_TILE_BYTES + 1024;  // Q0 Q1 | K (256 rows) | V (256 rows) | q_x k_x v_x (3 x 128 B), padded to
                                                        // 1 KB so that the second slot's swizzled tiles stay 1024-byte aligned
constexpr int VT_SMEM = 2 * VT_BUF_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 257 * 4 + 64 + 2 * 64 * 4 + 16 * 4;

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 128-byte global -> shared bulk copy signalled on an mbarrier
__device__ __forceinline__ void bulk_load_128(void* smem_dst, const void* gsrc, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// element offset helpers of a 128B-swizzled [rows][64 bf16] tile: 16-byte chunk c of row r
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ float dot8(const uint4& a, const uint4& b) {
  const uint32_t x[4] = {a.x, a.y, a.z, a.w}, y[4] = {b.x, b.y, b.z, b.w};
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) s += bf16_lo(x[e]) * bf16_lo(y[e]) + bf16_hi(x[e]) * bf16_hi(y[e]);
  return s;
}

__global__ void __launch_bounds__(VT_THREADS, 1)
vit_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmOut,
                        const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int images, int heads,
                        float scale, long long* trace) {
  constexpr int N = 257;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * VT_BUF_BYTES);
  uint64_t* full_bar = bars;          // [2] operands of ring slot b landed
  uint64_t* empty_bar = bars + 2;     // [2] slot b free again: MMA commit + 2 softmax groups + aux group
  uint64_t* s_full = bars + 4;        // [2] scores of tile t complete
  uint64_t* p_full = bars + 6;        // [2] probabilities of tile t written (4 warps)
  uint64_t* o_full = bars + 8;        // [2] P V of tile t complete
  uint64_t* t_empty = bars + 10;      // [2] TMEM region of tile t read out (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* aux_p = reinterpret_cast<float*>(bars + 32);  // [257 + pad] probabilities of the left-over query row
  float* aux_red = aux_p + 272;                        // [2][64] PV halves + [16] reductions

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* tr = (blockIdx.x == 0 && lane == 0) ? trace : nullptr;
  if (warp == 0 && lane == 0) {
    trace_stamp(tr, 0);
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 4);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&t_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();  // the q|k|v projection is complete
  if (warp == 0) trace_stamp(tr, 1);

  const int items = images * heads;
  const int ld = 3 * heads * VT_HD;          // row pitch of the packed q|k|v buffer (elements)
  const float sl2 = scale * 1.4426950408889634f;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int k = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {
        const int b = k & 1;
        const uint32_t ph = (k >> 1) & 1;
        const int img = item / heads, h = item - img * heads;
        const int r0 = img * N;
        uint8_t* buf = smem + b * VT_BUF_BYTES;
        mbar_wait(&empty_bar[b], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[b], 6 * VT_TILE_BYTES + 3 * 128);
        const int cq = h * VT_HD, ck = (heads + h) * VT_HD, cv = (2 * heads + h) * VT_HD;
        tma_load_2d(buf + 2 * VT_TILE_BYTES, &tmQKV, &full_bar[b], ck, r0);
        tma_load_2d(buf + 3 * VT_TILE_BYTES, &tmQKV, &full_bar[b], ck, r0 + 128);
        tma_load_2d(buf, &tmQKV, &full_bar[b], cq, r0);
        tma_load_2d(buf + VT_TILE_BYTES, &tmQKV, &full_bar[b], cq, r0 + 128);
        tma_load_2d(buf + 4 * VT_TILE_BYTES, &tmQKV, &full_bar[b], cv, r0);
        tma_load_2d(buf + 5 * VT_TILE_BYTES, &tmQKV, &full_bar[b], cv, r0 + 128);
        const __nv_bfloat16* xrow = qkv + (size_t)(r0 + 256) * ld;
        bulk_load_128(buf + 6 * VT_TILE_BYTES, xrow + cq, &full_bar[b]);
        bulk_load_128(buf + 6 * VT_TILE_BYTES + 128, xrow + ck, &full_bar[b]);
        bulk_load_128(buf + 6 * VT_TILE_BYTES + 256, xrow + cv, &full_bar[b]);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 256);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64) | (1u << 16);  // B (= V) is MN-major
      int k = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {
        const int b = k & 1;
        const uint32_t ph = (k >> 1) & 1, tp = k & 1;
        const uint32_t sb = smem_u32(smem + b * VT_BUF_BYTES);
        mbar_wait(&full_bar[b], ph);
        tc_fence_after();
        const uint64_t dk = make_umma_desc_sw128(sb + 2 * VT_TILE_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&t_empty[t], tp ^ 1);
          tc_fence_after();
          const uint64_t dq = make_umma_desc_sw128(sb + t * VT_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < VT_HD / 16; ++ks) umma_bf16(tmem_base + t * 256, dq + 2 * ks, dk + 2 * ks, idesc_s, ks > 0);
          umma_commit(&s_full[t]);
          if (t == 0 && k == 0) trace_stamp(tr, 2);
        }
        // V [256 keys][64] as the MN-major B operand: 8-key groups 1024 B apart, a k-step of 16 keys is 2048 B
        const uint64_t dv = make_umma_desc_sw128(sb + 4 * VT_TILE_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t], tp);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)
            umma_bf16_ts(tmem_base + t * 256 + 192, tmem_base + t * 256 + 8 * ks, dv + 128 * ks, idesc_o, ks > 0);
          umma_commit(&o_full[t]);
        }
        umma_commit(&empty_bar[b]);  // every MMA that reads this ring slot has retired
      }
      trace_stamp(tr, 3);
    }
  } else if (warp < 10) {
    // ================================ softmax groups (one per query tile) ================================
    const int t = (warp - 2) >> 2;               // query tile
    const int quarter = warp & 3;                // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;         // query row within the tile
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
    const int gtid = threadIdx.x - (2 + 4 * t) * 32;  // 0..127 within the group
    int k = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {
      const int b = k & 1;
      const uint32_t ph = (k >> 1) & 1, tp = k & 1;
      const int img = item / heads, h = item - img * heads;
      uint8_t* buf = smem + b * VT_BUF_BYTES;
      uint8_t* qt = buf + t * VT_TILE_BYTES;
      const uint8_t* xrows = buf + 6 * VT_TILE_BYTES;
      mbar_wait(&full_bar[b], ph);
      // score of the left-over key: this row's q (swizzled tile) against k_x, on CUDA cores while the MMA runs
      float sx = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        sx += dot8(*reinterpret_cast<const uint4*>(qt + sw_off(row, c)), *reinterpret_cast<const uint4*>(xrows + 128 + c * 16));
      mbar_wait(&s_full[t], tp);
      tc_fence_after();
      if (t == 0 && k == 0 && warp == 2) trace_stamp(tr, 4);
      // ---- pass 1: row max
      float mx = sx;
#pragma unroll 1
      for (int c = 0; c < 8; c += 2) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(taddr + c * 32, r0);
        tmem_ld_32x32(taddr + c * 32 + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[j]), __uint_as_float(r1[j])));
      }
      const float moff = mx * sl2;
      // ---- pass 2: p = exp2(s * sl2 - moff), row sum, bf16 pack written back over the scores
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; c += 2) {  // two chunks per TMEM round trip (~0.3 us each)
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(taddr + c * 32, r0);
        tmem_ld_32x32(taddr + c * 32 + 32, r1);
        tmem_ld_wait();
        uint32_t pk0[16], pk1[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(r0[2 * j]), sl2, -moff));
          const float p1 = ex2_approx(fmaf(__uint_as_float(r0[2 * j + 1]), sl2, -moff));
          const float p2 = ex2_approx(fmaf(__uint_as_float(r1[2 * j]), sl2, -moff));
          const float p3 = ex2_approx(fmaf(__uint_as_float(r1[2 * j + 1]), sl2, -moff));
          sum += (p0 + p1) + (p2 + p3);
          pk0[j] = pack_bf16(p0, p1);
          pk1[j] = pack_bf16(p2, p3);
        }
        tmem_st_32x16(taddr + c * 16, pk0);
        tmem_st_32x16(taddr + c * 16 + 16, pk1);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      const float px = ex2_approx(fmaf(sx, sl2, -moff));
      sum += px;
      const float inv = 1.f / sum;
      // the left-over key's value row, fetched while P V runs
      uint4 vx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) vx[c] = *reinterpret_cast<const uint4*>(xrows + 256 + c * 16);
      mbar_wait(&o_full[t], tp);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(taddr + 192, o0);
      tmem_ld_32x32(taddr + 224, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[t]);  // the tile's TMEM columns may take the next item's scores
      // ---- O = (P V + p_x v_x) / sum -> bf16 -> this row of the (dead) Q tile, swizzled as the TMA store expects
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t vw[4] = {vx[c].x, vx[c].y, vx[c].z, vx[c].w};
        uint32_t ow[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = c * 8 + 2 * e;
          const float a0 = __uint_as_float(j < 32 ? o0[j] : o1[j - 32]);
          const float a1 = __uint_as_float(j < 32 ? o0[j + 1] : o1[j - 31]);
          ow[e] = pack_bf16((a0 + px * bf16_lo(vw[e])) * inv, (a1 + px * bf16_hi(vw[e])) * inv);
        }
        *reinterpret_cast<uint4*>(qt + sw_off(row, c)) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
      fence_proxy_async();
      named_bar_sync(1 + t, 128);
      if (gtid == 0) {
        tma_store_2d(&tmOut, qt, h * VT_HD, img * N + t * 128);
        bulk_commit();
        bulk_wait_read<0>();            // the tile has been read out of shared memory
        mbar_arrive(&empty_bar[b]);
      }
    }
    if (gtid == 0) bulk_wait<0>();
    if (t == 0 && warp == 2) trace_stamp(tr, 5);
  } else {
    // ================================ left-over query row (token 256) on CUDA cores ================================
    const int atid = threadIdx.x - 10 * 32;  // 0..127
    const int awarp = atid >> 5;
    int k = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {
      const int b = k & 1;
      const uint32_t ph = (k >> 1) & 1;
      const int img = item / heads, h = item - img * heads;
      const uint8_t* buf = smem + b * VT_BUF_BYTES;
      const uint8_t* Ks = buf + 2 * VT_TILE_BYTES;
      const uint8_t* Vs = buf + 4 * VT_TILE_BYTES;
      const uint8_t* xrows = buf + 6 * VT_TILE_BYTES;
      mbar_wait(&full_bar[b], ph);
      uint4 qx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) qx[c] = *reinterpret_cast<const uint4*>(xrows + c * 16);
      float s0 = 0.f, s1 = 0.f, sx = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        s0 += dot8(qx[c], *reinterpret_cast<const uint4*>(Ks + sw_off(atid, c)));
        s1 += dot8(qx[c], *reinterpret_cast<const uint4*>(Ks + sw_off(atid + 128, c)));
        sx += dot8(qx[c], *reinterpret_cast<const uint4*>(xrows + 128 + c * 16));
      }
      float mx = warp_max(fmaxf(fmaxf(s0, s1), sx));
      if (lane == 0) aux_red[128 + awarp] = mx;
      named_bar_sync(3, 128);
      mx = fmaxf(fmaxf(aux_red[128], aux_red[129]), fmaxf(aux_red[130], aux_red[131]));
      const float moff = mx * sl2;
      const float p0 = ex2_approx(fmaf(s0, sl2, -moff)), p1 = ex2_approx(fmaf(s1, sl2, -moff));
      const float px = ex2_approx(fmaf(sx, sl2, -moff));
      aux_p[atid] = p0;
      aux_p[atid + 128] = p1;
      float sum = warp_sum(p0 + p1);
      if (lane == 0) aux_red[132 + awarp] = sum;
      named_bar_sync(3, 128);
      sum = (aux_red[132] + aux_red[133]) + (aux_red[134] + aux_red[135]) + px;
      // P V: thread (dim d, key parity)
      const int d = atid & 63, par = atid >> 6;
      float acc = 0.f;
#pragma unroll 8
      for (int j = par; j < 256; j += 2)
        acc += aux_p[j] * __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(Vs + sw_off(j, d >> 3) + (d & 7) * 2));
      aux_red[par * 64 + d] = acc;
      named_bar_sync(3, 128);
      if (atid < 64) {
        const float vxd = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(xrows + 256 + atid * 2));
        const float o = (aux_red[atid] + aux_red[64 + atid] + px * vxd) / sum;
        out[((size_t)img * N + 256) * (heads * VT_HD) + h * VT_HD + atid] = __float2bfloat16(o);
      }
      named_bar_sync(3, 128);  // every thread is done with the slot (and with the scratch) before it is released
      if (atid == 0) mbar_arrive(&empty_bar[b]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
  if (warp == 2) trace_stamp(tr, 6);
}


// ================================================================================================================
// InternViT self-attention of long token rows on the tensor cores: N = 1 + Nk tokens per image with Nk a multiple of 256
// up to 1024 (448 px tiles: 1024 patches + CLS; modeling_intern_vit.py:210-227, pizero_internvl.py:739-743 image_448).
// The Nk x Nk block runs here; the left-over token is the last row: as a KEY it is a rank-1 term on CUDA cores (like the
// 257-token kernel), as a QUERY it is vit_attention_last_row_kernel below.
// The scores of a query row no longer fit in TMEM, so softmax is exact in TWO passes instead of online: pass A computes
// S = Q K^T block by block only for the row maximum, pass B recomputes S, turns it into probabilities with the final
// maximum and accumulates O += P V in TMEM -- no rescaling of O, no correction step; Q K^T is a quarter of the MMA work
// per pass (head_dim 64) and the kernel is bound by the exponentials (MUFU), not by the tensor core.
// Work item = (image, head, pair of 128-row query tiles), handed out in contiguous ranges so that consecutive items of a
// CTA share the head's K, which stays resident in shared memory (128 KB at Nk = 1024); V streams through a 3-slot ring.
//   warp 0      TMA producer (K once per head, Q pair per item, V blocks of 128 keys)
//   warp 1      MMA issuer: pass A S_t = Q_t K_j^T with N = 256; pass B per 128 keys S_t (N = 128), then O_t += P_t V_j with P
//               read from TMEM and V as MN-major operand
//   warps 2-5 / 6-9  softmax group of query tile 0 / 1: one thread per row
// TMEM: tile t owns columns [256 t, 256 t + 256): pass A scores in all of them; pass B scores in [0, 128), P (bf16) over
// them in [0, 64), O in [192, 256).
// ================================================================================================================
constexpr int VL_THREADS = 320;
constexpr int VL_VSLOTS = 3;
constexpr int VL_MAX_KT = 8;  // resident K tiles of 128 keys
constexpr int VL_SMEM = (2 + VL_MAX_KT + VL_VSLOTS) * VT_TILE_BYTES + 1024 /*k_x v_x rows*/ + 1024 /*align*/ + 256 /*barriers*/;

__global__ void __launch_bounds__(VL_THREADS, 1)
vit_attention_tc_long_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmOut,
                             const __nv_bfloat16* __restrict__ qkv, int images, int heads, int N, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;                                   // [2 tiles][128][128 B]
  uint8_t* Ks = smem + 2 * VT_TILE_BYTES;               // [nk128 tiles][128][128 B]
  uint8_t* Vs = Ks + VL_MAX_KT * VT_TILE_BYTES;         // [3 slots][128][128 B]
  uint8_t* xrows = Vs + VL_VSLOTS * VT_TILE_BYTES;      // k_x, v_x (128 B each)
  uint64_t* bars = reinterpret_cast<uint64_t*>(xrows + 1024);
  uint64_t* q_full = bars;            // [1]
  uint64_t* q_empty = bars + 1;       // [1] both groups have stored their O tile
  uint64_t* k_full = bars + 2;        // [1]
  uint64_t* k_empty = bars + 3;       // [1] MMA commit + both groups
  uint64_t* v_full = bars + 4;        // [3]
  uint64_t* v_empty = bars + 7;       // [3]
  uint64_t* s_full = bars + 10;       // [2]
  uint64_t* s_free = bars + 12;       // [2] pass-A scores read (4 warps)
  uint64_t* p_full = bars + 14;       // [2] probabilities written (4 warps)
  uint64_t* o_full = bars + 16;       // [2]
  uint64_t* o_free = bars + 18;       // [2] O read out (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Nk = N - 1, nk128 = Nk >> 7, nk256 = Nk >> 8, qpairs = Nk >> 8;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmOut);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 3);
    for (int i = 0; i < VL_VSLOTS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();  // the q|k|v projection is complete

  // contiguous item ranges: items of one (image, head) stay on one CTA where possible
  const int items = images * heads * qpairs;
  const int per = (items + gridDim.x - 1) / gridDim.x;
  const int i0 = min(items, (int)blockIdx.x * per), i1 = min(items, i0 + per);
  const int ld = 3 * heads * VT_HD;
  const float sl2 = scale * 1.4426950408889634f;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int prev = -1, vc = 0;
      for (int item = i0, it = 0; item < i1; ++item, ++it) {
        const int hd = item / qpairs, qp = item - hd * qpairs;
        const int img = hd / heads, h = hd - img * heads;
        const int r0 = img * N;
        const int cq = h * VT_HD, ck = (heads + h) * VT_HD, cv = (2 * heads + h) * VT_HD;
        mbar_wait(k_empty, (it & 1) ^ 1);
        if (hd != prev) {
          mbar_arrive_expect_tx(k_full, (uint32_t)(nk128 * VT_TILE_BYTES + 256));
          for (int j = 0; j < nk128; ++j) tma_load_2d(Ks + j * VT_TILE_BYTES, &tmQKV, k_full, ck, r0 + j * 128);
          const __nv_bfloat16* xrow = qkv + (size_t)(r0 + Nk) * ld;
          bulk_load_128(xrows, xrow + ck, k_full);
          bulk_load_128(xrows + 128, xrow + cv, k_full);
          prev = hd;
        } else {
          mbar_arrive(k_full);  // the head's K is still resident
        }
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, 2 * VT_TILE_BYTES);
        tma_load_2d(Qs, &tmQKV, q_full, cq, r0 + qp * 256);
        tma_load_2d(Qs + VT_TILE_BYTES, &tmQKV, q_full, cq, r0 + qp * 256 + 128);
        for (int j = 0; j < nk128; ++j, ++vc) {
          const int slot = vc % VL_VSLOTS;
          mbar_wait(&v_empty[slot], ((vc / VL_VSLOTS) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[slot], VT_TILE_BYTES);
          tma_load_2d(Vs + slot * VT_TILE_BYTES, &tmQKV, &v_full[slot], cv, r0 + j * 128);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc_a = make_idesc_bf16(128, 256);
      constexpr uint32_t idesc_b = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64) | (1u << 16);  // B (= V) is MN-major
      const uint32_t qa = smem_u32(Qs), ka = smem_u32(Ks), va = smem_u32(Vs);
      uint32_t n_sfree[2] = {0, 0}, n_pfull[2] = {0, 0};
      int vc = 0;
      for (int item = i0, it = 0; item < i1; ++item, ++it) {
        mbar_wait(q_full, it & 1);
        mbar_wait(k_full, it & 1);
        tc_fence_after();
        const uint64_t dq[2] = {make_umma_desc_sw128(qa), make_umma_desc_sw128(qa + VT_TILE_BYTES)};
        // ---- pass A: scores for the row maximum, 256 keys per step
        for (int j = 0; j < nk256; ++j)
          for (int t = 0; t < 2; ++t) {
            if (j == 0) mbar_wait(&o_free[t], (it & 1) ^ 1);
            else mbar_wait(&s_free[t], n_sfree[t]++ & 1);
            tc_fence_after();
            const uint64_t dk = make_umma_desc_sw128(ka + j * 2 * VT_TILE_BYTES);
#pragma unroll
            for (int ks = 0; ks < VT_HD / 16; ++ks) umma_bf16(tmem_base + t * 256, dq[t] + 2 * ks, dk + 2 * ks, idesc_a, ks > 0);
            umma_commit(&s_full[t]);
          }
        // ---- pass B: probabilities and O += P V, 128 keys per step
        auto pv = [&](int jb, int t) {  // O_t += P_t V_jb (P of step jb is in TMEM, V block jb in ring slot vc % 3)
          const int slot = vc % VL_VSLOTS;
          mbar_wait(&p_full[t], n_pfull[t]++ & 1);
          if (t == 0) mbar_wait(&v_full[slot], (vc / VL_VSLOTS) & 1);
          tc_fence_after();
          const uint64_t dv = make_umma_desc_sw128(va + slot * VT_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_bf16_ts(tmem_base + t * 256 + 192, tmem_base + t * 256 + 8 * ks, dv + 128 * ks, idesc_o, (jb > 0 || ks > 0) ? 1u : 0u);
          if (t == 1) {
            umma_commit(&v_empty[slot]);
            ++vc;
          }
        };
        for (int j = 0; j < nk128; ++j)
          for (int t = 0; t < 2; ++t) {
            if (j == 0) {
              mbar_wait(&s_free[t], n_sfree[t]++ & 1);  // the last pass-A block has been read
              tc_fence_after();
            } else {
              pv(j - 1, t);
            }
            const uint64_t dk = make_umma_desc_sw128(ka + j * VT_TILE_BYTES);
#pragma unroll
            for (int ks = 0; ks < VT_HD / 16; ++ks) umma_bf16(tmem_base + t * 256, dq[t] + 2 * ks, dk + 2 * ks, idesc_b, ks > 0);
            umma_commit(&s_full[t]);
          }
        for (int t = 0; t < 2; ++t) {
          pv(nk128 - 1, t);
          umma_commit(&o_full[t]);
        }
        umma_commit(k_empty);  // every MMA that reads this item's Q / K has retired
      }
    }
  } else {
    // ================================ softmax groups (one per query tile) ================================
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
    const int gtid = threadIdx.x - (2 + 4 * t) * 32;
    uint8_t* qt = Qs + t * VT_TILE_BYTES;
    uint32_t n_sfull = 0;
    for (int item = i0, it = 0; item < i1; ++item, ++it) {
      const int hd = item / qpairs, qp = item - hd * qpairs;
      const int img = hd / heads, h = hd - img * heads;
      mbar_wait(q_full, it & 1);
      mbar_wait(k_full, it & 1);
      float sx = 0.f;  // score of the left-over key
#pragma unroll
      for (int c = 0; c < 8; ++c)
        sx += dot8(*reinterpret_cast<const uint4*>(qt + sw_off(row, c)), *reinterpret_cast<const uint4*>(xrows + c * 16));
      // ---- pass A: row maximum
      float mx = sx;
      for (int j = 0; j < nk256; ++j) {
        mbar_wait(&s_full[t], n_sfull++ & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 8; c += 2) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr + c * 32, r0);
          tmem_ld_32x32(taddr + c * 32 + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[e]), __uint_as_float(r1[e])));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
      }
      const float moff = mx * sl2;
      // ---- pass B: p = exp2(s * sl2 - moff), row sum, bf16 pack written over the scores
      float sum = 0.f;
      for (int j = 0; j < nk128; ++j) {
        mbar_wait(&s_full[t], n_sfull++ & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; c += 2) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr + c * 32, r0);
          tmem_ld_32x32(taddr + c * 32 + 32, r1);
          tmem_ld_wait();
          uint32_t pk0[16], pk1[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(r0[2 * e]), sl2, -moff));
            const float p1 = ex2_approx(fmaf(__uint_as_float(r0[2 * e + 1]), sl2, -moff));
            const float p2 = ex2_approx(fmaf(__uint_as_float(r1[2 * e]), sl2, -moff));
            const float p3 = ex2_approx(fmaf(__uint_as_float(r1[2 * e + 1]), sl2, -moff));
            sum += (p0 + p1) + (p2 + p3);
            pk0[e] = pack_bf16(p0, p1);
            pk1[e] = pack_bf16(p2, p3);
          }
          tmem_st_32x16(taddr + c * 16, pk0);
          tmem_st_32x16(taddr + c * 16 + 16, pk1);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      const float px = ex2_approx(fmaf(sx, sl2, -moff));
      sum += px;
      const float inv = 1.f / sum;
      uint4 vx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) vx[c] = *reinterpret_cast<const uint4*>(xrows + 128 + c * 16);
      mbar_wait(&o_full[t], it & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(taddr + 192, o0);
      tmem_ld_32x32(taddr + 224, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);
      // O = (P V + p_x v_x) / sum -> bf16 -> this row of the (dead) Q tile, swizzled as the TMA store expects
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t vw[4] = {vx[c].x, vx[c].y, vx[c].z, vx[c].w};
        uint32_t ow[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int jj = c * 8 + 2 * e;
          const float a0 = __uint_as_float(jj < 32 ? o0[jj] : o1[jj - 32]);
          const float a1 = __uint_as_float(jj < 32 ? o0[jj + 1] : o1[jj - 31]);
          ow[e] = pack_bf16((a0 + px * bf16_lo(vw[e])) * inv, (a1 + px * bf16_hi(vw[e])) * inv);
        }
        *reinterpret_cast<uint4*>(qt + sw_off(row, c)) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
      fence_proxy_async();
      named_bar_sync(1 + t, 128);
      if (gtid == 0) {
        tma_store_2d(&tmOut, qt, h * VT_HD, img * N + qp * 256 + t * 128);
        bulk_commit();
        bulk_wait_read<0>();  // the tile has been read out of shared memory
        mbar_arrive(q_empty);
        mbar_arrive(k_empty);
      }
    }
    if (gtid == 0) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// The left-over query row (token N - 1) of every (image, head) on CUDA cores: N scores, block softmax, P V. One CTA of 128
// threads per (image, head); K / V rows are read once from global memory (128 B each, L2 resident after the main kernel).
__global__ void __launch_bounds__(128) vit_attention_last_row_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                      __nv_bfloat16* __restrict__ out, int heads, int N,
                                                                      float scale) {
  extern __shared__ float lr_p[];  // [N] probabilities, then [2][64] partial outputs + [8] reductions
  float* red = lr_p + ((N + 3) & ~3);
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.x, img = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const int ld = 3 * heads * VT_HD;
  const __nv_bfloat16* base = qkv + (size_t)img * N * ld + h * VT_HD;
  const __nv_bfloat16* kb = base + heads * VT_HD;
  const __nv_bfloat16* vb = base + 2 * heads * VT_HD;
  const float sl2 = scale * 1.4426950408889634f;
  uint4 qx[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) qx[c] = *reinterpret_cast<const uint4*>(base + (size_t)(N - 1) * ld + c * 8);
  float mx = -INFINITY;
  for (int j = tid; j < N; j += 128) {
    float sc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) sc += dot8(qx[c], *reinterpret_cast<const uint4*>(kb + (size_t)j * ld + c * 8));
    lr_p[j] = sc;
    mx = fmaxf(mx, sc);
  }
  mx = warp_max(mx);
  if (lane == 0) red[128 + wp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[128], red[129]), fmaxf(red[130], red[131]));
  const float moff = mx * sl2;
  float sum = 0.f;
  for (int j = tid; j < N; j += 128) {
    const float pj = ex2_approx(fmaf(lr_p[j], sl2, -moff));
    lr_p[j] = pj;
    sum += pj;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[132 + wp] = sum;
  __syncthreads();
  sum = (red[132] + red[133]) + (red[134] + red[135]);
  const int d = tid & 63, par = tid >> 6;
  float acc = 0.f;
  for (int j = par; j < N; j += 2) acc += lr_p[j] * __bfloat162float(vb[(size_t)j * ld + d]);
  red[par * 64 + d] = acc;
  __syncthreads();
  if (tid < 64) out[((size_t)img * N + N - 1) * (heads * VT_HD) + h * VT_HD + tid] = __float2bfloat16((red[tid] + red[64 + tid]) / sum);
}


// ================================================================================================================
// Prefix-prefill attention of the LLM mixture on the tensor cores (joint_model.py:631-656 over the valid prefix of every
// sample; eager_attention_forward + repeat_kv): head_dim 128, GQA (the `group` query heads of a kv head are packed into
// the query tile: row = (token, head-in-group), so K / V of the kv head are loaded once), bidirectional over the L valid
// tokens of the sample, L <= 256 on this path (L comes from device memory; longer prefixes are left to the generic
// kernel, which is launched behind this one and skips the short ones).
// One CTA walks (sample, kv head) items; an item is ceil(L / TOK) query tiles of TOK = 128 / group tokens (126 rows at
// group 6) against the same L keys:
//   warp 0       TMA producer: K and V of the item as 32-row boxes of the KV cache (two 64-dim halves each, 128B swizzle),
//                then the query tiles through a 3-D box {64 dims, group heads, TOK tokens} of the q buffer, 2-deep ring
//   warp 1       MMA issuer: S = Q K^T (M 128, N = L rounded to 32, K 128 = 8 k-steps), O = P V (P from TMEM, V as MN-major
//                operand, two 64-dim halves per k-step)
//   warps 2..5   softmax / epilogue group of the even tiles, warps 6..9 of the odd tiles: one thread per query row; S, P
//                (bf16, over S) and O (columns 128..255 of the slot) live in a 256-column TMEM slot per group; O leaves
//                through a transposition tile in the (dead) Q slot and row-contiguous 256-byte global stores
// ================================================================================================================
constexpr int PT_THREADS = 320;
constexpr int PT_KMAX = 256;                         // keys on this path
constexpr int PT_HALF_BYTES = PT_KMAX * 128;         // one 64-dim half of K or V: [256 rows][128 B]
constexpr int PT_Q_BYTES = 2 * VT_TILE_BYTES;        // one query tile: two 64-dim halves of [128 rows][128 B]
constexpr int PT_QSLOTS = 3;                         // query-tile ring: the producer runs a tile ahead of the two softmax groups
constexpr int PT_SMEM = 4 * PT_HALF_BYTES + PT_QSLOTS * PT_Q_BYTES + 1024 + 256;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

struct PrefillTcParams {
  const int* q_start;    // [B] first packed row of the sample
  const int* q_len;      // [B] valid tokens L
  const int* kv_start;   // [B] first KV-cache row of the sample
  __nv_bfloat16* out;    // same addressing as q: row r, head h at out + r * out_row_stride + h * 128
  long long out_row_stride;
  long long kv_layer_row0;  // row offset of this layer's kv head 0 in the [layers * kv_heads * kv_rows, 128] cache view
  long long kv_rows;        // rows per kv head
  int batch, n_kv_heads, group;
  int tsplit;            // CTAs that share the query tiles of one (sample, kv head): CTA `sp` takes tiles sp, sp + tsplit, ...
  float scale;
  long long* trace;
};

__global__ void __launch_bounds__(PT_THREADS, 1)
prefill_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                            const __grid_constant__ CUtensorMap tmV, const PrefillTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Ks = smem;                          // [2 halves][256][128 B]
  uint8_t* Vs = smem + 2 * PT_HALF_BYTES;      // [2 halves][256][128 B]
  uint8_t* Qs = smem + 4 * PT_HALF_BYTES;      // [3 slots][2 halves][128][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Qs + PT_QSLOTS * PT_Q_BYTES);
  uint64_t* kv_full = bars;        // [1]
  uint64_t* kv_empty = bars + 1;   // [1]
  uint64_t* q_full = bars + 2;     // [3] query slot (tile counter % 3)
  uint64_t* q_empty = bars + 5;    // [3]
  uint64_t* s_full = bars + 8;     // [2] TMEM slot / softmax group (tile counter & 1)
  uint64_t* p_full = bars + 10;    // [2]
  uint64_t* o_full = bars + 12;    // [2]
  uint64_t* t_empty = bars + 14;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* tr = (blockIdx.x == 0 && lane == 0) ? p.trace : nullptr;
  if (warp == 0 && lane == 0) {
    trace_stamp(tr, 0);
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int i = 0; i < PT_QSLOTS; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&t_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();  // q and the K / V rows of this layer are complete (and the plan, from which L is read)
  if (warp == 0) trace_stamp(tr, 1);

  const int group = p.group, TOK = 128 / group, rows_tile = TOK * group;
  const int tsplit = p.tsplit;
  const int items = p.batch * p.n_kv_heads * tsplit;
  const float sl2 = p.scale * 1.4426950408889634f;
  // item -> (sample b, kv head g, tile phase sp); tiles of the item: sp, sp + tsplit, ... (count below)
#define PT_ITEM(item)                                                       \
  const int sp = (item) % tsplit, bg = (item) / tsplit;                     \
  const int b = bg / p.n_kv_heads, g = bg - b * p.n_kv_heads;               \
  const int L = p.q_len[b];                                                 \
  (void)g;                                                                  \
  const int all_tiles = (L + TOK - 1) / TOK;                                \
  const int n_tiles = all_tiles > sp ? (all_tiles - sp + tsplit - 1) / tsplit : 0; \
  if (L <= 0 || L > PT_KMAX || n_tiles == 0) continue; /* empty, or left to the generic kernel */

  if (warp == 0) {
    // ================================ TMA producer ================================
    // The whole warp walks the items; lane 0 waits and arms the barriers, then one lane per box issues the loads -- the
    // 4 x ceil(L / 32) K / V boxes + 2 query boxes of an item took one thread 3.8 us (0.27 us per issue, in-situ stamps
    // profiles/r02s3q), in front of the first score tile of a CTA that owns exactly one item
    {
      int ic = 0, tcn = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        PT_ITEM(item)
        const int nblk = (L + 31) >> 5;
        const int krow = (int)(p.kv_layer_row0 + (long long)g * p.kv_rows) + p.kv_start[b];
        if (lane == 0) {
          mbar_wait(kv_empty, (ic & 1) ^ 1);
          mbar_arrive_expect_tx(kv_full, (uint32_t)(4 * nblk * 4096));
        }
        __syncwarp();
        if (lane < 4 * nblk) {  // lane = 4 j + {K lo, K hi, V lo, V hi}
          const int j = lane >> 2, hi = lane & 1, isv = (lane >> 1) & 1;
          tma_load_2d((isv ? Vs : Ks) + hi * PT_HALF_BYTES + j * 4096, isv ? &tmV : &tmK, kv_full, hi * 64, krow + j * 32);
        }
        ++ic;
        const int q0 = p.q_start[b];
        for (int ti = 0; ti < n_tiles; ++ti, ++tcn) {
          const int t = sp + ti * tsplit;
          const int slot = tcn % PT_QSLOTS;
          uint8_t* qd = Qs + slot * PT_Q_BYTES;
          if (lane == 0) {
            mbar_wait(&q_empty[slot], ((tcn / PT_QSLOTS) & 1) ^ 1);
            mbar_arrive_expect_tx(&q_full[slot], (uint32_t)(2 * rows_tile * 128));
          }
          __syncwarp();
          if (lane < 2) tma_load_3d(qd + lane * VT_TILE_BYTES, &tmQ, &q_full[slot], lane * 64, g * group, q0 + t * TOK);
          if (tcn == 0) trace_stamp(tr, 2);  // K / V and the first query tile are on their way
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64) | (1u << 16);  // B (= a 64-dim half of V) is MN-major
      const uint64_t dk0 = make_umma_desc_sw128(smem_u32(Ks)), dk1 = make_umma_desc_sw128(smem_u32(Ks + PT_HALF_BYTES));
      const uint64_t dv0 = make_umma_desc_sw128(smem_u32(Vs)), dv1 = make_umma_desc_sw128(smem_u32(Vs + PT_HALF_BYTES));
      int ic = 0, tcn = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        PT_ITEM(item)
        const int Lp = (L + 31) & ~31;
        const uint32_t idesc_s = make_idesc_bf16(128, Lp);
        mbar_wait(kv_full, ic & 1);
        tc_fence_after();
        if (ic == 0) trace_stamp(tr, 3);  // K / V landed
        auto issue_s = [&](int tc_i) {
          const int slot = tc_i & 1, qs = tc_i % PT_QSLOTS;
          const uint32_t ph = (tc_i >> 1) & 1;
          mbar_wait(&q_full[qs], (tc_i / PT_QSLOTS) & 1);
          mbar_wait(&t_empty[slot], ph ^ 1);
          tc_fence_after();
          const uint32_t qa = smem_u32(Qs + qs * PT_Q_BYTES);
          const uint64_t dq0 = make_umma_desc_sw128(qa), dq1 = make_umma_desc_sw128(qa + VT_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_base + slot * 256, dq0 + 2 * ks, dk0 + 2 * ks, idesc_s, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_base + slot * 256, dq1 + 2 * ks, dk1 + 2 * ks, idesc_s, 1u);
          umma_commit(&s_full[slot]);
        };
        issue_s(tcn);
        for (int t = 0; t < n_tiles; ++t) {
          const int tc_i = tcn + t, slot = tc_i & 1;
          if (t + 1 < n_tiles) issue_s(tc_i + 1);  // the other group's scores while this group's softmax runs
          mbar_wait(&p_full[slot], (tc_i >> 1) & 1);
          tc_fence_after();
          const uint32_t tb = tmem_base + slot * 256;
          for (int ks = 0; ks < Lp / 16; ++ks) {
            umma_bf16_ts(tb + 128, tb + 8 * ks, dv0 + 128 * ks, idesc_o, ks > 0);
            umma_bf16_ts(tb + 192, tb + 8 * ks, dv1 + 128 * ks, idesc_o, ks > 0);
          }
          umma_commit(&o_full[slot]);
        }
        umma_commit(kv_empty);  // every MMA that reads this item's K / V has retired
        tcn += n_tiles;
        ++ic;
      }
    }
  } else {
    // ================================ softmax / epilogue groups ================================
    const int grp = (warp - 2) >> 2;             // tiles with (tile counter & 1) == grp
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;         // packed query row within the tile: token row / group, head row % group
    const int gtid = threadIdx.x - (2 + 4 * grp) * 32;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + grp * 256;
    int tcn = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      PT_ITEM(item)
      const int Lp = (L + 31) & ~31;
      const int q0 = p.q_start[b];
      for (int ti = 0; ti < n_tiles; ++ti) {
        const int tc_i = tcn + ti;
        if ((tc_i & 1) != grp) continue;
        const int t = sp + ti * tsplit;
        const uint32_t ph = (tc_i >> 1) & 1;
        uint8_t* stage = Qs + (tc_i % PT_QSLOTS) * PT_Q_BYTES;  // [128 rows][256 B], 16-byte chunks XOR-swizzled with the row
        const int tok = t * TOK + row / group, hh = row % group;
        const bool row_ok = row < rows_tile && tok < L;
        mbar_wait(&s_full[grp], ph);
        tc_fence_after();
        if (tc_i == 0 && warp == 2) trace_stamp(tr, 4);
        float sum = 0.f;
        if (Lp <= 96) {
          // up to 3 chunks: the whole score row in registers with ONE TMEM round trip (a tcgen05.ld -> wait is ~0.3 us;
          // the prefixes of this path are ~90 tokens)
          uint32_t sreg[3][32];
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c * 32 < Lp) tmem_ld_32x32(taddr + c * 32, sreg[c]);
          tmem_ld_wait();
          float mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c * 32 < Lp) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c * 32 + j < L) mx = fmaxf(mx, __uint_as_float(sreg[c][j]));
            }
          const float moff = mx * sl2;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c * 32 < Lp) {
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float p0 = ex2_approx(fmaf(__uint_as_float(sreg[c][2 * j]), sl2, -moff));
                float p1 = ex2_approx(fmaf(__uint_as_float(sreg[c][2 * j + 1]), sl2, -moff));
                if (c * 32 + 2 * j >= L) p0 = 0.f;
                if (c * 32 + 2 * j + 1 >= L) p1 = 0.f;
                sum += p0 + p1;
                pk[j] = pack_bf16(p0, p1);
              }
              tmem_st_32x16(taddr + c * 16, pk);
            }
        } else {
          // ---- pass 1: row max over the live keys (two chunks per round trip)
          float mx = -INFINITY;
#pragma unroll 1
          for (int c = 0; c < Lp; c += 64) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(taddr + c, r0);
            tmem_ld_32x32(taddr + c + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (c + j < L) mx = fmaxf(mx, __uint_as_float(r0[j]));
              if (c + 32 + j < L) mx = fmaxf(mx, __uint_as_float(r1[j]));
            }
          }
          const float moff = mx * sl2;
          // ---- pass 2: probabilities (0 beyond the live keys), row sum, bf16 pack over the scores
#pragma unroll 1
          for (int c = 0; c < Lp; c += 64) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(taddr + c, r0);
            tmem_ld_32x32(taddr + c + 32, r1);
            tmem_ld_wait();
            uint32_t pk0[16], pk1[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float p0 = ex2_approx(fmaf(__uint_as_float(r0[2 * j]), sl2, -moff));
              float p1 = ex2_approx(fmaf(__uint_as_float(r0[2 * j + 1]), sl2, -moff));
              float p2 = ex2_approx(fmaf(__uint_as_float(r1[2 * j]), sl2, -moff));
              float p3 = ex2_approx(fmaf(__uint_as_float(r1[2 * j + 1]), sl2, -moff));
              if (c + 2 * j >= L) p0 = 0.f;
              if (c + 2 * j + 1 >= L) p1 = 0.f;
              if (c + 32 + 2 * j >= L) p2 = 0.f;
              if (c + 32 + 2 * j + 1 >= L) p3 = 0.f;
              sum += (p0 + p1) + (p2 + p3);
              pk0[j] = pack_bf16(p0, p1);
              pk1[j] = pack_bf16(p2, p3);
            }
            tmem_st_32x16(taddr + (c >> 1), pk0);
            tmem_st_32x16(taddr + (c >> 1) + 16, pk1);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[grp]);
        long long* trg = (warp == 2 && tc_i <= 2) ? tr : nullptr;  // group 0: its first two tiles
        trace_stamp(trg, tc_i == 0 ? 8 : 12);    // probabilities written
        const float inv = 1.f / sum;
        mbar_wait(&o_full[grp], ph);
        tc_fence_after();
        trace_stamp(trg, tc_i == 0 ? 9 : 13);    // O = P V complete
        // ---- O / sum -> bf16 -> this row of the transposition tile (the query slot is dead: S is complete)
#pragma unroll 1
        for (int c0 = 0; c0 < 4; c0 += 2) {
          uint32_t o[2][32];
          tmem_ld_32x32(taddr + 128 + c0 * 32, o[0]);
          tmem_ld_32x32(taddr + 128 + c0 * 32 + 32, o[1]);
          tmem_ld_wait();
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 u;
              u.x = pack_bf16(__uint_as_float(o[cc][8 * q]) * inv, __uint_as_float(o[cc][8 * q + 1]) * inv);
              u.y = pack_bf16(__uint_as_float(o[cc][8 * q + 2]) * inv, __uint_as_float(o[cc][8 * q + 3]) * inv);
              u.z = pack_bf16(__uint_as_float(o[cc][8 * q + 4]) * inv, __uint_as_float(o[cc][8 * q + 5]) * inv);
              u.w = pack_bf16(__uint_as_float(o[cc][8 * q + 6]) * inv, __uint_as_float(o[cc][8 * q + 7]) * inv);
              const int ch = (c0 + cc) * 4 + q;  // 16-byte chunk of the 256-byte row
              *reinterpret_cast<uint4*>(stage + row * 256 + ((ch ^ (row & 15)) << 4)) = u;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[grp]);  // the slot's TMEM columns may take the group's next scores
        trace_stamp(trg, tc_i == 0 ? 10 : 14);   // O read out of TMEM
        // ---- row-contiguous stores: 16 lanes x 16 B = one 256-byte row, two rows per instruction
        const unsigned ok_mask = __ballot_sync(0xffffffffu, row_ok);
#pragma unroll 1
        for (int it = 0; it < 16; ++it) {
          const int rl = it * 2 + (lane >> 4), ch = lane & 15;     // row within this warp's 32 rows
          const int r = quarter * 32 + rl;
          if ((ok_mask >> rl) & 1u) {
            const uint4 u = *reinterpret_cast<const uint4*>(stage + r * 256 + ((ch ^ (r & 15)) << 4));
            const int tk = t * TOK + r / group, h2 = r % group;
            *reinterpret_cast<uint4*>(p.out + (size_t)(q0 + tk) * p.out_row_stride + (size_t)(g * group + h2) * 128 + ch * 8) = u;
          }
        }
        (void)tok;
        (void)hh;
        named_bar_sync(1 + grp, 128);  // every warp of the group has read its rows out of the slot
        trace_stamp(trg, tc_i == 0 ? 11 : 15);   // tile stored
        if (gtid == 0) mbar_arrive(&q_empty[tc_i % PT_QSLOTS]);
      }
      tcn += n_tiles;
    }
    if (grp == 0 && warp == 2) trace_stamp(tr, 5);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
  if (warp == 2) trace_stamp(tr, 6);
}

// ================================================================================================================
// Long prefixes (more than 256 tokens: 448 px tiles, several cameras, the dynamic tiles of BASELINE config 5, the causal
// prompt of the chat VLM) on the tensor cores. Same GQA packing and TMEM slots as the kernel above, but the keys of a sample
// no longer fit in one score tile, so K and V stream through shared memory in blocks of 128 keys and softmax is exact in two
// passes like vit_attention_tc_long_kernel: pass A computes S = Q K^T block by block for the row maximum only, pass B
// recomputes S, writes p = exp2(s - max) over it (bf16) and accumulates O += P V in TMEM -- no rescaling of O.
// `causal`: query token t sees keys [0, t] (HF causal mask of the text prefill); tile pairs stop at the last key their last
// token sees. Work item = (sample, kv head, pair of query tiles); group t of the CTA owns tile 2 * pair + t.
// ================================================================================================================
constexpr int PL_THREADS = 320;
constexpr int PL_BLK_BYTES = 2 * VT_TILE_BYTES;  // one block of 128 keys: two 64-dim halves of [128 rows][128 B]
constexpr int PL_SMEM = 2 * PT_Q_BYTES + 4 * PL_BLK_BYTES + 1024 + 256;

struct PrefillLongParams {
  const int* q_start;
  const int* q_len;
  const int* kv_start;
  __nv_bfloat16* out;
  long long out_row_stride;
  long long kv_layer_row0, kv_rows;
  int batch, n_kv_heads, group;
  int max_pairs;   // ceil(ceil(max tokens / TOK) / 2)
  int min_len;     // samples with fewer tokens are left to the short kernel
  int causal;
  float scale;
};

__global__ void __launch_bounds__(PL_THREADS, 1)
prefill_attention_tc_long_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                 const __grid_constant__ CUtensorMap tmV, const PrefillLongParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;                          // [2 tiles][2 halves][128][128 B]
  uint8_t* Ks = smem + 2 * PT_Q_BYTES;         // [2 slots][2 halves][128][128 B]
  uint8_t* Vs = Ks + 2 * PL_BLK_BYTES;         // [2 slots][2 halves][128][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + 2 * PL_BLK_BYTES);
  uint64_t* q_full = bars;         // [1]
  uint64_t* q_empty = bars + 1;    // [1] both groups are done with their query / staging tile
  uint64_t* k_full = bars + 2;     // [2]
  uint64_t* k_empty = bars + 4;    // [2]
  uint64_t* v_full = bars + 6;     // [2]
  uint64_t* v_empty = bars + 8;    // [2]
  uint64_t* s_full = bars + 10;    // [2]
  uint64_t* s_free = bars + 12;    // [2] pass-A scores read (4 warps)
  uint64_t* p_full = bars + 14;    // [2] (4 warps)
  uint64_t* o_full = bars + 16;    // [2]
  uint64_t* o_free = bars + 18;    // [2] (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    // Nothing to do? (The common policy call: max_image_text_tokens allows long prefixes, but every sample of the batch fits
    // the short kernel.) The lengths were written by the plan kernels at the start of the call -- every launch since has
    // waited for its predecessor, so they are visible here without the dependent-launch wait -- and leaving before the
    // barrier / TMEM set-up makes the empty launch a bare hop of the chain.
    int any = 0;
    for (int i = threadIdx.x; i < p.batch; i += PL_THREADS) any |= (p.q_len[i] >= p.min_len) ? 1 : 0;
    if (!__syncthreads_or(any)) return;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int group = p.group, TOK = 128 / group, rows_tile = TOK * group;
  const int bgs = p.batch * p.n_kv_heads;
  const int items = bgs * p.max_pairs;
  const float sl2 = p.scale * 1.4426950408889634f;
  // item -> (pair tp, sample b, kv head g), pair-major so that neighbouring CTAs take different samples. nt: tiles of the
  // pair that exist (1 | 2); nkb: key blocks the pair needs
#define PL_ITEM(item)                                                                   \
  const int tp = (item) / bgs, bg = (item) - tp * bgs;                                  \
  const int b = bg / p.n_kv_heads, g = bg - b * p.n_kv_heads;                           \
  const int L = p.q_len[b];                                                             \
  (void)g;                                                                              \
  const int all_tiles = (L + TOK - 1) / TOK;                                            \
  if (L < p.min_len || 2 * tp >= all_tiles) continue;                                   \
  const int nt = min(2, all_tiles - 2 * tp);                                            \
  const int kmax = p.causal ? min(L, (2 * tp + nt) * TOK) : L;                          \
  const int nkb = (kmax + 127) >> 7;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int it = 0, kc = 0, vc = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        PL_ITEM(item)
        const int krow = (int)(p.kv_layer_row0 + (long long)g * p.kv_rows) + p.kv_start[b];
        const int q0 = p.q_start[b];
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, (uint32_t)(nt * 2 * rows_tile * 128));
        for (int t = 0; t < nt; ++t) {
          tma_load_3d(Qs + t * PT_Q_BYTES, &tmQ, q_full, 0, g * group, q0 + (2 * tp + t) * TOK);
          tma_load_3d(Qs + t * PT_Q_BYTES + VT_TILE_BYTES, &tmQ, q_full, 64, g * group, q0 + (2 * tp + t) * TOK);
        }
        auto load_blk = [&](uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int j) {
          mbar_arrive_expect_tx(bar, PL_BLK_BYTES);
          for (int jj = 0; jj < 4; ++jj) {
            tma_load_2d(dst + jj * 4096, tm, bar, 0, krow + j * 128 + jj * 32);
            tma_load_2d(dst + VT_TILE_BYTES + jj * 4096, tm, bar, 64, krow + j * 128 + jj * 32);
          }
        };
        for (int j = 0; j < nkb; ++j, ++kc) {  // pass A: keys only
          mbar_wait(&k_empty[kc & 1], ((kc >> 1) & 1) ^ 1);
          load_blk(Ks + (kc & 1) * PL_BLK_BYTES, &tmK, &k_full[kc & 1], j);
        }
        for (int j = 0; j < nkb; ++j, ++kc, ++vc) {  // pass B: keys and values
          mbar_wait(&k_empty[kc & 1], ((kc >> 1) & 1) ^ 1);
          load_blk(Ks + (kc & 1) * PL_BLK_BYTES, &tmK, &k_full[kc & 1], j);
          mbar_wait(&v_empty[vc & 1], ((vc >> 1) & 1) ^ 1);
          load_blk(Vs + (vc & 1) * PL_BLK_BYTES, &tmV, &v_full[vc & 1], j);
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64) | (1u << 16);  // B (= a 64-dim half of V) is MN-major
      const uint32_t qa = smem_u32(Qs), ka = smem_u32(Ks), va = smem_u32(Vs);
      uint32_t n_sfree[2] = {0, 0}, n_pfull[2] = {0, 0}, n_ofree[2] = {0, 0};
      int it = 0, kc = 0, vc = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        PL_ITEM(item)
        mbar_wait(q_full, it & 1);
        tc_fence_after();
        auto issue_s = [&](int t, int kslot) {
          const uint64_t dq0 = make_umma_desc_sw128(qa + t * PT_Q_BYTES), dq1 = make_umma_desc_sw128(qa + t * PT_Q_BYTES + VT_TILE_BYTES);
          const uint64_t dk0 = make_umma_desc_sw128(ka + kslot * PL_BLK_BYTES);
          const uint64_t dk1 = make_umma_desc_sw128(ka + kslot * PL_BLK_BYTES + VT_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_base + t * 256, dq0 + 2 * ks, dk0 + 2 * ks, idesc_s, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_base + t * 256, dq1 + 2 * ks, dk1 + 2 * ks, idesc_s, 1u);
          umma_commit(&s_full[t]);
        };
        // ---- pass A
        for (int j = 0; j < nkb; ++j, ++kc) {
          mbar_wait(&k_full[kc & 1], (kc >> 1) & 1);
          tc_fence_after();
          for (int t = 0; t < nt; ++t) {
            if (j == 0) mbar_wait(&o_free[t], (n_ofree[t] & 1) ^ 1);  // O of the group's previous tile has been read out
            else mbar_wait(&s_free[t], n_sfree[t]++ & 1);
            tc_fence_after();
            issue_s(t, kc & 1);
          }
          umma_commit(&k_empty[kc & 1]);
        }
        // ---- pass B
        auto pv = [&](int jb, int t, bool last_tile) {  // O_t += P_t V_jb; V block jb sits in ring slot vc & 1
          mbar_wait(&p_full[t], n_pfull[t]++ & 1);
          if (t == 0) mbar_wait(&v_full[vc & 1], (vc >> 1) & 1);
          tc_fence_after();
          const uint32_t tb = tmem_base + t * 256;
          const uint64_t dv0 = make_umma_desc_sw128(va + (vc & 1) * PL_BLK_BYTES);
          const uint64_t dv1 = make_umma_desc_sw128(va + (vc & 1) * PL_BLK_BYTES + VT_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t acc = (jb > 0 || ks > 0) ? 1u : 0u;
            umma_bf16_ts(tb + 128, tb + 8 * ks, dv0 + 128 * ks, idesc_o, acc);
            umma_bf16_ts(tb + 192, tb + 8 * ks, dv1 + 128 * ks, idesc_o, acc);
          }
          if (last_tile) {
            umma_commit(&v_empty[vc & 1]);
            ++vc;
          }
        };
        for (int j = 0; j < nkb; ++j, ++kc) {
          mbar_wait(&k_full[kc & 1], (kc >> 1) & 1);
          tc_fence_after();
          for (int t = 0; t < nt; ++t) {
            if (j == 0) {
              mbar_wait(&s_free[t], n_sfree[t]++ & 1);  // the last pass-A block has been read
              tc_fence_after();
            } else {
              pv(j - 1, t, t == nt - 1);
            }
            issue_s(t, kc & 1);
          }
          umma_commit(&k_empty[kc & 1]);
        }
        for (int t = 0; t < nt; ++t) {
          pv(nkb - 1, t, t == nt - 1);
          umma_commit(&o_full[t]);
          ++n_ofree[t];
        }
        ++it;
      }
    }
  } else {
    // ================================ softmax / epilogue groups ================================
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int gtid = threadIdx.x - (2 + 4 * t) * 32;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
    uint8_t* stage = Qs + t * PT_Q_BYTES;  // the group's query tile, dead once its last S is complete
    uint32_t n_sfull = 0, n_ofull = 0;
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      PL_ITEM(item)
      const int q0 = p.q_start[b];
      if (t >= nt) {  // odd tile count: this group has no tile in the pair
        mbar_wait(q_full, it & 1);
        if (gtid == 0) mbar_arrive(q_empty);
        ++it;
        continue;
      }
      const int tile = 2 * tp + t;
      const int tok = min(tile * TOK + row / group, L - 1);   // (rows past the sample: any valid token, never stored)
      const bool row_ok = row < rows_tile && tile * TOK + row / group < L;
      const int lim = p.causal ? min(L, tok + 1) : L;          // keys [0, lim) are visible
      // ---- pass A: row maximum over the visible keys
      float mx = -INFINITY;
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(&s_full[t], n_sfull++ & 1);
        tc_fence_after();
        const int kb = j * 128;
#pragma unroll 1
        for (int c = 0; c < 128; c += 64) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr + c, r0);
          tmem_ld_32x32(taddr + c + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (kb + c + e < lim) mx = fmaxf(mx, __uint_as_float(r0[e]));
            if (kb + c + 32 + e < lim) mx = fmaxf(mx, __uint_as_float(r1[e]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
      }
      const float moff = mx * sl2;
      // ---- pass B: probabilities (0 for hidden keys), row sum, bf16 pack over the scores
      float sum = 0.f;
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(&s_full[t], n_sfull++ & 1);
        tc_fence_after();
        const int kb = j * 128;
#pragma unroll 1
        for (int c = 0; c < 128; c += 64) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr + c, r0);
          tmem_ld_32x32(taddr + c + 32, r1);
          tmem_ld_wait();
          uint32_t pk0[16], pk1[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float p0 = ex2_approx(fmaf(__uint_as_float(r0[2 * e]), sl2, -moff));
            float p1 = ex2_approx(fmaf(__uint_as_float(r0[2 * e + 1]), sl2, -moff));
            float p2 = ex2_approx(fmaf(__uint_as_float(r1[2 * e]), sl2, -moff));
            float p3 = ex2_approx(fmaf(__uint_as_float(r1[2 * e + 1]), sl2, -moff));
            if (kb + c + 2 * e >= lim) p0 = 0.f;
            if (kb + c + 2 * e + 1 >= lim) p1 = 0.f;
            if (kb + c + 32 + 2 * e >= lim) p2 = 0.f;
            if (kb + c + 32 + 2 * e + 1 >= lim) p3 = 0.f;
            sum += (p0 + p1) + (p2 + p3);
            pk0[e] = pack_bf16(p0, p1);
            pk1[e] = pack_bf16(p2, p3);
          }
          tmem_st_32x16(taddr + (c >> 1), pk0);
          tmem_st_32x16(taddr + (c >> 1) + 16, pk1);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      const float inv = 1.f / sum;
      mbar_wait(&o_full[t], n_ofull++ & 1);
      tc_fence_after();
      // ---- O / sum -> bf16 -> this row of the transposition tile
#pragma unroll 1
      for (int c0 = 0; c0 < 4; c0 += 2) {
        uint32_t o[2][32];
        tmem_ld_32x32(taddr + 128 + c0 * 32, o[0]);
        tmem_ld_32x32(taddr + 128 + c0 * 32 + 32, o[1]);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(o[cc][8 * q]) * inv, __uint_as_float(o[cc][8 * q + 1]) * inv);
            u.y = pack_bf16(__uint_as_float(o[cc][8 * q + 2]) * inv, __uint_as_float(o[cc][8 * q + 3]) * inv);
            u.z = pack_bf16(__uint_as_float(o[cc][8 * q + 4]) * inv, __uint_as_float(o[cc][8 * q + 5]) * inv);
            u.w = pack_bf16(__uint_as_float(o[cc][8 * q + 6]) * inv, __uint_as_float(o[cc][8 * q + 7]) * inv);
            const int ch = (c0 + cc) * 4 + q;
            *reinterpret_cast<uint4*>(stage + row * 256 + ((ch ^ (row & 15)) << 4)) = u;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);
      const unsigned ok_mask = __ballot_sync(0xffffffffu, row_ok);
#pragma unroll 1
      for (int i2 = 0; i2 < 16; ++i2) {
        const int rl = i2 * 2 + (lane >> 4), ch = lane & 15;
        const int r = quarter * 32 + rl;
        if ((ok_mask >> rl) & 1u) {
          const uint4 u = *reinterpret_cast<const uint4*>(stage + r * 256 + ((ch ^ (r & 15)) << 4));
          const int tk = tile * TOK + r / group, h2 = r % group;
          *reinterpret_cast<uint4*>(p.out + (size_t)(q0 + tk) * p.out_row_stride + (size_t)(g * group + h2) * 128 + ch * 8) = u;
        }
      }
      named_bar_sync(1 + t, 128);  // every warp of the group has read its rows out of the tile
      if (gtid == 0) mbar_arrive(q_empty);
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int vit_attention_tc_launch(const CUtensorMap& tmQKV, const CUtensorMap& tmOut, const __nv_bfloat16* qkv,
                            __nv_bfloat16* out, int images, int tokens, int heads, float scale, cudaStream_t stream,
                            long long* trace) {
  VL_CHECK(images > 0 && heads > 0, "vit_attention_tc: empty problem");
  if (tokens != 257) {
    // long rows (448 px: 1025 tokens): two-pass kernel for the Nk x Nk block + the left-over query row
    const int Nk = tokens - 1;
    VL_CHECK(Nk >= 256 && Nk % 256 == 0 && Nk <= 128 * VL_MAX_KT, "vit_attention_tc: tokens must be 257, or 1 + a multiple of 256 up to 1025");
    static unsigned long long attr_devs_l = 0;
    if (vl_first_on_device(attr_devs_l))
      VL_CUDA_CHECK(cudaFuncSetAttribute(vit_attention_tc_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VL_SMEM));
    const int items = images * heads * (Nk / 256);
    const int per = (items + gemm_num_sms() - 1) / gemm_num_sms();
    const int grid = (items + per - 1) / per;
    VL_CUDA_CHECK(launch_kernel(vit_attention_tc_long_kernel, dim3(grid), dim3(VL_THREADS), VL_SMEM, stream, tmQKV, tmOut, qkv,
                                images, heads, tokens, scale));
    const size_t lr_smem = (size_t)(((tokens + 3) & ~3) + 144) * sizeof(float);
    VL_CUDA_CHECK(launch_kernel(vit_attention_last_row_kernel, dim3(heads, images), dim3(128), lr_smem, stream, qkv, out, heads,
                                tokens, scale));
    return 0;
  }
  static unsigned long long attr_devs = 0;  // devices this kernel's attributes were set on
  if (vl_first_on_device(attr_devs)) {
    VL_CUDA_CHECK(cudaFuncSetAttribute(vit_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VT_SMEM));
  }
  const int grid = std::min(images * heads, gemm_num_sms());
  VL_CUDA_CHECK(launch_kernel(vit_attention_tc_kernel, dim3(grid), dim3(VT_THREADS), VT_SMEM, stream, tmQKV, tmOut, qkv, out,
                              images, heads, scale, trace));
  return 0;
}

int prefill_attention_tc_launch(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const int* q_start,
                                const int* q_len, const int* kv_start, __nv_bfloat16* out, long long out_row_stride,
                                long long kv_layer_row0, long long kv_rows, int batch, int n_kv_heads, int group, float scale,
                                cudaStream_t stream, long long* trace) {
  VL_CHECK(batch > 0 && n_kv_heads > 0 && group >= 1 && group <= 8, "prefill_attention_tc: bad shape");
  static unsigned long long attr_devs = 0;  // devices this kernel's attributes were set on
  if (vl_first_on_device(attr_devs)) {
    VL_CUDA_CHECK(cudaFuncSetAttribute(prefill_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM));
  }
  PrefillTcParams p;
  p.q_start = q_start, p.q_len = q_len, p.kv_start = kv_start;
  p.out = out, p.out_row_stride = out_row_stride;
  p.kv_layer_row0 = kv_layer_row0, p.kv_rows = kv_rows;
  p.batch = batch, p.n_kv_heads = n_kv_heads, p.group = group, p.scale = scale, p.trace = trace;
  // few (sample, kv head) pairs: their query tiles are shared out between up to 8 CTAs each (K / V are then loaded per CTA)
  const int pairs = batch * n_kv_heads;
  p.tsplit = std::max(1, std::min(8, (2 * gemm_num_sms()) / std::max(pairs, 1)));
  const int grid = std::min(pairs * p.tsplit, gemm_num_sms());
  VL_CUDA_CHECK(launch_kernel(prefill_attention_tc_kernel, dim3(grid), dim3(PT_THREADS), PT_SMEM, stream, tmQ, tmK, tmV, p));
  return 0;
}

// q [rows, n_heads, 128] bf16 as a 3-D map with boxes {64 dims, group heads, 128 / group tokens}, 128B swizzle
int prefill_attention_tc_long_launch(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const int* q_start,
                                     const int* q_len, const int* kv_start, __nv_bfloat16* out, long long out_row_stride,
                                     long long kv_layer_row0, long long kv_rows, int batch, int n_kv_heads, int group, int max_tokens,
                                     int min_len, int causal, float scale, cudaStream_t stream) {
  VL_CHECK(batch > 0 && n_kv_heads > 0 && group >= 1 && group <= 8 && max_tokens > 0, "prefill_attention_tc_long: bad shape");
  static unsigned long long attr_devs = 0;
  if (vl_first_on_device(attr_devs))
    VL_CUDA_CHECK(cudaFuncSetAttribute(prefill_attention_tc_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PL_SMEM));
  PrefillLongParams p;
  p.q_start = q_start, p.q_len = q_len, p.kv_start = kv_start;
  p.out = out, p.out_row_stride = out_row_stride;
  p.kv_layer_row0 = kv_layer_row0, p.kv_rows = kv_rows;
  p.batch = batch, p.n_kv_heads = n_kv_heads, p.group = group;
  const int TOK = 128 / group;
  p.max_pairs = ((max_tokens + TOK - 1) / TOK + 1) / 2;
  p.min_len = min_len, p.causal = causal, p.scale = scale;
  const long long items = (long long)batch * n_kv_heads * p.max_pairs;
  const int grid = (int)std::min<long long>(items, gemm_num_sms());
  VL_CUDA_CHECK(launch_kernel(prefill_attention_tc_long_kernel, dim3(grid), dim3(PL_THREADS), PL_SMEM, stream, tmQ, tmK, tmV, p));
  return 0;
}

int make_tmap_q3d(CUtensorMap* out, const void* ptr, uint64_t rows, int n_heads, int group) {
  PFN_encodeTiled_tc fn = get_encode_fn_tc();
  VL_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {128, (cuuint64_t)n_heads, rows};
  cuuint64_t strides[2] = {256, (cuuint64_t)n_heads * 256};
  cuuint32_t box[3] = {64, (cuuint32_t)group, (cuuint32_t)(128 / group)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (3-D q map) failed with CUresult " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

}  // namespace vlaser
