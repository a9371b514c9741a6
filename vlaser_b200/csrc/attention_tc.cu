// InternViT self-attention on the 5th-generation tensor cores (modeling_intern_vit.py:210-227, _naive_attn; no mask):
// head_dim 64, N = 257 tokens per image (224 px: 256 patches + CLS). See attention.cuh (vit_attention_tc_launch).
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

namespace vlaser {
namespace {

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
      for (int c = 0; c < 8; ++c) {
        uint32_t r0[32];
        tmem_ld_32x32(taddr + c * 32, r0);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(r0[2 * j]), sl2, -moff));
          const float p1 = ex2_approx(fmaf(__uint_as_float(r0[2 * j + 1]), sl2, -moff));
          sum += p0 + p1;
          pk[j] = pack_bf16(p0, p1);
        }
        tmem_st_32x16(taddr + c * 16, pk);
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

}  // namespace

int vit_attention_tc_launch(const CUtensorMap& tmQKV, const CUtensorMap& tmOut, const __nv_bfloat16* qkv,
                            __nv_bfloat16* out, int images, int tokens, int heads, float scale, cudaStream_t stream,
                            long long* trace) {
  VL_CHECK(tokens == 257, "vit_attention_tc: 257 tokens per image (224 px)");
  VL_CHECK(images > 0 && heads > 0, "vit_attention_tc: empty problem");
  static bool attr_set = false;
  if (!attr_set) {
    VL_CUDA_CHECK(cudaFuncSetAttribute(vit_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VT_SMEM));
    attr_set = true;
  }
  const int grid = std::min(images * heads, gemm_num_sms());
  VL_CUDA_CHECK(launch_kernel(vit_attention_tc_kernel, dim3(grid), dim3(VT_THREADS), VT_SMEM, stream, tmQKV, tmOut, qkv, out,
                              images, heads, scale, trace));
  return 0;
}

}  // namespace vlaser
