"""Thin torch wrappers over the single-kernel C-ABI entry points + torch fp32 references.
Used by the `-m gpu` kernel tests and by tools/gpu_diag.py."""
from __future__ import annotations

import ctypes as C
import math

import torch

from vlaser_b200 import _lib
from vlaser_b200._lib import (AttnArgs, GemmArgs, EPI_GELU_BF16, EPI_PATCH_EMBED, EPI_QKV_ROPE, EPI_RESID_F32,
                              EPI_SILU_BF16, EPI_STORE_BF16, EPI_STORE_F32, EPI_SWIGLU_BF16)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(A, W, epi, block_n=128, simt=False, out=None, ldo=None, bias=None, scale=None, M_dev=None, max_splits=1,
         row_map=None, row_pos=None, row_kvrow=None, cos_tab=None, sin_tab=None, k_cache=None, v_cache=None,
         kv_head_stride=0, n_q_heads=0, n_kv_heads=0, pos_emb=None, patches=0, gamma_next=None, a_next=None,
         ss_next=None, tile_counters=None, row_ss=None, inv_width=0.0, eps=0.0, stages=0, pair=False,
         partial=None, partial_stride=0, partial_splits=None, ss_slots=0):
    lib = _lib.load()
    tma_out = False
    if block_n in (512, 513, 515, 517, 521, 525):  # test shorthand for the CTA-pair kernel; 513 = TMA epilogue, 515 = + n-fastest
        # tiles, 517 = TMA epilogue with 256 x 128 tiles (residual epilogue); 521 / 525 = the half-SM variant (one tile per CTA
        # pair, static M) with 256- / 128-column tiles
        block_n, pair, tma_out = 256, {515: 3, 517: 5, 521: 9, 525: 13}.get(block_n, 1), block_n >= 513
    a = GemmArgs()
    a.A, a.W = _p(A), _p(W)
    a.M, a.K = A.shape
    a.N = W.shape[0]
    a.M_dev = _p(M_dev)
    a.epi, a.block_n, a.max_splits, a.simt = epi, block_n, max_splits, int(simt)
    a.out, a.ldo = _p(out), (ldo if ldo is not None else out.shape[-1])
    a.bias, a.scale, a.row_map = _p(bias), _p(scale), _p(row_map)
    a.row_pos, a.row_kvrow, a.cos_tab, a.sin_tab = _p(row_pos), _p(row_kvrow), _p(cos_tab), _p(sin_tab)
    a.k_cache, a.v_cache, a.kv_head_stride = _p(k_cache), _p(v_cache), kv_head_stride
    a.n_q_heads, a.n_kv_heads = n_q_heads, n_kv_heads
    a.pos_emb, a.patches = _p(pos_emb), patches
    a.gamma_next, a.a_next, a.ss_next, a.tile_counters = _p(gamma_next), _p(a_next), _p(ss_next), _p(tile_counters)
    a.lda_next = a_next.shape[-1] if a_next is not None else 0
    a.row_ss, a.inv_width, a.eps, a.stages = _p(row_ss), inv_width, eps, stages
    a.pair, a.tma_out = int(pair), int(tma_out)
    a.partial, a.partial_stride, a.partial_splits = _p(partial), partial_stride, _p(partial_splits)
    a.ss_slots = ss_slots
    _lib.check(lib.vlaser_op_gemm(C.byref(a), _stream()), "op_gemm")
    return out


def interleave_gate_up(gate, up):
    """Row layout the SwiGLU epilogue expects: 32 gate rows, 32 up rows, ..."""
    n, k = gate.shape
    return torch.stack([gate.view(n // 32, 32, k), up.view(n // 32, 32, k)], dim=1).reshape(2 * n, k).contiguous()


def rope_tables(max_pos, theta=1e6, device="cuda"):
    inv = 1.0 / (theta ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))
    ang = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv[None, :]
    return ang.cos().to(device).contiguous(), ang.sin().to(device).contiguous()


def rope_ref(x, cos, sin):
    """x [..., 128] fp32; cos/sin [..., 64] (rotate-half)."""
    x1, x2 = x[..., :64], x[..., 64:]
    return torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1)


def attention(q, k, v, out, head_dim, group, n_kv_heads, batch, scale, q_row_stride, kv_row_stride, kv_head_stride,
              out_row_stride, uni_q_len=0, uni_kv_len=0, max_q_len=0, q_start=None, q_len=None, kv_start=None,
              kv_len=None, q2=None, q2_row_stride=0, out2=None, out2_row_stride=0, causal=0, small_rows=0):
    lib = _lib.load()
    a = AttnArgs()
    a.causal = causal
    a.small_rows = small_rows
    a.q, a.q_row_stride, a.q2, a.q2_row_stride = _p(q), q_row_stride, _p(q2), q2_row_stride
    a.k, a.v, a.kv_row_stride, a.kv_head_stride = _p(k), _p(v), kv_row_stride, kv_head_stride
    a.out, a.out_row_stride, a.out2, a.out2_row_stride = _p(out), out_row_stride, _p(out2), out2_row_stride
    a.q_start, a.q_len, a.kv_start, a.kv_len = _p(q_start), _p(q_len), _p(kv_start), _p(kv_len)
    a.uni_q_len, a.uni_kv_len, a.max_q_len = uni_q_len, uni_kv_len, max_q_len
    a.group, a.n_kv_heads, a.batch, a.head_dim, a.scale = group, n_kv_heads, batch, head_dim, scale
    _lib.check(lib.vlaser_op_attention(C.byref(a), _stream()), "op_attention")
    return out


def decode_attention(q, k_cache, v_cache, kv_start, kv_len, n_heads, n_kv_heads, scale, chunk=256):
    """q bf16 [B, n_heads*128] (one query row per sample), caches bf16 [n_kv_heads, kv_rows, 128] -> out like q."""
    lib = _lib.load()
    B = q.shape[0]
    out = torch.zeros_like(q)
    max_kv = int(kv_len.max())
    splits = (max_kv + chunk - 1) // chunk
    group = n_heads // n_kv_heads
    ws_o = torch.empty(B * n_kv_heads * splits * group * 128, dtype=torch.float32, device=q.device)
    ws_ml = torch.empty(B * n_kv_heads * splits * group * 2, dtype=torch.float32, device=q.device)
    q_start = torch.arange(B, dtype=torch.int32, device=q.device)
    _lib.check(lib.vlaser_op_decode_attention(_p(q), _p(k_cache), _p(v_cache), _p(out), _p(kv_start), _p(kv_len), B, n_heads,
                                              n_kv_heads, k_cache.shape[1], max_kv, chunk, scale, _p(ws_o), _p(ws_ml), _p(q_start),
                                              _stream()), "op_decode_attention")
    return out


def vit_attention(qkv, images, tokens, heads, scale, tc=False):
    """qkv bf16 [images*tokens, 3*heads*64] -> out bf16 [images*tokens, heads*64] (InternViT fast path; tc: the tcgen05
    kernel, 257 tokens)."""
    lib = _lib.load()
    out = torch.zeros(images * tokens, heads * 64, dtype=torch.bfloat16, device=qkv.device)
    fn = lib.vlaser_op_vit_attention_tc if tc else lib.vlaser_op_vit_attention
    _lib.check(fn(_p(qkv), _p(out), images, tokens, heads, scale, _stream()), "op_vit_attention")
    return out


def prefill_attention_tc(q, k_cache, v_cache, q_start, q_len, kv_start, n_heads, n_kv_heads, scale):
    """q bf16 [rows, n_heads*128], caches bf16 [n_kv_heads, kv_rows, 128] -> out like q (tcgen05 prefill kernel)."""
    lib = _lib.load()
    out = torch.zeros_like(q)
    _lib.check(lib.vlaser_op_prefill_attention_tc(_p(q), _p(k_cache), _p(v_cache), _p(out), _p(q_start), _p(q_len), _p(kv_start),
                                                  q.shape[0], n_heads, n_kv_heads, k_cache.shape[1], q_len.numel(), scale,
                                                  _stream()), "op_prefill_attention_tc")
    return out


def prefill_attention_tc_long(q, k_cache, v_cache, q_start, q_len, kv_start, n_heads, n_kv_heads, scale, max_tokens, min_len=1,
                              causal=0):
    """Two-pass tcgen05 prefill kernel for samples of any length (bidirectional or causal); out like q, untouched rows 0."""
    lib = _lib.load()
    out = torch.zeros_like(q)
    _lib.check(lib.vlaser_op_prefill_attention_tc_long(_p(q), _p(k_cache), _p(v_cache), _p(out), _p(q_start), _p(q_len), _p(kv_start),
                                                       q.shape[0], n_heads, n_kv_heads, k_cache.shape[1], q_len.numel(), max_tokens,
                                                       min_len, causal, scale, _stream()), "op_prefill_attention_tc_long")
    return out


def layernorm(x, gamma, beta, eps, shuffle_grid=0, shuffle_channels=0, rows=None, width=None):
    lib = _lib.load()
    rows = x.shape[0] if rows is None else rows
    width = x.shape[1] if width is None else width
    out = torch.empty(rows, width, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.vlaser_op_layernorm(_p(x), _p(gamma), _p(beta), _p(out), rows, width, eps, shuffle_grid,
                                       shuffle_channels, _stream()), "op_layernorm")
    return out


def rmsnorm(x, gamma, eps):
    lib = _lib.load()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.vlaser_op_rmsnorm(_p(x), _p(gamma), _p(out), x.shape[0], x.shape[1], eps, _stream()), "op_rmsnorm")
    return out


def splitk_reduce(x, partial, partial_stride, partial_splits, gamma=None, eps=0.0):
    """x += partials in split order; with gamma also returns bf16 RMSNorm(x) * gamma (the fused consumer)."""
    lib = _lib.load()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if gamma is not None else None
    _lib.check(lib.vlaser_op_splitk_reduce(_p(x), x.shape[0], x.shape[1], _p(partial), partial_stride,
                                           _p(partial_splits), _p(gamma), _p(out), eps, _stream()), "op_splitk_reduce")
    return out


def pixel_shuffle(x):
    """x [n,g,g,C] fp32/bf16 -> [n,g/2,g/2,4C]"""
    lib = _lib.load()
    n, g, _, Cc = x.shape
    out = torch.empty(n, g // 2, g // 2, 4 * Cc, dtype=x.dtype, device=x.device)
    _lib.check(lib.vlaser_op_pixel_shuffle(_p(x), _p(out), n, g, Cc, x.element_size(), _stream()), "op_pixel_shuffle")
    return out


def im2col(pix, ps, kpad):
    lib = _lib.load()
    n, _, px, _ = pix.shape
    g = px // ps
    out = torch.empty(n * g * g, kpad, dtype=torch.bfloat16, device=pix.device)
    dt = _lib.BF16 if pix.dtype == torch.bfloat16 else _lib.F32
    _lib.check(lib.vlaser_op_im2col(_p(pix), dt, _p(out), n, px, ps, kpad, _stream()), "op_im2col")
    return out


def euler_step(x, gamma, Wdec, bdec, action, out, H, horizon, dt, eps, last_step, clip):
    lib = _lib.load()
    rows, width = x.shape
    A = Wdec.shape[0]
    odt = _lib.BF16 if out.dtype == torch.bfloat16 else _lib.F32
    _lib.check(lib.vlaser_op_euler_step(_p(x), _p(gamma), _p(Wdec), _p(bdec), _p(action), _p(out), odt, rows, width, A,
                                        H, horizon, dt, eps, int(last_step), int(clip is not None),
                                        float(clip or 0.0), _stream()), "op_euler_step")


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_err(a, b):
    return float((a.double() - b.double()).abs().max())
