"""-m gpu: every CUDA kernel behind the C ABI against a plain torch fp32 reference of the same op
(bit-exact for the indexing kernels)."""
from __future__ import annotations

import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import ops  # noqa: E402
from tests.ops import (EPI_GELU_BF16, EPI_PATCH_EMBED, EPI_QKV_ROPE, EPI_RESID_F32, EPI_SILU_BF16,  # noqa: E402
                       EPI_STORE_BF16, EPI_STORE_F32, EPI_SWIGLU_BF16)

DEV = "cuda"
BF16_TOL = 6e-3  # relative L2 of a bf16-rounded output vs the fp32 reference


def _rand(*shape, scale=1.0, dtype=torch.bfloat16, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


def _ref_mm(A, W):
    return A.float() @ W.float().t()


VARIANTS = [("tcgen05", 128, False), ("tcgen05", 256, False), ("pair", 512, False), ("pair_tma", 513, False),
            ("simt", 128, True)]


@pytest.mark.parametrize("name,bn,simt", VARIANTS)
@pytest.mark.parametrize("M,N,K", [(300, 256, 128), (1000, 1024, 512), (129, 768, 768), (4096, 2048, 1024)])
def test_gemm_store_bias(name, bn, simt, M, N, K):
    if N % min(bn, 256):
        pytest.skip("N not a multiple of block_n")
    A, W = _rand(M, K), _rand(N, K, scale=0.05)
    bias = _rand(N, dtype=torch.float32, seed=3)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, W, EPI_STORE_BF16, bn, simt, out=out, bias=bias)
    ref = _ref_mm(A, W) + bias
    err = ops.rel_err(out, ref)
    print(f"gemm_store {name} bn={bn} {M}x{N}x{K}: rel={err:.2e}")
    assert err < BF16_TOL


@pytest.mark.parametrize("name,bn,simt", VARIANTS)
def test_gemm_gelu_silu(name, bn, simt):
    M, N, K = 777, 512, 256
    A, W = _rand(M, K), _rand(N, K, scale=0.1)
    bias = _rand(N, dtype=torch.float32, seed=5)
    for epi, fn in ((EPI_GELU_BF16, torch.nn.functional.gelu), (EPI_SILU_BF16, torch.nn.functional.silu)):
        out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        ops.gemm(A, W, epi, bn, simt, out=out, bias=bias)
        ref = fn(_ref_mm(A, W) + bias)
        err = ops.rel_err(out, ref)
        print(f"gemm_act epi={epi} {name} bn={bn}: rel={err:.2e}")
        assert err < BF16_TOL


@pytest.mark.parametrize("name,bn,simt", VARIANTS)
@pytest.mark.parametrize("splits", [1, 8])
def test_gemm_resid(name, bn, simt, splits):
    if bn == 512 and splits > 1:
        pytest.skip("pair-kernel split-K needs the TMA reduce-add epilogue")
    M, N, K = 256, 768, 2048
    A, W = _rand(M, K), _rand(N, K, scale=0.05)
    bias = _rand(N, dtype=torch.float32, seed=7)
    scale = _rand(N, dtype=torch.float32, seed=8) * 0.1
    x0 = _rand(M, N, dtype=torch.float32, seed=9)
    out = x0.clone()
    ops.gemm(A, W, EPI_RESID_F32, bn, simt, out=out, bias=bias, scale=scale, max_splits=splits)
    ref = x0 + scale * (_ref_mm(A, W) + bias)
    err = ops.rel_err(out, ref)
    print(f"gemm_resid {name} bn={bn} splits={splits}: rel={err:.2e} max={ops.max_err(out, ref):.2e}")
    assert err < 1e-5


@pytest.mark.parametrize("name,simt", [("tcgen05", False), ("simt", True)])
@pytest.mark.parametrize("splits", [1, 32])
@pytest.mark.parametrize("M", [64, 256])
def test_gemm_fused_rmsnorm_chain(name, simt, splits, M):
    """Producer (o/down-proj + residual) emits bf16(x*gamma) and row sums of squares; the consumer (gate/up with
    SwiGLU) scales its accumulator by rsqrt(ss/width + eps): together == RMSNorm between the two GEMMs
    (HF Qwen2RMSNorm as used at joint_model.py:166,209)."""
    if simt and splits > 1:
        pytest.skip("the cross-check kernel has no split-K")
    N, K, F, eps = 768, 1536, 512, 1e-6
    A, W = _rand(M, K), _rand(N, K, scale=0.05)
    x0 = _rand(M, N, dtype=torch.float32, seed=9)
    gamma = 1.0 + 0.1 * _rand(N, dtype=torch.float32, seed=10)
    x = x0.clone()
    a_next = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    ss = torch.zeros(M, dtype=torch.float32, device=DEV)
    ctr = torch.zeros(64, dtype=torch.int32, device=DEV)
    for rep in range(2):  # second pass checks that the tile counters reset themselves
        x.copy_(x0)
        ss.zero_()
        ops.gemm(A, W, EPI_RESID_F32, 128, simt, out=x, max_splits=splits, gamma_next=gamma, a_next=a_next, ss_next=ss,
                 tile_counters=ctr, stages=3)
        x_ref = x0 + _ref_mm(A, W)
        assert ops.rel_err(x, x_ref) < 1e-5
        assert ops.rel_err(ss, (x_ref * x_ref).sum(-1)) < 1e-5
        assert ops.rel_err(a_next, x_ref * gamma) < BF16_TOL
        assert int(ctr.abs().sum()) == 0
    gate, up = _rand(F, N, scale=0.1, seed=1), _rand(F, N, scale=0.1, seed=2)
    out = torch.empty(M, F, dtype=torch.bfloat16, device=DEV)
    ops.gemm(a_next, ops.interleave_gate_up(gate, up), EPI_SWIGLU_BF16, 128, simt, out=out, row_ss=ss,
             inv_width=1.0 / N, eps=eps, stages=3)
    xn = x_ref * torch.rsqrt((x_ref * x_ref).mean(-1, keepdim=True) + eps) * gamma
    ref = torch.nn.functional.silu(xn @ gate.float().t()) * (xn @ up.float().t())
    err = ops.rel_err(out, ref)
    print(f"fused_rmsnorm_chain {name} M={M} splits={splits}: rel={err:.2e}")
    assert err < 1.5 * BF16_TOL


@pytest.mark.parametrize("M,N,K,F", [(700, 1536, 512, 512), (256, 768, 1024, 256), (5570, 1536, 1536, 1024)])
def test_gemm_pair_fused_rmsnorm_chain_slots(M, N, K, F):
    """Large-M form of the fused RMSNorm chain on the CTA-pair kernel (the LLM prefill): the residual projection reads
    x, writes x_new, bf16(x_new * gamma) and one partial sum of squares per (256-column tile, column half) with plain
    stores; the consumer (gate/up with SwiGLU) adds the slots in order and scales its accumulator rows by rstd.
    Together == RMSNorm between the two GEMMs (HF Qwen2RMSNorm, joint_model.py:166,209); bit-identical from call to call."""
    eps = 1e-6
    A, W = _rand(M, K), _rand(N, K, scale=0.05)
    bias = _rand(N, dtype=torch.float32, seed=7)
    x0 = _rand(M, N, dtype=torch.float32, seed=9)
    gamma = 1.0 + 0.1 * _rand(N, dtype=torch.float32, seed=10)
    slots = N // 128
    gate, up = _rand(F, N, scale=0.1, seed=1), _rand(F, N, scale=0.1, seed=2)
    runs = []
    for rep in range(2):
        x = x0.clone()
        a_next = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        ss = torch.full((M, slots), float("nan"), dtype=torch.float32, device=DEV)
        ops.gemm(A, W, EPI_RESID_F32, 513, False, out=x, bias=bias, gamma_next=gamma, a_next=a_next, ss_next=ss, ss_slots=slots)
        x_ref = x0 + _ref_mm(A, W) + bias
        assert ops.rel_err(x, x_ref) < 1e-5
        assert ops.rel_err(ss.sum(-1), (x_ref * x_ref).sum(-1)) < 1e-5
        assert ops.rel_err(a_next, x_ref * gamma) < BF16_TOL
        out = torch.empty(M, F, dtype=torch.bfloat16, device=DEV)
        ops.gemm(a_next, ops.interleave_gate_up(gate, up), EPI_SWIGLU_BF16, 513, False, out=out, row_ss=ss,
                 inv_width=1.0 / N, eps=eps, ss_slots=slots)
        xn = x_ref * torch.rsqrt((x_ref * x_ref).mean(-1, keepdim=True) + eps) * gamma
        ref = torch.nn.functional.silu(xn @ gate.float().t()) * (xn @ up.float().t())
        err = ops.rel_err(out, ref)
        print(f"pair fused_rmsnorm_chain M={M} N={N}: rel={err:.2e}")
        assert err < 1.5 * BF16_TOL
        runs.append((x, ss, a_next, out))
    for a, b in zip(runs[0], runs[1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("bn", [128, 513, 517])
@pytest.mark.parametrize("M,N,K", [(256, 768, 8960), (64, 768, 1536), (257, 1024, 4096), (200, 1536, 1536)])
def test_gemm_splitk_partials_deterministic(bn, M, N, K):
    """Deterministic split-K: the splits leave partial tiles in a workspace, the consumer adds them to the residual
    stream in split order -- alone (splitk_reduce) or fused into the RMSNorm that follows. The result equals the
    un-split product to fp32 rounding and is bit-identical from call to call (the fp32 atomics it replaces are not)."""
    A, W = _rand(M, K), _rand(N, K, scale=0.05)
    bias = torch.randn(N, device=DEV)
    scale = torch.rand(N, device=DEV) + 0.5
    gamma = torch.rand(N, device=DEV) + 0.5
    x0 = torch.randn(M, N, device=DEV)
    rows_p = (M + 255) // 256 * 256
    splits = 16
    part = torch.full((splits, rows_p, N), float("nan"), device=DEV)
    nsp = torch.full((1,), -1, dtype=torch.int32, device=DEV)
    ref = x0 + (_ref_mm(A, W) + bias) * scale
    outs = []
    for rep in range(3):
        x = x0.clone()
        ops.gemm(A, W, EPI_RESID_F32, bn, False, out=x, bias=bias, scale=scale, max_splits=splits, partial=part,
                 partial_stride=rows_p * N, partial_splits=nsp)
        used = int(nsp.item())
        assert 2 <= used <= splits, used
        assert torch.equal(x, x0), "with partials the residual stream is only touched by the consumer"
        if rep < 2:
            ops.splitk_reduce(x, part, rows_p * N, nsp)
            assert ops.rel_err(x, ref) < 1e-5
            outs.append(x)
        else:
            xn = ops.splitk_reduce(x, part, rows_p * N, nsp, gamma=gamma, eps=1e-6)
            assert torch.equal(x, outs[0])
            want = ref * torch.rsqrt((ref * ref).mean(-1, keepdim=True) + 1e-6) * gamma
            assert ops.rel_err(xn, want) < BF16_TOL
    assert torch.equal(outs[0], outs[1])
    # one split: the GEMM updates the stream itself and reports 0 partials
    x = x0.clone()
    ops.gemm(A, W, EPI_RESID_F32, bn, False, out=x, bias=bias, scale=scale, max_splits=1, partial=part,
             partial_stride=rows_p * N, partial_splits=nsp)
    assert int(nsp.item()) == 0 and ops.rel_err(x, ref) < 1e-5
    x1 = x.clone()
    ops.splitk_reduce(x, part, rows_p * N, nsp)
    assert torch.equal(x, x1)


@pytest.mark.parametrize("name,bn,simt", VARIANTS)
def test_gemm_dynamic_m(name, bn, simt):
    """Live row count from device memory: rows >= M_dev must stay untouched (the TMA epilogue of the pair kernel
    stores whole 32-row groups: there only rows past the last live 256-row tile are guaranteed untouched)."""
    cap, M, N, K = 640, 333, 512, 256
    A, W = _rand(cap, K), _rand(N, K, scale=0.1)
    m_dev = torch.tensor([M], dtype=torch.int32, device=DEV)
    out = torch.full((cap, N), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, W, EPI_STORE_BF16, bn, simt, out=out, M_dev=m_dev)
    ref = _ref_mm(A[:M], W)
    assert ops.rel_err(out[:M], ref) < BF16_TOL
    lim = M if bn != 513 else (M + 255) // 256 * 256
    assert torch.all(out[lim:] == 7.0)


@pytest.mark.parametrize("M,N,K,splits", [(256, 768, 8960, 12), (256, 768, 1536, 12), (200, 768, 1536, 12), (130, 1024, 4096, 9)])
def test_gemm_half_sm_resid_partials(M, N, K, splits):
    """Half-SM variant of the pair kernel (two CTAs fit on an SM: 4 epilogue warps, packed ring, one accumulator stage,
    staging in the idle ring), residual epilogue with deterministic split-K partials at one tile per CTA pair: same
    numbers as the full-SM kernel with the same split count, bit-identical from call to call."""
    A, W = _rand(M, K), _rand(N, K, scale=0.05)
    bias = torch.randn(N, device=DEV)
    scale = torch.rand(N, device=DEV) + 0.5
    x0 = torch.randn(M, N, device=DEV)
    rows_p = (M + 255) // 256 * 256
    ref = x0 + (_ref_mm(A, W) + bias) * scale
    res = {}
    for bn in (525, 517):
        part = torch.full((splits, rows_p, N), float("nan"), device=DEV)
        nsp = torch.full((1,), -1, dtype=torch.int32, device=DEV)
        outs = []
        for rep in range(2):
            x = x0.clone()
            ops.gemm(A, W, EPI_RESID_F32, bn, False, out=x, bias=bias, scale=scale, max_splits=splits, partial=part,
                     partial_stride=rows_p * N, partial_splits=nsp)
            assert 2 <= int(nsp.item()) <= splits
            assert torch.equal(x, x0)
            ops.splitk_reduce(x, part, rows_p * N, nsp)
            outs.append(x)
        assert torch.equal(outs[0], outs[1])
        assert ops.rel_err(outs[0], ref) < 1e-5
        res[bn] = (outs[0], int(nsp.item()))
    assert res[525][1] == res[517][1]
    assert torch.equal(res[525][0], res[517][0]), "same splits, same association: the two variants agree bit for bit"


@pytest.mark.parametrize("M,F,K", [(256, 8960, 768), (256, 512, 256), (140, 2816, 768)])
def test_gemm_half_sm_swiglu(M, F, K):
    """Half-SM variant, SwiGLU epilogue (the gate/up launch of a denoise step: N = 2 F accumulator columns, one 256 x 256
    tile per CTA pair): equal to the full-SM pair kernel bit for bit."""
    A = _rand(M, K)
    gate, up = _rand(F, K, scale=0.1, seed=1), _rand(F, K, scale=0.1, seed=2)
    W = ops.interleave_gate_up(gate, up)
    out = torch.empty(M, F, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, W, EPI_SWIGLU_BF16, 521, False, out=out)
    ref = torch.nn.functional.silu(_ref_mm(A, gate)) * _ref_mm(A, up)
    assert ops.rel_err(out, ref) < BF16_TOL
    out2 = torch.empty(M, F, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, W, EPI_SWIGLU_BF16, 513, False, out=out2)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("name,bn,simt", VARIANTS)
def test_gemm_swiglu(name, bn, simt):
    M, F, K = 500, 512, 256
    A = _rand(M, K)
    gate, up = _rand(F, K, scale=0.1, seed=1), _rand(F, K, scale=0.1, seed=2)
    W = ops.interleave_gate_up(gate, up)
    out = torch.empty(M, F, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, W, EPI_SWIGLU_BF16, bn, simt, out=out)
    ref = torch.nn.functional.silu(_ref_mm(A, gate)) * _ref_mm(A, up)
    err = ops.rel_err(out, ref)
    print(f"gemm_swiglu {name} bn={bn}: rel={err:.2e}")
    assert err < BF16_TOL


@pytest.mark.parametrize("name,simt", [("tcgen05", False), ("pair", False), ("simt", True)])
def test_gemm_qkv_rope(name, simt):
    M, K, nq, nkv = 200, 256, 6, 1
    N = (nq + 2 * nkv) * 128
    A, W = _rand(M, K), _rand(N, K, scale=0.1)
    bias = _rand(N, dtype=torch.float32, seed=4)
    cos, sin = ops.rope_tables(512)
    g = torch.Generator().manual_seed(0)
    pos = torch.randint(1, 400, (M,), generator=g, dtype=torch.int32).to(DEV)
    kvrow = torch.randperm(300, generator=g)[:M].to(torch.int32).to(DEV)
    rows_cap = 300
    q = torch.zeros(M, nq * 128, dtype=torch.bfloat16, device=DEV)
    kc = torch.zeros(nkv, rows_cap, 128, dtype=torch.bfloat16, device=DEV)
    vc = torch.zeros(nkv, rows_cap, 128, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, W, EPI_QKV_ROPE, 512 if name == "pair" else 128, simt, out=q, bias=bias, row_pos=pos, row_kvrow=kvrow, cos_tab=cos, sin_tab=sin,
             k_cache=kc, v_cache=vc, kv_head_stride=rows_cap * 128, n_q_heads=nq, n_kv_heads=nkv)
    y = (_ref_mm(A, W) + bias).view(M, nq + 2 * nkv, 128)
    c, s = cos[pos.long()][:, None, :], sin[pos.long()][:, None, :]
    q_ref = ops.rope_ref(y[:, :nq], c, s).reshape(M, -1)
    k_ref = ops.rope_ref(y[:, nq:nq + nkv], c, s)
    v_ref = y[:, nq + nkv:]
    print(f"qkv_rope {name}: q={ops.rel_err(q, q_ref):.2e} k={ops.rel_err(kc[:, kvrow.long()].transpose(0, 1), k_ref):.2e}")
    assert ops.rel_err(q, q_ref) < BF16_TOL
    assert ops.rel_err(kc[:, kvrow.long()].transpose(0, 1), k_ref) < BF16_TOL
    assert ops.rel_err(vc[:, kvrow.long()].transpose(0, 1), v_ref) < BF16_TOL


@pytest.mark.parametrize("name,bn,simt", VARIANTS)
def test_gemm_patch_embed_and_store_f32(name, bn, simt):
    n_img, P, N, K = 3, 256, 256, 640
    A, W = _rand(n_img * P, K), _rand(N, K, scale=0.05)
    bias = _rand(N, dtype=torch.float32, seed=2)
    pos = _rand(P + 1, N, dtype=torch.float32, seed=3)
    out = torch.zeros(n_img * (P + 1), N, dtype=torch.float32, device=DEV)
    ops.gemm(A, W, EPI_PATCH_EMBED, bn, simt, out=out, bias=bias, pos_emb=pos, patches=P)
    ref = (_ref_mm(A, W) + bias).view(n_img, P, N) + pos[1:]
    assert ops.rel_err(out.view(n_img, P + 1, N)[:, 1:], ref) < 1e-5
    assert torch.all(out.view(n_img, P + 1, N)[:, 0] == 0)
    # STORE_F32 with a scatter map (row -> packed row, -1 = drop)
    M = n_img * P
    g = torch.Generator().manual_seed(1)
    rmap = torch.randperm(2 * M, generator=g)[:M].to(torch.int32)
    rmap[::7] = -1
    rmap = rmap.to(DEV)
    out2 = torch.zeros(2 * M, N, dtype=torch.float32, device=DEV)
    ops.gemm(A, W, EPI_STORE_F32, bn, simt, out=out2, bias=bias, row_map=rmap)
    ref2 = torch.zeros_like(out2)
    keep = rmap >= 0
    ref2[rmap[keep].long()] = (_ref_mm(A, W) + bias)[keep]
    assert ops.rel_err(out2, ref2) < 1e-5


def _attn_ref(q, k, v, scale, kv_limit):
    """q [Lq,hd] k,v [Lk,hd] fp32, keys < kv_limit visible."""
    s = (q @ k.t()) * scale
    s[:, kv_limit:] = float("-inf")
    return torch.softmax(s, dim=-1) @ v


def test_attention_vit():
    B, heads, N, hd, d = 3, 4, 257, 64, 256
    qkv = _rand(B * N, 3 * d)
    out = torch.zeros(B * N, d, dtype=torch.bfloat16, device=DEV)
    ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], out, hd, 1, heads, B, hd ** -0.5, 3 * d, 3 * d, hd, d,
                  uni_q_len=N, uni_kv_len=N, max_q_len=N)
    x = qkv.float().view(B, N, 3, heads, hd)
    ref = torch.empty(B, N, heads, hd, device=DEV)
    for b in range(B):
        for h in range(heads):
            ref[b, :, h] = _attn_ref(x[b, :, 0, h], x[b, :, 1, h], x[b, :, 2, h], hd ** -0.5, N)
    err = ops.rel_err(out.view(B, N, heads, hd), ref)
    print(f"attention_vit rel={err:.2e}")
    assert err < 1e-2


@pytest.mark.parametrize("N", [257, 200, 128, 320])
def test_vit_attention_fast_path(N):
    """One-CTA-per-(image, head) kernel: 257 = 16 MMA row tiles + 1 CUDA-core row, 4 key blocks + 1 ragged key;
    200 / 128 / 320 exercise the other tail shapes."""
    B, heads, hd = 3, 4, 64
    d = heads * hd
    qkv = _rand(B * N, 3 * d, seed=N)
    out = ops.vit_attention(qkv, B, N, heads, hd ** -0.5)
    x = qkv.float().view(B, N, 3, heads, hd)
    ref = torch.empty(B, N, heads, hd, device=DEV)
    for b in range(B):
        for h in range(heads):
            ref[b, :, h] = _attn_ref(x[b, :, 0, h], x[b, :, 1, h], x[b, :, 2, h], hd ** -0.5, N)
    err = ops.rel_err(out.view(B, N, heads, hd), ref)
    rows = (out.view(B, N, heads, hd).float() - ref).norm(dim=-1) / ref.norm(dim=-1)
    print(f"vit_attention N={N} rel={err:.2e} worst row={float(rows.max()):.2e}")
    assert err < 1e-2 and float(rows.max()) < 3e-2


@pytest.mark.parametrize("B,heads", [(1, 4), (3, 16), (40, 16)])
def test_vit_attention_tcgen05(B, heads):
    """tcgen05 kernel (attention_tc.cu), 257 tokens: two 128-row query tiles on the tensor cores with the scores /
    probabilities / output in TMEM, the left-over key as a rank-1 CUDA-core term, the left-over query row on CUDA cores.
    B * heads below, at and well above the SM count (one item per CTA, a persistent loop of several items)."""
    N, hd = 257, 64
    d = heads * hd
    qkv = _rand(B * N, 3 * d, seed=B)
    out = ops.vit_attention(qkv, B, N, heads, hd ** -0.5, tc=True)
    x = qkv.float().view(B, N, 3, heads, hd).permute(0, 3, 2, 1, 4)          # [B, heads, 3, N, hd]
    s = torch.einsum("bhqd,bhkd->bhqk", x[:, :, 0], x[:, :, 1]) * hd ** -0.5
    ref = torch.einsum("bhqk,bhkd->bhqd", torch.softmax(s, dim=-1), x[:, :, 2]).permute(0, 2, 1, 3)  # [B, N, heads, hd]
    got = out.view(B, N, heads, hd).float()
    err = ops.rel_err(got, ref)
    rows = (got - ref).norm(dim=-1) / ref.norm(dim=-1)
    print(f"vit_attention_tcgen05 B={B} heads={heads} rel={err:.2e} worst row={float(rows.max()):.2e}")
    assert err < 1e-2 and float(rows.max()) < 3e-2
    legacy = ops.vit_attention(qkv, B, N, heads, hd ** -0.5).view(B, N, heads, hd).float()
    assert ops.rel_err(got, legacy) < 1e-2
    assert torch.equal(out, ops.vit_attention(qkv, B, N, heads, hd ** -0.5, tc=True)), "repeat call differs"


@pytest.mark.parametrize("B,heads,N", [(1, 4, 1025), (3, 16, 1025), (21, 16, 1025), (2, 3, 513)])
def test_vit_attention_tcgen05_long_rows(B, heads, N):
    """tcgen05 two-pass kernel (attention_tc.cu), 448 px tiles: 1025 tokens = 8 query tiles of 128 rows against 1024 keys
    resident in shared memory + the left-over key (rank-1 term) + the left-over query row (CUDA-core kernel). Items below,
    at and above the SM count (contiguous item ranges: several query pairs of a head per CTA, heads split between CTAs)."""
    hd = 64
    d = heads * hd
    qkv = _rand(B * N, 3 * d, seed=B + N)
    out = ops.vit_attention(qkv, B, N, heads, hd ** -0.5, tc=True)
    x = qkv.float().view(B, N, 3, heads, hd).permute(0, 3, 2, 1, 4)          # [B, heads, 3, N, hd]
    s = torch.einsum("bhqd,bhkd->bhqk", x[:, :, 0], x[:, :, 1]) * hd ** -0.5
    ref = torch.einsum("bhqk,bhkd->bhqd", torch.softmax(s, dim=-1), x[:, :, 2]).permute(0, 2, 1, 3)  # [B, N, heads, hd]
    got = out.view(B, N, heads, hd).float()
    err = ops.rel_err(got, ref)
    rows = (got - ref).norm(dim=-1) / ref.norm(dim=-1)
    print(f"vit_attention_tcgen05 N={N} B={B} heads={heads} rel={err:.2e} worst row={float(rows.max()):.2e} "
          f"last row={float(rows[:, -1].max()):.2e}")
    assert err < 1e-2 and float(rows.max()) < 3e-2
    assert torch.equal(out, ops.vit_attention(qkv, B, N, heads, hd ** -0.5, tc=True)), "repeat call differs"
    # peaked rows (one dominant key per query): the two-pass maximum must be the exact one
    qkv2 = qkv.clone().view(B, N, 3, heads, hd)
    qkv2[:, :, 0] *= 6.0
    qkv2 = qkv2.view(B * N, 3 * d).contiguous()
    out2 = ops.vit_attention(qkv2, B, N, heads, hd ** -0.5, tc=True).view(B, N, heads, hd).float()
    x2 = qkv2.float().view(B, N, 3, heads, hd).permute(0, 3, 2, 1, 4)
    s2 = torch.einsum("bhqd,bhkd->bhqk", x2[:, :, 0], x2[:, :, 1]) * hd ** -0.5
    ref2 = torch.einsum("bhqk,bhkd->bhqd", torch.softmax(s2, dim=-1), x2[:, :, 2]).permute(0, 2, 1, 3)
    assert ops.rel_err(out2, ref2) < 1e-2


@pytest.mark.parametrize("nkv,group,lens", [(2, 6, [85, 1, 130, 256, 87, 21, 22, 300]), (4, 7, [60, 200]), (1, 1, [256, 33]),
                                            (2, 6, [87, 64, 100, 41] * 20)])
def test_prefill_attention_tcgen05(nkv, group, lens):
    """tcgen05 prefill kernel: GQA-packed query tiles, ragged valid lengths from device memory (1 .. 256 tokens; the
    300-token sample is left untouched for the generic kernel), several items per CTA when there are more than SMs."""
    hd, nq = 128, nkv * group
    g = torch.Generator().manual_seed(len(lens))
    B = len(lens)
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    kvoff = [cu[b] + 5 * b for b in range(B)]          # KV rows of a sample: its tokens + room for others (as the engine)
    rows, kv_rows = cu[-1], cu[-1] + 5 * B + 8
    q = (torch.randn(rows, nq * hd, generator=g) * 1.5).bfloat16().to(DEV)
    kc = torch.randn(nkv, kv_rows, hd, generator=g).bfloat16().to(DEV)
    vc = torch.randn(nkv, kv_rows, hd, generator=g).bfloat16().to(DEV)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    out = ops.prefill_attention_tc(q, kc, vc, i32(cu[:-1]), i32(lens), i32(kvoff), nq, nkv, hd ** -0.5)
    worst = 0.0
    for b, n in enumerate(lens):
        got = out[cu[b]: cu[b] + n].float().view(n, nq, hd)
        if n > 256:
            assert torch.all(got == 0), "samples beyond 256 tokens belong to the generic kernel"
            continue
        qq = q[cu[b]: cu[b] + n].float().view(n, nq, hd)
        for h in range(nq):
            kv = h // group
            ref = _attn_ref(qq[:, h], kc[kv, kvoff[b]: kvoff[b] + n].float(), vc[kv, kvoff[b]: kvoff[b] + n].float(), hd ** -0.5, n)
            worst = max(worst, ops.rel_err(got[:, h], ref))
    print(f"prefill_attention_tcgen05 nkv={nkv} group={group}: worst head rel={worst:.2e}")
    assert worst < 1e-2
    assert torch.equal(out, ops.prefill_attention_tc(q, kc, vc, i32(cu[:-1]), i32(lens), i32(kvoff), nq, nkv, hd ** -0.5))


@pytest.mark.parametrize("nkv,group,lens,causal,min_len", [
    (2, 6, [300, 85, 257, 1000, 21, 640], 0, 257), (4, 7, [277, 1302, 2841], 0, 257), (2, 6, [300, 85, 1, 129, 1000], 1, 1),
    (4, 7, [3353, 277, 18, 19], 1, 1), (1, 1, [256, 513], 1, 1), (2, 8, [400] * 40, 0, 1)])
def test_prefill_attention_tcgen05_long(nkv, group, lens, causal, min_len):
    """Two-pass tcgen05 prefill kernel: samples of 1 .. 3353 tokens, bidirectional (samples shorter than min_len are left
    to the short kernel) and causal, GQA groups 1 / 6 / 7 / 8, odd tile counts, more items than SMs."""
    hd, nq = 128, nkv * group
    g = torch.Generator().manual_seed(len(lens) + causal)
    B = len(lens)
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    kvoff = [cu[b] + 5 * b for b in range(B)]
    rows, kv_rows = cu[-1], cu[-1] + 5 * B + 8
    q = (torch.randn(rows, nq * hd, generator=g) * 1.5).bfloat16().to(DEV)
    kc = torch.randn(nkv, kv_rows, hd, generator=g).bfloat16().to(DEV)
    vc = torch.randn(nkv, kv_rows, hd, generator=g).bfloat16().to(DEV)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    run = lambda: ops.prefill_attention_tc_long(q, kc, vc, i32(cu[:-1]), i32(lens), i32(kvoff), nq, nkv, hd ** -0.5, max(lens),
                                                min_len, causal)
    out = run()
    worst = 0.0
    for b, n in enumerate(lens):
        got = out[cu[b]: cu[b] + n].float().view(n, nq, hd)
        if n < min_len:
            assert torch.all(got == 0), "samples below min_len belong to the short kernel"
            continue
        qq = q[cu[b]: cu[b] + n].float().view(n, nq, hd)
        mask = torch.full((n, n), float("-inf"), device=DEV).triu(1) if causal else 0.0
        for h in range(nq):
            kv = h // group
            K, V = kc[kv, kvoff[b]: kvoff[b] + n].float(), vc[kv, kvoff[b]: kvoff[b] + n].float()
            ref = torch.softmax(qq[:, h] @ K.T * hd ** -0.5 + mask, -1) @ V
            worst = max(worst, ops.rel_err(got[:, h], ref))
    print(f"prefill_attention_tcgen05_long nkv={nkv} group={group} causal={causal}: worst head rel={worst:.2e}")
    assert worst < 1e-2
    assert torch.equal(out, run())


def test_attention_prefill_and_denoise():
    """GQA 6:1, ragged lengths, proprio extra row (prefill) then H action queries (denoise)."""
    B, nkv, group, hd, H = 3, 2, 6, 128, 4
    nq = nkv * group
    lens = [85, 1, 130]
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    kvoff = [cu[b] + b * (1 + H) for b in range(B)]
    rows_cap = cu[-1] + B * (1 + H) + 3
    q = _rand(cu[-1], nq * hd, seed=1)
    q2 = _rand(B, nq * hd, seed=2)
    kc = _rand(nkv, rows_cap, hd, seed=3)
    vc = _rand(nkv, rows_cap, hd, seed=4)
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=DEV)
    out = torch.zeros_like(q)
    out2 = torch.zeros_like(q2)
    scale = hd ** -0.5
    ops.attention(q, kc, vc, out, hd, group, nkv, B, scale, nq * hd, hd, rows_cap * hd, nq * hd,
                  max_q_len=max(lens) + 1, q_start=i32(cu[:-1]), q_len=i32(lens), kv_start=i32(kvoff), kv_len=i32(lens),
                  q2=q2, q2_row_stride=nq * hd, out2=out2, out2_row_stride=nq * hd)
    errs = []
    for b in range(B):
        L = lens[b]
        for h in range(nq):
            g = h // group
            K = kc[g, kvoff[b]:kvoff[b] + L + 1].float()
            V = vc[g, kvoff[b]:kvoff[b] + L + 1].float()
            r = _attn_ref(q[cu[b]:cu[b] + L, h * hd:(h + 1) * hd].float(), K, V, scale, L)
            errs.append(ops.rel_err(out[cu[b]:cu[b] + L, h * hd:(h + 1) * hd], r))
            r2 = _attn_ref(q2[b:b + 1, h * hd:(h + 1) * hd].float(), K, V, scale, L + 1)
            errs.append(ops.rel_err(out2[b:b + 1, h * hd:(h + 1) * hd], r2))
    print(f"attention_prefill max rel={max(errs):.2e}")
    assert max(errs) < 1e-2
    # denoise: H queries per sample over prefix + proprio + H action keys
    qa = _rand(B * H, nq * hd, seed=5)
    outa = torch.zeros_like(qa)
    ops.attention(qa, kc, vc, outa, hd, group, nkv, B, scale, nq * hd, hd, rows_cap * hd, nq * hd, max_q_len=H,
                  q_start=i32([b * H for b in range(B)]), q_len=i32([H] * B), kv_start=i32(kvoff),
                  kv_len=i32([L + 1 + H for L in lens]))
    errs = []
    for b in range(B):
        Lk = lens[b] + 1 + H
        for h in range(nq):
            g = h // group
            r = _attn_ref(qa[b * H:(b + 1) * H, h * hd:(h + 1) * hd].float(), kc[g, kvoff[b]:kvoff[b] + Lk].float(),
                          vc[g, kvoff[b]:kvoff[b] + Lk].float(), scale, Lk)
            errs.append(ops.rel_err(outa[b * H:(b + 1) * H, h * hd:(h + 1) * hd], r))
    print(f"attention_denoise max rel={max(errs):.2e}")
    assert max(errs) < 1e-2
    # the same step on the small-row kernel (key blocks split between the warp pairs): windows of 1, 2 and 3+ key blocks
    for lens2 in (lens, [30, 64, 59], [200, 129, 250]):
        kvoff2 = [sum(lens2[:b]) + b * (1 + H) for b in range(B)]
        cap2 = sum(lens2) + B * (1 + H) + 3
        kc2, vc2 = _rand(nkv, cap2, hd, seed=6), _rand(nkv, cap2, hd, seed=7)
        kw = dict(max_q_len=H, q_start=i32([b * H for b in range(B)]), q_len=i32([H] * B), kv_start=i32(kvoff2),
                  kv_len=i32([L + 1 + H for L in lens2]))
        o_small, o_gen = torch.zeros_like(qa), torch.zeros_like(qa)
        ops.attention(qa, kc2, vc2, o_small, hd, group, nkv, B, scale, nq * hd, hd, cap2 * hd, nq * hd, small_rows=1, **kw)
        ops.attention(qa, kc2, vc2, o_gen, hd, group, nkv, B, scale, nq * hd, hd, cap2 * hd, nq * hd, **kw)
        errs = []
        for b in range(B):
            Lk = lens2[b] + 1 + H
            for h in range(nq):
                g = h // group
                r = _attn_ref(qa[b * H:(b + 1) * H, h * hd:(h + 1) * hd].float(), kc2[g, kvoff2[b]:kvoff2[b] + Lk].float(),
                              vc2[g, kvoff2[b]:kvoff2[b] + Lk].float(), scale, Lk)
                errs.append(ops.rel_err(o_small[b * H:(b + 1) * H, h * hd:(h + 1) * hd], r))
        print(f"attention_denoise small-row kernel lens={lens2} max rel={max(errs):.2e}")
        assert max(errs) < 1e-2 and ops.rel_err(o_small, o_gen.float()) < 1e-2
        o_again = torch.zeros_like(qa)
        ops.attention(qa, kc2, vc2, o_again, hd, group, nkv, B, scale, nq * hd, hd, cap2 * hd, nq * hd, small_rows=1, **kw)
        assert torch.equal(o_small, o_again)


def test_attention_causal():
    """Text prefill of the chat VLM (HF causal mask): GQA 6:1 and 7:1, ragged lengths incl. several query blocks, and the
    decode shape (one query row over kv_len keys)."""
    for group, lens in ((6, [85, 1, 130, 300]), (7, [64, 129])):
        B, nkv, hd = len(lens), 2, 128
        nq = nkv * group
        cu = [0]
        for L in lens:
            cu.append(cu[-1] + L)
        window = max(lens) + 8
        kvoff = [b * window for b in range(B)]
        q = _rand(cu[-1], nq * hd, seed=11)
        kc = _rand(nkv, B * window, hd, seed=12)
        vc = _rand(nkv, B * window, hd, seed=13)
        i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=DEV)
        out = torch.zeros_like(q)
        scale = hd ** -0.5
        ops.attention(q, kc, vc, out, hd, group, nkv, B, scale, nq * hd, hd, B * window * hd, nq * hd, max_q_len=max(lens),
                      q_start=i32(cu[:-1]), q_len=i32(lens), kv_start=i32(kvoff), kv_len=i32(lens), causal=1)
        errs = []
        for b in range(B):
            L = lens[b]
            mask = torch.full((L, L), float("-inf"), device=DEV).triu(1)
            for h in range(nq):
                g = h // group
                K = kc[g, kvoff[b]:kvoff[b] + L].float()
                V = vc[g, kvoff[b]:kvoff[b] + L].float()
                Q = q[cu[b]:cu[b] + L, h * hd:(h + 1) * hd].float()
                r = torch.softmax(Q @ K.T * scale + mask, -1) @ V
                errs.append(ops.rel_err(out[cu[b]:cu[b] + L, h * hd:(h + 1) * hd], r))
        print(f"attention_causal group={group} max rel={max(errs):.2e}")
        assert max(errs) < 1e-2
        # decode: one new query per sample sees its kv_len keys (no mask needed, causal flag harmless)
        qd = _rand(B, nq * hd, seed=14)
        outd = torch.zeros_like(qd)
        ops.attention(qd, kc, vc, outd, hd, group, nkv, B, scale, nq * hd, hd, B * window * hd, nq * hd, max_q_len=1,
                      q_start=i32(list(range(B))), q_len=i32([1] * B), kv_start=i32(kvoff), kv_len=i32(lens), causal=1)
        errs = []
        for b in range(B):
            for h in range(nq):
                g = h // group
                r = _attn_ref(qd[b:b + 1, h * hd:(h + 1) * hd].float(), kc[g, kvoff[b]:kvoff[b] + lens[b]].float(),
                              vc[g, kvoff[b]:kvoff[b] + lens[b]].float(), scale, lens[b])
                errs.append(ops.rel_err(outd[b:b + 1, h * hd:(h + 1) * hd], r))
        assert max(errs) < 1e-2


@pytest.mark.parametrize("nkv,group,lens,chunk", [(2, 6, [85, 1, 130, 300, 3353, 64, 257], 256), (4, 7, [3353, 277, 1302, 2000], 256),
                                                  (1, 1, [100, 63], 64), (2, 8, [129, 128, 1000], 128)])
def test_decode_attention_split_kv(nkv, group, lens, chunk):
    """Split-KV decode-step attention: one query row per sample over windows of 1 .. 3353 keys (1 .. 14 chunks), GQA groups
    1 / 6 / 7 / 8, against the softmax over the whole window; call-to-call bit-identical (fixed merge order)."""
    hd, nq, B = 128, nkv * group, len(lens)
    window = max(lens) + 5
    q = _rand(B, nq * hd, seed=21)
    kc = _rand(nkv, B * window, hd, seed=22)
    vc = _rand(nkv, B * window, hd, seed=23)
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=DEV)
    kvoff = [b * window for b in range(B)]
    scale = hd ** -0.5
    out = ops.decode_attention(q, kc, vc, i32(kvoff), i32(lens), nq, nkv, scale, chunk)
    errs = []
    for b in range(B):
        for h in range(nq):
            g = h // group
            r = _attn_ref(q[b:b + 1, h * hd:(h + 1) * hd].float(), kc[g, kvoff[b]:kvoff[b] + lens[b]].float(),
                          vc[g, kvoff[b]:kvoff[b] + lens[b]].float(), scale, lens[b])
            errs.append(ops.rel_err(out[b:b + 1, h * hd:(h + 1) * hd], r))
    print(f"decode_attention nkv={nkv} group={group} chunk={chunk} max rel={max(errs):.2e}")
    assert max(errs) < 1e-2
    assert torch.equal(out, ops.decode_attention(q, kc, vc, i32(kvoff), i32(lens), nq, nkv, scale, chunk))


@pytest.mark.parametrize("width", [128, 768, 1024, 1536, 4096])
def test_norms(width):
    rows = 515
    x = _rand(rows, width, dtype=torch.float32, seed=1) * 3 + 0.5
    gamma = _rand(width, dtype=torch.float32, seed=2)
    beta = _rand(width, dtype=torch.float32, seed=3)
    ln = ops.layernorm(x, gamma, beta, 1e-6)
    ref = torch.nn.functional.layer_norm(x, (width,), gamma, beta, 1e-6)
    assert ops.rel_err(ln, ref) < 4e-3
    rms = ops.rmsnorm(x, gamma, 1e-6)
    ref = gamma * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6))
    assert ops.rel_err(rms, ref) < 4e-3


def test_pixel_shuffle_bit_exact():
    """Integer/indexing work: bit-exact against the closed-form index map and the reference op."""
    from oracle.vlaser_oracle import pixel_shuffle as ref_ps
    for g, Cc, dtype in ((16, 1024, torch.float32), (32, 128, torch.float32), (16, 256, torch.bfloat16)):
        n = 3
        x = torch.arange(n * g * g * Cc, dtype=torch.float32).remainder(65521).to(dtype).view(n, g, g, Cc)
        got = ops.pixel_shuffle(x.to(DEV)).cpu()
        assert torch.equal(got, ref_ps(x))


def test_pixel_shuffle_layernorm_gather():
    from oracle.vlaser_oracle import pixel_shuffle as ref_ps
    n, g, Cc = 2, 16, 256
    x = _rand(n * (g * g + 1), Cc, dtype=torch.float32, seed=1)
    gamma = _rand(4 * Cc, dtype=torch.float32, seed=2)
    beta = _rand(4 * Cc, dtype=torch.float32, seed=3)
    got = ops.layernorm(x, gamma, beta, 1e-5, g, Cc, rows=n * g * g // 4, width=4 * Cc)
    feats = x.view(n, g * g + 1, Cc)[:, 1:].reshape(n, g, g, Cc).cpu()
    sh = ref_ps(feats).reshape(-1, 4 * Cc).to(DEV)
    ref = torch.nn.functional.layer_norm(sh, (4 * Cc,), gamma, beta, 1e-5)
    assert ops.rel_err(got, ref) < 4e-3


def test_im2col_exact():
    n, px, ps = 2, 224, 14
    pix = _rand(n, 3, px, px, dtype=torch.bfloat16, seed=1)
    cols = ops.im2col(pix, ps, 640)
    ref = torch.nn.functional.unfold(pix.float(), ps, stride=ps).transpose(1, 2).reshape(n * 256, 588)
    assert torch.equal(cols[:, :588].float(), ref)
    assert torch.all(cols[:, 588:] == 0)


def test_euler_step():
    B, H, A, width, horizon = 5, 4, 7, 768, 4
    x = _rand(B * H, width, dtype=torch.float32, seed=1)
    gamma = _rand(width, dtype=torch.float32, seed=2)
    Wd = _rand(A, width, dtype=torch.float32, seed=3) * 0.1
    bd = _rand(A, dtype=torch.float32, seed=4)
    act = _rand(B * H, A, dtype=torch.float32, seed=5)
    ref = act + 0.1 * ((gamma * x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)) @ Wd.t() + bd)
    out = torch.zeros(B, horizon, A, dtype=torch.float32, device=DEV)
    ops.euler_step(x, gamma, Wd, bd, act, out, H, horizon, 0.1, 1e-6, True, 1.0)
    assert ops.max_err(act, ref) < 1e-5
    assert ops.max_err(out.view(-1, A), ref.clamp(-1, 1)) < 1e-5
