"""Per-shape micro-benchmark of the tcgen05 GEMM at the shapes of one Vlaser-2B-VLA call (B=64):
CUDA events, L2 flushed between iterations, prints time, TFLOP/s and the share of the call."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests import ops
from tests.ops import (EPI_GELU_BF16, EPI_QKV_ROPE, EPI_RESID_F32, EPI_SILU_BF16, EPI_STORE_BF16, EPI_STORE_F32,
                       EPI_SWIGLU_BF16)

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--bn", type=int, default=0, help="force block_n (0 = engine heuristic)")
ap.add_argument("--epi", type=int, default=-1, help="override the epilogue of bf16-output shapes (0 store, 1 GELU, 2 SiLU): what an epilogue costs")
ap.add_argument("--pair", type=int, default=1, help="0 single-CTA kernels, 1 CTA-pair kernel + TMA epilogue, 2 pair w/o TMA epilogue")
a = ap.parse_args()
B = a.batch
MV, MP, MA = B * 257, B * 87, B * 4
SHAPES = [  # name, M, N, K, epi, count per call, max_splits
    ("vit_qkv", MV, 3072, 1024, EPI_STORE_BF16, 24, 1), ("vit_proj", MV, 1024, 1024, EPI_RESID_F32, 24, 1),
    ("vit_fc1", MV, 4096, 1024, EPI_GELU_BF16, 24, 1), ("vit_fc2", MV, 1024, 4096, EPI_RESID_F32, 24, 1),
    ("llm_qkv", MP, 2048, 1536, EPI_QKV_ROPE, 28, 1), ("llm_o", MP, 1536, 1536, EPI_RESID_F32, 27, 1),
    ("llm_gu", MP, 17920, 1536, EPI_SWIGLU_BF16, 27, 1), ("llm_down", MP, 1536, 8960, EPI_RESID_F32, 27, 1),
    ("pro_qkv", B, 2048, 768, EPI_QKV_ROPE, 28, 1), ("pro_o", B, 768, 1536, EPI_RESID_F32, 27, 32),
    ("pro_gu", B, 17920, 768, EPI_SWIGLU_BF16, 27, 1), ("pro_down", B, 768, 8960, EPI_RESID_F32, 27, 32),
    ("act_qkv", MA, 2048, 768, EPI_QKV_ROPE, 280, 1), ("act_o", MA, 768, 1536, EPI_RESID_F32, 280, 32),
    ("act_gu", MA, 17920, 768, EPI_SWIGLU_BF16, 280, 1), ("act_down", MA, 768, 8960, EPI_RESID_F32, 280, 32),
    ("act_enc", MA, 768, 768, EPI_SILU_BF16, 20, 1),
    # the same large shapes behind the plain bf16 store: what the fused epilogue costs (not part of a call: count 0)
    ("llm_gu_st", MP, 17920, 1536, EPI_STORE_BF16, 0, 1), ("llm_qkv_st", MP, 2048, 1536, EPI_STORE_BF16, 0, 1),
    ("fc2_st", MV, 1024, 4096, EPI_STORE_BF16, 0, 1), ("llm_down_st", MP, 1536, 8960, EPI_STORE_BF16, 0, 1),
]
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
cos, sin = ops.rope_tables(512)
total = 0.0
rows = []
for name, M, N, K, epi, cnt, splits in SHAPES:
    if a.only and a.only != name:
        continue
    if a.epi >= 0 and epi in (EPI_STORE_BF16, EPI_GELU_BF16, EPI_SILU_BF16):
        epi = a.epi
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    kw = {}
    bn = 128
    if epi != EPI_QKV_ROPE and N % 256 == 0 and ((M + 127) // 128) * (N // 256) >= 148:
        bn = 256
    if a.bn and epi != EPI_QKV_ROPE and N % a.bn == 0:
        bn = a.bn
    if a.pair and not a.bn and N % 256 == 0 and splits <= 1 and ((M + 255) // 256) * (N // 256) >= 74:
        bn = {1: 513, 2: 512, 3: 515}[a.pair]
    if epi in (EPI_STORE_BF16, EPI_GELU_BF16, EPI_SILU_BF16):
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    elif epi == EPI_SWIGLU_BF16:
        out = torch.empty(M, N // 2, dtype=torch.bfloat16, device=dev)
    elif epi == EPI_RESID_F32:
        out = torch.zeros(M, N, dtype=torch.float32, device=dev)
        kw["max_splits"] = splits
    else:
        out = torch.empty(M, 1536, dtype=torch.bfloat16, device=dev)
        kw.update(row_pos=torch.ones(M, dtype=torch.int32, device=dev),
                  row_kvrow=torch.arange(M, dtype=torch.int32, device=dev), cos_tab=cos, sin_tab=sin,
                  k_cache=torch.empty(2, M, 128, dtype=torch.bfloat16, device=dev),
                  v_cache=torch.empty(2, M, 128, dtype=torch.bfloat16, device=dev), kv_head_stride=M * 128,
                  n_q_heads=12, n_kv_heads=2)
    ts = []
    for i in range(a.iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(A, W, epi, bn, False, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    tf = 2.0 * M * N * K / (t * 1e-3) / 1e12
    gbs = (M * K * 2 + N * K * 2) / (t * 1e-3) / 1e9
    total += t * cnt
    rows.append((name, M, N, K, bn, t, tf, gbs, cnt, t * cnt))
for r in rows:
    print("%-11s M=%6d N=%6d K=%5d bn=%3d  %8.3f ms  %7.1f TF/s  %7.0f GB/s(in)  x%3d = %7.2f ms (%4.1f%%)" % (*r, 100 * r[-1] / max(total, 1e-9)))
print("total GEMM ms per call (cold L2): %.2f" % total)
