"""Micro-benchmark of the InternViT attention kernels through the single-op C ABI entry points: the tcgen05 kernels
(257 tokens: single pass; 1025 tokens: two-pass, K resident) against the generic mma.sync flash kernel on the same packed
q|k|v buffer.    python tools/vit_attn_bench.py [images]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests import ops

heads, hd = 16, 64
d = heads * hd


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


for N, images in ((257, 64), (1025, int(sys.argv[1]) if len(sys.argv) > 1 else 24), (1025, 64)):
    g = torch.Generator(device="cuda").manual_seed(N)
    qkv = torch.randn((images * N, 3 * d), device="cuda", generator=g).bfloat16()
    flops = 4.0 * N * N * hd * heads * images
    out_g = torch.zeros(images * N, d, dtype=torch.bfloat16, device="cuda")

    def generic():
        ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], out_g, hd, 1, heads, images, hd ** -0.5, 3 * d, 3 * d, hd, d,
                      uni_q_len=N, uni_kv_len=N, max_q_len=N)

    lib = ops._lib.load()
    out_t = torch.zeros(images * N, d, dtype=torch.bfloat16, device="cuda")

    def tc():
        ops._lib.check(lib.vlaser_op_vit_attention_tc(ops._p(qkv), ops._p(out_t), images, N, heads, hd ** -0.5, ops._stream()), "tc")

    t_tc = timed(tc)
    t_g = timed(generic)
    out_tc = ops.vit_attention(qkv, images, N, heads, hd ** -0.5, tc=True)
    generic()
    print(f"N={N} images={images}: tcgen05 {t_tc:8.1f} us = {flops / t_tc / 1e6:7.1f} TFLOP/s | generic mma.sync {t_g:8.1f} us = "
          f"{flops / t_g / 1e6:7.1f} TFLOP/s | rel diff {ops.rel_err(out_tc, out_g.float()):.2e}")
