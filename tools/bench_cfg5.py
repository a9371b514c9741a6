"""BASELINE configs[4] on ONE GPU: Vlaser-8B widths (InternViT-300M + Qwen2.5-7B dims, expert MLP
5600) prefix encode + bidirectional block prefill only (`stop_after=3`: ViT -> pixel_shuffle + mlp1 -> token merge ->
28-layer prefill writing the KV cache), 448 px dynamic tiles: the 4 samples of a GPU's shard hold 1 / 5 / 7 / 11 tiles
(24 = B * cond_steps, 256 image tokens each). Exercised as the 8B VLA (bidirectional prefix), not the causal chat VLM.
Random-init weights generated on the device; prints one JSON line (time per call, tiles/s, prefix tokens/s, TFLOP/s of
the ViT + projector + prefill FLOPs of SURVEY.md §8d with valid-token lengths)."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import _device_state_dict_stream
from vlaser_b200.config import vlaser_8b_vla
from vlaser_b200.policy import PiZeroInference
from vlaser_b200.synth import make_inputs

TILES = [1, 5, 7, 11]
B = len(TILES)
cfg = vlaser_8b_vla().with_(cond_steps=sum(TILES) // B, max_image_text_tokens=2880)
dev = torch.device("cuda", 0)
pol = PiZeroInference(cfg, max_batch=B)
for k, t in _device_state_dict_stream(cfg, 0, dev):
    pol.load_state_dict({k: t}, strict=False)
pol.set_option("stop_after", 3)
inp = make_inputs(cfg, B, 5, tiles_per_sample=TILES)
ids, pix = inp["input_ids"].to(dev), inp["pixel_values"].to(torch.bfloat16).to(dev)
prop, noise = inp["proprios"].to(torch.bfloat16).to(dev), inp["noise"].to(torch.bfloat16).to(dev)
for _ in range(3):
    pol.infer_action(ids, pix, proprios=prop, noise=noise)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    pol.infer_action(ids, pix, proprios=prop, noise=noise)
    a1.record()
    torch.cuda.synchronize()
    ts.append(a0.elapsed_time(a1))
ms = statistics.median(ts)
# algorithmic FLOPs: ViT + projector per tile, prefill with valid lengths (bidirectional block prefix)
N, dv, fv, P = cfg.vit_tokens, cfg.vit_hidden, cfg.vit_mlp, cfg.patches
f_vit = cfg.vit_layers * (2 * N * dv * 3 * dv + 2 * N * dv * dv + 4 * N * dv * fv + 4 * N * N * dv) + 2 * P * 588 * dv
f_proj = cfg.image_tokens * (2 * 4 * dv * cfg.hidden + 2 * cfg.hidden ** 2)
W = lambda h, f: h * cfg.q_dim + 2 * h * cfg.kv_dim + cfg.q_dim * h + 3 * h * f
flops = sum(TILES) * (f_vit + f_proj)
for sx in inp["valid_len"].tolist():
    flops += cfg.layers * (2 * sx * W(cfg.hidden, cfg.mlp) + 2 * W(cfg.hidden_e, cfg.mlp_e) + 4 * (sx * sx + sx + 1) * cfg.q_dim) \
        - 2 * sx * (cfg.q_dim * cfg.hidden + 3 * cfg.hidden * cfg.mlp) - 2 * (cfg.q_dim * cfg.hidden_e + 3 * cfg.hidden_e * cfg.mlp_e)
print(json.dumps({
    "workload": "Vlaser-8B widths, prefix encode + prefill (stop_after=3), 448 px dynamic tiles %s per sample, B=%d on one B200 "
                "(the per-GPU shard of bs=32 on 8 GPUs), random-init bf16" % (TILES, B),
    "ms_per_call_p50": ms, "ms_min": min(ts), "tiles_per_s": sum(TILES) / (ms * 1e-3),
    "prefix_tokens_per_s": float(inp["valid_len"].sum()) / (ms * 1e-3), "valid_len": inp["valid_len"].tolist(),
    "algorithmic_tflop_per_call": flops / 1e12, "tflops": flops / (ms * 1e-3) / 1e12, "launches": pol.last_launch_count(),
}))
