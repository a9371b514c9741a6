"""BASELINE configs[4] on ONE GPU: Vlaser-8B widths (InternViT-300M + Qwen2.5-7B dims, expert MLP
5600) prefix encode + bidirectional block prefill only (`stop_after=3`: ViT -> pixel_shuffle + mlp1 -> token merge ->
28-layer prefill writing the KV cache), 448 px dynamic tiles: the 4 samples of a GPU's shard hold 1 / 5 / 7 / 11 tiles
(24 = B * cond_steps, 256 image tokens each). Exercised as the 8B VLA (bidirectional prefix), not the causal chat VLM.
Random-init weights generated on the device; prints one JSON line (time per call, tiles/s, prefix tokens/s, TFLOP/s of
the ViT + projector + prefill FLOPs of SURVEY.md §8d with valid-token lengths).

    python tools/bench_cfg5.py --causal   # the same widths as the chat VLM (eval_example.py:38-82): tiles 1 / 5 / 13 / 13
                                          # (12 tiles + thumbnail), causal prefill + greedy decode through vlaser_generate"""
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

CAUSAL = "--causal" in sys.argv
TILES = [1, 5, 13, 13] if CAUSAL else [1, 5, 7, 11]
B = len(TILES)
cfg = vlaser_8b_vla().with_(cond_steps=sum(TILES) // B, max_image_text_tokens=3456 if CAUSAL else 2880)
dev = torch.device("cuda", 0)
pol = PiZeroInference(cfg, max_batch=B)
for k, t in _device_state_dict_stream(cfg, 0, dev):
    pol.load_state_dict({k: t}, strict=False)
if CAUSAL:
    # chat VLM: causal prefill over the merged prompt + greedy decode (untied output head, as Qwen2.5-7B)
    g = torch.Generator(device=dev).manual_seed(7)
    pol.load_state_dict({"joint_model.mixtures.vlm.norm.weight": torch.ones(cfg.hidden, device=dev)}, strict=False)
    pol.load_state_dict({"lm_head.weight": (torch.randn((cfg.embed_rows, cfg.hidden), device=dev, generator=g) * 0.02).bfloat16()},
                        strict=False)
    inp = make_inputs(cfg, B, 5, tiles_per_sample=TILES)
    ids, pix = inp["input_ids"].to(dev), inp["pixel_values"].to(torch.bfloat16).to(dev)

    def timed(n_new, reps):
        for _ in range(2):
            pol.generate(ids, pix, max_new_tokens=n_new)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            pol.generate(ids, pix, max_new_tokens=n_new)
            a1.record()
            torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
        return statistics.median(ts)

    t1, t65 = timed(1, 5), timed(65, 3)
    N, dv, fv, P = cfg.vit_tokens, cfg.vit_hidden, cfg.vit_mlp, cfg.patches
    f_vit = cfg.vit_layers * (2 * N * dv * 3 * dv + 2 * N * dv * dv + 4 * N * dv * fv + 4 * N * N * dv) + 2 * P * 588 * dv
    f_proj = cfg.image_tokens * (2 * 4 * dv * cfg.hidden + 2 * cfg.hidden ** 2)
    Wl = cfg.hidden * cfg.q_dim + 2 * cfg.hidden * cfg.kv_dim + cfg.q_dim * cfg.hidden + 3 * cfg.hidden * cfg.mlp
    flops = sum(TILES) * (f_vit + f_proj)
    for sx in inp["valid_len"].tolist():  # causal: half of the score matrix
        flops += cfg.layers * (2 * sx * Wl + 2 * sx * sx * cfg.q_dim) + 2 * cfg.hidden * cfg.embed_rows
    dec_ms = (t65 - t1) / 64
    wbytes = 2.0 * (cfg.layers * Wl + cfg.hidden * cfg.embed_rows)
    print(json.dumps({
        "workload": "Vlaser-8B widths as the chat VLM (InternVLChatModel.generate, greedy): 448 px dynamic tiles %s per sample, "
                    "B=%d on one B200, causal prefill + decode, random-init bf16" % (TILES, B),
        "valid_len": inp["valid_len"].tolist(), "prefill_plus_first_token_ms": t1, "tiles_per_s": sum(TILES) / (t1 * 1e-3),
        "prompt_tokens_per_s": float(inp["valid_len"].sum()) / (t1 * 1e-3), "algorithmic_tflop_prefill": flops / 1e12,
        "prefill_tflops": flops / (t1 * 1e-3) / 1e12, "decode_ms_per_token": dec_ms, "decode_tokens_per_s": B / (dec_ms * 1e-3),
        "decode_weight_stream_gbs": wbytes / (dec_ms * 1e-3) / 1e9, "decode_weight_bytes_gb": wbytes / 1e9,
    }))
    sys.exit(0)
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
