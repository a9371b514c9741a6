"""One profiled chat-VLM generate call (for ncu --profile-from-start off): Vlaser-8B widths, 448 px tiles 1/5/13/13, B = 4,
max_new_tokens = 2 (causal prefill + ONE decode step, direct launches so that every kernel is listed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import _device_state_dict_stream
from vlaser_b200.config import vlaser_8b_vla
from vlaser_b200.policy import PiZeroInference
from vlaser_b200.synth import make_inputs

TILES = [1, 5, 13, 13]
B = len(TILES)
cfg = vlaser_8b_vla().with_(cond_steps=sum(TILES) // B, max_image_text_tokens=3456)
dev = torch.device("cuda", 0)
pol = PiZeroInference(cfg, max_batch=B)
for k, t in _device_state_dict_stream(cfg, 0, dev):
    pol.load_state_dict({k: t}, strict=False)
g = torch.Generator(device=dev).manual_seed(7)
pol.load_state_dict({"joint_model.mixtures.vlm.norm.weight": torch.ones(cfg.hidden, device=dev)}, strict=False)
pol.load_state_dict({"lm_head.weight": (torch.randn((cfg.embed_rows, cfg.hidden), device=dev, generator=g) * 0.02).bfloat16()}, strict=False)
pol.set_option("use_graph", 0)
inp = make_inputs(cfg, B, 5, tiles_per_sample=TILES)
ids, pix = inp["input_ids"].to(dev), inp["pixel_values"].to(torch.bfloat16).to(dev)
pol.generate(ids, pix, max_new_tokens=2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
pol.generate(ids, pix, max_new_tokens=2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches", pol.last_launch_count())
