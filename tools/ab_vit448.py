"""A/B of the 448 px ViT encode (stop_after=1) with the tcgen05 two-pass attention kernel (vit_tc=1) and the generic
mma.sync kernel (vit_tc=0) on one policy object: Vlaser-2B widths, 448 px tiles.   python tools/ab_vit448.py [tiles]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import _device_state_dict_stream
from vlaser_b200.config import vlaser_2b_bridge
from vlaser_b200.policy import PiZeroInference
from vlaser_b200.synth import make_inputs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cfg = vlaser_2b_bridge().with_(image_size=448, max_image_text_tokens=384)
dev = torch.device("cuda", 0)
pol = PiZeroInference(cfg, max_batch=B)
for k, t in _device_state_dict_stream(cfg, 0, dev):
    pol.load_state_dict({k: t}, strict=False)
inp = make_inputs(cfg, B, 1)
ids, pix = inp["input_ids"].to(dev), inp["pixel_values"].to(torch.bfloat16).to(dev)
prop, noise = inp["proprios"].to(torch.bfloat16).to(dev), inp["noise"].to(torch.bfloat16).to(dev)


def timed(iters=10):
    for _ in range(3):
        pol.infer_action(ids, pix, proprios=prop, noise=noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        pol.infer_action(ids, pix, proprios=prop, noise=noise)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for stop in (1, 0):
    pol.set_option("stop_after", stop)
    for rep in range(2):
        for tc in (1, 0):
            pol.set_option("vit_tc", tc)
            print(f"448 px, {B} tiles, stop_after={stop} vit_tc={tc}: {timed():8.3f} ms", flush=True)
