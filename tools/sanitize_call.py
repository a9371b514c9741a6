"""Small calls for compute-sanitizer (memcheck / racecheck): tiny config at B = 2 (per-kernel schedule), B = 1
(weight-streaming expert kernel) and the dynamic-tile layout, direct launches (no CUDA graph)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vlaser_b200.config import tiny
from vlaser_b200.policy import PiZeroInference
from vlaser_b200.synth import make_inputs
from vlaser_b200.weights import random_state_dict

for name, cfg, B, tiles in (("b2", tiny(), 2, None), ("b1", tiny(), 1, None),
                            ("tiles", tiny().with_(cond_steps=2, max_image_text_tokens=224), 4, [1, 3, 2, 2])):
    sd = random_state_dict(cfg, 0)
    inp = make_inputs(cfg, B, 7, tiles_per_sample=tiles)
    pol = PiZeroInference(cfg, max_batch=B)
    pol.load_state_dict(sd, strict=True)
    pol.set_option("use_graph", 0)
    out = pol.infer_action(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), proprios=inp["proprios"].cuda(),
                           noise=inp["noise"].cuda()).float().cpu()
    print(name, tuple(out.shape), float(out.abs().max()), "launches", pol.last_launch_count(), flush=True)
    del pol
