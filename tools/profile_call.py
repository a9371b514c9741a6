"""One profiled policy call (for ncu --profile-from-start off): builds Vlaser-2B-VLA with device-generated
random weights, warms up, then brackets exactly one call with cudaProfilerStart/Stop."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import _device_state_dict_stream
from vlaser_b200.config import PRESETS
from vlaser_b200.policy import PiZeroInference
from vlaser_b200.synth import make_inputs

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--graph", type=int, default=0)
ap.add_argument("--config", default="vlaser_2b_bridge")
ap.add_argument("--stop-after", type=int, default=0)
ap.add_argument("--px", type=int, default=0, help="image size override (448: the 1025-token ViT rows)")
ap.add_argument("--opt", action="append", default=[], help="schedule option k=v (repeatable)")
a = ap.parse_args()
cfg = PRESETS[a.config]()
if a.px:
    cfg = cfg.with_(image_size=a.px)
dev = torch.device("cuda", 0)
pol = PiZeroInference(cfg, max_batch=a.batch)
for k, t in _device_state_dict_stream(cfg, 0, dev):
    pol.load_state_dict({k: t}, strict=False)
pol.set_option("use_graph", a.graph)
pol.set_option("stop_after", a.stop_after)
for kv in a.opt:
    pol.set_option(kv.split("=")[0], int(kv.split("=")[1]))
inp = make_inputs(cfg, a.batch, 1)
ids, pix = inp["input_ids"].to(dev), inp["pixel_values"].to(torch.bfloat16).to(dev)
prop, noise = inp["proprios"].to(torch.bfloat16).to(dev), inp["noise"].to(torch.bfloat16).to(dev)
for _ in range(2):
    pol.infer_action(ids, pix, proprios=prop, noise=noise)
torch.cuda.synchronize()
torch.cuda.profiler.start()
pol.infer_action(ids, pix, proprios=prop, noise=noise)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches", pol.last_launch_count())
