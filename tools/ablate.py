"""In-situ cost of each launch kind of the denoise steps (B = 64 by default): the whole call is replayed from its CUDA
graph with one kind of launch left out (option "ablate"; the actions are then garbage, only the time is read), so the
difference to the full call is what that kind contributes to the dependent chain -- unlike an ncu launch list, which
serialises the launches with cold caches. Also prints the stage split through `stop_after`.
    python tools/ablate.py [--batch 64] [--opts "pair_min_tiles=70"] """
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
ap.add_argument("--config", default="vlaser_2b_bridge")
ap.add_argument("--opts", default="", help="comma list key=value applied first")
ap.add_argument("--masks", default="0,1,2,4,8,16,32,64,17,127")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
cfg = PRESETS[a.config]()
dev = torch.device("cuda", 0)
pol = PiZeroInference(cfg, max_batch=a.batch)
for k, t in _device_state_dict_stream(cfg, 0, dev):
    pol.load_state_dict({k: t}, strict=False)
for kv in filter(None, a.opts.split(",")):
    k, v = kv.split("=")
    pol.set_option(k, int(v))
inp = make_inputs(cfg, a.batch, 1)
ids, pix = inp["input_ids"].to(dev), inp["pixel_values"].to(torch.bfloat16).to(dev)
prop, noise = inp["proprios"].to(torch.bfloat16).to(dev), inp["noise"].to(torch.bfloat16).to(dev)


def timed():
    for _ in range(3):
        pol.infer_action(ids, pix, proprios=prop, noise=noise)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pol.infer_action(ids, pix, proprios=prop, noise=noise)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


print(f"config {a.config} B={a.batch} opts [{a.opts}]")
prev = 0.0
for stop, name in ((1, "ViT"), (2, "+projector/embed"), (3, "+prefill"), (0, "whole call")):
    pol.set_option("stop_after", stop)
    t = timed()
    print(f"stop_after={stop} {name:18s} {t:8.3f} ms  (+{t - prev:7.3f})", flush=True)
    prev = t
full = prev
names = {1: "norm_in", 2: "qkv", 4: "attention", 8: "o", 16: "norm_post", 32: "gate_up", 64: "down"}
n_layer_runs = cfg.layers * cfg.num_inference_steps
for m in [int(x) for x in a.masks.split(",")]:
    pol.set_option("ablate", m)
    t = timed()
    label = "+".join(v for k, v in names.items() if m & k) or "nothing"
    print(f"ablate={m:3d} without {label:50s} {t:8.3f} ms  saves {full - t:7.3f} ms = {1e3 * (full - t) / n_layer_runs:6.2f} us per layer",
          flush=True)
pol.set_option("ablate", 0)
