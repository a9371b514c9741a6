"""Timeline of the two sub-batch denoise chains (option den_chains=2): every traced launch of both streams inside a
window of one denoise step, sorted by entry time (us), to see whether the chains overlap or queue behind each other.
    python tools/timeline_chains.py [--opts k=v,...] [--layers 10,11] [--occurrence 5]"""
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
ap.add_argument("--opts", default="den_chains=2")
ap.add_argument("--layers", default="10,11")
ap.add_argument("--occurrence", type=int, default=5, help="which denoise step (per stream) to show")
a = ap.parse_args()
cfg = PRESETS["vlaser_2b_bridge"]()
dev = torch.device("cuda", 0)
pol = PiZeroInference(cfg, max_batch=64)
for k, t in _device_state_dict_stream(cfg, 0, dev):
    pol.load_state_dict({k: t}, strict=False)
for kv in filter(None, a.opts.split(",")):
    k, v = kv.split("=")
    pol.set_option(k, int(v))
pol.set_option("trace", 1)
inp = make_inputs(cfg, 64, 1)
ids, pix = inp["input_ids"].to(dev), inp["pixel_values"].to(torch.bfloat16).to(dev)
prop, noise = inp["proprios"].to(torch.bfloat16).to(dev), inp["noise"].to(torch.bfloat16).to(dev)
for _ in range(4):
    pol.infer_action(ids, pix, proprios=prop, noise=noise)
torch.cuda.synchronize()
names = bytes(pol.debug_buffer("trace_names", torch.uint8).tolist()).decode().strip().split("\n")
tr = pol.debug_buffer("trace", torch.int64).view(-1, 16)[: len(names)].double()
layers = [int(x) for x in a.layers.split(",")]
step = {False: -1, True: -1}
sel = []
for i, nm in enumerate(names):
    side = nm.endswith("[side]")
    base = nm.replace(" [side]", "")
    if base.startswith("act_enc2"):
        step[side] += 1
    kind, _, lay = base.partition(" L")
    if step[side] == a.occurrence and lay and int(lay) in layers:
        sel.append((float(tr[i][0]), float(tr[i][1]), float(tr[i][6]), ("B " if side else "A ") + base))
sel.sort()
t0 = sel[0][0]
print(f"opts [{a.opts}]: denoise step {a.occurrence}, layers {layers}: entry / post-wait / exit (us)")
for e, w, x, nm in sel:
    print(f"{nm:22s} {(e - t0) / 1e3:9.2f} {(w - t0) / 1e3:9.2f} {(x - t0) / 1e3:9.2f}")
print(f"window {(max(s[2] for s in sel) - t0) / 1e3:.2f} us for {len(sel)} launches")
