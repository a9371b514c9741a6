"""Per-kind cost of the launches of the ViT and prefill stages of one infer_action call (main stream), from the in-situ
stamps of option "trace" (see tools/timeline.py): mean distance from a launch's post-wait stamp to the next main-stream
launch's post-wait stamp = what each launch kind costs on the critical path of the graph replay.
    python tools/timeline_stages.py [--batch 64] [--opts k=v,...] """
import argparse
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import _device_state_dict_stream
from vlaser_b200.config import PRESETS
from vlaser_b200.policy import PiZeroInference
from vlaser_b200.synth import make_inputs

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--config", default="vlaser_2b_bridge")
ap.add_argument("--opts", default="")
ap.add_argument("--layer", type=int, default=10)
a = ap.parse_args()
cfg = PRESETS[a.config]()
dev = torch.device("cuda", 0)
pol = PiZeroInference(cfg, max_batch=a.batch)
for k, t in _device_state_dict_stream(cfg, 0, dev):
    pol.load_state_dict({k: t}, strict=False)
for kv in filter(None, a.opts.split(",")):
    k, v = kv.split("=")
    pol.set_option(k, int(v))
pol.set_option("trace", 1)
inp = make_inputs(cfg, a.batch, 1)
ids, pix = inp["input_ids"].to(dev), inp["pixel_values"].to(torch.bfloat16).to(dev)
prop, noise = inp["proprios"].to(torch.bfloat16).to(dev), inp["noise"].to(torch.bfloat16).to(dev)
for _ in range(4):
    pol.infer_action(ids, pix, proprios=prop, noise=noise)
torch.cuda.synchronize()
names = bytes(pol.debug_buffer("trace_names", torch.uint8).tolist()).decode().strip().split("\n")
tr = pol.debug_buffer("trace", torch.int64).view(-1, 16)[: len(names)].double()
print(f"config {a.config} B={a.batch} opts [{a.opts}]: {len(names)} traced launches")
recs = []
for i, nm in enumerate(names):
    if nm.startswith("act_enc2"):
        break
    side = nm.endswith("[side]")
    base = nm.replace(" [side]", "")
    kind, _, lay = base.partition(" L")
    recs.append(dict(kind=kind, layer=int(lay) if lay else -1, side=side, t=tr[i]))
SLOTS = ["entry", "wait", "1st-stage", "mma-issued", "acc-ready", "epi-done", "exit"]
for title, sel in (("ViT", [r for r in recs if r["kind"].startswith("vit") and r["layer"] in (a.layer, a.layer + 1)]),
                   ("prefill", [r for r in recs if not r["kind"].startswith("vit") and r["layer"] in (a.layer, a.layer + 1)])):
    if not sel:
        continue
    t0 = min(float(r["t"][0]) for r in sel if r["t"][0] > 0)
    print(f"---- {title} layers {a.layer}..{a.layer + 1} (us relative to the first entry)")
    print(f"{'launch':26s} " + " ".join(f"{s:>10s}" for s in SLOTS))
    for r in sel:
        vals = " ".join(f"{(float(v) - t0) / 1e3:10.2f}" if v > 0 else f"{'-':>10s}" for v in r["t"][:7])
        fine = ""
        if r["kind"] == "attention_tc":  # kernel-specific stamps 8..15: first two tiles of softmax group 0
            fine = "   tile0 p/o/read/stored, tile2 ...: " + " ".join(f"{(float(v) - t0) / 1e3:.2f}" for v in r["t"][8:16] if v > 0)
        print(f"{(r['kind'] + ' L' + str(r['layer']) + (' [side]' if r['side'] else '')):26s} {vals}{fine}")
main = [r for r in recs if not r["side"] and r["t"][1] > 0]
cost = defaultdict(list)
for r, nx in zip(main[:-1], main[1:]):
    cost[r["kind"]].append((float(nx["t"][1]) - float(r["t"][1])) / 1e3)
print("---- ViT + prefill, main stream: mean distance from a launch's post-wait stamp to the next launch's (us)")
tot = 0.0
for k, v in cost.items():
    m = sum(v) / len(v)
    tot += m * len(v)
    print(f"{k:16s} x{len(v):5d}  mean {m:8.2f}  min {min(v):8.2f}  max {max(v):8.2f}  total {m * len(v) / 1e3:7.3f} ms")
print(f"sum: {tot / 1e3:.3f} ms")
