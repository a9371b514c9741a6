"""In-situ timeline of the launches of one infer_action call replayed from its CUDA graph (option "trace"): block 0 of
every instrumented launch (tcgen05 GEMMs, attention, RMSNorm) stamps %globaltimer at entry / after its dependent-launch
wait / first operand stage landed / last MMA issued / first accumulator ready / epilogue done / exit. Prints, for one
layer of one denoise step (and optionally the prefill), every launch relative to the first one, plus the mean
wait-to-wait distance per launch kind over all layers of the denoise steps = what each kind costs on the chain.
    python tools/timeline.py [--batch 64] [--opts k=v,...] """
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
ap.add_argument("--show", default="den:5:10", help="stage:step:layer to print launch by launch (stage den|pre)")
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

# split the records of the main stream into stages: everything after the first "act_enc2" is a denoise step
recs = []
step = -1
for i, nm in enumerate(names):
    side = nm.endswith("[side]")
    base = nm.replace(" [side]", "")
    if base.startswith("act_enc2"):
        step += 1
    kind, _, lay = base.partition(" L")
    recs.append(dict(i=i, kind=kind, layer=int(lay) if lay else -1, side=side, step=step, t=tr[i]))

SLOTS = ["entry", "wait", "1st-stage", "mma-issued", "acc-ready", "epi-done", "exit"]


def show(sel, title):
    if not sel:
        return
    t0 = min(float(r["t"][0]) for r in sel if r["t"][0] > 0)
    print(f"---- {title} (us relative to the first entry)")
    print(f"{'launch':22s} " + " ".join(f"{s:>10s}" for s in SLOTS))
    for r in sel:
        vals = " ".join(f"{(float(v) - t0) / 1e3:10.2f}" if v > 0 else f"{'-':>10s}" for v in r["t"][:7])
        fine = " ".join(f"{(float(v) - t0) / 1e3:.2f}" for v in r["t"][8:14] if v > 0)
        allc = " ".join(f"{n} {(float(r['t'][i]) - t0) / 1e3:.2f}" for n, i in (("entry<=", 14), ("ready<=", 15), ("exit<=", 7))
                        if r["t"][i] > 0)
        print(f"{(r['kind'] + ' L' + str(r['layer'])):22s} {vals}" + (f"   all CTAs: {allc}" if allc else "")
              + (f"   fine: {fine}" if fine else ""))


stage, sstep, slayer = a.show.split(":")
sstep, slayer = int(sstep), int(slayer)
if stage == "den":
    show([r for r in recs if r["step"] == sstep and r["layer"] in (slayer, slayer + 1) and not r["side"]],
         f"denoise step {sstep}, layers {slayer}..{slayer + 1}")
else:
    show([r for r in recs if r["step"] == -1 and r["layer"] in (slayer, slayer + 1) and not r["kind"].startswith("vit")],
         f"prefill layers {slayer}..{slayer + 1} (main + side stream)")

# chain cost per kind over the denoise steps: distance from this launch's post-wait stamp to the next launch's
den = [r for r in recs if r["step"] >= 0 and not r["side"]]
cost = defaultdict(list)
for r, nx in zip(den[:-1], den[1:]):
    if r["t"][1] > 0 and nx["t"][1] > 0:
        cost[r["kind"]].append((float(nx["t"][1]) - float(r["t"][1])) / 1e3)
print("---- denoise steps: mean distance from a launch's post-wait stamp to the next launch's (us)")
tot = 0.0
for k, v in cost.items():
    m = sum(v) / len(v)
    tot += m * len(v)
    print(f"{k:12s} x{len(v):5d}  mean {m:7.2f}  min {min(v):7.2f}  max {max(v):7.2f}")
print(f"sum over the traced chain: {tot / 1e3:.3f} ms")
