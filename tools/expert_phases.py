"""Per-phase timing of the persistent expert-stream kernel (option "expert_timing"): CTA 0 stamps %globaltimer when it
sees each phase complete grid-wide. Prints the mean duration of each of the 7 phase kinds over the layers of one
denoise pass and of the proprio pass (B=64, Vlaser-2B-VLA shapes, device-generated random weights)."""
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
ap.add_argument("--flags", default="0", help="comma list of expert_flags values to time (profiling aids)")
ap.add_argument("--streams", default="", help="comma list of expert_streams values (sub-grids); stamps are those of sub-grid 0")
a = ap.parse_args()
cfg = PRESETS[a.config]()
dev = torch.device("cuda", 0)
pol = PiZeroInference(cfg, max_batch=a.batch)
for k, t in _device_state_dict_stream(cfg, 0, dev):
    pol.load_state_dict({k: t}, strict=False)
inp = make_inputs(cfg, a.batch, 1)
ids, pix = inp["input_ids"].to(dev), inp["pixel_values"].to(torch.bfloat16).to(dev)
prop, noise = inp["proprios"].to(torch.bfloat16).to(dev), inp["noise"].to(torch.bfloat16).to(dev)
L, P = cfg.layers, 7
names = ["norm1", "qkv", "attn", "o", "norm2", "gate_up", "down"]
runs = [(int(f), None) for f in a.flags.split(",")] if not a.streams else \
    [(int(f), int(st)) for st in a.streams.split(",") for f in a.flags.split(",")]
for fl, st in runs:
  if st is not None:
      pol.set_option("expert_streams", st)
      print(f"######## expert_streams = {st}")
  pol.set_option("expert_timing", 1)
  pol.set_option("expert_flags", fl)
  for _ in range(3):
    pol.infer_action(ids, pix, proprios=prop, noise=noise)
  torch.cuda.synchronize()
  print(f"==== expert_flags = {fl}")
  t = pol.debug_buffer("expert_timing", torch.int64).view(cfg.num_inference_steps + 1, L * P).double()
  for ps, label in ((0, "proprio pass (M=B)"), (1, "denoise step 0 (M=B*H)"), (cfg.num_inference_steps, "last denoise step")):
      row = t[ps]
      start = row[L * P - 1]
      n_ph = (L - 1) * P + (2 if ps == 0 else P)
      stamps = torch.cat([start.view(1), row[: n_ph - 1]])
      dur = (stamps[1:] - stamps[:-1]) / 1e3   # us; phase i ends at stamps[i+1]
      print(f"{label}: {n_ph} phases, {float(stamps[-1] - stamps[0]) / 1e3:.1f} us to the last barrier")
      for j, nm in enumerate(names):
          d = dur[j::P]
          if len(d):
              print(f"   {nm:8s} mean {float(d.mean()):7.2f} us  min {float(d.min()):7.2f}  max {float(d.max()):7.2f}   x{len(d)}")
