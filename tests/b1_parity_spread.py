"""Test infrastructure (run by hand / tools/gpu_r01aa.sh, not collected by pytest). Spread of the B = 1 full-size parity error over repeated calls (the split-K partial sums land with fp32 atomics, so the
summation order -- not the math -- changes from call to call): max-abs error of the clamped actions vs the fp32 oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import torch

from oracle import vlaser_oracle as O  # checker only
from vlaser_b200.config import vlaser_2b_bridge
from vlaser_b200.policy import PiZeroInference
from vlaser_b200.synth import make_inputs
from vlaser_b200.weights import random_state_dict

cfg = vlaser_2b_bridge().with_(final_action_clip_value=None)
sd = random_state_dict(cfg, 0)
for seed in (17, 18):
    inp = make_inputs(cfg, 1, seed)
    ref = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"]).clamp(-1, 1)
    for opts in ({}, {"small_m_split": 0}, {"small_m_split": 0, "expert_gemv": 0}):
        pol = PiZeroInference(cfg, max_batch=1)
        pol.load_state_dict(sd, strict=True)
        for k, v in opts.items():
            pol.set_option(k, v)
        errs, outs = [], []
        for i in range(6):
            out = pol.infer_action(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), proprios=inp["proprios"].cuda(),
                                   noise=inp["noise"].cuda()).float().cpu().clamp(-1, 1)
            errs.append(float((out - ref).abs().max()))
            outs.append(out)
        spread = max(float((a - outs[0]).abs().max()) for a in outs)
        print(f"seed {seed} {opts}: err vs oracle " + " ".join(f"{e:.2e}" for e in errs) + f" | run-to-run {spread:.2e}", flush=True)
        del pol
