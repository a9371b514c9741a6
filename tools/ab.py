"""A/B timing of schedule options on one policy object (graph replay, CUDA events, B=64 and B=1).
    python tools/ab.py "pdl=0" "pdl=1" "pdl=2" "pdl=1,exp_stages=3" """
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import _device_state_dict_stream
from vlaser_b200.config import PRESETS
from vlaser_b200.policy import PiZeroInference
from vlaser_b200.synth import make_inputs

cfg = PRESETS["vlaser_2b_bridge"]()
dev = torch.device("cuda", 0)
pol = PiZeroInference(cfg, max_batch=64)
for k, t in _device_state_dict_stream(cfg, 0, dev):
    pol.load_state_dict({k: t}, strict=False)


def timed(B, iters=10):
    inp = make_inputs(cfg, B, 1)
    ids, pix = inp["input_ids"].to(dev), inp["pixel_values"].to(torch.bfloat16).to(dev)
    prop, noise = inp["proprios"].to(torch.bfloat16).to(dev), inp["noise"].to(torch.bfloat16).to(dev)
    for _ in range(3):
        out = pol.infer_action(ids, pix, proprios=prop, noise=noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = pol.infer_action(ids, pix, proprios=prop, noise=noise)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out.float().cpu()


ref = None
for spec in sys.argv[1:] or ["pdl=1"]:
    for kv in spec.split(","):
        k, v = kv.split("=")
        if k == "stop_after":
            continue
        try:
            pol.set_option(k, int(v))
        except RuntimeError as ex:
            print("option ignored:", kv, ex)
    stops = [int(kv.split("=")[1]) for kv in spec.split(",") if kv.startswith("stop_after=")] or [0]
    res = []
    d = float("nan")
    for stop in stops:
        pol.set_option("stop_after", stop)
        t64, out = timed(64)
        t1, _ = timed(1, 20)
        res.append((stop, t64, t1))
        if stop == 0:
            if ref is None:
                ref = out
            d = float((out - ref).abs().max())
    pol.set_option("stop_after", 0)
    print(spec, " ".join(f"[stop={s} B64 {a:.3f} ms B1 {b:.3f} ms]" for s, a, b in res), f"maxdiff_vs_first={d:.2e}",
          f"launches(last call)={pol.last_launch_count()}", flush=True)
