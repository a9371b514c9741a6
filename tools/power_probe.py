"""Sustained-regime A/B: every spec runs back to back for --secs seconds (graph replay) while a thread samples SM clock
and board power through NVML; prints ms per call, mean / min SM clock and mean power per spec. A 10-call A/B on a
power-capped board compares clocks as much as code; this one holds each variant long enough for the cap to settle.
    python tools/power_probe.py --secs 3 "known_splits=1" "known_splits=0" "known_splits=1,stop_after=3" """
import argparse
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
import torch

from bench import _device_state_dict_stream
from vlaser_b200.config import PRESETS
from vlaser_b200.policy import PiZeroInference
from vlaser_b200.synth import make_inputs

ap = argparse.ArgumentParser()
ap.add_argument("specs", nargs="*", default=["pdl=1"])
ap.add_argument("--secs", type=float, default=3.0)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--rounds", type=int, default=2)
a = ap.parse_args()
cfg = PRESETS["vlaser_2b_bridge"]()
dev = torch.device("cuda", 0)
pol = PiZeroInference(cfg, max_batch=a.batch)
for k, t in _device_state_dict_stream(cfg, 0, dev):
    pol.load_state_dict({k: t}, strict=False)
inp = make_inputs(cfg, a.batch, 1)
ids, pix = inp["input_ids"].to(dev), inp["pixel_values"].to(torch.bfloat16).to(dev)
prop, noise = inp["proprios"].to(torch.bfloat16).to(dev), inp["noise"].to(torch.bfloat16).to(dev)
pynvml.nvmlInit()
hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False


def sampler():
    while not stop:
        try:
            samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetPowerUsage(hnd) / 1e3))
        except pynvml.NVMLError:
            pass
        time.sleep(0.005)


th = threading.Thread(target=sampler, daemon=True)
th.start()
for rnd in range(a.rounds):
    for spec in a.specs:
        for kv in spec.split(","):
            k, v = kv.split("=")
            pol.set_option(k, int(v))
        for _ in range(3):
            pol.infer_action(ids, pix, proprios=prop, noise=noise)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.perf_counter() - t0 < a.secs:
            for _ in range(5):
                pol.infer_action(ids, pix, proprios=prop, noise=noise)
            n += 5
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms = e0.elapsed_time(e1) / n
        # the second half of the window: the cap has settled
        win = [s for s in samples if t0 + 0.5 * (t1 - t0) <= s[0] <= t1]
        clk = [s[1] for s in win] or [0]
        pw = [s[2] for s in win] or [0]
        print(f"round {rnd} {spec:40s} {ms:8.3f} ms/call  sm clock mean {sum(clk) / len(clk):6.0f} min {min(clk)} max {max(clk)} MHz"
              f"  power mean {sum(pw) / len(pw):6.0f} W  ({n} calls)", flush=True)
        pol.set_option("stop_after", 0)
stop = True
