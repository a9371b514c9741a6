"""B = 1 latency of Vlaser-2B-VLA infer_action: p50 over graph replays for the two expert schedules (weight-streaming
single-observation kernel on / off), the stage split through `stop_after`, and the in-situ time of the expert passes."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import _device_state_dict_stream
from vlaser_b200.config import PRESETS
from vlaser_b200.policy import PiZeroInference
from vlaser_b200.synth import make_inputs

cfg = PRESETS[sys.argv[1] if len(sys.argv) > 1 else "vlaser_2b_bridge"]()
dev = torch.device("cuda", 0)
pol = PiZeroInference(cfg, max_batch=1)
for k, t in _device_state_dict_stream(cfg, 0, dev):
    pol.load_state_dict({k: t}, strict=False)
inp = make_inputs(cfg, 1, 999)
ids, pix = inp["input_ids"].to(dev), inp["pixel_values"].to(torch.bfloat16).to(dev)
prop, noise = inp["proprios"].to(torch.bfloat16).to(dev), inp["noise"].to(torch.bfloat16).to(dev)


def p50(n=21):
    for _ in range(3):
        pol.infer_action(ids, pix, proprios=prop, noise=noise)
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        pol.infer_action(ids, pix, proprios=prop, noise=noise)
        a1.record()
        torch.cuda.synchronize()
        ts.append(a0.elapsed_time(a1))
    return statistics.median(ts), min(ts)


for gemv in (1, 0):
    pol.set_option("expert_gemv", gemv)
    for stop, name in ((1, "ViT"), (2, "+projector/embed"), (3, "+prefill"), (4, "+1 denoise step"), (0, "whole call")):
        pol.set_option("stop_after", stop)
        med, mn = p50()
        print(f"expert_gemv={gemv} stop_after={stop} ({name}): p50 {med:.3f} ms  min {mn:.3f} ms  launches {pol.last_launch_count()}", flush=True)
    pol.set_option("stop_after", 0)
    pol.set_option("use_graph", 0)
    pol.infer_action(ids, pix, proprios=prop, noise=noise)
    torch.cuda.synchronize()
    pol.set_option("profile_mtot", int(inp["valid_len"].sum()))
    pol.set_option("profile_gemm", 1)
    for _ in range(3):
        pol.infer_action(ids, pix, proprios=prop, noise=noise)
    cls = pol.profile_read_classes()
    pol.set_option("profile_gemm", 0)
    pol.set_option("use_graph", 1)
    for name, c in cls.items():
        if c["launches"]:
            print(f"   class {name}: {c['launches'] // 3} launches/call, {c['ms'] / 3:.3f} ms/call, "
                  f"{c['bytes'] / (c['ms'] * 1e-3) / 1e9:.0f} GB/s algorithmic", flush=True)

# ---- per-phase timing of the weight-streaming kernel (CTA 0 stamps %globaltimer when it sees a phase complete grid-wide)
pol.set_option("expert_gemv", 1)
L, P = cfg.layers, 5
names = ["norm+qkv", "attention", "o-proj", "norm+gate/up", "down"]
for fl in (0, 32, 16, 4):
    pol.set_option("expert_timing", 1)
    pol.set_option("expert_flags", fl)
    for _ in range(3):
        pol.infer_action(ids, pix, proprios=prop, noise=noise)
    torch.cuda.synchronize()
    t = pol.debug_buffer("expert_timing", torch.int64).view(cfg.num_inference_steps + 1, L * 7).double()
    print(f"==== expert_flags = {fl} (2: no attention, 4: no weight units, 16: no L2 prefetch; 6 = barriers + activation staging only)")
    for ps, label in ((0, "proprio pass (1 row)"), (1, "denoise step 0"), (cfg.num_inference_steps, "last denoise step")):
        row = t[ps]
        n_ph = (L - 1) * P + (1 if ps == 0 else P)
        stamps = torch.cat([row[L * 7 - 1].view(1), row[: n_ph - 1]])
        dur = (stamps[1:] - stamps[:-1]) / 1e3
        print(f"{label}: {n_ph} phases, {float(stamps[-1] - stamps[0]) / 1e3:.1f} us to the last barrier")
        for j, nm in enumerate(names):
            dd = dur[j::P]
            if len(dd):
                print(f"   {nm:13s} mean {float(dd.mean()):7.2f} us  min {float(dd.min()):7.2f}  max {float(dd.max()):7.2f}   x{len(dd)}")
    row = t[1]
    sub = (row[L * P: L * P + 14] - row[L * P]) / 1e3
    print("   layer-1 stamps of CTA 0 (us since its phase-0 barrier wait ended): qkv[wait, staged, units done, synced] = "
          + " ".join(f"{float(v):.2f}" for v in sub[0:4]) + " | attn[wait, q+fresh, scores, softmax, pv, end] = "
          + " ".join(f"{float(v):.2f}" for v in sub[4:10]) + " | gate/up[wait, staged, units done, synced] = "
          + " ".join(f"{float(v):.2f}" for v in sub[10:14]))
pol.set_option("expert_flags", 0)
pol.set_option("expert_timing", 0)
