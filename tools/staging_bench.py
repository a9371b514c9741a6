"""Times the input-staging kernel (vlaser_op_stage_frames: u8 640x480 frames -> Lanczos4 resize -> normalise -> bf16
NCHW) and the action post-processing kernel with CUDA events on the launch stream. Inputs rotate over buffers whose
total size exceeds the 126 MB L2, so every launch reads its frames from HBM. Algorithmic bytes per frame = H*W*3 read
once + 3*px*px*2 written."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vlaser_b200.adapter import ActionPostprocessor, FrameStager

dev = torch.device("cuda", 0)
peak = 6551.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
res = {}
for (n, h, w, px) in [(64, 480, 640, 224), (64, 512, 640, 224), (64, 224, 224, 224), (192, 480, 640, 224)]:
    bufs = [torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=dev) for _ in range(max(2, int(300e6 // (n * h * w * 3)) + 1))]
    st = FrameStager(px)
    out = torch.empty((n, 3, px, px), dtype=torch.bfloat16, device=dev)
    for b in bufs[:3]:
        st(b, out)
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        st(bufs[i % len(bufs)], out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    byts = n * (h * w * 3 + 3 * px * px * 2)
    res[f"stage_{n}x{h}x{w}_to_{px}"] = {"us_per_launch": us, "algorithmic_mbytes": byts / 1e6, "achieved_gbs": byts / us / 1e3,
                                        "peak_gbs": peak, "frac": byts / us / 1e3 / peak}
stats = {"p01": [-0.03] * 7, "p99": [0.03] * 7, "mean": [0.0] * 7, "std": [0.01] * 7}
for n_env in (64, 512):
    pp = ActionPostprocessor(stats, "bound", "sticky", n_env=n_env)
    a = torch.rand(n_env, 4, 7, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        pp(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        pp(a)
    e1.record()
    torch.cuda.synchronize()
    res[f"postprocess_{n_env}env_x4"] = {"us_per_launch": e0.elapsed_time(e1) * 1e3 / 50}
print(json.dumps(res, indent=1))
