"""Summarises an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch list of
ONE policy call into per-kernel-family DRAM traffic per launch (read + write) and writes profiles/<tag>_traffic.json,
which bench.py quotes as `roofline.traffic` (ncu cannot run inside the timed bench)."""
import csv
import json
import re
import sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
FULL_GRID = int(sys.argv[3]) if len(sys.argv) > 3 else 148
with open(src, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
acc = defaultdict(lambda: defaultdict(float))
cnt = defaultdict(set)
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
         "ms": 1e3, "msecond": 1e3}
for r in csv.DictReader(lines):
    name = re.sub(r"[<(].*$", "", r["Kernel Name"]).replace("vlaser::(anonymous namespace)::", "").replace("void ", "")
    name = name.split("::")[-1]
    # the CTA-pair GEMM runs a full grid (one pair per SM pair) only when it has at least that many 256x256 tiles, i.e.
    # at large M (ViT / LLM-prefill projections); the small-M launches of the action expert use fewer CTAs. bench.py's
    # roofline line is about the large-M launches, so they get their own family.
    if name == "gemm_tcgen05_pair_kernel" and r.get("Grid Size", "").replace(" ", "") == f"({FULL_GRID},1,1)":
        name = "gemm_tcgen05_pair_kernel@full_grid"
    v = float(r["Metric Value"].replace(",", "")) * scale.get(r.get("Metric Unit", ""), 1.0)
    acc[name][r["Metric Name"]] += v
    cnt[name].add(r["ID"])
out = {}
for k, m in acc.items():
    n = len(cnt[k])
    out[k] = {"launches_per_call": n,
              "dram_read_bytes_per_launch": m.get("dram__bytes_read.sum", 0.0) / n,
              "dram_write_bytes_per_launch": m.get("dram__bytes_write.sum", 0.0) / n,
              "traffic_bytes_per_launch": (m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)) / n,
              "avg_launch_us_under_ncu": m.get("gpu__time_duration.sum", 0.0) / n}
json.dump({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over one policy call "
                     "(tools/profile_call.py --graph 0, B=64, Vlaser-2B-VLA)", "kernels": out}, open(dst, "w"), indent=1)
for k, v in sorted(out.items(), key=lambda kv: -kv[1]["traffic_bytes_per_launch"] * kv[1]["launches_per_call"]):
    print(f"{k:32s} x{v['launches_per_call']:5d}  traffic/launch {v['traffic_bytes_per_launch'] / 1e6:10.2f} MB  avg {v['avg_launch_us_under_ncu']:9.1f} us")
