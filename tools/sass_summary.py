"""Instruction histogram of every kernel in the shipped library (`cuobjdump -sass`), reduced to the mnemonics that show
what kind of kernel it is: tcgen05 MMA (UTCHMMA, .2CTA = cta_group::2), TMEM loads / stores (LDTM / STTM), TMA tensor
loads / stores / reduce-adds (UTMALDG / UTMASTG / UTMAREDG), bulk copies (UBLKCP), legacy tensor-core MMA (HMMA),
cp.async (LDGSTS), MUFU, local-memory spills (STL / LDL). Regenerate profiles/r02_sass_summary.txt with:
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vlaser_b200", "libvlaser_b200.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTMAPF", "HMMA", "LDGSTS",
        "LDSM", "MUFU", "RED", "ATOM", "STL", "LDL", "SYNCS", "UCGABAR"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
names = iter(demangle)
kernels = OrderedDict()
cur = None
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        nm = next(names)
        nm = re.sub(r"vlaser::\(anonymous namespace\)::", "", nm)
        nm = re.sub(r"\(.*$", "", nm).replace("void ", "")
        cur = kernels.setdefault(nm, Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                if k == "UTCHMMA" and ".2CTA" in op:
                    continue
                cur[k] += 1
                break
print(f"# {os.path.relpath(lib, ROOT)}: SASS instruction counts per kernel (static), sm_100a")
print(f"{'kernel':78s} {'instrs':>7s}  " + " ".join(f"{k}" for k in KEYS))
for nm, c in kernels.items():
    row = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
    print(f"{nm[:78]:78s} {c['_total']:7d}  {row}")
