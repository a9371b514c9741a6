"""Top instructions by warp-stall samples of an `ncu --page source --csv` export (one section per profiled launch)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        secs.append(cur)
        continue
    if cur is None:
        continue
    if cur["hdr"] is None:
        cur["hdr"] = r
        continue
    cur["rows"].append(r)
for sec in secs:
    h = {k: i for i, k in enumerate(sec["hdr"])}
    tot = sum(int(r[h["# Samples"]]) for r in sec["rows"] if len(r) > h["# Samples"])
    print(f"==== {sec['name'][:90]}: {len(sec['rows'])} instructions, {tot} samples")
    stall_cols = [k for k in sec["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
    for r in sorted(sec["rows"], key=lambda r: -int(r[h["# Samples"]]))[:topn]:
        st = sorted(((int(r[h[k]]), k[6:]) for k in stall_cols), reverse=True)[:2]
        print(r[h["# Samples"]].rjust(6), r[h["Address"]][-5:], r[h["Source"]].strip()[:64].ljust(64), st)
