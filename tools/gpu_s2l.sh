#!/bin/bash
OUT=gpurun_out/r02s2l
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_generate_8b.csv python tools/profile_generate.py > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu.log
python - <<'PY'
import csv,re
from collections import defaultdict
lines=[l for l in open("gpurun_out/r02s2l/launches_generate_8b.csv") if not l.startswith("==")]
rows=[r for r in csv.DictReader(lines) if r.get("Metric Name")=="gpu__time_duration.sum"]
# the decode step = everything after the last argmax_rows_kernel but one
names=[re.sub(r"\(.*$","",r["Kernel Name"]).replace("vlaser::(anonymous namespace)::","").replace("void ","") for r in rows]
val=[float(r["Metric Value"].replace(",",""))/1e3 for r in rows]
idx=[i for i,n in enumerate(names) if "argmax_rows" in n]
print("argmax launches at", idx, "of", len(names))
lo=idx[0]+1
tot=defaultdict(float); cnt=defaultdict(int)
for n,v in zip(names[lo:],val[lo:]):
    tot[n]+=v; cnt[n]+=1
print("decode step: %.1f us over %d launches"%(sum(tot.values()), sum(cnt.values())))
for k in sorted(tot,key=tot.get,reverse=True): print("%9.1f us x%4d avg %7.1f  %s"%(tot[k],cnt[k],tot[k]/cnt[k],k))
print("prefill + first token: %.1f ms"%(sum(val[:lo])/1e3))
PY
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention" 2>&1 | tail -2
