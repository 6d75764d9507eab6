#!/bin/bash
mkdir -p gpurun_out
for w in c3 c4; do
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$w.csv python scratch/c3_profile.py $w > gpurun_out/prof_$w.log 2>&1
echo "$w exit $?"
python - <<PY
import csv, collections, re
rows=list(csv.reader(open("gpurun_out/launches_$w.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[hdr+2:]:
    if len(r)<=vi: continue
    k=re.sub(r'\(.*','',r[ki]); k=re.sub(r'^void |d3b::|\(anonymous namespace\)::','',k)[:70]
    agg.setdefault(k,[0,0]); agg[k][0]+=float(r[vi].replace(',','')); agg[k][1]+=1
tot=sum(v[0] for v in agg.values())
print("$w total %.2f ms, %d launches"%(tot/1e6, sum(v[1] for v in agg.values())))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:14]: print("  %8.1f us x%-3d %s"%(v[0]/1000,v[1],k))
PY
done
