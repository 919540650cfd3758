"""Top stall lines of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv): SASS rows sorted by samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_")]
print("total samples", tot)
agg = {}
for h in stall_cols:
    agg[h] = sum(int(r[ix[h]] or 0) for r in data if len(r) > ix[h])
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    st = {h[6:]: int(r[ix[h]]) for h in stall_cols if len(r) > ix[h] and r[ix[h]] and int(r[ix[h]]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{int(r[ix['# Samples']]):6d} {100*int(r[ix['# Samples']])/tot:5.1f}%  {r[ix['Source']][:90]:90s} {st}")
