"""Top stall hot spots of an ncu report (SASS level with the dominant stall reason), one table per kernel in the file:
   python tools/ncu_hotspots.py file.ncu-rep [topN] [kernel_index]"""
import csv
import subprocess
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
only = int(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the csv holds one block per kernel: a "Kernel Name" row, a header row, then the SASS lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1] if len(r) > 1 else "?", "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for bi, b in enumerate(blocks):
    if only is not None and bi != only:
        continue
    hdr, data = b["hdr"], b["data"]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print(f"== kernel {bi}: {b['name'][:100]}: {len(data)} SASS lines, {tot} samples")
    agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stall_cols}
    print("  by reason: " + ", ".join(f"{k[6:]}={100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:topn]
    for i in sorted(order):
        r = data[i]
        n = int(r[ix["# Samples"]] or 0)
        top = max(stall_cols, key=lambda s: int(r[ix[s]] or 0))
        print(f"  [{i:5d}] {100 * n / max(tot, 1):5.1f}%  {top[6:]:14s} x{r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:90]}")
