"""Per CUDA source line stall samples from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
cur_file = ""
out = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix_s = hdr.index("# Samples")
        stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or r[0] in ("Function Name", ""):
        continue
    try:
        n = int(r[ix_s])
    except (ValueError, IndexError):
        continue
    st = {h[6:]: int(r[i]) for i, h in stall if i < len(r) and r[i].isdigit() and int(r[i]) > 0}
    out.append((n, cur_file, r[0], r[1].strip()[:100], dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])))
tot = sum(o[0] for o in out)
print("total samples", tot)
for n, f, ln, src, st in sorted(out, key=lambda o: -o[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{n:6d} {100*n/tot:5.1f}% {f}:{ln:>4s} {src:100s} {st}")
