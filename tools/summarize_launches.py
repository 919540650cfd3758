"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel total / share / count / average."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
per = collections.OrderedDict()
tot = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:110]
    d = per.setdefault(name, [0.0, 0])
    d[0] += us; d[1] += 1; tot += us
print(f"total {tot/1e3:.2f} ms over {sum(d[1] for d in per.values())} launches")
for name, (us, n) in sorted(per.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{us/1e3:9.3f} ms {100*us/tot:5.1f}% n={n:4d} avg={us/n:8.1f} us  {name}")
