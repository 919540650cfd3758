"""Per-kernel counts of the Blackwell-only SASS mnemonics in liboctic_b200.so (tcgen05 MMA / TMEM / TMA / cluster barriers):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
Evidence that the hot kernels are tcgen05 / TMEM / TMA code (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTMALDG = cp.async.bulk.tensor load, UTCBAR = tcgen05.commit, SYNCS = mbarrier) and which ones still use HMMA (mma.sync)."""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "octic_vits_b200" / "lib" / "liboctic_b200.so"
PATS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA2", "MUFU.EX2",
        "REDG", "ATOMG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", out)
    if names:
        dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    cur, counts = None, OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = dict.fromkeys(PATS, 0)
            continue
        if cur is None or "/*" not in line:
            continue
        for p in PATS:
            if p == "UTCHMMA":
                if re.search(r"\bUTC[A-Z]*MMA", line):
                    counts[cur][p] += 1
            elif p == "UTCHMMA.2CTA":
                if re.search(r"\bUTC[A-Z]*MMA\S*\.2CTA", line):
                    counts[cur][p] += 1
            elif p == "HMMA":
                if re.search(r"\bHMMA", line):
                    counts[cur][p] += 1
            elif p in line:
                counts[cur][p] += 1
    print("# " + " ".join(f"{p:>12s}" for p in PATS) + "   kernel")
    for k, c in counts.items():
        if not any(c.values()):
            continue
        name = demangle.get(k, k).replace("(int)", "")
        name = (name[:name.index(">(") + 1] if ">(" in name else re.sub(r"\(.*", "", name))[:110]
        print("  " + " ".join(f"{c[p]:12d}" for p in PATS) + "   " + name)


if __name__ == "__main__":
    sys.exit(main())
