"""Per-kernel SASS mnemonic histogram of libts_b200.so (cuobjdump -sass), written as a CSV + a markdown table of the
Blackwell-specific mnemonics (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP/UTMALDG = TMA, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, *.STRONG.SYS = system-scope peer accesses).  No GPU needed.

    python tools/sass_summary.py [profiles/r2_sass]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tianshou_b200", "libts_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass")
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda names: subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()  # noqa: E731
kernels: dict[str, collections.Counter] = {}
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
names = list(kernels)
pretty = dict(zip(names, demangle(names), strict=True))
short = lambda n: re.sub(r"\(.*", "", re.sub(r"\(anonymous namespace\)::", "", pretty[n]))  # noqa: E731
allm = sorted({m for c in kernels.values() for m in c})
with open(out + "_mnemonics.csv", "w") as f:
    f.write("kernel,instructions," + ",".join(allm) + "\n")
    for n in names:
        f.write(f"\"{short(n)}\",{sum(kernels[n].values())}," + ",".join(str(kernels[n].get(m, 0)) for m in allm) + "\n")
groups = {"UTC*MMA (tcgen05.mma)": r"^UTC[A-Z]*MMA", "LDTM (tcgen05.ld)": r"^LDTM", "STTM (tcgen05.st)": r"^STTM",
          "UBLKCP / UTMA* (TMA)": r"^(UBLKCP|UTMA)", "UTCBAR (tcgen05.commit)": r"^UTCBAR", "SYNCS (mbarrier)": r"^SYNCS",
          "MUFU": r"^MUFU", "*.STRONG.SYS (peer)": r"STRONG\.SYS", "HMMA/HGMMA (legacy)": r"^(HMMA|HGMMA|IMMA)"}
with open(out + "_blackwell.md", "w") as f:
    f.write("# SASS mnemonic counts per kernel (`cuobjdump -sass tianshou_b200/libts_b200.so`, static counts; full histogram: "
            f"`{os.path.basename(out)}_mnemonics.csv`, regenerate with `python tools/sass_summary.py`)\n\n")
    f.write("| kernel | instructions | " + " | ".join(groups) + " |\n|---|---:|" + "---:|" * len(groups) + "\n")
    for n in sorted(names, key=lambda k: -sum(kernels[k].values())):
        row = [sum(v for m, v in kernels[n].items() if re.search(rx, m)) for rx in groups.values()]
        f.write(f"| `{short(n)}` | {sum(kernels[n].values())} | " + " | ".join(str(x) if x else "—" for x in row) + " |\n")
print(f"{len(names)} kernels -> {out}_mnemonics.csv, {out}_blackwell.md")
