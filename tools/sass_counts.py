"""Per-kernel SASS instruction counts of the built library (tcgen05 / TMEM / TMA / legacy tensor-core mnemonics):
    python tools/sass_counts.py > profiles/<tag>_sass_counts.txt
UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk,
HMMA = mma.sync (legacy path), SYNCS = mbarrier ops. B200_PROFILING.md lists the mnemonics."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
so = ROOT / "pocket-tts.c_b200" / "libptts_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "LDSM", "FFMA", "BAR.SYNC"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    f = re.match(r"\s*Function : (\S+)", line)
    if f:
        name = subprocess.run(["cu++filt", f.group(1)], capture_output=True, text=True).stdout.strip()
        depth, cut = 0, len(name)
        for i, ch in enumerate(name):     # drop the parameter list, keep the template arguments
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                cut = i
                break
        cur = name[:cut].replace("(int)", "").replace("void ", "").replace("ptts::", "")
        counts[cur] = counts.get(cur, collections.Counter())
        continue
    if cur is None or "/*" not in line:
        continue
    ins = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
    if not ins:
        continue
    op = ins.group(1)
    counts[cur]["_all"] += 1
    for mn in MN:
        if op == mn or op.startswith(mn + ".") or (mn == "BAR.SYNC" and op.startswith("BAR.SYNC")):
            counts[cur][mn] += 1
            total[mn] += 1
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
print(f"SASS instruction counts per kernel of libptts_b200.so (sm_100a), kernel sources {bench.src_sha16()}")
print(f"{'kernel':64s} {'instr':>7s} " + " ".join(f"{m:>8s}" for m in MN))
for k, c in counts.items():
    print(f"{k[:64]:64s} {c['_all']:7d} " + " ".join(f"{c[m]:8d}" for m in MN))
print(f"{'TOTAL':64s} {sum(c['_all'] for c in counts.values()):7d} " + " ".join(f"{total[m]:8d}" for m in MN))
