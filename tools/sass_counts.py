"""profiles/r02_sass_counts.txt: per kernel of libgar_b200.so, how many tcgen05 / TMEM / TMA / mma.sync instructions its
SASS holds (cuobjdump -sass; B200_PROFILING.md mnemonics: UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG
= TMA tensor load/store, UBLKCP = bulk copy, HMMA = mma.sync)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "grasp_any_region_b200", "libgar_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangled = {}
MN = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "MUFU.EX2", "LDG", "STG"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for mn in MN:
        if re.search(r"\b" + re.escape(mn) + r"\b", line) or (mn + ".") in line:
            counts[cur][mn] += 1
            break
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
out = ["# cuobjdump -sass grasp_any_region_b200/libgar_b200.so | per-kernel instruction counts (sm_100a)",
       "# " + "  ".join(f"{m:>8s}" for m in MN) + "  kernel"]
for (mangled, c), name in zip(counts.items(), names):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*$", "", name)[:110]
    out.append("  " + "  ".join(f"{c[m]:8d}" for m in MN) + "  " + name)
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
out.append("# total " + "  ".join(f"{tot[m]:>8d}" for m in MN))
text = "\n".join(out) + "\n"
dst = os.path.join(ROOT, "profiles", "r02_sass_counts.txt")
open(dst, "w").write(text)
sys.stdout.write(text)
