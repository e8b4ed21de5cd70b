#!/usr/bin/env python
"""Top stall reasons and hottest SASS lines of one kernel in an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
idx = sys.argv[3] if len(sys.argv) > 3 else "1"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{pat}:{idx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:120])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data) or 1
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print("samples", tot, " ".join(f"{s[6:]}={100*v/tot:.0f}%" for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print("warp-instr executed:", int(sum(f(r, "Instructions Executed") for r in data)))
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:int(sys.argv[4]) if len(sys.argv) > 4 else 16]:
    st = sorted(((s[6:], f(r, s)) for s in stalls), key=lambda kv: -kv[1])[:2]
    print(f"{100*f(r,'# Samples')/tot:5.1f}% exec={int(f(r,'Instructions Executed')):9d} {r[ix['Source']].strip()[:64]:64s} {st[0][0]}:{int(st[0][1])} {st[1][0]}:{int(st[1][1])}")
