"""Summarise an `ncu --page source --csv --print-source sass` dump: LSU wavefronts / tag requests / stall samples
by instruction class (and top instructions).  usage: python profiles/src_summary.py dump.csv [topN]"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = rows[1]
ix = {k: i for i, k in enumerate(h)}


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", "") or 0)
    except (ValueError, KeyError, IndexError):
        return 0.0


data = [r for r in rows[2:] if len(r) > 10]
tot = defaultdict(float)
by = defaultdict(lambda: defaultdict(float))
for r in data:
    src = r[ix["Source"]]
    op = re.sub(r"@!?U?P\d+\s+", "", src).split()[0] if src.split() else "?"
    cls = op.split(".")[0]
    if cls in ("LDS", "STS", "LDG", "STG", "RED", "REDG", "ATOMS", "ATOMG", "LD", "ST", "UBLKCP", "SYNCS", "LDSM"):
        key = op if cls in ("LDS", "STS") else cls
    else:
        key = "other"
    for k in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Tag Requests Global", "L2 Theoretical Sectors Global", "Instructions Executed", "# Samples"):
        by[key][k] += f(r, k)
        tot[k] += f(r, k)
print("totals:", {k: int(v) for k, v in tot.items()})
print(f"{'class':24s} {'inst':>10s} {'wf_shared':>10s} {'ideal':>10s} {'tag_glob':>10s} {'l2_sect':>10s} {'samples':>8s}")
for key, d in sorted(by.items(), key=lambda kv: -(kv[1]['L1 Wavefronts Shared'] + kv[1]['L1 Tag Requests Global'])):
    print(f"{key:24s} {int(d['Instructions Executed']):10d} {int(d['L1 Wavefronts Shared']):10d} {int(d['L1 Wavefronts Shared Ideal']):10d} {int(d['L1 Tag Requests Global']):10d} {int(d['L2 Theoretical Sectors Global']):10d} {int(d['# Samples']):8d}")
print("\ntop instructions by shared wavefronts + global tag requests:")
data.sort(key=lambda r: -(f(r, "L1 Wavefronts Shared") + f(r, "L1 Tag Requests Global")))
for r in data[:top]:
    print(f"{r[ix['Address']][-5:]} {r[ix['Source']][:70]:70s} inst={int(f(r,'Instructions Executed')):8d} wf={int(f(r,'L1 Wavefronts Shared')):8d} ideal={int(f(r,'L1 Wavefronts Shared Ideal')):8d} tag={int(f(r,'L1 Tag Requests Global')):8d} smp={int(f(r,'# Samples'))}")
print("\ntop instructions by stall samples:")
data.sort(key=lambda r: -f(r, "# Samples"))
for r in data[:top]:
    st = {k: int(f(r, k)) for k in h if k.startswith("stall_") and "Not Issued" not in k and f(r, k) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{r[ix['Address']][-5:]} {r[ix['Source']][:60]:60s} smp={int(f(r,'# Samples')):6d} {st}")
