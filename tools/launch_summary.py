"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py launches.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as fh:
    lines = [ln for ln in fh if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rd:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    name = re.sub(r"\(.*$", "", r[ik]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"total device time {total:.1f} ms over {sum(cnt.values())} launches\n")
print("| kernel | launches | time ms | share % | avg us |\n|---|---|---|---|---|")
for k in sorted(tot, key=tot.get, reverse=True)[:25]:
    print(f"| `{k}` | {cnt[k]} | {tot[k]:.3f} | {100 * tot[k] / total:.1f} | {1e3 * tot[k] / cnt[k]:.1f} |")
