"""Sums gpu__time_duration per kernel name from an ncu --csv launch list."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[hdr + 1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    name = r[ki][:90]
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
for k in sorted(tot, key=tot.get, reverse=True):
    print("%9.1f us %5.1f%% n=%4d avg %8.1f  %s" % (tot[k], 100 * tot[k] / s, cnt[k], tot[k] / cnt[k], k))
print("total %.1f us" % s)
