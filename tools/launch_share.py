#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of GPU time)."""
import collections, csv, sys
path = sys.argv[1]; start_at = sys.argv[2] if len(sys.argv) > 2 else None
rows = list(csv.reader(open(path, errors="replace")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
data = [r for r in rows[h + 1:] if len(r) > mv]
if start_at:  # keep launches from the first occurrence of a kernel name substring
    first = next((i for i, r in enumerate(data) if start_at in r[kn]), 0)
    data = data[first:]
agg = collections.OrderedDict()
for r in data:
    name = r[kn].split("(")[0].replace("void ", "").replace("im2im::<unnamed>::", "")[:64]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"{len(data)} launches, {tot / 1e3:.1f} us total (per-launch times are cold-cache and serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1] / 1e3:10.1f} us {100 * v[1] / tot:6.2f}%  n={v[0]:4d}  avg {v[1] / v[0] / 1e3:8.1f} us  {k}")
