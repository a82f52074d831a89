"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) != len(hdr): continue
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]); name = name.split("::")[-1]
    v = float(r[ci["Metric Value"]].replace(",", "")); u = r[ci["Metric Unit"]]
    us = v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"  {k:44s} {n:5d} launches {t:10.1f} us  {100*t/tot:5.1f}%  ({t/n:7.1f} us each)")
