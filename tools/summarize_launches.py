"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel.  usage: summarize_launches.py file.csv [last_n]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
recs = list(csv.DictReader(lines))
if len(sys.argv) > 2:
    recs = recs[-int(sys.argv[2]):]
agg = collections.OrderedDict()
for x in recs:
    name = re.sub(r"\(.*", "", x["Kernel Name"])[:64]
    try:
        t = float(x["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    if t != t:
        continue
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += t
tot = sum(v[1] for v in agg.values())
print(f"{len(recs)} launches, {tot / 1e3:.1f} us total")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:66s} n={v[0]:4d} total={v[1] / 1e3:9.1f} us  share={v[1] / tot * 100:5.1f}%  avg={v[1] / v[0] / 1e3:8.1f} us")
