"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, max, share."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, data = None, []
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr:
        data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data:
    k = d["Kernel Name"].split("(")[0][:48]
    v = float(d["Metric Value"].replace(",", ""))
    v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print(f"{len(data)} launches, {tot / 1e3:.2f} ms of GPU time")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} n={a[0]:4d} total={a[1]:10.1f} us  max={a[2]:10.1f} us  share={a[1] / tot * 100:5.1f}%")
