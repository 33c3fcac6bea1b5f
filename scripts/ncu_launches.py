"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel. Usage: ncu_launches.py file.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            name = d["Kernel Name"]
            if any(t in name for t in ("at::", "elementwise", "reduce_kernel", "distribution_")):   # torch's own kernels
                continue
            k = name.split("(")[0][-48:] + " grid=" + d["Grid Size"] + " blk=" + d["Block Size"]
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += float(d["Metric Value"].replace(",", ""))
tot = 0.0
for k, (n, t) in agg.items():
    print(f"{n:4d} x {t / n / 1000:9.2f} us  {k}")
