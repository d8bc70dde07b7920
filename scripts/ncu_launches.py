"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python scripts/ncu_launches.py file.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1e-6)
        k = d["Kernel Name"].split("(")[0]
        tot[k] += v
        cnt[k] += 1
allms = sum(tot.values())
print(f"{sum(cnt.values())} launches, {allms:.2f} ms in kernels (each launch profiled alone: cold caches, serialised)")
for k, v in tot.most_common():
    print(f"{100 * v / allms:5.1f} %  {v:10.3f} ms  {cnt[k]:5d} launches  {v / cnt[k] * 1e3:9.1f} us each  {k}")
