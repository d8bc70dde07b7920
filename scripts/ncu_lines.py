"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1] if len(r) > 1 else "?"; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] != "":   # a source line summary row
        try:
            key = (cur_file, int(r[0]))
        except ValueError:
            continue
        d = dict(zip(hdr[4:], r[4:]))
        def _i(x):
            try: return int(x)
            except (TypeError, ValueError): return 0
        inst = _i(d.get("Instructions Executed"))
        samp = _i(d.get("# Samples"))
        a = agg.setdefault(key, [0, 0, r[1][:90]])
        a[0] += inst; a[1] += samp
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"total inst {tot_i:,}  samples {tot_s:,}")
for (f, ln), (i, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{ln:4d} inst {100*i/tot_i:5.1f}%  samples {100*s/tot_s:5.1f}%  | {src}")
