"""Per CUDA source line: shared-memory wavefronts and the excess caused by bank conflicts, from an
`ncu --page source --csv --print-source cuda,sass` dump."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1] if len(r) > 1 else "?"; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue
    try: key = (cur, int(r[0]))
    except ValueError: continue
    d = dict(zip(hdr, r))
    def f(k):
        try: return float(d.get(k) or 0)
        except ValueError: return 0.0
    a = agg.setdefault(key, [0.0, 0.0, r[1][:80]])
    a[0] += f("L1 Wavefronts Shared"); a[1] += f("L1 Wavefronts Shared Excessive")
tot = sum(a[0] for a in agg.values()) or 1.0
print(f"shared wavefronts {tot:,.0f}, excessive {sum(a[1] for a in agg.values()):,.0f}")
for (fn, ln), (w, e, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{fn}:{ln:4d} wavefronts {100*w/tot:5.1f}%  excessive {100*e/tot:5.1f}%  | {src}")
