import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
    except Exception:
        print(l, end=""); continue
    r = d.get("roofline", {})
    print("value %.4g  e2e %.4g  ms/step %.3f  frac %.3f  achieved %.2f/%.2f TF  phases %s  clocks %s  cpu %s" % (
        d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("frac") or 0, r.get("achieved") or 0, r.get("peak") or 0,
        d.get("phases_ms"), d.get("clocks"), (d.get("cpu_baseline") or {}).get("value")))
