import sys, json
# usage: show_bench.py FILE   or   ... | show_bench.py -   (never waits on a terminal/inherited stdin by accident)
src = open(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1] != "-" else sys.stdin
for l in src:
    try:
        d = json.loads(l)
    except Exception:
        print(l, end=""); continue
    r = d.get("roofline", {})
    print("value %.4g  e2e %.4g  ms/step %.3f  frac %.3f  achieved %.2f/%.2f TF  phases %s  clocks %s  cpu %s" % (
        d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("frac") or 0, r.get("achieved") or 0, r.get("peak") or 0,
        d.get("phases_ms"), d.get("clocks"), (d.get("cpu_baseline") or {}).get("value")))
