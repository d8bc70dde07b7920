import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import gskrige
name = sys.argv[1] if len(sys.argv) > 1 else "C5"
spec = gskrige.synth.config_spec(name)
ctx = gskrige.Context(0)
for i in range(3):
    t0 = time.perf_counter(); ctx.plan(spec); t1 = time.perf_counter()
    print(f"plan {i}: wall {1e3*(t1-t0):.1f} ms, events {ctx.timing()['ms_plan']:.1f} ms", flush=True)
