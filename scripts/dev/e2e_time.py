import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import gskrige
name = sys.argv[1] if len(sys.argv) > 1 else "C5"
count = int(sys.argv[2]) if len(sys.argv) > 2 else 2097152
use_torch = len(sys.argv) > 3
spec = gskrige.synth.config_spec(name).with_slab(0, count)
if use_torch:
    import torch
    torch.cuda.set_device(0)
    x = torch.empty(256*1024*1024//4, device="cuda")
    hm = torch.empty(count, dtype=torch.float64).pin_memory().numpy(); hv = torch.empty(count, dtype=torch.float64).pin_memory().numpy()
else:
    hm = np.empty(count); hv = np.empty(count)
ctx = gskrige.Context(0)
for i in range(4):
    t0 = time.perf_counter(); ctx.krige_into(spec, hm, hv); t1 = time.perf_counter()
    tm = ctx.timing()
    print(f"krige {i}: wall {1e3*(t1-t0):.1f} ms, plan {tm['ms_plan']:.1f} ms, exec total {tm['ms_total']:.1f}", flush=True)
