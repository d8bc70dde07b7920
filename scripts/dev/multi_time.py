"""Where does gsk_krige_multi lose time? (2 GPUs) python scripts/dev/multi_time.py"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import gskrige
spec = gskrige.synth.config_spec("C5")
T = spec.n_targets
count = T // 8
half = count // 2
def pinned(n):
    return torch.empty(n, dtype=torch.float64).pin_memory().numpy()
hm, hv = pinned(count), pinned(count)
for dev in (0, 1):
    c = gskrige.Context(dev)
    sl = spec.with_slab(dev * half, half)
    c.krige_into(sl, hm[:half], hv[:half])
    t0 = time.perf_counter(); c.krige_into(sl, hm[:half], hv[:half]); print(f"device {dev}: gsk_krige half slab {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
    c.close()
whole = spec.with_slab(0, count)
for ids in ([0], [0, 1], [0, 1], [1, 0]):
    gskrige.krige_multi(whole, ids, out=(hm, hv))
    t0 = time.perf_counter(); gskrige.krige_multi(whole, ids, out=(hm, hv)); print(f"krige_multi {ids}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
# unpinned outputs
um, uv = np.empty(count), np.empty(count)
t0 = time.perf_counter(); gskrige.krige_multi(whole, [0, 1], out=(um, uv)); print(f"krige_multi [0,1] pageable outputs: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
# two python threads, one context each
import threading
cs = [gskrige.Context(0), gskrige.Context(1)]
def work(i):
    cs[i].krige_into(spec.with_slab(i * half, half), hm[i*half:(i+1)*half], hv[i*half:(i+1)*half])
for rep in range(2):
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]
    print(f"two python threads, two contexts: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
