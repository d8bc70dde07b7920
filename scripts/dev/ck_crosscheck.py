"""Compact-key search against the exact-key search at full size (development check, needs a -DGSK_DEV_TUNABLES build):
    GSK_NO_COMPACT_KEYS=1 python scripts/dev/ck_crosscheck.py write /tmp/exact
    python scripts/dev/ck_crosscheck.py write /tmp/ck
    python scripts/dev/ck_crosscheck.py compare /tmp/exact /tmp/ck
Every target's neighbour list (order included) is hashed on the device; natural ties in the truncated distance occur at
~4e-7 per target at n = 1e6, so a 16.7M-target slab holds a handful of them — the redo pass must make the two runs equal."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import gskrige as gsk  # noqa: E402

CASES = [("C5", 1 << 24), ("C3a", 1 << 24), ("C2", 1_000_000)]


def hashes(name, count):
    spec = gsk.synth.config_spec(name)
    T = spec.n_targets
    count = min(count, T)
    first = (T - count) // 2
    k = spec.params["max_neighbors"]
    ctx = gsk.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.plan(spec)
    dm = torch.empty(count, dtype=torch.float64, device="cuda")
    dv = torch.empty(count, dtype=torch.float64, device="cuda")
    nn = torch.empty(count, dtype=torch.int32, device="cuda")
    idx = torch.empty((count, k), dtype=torch.int32, device="cuda")
    ctx.execute(first, count, dm.data_ptr(), dv.data_ptr(), nn.data_ptr(), idx.data_ptr())
    torch.cuda.synchronize()
    h = torch.zeros(count, dtype=torch.int64, device="cuda")
    for j in range(k):  # order-sensitive polynomial hash (wraps modulo 2^64)
        h = h * 1000003 + idx[:, j].to(torch.int64) + 7
    h = h * 31 + nn.to(torch.int64)
    ctx.close()
    return h.cpu().numpy(), dm.cpu().numpy(), dv.cpu().numpy()


if sys.argv[1] == "write":
    for name, count in CASES:
        h, m, v = hashes(name, count)
        np.save(f"{sys.argv[2]}_{name}_h.npy", h)
        np.save(f"{sys.argv[2]}_{name}_m.npy", m)
        print(name, "targets", h.size, "hash of hashes", int(np.bitwise_xor.reduce(h)), flush=True)
else:
    ok = True
    for name, _ in CASES:
        a = np.load(f"{sys.argv[2]}_{name}_h.npy"); b = np.load(f"{sys.argv[3]}_{name}_h.npy")
        ma = np.load(f"{sys.argv[2]}_{name}_m.npy"); mb = np.load(f"{sys.argv[3]}_{name}_m.npy")
        nd = int((a != b).sum())
        print(f"{name}: {a.size} targets, neighbour lists differing: {nd}, means bitwise equal: {bool(np.array_equal(ma, mb))}")
        ok = ok and nd == 0
    print("CROSSCHECK", "OK" if ok else "FAILED")
