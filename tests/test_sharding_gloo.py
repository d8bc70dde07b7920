"""N>1 host logic on CPU: world_size-2 gloo. Each rank computes its slab of a problem with the
oracle standing in for the device call (checker role only), the slabs are gathered with the same
collective bench.py uses, and the result must equal the single-process answer bit for bit."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, T_shape, q):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
    import gskrige
    import oracle_py as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = gskrige.synth.config_spec("C2", grid=T_shape, n=300)
        T = spec.n_targets
        first, count = gskrige.slab_bounds(T, rank, world)
        mean, var = O.krige(spec.with_slab(first, count), nthreads=1)
        gm = gskrige.gather_slabs(torch.from_numpy(mean), T)
        gv = gskrige.gather_slabs(torch.from_numpy(var), T)
        if rank == 0:
            q.put((gm.numpy(), gv.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(40, 30), (37, 29)])   # equal and ragged slabs
def test_two_rank_slabs_equal_single_process(gsk, oracle, shape):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + (7 if shape[0] == 37 else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shape, q)) for r in range(2)]
    for p in procs:
        p.start()
    gm, gv = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spec = gsk.synth.config_spec("C2", grid=shape, n=300)
    mean, var = oracle.krige(spec, nthreads=1)
    assert np.array_equal(gm, mean) and np.array_equal(gv, var)


def test_slab_bounds_cover_domain(gsk):
    for T in (1, 7, 1000, 1 << 20):
        for world in (1, 2, 4, 8):
            spans = [gsk.slab_bounds(T, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == T
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
