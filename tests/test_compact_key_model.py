"""The compact-key selection rule of the search kernel (csrc/search.cu, template parameter CK), restated in Python and
checked against the exact (d², index) order on adversarial candidate streams: many candidates whose d² differ only in
the mantissa bits the key drops, exact ties, streams shorter than k.

The kernel keeps, per target, the k smallest 64-bit keys  (bits(d²) & ~LOW) | index  and claims: if the thread does NOT
flag its tile, its list equals the exact top-k in exact order; a flagged tile is searched again with exact keys. This
file checks the first half of that claim (the second half — the redo pass — is a GPU test:
tests/test_gpu_parity.py::test_distances_that_differ_only_in_the_dropped_key_bits)."""
import struct

import numpy as np
import pytest


def bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def from_bits(b: int) -> float:
    return struct.unpack("<d", struct.pack("<Q", b))[0]


def compact_select(d2s, idxs, k, keybits):
    """One thread of search_kernel<…, CK = true>: returns (list of indices in key order, flagged)."""
    low = (1 << keybits) - 1
    inv = ~low & 0xFFFFFFFFFFFFFFFF
    keys = []
    worst = float("inf")      # largest d² sharing the k-th best's distance bits, once the list is full
    wkey = 0xFFFFFFFFFFFFFFFF
    lowbits = 0
    amb = 0xFFFFFFFFFFFFFFFF
    for d2, oi in zip(d2s, idxs):
        if d2 > worst:
            continue
        b = bits(d2)
        kc = (b & inv) | oi
        lowbits |= b
        if len(keys) == k:
            if kc > wkey:
                amb = wkey & inv
                continue
            keys.remove(max(keys))
            keys.append(kc)
            nw = max(keys)
            if (nw ^ wkey) & inv == 0:
                amb = nw & inv
            wkey = nw
            worst = from_bits(wkey | low)
        else:
            keys.append(kc)
            if len(keys) == k:
                wkey = max(keys)
                worst = from_bits(wkey | low)
    keys.sort()
    tie = any(((a ^ b) & inv) == 0 for a, b in zip(keys[:-1], keys[1:]))
    flagged = (lowbits & low) != 0 and (tie or (len(keys) == k and amb == (wkey & inv)))
    return [kk & low for kk in keys], flagged


def exact_select(d2s, idxs, k):
    order = sorted(zip(d2s, idxs))
    return [i for _, i in order[:k]]


def adversarial_stream(rng, n, keybits, exact_tie_rate):
    """d² values clustered in a few buckets of the truncated key, differing in the dropped bits (or not at all)."""
    low = (1 << keybits) - 1
    nb = max(2, n // 3)
    bases = [bits(float(v)) & ~low for v in rng.uniform(0.5, 4.0, nb)]
    d2s = []
    for _ in range(n):
        b = bases[rng.integers(nb)]
        if rng.random() < exact_tie_rate:
            d2s.append(from_bits(b))                       # dropped bits all zero: an exact tie inside its bucket
        else:
            d2s.append(from_bits(b | int(rng.integers(0, low + 1))))
    idxs = rng.permutation(1 << keybits)[:n].tolist()
    return d2s, idxs


@pytest.mark.parametrize("keybits", [4, 7, 12])
def test_unflagged_lists_are_exact(keybits):
    rng = np.random.default_rng(100 + keybits)
    unflagged = flagged = 0
    for trial in range(600):
        n = int(rng.integers(1, min(60, 1 << keybits)))
        k = int(rng.integers(1, 12))
        d2s, idxs = adversarial_stream(rng, n, keybits, exact_tie_rate=float(rng.choice([0.0, 0.3, 1.0])))
        got, flag = compact_select(d2s, idxs, k, keybits)
        if flag:
            flagged += 1
            continue
        unflagged += 1
        assert got == exact_select(d2s, idxs, k), (trial, n, k)
    assert unflagged > 50 and flagged > 50      # both outcomes occur on these streams


def test_exactly_representable_distances_never_flag():
    """Lattice data whose d² have no dropped bits: ties fall to the lower index and nothing is flagged."""
    rng = np.random.default_rng(7)
    for _ in range(200):
        n, k, keybits = 40, int(rng.integers(1, 10)), 8
        d2s = [float(v) * 0.25 for v in rng.integers(1, 12, n)]     # multiples of 1/4: low mantissa bits are zero
        idxs = rng.permutation(256)[:n].tolist()
        got, flag = compact_select(d2s, idxs, k, keybits)
        assert not flag and got == exact_select(d2s, idxs, k)


def test_well_separated_distances_never_flag():
    rng = np.random.default_rng(8)
    for _ in range(200):
        n, k, keybits = 50, int(rng.integers(1, 20)), 10
        d2s = rng.uniform(0.0, 100.0, n).tolist()                   # 42 kept mantissa bits: no two share a bucket
        idxs = rng.permutation(1024)[:n].tolist()
        got, flag = compact_select(d2s, idxs, k, keybits)
        assert not flag and got == exact_select(d2s, idxs, k)
