"""Slot table of local_solve_wpt.cuh: interval colouring of the 8x8 blocks (I, J) of the packed factor (block (I, J),
J <= min(I, 7), I <= 8, lives over the steps [J, I]; I = 8 — the extra rows — lives to the end) with the side
condition parity(slot(I, J)) == (I + p_J) mod 2, so that vertically adjacent blocks of a column sit in slots of
different parity (the kernel XORs the in-block position with 8·parity: two row blocks read by one half-warp then
fall into different halves of the 32 banks)."""
import itertools
import sys

NBM = int(sys.argv[1]) if len(sys.argv) > 1 else 8     # neighbour column blocks: 8 (one warp per target), 4 (half a warp)
blocks = [(I, J) for J in range(NBM) for I in range(J, NBM + 1)]
life = {(I, J): (J, I if I < NBM else NBM) for (I, J) in blocks}

def colour(pvec, nmax=64):
    slot, owner_end = {}, {}          # owner_end[s] = last step the slot is busy
    n = [0, 0]                        # slots allocated per parity: even ids 0,2,4…, odd ids 1,3,5…
    for step in range(NBM + 1):
        for b in [b for b in blocks if life[b][0] == step]:
            par = (b[0] + pvec[b[1]]) & 1
            free = sorted(s for s, e in owner_end.items() if e < step and (s & 1) == par)
            if free:
                s = free[0]
            else:
                s = 2 * n[par] + par
                n[par] += 1
            slot[b] = s
            owner_end[s] = life[b][1]
    return slot, max(slot.values()) + 1

best = None
for pvec in itertools.product((0, 1), repeat=NBM):
    slot, n = colour(pvec)
    if best is None or n < best[1]:
        best = (slot, n, pvec)
slot, n, pvec = best
print("slots (max id + 1):", n, "p_J:", pvec)
for step in range(NBM + 1):
    live = [b for b in blocks if life[b][0] <= step <= life[b][1]]
    assert len({slot[b] for b in live}) == len(live)
for (I, J) in blocks:
    if (I + 1, J) in slot:
        assert (slot[(I, J)] ^ slot[(I + 1, J)]) & 1
print("{" + ",\n ".join("{" + ", ".join(str(slot.get((I, J), 0)) for J in range(NBM)) + "}" for I in range(NBM + 1)) + "}")
