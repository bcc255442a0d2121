"""CPU model of the window walk of the decrypt fold (csrc/decrypt.cu, decrypt_count_window_kernel).

The kernel takes the block structure off the load pattern: a warp walks a contiguous run of double blocks in windows of
64 words, a window's verdicts are two 32-bit ballots (even words, odd words), and tables indexed by the step (period L)
say which bits belong to the blocks that end inside the window and what carries into the next one.  This file restates
those tables and the step exactly as the kernel computes them and checks the walk against the predicate of the reference
(src/SecretKey.cpp:131-137: a block is satisfied iff every key bit is set) for every block length class, ragged runs,
runs split over any number of warps, and failing words placed on window boundaries.  The GPU suite checks the kernel
itself against the oracle (tests/test_gpu_fused.py); this model pins the arithmetic on the CPU.
"""
import numpy as np
import pytest

M32 = 0xFFFFFFFF
M64 = (1 << 64) - 1


def bits_below(n):
    return M32 if n >= 32 else (1 << n) - 1


def tables(L):
    """(E, carry table [L], end table [L*E]) -- the kernel's shared-memory image"""
    E = 64 // L + 1
    carry, ends = [None] * L, [None] * (L * E)
    for i in range(L * E):
        s, j = divmod(i, E)
        r = (s * 64) % L                      # words of the open block consumed before window s
        n_ends = (64 + r) // L
        en = (0, 0, M32, 0)                   # no such end in this window: never counted
        if j < n_ends:
            e = (L - r) + j * L
            a = j * L - r if j * L > r else 0
            en = (bits_below((e + 1) >> 1) & ~bits_below((a + 1) >> 1) & M32,
                  bits_below(e >> 1) & ~bits_below(a >> 1) & M32, 0, M32 if j == 0 else 0)
        ends[i] = en
        if j == 0:
            c = (M32, M32, M32, 0)            # no end: the open block stays open
            if n_ends:
                a2 = (L - r) + (n_ends - 1) * L
                c = (~bits_below((a2 + 1) >> 1) & M32, ~bits_below(a2 >> 1) & M32, 0, n_ends)
            carry[s] = c
    return E, carry, ends


def window_count(words, L, mask, n_warps, unroll):
    T = len(words) // L
    n_pairs = T // 2
    E, tab_c, tab_e = tables(L)
    per, extra = divmod(n_pairs, n_warps)     # divided on the host
    total = 0
    for w in range(n_warps):
        pair0 = w * per + min(w, extra)
        my_pairs = per + (1 if w < extra else 0)
        if not my_pairs:
            continue
        nblk, n_units = 2 * my_pairs, my_pairs * L
        steps, full_steps = (n_units + 31) // 32, n_units // 32
        base = pair0 * 2 * L
        state = {"t": [lane % L for lane in range(32)], "sm": 0, "carry": 0, "cnt": 0, "cb": 0}

        def step(s, check):
            flo = fhi = 0
            for lane in range(32):
                unit = s * 32 + lane
                lo, hi = (int(words[base + 2 * unit]), int(words[base + 2 * unit + 1])) if unit < n_units else (0, 0)
                t = state["t"][lane]
                if ~lo & int(mask[(2 * t) % L]) & M64:
                    flo |= 1 << lane
                if ~hi & int(mask[(2 * t + 1) % L]) & M64:
                    fhi |= 1 << lane
            c = tab_c[state["sm"]]
            for lane in range(32):
                en = tab_e[state["sm"] * E + lane] if lane < E else (0, 0, M32, 0)
                bad = (flo & en[0]) | (fhi & en[1]) | en[2] | (state["carry"] & en[3])
                if bad == 0 and (not check or state["cb"] + lane < nblk):
                    state["cnt"] += 1
            if check:
                state["cb"] += c[3]
            state["carry"] = (state["carry"] & c[2]) | (flo & c[0]) | (fhi & c[1])
            state["t"] = [(t + 32 % L) % L for t in state["t"]]
            state["sm"] = 0 if state["sm"] + 1 == L else state["sm"] + 1

        s0 = 0
        while s0 + 2 * unroll <= full_steps:  # the refilling rounds: whole steps only
            for u in range(unroll):
                step(s0 + u, False)
            s0 += unroll
        state["cb"] = (s0 * 64) // L
        while s0 < steps:
            for u in range(unroll):
                if s0 + u < steps:
                    step(s0 + u, True)
            s0 += unroll
        total += state["cnt"]
    if T & 1:
        total += int(np.all((words[(T - 1) * L:] & mask) == mask))
    return total


def direct(words, L, mask):
    T = len(words) // L
    return int(np.sum(np.all((words.reshape(T, L) & mask) == mask, axis=1)))


@pytest.mark.parametrize("L", [3, 5, 9, 15, 17, 19, 31, 33, 34, 50, 63, 65, 97, 100, 129, 193])
def test_window_walk_model_matches_the_block_predicate(L):
    rng = np.random.default_rng(L)
    mask = np.zeros(L, dtype=np.uint64)
    for p in rng.integers(0, 64 * L, 3 + 2 * L):
        mask[p >> 6] |= np.uint64(1) << np.uint64(63 - (p & 63))
    for T in (0, 1, 2, 3, 33, 64, 65, 257, 700):
        w = rng.integers(0, 2**64, T * L, dtype=np.uint64)
        for b in range(T):
            if rng.random() < 0.8:
                w[b * L:(b + 1) * L] |= mask
                if rng.random() < 0.4:          # exactly one failing word: first, last, or anywhere
                    k = (0, L - 1, int(rng.integers(0, L)))[int(rng.integers(0, 3))]
                    w[b * L + k] &= ~mask[k]
        want = direct(w, L, mask)
        for n_warps, unroll in ((1, 8), (3, 6), (8, 12)):
            assert window_count(w, L, mask, n_warps, unroll) == want, (L, T, n_warps, unroll)


def test_window_tables_cover_every_block_end_exactly_once():
    """Over one period (L windows = 64 blocks) every block ends in exactly one window, and the masks of the blocks of a
    window plus the carry mask tile its 64 words."""
    for L in (3, 7, 19, 33, 64, 65, 200, 999):
        E, tab_c, tab_e = tables(L)
        assert sum(c[3] for c in tab_c) == 64
        for s in range(L):
            lo = hi = 0
            n_ends = tab_c[s][3]
            for j in range(E):
                en = tab_e[s * E + j]
                if j < n_ends:
                    assert en[2] == 0 and (lo & en[0]) == 0 and (hi & en[1]) == 0
                    lo |= en[0]
                    hi |= en[1]
                else:
                    assert en[2] == M32
            if n_ends:
                assert (lo & tab_c[s][0]) == 0 and (hi & tab_c[s][1]) == 0
                assert (lo | tab_c[s][0]) == M32 and (hi | tab_c[s][1]) == M32
            else:
                assert tab_c[s][:3] == (M32, M32, M32)
