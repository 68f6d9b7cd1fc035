"""Brute-force model check (pure Python, no GPU) of the two incremental rules the fused rollouts use:

* binary (csrc/pcgrl_device.cuh binary_stats_update): regions / longest path after a single-cell edit from the pieces next to the
  cell, the carried set of known-best cells and the size bound; falls back to a full re-measurement only when it must;
* zelda (local_piece_count): how many regions a toggled cell touches, decided from its 3x3 window, or "undecided".

The CUDA code is a transcription of `update` / `local` below; the GPU parity tests check the kernels, this file checks the rules
themselves against a from-scratch computation on every edit (calc_num_regions / calc_longest_path semantics of
gym_pcgrl/envs/helper.py:197-264: double sweep from the row-major-first cell, np.argmax tie-break)."""
import random
from collections import deque

NB = ((-1, 0), (1, 0), (0, -1), (0, 1))


def comps(P):
    seen, out = set(), []
    for c in sorted(P):
        if c in seen:
            continue
        q, comp = deque([c]), {c}
        seen.add(c)
        while q:
            y, x = q.popleft()
            for dy, dx in NB:
                d = (y + dy, x + dx)
                if d in P and d not in seen:
                    seen.add(d); comp.add(d); q.append(d)
        out.append(comp)
    return out


def bfs(s, comp):
    dist, q = {s: 0}, deque([s])
    while q:
        y, x = q.popleft()
        for dy, dx in NB:
            d = (y + dy, x + dx)
            if d in comp and d not in dist:
                dist[d] = dist[(y, x)] + 1; q.append(d)
    m = max(dist.values())
    return m, min(c for c in dist if dist[c] == m), set(dist)


def value(comp):
    d1, far, _ = bfs(min(comp), comp)
    return d1, bfs(far, comp)[0]


def full(P):
    """regions_and_longest_path incl. its shortcuts; bm = cells of components known to hold the maximum (a subset)"""
    cs = comps(P)
    iso, dom = [c for c in cs if len(c) == 1], [c for c in cs if len(c) == 2]
    R, B = len(iso) + len(dom), (1 if dom else 0)
    bm = set().union(*dom) if dom else (set().union(*iso) if iso else set())
    for c in cs:
        if len(c) < 3:
            continue
        R += 1
        d1, d2 = value(c)
        if 2 * d1 > B:
            if d2 > B:
                B, bm = d2, set(c)
            elif d2 == B:
                bm |= c
    return R, B, bm


def update(P, c, grew, R, B, bm, stats):
    """binary_stats_update"""
    P0 = P - {c}
    y, x = c
    nbs = {(y + dy, x + dx) for dy, dx in NB if (y + dy, x + dx) in P0}
    pieces, touched = [], set()
    while nbs:
        p = bfs(min(nbs), P0)[2]
        pieces.append(p); nbs -= p; touched |= p
    m = len(pieces)
    R += (1 - m) if grew else (m - 1)
    largest_old = max([len(p) for p in pieces], default=0) if grew else len(touched) + 1
    keep = bm - (touched if grew else touched | {c})
    rest_is_best = bool(keep) or largest_old - 1 < B
    cur, cc = (B if rest_is_best else B - 1), set()
    for comp in ([touched | {c}] if grew else pieces):
        if len(comp) - 1 > cur:
            d1, d2 = value(comp)
            if 2 * d1 > cur:
                if d2 > cur:
                    cur, cc = d2, set(comp)
                elif d2 == cur:
                    cc |= comp
    if cur >= B:
        bm = (keep if cur == B else set()) | cc
        B = cur
        stats[0] += 1
    else:
        R2, B, bm = full(P)
        assert R2 == R
        stats[1] += 1
    return R, B, bm


def local(P0, c):
    """local_piece_count"""
    y, x = c
    g = lambda dy, dx: 1 if (y + dy, x + dx) in P0 else 0
    N, S, W, E, NW, NE, SW, SE = g(-1, 0), g(1, 0), g(0, -1), g(0, 1), g(-1, -1), g(-1, 1), g(1, -1), g(1, 1)
    k = N + S + W + E
    if k <= 1:
        return k
    links = (N & NE & E) + (E & SE & S) + (S & SW & W) + (W & NW & N)
    return 1 if k - links <= 1 else -1


def truth(P):
    cs = comps(P)
    return len(cs), max([value(c)[1] for c in cs], default=0)


def test_incremental_binary_rule_equals_brute_force():
    rnd = random.Random(11)
    stats = [0, 0]
    for trial in range(60):
        H, W, p = rnd.choice([1, 2, 3, 5, 8, 16]), rnd.choice([1, 2, 3, 7, 16]), rnd.choice([0.2, 0.5, 0.8, 0.95])
        P = {(y, x) for y in range(H) for x in range(W) if rnd.random() < p}
        R, B, bm = full(P)
        assert (R, B) == truth(P)
        for step in range(120):
            if rnd.random() < 0.1:
                bm = set()          # a new launch starts with nothing known
            c = (rnd.randrange(H), rnd.randrange(W))
            grew = c not in P
            P.add(c) if grew else P.remove(c)
            R, B, bm = update(P, c, grew, R, B, bm, stats)
            assert (R, B) == truth(P), (trial, step)
            for comp in comps(P):
                if comp & bm:
                    assert value(comp)[1] == B      # invariant: bm only holds cells of components that reach the maximum
    assert stats[0] > 2 * stats[1]                  # the fallback is the exception (tiny maps included: 16x16 alone has 6-10 %)


def test_zelda_window_rule_equals_brute_force():
    rnd = random.Random(5)
    decided = total = 0
    for trial in range(120):
        H, W, p = rnd.choice([1, 2, 3, 5, 8, 16]), rnd.choice([1, 2, 3, 7, 16]), rnd.choice([0.3, 0.5, 0.68, 0.9])
        P = {(y, x) for y in range(H) for x in range(W) if rnd.random() < p}
        for step in range(60):
            c = (rnd.randrange(H), rnd.randrange(W))
            R0, grew = len(comps(P)), c not in P
            m = local(P - {c}, c)
            P.add(c) if grew else P.remove(c)
            total += 1
            if m >= 0:
                decided += 1
                assert len(comps(P)) == R0 + ((1 - m) if grew else (m - 1)), (trial, step)
    assert decided > total // 2
