"""A plain-Python model of the device algorithm of the rank phase (csrc/rb3b_merge.cu), kernel by kernel, so that the
ALGORITHM -- not only its CUDA implementation -- is pinned to the reference's interleave arrays on a CPU-only box:
walk order from the batch's LF mapping, equal slices walked with a bracket [lo, hi], fix-up from the exact arrival of
the previous slice, sorted-order heads (RLO/RCLO), scatter back to row order.  Small inputs only."""
import numpy as np


def _occ(bwt):
    n = len(bwt)
    occ = np.zeros((6, n + 1), np.int64)
    for c in range(6):
        occ[c, 1:] = np.cumsum(bwt == c)
    acc = np.concatenate([[0], np.cumsum(occ[:, n])]).astype(np.int64)
    return occ, acc


def walk_order(b):
    """k_prep_lf + k_fine_walk/k_list_rank/k_write_walk: the batch in the order the reference's loop meets its rows
    (fm-index.c:165-174), all sequences one after the other.  -> (wrow, wsym, chain_base, chain_len)"""
    occ, acc = _occ(b)
    lf = np.array([acc[c] + occ[c, i] for i, c in enumerate(b)], np.int64)
    wrow, base, length = [], [], []
    for p in range(int(acc[1])):          # chain p starts at sentinel row p
        base.append(len(wrow))
        k = p
        while True:
            wrow.append(k)
            if b[k] == 0:
                break
            k = int(lf[k])
        length.append(len(wrow) - base[-1])
    assert len(wrow) == len(b), "not the BWT of a sentinel-terminated string set"
    wrow = np.array(wrow, np.int64)
    return wrow, b[wrow], np.array(base, np.int64), np.array(length, np.int64), lf


def interleave(a_bwt, b_bwt, seg_len=16, so=0, part=0, n_parts=1, halo=2, warm=0, masks=False):
    """-> ka[len(b)] (or -1 for rows another part resolves) and the number of rows left unresolved.
    warm: rows walked before every slice to narrow its bracket (k_walk_pair).  masks: the walk leaves, for every
    unresolved row whose bracket is at most 127 wide, the 128 bits of the row symbol's plane from `lo` on, and the fix-up
    (k_fix_chain) advances by x' = popcount(mask below x) without looking at the index; rows without a mask take the
    general step.  Both must give what the plain rank chain gives."""
    A, Bv = np.asarray(a_bwt, np.uint8), np.asarray(b_bwt, np.uint8)
    occ, acc = _occ(A)
    nA = len(A)

    def rank(c, k):
        return int(occ[c, min(max(k, 0), nA)])

    wrow, wsym, cbase, clen, _ = walk_order(Bv)
    n = len(Bv)
    UNRES, HEAD, SEED = 1 << 62, 1 << 60, 1 << 59
    kseq = np.zeros(n, np.int64)
    pi = (lambda c: 5 - c if 1 <= c <= 4 else c) if so == 2 else (lambda c: c)
    if so:   # k_so_heads (mr_insert_multi_aux, mrope.c:226-275)
        for p in range(len(cbase)):
            l, u, P, t, ended = 0, int(acc[1]), 0, 0, False
            p0, L = int(cbase[p]), int(clen[p])
            raw = []
            while t < L and l < u:
                c = int(wsym[p0 + t])
                tl = [rank(x, l) for x in range(6)]
                tu = [rank(x, u) for x in range(6)]
                less = sum(tu[x] - tl[x] for x in range(6) if pi(x) < pi(c))
                raw.append(l - P)
                P += less
                t += 1
                if c == 0:
                    ended = True
                    break
                l, u = int(acc[c]) + tl[c], int(acc[c]) + tu[c]
            for s_, r in enumerate(raw):
                kseq[p0 + s_] = (r + P) | HEAD
            if not ended and t < L:
                kseq[p0 + t] = l | SEED
    n_seg = (n + seg_len - 1) // seg_len
    own_lo, own_hi = n_seg * part // n_parts, n_seg * (part + 1) // n_parts
    walk_lo = max(0, own_lo - halo) if n_parts > 1 else 0
    d = np.zeros(n_seg, np.int64)
    arr = [None] * n_seg
    tmask = {}                            # p -> (lo, 128-bit transfer mask) of the tight unresolved rows
    for s in range(walk_lo, own_hi):      # k_walk_first / k_walk_pair
        p0 = s * seg_len
        q0 = p0 if (so or s == 0) else max(0, p0 - warm)
        c0 = 0 if q0 == 0 else int(wsym[q0 - 1])
        if c0 == 0:
            lo = hi = 0 if so else int(acc[1])
        else:
            lo, hi = int(acc[c0]), int(acc[c0 + 1])
        for p in range(q0, p0):           # warm-up rows: results discarded
            c = int(wsym[p])
            if c == 0:
                lo = hi = int(acc[1])
            else:
                lo, hi = int(acc[c]) + rank(c, lo), int(acc[c]) + rank(c, hi)
        for p in range(p0, min(n, p0 + seg_len)):
            c = int(wsym[p])
            f = int(kseq[p]) if so else 0
            if masks and not so and lo != hi and hi - lo <= 127 and c != 0:
                tmask[p] = (lo, sum(1 << i for i in range(128) if lo + i < nA and A[lo + i] == c))
            if f & HEAD:
                kseq[p] = f & ~HEAD
                lo = hi = 0
                continue
            if f & SEED:
                lo = hi = f & ((1 << 42) - 1)
            if lo == hi:
                kseq[p] = lo
            else:
                kseq[p] = lo | UNRES
                d[s] += 1
            if c == 0:
                lo = hi = 0 if so else int(acc[1])
            else:
                lo, hi = int(acc[c]) + rank(c, lo), int(acc[c]) + rank(c, hi)
        arr[s] = (lo, hi)
    for s in range(walk_lo, own_hi - 1):  # k_collect_first + k_walk_fix[_log], cascading
        t, v = s + 1, None
        if arr[s][0] == arr[s][1] and d[t] > 0:
            v = arr[s][0]
        while v is not None:
            p0, cnt, c = t * seg_len, int(d[t]), 1
            full = cnt == min(seg_len, n - p0)
            for p in range(p0, p0 + cnt):
                c = int(wsym[p])
                lo_p = int(kseq[p]) & ((1 << 42) - 1)
                kseq[p] = v
                if c == 0:
                    break
                nv = int(acc[c]) + rank(c, v)
                if p in tmask:            # k_fix_chain: no index access
                    lo_m, m = tmask[p]
                    x = v - lo_p
                    assert lo_m == lo_p and 0 <= x <= 127
                    lo_next = (int(kseq[p + 1]) & ((1 << 42) - 1)) if p + 1 < p0 + cnt else None
                    xn = bin(m & ((1 << x) - 1)).count("1")
                    if lo_next is not None:
                        assert lo_next + xn == nv, "transfer mask disagrees with the rank chain"
                    else:                 # last unresolved row: the next row's low end is where the walk went on from
                        assert int(acc[c]) + rank(c, lo_p) + xn == nv
                v = nv
            d[t] = 0
            if full and c != 0:
                arr[t] = (v, v)
                t += 1
                if not (t < own_hi and d[t] > 0):
                    v = None
            else:
                v = None
    ka = np.full(n, -1, np.int64)         # k_scatter_ka
    unres = 0
    for p in range(own_lo * seg_len, min(n, own_hi * seg_len)):
        if int(kseq[p]) & UNRES:
            unres += 1
        else:
            ka[wrow[p]] = kseq[p]
    return ka, unres


def pack_rb(ka, b_bwt):
    """fm-index.c:168: (ka + i) << 6 | B[i] << 3 | first symbol of suffix i"""
    b = np.asarray(b_bwt, np.uint8)
    acc = np.concatenate([[0], np.cumsum(np.bincount(b, minlength=6)[:6])])
    bucket = np.searchsorted(acc[1:], np.arange(len(b)), side="right")
    return (ka + np.arange(len(b))) << 6 | b.astype(np.int64) << 3 | bucket


def ssa_image(bwt, ss):
    """k_ssa_emit: the sampled suffix array read off the walk order of the index's own BWT (no per-string walk)."""
    import struct
    b = np.asarray(bwt, np.uint8)
    wrow, _, cbase, clen, lf = walk_order(b)
    acc1 = int(np.count_nonzero(b == 0))
    m, n = acc1, len(b)
    ms = 1
    while (1 << ms) < m:
        ms += 1
    n_ssa = (n - m + (1 << ss) - 1) >> ss
    r2i = np.zeros(m, np.uint64)
    ssa = np.zeros(n_ssa, np.uint64)
    chain = np.searchsorted(cbase, np.arange(n), side="right") - 1
    for p in range(n):
        k0 = int(chain[p])
        l, L, row = p - int(cbase[k0]), int(clen[k0]), int(wrow[p])
        if l > 0 and row >= acc1 and ((row - acc1) & ((1 << ss) - 1)) == 0:
            ssa[(row - acc1) >> ss] = ((L - 1 - l) << ms) | k0
        if l == L - 1:
            r2i[int(lf[row])] = k0
    return b"SSA\x01" + struct.pack("<IIqq", ss, ms, m, n_ssa) + r2i.tobytes() + ssa.tobytes()
