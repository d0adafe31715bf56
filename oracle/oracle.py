"""ctypes front-end of oracle/librb3oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module (see the header of rb3_oracle.c).
All arrays are numpy; a BWT is a run list (sym uint8[n], len int64[n]).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build():
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "librb3oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.ora_plain2runs.restype = C.c_int64
        L.ora_plain2runs.argtypes = [C.c_int64, u8p, C.c_void_p, C.c_void_p]
        L.ora_rank1a.argtypes = [C.c_int64, u8p, i64p, C.c_int64, i64p, i64p, i8p]
        L.ora_build_bwt.argtypes = [C.c_int64, u8p, C.c_void_p]
        L.ora_mg_rank_plain.argtypes = [C.c_int64, u8p, i64p, C.c_int64, u8p, i64p, i64p]
        L.ora_merge_runs.restype = C.c_int64
        L.ora_merge_runs.argtypes = [C.c_int64, u8p, i64p, C.c_int64, i64p, C.c_void_p, C.c_void_p]
        L.ora_fmd_encode.restype = C.c_int64
        L.ora_fmd_encode.argtypes = [C.c_int64, u8p, i64p, C.POINTER(C.c_void_p)]
        L.ora_fmd_decode.restype = C.c_int64
        L.ora_fmd_decode.argtypes = [C.c_int64, u8p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_fmr_decode.restype = C.c_int64
        L.ora_fmr_decode.argtypes = [C.c_int64, u8p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_fmr_encode.restype = C.c_int64
        L.ora_fmr_encode.argtypes = [C.c_int64, u8p, i64p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.ora_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def coalesce(sym, ln):
    """Canonical run list: drop empty runs, merge equal neighbours."""
    sym = np.asarray(sym, np.uint8)
    ln = np.asarray(ln, np.int64)
    keep = ln > 0
    sym, ln = sym[keep], ln[keep]
    if len(sym) == 0:
        return sym, ln
    head = np.ones(len(sym), bool)
    head[1:] = sym[1:] != sym[:-1]
    idx = np.flatnonzero(head)
    return sym[idx].copy(), np.add.reduceat(ln, idx).astype(np.int64)


def plain2runs(bwt):
    bwt = _c(bwt, np.uint8)
    n = lib().ora_plain2runs(len(bwt), bwt, None, None)
    sym = np.empty(n, np.uint8)
    ln = np.empty(n, np.int64)
    if n:
        lib().ora_plain2runs(len(bwt), bwt, sym.ctypes.data, ln.ctypes.data)
    return sym, ln


def runs2plain(sym, ln):
    return np.repeat(np.asarray(sym, np.uint8), np.asarray(ln, np.int64))


def rank1a(sym, ln, k):
    """-> (ok int64[nq,6], ret int8[nq]) with ret = B[k] or -1 (mr_rank1a / rld_rank1a contract)."""
    sym, ln, k = _c(sym, np.uint8), _c(ln, np.int64), _c(k, np.int64)
    ok = np.zeros((len(k), 6), np.int64)
    ret = np.zeros(len(k), np.int8)
    lib().ora_rank1a(len(sym), sym, ln, len(k), k, ok, ret)
    return ok, ret


def build_bwt(text, want_sa=False):
    """text: uint8 nt6 codes, every sequence 0-terminated -> BWT (sais-ss.c semantics)."""
    t = _c(text, np.uint8).copy()
    sa = np.empty(len(t), np.int64) if want_sa else None
    rc = lib().ora_build_bwt(len(t), t, sa.ctypes.data if want_sa else None)
    if rc != 0:
        raise ValueError("text must be non-empty and end with a sentinel")
    return (t, sa) if want_sa else t


def mg_rank_plain(asym, alen, seq):
    """-> (rb int64[len], acc int64[7]); rb packs (ka+i)<<6 | B[i]<<3 | bucket (fm-index.c:168)."""
    asym, alen, seq = _c(asym, np.uint8), _c(alen, np.int64), _c(seq, np.uint8)
    rb = np.zeros(len(seq), np.int64)
    acc = np.zeros(7, np.int64)
    if lib().ora_mg_rank_plain(len(asym), asym, alen, len(seq), seq, rb, acc) != 0:
        raise ValueError("symbol >= 6 in the batch BWT")
    return rb, acc


def merge_runs(asym, alen, rb):
    asym, alen, rb = _c(asym, np.uint8), _c(alen, np.int64), _c(rb, np.int64)
    n = lib().ora_merge_runs(len(asym), asym, alen, len(rb), rb, None, None)
    sym = np.empty(n, np.uint8)
    ln = np.empty(n, np.int64)
    lib().ora_merge_runs(len(asym), asym, alen, len(rb), rb, sym.ctypes.data, ln.ctypes.data)
    return sym, ln


def merge_plain(asym, alen, seq):
    """rb3_fmi_merge_plain on run lists: A (runs) + batch BWT -> merged canonical runs."""
    rb, _ = mg_rank_plain(asym, alen, seq)
    return merge_runs(asym, alen, rb)


def fmd_encode(sym, ln):
    sym, ln = _c(sym, np.uint8), _c(ln, np.int64)
    p = C.c_void_p()
    n = lib().ora_fmd_encode(len(sym), sym, ln, C.byref(p))
    out = C.string_at(p, n)
    lib().ora_free(p)
    return out


def fmd_decode(img):
    a = np.frombuffer(img, np.uint8)
    n = lib().ora_fmd_decode(len(a), a, None, None, None)
    if n < 0:
        raise ValueError("not an FMD image")
    sym = np.empty(n, np.uint8)
    ln = np.empty(n, np.int64)
    mc = np.zeros(6, np.int64)
    lib().ora_fmd_decode(len(a), a, sym.ctypes.data, ln.ctypes.data, mc.ctypes.data)
    return sym, ln, mc


def fmr_decode(img):
    """-> (sym, len, rope_cnt int64[6,6], (so, max_nodes, block_len)); runs as stored, not coalesced."""
    a = np.frombuffer(img, np.uint8)
    n = lib().ora_fmr_decode(len(a), a, None, None, None, None)
    if n < 0:
        raise ValueError("bad FMR image (code %d)" % n)
    sym = np.empty(n, np.uint8)
    ln = np.empty(n, np.int64)
    rc = np.zeros((6, 6), np.int64)
    geom = np.zeros(3, np.int32)
    lib().ora_fmr_decode(len(a), a, sym.ctypes.data, ln.ctypes.data, rc.ctypes.data, geom.ctypes.data)
    return sym, ln, rc, tuple(int(x) for x in geom)


def fmr_encode(sym, ln, max_nodes=64, block_len=512, so=0):
    sym, ln = _c(sym, np.uint8), _c(ln, np.int64)
    p = C.c_void_p()
    n = lib().ora_fmr_encode(len(sym), sym, ln, max_nodes, block_len, so, C.byref(p))
    out = C.string_at(p, n)
    lib().ora_free(p)
    return out


# ---------------------------------------------------------------- text helpers
NT6 = np.full(256, 5, np.uint8)
for _ch, _v in zip("ACGTacgt", [1, 2, 3, 4, 1, 2, 3, 4]):
    NT6[ord(_ch)] = _v
NT6[0:5] = [0, 1, 2, 3, 4]  # io.c:12-21 maps raw bytes 0..4 to themselves


def encode_batch(seqs, fwd=True, rev=True):
    """List of ASCII strings -> one batch text (io.c:84-102: forward then revcomp, each 0-terminated)."""
    parts = []
    for s in seqs:
        x = NT6[np.frombuffer(s.encode() if isinstance(s, str) else s, np.uint8)]
        if fwd:
            parts += [x, np.zeros(1, np.uint8)]
        if rev:
            r = x[::-1].copy()
            m = (r >= 1) & (r <= 4)
            r[m] = 5 - r[m]
            parts += [r, np.zeros(1, np.uint8)]
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def to_ascii(bwt):
    return np.frombuffer(b"$ACGTN", np.uint8)[np.asarray(bwt, np.uint8)].tobytes().decode()


# ---------------------------------------------------------------- ropebwt2 insertion (build -2 / -s / -r)
def _split_strings(text):
    t = np.asarray(text, np.uint8)
    z = np.flatnonzero(t == 0)
    out, s = [], 0
    for e in z:
        out.append([int(x) for x in t[s:e]])
        s = int(e) + 1
    return out


def _comp6(c):
    return 5 - c if 1 <= c <= 4 else c  # rope_comp6, mrope.c:224


def insert_multi(ropes, text, so):
    """mr_insert_multi + mr_insert_multi_aux (mrope.c:226-385) restated on six plain Python lists: ropes[b] holds the
    BWT symbols of the rows whose suffix starts with symbol b (mrope.h:10-14).  `text`: the reads as read, 0-terminated;
    they are reversed here as build.c:216 (rb3_reverse_all) does.  so: 0 input order, 1 RLO, 2 RCLO.  Pure Python, small
    cases only.  Returns ropes (modified in place)."""
    strs = [s[::-1] + [0] for s in _split_strings(text)]  # reversed, NUL-terminated (mrope.c:310-319)
    m = len(strs)
    is_srt, is_comp = so != 0, so == 2
    n0 = sum(r.count(0) for r in ropes)  # mrope.c:321
    # triple64_t (mrope.c:216-220): [l, u, c, string, read offset]
    prev = [[0, n0, 0, s, 0] if is_srt else [n0 + k, n0 + k, 0, s, 0] for k, s in enumerate(strs)]  # mrope.c:322-326

    def insert_run(rope, x, b, n):  # rope_insert_run (rope.c:114): returns rank(b, x)
        r = rope[:x].count(b)
        rope[x:x] = [b] * n
        return r

    def aux(rope, a):  # mr_insert_multi_aux, mrope.c:226-275
        for t in a:
            t[2] = t[3][t[4]]
            t[4] += 1
        beg = 0
        for k in range(1, len(a) + 1):
            if k == len(a) or a[k][1] != a[k - 1][1]:
                l, u = a[beg][0], a[beg][1]
                if l == u and k == beg + 1:
                    a[beg][0] = a[beg][1] = insert_run(rope, l, a[beg][2], 1)
                    beg = k
                    continue
                tl = [rope[:l].count(b) for b in range(6)] if l != u else [0] * 6
                tu = [rope[:u].count(b) for b in range(6)] if l != u else [0] * 6
                c = [0] * 6
                for i in range(beg, k):
                    c[a[i][2]] += 1
                if c[0]:
                    insert_run(rope, l, 0, c[0])
                x = l + c[0] + (tu[0] - tl[0])
                for b in ([4, 3, 2, 1] if is_comp else [1, 2, 3, 4]):  # mrope.c:249-260
                    size = tu[b] - tl[b]
                    if c[b]:
                        tl[b] = insert_run(rope, x, b, c[b])
                        tu[b] = tl[b] + size
                    x += c[b] + size
                if c[5]:
                    tu[5] -= tl[5]
                    tl[5] = insert_run(rope, x, 5, c[5])
                    tu[5] += tl[5]
                for i in range(beg, k):
                    a[i][0], a[i][1] = tl[a[i][2]], tu[a[i][2]]
                beg = k

    aux(ropes[0], prev)  # the first (actually the last) column, mrope.c:327
    live = prev
    while live:
        buckets = [[t for t in live if t[2] == b] for b in range(6)]  # stable counting sort, mrope.c:343-349
        for b in range(1, 6):
            if buckets[b]:
                aux(ropes[b], buckets[b])
        ac = [0] * 6
        for b in range(1, 6):  # mrope.c:372-380
            for a in range(6):
                ac[a] += ropes[b - 1].count(a)
            for t in buckets[b]:
                t[0] += ac[t[2]]
                t[1] += ac[t[2]]
        live = [t for b in range(1, 6) for t in buckets[b]]
    return ropes


def rb2_bwt(texts, so):
    """BWT (uint8 array) after inserting the batches `texts` one after the other with insert_multi."""
    ropes = [[] for _ in range(6)]
    for t in texts:
        insert_multi(ropes, t, so)
    return np.array([c for r in ropes for c in r], np.uint8)


def sorted_bwt(text, so):
    """The same BWT by its closed form: suffix t of string S sorts by S[t:] + '$' and then by the symbols that precede
    it, S[t-1], S[t-2], ..., under the RLO ($ACGTN) or RCLO ($TGCAN) order with the start of the string smallest; in
    input order (so == 0) equal suffixes sort by string number (sais-ss.c)."""
    strs = _split_strings(text)
    keys = []
    for i, s in enumerate(strs):
        for t in range(len(s) + 1):
            if so == 0:
                tb = (i,)
            else:
                tb = tuple((_comp6(c) if so == 2 else c) + 1 for c in s[:t][::-1]) + (0,)
            keys.append((tuple(s[t:]) + (0,) + tb, s[t - 1] if t > 0 else 0))
    keys.sort(key=lambda kv: kv[0])
    return np.array([kv[1] for kv in keys], np.uint8)


# ---------------------------------------------------------------- sampled suffix array (ropebwt3 ssa)
def ssa_image(bwt, ss):
    """rb3_ssa_gen + ssa_gen1 + rb3_ssa_dump (ssa.c:17-81,198-213) restated on a plain BWT: bytes of the .ssa file."""
    b = np.asarray(bwt, np.uint8)
    n = len(b)
    cnt = np.bincount(b, minlength=6)[:6]
    acc = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    occ = np.zeros((6, n + 1), np.int64)  # occ[c][k] = #c in b[0:k]
    for c in range(6):
        occ[c, 1:] = np.cumsum(b == c)
    m = int(acc[1])
    ms = 1
    while (1 << ms) < m:
        ms += 1
    n_ssa = (n - m + (1 << ss) - 1) >> ss
    r2i = np.zeros(m, np.uint64)
    ssa = np.zeros(n_ssa, np.uint64)
    mask = (1 << ss) - 1
    for k0 in range(m):  # ssa_gen1, ssa.c:17-41
        k, l, buf = k0, 0, []
        while True:
            l += 1
            c = int(b[k])
            k = int(acc[c] + occ[c, k])
            if c:
                if ((k - m) & mask) == 0:
                    x = (k - m) >> ss
                    ssa[x] = l
                    buf.append(x)
            else:
                r2i[k] = k0
                break
        for x in buf:
            ssa[x] = ((l - 1 - int(ssa[x])) << ms) | k0
    import struct
    return b"SSA\x01" + struct.pack("<IIqq", ss, ms, m, n_ssa) + r2i.tobytes() + ssa.tobytes()
