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
