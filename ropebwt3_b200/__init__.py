"""ropebwt3_b200 -- B200-native BWT merge engine behind the `ropebwt3 build` merge seam.

Host-side mirror of the reference interface for the merge path (fm-index.h:53-66):
the functions keep the reference's names and argument meaning, the rope
(`mrope_t`) is replaced by an opaque device index.  All compute happens in
librb3b200.so (hand-written sm_100a CUDA); nothing here falls back to the CPU.
"""
import numpy as np

from . import capi
from .capi import Rb3bError  # noqa: F401

RB3_ASIZE = 6


def init(device=0):
    capi.check(capi.lib().rb3b_init(device))


def set_param(key, value):
    capi.check(capi.lib().rb3b_set_param(key.encode(), int(value)))


def get_stat(key):
    return int(capi.lib().rb3b_get_stat(key.encode()))


def sync():
    capi.check(capi.lib().rb3b_sync())


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


class Index:
    """Device-resident run-length BWT (replaces mrope_t for the merge path)."""

    def __init__(self):
        self._L = capi.lib()
        self.h = self._L.rb3b_index_create()
        if not self.h:
            raise Rb3bError(-1, self._L.rb3b_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self._L.rb3b_index_destroy(self.h)
            self.h = None

    __del__ = close

    # ---- construction -------------------------------------------------------
    @classmethod
    def from_plain(cls, bwt):
        x = cls()
        b = _u8(bwt)
        capi.check(x._L.rb3b_index_from_plain(x.h, len(b), capi.ptr(b)))
        return x

    @classmethod
    def from_plain_dev(cls, d_bwt, n):
        x = cls()
        capi.check(x._L.rb3b_index_from_plain_dev(x.h, n, capi.ptr(d_bwt)))
        return x

    @classmethod
    def from_runs(cls, sym, ln):
        x = cls()
        s, l = _u8(sym), np.ascontiguousarray(ln, dtype=np.int64)
        assert len(s) == len(l)
        capi.check(x._L.rb3b_index_from_runs(x.h, len(s), capi.ptr(s), capi.ptr(l)))
        return x

    @classmethod
    def restore(cls, fn):
        x = cls()
        capi.check(x._L.rb3b_restore(x.h, fn.encode()))
        return x

    def set_order(self, so):
        """mr_init's sorting order: 0 input order, 1 RLO, 2 RCLO (mrope.h:6-8)."""
        capi.check(self._L.rb3b_index_set_order(self.h, int(so)))

    def get_order(self):
        return int(self._L.rb3b_index_get_order(self.h))

    def insert_multi(self, text):
        """mr_insert_multi (mrope.c:300): insert the 0-terminated strings of `text` in the index's sorting order."""
        t = _u8(text)
        capi.check(self._L.rb3b_insert_multi(self.h, len(t), capi.ptr(t)))

    def reserve(self, n_symbols):
        """Size hint: the index will grow to about n_symbols (avoids regrowing device buffers merge after merge)."""
        capi.check(self._L.rb3b_index_reserve(self.h, int(n_symbols)))

    # ---- the merge path -----------------------------------------------------
    def merge_plain(self, bwt):
        b = _u8(bwt)
        capi.check(self._L.rb3b_merge_plain(self.h, len(b), capi.ptr(b)))

    def merge_plain_dev(self, d_bwt, n):
        capi.check(self._L.rb3b_merge_plain_dev(self.h, n, capi.ptr(d_bwt)))

    def mg_rank_plain(self, bwt):
        b = _u8(bwt)
        rb = np.empty(len(b), np.int64)
        acc = np.zeros(7, np.int64)
        capi.check(self._L.rb3b_mg_rank_plain(self.h, len(b), capi.ptr(b), capi.ptr(rb), capi.ptr(acc)))
        return rb, acc

    def mg_rank_plain_dev(self, d_bwt, n, d_rb):
        acc = np.zeros(7, np.int64)
        capi.check(self._L.rb3b_mg_rank_plain_dev(self.h, n, capi.ptr(d_bwt), capi.ptr(d_rb), capi.ptr(acc)))
        return acc

    # ---- queries ------------------------------------------------------------
    def rank1a(self, k):
        k = np.ascontiguousarray(k, dtype=np.int64)
        ok = np.zeros((len(k), 6), np.int64)
        sym = np.zeros(len(k), np.int8)
        capi.check(self._L.rb3b_rank1a(self.h, len(k), capi.ptr(k), capi.ptr(ok), capi.ptr(sym)))
        return ok, sym

    def rank1a_dev(self, nq, d_k, d_ok, d_sym):
        capi.check(self._L.rb3b_rank1a_dev(self.h, nq, capi.ptr(d_k), capi.ptr(d_ok), capi.ptr(d_sym)))

    def lf_dev(self, nq, d_k, d_c, d_out, variant=0):
        capi.check(self._L.rb3b_lf_dev(self.h, nq, capi.ptr(d_k), capi.ptr(d_c), capi.ptr(d_out), variant))

    def acc(self):
        a = np.zeros(7, np.int64)
        self._L.rb3b_get_acc(self.h, capi.ptr(a))
        return a

    def __len__(self):
        return int(self.acc()[6])

    def nbytes(self):
        return int(self._L.rb3b_index_bytes(self.h))

    # ---- export -------------------------------------------------------------
    def export_runs(self):
        n = capi.check(self._L.rb3b_export_runs(self.h, None, None, 0))
        sym = np.empty(n, np.uint8)
        ln = np.empty(n, np.int64)
        if n:
            capi.check(self._L.rb3b_export_runs(self.h, capi.ptr(sym), capi.ptr(ln), n))
        return sym, ln

    def dump_fmd(self, fn):
        capi.check(self._L.rb3b_dump_fmd(self.h, fn.encode()))

    def dump_fmr(self, fn, max_nodes=64, block_len=512):
        capi.check(self._L.rb3b_dump_fmr(self.h, fn.encode(), max_nodes, block_len))

    def ssa_dump(self, fn, ssa_shift=8):
        """rb3_ssa_gen + rb3_ssa_dump (ssa.c:55-81,198-213)."""
        capi.check(self._L.rb3b_ssa_dump(self.h, int(ssa_shift), fn.encode()))

    def dump_plain(self, fn):
        capi.check(self._L.rb3b_dump_plain(self.h, fn.encode()))


class Batch:
    """A batch prepared for merging: step 0 of the reference's pipeline (rb3_build_sais, build.c:55-83) plus the batch-only
    half of the rank phase -- partial BWT and walk order, device resident."""

    def __init__(self, h):
        self._L = capi.lib()
        if not h:
            raise Rb3bError(-1, self._L.rb3b_last_error().decode())
        self.h = h

    @classmethod
    def prepare(cls, text):
        t = _u8(text)
        return cls(capi.lib().rb3b_batch_prepare(len(t), capi.ptr(t)))

    @classmethod
    def prepare_dev(cls, d_text, n):
        return cls(capi.lib().rb3b_batch_prepare_dev(n, capi.ptr(d_text)))

    def __len__(self):
        return int(self._L.rb3b_batch_len(self.h))

    def bwt(self):
        out = np.empty(len(self), np.uint8)
        capi.check(self._L.rb3b_d2h(capi.ptr(out), self._L.rb3b_batch_bwt_dev(self.h), len(self)))
        return out

    def close(self):
        if getattr(self, "h", None):
            self._L.rb3b_batch_destroy(self.h)
            self.h = None

    __del__ = close


def merge_prepared(index, batch):
    """rb3_fmi_merge_plain (or rb3_enc_plain2fmr for an empty index) on a prepared batch."""
    capi.check(capi.lib().rb3b_merge_prepared(index.h, batch.h))


# ---- the reference's names for the seam (fm-index.h:53-66) ------------------

def rb3_enc_plain2fmr(bwt, max_nodes=0, block_len=0, n_threads=1):
    """fm-index.c:114-137.  Tree geometry and thread count have no meaning for the device index."""
    return Index.from_plain(bwt)


def rb3_fmi_merge_plain(r, bwt, n_threads=1):
    """fm-index.c:279-303: merge the BWT of a new batch into index `r` in place."""
    r.merge_plain(bwt)


def rb3_mg_rank_plain(fa, bwt, n_threads=1):
    """fm-index.c:202-225 -> (rb, acc)."""
    return fa.mg_rank_plain(bwt)


def rb3_build_sais(text):
    """sais-ss.c:50-56: concatenated 0-terminated nt6 strings -> BWT (on the device)."""
    t = _u8(text)
    out = np.empty_like(t)
    capi.check(capi.lib().rb3b_build_bwt(len(t), capi.ptr(t), capi.ptr(out)))
    return out


def build_bwt_so(text, so):
    """BWT of the batch in input (0), RLO (1) or RCLO (2) order (what mr_insert_multi builds on an empty rope)."""
    t = _u8(text)
    out = np.empty_like(t)
    capi.check(capi.lib().rb3b_build_bwt_so(len(t), capi.ptr(t), int(so), capi.ptr(out)))
    return out


def mr_insert_multi(mr, text, is_thr=0):
    """mrope.c:300-385.  `text`: the reads as read (not reversed)."""
    mr.insert_multi(text)


def fmd_image(sym, ln):
    """Host-only: canonical run list -> bytes of the .fmd file (rld0.c:137-243)."""
    import ctypes as C
    s, l = _u8(sym), np.ascontiguousarray(ln, dtype=np.int64)
    p = C.c_void_p()
    n = capi.lib().rb3b_fmd_image(len(s), capi.ptr(s), capi.ptr(l), C.byref(p))
    out = C.string_at(p, n)
    capi.lib().rb3b_host_free(p)
    return out


def fmr_image(sym, ln, max_nodes=64, block_len=512):
    """Host-only: run list -> bytes of a legal .fmr file (mrope.c:152, rope.c:265-287)."""
    import ctypes as C
    s, l = _u8(sym), np.ascontiguousarray(ln, dtype=np.int64)
    p = C.c_void_p()
    n = capi.lib().rb3b_fmr_image(len(s), capi.ptr(s), capi.ptr(l), max_nodes, block_len, C.byref(p))
    out = C.string_at(p, n)
    capi.lib().rb3b_host_free(p)
    return out


def runs_from_image(img):
    """Host-only: bytes of a .fmd / .fmr file -> (sym, len, sorting_order), the canonical run list rb3b_restore loads."""
    import ctypes as C
    a = np.frombuffer(img, np.uint8)
    ps, pl, so = C.c_void_p(), C.c_void_p(), C.c_int(0)
    n = capi.check(capi.lib().rb3b_runs_from_image(capi.ptr(a), len(a), C.byref(ps), C.byref(pl), C.byref(so)))
    sym = np.frombuffer(C.string_at(ps, n), np.uint8).copy()
    ln = np.frombuffer(C.string_at(pl, n * 8), np.int64).copy()
    capi.lib().rb3b_host_free(ps)
    capi.lib().rb3b_host_free(pl)
    return sym, ln, so.value
