"""Driver for the UNMODIFIED reference compiled into oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

oracle/_ref/ropebwt3      the reference CLI (subprocess)
oracle/_ref/librb3ref.so  the same objects as a shared library (ctypes)

Built by `make -C oracle ref` from the sources where they lie under
/root/reference (never copied into this repo).  oracle/_ref/ is git-ignored
but travels to the GPU box with the snapshot, so these helpers also work there.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(_HERE, "_ref", "ropebwt3")
SO = os.path.join(_HERE, "_ref", "librb3ref.so")
_LIB = None
_LIBC = None


def available():
    return os.path.exists(BIN) and os.path.exists(SO)


def run(args, stdin=None, check=True):
    """Run `ropebwt3 <args>`; returns stdout bytes."""
    p = subprocess.run([BIN] + [str(a) for a in args], input=stdin, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, check=False)
    if check and p.returncode != 0:
        raise RuntimeError("ropebwt3 %s failed: %s" % (args, p.stderr.decode()[-500:]))
    return p.stdout


class _Fmi(C.Structure):  # rb3_fmi_t, fm-index.h:42-49
    _fields_ = [("is_fmd", C.c_int32), ("e", C.c_void_p), ("r", C.c_void_p),
                ("ssa", C.c_void_p), ("sid", C.c_void_p), ("acc", C.c_int64 * 7)]


def lib():
    global _LIB, _LIBC
    if _LIB is None:
        L = C.CDLL(SO)
        C.c_int.in_dll(L, "rb3_verbose").value = 1
        L.rb3_build_sais.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        L.rb3_build_sais.restype = None
        L.rb3_enc_plain2fmr.restype = C.c_void_p
        L.rb3_enc_plain2fmr.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int32]
        L.rb3_fmi_merge_plain.restype = None
        L.rb3_fmi_merge_plain.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.rb3_mg_rank_plain.restype = None
        L.rb3_mg_rank_plain.argtypes = [C.POINTER(_Fmi), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.rb3_fmi_get_acc.restype = C.c_int64
        L.rb3_fmi_get_acc.argtypes = [C.POINTER(_Fmi), C.c_void_p]
        L.mr_rank2a.restype = C.c_int
        L.mr_rank2a.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.mr_dump.argtypes = [C.c_void_p, C.c_void_p]
        L.mr_dump.restype = None
        L.mr_destroy.argtypes = [C.c_void_p]
        L.mr_destroy.restype = None
        L.mr_restore_file.restype = C.c_void_p
        L.mr_restore_file.argtypes = [C.c_char_p]
        L.rb3_enc_fmr2fmd.restype = C.c_void_p
        L.rb3_enc_fmr2fmd.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.rld_dump.argtypes = [C.c_void_p, C.c_char_p]
        L.rld_destroy.argtypes = [C.c_void_p]
        L.rld_destroy.restype = None
        L.rld_restore.restype = C.c_void_p
        L.rld_restore.argtypes = [C.c_char_p]
        L.rld_rank1a.restype = C.c_int
        L.rld_rank1a.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        _LIBC = C.CDLL(None)
        _LIBC.fopen.restype = C.c_void_p
        _LIBC.fopen.argtypes = [C.c_char_p, C.c_char_p]
        _LIBC.fclose.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def build_sais(text, n_seq, n_threads=1):
    """rb3_build_sais (sais-ss.c:50): text (0-terminated nt6 strings) -> BWT."""
    t = np.ascontiguousarray(text, np.uint8).copy()
    lib().rb3_build_sais(n_seq, len(t), t.ctypes.data, n_threads)
    return t


class Rope:
    """An mrope_t* owned by the reference library."""

    def __init__(self, ptr):
        self.ptr = ptr

    @classmethod
    def from_plain(cls, bwt, n_threads=1, max_nodes=0, block_len=0):
        b = np.ascontiguousarray(bwt, np.uint8)
        return cls(lib().rb3_enc_plain2fmr(len(b), b.ctypes.data, max_nodes, block_len, n_threads))

    @classmethod
    def from_file(cls, fn):
        p = lib().mr_restore_file(fn.encode())
        if not p:
            raise IOError("mr_restore_file failed for " + fn)
        return cls(p)

    def merge_plain(self, bwt, n_threads=1):
        b = np.ascontiguousarray(bwt, np.uint8)
        lib().rb3_fmi_merge_plain(self.ptr, len(b), b.ctypes.data, n_threads)

    def _fmi(self):
        f = _Fmi()
        f.is_fmd, f.e, f.r, f.ssa, f.sid = 0, None, self.ptr, None, None
        acc = np.zeros(7, np.int64)
        lib().rb3_fmi_get_acc(C.byref(f), acc.ctypes.data)
        for i in range(7):
            f.acc[i] = int(acc[i])
        return f, acc

    def acc(self):
        return self._fmi()[1]

    def mg_rank_plain(self, bwt, n_threads=1):
        """rb3_mg_rank_plain (fm-index.c:202) -> (rb, accB)."""
        b = np.ascontiguousarray(bwt, np.uint8)
        f, _ = self._fmi()
        rb = np.zeros(len(b), np.int64)
        acc = np.zeros(7, np.int64)
        lib().rb3_mg_rank_plain(C.byref(f), len(b), b.ctypes.data, rb.ctypes.data, acc.ctypes.data, n_threads)
        return rb, acc

    def rank1a(self, k):
        k = np.asarray(k, np.int64)
        ok = np.zeros((len(k), 6), np.int64)
        ret = np.zeros(len(k), np.int8)
        L = lib()
        for i in range(len(k)):
            ret[i] = L.mr_rank2a(self.ptr, int(k[i]), -1, ok[i].ctypes.data, None)
        return ok, ret

    def dump_fmr(self):
        with tempfile.NamedTemporaryFile(suffix=".fmr", delete=False) as t:
            fn = t.name
        fp = _LIBC.fopen(fn.encode(), b"wb")
        lib().mr_dump(self.ptr, fp)
        _LIBC.fclose(fp)
        data = open(fn, "rb").read()
        os.unlink(fn)
        return data

    def to_fmd(self):
        """rb3_enc_fmr2fmd(is_free=1) + rld_dump; consumes the rope."""
        with tempfile.NamedTemporaryFile(suffix=".fmd", delete=False) as t:
            fn = t.name
        e = lib().rb3_enc_fmr2fmd(self.ptr, 0, 1)
        self.ptr = None
        lib().rld_dump(e, fn.encode())
        lib().rld_destroy(e)
        data = open(fn, "rb").read()
        os.unlink(fn)
        return data

    def close(self):
        if self.ptr:
            lib().mr_destroy(self.ptr)
            self.ptr = None


class Fmd:
    """An rld_t* restored from an .fmd image by the reference (rld0.c:295)."""

    def __init__(self, img):
        with tempfile.NamedTemporaryFile(suffix=".fmd", delete=False) as t:
            t.write(img)
            self.fn = t.name
        self.ptr = lib().rld_restore(self.fn.encode())
        os.unlink(self.fn)
        if not self.ptr:
            raise IOError("rld_restore failed")

    def rank1a(self, k):
        k = np.asarray(k, np.int64)
        ok = np.zeros((len(k), 6), np.uint64)
        ret = np.zeros(len(k), np.int8)
        for i in range(len(k)):
            ret[i] = lib().rld_rank1a(self.ptr, int(k[i]), ok[i].ctypes.data)
        return ok.astype(np.int64), ret

    def close(self):
        if self.ptr:
            lib().rld_destroy(self.ptr)
            self.ptr = None
