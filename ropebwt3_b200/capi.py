"""ctypes binding of librb3b200.so (the C ABI declared in include/rb3_b200.h).

The shared library is the product; this module only marshals numpy arrays and
raw device pointers into it.  There is no CPU fallback: if the library is not
built, or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librb3b200.so")
_LIB = None

_vp, _i64, _int = C.c_void_p, C.c_int64, C.c_int

# name -> (restype, argtypes); every symbol declared in include/rb3_b200.h
SIGNATURES = {
    "rb3b_init": (_int, [_int]),
    "rb3b_last_error": (C.c_char_p, []),
    "rb3b_version": (C.c_char_p, []),
    "rb3b_set_stream": (_int, [_vp]),
    "rb3b_sync": (_int, []),
    "rb3b_trim": (_int, []),
    "rb3b_ctx_create": (_vp, [_int]),
    "rb3b_ctx_make_current": (_int, [_vp]),
    "rb3b_ctx_destroy": (None, [_vp]),
    "rb3b_host_alloc_pinned": (_vp, [_i64]),
    "rb3b_host_free_pinned": (None, [_vp]),
    "rb3b_set_param": (_int, [C.c_char_p, _i64]),
    "rb3b_get_stat": (_i64, [C.c_char_p]),
    "rb3b_index_create": (_vp, []),
    "rb3b_index_destroy": (None, [_vp]),
    "rb3b_index_reserve": (_int, [_vp, _i64]),
    "rb3b_index_set_order": (_int, [_vp, _int]),
    "rb3b_index_get_order": (_int, [_vp]),
    "rb3b_insert_multi": (_int, [_vp, _i64, _vp]),
    "rb3b_insert_multi_dev": (_int, [_vp, _i64, _vp]),
    "rb3b_build_bwt_so": (_int, [_i64, _vp, _int, _vp]),
    "rb3b_build_bwt_so_dev": (_int, [_i64, _vp, _int, _vp]),
    "rb3b_index_from_plain": (_int, [_vp, _i64, _vp]),
    "rb3b_index_from_plain_dev": (_int, [_vp, _i64, _vp]),
    "rb3b_index_from_runs": (_int, [_vp, _i64, _vp, _vp]),
    "rb3b_index_from_runs_device": (_int, [_vp, _i64, _vp, _vp]),
    "rb3b_merge_plain": (_int, [_vp, _i64, _vp]),
    "rb3b_merge_plain_dev": (_int, [_vp, _i64, _vp]),
    "rb3b_prefetch_batch": (_int, [_i64, _vp]),
    "rb3b_mg_rank_plain": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "rb3b_mg_rank_plain_dev": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "rb3b_mg_rank_part": (_int, [_vp, _i64, _vp, _int, _int, _vp]),
    "rb3b_merge_with_ka": (_int, [_vp, _i64, _vp, _vp]),
    "rb3b_merge_index": (_int, [_vp, _vp]),
    "rb3b_merge_plain_dist_dev": (_int, [_vp, _i64, _vp]),
    "rb3b_merge_plain_dist": (_int, [_vp, _i64, _vp]),
    "rb3b_dist_unique_id": (_int, [_vp]),
    "rb3b_dist_init": (_int, [_int, _int, _vp]),
    "rb3b_dist_finalize": (_int, []),
    "rb3b_dist_rank": (_int, []),
    "rb3b_dist_world": (_int, []),
    "rb3b_dist_nccl_version": (_int, []),
    "rb3b_rank1a": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "rb3b_rank1a_dev": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "rb3b_lf_dev": (_int, [_vp, _i64, _vp, _vp, _vp, _int]),
    "rb3b_get_acc": (_i64, [_vp, _vp]),
    "rb3b_index_bytes": (_i64, [_vp]),
    "rb3b_index_wait": (_int, [_vp]),
    "rb3b_export_runs": (_i64, [_vp, _vp, _vp, _i64]),
    "rb3b_dump_fmd": (_int, [_vp, C.c_char_p]),
    "rb3b_dump_fmr": (_int, [_vp, C.c_char_p, _int, _int]),
    "rb3b_dump_plain": (_int, [_vp, C.c_char_p]),
    "rb3b_restore": (_int, [_vp, C.c_char_p]),
    "rb3b_fmd_image": (_i64, [_i64, _vp, _vp, C.POINTER(_vp)]),
    "rb3b_fmr_image": (_i64, [_i64, _vp, _vp, _int, _int, C.POINTER(_vp)]),
    "rb3b_runs_from_image": (_i64, [_vp, _i64, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_int)]),
    "rb3b_host_free": (None, [_vp]),
    "rb3b_build_bwt": (_int, [_i64, _vp, _vp]),
    "rb3b_build_bwt_dev": (_int, [_i64, _vp, _vp]),
    "rb3b_max_batch_symbols": (_i64, [_i64]),
    "rb3b_batch_prepare": (_vp, [_i64, _vp]),
    "rb3b_batch_prepare_dev": (_vp, [_i64, _vp]),
    "rb3b_batch_destroy": (None, [_vp]),
    "rb3b_batch_len": (_i64, [_vp]),
    "rb3b_batch_bwt_dev": (_vp, [_vp]),
    "rb3b_merge_prepared": (_int, [_vp, _vp]),
    "rb3b_ssa_sizes": (_int, [_vp, _int, _vp, _vp, _vp]),
    "rb3b_ssa_gen_dev": (_int, [_vp, _int, _vp, _vp]),
    "rb3b_ssa_dump": (_int, [_vp, _int, C.c_char_p]),
    "rb3b_dev_alloc": (_vp, [_i64]),
    "rb3b_dev_free": (None, [_vp]),
    "rb3b_h2d": (_int, [_vp, _vp, _i64]),
    "rb3b_d2h": (_int, [_vp, _vp, _i64]),
}


class Rb3bError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "rb3b error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load librb3b200.so; raises if it was not built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C ropebwt3_b200/csrc). There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _LIB = L
    return _LIB


def check(rc):
    if rc < 0:
        raise Rb3bError(rc, lib().rb3b_last_error().decode())
    return rc


def ptr(a):
    """numpy array -> void* (must be C-contiguous), int -> device pointer as is."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return int(a)
