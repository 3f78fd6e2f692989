"""ctypes binding of the C ABI (include/fmx.h).  Fails loudly when the CUDA library is missing."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_vp, _u64, _u32, _int = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
_pp = C.POINTER(C.c_void_p)
_u64p = C.POINTER(C.c_uint64)

# name -> (restype, argtypes): exactly the symbols include/fmx.h declares
SIGNATURES = {
    "fmx_last_error": (C.c_char_p, []),
    "fmx_version": (_int, []),
    "fmx_device_count": (_int, []),
    "fmx_free": (None, [_vp]),
    "fmx_index_build": (_int, [_vp, _u64, _u32, _u64, _int, _int, _int, _pp]),
    "fmx_index_build_ex": (_int, [_vp, _u64, _u32, _u64, _int, _int, _int, _int, _pp]),
    "fmx_index_mode_of": (_int, [_vp]),
    "fmx_blob_build": (_int, [_vp, _u64, _u32, _u64, _int, _int, _pp, _u64p]),
    "fmx_blob_build_ex": (_int, [_vp, _u64, _u32, _u64, _int, _int, _int, _pp, _u64p]),
    "fmx_index_from_blob": (_int, [_vp, _u64, _int, _pp]),
    "fmx_index_save": (_int, [_vp, C.c_char_p]),
    "fmx_index_load": (_int, [C.c_char_p, _int, _pp]),
    "fmx_index_free": (None, [_vp]),
    "fmx_build_suffix_array": (_int, [_vp, _u64, _u32, _vp]),
    "fmx_build_suffix_array_device": (_int, [_vp, _u64, _u32, _u64, _int, _vp, C.POINTER(C.c_int)]),
    "fmx_index_set_option": (_int, [_vp, C.c_char_p, C.c_int64]),
    "fmx_index_len": (_u64, [_vp]),
    "fmx_index_device_bytes": (_u64, [_vp]),
    "fmx_index_pieces_count": (_u64, [_vp]),
    "fmx_index_kind": (_int, [_vp]),
    "fmx_index_has_locate": (_int, [_vp]),
    "fmx_index_device": (_int, [_vp]),
    "fmx_index_wavelet_levels": (_u32, [_vp]),
    "fmx_index_sample_level": (_u32, [_vp]),
    "fmx_index_sectors_per_rank": (_u32, [_vp]),
    "fmx_index_kmer_k": (_u32, [_vp, _int]),
    "fmx_index_kmer_entry_bytes": (_u32, [_vp]),
    "fmx_index_layout": (_u32, [_vp]),
    "fmx_index_char_width": (_u32, [_vp]),
    "fmx_index_has_text": (_int, [_vp]),
    "fmx_search_batch": (_int, [_vp, _int, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp]),
    "fmx_search_batch_device": (_int, [_vp, _int, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp]),
    "fmx_search_check": (_int, [_vp, _vp]),
    "fmx_locate_batch": (_int, [_vp, _int, _vp, _vp, _u64, _vp, _pp, _pp]),
    "fmx_search_locate_batch": (_int, [_vp, _int, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _u64, _u64p]),
    "fmx_query_batch": (_int, [_vp, _vp, _u64p]),
    "fmx_query_batch_device": (_int, [_vp, _vp, _vp]),
    "fmx_csr_merge_device": (_int, [_int, _vp, _int, _u64, _u32, _vp, _vp, _vp, _u64, _vp]),
    "fmx_index_clone": (_int, [_vp, _int, _pp]),
    "fmx_group_create": (_int, [_vp, _int, _int, _vp, _u64, _u32, _u64, _int, _int, _int, _pp]),
    "fmx_group_free": (None, [_vp]),
    "fmx_group_size": (_int, [_vp]),
    "fmx_group_index": (_vp, [_vp, _int]),
    "fmx_group_len": (_u64, [_vp]),
    "fmx_group_pieces_count": (_u64, [_vp]),
    "fmx_group_query_batch": (_int, [_vp, _vp, _u64p]),
    "fmx_locate_count_device": (_int, [_vp, _int, _vp, _vp, _u64, _vp, _u64p, _vp]),
    "fmx_locate_fill_device": (_int, [_vp, _int, _vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp]),
    "fmx_locate_page": (_int, [_vp, _vp, _vp, _u64, _u64, _u64, _vp, _vp, _u64p]),
    "fmx_locate_page_device": (_int, [_vp, _vp, _vp, _u64, _u64, _u64, _vp, _vp, _vp]),
    "fmx_locate_batch_device": (_int, [_vp, _int, _vp, _vp, _u64, _vp, _vp, _vp, _u64, _vp]),
    "fmx_extract_batch": (_int, [_vp, _vp, _u64, _u32, _int, _vp, _vp]),
    "fmx_extract_batch_device": (_int, [_vp, _vp, _u64, _u32, _int, _vp, _vp, _vp]),
    "fmx_rows_op": (_int, [_vp, _int, _vp, _u64, _vp]),
    "fmx_lf_map2_batch": (_int, [_vp, _vp, _vp, _u64, _vp]),
    "fmx_last_work": (_int, [_vp, _vp, _u64p, _u64p]),
    "fmx_last_requests": (_int, [_vp, _vp, _u64p, _u64p]),
    "fmx_last_phase_ms": (_int, [_vp, _vp, C.POINTER(C.c_float), _int]),
    "fmx_random_gather_bench": (_int, [_int, _u64, _u64, _int, _u32, C.POINTER(C.c_double)]),
    "fmx_launch_count": (_u64, []),
}



class Query(C.Structure):
    """struct fmx_query (include/fmx.h)"""
    _fields_ = [("mode", C.c_int), ("packed_bits", C.c_uint32), ("patterns", C.c_void_p), ("pat_off", C.c_void_p),
                ("fixed_len", C.c_uint64), ("npat", C.c_uint64), ("out_width", C.c_uint32), ("reserved", C.c_uint32),
                ("out_s", C.c_void_p), ("out_e", C.c_void_p), ("counts", C.c_void_p), ("hit_off", C.c_void_p),
                ("positions", C.c_void_p), ("piece_ids", C.c_void_p), ("capacity", C.c_uint64)]


class CsrPart(C.Structure):
    """struct fmx_csr_part (include/fmx.h)"""
    _fields_ = [("hit_off", C.c_void_p), ("positions", C.c_void_p), ("piece_ids", C.c_void_p),
                ("position_base", C.c_uint64), ("piece_base", C.c_uint64)]


_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load libfmx_b200.so.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing and _build.needs_build():
        _build.build()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: the fmx engine is CUDA-only and has no CPU fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # raises AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
