"""ctypes binding of the CPU oracle (oracle/fmx_oracle.c).

TEST INFRASTRUCTURE ONLY: may be imported from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from fm-index_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfmx_oracle.so")

FM, RLFM, MULTI = 0, 1, 2
SEARCH, SEARCH_PREFIX, SEARCH_SUFFIX, SEARCH_EXACT = 0, 1, 2, 3
NONE = (1 << 64) - 1

_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)


def build_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "fmx_oracle.c")
    hdr = os.path.join(_HERE, "fmx_oracle.h")
    if (
        force
        or not os.path.exists(_SO)
        or (os.path.exists(src) and os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfmx_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build_oracle()
    L = C.CDLL(_SO)
    vp = C.c_void_p
    u64 = C.c_uint64
    L.orc_build.restype = vp
    L.orc_build.argtypes = [vp, u64, u64, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    L.orc_build_from_sa.restype = vp
    L.orc_build_from_sa.argtypes = [vp, u64, u64, C.c_int, C.c_int, vp, C.c_char_p, C.c_size_t]
    L.orc_free.argtypes = [vp]
    L.orc_check_suffix_array.restype = C.c_int
    L.orc_check_suffix_array.argtypes = [vp, u64, vp]
    L.orc_suffix_array.restype = C.c_int
    L.orc_suffix_array.argtypes = [vp, u64, vp, C.c_char_p, C.c_size_t]
    for name in ("orc_len", "orc_pieces_count", "orc_heap_bits", "orc_cs_len", "orc_rlfm_runs", "orc_first_row"):
        getattr(L, name).restype = u64
        getattr(L, name).argtypes = [vp]
    for name in ("orc_get_l", "orc_lf_map", "orc_get_f", "orc_fl_map", "orc_piece_id", "orc_sample_get",
                 "orc_cs", "orc_rlfm_s", "orc_doc"):
        getattr(L, name).restype = u64
        getattr(L, name).argtypes = [vp, u64]
    L.orc_lf_map2.restype = u64
    L.orc_lf_map2.argtypes = [vp, u64, u64]
    L.orc_get_sa.restype = u64
    L.orc_get_sa.argtypes = [vp, u64, _u64p]
    for name in ("orc_rlfm_b", "orc_rlfm_bp"):
        getattr(L, name).restype = C.c_int
        getattr(L, name).argtypes = [vp, u64]
    for name in ("orc_sample_level", "orc_sample_word_size"):
        getattr(L, name).restype = C.c_uint32
        getattr(L, name).argtypes = [vp]
    L.orc_search.restype = C.c_int64
    L.orc_search.argtypes = [vp, C.c_int, vp, u64, C.c_int, u64, u64, _u64p, _u64p]
    L.orc_search_batch.restype = C.c_int
    L.orc_search_batch.argtypes = [vp, C.c_int, vp, vp, u64, vp, vp, vp, vp, vp, C.c_int]
    L.orc_locate_batch.restype = C.c_int
    L.orc_locate_batch.argtypes = [vp, C.c_int, vp, vp, u64, vp, vp, vp, vp, C.c_int]
    L.orc_extract_batch.restype = None
    L.orc_extract_batch.argtypes = [vp, vp, u64, C.c_uint32, C.c_int, vp, vp, C.c_int]
    L.orc_max_threads.restype = C.c_int
    L.orc_build_w.restype = vp
    L.orc_build_w.argtypes = [vp, C.c_uint32, u64, u64, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    L.orc_suffix_array_w.restype = C.c_int
    L.orc_suffix_array_w.argtypes = [vp, C.c_uint32, u64, u64, vp, C.c_char_p, C.c_size_t]
    L.orc_search_batch_w.restype = C.c_int
    L.orc_search_batch_w.argtypes = [vp, C.c_int, vp, C.c_uint32, vp, u64, vp, vp, vp, vp, vp, C.c_int]
    L.orc_extract_batch_w.restype = None
    L.orc_extract_batch_w.argtypes = [vp, vp, u64, C.c_uint32, C.c_int, vp, C.c_uint32, vp, C.c_int]
    _lib = L
    return L


class InvalidText(Exception):
    """Error::InvalidText (src/error.rs:3-6)."""


def _as_u8(x) -> np.ndarray:
    if isinstance(x, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    if isinstance(x, str):
        return np.frombuffer(x.encode(), dtype=np.uint8)
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint8))


WIDE_DTYPES = (np.uint16, np.uint32, np.uint64)   # character.rs:38-42 (usize = u64)


def _char_dtype(x):
    """numpy dtype of a text / pattern: the array's own when it is one of the wide character types, else u8"""
    dt = getattr(x, "dtype", None)
    if dt is not None and np.dtype(dt) in [np.dtype(d) for d in WIDE_DTYPES]:
        return np.dtype(dt)
    return np.dtype(np.uint8)


def _as_chars(x, dtype) -> np.ndarray:
    dtype = np.dtype(dtype)
    if dtype == np.uint8:
        return _as_u8(x)
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


def pack_patterns(patterns, dtype=np.uint8):
    """list of patterns -> (flat character array, u64 offsets[npat+1] in characters)."""
    arrs = [_as_chars(p, dtype) for p in patterns]
    off = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        off[1:] = np.cumsum([a.size for a in arrs], dtype=np.uint64)
    flat = np.concatenate(arrs) if arrs and off[-1] > 0 else np.zeros(0, dtype=dtype)
    return np.ascontiguousarray(flat), off


def suffix_array(text, max_character=None) -> np.ndarray:
    dt = _char_dtype(text)
    t = _as_chars(text, dt)
    sa = np.zeros(max(t.size, 1), dtype=np.uint64)
    err = C.create_string_buffer(256)
    if dt == np.uint8:
        rc = lib().orc_suffix_array(t.ctypes.data, t.size, sa.ctypes.data, err, 256)
    else:
        mc = int(t.max()) if max_character is None and t.size else int(max_character or 1)
        rc = lib().orc_suffix_array_w(t.ctypes.data, dt.itemsize, t.size, mc, sa.ctypes.data, err, 256)
    if rc != 0:
        raise InvalidText(err.value.decode())
    return sa[: t.size]


def check_suffix_array(text, sa) -> bool:
    """independent linear-time check that `sa` is the suffix array of `text`"""
    t = _as_u8(text)
    sa = np.ascontiguousarray(sa, dtype=np.uint64)
    if sa.size != t.size:
        return False
    return lib().orc_check_suffix_array(t.ctypes.data, t.size, sa.ctypes.data) == 0


class OracleIndex:
    """One of the reference's six index types, by (kind, level)."""

    def __init__(self, text, kind=FM, level=None, max_character=255, sa=None):
        self.dtype = _char_dtype(text)   # u8 unless the text is a numpy array of u16 / u32 / u64 characters
        self._t = _as_chars(text, self.dtype)
        self.kind = kind
        err = C.create_string_buffer(256)
        lvl = -1 if level is None else int(level)
        if self.dtype != np.uint8:
            if sa is not None:
                raise ValueError("a supplied suffix array is only taken for u8 texts")
            h = lib().orc_build_w(self._t.ctypes.data, self.dtype.itemsize, self._t.size, max_character, kind, lvl, err, 256)
        elif sa is None:
            h = lib().orc_build(self._t.ctypes.data, self._t.size, max_character, kind, lvl, err, 256)
        else:
            sa = np.ascontiguousarray(sa, dtype=np.uint64)
            h = lib().orc_build_from_sa(self._t.ctypes.data, self._t.size, max_character, kind, lvl,
                                        sa.ctypes.data, err, 256)
        if not h:
            raise InvalidText(err.value.decode())
        self._h = h
        self.max_character = max_character

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().orc_free(h)
            self._h = None

    # ---- scalars
    def __len__(self):
        return lib().orc_len(self._h)

    def len(self):
        return lib().orc_len(self._h)

    def pieces_count(self):
        return lib().orc_pieces_count(self._h)

    def get_l(self, i):
        return lib().orc_get_l(self._h, i)

    def lf_map(self, i):
        return lib().orc_lf_map(self._h, i)

    def lf_map2(self, c, i):
        return lib().orc_lf_map2(self._h, c, i)

    def get_f(self, i):
        return lib().orc_get_f(self._h, i)

    def fl_map(self, i):
        v = lib().orc_fl_map(self._h, i)
        return None if v == NONE else v

    def get_sa(self, i):
        return lib().orc_get_sa(self._h, i, None)

    def piece_id(self, i):
        return lib().orc_piece_id(self._h, i)

    def sample_get(self, i):
        v = lib().orc_sample_get(self._h, i)
        return None if v == NONE else v

    def cs(self, c):
        return lib().orc_cs(self._h, c)

    # ---- wrapper.rs
    def search(self, pattern, mode=SEARCH, init=None):
        if self.dtype != np.uint8:
            p = _as_chars(pattern, self.dtype)
            s, e = self.search_batch(p, np.array([0, p.size], dtype=np.uint64), mode,
                                     None if init is None else [init[0]], None if init is None else [init[1]])
            return int(s[0]), int(e[0])
        p = _as_u8(pattern)
        s, e = C.c_uint64(0), C.c_uint64(0)
        it = lib().orc_search(self._h, mode, p.ctypes.data, p.size, init is not None,
                              init[0] if init else 0, init[1] if init else 0, C.byref(s), C.byref(e))
        if it < 0:
            raise IndexError("pattern character exceeds max_character (the reference panics)")
        return s.value, e.value

    def search_batch(self, flat, off, mode=SEARCH, init_s=None, init_e=None, nthreads=0, want_steps=False):
        flat = np.ascontiguousarray(flat, dtype=self.dtype)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        npat = off.size - 1
        s = np.zeros(npat, dtype=np.uint64)
        e = np.zeros(npat, dtype=np.uint64)
        steps = np.zeros(npat, dtype=np.uint32) if want_steps else None
        if self.dtype != np.uint8:
            def call(*a):
                return lib().orc_search_batch_w(a[0], a[1], a[2], self.dtype.itemsize, *a[3:])
        else:
            call = lib().orc_search_batch
        rc = call(
            self._h, mode, flat.ctypes.data, off.ctypes.data, npat,
            None if init_s is None else np.ascontiguousarray(init_s, dtype=np.uint64).ctypes.data,
            None if init_e is None else np.ascontiguousarray(init_e, dtype=np.uint64).ctypes.data,
            s.ctypes.data, e.ctypes.data, None if steps is None else steps.ctypes.data, nthreads)
        if rc != 0:
            raise IndexError("pattern character exceeds max_character (the reference panics)")
        return (s, e, steps) if want_steps else (s, e)

    def locate_batch(self, s, e, prefix_only=False, want_piece_ids=False, want_positions=True, nthreads=0):
        s = np.ascontiguousarray(s, dtype=np.uint64)
        e = np.ascontiguousarray(e, dtype=np.uint64)
        npat = s.size
        off = np.zeros(npat + 1, dtype=np.uint64)
        lib().orc_locate_batch(self._h, int(prefix_only), s.ctypes.data, e.ctypes.data, npat,
                               off.ctypes.data, None, None, None, nthreads)
        total = int(off[-1])
        pos = np.zeros(total, dtype=np.uint64) if want_positions else None
        pid = np.zeros(total, dtype=np.uint64) if want_piece_ids else None
        steps = C.c_uint64(0)
        if total and (want_positions or want_piece_ids):
            lib().orc_locate_batch(self._h, int(prefix_only), s.ctypes.data, e.ctypes.data, npat,
                                   off.ctypes.data, None if pos is None else pos.ctypes.data,
                                   None if pid is None else pid.ctypes.data, C.byref(steps), nthreads)
        self.last_lf_steps = steps.value
        return off, pos, pid

    def extract_batch(self, rows, k, forward, nthreads=0):
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.zeros((rows.size, k), dtype=self.dtype)
        out_len = np.zeros(rows.size, dtype=np.uint32)
        lib().orc_extract_batch_w(self._h, rows.ctypes.data, rows.size, k, int(forward), out.ctypes.data,
                                  self.dtype.itemsize, out_len.ctypes.data, nthreads)
        return out, out_len

    # convenience mirroring the crate's Search/Match use in its tests
    def count(self, pattern, mode=SEARCH):
        s, e = self.search(pattern, mode)
        return e - s

    def locate_all(self, pattern, mode=SEARCH):
        s, e = self.search(pattern, mode)
        po = mode in (SEARCH_PREFIX, SEARCH_EXACT)
        _, pos, _ = self.locate_batch([s], [e], prefix_only=po)
        return [int(v) for v in pos]

    def piece_ids_all(self, pattern, mode=SEARCH):
        s, e = self.search(pattern, mode)
        po = mode in (SEARCH_PREFIX, SEARCH_EXACT)
        _, _, pid = self.locate_batch([s], [e], prefix_only=po, want_piece_ids=True, want_positions=False)
        return [int(v) for v in pid]

    def rows(self, pattern, mode=SEARCH):
        s, e = self.search(pattern, mode)
        po = mode in (SEARCH_PREFIX, SEARCH_EXACT)
        return [i for i in range(s, e) if not po or self.get_l(i) == 0]


def max_threads() -> int:
    return lib().orc_max_threads()
