"""fmx -- B200-native batched FM-index query engine.

Host-side mirror of the reference crate's public API (ajalab/fm-index 0.3.1, src/frontend.rs,
src/text.rs) over the C ABI in include/fmx.h: same type names, same argument meaning, same
error behaviour, plus the batched entries the GPU path exists for.  All queries run as CUDA
kernels on a B200; there is no CPU fallback (a missing CUDA library or device is an error).

    text  = Text.new(b"mississippi\\0")
    index = FMIndexWithLocate.new(text, 2)
    s = index.search(b"ssi")
    s.count(); [m.locate() for m in s.iter_matches()]
    batch = index.search_batch([b"ssi", b"pp"])          # many patterns, one launch
    batch.count(); batch.locate()
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import load as load_library

__all__ = [
    "Text", "Error", "InvalidText", "PieceId", "FMIndex", "FMIndexWithLocate", "RLFMIndex",
    "RLFMIndexWithLocate", "FMIndexMultiPieces", "FMIndexMultiPiecesWithLocate", "Search", "Match",
    "SearchBatch", "load_library", "suffix_array", "random_gather_peak", "pack_patterns", "MODE_AUTO", "MODE_COMPACT",
    "MODE_RICH", "IndexGroup", "GROUP_REPLICATE", "GROUP_BY_PIECE",
]

KIND_FM, KIND_RLFM, KIND_MULTI = 0, 1, 2
MODE_AUTO, MODE_COMPACT, MODE_RICH = 0, 1, 2
SEARCH, SEARCH_PREFIX, SEARCH_SUFFIX, SEARCH_EXACT = 0, 1, 2, 3
_NONE = (1 << 64) - 1


class Error(Exception):
    """src/error.rs:1-20"""

    def __init__(self, msg, code=0):
        super().__init__(msg)
        self.code = code


class InvalidText(Error):
    """Error::InvalidText (src/error.rs:5)"""

    def __str__(self):
        return "invalid text: " + super().__str__()


def _check(rc):
    if rc == 0:
        return
    msg = (load_library().fmx_last_error() or b"").decode()
    if rc == -1:
        raise InvalidText(msg, rc)
    if rc == -5:
        raise IndexError(msg)  # the reference panics with an index-out-of-bounds (fm_index.rs:94)
    raise Error(msg, rc)


def _as_u8(x) -> np.ndarray:
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.uint8)
    if isinstance(x, str):
        x = x.encode()
    if isinstance(x, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    return np.ascontiguousarray(np.asarray(list(x), dtype=np.uint8))


WIDE_DTYPES = (np.dtype(np.uint16), np.dtype(np.uint32), np.dtype(np.uint64))   # character.rs:38-42 (usize = u64)


def _char_dtype(x):
    """character type of a text: the numpy dtype of a u16 / u32 / u64 array, u8 for everything else"""
    dt = getattr(x, "dtype", None)
    return np.dtype(dt) if dt is not None and np.dtype(dt) in WIDE_DTYPES else np.dtype(np.uint8)


def _as_chars(x, dtype) -> np.ndarray:
    dtype = np.dtype(dtype)
    if dtype == np.uint8:
        return _as_u8(x)
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


def _ptr(a):
    return None if a is None else a.ctypes.data


class PieceId(int):
    """src/piece.rs:1-15"""

    def __repr__(self):
        return f"PieceId({int(self)})"


class Text:
    """src/text.rs:10-64.  Characters are u8 unless `text` is a numpy array of u16 / u32 / u64 (character.rs:38-42)."""

    def __init__(self, text, max_character=None):
        self._dtype = _char_dtype(text)
        self._text = _as_chars(text, self._dtype)
        self._max_character = int(np.iinfo(self._dtype).max if max_character is None else max_character)

    @classmethod
    def new(cls, text):
        """Text::new: max_character = C::max_value() (text.rs:28-33).  For u32 / u64 texts that is beyond what any
        index can hold (the reference would allocate max_character + 1 counters): use with_max_character there."""
        return cls(text, None)

    def dtype(self):
        return self._dtype

    @classmethod
    def with_max_character(cls, text, max_character):
        """text.rs:44-49"""
        return cls(text, max_character)

    def text(self):
        return self._text

    def max_character(self):
        return self._max_character

    def max_bits(self):
        """text.rs:61-63"""
        return int(self._max_character).bit_length()


def _pack(patterns, dtype=np.uint8):
    """-> (flat characters of `dtype`, offsets u64 (in characters) or None, fixed_len, npat)"""
    if isinstance(patterns, np.ndarray) and patterns.ndim == 2:
        p = np.ascontiguousarray(patterns, dtype=dtype)
        return p.reshape(-1), None, p.shape[1], p.shape[0]
    if isinstance(patterns, tuple) and len(patterns) == 2:
        flat = np.ascontiguousarray(patterns[0], dtype=dtype)
        off = np.ascontiguousarray(patterns[1], dtype=np.uint64)
        return flat, off, 0, off.size - 1
    arrs = [_as_chars(p, dtype) for p in patterns]
    off = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        off[1:] = np.cumsum([a.size for a in arrs], dtype=np.uint64)
    flat = np.concatenate(arrs) if arrs and int(off[-1]) > 0 else np.zeros(0, dtype=dtype)
    return np.ascontiguousarray(flat), off, 0, len(arrs)


def _run_query(call, patterns, mode, rows, counts, locate, piece_ids, width, packed_bits, fixed_len, capacity, dtype=np.uint8):
    """fills a struct fmx_query, runs call(byref(query), byref(total)) -> rc, returns the requested arrays"""
    dt = np.uint64 if width == 8 else np.uint32
    if packed_bits:
        words = np.ascontiguousarray(patterns, dtype=np.uint64).reshape(-1)
        wpp = (int(fixed_len) * packed_bits + 63) // 64
        npat, flat, off, fixed = words.size // wpp, words, None, int(fixed_len)
    else:
        flat, off, fixed, npat = _pack(patterns, dtype)
    q = _lib.Query()
    q.mode, q.packed_bits, q.patterns, q.pat_off, q.fixed_len, q.npat = mode, packed_bits, _ptr(flat), _ptr(off), fixed, npat
    q.out_width = width
    out = {}
    if rows:
        out["s"], out["e"] = np.zeros(npat, dtype=np.uint64), np.zeros(npat, dtype=np.uint64)
        q.out_s, q.out_e = _ptr(out["s"]), _ptr(out["e"])
    if counts:
        out["counts"] = np.zeros(npat, dtype=dt)
        q.counts = _ptr(out["counts"])
    cap = 0
    if locate or piece_ids:
        out["hit_off"] = np.zeros(npat + 1, dtype=dt)
        q.hit_off = _ptr(out["hit_off"])
        cap = int(capacity) if capacity is not None else max(1024, 2 * npat)
        if locate:
            out["positions"] = np.zeros(cap, dtype=dt)
            q.positions = _ptr(out["positions"])
        if piece_ids:
            out["piece_ids"] = np.zeros(cap, dtype=dt)
            q.piece_ids = _ptr(out["piece_ids"])
    q.capacity = cap
    total = C.c_uint64(0)
    rc = call(C.byref(q), C.byref(total))
    t = int(total.value)
    if rc == -9 and capacity is None and t < (1 << 32 if width == 4 else 1 << 62):  # retry with exact-size buffers
        return _run_query(call, patterns, mode, rows, counts, locate, piece_ids, width, packed_bits, fixed_len, t, dtype)
    _check(rc)
    for k in ("positions", "piece_ids"):
        if k in out:
            out[k] = out[k][:t]
    out["total"] = t
    return out


class _Index:
    """Common part of the six index types (SearchIndex trait, frontend.rs:26-45)."""

    _kind = KIND_FM
    _locate = False

    def __init__(self, text: Text, level=None, device=0, _handle=None, mode=MODE_AUTO):
        L = load_library()
        self._L = L
        if _handle is not None:
            self._h = _handle
            self._dtype = np.dtype({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[int(L.fmx_index_char_width(_handle))])
            return
        if not isinstance(text, Text):
            text = Text.new(text)
        lvl = -1
        if self._locate:
            if level is None:
                raise TypeError("a sampling level is required (FMIndexWithLocate::new(&text, level))")
            lvl = int(level)
        t = text.text()
        self._dtype = text.dtype()   # width of the text's characters = width of patterns and extracted characters
        h = C.c_void_p()
        _check(L.fmx_index_build_ex(_ptr(t), t.size, self._dtype.itemsize, text.max_character(), self._kind, lvl, device, int(mode),
                                    C.byref(h)))
        self._h = h

    @classmethod
    def new(cls, text, *args, **kw):
        return cls(text, *args, **kw)

    @classmethod
    def load(cls, path, device=0):
        L = load_library()
        h = C.c_void_p()
        _check(L.fmx_index_load(str(path).encode(), device, C.byref(h)))
        self = cls(None, _handle=h)
        if L.fmx_index_kind(h) != cls._kind or bool(L.fmx_index_has_locate(h)) != cls._locate:
            L.fmx_index_free(h)
            self._h = None
            raise Error("index file holds a different index type")
        return self

    def save(self, path):
        _check(self._L.fmx_index_save(self._h, str(path).encode()))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.fmx_index_free(h)
            self._h = None

    # ---- SearchIndex
    def len(self):
        return int(self._L.fmx_index_len(self._h))

    __len__ = len

    def heap_size(self):
        """frontend.rs:41-44; here: bytes of the device-resident index."""
        return int(self._L.fmx_index_device_bytes(self._h))

    def search(self, pattern):
        return Search(self, SEARCH)._refine(pattern)

    # ---- batched entries
    def search_batch(self, patterns, mode=SEARCH, init=None):
        """Batched SearchIndex::search over many patterns (one kernel launch)."""
        flat, off, fixed, npat = _pack(patterns, self._dtype)
        s = np.zeros(npat, dtype=np.uint64)
        e = np.zeros(npat, dtype=np.uint64)
        ins = ine = None
        if init is not None:
            ins = np.ascontiguousarray(init[0], dtype=np.uint64)
            ine = np.ascontiguousarray(init[1], dtype=np.uint64)
        _check(self._L.fmx_search_batch(self._h, mode, _ptr(flat), _ptr(off), fixed, npat, _ptr(ins), _ptr(ine),
                                        _ptr(s), _ptr(e)))
        return SearchBatch(self, mode, s, e)

    def search_locate_batch(self, patterns, mode=SEARCH, piece_ids=False, capacity=None):
        """Fused, pipelined batched search + locate (fmx_search_locate_batch): the batched form of
        `index.search(p).iter_matches().map(|m| m.locate())`.
        -> (SearchBatch, hit_off[npat+1], positions[, piece_ids])"""
        flat, off, fixed, npat = _pack(patterns, self._dtype)
        s = np.zeros(npat, dtype=np.uint64)
        e = np.zeros(npat, dtype=np.uint64)
        hit_off = np.zeros(npat + 1, dtype=np.uint64)
        cap = int(capacity) if capacity is not None else max(1024, 2 * npat)
        pos = np.zeros(cap, dtype=np.uint64)
        pid = np.zeros(cap, dtype=np.uint64) if piece_ids else None
        total = C.c_uint64(0)
        rc = self._L.fmx_search_locate_batch(self._h, mode, _ptr(flat), _ptr(off), fixed, npat, _ptr(s), _ptr(e),
                                             _ptr(hit_off), _ptr(pos), _ptr(pid), cap, C.byref(total))
        batch = SearchBatch(self, mode, s, e)
        if rc == -9:  # FMX_ERR_CAPACITY: hit_off and total are valid; fetch the hits with exact-size buffers
            res = batch.locate(piece_ids)
            return (batch,) + tuple(res)
        _check(rc)
        t = int(total.value)
        return (batch, hit_off, pos[:t].copy(), pid[:t].copy()) if piece_ids else (batch, hit_off, pos[:t].copy())

    def query_batch(self, patterns, mode=SEARCH, rows=False, counts=False, locate=True, piece_ids=False, width=8,
                    packed_bits=0, fixed_len=None, capacity=None):
        """fmx_query_batch: count + locate of a batch in one call, byte or packed patterns, 64- or 32-bit outputs.
        `patterns`: as for search_batch, or (packed_bits != 0) a uint64 array of packed words with `fixed_len`
        characters per pattern (see pack_patterns).  -> dict with the requested arrays + "total"."""
        return _run_query(lambda q, t: self._L.fmx_query_batch(self._h, q, t), patterns, mode, rows, counts, locate, piece_ids,
                          width, packed_bits, fixed_len, capacity, self._dtype)

    def mode(self):
        """MODE_COMPACT or MODE_RICH: what the index holds (fmx_index_mode_of)"""
        return int(self._L.fmx_index_mode_of(self._h))

    def locate_page(self, s, e, first_hit, nhits, piece_ids=False):
        """One page of the hit list of the ranges (s, e): hits [first_hit, first_hit + nhits) in the reference's
        iteration order (fmx_locate_page).  -> (positions[, piece_ids], total_hits)"""
        s = np.ascontiguousarray(s, dtype=np.uint64)
        e = np.ascontiguousarray(e, dtype=np.uint64)
        pos = np.zeros(max(int(nhits), 1), dtype=np.uint64)
        pid = np.zeros(max(int(nhits), 1), dtype=np.uint64) if piece_ids else None
        total = C.c_uint64(0)
        _check(self._L.fmx_locate_page(self._h, _ptr(s), _ptr(e), s.size, int(first_hit), int(nhits), _ptr(pos), _ptr(pid),
                                       C.byref(total)))
        got = max(0, min(int(nhits), int(total.value) - int(first_hit)))
        return (pos[:got], pid[:got], int(total.value)) if piece_ids else (pos[:got], int(total.value))

    def locate_batch(self, s, e, prefix_only=False, piece_ids=False):
        s = np.ascontiguousarray(s, dtype=np.uint64)
        e = np.ascontiguousarray(e, dtype=np.uint64)
        npat = s.size
        off = np.zeros(npat + 1, dtype=np.uint64)
        ppos, ppid = C.c_void_p(), C.c_void_p()
        _check(self._L.fmx_locate_batch(self._h, int(prefix_only), _ptr(s), _ptr(e), npat, _ptr(off), C.byref(ppos),
                                        C.byref(ppid) if piece_ids else None))
        total = int(off[-1])

        def take(p):
            if not p.value:
                return np.zeros(0, dtype=np.uint64)
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(total,)).copy()
            self._L.fmx_free(p)
            return a

        pos = take(ppos)
        return (off, pos, take(ppid)) if piece_ids else (off, pos)

    def extract_batch(self, rows, k, forward):
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.zeros((rows.size, k), dtype=self._dtype)
        out_len = np.zeros(rows.size, dtype=np.uint32)
        _check(self._L.fmx_extract_batch(self._h, _ptr(rows), rows.size, k, int(forward), _ptr(out), _ptr(out_len)))
        return out, out_len

    def rows_op(self, op, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.zeros(rows.size, dtype=np.uint64)
        _check(self._L.fmx_rows_op(self._h, op, _ptr(rows), rows.size, _ptr(out)))
        return out

    def lf_map2_batch(self, c, i):
        c = np.ascontiguousarray(c, dtype=self._dtype)
        i = np.ascontiguousarray(i, dtype=np.uint64)
        out = np.zeros(i.size, dtype=np.uint64)
        _check(self._L.fmx_lf_map2_batch(self._h, _ptr(c), _ptr(i), i.size, _ptr(out)))
        return out

    def sectors_per_rank(self):
        """32-byte sectors per rank/access probe in this index's device layout (L, or 1 for Q4)."""
        return int(self._L.fmx_index_sectors_per_rank(self._h))

    def layout_name(self):
        """which device layout the builder chose (fmx_layout.h)"""
        return ["binary wavelet matrix (L sectors per rank)", "Q4: one quaternary level (1 sector per rank)",
                "WM4: quaternary wavelet matrix (ceil(L/2) sectors per rank)",
                "SYM: one bit vector per symbol (1 sector per rank)",
                "WIDE: binary wavelet matrix of up to 32 levels, character tables in global memory"][int(self._L.fmx_index_layout(self._h))]

    def kmer_k(self, big=False):
        """characters memoised by the small / large k-mer table (0 = none)"""
        return int(self._L.fmx_index_kmer_k(self._h, int(big)))

    def set_option(self, key, value):
        """tuning knobs for A/B measurements; results never change"""
        _check(self._L.fmx_index_set_option(self._h, key.encode(), int(value)))

    def last_work(self, stream=None):
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(self._L.fmx_last_work(self._h, stream, C.byref(a), C.byref(b)))
        return a.value, b.value


class _MultiMixin:
    """SearchIndexWithMultiPieces (frontend.rs:47-63)."""

    def search_prefix(self, pattern):
        return Search(self, SEARCH_PREFIX)._refine(pattern)

    def search_suffix(self, pattern):
        return Search(self, SEARCH_SUFFIX)._refine(pattern)

    def search_exact(self, pattern):
        return Search(self, SEARCH_EXACT)._refine(pattern)

    def pieces_count(self):
        return int(self._L.fmx_index_pieces_count(self._h))


class FMIndex(_Index):
    """frontend.rs:110, :195-203"""
    _kind, _locate = KIND_FM, False


class FMIndexWithLocate(_Index):
    """frontend.rs:124-126, :205-218"""
    _kind, _locate = KIND_FM, True


class RLFMIndex(_Index):
    """frontend.rs:139, :220-228"""
    _kind, _locate = KIND_RLFM, False


class RLFMIndexWithLocate(_Index):
    """frontend.rs:153-155, :230-243"""
    _kind, _locate = KIND_RLFM, True


class FMIndexMultiPieces(_MultiMixin, _Index):
    """frontend.rs:168-170, :245-252"""
    _kind, _locate = KIND_MULTI, False


class FMIndexMultiPiecesWithLocate(_MultiMixin, _Index):
    """frontend.rs:184-186, :254-267"""
    _kind, _locate = KIND_MULTI, True


class Search:
    """Search trait (frontend.rs:65-83) over wrapper.rs:14-23."""

    def __init__(self, index, mode, s=None, e=None):
        self._index = index
        self._mode = mode
        self._s = s
        self._e = e
        self._cache = None

    def _refine(self, pattern):
        p = _as_chars(pattern, self._index._dtype)
        init = None if self._s is None else (np.array([self._s], dtype=np.uint64), np.array([self._e], dtype=np.uint64))
        b = self._index.search_batch([p], self._mode, init)
        return Search(self._index, self._mode, int(b.s[0]), int(b.e[0]))

    def search(self, pattern):
        """Refine: prepends `pattern` (wrapper.rs:99-124)."""
        return self._refine(pattern)

    def get_range(self):
        """wrapper.rs:126-129 (test-only in the reference)"""
        return self._s, self._e

    def count(self):
        """wrapper.rs:132-134: e - s, ignoring the prefix filter."""
        return self._e - self._s

    @property
    def _prefix_only(self):
        return self._mode in (SEARCH_PREFIX, SEARCH_EXACT)

    def _rows(self):
        if self._e <= self._s:
            return np.zeros(0, dtype=np.uint64)
        rows = np.arange(self._s, self._e, dtype=np.uint64)
        if self._prefix_only:  # wrapper.rs:208
            rows = rows[self._index.rows_op(0, rows) == 0]
        return rows

    def _hits(self):
        if self._cache is None:
            rows = self._rows()
            pos = pid = None
            if self._index._locate and rows.size:
                if self._index._kind == KIND_MULTI:
                    _, pos, pid = self._index.locate_batch([self._s], [self._e], self._prefix_only, piece_ids=True)
                else:
                    _, pos = self._index.locate_batch([self._s], [self._e], self._prefix_only)
            self._cache = (rows, pos, pid)
        return self._cache

    def iter_matches(self):
        """wrapper.rs:137-139, 203-217: ascending SA-row order."""
        rows, pos, pid = self._hits()
        for k in range(rows.size):
            yield Match(self._index, int(rows[k]), None if pos is None else int(pos[k]),
                        None if pid is None else int(pid[k]))


class Match:
    """Match / MatchWithLocate / MatchWithPieceId (frontend.rs:85-104) over wrapper.rs:219-248."""

    def __init__(self, index, row, position=None, piece=None):
        self._index = index
        self._i = row
        self._position = position
        self._piece = piece

    def row(self):
        return self._i

    def locate(self):
        if not self._index._locate:
            raise AttributeError("locate() needs a ...WithLocate index (MatchWithLocate, frontend.rs:94-98)")
        if self._position is None:
            self._position = int(self._index.rows_op(4, [self._i])[0])
        return self._position

    def piece_id(self):
        if not (self._index._locate and self._index._kind == KIND_MULTI):
            raise AttributeError("piece_id() needs FMIndexMultiPiecesWithLocate (frontend.rs:542)")
        if self._piece is None:
            _, _, pid = self._index.locate_batch([self._i], [self._i + 1], False, piece_ids=True)
            self._piece = int(pid[0])
        return PieceId(self._piece)

    def _iter(self, forward):
        k, done = 32, 0
        while True:
            out, ln = self._index.extract_batch([self._i], k, forward)
            n = int(ln[0])
            for t in range(done, n):
                yield int(out[0, t])
            if n < k:  # forward ended (fl_map None, multi_pieces.rs:171-181)
                return
            done, k = n, k * 2

    def iter_chars_forward(self):
        """wrapper.rs:229-231, 175-183"""
        return self._iter(True)

    def iter_chars_backward(self):
        """wrapper.rs:233-235, 154-161 (never ends: the walk is cyclic)"""
        return self._iter(False)


class SearchBatch:
    """Result of a batched search: SA ranges of many patterns, resident on the host."""

    def __init__(self, index, mode, s, e):
        self._index, self._mode, self.s, self.e = index, mode, s, e

    def __len__(self):
        return self.s.size

    def count(self):
        return self.e - self.s

    def search_batch(self, patterns):
        return self._index.search_batch(patterns, self._mode, init=(self.s, self.e))

    def locate(self, piece_ids=False):
        """-> (hit_off[npat+1], positions[, piece_ids]); matches of pattern p are
        positions[hit_off[p]:hit_off[p+1]] in the reference's iteration order."""
        return self._index.locate_batch(self.s, self.e, self._mode in (SEARCH_PREFIX, SEARCH_EXACT), piece_ids)

    def iter_locate_pages(self, page_hits):
        """Lazily yields (first_hit, positions) pages of at most `page_hits` matches, in the reference's
        iteration order: bounded memory for match sets too large to hold at once."""
        if self._mode in (SEARCH_PREFIX, SEARCH_EXACT):
            raise Error("paged locate supports the unfiltered modes (search, search_suffix)")
        first = 0
        while True:
            pos, total = self._index.locate_page(self.s, self.e, first, page_hits)
            if pos.size:
                yield first, pos
            first += int(page_hits)
            if first >= total:
                return


GROUP_REPLICATE, GROUP_BY_PIECE = 0, 1


class IndexGroup:
    """fmx_group: several GPUs of this process behind one handle (include/fmx.h, "multi-GPU").
    GROUP_REPLICATE: the index copied to every device, batches cut into shards (input order and the reference's
    iteration order preserved).  GROUP_BY_PIECE (MultiPieces): pieces partitioned over the devices, every device
    answers the whole batch, parts gathered and merged on the first device."""

    def __init__(self, text, kind, level, devices, group_mode=GROUP_REPLICATE, mode=MODE_AUTO):
        self._L = load_library()
        if not isinstance(text, Text):
            text = Text.new(text)
        t = text.text()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        _check(self._L.fmx_group_create(devs, len(devices), int(group_mode), _ptr(t), t.size, 1, text.max_character(), int(kind),
                                        -1 if level is None else int(level), int(mode), C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.fmx_group_free(h)
            self._h = None

    def size(self):
        return int(self._L.fmx_group_size(self._h))

    def len(self):
        return int(self._L.fmx_group_len(self._h))

    def pieces_count(self):
        return int(self._L.fmx_group_pieces_count(self._h))

    def query_batch(self, patterns, mode=SEARCH, rows=False, counts=False, locate=True, piece_ids=False, width=8,
                    packed_bits=0, fixed_len=None, capacity=None):
        return _run_query(lambda q, t: self._L.fmx_group_query_batch(self._h, q, t), patterns, mode, rows, counts, locate,
                          piece_ids, width, packed_bits, fixed_len, capacity)


def suffix_array(text) -> np.ndarray:
    """sais::build_suffix_array (sais.rs:115-144) through the C ABI (host side)."""
    t = _as_chars(text, _char_dtype(text))
    sa = np.zeros(max(t.size, 1), dtype=np.uint64)
    _check(load_library().fmx_build_suffix_array(_ptr(t), t.size, t.dtype.itemsize, _ptr(sa)))
    return sa[: t.size]


def suffix_array_device(text, max_character=255, device=0):
    """The same suffix array built on the GPU (gpu_sa.cu). -> (sa, doubling rounds)"""
    t = _as_u8(text)
    sa = np.zeros(max(t.size, 1), dtype=np.uint64)
    rounds = C.c_int(0)
    _check(load_library().fmx_build_suffix_array_device(_ptr(t), t.size, 1, max_character, device, _ptr(sa),
                                                        C.byref(rounds)))
    return sa[: t.size], rounds.value


def pack_patterns(patterns, bits=2) -> np.ndarray:
    """[npat, m] characters 1..2^bits -> packed words as fmx_query wants them: pattern p occupies
    ceil(m * bits / 64) uint64 words, character k in bits [k*bits, (k+1)*bits) of that stream, stored as c - 1."""
    p = np.ascontiguousarray(patterns, dtype=np.uint8)
    npat, m = p.shape
    per = 64 // bits
    wpp = (m + per - 1) // per
    codes = np.zeros((npat, wpp * per), dtype=np.uint64)
    codes[:, :m] = p.astype(np.uint64) - 1
    shifts = (np.arange(per, dtype=np.uint64) * np.uint64(bits))[None, None, :]
    return np.bitwise_or.reduce(codes.reshape(npat, wpp, per) << shifts, axis=2).reshape(-1)


def blob_build(text: Text, kind, level=None, mode=MODE_AUTO) -> np.ndarray:
    """Host-only half of construction: the device-layout blob as bytes."""
    L = load_library()
    t = text.text()
    p, nb = C.c_void_p(), C.c_uint64(0)
    _check(L.fmx_blob_build_ex(_ptr(t), t.size, t.dtype.itemsize, text.max_character(), kind, -1 if level is None else level, int(mode),
                               C.byref(p), C.byref(nb)))
    a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nb.value,)).copy()
    L.fmx_free(p)
    return a


def random_gather_peak(device=0, nbytes=4 << 30, nloads=1 << 28, iters=3, load_bytes=32) -> float:
    """Measured peak random gather rate in 32-byte sectors/s: the random-access roofline."""
    v = C.c_double(0)
    _check(load_library().fmx_random_gather_bench(device, nbytes, nloads, iters, load_bytes, C.byref(v)))
    return v.value
