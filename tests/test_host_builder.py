"""Host-side logic without a GPU: the builder's device-layout blob, read back on the CPU,
agrees with the oracle; the C-ABI library loads and exports every declared symbol."""
import os
import re

import numpy as np
import pytest

import fmx_pkg
from oracle import oracle as orc
from refutil import TEXT_README, TEXT_TWINKLE, build_text
from blobreader import Blob

fmx = fmx_pkg.load()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = fmx.load_library()
    hdr = open(os.path.join(ROOT, "include", "fmx.h")).read()
    declared = set(re.findall(r"\b(fmx_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"fmx_index", "fmx_status", "fmx_kind", "fmx_mode"}
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/fmx.h but not exported"
    assert set(fmx._lib.SIGNATURES) == declared
    assert L.fmx_version() == 100


def test_no_gpu_means_loud_failure():
    L = fmx.load_library()
    if L.fmx_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fmx.Error, match="no CUDA device"):
        fmx.FMIndex.new(fmx.Text.new(b"abc\0"))


@pytest.mark.parametrize("text,msg", [
    (b"\0abc\0", "must not start with zero"),
    (b"abc", "must end with exactly one zero"),
    (b"abc\0\0", "must end with exactly one zero"),
])
def test_invalid_text_messages(text, msg):
    with pytest.raises(fmx.InvalidText, match=msg):      # sais.rs:128-139 through the C ABI
        fmx.blob_build(fmx.Text.new(text), fmx.KIND_FM, 2)
    with pytest.raises(fmx.InvalidText, match=msg):
        fmx.suffix_array(text)


def test_char_above_max_character_rejected():
    with pytest.raises(fmx.Error):
        fmx.blob_build(fmx.Text.with_max_character(bytes([1, 2, 9, 0]), 4), fmx.KIND_FM)


def test_suffix_array_matches_oracle():
    rng = np.random.default_rng(3)
    for t in range(60):
        n = int(rng.integers(2, 2000))
        text = build_text(rng, n, int(rng.choice([2, 4, 8, 200])), multi_pieces=bool(t & 1))
        assert np.array_equal(fmx.suffix_array(text), orc.suffix_array(text))
    big = build_text(rng, 300_000, 4, False)
    assert np.array_equal(fmx.suffix_array(big), orc.suffix_array(big))
    rep = (build_text(rng, 5000, 4, False)[:-1] * 40) + b"\0"      # highly repetitive
    assert np.array_equal(fmx.suffix_array(rep), orc.suffix_array(rep))


CASES = [
    (b"mississippi\0", 255, 2), (TEXT_README, 255, 2), (TEXT_TWINKLE, 255, 2), (b"a\0", 255, 2),
]


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_blob_reader_matches_oracle(kind):
    rng = np.random.default_rng(kind + 40)
    cases = list(CASES)
    for _ in range(6):
        mc = int(rng.choice([4, 7, 255]))
        text = build_text(rng, int(rng.integers(2, 700)), min(mc + 1, 8) if kind == 2 else min(mc, 7), kind == 2)
        cases.append((text, mc, int(rng.integers(0, 4))))
    for text, mc, level in cases:
        if kind != 2 and 0 in text[:-1]:
            continue
        o = orc.OracleIndex(text, kind, level=level, max_character=mc)
        b = Blob(fmx.blob_build(fmx.Text.with_max_character(text, mc), kind, level))
        n = len(text)
        assert (b.n, b.kind, b.levels, b.cs_len) == (n, kind, int(mc).bit_length(), mc + 1)
        zeros = text.count(0) if kind != 1 else 1
        assert b.layout == (1 if mc <= 4 and zeros <= 1024 else 3)
        assert b.sa_level == orc.lib().orc_sample_level(o._h)
        assert b.sa_word_size == orc.lib().orc_sample_word_size(o._h)
        rows = range(n) if n < 300 else [int(v) for v in rng.integers(0, n, 200)]
        for i in rows:
            c, nx = b.lf_step(i)
            assert c == o.get_l(i) and nx == o.lf_map(i)
            assert b.get_sa(i) == o.get_sa(i)
        chars = sorted(set(text)) + [c for c in (1, mc) if c not in text]
        for c in chars:
            for i in list(rows)[:80] + [n]:
                assert b.lf_map2(c, i) == o.lf_map2(c, i), (c, i)
        for _ in range(30):
            m = int(rng.integers(1, 8))
            p0 = int(rng.integers(0, max(1, n - m)))
            pat = text[p0:p0 + m] if rng.random() < 0.7 else bytes(int(x) for x in rng.integers(1, mc + 1, m))
            assert b.search(pat) == o.search(pat)
            if kind == 2:
                for mode in (1, 2, 3):
                    assert b.search(pat, mode) == o.search(pat, mode)
        if kind == 2:
            assert b.ndoc == o.pieces_count() and b.first_row == orc.lib().orc_first_row(o._h)
            assert [int(v) for v in b.doc] == [orc.lib().orc_doc(o._h, k) for k in range(b.ndoc)]


def test_q4_falls_back_to_wavelet_with_many_zeros(monkeypatch):
    rng = np.random.default_rng(77)
    text = build_text(rng, 9000, 5, True)          # ~17 % zeros > FMX_MAX_EXC
    assert text.count(0) > 1024
    b = Blob(fmx.blob_build(fmx.Text.with_max_character(text, 4), 2, 1))
    assert b.layout == 3                           # per-symbol bit vectors take every other case within the budget
    o = orc.OracleIndex(text, 2, level=1, max_character=4)
    for i in rng.integers(0, len(text), 100):
        c, nx = b.lf_step(int(i))
        assert c == o.get_l(int(i)) and nx == o.lf_map(int(i))
    monkeypatch.setenv("FMX_FORCE_WAVELET", "1")
    small = build_text(rng, 500, 4, False)
    assert Blob(fmx.blob_build(fmx.Text.with_max_character(small, 4), 0, 1)).layout == 0


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("env,layout", [(("FMX_FORCE_WAVELET", "1"), 0), (("FMX_SYM_BUDGET_MB", "0"), 2)])
def test_fallback_layout_blobs_match_oracle(kind, env, layout, monkeypatch):
    """FMX_FORCE_WAVELET=1 keeps the binary wavelet matrix (layout 0); a zero SYM budget gives the quaternary
    wavelet matrix (layout 2, what alphabets too large for the per-symbol vectors get): same answers"""
    monkeypatch.setenv(*env)
    rng = np.random.default_rng(kind + 400)
    for mc in (4, 37, 255):
        text = build_text(rng, 500, min(mc, 8), kind == 2)
        o = orc.OracleIndex(text, kind, level=1, max_character=mc)
        b = Blob(fmx.blob_build(fmx.Text.with_max_character(text, mc), kind, 1))
        zeros = text.count(0) if kind != 1 else 1
        assert b.layout == (1 if layout == 2 and mc <= 4 and zeros <= 1024 else layout)
        n = len(text)
        for i in range(0, n, 3):
            c, nx = b.lf_step(i)
            assert c == o.get_l(i) and nx == o.lf_map(i)
        for c in sorted(set(text)):
            for i in list(range(0, n, 17)) + [n]:
                assert b.lf_map2(c, i) == o.lf_map2(c, i)


@pytest.mark.parametrize("dense", [True, False])
@pytest.mark.parametrize("mc,level,kind", [(4, 2, 0), (4, None, 0), (255, 1, 0), (255, 5, 0), (4, 2, 2), (255, None, 2)])
def test_verify_structures_and_tail_logic(mc, level, kind, dense, monkeypatch):
    """SEC_TEXT / SEC_ISA / verify samples of the blob, and the seed-and-verify tail arithmetic (mirrored in
    blobreader.search_verify) against the oracle's plain loop: same (s, e), same executed step count"""
    monkeypatch.setenv("FMX_VERIFY_MIN_RANK_MB", "0")    # by default only indexes beyond the L2 carry the structures
    if not dense:
        monkeypatch.setenv("FMX_VERIFY_BUDGET_MB", "0")   # the sampled form (what texts beyond the budget get)
    rng = np.random.default_rng(500 + mc + (level or 0))
    n = 6000
    body = rng.integers(1, min(mc, 4) + 1, n)
    if kind == 2:
        body[rng.integers(5, n - 5, 12) // 2 * 2] = 0     # pieces (never two \0 in a row, none at either end)
    text = bytes(int(x) for x in body) + b"\0"
    b = Blob(fmx.blob_build(fmx.Text.with_max_character(text, mc), kind, level))
    o = orc.OracleIndex(text, kind, level=level, max_character=mc)
    sa = orc.suffix_array(text)
    if not dense and (kind == 2 or mc <= 4 or (level is not None and level > 3)):
        assert not b.verify                       # sampled form: only where an LF step is expensive and walks are short
        return
    assert b.verify and bytes(b.text) == text
    if dense:
        assert b.isa_level == 0 and b.vsa_level == 0 and np.array_equal(b.vsa, sa.astype(np.uint32))
    else:
        assert b.isa_level == 2 and b.vsa_level == b.sa_level
    for k in range(0, len(text), 1 << b.isa_level):
        assert int(sa[int(b.isa[k >> b.isa_level])]) == k
    assert b.has_locate == (0 if level is None else 1)
    flat_steps = []
    for t in range(400):
        m = int(rng.integers(10, 60))
        p0 = int(rng.integers(0, n - m))
        if t % 7 == 0:
            m = int(rng.integers(6, 12))               # around the shortest tail the path takes
        pat = list(text[p0:p0 + m])
        variant = t % 4
        if variant == 1:                           # a mismatch somewhere in the tail
            j = int(rng.integers(0, m))
            pat[j] = pat[j] % min(mc, 4) + 1
        elif variant == 2 and t % 8 == 2:          # runs off the start of the text
            pat = [1] * 30 + list(text[:m])
        elif variant == 3:                         # several mismatches
            for j in rng.integers(0, m, 3):
                pat[int(j)] = int(rng.integers(1, min(mc, 4) + 1))
        for mode in ((0, 1, 2, 3) if kind == 2 else (0,)):
            s, e, it = b.search_verify(pat, mode)
            os_, oe, osteps = o.search_batch(*orc.pack_patterns([bytes(pat)]), mode, want_steps=True)
            assert (s, e) == (int(os_[0]), int(oe[0])), (t, variant, mode)
            assert it == int(osteps[0]), (t, variant, mode)


# ---- texts of wide characters (character.rs:38-42: u16 / u32 / u64 / usize)

def _wide_text(rng, n, mc, dtype, multi):
    """random text over 1..=mc (a small set of symbols actually used, so patterns recur), \\0 pieces when multi"""
    used = np.unique(np.concatenate([rng.integers(1, mc + 1, 12), [mc]]))
    body = rng.choice(used, n)
    if multi:
        body[rng.integers(2, n - 2, max(1, n // 60)) // 2 * 2] = 0   # never two \0 in a row, none at either end
    return np.append(body, 0).astype(dtype)


def test_oracle_wide_characters_are_pinned_to_the_u8_oracle():
    """The reference holds no wide-character vectors; the oracle's wide path is pinned to its own u8 path (which the
    reference's golden vectors pin, test_oracle_golden.py): the same text widened gives the same index, answer for answer."""
    rng = np.random.default_rng(11)
    for kind in (orc.FM, orc.RLFM, orc.MULTI):
        t8 = np.frombuffer(build_text(rng, 1500, 7, kind == orc.MULTI), dtype=np.uint8)
        a = orc.OracleIndex(t8, kind, 2, max_character=255)
        pats = [t8[i:i + int(rng.integers(1, 7))] for i in rng.integers(0, 1400, 150)]
        pats = [p for p in pats if 0 not in p or kind == orc.MULTI]
        f, o = orc.pack_patterns(pats)
        for dt in (np.uint16, np.uint32, np.uint64):
            b = orc.OracleIndex(t8.astype(dt), kind, 2, max_character=255)
            s1, e1, st1 = a.search_batch(f, o, want_steps=True)
            s2, e2, st2 = b.search_batch(f.astype(dt), o, want_steps=True)
            assert np.array_equal(s1, s2) and np.array_equal(e1, e2) and np.array_equal(st1, st2)
            assert np.array_equal(a.locate_batch(s1, e1)[1], b.locate_batch(s2, e2)[1])
            for fwd in (False, True):
                x1, l1 = a.extract_batch(np.arange(0, 1500, 7), 9, fwd)
                x2, l2 = b.extract_batch(np.arange(0, 1500, 7), 9, fwd)
                assert x2.dtype == dt and np.array_equal(x1, x2) and np.array_equal(l1, l2)
            assert np.array_equal(orc.suffix_array(t8), orc.suffix_array(t8.astype(dt)))


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("dtype,mc", [(np.uint16, 300), (np.uint16, 65535), (np.uint32, 1_000_003), (np.uint64, 70_000)])
def test_wide_blob_matches_oracle(kind, dtype, mc):
    """max_character > 255: the WIDE layout (all levels in one section, cs / adj of max_character + 1 words),
    read back on the CPU against the oracle: primitives, walks to the samples, searches in every mode"""
    rng = np.random.default_rng(kind * 10 + mc % 97)
    n = 900
    text = _wide_text(rng, n, mc, dtype, kind == 2)
    level = int(rng.integers(0, 4))
    o = orc.OracleIndex(text, kind, level=level, max_character=mc)
    b = Blob(fmx.blob_build(fmx.Text.with_max_character(text, mc), kind, level))
    n = text.size
    assert (b.n, b.kind, b.levels, b.cs_len, b.layout, b.char_width) == (n, kind, int(mc).bit_length(), mc + 1, 4, np.dtype(dtype).itemsize)
    assert not b.verify and b.sa_level == orc.lib().orc_sample_level(o._h)
    assert np.array_equal(fmx.suffix_array(text), orc.suffix_array(text, mc))
    rows = [int(v) for v in rng.integers(0, n, 150)] + [0, n - 1]
    for i in rows:
        c, nx = b.lf_step(i)
        assert c == o.get_l(i) and nx == o.lf_map(i)
        assert b.get_sa(i) == o.get_sa(i)
    present = [int(c) for c in np.unique(text)]
    absent = [c for c in (1, 2, mc - 1, mc, present[1] + 1) if c not in present and c <= mc]
    for c in present + absent:
        for i in rows[:40] + [n]:
            assert b.lf_map2(c, i) == o.lf_map2(c, i), (c, i)
    for _ in range(60):
        m = int(rng.integers(1, 6))
        p0 = int(rng.integers(0, n - m))
        pat = text[p0:p0 + m] if rng.random() < 0.7 else rng.integers(1, mc + 1, m).astype(dtype)
        if kind != 2 and 0 in pat:
            continue
        for mode in ((0, 1, 2, 3) if kind == 2 else (0,)):
            assert b.search([int(c) for c in pat], mode) == o.search(pat, mode)
    if kind == 2:
        assert b.ndoc == o.pieces_count() and b.first_row == orc.lib().orc_first_row(o._h)
        assert [int(v) for v in b.doc] == [orc.lib().orc_doc(o._h, k) for k in range(b.ndoc)]


def test_wide_characters_over_a_small_alphabet_take_the_u8_layouts():
    """u16 / u32 / u64 texts with max_character <= 255: narrowed, same blob as the u8 text except the header's char_width"""
    rng = np.random.default_rng(5)
    t8 = np.frombuffer(build_text(rng, 3000, 4, False), dtype=np.uint8)
    ref = fmx.blob_build(fmx.Text.with_max_character(t8, 4), 0, 2, mode=fmx.MODE_COMPACT)
    for dt in (np.uint16, np.uint32, np.uint64):
        w = fmx.blob_build(fmx.Text.with_max_character(t8.astype(dt), 4), 0, 2)
        bw = Blob(w)
        assert bw.char_width == np.dtype(dt).itemsize and bw.layout == 1 and not bw.verify
        diff = np.nonzero(w != ref)[0]
        assert diff.size <= 1      # the char_width field of the header only
    with pytest.raises(fmx.Error, match="larger than max_character"):
        fmx.blob_build(fmx.Text.with_max_character(np.array([1, 300, 2, 0], dtype=np.uint16), 255), 0, 2)
    with pytest.raises(fmx.InvalidText, match="must end with exactly one zero"):
        fmx.blob_build(fmx.Text.with_max_character(np.array([1, 300, 2], dtype=np.uint16), 300), 0, 2)
    with pytest.raises(fmx.Error, match="out of range"):
        fmx.blob_build(fmx.Text.new(np.array([1, 300, 2, 0], dtype=np.uint32)), 0, 2)   # Text::new on u32: C::max_value()
