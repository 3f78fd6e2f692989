"""GPU parity tests: the CUDA path, called through the C ABI (include/fmx.h) via the Python mirror of
the reference API, against (1) the reference's golden vectors, (2) the CPU oracle on seeded inputs,
(3) a naive scan (the reference's own differential method, tests/testutil/mod.rs).
Bit-exact: SA ranges, counts, locate positions IN THE REFERENCE'S ORDER, piece ids, extracted chars."""
import os

import numpy as np
import pytest

import fmx_pkg
from oracle import oracle as orc
from refutil import (TEXT_README, TEXT_TWINKLE, build_text, naive_search, random_cases)

pytestmark = pytest.mark.gpu
fmx = fmx_pkg.load()

MISS = b"mississippi\0"
KINDS = {
    orc.FM: (fmx.FMIndex, fmx.FMIndexWithLocate),
    orc.RLFM: (fmx.RLFMIndex, fmx.RLFMIndexWithLocate),
    orc.MULTI: (fmx.FMIndexMultiPieces, fmx.FMIndexMultiPiecesWithLocate),
}


def chain(index, op, start, n):
    i, out = start, []
    for _ in range(n):
        i = int(index.rows_op(op, [i])[0])
        out.append(i)
    return out


# ---- src/fm_index.rs:148-173, src/rlfmi.rs:255-351
@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
def test_mississippi_golden(kind):
    index = KINDS[kind][1].new(fmx.Text.new(MISS), 2)
    assert chain(index, 1, 0, 12) == [1, 6, 7, 2, 8, 10, 3, 9, 11, 4, 5, 0]
    rows = np.arange(12)
    assert bytes(index.rows_op(0, rows).astype(np.uint8)) == b"ipssm\0pissii"
    assert bytes(index.rows_op(2, rows).astype(np.uint8)) == bytes(sorted(MISS))
    if kind != orc.MULTI:
        assert list(index.rows_op(3, rows)) == [5, 0, 7, 10, 11, 4, 1, 6, 2, 3, 8, 9]
    n = index.len()
    for c, r in [(0, (0, 1)), (ord("i"), (1, 5)), (ord("m"), (5, 6)), (ord("p"), (6, 8)), (ord("s"), (8, 12))]:
        if kind == orc.MULTI and c == 0:
            continue
        assert tuple(index.lf_map2_batch([c, c], [0, n])) == r
    for pat, r in [(b"iss", (3, 5)), (b"ppi", (7, 8)), (b"si", (8, 10)), (b"ssi", (10, 12))]:
        assert index.search(pat).get_range() == r


# ---- README.md:49-85
def test_readme_doctest():
    index = fmx.FMIndexWithLocate.new(fmx.Text.new(TEXT_README), 2)
    search = index.search("dolor")
    assert search.count() == 4
    assert [m.locate() for m in search.iter_matches()] == [246, 12, 300, 103]
    import itertools
    prefix = list(itertools.islice(next(iter(search.iter_matches())).iter_chars_backward(), 16))
    assert bytes(reversed(prefix)) == b"Duis aute irure "
    m3 = list(search.iter_matches())[3]
    assert bytes(itertools.islice(m3.iter_chars_forward(), 20)) == b"dolore magna aliqua."


# ---- examples/multi_pieces.rs:34-88
def test_multi_pieces_example():
    import itertools
    index = fmx.FMIndexMultiPiecesWithLocate.new(fmx.Text.new(TEXT_TWINKLE), 2)
    assert index.search("star").count() == 4
    assert sorted(int(m.piece_id()) for m in index.search("How I wonder").iter_matches()) == [0, 0, 1, 2]
    pre = [bytes(itertools.takewhile(lambda c: c != ord(" "), m.iter_chars_backward()))
           for m in index.search(" in the dark").iter_matches()]
    assert pre == [b"rellevart"]
    suc = [bytes(itertools.takewhile(lambda c: c != ord(","), m.iter_chars_forward()))
           for m in index.search("ing ").iter_matches()]
    assert suc == [b"ing shines upon", b"ing sun is gone"]
    assert sorted(int(m.piece_id()) for m in index.search_prefix("Twinkle").iter_matches()) == [0]
    assert sorted(int(m.piece_id()) for m in index.search_suffix("what you are!\n").iter_matches()) == [0, 1, 2]


# ---- tests/test_fmindex.rs:6-24, test_rlfmindex.rs, test_multi_pieces.rs:8-41, test_api.rs
@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
def test_small_and_api(kind):
    index = KINDS[kind][1].new(fmx.Text.new(b"a\0"), 2)
    assert index.search("a").count() == 1
    assert [m.locate() for m in index.search("a").iter_matches()] == [0]
    if kind == orc.MULTI:
        assert [int(m.piece_id()) for m in index.search("a").iter_matches()] == [0]
    count_only = KINDS[kind][0].new(fmx.Text.new(b"text\0"))
    assert count_only.len() == 5 and count_only.heap_size() > 0
    assert count_only.search("t").count() == 2
    with pytest.raises(fmx.Error):
        count_only.locate_batch([0], [1])                 # DiscardedSuffixArray: no locate
    with pytest.raises(fmx.InvalidText, match="must not start with zero"):
        KINDS[kind][0].new(fmx.Text.new(b"\0a\0"))


# ---- the reference's randomized differential tests, against BOTH the oracle and the naive scan
@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
def test_random_differential(kind):
    multi = kind == orc.MULTI
    for text, level, pats in random_cases(seed=100 + kind, texts=40, patterns=100, text_size_max=1024,
                                          alphabet_size=8, level_max=3, pattern_size_max=10, multi_pieces=multi):
        index = KINDS[kind][1].new(fmx.Text.new(text), level)
        oracle = orc.OracleIndex(text, kind, level=level)
        flat, off = orc.pack_patterns(pats)
        modes = [fmx.SEARCH] + ([fmx.SEARCH_PREFIX, fmx.SEARCH_SUFFIX, fmx.SEARCH_EXACT] if multi else [])
        for mode in modes:
            b = index.search_batch(pats, mode)
            os_, oe = oracle.search_batch(flat, off, mode)
            assert np.array_equal(b.s, os_) and np.array_equal(b.e, oe)
            po = mode in (fmx.SEARCH_PREFIX, fmx.SEARCH_EXACT)
            if multi:
                hoff, pos, pid = b.locate(piece_ids=True)
                ooff, opos, opid = oracle.locate_batch(os_, oe, prefix_only=po, want_piece_ids=True)
                assert np.array_equal(pid, opid)
            else:
                hoff, pos = b.locate()
                ooff, opos, _ = oracle.locate_batch(os_, oe, prefix_only=po)
            assert np.array_equal(hoff, ooff) and np.array_equal(pos, opos)     # order included
            for k, pat in enumerate(pats[:25]):
                exp = naive_search(text, pat, mode in (1, 3), mode in (2, 3))
                got = sorted(int(v) for v in pos[int(hoff[k]):int(hoff[k + 1])])
                assert got == [p for p, _ in exp]
                if mode == fmx.SEARCH:
                    assert int(b.e[k] - b.s[k]) == len(exp)
                if multi:
                    assert sorted(int(v) for v in pid[int(hoff[k]):int(hoff[k + 1])]) == [d for _, d in exp]


@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
def test_backend_primitives_vs_oracle(kind):
    rng = np.random.default_rng(kind)
    for mc in (4, 7, 255):
        text = build_text(rng, 3000, min(mc, 8), kind == orc.MULTI)
        index = KINDS[kind][1].new(fmx.Text.with_max_character(text, mc), 3)
        oracle = orc.OracleIndex(text, kind, level=3, max_character=mc)
        n = len(text)
        rows = np.arange(n, dtype=np.uint64)
        assert list(index.rows_op(0, rows)) == [oracle.get_l(i) for i in range(n)]
        assert list(index.rows_op(1, rows)) == [oracle.lf_map(i) for i in range(n)]
        assert list(index.rows_op(2, rows)) == [oracle.get_f(i) for i in range(n)]
        exp_fl = [oracle.fl_map(i) for i in range(n)]
        assert list(index.rows_op(3, rows)) == [(1 << 64) - 1 if v is None else v for v in exp_fl]
        assert list(index.rows_op(4, rows)) == [oracle.get_sa(i) for i in range(n)]
        if kind == orc.MULTI:
            assert list(index.rows_op(5, rows)) == [oracle.piece_id(i) for i in range(n)]   # literal walk
        cs, is_ = [], []
        for c in sorted(set(text)) + [mc]:
            for i in list(rng.integers(0, n + 1, 300)) + [0, n]:
                cs.append(c)
                is_.append(int(i))
        assert list(index.lf_map2_batch(cs, is_)) == [oracle.lf_map2(c, i) for c, i in zip(cs, is_)]


@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
def test_extraction_vs_oracle(kind):
    rng = np.random.default_rng(50 + kind)
    text = build_text(rng, 5000, 8, kind == orc.MULTI)
    index = KINDS[kind][0].new(fmx.Text.new(text))
    oracle = orc.OracleIndex(text, kind)
    rows = rng.integers(0, len(text), 500).astype(np.uint64)
    for fwd in (False, True):
        out, ln = index.extract_batch(rows, 37, fwd)
        oout, oln = oracle.extract_batch(rows, 37, fwd)
        assert np.array_equal(ln, oln) and np.array_equal(out, oout)


@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
def test_extraction_from_the_text_in_rich_mode(kind):
    """HBM-rich indexes read extracted characters from the text at SA[row] (k_extract_text): same characters and
    lengths as the oracle's LF / FL walks, forward and backward, incl. walks that wrap around the text, odd lengths,
    pieces that end early, and the A/B option that forces the walking kernel"""
    rng = np.random.default_rng(70 + kind)
    L = fmx.load_library()
    for n, mc in ((70, 4), (5000, 4), (40_000, 200)):
        text = build_text(rng, n, min(mc + 1, 8) if kind == orc.MULTI else min(mc, 7), kind == orc.MULTI)
        if kind != orc.MULTI and 0 in text[:-1]:
            continue
        index = KINDS[kind][1].new(fmx.Text.with_max_character(text, mc), 2, mode=fmx.MODE_RICH)
        oracle = orc.OracleIndex(text, kind, level=2, max_character=mc)
        assert L.fmx_index_has_text(index._h) == 1
        rows = np.concatenate([np.arange(min(len(text), 300)), rng.integers(0, len(text), 700)]).astype(np.uint64)
        for k in (1, 4, 7, 32, 37, 100):
            for fwd in (False, True):
                oout, oln = oracle.extract_batch(rows, k, fwd)
                for opt in (1, 0):
                    index.set_option("extract_text", opt)
                    out, ln = index.extract_batch(rows, k, fwd)
                    assert np.array_equal(ln, oln) and np.array_equal(out, oout), (n, mc, k, fwd, opt)
    compact = KINDS[kind][1].new(fmx.Text.with_max_character(text, mc), 2, mode=fmx.MODE_COMPACT)
    assert L.fmx_index_has_text(compact._h) == 0


# ---- the four device layouts (fmx_layout.h) answer identically: the builder picks Q4 for DNA-coded
# texts and per-symbol bit vectors (SYM) otherwise; a zero SYM budget gives the quaternary wavelet
# matrix (what alphabets too large for the budget get); FMX_FORCE_WAVELET=1 keeps the binary matrix
@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
@pytest.mark.parametrize("env", [("FMX_FORCE_WAVELET", "1"), ("FMX_SYM_BUDGET_MB", "0")])
def test_fallback_layouts_parity(kind, env, monkeypatch):
    monkeypatch.setenv(*env)
    rng = np.random.default_rng(300 + kind)
    multi = kind == orc.MULTI
    for mc in (4, 37, 255):
        text = build_text(rng, 4000, min(mc, 20), multi)
        index = KINDS[kind][1].new(fmx.Text.with_max_character(text, mc), 2)
        bits = int(mc).bit_length()
        if env[0] == "FMX_FORCE_WAVELET":
            assert index.sectors_per_rank() == bits
        else:
            q4 = mc <= 4 and (text.count(0) if kind != orc.RLFM else 1) <= 1024
            assert index.sectors_per_rank() == (1 if q4 else (bits + 1) // 2)
        oracle = orc.OracleIndex(text, kind, level=2, max_character=mc)
        n = len(text)
        rows = np.arange(n, dtype=np.uint64)
        assert list(index.rows_op(0, rows)) == [oracle.get_l(i) for i in range(n)]
        assert list(index.rows_op(1, rows)) == [oracle.lf_map(i) for i in range(n)]
        exp_fl = [oracle.fl_map(i) for i in range(n)]
        assert list(index.rows_op(3, rows)) == [(1 << 64) - 1 if v is None else v for v in exp_fl]
        pats = [bytes(text[p:p + m]) if k & 1 else bytes(int(x) for x in rng.integers(1, min(mc, 20) + 1, m))
                for k, (p, m) in enumerate(zip(rng.integers(0, n - 12, 400), rng.integers(1, 12, 400)))]
        pats = [p for p in pats if multi or 0 not in p]
        flat, off = orc.pack_patterns(pats)
        b = index.search_batch(pats)
        s, e = oracle.search_batch(flat, off)
        assert np.array_equal(b.s, s) and np.array_equal(b.e, e)
        hoff, pos = b.locate()
        ooff, opos, _ = oracle.locate_batch(s, e)
        assert np.array_equal(hoff, ooff) and np.array_equal(pos, opos)


def test_one_sector_per_rank_is_the_default():
    rng = np.random.default_rng(310)
    for mc in (5, 37, 255):
        text = build_text(rng, 2000, min(mc, 30), False)
        index = fmx.FMIndex.new(fmx.Text.with_max_character(text, mc))
        assert index.sectors_per_rank() == 1           # per-symbol bit vectors
    dna_like = build_text(rng, 2000, 4, False)
    assert fmx.FMIndex.new(fmx.Text.with_max_character(dna_like, 4)).sectors_per_rank() == 1


def test_refinement_and_empty_pattern():
    index = fmx.FMIndexWithLocate.new(fmx.Text.new(MISS), 0)
    assert index.search(b"").get_range() == (0, 12)
    assert index.search(b"si").search(b"s").get_range() == index.search(b"ssi").get_range()
    b = index.search_batch([b"si", b"pi", b"i"]).search_batch([b"s", b"p", b"ss"])
    assert [tuple(x) for x in zip(b.s, b.e)] == [index.search(p).get_range() for p in (b"ssi", b"ppi", b"ssi")]
    empty = index.search_batch([])
    assert len(empty) == 0
    hoff, pos = empty.locate()
    assert list(hoff) == [0] and pos.size == 0


def test_pattern_char_above_max_character():
    text = bytes([1, 2, 3, 4, 1, 2, 0])
    index = fmx.FMIndex.new(fmx.Text.with_max_character(text, 4))
    with pytest.raises(IndexError):
        index.search(bytes([1, 5]))
    r = index.search(bytes([5, 4, 4])).get_range()       # range empties before the bad char is reached
    assert r[0] == r[1]
    assert index.search(bytes([1, 2])).count() == 2       # the handle stays usable after the error


def test_save_load_roundtrip(tmp_path):
    rng = np.random.default_rng(9)
    text = build_text(rng, 20000, 4, False)
    index = fmx.RLFMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 2)
    p = tmp_path / "index.fmx"
    index.save(p)
    again = fmx.RLFMIndexWithLocate.load(p)
    pats = [text[i:i + 12] for i in range(0, 5000, 97)]
    a, b = index.search_batch(pats), again.search_batch(pats)
    assert np.array_equal(a.s, b.s) and np.array_equal(a.e, b.e)
    assert all(np.array_equal(x, y) for x, y in zip(a.locate(), b.locate()))
    with pytest.raises(fmx.Error):
        fmx.FMIndex.load(p)


def dna(n, seed):
    rng = np.random.default_rng(seed)
    return np.append(rng.integers(1, 5, n, dtype=np.uint8), np.uint8(0))


def mixed_patterns(text, npat, m, seed, sigma=4):
    rng = np.random.default_rng(seed)
    pats = rng.integers(1, sigma + 1, (npat, m), dtype=np.uint8)
    starts = rng.integers(0, text.size - 1 - m, npat)
    idx = np.arange(m)
    sampled = text[starts[:, None] + idx[None, :]]
    pats[::2] = sampled[::2]
    return pats, starts


# ---- BASELINE config 1: 10k random 20-mers over 1 MB random ACGT + \0 (+ the reference's own
# bench shape, benches/count.rs:23-24: 50 000 binary symbols, all 256 8-mers)
def test_config1_count_parity():
    text = dna(1_000_000, 1)
    pats, _ = mixed_patterns(text, 10_000, 20, 2)
    index = fmx.FMIndex.new(fmx.Text.with_max_character(text, 4))
    index.set_option("count_work", 1)
    oracle = orc.OracleIndex(text, orc.FM, max_character=4)
    b = index.search_batch(pats)
    s, e, steps = oracle.search_batch(pats.reshape(-1), np.arange(10_001, dtype=np.uint64) * 20, want_steps=True)
    assert np.array_equal(b.s, s) and np.array_equal(b.e, e)
    assert index.last_work()[0] == int(steps.sum())           # executed iterations: the roofline numerator

    rng = np.random.default_rng(0)
    btext = np.append(np.where(rng.random(50_000) < 0.5, ord("0"), ord("1")).astype(np.uint8), np.uint8(0))
    bpats = np.array([[ord("0") + ((v >> (7 - k)) & 1) for k in range(8)] for v in range(256)], dtype=np.uint8)
    for cls, kind in ((fmx.FMIndexWithLocate, orc.FM), (fmx.RLFMIndexWithLocate, orc.RLFM)):
        ix = cls.new(fmx.Text.with_max_character(btext, ord("1")), 2)
        oc = orc.OracleIndex(btext, kind, level=2, max_character=ord("1"))
        bb = ix.search_batch(bpats)
        s, e = oc.search_batch(bpats.reshape(-1), np.arange(257, dtype=np.uint64) * 8)
        assert np.array_equal(bb.s, s) and np.array_equal(bb.e, e)
        hoff, pos = bb.locate()
        ooff, opos, _ = oc.locate_batch(s, e)
        assert np.array_equal(hoff, ooff) and np.array_equal(pos, opos)
        assert int(hoff[-1]) == 50_000 - 7


# ---- config 2 shape at a size the oracle finishes in seconds, + size-independent properties
@pytest.mark.parametrize("n,npat", [(4_000_000, 200_000)])
def test_config2_locate_parity_and_properties(n, npat):
    text = dna(n, 3)
    pats, starts = mixed_patterns(text, npat, 32, 4)
    index = fmx.FMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 2)
    oracle = orc.OracleIndex(text, orc.FM, level=2, max_character=4)
    b = index.search_batch(pats)
    hoff, pos = b.locate()
    s, e = oracle.search_batch(pats.reshape(-1), np.arange(npat + 1, dtype=np.uint64) * 32)
    ooff, opos, _ = oracle.locate_batch(s, e)
    assert np.array_equal(b.s, s) and np.array_equal(b.e, e)
    assert np.array_equal(hoff, ooff) and np.array_equal(pos, opos)
    check_locate_properties(text, pats, starts, b, hoff, pos)


def check_locate_properties(text, pats, starts, b, hoff, pos):
    """Size-independent properties: every located position really holds the pattern; every sampled
    pattern finds its own origin; counts equal hit counts; positions are distinct per pattern."""
    npat, m = pats.shape
    cnt = (b.e - b.s).astype(np.int64)
    assert np.array_equal(np.diff(hoff.astype(np.int64)), cnt)
    owner = np.repeat(np.arange(npat), cnt)
    got = text[pos.astype(np.int64)[:, None] + np.arange(m)[None, :]]
    assert np.array_equal(got, pats[owner])
    sampled = np.arange(0, npat, 2)
    assert np.all(cnt[sampled] >= 1)
    # origin among the hits of each sampled pattern
    first = hoff[sampled].astype(np.int64)
    found = np.zeros(sampled.size, dtype=bool)
    for k in range(int(cnt[sampled].max())):
        ok = k < cnt[sampled]
        idx = np.where(ok, first + k, 0)
        found |= ok & (pos[idx].astype(np.int64) == starts[sampled])
    assert found.all()


@pytest.mark.parametrize("kind", [orc.RLFM, orc.MULTI])
def test_medium_repetitive_and_multi(kind):
    rng = np.random.default_rng(21)
    if kind == orc.RLFM:      # config 3 shape, scaled: mutated copies of one haplotype
        base = rng.integers(1, 5, 50_000, dtype=np.uint8)
        copies = []
        for _ in range(16):
            c = base.copy()
            mut = rng.random(c.size) < 0.001
            c[mut] = rng.integers(1, 5, int(mut.sum()), dtype=np.uint8)
            copies.append(c)
        text = np.append(np.concatenate(copies), np.uint8(0))
    else:                     # config 4 shape, scaled: 24 chromosome-like pieces
        lens = (rng.integers(20_000, 60_000, 24))
        text = np.concatenate([np.append(rng.integers(1, 5, int(l), dtype=np.uint8), np.uint8(0)) for l in lens])
    pats, starts = mixed_patterns(text, 20_000, 32, 5)
    index = KINDS[kind][1].new(fmx.Text.with_max_character(text, 4), 2)
    oracle = orc.OracleIndex(text, kind, level=2, max_character=4)
    b = index.search_batch(pats)
    s, e = oracle.search_batch(pats.reshape(-1), np.arange(20_001, dtype=np.uint64) * 32)
    assert np.array_equal(b.s, s) and np.array_equal(b.e, e)
    if kind == orc.MULTI:
        hoff, pos, pid = b.locate(piece_ids=True)
        ooff, opos, opid = oracle.locate_batch(s, e, want_piece_ids=False)
        assert np.array_equal(hoff, ooff) and np.array_equal(pos, opos)
        ends = np.flatnonzero(text == 0)
        assert np.array_equal(pid, np.searchsorted(ends, pos, side="left"))   # piece id = zeros before position
    else:
        hoff, pos = b.locate()
        ooff, opos, _ = oracle.locate_batch(s, e)
        assert np.array_equal(hoff, ooff) and np.array_equal(pos, opos)


# ---- config 5 shape, scaled: byte alphabet (L = 8), ragged pattern lengths 8..64
def test_byte_alphabet_ragged():
    rng = np.random.default_rng(80)
    n = 2_000_000
    text = np.append(rng.integers(1, 256, n, dtype=np.uint8), np.uint8(0))
    lens = rng.integers(8, 65, 50_000)
    starts = rng.integers(0, n - 64, 50_000)
    pats = []
    for k in range(50_000):
        pats.append(text[starts[k]:starts[k] + lens[k]].tobytes() if k & 1 else
                    rng.integers(1, 256, int(lens[k]), dtype=np.uint8).tobytes())
    flat, off = orc.pack_patterns(pats)
    index = fmx.FMIndexWithLocate.new(fmx.Text.new(text), 2)
    oracle = orc.OracleIndex(text, orc.FM, level=2)
    b = index.search_batch((flat, off))
    s, e = oracle.search_batch(flat, off)
    assert np.array_equal(b.s, s) and np.array_equal(b.e, e)
    hoff, pos = b.locate()
    ooff, opos, _ = oracle.locate_batch(s, e)
    assert np.array_equal(hoff, ooff) and np.array_equal(pos, opos)


@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
@pytest.mark.parametrize("level", [0, 2, 5])
def test_range_walk_locate_on_repetitive_text(kind, level):
    """k_locate_ranges (whole SA sub-ranges stepped by one lf_map2 pair while their BWT symbols agree) against
    the oracle's per-row walks: positions in the reference's order, piece ids, and the executed LF steps"""
    rng = np.random.default_rng(900 + kind + level)
    base = rng.integers(1, 5, 3000, dtype=np.uint8)
    copies = []
    for c in range(40):                                    # 40 mutated copies: every 12-mer has ~40 matches
        x = base.copy()
        mut = rng.random(x.size) < 0.01
        x[mut] = rng.integers(1, 5, int(mut.sum()), dtype=np.uint8)
        copies.append(x)
        if kind == orc.MULTI:
            copies.append(np.zeros(1, dtype=np.uint8))
    text = np.concatenate(copies + ([] if kind == orc.MULTI else [np.zeros(1, dtype=np.uint8)]))
    index = KINDS[kind][1].new(fmx.Text.with_max_character(text, 4), level)
    index.set_option("count_work", 1)
    oracle = orc.OracleIndex(text, kind, level=level, max_character=4)
    starts = rng.integers(0, base.size - 12, 300)
    pats = [bytes(base[p:p + int(m)]) for p, m in zip(starts, rng.integers(1, 13, 300))]
    if kind == orc.MULTI:
        pats = [p for p in pats if len(p) >= 3]        # the oracle's literal piece_id walk is slow: keep the hit count moderate
    else:
        pats += [b"", bytes([3])]
    flat, off = orc.pack_patterns(pats)
    s, e = oracle.search_batch(flat, off)
    want = oracle.locate_batch(s, e, want_piece_ids=kind == orc.MULTI)
    for mode in (1, 0):
        index.set_option("locate_ranges", mode)
        b = index.search_batch(pats)
        got = b.locate(piece_ids=kind == orc.MULTI)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), mode
        if kind == orc.MULTI:
            assert np.array_equal(got[2], want[2])
        if mode == 1:
            lf_ranges = index.last_work()[1]
        else:
            assert index.last_work()[1] == lf_ranges
        fused = index.search_locate_batch(pats, capacity=int(want[0][-1]) + 8)
        assert np.array_equal(fused[1], want[0]) and np.array_equal(fused[2], want[1])
        page, total = index.locate_page(b.s, b.e, 1000, 5000)
        assert total == int(want[0][-1]) and np.array_equal(page, want[1][1000:6000])


@pytest.mark.parametrize("m", [16, 32, 48, 64, 80, 20])
def test_fixed_length_patterns_staged_in_shared_memory(m):
    """fixed-length batches whose length is a multiple of 16 read their characters through a shared-memory
    tile (32-character windows, refilled for longer patterns); everything else through the word reader"""
    text = dna(300_000, 61)
    pats, _ = mixed_patterns(text, 20_000, m, 62)
    index = fmx.FMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 2)
    oracle = orc.OracleIndex(text, orc.FM, level=2, max_character=4)
    s, e = oracle.search_batch(pats.reshape(-1), np.arange(pats.shape[0] + 1, dtype=np.uint64) * m)
    for stage in (1, 0):
        index.set_option("stage_patterns", stage)
        for kmer in (1, 0):
            index.set_option("kmer", kmer)
            b = index.search_batch(pats)
            assert np.array_equal(b.s, s) and np.array_equal(b.e, e), (stage, kmer)


@pytest.mark.parametrize("mc,level,dense,kind", [(4, 2, True, orc.FM), (4, None, True, orc.FM), (255, 1, True, orc.FM),
                                                 (255, 3, False, orc.FM), (255, None, False, orc.FM),
                                                 (4, 2, True, orc.MULTI), (255, None, True, orc.MULTI)])
def test_seed_and_verify_tail(mc, level, dense, kind, monkeypatch):
    """one-row ranges finish by locate + text comparison + sampled-ISA jump (verify_tail): SA ranges and
    executed step counts equal the oracle's plain loop -- matches, mismatches anywhere in the tail, patterns
    running off the start of the text, invalid characters inside the compared region"""
    monkeypatch.setenv("FMX_VERIFY_MIN_RANK_MB", "0")    # by default only indexes beyond the L2 carry the structures
    if not dense:
        monkeypatch.setenv("FMX_VERIFY_BUDGET_MB", "0")   # sampled structures (texts beyond the dense budget, SYM layout)
    rng = np.random.default_rng(700 + mc + (level or 0))
    n = 50_000
    sigma = min(mc, 4)
    text = np.append(rng.integers(1, sigma + 1, n, dtype=np.uint8), np.uint8(0))
    if kind == orc.MULTI:
        text[rng.integers(5, n - 5, 40) // 2 * 2] = 0     # pieces: never two \0 in a row, none at either end
    cls = KINDS[kind][0 if level is None else 1]
    index = cls.new(fmx.Text.with_max_character(text, mc)) if level is None else cls.new(fmx.Text.with_max_character(text, mc), level)
    index.set_option("count_work", 1)
    oracle = orc.OracleIndex(text, kind, level=level, max_character=mc)
    pats = []
    for t in range(6000):
        m = int(rng.integers(10, 70)) if t % 7 else int(rng.integers(6, 14))
        p0 = int(rng.integers(0, n - m))
        pat = text[p0:p0 + m].copy()
        kind = t % 5
        if kind == 1:
            j = int(rng.integers(0, m))
            pat[j] = pat[j] % sigma + 1
        elif kind == 2:
            pat = np.concatenate([rng.integers(1, sigma + 1, 25, dtype=np.uint8), text[:m]])
        elif kind == 3:
            for j in rng.integers(0, m, 3):
                pat[int(j)] = rng.integers(1, sigma + 1)
        elif kind == 4 and mc == 255:
            pat[int(rng.integers(0, m))] = 0          # a \0 inside the pattern: valid character, never matches here
        pats.append(pat.tobytes())
    flat, off = orc.pack_patterns(pats)
    for mode in ((fmx.SEARCH, fmx.SEARCH_PREFIX, fmx.SEARCH_SUFFIX, fmx.SEARCH_EXACT) if kind == orc.MULTI else (fmx.SEARCH,)):
        s, e, steps = oracle.search_batch(flat, off, mode, want_steps=True)
        for verify, phased in ((1, 2), (1, 1), (1, 0), (0, 2)):   # fused kernel / phased kernels / verify tail inside k_search / plain loop
            index.set_option("verify", verify)
            index.set_option("search_phased", phased)
            b = index.search_batch(pats, mode)
            assert np.array_equal(b.s, s) and np.array_equal(b.e, e), (mode, verify, phased)
            assert index.last_work()[0] == int(steps.sum()), (mode, verify, phased)
    if mc == 4 and kind == orc.FM:                                     # invalid character deep inside an otherwise unique match
        bad = bytearray(text[1000:1040].tobytes())
        bad[5] = 9
        with pytest.raises(IndexError):
            index.search_batch([bytes(bad)])
        bad2 = bytearray(text[1000:1040].tobytes())
        bad2[5] = 9
        bad2[30] = bad2[30] % 4 + 1                   # ... but the range empties first: no error, as in the reference
        assert index.search_batch([bytes(bad2)]).count()[0] == 0


def test_paged_locate_equals_full_locate():
    """fmx_locate_page: the hit list in pages (bounded memory) equals the one-shot list, order included"""
    text = dna(200_000, 55)
    index = fmx.FMIndexMultiPiecesWithLocate.new(fmx.Text.with_max_character(text, 4), 2)
    pats = [bytes([1]), bytes([2, 3]), bytes([4, 4, 1]), bytes([1, 2, 3, 4, 1, 2, 3]), b""]
    b = index.search_batch(pats)
    hoff, pos, pid = b.locate(piece_ids=True)
    total = int(hoff[-1])
    assert total > 250_000
    got = np.concatenate([p for _, p in b.iter_locate_pages(30_011)])
    assert np.array_equal(got, pos)
    ppos, ppid, t = index.locate_page(b.s, b.e, 123_456, 1000, piece_ids=True)
    assert t == total and np.array_equal(ppos, pos[123_456:124_456]) and np.array_equal(ppid, pid[123_456:124_456])
    tail, t = index.locate_page(b.s, b.e, total - 5, 100)
    assert np.array_equal(tail, pos[-5:])
    none, t = index.locate_page(b.s, b.e, total + 7, 100)
    assert none.size == 0 and t == total


def test_huge_range_locate():
    """A pattern matching a large share of the text: one pattern, many hits (load balance path)."""
    text = dna(300_000, 77)
    index = fmx.FMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 3)
    oracle = orc.OracleIndex(text, orc.FM, level=3, max_character=4)
    b = index.search_batch([bytes([1]), bytes([2, 3]), b""])
    hoff, pos = b.locate()
    s, e = oracle.search_batch(*orc.pack_patterns([bytes([1]), bytes([2, 3]), b""]))
    ooff, opos, _ = oracle.locate_batch(s, e)
    assert np.array_equal(hoff, ooff) and np.array_equal(pos, opos)
    assert int(hoff[-1]) > 300_000 and sorted(pos[int(hoff[2]):]) == list(range(300_001))


@pytest.mark.skipif(os.environ.get("FMX_SKIP_FULL") == "1", reason="full-size run disabled")
def test_config2_full_size_properties():
    """BASELINE config 2 at full size (100 MB DNA, 1M 32-mers, level 2): size-independent
    properties (bench.py checks bit-exact parity with the oracle on a sample of this workload)."""
    n, npat = 100_000_000, 1_000_000
    text = dna(n, 3)
    pats, starts = mixed_patterns(text, npat, 32, 4)
    index = fmx.FMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 2)
    b = index.search_batch(pats)
    hoff, pos = b.locate()
    check_locate_properties(text, pats, starts, b, hoff, pos)
    # checksum of checksums: the batch split in two halves gives the same answers
    b1, b2 = index.search_batch(pats[: npat // 2]), index.search_batch(pats[npat // 2:])
    assert np.array_equal(np.concatenate([b1.s, b2.s]), b.s) and np.array_equal(np.concatenate([b1.e, b2.e]), b.e)


# ---- the fused, pipelined entry (fmx_search_locate_batch) must agree with the two-call path
@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
def test_fused_search_locate_matches_two_phase(kind):
    rng = np.random.default_rng(300 + kind)
    if kind == orc.MULTI:
        text = np.concatenate([np.append(rng.integers(1, 5, int(l), dtype=np.uint8), np.uint8(0))
                               for l in rng.integers(10_000, 30_000, 12)])
    else:
        text = dna(300_000, 31 + kind)
    npat = 50_000
    pats, _ = mixed_patterns(text, npat, 24, 9)
    index = KINDS[kind][1].new(fmx.Text.with_max_character(text, 4), 2)
    index.set_option("pipeline_chunk", 7_001)    # 8 ragged chunks: exercises both pipeline lanes
    modes = [fmx.SEARCH] + ([fmx.SEARCH_PREFIX, fmx.SEARCH_SUFFIX, fmx.SEARCH_EXACT] if kind == orc.MULTI else [])
    for mode in modes:
        ref = index.search_batch(pats, mode)
        if kind == orc.MULTI:
            rh, rp, rd = ref.locate(piece_ids=True)
            b, h, p, d = index.search_locate_batch(pats, mode, piece_ids=True)
            assert np.array_equal(d, rd)
        else:
            rh, rp = ref.locate()
            b, h, p = index.search_locate_batch(pats, mode)
        assert np.array_equal(b.s, ref.s) and np.array_equal(b.e, ref.e)
        assert np.array_equal(h, rh) and np.array_equal(p, rp)
    # ragged patterns + a capacity that is too small (falls back to exact-size buffers)
    lens = rng.integers(1, 40, 40_000)
    flat = rng.integers(1, 5, int(lens.sum()), dtype=np.uint8)
    off = np.zeros(lens.size + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    ref = index.search_batch((flat, off))
    rh, rp = ref.locate()
    b, h, p = index.search_locate_batch((flat, off), capacity=1000)
    assert np.array_equal(b.s, ref.s) and np.array_equal(h, rh) and np.array_equal(p, rp)
    b, h, p = index.search_locate_batch((flat, off), capacity=int(rh[-1]) + 5)
    assert np.array_equal(b.e, ref.e) and np.array_equal(h, rh) and np.array_equal(p, rp)


# ---- index construction on the GPU (gpu_sa.cu): same suffix array, byte-identical blob
def test_gpu_suffix_array_matches_oracle():
    rng = np.random.default_rng(404)
    cases = []
    for t in range(12):
        n = int(rng.integers(2, 5000))
        mc = int(rng.choice([1, 4, 7, 255]))
        cases.append((build_text(rng, n, min(mc + 1, 256) if t & 1 else max(2, min(mc, 255)), bool(t & 1)), mc))
    cases.append((bytes([3, 0, 0, 0, 2, 0, 0, 1, 0]), 3))                        # runs of \0 inside the text
    cases.append((bytes([1] * 3000 + [0]), 1))                                     # one long run: many doubling rounds
    rep = build_text(rng, 4000, 4, False)[:-1]
    cases.append((rep * 50 + b"\0", 4))                                            # highly repetitive
    cases.append((dna(400_000, 8).tobytes(), 4))
    cases.append((np.append(rng.integers(1, 256, 300_000, dtype=np.uint8), np.uint8(0)).tobytes(), 255))
    for text, mc in cases:
        mc = max(mc, max(text))
        sa, rounds = fmx.suffix_array_device(text, mc)
        assert np.array_equal(sa, orc.suffix_array(text)), (len(text), mc, rounds)


@pytest.mark.parametrize("mc", [4, 255])
@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
def test_gpu_built_blob_identical_to_host_built(kind, mc, tmp_path):
    """FM / MultiPieces indexes in the Q4 (mc = 4) and SYM (mc = 255) layouts are built entirely in device memory
    (gpu_build.cu): the saved index equals the host builder's blob byte for byte, in both modes"""
    rng = np.random.default_rng(500 + kind + mc)
    hi = 5 if mc == 4 else 256
    if kind == orc.MULTI:
        text = np.concatenate([np.append(rng.integers(1, hi, int(l), dtype=np.uint8), np.uint8(0))
                               for l in rng.integers(5_000, 40_000, 9)])
    else:
        text = np.append(rng.integers(1, hi, 200_000, dtype=np.uint8), np.uint8(0))
    t = fmx.Text.with_max_character(text, mc)
    for mode in (fmx.MODE_COMPACT, fmx.MODE_RICH):
        index = KINDS[kind][1].new(t, 2, mode=mode)      # >= 2^16 symbols: built on the GPU
        p = tmp_path / f"gpu_built_{mode}.fmx"
        index.save(p)
        host = fmx.blob_build(t, kind, 2, mode=mode)     # host SA-IS, host rank structures
        assert np.array_equal(np.fromfile(p, dtype=np.uint8), host)
        del index


def test_search_options_do_not_change_results():
    """the tuning knobs (persistent kernels, k-mer table off, bucketed order) are result-neutral"""
    text = dna(500_000, 91)
    pats, _ = mixed_patterns(text, 60_000, 28, 92)
    index = fmx.FMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 1)
    index.set_option("count_work", 1)
    ref = index.search_batch(pats)
    steps = index.last_work()[0]
    for key, val in (("search_persistent", 1), ("kmer", 0), ("search_persistent", 0), ("kmer", 1), ("bucket", 1)):
        index.set_option(key, val)
        b = index.search_batch(pats)
        assert np.array_equal(b.s, ref.s) and np.array_equal(b.e, ref.e), (key, val)
        assert index.last_work()[0] == steps
    hoff, pos = ref.locate()
    index.set_option("locate_refill", 0)                  # one hit per thread instead of per-lane refill
    hoff2, pos2 = index.search_batch(pats).locate()
    lf_simple = index.last_work()[1]
    for mode in (1, 2):                                   # unstable refill / stable warp-level compaction
        index.set_option("locate_refill", mode)
        hoff3, pos3 = index.search_batch(pats).locate()
        assert np.array_equal(hoff, hoff2) and np.array_equal(pos, pos2) and np.array_equal(pos, pos3), mode
        assert index.last_work()[1] == lf_simple
    index.set_option("locate_refill", 0)
    for mode in (1, 2, 0):                                # rows expanded by scans / found by binary search / auto
        index.set_option("locate_expand", mode)
        b = index.search_batch(pats)
        h4, p4 = b.locate()
        assert np.array_equal(hoff, h4) and np.array_equal(pos, p4), mode
        h5, p5 = index.search_locate_batch(pats)[1:3]
        assert np.array_equal(hoff, h5) and np.array_equal(pos, p5), mode
    for mode in (1, 0, -1):                               # sub-range walks instead of row walks / rows / auto
        index.set_option("locate_ranges", mode)
        h6, p6 = index.search_batch(pats).locate()
        assert np.array_equal(hoff, h6) and np.array_equal(pos, p6), mode
        assert index.last_work()[1] == lf_simple
    with pytest.raises(fmx.Error):
        index.set_option("no_such_option", 1)


def test_cpp_facade(tmp_path):
    """include/fmx.hpp (the compiled host-side mirror of the crate's API) against the reference's
    README doctest and multi_pieces example, as a C++ program linked to libfmx_b200.so"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "test_api"
    libdir = os.path.join(root, "fm-index_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "test_api.cpp"), "-o", str(exe),
                           "-L", libdir, "-lfmx_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "cpp facade ok" in out.stdout, out.stdout + out.stderr


def test_async_device_locate_matches_host_api():
    """fmx_locate_batch_device (no host synchronisation; the total stays on the device) against the
    host-buffer API, for an exact, a generous and a too-small capacity"""
    import ctypes as C
    import torch
    text = dna(400_000, 17)
    pats, _ = mixed_patterns(text, 30_000, 20, 18)
    index = fmx.FMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 2)
    ref = index.search_batch(pats)
    rh, rp = ref.locate()
    total = int(rh[-1])
    L, h = index._L, index._h
    d_pat = torch.from_numpy(pats).cuda()
    npat, m = pats.shape
    d_s = torch.empty(npat, dtype=torch.int64, device="cuda")
    d_e = torch.empty_like(d_s)
    d_off = torch.empty(npat + 1, dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    sp = C.c_void_p(st.cuda_stream)
    assert L.fmx_search_batch_device(h, 0, d_pat.data_ptr(), None, m, npat, None, None, d_s.data_ptr(), d_e.data_ptr(), sp) == 0
    for cap in (total, total + 5000, total // 2):
        d_pos = torch.full((max(cap, 1),), -1, dtype=torch.int64, device="cuda")
        assert L.fmx_locate_batch_device(h, 0, d_s.data_ptr(), d_e.data_ptr(), npat, d_off.data_ptr(), d_pos.data_ptr(),
                                         None, cap, sp) == 0, L.fmx_last_error()
        st.synchronize()
        assert np.array_equal(d_off.cpu().numpy().view(np.uint64), rh)
        k = min(cap, total)
        assert np.array_equal(d_pos[:k].cpu().numpy().view(np.uint64), rp[:k])
        if cap > total:
            assert bool((d_pos[total:] == -1).all())


# ---------------------------------------------------------------- round 2: HBM-rich mode, fused query, packed patterns

def _rich_case(kind, mc, seed, n=60_000, pieces=30):
    rng = np.random.default_rng(seed)
    sigma = min(mc, 4) if mc <= 4 else 6
    text = np.append(rng.integers(1, sigma + 1, n, dtype=np.uint8), np.uint8(0))
    if kind == orc.MULTI:
        text[rng.integers(5, n - 5, pieces) // 2 * 2] = 0
    pats = []
    for t in range(5000):
        m = int(rng.integers(1, 60))
        p0 = int(rng.integers(0, n - m))
        pat = text[p0:p0 + m].copy()
        v = t % 6
        if v == 1:
            j = int(rng.integers(0, m))
            pat[j] = pat[j] % sigma + 1
        elif v == 2:
            pat = np.concatenate([rng.integers(1, sigma + 1, 9, dtype=np.uint8), text[:m]])   # runs off the start of the text
        elif v == 3:
            pat = rng.integers(1, sigma + 1, m, dtype=np.uint8)
        elif v == 4 and kind == orc.MULTI:
            pat[int(rng.integers(0, m))] = 0
        pats.append(pat.tobytes())
    pats += [b"", bytes([1])]
    return text, pats


@pytest.mark.parametrize("kind", [orc.FM, orc.MULTI])
@pytest.mark.parametrize("mc,env", [(4, None), (255, None), (37, ("FMX_SYM_BUDGET_MB", "0")), (6, ("FMX_FORCE_WAVELET", "1"))])
def test_rich_mode_query_parity(kind, mc, env, monkeypatch):
    """FMX_MODE_RICH: phased search (table-embedded positions, verify queue, hints) + locate by the resident suffix
    array, through fmx_query_batch in every output combination, against the oracle: SA ranges, counts, ordered
    positions, piece ids, executed step counts; and against the same index with the rich paths switched off"""
    if env:
        monkeypatch.setenv(*env)
    text, pats = _rich_case(kind, mc, 4000 + kind + mc)
    index = KINDS[kind][1].new(fmx.Text.with_max_character(text, mc), 2, mode=fmx.MODE_RICH)
    assert index.mode() == fmx.MODE_RICH
    index.set_option("count_work", 1)
    oracle = orc.OracleIndex(text, kind, level=2, max_character=mc)
    flat, off = orc.pack_patterns(pats)
    modes = [fmx.SEARCH] + ([fmx.SEARCH_PREFIX, fmx.SEARCH_SUFFIX, fmx.SEARCH_EXACT] if kind == orc.MULTI else [])
    for mode in modes:
        po = mode in (fmx.SEARCH_PREFIX, fmx.SEARCH_EXACT)
        s, e, steps = oracle.search_batch(flat, off, mode, want_steps=True)
        ooff, opos, _ = oracle.locate_batch(s, e, prefix_only=po)
        opid = None
        if kind == orc.MULTI:
            # piece of a position = zeros before it (multi_pieces.rs:287-296).  The oracle's literal walk to the next \0
            # costs piece-length LF steps per hit (40 s for this batch): it pins 300 patterns with few matches here
            # and every hit in the smaller differential tests
            opid = np.searchsorted(np.nonzero(text == 0)[0], opos, side="left").astype(np.uint64)
            sel = np.nonzero(e - np.minimum(s, e) <= 64)[0][:300]
            woff, wpos, wpid = oracle.locate_batch(s[sel], e[sel], prefix_only=po, want_piece_ids=True)
            seg = np.concatenate([np.arange(int(ooff[p]), int(ooff[p + 1])) for p in sel] + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
            assert np.array_equal(wpos, opos[seg]) and np.array_equal(wpid, opid[seg])
        cnt = np.where(e > s, e - s, 0)
        b = index.search_batch(pats, mode)                      # phased, rows wanted
        assert np.array_equal(b.s, s) and np.array_equal(b.e, e), mode
        assert index.last_work()[0] == int(steps.sum())
        got = b.locate(piece_ids=kind == orc.MULTI)             # two-call path: dense k_locate_simple
        assert np.array_equal(got[0], ooff) and np.array_equal(got[1], opos)
        for rows, width, phased, fused in ((True, 8, 2, 0), (False, 8, 2, 1), (True, 4, 2, 1), (False, 4, 2, 0), (True, 8, 1, 1), (False, 4, 1, 0)):
            if True:
                index.set_option("search_phased", phased)       # 2: one fused kernel (default); 1: seed / steps / verify kernels
                index.set_option("emit_fused", fused)           # 1: hit offsets + positions in one pass over the ranges (k_offsets_emit)
                r = index.query_batch(pats, mode, rows=rows, counts=True, piece_ids=kind == orc.MULTI, width=width)
                if rows:
                    assert np.array_equal(r["s"], s) and np.array_equal(r["e"], e)
                assert np.array_equal(r["counts"].astype(np.uint64), cnt), (mode, rows, width)
                assert np.array_equal(r["hit_off"].astype(np.uint64), ooff), (mode, rows, width)
                assert np.array_equal(r["positions"].astype(np.uint64), opos), (mode, rows, width)
                if kind == orc.MULTI:
                    assert np.array_equal(r["piece_ids"].astype(np.uint64), opid)
        for phased in (1, 2):
            index.set_option("search_phased", phased)
            r = index.query_batch(pats, mode, locate=False, counts=True)       # count only: no ISA request, no locate
            assert np.array_equal(r["counts"], cnt)
            if not po:
                assert index.last_work()[0] == int(steps.sum()), phased        # the shortcut counts the reference's last iteration too
    # the rich paths off: same answers from the plain kernels
    ref = index.query_batch(pats, rows=True, counts=True)
    for key in ("search_phased", "locate_dense", "verify"):
        index.set_option(key, 0)
        r = index.query_batch(pats, rows=True, counts=True)
        for k in ("s", "e", "counts", "hit_off", "positions"):
            assert np.array_equal(r[k], ref[k]), (key, k)
    for key, v in (("search_phased", 2), ("locate_dense", 1), ("verify", 1)):
        index.set_option(key, v)
    if mc < 255:                                          # invalid characters: first processed, and deep inside a unique match
        bad = bytearray(text[1000:1040].tobytes())
        bad[5] = mc + 1
        for rows in (True, False):
            for pat in (bytes([1, 2, mc + 1]), bytes(bad)):
                with pytest.raises(IndexError):
                    index.query_batch([pat], rows=rows)
        assert index.query_batch([bytes([1])], locate=False, counts=True)["counts"][0] > 0   # still usable


@pytest.mark.parametrize("kind", [orc.FM, orc.MULTI])
def test_table_entries_with_text_context(kind):
    """The large k-mer table of HBM-rich DNA indexes has 16-byte entries whose one-row ranges carry the 16 text characters
    in front of the row (SearchArgs::big_tab4): patterns with <= 16 characters left after the table are finished without a
    text request.  Same ranges, counts, ordered positions, piece ids and executed step counts as the oracle -- with the
    context entries (default), and with 8-byte entries (option table_ctx = 0) -- for pattern lengths on both sides of
    k and k + 16, mismatches at every depth, patterns that run off the start of the text or across a piece boundary,
    and invalid characters."""
    rng = np.random.default_rng(900 + kind)
    n = 4_300_000                                               # >= 2^22: the large table exists
    text = np.append(rng.integers(1, 5, n, dtype=np.uint8), np.uint8(0))
    if kind == orc.MULTI:
        text[rng.integers(40, n - 40, 300) // 2 * 2] = 0
    index = KINDS[kind][1].new(fmx.Text.with_max_character(text, 4), 2, mode=fmx.MODE_RICH)
    index.set_option("kmer_budget_mb", 200)                     # k = 12: 16.7 M entries, 0.26 occurrences per 12-mer
    assert index.kmer_k(True) == 12
    oracle = _oracle_from_gpu_sa(text, kind, 2, 4)
    pats = []
    for t in range(60_000):
        m = int(rng.integers(9, 50))                              # >= 9: a few dozen matches at most (the oracle locates every one)
        p0 = int(rng.integers(0, n - m))
        pat = text[p0:p0 + m].copy()
        v = t % 8
        if v in (1, 5):                                          # one mismatch, anywhere
            j = int(rng.integers(0, m))
            pat[j] = pat[j] % 4 + 1
        elif v == 2:                                             # runs off the start of the text
            pat = np.concatenate([rng.integers(1, 5, int(rng.integers(1, 20)), dtype=np.uint8), text[:m]])
        elif v == 3:
            pat = rng.integers(1, 5, m, dtype=np.uint8)
        elif v == 4 and kind == orc.MULTI and m > 13:            # a \0 inside the part the context would cover
            pat[int(rng.integers(0, m - 12))] = 0
        pats.append(pat.tobytes())
    pats += [text[:30].tobytes(), text[:12].tobytes(), text[1:29].tobytes()]
    flat, off = orc.pack_patterns(pats)
    zeros = np.nonzero(text == 0)[0]
    modes = [fmx.SEARCH] + ([fmx.SEARCH_PREFIX] if kind == orc.MULTI else [])
    index.set_option("count_work", 1)
    for ctx in (1, 0):
        index.set_option("table_ctx", ctx)
        assert index.kmer_k(True) == 12
        for mode in modes:
            po = mode == fmx.SEARCH_PREFIX
            s, e, steps = oracle.search_batch(flat, off, mode, want_steps=True)
            ooff, opos, _ = oracle.locate_batch(s, e, prefix_only=po)
            # piece of a position = zeros before it (multi_pieces.rs:287-296; the literal walk is pinned in other tests and
            # would take piece-length LF steps per hit here)
            opid = np.searchsorted(zeros, opos, side="left").astype(np.uint64)
            # 1: block-local second pass for the patterns the head does not finish; 3 .. 7: its occupancy / queue variants
            # (60 k patterns are one round of 235 blocks; 7 blocks make it 34 rounds: queue drains and leftovers as at scale)
            for rows, defer, nblocks in ((True, 1, 0), (False, 1, 7), (True, 0, 0), (False, 0, 7), (True, 3, 7), (False, 5, 7), (True, 6, 7),
                                        (False, 7, 7), (True, 7, 3)):
                index.set_option("fused_defer", defer)
                index.set_option("query_blocks", nblocks)
                r = index.query_batch(pats, mode, rows=rows, counts=True, piece_ids=kind == orc.MULTI, capacity=int(ooff[-1]) + 8)
                if rows:
                    assert np.array_equal(r["s"], s) and np.array_equal(r["e"], e), (ctx, mode, defer)
                assert np.array_equal(r["counts"], np.where(e > s, e - s, 0)), (ctx, mode, rows, defer)
                assert np.array_equal(r["hit_off"], ooff) and np.array_equal(r["positions"], opos), (ctx, mode, rows, defer)
                if kind == orc.MULTI:
                    assert np.array_equal(r["piece_ids"], opid)
                if not po:
                    assert index.last_work()[0] == int(steps.sum()), (ctx, mode, rows, defer)
            index.set_option("fused_defer", 1)
            index.set_option("query_blocks", 0)
        bad = bytearray(text[1000:1028].tobytes())               # an invalid character inside the context window
        bad[5] = 5
        for rows in (True, False):
            if 0 not in bad:
                with pytest.raises(IndexError):
                    index.query_batch([bytes(bad)], rows=rows)
    fixed = np.stack([text[i:i + 28] for i in rng.integers(0, n - 28, 20_000)])
    fixed = fixed[~(fixed == 0).any(axis=1)]
    ref = index.query_batch(fixed, rows=True, counts=True)
    index.set_option("table_ctx", 1)
    packed = fmx.pack_patterns(fixed, 2)
    r = index.query_batch(packed, rows=True, counts=True, packed_bits=2, fixed_len=28)   # the packed reader through the same path
    for k in ("s", "e", "counts", "hit_off", "positions"):
        assert np.array_equal(r[k], ref[k]), k


@pytest.mark.parametrize("kind,mc", [(orc.FM, 4), (orc.FM, 255), (orc.MULTI, 4)])
def test_ragged_batches_in_order_of_length(kind, mc):
    """option order_by_length: the fused kernel visits a ragged batch sorted by pattern length (k_len_hist / k_len_scatter);
    results stay in input order and equal the oracle's, lengths beyond the last bucket included"""
    text, pats = _rich_case(kind, mc, 7000 + kind + mc)
    rng = np.random.default_rng(3)
    pats = pats + [text[i:i + int(m)].tobytes() for i, m in zip(rng.integers(0, 50_000, 40), rng.integers(120, 400, 40))]
    index = KINDS[kind][1].new(fmx.Text.with_max_character(text, mc), 2, mode=fmx.MODE_RICH)
    oracle = orc.OracleIndex(text, kind, level=2, max_character=mc)
    flat, off = orc.pack_patterns(pats)
    s, e = oracle.search_batch(flat, off)
    ooff, opos, _ = oracle.locate_batch(s, e)
    for opt in (1, 0):
        index.set_option("order_by_length", opt)
        for rows in (True, False):
            r = index.query_batch(pats, rows=rows, counts=True)
            if rows:
                assert np.array_equal(r["s"], s) and np.array_equal(r["e"], e), opt
            assert np.array_equal(r["counts"], np.where(e > s, e - s, 0)), opt
            assert np.array_equal(r["hit_off"], ooff) and np.array_equal(r["positions"], opos), opt


def test_rich_mode_many_hits_and_pipeline():
    """patterns with thousands of matches (k_emit_big), a chunked pipeline with a position-capacity guess that is too
    small for the first chunks, a too-small caller capacity"""
    text = dna(300_000, 123)
    index = fmx.FMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 2, mode=fmx.MODE_RICH)
    oracle = orc.OracleIndex(text, orc.FM, level=2, max_character=4)
    rng = np.random.default_rng(5)
    pats = [bytes(rng.integers(1, 5, int(m), dtype=np.uint8)) for m in rng.integers(1, 9, 30_000)] + [b""]
    flat, off = orc.pack_patterns(pats)
    s, e = oracle.search_batch(flat, off)
    ooff, opos, _ = oracle.locate_batch(s, e)
    index.set_option("pipeline_chunk", 4_001)
    for width, fused in ((8, 0), (4, 0), (8, 1), (4, 1)):        # emit_fused: k_offsets_emit instead of scan apply + k_emit_small
        index.set_option("emit_fused", fused)
        r = index.query_batch(pats, width=width, capacity=int(ooff[-1]) + 3)
        assert r["total"] == int(ooff[-1])
        assert np.array_equal(r["hit_off"].astype(np.uint64), ooff) and np.array_equal(r["positions"].astype(np.uint64), opos)
        with pytest.raises(fmx.Error, match="too small"):
            index.query_batch(pats, width=width, capacity=int(ooff[-1]) - 5)
    r = index.query_batch(pats)                                  # default capacity is too small: retried with the exact size
    assert np.array_equal(r["positions"], opos)
    b, h, p = index.search_locate_batch(pats, capacity=int(ooff[-1]) + 3)
    assert np.array_equal(b.s, s) and np.array_equal(h, ooff) and np.array_equal(p, opos)


@pytest.mark.parametrize("mode", [fmx.MODE_RICH, fmx.MODE_COMPACT])
@pytest.mark.parametrize("m", [12, 32, 33, 70])
def test_packed_patterns(mode, m):
    """2-bit packed patterns (fmx_query.packed_bits = 2): same answers as the byte form, on rich and compact indexes"""
    text = dna(200_000, 77 + m)
    pats, _ = mixed_patterns(text, 20_000, m, 78)
    index = fmx.FMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 2, mode=mode)
    assert index.mode() == mode
    ref = index.query_batch(pats, rows=True, counts=True)
    packed = fmx.pack_patterns(pats, 2)
    assert packed.size == pats.shape[0] * ((2 * m + 63) // 64)
    for width in (8, 4):
        r = index.query_batch(packed, rows=True, counts=True, width=width, packed_bits=2, fixed_len=m)
        for k in ("s", "e", "counts", "hit_off", "positions"):
            assert np.array_equal(r[k].astype(np.uint64), ref[k]), (k, width)
    r = index.query_batch(packed, packed_bits=2, fixed_len=m)            # no rows: hints only
    assert np.array_equal(r["positions"], ref["positions"])
    oracle = orc.OracleIndex(text, orc.FM, level=2, max_character=4)
    s, e = oracle.search_batch(pats.reshape(-1), np.arange(pats.shape[0] + 1, dtype=np.uint64) * m)
    ooff, opos, _ = oracle.locate_batch(s, e)
    assert np.array_equal(ref["s"], s) and np.array_equal(ref["e"], e) and np.array_equal(ref["positions"], opos)
    with pytest.raises(fmx.Error):                                        # 4 < max_character: cannot be packed in 2 bits
        fmx.FMIndex.new(fmx.Text.with_max_character(text, 9)).query_batch(packed, packed_bits=2, fixed_len=m, locate=False, counts=True)


def test_rlfm_rich_mode_locates_from_the_suffix_array():
    rng = np.random.default_rng(31)
    base = rng.integers(1, 5, 5000, dtype=np.uint8)
    copies = []
    for _ in range(20):
        x = base.copy()
        mut = rng.random(x.size) < 0.005
        x[mut] = rng.integers(1, 5, int(mut.sum()), dtype=np.uint8)
        copies.append(x)
    text = np.append(np.concatenate(copies), np.uint8(0))
    index = fmx.RLFMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 3, mode=fmx.MODE_RICH)
    assert index.mode() == fmx.MODE_RICH
    oracle = orc.OracleIndex(text, orc.RLFM, level=3, max_character=4)
    pats = [bytes(base[p:p + int(m)]) for p, m in zip(rng.integers(0, 4900, 2000), rng.integers(1, 40, 2000))]
    flat, off = orc.pack_patterns(pats)
    s, e = oracle.search_batch(flat, off)
    ooff, opos, _ = oracle.locate_batch(s, e)
    r = index.query_batch(pats, rows=True)
    assert np.array_equal(r["s"], s) and np.array_equal(r["hit_off"], ooff) and np.array_equal(r["positions"], opos)
    h, p = index.search_batch(pats).locate()
    assert np.array_equal(h, ooff) and np.array_equal(p, opos)
    index.set_option("locate_dense", 0)
    h, p = index.search_batch(pats).locate()
    assert np.array_equal(p, opos)


@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM])
def test_single_text_with_interior_zeros(kind):
    """the reference accepts single-text indexes over texts with interior \\0 (sais.rs:128-139); its
    lf_map2(0, i) is a plain rank there, which is not the true LF row.  SA ranges of patterns with and without \\0
    must equal the oracle's in every mode the index can be asked for."""
    rng = np.random.default_rng(17 + kind)
    n = 30_000
    text = np.append(rng.integers(1, 5, n, dtype=np.uint8), np.uint8(0))
    text[rng.integers(5, n - 5, 60) // 2 * 2] = 0
    oracle = orc.OracleIndex(text, kind, max_character=4)
    pats = []
    for t in range(4000):
        m = int(rng.integers(1, 50))
        p0 = int(rng.integers(0, n - m))
        pat = text[p0:p0 + m].copy()
        if t % 3 == 1:
            pat[int(rng.integers(0, m))] = rng.integers(0, 5)
        pats.append(pat.tobytes())
    flat, off = orc.pack_patterns(pats)
    s, e = oracle.search_batch(flat, off)
    for mode in (fmx.MODE_AUTO, fmx.MODE_RICH, fmx.MODE_COMPACT):
        index = KINDS[kind][0].new(fmx.Text.with_max_character(text, 4), mode=mode)
        assert index.mode() == fmx.MODE_COMPACT          # nothing that answers from the suffix array is built
        b = index.search_batch(pats)
        assert np.array_equal(b.s, s) and np.array_equal(b.e, e), mode


def test_two_phase_device_locate_rejects_a_mismatched_fill():
    import ctypes as C
    import torch
    text = dna(100_000, 3)
    pats, _ = mixed_patterns(text, 5000, 16, 4)
    index = fmx.FMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 2)
    L, h = index._L, index._h
    d_pat = torch.from_numpy(pats).cuda()
    npat = pats.shape[0]
    d_s = torch.empty(npat, dtype=torch.int64, device="cuda")
    d_e = torch.empty_like(d_s)
    d_off = torch.empty(npat + 1, dtype=torch.int64, device="cuda")
    assert L.fmx_search_batch_device(h, 0, d_pat.data_ptr(), None, 16, npat, None, None, d_s.data_ptr(), d_e.data_ptr(), None) == 0
    total = C.c_uint64(0)
    assert L.fmx_locate_count_device(h, 0, d_s.data_ptr(), d_e.data_ptr(), npat, d_off.data_ptr(), C.byref(total), None) == 0
    d_pos = torch.empty(total.value + 1, dtype=torch.int64, device="cuda")
    assert L.fmx_locate_fill_device(h, 0, d_s.data_ptr(), d_e.data_ptr(), npat - 1, d_off.data_ptr(), total.value, d_pos.data_ptr(), None, None) == -2
    assert L.fmx_locate_fill_device(h, 0, d_s.data_ptr(), d_e.data_ptr(), npat, d_off.data_ptr(), total.value, d_pos.data_ptr(), None, None) == 0
    torch.cuda.synchronize()
    rh, rp = index.search_batch(pats).locate()
    assert np.array_equal(d_pos[: total.value].cpu().numpy().view(np.uint64), rp)


# ---------------------------------------------------------------- at-scale code paths
# The paths that only switch on at scale -- the k = 16 table sized by sigma^k >= 4n, the default-on dense verify
# structures (rank structure >= 192 MB), SYM at tens of GB, beyond-L2 behaviour -- compared with the oracle by the
# driver's own pytest run.  The oracle's single-threaded SA-IS would take minutes here, so it is handed the
# GPU-built suffix array, which it first VERIFIES with its own linear-time checker (tests/test_oracle_golden.py::
# test_suffix_array_checker_and_build_from_sa pins that checker).

def _oracle_from_gpu_sa(text, kind, level, mc):
    sa, _ = fmx.suffix_array_device(text, mc)
    return orc.OracleIndex(text, kind, level=level, max_character=mc, sa=sa)


def _scale_check(index, oracle, text, pats, starts, m, sample=100_000, packed=False):
    npat = pats.shape[0]
    r = index.query_batch(pats, rows=True, counts=True, capacity=4 * npat)
    flat, off = pats[:sample].reshape(-1), np.arange(sample + 1, dtype=np.uint64) * m
    s, e, steps = oracle.search_batch(flat, off, want_steps=True)
    ooff, opos, _ = oracle.locate_batch(s, e)
    assert np.array_equal(r["s"][:sample], s) and np.array_equal(r["e"][:sample], e)
    assert np.array_equal(r["hit_off"][: sample + 1], ooff) and np.array_equal(r["positions"][: int(ooff[-1])], opos)
    # size-independent properties on the whole batch
    class B:
        pass
    b = B()
    b.s, b.e = r["s"], r["e"]
    check_locate_properties(text, pats, starts, b, r["hit_off"], r["positions"])
    # the fast form (no rows: hints, no inverse-suffix-array request) and 32-bit outputs give the same lists
    r2 = index.query_batch(pats, width=4, capacity=4 * npat)
    assert np.array_equal(r2["hit_off"].astype(np.uint64), r["hit_off"]) and np.array_equal(r2["positions"].astype(np.uint64), r["positions"])
    if packed:
        r3 = index.query_batch(fmx.pack_patterns(pats, 2), packed_bits=2, fixed_len=m, width=4, capacity=4 * npat)
        assert np.array_equal(r3["positions"], r2["positions"])
    # executed iterations on the sample: the table's early-break step counts included
    index.set_option("count_work", 1)
    index.set_option("pipeline_chunk", sample)                 # one chunk: the host call counts work
    for rows in (True, False):
        index.query_batch(pats[:sample], rows=rows, locate=False, counts=True)
        assert index.last_work()[0] == int(steps.sum()), rows
    index.set_option("pipeline_chunk", 0)
    index.set_option("count_work", 0)
    return r


@pytest.mark.skipif(os.environ.get("FMX_SKIP_FULL") == "1", reason="full-size run disabled")
def test_scale_dna_420m_rich_and_compact():
    n, npat, m = 420_000_000, 1_000_000, 32           # rank structure 200 MiB: past the 192 MiB threshold of FMX_MODE_AUTO
    text = dna(n, 1003)
    pats, starts = mixed_patterns(text, npat, m, 1004)
    index = fmx.FMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 2)
    assert index.mode() == fmx.MODE_RICH and index.kmer_k(True) == 16        # default-on at this size
    oracle = _oracle_from_gpu_sa(text, orc.FM, 2, 4)
    r = _scale_check(index, oracle, text, pats, starts, m, packed=True)
    for key in ("search_phased", "locate_dense"):                             # the round-1 kernels on the same index
        index.set_option(key, 0)
    r0 = index.query_batch(pats, rows=True, capacity=4 * npat)
    for k in ("s", "e", "hit_off", "positions"):
        assert np.array_equal(r0[k], r[k]), k
    del index
    compact = fmx.FMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 2, mode=fmx.MODE_COMPACT)
    assert compact.mode() == fmx.MODE_COMPACT and compact.heap_size() < 2 * n
    rc = compact.query_batch(pats, rows=True, capacity=4 * npat)
    for k in ("s", "e", "hit_off", "positions"):
        assert np.array_equal(rc[k], r[k]), k


@pytest.mark.skipif(os.environ.get("FMX_SKIP_FULL") == "1", reason="full-size run disabled")
def test_scale_bytes_64m_sym():
    n, npat, m = 64_000_000, 400_000, 24
    rng = np.random.default_rng(1005)
    text = np.append(rng.integers(1, 256, n, dtype=np.uint8), np.uint8(0))
    pats, starts = mixed_patterns(text, npat, m, 1006, sigma=255)
    index = fmx.FMIndexWithLocate.new(fmx.Text.new(text), 2)
    assert index.sectors_per_rank() == 1 and index.mode() == fmx.MODE_RICH    # SYM (2.7 GB) + dense verify structures
    oracle = _oracle_from_gpu_sa(text, orc.FM, 2, 255)
    _scale_check(index, oracle, text, pats, starts, m)


@pytest.mark.skipif(os.environ.get("FMX_SKIP_FULL") == "1", reason="full-size run disabled")
def test_scale_rlfm_256m_16_copies():
    rng = np.random.default_rng(1007)
    base = rng.integers(1, 5, 16 << 20, dtype=np.uint8)
    parts = []
    for _ in range(16):
        x = base.copy()
        mut = rng.random(x.size) < 0.001
        x[mut] = rng.integers(1, 5, int(mut.sum()), dtype=np.uint8)
        parts.append(x)
    text = np.append(np.concatenate(parts), np.uint8(0))
    npat, m = 200_000, 32
    starts = rng.integers(0, text.size - 1 - m, npat)
    pats = text[starts[:, None] + np.arange(m)[None, :]]
    oracle = _oracle_from_gpu_sa(text, orc.RLFM, 2, 4)
    flat, off = pats[:50_000].reshape(-1), np.arange(50_001, dtype=np.uint64) * m
    s, e = oracle.search_batch(flat, off)
    ooff, opos, _ = oracle.locate_batch(s, e)
    for mode in (fmx.MODE_COMPACT, fmx.MODE_AUTO):
        index = fmx.RLFMIndexWithLocate.new(fmx.Text.with_max_character(text, 4), 2, mode=mode)
        # AUTO keeps the suffix array resident from 2^27 symbols on (locate = SA[row]); COMPACT walks to the samples
        assert index.mode() == (fmx.MODE_COMPACT if mode == fmx.MODE_COMPACT else fmx.MODE_RICH)
        r = index.query_batch(pats, rows=True, capacity=40 * npat)
        assert np.array_equal(r["s"][:50_000], s) and np.array_equal(r["e"][:50_000], e)
        assert np.array_equal(r["hit_off"][:50_001], ooff) and np.array_equal(r["positions"][: int(ooff[-1])], opos)
        cnt = (r["e"] - r["s"]).astype(np.int64)
        assert cnt.min() >= 1 and np.array_equal(np.diff(r["hit_off"].astype(np.int64)), cnt)
        owner = np.repeat(np.arange(npat), cnt)
        sel = rng.integers(0, owner.size, 300_000)
        got = text[r["positions"][sel].astype(np.int64)[:, None] + np.arange(m)[None, :]]
        assert np.array_equal(got, pats[owner[sel]])
        del index


# ---------------------------------------------------------------- multi-GPU forms (config 4)

def _multi_case(seed, pieces=9, lo=3_000, hi=20_000, npat=20_000, m=14):
    rng = np.random.default_rng(seed)
    text = np.concatenate([np.append(rng.integers(1, 5, int(l), dtype=np.uint8), np.uint8(0)) for l in rng.integers(lo, hi, pieces)])
    starts = rng.integers(0, text.size - m - 1, npat)
    pats = rng.integers(1, 5, (npat, m), dtype=np.uint8)
    samp = text[starts[:, None] + np.arange(m)[None, :]]
    pats[::2] = samp[::2]
    pats = pats[(pats != 0).all(axis=1)]
    pats[5, :6] = pats[5, 6:12]                       # a few short-period patterns: many hits in some pieces
    return text, np.ascontiguousarray(pats)


def _same_match_sets(hoff, pos, pid, rh, rp, rd):
    """equal CSR offsets and, per pattern, equal multisets of (position, piece id)"""
    assert np.array_equal(np.asarray(hoff, dtype=np.uint64), np.asarray(rh, dtype=np.uint64))
    owner = np.repeat(np.arange(len(rh) - 1), np.diff(np.asarray(rh, dtype=np.int64)))
    a = np.lexsort((np.asarray(pos, dtype=np.int64), owner))
    b = np.lexsort((np.asarray(rp, dtype=np.int64), owner))
    assert np.array_equal(np.asarray(pos)[a], np.asarray(rp)[b]) and np.array_equal(np.asarray(pid)[a], np.asarray(rd)[b])


@pytest.mark.parametrize("ndev", [1, 3])
@pytest.mark.parametrize("imode", [fmx.MODE_COMPACT, fmx.MODE_RICH])
def test_group_by_piece_and_replicated(ndev, imode):
    """fmx_group on ONE physical GPU used `ndev` times (the code path is the same: per-member streams and buffers,
    peer copies to the first member, fmx_csr_merge_device): BY_PIECE gives the replicated index's match sets, counts,
    positions and piece ids; REPLICATE gives its exact lists, order included"""
    text, pats = _multi_case(77 + ndev)
    t = fmx.Text.with_max_character(text, 4)
    single = fmx.FMIndexMultiPiecesWithLocate.new(t, 2, mode=imode)
    oracle = orc.OracleIndex(text, orc.MULTI, level=2, max_character=4)
    flat, off = pats.reshape(-1), np.arange(pats.shape[0] + 1, dtype=np.uint64) * pats.shape[1]
    for mode in (fmx.SEARCH, fmx.SEARCH_PREFIX, fmx.SEARCH_SUFFIX, fmx.SEARCH_EXACT):
        s, e = oracle.search_batch(flat, off, mode)
        rh, rp, rd = oracle.locate_batch(s, e, prefix_only=mode in (1, 3), want_piece_ids=True)
        ref = single.query_batch(pats, mode, piece_ids=True, capacity=int(rh[-1]) + 16)
        assert np.array_equal(ref["hit_off"], rh) and np.array_equal(ref["positions"], rp) and np.array_equal(ref["piece_ids"], rd)
        bp = fmx.IndexGroup(t, fmx.KIND_MULTI, 2, [0] * ndev, fmx.GROUP_BY_PIECE, mode=imode)
        assert bp.size() == ndev and bp.len() == text.size and bp.pieces_count() == int((text == 0).sum())
        r = bp.query_batch(pats, mode, piece_ids=True, counts=mode in (0, 2))
        _same_match_sets(r["hit_off"], r["positions"], r["piece_ids"], rh, rp, rd)
        if mode in (0, 2):
            assert np.array_equal(r["counts"], np.where(e > s, e - s, 0))
        if int(rh[-1]) > 5:
            with pytest.raises(fmx.Error, match="too small"):      # an explicit capacity that is too small is an error ..
                bp.query_batch(pats, mode, capacity=5)
        r = bp.query_batch(pats[:200], mode, capacity=None)        # .. the default one is retried with the exact size
        assert r["total"] == int(rh[200])
        del bp
    bad = pats.copy()
    bad[3, 2] = 0
    bp = fmx.IndexGroup(t, fmx.KIND_MULTI, 2, [0] * ndev, fmx.GROUP_BY_PIECE, mode=imode)
    with pytest.raises(fmx.Error):
        bp.query_batch(bad)
    del bp
    rep = fmx.IndexGroup(t, fmx.KIND_MULTI, 2, [0] * ndev, fmx.GROUP_REPLICATE, mode=imode)
    s, e = oracle.search_batch(flat, off)
    rh, rp, rd = oracle.locate_batch(s, e, want_piece_ids=True)
    for width in (8, 4):
        r = rep.query_batch(pats, rows=True, counts=True, piece_ids=True, width=width)
        assert np.array_equal(r["s"], s) and np.array_equal(r["e"], e) and np.array_equal(r["counts"].astype(np.uint64), e - s)
        assert np.array_equal(r["hit_off"].astype(np.uint64), rh) and np.array_equal(r["positions"].astype(np.uint64), rp)
        assert np.array_equal(r["piece_ids"].astype(np.uint64), rd)
    lens = np.random.default_rng(3).integers(1, 30, 5_000)         # ragged patterns through the sharded form
    fl = np.random.default_rng(4).integers(1, 5, int(lens.sum()), dtype=np.uint8)
    of = np.zeros(lens.size + 1, dtype=np.uint64)
    of[1:] = np.cumsum(lens)
    a, b = rep.query_batch((fl, of), rows=True), single.query_batch((fl, of), rows=True)
    for k in ("s", "e", "hit_off", "positions"):
        assert np.array_equal(a[k], b[k]), k
    with pytest.raises(fmx.Error):
        fmx.IndexGroup(fmx.Text.with_max_character(dna(5000, 1), 4), fmx.KIND_FM, 2, [0], fmx.GROUP_BY_PIECE)


@pytest.mark.skipif(os.environ.get("FMX_SKIP_FULL") == "1", reason="full-size run disabled")
def test_group_by_piece_over_a_text_beyond_2_32_symbols():
    """The reference's positions are usize (multi_pieces.rs:188-223); one fmx index addresses rows with u32.  A
    MultiPieces text of 4.4e9 symbols is served as a BY_PIECE group whose partitions (here two, both on device 0) stay
    below 2^32 each: hit offsets, positions (beyond 2^32) and piece ids come back as u64 in the coordinates of the
    WHOLE text.  32-mers cut from random DNA of this size occur once (4^32 >> n), so the expected answer of a
    sampled pattern is exactly its origin, and of a random one (almost surely) nothing; every hit is checked
    against the text."""
    pieces, plen, npat, m = 8, 550_000_000, 200_000, 32
    n = pieces * plen
    assert n > (1 << 32)
    import torch
    rng = np.random.default_rng(2032)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(2032)
    text = np.empty(n, dtype=np.uint8)                         # random DNA, generated on the GPU piece by piece
    for k in range(pieces):
        text[k * plen:(k + 1) * plen] = torch.randint(1, 5, (plen,), dtype=torch.uint8, device="cuda", generator=gen).cpu().numpy()
    text[plen - 1::plen] = 0
    torch.cuda.empty_cache()
    starts = rng.integers(0, n - m - 1, npat)
    starts[:8:2] = [0, n - m - 1, plen, (1 << 32) - 7]        # first / last piece, a piece start, across the u32 wall
    pats = rng.integers(1, 5, (npat, m), dtype=np.uint8)
    sampled = text[starts[:, None] + np.arange(m)[None, :]]
    inside = (sampled != 0).all(axis=1)                        # a pattern may not span a piece boundary
    inside[1::2] = False
    pats[inside] = sampled[inside]
    group = fmx.IndexGroup(fmx.Text.with_max_character(text, 4), fmx.KIND_MULTI, 4, [0, 0], fmx.GROUP_BY_PIECE, mode=fmx.MODE_COMPACT)
    assert group.len() == n and group.pieces_count() == pieces
    r = group.query_batch(pats, piece_ids=True, counts=True)
    hoff, pos, pid = r["hit_off"].astype(np.int64), r["positions"].astype(np.int64), r["piece_ids"].astype(np.int64)
    cnt = np.diff(hoff)
    assert np.array_equal(cnt, r["counts"].astype(np.int64)) and r["total"] == int(hoff[-1]) == pos.size
    assert np.all(cnt[inside] >= 1) and int(cnt[~inside].sum()) <= 2
    owner = np.repeat(np.arange(npat), cnt)
    assert np.array_equal(text[pos[:, None] + np.arange(m)[None, :]], pats[owner])          # every hit holds its pattern
    assert np.array_equal(pid, pos // plen)
    first = hoff[:-1][inside]
    once = cnt[inside] == 1
    assert once.mean() > 0.999 and np.array_equal(pos[first[once]], starts[inside][once])  # .. and is the origin
    assert int(pos.max()) > (1 << 32) and pos[hoff[6]] == (1 << 32) - 7
    del group


def test_partitioned_cuda_engine_and_merge_kernel():
    """fm-index_b200/partitioned.py on one rank with the CUDA engine (fmx_query_batch_device, 32-bit CSR) and
    fmx_csr_merge_device against (a) the replicated index and (b) the host merge the gloo tests pin"""
    import torch
    from fm_index_b200 import partitioned as part
    text, pats = _multi_case(91, pieces=12)
    idx = part.PartitionedMultiPieces(text, 2, 4, device=0)
    counts, hoff, pos, pid = idx.search_locate(torch.from_numpy(pats))
    single = fmx.FMIndexMultiPiecesWithLocate.new(fmx.Text.with_max_character(text, 4), 2)
    ref = single.query_batch(pats, piece_ids=True, counts=True, capacity=int(hoff[-1].item()) + 8)
    assert np.array_equal(hoff.cpu().numpy().view(np.uint64), ref["hit_off"])
    assert np.array_equal(pos.cpu().numpy().view(np.uint64), ref["positions"])          # one partition: same order too
    assert np.array_equal(pid.cpu().numpy().view(np.uint64), ref["piece_ids"])
    assert np.array_equal(counts.cpu().numpy().view(np.uint64), ref["counts"])
    # three partitions answered by three engines on this GPU, merged by the kernel and by the host merge
    starts, ends = part.piece_bounds(text)
    ranges = part.partition_pieces(ends - starts + 1, 3)
    parts, bases = [], []
    for a, b in ranges:
        base = int(starts[a])
        eng = part.CudaEngine(text[base:int(ends[b - 1]) + 1], 2, 4, 0)
        off_r, pos_r, pid_r = eng.search_locate(torch.from_numpy(pats))
        parts.append((off_r.clone(), pos_r.clone(), pid_r.clone()))
        bases.append((base, int(a)))
    dev = torch.device("cuda", 0)
    k_off, k_pos, k_pid = part.merge_parts_cuda(parts, bases, pats.shape[0], dev)
    h_off, h_pos, h_pid = part.merge_parts_host([tuple(t.cpu() for t in p) for p in parts], bases, pats.shape[0], None)
    assert torch.equal(k_off.cpu(), h_off) and torch.equal(k_pos.cpu(), h_pos) and torch.equal(k_pid.cpu(), h_pid)
    _same_match_sets(k_off.cpu().numpy(), k_pos.cpu().numpy(), k_pid.cpu().numpy(), ref["hit_off"], ref["positions"], ref["piece_ids"])
    with pytest.raises(ValueError):
        bad = pats.copy()
        bad[0, 0] = 0
        idx.search_locate(torch.from_numpy(bad))


# ---- texts of wide characters (character.rs:38-42): u16 / u32 / u64, WIDE layout (max_character > 255) and the u8
# layouts behind wide patterns (max_character <= 255); every entry point against the oracle
def _wide_text(rng, n, mc, dtype, multi, nsym=40):
    used = np.unique(np.concatenate([rng.integers(1, mc + 1, nsym), [mc]]))
    body = rng.choice(used, n)
    if multi:
        body[rng.integers(2, n - 2, max(1, n // 400)) // 2 * 2] = 0
    return np.append(body, 0).astype(dtype)


@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
@pytest.mark.parametrize("dtype,mc", [(np.uint16, 4), (np.uint16, 300), (np.uint16, 65535), (np.uint32, 5_000_011), (np.uint64, 70_000)])
def test_wide_character_texts(kind, dtype, mc, tmp_path):
    rng = np.random.default_rng(kind * 31 + mc % 101)
    multi = kind == orc.MULTI
    n = 40_000
    text = _wide_text(rng, n, mc, dtype, multi, nsym=3 if mc == 4 else 40)
    level = 2
    index = KINDS[kind][1].new(fmx.Text.with_max_character(text, mc), level)
    oracle = orc.OracleIndex(text, kind, level=level, max_character=mc)
    assert index.len() == text.size and index.layout_name().startswith("WIDE" if mc > 255 else "Q4")
    pats = []
    for t in range(3000):
        m = int(rng.integers(1, 12))
        p0 = int(rng.integers(0, n - m))
        p = text[p0:p0 + m].copy() if t % 3 else rng.integers(1, mc + 1, m).astype(dtype)
        if t % 7 == 0 and m > 2:
            p[int(rng.integers(0, m))] = rng.integers(1, mc + 1)       # a mismatch, often a symbol that never occurs
        if not multi and 0 in p:
            continue
        pats.append(p)
    flat, off = orc.pack_patterns(pats, dtype)
    modes = [fmx.SEARCH] + ([fmx.SEARCH_PREFIX, fmx.SEARCH_SUFFIX, fmx.SEARCH_EXACT] if multi else [])
    for mode in modes:
        b = index.search_batch((flat, off), mode)
        os_, oe = oracle.search_batch(flat, off, mode)
        assert np.array_equal(b.s, os_) and np.array_equal(b.e, oe)
        po = mode in (fmx.SEARCH_PREFIX, fmx.SEARCH_EXACT)
        if multi:
            hoff, pos, pid = b.locate(piece_ids=True)
            ooff, opos, opid = oracle.locate_batch(os_, oe, prefix_only=po, want_piece_ids=True)
            assert np.array_equal(pid, opid)
        else:
            hoff, pos = b.locate()
            ooff, opos, _ = oracle.locate_batch(os_, oe, prefix_only=po)
        assert np.array_equal(hoff, ooff) and np.array_equal(pos, opos)         # order included
        r = index.query_batch((flat, off), mode, rows=True, counts=True, locate=True, width=4)
        assert np.array_equal(r["s"], os_) and np.array_equal(r["e"], oe)
        assert np.array_equal(r["hit_off"], ooff) and np.array_equal(r["positions"], opos)
        fb, fh, fp = index.search_locate_batch((flat, off), mode)[:3]
        assert np.array_equal(fh, ooff) and np.array_equal(fp, opos)
    # fixed-length patterns, refinement, the Search / Match objects
    fixed = np.stack([text[i:i + 8] for i in rng.integers(0, n - 8, 500)])
    fixed = fixed[~(fixed == 0).any(axis=1)]
    fb = index.search_batch(fixed)
    fs, fe = oracle.search_batch(fixed.reshape(-1), np.arange(fixed.shape[0] + 1, dtype=np.uint64) * 8)
    assert np.array_equal(fb.s, fs) and np.array_equal(fb.e, fe)
    head, tail = fixed[:, :3], fixed[:, 3:]
    rb = index.search_batch(tail).search_batch(head)                             # Search::search refines by prepending
    assert np.array_equal(rb.s, fs) and np.array_equal(rb.e, fe)
    one = index.search(fixed[0])
    assert one.count() == int(fe[0] - fs[0])
    m0 = next(iter(one.iter_matches()))
    assert m0.locate() == oracle.get_sa(int(fs[0]))
    got = [c for _, c in zip(range(8), m0.iter_chars_forward())]
    assert got[:len(fixed[0])] == [int(c) for c in fixed[0]][:len(got)]
    # primitives and extraction, in the text's character width
    rows = rng.integers(0, text.size, 400).astype(np.uint64)
    assert list(index.rows_op(0, rows)) == [oracle.get_l(int(i)) for i in rows]
    assert list(index.rows_op(1, rows)) == [oracle.lf_map(int(i)) for i in rows]
    assert list(index.rows_op(2, rows)) == [oracle.get_f(int(i)) for i in rows]
    assert list(index.rows_op(3, rows)) == [(1 << 64) - 1 if oracle.fl_map(int(i)) is None else oracle.fl_map(int(i)) for i in rows]
    assert list(index.rows_op(4, rows)) == [oracle.get_sa(int(i)) for i in rows]
    cs_, is_ = [], []
    for c in [int(v) for v in np.unique(text)[:20]] + [mc, 1]:
        for i in list(rng.integers(0, text.size + 1, 40)) + [0, text.size]:
            cs_.append(c)
            is_.append(int(i))
    assert list(index.lf_map2_batch(cs_, is_)) == [oracle.lf_map2(c, i) for c, i in zip(cs_, is_)]
    for fwd in (False, True):
        out, ln = index.extract_batch(rows, 21, fwd)
        oout, oln = oracle.extract_batch(rows, 21, fwd)
        assert out.dtype == np.dtype(dtype) and np.array_equal(ln, oln) and np.array_equal(out, oout)
    # a character above max_character is the reference's panic (fm_index.rs:94)
    if mc < np.iinfo(dtype).max:
        with pytest.raises(IndexError):
            index.search_batch([np.array([int(text[0]), mc + 1], dtype=dtype)])
    # save / load keeps the character width
    p = tmp_path / "wide.fmx"
    index.save(p)
    again = KINDS[kind][1].load(p)
    lb = again.search_batch((flat, off))
    os_, oe = oracle.search_batch(flat, off)
    assert again._dtype == np.dtype(dtype) and np.array_equal(lb.s, os_) and np.array_equal(lb.e, oe)
