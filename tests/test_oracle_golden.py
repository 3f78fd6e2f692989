"""Pins the CPU oracle against every known-answer vector the reference's own tests hold
for the count/locate/extract path.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as orc
from refutil import (TEXT_README, TEXT_TWINKLE, build_text, naive_search, naive_suffix_array, random_cases)

MISS = b"mississippi\0"


# ---- src/fm_index.rs:148-173
def test_fm_lf_map_chain():
    idx = orc.OracleIndex(MISS, orc.FM, level=2)
    i, got = 0, []
    for _ in range(12):
        i = idx.lf_map(i)
        got.append(i)
    assert got == [1, 6, 7, 2, 8, 10, 3, 9, 11, 4, 5, 0]


def test_fm_fl_map():
    idx = orc.OracleIndex(MISS, orc.FM, level=2)
    assert [idx.fl_map(i) for i in range(12)] == [5, 0, 7, 10, 11, 4, 1, 6, 2, 3, 8, 9]


# ---- src/rlfmi.rs:197-351
def test_rlfm_s_b_bp_cs():
    idx = orc.OracleIndex(MISS, orc.RLFM)
    L = orc.lib()
    assert bytes(L.orc_rlfm_s(idx._h, i) for i in range(9)) == b"ipsm\0pisi"          # rlfmi.rs:201
    assert L.orc_rlfm_runs(idx._h) == 9
    assert [L.orc_rlfm_b(idx._h, i) for i in range(12)] == [1, 1, 1, 0, 1, 1, 1, 1, 1, 0, 1, 0]   # :213
    assert [L.orc_rlfm_bp(idx._h, i) for i in range(12)] == [1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 0]  # :236
    for c, a in [(0, 0), (ord("i"), 1), (ord("m"), 4), (ord("p"), 5), (ord("s"), 7)]:             # :252
        assert idx.cs(c) == a


def test_rlfm_get_l_lf_fl_getf():
    idx = orc.OracleIndex(MISS, orc.RLFM)
    assert bytes(idx.get_l(i) for i in range(12)) == b"ipssm\0pissii"                  # rlfmi.rs:262
    i, got = 0, []
    for _ in range(12):
        i = idx.lf_map(i)
        got.append(i)
    assert got == [1, 6, 7, 2, 8, 10, 3, 9, 11, 4, 5, 0]                                 # :273
    assert [idx.fl_map(i) for i in range(12)] == [5, 0, 7, 10, 11, 4, 1, 6, 2, 3, 8, 9]  # :346
    assert bytes(idx.get_f(i) for i in range(12)) == bytes(sorted(MISS))                 # :325-336


def test_rlfm_lf_map2_ranges_and_search():
    idx = orc.OracleIndex(MISS, orc.RLFM)
    n = idx.len()
    for c, r in [(0, (0, 1)), (ord("i"), (1, 5)), (ord("m"), (5, 6)), (ord("p"), (6, 8)), (ord("s"), (8, 12))]:
        assert (idx.lf_map2(c, 0), idx.lf_map2(c, n)) == r                               # rlfmi.rs:287-293
    for pat, r in [(b"iss", (3, 5)), (b"ppi", (7, 8)), (b"si", (8, 10)), (b"ssi", (10, 12))]:
        assert idx.search(pat) == r                                                      # :314-319
        assert orc.OracleIndex(MISS, orc.FM).search(pat) == r


# ---- src/suffix_array/sample.rs:96-136 (through a text whose SA is known)
@pytest.mark.parametrize("level,n", [(1, 10), (1, 25), (2, 8), (2, 9), (2, 10), (2, 25), (3, 24), (3, 25)])
def test_sample_regular(level, n):
    # a strictly decreasing text has SA = [n-1, n-2, ..., 0]; sample.rs tests SA = identity,
    # the sampling rule only looks at row indices so both pin `get(i) is Some iff i % 2^level == 0`.
    text = bytes(list(range(n - 1, 0, -1)) + [0])
    idx = orc.OracleIndex(text, orc.FM, level=level)
    for i in range(n):
        v = idx.sample_get(i)
        if i & ((1 << level) - 1) == 0:
            assert v == n - 1 - i
        else:
            assert v is None
    assert idx.sample_get(n) is None


def test_sample_level_too_high():
    text = bytes(list(range(9, 0, -1)) + [0])     # n = 10 <= 2^4  => level forced to 0 (sample.rs:28-31, :126-135)
    idx = orc.OracleIndex(text, orc.FM, level=4)
    assert orc.lib().orc_sample_level(idx._h) == 0
    assert [idx.sample_get(i) for i in range(10)] == [9 - i for i in range(10)]
    assert orc.lib().orc_sample_word_size(idx._h) == 4   # floor(log2 10) + 1


# ---- README.md:49-85 (the crate's doctest, src/lib.rs:148-150)
def test_readme_doctest():
    idx = orc.OracleIndex(TEXT_README, orc.FM, level=2)
    s, e = idx.search(b"dolor")
    assert e - s == 4
    assert idx.locate_all(b"dolor") == [246, 12, 300, 103]      # ORDER is pinned (README.md:64)
    rows = list(range(s, e))
    out, _ = idx.extract_batch([rows[0]], 16, forward=False)
    assert bytes(out[0][::-1]) == b"Duis aute irure "
    out, ln = idx.extract_batch([rows[3]], 20, forward=True)
    assert bytes(out[0]) == b"dolore magna aliqua." and ln[0] == 20


# ---- examples/multi_pieces.rs:34-88
def test_multi_pieces_example():
    idx = orc.OracleIndex(TEXT_TWINKLE, orc.MULTI, level=2)
    assert idx.count(b"star") == 4
    assert sorted(idx.piece_ids_all(b"How I wonder")) == [0, 0, 1, 2]

    def take_while(row, forward, stop):
        out, ln = idx.extract_batch([row], 64, forward=forward)
        b = bytes(out[0][: ln[0]])
        return b[: b.index(stop)] if stop in b else b

    assert [take_while(r, False, b" ") for r in idx.rows(b" in the dark")] == [b"rellevart"]
    assert [take_while(r, True, b",") for r in idx.rows(b"ing ")] == [b"ing shines upon", b"ing sun is gone"]
    assert sorted(idx.piece_ids_all(b"Twinkle", orc.SEARCH_PREFIX)) == [0]
    assert sorted(idx.piece_ids_all(b"what you are!\n", orc.SEARCH_SUFFIX)) == [0, 1, 2]


# ---- tests/test_fmindex.rs:6-24, tests/test_multi_pieces.rs:8-41, tests/test_api.rs:15-18
@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
def test_small(kind):
    idx = orc.OracleIndex(b"a\0", kind, level=2)
    assert idx.count(b"a") == 1
    assert idx.locate_all(b"a") == [0]
    if kind == orc.MULTI:
        assert idx.piece_ids_all(b"a") == [0]
    assert orc.OracleIndex(b"text\0", kind).len() == 5


# ---- src/suffix_array/sais.rs:401-425 (the three InvalidText errors)
@pytest.mark.parametrize("text,msg", [
    (b"\0abc\0", "must not start with zero"),
    (b"abc", "must end with exactly one zero"),
    (b"abc\0\0", "must end with exactly one zero"),
])
def test_invalid_text(text, msg):
    with pytest.raises(orc.InvalidText, match=msg):
        orc.OracleIndex(text, orc.FM)
    with pytest.raises(orc.InvalidText, match=msg):
        orc.suffix_array(text)


# ---- src/suffix_array/sais.rs:452-457, :546-557: SA == naive sort, incl. interior zeros
def test_sa_known_and_random():
    assert list(orc.suffix_array(b"mmiissiissiippii\0")) == naive_suffix_array(b"mmiissiissiippii\0")
    rng = np.random.default_rng(11)
    for t in range(200):
        n = int(rng.integers(2, 300))
        sigma = int(rng.choice([2, 3, 8, 255]))
        from refutil import build_text
        text = build_text(rng, n, sigma + 1 if sigma < 255 else 255, multi_pieces=bool(t & 1))
        assert list(orc.suffix_array(text)) == naive_suffix_array(text), text


# ---- src/multi_pieces.rs:250-325
def test_multi_lf_map_and_piece_id_random():
    rng = np.random.default_rng(0)
    from refutil import build_text
    texts = [b"foo\0bar\0baz\0"] + [build_text(rng, 512, 8, True) for _ in range(30)]
    for text in texts:
        n = len(text)
        sa = naive_suffix_array(text)
        isa = [0] * n
        for p, i in enumerate(sa):
            isa[i] = p
        idx = orc.OracleIndex(text, orc.MULTI, level=0)
        assert [idx.lf_map(i) for i in range(n)] == [isa[(sa[i] - 1) % n] for i in range(n)]
        assert [idx.piece_id(i) for i in range(n)] == [text[: sa[i]].count(0) for i in range(n)]


# ---- tests/test_fmindex.rs, test_rlfmindex.rs, test_multi_pieces.rs: differential vs naive scan
@pytest.mark.parametrize("kind", [orc.FM, orc.RLFM, orc.MULTI])
def test_random_differential(kind):
    multi = kind == orc.MULTI
    for text, level, pats in random_cases(seed=kind, texts=25, patterns=40, text_size_max=300 if not multi else 600,
                                          alphabet_size=8, level_max=3, pattern_size_max=10, multi_pieces=multi):
        idx = orc.OracleIndex(text, kind, level=level)
        for pat in pats:
            exp = naive_search(text, pat)
            assert idx.count(pat) == len(exp)
            assert sorted(idx.locate_all(pat)) == [p for p, _ in exp]
            if multi:
                assert sorted(idx.piece_ids_all(pat)) == [d for _, d in exp]
                for mode, pre, suf in [(orc.SEARCH_PREFIX, True, False), (orc.SEARCH_SUFFIX, False, True),
                                       (orc.SEARCH_EXACT, True, True)]:
                    exp2 = naive_search(text, pat, pre, suf)
                    assert sorted(idx.piece_ids_all(pat, mode)) == [d for _, d in exp2]
                    assert sorted(idx.locate_all(pat, mode)) == [p for p, _ in exp2]


def test_rlfm_matches_fm_everywhere():
    """RLFM lf_map2 == cs[c] + rank(bw, i, c) for every c and i in [0, n] (rlfmi.rs:285-309 edge)."""
    rng = np.random.default_rng(5)
    from refutil import build_text
    for _ in range(20):
        text = build_text(rng, int(rng.integers(2, 200)), 5, False)
        fm = orc.OracleIndex(text, orc.FM, level=1, max_character=7)
        rl = orc.OracleIndex(text, orc.RLFM, level=1, max_character=7)
        n = len(text)
        for c in range(8):
            for i in range(n + 1):
                assert fm.lf_map2(c, i) == rl.lf_map2(c, i)
        for i in range(n):
            assert fm.get_l(i) == rl.get_l(i) and fm.lf_map(i) == rl.lf_map(i)
            assert fm.get_f(i) == rl.get_f(i) and fm.fl_map(i) == rl.fl_map(i)
            assert fm.get_sa(i) == rl.get_sa(i)


def test_empty_pattern_and_refine():
    idx = orc.OracleIndex(MISS, orc.FM, level=0)
    assert idx.search(b"") == (0, 12)                      # wrapper.rs:103-124 with an empty loop
    r1 = idx.search(b"si")
    assert idx.search(b"s", init=r1) == idx.search(b"ssi")  # refinement prepends (wrapper.rs:99-124)


def test_pattern_char_above_max_character():
    idx = orc.OracleIndex(bytes([1, 2, 3, 4, 1, 2, 0]), orc.FM, max_character=4)
    with pytest.raises(IndexError):
        idx.search(bytes([1, 5]))
    # the reference breaks out before reaching the bad char when the range empties first
    assert idx.search(bytes([5, 4, 4]))[0] == idx.search(bytes([5, 4, 4]))[1]


def test_suffix_array_checker_and_build_from_sa():
    """orc_check_suffix_array accepts exactly the suffix array; orc_build_from_sa refuses anything else
    (bench.py hands the oracle a GPU-built SA for GB-scale texts: it is verified, not trusted)"""
    rng = np.random.default_rng(9)
    for multi in (False, True):
        for n in (2, 3, 50, 3000):
            text = build_text(rng, n, 4, multi)
            sa = orc.suffix_array(text)
            assert orc.check_suffix_array(text, sa)
            if n > 3:
                bad = sa.copy()
                i = int(rng.integers(0, n - 1))
                bad[i], bad[i + 1] = bad[i + 1], bad[i]
                assert not orc.check_suffix_array(text, bad)
                dup = sa.copy()
                dup[1] = dup[2]
                assert not orc.check_suffix_array(text, dup)
                oob = sa.copy()
                oob[0] = n
                assert not orc.check_suffix_array(text, oob)
                with pytest.raises(orc.InvalidText, match="not the suffix array"):
                    orc.OracleIndex(text, orc.MULTI if multi else orc.FM, level=1, sa=bad)
            a = orc.OracleIndex(text, orc.MULTI if multi else orc.FM, level=1, sa=sa)
            b = orc.OracleIndex(text, orc.MULTI if multi else orc.FM, level=1)
            assert [a.lf_map(i) for i in range(n)] == [b.lf_map(i) for i in range(n)]
            assert [a.get_sa(i) for i in range(n)] == [b.get_sa(i) for i in range(n)]
    rep = (build_text(rng, 700, 3, False)[:-1] * 30) + b"\0"      # repetitive: long equal prefixes
    assert orc.check_suffix_array(rep, orc.suffix_array(rep))
