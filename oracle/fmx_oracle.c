/*
 * fmx_oracle.c -- CPU oracle for the FM-index query path.  TEST INFRASTRUCTURE ONLY
 * (see fmx_oracle.h for the rules about who may load this).
 *
 * Restates, in plain C:
 *   - src/wrapper.rs            (backward search, match iteration, locate, extraction)
 *   - src/fm_index.rs:44-141    (FMIndexBackend)
 *   - src/rlfmi.rs:30-190       (RLFMIndexBackend)
 *   - src/multi_pieces.rs:31-223(FMIndexMultiPiecesBackend)
 *   - src/suffix_array/sample.rs:21-60 (SOSampledSuffixArray)
 *   - src/suffix_array/sais.rs:9-32, 115-144 (cs, text validation; SA = SA-IS,
 *     Nong/Zhang/Chan 2010, the algorithm sais.rs:147-307 implements)
 *   - vers-vecs =1.10.1 (absent third-party dependency, Cargo.toml:16): RsVec
 *     rank1/rank0/select1/select0/get with 512-bit blocks + 8192-bit super-blocks held
 *     in separate arrays, WaveletMatrix rank/get/select with one RsVec per bit level,
 *     MSB first, zeros stably partitioned before ones (its published design).  Only the
 *     VALUES rank/select/access return are observable through the reference; the
 *     edge rules the reference relies on (rank1 clamps past the end, select1 past the
 *     last one returns len: rlfmi.rs:285-309) are kept.
 *
 * Parity: pinned by the reference's own known-answer tests, see
 * tests/test_oracle_golden.py.  vers-internal memory layout is unpinned (and
 * irrelevant to results).
 */
#include "fmx_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ util */

/* util.rs:1-3 */
static inline uint32_t log2_u64(uint64_t x) { return 63u - (uint32_t)__builtin_clzll(x); }

static void set_err(char *err, size_t errlen, const char *msg) {
    if (err && errlen) {
        snprintf(err, errlen, "%s", msg);
    }
}

/* ------------------------------------------------------------------ characters
 * character.rs:24-42: u8, u16, u32, u64, usize texts.  A text (or pattern) is a pointer plus the width of one
 * character in bytes; every character is read through tget.  Values must stay below 2^32 (the reference sizes
 * its `cs` table by max_character + 1 words, so larger alphabets are out of reach there as well). */
typedef struct {
    const void *p;
    uint32_t w; /* 1, 2, 4 or 8 */
} tview;

static inline uint64_t tget(tview t, uint64_t i) {
    switch (t.w) {
        case 1: return ((const uint8_t *)t.p)[i];
        case 2: return ((const uint16_t *)t.p)[i];
        case 4: return ((const uint32_t *)t.p)[i];
        default: return ((const uint64_t *)t.p)[i];
    }
}
static inline tview tview_at(tview t, uint64_t i) {
    tview r;
    r.p = (const uint8_t *)t.p + i * t.w;
    r.w = t.w;
    return r;
}
static inline void tput(void *p, uint32_t w, uint64_t i, uint64_t v) {
    switch (w) {
        case 1: ((uint8_t *)p)[i] = (uint8_t)v; break;
        case 2: ((uint16_t *)p)[i] = (uint16_t)v; break;
        case 4: ((uint32_t *)p)[i] = (uint32_t)v; break;
        default: ((uint64_t *)p)[i] = v; break;
    }
}

/* ------------------------------------------------------------------ SA-IS
 * Suffix array of the text under plain lexicographic order with \0 an ordinary
 * (smallest) symbol, which is what sais.rs produces (pinned there against a naive
 * sort, sais.rs:546-557).  Core routine: classic SA-IS on a string whose last
 * symbol is a unique smallest sentinel.  The top level appends a virtual sentinel
 * (symbols shifted by +1) so arbitrary texts, incl. interior zeros, are handled. */

typedef struct {
    const uint8_t *t8;  /* mode 0: original u8 text, virtual sentinel at n-1 */
    const int64_t *t64; /* mode 1: reduced string */
    tview tw;           /* mode 2: original text of wider characters, virtual sentinel at n-1 */
    int64_t n;
    int mode;
} sstr;

static inline int64_t sch(const sstr *s, int64_t i) {
    if (s->mode == 0) return i == s->n - 1 ? 0 : (int64_t)s->t8[i] + 1;
    if (s->mode == 2) return i == s->n - 1 ? 0 : (int64_t)tget(s->tw, (uint64_t)i) + 1;
    return s->t64[i];
}

#define TGET(t, i) (((t)[(i) >> 3] >> ((i)&7)) & 1)
#define TSET(t, i) ((t)[(i) >> 3] |= (uint8_t)(1u << ((i)&7)))
#define IS_LMS(t, i) ((i) > 0 && TGET(t, i) && !TGET(t, (i)-1))

static void sais_buckets(const sstr *s, int64_t *bkt, int64_t K, int end) {
    memset(bkt, 0, (size_t)K * sizeof(int64_t));
    for (int64_t i = 0; i < s->n; i++) bkt[sch(s, i)]++;
    int64_t sum = 0;
    for (int64_t c = 0; c < K; c++) {
        sum += bkt[c];
        bkt[c] = end ? sum : sum - bkt[c];
    }
}

static void sais_induce(const sstr *s, const uint8_t *t, int64_t *SA, int64_t *bkt, int64_t K) {
    int64_t n = s->n;
    sais_buckets(s, bkt, K, 0);
    for (int64_t i = 0; i < n; i++) {
        int64_t j = SA[i];
        if (j > 0 && !TGET(t, j - 1)) SA[bkt[sch(s, j - 1)]++] = j - 1;
    }
    sais_buckets(s, bkt, K, 1);
    for (int64_t i = n - 1; i >= 0; i--) {
        int64_t j = SA[i];
        if (j > 0 && TGET(t, j - 1)) SA[--bkt[sch(s, j - 1)]] = j - 1;
    }
}

static void sais_core(const sstr *s, int64_t *SA, int64_t K) {
    int64_t n = s->n;
    if (n == 1) {
        SA[0] = 0;
        return;
    }
    uint8_t *t = (uint8_t *)calloc((size_t)(n + 7) / 8, 1);
    int64_t *bkt = (int64_t *)malloc((size_t)K * sizeof(int64_t));
    /* types: 1 = S, 0 = L */
    TSET(t, n - 1);
    for (int64_t i = n - 2; i >= 0; i--) {
        int64_t c0 = sch(s, i), c1 = sch(s, i + 1);
        if (c0 < c1 || (c0 == c1 && TGET(t, i + 1))) TSET(t, i);
    }
    /* stage 1: sort LMS substrings */
    sais_buckets(s, bkt, K, 1);
    for (int64_t i = 0; i < n; i++) SA[i] = -1;
    for (int64_t i = 1; i < n; i++)
        if (IS_LMS(t, i)) SA[--bkt[sch(s, i)]] = i;
    sais_induce(s, t, SA, bkt, K);
    int64_t n1 = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t p = SA[i];
        if (p > 0 && IS_LMS(t, p)) SA[n1++] = p;
    }
    for (int64_t i = n1; i < n; i++) SA[i] = -1;
    int64_t name = 0, prev = -1;
    for (int64_t i = 0; i < n1; i++) {
        int64_t pos = SA[i];
        int diff = 0;
        if (prev < 0) {
            diff = 1;
        } else {
            for (int64_t d = 0;; d++) {
                if (sch(s, pos + d) != sch(s, prev + d) || TGET(t, pos + d) != TGET(t, prev + d)) {
                    diff = 1;
                    break;
                }
                if (d > 0 && (IS_LMS(t, pos + d) || IS_LMS(t, prev + d))) break;
            }
        }
        if (diff) {
            name++;
            prev = pos;
        }
        SA[n1 + pos / 2] = name - 1;
    }
    {
        int64_t j = n - 1;
        for (int64_t i = n - 1; i >= n1; i--)
            if (SA[i] >= 0) SA[j--] = SA[i];
    }
    int64_t *s1 = SA + n - n1;
    /* stage 2: order of LMS suffixes */
    if (name < n1) {
        sstr r;
        r.t8 = NULL;
        r.t64 = s1;
        r.n = n1;
        r.mode = 1;
        sais_core(&r, SA, name);
    } else {
        for (int64_t i = 0; i < n1; i++) SA[s1[i]] = i;
    }
    /* stage 3: induce the full order */
    {
        int64_t j = 0;
        for (int64_t i = 1; i < n; i++)
            if (IS_LMS(t, i)) s1[j++] = i;
    }
    for (int64_t i = 0; i < n1; i++) SA[i] = s1[SA[i]];
    for (int64_t i = n1; i < n; i++) SA[i] = -1;
    sais_buckets(s, bkt, K, 1);
    for (int64_t i = n1 - 1; i >= 0; i--) {
        int64_t j = SA[i];
        SA[i] = -1;
        SA[--bkt[sch(s, j)]] = j;
    }
    sais_induce(s, t, SA, bkt, K);
    free(bkt);
    free(t);
}

/* sais.rs:115-144 */
static int suffix_array_view(tview text, uint64_t n, uint64_t mc, uint64_t *sa, char *err, size_t errlen) {
    if (n == 0) return 0;
    if (n == 1) {
        sa[0] = 0;
        return 0;
    }
    if (tget(text, 0) == 0) { /* sais.rs:128-133 */
        set_err(err, errlen, "the given text must not start with zero character");
        return -1;
    }
    { /* sais.rs:134-139: rposition of the last non-zero char must be n-2 */
        int64_t last_nz = -1;
        for (int64_t i = (int64_t)n - 1; i >= 0; i--)
            if (tget(text, (uint64_t)i) != 0) {
                last_nz = i;
                break;
            }
        if (last_nz != (int64_t)n - 2) {
            set_err(err, errlen, "the given text must end with exactly one zero character");
            return -1;
        }
    }
    int64_t *tmp = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
    sstr s;
    s.t8 = (const uint8_t *)text.p;
    s.t64 = NULL;
    s.tw = text;
    s.n = (int64_t)n + 1;
    s.mode = text.w == 1 ? 0 : 2;
    sais_core(&s, tmp, text.w == 1 ? 257 : (int64_t)mc + 2);
    for (uint64_t i = 0; i < n; i++) sa[i] = (uint64_t)tmp[i + 1];
    free(tmp);
    return 0;
}

int orc_suffix_array(const uint8_t *text, uint64_t n, uint64_t *sa, char *err, size_t errlen) {
    tview t = {text, 1};
    return suffix_array_view(t, n, 255, sa, err, errlen);
}

/* texts of u16 / u32 / u64 characters (character.rs:38-42); every character <= max_character < 2^32 */
int orc_suffix_array_w(const void *text, uint32_t char_width, uint64_t n, uint64_t max_character, uint64_t *sa, char *err,
                       size_t errlen) {
    tview t = {text, char_width};
    for (uint64_t i = 0; i < n; i++)
        if (tget(t, i) > max_character) {
            set_err(err, errlen, "text contains a character larger than max_character");
            return -1;
        }
    return suffix_array_view(t, n, max_character, sa, err, errlen);
}

/* ------------------------------------------------------------------ RsVec
 * vers-vecs RsVec: bit vector + rank directory (512-bit blocks with u16 counts
 * relative to 8192-bit super-blocks with u64 counts), in separate arrays. */

typedef struct {
    uint64_t len, ones, nwords, nblk, nsuper;
    uint64_t *w;
    uint64_t *super; /* ones before super-block k, k in 0..nsuper */
    uint16_t *blk;   /* ones from the super-block start to block b */
} rsvec;

static void rs_init(rsvec *v, uint64_t len) {
    memset(v, 0, sizeof(*v));
    v->len = len;
    v->nblk = (len >> 9) + 1;
    v->nsuper = (len >> 13) + 1;
    v->nwords = v->nblk * 8 + 1;
    v->w = (uint64_t *)calloc(v->nwords, 8);
    v->super = (uint64_t *)calloc(v->nsuper + 1, 8);
    v->blk = (uint16_t *)calloc(v->nblk + 1, 2);
}
static inline void rs_setbit(rsvec *v, uint64_t i) { v->w[i >> 6] |= 1ull << (i & 63); }
static void rs_finish(rsvec *v) {
    uint64_t total = 0, in_super = 0;
    for (uint64_t b = 0; b < v->nblk; b++) {
        if ((b & 15) == 0) {
            v->super[b >> 4] = total;
            in_super = 0;
        }
        v->blk[b] = (uint16_t)in_super;
        uint64_t c = 0;
        for (int k = 0; k < 8; k++) c += (uint64_t)__builtin_popcountll(v->w[b * 8 + k]);
        total += c;
        in_super += c;
    }
    v->super[v->nsuper] = total;
    v->ones = total;
}
static void rs_free(rsvec *v) {
    free(v->w);
    free(v->super);
    free(v->blk);
    memset(v, 0, sizeof(*v));
}
static inline uint64_t rs_get(const rsvec *v, uint64_t i) { return (v->w[i >> 6] >> (i & 63)) & 1; }
/* rank1(pos) = ones in [0,pos); positions past the end clamp (vers edge rule). */
static inline uint64_t rs_rank1(const rsvec *v, uint64_t pos) {
    if (pos >= v->len) return v->ones;
    uint64_t b = pos >> 9;
    uint64_t r = v->super[pos >> 13] + v->blk[b];
    uint64_t wi = pos >> 6;
    for (uint64_t k = b * 8; k < wi; k++) r += (uint64_t)__builtin_popcountll(v->w[k]);
    r += (uint64_t)__builtin_popcountll(v->w[wi] & ((1ull << (pos & 63)) - 1));
    return r;
}
static inline uint64_t rs_rank0(const rsvec *v, uint64_t pos) {
    if (pos >= v->len) return v->len - v->ones;
    return pos - rs_rank1(v, pos);
}
static inline uint32_t select_in_word(uint64_t w, uint32_t k) {
    for (uint32_t i = 0; i < k; i++) w &= w - 1;
    return (uint32_t)__builtin_ctzll(w);
}
/* select1(k) = position of the k-th (0-based) one; len if there is none (vers edge rule). */
static uint64_t rs_select1(const rsvec *v, uint64_t k) {
    if (k >= v->ones) return v->len;
    uint64_t lo = 0, hi = v->nsuper; /* largest sb with super[sb] <= k */
    while (hi - lo > 1) {
        uint64_t m = lo + (hi - lo) / 2;
        if (v->super[m] <= k) lo = m; else hi = m;
    }
    uint64_t b = lo << 4, bend = b + 16 < v->nblk ? b + 16 : v->nblk;
    uint64_t base = v->super[lo];
    while (b + 1 < bend && base + v->blk[b + 1] <= k) b++;
    uint64_t r = base + v->blk[b];
    uint64_t wi = b * 8;
    for (;;) {
        uint64_t c = (uint64_t)__builtin_popcountll(v->w[wi]);
        if (r + c > k) break;
        r += c;
        wi++;
    }
    return wi * 64 + select_in_word(v->w[wi], (uint32_t)(k - r));
}
static uint64_t rs_select0(const rsvec *v, uint64_t k) {
    if (k >= v->len - v->ones) return v->len;
    uint64_t lo = 0, hi = v->nsuper;
    while (hi - lo > 1) {
        uint64_t m = lo + (hi - lo) / 2;
        if ((m << 13) - v->super[m] <= k) lo = m; else hi = m;
    }
    uint64_t b = lo << 4, bend = b + 16 < v->nblk ? b + 16 : v->nblk;
    uint64_t base = (lo << 13) - v->super[lo];
    while (b + 1 < bend && base + (((b + 1) & 15) << 9) - v->blk[b + 1] <= k) b++;
    uint64_t r = base + ((b & 15) << 9) - v->blk[b];
    uint64_t wi = b * 8;
    for (;;) {
        uint64_t c = (uint64_t)__builtin_popcountll(~v->w[wi]);
        if (r + c > k) break;
        r += c;
        wi++;
    }
    return wi * 64 + select_in_word(~v->w[wi], (uint32_t)(k - r));
}

/* ------------------------------------------------------------------ WaveletMatrix
 * vers-vecs WaveletMatrix: one RsVec per bit level, level 0 = most significant bit. */

typedef struct {
    uint32_t L;
    uint64_t n;
    rsvec *lv;
    uint64_t *zeros;
} wmat;

/* seq: n symbols of `w` bytes each (1: the u8 texts of every BASELINE config; 4: wider alphabets) */
#define WM_BUILD_BODY(T)                                                                  \
    T *cur = (T *)malloc((n ? n : 1) * sizeof(T)), *nxt = (T *)malloc((n ? n : 1) * sizeof(T)); \
    memcpy(cur, seq, n * sizeof(T));                                                      \
    for (uint32_t l = 0; l < L; l++) {                                                    \
        uint32_t sh = L - 1 - l;                                                          \
        rs_init(&m->lv[l], n);                                                            \
        uint64_t z = 0;                                                                   \
        for (uint64_t i = 0; i < n; i++) {                                                \
            if ((cur[i] >> sh) & 1) rs_setbit(&m->lv[l], i); else z++;                    \
        }                                                                                 \
        rs_finish(&m->lv[l]);                                                             \
        m->zeros[l] = z;                                                                  \
        uint64_t p0 = 0, p1 = z;                                                          \
        for (uint64_t i = 0; i < n; i++) {                                                \
            if ((cur[i] >> sh) & 1) nxt[p1++] = cur[i]; else nxt[p0++] = cur[i];          \
        }                                                                                 \
        T *tmp = cur;                                                                     \
        cur = nxt;                                                                        \
        nxt = tmp;                                                                        \
    }                                                                                     \
    free(cur);                                                                            \
    free(nxt);

static void wm_build(wmat *m, const void *seq, uint32_t w, uint64_t n, uint32_t L) {
    m->L = L;
    m->n = n;
    m->lv = (rsvec *)calloc(L, sizeof(rsvec));
    m->zeros = (uint64_t *)calloc(L, 8);
    if (w == 1) {
        WM_BUILD_BODY(uint8_t)
    } else {
        WM_BUILD_BODY(uint32_t)
    }
}
static void wm_free(wmat *m) {
    for (uint32_t l = 0; l < m->L; l++) rs_free(&m->lv[l]);
    free(m->lv);
    free(m->zeros);
    memset(m, 0, sizeof(*m));
}
/* rank_u64_unchecked(i, sym) = #{ j < i : seq[j] == sym } (range [0,i) mapped down the levels) */
static inline uint64_t wm_rank(const wmat *m, uint64_t i, uint64_t sym) {
    uint64_t s = 0, e = i;
    for (uint32_t l = 0; l < m->L; l++) {
        const rsvec *v = &m->lv[l];
        if ((sym >> (m->L - 1 - l)) & 1) {
            s = m->zeros[l] + rs_rank1(v, s);
            e = m->zeros[l] + rs_rank1(v, e);
        } else {
            s = rs_rank0(v, s);
            e = rs_rank0(v, e);
        }
    }
    return e - s;
}
/* get_u64_unchecked(i) */
static inline uint64_t wm_get(const wmat *m, uint64_t i) {
    uint64_t sym = 0;
    for (uint32_t l = 0; l < m->L; l++) {
        const rsvec *v = &m->lv[l];
        uint64_t bit = rs_get(v, i);
        sym = (sym << 1) | bit;
        i = bit ? m->zeros[l] + rs_rank1(v, i) : rs_rank0(v, i);
    }
    return sym;
}
/* select_u64_unchecked(k, sym) = position of the k-th (0-based) sym */
static inline uint64_t wm_select(const wmat *m, uint64_t k, uint64_t sym) {
    uint64_t s = 0;
    for (uint32_t l = 0; l < m->L; l++) {
        const rsvec *v = &m->lv[l];
        if ((sym >> (m->L - 1 - l)) & 1) s = m->zeros[l] + rs_rank1(v, s); else s = rs_rank0(v, s);
    }
    uint64_t pos = s + k;
    for (uint32_t l = m->L; l-- > 0;) {
        const rsvec *v = &m->lv[l];
        if ((sym >> (m->L - 1 - l)) & 1) pos = rs_select1(v, pos - m->zeros[l]); else pos = rs_select0(v, pos);
    }
    return pos;
}

/* ------------------------------------------------------------------ sampled SA (sample.rs) */

typedef struct {
    uint32_t level, word_size;
    uint64_t len;
    uint64_t *bits; /* word_size-bit fields, LSB first (vers BitVec::append_bits) */
    int present;    /* 0 = DiscardedSuffixArray (discard.rs) */
} ssa;

/* sample.rs:21-44 */
static void ssa_sample(ssa *a, const uint64_t *sa, uint64_t n, uint32_t level) {
    memset(a, 0, sizeof(*a));
    a->present = 1;
    if (n == 0) return;
    a->word_size = log2_u64(n) + 1;
    if (level >= 63 || n <= (1ull << level)) level = 0; /* sample.rs:28-31 */
    a->level = level;
    a->len = n;
    uint64_t cnt = ((n - 1) >> level) + 1;
    a->bits = (uint64_t *)calloc((cnt * a->word_size + 63) / 64 + 2, 8);
    for (uint64_t i = 0; i < cnt; i++) {
        uint64_t v = sa[i << level], bit = i * a->word_size;
        a->bits[bit >> 6] |= v << (bit & 63);
        if ((bit & 63) + a->word_size > 64) a->bits[(bit >> 6) + 1] |= v >> (64 - (bit & 63));
    }
}
/* sample.rs:46-60 */
static inline uint64_t ssa_get(const ssa *a, uint64_t i) {
    if (i >= a->len) return ORC_NONE;
    if ((i & ((1ull << a->level) - 1)) != 0) return ORC_NONE;
    uint64_t bit = (i >> a->level) * a->word_size;
    uint64_t v = a->bits[bit >> 6] >> (bit & 63);
    if ((bit & 63) + a->word_size > 64) v |= a->bits[(bit >> 6) + 1] << (64 - (bit & 63));
    return a->word_size == 64 ? v : v & ((1ull << a->word_size) - 1);
}

/* ------------------------------------------------------------------ index */

struct orc_index {
    int kind;
    uint64_t n;
    uint64_t max_character;
    uint32_t L;
    uint64_t cs_len;
    uint64_t *cs;
    wmat bw; /* FM / MULTI: the BWT.  RLFM: run heads `s` */
    rsvec b, bp;
    uint64_t runs;
    ssa sa;
    uint64_t *doc;
    uint64_t ndoc, first;
};

void orc_free(orc_index *x) {
    if (!x) return;
    free(x->cs);
    wm_free(&x->bw);
    if (x->kind == ORC_RLFM) {
        rs_free(&x->b);
        rs_free(&x->bp);
    }
    free(x->sa.bits);
    free(x->doc);
    free(x);
}

/* Linear-time check that `sa` is THE suffix array of `text` (plain lexicographic suffix order,
 * a proper prefix sorts first -- the order sais.rs:115-307 produces, pinned by sais.rs:546-557):
 * (1) sa is a permutation of 0..n; (2) for neighbours a = sa[i], b = sa[i+1]: text[a] < text[b], or
 * text[a] == text[b] and suffix a+1 sorts before suffix b+1 (by its rank; the empty suffix first).
 * (1) + (2) for all i imply the order is the suffix order by induction on the suffix length.
 * returns 0 when it is.  Lets bench.py hand a GPU-built SA to the oracle without trusting it. */
static int check_suffix_array_view(tview text, uint64_t n, const uint64_t *sa) {
    if (n == 0) return 0;
    uint64_t *isa = (uint64_t *)malloc(n * 8);
    if (!isa) return -2;
    memset(isa, 0xFF, n * 8);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        if (sa[i] >= n) bad |= 1;
        else isa[sa[i]] = (uint64_t)i; /* a duplicate leaves some slot at UINT64_MAX */
    }
    if (!bad) {
#pragma omp parallel for schedule(static) reduction(| : bad)
        for (int64_t i = 0; i < (int64_t)n; i++)
            if (isa[i] == UINT64_MAX || sa[isa[i]] != (uint64_t)i) bad |= 1;
    }
    if (!bad) {
#pragma omp parallel for schedule(static) reduction(| : bad)
        for (int64_t i = 0; i < (int64_t)n - 1; i++) {
            uint64_t a = sa[i], b = sa[i + 1];
            uint64_t ca = tget(text, a), cb = tget(text, b);
            if (ca < cb) continue;
            if (ca > cb) { bad |= 1; continue; }
            if (a + 1 == n) continue;            /* suffix a is a proper prefix of suffix b */
            if (b + 1 == n) { bad |= 1; continue; }
            if (isa[a + 1] > isa[b + 1]) bad |= 1;
        }
    }
    free(isa);
    return bad ? -1 : 0;
}
int orc_check_suffix_array(const uint8_t *text, uint64_t n, const uint64_t *sa) {
    tview t = {text, 1};
    return check_suffix_array_view(t, n, sa);
}

static orc_index *build_impl(tview text, uint64_t n, uint64_t mc, int kind, int level,
                             const uint64_t *sa_in, char *err, size_t errlen) {
    if (text.w == 1 && (mc == 0 || mc > 255)) {
        set_err(err, errlen, "max_character must be in 1..=255 for u8 texts");
        return NULL;
    }
    if (text.w != 1 && (mc == 0 || mc >= 0xFFFFFFFFull || (text.w == 2 && mc > 0xFFFF))) {
        set_err(err, errlen, "max_character out of range for this character width (wide texts: below 2^32 - 1)");
        return NULL;
    }
    /* working width of the BWT / run heads: bytes for u8 texts, u32 for everything wider */
    const uint32_t sw = text.w == 1 ? 1u : 4u;
    for (uint64_t i = 0; i < n; i++) {
        if (tget(text, i) > mc) { /* sais.rs:16-18 would index out of bounds (panic) */
            set_err(err, errlen, "text contains a character larger than max_character");
            return NULL;
        }
    }
    uint64_t *sa = (uint64_t *)malloc((n ? n : 1) * 8);
    if (sa_in) {
        /* still run the validation rules (sais.rs:128-139) */
        if (n >= 2) {
            if (tget(text, 0) == 0) {
                set_err(err, errlen, "the given text must not start with zero character");
                free(sa);
                return NULL;
            }
            if (!(tget(text, n - 1) == 0 && tget(text, n - 2) != 0)) {
                set_err(err, errlen, "the given text must end with exactly one zero character");
                free(sa);
                return NULL;
            }
        }
        if (check_suffix_array_view(text, n, sa_in) != 0) {
            set_err(err, errlen, "the supplied array is not the suffix array of the text");
            free(sa);
            return NULL;
        }
        memcpy(sa, sa_in, n * 8);
    } else if (suffix_array_view(text, n, mc, sa, err, errlen) != 0) {
        free(sa);
        return NULL;
    }
    orc_index *x = (orc_index *)calloc(1, sizeof(orc_index));
    x->kind = kind;
    x->n = n;
    x->max_character = mc;
    x->L = log2_u64(mc) + 1; /* text.rs:61-63 */
    x->cs_len = mc + 1;
    x->cs = (uint64_t *)calloc(x->cs_len, 8);

    if (kind == ORC_FM || kind == ORC_MULTI) {
        /* sais.rs:9-32: cs[c] = #chars < c */
        uint64_t *occ = (uint64_t *)calloc(x->cs_len, 8);
        for (uint64_t i = 0; i < n; i++) occ[tget(text, i)]++;
        uint64_t sum = 0;
        for (uint64_t c = 0; c < x->cs_len; c++) {
            x->cs[c] = sum;
            sum += occ[c];
        }
        free(occ);
        /* fm_index.rs:44-58 / multi_pieces.rs:81-97: bw[i] = text[sa[i]-1], 0 if sa[i]==0 */
        void *bw = calloc(n ? n : 1, sw);
        for (uint64_t i = 0; i < n; i++)
            if (sa[i] > 0) tput(bw, sw, i, tget(text, sa[i] - 1));
        wm_build(&x->bw, bw, sw, n, x->L);
        free(bw);
        if (kind == ORC_MULTI) {
            /* multi_pieces.rs:53-79 */
            uint64_t zc = 0;
            for (uint64_t i = 0; i < n; i++) zc += tget(text, i) == 0;
            uint64_t *zeros_before = (uint64_t *)malloc((n + 1) * 8); /* rank1 of end-marker flags */
            uint64_t acc = 0;
            for (uint64_t i = 0; i < n; i++) {
                zeros_before[i] = acc;
                acc += tget(text, i) == 0;
            }
            zeros_before[n] = acc;
            x->ndoc = zc;
            x->doc = (uint64_t *)calloc(zc ? zc : 1, 8);
            uint64_t nz_l = wm_rank(&x->bw, n, 0);
            for (uint64_t k = 0; k < nz_l && k < zc; k++) {
                uint64_t p = wm_select(&x->bw, k, 0);
                uint64_t em = sa[p] >= 1 ? sa[p] - 1 : n + sa[p] - 1; /* modular_sub */
                uint64_t pid = zeros_before[em];
                if (pid == zc - 1) x->first = p;
                x->doc[k] = pid;
            }
            free(zeros_before);
        }
    } else {
        /* rlfmi.rs:37-96 */
        uint64_t m = x->cs_len;
        void *heads = malloc((n ? n : 1) * sw);
        const tview hv = {heads, sw};
        uint64_t r = 0;
        rs_init(&x->b, n);
        rs_init(&x->bp, n);
        /* per-character run lengths, in L order, kept as (char, len) then bucketed */
        uint64_t *runlen = (uint64_t *)malloc((n ? n : 1) * 8);
        uint64_t c0 = 0;
        for (uint64_t i = 0; i < n; i++) {
            uint64_t k = sa[i];
            uint64_t c = k > 0 ? tget(text, k - 1) : tget(text, n - 1);
            if (c0 != c) {
                tput(heads, sw, r, c);
                runlen[r] = 1;
                r++;
                rs_setbit(&x->b, i);
            } else {
                if (r == 0) { /* rlfmi.rs:62: unreachable!() in the reference */
                    set_err(err, errlen, "text not representable by RLFMIndex (leading zero run)");
                    free(heads);
                    free(runlen);
                    free(sa);
                    orc_free(x);
                    return NULL;
                }
                runlen[r - 1]++;
            }
            c0 = c;
        }
        rs_finish(&x->b);
        x->runs = r;
        wm_build(&x->bw, heads, sw, r, x->L);
        /* bp: runs grouped by head character (stable), each encoded 1 0^{len-1}; cs = #runs with head < c */
        uint64_t *cnt = (uint64_t *)calloc(m + 1, 8);
        for (uint64_t j = 0; j < r; j++) cnt[tget(hv, j)]++;
        uint64_t acc = 0;
        for (uint64_t c = 0; c < m; c++) {
            x->cs[c] = acc;
            acc += cnt[c];
        }
        /* start bit position of each char group in bp */
        uint64_t *len_by_c = (uint64_t *)calloc(m + 1, 8);
        for (uint64_t j = 0; j < r; j++) len_by_c[tget(hv, j)] += runlen[j];
        uint64_t *pos_c = (uint64_t *)calloc(m + 1, 8);
        acc = 0;
        for (uint64_t c = 0; c < m; c++) {
            pos_c[c] = acc;
            acc += len_by_c[c];
        }
        for (uint64_t j = 0; j < r; j++) {
            uint64_t c = tget(hv, j);
            rs_setbit(&x->bp, pos_c[c]);
            pos_c[c] += runlen[j];
        }
        rs_finish(&x->bp);
        free(cnt);
        free(len_by_c);
        free(pos_c);
        free(heads);
        free(runlen);
    }
    if (level >= 0) ssa_sample(&x->sa, sa, n, (uint32_t)level);
    free(sa);
    return x;
}

orc_index *orc_build(const uint8_t *text, uint64_t n, uint64_t mc, int kind, int level, char *err,
                     size_t errlen) {
    tview t = {text, 1};
    return build_impl(t, n, mc, kind, level, NULL, err, errlen);
}
orc_index *orc_build_from_sa(const uint8_t *text, uint64_t n, uint64_t mc, int kind, int level,
                             const uint64_t *sa, char *err, size_t errlen) {
    tview t = {text, 1};
    return build_impl(t, n, mc, kind, level, sa, err, errlen);
}
/* Text<C> for C = u16 / u32 / u64 / usize (character.rs:38-42, text.rs:28-49) */
orc_index *orc_build_w(const void *text, uint32_t char_width, uint64_t n, uint64_t mc, int kind, int level, char *err,
                       size_t errlen) {
    tview t = {text, char_width};
    if (char_width != 1 && char_width != 2 && char_width != 4 && char_width != 8) {
        set_err(err, errlen, "character width must be 1, 2, 4 or 8 bytes");
        return NULL;
    }
    return build_impl(t, n, mc, kind, level, NULL, err, errlen);
}

uint64_t orc_len(const orc_index *x) { return x->n; }
uint64_t orc_pieces_count(const orc_index *x) { return x->ndoc; }
uint64_t orc_cs(const orc_index *x, uint64_t c) { return x->cs[c]; }
uint64_t orc_cs_len(const orc_index *x) { return x->cs_len; }
uint64_t orc_rlfm_runs(const orc_index *x) { return x->runs; }
uint64_t orc_rlfm_s(const orc_index *x, uint64_t i) { return wm_get(&x->bw, i); }
int orc_rlfm_b(const orc_index *x, uint64_t i) { return (int)rs_get(&x->b, i); }
int orc_rlfm_bp(const orc_index *x, uint64_t i) { return (int)rs_get(&x->bp, i); }
uint64_t orc_doc(const orc_index *x, uint64_t k) { return x->doc[k]; }
uint64_t orc_first_row(const orc_index *x) { return x->first; }
uint32_t orc_sample_level(const orc_index *x) { return x->sa.level; }
uint32_t orc_sample_word_size(const orc_index *x) { return x->sa.word_size; }
uint64_t orc_sample_get(const orc_index *x, uint64_t i) {
    return x->sa.present ? ssa_get(&x->sa, i) : ORC_NONE;
}
uint64_t orc_heap_bits(const orc_index *x) {
    uint64_t bits = x->cs_len * 64 + x->ndoc * 64;
    for (uint32_t l = 0; l < x->bw.L; l++) bits += x->bw.lv[l].nwords * 64;
    if (x->kind == ORC_RLFM) bits += (x->b.nwords + x->bp.nwords) * 64;
    if (x->sa.present && x->sa.len) bits += ((((x->sa.len - 1) >> x->sa.level) + 1) * x->sa.word_size);
    return bits;
}

/* ---- backend primitives ---- */

/* fm_index.rs:82-84, multi_pieces.rs:121-123, rlfmi.rs:122-125 */
uint64_t orc_get_l(const orc_index *x, uint64_t i) {
    if (x->kind == ORC_RLFM) return wm_get(&x->bw, rs_rank1(&x->b, i + 1) - 1);
    return wm_get(&x->bw, i);
}

/* fm_index.rs:93-95, multi_pieces.rs:140-153, rlfmi.rs:135-143 */
uint64_t orc_lf_map2(const orc_index *x, uint64_t c, uint64_t i) {
    if (x->kind == ORC_FM) return x->cs[c] + wm_rank(&x->bw, i, c);
    if (x->kind == ORC_MULTI) {
        uint64_t rank = wm_rank(&x->bw, i, c);
        if (c == 0) {
            if (i < x->first) return rank + 1;
            if (i == x->first) return 0;
            return rank;
        }
        return rank + x->cs[c];
    }
    uint64_t j = rs_rank1(&x->b, i);
    uint64_t nr = wm_rank(&x->bw, j, c);
    if (orc_get_l(x, i) != c) return rs_select1(&x->bp, x->cs[c] + nr);
    return rs_select1(&x->bp, x->cs[c] + nr) + i - rs_select1(&x->b, j);
}

/* fm_index.rs:86-91, multi_pieces.rs:125-138, rlfmi.rs:127-133 */
uint64_t orc_lf_map(const orc_index *x, uint64_t i) {
    uint64_t c = orc_get_l(x, i);
    if (x->kind == ORC_RLFM) {
        uint64_t j = rs_rank1(&x->b, i);
        uint64_t nr = wm_rank(&x->bw, j, c);
        return rs_select1(&x->bp, x->cs[c] + nr) + i - rs_select1(&x->b, j);
    }
    return orc_lf_map2(x, c, i);
}

/* fm_index.rs:97-112, multi_pieces.rs:155-169, rlfmi.rs:145-158 */
uint64_t orc_get_f(const orc_index *x, uint64_t i) {
    uint64_t key = i;
    if (x->kind == ORC_RLFM) key = rs_rank1(&x->bp, i + 1) - 1;
    uint64_t s = 0, e = x->cs_len;
    while (e - s > 1) {
        uint64_t m = s + (e - s) / 2;
        if (x->cs[m] <= key) s = m; else e = m;
    }
    return s;
}

/* fm_index.rs:114-120, multi_pieces.rs:171-181, rlfmi.rs:160-169 */
uint64_t orc_fl_map(const orc_index *x, uint64_t i) {
    uint64_t c = orc_get_f(x, i);
    if (x->kind == ORC_RLFM) {
        uint64_t j = rs_rank1(&x->bp, i + 1) - 1;
        uint64_t p = rs_select1(&x->bp, j);
        uint64_t m = wm_select(&x->bw, j - x->cs[c], c);
        uint64_t nn = rs_select1(&x->b, m);
        return nn + i - p;
    }
    if (x->kind == ORC_MULTI && c == 0) return ORC_NONE;
    return wm_select(&x->bw, i - x->cs[c], c);
}

/* fm_index.rs:127-140, rlfmi.rs:176-189, multi_pieces.rs:188-201 */
uint64_t orc_get_sa(const orc_index *x, uint64_t i, uint64_t *steps_out) {
    uint64_t steps = 0;
    for (;;) {
        uint64_t v = ssa_get(&x->sa, i);
        if (v != ORC_NONE) {
            if (steps_out) *steps_out = steps;
            return (v + steps) % x->n;
        }
        i = orc_lf_map(x, i);
        steps++;
    }
}

/* multi_pieces.rs:208-218 */
uint64_t orc_piece_id(const orc_index *x, uint64_t i) {
    for (;;) {
        if (orc_get_l(x, i) == 0) {
            uint64_t prev = x->doc[wm_rank(&x->bw, i, 0)];
            return (prev + 1) % x->ndoc; /* modular_add */
        }
        i = orc_lf_map(x, i);
    }
}

/* ---- wrapper.rs ---- */

/* wrapper.rs:37-42, 61-82 (initial range) + 103-124 (loop) */
static int64_t search_view(const orc_index *x, int mode, tview pat, uint64_t m, int use_init,
                           uint64_t init_s, uint64_t init_e, uint64_t *so, uint64_t *eo) {
    uint64_t s, e;
    if (use_init) {
        s = init_s;
        e = init_e;
    } else {
        s = 0;
        e = (mode == ORC_SEARCH_SUFFIX || mode == ORC_SEARCH_EXACT) ? x->ndoc : x->n;
    }
    int64_t it = 0;
    for (uint64_t k = m; k-- > 0;) {
        uint64_t c = tget(pat, k);
        if (c > x->max_character) return -1; /* cs[c] out of bounds => panic in the reference */
        s = orc_lf_map2(x, c, s);
        e = orc_lf_map2(x, c, e);
        it++;
        if (s == e) break;
    }
    *so = s;
    *eo = e;
    return it;
}

int64_t orc_search(const orc_index *x, int mode, const uint8_t *pat, uint64_t m, int use_init,
                   uint64_t init_s, uint64_t init_e, uint64_t *so, uint64_t *eo) {
    tview t = {pat, 1};
    return search_view(x, mode, t, m, use_init, init_s, init_e, so, eo);
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int search_batch_view(const orc_index *x, int mode, tview pat, const uint64_t *off,
                             uint64_t npat, const uint64_t *is, const uint64_t *ie, uint64_t *s,
                             uint64_t *e, uint32_t *steps, int nthreads) {
    int bad = 0;
    if (nthreads <= 0) nthreads = orc_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(| : bad)
    for (int64_t p = 0; p < (int64_t)npat; p++) {
        int64_t it = search_view(x, mode, tview_at(pat, off[p]), off[p + 1] - off[p], is != NULL,
                                 is ? is[p] : 0, ie ? ie[p] : 0, &s[p], &e[p]);
        if (it < 0) {
            bad |= 1;
            it = 0;
        }
        if (steps) steps[p] = (uint32_t)it;
    }
    return bad ? -1 : 0;
}

int orc_search_batch(const orc_index *x, int mode, const uint8_t *pat, const uint64_t *off,
                     uint64_t npat, const uint64_t *is, const uint64_t *ie, uint64_t *s,
                     uint64_t *e, uint32_t *steps, int nthreads) {
    tview t = {pat, 1};
    return search_batch_view(x, mode, t, off, npat, is, ie, s, e, steps, nthreads);
}

/* patterns of u16 / u32 / u64 characters; `off` counts characters */
int orc_search_batch_w(const orc_index *x, int mode, const void *pat, uint32_t char_width, const uint64_t *off,
                       uint64_t npat, const uint64_t *is, const uint64_t *ie, uint64_t *s,
                       uint64_t *e, uint32_t *steps, int nthreads) {
    tview t = {pat, char_width};
    return search_batch_view(x, mode, t, off, npat, is, ie, s, e, steps, nthreads);
}

int orc_locate_batch(const orc_index *x, int prefix_only, const uint64_t *s, const uint64_t *e,
                     uint64_t npat, uint64_t *hit_off, uint64_t *positions, uint64_t *piece_ids,
                     uint64_t *lf_steps, int nthreads) {
    if (nthreads <= 0) nthreads = orc_max_threads();
    if (!positions && !piece_ids) {
        /* pass 1: counts (wrapper.rs:206-216: the L==0 filter applies only when iterating) */
#pragma omp parallel for schedule(static) num_threads(nthreads)
        for (int64_t p = 0; p < (int64_t)npat; p++) {
            uint64_t c = 0;
            if (e[p] > s[p]) {
                if (!prefix_only) {
                    c = e[p] - s[p];
                } else {
                    for (uint64_t i = s[p]; i < e[p]; i++) c += orc_get_l(x, i) == 0;
                }
            }
            hit_off[p + 1] = c;
        }
        hit_off[0] = 0;
        for (uint64_t p = 0; p < npat; p++) hit_off[p + 1] += hit_off[p];
        return 0;
    }
    uint64_t total_steps = 0;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads) reduction(+ : total_steps)
    for (int64_t p = 0; p < (int64_t)npat; p++) {
        uint64_t o = hit_off[p];
        for (uint64_t i = s[p]; i < e[p]; i++) {
            if (prefix_only && orc_get_l(x, i) != 0) continue;
            if (positions) {
                uint64_t st = 0;
                positions[o] = orc_get_sa(x, i, &st);
                total_steps += st;
            }
            if (piece_ids) piece_ids[o] = orc_piece_id(x, i);
            o++;
        }
    }
    if (lf_steps) *lf_steps = total_steps;
    return 0;
}

/* wrapper.rs:154-161 (backward: c = get_l(i); i = lf_map(i)) and :175-183
 * (forward: c = get_f(i); i = fl_map(i)?; yield c) */
static void extract_batch_w(const orc_index *x, const uint64_t *rows, uint64_t nrows, uint32_t k,
                            int forward, void *out, uint32_t w, uint32_t *out_len, int nthreads) {
    if (nthreads <= 0) nthreads = orc_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t r = 0; r < (int64_t)nrows; r++) {
        uint64_t i = rows[r];
        uint8_t *o = (uint8_t *)out + (uint64_t)r * k * w;
        uint32_t got = 0;
        memset(o, 0, (size_t)k * w);
        for (uint32_t t = 0; t < k; t++) {
            if (!forward) {
                tput(o, w, t, orc_get_l(x, i));
                i = orc_lf_map(x, i);
            } else {
                uint64_t c = orc_get_f(x, i);
                uint64_t nx = orc_fl_map(x, i);
                if (nx == ORC_NONE) break;
                i = nx;
                tput(o, w, t, c);
            }
            got++;
        }
        if (out_len) out_len[r] = got;
    }
}

void orc_extract_batch(const orc_index *x, const uint64_t *rows, uint64_t nrows, uint32_t k,
                       int forward, uint8_t *out, uint32_t *out_len, int nthreads) {
    extract_batch_w(x, rows, nrows, k, forward, out, 1, out_len, nthreads);
}

/* the same with characters of char_width bytes in `out` (k characters per row) */
void orc_extract_batch_w(const orc_index *x, const uint64_t *rows, uint64_t nrows, uint32_t k,
                         int forward, void *out, uint32_t char_width, uint32_t *out_len, int nthreads) {
    extract_batch_w(x, rows, nrows, k, forward, out, char_width, out_len, nthreads);
}
