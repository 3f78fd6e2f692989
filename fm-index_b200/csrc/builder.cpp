// builder.cpp -- host-side index construction: text -> device-layout blob.
//
// Restates the host-side producers of the hot path's inputs:
//   sais.rs:9-32 (cs), sais.rs:115-144 (validation + SA), fm_index.rs:44-58 (BWT into a
//   wavelet matrix), sample.rs:21-44 (suffix-order sampling), rlfmi.rs:37-96 (run heads, b, bp,
//   run-count cs), multi_pieces.rs:53-97 (doc array, sa_idx_first_text)
// and lays the result out for the GPU (fmx_layout.h).  Construction is not the hot path.
#include "builder.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/fmx.h"
#include "sais.hpp"

namespace fmx {

static inline uint32_t log2_u64(uint64_t x) { return 63u - (uint32_t)__builtin_clzll(x); }  // util.rs:1-3

// sais.rs:121-139
static int validate_text(const uint8_t *text, uint64_t n, std::string &err) {
    if (n <= 1) return 0;
    if (text[0] == 0) {
        err = "the given text must not start with zero character";
        return FMX_ERR_INVALID_TEXT;
    }
    if (!(text[n - 1] == 0 && text[n - 2] != 0)) {
        err = "the given text must end with exactly one zero character";
        return FMX_ERR_INVALID_TEXT;
    }
    return 0;
}

// Suffix array as u32 (n < 2^32).
static void suffix_array_u32(const uint8_t *text, uint64_t n, std::vector<uint32_t> &sa) {
    sa.resize(n);
    if (n == 0) return;
    // the trailing \0 is a unique minimum iff the text has no interior zero
    bool unique = text[n - 1] == 0 && std::memchr(text, 0, n - 1) == nullptr;
    if (n < (1ull << 31) - 2) {
        suffix_array_bytes<int32_t>(text, n, reinterpret_cast<int32_t *>(sa.data()), unique);
    } else {
        std::vector<int64_t> tmp(n);
        suffix_array_bytes<int64_t>(text, n, tmp.data(), unique);
        for (uint64_t i = 0; i < n; i++) sa[i] = (uint32_t)tmp[i];
    }
}

// Suffix array of a sequence of symbol ranks 0 .. nsym-1 (wide-character texts), n < 2^32 - 1.  The sequence need not
// end in a unique minimum: a virtual sentinel is appended, as for byte texts with interior zeros.
static void suffix_array_ranks(const uint32_t *rk, uint64_t n, uint64_t nsym, uint32_t *sa) {
    if (n + 1 < (1ull << 31) - 2) {
        std::vector<int32_t> tmp(n + 1);
        SentinelAcc32 a{rk, (int64_t)n + 1};
        sais_core<SentinelAcc32, int32_t>(a, tmp.data(), (int32_t)(n + 1), (int32_t)(nsym + 1));
        for (uint64_t i = 0; i < n; i++) sa[i] = (uint32_t)tmp[i + 1];
    } else {
        std::vector<int64_t> tmp(n + 1);
        SentinelAcc32 a{rk, (int64_t)n + 1};
        sais_core<SentinelAcc32, int64_t>(a, tmp.data(), (int64_t)(n + 1), (int64_t)(nsym + 1));
        for (uint64_t i = 0; i < n; i++) sa[i] = (uint32_t)tmp[i + 1];
    }
}

int build_suffix_array(const uint8_t *text, uint64_t n, uint64_t *sa_out, std::string &err) {
    int rc = validate_text(text, n, err);
    if (rc) return rc;
    if (n >= (1ull << 32) - 1) {
        err = "text length must be below 2^32 - 1";
        return FMX_ERR_UNSUPPORTED;
    }
    std::vector<uint32_t> sa;
    suffix_array_u32(text, n, sa);
    for (uint64_t i = 0; i < n; i++) sa_out[i] = sa[i];
    return 0;
}

// ascending positions i with seq[i] == 0 (the \0 terminators: few), found by all cores
static std::vector<uint32_t> zero_positions(const uint8_t *seq, uint64_t n) {
    const uint64_t chunk = 1ull << 22;
    const int64_t nchunks = (int64_t)((n + chunk - 1) / chunk);
    std::vector<std::vector<uint32_t>> part((size_t)nchunks);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c = 0; c < nchunks; c++) {
        const uint64_t lo = (uint64_t)c * chunk, hi = lo + chunk < n ? lo + chunk : n;
        const uint8_t *q = seq + lo;
        uint64_t len = hi - lo, off = 0;
        while (off < len) {
            const void *hit = std::memchr(q + off, 0, len - off);
            if (!hit) break;
            off = (uint64_t)(static_cast<const uint8_t *>(hit) - q);
            part[(size_t)c].push_back((uint32_t)(lo + off));
            off++;
        }
    }
    std::vector<uint32_t> out;
    for (auto &v : part) out.insert(out.end(), v.begin(), v.end());
    return out;
}

// One rank-able bit vector in RB192 form (fmx_layout.h).
struct RBVec {
    uint64_t nbits, nblk;
    std::vector<uint32_t> w;
    explicit RBVec(uint64_t nbits_ = 0) : nbits(nbits_), nblk(nbits_ / FMX_RB_BITS + 1), w(8 * (nbits_ / FMX_RB_BITS + 1), 0) {}
    inline void set(uint64_t pos) {
        uint64_t b = pos / FMX_RB_BITS;
        uint32_t r = (uint32_t)(pos - b * FMX_RB_BITS);
        w[b * 8 + 2 + (r >> 5)] |= 1u << (r & 31);
    }
    static inline uint32_t pop64(const uint32_t *p) { return (uint32_t)(__builtin_popcount(p[0]) + __builtin_popcount(p[1])); }
    uint64_t finish() {
        uint64_t acc = 0;
        for (uint64_t b = 0; b < nblk; b++) {
            uint32_t *blk = &w[b * 8];
            uint32_t p0 = pop64(blk + 2), p1 = pop64(blk + 4), p2 = pop64(blk + 6);
            blk[0] = (uint32_t)acc;
            blk[1] = (p0 << 8) | ((p0 + p1) << 16);
            acc += p0 + p1 + p2;
        }
        return acc;
    }
    uint64_t rank1(uint64_t pos) const {
        uint64_t b = pos / FMX_RB_BITS;
        uint32_t r = (uint32_t)(pos - b * FMX_RB_BITS);
        const uint32_t *blk = &w[b * 8];
        uint32_t j = r >> 6, x = r & 63;
        uint64_t word = (uint64_t)blk[2 + 2 * j] | ((uint64_t)blk[3 + 2 * j] << 32);
        uint64_t c = (uint64_t)blk[0] + ((blk[1] >> (8 * j)) & 0xFF);
        if (x) c += (uint64_t)__builtin_popcountll(word & ((1ull << x) - 1));
        return c;
    }
    uint64_t bytes() const { return w.size() * 4; }
};

struct WMat {
    uint32_t L = 0;
    uint64_t n = 0;
    std::vector<RBVec> lv;
    std::vector<uint64_t> zeros;
    // walk of position `pos` along symbol c's path
    uint64_t walk(uint64_t pos, uint32_t c) const {
        for (uint32_t l = 0; l < L; l++) {
            uint64_t ones = lv[l].rank1(pos);
            pos = ((c >> (L - 1 - l)) & 1) ? zeros[l] + ones : pos - ones;
        }
        return pos;
    }
};

// vers WaveletMatrix::from_slice semantics: level 0 = MSB, stable zero/one partition per level
template <class T>
static void build_wavelet(const T *seq, uint64_t n, uint32_t L, WMat &m) {
    m.L = L;
    m.n = n;
    m.lv.clear();
    m.lv.reserve(L);
    m.zeros.assign(L, 0);
    std::vector<T> cur(seq, seq + n), nxt(n);
    // chunks of whole rank blocks so that threads never share a payload word
    const uint64_t chunk = (uint64_t)FMX_RB_BITS * 4096;
    const int64_t nchunks = (int64_t)((n + chunk - 1) / chunk);
    std::vector<uint64_t> zc((size_t)nchunks + 1, 0);
    for (uint32_t l = 0; l < L; l++) {
        uint32_t sh = L - 1 - l;
        m.lv.emplace_back(n);
        RBVec &v = m.lv.back();
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < nchunks; c++) {
            uint64_t lo = (uint64_t)c * chunk, hi = lo + chunk < n ? lo + chunk : n, z = 0;
            for (uint64_t i = lo; i < hi; i++) {
                if ((cur[i] >> sh) & 1) v.set(i); else z++;
            }
            zc[(size_t)c + 1] = z;
        }
        v.finish();
        zc[0] = 0;
        for (int64_t c = 0; c < nchunks; c++) zc[(size_t)c + 1] += zc[(size_t)c];
        uint64_t z = zc[(size_t)nchunks];
        m.zeros[l] = z;
        if (l + 1 < L) {
#pragma omp parallel for schedule(static)
            for (int64_t c = 0; c < nchunks; c++) {
                uint64_t lo = (uint64_t)c * chunk, hi = lo + chunk < n ? lo + chunk : n;
                uint64_t p0 = zc[(size_t)c], p1 = z + (lo - zc[(size_t)c]);
                for (uint64_t i = lo; i < hi; i++) {
                    if ((cur[i] >> sh) & 1) nxt[p1++] = cur[i]; else nxt[p0++] = cur[i];
                }
            }
            cur.swap(nxt);
        }
    }
}

// One quaternary level (fmx_layout.h): cnt[4] + 64 two-bit codes (as two 64-bit planes) per 32-byte block.
struct Q4Vec {
    std::vector<uint32_t> w;
    std::vector<uint32_t> exc;
    void build(const uint8_t *seq, uint64_t n) {
        const uint64_t nblk = n / 64 + 1;
        w.assign(nblk * 8, 0);
        exc.clear();
        // pass 1: payload + per-block code counts (independent per block)
#pragma omp parallel for schedule(static)
        for (int64_t b = 0; b < (int64_t)nblk; b++) {
            uint32_t *blk = &w[(uint64_t)b * 8];
            uint32_t cnt[4] = {0, 0, 0, 0};
            uint64_t lo = (uint64_t)b * 64, hi = lo + 64 < n ? lo + 64 : n;
            for (uint64_t i = lo; i < hi; i++) {
                uint32_t sym = seq[i];
                uint32_t code = sym ? sym - 1 : 0;
                uint32_t t = (uint32_t)(i - lo);
                blk[4 + (t >> 5)] |= (code & 1u) << (t & 31);  // bit planes: words 4,5 = low bits, 6,7 = high bits
                blk[6 + (t >> 5)] |= (code >> 1) << (t & 31);
                cnt[code]++;
            }
            for (int c = 0; c < 4; c++) blk[c] = cnt[c];
        }
        // pass 2: exclusive prefix of the counts
        uint32_t acc[4] = {0, 0, 0, 0};
        for (uint64_t b = 0; b < nblk; b++) {
            uint32_t *blk = &w[b * 8];
            for (int c = 0; c < 4; c++) {
                uint32_t v = blk[c];
                blk[c] = acc[c];
                acc[c] += v;
            }
        }
        exc = zero_positions(seq, n);
    }
};

// Quaternary wavelet matrix (fmx_layout.h): one Q4-style level per two bits of the symbol.
struct WM4 {
    uint32_t Lq = 0;
    std::vector<std::vector<uint32_t>> lv;  // blocks per level
    uint64_t qoff[FMX_MAX_QLEVELS][4] = {};
    static inline uint32_t digit(uint32_t c, uint32_t Lq, uint32_t l) { return (c >> (2 * (Lq - 1 - l))) & 3u; }
    void build(const uint8_t *seq, uint64_t n, uint32_t L) {
        Lq = (L + 1) / 2;
        lv.assign(Lq, {});
        std::vector<uint8_t> cur(seq, seq + n), nxt(n);
        const uint64_t nblk = n / 64 + 1;
        const uint64_t chunk = 64ull * 8192;
        const int64_t nchunks = (int64_t)((n + chunk - 1) / chunk);
        std::vector<uint64_t> cc((size_t)(nchunks + 1) * 4, 0);
        for (uint32_t l = 0; l < Lq; l++) {
            std::vector<uint32_t> &w = lv[l];
            w.assign(nblk * 8, 0);
#pragma omp parallel for schedule(static)
            for (int64_t b = 0; b < (int64_t)nblk; b++) {
                uint32_t *blk = &w[(uint64_t)b * 8];
                uint32_t cnt[4] = {0, 0, 0, 0};
                uint64_t lo = (uint64_t)b * 64, hi = lo + 64 < n ? lo + 64 : n;
                for (uint64_t i = lo; i < hi; i++) {
                    uint32_t d = digit(cur[i], Lq, l), t = (uint32_t)(i - lo);
                    blk[4 + (t >> 5)] |= (d & 1u) << (t & 31);
                    blk[6 + (t >> 5)] |= (d >> 1) << (t & 31);
                    cnt[d]++;
                }
                for (int c = 0; c < 4; c++) blk[c] = cnt[c];
            }
            uint32_t acc[4] = {0, 0, 0, 0};
            for (uint64_t b = 0; b < nblk; b++) {
                uint32_t *blk = &w[b * 8];
                for (int c = 0; c < 4; c++) {
                    uint32_t v = blk[c];
                    blk[c] = acc[c];
                    acc[c] += v;
                }
            }
            uint64_t o = 0;
            for (int d = 0; d < 4; d++) {
                qoff[l][d] = o;
                o += acc[d];
            }
            if (l + 1 < Lq) {  // stable partition by digit
#pragma omp parallel for schedule(static)
                for (int64_t c = 0; c < nchunks; c++) {
                    uint64_t lo = (uint64_t)c * chunk, hi = lo + chunk < n ? lo + chunk : n, k[4] = {0, 0, 0, 0};
                    for (uint64_t i = lo; i < hi; i++) k[digit(cur[i], Lq, l)]++;
                    for (int d = 0; d < 4; d++) cc[(size_t)(c + 1) * 4 + d] = k[d];
                }
                for (int d = 0; d < 4; d++) cc[d] = qoff[l][d];
                for (int64_t c = 0; c < nchunks; c++)
                    for (int d = 0; d < 4; d++) cc[(size_t)(c + 1) * 4 + d] += cc[(size_t)c * 4 + d];
#pragma omp parallel for schedule(static)
                for (int64_t c = 0; c < nchunks; c++) {
                    uint64_t lo = (uint64_t)c * chunk, hi = lo + chunk < n ? lo + chunk : n;
                    uint64_t p[4] = {cc[(size_t)c * 4], cc[(size_t)c * 4 + 1], cc[(size_t)c * 4 + 2], cc[(size_t)c * 4 + 3]};
                    for (uint64_t i = lo; i < hi; i++) nxt[p[digit(cur[i], Lq, l)]++] = cur[i];
                }
                cur.swap(nxt);
            }
        }
    }
    // occurrences of digit d among the first pos entries of level l
    uint64_t rank(uint32_t l, uint64_t pos, uint32_t d) const {
        const uint32_t *blk = &lv[l][(pos / 64) * 8];
        uint64_t c = blk[d];
        for (uint32_t t = 0; t < pos % 64; t++)
            c += ((((blk[4 + (t >> 5)] >> (t & 31)) & 1u) | (((blk[6 + (t >> 5)] >> (t & 31)) & 1u) << 1)) == d);
        return c;
    }
    uint64_t walk(uint64_t pos, uint32_t c) const {
        for (uint32_t l = 0; l < Lq; l++) {
            uint32_t d = digit(c, Lq, l);
            pos = qoff[l][d] + rank(l, pos, d);
        }
        return pos;
    }
};

static bool force_binary_wavelet() {
    const char *force = std::getenv("FMX_FORCE_WAVELET");
    return force && force[0] && force[0] != '0';
}

// Free memory of the device the index is being built for (0 = unknown / host-only build): the HBM-for-requests trades
// below are sized for a B200's 180 GB and shrink with it on a smaller device instead of failing with FMX_ERR_OOM.
static std::atomic<uint64_t> g_device_free_hint{0};
void set_device_memory_hint(uint64_t free_bytes) { g_device_free_hint.store(free_bytes); }

// Per-symbol RB192 bit vectors (fmx_layout.h "SYM"), built in place inside the blob.
static uint64_t sym_budget_bytes() {
    const char *e = std::getenv("FMX_SYM_BUDGET_MB");
    if (e && e[0]) return (uint64_t)std::strtoull(e, nullptr, 10) << 20;
    uint64_t b = 49152ull << 20;
    const uint64_t hint = g_device_free_hint.load();
    if (hint && b > hint / 100 * 45) b = hint / 100 * 45;
    return b;
}
static uint64_t sym_bytes(uint64_t cs_len, uint64_t n) { return cs_len * (n / FMX_RB_BITS + 1) * 32; }
static void sym_build(const uint8_t *seq, uint64_t n, uint32_t cs_len, uint32_t *dst) {
    const uint64_t nblk = n / FMX_RB_BITS + 1;
    const uint64_t chunk_blocks = 2048;
    const int64_t nchunks = (int64_t)((nblk + chunk_blocks - 1) / chunk_blocks);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t ch = 0; ch < nchunks; ch++) {  // a thread owns whole blocks (of every symbol): no shared words
        uint64_t lo = (uint64_t)ch * chunk_blocks * FMX_RB_BITS, hi = lo + chunk_blocks * FMX_RB_BITS;
        if (hi > n) hi = n;
        for (uint64_t i = lo; i < hi; i++) {
            uint64_t b = i / FMX_RB_BITS;
            uint32_t r = (uint32_t)(i - b * FMX_RB_BITS);
            dst[((uint64_t)seq[i] * nblk + b) * 8 + 2 + (r >> 5)] |= 1u << (r & 31);
        }
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c = 0; c < (int64_t)cs_len; c++) {
        uint32_t *v = dst + (uint64_t)c * nblk * 8;
        uint64_t acc = 0;
        for (uint64_t b = 0; b < nblk; b++) {
            uint32_t *blk = v + b * 8;
            uint32_t p0 = RBVec::pop64(blk + 2), p1 = RBVec::pop64(blk + 4), p2 = RBVec::pop64(blk + 6);
            blk[0] = (uint32_t)acc;
            blk[1] = (p0 << 8) | ((p0 + p1) << 16);
            acc += p0 + p1 + p2;
        }
    }
}

static bool want_q4(uint64_t mc, const uint8_t *seq, uint64_t n) {
    if (force_binary_wavelet()) return false;
    if (mc > 4) return false;
    uint64_t zeros = 0;
#pragma omp parallel for schedule(static) reduction(+ : zeros)
    for (int64_t i = 0; i < (int64_t)n; i++) zeros += seq[i] == 0;
    return zeros <= FMX_MAX_EXC;
}

struct SectionData {
    const void *ptr = nullptr;
    uint64_t bytes = 0;
};

HostBlob::~HostBlob() { std::free(p); }

int HostBlob::alloc(uint64_t bytes, bool zero) {
    std::free(p);
    p = static_cast<uint8_t *>(std::malloc(bytes ? bytes : 1));
    n = p ? bytes : 0;
    if (!p) return FMX_ERR_OOM;
    if (zero) {
        const uint64_t piece = 32ull << 20;
        const int64_t np = (int64_t)((bytes + piece - 1) / piece);
#pragma omp parallel for schedule(static)
        for (int64_t q = 0; q < np; q++) {
            const uint64_t lo = (uint64_t)q * piece;
            std::memset(p + lo, 0, bytes - lo < piece ? bytes - lo : piece);
        }
    }
    return 0;
}

// FMX_BUILD_TIMING=1: phase times of build_blob on stderr
struct PhaseTimer {
    bool on;
    std::chrono::steady_clock::time_point t0;
    PhaseTimer() : on(false), t0(std::chrono::steady_clock::now()) {
        const char *e = std::getenv("FMX_BUILD_TIMING");
        on = e && e[0] && e[0] != '0';
    }
    void mark(const char *what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[fmx build] %-28s %8.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    }
};

int resolve_mode(int &mode, std::string &err) {
    if (mode == FMX_MODE_AUTO) {
        const char *m = std::getenv("FMX_MODE");
        if (m && !std::strcmp(m, "compact")) mode = FMX_MODE_COMPACT;
        else if (m && !std::strcmp(m, "rich")) mode = FMX_MODE_RICH;
    }
    if (mode < FMX_MODE_AUTO || mode > FMX_MODE_RICH) {
        err = "unknown index mode";
        return FMX_ERR_INVALID_ARG;
    }
    return 0;
}

// Which of the structures that answer from the suffix array an index gets (fmx_layout.h; include/fmx.h "modes").
VerifyPlan plan_verify(int kind, uint64_t n, int mode, int level, bool interior_zero, uint64_t rank_bytes, bool use_sym) {
    VerifyPlan p;
    bool verify = (kind == FMX_KIND_FM || kind == FMX_KIND_MULTI) && n >= (mode == FMX_MODE_RICH ? 64u : 4096u);
    const char *nv = std::getenv("FMX_NO_VERIFY");
    if (nv && nv[0] && nv[0] != '0') verify = false;
    if (mode == FMX_MODE_COMPACT || interior_zero) verify = false;
    uint64_t budget = 32768ull << 20;
    if (const uint64_t hint = g_device_free_hint.load()) budget = std::min<uint64_t>(budget, hint / 4);
    if (const char *vb = std::getenv("FMX_VERIFY_BUDGET_MB")) budget = std::strtoull(vb, nullptr, 10) << 20;
    // an index whose rank structure sits in the 126 MB L2 answers a step from L2; the tail's three or four
    // DRAM reads are slower than that (measured on the 100 MB DNA config), so it is not built there
    uint64_t min_rank = 192ull << 20;
    if (const char *mr = std::getenv("FMX_VERIFY_MIN_RANK_MB")) min_rank = std::strtoull(mr, nullptr, 10) << 20;
    if (rank_bytes < min_rank && mode != FMX_MODE_RICH) verify = false;
    p.dense = verify && 9 * n <= budget;
    if (verify && !p.dense) verify = kind == FMX_KIND_FM && use_sym && (level < 0 || level <= 3);
    p.verify = verify;
    p.dense_sa = p.dense;
    // RLFM in the HBM-rich mode: no verify tail (its ranges rarely narrow to one row), but locate by the
    // resident suffix array.  AUTO takes it once the text has 2^27 symbols: from there on the sampled walk
    // (2^level - 1 run-length LF steps per hit, each several requests) leaves L2, and on the 1 GiB config the
    // resident array answers 6.1e8 hits in 2.7 ms against 31.9 ms (profiles/r02_c7_bench_cfg3_rlfm_rich.json)
    const bool rl_rich = mode == FMX_MODE_RICH || (mode == FMX_MODE_AUTO && n >= (1ull << 27));
    if (kind == FMX_KIND_RLFM && rl_rich && level >= 0 && !interior_zero && n >= 64 && 4 * n <= budget) p.dense_sa = true;
    p.isa_level = p.dense ? 0u : 2u;
    return p;
}

// section offsets (256-byte aligned, in enum order) and the total size, from the section sizes
void layout_sections(FmxBlobHeader &hdr, const uint64_t bytes[SEC_COUNT]) {
    uint64_t off = (sizeof(FmxBlobHeader) + FMX_SECTION_ALIGN - 1) / FMX_SECTION_ALIGN * FMX_SECTION_ALIGN;
    for (int k = 0; k < (int)SEC_COUNT; k++) {
        hdr.sec[k].offset = bytes[k] ? off : 0;
        hdr.sec[k].bytes = bytes[k];
        off += (bytes[k] + FMX_SECTION_ALIGN - 1) / FMX_SECTION_ALIGN * FMX_SECTION_ALIGN;
    }
    hdr.total_bytes = off;
}

bool q4_forbidden_by_env() { return force_binary_wavelet(); }
// SYM is the layout of a sequence of `len` symbols over cs_len table entries that Q4 does not take
bool sym_layout_chosen(uint64_t cs_len, uint64_t len) { return !force_binary_wavelet() && sym_bytes(cs_len, len) + len <= sym_budget_bytes(); }
uint64_t sym_layout_bytes(uint64_t cs_len, uint64_t len) { return sym_bytes(cs_len, len); }

int build_blob(const uint8_t *text, uint64_t n, uint64_t mc, int kind, int level,
               HostBlob &blob, std::string &err, int sa_device, int mode) {
    PhaseTimer tm;
    if (int mrc = resolve_mode(mode, err)) return mrc;
    if (mc == 0 || mc > 255) {
        err = "max_character must be in 1..=255 for u8 texts";
        return FMX_ERR_INVALID_ARG;
    }
    if (kind != FMX_KIND_FM && kind != FMX_KIND_RLFM && kind != FMX_KIND_MULTI) {
        err = "unknown index kind";
        return FMX_ERR_INVALID_ARG;
    }
    if (n >= (1ull << 32) - 1) {
        err = "text length must be below 2^32 - 1 (u32 rank counts in the device layout)";
        return FMX_ERR_UNSUPPORTED;
    }
    {
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
        for (int64_t i = 0; i < (int64_t)n; i++) bad |= text[i] > mc;
        if (bad) {  // sais.rs:16-18 would index occs out of bounds (panic)
            err = "text contains a character larger than max_character";
            return FMX_ERR_INVALID_ARG;
        }
    }
    int rc = validate_text(text, n, err);
    if (rc) return rc;
    tm.mark("validate");

    const uint32_t L = log2_u64(mc) + 1;  // text.rs:61-63
    const uint32_t cs_len = (uint32_t)mc + 1;

    std::vector<uint32_t> sa;
    {
        const char *host_only = std::getenv("FMX_HOST_SA");
        bool use_gpu = sa_device >= 0 && n >= (1u << 16) && !(host_only && host_only[0] && host_only[0] != '0');
        if (use_gpu) {
            sa.resize(n);
            std::string gerr;
            int grc = gpu_suffix_array(text, n, L, sa_device, sa.data(), nullptr, gerr);
            if (grc == FMX_ERR_OOM) {  // the device has no room for the sorter's working set: sort on the host
                suffix_array_u32(text, n, sa);
            } else if (grc) {
                err = "GPU suffix array construction failed: " + gerr;
                return grc;
            }
        } else {
            suffix_array_u32(text, n, sa);
        }
    }

    tm.mark("suffix array");
    // BWT: bw[i] = text[sa[i]-1], 0 when sa[i] == 0 (fm_index.rs:48-55; rlfmi.rs:49-53 uses
    // text[n-1] there, which is the same \0 for every text that passes validation with n >= 2)
    std::vector<uint8_t> bwt(n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) bwt[i] = sa[i] ? text[sa[i] - 1] : (kind == FMX_KIND_RLFM && n ? text[n - 1] : 0);

    tm.mark("bwt");
    FmxBlobHeader hdr;
    std::memset(&hdr, 0, sizeof(hdr));
    hdr.magic = FMX_BLOB_MAGIC;
    hdr.version = FMX_BLOB_VERSION;
    hdr.kind = (uint32_t)kind;
    hdr.n = n;
    hdr.levels = L;
    hdr.max_character = (uint32_t)mc;
    hdr.cs_len = cs_len;
    hdr.char_width = 1;

    WMat wm;
    Q4Vec q4;
    WM4 wm4;
    bool use_q4 = false, use_sym = false;
    const bool use_wm4_else = !force_binary_wavelet();  // every alphabet Q4 / SYM do not take
    std::vector<uint8_t> heads;                          // RLFM: the run heads
    const uint8_t *seq_ptr = bwt.data();                 // the sequence the rank structure is over
    auto pick_sym = [&](uint64_t len) { return !use_q4 && use_wm4_else && sym_bytes(cs_len, len) + len <= sym_budget_bytes(); };
    std::vector<uint32_t> cs(cs_len + 1, 0), adj(cs_len, 0);
    std::vector<uint32_t> doc, piece_end, bsel, bpsel;
    RBVec rb_b(0), rb_bp(0);

    if (kind == FMX_KIND_FM || kind == FMX_KIND_MULTI) {
        std::vector<uint64_t> occ(cs_len, 0);
#pragma omp parallel
        {
            std::vector<uint64_t> mine(cs_len, 0);
#pragma omp for schedule(static) nowait
            for (int64_t i = 0; i < (int64_t)n; i++) mine[text[i]]++;
#pragma omp critical
            for (uint32_t c = 0; c < cs_len; c++) occ[c] += mine[c];
        }
        uint64_t sum = 0;
        for (uint32_t c = 0; c < cs_len; c++) {  // sais.rs:21-32
            cs[c] = (uint32_t)sum;
            sum += occ[c];
        }
        cs[cs_len] = (uint32_t)n;
        use_q4 = want_q4(mc, bwt.data(), n);
        use_sym = pick_sym(n);
        if (use_q4) q4.build(bwt.data(), n);
        else if (use_sym) {}  // built in place once the blob is allocated
        else if (use_wm4_else) wm4.build(bwt.data(), n, L);
        else build_wavelet(bwt.data(), n, L, wm);
        hdr.seq_len = n;
        if (kind == FMX_KIND_MULTI) {
            // multi_pieces.rs:53-79
            piece_end = zero_positions(text, n);
            uint64_t zc = piece_end.size();
            doc.assign(zc, 0);
            uint64_t k = 0;
            for (uint32_t p : zero_positions(bwt.data(), n)) {  // p = select(bw, k, 0)
                if (k >= zc) {               // the reference would index doc out of bounds (panic)
                    err = "text without a \\0 terminator cannot be indexed as multi-pieces";
                    return FMX_ERR_INVALID_TEXT;
                }
                uint64_t em = sa[p] ? sa[p] - 1 : n - 1;  // modular_sub(sa[p], 1, n)
                uint64_t pid = (uint64_t)(std::lower_bound(piece_end.begin(), piece_end.end(), (uint32_t)em) - piece_end.begin());
                if (pid == zc - 1) hdr.first_row = p;
                doc[k++] = (uint32_t)pid;
            }
            hdr.ndoc = zc;
        }
    } else {
        // rlfmi.rs:37-96
        std::vector<uint32_t> starts;
        rb_b = RBVec(n);
        rb_bp = RBVec(n);
        uint32_t c0 = 0;
        for (uint64_t i = 0; i < n; i++) {
            uint32_t c = bwt[i];
            if (c0 != c) {
                heads.push_back((uint8_t)c);
                starts.push_back((uint32_t)i);
                rb_b.set(i);
            } else if (heads.empty()) {
                err = "text not representable by RLFMIndex (the reference hits unreachable!() at rlfmi.rs:62)";
                return FMX_ERR_INVALID_TEXT;
            }
            c0 = c;
        }
        rb_b.finish();
        uint64_t r = heads.size();
        hdr.runs = r;
        hdr.seq_len = r;
        seq_ptr = heads.data();
        use_q4 = want_q4(mc, heads.data(), r);
        use_sym = pick_sym(r);
        if (use_q4) q4.build(heads.data(), r);
        else if (use_sym) {}
        else if (use_wm4_else) wm4.build(heads.data(), r, L);
        else build_wavelet(heads.data(), r, L, wm);
        bsel.assign(r + 1, (uint32_t)n);
        for (uint64_t j = 0; j < r; j++) bsel[j] = starts[j];
        // cs over run heads; bp groups runs by head character, stably
        std::vector<uint64_t> cnt(cs_len, 0), len_by_c(cs_len, 0);
        for (uint64_t j = 0; j < r; j++) {
            cnt[heads[j]]++;
            len_by_c[heads[j]] += (uint64_t)bsel[j + 1] - bsel[j];
        }
        std::vector<uint64_t> next_run(cs_len, 0), next_pos(cs_len, 0);
        uint64_t acc = 0, pacc = 0;
        for (uint32_t c = 0; c < cs_len; c++) {
            cs[c] = (uint32_t)acc;
            next_run[c] = acc;
            next_pos[c] = pacc;
            acc += cnt[c];
            pacc += len_by_c[c];
        }
        cs[cs_len] = (uint32_t)r;
        bpsel.assign(r + 1, (uint32_t)n);
        for (uint64_t j = 0; j < r; j++) {
            uint8_t c = heads[j];
            rb_bp.set(next_pos[c]);
            bpsel[next_run[c]++] = (uint32_t)next_pos[c];
            next_pos[c] += (uint64_t)bsel[j + 1] - bsel[j];
        }
        rb_bp.finish();
    }
    tm.mark("rank structure / runs / doc");
    if (use_q4) {
        hdr.layout = FMX_LAYOUT_QUAT;
        hdr.nexc = (uint32_t)q4.exc.size();
    } else if (use_sym) {
        hdr.layout = FMX_LAYOUT_SYM;
        hdr.sym_nblk = (uint32_t)(hdr.seq_len / FMX_RB_BITS + 1);
    } else if (use_wm4_else) {
        hdr.layout = FMX_LAYOUT_WM4;
        hdr.qlevels = wm4.Lq;
        for (uint32_t l = 0; l < wm4.Lq; l++)
            for (int d = 0; d < 4; d++) hdr.qoff[l][d] = wm4.qoff[l][d];
        for (uint32_t c = 0; c < cs_len; c++) adj[c] = cs[c] - (uint32_t)wm4.walk(0, c);
    } else {
        hdr.layout = FMX_LAYOUT_WAVELET;
        for (uint32_t l = 0; l < L; l++) hdr.zeros[l] = wm.zeros[l];
        for (uint32_t c = 0; c < cs_len; c++) adj[c] = cs[c] - (uint32_t)wm.walk(0, c);
    }

    // seed-and-verify structures (fmx_layout.h): dense (full SA + full ISA) within the budget, else sampled
    // for the SYM layout (text, ISA every 4 positions, SA samples of level <= 3), else none
    std::vector<uint32_t> isa_s;
    // A single-text index (FM, RLFM) over a text with INTERIOR zeros: the reference's lf_map2(0, i) = cs[0] + rank(i, 0)
    // (fm_index.rs:93-95) is not the true LF row there, so neither its walks nor its searches agree with the
    // suffix array; the structures that answer from the suffix array are not built for such texts.
    bool interior_zero = false;
    if (kind != FMX_KIND_MULTI && n >= 2) interior_zero = std::memchr(text, 0, n - 1) != nullptr;
    const uint64_t rank_bytes = use_q4 ? (n / 64 + 1) * 32 : (use_sym ? sym_bytes(cs_len, n) : (uint64_t)L * (n / FMX_RB_BITS + 1) * 32);
    const VerifyPlan vp = plan_verify(kind, n, mode, level, interior_zero, rank_bytes, use_sym);
    const bool verify = vp.verify, verify_dense = vp.dense, dense_sa = vp.dense_sa;
    const uint32_t isa_level = vp.isa_level;
    if (verify) {
        isa_s.assign(((n - 1) >> isa_level) + 1, 0);
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < (int64_t)n; r++)
            if ((sa[r] & ((1u << isa_level) - 1u)) == 0) isa_s[sa[r] >> isa_level] = (uint32_t)r;
        hdr.verify = 1;
        hdr.isa_level = isa_level;
    }
    const bool need_samples_for_verify = verify && !verify_dense;
    tm.mark("adj / inverse suffix array");

    // sample.rs:21-44
    std::vector<uint32_t> samples;
    if (level >= 0 || need_samples_for_verify) {
        hdr.has_locate = level >= 0 ? 1 : 0;  // count-only indexes keep level-2 samples for the sampled verify path only
        uint32_t lvl = level >= 0 ? (uint32_t)level : 2u;
        if (n > 0) {
            hdr.sa_word_size = log2_u64(n) + 1;
            if (lvl >= 63 || n <= (1ull << lvl)) lvl = 0;  // sample.rs:28-31
            hdr.sa_level = lvl;
            hdr.sa_count = ((n - 1) >> lvl) + 1;
            samples.resize(hdr.sa_count);
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < (int64_t)hdr.sa_count; i++) samples[i] = sa[(uint64_t)i << lvl];
        }
    }
    hdr.vsa_level = verify_dense ? 0u : hdr.sa_level;
    hdr.reserved[0] = (uint64_t)mode;
    if (!dense_sa) std::vector<uint32_t>().swap(sa);  // the dense paths keep the full array (SEC_VSA)

    tm.mark("samples");
    // ---- assemble
    SectionData sec[SEC_COUNT];
    if (use_q4) {
        sec[SEC_LEVEL0] = {q4.w.data(), q4.w.size() * 4};
        if (!q4.exc.empty()) sec[SEC_EXC] = {q4.exc.data(), q4.exc.size() * 4};
    } else if (use_sym) {
        sec[SEC_LEVEL0] = {nullptr, sym_bytes(cs_len, hdr.seq_len)};  // filled in place below
        sec[SEC_LEVEL0 + 1] = {seq_ptr, hdr.seq_len ? hdr.seq_len : 1};
    } else if (use_wm4_else) {
        for (uint32_t l = 0; l < wm4.Lq; l++) sec[SEC_LEVEL0 + l] = {wm4.lv[l].data(), wm4.lv[l].size() * 4};
    } else {
        for (uint32_t l = 0; l < L; l++) sec[SEC_LEVEL0 + l] = {wm.lv[l].w.data(), wm.lv[l].bytes()};
    }
    sec[SEC_ADJ] = {adj.data(), adj.size() * 4};
    sec[SEC_CS] = {cs.data(), cs.size() * 4};
    if (!samples.empty()) sec[SEC_SA] = {samples.data(), samples.size() * 4};
    if (verify) {
        sec[SEC_TEXT] = {text, n};
        sec[SEC_ISA] = {isa_s.data(), isa_s.size() * 4};
    } else if (dense_sa) {
        sec[SEC_TEXT] = {text, n};  // RLFM in the HBM-rich mode: extraction reads the text at SA[row] (k_extract_text)
    }
    if (dense_sa) sec[SEC_VSA] = {sa.data(), sa.size() * 4};
    if (!doc.empty()) sec[SEC_DOC] = {doc.data(), doc.size() * 4};
    if (!piece_end.empty()) sec[SEC_PIECE_END] = {piece_end.data(), piece_end.size() * 4};
    if (kind == FMX_KIND_RLFM) {
        sec[SEC_RL_B] = {rb_b.w.data(), rb_b.bytes()};
        sec[SEC_RL_BP] = {rb_bp.w.data(), rb_bp.bytes()};
        sec[SEC_RL_BSEL] = {bsel.data(), bsel.size() * 4};
        sec[SEC_RL_BPSEL] = {bpsel.data(), bpsel.size() * 4};
    }
    uint64_t sec_bytes[SEC_COUNT];
    for (int k = 0; k < (int)SEC_COUNT; k++) sec_bytes[k] = sec[k].bytes;
    layout_sections(hdr, sec_bytes);
    const uint64_t off = hdr.total_bytes;
    if (blob.alloc(off, true)) {
        err = "out of host memory for the index blob";
        return FMX_ERR_OOM;
    }
    tm.mark("allocate blob");
    std::memcpy(blob.p, &hdr, sizeof(hdr));
    for (int k = 0; k < (int)SEC_COUNT; k++) {
        if (!sec[k].bytes || !sec[k].ptr) continue;
        const uint64_t piece = 64ull << 20;  // large sections: copy in parallel
        const int64_t np = (int64_t)((sec[k].bytes + piece - 1) / piece);
#pragma omp parallel for schedule(static) if (np > 1)
        for (int64_t q = 0; q < np; q++) {
            uint64_t lo = (uint64_t)q * piece, len = sec[k].bytes - lo < piece ? sec[k].bytes - lo : piece;
            std::memcpy(blob.p + hdr.sec[k].offset + lo, static_cast<const uint8_t *>(sec[k].ptr) + lo, len);
        }
    }
    tm.mark("copy sections");
    if (use_sym) sym_build(seq_ptr, hdr.seq_len, cs_len, reinterpret_cast<uint32_t *>(blob.p + hdr.sec[SEC_LEVEL0].offset));
    tm.mark("per-symbol vectors");
    return 0;
}


// ------------------------------------------------------------------ texts of wide characters (character.rs:38-42)
// max_character > 255: the WIDE layout (fmx_layout.h).  Same producers as build_blob -- suffix array (sais.rs:115-144),
// cs (sais.rs:9-32), BWT (fm_index.rs:44-58), runs (rlfmi.rs:37-96), doc (multi_pieces.rs:53-97), samples
// (sample.rs:21-44) -- over u32 symbols.  No k-mer tables and no verify structures: those are sized for small alphabets.
int build_blob_wide(const void *text_in, uint32_t cw, uint64_t n, uint64_t mc, int kind, int level, HostBlob &blob,
                    std::string &err, int mode) {
    if (int mrc = resolve_mode(mode, err)) return mrc;
    if (cw != 2 && cw != 4 && cw != 8) {
        err = "character width must be 1, 2, 4 or 8 bytes";
        return FMX_ERR_INVALID_ARG;
    }
    if (mc <= 255 || mc >= 0xFFFFFFFFull || (cw == 2 && mc > 0xFFFF)) {
        err = "max_character out of range for this character width (wide texts: 256 ..= 2^32 - 2)";
        return FMX_ERR_INVALID_ARG;
    }
    if (kind != FMX_KIND_FM && kind != FMX_KIND_RLFM && kind != FMX_KIND_MULTI) {
        err = "unknown index kind";
        return FMX_ERR_INVALID_ARG;
    }
    if (n >= (1ull << 32) - 1) {
        err = "text length must be below 2^32 - 1 (u32 rank counts in the device layout)";
        return FMX_ERR_UNSUPPORTED;
    }
    std::vector<uint32_t> t(n);
    {
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
        for (int64_t i = 0; i < (int64_t)n; i++) {
            uint64_t v = cw == 2 ? static_cast<const uint16_t *>(text_in)[i]
                                 : (cw == 4 ? static_cast<const uint32_t *>(text_in)[i] : static_cast<const uint64_t *>(text_in)[i]);
            bad |= v > mc;
            t[(size_t)i] = (uint32_t)v;
        }
        if (bad) {  // sais.rs:16-18 would index occs out of bounds (panic)
            err = "text contains a character larger than max_character";
            return FMX_ERR_INVALID_ARG;
        }
    }
    if (n > 1) {  // sais.rs:121-139
        if (t[0] == 0) {
            err = "the given text must not start with zero character";
            return FMX_ERR_INVALID_TEXT;
        }
        if (!(t[n - 1] == 0 && t[n - 2] != 0)) {
            err = "the given text must end with exactly one zero character";
            return FMX_ERR_INVALID_TEXT;
        }
    }
    const uint32_t L = log2_u64(mc) + 1;  // text.rs:61-63
    const uint64_t cs_len = mc + 1;

    // the symbols that occur, ascending; suffix sorting runs over their ranks (order preserving, so the same array)
    std::vector<uint32_t> present(t);
    std::sort(present.begin(), present.end());
    present.erase(std::unique(present.begin(), present.end()), present.end());
    std::vector<uint32_t> sa(n);
    if (n == 1) sa[0] = 0;
    if (n > 1) {
        std::vector<uint32_t> rk(n);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; i++)
            rk[(size_t)i] = (uint32_t)(std::lower_bound(present.begin(), present.end(), t[(size_t)i]) - present.begin());
        suffix_array_ranks(rk.data(), n, (uint64_t)present.size(), sa.data());
    }
    std::vector<uint32_t> bwt(n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) bwt[(size_t)i] = sa[(size_t)i] ? t[sa[(size_t)i] - 1] : (kind == FMX_KIND_RLFM && n ? t[n - 1] : 0u);

    FmxBlobHeader hdr;
    std::memset(&hdr, 0, sizeof(hdr));
    hdr.magic = FMX_BLOB_MAGIC;
    hdr.version = FMX_BLOB_VERSION;
    hdr.kind = (uint32_t)kind;
    hdr.n = n;
    hdr.levels = L;
    hdr.max_character = (uint32_t)mc;
    hdr.cs_len = (uint32_t)cs_len;
    hdr.layout = FMX_LAYOUT_WIDE;
    hdr.char_width = cw;
    hdr.reserved[0] = (uint64_t)FMX_MODE_COMPACT;

    std::vector<uint32_t> cs, adj;
    try {
        cs.assign(cs_len + 1, 0);
        adj.assign(cs_len, 0);
    } catch (const std::bad_alloc &) {
        err = "out of host memory for the character tables (max_character + 1 words each)";
        return FMX_ERR_OOM;
    }
    WMat wm;
    std::vector<uint32_t> heads, doc, piece_end, bsel, bpsel;
    RBVec rb_b(0), rb_bp(0);
    auto fill_cs = [&](const std::vector<uint32_t> &seq) {  // cs[c] = #symbols of seq below c (sais.rs:21-32)
        for (uint32_t v : seq) cs[(size_t)v + 1]++;          // counts, shifted by one ...
        uint64_t sum = 0;
        for (uint64_t c = 0; c <= cs_len; c++) {             // ... then the running sum in place
            sum += cs[c];
            cs[c] = (uint32_t)sum;
        }
    };
    if (kind == FMX_KIND_FM || kind == FMX_KIND_MULTI) {
        fill_cs(t);
        build_wavelet(bwt.data(), n, L, wm);
        hdr.seq_len = n;
        if (kind == FMX_KIND_MULTI) {  // multi_pieces.rs:53-79
            for (uint64_t i = 0; i < n; i++)
                if (t[i] == 0) piece_end.push_back((uint32_t)i);
            const uint64_t zc = piece_end.size();
            doc.assign(zc, 0);
            uint64_t k = 0;
            for (uint64_t p = 0; p < n; p++) {
                if (bwt[p] != 0) continue;  // p = select(bw, k, 0)
                if (k >= zc) {
                    err = "text without a \\0 terminator cannot be indexed as multi-pieces";
                    return FMX_ERR_INVALID_TEXT;
                }
                const uint64_t em = sa[p] ? sa[p] - 1 : n - 1;  // modular_sub(sa[p], 1, n)
                const uint64_t pid = (uint64_t)(std::lower_bound(piece_end.begin(), piece_end.end(), (uint32_t)em) - piece_end.begin());
                if (pid == zc - 1) hdr.first_row = p;
                doc[k++] = (uint32_t)pid;
            }
            hdr.ndoc = zc;
        }
    } else {  // rlfmi.rs:37-96
        std::vector<uint32_t> starts;
        rb_b = RBVec(n);
        rb_bp = RBVec(n);
        uint32_t c0 = 0;
        for (uint64_t i = 0; i < n; i++) {
            const uint32_t c = bwt[i];
            if (c0 != c) {
                heads.push_back(c);
                starts.push_back((uint32_t)i);
                rb_b.set(i);
            } else if (heads.empty()) {
                err = "text not representable by RLFMIndex (the reference hits unreachable!() at rlfmi.rs:62)";
                return FMX_ERR_INVALID_TEXT;
            }
            c0 = c;
        }
        rb_b.finish();
        const uint64_t r = heads.size();
        hdr.runs = r;
        hdr.seq_len = r;
        build_wavelet(heads.data(), r, L, wm);
        bsel.assign(r + 1, (uint32_t)n);
        for (uint64_t j = 0; j < r; j++) bsel[j] = starts[j];
        fill_cs(heads);  // cs over run heads
        // bp groups the runs by head character, stably: position of the group of c = total length of the runs of smaller heads
        std::vector<uint64_t> len_of(present.size() + 1, 0), next_pos(present.size() + 1, 0), next_run(present.size() + 1, 0);
        auto slot = [&](uint32_t c) { return (size_t)(std::lower_bound(present.begin(), present.end(), c) - present.begin()); };
        for (uint64_t j = 0; j < r; j++) len_of[slot(heads[j])] += (uint64_t)bsel[j + 1] - bsel[j];
        uint64_t pacc = 0;
        for (size_t q = 0; q < present.size(); q++) {
            next_pos[q] = pacc;
            next_run[q] = cs[present[q]];
            pacc += len_of[q];
        }
        bpsel.assign(r + 1, (uint32_t)n);
        for (uint64_t j = 0; j < r; j++) {
            const size_t q = slot(heads[j]);
            rb_bp.set(next_pos[q]);
            bpsel[next_run[q]++] = (uint32_t)next_pos[q];
            next_pos[q] += (uint64_t)bsel[j + 1] - bsel[j];
        }
        rb_bp.finish();
    }
    // adj[c] = cs[c] - walk_c(0), for the symbols that occur (the kernels never read the others)
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < (int64_t)present.size(); q++) {
        const uint32_t c = present[(size_t)q];
        adj[c] = cs[c] - (uint32_t)wm.walk(0, c);
    }
    std::vector<uint32_t> wzeros(L);
    for (uint32_t l = 0; l < L; l++) wzeros[l] = (uint32_t)wm.zeros[l];

    std::vector<uint32_t> samples;  // sample.rs:21-44
    if (level >= 0) {
        hdr.has_locate = 1;
        uint32_t lvl = (uint32_t)level;
        if (n > 0) {
            hdr.sa_word_size = log2_u64(n) + 1;
            if (lvl >= 63 || n <= (1ull << lvl)) lvl = 0;  // sample.rs:28-31
            hdr.sa_level = lvl;
            hdr.sa_count = ((n - 1) >> lvl) + 1;
            samples.resize(hdr.sa_count);
            for (uint64_t i = 0; i < hdr.sa_count; i++) samples[i] = sa[i << lvl];
        }
    }
    hdr.vsa_level = hdr.sa_level;

    const uint64_t level_bytes = (hdr.seq_len / FMX_RB_BITS + 1) * 32;
    uint64_t sec_bytes[SEC_COUNT];
    std::memset(sec_bytes, 0, sizeof(sec_bytes));
    sec_bytes[SEC_LEVEL0] = level_bytes * L;
    sec_bytes[SEC_WZEROS] = (uint64_t)L * 4;
    sec_bytes[SEC_ADJ] = adj.size() * 4;
    sec_bytes[SEC_CS] = cs.size() * 4;
    sec_bytes[SEC_SA] = samples.size() * 4;
    sec_bytes[SEC_DOC] = doc.size() * 4;
    sec_bytes[SEC_PIECE_END] = piece_end.size() * 4;
    if (kind == FMX_KIND_RLFM) {
        sec_bytes[SEC_RL_B] = rb_b.bytes();
        sec_bytes[SEC_RL_BP] = rb_bp.bytes();
        sec_bytes[SEC_RL_BSEL] = bsel.size() * 4;
        sec_bytes[SEC_RL_BPSEL] = bpsel.size() * 4;
    }
    layout_sections(hdr, sec_bytes);
    if (blob.alloc(hdr.total_bytes, true)) {
        err = "out of host memory for the index blob";
        return FMX_ERR_OOM;
    }
    std::memcpy(blob.p, &hdr, sizeof(hdr));
    auto put = [&](int k, const void *src, uint64_t bytes) {
        if (bytes) std::memcpy(blob.p + hdr.sec[k].offset, src, bytes);
    };
    for (uint32_t l = 0; l < L; l++) std::memcpy(blob.p + hdr.sec[SEC_LEVEL0].offset + l * level_bytes, wm.lv[l].w.data(), level_bytes);
    put(SEC_WZEROS, wzeros.data(), wzeros.size() * 4);
    put(SEC_ADJ, adj.data(), adj.size() * 4);
    put(SEC_CS, cs.data(), cs.size() * 4);
    put(SEC_SA, samples.data(), samples.size() * 4);
    put(SEC_DOC, doc.data(), doc.size() * 4);
    put(SEC_PIECE_END, piece_end.data(), piece_end.size() * 4);
    if (kind == FMX_KIND_RLFM) {
        put(SEC_RL_B, rb_b.w.data(), rb_b.bytes());
        put(SEC_RL_BP, rb_bp.w.data(), rb_bp.bytes());
        put(SEC_RL_BSEL, bsel.data(), bsel.size() * 4);
        put(SEC_RL_BPSEL, bpsel.data(), bpsel.size() * 4);
    }
    return 0;
}

// Suffix array of a wide-character text, for fmx_build_suffix_array (tests).  Characters below 2^32.
int build_suffix_array_wide(const void *text_in, uint32_t cw, uint64_t n, uint64_t *sa_out, std::string &err) {
    if (cw != 2 && cw != 4 && cw != 8) {
        err = "character width must be 1, 2, 4 or 8 bytes";
        return FMX_ERR_INVALID_ARG;
    }
    if (n >= (1ull << 32) - 1) {
        err = "text length must be below 2^32 - 1";
        return FMX_ERR_UNSUPPORTED;
    }
    std::vector<uint32_t> t(n);
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t v = cw == 2 ? static_cast<const uint16_t *>(text_in)[i]
                                   : (cw == 4 ? static_cast<const uint32_t *>(text_in)[i] : static_cast<const uint64_t *>(text_in)[i]);
        if (v >= 0xFFFFFFFFull) {
            err = "characters must be below 2^32 - 1";
            return FMX_ERR_UNSUPPORTED;
        }
        t[i] = (uint32_t)v;
    }
    if (n > 1) {
        if (t[0] == 0) {
            err = "the given text must not start with zero character";
            return FMX_ERR_INVALID_TEXT;
        }
        if (!(t[n - 1] == 0 && t[n - 2] != 0)) {
            err = "the given text must end with exactly one zero character";
            return FMX_ERR_INVALID_TEXT;
        }
    }
    if (n == 1) sa_out[0] = 0;
    if (n <= 1) return 0;
    std::vector<uint32_t> present(t);
    std::sort(present.begin(), present.end());
    present.erase(std::unique(present.begin(), present.end()), present.end());
    std::vector<uint32_t> rk(n), sa(n);
    for (uint64_t i = 0; i < n; i++) rk[i] = (uint32_t)(std::lower_bound(present.begin(), present.end(), t[i]) - present.begin());
    suffix_array_ranks(rk.data(), n, (uint64_t)present.size(), sa.data());
    for (uint64_t i = 0; i < n; i++) sa_out[i] = sa[i];
    return 0;
}

int check_blob(const void *blob, uint64_t bytes, FmxBlobHeader &hdr, std::string &err) {
    if (!blob || bytes < sizeof(FmxBlobHeader)) {
        err = "blob too small";
        return FMX_ERR_INVALID_ARG;
    }
    std::memcpy(&hdr, blob, sizeof(hdr));
    if (hdr.magic != FMX_BLOB_MAGIC || hdr.version != FMX_BLOB_VERSION) {
        err = "not an fmx blob (bad magic or version)";
        return FMX_ERR_INVALID_ARG;
    }
    auto bad = [&](const char *what) {
        err = std::string("corrupt fmx blob: ") + what;
        return FMX_ERR_INVALID_ARG;
    };
    if (hdr.total_bytes != bytes) return bad("total size");
    const bool wide = hdr.layout == FMX_LAYOUT_WIDE;
    if (hdr.levels == 0 || hdr.levels > (wide ? FMX_MAX_WIDE_LEVELS : FMX_MAX_LEVELS) || hdr.kind > 2 || hdr.layout > FMX_LAYOUT_WIDE ||
        hdr.qlevels > FMX_MAX_QLEVELS || hdr.nexc > FMX_MAX_EXC)
        return bad("header");
    if (hdr.char_width != 0 && hdr.char_width != 1 && hdr.char_width != 2 && hdr.char_width != 4 && hdr.char_width != 8) return bad("character width");
    // the kernels copy cs_len + 1 words into shared tables of 256 / 257 entries (WIDE: read them from global memory)
    // and index rows with u32
    if (hdr.max_character == 0 || hdr.max_character == 0xFFFFFFFFu || (hdr.max_character > 255) != wide ||
        hdr.cs_len != hdr.max_character + 1)
        return bad("alphabet");
    if (wide && (hdr.char_width < 2 || hdr.verify || hdr.sec[SEC_VSA].bytes)) return bad("wide layout");
    if (hdr.levels != 64u - (uint32_t)__builtin_clzll((uint64_t)hdr.max_character)) return bad("levels");
    if (hdr.n >= 0xFFFFFFFFull || hdr.seq_len > hdr.n || hdr.runs > hdr.n || hdr.ndoc > hdr.n) return bad("lengths");
    if (hdr.sa_level >= 32 || hdr.isa_level >= 32 || hdr.vsa_level >= 32) return bad("sampling levels");
    if (hdr.kind == FMX_KIND_RLFM ? hdr.seq_len != hdr.runs : hdr.seq_len != hdr.n) return bad("sequence length");
    if (hdr.reserved[0] > FMX_MODE_RICH) return bad("mode");
    for (int k = 0; k < (int)SEC_COUNT; k++) {
        const uint64_t o = hdr.sec[k].offset, b = hdr.sec[k].bytes;
        if (b && (o % FMX_SECTION_ALIGN || o < sizeof(FmxBlobHeader) || b > bytes || o > bytes - b)) return bad("section table");
    }
    auto need = [&](int k, uint64_t min_bytes) { return hdr.sec[k].bytes >= min_bytes; };
    const uint64_t n = hdr.n, len = hdr.seq_len;
    if (!need(SEC_CS, (uint64_t)(hdr.cs_len + 1) * 4) || !need(SEC_ADJ, (uint64_t)hdr.cs_len * 4)) return bad("cs / adj");
    if (hdr.layout == FMX_LAYOUT_QUAT) {
        if (hdr.max_character > 4 || !need(SEC_LEVEL0, (len / 64 + 1) * 32) || (hdr.nexc && !need(SEC_EXC, (uint64_t)hdr.nexc * 4))) return bad("Q4 level");
    } else if (hdr.layout == FMX_LAYOUT_SYM) {
        if (hdr.sym_nblk != len / FMX_RB_BITS + 1 || !need(SEC_LEVEL0, (uint64_t)hdr.cs_len * hdr.sym_nblk * 32) || !need(SEC_LEVEL0 + 1, len))
            return bad("SYM vectors");
    } else if (hdr.layout == FMX_LAYOUT_WM4) {
        if (hdr.qlevels != (hdr.levels + 1) / 2) return bad("WM4 levels");
        for (uint32_t l = 0; l < hdr.qlevels; l++) {
            if (!need(SEC_LEVEL0 + l, (len / 64 + 1) * 32)) return bad("WM4 level");
            for (int d = 0; d < 4; d++)
                if (hdr.qoff[l][d] > len) return bad("WM4 offsets");
        }
    } else if (wide) {
        if (!need(SEC_LEVEL0, (uint64_t)hdr.levels * (len / FMX_RB_BITS + 1) * 32) || !need(SEC_WZEROS, (uint64_t)hdr.levels * 4))
            return bad("wide wavelet levels");
    } else {
        for (uint32_t l = 0; l < hdr.levels; l++)
            if (!need(SEC_LEVEL0 + l, (len / FMX_RB_BITS + 1) * 32) || hdr.zeros[l] > len) return bad("wavelet level");
    }
    if (hdr.has_locate || hdr.sec[SEC_SA].bytes) {
        if (n && (hdr.sa_count != ((n - 1) >> hdr.sa_level) + 1 || !need(SEC_SA, hdr.sa_count * 4))) return bad("suffix-array samples");
    }
    if (hdr.kind == FMX_KIND_MULTI && (!need(SEC_DOC, hdr.ndoc * 4) || !need(SEC_PIECE_END, hdr.ndoc * 4) || hdr.first_row >= (n ? n : 1)))
        return bad("pieces");
    if (hdr.kind == FMX_KIND_RLFM && (!need(SEC_RL_B, (n / FMX_RB_BITS + 1) * 32) || !need(SEC_RL_BP, (n / FMX_RB_BITS + 1) * 32) ||
                                      !need(SEC_RL_BSEL, (hdr.runs + 1) * 4) || !need(SEC_RL_BPSEL, (hdr.runs + 1) * 4)))
        return bad("run-length sections");
    if (hdr.verify) {
        if (!n || !need(SEC_TEXT, n) || !need(SEC_ISA, (((n - 1) >> hdr.isa_level) + 1) * 4)) return bad("verify sections");
        if (hdr.vsa_level == 0 && !need(SEC_VSA, n * 4)) return bad("verify suffix array");
        if (hdr.vsa_level != 0 && !hdr.sec[SEC_SA].bytes) return bad("verify samples");
    }
    if (hdr.sec[SEC_VSA].bytes && !need(SEC_VSA, n * 4)) return bad("dense suffix array");
    return 0;
}

}  // namespace fmx
