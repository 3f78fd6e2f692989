// kernels.cuh -- hand-written sm_100a kernels for the FM-index query path.
//
// Everything here is HBM-latency / sector-rate bound integer work: no tensor cores.
// One rank probe = one 32-byte sector (one 256-bit LDG) + one masked 64-bit popcount.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fmx_layout.h"

namespace fmx {

#define FMX_KIND_FM_ 0
#define FMX_KIND_RLFM_ 1
#define FMX_KIND_MULTI_ 2
#define FMX_NONE 0xFFFFFFFFFFFFFFFFull

// ------------------------------------------------------------------ rank blocks

struct RB {
    uint32_t w[8];  // w[0] = ones before the block, w[1] = in-block sub-counts, w[2..7] = 3 x 64 payload bits
};

// one 32-byte sector, one instruction (sm_100: LDG.E.256), read-only path
__device__ __forceinline__ RB rb_load(const uint4 *__restrict__ v, uint32_t blk) {
    RB b;
    const uint4 *p = v + 2ull * blk;
    // .L2::64B: on sm_100 a plain LDG makes L2 fill the whole 128-byte line from DRAM (125 B of DRAM
    // traffic per random 32 B read); this qualifier caps the fill at 64 B (63 B measured) --
    // tools/fetch_gran.cu, profiles/r01b_random_access_study.md
#ifndef FMX_NO_L2_64B
    asm("ld.global.nc.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
        : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]), "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]), "=r"(b.w[7])
        : "l"(p));
    return b;
}

// Scalar loads of randomly addressed index data (table entries, suffix-array / inverse entries, text words):
// the same .L2::64B qualifier, so a miss fills 64 bytes of the line from DRAM instead of all 128.
#ifndef FMX_NO_L2_64B
#define FMX_LD_S "ld.global.nc.L2::64B"
#else
#define FMX_LD_S "ld.global.nc"
#endif
__device__ __forceinline__ uint32_t ldg32_s(const uint32_t *p) {
    uint32_t v;
    asm(FMX_LD_S ".u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg64_s(const uint2 *p) {
    uint2 v;
    asm(FMX_LD_S ".v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg128_s(const uint4 *p) {
    uint4 v;
    asm volatile(FMX_LD_S ".v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg8_s(const uint8_t *p) {
    uint32_t v;
    asm(FMX_LD_S ".u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void rb_split(uint32_t pos, uint32_t &blk, uint32_t &r) {
    blk = pos / FMX_RB_BITS;
    r = pos - blk * FMX_RB_BITS;
}

// the 64-bit payload word holding in-block bit r
__device__ __forceinline__ void rb_word(const RB &b, uint32_t r, uint32_t &lo, uint32_t &hi) {
    uint32_t j = r >> 6;
    lo = j == 0 ? b.w[2] : (j == 1 ? b.w[4] : b.w[6]);
    hi = j == 0 ? b.w[3] : (j == 1 ? b.w[5] : b.w[7]);
}

// ones in [0, blk*192 + r), r in [0, 192): block count + stored sub-count + one masked 64-bit popcount
__device__ __forceinline__ uint32_t rb_rank(const RB &b, uint32_t r) {
    uint32_t lo, hi;
    rb_word(b, r, lo, hi);
    uint32_t sub = __byte_perm(b.w[1], 0, 0x4440u | (r >> 6));
    uint64_t keep = ~(~0ull << (r & 63u));  // low (r mod 64) bits
    return b.w[0] + sub + __popc(lo & (uint32_t)keep) + __popc(hi & (uint32_t)(keep >> 32));
}

__device__ __forceinline__ uint32_t rb_bit(const RB &b, uint32_t r) {
    uint32_t lo, hi;
    rb_word(b, r, lo, hi);
    uint32_t x = r & 63u;
    return ((x & 32u ? hi : lo) >> (x & 31u)) & 1u;
}

// rank and bit together (access + rank share the selected word)
__device__ __forceinline__ uint32_t rb_rank_bit(const RB &b, uint32_t r, uint32_t &bit) {
    uint32_t lo, hi;
    rb_word(b, r, lo, hi);
    uint32_t x = r & 63u;
    uint32_t sub = __byte_perm(b.w[1], 0, 0x4440u | (r >> 6));
    uint64_t keep = ~(~0ull << x);
    bit = ((x & 32u ? hi : lo) >> (x & 31u)) & 1u;
    return b.w[0] + sub + __popc(lo & (uint32_t)keep) + __popc(hi & (uint32_t)(keep >> 32));
}

// rank1 of a stand-alone RB192 vector
__device__ __forceinline__ uint32_t rbv_rank1(const uint4 *__restrict__ v, uint32_t pos) {
    uint32_t blk, r;
    rb_split(pos, blk, r);
    RB b = rb_load(v, blk);
    return rb_rank(b, r);
}

// position of the k-th (0-based) set bit of a 64-bit word given as (lo, hi)
__device__ __forceinline__ uint32_t select_in_word64(uint32_t lo, uint32_t hi, uint32_t k) {
    uint32_t c = __popc(lo);
    return k < c ? __fns(lo, 0, k + 1) : 32u + __fns(hi, 0, k - c + 1);
}

// select1: position of the k-th (0-based) one of an RB192 vector with nblk blocks.  k must exist.
__device__ __forceinline__ uint32_t rbv_select1(const uint4 *__restrict__ v, uint32_t nblk, uint32_t k) {
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(v);
    uint32_t lo = 0, hi = nblk;  // last block with count <= k
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(cw + 8ull * mid) <= k) lo = mid; else hi = mid;
    }
    RB b = rb_load(v, lo);
    uint32_t rem = k - b.w[0];
    uint32_t s1 = (b.w[1] >> 8) & 0xFFu, s2 = (b.w[1] >> 16) & 0xFFu;
    uint32_t j = rem >= s2 ? 2u : (rem >= s1 ? 1u : 0u);
    rem -= j == 2 ? s2 : (j == 1 ? s1 : 0u);
    uint32_t wl, wh;
    rb_word(b, j << 6, wl, wh);
    return lo * FMX_RB_BITS + (j << 6) + select_in_word64(wl, wh, rem);
}
__device__ __forceinline__ uint32_t rbv_select0(const uint4 *__restrict__ v, uint32_t nblk, uint32_t k) {
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(v);
    uint32_t lo = 0, hi = nblk;  // last block with zeros-before <= k
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (mid * FMX_RB_BITS - __ldg(cw + 8ull * mid) <= k) lo = mid; else hi = mid;
    }
    RB b = rb_load(v, lo);
    uint32_t rem = k - (lo * FMX_RB_BITS - b.w[0]);
    uint32_t z1 = 64u - ((b.w[1] >> 8) & 0xFFu), z2 = 128u - ((b.w[1] >> 16) & 0xFFu);
    uint32_t j = rem >= z2 ? 2u : (rem >= z1 ? 1u : 0u);
    rem -= j == 2 ? z2 : (j == 1 ? z1 : 0u);
    uint32_t wl, wh;
    rb_word(b, j << 6, wl, wh);
    return lo * FMX_RB_BITS + (j << 6) + select_in_word64(~wl, ~wh, rem);
}

// ------------------------------------------------------------------ shared-memory tables

#define FMX_LAYOUT_WM 0  // binary wavelet matrix, L levels (any alphabet)
#define FMX_LAYOUT_Q4 1  // one quaternary level (max_character <= 4, few \0): ONE sector per lf_map2
#define FMX_LAYOUT_W4 2  // quaternary wavelet matrix, ceil(L/2) levels (alphabets too large for SYM's budget)
#define FMX_LAYOUT_SY 3  // one RB192 bit vector per symbol + the raw sequence: ONE sector per lf_map2

#define FMX_LAYOUT_WW 4  // WIDE: binary wavelet matrix of up to 32 levels, cs / adj in global memory (max_character > 255)

template <int LAYOUT>
struct Tabs {
    uint32_t adj[256];  // WM: cs[c] - walk_c(0)
    uint32_t cs[257];   // cs[c], cs[cs_len] = sequence length
    uint32_t exc[LAYOUT == FMX_LAYOUT_Q4 ? FMX_MAX_EXC : 1];  // Q4: sorted positions whose symbol is 0
};
// wide alphabets: max_character + 1 entries each, far beyond shared memory -- the tables stay where they are
template <>
struct Tabs<FMX_LAYOUT_WW> {
    const uint32_t *adj;
    const uint32_t *cs;
    uint32_t exc[1];
};

template <int LAYOUT>
__device__ __forceinline__ void load_tables(const FmxDev &ix, Tabs<LAYOUT> &t) {
    for (uint32_t k = threadIdx.x; k < ix.cs_len; k += blockDim.x) t.adj[k] = __ldg(ix.adj + k);
    for (uint32_t k = threadIdx.x; k <= ix.cs_len; k += blockDim.x) t.cs[k] = __ldg(ix.cs + k);
    if (LAYOUT == FMX_LAYOUT_Q4)
        for (uint32_t k = threadIdx.x; k < ix.nexc; k += blockDim.x) t.exc[k] = __ldg(ix.exc + k);
    __syncthreads();
}
template <>
__device__ __forceinline__ void load_tables<FMX_LAYOUT_WW>(const FmxDev &ix, Tabs<FMX_LAYOUT_WW> &t) {
    if (threadIdx.x == 0) {
        t.adj = ix.adj;
        t.cs = ix.cs;
    }
    __syncthreads();
}

// level l of the binary wavelet matrix and its zero count: per-level sections (WM) or one section (WIDE)
template <int LAYOUT>
__device__ __forceinline__ const uint4 *wm_lv(const FmxDev &ix, uint32_t l) {
    return LAYOUT == FMX_LAYOUT_WW ? ix.lv[0] + 2ull * l * ix.wide_nblk : ix.lv[l];  // a block is two uint4
}
template <int LAYOUT>
__device__ __forceinline__ uint32_t wm_zeros(const FmxDev &ix, uint32_t l) {
    return LAYOUT == FMX_LAYOUT_WW ? __ldg(ix.wzeros + l) : ix.zeros[l];
}
// WIDE: symbols that do not occur have rank 0 everywhere and no adj entry (cs[c + 1] == cs[c] marks them)
template <int LAYOUT>
__device__ __forceinline__ bool sym_absent(const Tabs<LAYOUT> &t, uint32_t c) {
    return LAYOUT == FMX_LAYOUT_WW && t.cs[c + 1] == t.cs[c];
}

// ------------------------------------------------------------------ wavelet matrix (LAYOUT_WM)

// position of `pos` after walking symbol c's path down all levels (rank(i,c) = walk - walk_c(0))
template <int LAYOUT>
__device__ __forceinline__ uint32_t wm_walk(const FmxDev &ix, uint32_t c, uint32_t pos) {
    const uint32_t L = ix.levels;
#pragma unroll 1
    for (uint32_t l = 0; l < L; l++) {
        uint32_t blk, r;
        rb_split(pos, blk, r);
        RB b = rb_load(wm_lv<LAYOUT>(ix, l), blk);
        uint32_t ones = rb_rank(b, r);
        pos = ((c >> (L - 1 - l)) & 1u) ? wm_zeros<LAYOUT>(ix, l) + ones : pos - ones;
    }
    return pos;
}

// the same for the two ends of an SA range; shares the sector when both fall in one block
template <int LAYOUT>
__device__ __forceinline__ void wm_walk2(const FmxDev &ix, uint32_t c, uint32_t &s, uint32_t &e) {
    const uint32_t L = ix.levels;
#pragma unroll 1
    for (uint32_t l = 0; l < L; l++) {
        uint32_t bs, rs, be, re;
        rb_split(s, bs, rs);
        rb_split(e, be, re);
        const uint4 *v = wm_lv<LAYOUT>(ix, l);
        RB a = rb_load(v, bs);
        RB b = a;
        if (be != bs) b = rb_load(v, be);
        uint32_t os = rb_rank(a, rs), oe = rb_rank(b, re);
        if ((c >> (L - 1 - l)) & 1u) {
            uint32_t z = wm_zeros<LAYOUT>(ix, l);
            s = z + os;
            e = z + oe;
        } else {
            s -= os;
            e -= oe;
        }
    }
}

// access + walk fused: symbol at `pos` and its position at the bottom level (one sector / level;
// the reference does get_l then rank separately, fm_index.rs:87-89)
template <int LAYOUT>
__device__ __forceinline__ uint32_t wm_access_walk(const FmxDev &ix, uint32_t pos, uint32_t &sym) {
    const uint32_t L = ix.levels;
    uint32_t c = 0;
#pragma unroll 1
    for (uint32_t l = 0; l < L; l++) {
        uint32_t blk, r;
        rb_split(pos, blk, r);
        RB b = rb_load(wm_lv<LAYOUT>(ix, l), blk);
        uint32_t bit;
        uint32_t ones = rb_rank_bit(b, r, bit);
        c = (c << 1) | bit;
        pos = bit ? wm_zeros<LAYOUT>(ix, l) + ones : pos - ones;
    }
    sym = c;
    return pos;
}

// select_u64_unchecked(k, c): position of the k-th c.  base = walk_c(0).
template <int LAYOUT>
__device__ __forceinline__ uint32_t wm_select(const FmxDev &ix, uint32_t c, uint32_t k, uint32_t base) {
    const uint32_t L = ix.levels;
    const uint32_t nblk = ix.seq_len / FMX_RB_BITS + 1;
    uint32_t pos = base + k;
#pragma unroll 1
    for (uint32_t l = L; l-- > 0;) {
        if ((c >> (L - 1 - l)) & 1u) pos = rbv_select1(wm_lv<LAYOUT>(ix, l), nblk, pos - wm_zeros<LAYOUT>(ix, l));
        else pos = rbv_select0(wm_lv<LAYOUT>(ix, l), nblk, pos);
    }
    return pos;
}

// ------------------------------------------------------------------ quaternary level (LAYOUT_Q4)
// block = u32 cnt[4] (occurrences of codes 0..3 before the block) + 64 two-bit codes.
// code = symbol - 1; the rare symbol 0 (the \0 terminators) is stored as code 0 and listed in `exc`.

__device__ __forceinline__ uint32_t q4_lower_bound(const uint32_t *exc, uint32_t n, uint32_t i) {
    uint32_t lo = 0, hi = n;  // number of exceptions < i
    while (lo < hi) {
        uint32_t m = lo + ((hi - lo) >> 1);
        if (exc[m] < i) lo = m + 1; else hi = m;
    }
    return lo;
}
__device__ __forceinline__ uint32_t q4_exc_before(const FmxDev &ix, const uint32_t *exc, uint32_t i) {
    if (ix.nexc == 1) return exc[0] < i ? 1u : 0u;
    return q4_lower_bound(exc, ix.nexc, i);
}
__device__ __forceinline__ bool q4_is_exc(const FmxDev &ix, const uint32_t *exc, uint32_t i) {
    if (ix.nexc == 1) return exc[0] == i;
    uint32_t lb = q4_lower_bound(exc, ix.nexc, i);
    return lb < ix.nexc && exc[lb] == i;
}
// the block's payload is two 64-bit planes (w[4..5] = low bits, w[6..7] = high bits of the 64 codes):
// bit t of (lo, hi) is set where code t equals `code`
__device__ __forceinline__ void q4_match(const RB &b, uint32_t code, uint32_t &lo, uint32_t &hi) {
    const uint32_t a0 = (code & 1u) ? 0u : 0xFFFFFFFFu, a1 = (code & 2u) ? 0u : 0xFFFFFFFFu;
    lo = (b.w[4] ^ a0) & (b.w[6] ^ a1);
    hi = (b.w[5] ^ a0) & (b.w[7] ^ a1);
}
// occurrences of `code` among the first r codes of the block, r in [0, 64)
__device__ __forceinline__ uint32_t q4_count(const RB &b, uint32_t r, uint32_t code) {
    uint32_t lo, hi;
    q4_match(b, code, lo, hi);
    const uint64_t keep = ~(~0ull << r);
    return __popc(lo & (uint32_t)keep) + __popc(hi & (uint32_t)(keep >> 32));
}
__device__ __forceinline__ uint32_t q4_code(const RB &b, uint32_t r) {
    const uint32_t p0 = (r & 32u) ? b.w[5] : b.w[4], p1 = (r & 32u) ? b.w[7] : b.w[6];
    return ((p0 >> (r & 31u)) & 1u) | (((p1 >> (r & 31u)) & 1u) << 1);
}
__device__ __forceinline__ uint32_t q4_cnt(const RB &b, uint32_t code) {
    const uint32_t lo = (code & 1u) ? b.w[1] : b.w[0], hi = (code & 1u) ? b.w[3] : b.w[2];  // selects, no branches
    return (code & 2u) ? hi : lo;
}
// exceptions before i, counted only when `on` (lane-uniform control flow: the common single-\0 case is
// two predicated instructions instead of a divergent branch)
__device__ __forceinline__ uint32_t q4_exc_before_if(const FmxDev &ix, const uint32_t *exc, uint32_t i, bool on) {
    if (ix.nexc == 1) return (on && exc[0] < i) ? 1u : 0u;
    return on ? q4_lower_bound(exc, ix.nexc, i) : 0u;
}
// rank(i, c) given the block of i
__device__ __forceinline__ uint32_t q4_rank_in(const FmxDev &ix, const uint32_t *exc, const RB &b, uint32_t i,
                                               uint32_t c) {
    if (c == 0) return q4_exc_before(ix, exc, i);
    uint32_t code = c - 1u;
    uint32_t v = q4_cnt(b, code) + q4_count(b, i & 63u, code);
    return v - q4_exc_before_if(ix, exc, i, code == 0);
}
// the same for both ends of an SA range; bs / be = their blocks, b2 = the block of e when it differs.
// When both ends share the block (the usual case once the range is narrow) the match mask and the
// block counter are computed once.
__device__ __forceinline__ void q4_rank2_in(const FmxDev &ix, const uint32_t *exc, const uint4 *v, uint32_t c, uint32_t &s,
                                            uint32_t &e) {
    const uint32_t bs = s >> 6, be = e >> 6;
    RB a = rb_load(v, bs);
    if (c == 0) {
        s = q4_exc_before(ix, exc, s);
        e = q4_exc_before(ix, exc, e);
        return;
    }
    const uint32_t code = c - 1u;
    const uint32_t xs = q4_exc_before_if(ix, exc, s, code == 0), xe = q4_exc_before_if(ix, exc, e, code == 0);
    uint32_t ml, mh;
    q4_match(a, code, ml, mh);
    const uint32_t base = q4_cnt(a, code);
    const uint64_t ks = ~(~0ull << (s & 63u));
    s = base + __popc(ml & (uint32_t)ks) + __popc(mh & (uint32_t)(ks >> 32)) - xs;
    if (be == bs) {
        const uint64_t ke = ~(~0ull << (e & 63u));
        e = base + __popc(ml & (uint32_t)ke) + __popc(mh & (uint32_t)(ke >> 32)) - xe;
    } else {
        RB b = rb_load(v, be);
        e = q4_cnt(b, code) + q4_count(b, e & 63u, code) - xe;
    }
}
// select(k, c)
__device__ __forceinline__ uint32_t q4_select(const FmxDev &ix, const uint32_t *exc, uint32_t c, uint32_t k) {
    if (c == 0) return exc[k];
    const uint32_t code = c - 1u;
    const uint32_t nblk = (ix.seq_len >> 6) + 1u;
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(ix.lv[0]);
    uint32_t lo = 0, hi = nblk;  // last block with (occurrences of c before it) <= k
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        uint32_t before = __ldg(cw + 8ull * mid + code);
        if (code == 0) before -= q4_exc_before(ix, exc, mid << 6);
        if (before <= k) lo = mid; else hi = mid;
    }
    RB b = rb_load(ix.lv[0], lo);
    uint32_t before = q4_cnt(b, code);
    if (code == 0) before -= q4_exc_before(ix, exc, lo << 6);
    uint32_t rem = k - before, pos = lo << 6;
    uint32_t ml, mh;
    q4_match(b, code, ml, mh);
    if (code == 0) {  // positions holding \0 are stored as code 0: drop them
        uint32_t x0 = q4_exc_before(ix, exc, pos), x1 = q4_exc_before(ix, exc, pos + 64u);
        for (uint32_t x = x0; x < x1; x++) {
            uint32_t t = exc[x] - pos;
            if (t & 32u) mh &= ~(1u << (t & 31u)); else ml &= ~(1u << t);
        }
    }
    return pos + select_in_word64(ml, mh, rem);
}

// ------------------------------------------------------------------ quaternary wavelet matrix (LAYOUT_W4)
// Q4-style blocks (u32 cnt[4] + 64 two-bit codes), one level per two bits of the symbol, most
// significant digit first; the codes are plain digits (no exception list).  Symbols are stably
// partitioned by digit: group d of level l starts at qoff[l][d].

__device__ __forceinline__ uint32_t w4_digit(const FmxDev &ix, uint32_t c, uint32_t l) {
    return (c >> (2u * (ix.qlevels - 1u - l))) & 3u;
}
// position after descending one level along digit d, given the block of pos
__device__ __forceinline__ uint32_t w4_down(const FmxDev &ix, const RB &b, uint32_t l, uint32_t d, uint32_t pos) {
    return ix.qoff[l * 4u + d] + q4_cnt(b, d) + q4_count(b, pos & 63u, d);
}

__device__ __forceinline__ uint32_t w4_walk(const FmxDev &ix, uint32_t c, uint32_t pos) {
    const uint32_t Lq = ix.qlevels;
#pragma unroll 1
    for (uint32_t l = 0; l < Lq; l++) {
        RB b = rb_load(ix.lv[l], pos >> 6);
        pos = w4_down(ix, b, l, w4_digit(ix, c, l), pos);
    }
    return pos;
}

__device__ __forceinline__ void w4_walk2(const FmxDev &ix, uint32_t c, uint32_t &s, uint32_t &e) {
    const uint32_t Lq = ix.qlevels;
#pragma unroll 1
    for (uint32_t l = 0; l < Lq; l++) {
        const uint4 *v = ix.lv[l];
        uint32_t bs = s >> 6, be = e >> 6;
        RB a = rb_load(v, bs);
        RB b = a;
        if (be != bs) b = rb_load(v, be);
        uint32_t d = w4_digit(ix, c, l);
        s = w4_down(ix, a, l, d, s);
        e = w4_down(ix, b, l, d, e);
    }
}

__device__ __forceinline__ uint32_t w4_access_walk(const FmxDev &ix, uint32_t pos, uint32_t &sym) {
    const uint32_t Lq = ix.qlevels;
    uint32_t c = 0;
#pragma unroll 1
    for (uint32_t l = 0; l < Lq; l++) {
        RB b = rb_load(ix.lv[l], pos >> 6);
        uint32_t d = q4_code(b, pos & 63u);
        c = (c << 2) | d;
        pos = w4_down(ix, b, l, d, pos);
    }
    sym = c;
    return pos;
}

// position of the k-th (0-based) digit d in level l
__device__ __forceinline__ uint32_t w4_select_level(const FmxDev &ix, uint32_t l, uint32_t d, uint32_t k) {
    const uint32_t nblk = (ix.seq_len >> 6) + 1u;
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(ix.lv[l]);
    uint32_t lo = 0, hi = nblk;  // last block with (digits d before it) <= k
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(cw + 8ull * mid + d) <= k) lo = mid; else hi = mid;
    }
    RB b = rb_load(ix.lv[l], lo);
    uint32_t rem = k - q4_cnt(b, d);
    uint32_t ml, mh;
    q4_match(b, d, ml, mh);
    return (lo << 6) + select_in_word64(ml, mh, rem);
}

// select(k, c); base = walk_c(0)
__device__ __forceinline__ uint32_t w4_select(const FmxDev &ix, uint32_t c, uint32_t k, uint32_t base) {
    uint32_t pos = base + k;
#pragma unroll 1
    for (uint32_t l = ix.qlevels; l-- > 0;) {
        uint32_t d = w4_digit(ix, c, l);
        pos = w4_select_level(ix, l, d, pos - ix.qoff[l * 4u + d]);
    }
    return pos;
}

// ------------------------------------------------------------------ per-symbol bit vectors (LAYOUT_SY)
// vector c = RB192 over [seq[i] == c], starting at block c * sym_nblk of lv[0]; raw = the sequence bytes

__device__ __forceinline__ const uint4 *sy_vec(const FmxDev &ix, uint32_t c) {
    return ix.lv[0] + 2ull * (uint64_t)c * ix.sym_nblk;
}

// ------------------------------------------------------------------ sequence primitives, all layouts
// The "sequence" is the BWT (FM, MultiPieces) or the run heads (RLFM); cs is the matching C array.

// cs[c] + rank(i, c)
template <int LAYOUT>
__device__ __forceinline__ uint32_t seq_lf(const FmxDev &ix, const Tabs<LAYOUT> &t, uint32_t c, uint32_t i) {
    if (LAYOUT == FMX_LAYOUT_Q4) {
        RB b = rb_load(ix.lv[0], i >> 6);
        return t.cs[c] + q4_rank_in(ix, t.exc, b, i, c);
    } else if (LAYOUT == FMX_LAYOUT_SY) {
        return t.cs[c] + rbv_rank1(sy_vec(ix, c), i);
    } else if (LAYOUT == FMX_LAYOUT_W4) {
        return t.adj[c] + w4_walk(ix, c, i);
    } else {
        if (sym_absent<LAYOUT>(t, c)) return t.cs[c];
        return t.adj[c] + wm_walk<LAYOUT>(ix, c, i);
    }
}

// (s, e) <- (cs[c] + rank(s, c), cs[c] + rank(e, c)); one sector when both ends share a block
template <int LAYOUT>
__device__ __forceinline__ void seq_lf2(const FmxDev &ix, const Tabs<LAYOUT> &t, uint32_t c, uint32_t &s, uint32_t &e) {
    if (LAYOUT == FMX_LAYOUT_Q4) {
        q4_rank2_in(ix, t.exc, ix.lv[0], c, s, e);
        const uint32_t base = t.cs[c];
        s += base;
        e += base;
    } else if (LAYOUT == FMX_LAYOUT_SY) {
        const uint4 *v = sy_vec(ix, c);
        uint32_t bs, rs, be, re;
        rb_split(s, bs, rs);
        rb_split(e, be, re);
        RB a = rb_load(v, bs);
        RB b = a;
        if (be != bs) b = rb_load(v, be);
        uint32_t base = t.cs[c];
        s = base + rb_rank(a, rs);
        e = base + rb_rank(b, re);
    } else {
        if (sym_absent<LAYOUT>(t, c)) {
            s = e = t.cs[c];
            return;
        }
        if (LAYOUT == FMX_LAYOUT_W4) w4_walk2(ix, c, s, e);
        else wm_walk2<LAYOUT>(ix, c, s, e);
        uint32_t a = t.adj[c];
        s += a;
        e += a;
    }
}

// sym = seq[i]; returns cs[sym] + rank(i, sym)   (access + rank fused: same sector(s))
template <int LAYOUT>
__device__ __forceinline__ uint32_t seq_access_lf(const FmxDev &ix, const Tabs<LAYOUT> &t, uint32_t i, uint32_t &sym) {
    if (LAYOUT == FMX_LAYOUT_Q4) {
        RB b = rb_load(ix.lv[0], i >> 6);
        uint32_t c = q4_is_exc(ix, t.exc, i) ? 0u : q4_code(b, i & 63u) + 1u;
        sym = c;
        return t.cs[c] + q4_rank_in(ix, t.exc, b, i, c);
    } else if (LAYOUT == FMX_LAYOUT_SY) {
        uint32_t c = ldg8_s(ix.raw + i);
        sym = c;
        return t.cs[c] + rbv_rank1(sy_vec(ix, c), i);
    } else {
        uint32_t c;
        uint32_t w = LAYOUT == FMX_LAYOUT_W4 ? w4_access_walk(ix, i, c) : wm_access_walk<LAYOUT>(ix, i, c);
        sym = c;
        return w + t.adj[c];
    }
}

template <int LAYOUT>
__device__ __forceinline__ uint32_t seq_access(const FmxDev &ix, const Tabs<LAYOUT> &t, uint32_t i) {
    if (LAYOUT == FMX_LAYOUT_Q4) {
        if (q4_is_exc(ix, t.exc, i)) return 0u;
        RB b = rb_load(ix.lv[0], i >> 6);
        return q4_code(b, i & 63u) + 1u;
    } else if (LAYOUT == FMX_LAYOUT_SY) {
        return ldg8_s(ix.raw + i);
    } else {
        uint32_t c;
        if (LAYOUT == FMX_LAYOUT_W4) w4_access_walk(ix, i, c);
        else wm_access_walk<LAYOUT>(ix, i, c);
        return c;
    }
}

// select(k, c): position of the k-th (0-based) c
template <int LAYOUT>
__device__ __forceinline__ uint32_t seq_select(const FmxDev &ix, const Tabs<LAYOUT> &t, uint32_t c, uint32_t k) {
    if (LAYOUT == FMX_LAYOUT_Q4) return q4_select(ix, t.exc, c, k);
    if (LAYOUT == FMX_LAYOUT_SY) return rbv_select1(sy_vec(ix, c), ix.sym_nblk, k);
    if (LAYOUT == FMX_LAYOUT_W4) return w4_select(ix, c, k, t.cs[c] - t.adj[c]);
    return wm_select<LAYOUT>(ix, c, k, t.cs[c] - t.adj[c]);
}

// ------------------------------------------------------------------ backend primitives
// (crate-private seam src/backend.rs:5-40)

// MultiPieces rule for c == 0 (multi_pieces.rs:142-148); rank = rank(bw, i, 0)
__device__ __forceinline__ uint32_t multi_zero_rule(const FmxDev &ix, uint32_t i, uint32_t rank) {
    return i < ix.first_row ? rank + 1u : (i == ix.first_row ? 0u : rank);
}

// ---- RLFM pieces (rlfmi.rs:118-170)
// for row i and symbol c:
//   j   = rank1(b, i)
//   h   = index of the run containing row i  (= rank1(b, i+1) - 1, clamped at i == n)
//   nrc = cs[c] + rank(s, j, c)
//   hit = (s[h] == c)
// h is j or j-1, so the access of s[h] shares its sector(s) with the rank of j.
template <int LAYOUT>
__device__ __forceinline__ void rl_probe(const FmxDev &ix, const Tabs<LAYOUT> &t, uint32_t c, uint32_t i, uint32_t &j,
                                         uint32_t &nrc, bool &hit) {
    uint32_t blk, r;
    rb_split(i, blk, r);
    RB bb = rb_load(ix.rl_b, blk);
    j = rb_rank(bb, r);
    uint32_t starts_here = (i < ix.n) ? rb_bit(bb, r) : 0u;
    uint32_t h = starts_here ? j : j - 1u;  // b[0] == 1 so j >= 1 whenever !starts_here
    if (LAYOUT == FMX_LAYOUT_Q4) {
        uint32_t bj = j >> 6, bh = h >> 6;
        RB a = rb_load(ix.lv[0], bj);
        RB d = a;
        if (bh != bj) d = rb_load(ix.lv[0], bh);
        uint32_t sh = q4_is_exc(ix, t.exc, h) ? 0u : q4_code(d, h & 63u) + 1u;
        hit = sh == c;
        nrc = t.cs[c] + q4_rank_in(ix, t.exc, a, j, c);
    } else if (LAYOUT == FMX_LAYOUT_SY) {
        hit = ldg8_s(ix.raw + h) == c;
        nrc = t.cs[c] + rbv_rank1(sy_vec(ix, c), j);
    } else if (LAYOUT == FMX_LAYOUT_W4) {
        const uint32_t Lq = ix.qlevels;
        uint32_t p = j, q = h;
        bool alive = true;
#pragma unroll 1
        for (uint32_t l = 0; l < Lq; l++) {
            const uint4 *v = ix.lv[l];
            uint32_t d = w4_digit(ix, c, l);
            uint32_t bp_ = p >> 6, bq = q >> 6;
            RB a = rb_load(v, bp_);
            if (alive) {
                RB g = a;
                if (bq != bp_) g = rb_load(v, bq);
                alive = q4_code(g, q & 63u) == d;
                q = w4_down(ix, g, l, d, q);
            }
            p = w4_down(ix, a, l, d, p);
        }
        nrc = t.adj[c] + p;
        hit = alive;
    } else {
        if (sym_absent<LAYOUT>(t, c)) {  // no run of c: rank 0, and the run of row i has another head
            nrc = t.cs[c];
            hit = false;
            return;
        }
        const uint32_t L = ix.levels;
        uint32_t p = j, q = h;
        bool alive = true;
#pragma unroll 1
        for (uint32_t l = 0; l < L; l++) {
            uint32_t bit = (c >> (L - 1 - l)) & 1u;
            uint32_t bp_, rp, bq, rq;
            rb_split(p, bp_, rp);
            rb_split(q, bq, rq);
            const uint4 *v = wm_lv<LAYOUT>(ix, l);
            RB a = rb_load(v, bp_);
            uint32_t op = rb_rank(a, rp);
            if (alive) {
                RB d = a;
                if (bq != bp_) d = rb_load(v, bq);
                uint32_t qbit;
                uint32_t oq = rb_rank_bit(d, rq, qbit);
                alive = qbit == bit;
                q = bit ? wm_zeros<LAYOUT>(ix, l) + oq : q - oq;
            }
            p = bit ? wm_zeros<LAYOUT>(ix, l) + op : p - op;
        }
        nrc = t.adj[c] + p;
        hit = alive;
    }
}

// lf_map2 for every kind (fm_index.rs:93-95, multi_pieces.rs:140-153, rlfmi.rs:135-143)
template <int KIND, int LAYOUT>
__device__ __forceinline__ uint32_t lf_map2_dev(const FmxDev &ix, const Tabs<LAYOUT> &t, uint32_t c, uint32_t i) {
    if (KIND == FMX_KIND_RLFM_) {
        uint32_t j, nrc;
        bool hit;
        rl_probe<LAYOUT>(ix, t, c, i, j, nrc, hit);
        uint32_t v = ldg32_s(ix.rl_bpsel + nrc);
        if (!hit) return v;
        return v + i - ldg32_s(ix.rl_bsel + j);
    } else {
        uint32_t w = seq_lf<LAYOUT>(ix, t, c, i);
        if (KIND == FMX_KIND_MULTI_ && c == 0) return multi_zero_rule(ix, i, w);
        return w;
    }
}

// (s, e) <- (lf_map2(c, s), lf_map2(c, e))   (wrapper.rs:109-110)
template <int KIND, int LAYOUT>
__device__ __forceinline__ void lf_map2_pair(const FmxDev &ix, const Tabs<LAYOUT> &t, uint32_t c, uint32_t &s,
                                             uint32_t &e) {
    if (KIND == FMX_KIND_RLFM_) {
        s = lf_map2_dev<KIND, LAYOUT>(ix, t, c, s);
        e = lf_map2_dev<KIND, LAYOUT>(ix, t, c, e);
    } else {
        uint32_t s0 = s, e0 = e;
        seq_lf2<LAYOUT>(ix, t, c, s, e);
        if (KIND == FMX_KIND_MULTI_ && c == 0) {
            s = multi_zero_rule(ix, s0, s);
            e = multi_zero_rule(ix, e0, e);
        }
    }
}

// get_l + lf_map fused: returns lf_map(i), sym = get_l(i)
template <int KIND, int LAYOUT>
__device__ __forceinline__ uint32_t lf_step(const FmxDev &ix, const Tabs<LAYOUT> &t, uint32_t i, uint32_t &sym) {
    if (KIND == FMX_KIND_RLFM_) {
        // rlfmi.rs:127-133: c = get_l(i); j = rank1(b,i); nr = rank(s,j,c); select1(bp, cs[c]+nr) + i - select1(b, j)
        uint32_t blk, r;
        rb_split(i, blk, r);
        RB bb = rb_load(ix.rl_b, blk);
        uint32_t bit;
        uint32_t j = rb_rank_bit(bb, r, bit);
        uint32_t h = bit ? j : j - 1u;
        uint32_t c;
        uint32_t q = seq_access_lf<LAYOUT>(ix, t, h, c);  // cs[c] + rank(s, h, c); rank(s, j, c) adds (j > h)
        sym = c;
        return ldg32_s(ix.rl_bpsel + (q + (j - h))) + i - ldg32_s(ix.rl_bsel + j);
    } else {
        uint32_t c;
        uint32_t w = seq_access_lf<LAYOUT>(ix, t, i, c);
        sym = c;
        if (KIND == FMX_KIND_MULTI_ && c == 0) return multi_zero_rule(ix, i, w);
        return w;
    }
}

template <int KIND, int LAYOUT>
__device__ __forceinline__ uint32_t get_l_dev(const FmxDev &ix, const Tabs<LAYOUT> &t, uint32_t i) {
    if (KIND == FMX_KIND_RLFM_) {
        uint32_t h = rbv_rank1(ix.rl_b, i + 1 > ix.n ? ix.n : i + 1) - 1u;  // rank1 clamps past the end
        return seq_access<LAYOUT>(ix, t, h);
    }
    return seq_access<LAYOUT>(ix, t, i);
}

// greatest c with cs[c] <= key  (fm_index.rs:97-112)
__device__ __forceinline__ uint32_t cs_search(const uint32_t *s_cs, uint32_t cs_len, uint32_t key) {
    uint32_t s = 0, e = cs_len;
    while (e - s > 1) {
        uint32_t m = s + ((e - s) >> 1);
        if (s_cs[m] <= key) s = m; else e = m;
    }
    return s;
}

// get_f + fl_map fused (fm_index.rs:97-120, multi_pieces.rs:155-181, rlfmi.rs:145-169).
// returns false when fl_map is None (MultiPieces, F[i] == 0).
template <int KIND, int LAYOUT>
__device__ __forceinline__ bool fl_step(const FmxDev &ix, const Tabs<LAYOUT> &t, uint32_t i, uint32_t &sym,
                                        uint32_t &next) {
    if (KIND == FMX_KIND_RLFM_) {
        uint32_t jr = rbv_rank1(ix.rl_bp, i + 1) - 1u;
        uint32_t c = cs_search(t.cs, ix.cs_len, jr);
        uint32_t p = ldg32_s(ix.rl_bpsel + jr);
        uint32_t m = seq_select<LAYOUT>(ix, t, c, jr - t.cs[c]);
        sym = c;
        next = ldg32_s(ix.rl_bsel + m) + i - p;
        return true;
    } else {
        uint32_t c = cs_search(t.cs, ix.cs_len, i);
        sym = c;
        if (KIND == FMX_KIND_MULTI_ && c == 0) return false;
        next = seq_select<LAYOUT>(ix, t, c, i - t.cs[c]);
        return true;
    }
}

// ------------------------------------------------------------------ kernels

struct SearchArgs {
    const uint8_t *pat;
    const uint64_t *pat_off;  // NULL => fixed_len
    uint64_t fixed_len;
    uint64_t npat;
    const uint64_t *init_s;  // NULL => (s0, e0)
    const uint64_t *init_e;
    uint32_t s0, e0;
    uint64_t *out_s;
    uint64_t *out_e;
    uint32_t *err;                  // set to 1 if a processed char > max_character
    unsigned long long *work;       // [0] += executed search iterations
    // memoised top of the search: (s, e) after the reference loop has consumed the LAST kmer_k
    // characters of a pattern starting from (0, n) -- including its early break -- for every
    // k-mer over 0..=max_character.  NULL => not used.
    // Table indices are base-max_character numbers over the digits c-1 (patterns with a \0 or an
    // out-of-range character among those k take the ordinary path).  Two tables: a small L2-resident
    // one (kmer_k) and, HBM permitting, a large one (big_k > kmer_k) that replaces big_k iterations
    // -- and usually the whole search of an absent pattern -- by one DRAM access.
    const uint2 *kmer_tab;
    const uint8_t *kmer_steps;      // iterations the reference executes for that k-mer (<= kmer_k)
    uint32_t kmer_k;
    const uint2 *big_tab;
    const uint8_t *big_steps;
    uint32_t big_k;
    // the large table with 16-byte entries (alphabets of <= 4 symbols, HBM-rich indexes): {s, e | FLAG + SA[s], ctx, ctx_len}.
    // For a one-row entry, ctx holds the up to 16 text characters IN FRONT of the row's suffix -- character
    // text[SA[s] - 1 - j] as the 2-bit code c - 1 in bits [2j, 2j + 2) -- and ctx_len how many of them exist (the
    // window ends at the start of the text or in front of a \0).  A pattern with no more than ctx_len characters left is
    // then finished by one 32-bit comparison: the verify tail without its text request.  Replaces big_tab when set.
    const uint4 *big_tab4;
    uint8_t *steps_out;             // nullable: per-pattern executed iterations (table build)
    // nullable: order[t] = pattern handled by thread t.  Patterns bucketed by their k-mer table index
    // (= sorted by SA range start to within one k-mer range) touch the index quasi-sequentially, so
    // concurrently running threads share sectors in L2 and DRAM rows instead of hitting random ones.
    const uint32_t *order;
    // fixed_len is a positive multiple of 16 and pat is 16-byte aligned: stage pattern bytes in shared memory
    uint32_t staged;
    // the index carries the verify structures and the option is on: seed-and-verify tail for one-row ranges
    uint32_t verify;
    // v2 work queue: unfinished patterns after phase A, entry = {pattern id, remaining chars, s, e}
    uint4 *queue;
    unsigned long long *qcount;
    unsigned long long *qcursor;
    // table entries of one-row ranges carry the row's text position instead of e (y = FMX_TAB_POS_FLAG | SA[s];
    // only when n < 2^31 and the dense suffix array exists): the seed-and-verify path then skips the SA request
    uint32_t tab_embed;
    // packed patterns (fmx_query.packed_bits = 2 or 4): pattern p is the 64-bit words [p * packed_wpp, (p+1) * packed_wpp)
    // of `packed`, character k in bits [k * bits, (k+1) * bits) of that stream, stored as character - 1
    const uint64_t *packed;
    uint32_t packed_bits;
    uint32_t packed_wpp;
    // log2 of the bytes per pattern character (0 for u8 indexes): `pat` then holds characters of 2, 4 or 8 bytes,
    // offsets and lengths count characters (character.rs:38-42)
    uint32_t cw_shift;
};
#define FMX_TAB_POS_FLAG 0x80000000u
#define FMX_NOHINT 0xFFFFFFFFu

__device__ __forceinline__ void pattern_span(const SearchArgs &a, uint64_t p, uint64_t &beg, uint32_t &len) {
    if (a.pat_off) {
        beg = a.pat_off[p];
        len = (uint32_t)(a.pat_off[p + 1] - beg);
    } else {
        beg = p * a.fixed_len;
        len = (uint32_t)a.fixed_len;
    }
}

// Pattern bytes through a one-word register cache: consecutive characters of a pattern come from the
// same aligned 32-bit word, so a 32-mer costs 8-9 load instructions instead of 32 (the byte loads were
// ~20 % of the L1TEX wavefronts of k_search and, once index blocks had evicted the pattern lines from
// L1, extra L2 requests: profiles/r01b_target_ncu.txt)
template <bool SECT>
struct ByteReader {
    const uint8_t *q;
    const uint32_t *words;  // the aligned word holding the pattern's first byte
    uint32_t off;           // byte offset of the pattern inside that word
    uint32_t len;
    uint32_t cur = 0xFFFFFFFFu, w = 0;
    __device__ __forceinline__ ByteReader(const uint8_t *q_, uint32_t len_) : q(q_), len(len_) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(q_);
        words = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        off = (uint32_t)(a & 3u);
    }
    // aligned word wi of the stream; bytes outside the pattern read as 0 and are never touched
    __device__ __forceinline__ uint32_t word(uint32_t wi) {
        if (wi != cur) {
            cur = wi;
            const uint32_t first = wi << 2;  // the word covers pattern bytes [first - off, first - off + 4)
            if (first >= off && first + 4u <= off + len) {
                w = SECT ? ldg32_s(words + wi) : __ldg(words + wi);
            } else {  // a word that sticks out of the pattern is never read as a whole: no byte outside the
                      // caller's buffer is touched, whatever its alignment and size
                w = 0;
                for (uint32_t b = 0; b < 4; b++) {
                    const uint32_t pi = first + b;
                    if (pi >= off && pi < off + len) w |= (SECT ? ldg8_s(q + (pi - off)) : (uint32_t)__ldg(q + (pi - off))) << (8u * b);
                }
            }
        }
        return w;
    }
    // character k of the pattern, k < len
    __device__ __forceinline__ uint32_t get(uint32_t k) {
        const uint32_t idx = k + off;
        return (word(idx >> 2) >> (8u * (idx & 3u))) & 0xFFu;
    }
    // characters k .. k+3 (k + 4 <= len), character k in the low byte.  Walking a pattern downwards in steps of four,
    // the lower word of one call is the upper word of the next: one new load per call.
    __device__ __forceinline__ uint32_t get4(uint32_t k) {
        const uint32_t idx = k + off, sh = (idx & 3u) * 8u;
        const uint32_t hi = sh ? word((idx >> 2) + 1u) : 0u;
        const uint32_t lo = word(idx >> 2);
        return __funnelshift_r(lo, hi, sh);
    }
};
typedef ByteReader<false> PatReader;   // pattern bytes: streamed, default fill
typedef ByteReader<true> TextReader;   // text bytes at a random position: sector-granular fill

// Patterns in any input form behind one get(k): bytes (PatReader), packed codes (SearchArgs::packed) or characters of
// 2 / 4 / 8 bytes (SearchArgs::cw_shift; a u64 character beyond 32 bits reads as 0xFFFFFFFF, which is above every
// max_character the index accepts, so it raises the same error as in the reference)
struct AnyReader {
    PatReader pr;
    const uint64_t *pw;
    const uint8_t *wide;
    uint32_t bits, curw, cws;
    uint64_t w;
    __device__ __forceinline__ AnyReader(const SearchArgs &a, uint64_t p, uint64_t beg, uint32_t len)
        : pr(a.pat + ((a.packed_bits || a.cw_shift) ? 0 : beg), (a.packed_bits || a.cw_shift) ? 0u : len),
          pw(a.packed + p * a.packed_wpp), wide(a.pat + (beg << a.cw_shift)), bits(a.packed_bits), curw(0xFFFFFFFFu),
          cws(a.cw_shift), w(0) {}
    __device__ __forceinline__ uint32_t get(uint32_t k) {
        if (cws) {
            if (cws == 1) return __ldg(reinterpret_cast<const uint16_t *>(wide) + k);
            if (cws == 2) return __ldg(reinterpret_cast<const uint32_t *>(wide) + k);
            const unsigned long long v = __ldg(reinterpret_cast<const unsigned long long *>(wide) + k);
            return v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v;
        }
        if (bits == 0) return pr.get(k);
        const uint32_t bit = k * bits, wi = bit >> 6;
        if (wi != curw) {
            curw = wi;
            w = __ldg(pw + wi);
        }
        return ((uint32_t)(w >> (bit & 63u)) & ((1u << bits) - 1u)) + 1u;
    }
    __device__ __forceinline__ uint32_t get4(uint32_t k) {
        if (bits == 0) return pr.get4(k);
        return get(k) | (get(k + 1u) << 8) | (get(k + 2u) << 16) | (get(k + 3u) << 24);
    }
};

// table index of the last K characters of a pattern (base max_character, digit = c - 1);
// false if one of them is \0 or exceeds max_character
template <class Reader>
__device__ __forceinline__ bool kmer_index(Reader &pr, uint32_t len, uint32_t K, uint32_t maxc, uint32_t &idx) {
    uint32_t v = 0;
    bool valid = true;
    for (uint32_t j = 0; j < K; j++) {
        uint32_t c = pr.get(len - K + j);
        valid = valid && (c - 1u) < maxc;
        v = v * maxc + (c - 1u);
    }
    idx = v;
    return valid;
}

// The same for alphabets of four symbols (DNA coded 1..4), four characters per step: the
// bytes c - 1 of a word are two-bit digits d0 (first character) .. d3, and one multiply gathers them into
// d0 * 64 + d1 * 16 + d2 * 4 + d3 (the four partial products land in disjoint bit pairs of the top byte).
template <class Reader>
__device__ __forceinline__ bool kmer_index4(Reader &pr, uint32_t len, uint32_t K, uint32_t &idx) {
    uint32_t v = 0, bad = 0, j = 0;
    for (; j < (K & 3u); j++) {  // the K mod 4 leading characters one by one
        const uint32_t d = pr.get(len - K + j) - 1u;
        bad |= d & ~3u;
        v = (v << 2) | (d & 3u);
    }
    for (; j < K; j += 4u) {
        const uint32_t t = pr.get4(len - K + j) - 0x01010101u;  // per byte c - 1; a \0 character borrows and shows up as invalid
        bad |= t & 0xFCFCFCFCu;
        v = (v << 8) | (((t & 0x03030303u) * 0x40100401u) >> 24);
    }
    idx = v;
    return bad == 0;
}

// memoised start of a fresh search: returns true and sets (s, e, it, len) when a table serves the
// pattern; otherwise the caller walks from (s0, e0) so errors show up (or not) exactly as in the reference
template <class Reader>
__device__ __forceinline__ bool kmer_lookup(const SearchArgs &a, uint32_t maxc, Reader &q, uint32_t &len,
                                            uint32_t &s, uint32_t &e, uint32_t &it, uint32_t *pos = nullptr,
                                            uint32_t *ctx = nullptr, uint32_t *ctx_len = nullptr) {
    const uint2 *tab = nullptr;
    const uint8_t *stp = nullptr;
    uint32_t K = 0;
    if ((a.big_tab != nullptr || a.big_tab4 != nullptr) && len >= a.big_k) {
        tab = a.big_tab;
        stp = a.big_steps;
        K = a.big_k;
    } else if (a.kmer_tab != nullptr && len >= a.kmer_k) {
        tab = a.kmer_tab;
        stp = a.kmer_steps;
        K = a.kmer_k;
    } else {
        return false;
    }
    uint32_t idx;
    if (maxc == 4u) {
        if (!kmer_index4(q, len, K, idx)) return false;
    } else if (!kmer_index(q, len, K, maxc, idx)) {
        return false;
    }
    uint2 t;
    if (tab == a.big_tab && a.big_tab4 != nullptr) {  // 16-byte entries: the row's text context rides along
        const uint4 t4 = ldg128_s(a.big_tab4 + idx);
        t = make_uint2(t4.x, t4.y);
        if (ctx) *ctx = t4.z;
        if (ctx_len) *ctx_len = t4.w;
    } else {
        t = ldg64_s(tab + idx);
    }
    s = t.x;
    e = t.y;
    if (a.tab_embed && (e & FMX_TAB_POS_FLAG)) {  // one-row range: y carries SA[s]
        if (pos) *pos = e & ~FMX_TAB_POS_FLAG;
        e = s + 1u;
    }
    it = K;
    len -= K;
    if (s == e) {
        // the iterations the reference executed before the range emptied: a second random request, made only
        // when the caller asked for work counters (option "count_work") or per-pattern steps (table build)
        if (a.work || a.steps_out) it = ldg8_s(stp + idx);
        len = 0;
    }
    return true;
}

// Fixed-length patterns whose bytes are 16-byte aligned (the batched entry points' usual input): the thread
// pulls 32 characters of ITS pattern with two 128-bit loads into its own column of a shared-memory tile
// and reads characters from there.  The warp's loads are one request per 128-byte line (0.25 per
// pattern); word-by-word reads through L1 cost up to two L2 requests per pattern once index blocks had
// evicted the lines -- of the twelve a 32-mer search on the 1 GB index makes in total.
struct StagedReader {
    uint32_t (*sp)[256];
    const uint8_t *q;
    uint32_t len, wb;  // the tile holds characters [wb, wb + 32)
    __device__ __forceinline__ StagedReader(uint32_t (*sp_)[256], const uint8_t *q_, uint32_t len_) : sp(sp_), q(q_), len(len_) {
        load(len_ > 32u ? len_ - 32u : 0u);
    }
    __device__ __forceinline__ void load(uint32_t w0) {
        wb = w0;
        const uint4 *g = reinterpret_cast<const uint4 *>(q + w0);
        const uint4 v0 = __ldg(g);
        sp[0][threadIdx.x] = v0.x;
        sp[1][threadIdx.x] = v0.y;
        sp[2][threadIdx.x] = v0.z;
        sp[3][threadIdx.x] = v0.w;
        if (w0 + 16u < len) {
            const uint4 v1 = __ldg(g + 1);
            sp[4][threadIdx.x] = v1.x;
            sp[5][threadIdx.x] = v1.y;
            sp[6][threadIdx.x] = v1.z;
            sp[7][threadIdx.x] = v1.w;
        }
    }
    __device__ __forceinline__ uint32_t get(uint32_t k) {
        if (k < wb) load(wb >= 32u ? wb - 32u : 0u);
        const uint32_t idx = k - wb;
        return (sp[idx >> 2][threadIdx.x] >> (8u * (idx & 3u))) & 0xFFu;
    }
    // characters k .. k+3, character k in the low byte
    __device__ __forceinline__ uint32_t get4(uint32_t k) {
        if (k >= wb && k + 4u <= wb + 32u) {
            const uint32_t idx = k - wb, sh = (idx & 3u) * 8u, wi = idx >> 2;
            const uint32_t lo = sp[wi][threadIdx.x], hi = sh ? sp[wi + 1u][threadIdx.x] : 0u;  // wi + 1 <= 7 when sh != 0
            return __funnelshift_r(lo, hi, sh);
        }
        const uint32_t c3 = get(k + 3u), c2 = get(k + 2u), c1 = get(k + 1u), c0 = get(k);  // downwards: the tile moves once
        return c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
    }
};

// The comparison of the seed-and-verify tail: characters rd[0 .. rem) against text[pos - rem .. pos) (`tr` reads that
// window), from the END backwards, four characters per step; stops in front of a \0 of the text or at the first
// mismatch.  Returns the number of characters that match; `c` = the pattern character at the stop (unset when all match).
template <class Reader>
__device__ __forceinline__ uint32_t match_backward(Reader &rd, TextReader &tr, uint32_t rem, uint32_t &c) {
    uint32_t matched = 0;
    while (matched + 4u <= rem) {
        const uint32_t i = rem - 4u - matched;
        const uint32_t tw = tr.get4(i), pw = rd.get4(i);
        const uint32_t stop = __vcmpne4(tw, pw) | __vcmpeq4(tw, 0u);  // 0xFF per byte that ends the match
        if (stop) {
            const uint32_t good = (uint32_t)__clz((int)stop) >> 3;   // matching characters above the first stop byte
            matched += good;
            c = (pw >> (8u * (3u - good))) & 0xFFu;
            return matched;
        }
        matched += 4u;
    }
    while (matched < rem) {
        const uint32_t tc = tr.get(rem - 1u - matched);
        c = rd.get(rem - 1u - matched);
        if (tc == 0u || c != tc) break;
        matched++;
    }
    return matched;
}

// Seed-and-verify tail (fmx_layout.h): the range is the single row s and `rem` characters rd[0 .. rem) are
// still to be consumed.  Locates the row, compares the characters with the text, and jumps to the row the
// reference loop reaches after the characters that match, through the inverse suffix array.  Returns how
// many characters it consumed (s updated; the range is [s, s + 1)); 0 = nothing done (immediate mismatch,
// or the pattern would run off the start of the text).  The caller's ordinary loop takes the next
// character -- the mismatching one empties the range there exactly as in the reference, an invalid one
// raises the error there.  The comparison stops in front of a \0 of the text for every kind: the LF rule of
// MultiPieces for \0 (multi_pieces.rs:140-153) is not a plain rank, and the FM backend's plain rank for \0
// (fm_index.rs:93-95) is not the true LF row when a text holds interior zeros -- the ordinary step reproduces
// either.  Arithmetic mirrored in tests/blobreader.py.
#define FMX_VERIFY_MIN_DENSE 6u     /* fewest remaining characters worth the tail: dense structures (3-4 requests) */
#define FMX_VERIFY_MIN_SAMPLED 10u  /* sampled structures (two short LF walks on top) */
template <int KIND, int LAYOUT, class Reader>
__device__ __forceinline__ uint32_t verify_tail(const FmxDev &ix, const Tabs<LAYOUT> &tb, Reader &rd, uint32_t rem, uint32_t &s) {
    const uint32_t mask = (1u << ix.vsa_level) - 1u;  // 0 with the dense structures: no walk
    uint32_t row = s, st = 0, sym;
    while (row & mask) {
        row = lf_step<KIND, LAYOUT>(ix, tb, row, sym);
        st++;
    }
    uint32_t pos = ldg32_s(ix.vsa + (row >> ix.vsa_level)) + st;  // < 2n < 2^32
    if (pos >= ix.n) pos -= ix.n;
    if (pos < rem) return 0;
    // characters rd[rem-1], rd[rem-2], .. against text[pos-1], text[pos-2], ..
    TextReader tr(ix.text + (pos - rem), rem);
    uint32_t cstop;
    const uint32_t matched = match_backward(rd, tr, rem, cstop);
    if (!matched) return 0;
    const uint32_t q = pos - matched, step = 1u << ix.isa_level;
    uint32_t q4 = (q + step - 1u) & ~(step - 1u), r;
    if (q4 >= ix.n) {  // past the last sample: walk from the final suffix "\0", which is row 0
        q4 = ix.n - 1u;
        r = 0;
    } else {
        r = ldg32_s(ix.isa + (q4 >> ix.isa_level));
    }
    for (uint32_t d = q4 - q; d > 0; d--) r = lf_step<KIND, LAYOUT>(ix, tb, r, sym);
    s = r;
    return matched;
}

// the reference loop for one pattern (wrapper.rs:103-124), characters through `rd`
template <int KIND, int LAYOUT, class Reader>
__device__ __forceinline__ void search_one(const FmxDev &ix, const Tabs<LAYOUT> &tb, const SearchArgs &a, Reader &rd,
                                           uint32_t len, uint32_t &s, uint32_t &e, uint32_t &it) {
    if (a.init_s == nullptr) kmer_lookup(a, ix.max_character, rd, len, s, e, it);
    const uint32_t vmin = ix.isa_level == 0 ? FMX_VERIFY_MIN_DENSE : FMX_VERIFY_MIN_SAMPLED;
    bool armed = KIND != FMX_KIND_RLFM_ && a.verify != 0;
    uint32_t k = len;
    while (k > 0) {
        if (armed && e - s == 1u && k >= vmin) {
            const uint32_t got = verify_tail<KIND, LAYOUT>(ix, tb, rd, k, s);
            armed = got != 0;  // try again after the next ordinary step only if this attempt got somewhere
            if (got) {
                e = s + 1u;
                it += got;
                k -= got;
                continue;
            }
        }
        const uint32_t c = rd.get(k - 1u);
        if (c > ix.max_character) {  // the reference panics here (cs[c] out of bounds, fm_index.rs:94)
            atomicOr(a.err, 1u);
            break;
        }
        lf_map2_pair<KIND, LAYOUT>(ix, tb, c, s, e);
        it++;
        k--;
        if (s == e) break;
    }
}

// Backward search (wrapper.rs:103-124): one pattern per thread, grid-stride; the first kmer_k
// iterations of a fresh search are one table lookup.  This simple shape won the A/B against the
// persistent refill kernels below: once the index sits in L2 the kernel is bound by the number of
// lane-sector requests through L1TEX, which idle (diverged) lanes do not consume
// (profiles/r01_cfg2_search_variants_ncu.txt).
template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_search(const __grid_constant__ FmxDev ix, const __grid_constant__ SearchArgs a) {
    __shared__ Tabs<LAYOUT> tb;
    __shared__ uint32_t spat[8][256];
    load_tables<LAYOUT>(ix, tb);
    unsigned long long steps = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.npat; t += stride) {
        const uint64_t p = a.order ? (uint64_t)a.order[t] : t;
        uint64_t beg;
        uint32_t len;
        pattern_span(a, p, beg, len);
        uint32_t s = a.init_s ? (uint32_t)a.init_s[p] : a.s0;
        uint32_t e = a.init_e ? (uint32_t)a.init_e[p] : a.e0;
        const uint8_t *q = a.pat + beg;
        uint32_t it = 0;
        if (a.packed_bits || a.cw_shift) {
            AnyReader rd(a, p, beg, len);
            search_one<KIND, LAYOUT>(ix, tb, a, rd, len, s, e, it);
        } else if (a.staged) {
            StagedReader rd(spat, q, len);
            search_one<KIND, LAYOUT>(ix, tb, a, rd, len, s, e, it);
        } else {
            PatReader rd(q, len);  // bounds = the whole pattern
            search_one<KIND, LAYOUT>(ix, tb, a, rd, len, s, e, it);
        }
        steps += it;
        a.out_s[p] = s;
        a.out_e[p] = e;
        if (a.steps_out) a.steps_out[p] = (uint8_t)it;
    }
    if (a.work) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
        if ((threadIdx.x & 31) == 0 && steps) atomicAdd(a.work, steps);
    }
}

// patterns for the k-mer table: entry t is the k-mer whose FIRST character is the most significant
// base-sigma digit of t, so table order = lexicographic order = order of the SA ranges
__global__ void k_kmer_patterns(uint32_t k, uint32_t sigma, uint64_t first, uint64_t entries, uint8_t *pat) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= entries) return;
    uint64_t v = first + t;
    for (uint32_t j = k; j-- > 0;) {
        pat[t * k + j] = (uint8_t)(v % sigma + 1);  // digit d stands for the character d + 1
        v /= sigma;
    }
}

__global__ void k_kmer_pack(const uint64_t *s, const uint64_t *e, uint64_t entries, uint2 *tab) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < entries) tab[t] = make_uint2((uint32_t)s[t], (uint32_t)e[t]);
}

// bucketing pass 1: bucket[p] = k-mer table index of pattern p (the extra bucket `entries` takes
// patterns the table cannot serve); hist[bucket]++
__global__ void __launch_bounds__(256) k_bucket_count(const __grid_constant__ SearchArgs a, uint32_t maxc,
                                                       uint32_t entries, uint32_t *bucket, uint32_t *hist) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.npat) return;
    uint64_t beg;
    uint32_t len, idx = entries;
    pattern_span(a, p, beg, len);
    if (len >= a.kmer_k) {
        uint32_t v;
        PatReader rd(a.pat + beg, len);
        if (kmer_index(rd, len, a.kmer_k, maxc, v)) idx = v;
    }
    bucket[p] = idx;
    atomicAdd(hist + idx, 1u);
}
// bucketing pass 2: cursor[] holds the exclusive prefix sum of hist; order[slot] = p
__global__ void __launch_bounds__(256) k_bucket_scatter(const uint32_t *bucket, uint64_t npat, uint32_t *cursor,
                                                         uint32_t *order) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npat) order[atomicAdd(cursor + bucket[p], 1u)] = (uint32_t)p;
}

// Ragged batches (option "order_by_length"): patterns visited in order of their length, so that the 32 lanes of a warp run
// loops of the same trip count (table depth, comparison length).  Counting sort over FMX_LEN_BUCKETS lengths, two
// passes over the offsets; a block counts in shared memory and reserves its slots with one atomic per bucket.
#define FMX_LEN_BUCKETS 128u
__device__ __forceinline__ uint32_t len_bucket(const uint64_t *off, uint64_t p) {
    const uint64_t len = off[p + 1] - off[p];
    return len < FMX_LEN_BUCKETS ? (uint32_t)len : FMX_LEN_BUCKETS - 1u;
}
__global__ void __launch_bounds__(256) k_len_hist(const uint64_t *off, uint64_t npat, uint32_t *hist) {
    __shared__ uint32_t sh[FMX_LEN_BUCKETS];
    for (uint32_t k = threadIdx.x; k < FMX_LEN_BUCKETS; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npat; p += stride) atomicAdd(&sh[len_bucket(off, p)], 1u);
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < FMX_LEN_BUCKETS; k += blockDim.x)
        if (sh[k]) atomicAdd(&hist[k], sh[k]);
}
__global__ void k_len_cursor(uint32_t *hist) {  // exclusive prefix sums, in place (one thread: 128 counters)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t acc = 0;
        for (uint32_t k = 0; k < FMX_LEN_BUCKETS; k++) {
            const uint32_t v = hist[k];
            hist[k] = acc;
            acc += v;
        }
    }
}
#define FMX_LEN_ITEMS 8u
__global__ void __launch_bounds__(256) k_len_scatter(const uint64_t *off, uint64_t npat, uint32_t *cursor, uint32_t *order) {
    __shared__ uint32_t cnt[FMX_LEN_BUCKETS], base[FMX_LEN_BUCKETS];
    for (uint32_t k = threadIdx.x; k < FMX_LEN_BUCKETS; k += blockDim.x) cnt[k] = 0;
    __syncthreads();
    const uint64_t first = (uint64_t)blockIdx.x * (256u * FMX_LEN_ITEMS);
    uint32_t b[FMX_LEN_ITEMS], r[FMX_LEN_ITEMS];
#pragma unroll
    for (uint32_t i = 0; i < FMX_LEN_ITEMS; i++) {
        const uint64_t p = first + i * 256u + threadIdx.x;
        b[i] = 0xFFFFFFFFu;
        if (p < npat) {
            b[i] = len_bucket(off, p);
            r[i] = atomicAdd(&cnt[b[i]], 1u);  // rank inside the block's share of the bucket
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < FMX_LEN_BUCKETS; k += blockDim.x) base[k] = cnt[k] ? atomicAdd(&cursor[k], cnt[k]) : 0u;
    __syncthreads();
#pragma unroll
    for (uint32_t i = 0; i < FMX_LEN_ITEMS; i++)
        if (b[i] != 0xFFFFFFFFu) order[base[b[i]] + r[i]] = (uint32_t)(first + i * 256u + threadIdx.x);
}

// Backward search, persistent variant (option "search_persistent"), phase A: one pattern per thread, fully converged.
// Sets up (s, e) and the number of characters still to consume; the first kmer_k iterations of a
// fresh search are ONE table lookup.  Patterns the table already finishes (their range emptied
// within kmer_k characters, or nothing is left) are final here; the rest go to a work queue.
template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_search_init(const __grid_constant__ FmxDev ix, const __grid_constant__ SearchArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const bool use_tab = a.init_s == nullptr;
    unsigned long long steps = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (a.npat + stride - 1) / stride;
    for (uint64_t rd = 0; rd < rounds; rd++) {
        uint64_t p = rd * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        bool todo = false;
        uint32_t k = 0;
        uint4 ent = make_uint4(0, 0, 0, 0);
        if (p < a.npat) {
            uint64_t beg;
            pattern_span(a, p, beg, k);
            const uint8_t *q = a.pat + beg;
            uint32_t s = a.init_s ? (uint32_t)a.init_s[p] : a.s0;
            uint32_t e = a.init_e ? (uint32_t)a.init_e[p] : a.e0;
            bool done = k == 0;
            if (use_tab) {
                uint32_t it = 0;
                PatReader rd(q, k);
                if (kmer_lookup(a, ix.max_character, rd, k, s, e, it)) {
                    steps += it;
                    done = k == 0;
                }
            }
            todo = !done;
            if (done) {
                a.out_s[p] = s;
                a.out_e[p] = e;
            }
            ent = make_uint4((uint32_t)p, k, s, e);
        }
        unsigned m = __ballot_sync(0xffffffffu, todo);
        if (m) {
            unsigned long long base = 0;
            if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(a.qcount, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (todo) a.queue[base + __popc(m & ((1u << lane) - 1u))] = ent;
        }
    }
    if (a.work) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
        if (lane == 0 && steps) atomicAdd(a.work, steps);
    }
}

// phase B: persistent warps drain the work queue.  Every lane owns one pattern at a time and takes
// its next queue entry the moment it finishes (s == e breaks make run lengths very uneven: the
// one-pattern-per-thread kernel kept 23 of 32 lanes busy).  The next entry of every lane is
// already in registers (loaded one refill ahead), so taking it costs no memory round trip.  Warps
// grab 64-entry chunks from one global cursor, so SMs stay balanced to the end.
#define FMX_QCHUNK 64u
template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_search_steps(const __grid_constant__ FmxDev ix, const __grid_constant__ SearchArgs a) {
    __shared__ Tabs<LAYOUT> tb;
    load_tables<LAYOUT>(ix, tb);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const unsigned long long qn = *a.qcount;
    unsigned long long w_next = 0, w_end = 0;  // this warp's current chunk of the queue
    bool active = false, has_nxt = false, exhausted = qn == 0;
    uint4 nxt = make_uint4(0, 0, 0, 0);
    uint32_t p = 0, k = 0, s = 0, e = 0;
    const uint8_t *q = nullptr;
    unsigned long long steps = 0;
    for (;;) {
        unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle) {
            if (!active && has_nxt) {
                p = nxt.x;
                k = nxt.y;
                s = nxt.z;
                e = nxt.w;
                q = a.pat + (a.pat_off ? a.pat_off[p] : (uint64_t)p * a.fixed_len);
                active = true;
                has_nxt = false;
            }
            unsigned want = __ballot_sync(0xffffffffu, !has_nxt);
            if (want && !exhausted) {
                if (w_next >= w_end) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(a.qcursor, (unsigned long long)FMX_QCHUNK);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    w_next = base;
                    w_end = base + FMX_QCHUNK < qn ? base + FMX_QCHUNK : qn;
                    exhausted = base >= qn;
                }
                if (!exhausted) {
                    unsigned long long mine = w_next + __popc(want & lt);
                    if (!has_nxt && mine < w_end) {
                        nxt = __ldg(a.queue + mine);
                        has_nxt = true;
                    }
                    w_next += __popc(want);
                }
            }
            if (exhausted && __ballot_sync(0xffffffffu, active || has_nxt) == 0) break;
        }
        if (active) {
            uint32_t c = __ldg(q + --k);
            if (c > ix.max_character) {  // the reference panics here (cs[c] out of bounds, fm_index.rs:94)
                atomicOr(a.err, 1u);
                k = 0;
            } else {
                lf_map2_pair<KIND, LAYOUT>(ix, tb, c, s, e);
                steps++;
                if (s == e) k = 0;
            }
            if (k == 0) {
                a.out_s[p] = s;
                a.out_e[p] = e;
                active = false;
            }
        }
    }
    if (a.work) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
        if (lane == 0 && steps) atomicAdd(a.work, steps);
    }
}

// owner[off[p]] = p + 1 for every pattern with at least one candidate row
__global__ void k_mark_owners(const uint64_t *off, uint64_t npat, uint32_t *owner) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npat && off[p + 1] > off[p]) owner[off[p]] = (uint32_t)p + 1u;
}

// the same when only the first `cap` rows have room
__global__ void k_mark_owners_capped(const uint64_t *off, uint64_t npat, uint64_t cap, uint32_t *owner) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npat && off[p + 1] > off[p] && off[p] < cap) owner[off[p]] = (uint32_t)p + 1u;
}

// rows[h] = s[p] + (h - off[p]) with p = owner[h] - 1 (after the max-scan of owner)
// total_dev (nullable): the real number of rows, on the device (fully asynchronous locate)
__global__ void k_expand_rows(const uint64_t *s, const uint64_t *off, const uint32_t *owner, uint64_t total,
                              const uint64_t *total_dev, uint32_t *rows) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (total_dev && *total_dev < total) total = *total_dev;
    if (h < total) {
        uint32_t p = owner[h] - 1u;
        rows[h] = (uint32_t)(s[p] + (h - off[p]));
    }
}

// prefix filter (wrapper.rs:208): flag[h] = (get_l(rows[h]) == 0)
template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_flag_prefix(const __grid_constant__ FmxDev ix, const uint32_t *rows,
                                                     uint64_t total, uint32_t *flag) {
    __shared__ Tabs<LAYOUT> tb;
    load_tables<LAYOUT>(ix, tb);
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h < total) flag[h] = get_l_dev<KIND, LAYOUT>(ix, tb, rows[h]) == 0u ? 1u : 0u;
}

// stream compaction of the kept rows; fpos = exclusive scan of flag
__global__ void k_compact_rows(const uint32_t *rows, const uint32_t *flag, const uint64_t *fpos, uint64_t total,
                               uint32_t *out_rows) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h < total && flag[h]) out_rows[fpos[h]] = rows[h];
}

// out[i] = in[i] + base (chunked pipelines: chunk-local hit offsets -> batch-global ones)
__global__ void k_add_base(const uint64_t *in, uint64_t n, uint64_t base, uint64_t *out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + base;
}

// filtered hit offsets: hit_off[p] = fpos[off[p]]  (fpos has total+1 entries)
__global__ void k_filtered_offsets(const uint64_t *off, const uint64_t *fpos, uint64_t npat, uint64_t *hit_off) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p <= npat) hit_off[p] = fpos[off[p]];
}

struct LocateArgs {
    const uint32_t *rows;       // the hit rows; NULL => row h = s[p] + (h - hoff[p]) with p found by binary search
    const uint64_t *hoff;       // (rows == NULL) npat + 1 hit offsets
    const uint64_t *s;          // (rows == NULL) SA range starts
    uint64_t npat;
    uint64_t first;             // (rows == NULL) index of the first hit to produce: output slot h holds hit first + h
    uint64_t total;             // number of hits, or the capacity when total_dev is given
    const uint64_t *total_dev;  // nullable: the real number of hits, on the device
    uint64_t *positions;  // nullable
    uint64_t *piece_ids;  // nullable (MultiPieces)
    unsigned long long *work;  // [1] += executed LF steps
    uint64_t chunk;            // k_locate: consecutive hits per warp chunk
    uint32_t dense;            // the full suffix array is resident (SEC_VSA) and may be used: position = SA[row]
};

// piece id of a text position (multi_pieces.rs:208-218): the number of piece ends strictly before it
__device__ __forceinline__ uint32_t piece_of(const FmxDev &ix, uint64_t v) {
    uint32_t lo = 0, hi = ix.ndoc;
    while (lo < hi) {
        uint32_t m = lo + ((hi - lo) >> 1);
        if (__ldg(ix.piece_end + m) < v) lo = m + 1; else hi = m;
    }
    return lo;
}

// row of hit h (wrapper.rs:206-216: rows s..e of pattern p, patterns in order): from the expanded row
// list, or -- small batches, where the five expansion launches cost more than they save -- by binary
// search of the hit offsets (neighbouring hits take the same path, so the probes hit L1)
__device__ __forceinline__ uint32_t locate_row(const LocateArgs &a, uint64_t h) {
    if (a.rows) return a.rows[h];
    h += a.first;
    uint64_t lo = 0, hi = a.npat;  // last p with hoff[p] <= h
    while (hi - lo > 1) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(a.hoff + mid) <= h) lo = mid; else hi = mid;
    }
    return (uint32_t)(__ldg(a.s + lo) + (h - __ldg(a.hoff + lo)));
}

// get_sa (fm_index.rs:127-140 / rlfmi.rs:176-189 / multi_pieces.rs:188-201): LF-walk each row to a
// sampled row, one hit per thread.  piece id (multi_pieces.rs:208-218) is derived from the located
// position and the piece boundary table: it equals the number of \0 before the position, which is
// what the reference's walk to the piece start computes (pinned by multi_pieces.rs:287-296).
template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_locate_simple(const __grid_constant__ FmxDev ix, const __grid_constant__ LocateArgs a) {
    __shared__ Tabs<LAYOUT> tb;
    load_tables<LAYOUT>(ix, tb);
    unsigned long long steps = 0;
    const uint32_t mask = (1u << ix.sa_level) - 1u;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t total = a.total;
    if (a.total_dev && *a.total_dev < total) total = *a.total_dev;
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < total; h += stride) {
        uint32_t row = locate_row(a, h);
        uint32_t st = 0, sym;
        uint64_t v;
        if (a.dense) {  // HBM-rich mode: the walk's result is SA[row] by definition; one request, no walk
            v = ldg32_s(ix.vsa + row);
        } else {
            while (row & mask) {
                row = lf_step<KIND, LAYOUT>(ix, tb, row, sym);
                st++;
            }
            v = (uint64_t)ldg32_s(ix.sa + (row >> ix.sa_level)) + st;
            if (v >= ix.n) v -= ix.n;  // (sa + steps) % n; both terms are < n
        }
        if (a.positions) a.positions[h] = v;
        if (KIND == FMX_KIND_MULTI_ && a.piece_ids) a.piece_ids[h] = piece_of(ix, v);
        steps += st;
    }
    if (a.work) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
        if ((threadIdx.x & 31) == 0 && steps) atomicAdd(a.work + 1, steps);
    }
}

// The same, with per-lane refill (option "locate_refill"; NOT the default -- it lost the A/B).  Walk
// lengths are geometric (mean 2^level - 1, long tail), so with one hit per thread a warp waits for its
// slowest lane: ncu showed 8-12 of 32 lanes active and the kernel issue-bound on the RLFM config
// (profiles/r01b_cfg3_ncu.txt: ALU pipe 75 %, 11.7 active threads per instruction).  Here a warp owns
// chunks of `chunk` consecutive hits and a lane takes the next hit the moment its own walk ends, so
// nearly every issued LF step serves 32 lanes.  Measured: 15-20 % SLOWER on every workload
// (profiles/r01b_locate_refill_ab.json).  With one hit per thread the 32 lanes walk ADJACENT rows in
// lockstep, and LF keeps adjacent rows with equal symbols adjacent, so their probes share sectors;
// refilled lanes hold unrelated rows and every probe becomes its own memory request.
template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_locate(const __grid_constant__ FmxDev ix, const __grid_constant__ LocateArgs a) {
    __shared__ Tabs<LAYOUT> tb;
    load_tables<LAYOUT>(ix, tb);
    unsigned long long steps = 0;
    const uint32_t mask = (1u << ix.sa_level) - 1u;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint64_t total = a.total;
    if (a.total_dev && *a.total_dev < total) total = *a.total_dev;
    const uint64_t C = a.chunk;
    uint64_t chunk = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint64_t next = chunk * C < total ? chunk * C : total;
    uint64_t end = next + C < total ? next + C : total;
    bool active = false;
    uint32_t row = 0, st = 0;
    uint64_t h = 0;
    for (;;) {
        if (active && !(row & mask)) {  // sampled row reached: (sa + steps) % n, both terms < n
            uint64_t v = (uint64_t)ldg32_s(ix.sa + (row >> ix.sa_level)) + st;
            if (v >= ix.n) v -= ix.n;
            if (a.positions) a.positions[h] = v;
            if (KIND == FMX_KIND_MULTI_ && a.piece_ids) {
                uint32_t lo = 0, hi = ix.ndoc;  // number of piece ends strictly before v
                while (lo < hi) {
                    uint32_t m = lo + ((hi - lo) >> 1);
                    if (__ldg(ix.piece_end + m) < v) lo = m + 1; else hi = m;
                }
                a.piece_ids[h] = lo;
            }
            steps += st;
            active = false;
        }
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle) {
            if (next == end && next < total) {  // this warp's next chunk
                chunk += nwarps;
                next = chunk * C < total ? chunk * C : total;
                end = next + C < total ? next + C : total;
            }
            const uint64_t avail = end - next;
            if (avail) {
                const uint32_t r = __popc(idle & lt);
                if (!active && r < avail) {
                    h = next + r;
                    row = locate_row(a, h);
                    st = 0;
                    active = true;
                }
                const uint32_t want = __popc(idle);
                next += want < avail ? want : avail;
            } else if (idle == 0xffffffffu) {
                break;
            }
        }
        if (active && (row & mask)) {
            uint32_t sym;
            row = lf_step<KIND, LAYOUT>(ix, tb, row, sym);
            st++;
        }
    }
    if (a.work) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
        if (lane == 0 && steps) atomicAdd(a.work + 1, steps);
    }
}

// get_sa with STABLE warp-level compaction (option "locate_refill" = 2).  The refill kernel above lost
// because it drops new hits into whichever lanes are idle, so neighbouring lanes stop holding neighbouring
// rows and their probes stop sharing sectors.  Here a warp owns chunks of consecutive hits and keeps its
// live walks in HIT ORDER: when fewer than FMX_COMPACT_BELOW lanes are busy the survivors are shifted down
// (order preserved, three shuffles) and the free lanes on top take the next hits of the chunk, in order.
// Survivors of a run are its unsampled rows -- still near-adjacent -- so coalescing mostly survives while
// nearly every issued LF step serves a full warp.
#define FMX_COMPACT_BELOW 25
template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_locate_stable(const __grid_constant__ FmxDev ix, const __grid_constant__ LocateArgs a) {
    __shared__ Tabs<LAYOUT> tb;
    __shared__ uint32_t xch[8][3][32];
    const uint32_t warp = threadIdx.x >> 5;
    load_tables<LAYOUT>(ix, tb);
    unsigned long long steps = 0;
    const uint32_t mask = (1u << ix.sa_level) - 1u;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint64_t total = a.total;
    if (a.total_dev && *a.total_dev < total) total = *a.total_dev;
    const uint64_t C = a.chunk;
    for (uint64_t chunk = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk * C < total; chunk += nwarps) {
        const uint64_t base = chunk * C;
        const uint32_t cnt = (uint32_t)(base + C < total ? C : total - base);
        uint32_t next = 0;  // hits of the chunk handed out so far
        bool active = false;
        uint32_t row = 0, st = 0, hl = 0;  // hl = hit index inside the chunk
        for (;;) {
            if (active && !(row & mask)) {  // sampled row reached: (sa + steps) % n, both terms < n
                uint64_t v = (uint64_t)ldg32_s(ix.sa + (row >> ix.sa_level)) + st;
                if (v >= ix.n) v -= ix.n;
                if (a.positions) a.positions[base + hl] = v;
                if (KIND == FMX_KIND_MULTI_ && a.piece_ids) {
                    uint32_t lo = 0, hi = ix.ndoc;  // number of piece ends strictly before v
                    while (lo < hi) {
                        uint32_t m = lo + ((hi - lo) >> 1);
                        if (__ldg(ix.piece_end + m) < v) lo = m + 1; else hi = m;
                    }
                    a.piece_ids[base + hl] = lo;
                }
                steps += st;
                active = false;
            }
            const unsigned live = __ballot_sync(0xffffffffu, active);
            const uint32_t nlive = __popc(live);
            if (nlive < FMX_COMPACT_BELOW && next < cnt) {
                // stable compaction through shared memory: live lane -> slot popc(live lanes below it)
                // (a pull by shuffle needs __fns, which is a software loop)
                if (active) {
                    const uint32_t d = __popc(live & lt);
                    xch[warp][0][d] = row;
                    xch[warp][1][d] = st;
                    xch[warp][2][d] = hl;
                }
                __syncwarp();
                active = lane < nlive;
                row = xch[warp][0][lane];
                st = xch[warp][1][lane];
                hl = xch[warp][2][lane];
                __syncwarp();
                if (!active && next + (lane - nlive) < cnt) {  // the next hits of the chunk, in order
                    hl = next + (lane - nlive);
                    row = locate_row(a, base + hl);
                    st = 0;
                    active = true;
                }
                const uint32_t room = 32u - nlive, left = cnt - next;
                next += room < left ? room : left;
            } else if (live == 0) {
                break;
            }
            if (active && (row & mask)) {
                uint32_t sym;
                row = lf_step<KIND, LAYOUT>(ix, tb, row, sym);
                st++;
            }
        }
    }
    if (a.work) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
        if (lane == 0 && steps) atomicAdd(a.work + 1, steps);
    }
}

// get_sa for whole SA ranges (option "locate_ranges"; automatic when patterns have many matches each).
// The rows of one pattern are ADJACENT (s .. e-1).  While all rows of a sub-range [a, a+len) carry the same
// BWT symbol c -- the normal case on repetitive texts, which is what RLFMIndex is for -- LF maps them to
// the adjacent rows lf_map2(c, a) .. +len, so ONE lf_map2 pair steps the whole sub-range: the check is
// lf_map2(c, a + len) - lf_map2(c, a) == len.  At every step the members whose current row is sampled
// (row % 2^level == 0) take their position sa[row >> level] + t and leave; a sub-range that is not uniform
// is split in halves.  Size-1 sub-ranges are the reference's per-row walk, so the result (and the count of
// executed LF steps) is identical by construction; the per-row kernel above walks 64 copies of the same
// path, this one walks it once.  One thread handles LOCATE_GROUP consecutive hits of the hit list.
// Measured on BASELINE config 3 (6.1e8 hits, 64 copies): 31.0 ms against 35.1 ms for the per-row kernel --
// it executes ~10x fewer instructions but its loads and stores are one memory request per thread, where
// the per-row kernel's 32 adjacent rows share theirs (1.25 requests per hit), so it trades the issue
// bound for the request-rate bound.  Staging the positions in shared memory for a coalesced store was
// slower still (53 ms: occupancy).  Automatic for RLFM indexes only (a user picks those for repetitive
// texts); on non-repetitive texts every sub-range splits at once and this degenerates to row walks
// with overhead.
#define LOCATE_GROUP 64
#define LOCATE_STACK 8
#define LOCATE_RANGE_THREADS 256
template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(LOCATE_RANGE_THREADS) k_locate_ranges(const __grid_constant__ FmxDev ix,
                                                                        const __grid_constant__ LocateArgs a) {
    __shared__ Tabs<LAYOUT> tb;
    load_tables<LAYOUT>(ix, tb);
    unsigned long long steps = 0;
    uint64_t total = a.total;
    if (a.total_dev && *a.total_dev < total) total = *a.total_dev;
    const uint32_t M = 1u << ix.sa_level;  // host guarantees sa_level <= 6
    uint64_t cm = 0;                       // bit j set for j = 0, M, 2M, .. < 64
    for (uint32_t j = 0; j < 64; j += M) cm |= 1ull << j;
    const uint64_t H0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * LOCATE_GROUP;
    if (H0 < total) {
        const uint64_t H1 = H0 + LOCATE_GROUP < total ? H0 + LOCATE_GROUP : total;
        uint64_t p = 0;
        {
            uint64_t lo = 0, hi = a.npat;  // last p with hoff[p] <= first + H0
            while (hi - lo > 1) {
                uint64_t mid = lo + ((hi - lo) >> 1);
                if (__ldg(a.hoff + mid) <= a.first + H0) lo = mid; else hi = mid;
            }
            p = lo;
        }
        uint64_t h = H0;
        while (h < H1) {
            const uint64_t g = a.first + h;  // global hit number
            while (__ldg(a.hoff + p + 1) <= g) p++;
            const uint64_t pend = __ldg(a.hoff + p + 1) - a.first;  // this pattern's hits end here (local numbering)
            const uint64_t seg_end = pend < H1 ? pend : H1;
            // sub-range stack: rows [row, row + 64 - clz(pending)), output slot of member 0, steps so far, unresolved members
            uint32_t st_row[LOCATE_STACK], st_t[LOCATE_STACK];
            uint64_t st_out[LOCATE_STACK], st_pend[LOCATE_STACK];
            int sp = 0;
            {
                const uint32_t len = (uint32_t)(seg_end - h);
                st_row[0] = (uint32_t)(__ldg(a.s + p) + (g - __ldg(a.hoff + p)));
                st_out[0] = h;
                st_t[0] = 0;
                st_pend[0] = len >= 64 ? ~0ull : ((1ull << len) - 1ull);
                sp = 1;
            }
            while (sp > 0) {
                sp--;
                uint32_t row = st_row[sp], t = st_t[sp];
                uint64_t out = st_out[sp], pending = st_pend[sp];
                for (;;) {
                    // members on sampled rows leave with their position
                    const uint32_t r = (0u - row) & (M - 1u);
                    uint64_t hit = r < 64 ? (pending & (cm << r)) : 0ull;
                    pending &= ~hit;
                    while (hit) {
                        const uint32_t j = (uint32_t)__ffsll((long long)hit) - 1u;
                        hit &= hit - 1;
                        uint64_t v = (uint64_t)ldg32_s(ix.sa + ((row + j) >> ix.sa_level)) + t;
                        if (v >= ix.n) v -= ix.n;  // (sa + steps) % n; both terms are < n
                        if (a.positions) a.positions[out + j] = v;
                        if (KIND == FMX_KIND_MULTI_ && a.piece_ids) {
                            uint32_t lo = 0, hi = ix.ndoc;  // number of piece ends strictly before v
                            while (lo < hi) {
                                uint32_t m = lo + ((hi - lo) >> 1);
                                if (__ldg(ix.piece_end + m) < v) lo = m + 1; else hi = m;
                            }
                            a.piece_ids[out + j] = lo;
                        }
                        steps += t;
                    }
                    if (!pending) break;
                    // drop resolved members at both ends
                    const uint32_t tz = (uint32_t)__ffsll((long long)pending) - 1u;
                    pending >>= tz;
                    row += tz;
                    out += tz;
                    const uint32_t len = 64u - (uint32_t)__clzll((long long)pending);
                    uint32_t sym;
                    const uint32_t nrow = lf_step<KIND, LAYOUT>(ix, tb, row, sym);
                    bool uniform = len == 1;
                    if (!uniform && !(KIND == FMX_KIND_MULTI_ && sym == 0))
                        uniform = lf_map2_dev<KIND, LAYOUT>(ix, tb, sym, row + len) - nrow == len;
                    if (uniform) {
                        row = nrow;
                        t++;
                        continue;
                    }
                    // split in halves; the upper half waits on the stack (it cannot overflow: every split
                    // halves the length, 64 -> 1 in 6 splits, and a popped entry frees its slot first)
                    const uint32_t half = len >> 1;
                    const uint64_t up = pending >> half;
                    if (up) {
                        st_row[sp] = row + half;
                        st_out[sp] = out + half;
                        st_t[sp] = t;
                        st_pend[sp] = up;
                        sp++;
                    }
                    pending &= (1ull << half) - 1ull;
                    if (!pending) break;
                }
            }
            h = seg_end;
        }
    }
    if (a.work) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
        if ((threadIdx.x & 31) == 0 && steps) atomicAdd(a.work + 1, steps);
    }
}

// iter_chars_backward / iter_chars_forward, k characters per row (wrapper.rs:143-183)
template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_extract(const __grid_constant__ FmxDev ix, const uint64_t *rows,
                                                 uint64_t nrows, uint32_t k, int forward, uint8_t *out,
                                                 uint32_t *out_len) {
    __shared__ Tabs<LAYOUT> tb;
    load_tables<LAYOUT>(ix, tb);
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    uint32_t i = (uint32_t)rows[r], got = 0;
    const uint32_t cws = ix.cw_shift;  // characters leave in the width of the text the index was built from
    uint8_t *o = out + ((r * k) << cws);
    auto put = [&](uint32_t t, uint32_t c) {
        if (cws == 0) o[t] = (uint8_t)c;
        else if (cws == 1) reinterpret_cast<uint16_t *>(o)[t] = (uint16_t)c;
        else if (cws == 2) reinterpret_cast<uint32_t *>(o)[t] = c;
        else reinterpret_cast<unsigned long long *>(o)[t] = c;
    };
    for (uint32_t t = 0; t < k; t++) {
        uint32_t c, nx;
        if (!forward) {
            nx = lf_step<KIND, LAYOUT>(ix, tb, i, c);
        } else if (!fl_step<KIND, LAYOUT>(ix, tb, i, c, nx)) {
            break;
        }
        put(t, c);
        i = nx;
        got++;
    }
    for (uint32_t t = got; t < k; t++) put(t, 0u);
    if (out_len) out_len[r] = got;
}

// The same characters straight from the text, for indexes that hold the text and the full suffix array (HBM-rich mode).
// With p = SA[row]: the reference's forward iterator yields text[p], text[p + 1], .. (get_f(i) = text[SA[i]], fl_map(i) =
// row of suffix SA[i] + 1; a MultiPieces index stops IN FRONT of a \0, multi_pieces.rs:171-181), its backward iterator
// text[p - 1], text[p - 2], .. (get_l(i) = text[SA[i] - 1], lf_map(i) = row of suffix SA[i] - 1), both cyclic over the
// text: fm_index.rs:78-120, wrapper.rs:143-183.  (Single texts with interior \0, where the reference's lf_map2(0, .) is
// not the true LF row, never carry these structures.)  So k characters cost one suffix-array request plus the one or
// two text sectors they lie in, instead of k rank -- or k select -- probes.
// Eight lanes per row: four characters each per round, aligned 4-byte loads and stores.
#define FMX_XT_LANES 8u
template <int KIND>
__global__ void __launch_bounds__(256) k_extract_text(const __grid_constant__ FmxDev ix, const uint64_t *rows, uint64_t nrows,
                                                      uint32_t k, int forward, uint8_t *out, uint32_t *out_len) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / FMX_XT_LANES;
    const uint32_t j = threadIdx.x % FMX_XT_LANES;
    if (r >= nrows) return;
    const uint32_t n = ix.n;
    const uint32_t p = ldg32_s(ix.vsa + (uint32_t)rows[r]);
    uint8_t *o = out + r * k;
    uint32_t got = k;
    if (KIND == FMX_KIND_MULTI_ && forward) {  // the walk ends in front of the piece's \0
        const uint32_t end = __ldg(ix.piece_end + piece_of(ix, p));
        if (end - p < got) got = end - p;
    }
    if (j == 0 && out_len) out_len[r] = got;
    const uint32_t *tw = reinterpret_cast<const uint32_t *>(ix.text);
    const bool inside = forward ? (uint64_t)p + k <= n : p >= k;  // no wrap around the end of the text
    if (inside && (k & 3u) == 0 && (reinterpret_cast<uintptr_t>(o) & 3u) == 0) {
        for (uint32_t t = 4u * j; t < k; t += 4u * FMX_XT_LANES) {
            const uint32_t a = forward ? p + t : p - t - 4u;  // the four text bytes, ascending
            const uint32_t sh = (a & 3u) * 8u;
            const uint32_t lo = ldg32_s(tw + (a >> 2));
            const uint32_t hi = sh ? ldg32_s(tw + (a >> 2) + 1u) : 0u;
            uint32_t v = __funnelshift_r(lo, hi, sh);
            if (!forward) v = __byte_perm(v, 0u, 0x0123);  // out[t] = text[p - 1 - t]
            if (t + 4u > got) v = t >= got ? 0u : v & (0xFFFFFFFFu >> (8u * (t + 4u - got)));
            *reinterpret_cast<uint32_t *>(o + t) = v;
        }
        return;
    }
    for (uint32_t t = j; t < k; t += FMX_XT_LANES) {  // odd sizes and walks that wrap: byte by byte
        uint32_t c = 0;
        if (t < got) {
            uint64_t q = forward ? ((uint64_t)p + t) % n : ((uint64_t)p + (uint64_t)n - 1u - (t % n)) % n;
            c = ldg8_s(ix.text + q);
        }
        o[t] = (uint8_t)c;
    }
}

// backend primitives over a batch of rows, for parity tests (src/backend.rs:5-40)
template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_rows_op(const __grid_constant__ FmxDev ix, int op, const uint64_t *rows,
                                                 uint64_t nrows, uint64_t *out) {
    __shared__ Tabs<LAYOUT> tb;
    load_tables<LAYOUT>(ix, tb);
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    uint32_t i = (uint32_t)rows[r], c, nx;
    uint64_t res = 0;
    switch (op) {
        case 0: res = get_l_dev<KIND, LAYOUT>(ix, tb, i); break;
        case 1: res = lf_step<KIND, LAYOUT>(ix, tb, i, c); break;
        case 2:
            if (KIND == FMX_KIND_RLFM_) res = cs_search(tb.cs, ix.cs_len, rbv_rank1(ix.rl_bp, i + 1) - 1u);
            else res = cs_search(tb.cs, ix.cs_len, i);
            break;
        case 3: res = fl_step<KIND, LAYOUT>(ix, tb, i, c, nx) ? (uint64_t)nx : FMX_NONE; break;
        case 4: {
            const uint32_t mask = (1u << ix.sa_level) - 1u;
            uint32_t st = 0;
            while (i & mask) {
                i = lf_step<KIND, LAYOUT>(ix, tb, i, c);
                st++;
            }
            uint64_t v = (uint64_t)ldg32_s(ix.sa + (i >> ix.sa_level)) + st;
            res = v >= ix.n ? v - ix.n : v;
            break;
        }
        case 5: {  // the reference's literal piece_id walk (multi_pieces.rs:208-218)
            for (;;) {
                uint32_t nxt = lf_step<KIND, LAYOUT>(ix, tb, i, c);
                if (c == 0) {
                    uint32_t rank0 = seq_lf<LAYOUT>(ix, tb, 0, i);  // rank(bw, i, 0); cs[0] == 0
                    res = (uint64_t)((__ldg(ix.doc + rank0) + 1u) % ix.ndoc);
                    break;
                }
                i = nxt;
            }
            break;
        }
    }
    out[r] = res;
}

template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_lf_map2(const __grid_constant__ FmxDev ix, const uint32_t *c, const uint64_t *i,
                                                 uint64_t nrows, uint64_t *out) {
    __shared__ Tabs<LAYOUT> tb;
    load_tables<LAYOUT>(ix, tb);
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    out[r] = lf_map2_dev<KIND, LAYOUT>(ix, tb, c[r], (uint32_t)i[r]);
}

// ------------------------------------------------------------------ scans (3-phase, hand written)
// items per block = SCAN_THREADS * SCAN_ITEMS

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

struct OpSum {
    __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a + b; }
    __device__ __forceinline__ uint64_t identity() const { return 0; }
};
struct OpMax {
    __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a > b ? a : b; }
    __device__ __forceinline__ uint64_t identity() const { return 0; }
};

template <class Op>
__device__ __forceinline__ uint64_t block_scan_excl(uint64_t v, Op op, uint64_t &total, uint64_t *s_warp) {
    // inclusive scan inside the warp
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc = op(t, inc);
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = lane < (SCAN_THREADS / 32) ? s_warp[lane] : op.identity();
        uint64_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc = op(t, winc);
        }
        if (lane < (SCAN_THREADS / 32)) s_warp[lane] = winc;
    }
    __syncthreads();
    uint64_t warp_prefix = wid ? s_warp[wid - 1] : op.identity();
    total = s_warp[SCAN_THREADS / 32 - 1];
    uint64_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = op.identity();
    __syncthreads();
    return op(warp_prefix, excl);
}

// scan inputs: a plain array, or the candidate-row count e - s of every SA range computed on the fly
// (wrapper.rs:132-139; saves the k_range_counts pass and its array)
template <class T>
struct LoadPtr {
    const T *p;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return (uint64_t)p[i]; }
};
struct LoadRangeCount {
    const uint64_t *s, *e;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const {
        uint64_t a = s[i], b = e[i];
        return b > a ? b - a : 0;
    }
};

// phase 1: per-tile reduction
template <class Load, class Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(Load in, uint64_t n, uint64_t *tile_sum, Op op) {
    __shared__ uint64_t s_warp[SCAN_THREADS / 32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t acc = op.identity();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) acc = op(acc, in(base + k));
    uint64_t total;
    block_scan_excl(acc, op, total, s_warp);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// phase 3: per-tile scan with the carry of the preceding tiles.  EXCL: exclusive (out[i] excludes in[i]).
// The carry is tile_prefix[blockIdx.x] when a scanned prefix array is given; otherwise, when tile_sums is
// given, the block reduces the sums of the tiles before it itself (small inputs: no middle kernel).
// When `write_total` is set and EXCL, out[n] receives the grand total (out has n+1 entries).
template <class Load, class Tout, class Op, bool EXCL>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(Load in, uint64_t n, const uint64_t *tile_prefix,
                                                             const uint64_t *tile_sums, Tout *out, Op op, int write_total) {
    __shared__ uint64_t s_warp[SCAN_THREADS / 32];
    uint64_t tile_carry = op.identity();
    if (tile_prefix) {
        tile_carry = tile_prefix[blockIdx.x];
    } else if (tile_sums) {
        uint64_t part = op.identity();
        for (uint32_t i = threadIdx.x; i < blockIdx.x; i += SCAN_THREADS) part = op(part, tile_sums[i]);
        block_scan_excl(part, op, tile_carry, s_warp);
    }
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS];
    uint64_t acc = op.identity();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? in(base + k) : op.identity();
        acc = op(acc, v[k]);
    }
    uint64_t total;
    uint64_t pre = block_scan_excl(acc, op, total, s_warp);
    uint64_t carry = op(tile_carry, pre);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint64_t inc = op(carry, v[k]);
        if (base + k < n) out[base + k] = (Tout)(EXCL ? carry : inc);
        carry = inc;
    }
    if (EXCL && write_total && base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = (Tout)carry;
}

// ------------------------------------------------------------------ random gather microbenchmark
// independent random loads of WIDTH bytes (32 = one sector, 64, 128 = a full line), no dependent
// chain: the random-access roofline denominator
template <int WIDTH>
__global__ void __launch_bounds__(256) k_random_gather(const uint4 *buf, uint64_t nunits, uint64_t nloads,
                                                       uint64_t seed, uint32_t *sink) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
#pragma unroll 4
    for (uint64_t k = t; k < nloads; k += stride) {
        uint64_t z = (k + seed) * 0x9E3779B97F4A7C15ull;  // splitmix64 finaliser
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        uint64_t unit = __umul64hi(z, nunits);  // uniform in [0, nunits) without a 64-bit division
#pragma unroll
        for (int q = 0; q < WIDTH / 32; q++) {
            RB b = rb_load(buf, (uint32_t)(unit * (WIDTH / 32) + q));
            acc += b.w[0] ^ b.w[7];
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

}  // namespace fmx
