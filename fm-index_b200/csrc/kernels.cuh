// kernels.cuh -- hand-written sm_100a kernels for the FM-index query path.
//
// Everything here is HBM-latency / sector-rate bound integer work: no tensor cores.
// One rank probe = one 32-byte sector (one 256-bit LDG) + an in-register popcount.
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
    uint32_t w[8];  // w[0] = ones before the block, w[1..7] = 224 payload bits
};

// one 32-byte sector, one instruction (sm_100: LDG.E.256), read-only path
__device__ __forceinline__ RB rb_load(const uint4 *__restrict__ v, uint32_t blk) {
    RB b;
    const uint4 *p = v + 2ull * blk;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]), "=r"(b.w[4]), "=r"(b.w[5]),
                   "=r"(b.w[6]), "=r"(b.w[7])
                 : "l"(p));
    return b;
}

// ones in [0, blk*224 + r), r in [0, 224)
__device__ __forceinline__ uint32_t rb_rank(const RB &b, uint32_t r) {
    uint32_t full = r >> 5, rem = r & 31u, c = b.w[0];
    uint32_t part = (1u << rem) - 1u;
#pragma unroll
    for (uint32_t k = 0; k < 7; k++) {
        uint32_t m = k < full ? 0xFFFFFFFFu : (k == full ? part : 0u);
        c += __popc(b.w[k + 1] & m);
    }
    return c;
}

__device__ __forceinline__ uint32_t rb_bit(const RB &b, uint32_t r) {
    uint32_t full = r >> 5, word = b.w[1];
#pragma unroll
    for (uint32_t k = 1; k < 7; k++) word = (full == k) ? b.w[k + 1] : word;
    return (word >> (r & 31u)) & 1u;
}

__device__ __forceinline__ void rb_split(uint32_t pos, uint32_t &blk, uint32_t &r) {
    blk = pos / FMX_RB_BITS;
    r = pos - blk * FMX_RB_BITS;
}

// rank1 of a stand-alone RB32 vector
__device__ __forceinline__ uint32_t rbv_rank1(const uint4 *__restrict__ v, uint32_t pos) {
    uint32_t blk, r;
    rb_split(pos, blk, r);
    RB b = rb_load(v, blk);
    return rb_rank(b, r);
}

// select1: position of the k-th (0-based) one of an RB32 vector with nblk blocks.  k must exist.
__device__ __forceinline__ uint32_t rbv_select1(const uint4 *__restrict__ v, uint32_t nblk, uint32_t k) {
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(v);
    uint32_t lo = 0, hi = nblk;  // last block with count <= k
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(cw + 8ull * mid) <= k) lo = mid; else hi = mid;
    }
    RB b = rb_load(v, lo);
    uint32_t rem = k - b.w[0], pos = lo * FMX_RB_BITS;
#pragma unroll
    for (uint32_t j = 1; j < 8; j++) {
        uint32_t c = __popc(b.w[j]);
        if (rem < c) return pos + __fns(b.w[j], 0, rem + 1);
        rem -= c;
        pos += 32;
    }
    return pos;  // unreachable for valid k
}
__device__ __forceinline__ uint32_t rbv_select0(const uint4 *__restrict__ v, uint32_t nblk, uint32_t k) {
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(v);
    uint32_t lo = 0, hi = nblk;  // last block with zeros-before <= k
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (mid * FMX_RB_BITS - __ldg(cw + 8ull * mid) <= k) lo = mid; else hi = mid;
    }
    RB b = rb_load(v, lo);
    uint32_t rem = k - (lo * FMX_RB_BITS - b.w[0]), pos = lo * FMX_RB_BITS;
#pragma unroll
    for (uint32_t j = 1; j < 8; j++) {
        uint32_t z = ~b.w[j];
        uint32_t c = __popc(z);
        if (rem < c) return pos + __fns(z, 0, rem + 1);
        rem -= c;
        pos += 32;
    }
    return pos;
}

// ------------------------------------------------------------------ wavelet matrix

// position of `pos` after walking symbol c's path down all levels (rank(i,c) = walk - walk_c(0))
__device__ __forceinline__ uint32_t wm_walk(const FmxDev &ix, uint32_t c, uint32_t pos) {
    const uint32_t L = ix.levels;
#pragma unroll 1
    for (uint32_t l = 0; l < L; l++) {
        uint32_t blk, r;
        rb_split(pos, blk, r);
        RB b = rb_load(ix.lv[l], blk);
        uint32_t ones = rb_rank(b, r);
        pos = ((c >> (L - 1 - l)) & 1u) ? ix.zeros[l] + ones : pos - ones;
    }
    return pos;
}

// the same for the two ends of an SA range; shares the sector when both fall in one block
__device__ __forceinline__ void wm_walk2(const FmxDev &ix, uint32_t c, uint32_t &s, uint32_t &e) {
    const uint32_t L = ix.levels;
#pragma unroll 1
    for (uint32_t l = 0; l < L; l++) {
        uint32_t bs, rs, be, re;
        rb_split(s, bs, rs);
        rb_split(e, be, re);
        const uint4 *v = ix.lv[l];
        RB a = rb_load(v, bs);
        RB b = a;
        if (be != bs) b = rb_load(v, be);
        uint32_t os = rb_rank(a, rs), oe = rb_rank(b, re);
        if ((c >> (L - 1 - l)) & 1u) {
            uint32_t z = ix.zeros[l];
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
__device__ __forceinline__ uint32_t wm_access_walk(const FmxDev &ix, uint32_t pos, uint32_t &sym) {
    const uint32_t L = ix.levels;
    uint32_t c = 0;
#pragma unroll 1
    for (uint32_t l = 0; l < L; l++) {
        uint32_t blk, r;
        rb_split(pos, blk, r);
        RB b = rb_load(ix.lv[l], blk);
        uint32_t ones = rb_rank(b, r), bit = rb_bit(b, r);
        c = (c << 1) | bit;
        pos = bit ? ix.zeros[l] + ones : pos - ones;
    }
    sym = c;
    return pos;
}

// select_u64_unchecked(k, c): position of the k-th c.  base = walk_c(0).
__device__ __forceinline__ uint32_t wm_select(const FmxDev &ix, uint32_t c, uint32_t k, uint32_t base) {
    const uint32_t L = ix.levels;
    const uint32_t nblk = ix.seq_len / FMX_RB_BITS + 1;
    uint32_t pos = base + k;
#pragma unroll 1
    for (uint32_t l = L; l-- > 0;) {
        if ((c >> (L - 1 - l)) & 1u) pos = rbv_select1(ix.lv[l], nblk, pos - ix.zeros[l]);
        else pos = rbv_select0(ix.lv[l], nblk, pos);
    }
    return pos;
}

// ------------------------------------------------------------------ backend primitives
// (crate-private seam src/backend.rs:5-40)

// MultiPieces rule for c == 0 (multi_pieces.rs:142-148); rank = rank(bw, i, 0)
__device__ __forceinline__ uint32_t multi_zero_rule(const FmxDev &ix, uint32_t i, uint32_t rank) {
    return i < ix.first_row ? rank + 1u : (i == ix.first_row ? 0u : rank);
}

// ---- RLFM pieces (rlfmi.rs:118-170)
// run index holding row i and whether that run's head is c, fused with rank(s, j, c):
//   j   = rank1(b, i)
//   h   = index of the run containing row i  (= rank1(b, i+1) - 1, clamped at i == n)
//   nr  = rank(s, j, c)
//   hit = (s[h] == c)
// The access of s[h] is done along c's path (h is j or j-1, so it shares sectors with the rank walk).
__device__ __forceinline__ void rl_probe(const FmxDev &ix, uint32_t c, uint32_t i, uint32_t &j, uint32_t &nr,
                                         bool &hit) {
    uint32_t blk, r;
    rb_split(i, blk, r);
    RB bb = rb_load(ix.rl_b, blk);
    j = rb_rank(bb, r);
    uint32_t starts_here = (i < ix.n) ? rb_bit(bb, r) : 0u;
    uint32_t h = starts_here ? j : j - 1u;  // b[0] == 1 so j >= 1 whenever !starts_here
    const uint32_t L = ix.levels;
    uint32_t p = j, q = h;
    bool alive = true;
#pragma unroll 1
    for (uint32_t l = 0; l < L; l++) {
        uint32_t bit = (c >> (L - 1 - l)) & 1u;
        uint32_t bp_, rp, bq, rq;
        rb_split(p, bp_, rp);
        rb_split(q, bq, rq);
        const uint4 *v = ix.lv[l];
        RB a = rb_load(v, bp_);
        uint32_t op = rb_rank(a, rp);
        if (alive) {
            RB d = a;
            if (bq != bp_) d = rb_load(v, bq);
            uint32_t oq = rb_rank(d, rq);
            alive = rb_bit(d, rq) == bit;
            q = bit ? ix.zeros[l] + oq : q - oq;
        }
        p = bit ? ix.zeros[l] + op : p - op;
    }
    nr = p;  // still offset by walk_c(0); the caller adds adj[c] = cs[c] - walk_c(0)
    hit = alive;
}

// lf_map2 for every kind; s_adj = shared-memory copy of adj[]
template <int KIND>
__device__ __forceinline__ uint32_t lf_map2_dev(const FmxDev &ix, const uint32_t *s_adj, uint32_t c, uint32_t i) {
    if (KIND == FMX_KIND_RLFM_) {
        uint32_t j, nr;
        bool hit;
        rl_probe(ix, c, i, j, nr, hit);
        uint32_t t = __ldg(ix.rl_bpsel + (s_adj[c] + nr));
        if (!hit) return t;
        return t + i - __ldg(ix.rl_bsel + j);
    } else {
        uint32_t w = s_adj[c] + wm_walk(ix, c, i);
        if (KIND == FMX_KIND_MULTI_ && c == 0) return multi_zero_rule(ix, i, w);
        return w;
    }
}

// (s, e) <- (lf_map2(c, s), lf_map2(c, e))   (wrapper.rs:109-110)
template <int KIND>
__device__ __forceinline__ void lf_map2_pair(const FmxDev &ix, const uint32_t *s_adj, uint32_t c, uint32_t &s,
                                             uint32_t &e) {
    if (KIND == FMX_KIND_RLFM_) {
        s = lf_map2_dev<KIND>(ix, s_adj, c, s);
        e = lf_map2_dev<KIND>(ix, s_adj, c, e);
    } else {
        uint32_t s0 = s, e0 = e;
        wm_walk2(ix, c, s, e);
        uint32_t a = s_adj[c];
        s += a;
        e += a;
        if (KIND == FMX_KIND_MULTI_ && c == 0) {
            s = multi_zero_rule(ix, s0, s);
            e = multi_zero_rule(ix, e0, e);
        }
    }
}

// get_l + lf_map fused: returns lf_map(i), sym = get_l(i)
template <int KIND>
__device__ __forceinline__ uint32_t lf_step(const FmxDev &ix, const uint32_t *s_adj, uint32_t i, uint32_t &sym) {
    if (KIND == FMX_KIND_RLFM_) {
        // rlfmi.rs:127-133: c = get_l(i); j = rank1(b,i); nr = rank(s,j,c); select1(bp, cs[c]+nr) + i - select1(b, j)
        uint32_t blk, r;
        rb_split(i, blk, r);
        RB bb = rb_load(ix.rl_b, blk);
        uint32_t j = rb_rank(bb, r);
        uint32_t h = rb_bit(bb, r) ? j : j - 1u;
        uint32_t c;
        uint32_t q = wm_access_walk(ix, h, c);  // q = walk_c(h); rank(s, j, c) = rank(s, h, c) + (j > h)
        uint32_t nr = q + (j - h);
        sym = c;
        return __ldg(ix.rl_bpsel + (s_adj[c] + nr)) + i - __ldg(ix.rl_bsel + j);
    } else {
        uint32_t c;
        uint32_t w = wm_access_walk(ix, i, c);
        sym = c;
        w += s_adj[c];
        if (KIND == FMX_KIND_MULTI_ && c == 0) return multi_zero_rule(ix, i, w);
        return w;
    }
}

template <int KIND>
__device__ __forceinline__ uint32_t get_l_dev(const FmxDev &ix, uint32_t i) {
    uint32_t c;
    if (KIND == FMX_KIND_RLFM_) {
        uint32_t h = rbv_rank1(ix.rl_b, i + 1 > ix.n ? ix.n : i + 1) - 1u;  // rank1 clamps past the end
        wm_access_walk(ix, h, c);
    } else {
        wm_access_walk(ix, i, c);
    }
    return c;
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
template <int KIND>
__device__ __forceinline__ bool fl_step(const FmxDev &ix, const uint32_t *s_adj, const uint32_t *s_cs, uint32_t i,
                                        uint32_t &sym, uint32_t &next) {
    if (KIND == FMX_KIND_RLFM_) {
        uint32_t jr = rbv_rank1(ix.rl_bp, i + 1) - 1u;
        uint32_t c = cs_search(s_cs, ix.cs_len, jr);
        uint32_t p = __ldg(ix.rl_bpsel + jr);
        uint32_t base = s_cs[c] - s_adj[c];
        uint32_t m = wm_select(ix, c, jr - s_cs[c], base);
        sym = c;
        next = __ldg(ix.rl_bsel + m) + i - p;
        return true;
    } else {
        uint32_t c = cs_search(s_cs, ix.cs_len, i);
        sym = c;
        if (KIND == FMX_KIND_MULTI_ && c == 0) return false;
        uint32_t base = s_cs[c] - s_adj[c];
        next = wm_select(ix, c, i - s_cs[c], base);
        return true;
    }
}

__device__ __forceinline__ void load_tables(const FmxDev &ix, uint32_t *s_adj, uint32_t *s_cs) {
    for (uint32_t k = threadIdx.x; k < ix.cs_len; k += blockDim.x) s_adj[k] = __ldg(ix.adj + k);
    if (s_cs)
        for (uint32_t k = threadIdx.x; k <= ix.cs_len; k += blockDim.x) s_cs[k] = __ldg(ix.cs + k);
    __syncthreads();
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
};

// Backward search (wrapper.rs:103-124), one pattern per thread.
template <int KIND>
__global__ void __launch_bounds__(256) k_search(const __grid_constant__ FmxDev ix, const __grid_constant__ SearchArgs a) {
    __shared__ uint32_t s_adj[256];
    load_tables(ix, s_adj, nullptr);
    unsigned long long steps = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.npat; p += stride) {
        uint64_t beg, len;
        if (a.pat_off) {
            beg = a.pat_off[p];
            len = a.pat_off[p + 1] - beg;
        } else {
            beg = p * a.fixed_len;
            len = a.fixed_len;
        }
        uint32_t s = a.init_s ? (uint32_t)a.init_s[p] : a.s0;
        uint32_t e = a.init_e ? (uint32_t)a.init_e[p] : a.e0;
        const uint8_t *q = a.pat + beg;
        for (uint64_t k = len; k-- > 0;) {
            uint32_t c = __ldg(q + k);
            if (c > ix.max_character) {  // the reference panics here (cs[c] out of bounds, fm_index.rs:94)
                atomicOr(a.err, 1u);
                break;
            }
            lf_map2_pair<KIND>(ix, s_adj, c, s, e);
            steps++;
            if (s == e) break;
        }
        a.out_s[p] = s;
        a.out_e[p] = e;
    }
    if (a.work) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
        if ((threadIdx.x & 31) == 0 && steps) atomicAdd(a.work, steps);
    }
}

// hit counts per pattern (wrapper.rs:132-139, 203-217): e - s, unfiltered (the L == 0 filter is
// applied on the expanded candidate rows)
__global__ void k_range_counts(const uint64_t *s, const uint64_t *e, uint64_t npat, uint64_t *cnt) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npat) cnt[p] = e[p] > s[p] ? e[p] - s[p] : 0;
}

// owner[off[p]] = p + 1 for every pattern with at least one candidate row
__global__ void k_mark_owners(const uint64_t *off, uint64_t npat, uint32_t *owner) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npat && off[p + 1] > off[p]) owner[off[p]] = (uint32_t)p + 1u;
}

// rows[h] = s[p] + (h - off[p]) with p = owner[h] - 1 (after the max-scan of owner)
__global__ void k_expand_rows(const uint64_t *s, const uint64_t *off, const uint32_t *owner, uint64_t total,
                              uint32_t *rows) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h < total) {
        uint32_t p = owner[h] - 1u;
        rows[h] = (uint32_t)(s[p] + (h - off[p]));
    }
}

// prefix filter (wrapper.rs:208): flag[h] = (get_l(rows[h]) == 0)
template <int KIND>
__global__ void __launch_bounds__(256) k_flag_prefix(const __grid_constant__ FmxDev ix, const uint32_t *rows,
                                                     uint64_t total, uint32_t *flag) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h < total) flag[h] = get_l_dev<KIND>(ix, rows[h]) == 0u ? 1u : 0u;
}

// stream compaction of the kept rows; fpos = exclusive scan of flag
__global__ void k_compact_rows(const uint32_t *rows, const uint32_t *flag, const uint64_t *fpos, uint64_t total,
                               uint32_t *out_rows) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h < total && flag[h]) out_rows[fpos[h]] = rows[h];
}

// filtered hit offsets: hit_off[p] = fpos[off[p]]  (fpos has total+1 entries)
__global__ void k_filtered_offsets(const uint64_t *off, const uint64_t *fpos, uint64_t npat, uint64_t *hit_off) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p <= npat) hit_off[p] = fpos[off[p]];
}

struct LocateArgs {
    const uint32_t *rows;
    uint64_t total;
    uint64_t *positions;  // nullable
    uint64_t *piece_ids;  // nullable (MultiPieces)
    unsigned long long *work;  // [1] += executed LF steps
};

// get_sa (fm_index.rs:127-140 / rlfmi.rs:176-189 / multi_pieces.rs:188-201): LF-walk each row to a
// sampled row, one hit per thread.  piece id (multi_pieces.rs:208-218) is derived from the located
// position and the piece boundary table: it equals the number of \0 before the position, which is
// what the reference's walk to the piece start computes (pinned by multi_pieces.rs:287-296).
template <int KIND>
__global__ void __launch_bounds__(256) k_locate(const __grid_constant__ FmxDev ix, const __grid_constant__ LocateArgs a) {
    __shared__ uint32_t s_adj[256];
    load_tables(ix, s_adj, nullptr);
    unsigned long long steps = 0;
    const uint32_t mask = (1u << ix.sa_level) - 1u;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < a.total; h += stride) {
        uint32_t row = a.rows[h];
        uint32_t st = 0, sym;
        while (row & mask) {
            row = lf_step<KIND>(ix, s_adj, row, sym);
            st++;
        }
        uint64_t v = (uint64_t)__ldg(ix.sa + (row >> ix.sa_level)) + st;
        if (v >= ix.n) v -= ix.n;  // (sa + steps) % n; both terms are < n
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
    }
    if (a.work) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
        if ((threadIdx.x & 31) == 0 && steps) atomicAdd(a.work + 1, steps);
    }
}

// iter_chars_backward / iter_chars_forward, k characters per row (wrapper.rs:143-183)
template <int KIND>
__global__ void __launch_bounds__(256) k_extract(const __grid_constant__ FmxDev ix, const uint64_t *rows,
                                                 uint64_t nrows, uint32_t k, int forward, uint8_t *out,
                                                 uint32_t *out_len) {
    __shared__ uint32_t s_adj[256];
    __shared__ uint32_t s_cs[257];
    load_tables(ix, s_adj, s_cs);
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    uint32_t i = (uint32_t)rows[r], got = 0;
    uint8_t *o = out + r * k;
    for (uint32_t t = 0; t < k; t++) {
        uint32_t c, nx;
        if (!forward) {
            nx = lf_step<KIND>(ix, s_adj, i, c);
        } else if (!fl_step<KIND>(ix, s_adj, s_cs, i, c, nx)) {
            break;
        }
        o[t] = (uint8_t)c;
        i = nx;
        got++;
    }
    for (uint32_t t = got; t < k; t++) o[t] = 0;
    if (out_len) out_len[r] = got;
}

// backend primitives over a batch of rows, for parity tests (src/backend.rs:5-40)
template <int KIND>
__global__ void __launch_bounds__(256) k_rows_op(const __grid_constant__ FmxDev ix, int op, const uint64_t *rows,
                                                 uint64_t nrows, uint64_t *out) {
    __shared__ uint32_t s_adj[256];
    __shared__ uint32_t s_cs[257];
    load_tables(ix, s_adj, s_cs);
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    uint32_t i = (uint32_t)rows[r], c, nx;
    uint64_t res = 0;
    switch (op) {
        case 0: res = get_l_dev<KIND>(ix, i); break;
        case 1: res = lf_step<KIND>(ix, s_adj, i, c); break;
        case 2:
            if (KIND == FMX_KIND_RLFM_) res = cs_search(s_cs, ix.cs_len, rbv_rank1(ix.rl_bp, i + 1) - 1u);
            else res = cs_search(s_cs, ix.cs_len, i);
            break;
        case 3: res = fl_step<KIND>(ix, s_adj, s_cs, i, c, nx) ? (uint64_t)nx : FMX_NONE; break;
        case 4: {
            const uint32_t mask = (1u << ix.sa_level) - 1u;
            uint32_t st = 0;
            while (i & mask) {
                i = lf_step<KIND>(ix, s_adj, i, c);
                st++;
            }
            uint64_t v = (uint64_t)__ldg(ix.sa + (i >> ix.sa_level)) + st;
            res = v >= ix.n ? v - ix.n : v;
            break;
        }
        case 5: {  // the reference's literal piece_id walk (multi_pieces.rs:208-218)
            for (;;) {
                uint32_t nxt = lf_step<KIND>(ix, s_adj, i, c);
                if (c == 0) {
                    uint32_t rank0 = wm_walk(ix, 0, i) + s_adj[0];  // rank(bw, i, 0); cs[0] == 0
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

template <int KIND>
__global__ void __launch_bounds__(256) k_lf_map2(const __grid_constant__ FmxDev ix, const uint8_t *c, const uint64_t *i,
                                                 uint64_t nrows, uint64_t *out) {
    __shared__ uint32_t s_adj[256];
    load_tables(ix, s_adj, nullptr);
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    out[r] = lf_map2_dev<KIND>(ix, s_adj, c[r], (uint32_t)i[r]);
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

// phase 1: per-tile reduction
template <class Tin, class Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const Tin *in, uint64_t n, uint64_t *tile_sum, Op op) {
    __shared__ uint64_t s_warp[SCAN_THREADS / 32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t acc = op.identity();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) acc = op(acc, (uint64_t)in[base + k]);
    uint64_t total;
    block_scan_excl(acc, op, total, s_warp);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// phase 3: per-tile scan with the carry of the preceding tiles.  EXCL: exclusive (out[i] excludes in[i]).
// When `out_total` is set and EXCL, out[n] receives the grand total (out has n+1 entries).
template <class Tin, class Tout, class Op, bool EXCL>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const Tin *in, uint64_t n, const uint64_t *tile_prefix,
                                                             Tout *out, Op op, int write_total) {
    __shared__ uint64_t s_warp[SCAN_THREADS / 32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS];
    uint64_t acc = op.identity();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? (uint64_t)in[base + k] : op.identity();
        acc = op(acc, v[k]);
    }
    uint64_t total;
    uint64_t pre = block_scan_excl(acc, op, total, s_warp);
    uint64_t carry = op(tile_prefix ? tile_prefix[blockIdx.x] : op.identity(), pre);
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
        uint64_t unit = z % nunits;
#pragma unroll
        for (int q = 0; q < WIDTH / 32; q++) {
            RB b = rb_load(buf, (uint32_t)(unit * (WIDTH / 32) + q));
            acc += b.w[0] ^ b.w[7];
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

}  // namespace fmx
