// phased.cuh -- the HBM-rich query pipeline: backward search split into converged phases, and locate by the
// resident suffix array.
//
// k_search (kernels.cuh) runs the whole reference loop (wrapper.rs:103-124) of one pattern in one thread.  On
// indexes that carry the dense seed-and-verify structures (fmx_layout.h: text, full suffix array, inverse) the
// work of a pattern is three very different things -- one table lookup, a few rank steps, one verify tail --
// and the patterns of a warp disagree about which of them they are in: ncu showed 6-10 of 32 lanes active
// (profiles/r01b_*_ncu.txt).  Here every phase is its own kernel over a compacted queue, so a warp's lanes do
// the same thing at the same time and every memory instruction serves 32 patterns:
//
//   k_ph_seed    one pattern per thread: the k-mer table lookup (ONE request).  Finished patterns (range
//                emptied inside the table, or nothing left) are written out; one-row ranges with enough
//                characters left go to the VERIFY queue; the rest to the STEPS queue.
//   k_ph_steps   ordinary reference iterations (lf_map2 on both range ends) until the range empties, the pattern
//                ends, or the range is one row with enough characters left (-> VERIFY queue).
//   k_ph_verify  position of the row (from the table entry, else SA[s]), comparison of the remaining characters
//                with the text, and -- only when the caller wants SA rows -- ISA[pos - matched].
//   k_ph_steps   again, over the patterns whose comparison stopped early and whose exact (s, e) are wanted
//                (second = 1: the verify tail is not tried again, as in search_one).
//
// Results are those of search_one, bit for bit: (s, e) when rows are wanted, e - s otherwise, plus a HINT: the
// text position of the final row when the verify phase produced it, which is exactly what locate returns for
// that row, so locate of a verified one-row range costs no memory request at all.
//
// Locate in this mode (k_emit_small / k_emit_big) reads positions from the resident suffix array: the
// reference's walk (fm_index.rs:127-140) ends in (sa[row'] + steps) % n = SA[row] by construction.
#pragma once
#include "kernels.cuh"

namespace fmx {

struct PhasedArgs {
    SearchArgs a;             // patterns, (s0, e0), tables, err, work
    uint32_t *rs, *re;        // [npat] final range; with want_rows == 0 only re - rs is meaningful for verified patterns
    uint32_t *hint;           // [npat] text position of row rs when the verify phase knows it, else FMX_NOHINT
    uint32_t want_rows;       // 1: exact (rs, re) for every pattern (one more request per verified pattern)
    uint4 *q_steps;           // STEPS queue: {pattern, characters left, s, e}
    uint4 *q_verify;          // VERIFY queue: {pattern, characters left, s, position of row s or FMX_NOHINT}
    unsigned long long *qn;   // [0] STEPS entries, [1] VERIFY entries, [2] STEPS entries of the second pass
};

// order-preserving append of the flagged lanes' entries (one atomic per warp); all 32 lanes must call it
__device__ __forceinline__ void warp_push(bool want, const uint4 &ent, uint4 *queue, unsigned long long *counter) {
    const uint32_t lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) queue[base + __popc(m & ((1u << lane) - 1u))] = ent;
}

// rank blocks lf_map2_pair loads for the range (s, e): the two ends share the block once the range is narrow
template <int LAYOUT>
__device__ __forceinline__ uint32_t pair_requests(const FmxDev &ix, uint32_t s, uint32_t e) {
    if (LAYOUT == FMX_LAYOUT_Q4) return (s >> 6) == (e >> 6) ? 1u : 2u;
    if (LAYOUT == FMX_LAYOUT_SY) return s / FMX_RB_BITS == e / FMX_RB_BITS ? 1u : 2u;
    if (LAYOUT == FMX_LAYOUT_W4) return ix.qlevels * ((s >> 6) == (e >> 6) ? 1u : 2u);
    return ix.levels * (s / FMX_RB_BITS == e / FMX_RB_BITS ? 1u : 2u);
}

__device__ __forceinline__ void warp_add(unsigned long long v, unsigned long long *counter) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(counter, v);
}

template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_ph_seed(const __grid_constant__ FmxDev ix, const __grid_constant__ PhasedArgs g) {
    const SearchArgs &a = g.a;
    unsigned long long steps = 0, reqs = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (a.npat + stride - 1) / stride;
    for (uint64_t r = 0; r < rounds; r++) {
        const uint64_t p = r * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int dest = 0;  // 1 = STEPS, 2 = VERIFY
        uint4 ent = make_uint4(0, 0, 0, 0);
        if (p < a.npat) {
            uint64_t beg;
            uint32_t len;
            pattern_span(a, p, beg, len);
            uint32_t s = a.s0, e = a.e0, it = 0, pos = FMX_NOHINT, rem = len;
            AnyReader rd(a, p, beg, len);
            if (kmer_lookup(a, ix.max_character, rd, rem, s, e, it, &pos)) reqs += 1u;  // the entry (the step byte is read by counting passes only)
            steps += it;
            if (rem == 0 || s == e) {
                g.rs[p] = s;
                g.re[p] = e;
                g.hint[p] = (e - s == 1u) ? pos : FMX_NOHINT;
            } else if (a.verify && e - s == 1u && rem >= FMX_VERIFY_MIN_DENSE) {
                dest = 2;
                ent = make_uint4((uint32_t)p, rem, s, pos);
            } else {
                dest = 1;
                ent = make_uint4((uint32_t)p, rem, s, e);
            }
        }
        warp_push(dest == 1, ent, g.q_steps, g.qn + 0);
        warp_push(dest == 2, ent, g.q_verify, g.qn + 1);
    }
    if (a.work) {
        warp_add(steps, a.work);
        warp_add(reqs, a.work + 2);
    }
}

template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_ph_steps(const __grid_constant__ FmxDev ix, const __grid_constant__ PhasedArgs g,
                                                  int second) {
    __shared__ Tabs<LAYOUT> tb;
    load_tables<LAYOUT>(ix, tb);
    const SearchArgs &a = g.a;
    const unsigned long long qn = g.qn[second ? 2 : 0];
    unsigned long long steps = 0, reqs = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (qn + stride - 1) / stride;
    const bool armed = !second && a.verify;
    for (uint64_t r = 0; r < rounds; r++) {
        const uint64_t t = r * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        bool to_verify = false;
        uint4 ent = make_uint4(0, 0, 0, 0);
        if (t < qn) {
            ent = g.q_steps[t];
            const uint64_t p = ent.x;
            uint32_t rem = ent.y, s = ent.z, e = ent.w;
            uint64_t beg;
            uint32_t len;
            pattern_span(a, p, beg, len);
            AnyReader rd(a, p, beg, len);
            while (rem > 0) {
                if (armed && e - s == 1u && rem >= FMX_VERIFY_MIN_DENSE) {
                    to_verify = true;
                    break;
                }
                const uint32_t c = rd.get(rem - 1u);
                if (c > ix.max_character) {  // the reference panics here (cs[c] out of bounds, fm_index.rs:94)
                    atomicOr(a.err, 1u);
                    break;
                }
                if (a.work) reqs += pair_requests<LAYOUT>(ix, s, e);
                lf_map2_pair<KIND, LAYOUT>(ix, tb, c, s, e);
                steps++;
                rem--;
                if (s == e) break;
            }
            if (to_verify) {
                ent = make_uint4((uint32_t)p, rem, s, FMX_NOHINT);
            } else {
                g.rs[p] = s;
                g.re[p] = e;
                g.hint[p] = FMX_NOHINT;
            }
        }
        warp_push(to_verify, ent, g.q_verify, g.qn + 1);
    }
    if (a.work) {
        warp_add(steps, a.work);
        warp_add(reqs, a.work + 2);
    }
}

// verify phase.  Entry: the range is the single row s, `rem` characters rd[0 .. rem) are still to be consumed.
// What search_one's verify_tail + following iterations do, case by case (kernels.cuh):
//   all rem characters match            -> row ISA[pos - rem], range of one row, pattern consumed
//   `matched` < rem match, then a true mismatch with a character c != 0, c <= max_character
//                                       -> the next ordinary iteration empties the range (BWT of the row != c)
//   the comparison stops for another reason (text \0, pattern \0 or invalid character, start of the text)
//                                       -> ordinary iterations from row ISA[pos - matched], tail not tried again
template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) k_ph_verify(const __grid_constant__ FmxDev ix, const __grid_constant__ PhasedArgs g) {
    const SearchArgs &a = g.a;
    const unsigned long long qn = g.qn[1];
    unsigned long long steps = 0, reqs = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (qn + stride - 1) / stride;
    for (uint64_t r = 0; r < rounds; r++) {
        const uint64_t t = r * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        bool again = false;
        uint4 ent = make_uint4(0, 0, 0, 0);
        if (t < qn) {
            ent = g.q_verify[t];
            const uint64_t p = ent.x;
            const uint32_t rem = ent.y, s = ent.z;
            uint32_t pos = ent.w;
            if (pos == FMX_NOHINT) {
                pos = ldg32_s(ix.vsa + s);
                reqs++;
            }
            uint32_t matched = 0, c = 0;
            if (pos >= rem) {
                uint64_t beg;
                uint32_t len;
                pattern_span(a, p, beg, len);
                AnyReader rd(a, p, beg, len);
                TextReader tr(ix.text + (pos - rem), rem);
                matched = match_backward(rd, tr, rem, c);
            }
            steps += matched;
            if (pos >= rem) {  // 32-byte sectors of the text the comparison read
                const uint32_t last = pos - 1u, first = pos - (matched < rem ? matched + 1u : rem);
                reqs += (last >> 5) - (first >> 5) + 1u;
            }
            if (pos >= rem && matched == rem) {
                const uint32_t q = pos - rem;
                g.hint[p] = q;
                uint32_t row = 0;
                if (g.want_rows) {
                    row = ldg32_s(ix.isa + q);
                    reqs++;
                }
                g.rs[p] = row;
                g.re[p] = row + 1u;
            } else if (pos >= rem && !g.want_rows && c != 0u && c <= ix.max_character) {
                // true mismatch: the reference's next iteration leaves s == e; count 0, rows not wanted
                steps++;
                g.rs[p] = 0;
                g.re[p] = 0;
                g.hint[p] = FMX_NOHINT;
            } else {
                const uint32_t row = matched ? ldg32_s(ix.isa + (pos - matched)) : s;
                reqs += matched ? 1u : 0u;
                again = true;
                ent = make_uint4((uint32_t)p, rem - matched, row, row + 1u);
            }
        }
        warp_push(again, ent, g.q_steps, g.qn + 2);
    }
    if (a.work) {
        warp_add(steps, a.work);
        warp_add(reqs, a.work + 2);
    }
}

// ---- the same pipeline as ONE kernel (option "search_phased" = 2, the default).
// Measured on the 1 GB DNA target (profiles/r02_*): the three-kernel form keeps every warp converged but pays for it --
// every pattern that reaches the verify phase is read from HBM a second time (one more DRAM request, of the three it
// makes in total), and the two queue appends of every warp are atomics on two hot addresses.  What bounds this
// workload is the number of DRAM requests, which idle lanes do not issue, so here one thread takes one pattern
// through table lookup, the few ordinary iterations and the text comparison without leaving the registers its
// characters already sit in.  Same results, same hints, same step counts.
// HEAD of a query: the table lookup and -- when the entry is a one-row range that carries enough of the text in front of
// its row (SearchArgs::big_tab4) -- the comparison of the rest of the pattern against that context.  Returns true when
// (s, e, hint) are final; otherwise (k, s, e, pos, armed) is the state query_tail continues from.
template <int KIND, int LAYOUT, class Reader>
__device__ __forceinline__ bool query_head(const FmxDev &ix, const PhasedArgs &g, Reader &rd, uint32_t len, uint32_t &k, uint32_t &s,
                                           uint32_t &e, uint32_t &pos, bool &armed, uint32_t &hint, unsigned long long &steps,
                                           unsigned long long &reqs) {
    const SearchArgs &a = g.a;
    uint32_t it = 0, ctx = 0, ctx_len = 0;
    pos = FMX_NOHINT;
    k = len;
    // one request: the entry.  (A pass that counts work also reads the step byte of an emptied range; the timed pass
    // does not, and the request counter reports what the timed pass issues.)
    if (kmer_lookup(a, ix.max_character, rd, k, s, e, it, &pos, &ctx, &ctx_len)) reqs += 1u;
    armed = a.verify != 0;
    hint = FMX_NOHINT;
    steps += it;
    if (k == 0 || s == e) {  // the table finished the pattern (the reference loop tests the range after a step, never before)
        if (k == 0 && e - s == 1u) hint = pos;  // a one-row entry that consumed the whole pattern
        return it != 0 || k == 0;  // nothing consumed and characters left: an empty INITIAL range still runs the loop
    }
    uint32_t pk;
    if (armed && pos != FMX_NOHINT && e - s == 1u && k <= ctx_len && kmer_index4(rd, k, k, pk)) {
        // the rest of the pattern (as 2-bit codes, last character lowest -- exactly its k-mer index) against the
        // characters in front of the row: no text request
        const uint32_t diff = (pk ^ ctx) & (k < 16u ? (1u << (2u * k)) - 1u : 0xFFFFFFFFu);
        const uint32_t matched = diff ? (uint32_t)(__ffs((int)((diff | (diff >> 1)) & 0x55555555u)) - 1) >> 1 : k;
        steps += matched;
        if (matched == k) {  // the whole rest of the pattern stands in the text in front of the row
            hint = pos - k;
            if (g.want_rows) {
                s = ldg32_s(ix.isa + hint);
                reqs++;
            } else {
                s = 0;
            }
            e = s + 1u;
            k = 0;
            return true;
        }
        if (!g.want_rows) {  // a true mismatch (the characters are valid): the next iteration empties the range
            steps++;
            s = e = 0;
            return true;
        }
        if (matched) {  // exact rows wanted: go on from the row the reference reaches after the characters that match
            s = ldg32_s(ix.isa + (pos - matched));
            reqs++;
            e = s + 1u;
            k -= matched;
        }
        armed = false;  // as search_one: the tail is not tried again once an attempt stopped early
    }
    return false;
}

// TAIL of a query: the reference loop (wrapper.rs:103-124) from the state the head left, with the seed-and-verify
// tail against the text for one-row ranges.
template <int KIND, int LAYOUT, class Reader>
__device__ __forceinline__ void query_tail(const FmxDev &ix, const Tabs<LAYOUT> &tb, const PhasedArgs &g, Reader &rd, uint32_t k,
                                           uint32_t &s, uint32_t &e, uint32_t pos, bool armed, uint32_t &hint,
                                           unsigned long long &steps, unsigned long long &reqs) {
    const SearchArgs &a = g.a;
    uint32_t it = 0;
    while (k > 0) {  // the range is tested after a step, never before
        if (armed && e - s == 1u && k >= FMX_VERIFY_MIN_DENSE) {
            if (pos == FMX_NOHINT) {
                pos = ldg32_s(ix.vsa + s);
                reqs++;
            }
            uint32_t matched = 0, c = 0;
            if (pos >= k) {
                TextReader tr(ix.text + (pos - k), k);
                matched = match_backward(rd, tr, k, c);
                const uint32_t last = pos - 1u, first = pos - (matched < k ? matched + 1u : k);
                reqs += (last >> 5) - (first >> 5) + 1u;
            }
            it += matched;
            if (pos >= k && matched == k) {  // the whole rest of the pattern stands in the text in front of the row
                hint = pos - k;
                if (g.want_rows) {
                    s = ldg32_s(ix.isa + hint);
                    reqs++;
                } else {
                    s = 0;
                }
                e = s + 1u;
                k = 0;
                break;
            }
            if (pos >= k && !g.want_rows && c != 0u && c <= ix.max_character) {  // true mismatch: the next iteration empties the range
                it++;
                s = e = 0;
                break;
            }
            if (matched) {
                s = ldg32_s(ix.isa + (pos - matched));
                reqs++;
                e = s + 1u;
                k -= matched;
            }
            armed = false;  // as search_one: the tail is not tried again once an attempt stopped early
            continue;
        }
        const uint32_t c = rd.get(k - 1u);
        if (c > ix.max_character) {  // the reference panics here (cs[c] out of bounds, fm_index.rs:94)
            atomicOr(a.err, 1u);
            break;
        }
        if (a.work) reqs += pair_requests<LAYOUT>(ix, s, e);
        lf_map2_pair<KIND, LAYOUT>(ix, tb, c, s, e);
        pos = FMX_NOHINT;
        it++;
        k--;
        if (s == e) break;
    }
    if (hint == FMX_NOHINT && e - s == 1u && k == 0) hint = pos;  // a one-row range whose position is known
    steps += it;
}

template <int KIND, int LAYOUT, class Reader>
__device__ __forceinline__ void query_one(const FmxDev &ix, const Tabs<LAYOUT> &tb, const PhasedArgs &g, Reader &rd, uint32_t len,
                                          uint32_t &s, uint32_t &e, uint32_t &hint, unsigned long long &steps,
                                          unsigned long long &reqs) {
    uint32_t k, pos;
    bool armed;
    if (!query_head<KIND, LAYOUT>(ix, g, rd, len, k, s, e, pos, armed, hint, steps, reqs))
        query_tail<KIND, LAYOUT>(ix, tb, g, rd, k, s, e, pos, armed, hint, steps, reqs);
}

// five resident blocks per SM (<= 48 registers): this kernel lives on memory requests in flight; at 56 registers (four
// blocks) the MultiPieces instance ran 14 % slower (config 4: 1.25 -> 1.43 ms)
template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256, 5) k_query_fused(const __grid_constant__ FmxDev ix, const __grid_constant__ PhasedArgs g) {
    __shared__ Tabs<LAYOUT> tb;
    __shared__ uint32_t spat[8][256];
    load_tables<LAYOUT>(ix, tb);
    const SearchArgs &a = g.a;
    unsigned long long steps = 0, reqs = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.npat; t += stride) {
        const uint64_t p = a.order ? (uint64_t)a.order[t] : t;  // ragged batches: in order of length (k_len_scatter)
        uint64_t beg;
        uint32_t len;
        pattern_span(a, p, beg, len);
        uint32_t s = a.s0, e = a.e0, hint;
        if (a.packed_bits) {
            AnyReader rd(a, p, beg, len);
            query_one<KIND, LAYOUT>(ix, tb, g, rd, len, s, e, hint, steps, reqs);
        } else if (a.staged) {
            StagedReader rd(spat, a.pat + beg, len);
            query_one<KIND, LAYOUT>(ix, tb, g, rd, len, s, e, hint, steps, reqs);
        } else {
            PatReader rd(a.pat + beg, len);
            query_one<KIND, LAYOUT>(ix, tb, g, rd, len, s, e, hint, steps, reqs);
        }
        g.rs[p] = s;
        g.re[p] = e;
        g.hint[p] = hint;
    }
    if (a.work) {
        warp_add(steps, a.work);
        warp_add(reqs, a.work + 2);
    }
}

// ---- the fused kernel with a block-local second pass (option "fused_defer").
// With the 16-byte table entries nine patterns in ten are finished by query_head (one request, ~300 instructions, all 32
// lanes of the warp in step); the tenth -- a k-mer that occurs more than once, a pattern longer than the context --
// needs rank steps, SA[row] and a text comparison: ~1100 instructions that the plain fused kernel executes with ~3 of
// 32 lanes active, which is 80 % of its issue slots (profiles/r02_c10_target_k_query_fused_ctx_ncu.txt: 9.4 lanes,
// issue-active 64 %).  Here a thread whose pattern is not finished by the head parks its state (pattern, characters
// left, range, position: five words) in a shared-memory queue; whenever the block holds 256 parked states all of its
// threads take one each through query_tail, so the expensive path runs with full warps.  No global queue, no second
// kernel; the only extra traffic is re-reading the parked patterns' characters.
template <int KIND, int LAYOUT, uint32_t CAP>
__device__ __forceinline__ void defer_tail(const FmxDev &ix, const Tabs<LAYOUT> &tb, const PhasedArgs &g, uint32_t (*spat)[256],
                                           const uint32_t (*dq)[CAP], uint32_t slot, unsigned long long &steps,
                                           unsigned long long &reqs) {
    const SearchArgs &a = g.a;
    const uint64_t p = dq[0][slot];
    const uint32_t k = dq[1][slot] & 0x7FFFFFFFu;
    const bool armed = (dq[1][slot] >> 31) != 0;
    uint32_t s = dq[2][slot], e = dq[3][slot], hint = FMX_NOHINT;
    const uint32_t pos = dq[4][slot];
    uint64_t beg;
    uint32_t len;
    pattern_span(a, p, beg, len);
    if (a.packed_bits) {
        AnyReader rd(a, p, beg, len);
        query_tail<KIND, LAYOUT>(ix, tb, g, rd, k, s, e, pos, armed, hint, steps, reqs);
    } else if (a.staged) {
        StagedReader rd(spat, a.pat + beg, len);
        query_tail<KIND, LAYOUT>(ix, tb, g, rd, k, s, e, pos, armed, hint, steps, reqs);
    } else {
        PatReader rd(a.pat + beg, len);
        query_tail<KIND, LAYOUT>(ix, tb, g, rd, k, s, e, pos, armed, hint, steps, reqs);
    }
    g.rs[p] = s;
    g.re[p] = e;
    g.hint[p] = hint;
}

// Two knobs, both A/B'd on the target (option fused_defer = 3 .. 6, profiles/r02_c18_*):
//   MINB    resident blocks per SM.  The kernel lives on requests in flight: 5 blocks (<= 48 registers, no spills)
//           against 6 (<= 40 registers, 40-400 bytes of spills per thread).
//   ROUNDS  rounds of 256 patterns between two looks at the queue.  One pattern in ten is parked, so the queue needs
//           ~10 rounds to fill a drain of 256; looking after every round costs two block barriers per round, each
//           making every warp wait for the slowest table lookup of the block.  The queue holds 256 * (ROUNDS + 1)
//           states: fewer than 256 left over plus at most 256 per round.
//   grid    blocks per SM of this kernel's grid-stride loop (FMX_QUERY_BLOCKS_PER_SM x SMs at most; option query_blocks
//           overrides; k_query_fused keeps 64): many short-lived blocks against exactly the resident ones.
#define FMX_DEFER_BLOCKS 5
#define FMX_DEFER_ROUNDS 1
#define FMX_QUERY_BLOCKS_PER_SM 64
template <int KIND, int LAYOUT, int MINB, int ROUNDS>
__global__ void __launch_bounds__(256, MINB) k_query_fused_defer(const __grid_constant__ FmxDev ix, const __grid_constant__ PhasedArgs g) {
    constexpr uint32_t CAP = 256u * (ROUNDS + 1);
    __shared__ Tabs<LAYOUT> tb;
    __shared__ uint32_t spat[8][256];
    __shared__ uint32_t dq[5][CAP];
    __shared__ uint32_t dq_n;
    load_tables<LAYOUT>(ix, tb);
    if (threadIdx.x == 0) dq_n = 0;
    __syncthreads();
    const SearchArgs &a = g.a;
    unsigned long long steps = 0, reqs = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (a.npat + stride - 1) / stride;
    for (uint64_t r = 0; r < rounds; r++) {
        const uint64_t p = r * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (p < a.npat) {
            uint64_t beg;
            uint32_t len;
            pattern_span(a, p, beg, len);
            uint32_t s = a.s0, e = a.e0, hint, k, pos;
            bool armed, done;
            if (a.packed_bits) {
                AnyReader rd(a, p, beg, len);
                done = query_head<KIND, LAYOUT>(ix, g, rd, len, k, s, e, pos, armed, hint, steps, reqs);
            } else if (a.staged) {
                StagedReader rd(spat, a.pat + beg, len);
                done = query_head<KIND, LAYOUT>(ix, g, rd, len, k, s, e, pos, armed, hint, steps, reqs);
            } else {
                PatReader rd(a.pat + beg, len);
                done = query_head<KIND, LAYOUT>(ix, g, rd, len, k, s, e, pos, armed, hint, steps, reqs);
            }
            if (done) {
                g.rs[p] = s;
                g.re[p] = e;
                g.hint[p] = hint;
            } else {  // park the state
                const uint32_t slot = atomicAdd(&dq_n, 1u);
                dq[0][slot] = (uint32_t)p;
                dq[1][slot] = k | (armed ? 0x80000000u : 0u);
                dq[2][slot] = s;
                dq[3][slot] = e;
                dq[4][slot] = pos;
            }
        }
        if (ROUNDS > 1 && (r + 1) % ROUNDS != 0) continue;  // r is the same for every thread of the block
        __syncthreads();
        while (dq_n >= 256u) {  // the same value for every thread: read between two barriers
            const uint32_t base = dq_n - 256u;
            defer_tail<KIND, LAYOUT, CAP>(ix, tb, g, spat, dq, base + threadIdx.x, steps, reqs);
            __syncthreads();
            if (threadIdx.x == 0) dq_n = base;
            __syncthreads();
        }
    }
    __syncthreads();
    for (uint32_t slot = threadIdx.x; slot < dq_n; slot += 256u) defer_tail<KIND, LAYOUT, CAP>(ix, tb, g, spat, dq, slot, steps, reqs);
    if (a.work) {
        warp_add(steps, a.work);
        warp_add(reqs, a.work + 2);
    }
}

// (rs, re) as the u64 arrays of the C ABI
__global__ void k_ph_widen(const uint32_t *rs, const uint32_t *re, uint64_t npat, uint64_t *out_s, uint64_t *out_e) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npat) {
        out_s[p] = rs[p];
        out_e[p] = re[p];
    }
}

// scan input: matches per pattern (wrapper.rs:132-134)
struct LoadCount32 {
    const uint32_t *s, *e;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const {
        const uint32_t a = s[i], b = e[i];
        return b > a ? (uint64_t)(b - a) : 0ull;
    }
};

// ---- locate by the resident suffix array.  positions[off[p] + j] = SA[rs[p] + j], rows ascending (wrapper.rs:206-216)
template <class Toff, class Tout>
struct EmitArgs {
    const uint32_t *rs, *re, *hint;  // hint nullable
    const Toff *off;                 // npat + 1 hit offsets
    uint64_t npat;
    uint64_t capacity;               // entries of positions / piece_ids
    Tout *positions;                 // nullable
    Tout *piece_ids;                 // nullable (MultiPieces)
    uint32_t *bigq;                  // patterns with more than FMX_EMIT_SMALL matches
    unsigned long long *bign;
    unsigned long long *req;         // nullable: += suffix-array requests issued (hinted positions cost none)
};
#define FMX_EMIT_SMALL 4u

template <int KIND, class Toff, class Tout>
__global__ void __launch_bounds__(256) k_emit_small(const __grid_constant__ FmxDev ix, const __grid_constant__ EmitArgs<Toff, Tout> g) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (g.npat + stride - 1) / stride;
    const uint32_t lane = threadIdx.x & 31;
    unsigned long long reqs = 0;
    for (uint64_t r = 0; r < rounds; r++) {
        const uint64_t p = r * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        bool big = false;
        if (p < g.npat) {
            const uint32_t s = g.rs[p], e = g.re[p];
            const uint32_t cnt = e > s ? e - s : 0u;
            if (cnt > FMX_EMIT_SMALL) {
                big = true;
            } else if (cnt) {
                const uint64_t o = (uint64_t)g.off[p];
                const uint32_t h = (g.hint && cnt == 1u) ? g.hint[p] : FMX_NOHINT;
                for (uint32_t j = 0; j < cnt; j++) {
                    if (o + j >= g.capacity) break;
                    const uint32_t v = h != FMX_NOHINT ? h : ldg32_s(ix.vsa + s + j);
                    reqs += h != FMX_NOHINT ? 0u : 1u;
                    if (g.positions) g.positions[o + j] = (Tout)v;
                    if (KIND == FMX_KIND_MULTI_ && g.piece_ids) g.piece_ids[o + j] = (Tout)piece_of(ix, v);
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, big);
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(g.bign, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (big) g.bigq[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)p;
        }
    }
    if (g.req) warp_add(reqs, g.req);
}

#define FMX_EMIT_FUSED_DEFAULT 1
// ---- counts -> offsets -> positions in ONE pass over the ranges (option "emit_fused"): the last phase of the scan
// (k_scan_apply<LoadCount32>) and k_emit_small as one kernel.  The separate kernels read every range twice more and the
// offsets once more than needed (1.6 GB per 100 M patterns).  Here a warp owns 256 consecutive patterns and visits them
// 32 at a time, lane = pattern, so the loads of (rs, re, hint) and the stores of offsets and positions are the coalesced
// ones of k_emit_small; the match counts stay in registers between the two sweeps (warp total -> warp base -> offsets).
// tile_prefix / tile_sums: as in k_scan_apply.  Patterns with more than FMX_EMIT_SMALL matches go to k_emit_big's queue.
template <int KIND, class Tw>
__global__ void __launch_bounds__(SCAN_THREADS) k_offsets_emit(const __grid_constant__ FmxDev ix, const __grid_constant__ EmitArgs<Tw, Tw> g,
                                                              Tw *off_out, const uint64_t *tile_prefix, const uint64_t *tile_sums) {
    __shared__ uint64_t s_warp[SCAN_THREADS / 32];
    OpSum op;
    uint64_t tile_carry = 0;
    if (tile_prefix) {
        tile_carry = tile_prefix[blockIdx.x];
    } else if (tile_sums) {
        uint64_t part = 0;
        for (uint32_t i = threadIdx.x; i < blockIdx.x; i += SCAN_THREADS) part += tile_sums[i];
        block_scan_excl(part, op, tile_carry, s_warp);   // ends with a barrier: s_warp is free again
    }
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t wbase = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)wid * (32u * SCAN_ITEMS);
    uint32_t cnt[SCAN_ITEMS];
    uint64_t wtot = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        const uint64_t p = wbase + (uint64_t)j * 32u + lane;
        uint32_t c = 0;
        if (p < g.npat) {
            const uint32_t s = g.rs[p], e = g.re[p];
            c = e > s ? e - s : 0u;
        }
        cnt[j] = c;
        wtot += c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wtot += __shfl_xor_sync(0xffffffffu, wtot, o);
    if (lane == 0) s_warp[wid] = wtot;
    __syncthreads();
    uint64_t carry = tile_carry;
    for (uint32_t w = 0; w < wid; w++) carry += s_warp[w];
    unsigned long long reqs = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        const uint64_t p = wbase + (uint64_t)j * 32u + lane;
        const uint32_t c = cnt[j];
        uint64_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        const uint64_t o = carry + inc - c;
        carry += __shfl_sync(0xffffffffu, inc, 31);
        bool big = false;
        if (p < g.npat) {
            off_out[p] = (Tw)o;
            if (p == g.npat - 1) off_out[g.npat] = (Tw)(o + c);
            if (c > FMX_EMIT_SMALL) {
                big = true;
            } else if (c) {
                const uint32_t s = g.rs[p];
                const uint32_t h = (g.hint && c == 1u) ? g.hint[p] : FMX_NOHINT;
                for (uint32_t i = 0; i < c; i++) {
                    if (o + i >= g.capacity) break;
                    const uint32_t v = h != FMX_NOHINT ? h : ldg32_s(ix.vsa + s + i);
                    reqs += h != FMX_NOHINT ? 0u : 1u;
                    if (g.positions) g.positions[o + i] = (Tw)v;
                    if (KIND == FMX_KIND_MULTI_ && g.piece_ids) g.piece_ids[o + i] = (Tw)piece_of(ix, v);
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, big);
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(g.bign, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (big) g.bigq[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)p;
        }
    }
    if (g.req) warp_add(reqs, g.req);
}

// one warp per pattern with many matches: SA[s .. e) is a contiguous read, the output a contiguous write
template <int KIND, class Toff, class Tout>
__global__ void __launch_bounds__(256) k_emit_big(const __grid_constant__ FmxDev ix, const __grid_constant__ EmitArgs<Toff, Tout> g) {
    const unsigned long long qn = *g.bign;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < qn; w += nwarps) {
        const uint32_t p = g.bigq[w];
        const uint32_t s = g.rs[p], e = g.re[p];
        const uint64_t o = (uint64_t)g.off[p];
        for (uint32_t j = lane; j < e - s; j += 32) {
            if (o + j >= g.capacity) break;
            const uint32_t v = __ldg(ix.vsa + s + j);
            if (g.positions) g.positions[o + j] = (Tout)v;
            if (KIND == FMX_KIND_MULTI_ && g.piece_ids) g.piece_ids[o + j] = (Tout)piece_of(ix, v);
        }
    }
}

// the k-mer tables of an index with a resident suffix array: one-row entries carry the row's text position
__global__ void k_table_embed(uint2 *tab, uint64_t entries, const uint32_t *vsa) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= entries) return;
    const uint2 v = tab[t];
    if (v.y == v.x + 1u) tab[t] = make_uint2(v.x, FMX_TAB_POS_FLAG | ldg32_s(vsa + v.x));
}

// 16-byte entries from 8-byte ones (SearchArgs::big_tab4): a one-row entry also gets the up to 16 text characters in
// front of its row as 2-bit codes (alphabets of <= 4 symbols), stopping at the start of the text or in front of a \0
__global__ void k_table_widen(const uint2 *tab, uint64_t entries, const uint8_t *text, uint4 *tab4) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= entries) return;
    const uint2 v = tab[t];
    uint32_t ctx = 0, len = 0;
    if (v.y & FMX_TAB_POS_FLAG) {
        const uint32_t pos = v.y & ~FMX_TAB_POS_FLAG;
        while (len < 16u && len < pos) {
            const uint32_t ch = ldg8_s(text + (pos - 1u - len));
            if (ch == 0u || ch > 4u) break;
            ctx |= (ch - 1u) << (2u * len);
            len++;
        }
    }
    tab4[t] = make_uint4(v.x, v.y, ctx, len);
}

}  // namespace fmx
