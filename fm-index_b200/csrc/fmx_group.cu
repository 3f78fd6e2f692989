// fmx_group.cu -- several GPUs of ONE process behind one handle (include/fmx.h, "multi-GPU"), built on the public
// C ABI only (fmx_index_build_ex / fmx_index_clone / fmx_query_batch / fmx_query_batch_device / fmx_csr_merge_device).
//
//   FMX_GROUP_REPLICATE  queries are independent: the index is copied to every device, the batch is
//                        cut into contiguous shards, one host thread per device drives fmx_query_batch on its shard
//                        (its own PCIe link, its own copy/compute pipeline); the shard CSRs are stitched in input order.
//   FMX_GROUP_BY_PIECE   MultiPieces partitioned by piece (multi_pieces.rs:188-223 over shards): every device indexes
//                        a contiguous, length-balanced group of pieces and answers the WHOLE batch; the parts are
//                        gathered on the first device with peer copies (NVLink) and merged there by one scan + one
//                        scatter kernel (fmx_csr_merge_device).  One process per GPU does the same exchange with NCCL
//                        (fm-index_b200/partitioned.py).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fmx.h"

namespace {

struct Grow {  // grow-only buffer, device or pinned host
    void *p = nullptr;
    size_t cap = 0;
    bool host = false;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        release();
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = host ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return FMX_ERR_OOM;
        }
        cap = want;
        return 0;
    }
    void release() {
        if (p) {
            if (host) cudaFreeHost(p); else cudaFree(p);
        }
        p = nullptr;
        cap = 0;
    }
};

struct Member {
    int device = 0;
    fmx_index *idx = nullptr;
    uint64_t pos_base = 0, pid_base = 0;  // BY_PIECE: text offset and first piece of the partition
    cudaStream_t st = nullptr;
    Grow d_pat, d_poff, d_hoff, d_pos, d_pid;       // BY_PIECE: the batch and the part on the member's device
    Grow s_hoff, s_pos, s_pid;                      // BY_PIECE: the part staged on the first device
    Grow h_off, h_pos, h_pid;                       // REPLICATE: the shard's CSR on the host (pinned)
    uint64_t *h_total = nullptr;                    // pinned
    Member() {
        h_off.host = h_pos.host = h_pid.host = true;
    }
};

}  // namespace

struct fmx_group {
    int mode = FMX_GROUP_REPLICATE;
    uint64_t n_total = 0, pieces_total = 0;
    std::vector<Member> m;
    Grow o_hoff, o_pos, o_pid;  // BY_PIECE: the merged CSR on the first device
    std::mutex mu;
};

// fmx_last_error() is per thread and owned by fmx_api.cu: the group's own errors are stored there too
extern "C" void fmx_internal_set_error(const char *msg);
static int gfail(int code, const char *msg) {
    fmx_internal_set_error(msg);
    return code;
}

// contiguous groups of pieces balanced by symbol count; every group keeps at least one piece
static std::vector<std::pair<size_t, size_t>> partition_pieces(const std::vector<uint64_t> &len, int world) {
    const size_t d = len.size();
    std::vector<uint64_t> cs(d + 1, 0);
    for (size_t k = 0; k < d; k++) cs[k + 1] = cs[k] + len[k];
    std::vector<size_t> cuts{0};
    for (int r = 1; r < world; r++) {
        const double target = (double)cs[d] * r / world;
        size_t k = (size_t)(std::lower_bound(cs.begin(), cs.end(), (uint64_t)target) - cs.begin());
        k = std::min(std::max(k, cuts.back() + 1), d - (size_t)(world - r));
        if (k - 1 > cuts.back() && std::abs((double)cs[k - 1] - target) < std::abs((double)cs[k] - target)) k--;
        cuts.push_back(k);
    }
    cuts.push_back(d);
    std::vector<std::pair<size_t, size_t>> out;
    for (int r = 0; r < world; r++) out.push_back({cuts[r], cuts[r + 1]});
    return out;
}

extern "C" {

void fmx_group_free(fmx_group *g) {
    if (!g) return;
    for (auto &mb : g->m) {
        cudaSetDevice(mb.device);
        if (mb.idx) fmx_index_free(mb.idx);
        for (Grow *b : {&mb.d_pat, &mb.d_poff, &mb.d_hoff, &mb.d_pos, &mb.d_pid, &mb.h_off, &mb.h_pos, &mb.h_pid}) b->release();
        if (mb.st) cudaStreamDestroy(mb.st);
        if (mb.h_total) cudaFreeHost(mb.h_total);
    }
    if (!g->m.empty()) {
        cudaSetDevice(g->m[0].device);
        for (auto &mb : g->m)
            for (Grow *b : {&mb.s_hoff, &mb.s_pos, &mb.s_pid}) b->release();
        g->o_hoff.release();
        g->o_pos.release();
        g->o_pid.release();
    }
    delete g;
}

int fmx_group_create(const int *device_ids, int ndev, int group_mode, const void *text, uint64_t n, uint32_t char_width,
                     uint64_t max_character, int kind, int level, int index_mode, fmx_group **out) {
    if (!out || !device_ids || ndev < 1 || ndev > 16 || (!text && n)) return gfail(FMX_ERR_INVALID_ARG, "bad argument");
    if (group_mode != FMX_GROUP_REPLICATE && group_mode != FMX_GROUP_BY_PIECE) return gfail(FMX_ERR_INVALID_ARG, "unknown group mode");
    if (group_mode == FMX_GROUP_BY_PIECE && kind != FMX_KIND_MULTI)
        return gfail(FMX_ERR_UNSUPPORTED, "FMX_GROUP_BY_PIECE needs a MultiPieces index");
    if (char_width != 1) return gfail(FMX_ERR_UNSUPPORTED, "only u8 texts (char_width == 1) are supported");
    fmx_group *g = new fmx_group();
    g->mode = group_mode;
    g->n_total = n;
    g->m.resize((size_t)ndev);
    for (int k = 0; k < ndev; k++) g->m[(size_t)k].device = device_ids[k];
    const uint8_t *t = static_cast<const uint8_t *>(text);
    int rc = FMX_OK;
    if (group_mode == FMX_GROUP_REPLICATE) {
        rc = fmx_index_build_ex(text, n, char_width, max_character, kind, level, device_ids[0], index_mode, &g->m[0].idx);
        for (int k = 1; k < ndev && rc == FMX_OK; k++) rc = fmx_index_clone(g->m[0].idx, device_ids[k], &g->m[(size_t)k].idx);
        if (rc == FMX_OK) g->pieces_total = fmx_index_pieces_count(g->m[0].idx);
    } else {
        // piece k is text[start_k, end_k] with text[end_k] == 0 (multi_pieces.rs:53-79)
        std::vector<uint64_t> ends;
        for (const uint8_t *p = t, *stop = t + n; p < stop;) {
            const uint8_t *z = static_cast<const uint8_t *>(std::memchr(p, 0, (size_t)(stop - p)));
            if (!z) break;
            ends.push_back((uint64_t)(z - t));
            p = z + 1;
        }
        if (ends.empty() || ends.back() != n - 1) {
            fmx_group_free(g);
            return gfail(FMX_ERR_INVALID_TEXT, "the given text must end with exactly one zero character");
        }
        if ((int)ends.size() < ndev) {
            fmx_group_free(g);
            return gfail(FMX_ERR_INVALID_ARG, "fewer pieces than devices");
        }
        g->pieces_total = ends.size();
        std::vector<uint64_t> len(ends.size());
        for (size_t k = 0; k < ends.size(); k++) len[k] = ends[k] + 1 - (k ? ends[k - 1] + 1 : 0);
        auto ranges = partition_pieces(len, ndev);
        for (int k = 0; k < ndev && rc == FMX_OK; k++) {
            const size_t first = ranges[(size_t)k].first, last = ranges[(size_t)k].second;
            const uint64_t base = first ? ends[first - 1] + 1 : 0;
            g->m[(size_t)k].pos_base = base;
            g->m[(size_t)k].pid_base = first;
            // one index addresses rows with u32; the WHOLE text may be longer (positions are base + local, u64)
            if (ends[last - 1] + 1 - base >= 0xFFFFFFFFull) {
                fmx_group_free(g);
                return gfail(FMX_ERR_UNSUPPORTED, "a partition holds 2^32 - 1 symbols or more: use more partitions (device ids may repeat)");
            }
            rc = fmx_index_build_ex(t + base, ends[last - 1] + 1 - base, char_width, max_character, kind, level, device_ids[k],
                                    index_mode, &g->m[(size_t)k].idx);
        }
    }
    for (int k = 0; k < ndev && rc == FMX_OK; k++) {
        Member &mb = g->m[(size_t)k];
        if (cudaSetDevice(mb.device) != cudaSuccess || cudaStreamCreateWithFlags(&mb.st, cudaStreamNonBlocking) != cudaSuccess ||
            cudaMallocHost(reinterpret_cast<void **>(&mb.h_total), 64) != cudaSuccess) {
            cudaGetLastError();
            rc = gfail(FMX_ERR_CUDA, "stream / pinned memory creation failed");
        }
    }
    if (rc == FMX_OK && group_mode == FMX_GROUP_BY_PIECE) {
        // peer access so that the gather runs over NVLink when the devices are peers (a plain staged copy otherwise)
        for (int k = 1; k < ndev; k++) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, device_ids[0], device_ids[k]) == cudaSuccess && can) {
                cudaSetDevice(device_ids[0]);
                if (cudaDeviceEnablePeerAccess(device_ids[k], 0) != cudaSuccess) cudaGetLastError();
            }
        }
    }
    if (rc != FMX_OK) {
        fmx_group_free(g);
        return rc;
    }
    *out = g;
    return FMX_OK;
}

int fmx_group_size(const fmx_group *g) { return g ? (int)g->m.size() : 0; }
const fmx_index *fmx_group_index(const fmx_group *g, int k) {
    return (g && k >= 0 && k < (int)g->m.size()) ? g->m[(size_t)k].idx : nullptr;
}
uint64_t fmx_group_len(const fmx_group *g) { return g ? g->n_total : 0; }
uint64_t fmx_group_pieces_count(const fmx_group *g) { return g ? g->pieces_total : 0; }

}  // extern "C"

// ---- REPLICATE: shards in parallel, stitched in input order
static int query_replicated(fmx_group *g, const fmx_query *q, uint64_t *total_hits) {
    const int R = (int)g->m.size();
    const uint64_t npat = q->npat;
    const uint32_t W = q->out_width;
    const bool want_hits = q->positions || q->piece_ids;
    const uint64_t wpp = q->packed_bits ? (q->fixed_len * q->packed_bits + 63) / 64 : 0;
    std::vector<uint64_t> lo((size_t)R + 1);
    for (int k = 0; k <= R; k++) lo[(size_t)k] = npat * (uint64_t)k / (uint64_t)R;
    std::vector<int> rcs((size_t)R, FMX_OK);
    std::vector<uint64_t> totals((size_t)R, 0);
    std::vector<std::string> errs((size_t)R);
    auto work = [&](int k) {
        Member &mb = g->m[(size_t)k];
        const uint64_t a = lo[(size_t)k], n = lo[(size_t)k + 1] - a;
        if (n == 0) return;
        cudaSetDevice(mb.device);
        fmx_query s = *q;
        s.npat = n;
        const uint8_t *p8 = static_cast<const uint8_t *>(q->patterns);
        if (q->packed_bits) s.patterns = p8 + a * wpp * 8;
        else if (q->pat_off) s.pat_off = q->pat_off + a;  // offsets stay batch-global: fmx_query_batch subtracts its own first
        else s.patterns = p8 + a * q->fixed_len;
        if (q->out_s) {
            s.out_s = q->out_s + a;
            s.out_e = q->out_e + a;
        }
        if (q->counts) s.counts = static_cast<uint8_t *>(q->counts) + a * W;
        uint64_t cap = 0;
        if (q->hit_off) {
            if (mb.h_off.ensure((n + 1) * W)) { rcs[(size_t)k] = FMX_ERR_OOM; return; }
            s.hit_off = mb.h_off.p;
        }
        for (int attempt = 0; attempt < 2; attempt++) {
            if (want_hits) {
                if (attempt == 0) cap = q->capacity / (uint64_t)R + q->capacity / (uint64_t)(4 * R) + 4096;
                if (cap > q->capacity) cap = q->capacity;
                if (q->positions) {
                    if (mb.h_pos.ensure(cap * W + 16)) { rcs[(size_t)k] = FMX_ERR_OOM; return; }
                    s.positions = mb.h_pos.p;
                }
                if (q->piece_ids) {
                    if (mb.h_pid.ensure(cap * W + 16)) { rcs[(size_t)k] = FMX_ERR_OOM; return; }
                    s.piece_ids = mb.h_pid.p;
                }
                s.capacity = cap;
            }
            uint64_t tot = 0;
            int rc = fmx_query_batch(mb.idx, &s, &tot);
            totals[(size_t)k] = tot;
            rcs[(size_t)k] = rc;
            if (rc == FMX_ERR_CAPACITY && attempt == 0 && tot <= q->capacity) {
                cap = tot;  // the even split was too small for this shard: once more with room for every hit
                continue;
            }
            if (rc != FMX_OK) errs[(size_t)k] = fmx_last_error();
            break;
        }
    };
    std::vector<std::thread> th;
    for (int k = 0; k < R; k++) th.emplace_back(work, k);
    for (auto &t : th) t.join();
    for (int k = 0; k < R; k++)
        if (rcs[(size_t)k] != FMX_OK && rcs[(size_t)k] != FMX_ERR_CAPACITY) {
            std::fprintf(stderr, "fmx_group: device %d: %s\n", g->m[(size_t)k].device, errs[(size_t)k].c_str());
            return rcs[(size_t)k];
        }
    uint64_t running = 0;
    bool overflow = false;
    for (int k = 0; k < R; k++) {
        if (rcs[(size_t)k] == FMX_ERR_CAPACITY) overflow = true;
        running += totals[(size_t)k];
    }
    if (total_hits) *total_hits = running;
    if (want_hits && running > q->capacity) overflow = true;
    if (W == 4 && running > 0xFFFFFFFFull) overflow = true;
    if (!q->hit_off) return overflow ? gfail(FMX_ERR_CAPACITY, "output buffer too small: total_hits holds the size needed") : FMX_OK;
    // stitch: offsets rebased by the hits of the shards before, hit lists copied behind one another
    uint64_t base = 0;
    for (int k = 0; k < R; k++) {
        Member &mb = g->m[(size_t)k];
        const uint64_t a = lo[(size_t)k], n = lo[(size_t)k + 1] - a, tot = totals[(size_t)k];
        if (W == 8) {
            const uint64_t *src = static_cast<const uint64_t *>(mb.h_off.p);
            uint64_t *dst = static_cast<uint64_t *>(q->hit_off) + a;
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < (int64_t)n; i++) dst[i] = src[i] + base;
        } else {
            const uint32_t *src = static_cast<const uint32_t *>(mb.h_off.p);
            uint32_t *dst = static_cast<uint32_t *>(q->hit_off) + a;
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < (int64_t)n; i++) dst[i] = (uint32_t)(src[i] + base);
        }
        if (want_hits && !overflow && tot) {
            const uint64_t piece = 8ull << 20;
            const int64_t np = (int64_t)((tot * W + piece - 1) / piece);
#pragma omp parallel for schedule(static)
            for (int64_t c = 0; c < np; c++) {
                const uint64_t o = (uint64_t)c * piece, len = std::min(piece, tot * W - o);
                if (q->positions) std::memcpy(static_cast<uint8_t *>(q->positions) + base * W + o, static_cast<uint8_t *>(mb.h_pos.p) + o, len);
                if (q->piece_ids) std::memcpy(static_cast<uint8_t *>(q->piece_ids) + base * W + o, static_cast<uint8_t *>(mb.h_pid.p) + o, len);
            }
        }
        base += tot;
    }
    if (W == 8) static_cast<uint64_t *>(q->hit_off)[npat] = running;
    else static_cast<uint32_t *>(q->hit_off)[npat] = (uint32_t)running;
    return overflow ? gfail(FMX_ERR_CAPACITY, "output buffer too small: total_hits holds the size needed") : FMX_OK;
}

#define G_TRY(expr)                                   \
    do {                                              \
        if ((expr) != cudaSuccess) {                  \
            cudaGetLastError();                       \
            return gfail(FMX_ERR_CUDA, #expr);        \
        }                                             \
    } while (0)

// ---- BY_PIECE: every member answers the whole batch; gather on the first device; one merge
static int query_by_piece(fmx_group *g, const fmx_query *q, uint64_t *total_hits) {
    const int R = (int)g->m.size();
    const uint64_t npat = q->npat;
    if (q->out_width != 8) return gfail(FMX_ERR_UNSUPPORTED, "FMX_GROUP_BY_PIECE produces uint64 outputs (out_width 8)");
    if (q->out_s || q->out_e) return gfail(FMX_ERR_UNSUPPORTED, "SA rows of different partitions are unrelated: out_s / out_e are not available");
    // search_prefix / search_exact: "starts a piece" is a piece-local property, so the L == 0 filter of every partition is
    // the filter of the whole text; but then the merged offsets count FILTERED hits, not e - s
    if (q->counts && (q->mode == FMX_SEARCH_PREFIX || q->mode == FMX_SEARCH_EXACT))
        return gfail(FMX_ERR_UNSUPPORTED, "BY_PIECE counts (e - s, unfiltered) are not available in the filtered modes");
    const bool want_pos = q->positions != nullptr, want_pid = q->piece_ids != nullptr;
    const uint64_t wpp = q->packed_bits ? (q->fixed_len * q->packed_bits + 63) / 64 : 0;
    const uint64_t pat_bytes = q->packed_bits ? npat * wpp * 8 : (q->pat_off ? q->pat_off[npat] - q->pat_off[0] : npat * q->fixed_len);
    const uint8_t *p8 = static_cast<const uint8_t *>(q->patterns) + (q->packed_bits || !q->pat_off ? 0 : q->pat_off[0]);
    if (!q->packed_bits) {  // a pattern holding \0 could span the partition
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
        for (int64_t i = 0; i < (int64_t)pat_bytes; i++) bad |= p8[i] == 0;
        if (bad) return gfail(FMX_ERR_INVALID_ARG, "patterns containing \\0 are not supported by the piece-partitioned group");
    }
    uint64_t cap = (want_pos || want_pid) ? q->capacity : 0;
    if (cap > 0xFFFFFFF0ull) cap = 0xFFFFFFF0ull;
    for (int attempt = 0; attempt < 2; attempt++) {
        for (int k = 0; k < R; k++) {
            Member &mb = g->m[(size_t)k];
            G_TRY(cudaSetDevice(mb.device));
            if (mb.d_pat.ensure(pat_bytes + 16) || mb.d_hoff.ensure((npat + 1) * 4) || (want_pos && mb.d_pos.ensure(cap * 4 + 16)) ||
                (want_pid && mb.d_pid.ensure(cap * 4 + 16)))
                return gfail(FMX_ERR_OOM, "out of device memory for the batch");
            if (pat_bytes) G_TRY(cudaMemcpyAsync(mb.d_pat.p, p8, pat_bytes, cudaMemcpyHostToDevice, mb.st));
            fmx_query s;
            std::memset(&s, 0, sizeof(s));
            s.mode = q->mode;
            s.packed_bits = q->packed_bits;
            s.patterns = mb.d_pat.p;
            s.fixed_len = q->fixed_len;
            s.npat = npat;
            s.out_width = 4;
            if (q->pat_off) {
                if (mb.d_poff.ensure((npat + 1) * 8)) return gfail(FMX_ERR_OOM, "out of device memory for the pattern offsets");
                G_TRY(cudaMemcpyAsync(mb.d_poff.p, q->pat_off, (npat + 1) * 8, cudaMemcpyHostToDevice, mb.st));
                s.pat_off = static_cast<const uint64_t *>(mb.d_poff.p);
                s.patterns = static_cast<const uint8_t *>(mb.d_pat.p) - q->pat_off[0];
            }
            s.hit_off = mb.d_hoff.p;
            s.positions = want_pos ? mb.d_pos.p : nullptr;
            s.piece_ids = want_pid ? mb.d_pid.p : nullptr;
            s.capacity = cap;
            int rc = fmx_query_batch_device(mb.idx, &s, mb.st);
            if (rc) return rc;
            *mb.h_total = 0;
            G_TRY(cudaMemcpyAsync(mb.h_total, static_cast<uint32_t *>(mb.d_hoff.p) + npat, 4, cudaMemcpyDeviceToHost, mb.st));
        }
        uint64_t need = 0;
        for (int k = 0; k < R; k++) {
            Member &mb = g->m[(size_t)k];
            G_TRY(cudaSetDevice(mb.device));
            G_TRY(cudaStreamSynchronize(mb.st));
            int rc = fmx_search_check(mb.idx, mb.st);
            if (rc) return rc;
            need = std::max<uint64_t>(need, *mb.h_total);
        }
        if (!(want_pos || want_pid) || need <= cap || attempt == 1) break;
        cap = need;  // a partition found more hits than there was room for: once more
    }
    uint64_t total = 0;
    for (int k = 0; k < R; k++) total += *g->m[(size_t)k].h_total;
    if (total_hits) *total_hits = total;
    const bool overflow = (want_pos || want_pid) && total > q->capacity;
    // gather on the first device and merge there
    Member &root = g->m[0];
    G_TRY(cudaSetDevice(root.device));
    std::vector<fmx_csr_part> parts((size_t)R);
    for (int k = 0; k < R; k++) {
        Member &mb = g->m[(size_t)k];
        const uint64_t tk = std::min<uint64_t>(*mb.h_total, cap);
        const void *hoff = mb.d_hoff.p, *pos = mb.d_pos.p, *pid = mb.d_pid.p;
        if (k > 0) {
            if (mb.s_hoff.ensure((npat + 1) * 4) || (want_pos && mb.s_pos.ensure(tk * 4 + 16)) || (want_pid && mb.s_pid.ensure(tk * 4 + 16)))
                return gfail(FMX_ERR_OOM, "out of device memory for the gathered parts");
            G_TRY(cudaMemcpyPeerAsync(mb.s_hoff.p, root.device, mb.d_hoff.p, mb.device, (npat + 1) * 4, root.st));
            if (want_pos && tk) G_TRY(cudaMemcpyPeerAsync(mb.s_pos.p, root.device, mb.d_pos.p, mb.device, tk * 4, root.st));
            if (want_pid && tk) G_TRY(cudaMemcpyPeerAsync(mb.s_pid.p, root.device, mb.d_pid.p, mb.device, tk * 4, root.st));
            hoff = mb.s_hoff.p;
            pos = mb.s_pos.p;
            pid = mb.s_pid.p;
        }
        parts[(size_t)k].hit_off = hoff;
        parts[(size_t)k].positions = want_pos ? pos : nullptr;
        parts[(size_t)k].piece_ids = want_pid ? pid : nullptr;
        parts[(size_t)k].position_base = mb.pos_base;
        parts[(size_t)k].piece_base = mb.pid_base;
    }
    const uint64_t out_cap = overflow ? 0 : total;
    if (g->o_hoff.ensure((npat + 1) * 8) || (want_pos && g->o_pos.ensure(out_cap * 8 + 16)) || (want_pid && g->o_pid.ensure(out_cap * 8 + 16)))
        return gfail(FMX_ERR_OOM, "out of device memory for the merged CSR");
    int rc = fmx_csr_merge_device(root.device, parts.data(), R, npat, 4, static_cast<uint64_t *>(g->o_hoff.p),
                                  want_pos ? static_cast<uint64_t *>(g->o_pos.p) : nullptr,
                                  want_pid ? static_cast<uint64_t *>(g->o_pid.p) : nullptr, out_cap, root.st);
    if (rc) return rc;
    if (q->hit_off) G_TRY(cudaMemcpyAsync(q->hit_off, g->o_hoff.p, (npat + 1) * 8, cudaMemcpyDeviceToHost, root.st));
    if (!overflow && total) {
        if (want_pos) G_TRY(cudaMemcpyAsync(q->positions, g->o_pos.p, total * 8, cudaMemcpyDeviceToHost, root.st));
        if (want_pid) G_TRY(cudaMemcpyAsync(q->piece_ids, g->o_pid.p, total * 8, cudaMemcpyDeviceToHost, root.st));
    }
    G_TRY(cudaStreamSynchronize(root.st));
    if (q->counts && q->hit_off) {  // e - s summed over the partitions = the hit counts (unfiltered modes)
        const uint64_t *off = static_cast<const uint64_t *>(q->hit_off);
        uint64_t *cnt = static_cast<uint64_t *>(q->counts);
#pragma omp parallel for schedule(static)
        for (int64_t p = 0; p < (int64_t)npat; p++) cnt[p] = off[p + 1] - off[p];
    }
    return overflow ? gfail(FMX_ERR_CAPACITY, "output buffer too small: total_hits holds the size needed") : FMX_OK;
}

extern "C" int fmx_group_query_batch(const fmx_group *gc, const fmx_query *q, uint64_t *total_hits) {
    if (!gc || !q) return gfail(FMX_ERR_INVALID_ARG, "null argument");
    fmx_group *g = const_cast<fmx_group *>(gc);
    if (total_hits) *total_hits = 0;
    if (q->out_width != 8 && q->out_width != 4) return gfail(FMX_ERR_INVALID_ARG, "out_width must be 8 or 4");
    if ((q->positions || q->piece_ids) && !q->hit_off) return gfail(FMX_ERR_INVALID_ARG, "positions / piece_ids need hit_off");
    if (q->counts && g->mode == FMX_GROUP_BY_PIECE && !q->hit_off) return gfail(FMX_ERR_INVALID_ARG, "BY_PIECE counts need hit_off");
    std::lock_guard<std::mutex> lk(g->mu);
    if (q->npat == 0) {
        if (q->hit_off) std::memset(q->hit_off, 0, q->out_width);
        return FMX_OK;
    }
    return g->mode == FMX_GROUP_REPLICATE ? query_replicated(g, q, total_hits) : query_by_piece(g, q, total_hits);
}
