// gpu_build.cu -- index construction on the GPU, next to the suffix array.
//
// Round 1 built only the suffix array on the device; the BWT, the rank blocks, the inverse suffix array, the samples
// and the blob assembly stayed host work (13 of the 14 s of the 1 GB DNA build, plus a 10.5 GB host-to-device copy).
// Here the whole device-layout blob (fmx_layout.h) of an FMIndex / FMIndexMultiPieces index in the Q4 layout
// (max_character <= 4: the DNA configs) or the SYM layout (every other u8 alphabet within the budget: the byte
// config) is produced in device memory:
//
//   reference producer (file:line)                         here
//   suffix array            sais.rs:115-144                gpu_sa.cu (prefix doubling), left on the device
//   BWT                     fm_index.rs:44-58              k_bwt: bw[i] = text[sa[i] - 1], 0 when sa[i] == 0
//   rank structure          (vers-vecs WaveletMatrix)      k_q4_pack -> own scans (scan3.cuh) -> k_q4_counts: Q4 blocks
//                                                          k_sym_pack -> one scan -> k_sym_counts: per-symbol RB192 vectors
//   SA samples              sample.rs:21-44                k_samples: sa[i << level]
//   inverse SA / full SA / text (verify, rich locate)      k_isa (scatter), device copies
//   cs                      sais.rs:9-32                   host histogram of the text (5 counters)
//   doc / first row         multi_pieces.rs:53-97          host, from the <= 1024 zero rows gathered from the device
//
// The result is byte-identical to the host builder's blob (tests/test_gpu_parity.py::
// test_gpu_built_blob_identical_to_host_built): same header, same sections, same padding.
// RLFMIndex: gpu_build_rlfm_blob below (run encoding rlfmi.rs:37-96 on the device).
// Other layouts (WM4, binary wavelet, WIDE) keep the host builder (builder.cpp).
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fmx.h"
#include "builder.h"
#include "scan3.cuh"

namespace fmx {

__global__ void k_bwt(const uint8_t *text, const uint32_t *sa, uint64_t n, uint8_t *bwt) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t p = sa[i];
        bwt[i] = p ? text[p - 1] : (uint8_t)0;
    }
}

// one thread per Q4 block: 64 symbols -> two bit planes of the codes (symbol - 1; \0 stored as code 0) in words 4..7,
// the block's own code counts to cnt[c * nblk + b] (prefix sums follow)
__global__ void k_q4_pack(const uint8_t *bwt, uint64_t n, uint64_t nblk, uint32_t *blocks, uint32_t *cnt) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint64_t lo = b * 64, hi = lo + 64 < n ? lo + 64 : n;
    uint32_t p0[2] = {0, 0}, p1[2] = {0, 0}, c[4] = {0, 0, 0, 0};
    for (uint64_t i = lo; i < hi; i++) {
        const uint32_t sym = bwt[i], code = sym ? sym - 1u : 0u, t = (uint32_t)(i - lo);
        p0[t >> 5] |= (code & 1u) << (t & 31u);
        p1[t >> 5] |= (code >> 1) << (t & 31u);
        c[code]++;
    }
    uint32_t *blk = blocks + b * 8;
    blk[4] = p0[0];
    blk[5] = p0[1];
    blk[6] = p1[0];
    blk[7] = p1[1];
    for (int k = 0; k < 4; k++) cnt[(uint64_t)k * nblk + b] = c[k];
}

__global__ void k_q4_counts(uint32_t *blocks, const uint32_t *cnt, uint64_t nblk) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    for (int k = 0; k < 4; k++) blocks[b * 8 + k] = cnt[(uint64_t)k * nblk + b];
}

// SYM layout: thread (b, c) builds block b of symbol c's RB192 vector -- payload bit t = [bwt[192 b + t] == c], the two
// in-block sub-counts -- and leaves the block's popcount in cnt[c * nblk + b] for the prefix sums.  The 192 bytes of a
// block are compared four at a time; threads of a warp share b's bytes through L1.
__global__ void __launch_bounds__(256) k_sym_pack(const uint8_t *bwt, uint64_t n, uint64_t nblk, uint32_t cs_len, uint32_t *blocks,
                                                  uint32_t *cnt) {
    const uint64_t b = (uint64_t)blockIdx.x * (256 / 32) + threadIdx.x / 32;  // a warp per block b, lanes over symbols
    if (b >= nblk) return;
    const uint64_t lo = b * FMX_RB_BITS;
    const uint32_t *w4 = reinterpret_cast<const uint32_t *>(bwt + lo);  // 192 b is a multiple of 4; bwt is cudaMalloc'd
    const uint64_t avail = lo < n ? (n - lo < FMX_RB_BITS ? n - lo : FMX_RB_BITS) : 0;
    for (uint32_t c = threadIdx.x & 31; c < cs_len; c += 32) {
        uint32_t pay[6];
        const uint32_t pat = c * 0x01010101u;
#pragma unroll
        for (uint32_t w = 0; w < 6; w++) {  // payload word w = symbols 32 w .. 32 w + 31 = text words 8 w .. 8 w + 7
            uint32_t acc = 0;
#pragma unroll
            for (uint32_t q = 0; q < 8; q++) {
                const uint32_t k = w * 8 + q;  // 4 symbols per text word
                if ((uint64_t)k * 4 < avail) {
                    uint32_t m = __vcmpeq4(__ldg(w4 + k), pat) & 0x01010101u;
                    const uint64_t left = avail - (uint64_t)k * 4;
                    if (left < 4) m &= 0x01010101u >> (8u * (4u - (uint32_t)left));
                    acc |= ((m * 0x01020408u) >> 24) << (q * 4u);  // byte j of the word -> bit j
                }
            }
            pay[w] = acc;
        }
        const uint32_t p0 = __popc(pay[0]) + __popc(pay[1]), p1 = __popc(pay[2]) + __popc(pay[3]), p2 = __popc(pay[4]) + __popc(pay[5]);
        uint32_t *blk = blocks + ((uint64_t)c * nblk + b) * 8;
        blk[1] = (p0 << 8) | ((p0 + p1) << 16);
        for (int k = 0; k < 6; k++) blk[2 + k] = pay[k];
        cnt[(uint64_t)c * nblk + b] = p0 + p1 + p2;
    }
}

// word 0 of every block = ones of its vector before it: the running sum over the whole (symbol-major) count array minus
// the sum at the start of the symbol's vector
__global__ void k_sym_counts(uint32_t *blocks, const uint32_t *scan, uint64_t nblk, uint64_t total_blocks) {
    const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= total_blocks) return;
    const uint64_t c = x / nblk;
    blocks[x * 8] = scan[x] - scan[c * nblk];
}

// rows whose BWT symbol is \0 (at most `cap`; *count keeps counting past it)
__global__ void k_zero_rows(const uint8_t *bwt, uint64_t n, uint32_t *rows, uint32_t cap, unsigned *count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && bwt[i] == 0) {
        const unsigned k = atomicAdd(count, 1u);
        if (k < cap) rows[k] = (uint32_t)i;
    }
}

__global__ void k_gather32(const uint32_t *src, const uint32_t *idx, uint32_t m, uint32_t *out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) out[k] = src[idx[k]];
}

__global__ void k_isa(const uint32_t *sa, uint64_t n, uint32_t *isa) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) isa[sa[i]] = (uint32_t)i;
}

__global__ void k_samples(const uint32_t *sa, uint64_t count, uint32_t level, uint32_t *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = sa[i << level];
}

// ---- run-length encoding (rlfmi.rs:37-96) on the device
// a run starts at row i when the BWT symbol differs from the previous one (the symbol "before" row 0 is \0, rlfmi.rs:47)
__global__ void k_run_flags(const uint8_t *bwt, uint64_t n, uint8_t *flag, uint32_t *flag32) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint8_t f = bwt[i] != (i ? bwt[i - 1] : (uint8_t)0);
        flag[i] = f;
        flag32[i] = f;
    }
}
// RB192 vector from a 0/1 byte per position: one thread per block (payload + in-block sub-counts; cnt[b] = its ones)
__global__ void k_rb_pack(const uint8_t *flag, uint64_t n, uint64_t nblk, uint32_t *blocks, uint32_t *cnt) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint64_t lo = b * FMX_RB_BITS;
    uint32_t w[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        uint32_t acc = 0;
        for (uint32_t t = 0; t < 32; t++) {
            const uint64_t i = lo + (uint64_t)k * 32 + t;
            if (i < n && flag[i]) acc |= 1u << t;
        }
        w[k] = acc;
    }
    const uint32_t p0 = __popc(w[0]) + __popc(w[1]), p1 = __popc(w[2]) + __popc(w[3]), p2 = __popc(w[4]) + __popc(w[5]);
    uint32_t *blk = blocks + b * 8;
    blk[1] = (p0 << 8) | ((p0 + p1) << 16);
#pragma unroll
    for (int k = 0; k < 6; k++) blk[2 + k] = w[k];
    cnt[b] = p0 + p1 + p2;
}
__global__ void k_rb_counts(uint32_t *blocks, const uint32_t *scan, uint64_t nblk) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nblk) blocks[b * 8] = scan[b];
}
// run j = (number of run starts before row i) for every flagged row: its first row and its head symbol
__global__ void k_runs_scatter(const uint8_t *flag, const uint32_t *ridx, const uint8_t *bwt, uint64_t n, uint32_t *starts,
                               uint8_t *heads, uint32_t *order) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) {
        const uint32_t j = ridx[i];
        starts[j] = (uint32_t)i;
        heads[j] = bwt[i];
        order[j] = j;
    }
}
// length of the run that is k-th in (head, run index) order
__global__ void k_sorted_run_len(const uint32_t *order, const uint32_t *starts, uint64_t r, uint64_t n, uint32_t *len) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < r) {
        const uint32_t j = order[k];
        len[k] = (j + 1 < r ? starts[j + 1] : (uint32_t)n) - starts[j];
    }
}
// bp holds, head by head, every run as 1 0^{len-1} (rlfmi.rs:70-83): the k-th sorted run starts at pos[k]
__global__ void k_mark_bp(const uint32_t *pos, uint64_t r, uint8_t *flag) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < r) flag[pos[k]] = 1;
}

static inline uint32_t log2_u64(uint64_t x) { return 63u - (uint32_t)__builtin_clzll(x); }

#define GB_TRY(expr)                                                                  \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            err = std::string("GPU index build: ") + #expr + ": " + cudaGetErrorString(_e); \
            rc = _e == cudaErrorMemoryAllocation ? FMX_ERR_OOM : FMX_ERR_CUDA;        \
            goto fail;                                                                \
        }                                                                             \
    } while (0)

int gpu_build_blob(const uint8_t *text, uint64_t n, uint64_t mc, int kind, int level, int mode, int device,
                   void **d_blob_out, FmxBlobHeader *hdr_out, std::string &err) {
    *d_blob_out = nullptr;
    if (mc == 0 || mc > 255 || (kind != FMX_KIND_FM && kind != FMX_KIND_MULTI) || n < (1u << 16) || n >= (1ull << 32) - 1)
        return FMX_ERR_UNSUPPORTED;
    // the text's zeros: the Q4 exception list holds as many, and they are the piece ends of a multi-piece text
    std::vector<uint32_t> zeros;
    {
        const uint64_t chunk = 1ull << 22;
        const int64_t nchunks = (int64_t)((n + chunk - 1) / chunk);
        std::vector<std::vector<uint32_t>> part((size_t)nchunks);
        int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(| : bad)
        for (int64_t c = 0; c < nchunks; c++) {
            const uint64_t lo = (uint64_t)c * chunk, hi = lo + chunk < n ? lo + chunk : n;
            for (uint64_t i = lo; i < hi; i++) {
                bad |= text[i] > mc;
                if (text[i] == 0) part[(size_t)c].push_back((uint32_t)i);
            }
        }
        if (bad) return FMX_ERR_UNSUPPORTED;  // the host builder reports it
        for (auto &v : part) zeros.insert(zeros.end(), v.begin(), v.end());
    }
    if (zeros.empty() || zeros.back() != n - 1 || text[0] == 0 || (n >= 2 && text[n - 2] == 0))
        return FMX_ERR_UNSUPPORTED;  // invalid texts: the host builder's business
    const uint32_t L = log2_u64(mc) + 1, cs_len = (uint32_t)mc + 1;
    // the layout, by the host builder's rules (builder.cpp: want_q4, then SYM within its budget)
    const bool use_q4 = mc <= 4 && !q4_forbidden_by_env() && zeros.size() <= FMX_MAX_EXC;
    const bool use_sym = !use_q4 && sym_layout_chosen(cs_len, n);
    if (!use_q4 && !use_sym) return FMX_ERR_UNSUPPORTED;
    const bool need_rows = use_q4 || kind == FMX_KIND_MULTI;  // rows of the \0 symbols: exception list, doc
    const uint32_t row_cap = use_q4 ? FMX_MAX_EXC : (1u << 22);
    if (need_rows && zeros.size() > row_cap) return FMX_ERR_UNSUPPORTED;
    const bool interior_zero = kind != FMX_KIND_MULTI && zeros.size() > 1;
    if (int mrc = resolve_mode(mode, err)) return mrc;

    const uint64_t nblk = use_q4 ? n / 64 + 1 : n / FMX_RB_BITS + 1;
    const uint64_t rank_bytes = use_q4 ? nblk * 32 : sym_layout_bytes(cs_len, n);
    const VerifyPlan vp = plan_verify(kind, n, mode, level, interior_zero, rank_bytes, use_sym);
    if (vp.verify && !vp.dense) return FMX_ERR_UNSUPPORTED;  // the sampled verify form: host builder

    FmxBlobHeader hdr;
    std::memset(&hdr, 0, sizeof(hdr));
    hdr.magic = FMX_BLOB_MAGIC;
    hdr.version = FMX_BLOB_VERSION;
    hdr.kind = (uint32_t)kind;
    hdr.n = n;
    hdr.seq_len = n;
    hdr.levels = L;
    hdr.max_character = (uint32_t)mc;
    hdr.cs_len = cs_len;
    hdr.layout = use_q4 ? FMX_LAYOUT_QUAT : FMX_LAYOUT_SYM;
    hdr.char_width = 1;
    hdr.nexc = use_q4 ? (uint32_t)zeros.size() : 0u;
    hdr.sym_nblk = use_sym ? (uint32_t)nblk : 0u;
    hdr.reserved[0] = (uint64_t)mode;
    if (vp.verify) {
        hdr.verify = 1;
        hdr.isa_level = vp.isa_level;
    }
    if (level >= 0) {
        hdr.has_locate = 1;
        uint32_t lvl = (uint32_t)level;
        hdr.sa_word_size = log2_u64(n) + 1;
        if (lvl >= 63 || n <= (1ull << lvl)) lvl = 0;  // sample.rs:28-31
        hdr.sa_level = lvl;
        hdr.sa_count = ((n - 1) >> lvl) + 1;
    }
    hdr.vsa_level = vp.dense ? 0u : hdr.sa_level;
    if (kind == FMX_KIND_MULTI) hdr.ndoc = zeros.size();

    // cs (sais.rs:21-32)
    std::vector<uint32_t> cs(cs_len + 1, 0), adj(cs_len, 0);
    {
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
        for (uint32_t c = 0; c < cs_len; c++) {
            cs[c] = (uint32_t)sum;
            sum += occ[c];
        }
        cs[cs_len] = (uint32_t)n;
    }

    uint64_t bytes[SEC_COUNT];
    std::memset(bytes, 0, sizeof(bytes));
    bytes[SEC_LEVEL0] = rank_bytes;
    if (use_sym) bytes[SEC_LEVEL0 + 1] = n;  // the raw sequence
    if (use_q4) bytes[SEC_EXC] = zeros.size() * 4;
    bytes[SEC_ADJ] = (uint64_t)cs_len * 4;
    bytes[SEC_CS] = (uint64_t)(cs_len + 1) * 4;
    if (hdr.has_locate) bytes[SEC_SA] = hdr.sa_count * 4;
    if (vp.verify) {
        bytes[SEC_TEXT] = n;
        bytes[SEC_ISA] = n * 4;
    }
    if (vp.dense_sa) bytes[SEC_VSA] = n * 4;
    if (kind == FMX_KIND_MULTI) {
        bytes[SEC_DOC] = zeros.size() * 4;
        bytes[SEC_PIECE_END] = zeros.size() * 4;
    }
    layout_sections(hdr, bytes);

    int rc = FMX_ERR_CUDA;
    uint8_t *d_text = nullptr, *d_bwt = nullptr, *blob = nullptr;
    uint32_t *d_sa = nullptr, *d_cnt = nullptr, *d_rows = nullptr, *d_rowsa = nullptr;
    unsigned *d_count = nullptr;
    const unsigned T = 256;
    auto grid = [&](uint64_t m) { return (unsigned)((m + T - 1) / T); };
    auto at = [&](int k) { return blob + hdr.sec[k].offset; };
    std::vector<uint32_t> rows, rowsa, doc;
    unsigned nz = 0;
    const uint64_t cnt_entries = use_q4 ? nblk * 4 : nblk * cs_len;

    GB_TRY(cudaSetDevice(device));
    GB_TRY(cudaMalloc(&d_text, n));
    GB_TRY(cudaMemcpy(d_text, text, n, cudaMemcpyHostToDevice));
    {
        std::string serr;
        int src = gpu_suffix_array_device(d_text, n, L, device, &d_sa, nullptr, serr);
        if (src) {
            err = "GPU suffix array construction failed: " + serr;
            rc = src;
            goto fail;
        }
    }
    GB_TRY(cudaMalloc(&blob, hdr.total_bytes));
    GB_TRY(cudaMemset(blob, 0, hdr.total_bytes));
    GB_TRY(cudaMalloc(&d_bwt, n + 16));  // k_sym_pack reads whole words
    GB_TRY(cudaMemset(d_bwt + n, 0, 16));
    GB_TRY(cudaMalloc(&d_cnt, cnt_entries * 4));
    GB_TRY(cudaMalloc(&d_rows, (uint64_t)row_cap * 4));
    GB_TRY(cudaMalloc(&d_rowsa, (uint64_t)row_cap * 4));
    GB_TRY(cudaMalloc(&d_count, 4));
    GB_TRY(cudaMemset(d_count, 0, 4));

    k_bwt<<<grid(n), T>>>(d_text, d_sa, n, d_bwt);
    if (use_q4) {
        k_q4_pack<<<grid(nblk), T>>>(d_bwt, n, nblk, reinterpret_cast<uint32_t *>(at(SEC_LEVEL0)), d_cnt);
        GB_TRY(cudaGetLastError());
        for (int c = 0; c < 4; c++)  // occurrences of code c before every block
            GB_TRY((scan3::run<uint32_t, scan3::Sum32, true>(d_cnt + (uint64_t)c * nblk, nblk, d_cnt + (uint64_t)c * nblk, scan3::Sum32(), 0u, 0)));
        k_q4_counts<<<grid(nblk), T>>>(reinterpret_cast<uint32_t *>(at(SEC_LEVEL0)), d_cnt, nblk);
    } else {
        k_sym_pack<<<(unsigned)((nblk + 7) / 8), 256>>>(d_bwt, n, nblk, cs_len, reinterpret_cast<uint32_t *>(at(SEC_LEVEL0)), d_cnt);
        GB_TRY(cudaGetLastError());
        // one running sum over all vectors' block popcounts (symbol-major; the grand total is n < 2^32)
        GB_TRY((scan3::run<uint32_t, scan3::Sum32, true>(d_cnt, cnt_entries, d_cnt, scan3::Sum32(), 0u, 0)));
        k_sym_counts<<<grid(cnt_entries), T>>>(reinterpret_cast<uint32_t *>(at(SEC_LEVEL0)), d_cnt, nblk, cnt_entries);
        GB_TRY(cudaMemcpy(at(SEC_LEVEL0 + 1), d_bwt, n, cudaMemcpyDeviceToDevice));
    }
    GB_TRY(cudaGetLastError());
    if (need_rows) {
        k_zero_rows<<<grid(n), T>>>(d_bwt, n, d_rows, row_cap, d_count);
        GB_TRY(cudaGetLastError());
        GB_TRY(cudaMemcpy(&nz, d_count, 4, cudaMemcpyDeviceToHost));
        if (nz != zeros.size()) {
            err = "GPU index build: the BWT holds another number of zeros than the text";
            rc = FMX_ERR_CUDA;
            goto fail;
        }
        rows.resize(nz);
        GB_TRY(cudaMemcpy(rows.data(), d_rows, (uint64_t)nz * 4, cudaMemcpyDeviceToHost));
        std::sort(rows.begin(), rows.end());  // ascending rows whose symbol is \0: SEC_EXC, and select(bw, k, 0)
        if (use_q4) GB_TRY(cudaMemcpy(at(SEC_EXC), rows.data(), (uint64_t)nz * 4, cudaMemcpyHostToDevice));
    }
    if (kind == FMX_KIND_MULTI) {  // multi_pieces.rs:53-79
        GB_TRY(cudaMemcpy(d_rows, rows.data(), (uint64_t)nz * 4, cudaMemcpyHostToDevice));
        k_gather32<<<(nz + T - 1) / T, T>>>(d_sa, d_rows, nz, d_rowsa);
        GB_TRY(cudaGetLastError());
        rowsa.resize(nz);
        GB_TRY(cudaMemcpy(rowsa.data(), d_rowsa, (uint64_t)nz * 4, cudaMemcpyDeviceToHost));
        doc.assign(nz, 0);
        for (unsigned k = 0; k < nz; k++) {  // rows[k] = select(bw, k, 0)
            const uint64_t em = rowsa[k] ? rowsa[k] - 1 : n - 1;  // modular_sub(sa[p], 1, n)
            const uint64_t pid = (uint64_t)(std::lower_bound(zeros.begin(), zeros.end(), (uint32_t)em) - zeros.begin());
            if (pid == nz - 1) hdr.first_row = rows[k];
            doc[k] = (uint32_t)pid;
        }
        GB_TRY(cudaMemcpy(at(SEC_DOC), doc.data(), (uint64_t)nz * 4, cudaMemcpyHostToDevice));
        GB_TRY(cudaMemcpy(at(SEC_PIECE_END), zeros.data(), (uint64_t)nz * 4, cudaMemcpyHostToDevice));
    }
    GB_TRY(cudaMemcpy(at(SEC_ADJ), adj.data(), adj.size() * 4, cudaMemcpyHostToDevice));
    GB_TRY(cudaMemcpy(at(SEC_CS), cs.data(), cs.size() * 4, cudaMemcpyHostToDevice));
    if (hdr.has_locate) k_samples<<<grid(hdr.sa_count), T>>>(d_sa, hdr.sa_count, hdr.sa_level, reinterpret_cast<uint32_t *>(at(SEC_SA)));
    if (vp.verify) {
        GB_TRY(cudaMemcpy(at(SEC_TEXT), d_text, n, cudaMemcpyDeviceToDevice));
        k_isa<<<grid(n), T>>>(d_sa, n, reinterpret_cast<uint32_t *>(at(SEC_ISA)));
    }
    if (vp.dense_sa) GB_TRY(cudaMemcpy(at(SEC_VSA), d_sa, n * 4, cudaMemcpyDeviceToDevice));
    GB_TRY(cudaGetLastError());
    GB_TRY(cudaMemcpy(blob, &hdr, sizeof(hdr), cudaMemcpyHostToDevice));
    GB_TRY(cudaDeviceSynchronize());
    *d_blob_out = blob;
    *hdr_out = hdr;
    blob = nullptr;
    rc = 0;
fail:
    cudaFree(d_text);
    cudaFree(d_bwt);
    cudaFree(d_sa);
    cudaFree(d_cnt);
    cudaFree(d_rows);
    cudaFree(d_rowsa);
    cudaFree(d_count);
    cudaFree(blob);
    if (rc) cudaGetLastError();
    return rc;
}

// RLFMIndex (rlfmi.rs:37-96) in device memory: suffix array, BWT, run starts -> b, run heads -> their rank structure
// (Q4 or SYM over `runs` symbols), runs sorted stably by head -> bp and the two select tables, samples; in the HBM-rich
// mode also the full suffix array and the text.  Byte-identical to the host builder's blob.
int gpu_build_rlfm_blob(const uint8_t *text, uint64_t n, uint64_t mc, int level, int mode, int device, void **d_blob_out,
                        FmxBlobHeader *hdr_out, std::string &err) {
    *d_blob_out = nullptr;
    if (mc == 0 || mc > 255 || n < (1u << 16) || n >= (1ull << 32) - 1) return FMX_ERR_UNSUPPORTED;
    if (text[0] == 0 || text[n - 1] != 0 || text[n - 2] == 0) return FMX_ERR_UNSUPPORTED;  // invalid texts: the host builder reports
    bool interior_zero = false, bad_char = false;
    {
        int iz = 0, bad = 0;
#pragma omp parallel for schedule(static) reduction(| : iz, bad)
        for (int64_t i = 0; i < (int64_t)n - 1; i++) {
            iz |= text[i] == 0;
            bad |= text[i] > mc;
        }
        interior_zero = iz != 0;
        bad_char = bad != 0;
    }
    if (bad_char) return FMX_ERR_UNSUPPORTED;
    if (int mrc = resolve_mode(mode, err)) return mrc;
    const uint32_t L = log2_u64(mc) + 1, cs_len = (uint32_t)mc + 1;
    const VerifyPlan vp = plan_verify(FMX_KIND_RLFM, n, mode, level, interior_zero, 0, false);

    int rc = FMX_ERR_CUDA;
    uint8_t *d_text = nullptr, *d_bwt = nullptr, *d_flag = nullptr, *d_heads = nullptr, *d_heads_sorted = nullptr, *blob = nullptr;
    uint32_t *d_sa = nullptr, *d_ridx = nullptr, *d_starts = nullptr, *d_order = nullptr, *d_order2 = nullptr, *d_len = nullptr,
             *d_cnt = nullptr, *d_rows = nullptr;
    unsigned *d_count = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    const unsigned T = 256;
    auto grid = [&](uint64_t m) { return (unsigned)((m + T - 1) / T); };
    FmxBlobHeader hdr;
    std::memset(&hdr, 0, sizeof(hdr));
    auto at = [&](int k) { return blob + hdr.sec[k].offset; };
    const uint64_t nblk_n = n / FMX_RB_BITS + 1;
    uint64_t r = 0, nblk = 0, cnt_entries = 0;
    uint32_t last_idx = 0;
    uint8_t last_flag = 0;
    bool use_q4 = false, use_sym = false;
    std::vector<uint8_t> heads;
    std::vector<uint32_t> cs(cs_len + 1, 0), adj(cs_len, 0), rows;
    std::vector<uint64_t> occ(cs_len, 0);
    uint64_t bytes[SEC_COUNT];
    unsigned nz = 0;
    const uint32_t n32 = (uint32_t)n;

    GB_TRY(cudaSetDevice(device));
    GB_TRY(cudaMalloc(&d_text, n));
    GB_TRY(cudaMemcpy(d_text, text, n, cudaMemcpyHostToDevice));
    {
        std::string serr;
        int src = gpu_suffix_array_device(d_text, n, L, device, &d_sa, nullptr, serr);
        if (src) {
            err = "GPU suffix array construction failed: " + serr;
            rc = src;
            goto fail;
        }
    }
    GB_TRY(cudaMalloc(&d_bwt, n + 16));
    GB_TRY(cudaMemset(d_bwt + n, 0, 16));
    GB_TRY(cudaMalloc(&d_flag, n));
    GB_TRY(cudaMalloc(&d_ridx, n * 4));
    // rlfmi.rs:49-53 takes text[n - 1] where fm_index.rs takes 0: the same \0 for every valid text
    k_bwt<<<grid(n), T>>>(d_text, d_sa, n, d_bwt);
    k_run_flags<<<grid(n), T>>>(d_bwt, n, d_flag, d_ridx);
    GB_TRY(cudaGetLastError());
    GB_TRY(cudaMemcpy(&last_flag, d_flag + n - 1, 1, cudaMemcpyDeviceToHost));
    GB_TRY((scan3::run<uint32_t, scan3::Sum32, true>(d_ridx, n, d_ridx, scan3::Sum32(), 0u, 0)));
    GB_TRY(cudaMemcpy(&last_idx, d_ridx + n - 1, 4, cudaMemcpyDeviceToHost));
    r = (uint64_t)last_idx + last_flag;
    if (r == 0) {
        rc = FMX_ERR_UNSUPPORTED;
        goto fail;
    }
    GB_TRY(cudaMalloc(&d_starts, r * 4));
    GB_TRY(cudaMalloc(&d_heads, r + 16));
    GB_TRY(cudaMemset(d_heads + r, 0, 16));
    GB_TRY(cudaMalloc(&d_heads_sorted, r));
    GB_TRY(cudaMalloc(&d_order, r * 4));
    GB_TRY(cudaMalloc(&d_order2, r * 4));
    GB_TRY(cudaMalloc(&d_len, r * 4));
    k_runs_scatter<<<grid(n), T>>>(d_flag, d_ridx, d_bwt, n, d_starts, d_heads, d_order);
    GB_TRY(cudaGetLastError());
    heads.resize(r);
    GB_TRY(cudaMemcpy(heads.data(), d_heads, r, cudaMemcpyDeviceToHost));
    if (heads[0] == 0) {  // the reference hits unreachable!() (rlfmi.rs:62); cannot happen for a valid text
        rc = FMX_ERR_UNSUPPORTED;
        goto fail;
    }
    for (uint64_t j = 0; j < r; j++) occ[heads[j]]++;
    {
        uint64_t sum = 0;
        for (uint32_t c = 0; c < cs_len; c++) {  // cs over the run heads
            cs[c] = (uint32_t)sum;
            sum += occ[c];
        }
        cs[cs_len] = (uint32_t)r;
    }
    use_q4 = mc <= 4 && !q4_forbidden_by_env() && occ[0] <= FMX_MAX_EXC;
    use_sym = !use_q4 && sym_layout_chosen(cs_len, r);
    if (!use_q4 && !use_sym) {
        rc = FMX_ERR_UNSUPPORTED;
        goto fail;
    }
    nblk = use_q4 ? r / 64 + 1 : r / FMX_RB_BITS + 1;
    cnt_entries = use_q4 ? nblk * 4 : nblk * cs_len;
    if (cnt_entries < nblk_n) cnt_entries = nblk_n;  // the same buffer serves the block counts of b and bp

    hdr.magic = FMX_BLOB_MAGIC;
    hdr.version = FMX_BLOB_VERSION;
    hdr.kind = FMX_KIND_RLFM;
    hdr.n = n;
    hdr.seq_len = r;
    hdr.runs = r;
    hdr.levels = L;
    hdr.max_character = (uint32_t)mc;
    hdr.cs_len = cs_len;
    hdr.layout = use_q4 ? FMX_LAYOUT_QUAT : FMX_LAYOUT_SYM;
    hdr.char_width = 1;
    hdr.nexc = use_q4 ? (uint32_t)occ[0] : 0u;
    hdr.sym_nblk = use_sym ? (uint32_t)nblk : 0u;
    hdr.reserved[0] = (uint64_t)mode;
    if (level >= 0) {
        hdr.has_locate = 1;
        uint32_t lvl = (uint32_t)level;
        hdr.sa_word_size = log2_u64(n) + 1;
        if (lvl >= 63 || n <= (1ull << lvl)) lvl = 0;  // sample.rs:28-31
        hdr.sa_level = lvl;
        hdr.sa_count = ((n - 1) >> lvl) + 1;
    }
    hdr.vsa_level = hdr.sa_level;
    std::memset(bytes, 0, sizeof(bytes));
    bytes[SEC_LEVEL0] = use_q4 ? nblk * 32 : sym_layout_bytes(cs_len, r);
    if (use_sym) bytes[SEC_LEVEL0 + 1] = r;
    if (use_q4) bytes[SEC_EXC] = occ[0] * 4;
    bytes[SEC_ADJ] = (uint64_t)cs_len * 4;
    bytes[SEC_CS] = (uint64_t)(cs_len + 1) * 4;
    if (hdr.has_locate) bytes[SEC_SA] = hdr.sa_count * 4;
    bytes[SEC_RL_B] = nblk_n * 32;
    bytes[SEC_RL_BP] = nblk_n * 32;
    bytes[SEC_RL_BSEL] = (r + 1) * 4;
    bytes[SEC_RL_BPSEL] = (r + 1) * 4;
    if (vp.dense_sa) {
        bytes[SEC_VSA] = n * 4;
        bytes[SEC_TEXT] = n;
    }
    layout_sections(hdr, bytes);
    GB_TRY(cudaMalloc(&blob, hdr.total_bytes));
    GB_TRY(cudaMemset(blob, 0, hdr.total_bytes));
    GB_TRY(cudaMalloc(&d_cnt, cnt_entries * 4));

    // b: run starts in L order
    k_rb_pack<<<grid(nblk_n), T>>>(d_flag, n, nblk_n, reinterpret_cast<uint32_t *>(at(SEC_RL_B)), d_cnt);
    GB_TRY(cudaGetLastError());
    GB_TRY((scan3::run<uint32_t, scan3::Sum32, true>(d_cnt, nblk_n, d_cnt, scan3::Sum32(), 0u, 0)));
    k_rb_counts<<<grid(nblk_n), T>>>(reinterpret_cast<uint32_t *>(at(SEC_RL_B)), d_cnt, nblk_n);
    // select1(b, j) = first row of run j; [runs] = n (vers: select past the last one returns len, rlfmi.rs:285-309)
    GB_TRY(cudaMemcpy(at(SEC_RL_BSEL), d_starts, r * 4, cudaMemcpyDeviceToDevice));
    GB_TRY(cudaMemcpy(at(SEC_RL_BSEL) + r * 4, &n32, 4, cudaMemcpyHostToDevice));
    // bp: runs grouped by head, stably (rlfmi.rs:70-83) = a stable sort of the runs by their head symbol
    {
        GB_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_heads, d_heads_sorted, d_order, d_order2, (long long)r, 0, 8));
        GB_TRY(cudaMalloc(&d_tmp, tmp_bytes));
        GB_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_heads, d_heads_sorted, d_order, d_order2, (long long)r, 0, 8));
    }
    k_sorted_run_len<<<grid(r), T>>>(d_order2, d_starts, r, n, d_len);
    GB_TRY(cudaGetLastError());
    GB_TRY((scan3::run<uint32_t, scan3::Sum32, true>(d_len, r, d_len, scan3::Sum32(), 0u, 0)));  // position of every sorted run in bp
    GB_TRY(cudaMemcpy(at(SEC_RL_BPSEL), d_len, r * 4, cudaMemcpyDeviceToDevice));
    GB_TRY(cudaMemcpy(at(SEC_RL_BPSEL) + r * 4, &n32, 4, cudaMemcpyHostToDevice));
    GB_TRY(cudaMemset(d_flag, 0, n));
    k_mark_bp<<<grid(r), T>>>(d_len, r, d_flag);
    k_rb_pack<<<grid(nblk_n), T>>>(d_flag, n, nblk_n, reinterpret_cast<uint32_t *>(at(SEC_RL_BP)), d_cnt);
    GB_TRY(cudaGetLastError());
    GB_TRY((scan3::run<uint32_t, scan3::Sum32, true>(d_cnt, nblk_n, d_cnt, scan3::Sum32(), 0u, 0)));
    k_rb_counts<<<grid(nblk_n), T>>>(reinterpret_cast<uint32_t *>(at(SEC_RL_BP)), d_cnt, nblk_n);
    // the run heads' own rank structure
    if (use_q4) {
        k_q4_pack<<<grid(nblk), T>>>(d_heads, r, nblk, reinterpret_cast<uint32_t *>(at(SEC_LEVEL0)), d_cnt);
        GB_TRY(cudaGetLastError());
        for (int c = 0; c < 4; c++)
            GB_TRY((scan3::run<uint32_t, scan3::Sum32, true>(d_cnt + (uint64_t)c * nblk, nblk, d_cnt + (uint64_t)c * nblk, scan3::Sum32(), 0u, 0)));
        k_q4_counts<<<grid(nblk), T>>>(reinterpret_cast<uint32_t *>(at(SEC_LEVEL0)), d_cnt, nblk);
        if (occ[0]) {
            GB_TRY(cudaMalloc(&d_rows, FMX_MAX_EXC * 4));
            GB_TRY(cudaMalloc(&d_count, 4));
            GB_TRY(cudaMemset(d_count, 0, 4));
            k_zero_rows<<<grid(r), T>>>(d_heads, r, d_rows, FMX_MAX_EXC, d_count);
            GB_TRY(cudaGetLastError());
            GB_TRY(cudaMemcpy(&nz, d_count, 4, cudaMemcpyDeviceToHost));
            if (nz != occ[0]) {
                err = "GPU index build: run heads hold another number of zeros than counted";
                rc = FMX_ERR_CUDA;
                goto fail;
            }
            rows.resize(nz);
            GB_TRY(cudaMemcpy(rows.data(), d_rows, (uint64_t)nz * 4, cudaMemcpyDeviceToHost));
            std::sort(rows.begin(), rows.end());
            GB_TRY(cudaMemcpy(at(SEC_EXC), rows.data(), (uint64_t)nz * 4, cudaMemcpyHostToDevice));
        }
    } else {
        const uint64_t sym_entries = nblk * cs_len;
        k_sym_pack<<<(unsigned)((nblk + 7) / 8), 256>>>(d_heads, r, nblk, cs_len, reinterpret_cast<uint32_t *>(at(SEC_LEVEL0)), d_cnt);
        GB_TRY(cudaGetLastError());
        GB_TRY((scan3::run<uint32_t, scan3::Sum32, true>(d_cnt, sym_entries, d_cnt, scan3::Sum32(), 0u, 0)));
        k_sym_counts<<<grid(sym_entries), T>>>(reinterpret_cast<uint32_t *>(at(SEC_LEVEL0)), d_cnt, nblk, sym_entries);
        GB_TRY(cudaMemcpy(at(SEC_LEVEL0 + 1), d_heads, r, cudaMemcpyDeviceToDevice));
    }
    GB_TRY(cudaGetLastError());
    GB_TRY(cudaMemcpy(at(SEC_ADJ), adj.data(), adj.size() * 4, cudaMemcpyHostToDevice));
    GB_TRY(cudaMemcpy(at(SEC_CS), cs.data(), cs.size() * 4, cudaMemcpyHostToDevice));
    if (hdr.has_locate) k_samples<<<grid(hdr.sa_count), T>>>(d_sa, hdr.sa_count, hdr.sa_level, reinterpret_cast<uint32_t *>(at(SEC_SA)));
    if (vp.dense_sa) {
        GB_TRY(cudaMemcpy(at(SEC_VSA), d_sa, n * 4, cudaMemcpyDeviceToDevice));
        GB_TRY(cudaMemcpy(at(SEC_TEXT), d_text, n, cudaMemcpyDeviceToDevice));
    }
    GB_TRY(cudaGetLastError());
    GB_TRY(cudaMemcpy(blob, &hdr, sizeof(hdr), cudaMemcpyHostToDevice));
    GB_TRY(cudaDeviceSynchronize());
    *d_blob_out = blob;
    *hdr_out = hdr;
    blob = nullptr;
    rc = 0;
fail:
    cudaFree(d_text);
    cudaFree(d_bwt);
    cudaFree(d_flag);
    cudaFree(d_heads);
    cudaFree(d_heads_sorted);
    cudaFree(d_sa);
    cudaFree(d_ridx);
    cudaFree(d_starts);
    cudaFree(d_order);
    cudaFree(d_order2);
    cudaFree(d_len);
    cudaFree(d_cnt);
    cudaFree(d_rows);
    cudaFree(d_count);
    cudaFree(d_tmp);
    cudaFree(blob);
    if (rc) cudaGetLastError();
    return rc;
}

}  // namespace fmx
