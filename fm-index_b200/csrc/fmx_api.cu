// fmx_api.cu -- the C ABI (include/fmx.h) over the CUDA kernels.  No CPU fallback anywhere:
// every query entry point launches kernels on the index's GPU or fails with FMX_ERR_CUDA.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/fmx.h"
#include "builder.h"
#include "kernels.cuh"
#include "phased.cuh"

using namespace fmx;

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(FMX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
    } while (0)
#define LAUNCH_CHECK()                                                                             \
    do {                                                                                           \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                        \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess) return fail(FMX_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    } while (0)

// grow-only device scratch buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(FMX_ERR_OOM, std::string("cudaMalloc scratch: ") + cudaGetErrorString(e));
        }
        cap = want;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

enum { B_QUEUE, B_BUCKET, B_BHIST, B_ORDER, B_PAT, B_OFF, B_S, B_E, B_INS, B_INE, B_CNT, B_HOFF, B_OWNER, B_ROWS, B_ROWS2, B_FLAG, B_FPOS, B_TILES, B_POS, B_PID,
       B_OUT8, B_OUT32, B_RS, B_RE, B_HINT, B_QS, B_QV, B_QN, B_BIGQ, B_W64, B_COUNT };

// one lane of the chunked H2D / kernels / D2H pipeline (fmx_search_locate_batch)
#define FMX_LANES 4
struct Lane {
    DevBuf buf[B_COUNT];
    cudaStream_t st = nullptr;
    cudaEvent_t ev = nullptr;
    uint64_t *h_total = nullptr;  // pinned
    uint64_t pos_cap = 0;         // entries the lane's position buffers were sized for when the chunk was emitted
};

#define FMX_NPHASE 6  /* seed, steps, verify, steps2 (+ widen), offsets (counts + scan), emit / locate */

// what fmx_locate_count_device left behind for the fmx_locate_fill_device call that completes it
struct CountToken {
    bool valid = false;
    int prefix_only = 0;
    uint64_t npat = 0, total = 0;
    const uint64_t *d_s = nullptr, *d_hit_off = nullptr;
    const uint32_t *rows = nullptr;
    bool ranges = false;
};

struct fmx_index {
    FmxBlobHeader hdr;
    FmxDev dev;
    void *d_blob = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t *d_err = nullptr;             // [0] pattern-char error flag
    unsigned long long *d_work = nullptr;  // [0] search iterations, [1] LF steps
    uint2 *d_kmer_tab = nullptr;           // memoised first kmer_k search iterations (see SearchArgs)
    uint8_t *d_kmer_steps = nullptr;
    uint32_t kmer_k = 0;
    uint64_t kmer_entries = 0;
    uint4 *d_big_tab4 = nullptr;           // the large table with 16-byte entries (text context of one-row entries); replaces d_big_tab
    int opt_table_ctx = 1;                 // 0: keep 8-byte entries (A/B)
    uint2 *d_big_tab = nullptr;            // the large (HBM-resident) table
    uint8_t *d_big_steps = nullptr;
    uint32_t big_k = 0;
    uint64_t big_entries = 0;
    int opt_kmer_big = 1;
    int sms = 148;
    int persist_blocks_per_sm = 4;
    // tuning knobs (fmx_index_set_option)
    int opt_persistent = 0;  // 1: persistent refill search kernels instead of one pattern per thread
    int opt_kmer = 1;
    uint64_t opt_pipeline_chunk = 0;  // patterns per pipeline chunk (0 = automatic)
    int opt_verify = 1;               // 0: never take the seed-and-verify tail (A/B)
    int opt_stage_patterns = 1;       // 0: never stage pattern bytes in shared memory (A/B)
    int opt_locate_ranges = -1;       // -1 auto (RLFM with >= 8 matches per pattern), 0 never, 1 always: k_locate_ranges
    int opt_locate_expand = 0;        // 0 auto, 1 always expand the rows first, 2 always binary-search in k_locate
    int opt_locate_refill = 0;        // 1: per-lane refill k_locate (lost the A/B: it breaks the coalescing of adjacent rows)
    int opt_bucket = 0;               // 1: visit the batch in k-mer bucket order (lost the A/B, kept for experiments); -1 auto
    int opt_count_work = 0;           // 1: kernels count executed search iterations / LF steps (fmx_last_work)
    int opt_phased = 2;               // dense verify structures: 2 = k_query_fused (one kernel), 1 = the three phased kernels, 0 = k_search
    int opt_order_by_length = 0;      // 1: ragged batches through the fused kernel in order of pattern length (A/B: lost, see search_phased)
    int opt_fused_defer = 1;          // fused kernel with the block-local second pass when the 16-byte table is in use (phased.cuh)
    int opt_emit_fused = FMX_EMIT_FUSED_DEFAULT;  // rich locate: offsets + positions in one pass over the ranges (k_offsets_emit); FMX_EMIT_FUSED=0|1 at construction
    int opt_query_blocks = 0;         // > 0: at most this many blocks for the fused query kernels (tests: many rounds per block on small batches)
    int opt_extract_text = 1;         // 0: extraction by LF / FL steps even when text and suffix array are resident (A/B)
    int opt_locate_dense = 1;         // 0: LF walks to the samples even when the full suffix array is resident (A/B)
    uint32_t tab_embed = 0;           // one-row k-mer table entries carry the row's text position (SearchArgs::tab_embed)
    mutable DevBuf buf[B_COUNT];
    mutable Lane lane[FMX_LANES];
    mutable bool lanes_ready = false;
    mutable std::mutex mu;
    mutable CountToken count_token;
    // option "phase_timing": CUDA events between the phases of a query (fmx_last_phase_ms)
    int opt_phase_timing = 0;
    mutable cudaEvent_t pev[FMX_NPHASE + 1] = {};
    mutable uint32_t pev_mask = 0;
};

static bool env_flag(const char *name) {
    const char *v = std::getenv(name);
    return v && v[0] && v[0] != '0';
}

// bytes per character of patterns and extracted characters (1 for every u8 index)
static inline uint64_t char_width_of(const fmx_index *idx) { return 1ull << idx->dev.cw_shift; }

static int build_kmer_table(fmx_index *idx);
static int build_big_table(fmx_index *idx, uint64_t budget_bytes);

static inline cudaStream_t pick_stream(const fmx_index *idx, void *stream) {
    return stream ? reinterpret_cast<cudaStream_t>(stream) : idx->stream;
}
static inline unsigned grid_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }
// phase boundary k (0 = start of the query) on the query's stream; only with option "phase_timing"
static void phase_mark(const fmx_index *idx, int k, cudaStream_t st) {
    if (!idx->opt_phase_timing) return;
    if (!idx->pev[k] && cudaEventCreate(&idx->pev[k]) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    if (k == 0) idx->pev_mask = 0;
    if (cudaEventRecord(idx->pev[k], st) == cudaSuccess) idx->pev_mask |= 1u << k;
    else cudaGetLastError();
}
// the full suffix array is resident (SEC_VSA): HBM-rich locate, position = SA[row]
static inline bool has_dense_sa(const fmx_index *idx) { return idx->hdr.sec[SEC_VSA].bytes != 0 && idx->dev.vsa != nullptr; }
static inline bool locate_dense(const fmx_index *idx) { return has_dense_sa(idx) && idx->opt_locate_dense && idx->hdr.has_locate; }

// run f(K, LY) with the index's (kind, layout) as compile-time constants
template <class F>
static void dispatch(const fmx_index *idx, F &&f) {
    using std::integral_constant;
    switch (idx->hdr.kind * 8 + idx->hdr.layout) {
        case 0: f(integral_constant<int, 0>{}, integral_constant<int, 0>{}); break;
        case 1: f(integral_constant<int, 0>{}, integral_constant<int, 1>{}); break;
        case 2: f(integral_constant<int, 0>{}, integral_constant<int, 2>{}); break;
        case 3: f(integral_constant<int, 0>{}, integral_constant<int, 3>{}); break;
        case 4: f(integral_constant<int, 0>{}, integral_constant<int, 4>{}); break;
        case 8: f(integral_constant<int, 1>{}, integral_constant<int, 0>{}); break;
        case 9: f(integral_constant<int, 1>{}, integral_constant<int, 1>{}); break;
        case 10: f(integral_constant<int, 1>{}, integral_constant<int, 2>{}); break;
        case 11: f(integral_constant<int, 1>{}, integral_constant<int, 3>{}); break;
        case 12: f(integral_constant<int, 1>{}, integral_constant<int, 4>{}); break;
        case 16: f(integral_constant<int, 2>{}, integral_constant<int, 0>{}); break;
        case 17: f(integral_constant<int, 2>{}, integral_constant<int, 1>{}); break;
        case 18: f(integral_constant<int, 2>{}, integral_constant<int, 2>{}); break;
        case 19: f(integral_constant<int, 2>{}, integral_constant<int, 3>{}); break;
        default: f(integral_constant<int, 2>{}, integral_constant<int, 4>{}); break;
    }
}

extern "C" {

const char *fmx_last_error(void) { return g_err.c_str(); }
// fmx_group.cu reports its own errors through the same thread-local message
void fmx_internal_set_error(const char *msg) { g_err = msg ? msg : ""; }
int fmx_version(void) { return 100; }
uint64_t fmx_launch_count(void) { return g_launches.load(); }
void fmx_free(void *p) { std::free(p); }

int fmx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int fmx_build_suffix_array(const void *text, uint64_t n, uint32_t char_width, uint64_t *sa_out) {
    std::string err;
    int rc = char_width == 1 ? build_suffix_array(static_cast<const uint8_t *>(text), n, sa_out, err)
                             : build_suffix_array_wide(text, char_width, n, sa_out, err);
    return rc ? fail(rc, err) : FMX_OK;
}

int fmx_build_suffix_array_device(const void *text, uint64_t n, uint32_t char_width, uint64_t max_character, int device,
                                  uint64_t *sa_out, int *rounds) {
    if (char_width != 1) return fail(FMX_ERR_UNSUPPORTED, "only u8 texts (char_width == 1) are supported");
    if (max_character == 0 || max_character > 255) return fail(FMX_ERR_INVALID_ARG, "max_character must be in 1..=255");
    if ((!text || !sa_out) && n) return fail(FMX_ERR_INVALID_ARG, "null argument");
    const uint8_t *t = static_cast<const uint8_t *>(text);
    for (uint64_t i = 0; i < n; i++)
        if (t[i] > max_character) return fail(FMX_ERR_INVALID_ARG, "text contains a character larger than max_character");
    uint32_t bits = 64 - (uint32_t)__builtin_clzll(max_character);
    std::vector<uint32_t> sa(n);
    std::string err;
    int rc = gpu_suffix_array(t, n, bits, device, sa.data(), rounds, err);
    if (rc) return fail(rc, err);
    for (uint64_t i = 0; i < n; i++) sa_out[i] = sa[i];
    return FMX_OK;
}

// Host blob of a text of any character width (character.rs:38-42).  u8: build_blob.  u16 / u32 / u64 with
// max_character > 255: the WIDE layout (build_blob_wide).  Wide characters over a small alphabet (max_character <= 255):
// the text is narrowed and takes the u8 layouts; only the header's char_width -- the width of patterns and extracted
// characters at the ABI -- differs, and the index is a compact one (no tables, no verify structures: kernels.cuh AnyReader).
static int build_blob_any(const void *text, uint64_t n, uint32_t char_width, uint64_t max_character, int kind, int level,
                          HostBlob &b, std::string &err, int sa_device, int mode) {
    if (char_width == 1) return build_blob(static_cast<const uint8_t *>(text), n, max_character, kind, level, b, err, sa_device, mode);
    if (char_width != 2 && char_width != 4 && char_width != 8) {
        err = "character width must be 1, 2, 4 or 8 bytes";
        return FMX_ERR_INVALID_ARG;
    }
    if (max_character > 255) return build_blob_wide(text, char_width, n, max_character, kind, level, b, err, mode);
    std::vector<uint8_t> narrow(n);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const uint64_t v = char_width == 2 ? static_cast<const uint16_t *>(text)[i]
                         : char_width == 4 ? static_cast<const uint32_t *>(text)[i] : static_cast<const uint64_t *>(text)[i];
        bad |= v > max_character;
        narrow[(size_t)i] = (uint8_t)v;
    }
    if (bad) {
        err = "text contains a character larger than max_character";
        return FMX_ERR_INVALID_ARG;
    }
    if (int mrc = resolve_mode(mode, err)) return mrc;
    int rc = build_blob(narrow.data(), n, max_character, kind, level, b, err, sa_device, FMX_MODE_COMPACT);
    if (rc) return rc;
    reinterpret_cast<FmxBlobHeader *>(b.p)->char_width = char_width;
    return 0;
}

int fmx_blob_build_ex(const void *text, uint64_t n, uint32_t char_width, uint64_t max_character, int kind, int level,
                      int mode, void **blob, uint64_t *blob_bytes) {
    if (!blob || !blob_bytes || (!text && n)) return fail(FMX_ERR_INVALID_ARG, "null argument");
    set_device_memory_hint(0);  // a host-only build knows of no device
    HostBlob b;
    std::string err;
    int rc = build_blob_any(text, n, char_width, max_character, kind, level, b, err, -1, mode);
    if (rc) return fail(rc, err);
    *blob_bytes = b.n;
    *blob = b.release();  // malloc'd: the caller frees it with fmx_free
    return FMX_OK;
}

int fmx_blob_build(const void *text, uint64_t n, uint32_t char_width, uint64_t max_character, int kind, int level,
                   void **blob, uint64_t *blob_bytes) {
    return fmx_blob_build_ex(text, n, char_width, max_character, kind, level, FMX_MODE_AUTO, blob, blob_bytes);
}

// uploads a blob the caller owns (nothing of it is kept on the host: save() reads the device copy back)
// the kernels' view of the device-resident blob (FmxDev) from its header
static void bind_sections(fmx_index *idx) {
    const FmxBlobHeader &hdr = idx->hdr;
    const int device = idx->device;
    const char *base = static_cast<const char *>(idx->d_blob);
    auto sec = [&](int k) -> const void * { return hdr.sec[k].bytes ? base + hdr.sec[k].offset : nullptr; };
    FmxDev &d = idx->dev;
    std::memset(&d, 0, sizeof(d));
    for (uint32_t l = 0; l < hdr.levels && l < FMX_MAX_LEVELS && hdr.layout != FMX_LAYOUT_WIDE; l++) {
        d.lv[l] = static_cast<const uint4 *>(sec(SEC_LEVEL0 + l));
        d.zeros[l] = (uint32_t)hdr.zeros[l];
    }
    if (hdr.layout == FMX_LAYOUT_WIDE) {  // all levels in SEC_LEVEL0, zeros per level in SEC_WZEROS
        d.lv[0] = static_cast<const uint4 *>(sec(SEC_LEVEL0));
        d.wzeros = static_cast<const uint32_t *>(sec(SEC_WZEROS));
        d.wide_nblk = (uint32_t)(hdr.seq_len / FMX_RB_BITS + 1);
    }
    {
        const uint32_t cw = hdr.char_width ? hdr.char_width : 1u;
        d.cw_shift = cw == 8 ? 3u : (cw == 4 ? 2u : (cw == 2 ? 1u : 0u));
    }
    if (hdr.layout == FMX_LAYOUT_SYM) {
        d.lv[0] = static_cast<const uint4 *>(sec(SEC_LEVEL0));
        d.raw = static_cast<const uint8_t *>(sec(SEC_LEVEL0 + 1));
        d.sym_nblk = hdr.sym_nblk;
    }
    d.qlevels = hdr.qlevels;
    for (uint32_t l = 0; l < hdr.qlevels; l++)
        for (int g = 0; g < 4; g++) d.qoff[l * 4 + g] = (uint32_t)hdr.qoff[l][g];
    d.adj = static_cast<const uint32_t *>(sec(SEC_ADJ));
    d.cs = static_cast<const uint32_t *>(sec(SEC_CS));
    d.sa = static_cast<const uint32_t *>(sec(SEC_SA));
    d.doc = static_cast<const uint32_t *>(sec(SEC_DOC));
    d.piece_end = static_cast<const uint32_t *>(sec(SEC_PIECE_END));
    d.rl_b = static_cast<const uint4 *>(sec(SEC_RL_B));
    d.rl_bp = static_cast<const uint4 *>(sec(SEC_RL_BP));
    d.rl_bsel = static_cast<const uint32_t *>(sec(SEC_RL_BSEL));
    d.rl_bpsel = static_cast<const uint32_t *>(sec(SEC_RL_BPSEL));
    d.exc = static_cast<const uint32_t *>(sec(SEC_EXC));
    d.text = static_cast<const uint8_t *>(sec(SEC_TEXT));
    d.isa = static_cast<const uint32_t *>(sec(SEC_ISA));
    d.vsa = hdr.sec[SEC_VSA].bytes ? static_cast<const uint32_t *>(sec(SEC_VSA)) : d.sa;
    d.vsa_level = hdr.sec[SEC_VSA].bytes ? 0u : hdr.vsa_level;
    d.verify = hdr.verify && d.text && d.isa && d.vsa ? 1u : 0u;
    idx->tab_embed = (d.verify && hdr.sec[SEC_VSA].bytes && hdr.isa_level == 0 && hdr.n < (1ull << 31)) ? 1u : 0u;
    d.isa_level = hdr.isa_level;
    d.layout = hdr.layout;
    d.nexc = hdr.nexc;
    d.n = (uint32_t)hdr.n;
    d.seq_len = (uint32_t)hdr.seq_len;
    d.levels = hdr.levels;
    d.max_character = hdr.max_character;
    d.cs_len = hdr.cs_len;
    d.kind = hdr.kind;
    d.has_locate = hdr.has_locate;
    d.sa_level = hdr.sa_level;
    d.ndoc = (uint32_t)hdr.ndoc;
    d.first_row = (uint32_t)hdr.first_row;
    d.runs = (uint32_t)hdr.runs;
    cudaDeviceGetAttribute(&idx->sms, cudaDevAttrMultiProcessorCount, device);
    dispatch(idx, [&](auto K, auto LY) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&idx->persist_blocks_per_sm, k_search_steps<K(), LY()>, 256, 0);
    });
    if (idx->persist_blocks_per_sm < 1) idx->persist_blocks_per_sm = 1;
}

// common tail of every constructor: the blob is resident in idx->d_blob
static int finish_index(fmx_index *idx, fmx_index **out) {
    cudaError_t e = cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&idx->d_err, 256);
    if (e == cudaSuccess) e = cudaMemset(idx->d_err, 0, 256);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fmx_index_free(idx);
        return fail(FMX_ERR_CUDA, std::string("index setup: ") + cudaGetErrorString(e));
    }
    idx->d_work = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(idx->d_err) + 64);
    if (const char *ef = std::getenv("FMX_EMIT_FUSED")) idx->opt_emit_fused = ef[0] && ef[0] != '0' ? 1 : 0;
    bind_sections(idx);
    int rc = build_kmer_table(idx);
    if (rc) {
        fmx_index_free(idx);
        return rc;
    }
    *out = idx;
    return FMX_OK;
}

// uploads a blob the caller owns (nothing of it is kept on the host: save() reads the device copy back)
static int upload(const uint8_t *blob, uint64_t blob_bytes, int device, fmx_index **out) {
    FmxBlobHeader hdr;
    std::string err;
    int rc = check_blob(blob, blob_bytes, hdr, err);
    if (rc) return fail(rc, err);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(FMX_ERR_CUDA, "no CUDA device available (the engine has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(FMX_ERR_INVALID_ARG, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device));
    fmx_index *idx = new fmx_index();
    idx->hdr = hdr;
    idx->device = device;
    cudaError_t e = cudaMalloc(&idx->d_blob, blob_bytes);
    if (e != cudaSuccess) {
        delete idx;
        cudaGetLastError();
        return fail(FMX_ERR_OOM, std::string("cudaMalloc index: ") + cudaGetErrorString(e));
    }
    e = cudaMemcpy(idx->d_blob, blob, blob_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(idx->d_blob);
        delete idx;
        return fail(FMX_ERR_CUDA, std::string("index upload: ") + cudaGetErrorString(e));
    }
    return finish_index(idx, out);
}

// takes over a blob that was built in device memory (gpu_build.cu)
static int adopt(const FmxBlobHeader &hdr, void *d_blob, int device, fmx_index **out) {
    CUDA_TRY(cudaSetDevice(device));
    fmx_index *idx = new fmx_index();
    idx->hdr = hdr;
    idx->device = device;
    idx->d_blob = d_blob;
    return finish_index(idx, out);
}

// A copy of an index on another device: the blob and the k-mer tables travel device to device (NVLink peer copies
// when the GPUs are peers), nothing is rebuilt.
int fmx_index_clone(const fmx_index *src, int device, fmx_index **out) {
    if (!src || !out) return fail(FMX_ERR_INVALID_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(FMX_ERR_INVALID_ARG, "device ordinal out of range");
    }
    CUDA_TRY(cudaSetDevice(device));
    fmx_index *idx = new fmx_index();
    idx->hdr = src->hdr;
    idx->device = device;
    auto dup = [&](const void *p, size_t bytes, void **q) -> cudaError_t {
        *q = nullptr;
        if (!p || !bytes) return cudaSuccess;
        cudaError_t e = cudaMalloc(q, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpyPeer(*q, device, p, src->device, bytes);
    };
    cudaError_t e = dup(src->d_blob, src->hdr.total_bytes, &idx->d_blob);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&idx->d_err, 256);
    if (e == cudaSuccess) e = cudaMemset(idx->d_err, 0, 256);
    if (e == cudaSuccess) e = dup(src->d_kmer_tab, src->kmer_entries * sizeof(uint2), reinterpret_cast<void **>(&idx->d_kmer_tab));
    if (e == cudaSuccess) e = dup(src->d_kmer_steps, src->kmer_entries, reinterpret_cast<void **>(&idx->d_kmer_steps));
    if (e == cudaSuccess) e = dup(src->d_big_tab, src->big_entries * sizeof(uint2), reinterpret_cast<void **>(&idx->d_big_tab));
    if (e == cudaSuccess) e = dup(src->d_big_tab4, src->d_big_tab4 ? src->big_entries * sizeof(uint4) : 0, reinterpret_cast<void **>(&idx->d_big_tab4));
    if (e == cudaSuccess) e = dup(src->d_big_steps, src->big_entries, reinterpret_cast<void **>(&idx->d_big_steps));
    if (e != cudaSuccess) {
        cudaGetLastError();
        fmx_index_free(idx);
        return fail(e == cudaErrorMemoryAllocation ? FMX_ERR_OOM : FMX_ERR_CUDA, std::string("index clone: ") + cudaGetErrorString(e));
    }
    idx->d_work = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(idx->d_err) + 64);
    idx->kmer_k = src->kmer_k;
    idx->kmer_entries = src->kmer_entries;
    idx->big_k = src->big_k;
    idx->big_entries = src->big_entries;
    bind_sections(idx);
    *out = idx;
    return FMX_OK;
}

int fmx_index_from_blob(const void *blob, uint64_t blob_bytes, int device, fmx_index **out) {
    if (!blob || !out) return fail(FMX_ERR_INVALID_ARG, "null argument");
    return upload(static_cast<const uint8_t *>(blob), blob_bytes, device, out);
}

int fmx_index_build(const void *text, uint64_t n, uint32_t char_width, uint64_t max_character, int kind, int level,
                    int device, fmx_index **out) {
    return fmx_index_build_ex(text, n, char_width, max_character, kind, level, device, FMX_MODE_AUTO, out);
}

int fmx_index_mode_of(const fmx_index *idx) {
    if (!idx) return -1;
    return (idx->hdr.verify || idx->hdr.sec[SEC_VSA].bytes) ? FMX_MODE_RICH : FMX_MODE_COMPACT;
}

int fmx_index_build_ex(const void *text, uint64_t n, uint32_t char_width, uint64_t max_character, int kind, int level,
                       int device, int mode, fmx_index **out) {
    if (!out || (!text && n)) return fail(FMX_ERR_INVALID_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(FMX_ERR_CUDA, "no CUDA device available (the engine has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(FMX_ERR_INVALID_ARG, "device ordinal out of range");
    {   // the layout budgets follow the device's free memory (a B200's 180 GB leaves them at their defaults)
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaSetDevice(device));
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) set_device_memory_hint(free_b);
        else cudaGetLastError();
    }
    std::string err;
    if (char_width == 1 && !env_flag("FMX_HOST_BUILD") && !env_flag("FMX_HOST_SA")) {
        // Q4 and SYM layouts of FM / MultiPieces indexes: the whole blob is built in device memory (gpu_build.cu)
        void *d_blob = nullptr;
        FmxBlobHeader hdr;
        int grc = kind == FMX_KIND_RLFM
                      ? gpu_build_rlfm_blob(static_cast<const uint8_t *>(text), n, max_character, level, mode, device, &d_blob, &hdr, err)
                      : gpu_build_blob(static_cast<const uint8_t *>(text), n, max_character, kind, level, mode, device, &d_blob, &hdr, err);
        if (grc == 0) return adopt(hdr, d_blob, device, out);
        // not its case, or not enough free HBM for the construction's working set (32 bytes per symbol while sorting):
        // the host builder takes over
        if (grc != FMX_ERR_UNSUPPORTED && grc != FMX_ERR_OOM) return fail(grc, err);
        err.clear();
    }
    HostBlob b;
    int rc = build_blob_any(text, n, char_width, max_character, kind, level, b, err, device, mode);
    if (rc) return fail(rc, err);
    return upload(b.p, b.n, device, out);
}

int fmx_index_save(const fmx_index *idx, const char *path) {
    if (!idx || !path) return fail(FMX_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(idx->device));
    FILE *f = std::fopen(path, "wb");
    if (!f) return fail(FMX_ERR_IO, std::string("cannot open ") + path);
    // the blob lives on the device only: stream it back through a bounded staging buffer
    const uint64_t total = idx->hdr.total_bytes, piece = 256ull << 20;
    std::vector<uint8_t> stage((size_t)(total < piece ? total : piece));
    bool ok = true;
    for (uint64_t off = 0; off < total && ok; off += piece) {
        const uint64_t len = total - off < piece ? total - off : piece;
        if (cudaMemcpy(stage.data(), static_cast<const uint8_t *>(idx->d_blob) + off, len, cudaMemcpyDeviceToHost) != cudaSuccess) {
            std::fclose(f);
            cudaGetLastError();
            return fail(FMX_ERR_CUDA, "index save: device to host copy failed");
        }
        ok = std::fwrite(stage.data(), 1, len, f) == len;
    }
    std::fclose(f);
    return ok ? FMX_OK : fail(FMX_ERR_IO, "short write");
}

int fmx_index_load(const char *path, int device, fmx_index **out) {
    if (!path || !out) return fail(FMX_ERR_INVALID_ARG, "null argument");
    FILE *f = std::fopen(path, "rb");
    if (!f) return fail(FMX_ERR_IO, std::string("cannot open ") + path);
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    HostBlob b;
    if (b.alloc((uint64_t)(sz > 0 ? sz : 0), false)) {
        std::fclose(f);
        return fail(FMX_ERR_OOM, "out of host memory reading the index file");
    }
    size_t r = std::fread(b.p, 1, b.n, f);
    std::fclose(f);
    if (r != b.n) return fail(FMX_ERR_IO, "short read");
    return upload(b.p, b.n, device, out);
}

void fmx_index_free(fmx_index *idx) {
    if (!idx) return;
    cudaSetDevice(idx->device);
    for (auto &b : idx->buf) b.release();
    for (auto &l : idx->lane) {
        for (auto &b : l.buf) b.release();
        if (l.st) cudaStreamDestroy(l.st);
        if (l.ev) cudaEventDestroy(l.ev);
        if (l.h_total) cudaFreeHost(l.h_total);
    }
    if (idx->d_blob) cudaFree(idx->d_blob);
    if (idx->d_err) cudaFree(idx->d_err);
    if (idx->d_kmer_tab) cudaFree(idx->d_kmer_tab);
    if (idx->d_kmer_steps) cudaFree(idx->d_kmer_steps);
    if (idx->d_big_tab) cudaFree(idx->d_big_tab);
    if (idx->d_big_tab4) cudaFree(idx->d_big_tab4);
    if (idx->d_big_steps) cudaFree(idx->d_big_steps);
    if (idx->stream) cudaStreamDestroy(idx->stream);
    for (auto &e : idx->pev)
        if (e) cudaEventDestroy(e);
    delete idx;
}

int fmx_index_set_option(fmx_index *idx, const char *key, int64_t value) {
    if (!idx || !key) return fail(FMX_ERR_INVALID_ARG, "null argument");
    std::string k(key);
    if (k == "search_persistent") idx->opt_persistent = value != 0;
    else if (k == "kmer") idx->opt_kmer = value != 0;
    else if (k == "kmer_big") idx->opt_kmer_big = value != 0;
    else if (k == "kmer_budget_mb") {  // rebuild the large table with another HBM budget
        CUDA_TRY(cudaSetDevice(idx->device));
        int rc = build_big_table(idx, (uint64_t)(value < 0 ? 0 : value) << 20);
        if (rc) return rc;
    }
    else if (k == "locate_refill") idx->opt_locate_refill = value < 0 || value > 2 ? 0 : (int)value;
    else if (k == "locate_expand") idx->opt_locate_expand = (int)value;
    else if (k == "stage_patterns") idx->opt_stage_patterns = value != 0;
    else if (k == "verify") idx->opt_verify = value != 0;
    else if (k == "locate_ranges") idx->opt_locate_ranges = value < 0 ? -1 : (value != 0);
    else if (k == "bucket") idx->opt_bucket = value < 0 ? -1 : (value != 0);
    else if (k == "count_work") idx->opt_count_work = value != 0;
    else if (k == "search_phased") idx->opt_phased = value < 0 || value > 2 ? 2 : (int)value;
    else if (k == "locate_dense") idx->opt_locate_dense = value != 0;
    else if (k == "extract_text") idx->opt_extract_text = value != 0;
    else if (k == "fused_defer") idx->opt_fused_defer = value < 0 || value > 7 ? 1 : (int)value;
    else if (k == "order_by_length") idx->opt_order_by_length = value > 0 ? 1 : 0;
    else if (k == "emit_fused") idx->opt_emit_fused = value > 0 ? 1 : 0;
    else if (k == "query_blocks") idx->opt_query_blocks = value > 0 && value < (1 << 20) ? (int)value : 0;
    else if (k == "table_ctx") {  // 16-byte table entries on / off: takes effect by rebuilding the large table
        idx->opt_table_ctx = value != 0;
        if (idx->big_entries) {
            CUDA_TRY(cudaSetDevice(idx->device));
            int rc = build_big_table(idx, idx->big_entries * 9);
            if (rc) return rc;
        }
    }
    else if (k == "phase_timing") idx->opt_phase_timing = value != 0;
    else if (k == "pipeline_chunk") idx->opt_pipeline_chunk = value > 0 ? (uint64_t)value : 0;
    else if (k == "l2_fetch_granularity") {
        if (value != 32 && value != 64 && value != 128) return fail(FMX_ERR_INVALID_ARG, "l2_fetch_granularity must be 32, 64 or 128");
        CUDA_TRY(cudaSetDevice(idx->device));
        CUDA_TRY(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
    }
    else if (k == "persist_blocks_per_sm") {
        if (value < 1 || value > 32) return fail(FMX_ERR_INVALID_ARG, "persist_blocks_per_sm out of range");
        idx->persist_blocks_per_sm = (int)value;
    } else return fail(FMX_ERR_INVALID_ARG, "unknown option: " + k);
    return FMX_OK;
}

uint64_t fmx_index_len(const fmx_index *idx) { return idx ? idx->hdr.n : 0; }
uint64_t fmx_index_device_bytes(const fmx_index *idx) {
    return idx ? idx->hdr.total_bytes + idx->kmer_entries * 9 + idx->big_entries * (idx->d_big_tab4 ? 17 : 9) : 0;
}
uint32_t fmx_index_kmer_k(const fmx_index *idx, int big) { return idx ? (big ? idx->big_k : idx->kmer_k) : 0; }
uint32_t fmx_index_kmer_entry_bytes(const fmx_index *idx) { return idx && idx->big_entries ? (idx->d_big_tab4 ? 16u : 8u) : 0u; }
uint64_t fmx_index_pieces_count(const fmx_index *idx) { return idx ? idx->hdr.ndoc : 0; }
int fmx_index_kind(const fmx_index *idx) { return idx ? (int)idx->hdr.kind : -1; }
int fmx_index_has_locate(const fmx_index *idx) { return idx ? (int)idx->hdr.has_locate : 0; }
int fmx_index_device(const fmx_index *idx) { return idx ? idx->device : -1; }
uint32_t fmx_index_wavelet_levels(const fmx_index *idx) { return idx ? idx->hdr.levels : 0; }
uint32_t fmx_index_sample_level(const fmx_index *idx) { return idx && idx->hdr.has_locate ? idx->hdr.sa_level : 0; }
uint32_t fmx_index_layout(const fmx_index *idx) { return idx ? idx->hdr.layout : 0; }
int fmx_index_has_text(const fmx_index *idx) { return idx && idx->dev.text && has_dense_sa(idx) && idx->dev.cw_shift == 0 ? 1 : 0; }
uint32_t fmx_index_char_width(const fmx_index *idx) { return idx ? (idx->hdr.char_width ? idx->hdr.char_width : 1u) : 0; }
uint32_t fmx_index_sectors_per_rank(const fmx_index *idx) {
    if (!idx) return 0;
    if (idx->hdr.layout == FMX_LAYOUT_QUAT || idx->hdr.layout == FMX_LAYOUT_SYM) return 1u;
    return idx->hdr.layout == FMX_LAYOUT_WM4 ? idx->hdr.qlevels : idx->hdr.levels;
}

}  // extern "C"

// ------------------------------------------------------------------ scans

// tiles below this count: each block of the last phase reduces the preceding tile sums itself
#define FMX_SCAN_SELF_CARRY_TILES 4096

// `last` (optional): launches the last phase itself -- (tiles, scanned tile prefixes or null, tile sums or null) as
// k_scan_apply takes them -- so that a consumer of the offsets can be fused into it (k_offsets_emit)
using ScanLast = std::function<int(unsigned, const uint64_t *, const uint64_t *)>;

template <class Load, class Tout, class Op, bool EXCL>
static int device_scan_load(Load in, uint64_t n, Tout *out, Op op, bool write_total, DevBuf &tiles, cudaStream_t st,
                            const ScanLast *last = nullptr) {
    if (n == 0) {
        if (EXCL && write_total) CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(Tout), st));
        return 0;
    }
    uint64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (ntiles == 1) {
        if (last) return (*last)(1u, nullptr, nullptr);
        k_scan_apply<Load, Tout, Op, EXCL><<<1, SCAN_THREADS, 0, st>>>(in, n, nullptr, nullptr, out, op, write_total ? 1 : 0);
        LAUNCH_CHECK();
        return 0;
    }
    // tile sums for every level live in one scratch buffer
    uint64_t need = 0;
    for (uint64_t m = ntiles; m > 1; m = (m + SCAN_TILE - 1) / SCAN_TILE) need += 2 * m;
    need += 2;
    int rc = tiles.ensure(need * sizeof(uint64_t));
    if (rc) return rc;
    uint64_t *sum0 = tiles.as<uint64_t>();
    uint64_t *pre0 = sum0 + ntiles;
    k_scan_reduce<Load, Op><<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(in, n, sum0, op);
    LAUNCH_CHECK();
    if (ntiles <= FMX_SCAN_SELF_CARRY_TILES) {
        if (last) return (*last)((unsigned)ntiles, nullptr, sum0);
        k_scan_apply<Load, Tout, Op, EXCL><<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(in, n, nullptr, sum0, out, op, write_total ? 1 : 0);
        LAUNCH_CHECK();
        return 0;
    }
    // exclusive scan of the tile sums (recursive on the same buffer tail)
    struct Level {
        uint64_t *sum, *pre, n;
    };
    std::vector<Level> lv;
    lv.push_back({sum0, pre0, ntiles});
    uint64_t *cursor = pre0 + ntiles;
    using L64 = LoadPtr<uint64_t>;
    while (lv.back().n > SCAN_TILE) {
        uint64_t m = (lv.back().n + SCAN_TILE - 1) / SCAN_TILE;
        Level nx{cursor, cursor + m, m};
        cursor += 2 * m;
        k_scan_reduce<L64, Op><<<(unsigned)m, SCAN_THREADS, 0, st>>>(L64{lv.back().sum}, lv.back().n, nx.sum, op);
        LAUNCH_CHECK();
        lv.push_back(nx);
    }
    for (size_t k = lv.size(); k-- > 0;) {
        const uint64_t *carry = (k + 1 < lv.size()) ? lv[k + 1].pre : nullptr;
        unsigned g = (unsigned)((lv[k].n + SCAN_TILE - 1) / SCAN_TILE);
        k_scan_apply<L64, uint64_t, Op, true><<<g, SCAN_THREADS, 0, st>>>(L64{lv[k].sum}, lv[k].n, carry, nullptr, lv[k].pre, op, 0);
        LAUNCH_CHECK();
    }
    if (last) return (*last)((unsigned)ntiles, pre0, nullptr);
    k_scan_apply<Load, Tout, Op, EXCL><<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(in, n, pre0, nullptr, out, op, write_total ? 1 : 0);
    LAUNCH_CHECK();
    return 0;
}

template <class Tin, class Tout, class Op, bool EXCL>
static int device_scan(const Tin *in, uint64_t n, Tout *out, Op op, bool write_total, DevBuf &tiles, cudaStream_t st) {
    return device_scan_load<LoadPtr<Tin>, Tout, Op, EXCL>(LoadPtr<Tin>{in}, n, out, op, write_total, tiles, st);
}

// ------------------------------------------------------------------ search

// k-mer bucketing of the batch (see SearchArgs::order): histogram, exclusive scan, scatter
static int bucket_patterns(const fmx_index *idx, DevBuf *buf, SearchArgs &a, cudaStream_t st) {
    const uint64_t nb = idx->kmer_entries + 1;
    int rc;
    if ((rc = buf[B_BUCKET].ensure(a.npat * 4))) return rc;
    if ((rc = buf[B_BHIST].ensure((nb + 1) * 4))) return rc;
    if ((rc = buf[B_ORDER].ensure(a.npat * 4))) return rc;
    uint32_t *d_bucket = buf[B_BUCKET].as<uint32_t>(), *d_hist = buf[B_BHIST].as<uint32_t>();
    CUDA_TRY(cudaMemsetAsync(d_hist, 0, (nb + 1) * 4, st));
    k_bucket_count<<<grid_for(a.npat, 256), 256, 0, st>>>(a, idx->hdr.max_character, (uint32_t)idx->kmer_entries, d_bucket,
                                                          d_hist);
    LAUNCH_CHECK();
    if ((rc = device_scan<uint32_t, uint32_t, OpSum, true>(d_hist, nb, d_hist, OpSum(), false, buf[B_TILES], st))) return rc;
    k_bucket_scatter<<<grid_for(a.npat, 256), 256, 0, st>>>(d_bucket, a.npat, d_hist, buf[B_ORDER].as<uint32_t>());
    LAUNCH_CHECK();
    a.order = buf[B_ORDER].as<uint32_t>();
    return 0;
}

static int dispatch_search(const fmx_index *idx, const SearchArgs &a_in, cudaStream_t st, bool force_simple = false,
                           DevBuf *buf = nullptr) {
    if (a_in.npat == 0) return 0;
    SearchArgs a = a_in;
    if (!buf) buf = idx->buf;
    if (force_simple || !idx->opt_persistent) {
        // big batches over an index that does not fit L2: visit the index in SA order
        const bool big = idx->hdr.total_bytes > (96ull << 20) && a.npat >= (1ull << 18);
        if (a.kmer_tab && a.npat < 0xFFFFFFFFull && !a.steps_out && (idx->opt_bucket == 1 || (idx->opt_bucket < 0 && big))) {
            int rc = bucket_patterns(idx, buf, a, st);
            if (rc) return rc;
        }
        uint64_t blocks = (a.npat + 255) / 256;
        uint64_t cap = (uint64_t)idx->sms * 8 * 8;  // grid-stride beyond 8 waves of 8 CTAs/SM
        if (blocks > cap) blocks = cap;
        dispatch(idx, [&](auto K, auto LY) { k_search<K(), LY()><<<(unsigned)blocks, 256, 0, st>>>(idx->dev, a); });
        LAUNCH_CHECK();
        return 0;
    }
    if (a.npat >= 0xFFFFFFFFull) return fail(FMX_ERR_UNSUPPORTED, "at most 2^32 - 2 patterns per search call");
    int rc = idx->buf[B_QUEUE].ensure(a.npat * 16);
    if (rc) return rc;
    SearchArgs b = a;
    b.queue = idx->buf[B_QUEUE].as<uint4>();
    b.qcount = idx->d_work + 4;
    b.qcursor = idx->d_work + 5;
    CUDA_TRY(cudaMemsetAsync(b.qcount, 0, 2 * sizeof(unsigned long long), st));
    // phase A: one pattern per thread
    uint64_t blocks = (a.npat + 255) / 256;
    uint64_t cap = (uint64_t)idx->sms * 8 * 4;
    if (blocks > cap) blocks = cap;
    dispatch(idx, [&](auto K, auto LY) { k_search_init<K(), LY()><<<(unsigned)blocks, 256, 0, st>>>(idx->dev, b); });
    LAUNCH_CHECK();
    // phase B: persistent, one resident wave of CTAs (a multiple of the SM count) drains the queue
    uint64_t pblocks = (uint64_t)idx->sms * idx->persist_blocks_per_sm;
    uint64_t need = (a.npat + 255) / 256;
    if (pblocks > need) pblocks = need;
    dispatch(idx, [&](auto K, auto LY) { k_search_steps<K(), LY()><<<(unsigned)pblocks, 256, 0, st>>>(idx->dev, b); });
    LAUNCH_CHECK();
    return 0;
}

// Memoise the first k iterations of every fresh search: run the search kernel itself over all
// base^k k-mers (base = max_character: the non-zero symbols), at index upload.  `accel` lets the
// build of the large table start from the small one.
static int build_one_table(fmx_index *idx, uint32_t k, uint64_t entries, bool accel, uint2 **tab_out, uint8_t **steps_out) {
    cudaStream_t st = idx->stream;
    const uint64_t chunk = entries < (1ull << 26) ? entries : (1ull << 26);
    uint8_t *d_pat = nullptr, *d_steps = nullptr;
    uint64_t *d_s = nullptr, *d_e = nullptr;
    uint2 *d_tab = nullptr;
    cudaError_t e = cudaMalloc(&d_tab, entries * sizeof(uint2));
    if (e == cudaSuccess) e = cudaMalloc(&d_steps, entries);
    if (e == cudaSuccess) e = cudaMalloc(&d_pat, chunk * k);
    if (e == cudaSuccess) e = cudaMalloc(&d_s, chunk * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_e, chunk * 8);
    int rc = 0;
    if (e != cudaSuccess) {
        cudaGetLastError();
        rc = fail(FMX_ERR_OOM, std::string("k-mer table: ") + cudaGetErrorString(e));
    }
    for (uint64_t first = 0; rc == 0 && first < entries; first += chunk) {
        const uint64_t m = entries - first < chunk ? entries - first : chunk;
        k_kmer_patterns<<<grid_for(m, 256), 256, 0, st>>>(k, idx->hdr.max_character, first, m, d_pat);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        SearchArgs a;
        std::memset(&a, 0, sizeof(a));
        a.pat = d_pat;
        a.fixed_len = k;
        a.npat = m;
        a.s0 = 0;
        a.e0 = (uint32_t)idx->hdr.n;
        a.out_s = d_s;
        a.out_e = d_e;
        a.err = idx->d_err;
        a.steps_out = d_steps + first;
        if (accel) {
            a.kmer_tab = idx->d_kmer_tab;
            a.kmer_steps = idx->d_kmer_steps;
            a.kmer_k = idx->kmer_k;
            a.tab_embed = idx->tab_embed;  // the small table was embedded when it was built
        }
        rc = dispatch_search(idx, a, st, true);
        if (rc) break;
        k_kmer_pack<<<grid_for(m, 256), 256, 0, st>>>(d_s, d_e, m, d_tab + first);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaStreamSynchronize(st) != cudaSuccess) rc = fail(FMX_ERR_CUDA, "k-mer table build failed");
    }
    cudaFree(d_pat);
    cudaFree(d_s);
    cudaFree(d_e);
    if (rc == 0 && idx->tab_embed) {  // one-row entries: y = FLAG | SA[s] (phased.cuh)
        k_table_embed<<<grid_for(entries, 256), 256, 0, st>>>(d_tab, entries, idx->dev.vsa);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaStreamSynchronize(st) != cudaSuccess) rc = fail(FMX_ERR_CUDA, "k-mer table embedding failed");
    }
    if (rc) {
        cudaFree(d_tab);
        cudaFree(d_steps);
        return rc;
    }
    *tab_out = d_tab;
    *steps_out = d_steps;
    return 0;
}

static void pick_k(uint64_t base, uint64_t max_entries, uint32_t &k, uint64_t &entries) {
    k = 0;
    entries = 1;
    while (base > 1 && entries * base <= max_entries && k < 32) {
        entries *= base;
        k++;
    }
}

// the large table: as many characters as `budget_bytes` of HBM buy (9 bytes per entry)
static int build_big_table(fmx_index *idx, uint64_t budget_bytes) {
    if (idx->d_big_tab) cudaFree(idx->d_big_tab);
    if (idx->d_big_tab4) cudaFree(idx->d_big_tab4);
    if (idx->d_big_steps) cudaFree(idx->d_big_steps);
    idx->d_big_tab = nullptr;
    idx->d_big_tab4 = nullptr;
    idx->d_big_steps = nullptr;
    idx->big_k = 0;
    idx->big_entries = 0;
    if (!idx->d_kmer_tab || budget_bytes == 0) return 0;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && budget_bytes > free_b / 3) budget_bytes = free_b / 3;
    uint64_t max_entries = budget_bytes / 9;
    if (max_entries > (1ull << 32)) max_entries = 1ull << 32;
    uint32_t k;
    uint64_t entries;
    pick_k(idx->hdr.max_character, max_entries, k, entries);
    if (k <= idx->kmer_k) return 0;
    int rc = build_one_table(idx, k, entries, true, &idx->d_big_tab, &idx->d_big_steps);
    if (rc) return rc;
    idx->big_k = k;
    idx->big_entries = entries;
    // DNA-sized alphabets on HBM-rich indexes: 16-byte entries that carry the text in front of one-row ranges, so that a
    // pattern with <= 16 characters left after the table is finished without a text request (kernels.cuh big_tab4).
    // Twice the table's HBM; only when half of what is free still holds it.
    if (idx->tab_embed && idx->opt_table_ctx && !env_flag("FMX_NO_TABLE_CTX") && idx->hdr.max_character <= 4 && idx->dev.text) {
        size_t f2 = 0, t2 = 0;
        const uint64_t wide_b = entries * sizeof(uint4), narrow_b = entries * sizeof(uint2);
        // both forms are resident while the entries are copied over; afterwards at least as much must stay free as the table takes
        // (FMX_TABLE_CTX_FORCE=1 drops the second condition: for callers who know nothing else will be built next to the index)
        if (cudaMemGetInfo(&f2, &t2) == cudaSuccess && wide_b + (8ull << 30) <= f2 &&
            (wide_b <= (f2 + narrow_b) / 2 || env_flag("FMX_TABLE_CTX_FORCE"))) {
            uint4 *t4 = nullptr;
            if (cudaMalloc(&t4, entries * sizeof(uint4)) == cudaSuccess) {
                k_table_widen<<<grid_for(entries, 256), 256, 0, idx->stream>>>(idx->d_big_tab, entries, idx->dev.text, t4);
                g_launches.fetch_add(1, std::memory_order_relaxed);
                if (cudaStreamSynchronize(idx->stream) == cudaSuccess) {
                    cudaFree(idx->d_big_tab);
                    idx->d_big_tab = nullptr;
                    idx->d_big_tab4 = t4;
                } else {
                    cudaGetLastError();
                    cudaFree(t4);
                }
            } else {
                cudaGetLastError();
            }
        } else {
            cudaGetLastError();
        }
    }
    return 0;
}

static int build_kmer_table(fmx_index *idx) {
    if (idx->dev.cw_shift) return 0;  // indexes over wide characters run the plain search kernel (no tables, no verify tail)
    if (env_flag("FMX_NO_KMER") || idx->hdr.n < 2) return 0;
    const bool large_text = idx->hdr.n >= (1ull << 22);
    uint32_t k;
    uint64_t entries;
    pick_k(idx->hdr.max_character, large_text ? (1ull << 21) : (1ull << 12), k, entries);
    if (k < 1) return 0;
    int rc = build_one_table(idx, k, entries, false, &idx->d_kmer_tab, &idx->d_kmer_steps);
    if (rc) return rc;
    idx->kmer_k = k;
    idx->kmer_entries = entries;
    if (!large_text || idx->hdr.reserved[0] == FMX_MODE_COMPACT) return 0;  // COMPACT: the L2-resident table only
    // default HBM budget of the large table (FMX_KMER_BUDGET_MB overrides): twice the index itself, at most
    // 8 GiB, while the rank structure sits in L2.  Beyond L2 every step of a search is a DRAM request, and
    // the table is sized to END the search of an absent pattern: the smallest k with sigma^k >= 4 n leaves
    // n / sigma^k <= 1/4 expected occurrences of a random k-mer (1 GB of DNA: k = 16, 38.7 GB; measured
    // 17.0 -> 11.6 ms per 100 M 32-mers against k = 14; the 3 GB MultiPieces config 2.03 -> 1.62 ms per 10 M).
    // At most 40 GiB, never more than a third of the free HBM (build_big_table), and only for small alphabets:
    // with 255 symbols one more character multiplies the table by 255 and bought 2.5 % on the byte config.
    uint64_t budget = 2 * idx->hdr.total_bytes;
    if (budget > (8ull << 30)) budget = 8ull << 30;
    if (idx->hdr.sec[SEC_LEVEL0].bytes >= (192ull << 20) && idx->hdr.max_character <= 16) {
        const uint64_t base = idx->hdr.max_character, want = 4 * idx->hdr.n;
        uint64_t entries_k = 1;
        while (entries_k < want && entries_k <= (1ull << 32) / base) entries_k *= base;
        uint64_t b2 = entries_k * 9;
        if (b2 > (40ull << 30)) b2 = 40ull << 30;
        if (b2 > budget) budget = b2;
    }
    if (const char *v = std::getenv("FMX_KMER_BUDGET_MB")) budget = std::strtoull(v, nullptr, 10) << 20;
    return build_big_table(idx, budget);
}

// where the patterns of a batch live on the device: bytes (+ offsets or a fixed length) or packed words
struct PatSrc {
    const uint8_t *pat = nullptr;
    const uint64_t *pat_off = nullptr;
    uint64_t fixed_len = 0;
    const uint64_t *packed = nullptr;
    uint32_t packed_bits = 0;
};

static int make_search_args(const fmx_index *idx, int mode, const PatSrc &ps, uint64_t npat, const uint64_t *d_is,
                            const uint64_t *d_ie, bool count_work, SearchArgs &a) {
    if (mode < FMX_SEARCH || mode > FMX_SEARCH_EXACT) return fail(FMX_ERR_INVALID_ARG, "bad search mode");
    if (mode != FMX_SEARCH && idx->hdr.kind != FMX_KIND_MULTI)
        return fail(FMX_ERR_UNSUPPORTED, "search_prefix/suffix/exact need a MultiPieces index (frontend.rs:369-390)");
    if ((d_is == nullptr) != (d_ie == nullptr)) return fail(FMX_ERR_INVALID_ARG, "init_s and init_e must both be given");
    std::memset(&a, 0, sizeof(a));
    a.pat = ps.pat;
    a.pat_off = ps.pat_off;
    a.fixed_len = ps.fixed_len;
    a.cw_shift = idx->dev.cw_shift;
    if (ps.packed_bits && a.cw_shift) return fail(FMX_ERR_UNSUPPORTED, "packed patterns are for u8 indexes");
    if (ps.packed_bits) {
        if (ps.packed_bits != 2 && ps.packed_bits != 4) return fail(FMX_ERR_INVALID_ARG, "packed_bits must be 0, 2 or 4");
        if (idx->hdr.max_character > (1u << ps.packed_bits))
            return fail(FMX_ERR_UNSUPPORTED, "packed patterns need max_character <= 2^packed_bits");
        if (ps.pat_off || ps.fixed_len == 0) return fail(FMX_ERR_INVALID_ARG, "packed patterns need a fixed length");
        if (ps.fixed_len > 0xFFFFFFFFull / ps.packed_bits) return fail(FMX_ERR_INVALID_ARG, "packed pattern too long");
        a.packed = ps.packed;
        a.packed_bits = ps.packed_bits;
        a.packed_wpp = (uint32_t)((ps.fixed_len * ps.packed_bits + 63) / 64);
    }
    a.npat = npat;
    a.init_s = d_is;
    a.init_e = d_ie;
    a.s0 = 0;  // wrapper.rs:41, 65, 73, 81
    a.e0 = (mode == FMX_SEARCH_SUFFIX || mode == FMX_SEARCH_EXACT) ? (uint32_t)idx->hdr.ndoc : (uint32_t)idx->hdr.n;
    a.err = idx->d_err;
    a.work = (count_work && idx->opt_count_work) ? idx->d_work : nullptr;
    a.staged = (!ps.packed_bits && !a.cw_shift && !ps.pat_off && ps.fixed_len > 0 && ps.fixed_len % 16 == 0 &&
                reinterpret_cast<uintptr_t>(ps.pat) % 16 == 0 && idx->opt_stage_patterns) ? 1u : 0u;
    a.verify = (idx->dev.verify && idx->opt_verify) ? 1u : 0u;
    // the table memoises searches that start from (0, n): fresh search / search_prefix
    const bool tab_ok = idx->d_kmer_tab && idx->opt_kmer && !d_is && (mode == FMX_SEARCH || mode == FMX_SEARCH_PREFIX);
    a.kmer_tab = tab_ok ? idx->d_kmer_tab : nullptr;
    a.kmer_steps = tab_ok ? idx->d_kmer_steps : nullptr;
    a.kmer_k = tab_ok ? idx->kmer_k : 0;
    const bool big_ok = tab_ok && (idx->d_big_tab || idx->d_big_tab4) && idx->opt_kmer_big;
    a.big_tab = big_ok ? idx->d_big_tab : nullptr;
    a.big_tab4 = big_ok ? idx->d_big_tab4 : nullptr;
    a.big_steps = big_ok ? idx->d_big_steps : nullptr;
    a.big_k = big_ok ? idx->big_k : 0;
    a.tab_embed = tab_ok ? idx->tab_embed : 0u;
    return 0;
}

// the phased pipeline (phased.cuh) needs the DENSE verify structures and a fresh search
static bool phased_ok(const fmx_index *idx, const SearchArgs &a) {
    return idx->opt_phased && a.verify && idx->hdr.isa_level == 0 && has_dense_sa(idx) && idx->hdr.kind != FMX_KIND_RLFM &&
           !a.init_s && !a.steps_out && a.npat < 0xFFFFFFFFull;
}

// Backward search of a batch by the phased kernels; leaves (rs, re, hint) in buf[B_RS / B_RE / B_HINT].
static int search_phased(const fmx_index *idx, DevBuf *buf, const SearchArgs &a, bool want_rows, cudaStream_t st) {
    int rc;
    const uint64_t npat = a.npat;
    if ((rc = buf[B_RS].ensure(npat * 4 + 16))) return rc;
    if ((rc = buf[B_RE].ensure(npat * 4 + 16))) return rc;
    if ((rc = buf[B_HINT].ensure(npat * 4 + 16))) return rc;
    if ((rc = buf[B_QN].ensure(64))) return rc;
    if (npat == 0) return 0;
    if (idx->opt_phased == 2) {  // everything in one kernel: no queues
        PhasedArgs g;
        std::memset(&g, 0, sizeof(g));
        g.a = a;
        // Option "order_by_length": ragged byte patterns visited in order of length, so that a warp's lanes run loops of
        // equal trip counts.  Measured on config 5 (100 M patterns of 8..64 bytes): 23.7 ms against 16.6 ms in input
        // order -- the scattered pattern reads and result writes cost more than the convergence buys
        // (profiles/r02_c10b_bench_cfg5_order{0,1}.json).  Off by default.
        const bool by_len = a.pat_off && !a.packed_bits && !a.order && npat < 0xFFFFFFFFull && idx->opt_order_by_length == 1;
        if (by_len) {
            if ((rc = buf[B_BHIST].ensure(FMX_LEN_BUCKETS * 4 + 16))) return rc;
            if ((rc = buf[B_ORDER].ensure(npat * 4))) return rc;
            uint32_t *d_hist = buf[B_BHIST].as<uint32_t>();
            CUDA_TRY(cudaMemsetAsync(d_hist, 0, FMX_LEN_BUCKETS * 4, st));
            uint64_t hb = (npat + 255) / 256;
            if (hb > (uint64_t)idx->sms * 16) hb = (uint64_t)idx->sms * 16;
            k_len_hist<<<(unsigned)hb, 256, 0, st>>>(a.pat_off, npat, d_hist);
            k_len_cursor<<<1, 32, 0, st>>>(d_hist);
            k_len_scatter<<<grid_for(npat, 256u * FMX_LEN_ITEMS), 256, 0, st>>>(a.pat_off, npat, d_hist, buf[B_ORDER].as<uint32_t>());
            g_launches.fetch_add(2, std::memory_order_relaxed);
            LAUNCH_CHECK();
            g.a.order = buf[B_ORDER].as<uint32_t>();
        }
        g.rs = buf[B_RS].as<uint32_t>();
        g.re = buf[B_RE].as<uint32_t>();
        g.hint = buf[B_HINT].as<uint32_t>();
        g.want_rows = want_rows ? 1u : 0u;
        // 1 (default): with the 16-byte table, where the head finishes nine patterns in ten; 2: always (A/B)
        // 3 .. 7: as 1 with (resident blocks per SM, rounds between two looks at the queue) = (6,1) (5,1) (5,2) (6,2) (5,4):
        // the A/B of FMX_DEFER_BLOCKS / FMX_DEFER_ROUNDS; the 16-byte table exists for the Q4 layout only
        const int od = idx->opt_fused_defer;
        const bool defer = ((g.a.big_tab4 != nullptr && od != 0) || od == 2) && !g.a.order && npat < 0xFFFFFFFFull;
        uint64_t blocks = (npat + 255) / 256;
        const uint64_t cap = idx->opt_query_blocks > 0 ? (uint64_t)idx->opt_query_blocks
                                                       : (uint64_t)idx->sms * (defer ? FMX_QUERY_BLOCKS_PER_SM : 64);
        if (blocks > cap) blocks = cap;
        dispatch(idx, [&](auto K, auto LY) {
            if constexpr (K() != FMX_KIND_RLFM_) {
                const unsigned nb = (unsigned)blocks;
                if (!defer) {
                    k_query_fused<K(), LY()><<<nb, 256, 0, st>>>(idx->dev, g);
                    return;
                }
                if constexpr (LY() == FMX_LAYOUT_Q4) {
                    if (od == 3) return (void)k_query_fused_defer<K(), LY(), 6, 1><<<nb, 256, 0, st>>>(idx->dev, g);
                    if (od == 4) return (void)k_query_fused_defer<K(), LY(), 5, 1><<<nb, 256, 0, st>>>(idx->dev, g);
                    if (od == 5) return (void)k_query_fused_defer<K(), LY(), 5, 2><<<nb, 256, 0, st>>>(idx->dev, g);
                    if (od == 6) return (void)k_query_fused_defer<K(), LY(), 6, 2><<<nb, 256, 0, st>>>(idx->dev, g);
                    if (od == 7) return (void)k_query_fused_defer<K(), LY(), 5, 4><<<nb, 256, 0, st>>>(idx->dev, g);
                }
                k_query_fused_defer<K(), LY(), FMX_DEFER_BLOCKS, FMX_DEFER_ROUNDS><<<nb, 256, 0, st>>>(idx->dev, g);
            }
        });
        LAUNCH_CHECK();
        phase_mark(idx, 1, st);
        return 0;
    }
    if ((rc = buf[B_QS].ensure(npat * 16 + 16))) return rc;
    if ((rc = buf[B_QV].ensure(npat * 16 + 16))) return rc;
    PhasedArgs g;
    g.a = a;
    g.rs = buf[B_RS].as<uint32_t>();
    g.re = buf[B_RE].as<uint32_t>();
    g.hint = buf[B_HINT].as<uint32_t>();
    g.want_rows = want_rows ? 1u : 0u;
    g.q_steps = buf[B_QS].as<uint4>();
    g.q_verify = buf[B_QV].as<uint4>();
    g.qn = buf[B_QN].as<unsigned long long>();
    CUDA_TRY(cudaMemsetAsync(g.qn, 0, 32, st));
    uint64_t blocks = (npat + 255) / 256;
    const uint64_t cap = (uint64_t)idx->sms * 8 * 4;
    if (blocks > cap) blocks = cap;
    // queue kernels: one resident wave (a multiple of the SM count); the queue lengths live on the device
    uint64_t qblocks = (uint64_t)idx->sms * 8;
    if (qblocks > (npat + 255) / 256) qblocks = (npat + 255) / 256;
    dispatch(idx, [&](auto K, auto LY) {
        if constexpr (K() != FMX_KIND_RLFM_) {
            k_ph_seed<K(), LY()><<<(unsigned)blocks, 256, 0, st>>>(idx->dev, g);
            phase_mark(idx, 1, st);
            k_ph_steps<K(), LY()><<<(unsigned)qblocks, 256, 0, st>>>(idx->dev, g, 0);
            phase_mark(idx, 2, st);
            k_ph_verify<K(), LY()><<<(unsigned)qblocks, 256, 0, st>>>(idx->dev, g);
            phase_mark(idx, 3, st);
            k_ph_steps<K(), LY()><<<(unsigned)qblocks, 256, 0, st>>>(idx->dev, g, 1);
        }
    });
    g_launches.fetch_add(3, std::memory_order_relaxed);
    LAUNCH_CHECK();
    return 0;
}

static int search_device(const fmx_index *idx, int mode, const PatSrc &ps, uint64_t npat, const uint64_t *d_is,
                         const uint64_t *d_ie, uint64_t *d_os, uint64_t *d_oe, cudaStream_t st, bool count_work = true,
                         bool force_simple = false, DevBuf *buf = nullptr) {
    SearchArgs a;
    int rc = make_search_args(idx, mode, ps, npat, d_is, d_ie, count_work, a);
    if (rc) return rc;
    a.out_s = d_os;
    a.out_e = d_oe;
    if (a.work) {
        CUDA_TRY(cudaMemsetAsync(idx->d_work, 0, sizeof(unsigned long long), st));
        CUDA_TRY(cudaMemsetAsync(idx->d_work + 2, 0, sizeof(unsigned long long), st));
    }
    if (!buf) buf = idx->buf;
    if (phased_ok(idx, a)) {
        if ((rc = search_phased(idx, buf, a, true, st))) return rc;
        if (npat) {
            k_ph_widen<<<grid_for(npat, 256), 256, 0, st>>>(buf[B_RS].as<uint32_t>(), buf[B_RE].as<uint32_t>(), npat, d_os, d_oe);
            LAUNCH_CHECK();
        }
        return 0;
    }
    return dispatch_search(idx, a, st, force_simple || a.packed_bits != 0 || a.cw_shift != 0, buf);
}

static int check_err_flag(const fmx_index *idx, cudaStream_t st) {
    uint32_t flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, idx->d_err, sizeof(flag), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (flag) {
        CUDA_TRY(cudaMemsetAsync(idx->d_err, 0, sizeof(uint32_t), st));
        return fail(FMX_ERR_PATTERN_CHAR,
                    "pattern contains a character larger than max_character (the reference panics: fm_index.rs:94)");
    }
    return FMX_OK;
}

extern "C" int fmx_search_batch_device(const fmx_index *idx, int mode, const uint8_t *d_pat, const uint64_t *d_pat_off,
                                       uint64_t fixed_len, uint64_t npat, const uint64_t *d_init_s,
                                       const uint64_t *d_init_e, uint64_t *d_out_s, uint64_t *d_out_e, void *stream) {
    if (!idx || !d_out_s || !d_out_e) return fail(FMX_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    PatSrc ps;
    ps.pat = d_pat;
    ps.pat_off = d_pat_off;
    ps.fixed_len = fixed_len;
    return search_device(idx, mode, ps, npat, d_init_s, d_init_e, d_out_s, d_out_e, pick_stream(idx, stream));
}

extern "C" int fmx_search_check(const fmx_index *idx, void *stream) {
    if (!idx) return fail(FMX_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(idx->device));
    return check_err_flag(idx, pick_stream(idx, stream));
}

extern "C" int fmx_search_batch(const fmx_index *idx, int mode, const uint8_t *pat, const uint64_t *pat_off,
                                uint64_t fixed_len, uint64_t npat, const uint64_t *init_s, const uint64_t *init_e,
                                uint64_t *out_s, uint64_t *out_e) {
    if (!idx || !out_s || !out_e) return fail(FMX_ERR_INVALID_ARG, "null argument");
    if (npat == 0) return FMX_OK;
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    uint64_t pat_bytes = (pat_off ? pat_off[npat] : npat * fixed_len) * char_width_of(idx);
    if (pat_bytes && !pat) return fail(FMX_ERR_INVALID_ARG, "null pattern buffer");
    int rc;
    if ((rc = idx->buf[B_PAT].ensure(pat_bytes + 16))) return rc;
    if ((rc = idx->buf[B_S].ensure(npat * 8))) return rc;
    if ((rc = idx->buf[B_E].ensure(npat * 8))) return rc;
    if (pat_bytes) CUDA_TRY(cudaMemcpyAsync(idx->buf[B_PAT].p, pat, pat_bytes, cudaMemcpyHostToDevice, st));
    const uint64_t *d_off = nullptr;
    if (pat_off) {
        if ((rc = idx->buf[B_OFF].ensure((npat + 1) * 8))) return rc;
        CUDA_TRY(cudaMemcpyAsync(idx->buf[B_OFF].p, pat_off, (npat + 1) * 8, cudaMemcpyHostToDevice, st));
        d_off = idx->buf[B_OFF].as<uint64_t>();
    }
    const uint64_t *d_is = nullptr, *d_ie = nullptr;
    if (init_s && init_e) {
        if ((rc = idx->buf[B_INS].ensure(npat * 8))) return rc;
        if ((rc = idx->buf[B_INE].ensure(npat * 8))) return rc;
        CUDA_TRY(cudaMemcpyAsync(idx->buf[B_INS].p, init_s, npat * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(idx->buf[B_INE].p, init_e, npat * 8, cudaMemcpyHostToDevice, st));
        d_is = idx->buf[B_INS].as<uint64_t>();
        d_ie = idx->buf[B_INE].as<uint64_t>();
    } else if (init_s || init_e) {
        return fail(FMX_ERR_INVALID_ARG, "init_s and init_e must both be given");
    }
    PatSrc ps;
    ps.pat = idx->buf[B_PAT].as<uint8_t>();
    ps.pat_off = d_off;
    ps.fixed_len = fixed_len;
    rc = search_device(idx, mode, ps, npat, d_is, d_ie, idx->buf[B_S].as<uint64_t>(), idx->buf[B_E].as<uint64_t>(), st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_s, idx->buf[B_S].p, npat * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_e, idx->buf[B_E].p, npat * 8, cudaMemcpyDeviceToHost, st));
    return check_err_flag(idx, st);
}

// ------------------------------------------------------------------ locate

// where the rows of the hits come from: the expanded list, or (rows == NULL) the ranges themselves
struct RowSource {
    const uint32_t *rows = nullptr;
    const uint64_t *hoff = nullptr;
    const uint64_t *s = nullptr;
    uint64_t npat = 0;
    uint64_t first = 0;
    bool ranges = false;  // walk sub-ranges (k_locate_ranges) instead of rows
};

// small batches skip the row expansion (memset + mark + 2-3 scan launches + expand): k_locate finds its row
// by binary search over the hit offsets instead
static bool expand_by_search(const fmx_index *idx, uint64_t total) {
    if (idx->opt_locate_expand == 1) return false;
    if (idx->opt_locate_expand == 2) return true;
    return total <= (1ull << 22);
}

// many matches per pattern (repetitive texts): walk whole sub-ranges instead of single rows (k_locate_ranges)
static bool locate_by_ranges(const fmx_index *idx, uint64_t total, uint64_t npat) {
    if (idx->hdr.sa_level > 6 || idx->opt_locate_ranges == 0 || idx->opt_locate_refill) return false;
    if (idx->opt_locate_ranges == 1) return true;
    return idx->hdr.kind == FMX_KIND_RLFM && npat > 0 && total >= 8 * npat;
}


// ---- locate, stage 1 (asynchronous): candidate-row counts of every range and their exclusive
// prefix sum.  d_off gets npat+1 entries; d_off[npat] is the number of candidate rows.
static int locate_counts(const fmx_index *idx, DevBuf *buf, const uint64_t *d_s, const uint64_t *d_e, uint64_t npat,
                         uint64_t *d_off, cudaStream_t st) {
    int rc;
    if (npat >= 0xFFFFFFFFull) return fail(FMX_ERR_UNSUPPORTED, "at most 2^32 - 2 patterns per locate call");
    (void)rc;
    // candidate rows per range (e - s) are computed inside the scan's loads
    return device_scan_load<LoadRangeCount, uint64_t, OpSum, true>(LoadRangeCount{d_s, d_e}, npat, d_off, OpSum(), true,
                                                                   buf[B_TILES], st);
}

// ---- locate, stage 2: the candidate rows themselves, in pattern order then ascending row order
// (wrapper.rs:206-216), given their count `total` (= d_off[npat], already read back by the caller).
// With prefix_only the L == 0 filter (wrapper.rs:208) is applied by flag / scan / compaction and
// d_hit_off receives the filtered offsets (one stream synchronisation to learn the kept count).
// Without it d_hit_off must be d_off.  *rows_out points at the hit rows in scratch.
static int locate_rows(const fmx_index *idx, DevBuf *buf, int prefix_only, const uint64_t *d_s, uint64_t npat,
                       const uint64_t *d_off, uint64_t total, uint64_t *d_hit_off, uint64_t *hits_out,
                       RowSource *rows_out, cudaStream_t st) {
    int rc;
    *rows_out = RowSource();
    if (total >= 0xFFFFFFFFull * 4) return fail(FMX_ERR_UNSUPPORTED, "too many candidate rows in one locate call");
    if (total == 0) {
        if (prefix_only) CUDA_TRY(cudaMemsetAsync(d_hit_off, 0, (npat + 1) * 8, st));
        *hits_out = 0;
        return 0;
    }
    if (!prefix_only && (locate_by_ranges(idx, total, npat) || expand_by_search(idx, total))) {
        *hits_out = total;
        rows_out->hoff = d_off;
        rows_out->s = d_s;
        rows_out->npat = npat;
        rows_out->ranges = locate_by_ranges(idx, total, npat);
        return 0;
    }
    if ((rc = buf[B_OWNER].ensure(total * 4))) return rc;
    if ((rc = buf[B_ROWS].ensure(total * 4))) return rc;
    uint32_t *d_owner = buf[B_OWNER].as<uint32_t>();
    uint32_t *d_rows = buf[B_ROWS].as<uint32_t>();
    CUDA_TRY(cudaMemsetAsync(d_owner, 0, total * 4, st));
    k_mark_owners<<<grid_for(npat, 256), 256, 0, st>>>(d_off, npat, d_owner);
    LAUNCH_CHECK();
    if ((rc = device_scan<uint32_t, uint32_t, OpMax, false>(d_owner, total, d_owner, OpMax(), false, buf[B_TILES], st))) return rc;
    k_expand_rows<<<grid_for(total, 256), 256, 0, st>>>(d_s, d_off, d_owner, total, nullptr, d_rows);
    LAUNCH_CHECK();
    if (!prefix_only) {
        *hits_out = total;
        rows_out->rows = d_rows;
        return 0;
    }
    if ((rc = buf[B_FLAG].ensure(total * 4))) return rc;
    if ((rc = buf[B_FPOS].ensure((total + 1) * 8))) return rc;
    uint32_t *d_flag = buf[B_FLAG].as<uint32_t>();
    uint64_t *d_fpos = buf[B_FPOS].as<uint64_t>();
    dispatch(idx, [&](auto K, auto LY) { k_flag_prefix<K(), LY()><<<grid_for(total, 256), 256, 0, st>>>(idx->dev, d_rows, total, d_flag); });
    LAUNCH_CHECK();
    if ((rc = device_scan<uint32_t, uint64_t, OpSum, true>(d_flag, total, d_fpos, OpSum(), true, buf[B_TILES], st))) return rc;
    uint64_t kept = 0;
    CUDA_TRY(cudaMemcpyAsync(&kept, d_fpos + total, 8, cudaMemcpyDeviceToHost, st));
    k_filtered_offsets<<<grid_for(npat + 1, 256), 256, 0, st>>>(d_off, d_fpos, npat, d_hit_off);
    LAUNCH_CHECK();
    CUDA_TRY(cudaStreamSynchronize(st));
    if (kept) {
        if ((rc = buf[B_ROWS2].ensure(kept * 4))) return rc;
        k_compact_rows<<<grid_for(total, 256), 256, 0, st>>>(d_rows, d_flag, d_fpos, total, buf[B_ROWS2].as<uint32_t>());
        LAUNCH_CHECK();
        rows_out->rows = buf[B_ROWS2].as<uint32_t>();
    }
    *hits_out = kept;
    return 0;
}

// both stages with a synchronisation in between (the two-phase device API and fmx_locate_batch)
static int locate_prepare(const fmx_index *idx, int prefix_only, const uint64_t *d_s, const uint64_t *d_e,
                          uint64_t npat, uint64_t *d_hit_off, uint64_t *total_out, RowSource *rows_out,
                          cudaStream_t st) {
    int rc;
    DevBuf *buf = idx->buf;
    uint64_t *d_off = d_hit_off;  // unfiltered offsets go to d_hit_off directly when there is no filter
    if (prefix_only) {
        if ((rc = buf[B_HOFF].ensure((npat + 1) * 8))) return rc;
        d_off = buf[B_HOFF].as<uint64_t>();
    }
    if ((rc = locate_counts(idx, buf, d_s, d_e, npat, d_off, st))) return rc;
    uint64_t total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, d_off + npat, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return locate_rows(idx, buf, prefix_only, d_s, npat, d_off, total, d_hit_off, total_out, rows_out, st);
}

static int locate_fill(const fmx_index *idx, const RowSource &src, uint64_t total, uint64_t *d_pos, uint64_t *d_pid,
                       cudaStream_t st, bool count_work = true, const uint64_t *total_dev = nullptr) {
    count_work = count_work && idx->opt_count_work;
    if (count_work) CUDA_TRY(cudaMemsetAsync(idx->d_work + 1, 0, sizeof(unsigned long long), st));
    if (total == 0) return 0;
    LocateArgs a;
    a.dense = locate_dense(idx) ? 1u : 0u;
    a.rows = src.rows;
    a.hoff = src.hoff;
    a.s = src.s;
    a.npat = src.npat;
    a.first = src.first;
    a.total = total;
    a.total_dev = total_dev;
    a.positions = d_pos;
    a.piece_ids = d_pid;
    a.work = count_work ? idx->d_work : nullptr;
    if (src.ranges && !src.rows && !a.dense) {
        const uint64_t threads = (total + LOCATE_GROUP - 1) / LOCATE_GROUP;
        dispatch(idx, [&](auto K, auto LY) {
            k_locate_ranges<K(), LY()><<<(unsigned)((threads + LOCATE_RANGE_THREADS - 1) / LOCATE_RANGE_THREADS),
                                         LOCATE_RANGE_THREADS, 0, st>>>(idx->dev, a);
        });
        LAUNCH_CHECK();
        return 0;
    }
    if (!idx->opt_locate_refill || a.dense) {
        uint64_t blocks = (total + 255) / 256;
        uint64_t cap = (uint64_t)idx->sms * 8 * 8;
        if (blocks > cap) blocks = cap;
        dispatch(idx, [&](auto K, auto LY) { k_locate_simple<K(), LY()><<<(unsigned)blocks, 256, 0, st>>>(idx->dev, a); });
        LAUNCH_CHECK();
        return 0;
    }
    // per-lane refill: a warp works through chunks of consecutive hits; enough chunks per warp to keep
    // its lanes fed, enough warps to fill the GPU (8 CTAs of 8 warps per SM)
    const uint64_t max_warps = (uint64_t)idx->sms * 8 * 8;
    uint64_t chunk = 64;
    while (chunk < 1024 && (total + chunk - 1) / chunk > 4 * max_warps) chunk *= 2;
    uint64_t warps = (total + chunk - 1) / chunk;
    if (warps > max_warps) warps = max_warps;
    a.chunk = chunk;
    const uint64_t blocks = (warps + 7) / 8;
    if (idx->opt_locate_refill == 2)
        dispatch(idx, [&](auto K, auto LY) { k_locate_stable<K(), LY()><<<(unsigned)blocks, 256, 0, st>>>(idx->dev, a); });
    else
        dispatch(idx, [&](auto K, auto LY) { k_locate<K(), LY()><<<(unsigned)blocks, 256, 0, st>>>(idx->dev, a); });
    LAUNCH_CHECK();
    return 0;
}

static int locate_args_ok(const fmx_index *idx, bool want_pid) {
    if (!idx->hdr.has_locate)
        return fail(FMX_ERR_NO_LOCATE, "index was built without a sampled suffix array (count-only variant)");
    if (want_pid && idx->hdr.kind != FMX_KIND_MULTI)
        return fail(FMX_ERR_UNSUPPORTED, "piece_id needs a MultiPieces index (frontend.rs:542)");
    return 0;
}

extern "C" int fmx_locate_count_device(const fmx_index *idx, int prefix_only, const uint64_t *d_s, const uint64_t *d_e,
                                       uint64_t npat, uint64_t *d_hit_off, uint64_t *total_hits, void *stream) {
    if (!idx || !d_hit_off || !total_hits) return fail(FMX_ERR_INVALID_ARG, "null argument");
    int rc = locate_args_ok(idx, false);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    RowSource rows;
    idx->count_token = CountToken();
    rc = locate_prepare(idx, prefix_only, d_s, d_e, npat, d_hit_off, total_hits, &rows, pick_stream(idx, stream));
    if (rc) return rc;
    // what the matching fmx_locate_fill_device call must present again (the rows may sit in scratch)
    idx->count_token.valid = true;
    idx->count_token.prefix_only = prefix_only;
    idx->count_token.npat = npat;
    idx->count_token.total = *total_hits;
    idx->count_token.d_s = d_s;
    idx->count_token.d_hit_off = d_hit_off;
    idx->count_token.rows = rows.rows;
    idx->count_token.ranges = rows.ranges;
    return FMX_OK;
}

extern "C" int fmx_locate_fill_device(const fmx_index *idx, int prefix_only, const uint64_t *d_s, const uint64_t *d_e,
                                      uint64_t npat, const uint64_t *d_hit_off, uint64_t total_hits,
                                      uint64_t *d_positions, uint64_t *d_piece_ids, void *stream) {
    if (!idx) return fail(FMX_ERR_INVALID_ARG, "null argument");
    int rc = locate_args_ok(idx, d_piece_ids != nullptr);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    (void)d_e;
    // the rows prepared by the matching fmx_locate_count_device call may still sit in scratch: the call must be
    // the continuation of exactly that count
    const CountToken &tk = idx->count_token;
    if (!tk.valid || tk.prefix_only != prefix_only || tk.npat != npat || tk.total != total_hits || tk.d_s != d_s ||
        tk.d_hit_off != d_hit_off)
        return fail(FMX_ERR_INVALID_ARG, "fmx_locate_fill_device must follow the fmx_locate_count_device call for the same ranges");
    RowSource rows;
    if (tk.rows) {
        rows.rows = tk.rows;
    } else {
        rows.hoff = d_hit_off;
        rows.s = d_s;
        rows.npat = npat;
        rows.ranges = tk.ranges;
    }
    return locate_fill(idx, rows, total_hits, d_positions, d_piece_ids, pick_stream(idx, stream));
}

// Fully asynchronous locate on device buffers (no host synchronisation, graph-capturable): the hit
// total stays on the device (d_hit_off[npat]); launches are sized by `capacity`.
static int locate_async(const fmx_index *idx, DevBuf *buf, int prefix_only, const uint64_t *d_s, const uint64_t *d_e,
                        uint64_t npat, uint64_t *d_hit_off, uint64_t *d_positions, uint64_t *d_piece_ids, uint64_t capacity,
                        cudaStream_t st, bool count_work) {
    int rc;
    if (prefix_only) {  // the L == 0 filter needs the kept count on the host: synchronous path
        uint64_t total = 0;
        RowSource rows;
        uint64_t *d_off = d_hit_off;
        if ((rc = buf[B_HOFF].ensure((npat + 1) * 8))) return rc;
        d_off = buf[B_HOFF].as<uint64_t>();
        if ((rc = locate_counts(idx, buf, d_s, d_e, npat, d_off, st))) return rc;
        uint64_t cand = 0;
        CUDA_TRY(cudaMemcpyAsync(&cand, d_off + npat, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if ((rc = locate_rows(idx, buf, prefix_only, d_s, npat, d_off, cand, d_hit_off, &total, &rows, st))) return rc;
        if (total > capacity) total = capacity;
        if (!d_positions && !d_piece_ids) return 0;
        return locate_fill(idx, rows, total, d_positions, d_piece_ids, st, count_work);
    }
    if ((rc = locate_counts(idx, buf, d_s, d_e, npat, d_hit_off, st))) return rc;
    if (capacity == 0 || npat == 0 || (!d_positions && !d_piece_ids)) return FMX_OK;
    if (capacity >= 0xFFFFFFFFull * 4) return fail(FMX_ERR_UNSUPPORTED, "capacity too large");
    const uint64_t *d_total0 = d_hit_off + npat;
    if (locate_by_ranges(idx, capacity, npat) || expand_by_search(idx, capacity)) {
        RowSource src;
        src.hoff = d_hit_off;
        src.s = d_s;
        src.npat = npat;
        src.ranges = locate_by_ranges(idx, capacity, npat);
        return locate_fill(idx, src, capacity, d_positions, d_piece_ids, st, count_work, d_total0);
    }
    if ((rc = buf[B_OWNER].ensure(capacity * 4))) return rc;
    if ((rc = buf[B_ROWS].ensure(capacity * 4))) return rc;
    uint32_t *d_owner = buf[B_OWNER].as<uint32_t>();
    uint32_t *d_rows = buf[B_ROWS].as<uint32_t>();
    const uint64_t *d_total = d_hit_off + npat;
    CUDA_TRY(cudaMemsetAsync(d_owner, 0, capacity * 4, st));
    k_mark_owners_capped<<<grid_for(npat, 256), 256, 0, st>>>(d_hit_off, npat, capacity, d_owner);
    LAUNCH_CHECK();
    if ((rc = device_scan<uint32_t, uint32_t, OpMax, false>(d_owner, capacity, d_owner, OpMax(), false, buf[B_TILES], st))) return rc;
    k_expand_rows<<<grid_for(capacity, 256), 256, 0, st>>>(d_s, d_hit_off, d_owner, capacity, d_total, d_rows);
    LAUNCH_CHECK();
    RowSource src;
    src.rows = d_rows;
    return locate_fill(idx, src, capacity, d_positions, d_piece_ids, st, count_work, d_total);
}

extern "C" int fmx_locate_batch_device(const fmx_index *idx, int prefix_only, const uint64_t *d_s, const uint64_t *d_e,
                                       uint64_t npat, uint64_t *d_hit_off, uint64_t *d_positions, uint64_t *d_piece_ids,
                                       uint64_t capacity, void *stream) {
    if (!idx || !d_hit_off || (!d_positions && !d_piece_ids)) return fail(FMX_ERR_INVALID_ARG, "null argument");
    int rc = locate_args_ok(idx, d_piece_ids != nullptr);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    return locate_async(idx, idx->buf, prefix_only, d_s, d_e, npat, d_hit_off, d_positions, d_piece_ids, capacity,
                        pick_stream(idx, stream), true);
}

// One page of the hit list: hits [first_hit, first_hit + nhits) of the CSR that fmx_locate_count_device
// (prefix_only == 0) left in d_hit_off, written to d_positions[0 .. nhits).  Bounded-memory iteration over
// huge match sets (lazy iter_matches, wrapper.rs:137-139, 203-217): asynchronous, no scratch.
extern "C" int fmx_locate_page_device(const fmx_index *idx, const uint64_t *d_s, const uint64_t *d_hit_off, uint64_t npat,
                                      uint64_t first_hit, uint64_t nhits, uint64_t *d_positions, uint64_t *d_piece_ids,
                                      void *stream) {
    if (!idx || !d_s || !d_hit_off || (!d_positions && !d_piece_ids)) return fail(FMX_ERR_INVALID_ARG, "null argument");
    int rc = locate_args_ok(idx, d_piece_ids != nullptr);
    if (rc) return rc;
    if (nhits == 0 || npat == 0) return FMX_OK;
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    RowSource src;
    src.hoff = d_hit_off;
    src.s = d_s;
    src.npat = npat;
    src.first = first_hit;
    src.ranges = locate_by_ranges(idx, nhits, npat);
    return locate_fill(idx, src, nhits, d_positions, d_piece_ids, pick_stream(idx, stream));
}

// host-buffer form: SA ranges in, one page of positions out; *total_hits = size of the whole hit list
extern "C" int fmx_locate_page(const fmx_index *idx, const uint64_t *s, const uint64_t *e, uint64_t npat, uint64_t first_hit,
                               uint64_t nhits, uint64_t *positions, uint64_t *piece_ids, uint64_t *total_hits) {
    if (!idx || !total_hits || (!s && npat) || (!e && npat)) return fail(FMX_ERR_INVALID_ARG, "null argument");
    int rc = locate_args_ok(idx, piece_ids != nullptr);
    if (rc) return rc;
    *total_hits = 0;
    if (npat == 0) return FMX_OK;
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    if ((rc = idx->buf[B_S].ensure(npat * 8 + 8))) return rc;
    if ((rc = idx->buf[B_E].ensure(npat * 8 + 8))) return rc;
    if ((rc = idx->buf[B_OFF].ensure((npat + 1) * 8))) return rc;
    CUDA_TRY(cudaMemcpyAsync(idx->buf[B_S].p, s, npat * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(idx->buf[B_E].p, e, npat * 8, cudaMemcpyHostToDevice, st));
    uint64_t *d_hoff = idx->buf[B_OFF].as<uint64_t>();
    if ((rc = locate_counts(idx, idx->buf, idx->buf[B_S].as<uint64_t>(), idx->buf[B_E].as<uint64_t>(), npat, d_hoff, st))) return rc;
    uint64_t total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, d_hoff + npat, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *total_hits = total;
    if (first_hit >= total || nhits == 0) return FMX_OK;
    if (nhits > total - first_hit) nhits = total - first_hit;
    uint64_t *d_pos = nullptr, *d_pid = nullptr;
    if (positions) {
        if ((rc = idx->buf[B_POS].ensure(nhits * 8))) return rc;
        d_pos = idx->buf[B_POS].as<uint64_t>();
    }
    if (piece_ids) {
        if ((rc = idx->buf[B_PID].ensure(nhits * 8))) return rc;
        d_pid = idx->buf[B_PID].as<uint64_t>();
    }
    if (!d_pos && !d_pid) return FMX_OK;
    RowSource src;
    src.hoff = d_hoff;
    src.s = idx->buf[B_S].as<uint64_t>();
    src.npat = npat;
    src.first = first_hit;
    src.ranges = locate_by_ranges(idx, total, npat);
    if ((rc = locate_fill(idx, src, nhits, d_pos, d_pid, st))) return rc;
    if (positions) CUDA_TRY(cudaMemcpyAsync(positions, d_pos, nhits * 8, cudaMemcpyDeviceToHost, st));
    if (piece_ids) CUDA_TRY(cudaMemcpyAsync(piece_ids, d_pid, nhits * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return FMX_OK;
}

extern "C" int fmx_locate_batch(const fmx_index *idx, int prefix_only, const uint64_t *s, const uint64_t *e, uint64_t npat,
                                uint64_t *hit_off, uint64_t **positions, uint64_t **piece_ids) {
    if (!idx || !hit_off || (!s && npat) || (!e && npat)) return fail(FMX_ERR_INVALID_ARG, "null argument");
    if (positions) *positions = nullptr;
    if (piece_ids) *piece_ids = nullptr;
    int rc = locate_args_ok(idx, piece_ids != nullptr);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    if ((rc = idx->buf[B_S].ensure(npat * 8 + 8))) return rc;
    if ((rc = idx->buf[B_E].ensure(npat * 8 + 8))) return rc;
    if ((rc = idx->buf[B_OFF].ensure((npat + 1) * 8))) return rc;
    if (npat) {
        CUDA_TRY(cudaMemcpyAsync(idx->buf[B_S].p, s, npat * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(idx->buf[B_E].p, e, npat * 8, cudaMemcpyHostToDevice, st));
    }
    uint64_t total = 0;
    RowSource rows;
    uint64_t *d_hoff = idx->buf[B_OFF].as<uint64_t>();
    if ((rc = locate_prepare(idx, prefix_only, idx->buf[B_S].as<uint64_t>(), idx->buf[B_E].as<uint64_t>(), npat, d_hoff,
                             &total, &rows, st)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(hit_off, d_hoff, (npat + 1) * 8, cudaMemcpyDeviceToHost, st));
    uint64_t *h_pos = nullptr, *h_pid = nullptr;
    if (total) {
        uint64_t *d_pos = nullptr, *d_pid = nullptr;
        if (positions) {
            if ((rc = idx->buf[B_POS].ensure(total * 8))) return rc;
            d_pos = idx->buf[B_POS].as<uint64_t>();
        }
        if (piece_ids) {
            if ((rc = idx->buf[B_PID].ensure(total * 8))) return rc;
            d_pid = idx->buf[B_PID].as<uint64_t>();
        }
        if ((rc = locate_fill(idx, rows, total, d_pos, d_pid, st))) return rc;
        if (positions) {
            h_pos = static_cast<uint64_t *>(std::malloc(total * 8));
            if (!h_pos) return fail(FMX_ERR_OOM, "malloc positions");
            CUDA_TRY(cudaMemcpyAsync(h_pos, d_pos, total * 8, cudaMemcpyDeviceToHost, st));
        }
        if (piece_ids) {
            h_pid = static_cast<uint64_t *>(std::malloc(total * 8));
            if (!h_pid) {
                std::free(h_pos);
                return fail(FMX_ERR_OOM, "malloc piece ids");
            }
            CUDA_TRY(cudaMemcpyAsync(h_pid, d_pid, total * 8, cudaMemcpyDeviceToHost, st));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (positions) *positions = h_pos;
    if (piece_ids) *piece_ids = h_pid;
    return FMX_OK;
}

// ------------------------------------------------------------------ fused query: count + locate, any input form

// out[i] = in[i] (+ base) in the caller's width
template <class Tin, class Tout>
__global__ void k_convert(const Tin *in, uint64_t n, uint64_t base, Tout *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (Tout)((uint64_t)in[i] + base);
}
// counts[p] = e - s (wrapper.rs:132-134; an inverted range counts 0 here, the reference would underflow)
template <class Tr, class Tout>
__global__ void k_range_counts(const Tr *s, const Tr *e, uint64_t n, Tout *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = e[i] > s[i] ? (Tout)(e[i] - s[i]) : (Tout)0;
}

struct QueryOut {           // device pointers, caller's width
    uint32_t width = 8;
    uint64_t *out_s = nullptr, *out_e = nullptr;
    void *counts = nullptr, *hit_off = nullptr, *positions = nullptr, *piece_ids = nullptr;
    uint64_t capacity = 0;
};

template <class Tw>
static int emit_dense(const fmx_index *idx, DevBuf *buf, uint64_t npat, const uint32_t *hint, const QueryOut &o, cudaStream_t st,
                      bool count_work) {
    int rc;
    if ((rc = buf[B_BIGQ].ensure(npat * 4 + 16))) return rc;
    if ((rc = buf[B_QN].ensure(64))) return rc;
    EmitArgs<Tw, Tw> g;
    g.rs = buf[B_RS].as<uint32_t>();
    g.re = buf[B_RE].as<uint32_t>();
    g.hint = hint;
    g.off = static_cast<const Tw *>(o.hit_off);
    g.npat = npat;
    g.capacity = o.capacity;
    g.positions = static_cast<Tw *>(o.positions);
    g.piece_ids = static_cast<Tw *>(o.piece_ids);
    g.bigq = buf[B_BIGQ].as<uint32_t>();
    g.bign = buf[B_QN].as<unsigned long long>() + 4;
    g.req = count_work ? idx->d_work + 3 : nullptr;
    CUDA_TRY(cudaMemsetAsync(g.bign, 0, 8, st));
    uint64_t blocks = (npat + 255) / 256;
    const uint64_t cap = (uint64_t)idx->sms * 8 * 4;
    if (blocks > cap) blocks = cap;
    uint64_t bblocks = (uint64_t)idx->sms * 8;
    if (bblocks > (npat + 7) / 8) bblocks = (npat + 7) / 8;
    if (idx->hdr.kind == FMX_KIND_MULTI) {
        k_emit_small<FMX_KIND_MULTI_, Tw, Tw><<<(unsigned)blocks, 256, 0, st>>>(idx->dev, g);
        k_emit_big<FMX_KIND_MULTI_, Tw, Tw><<<(unsigned)bblocks, 256, 0, st>>>(idx->dev, g);
    } else {
        k_emit_small<FMX_KIND_FM_, Tw, Tw><<<(unsigned)blocks, 256, 0, st>>>(idx->dev, g);
        k_emit_big<FMX_KIND_FM_, Tw, Tw><<<(unsigned)bblocks, 256, 0, st>>>(idx->dev, g);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    LAUNCH_CHECK();
    return 0;
}

// counts -> offsets -> positions of a rich index in one pass over the ranges (k_offsets_emit as the last phase of the
// scan) + k_emit_big for the patterns with many matches
template <class Tw>
static int offsets_emit_fused(const fmx_index *idx, DevBuf *buf, uint64_t npat, const uint32_t *hint, const QueryOut &o, cudaStream_t st,
                              bool count_work) {
    int rc;
    if ((rc = buf[B_BIGQ].ensure(npat * 4 + 16))) return rc;
    if ((rc = buf[B_QN].ensure(64))) return rc;
    EmitArgs<Tw, Tw> g;
    g.rs = buf[B_RS].as<uint32_t>();
    g.re = buf[B_RE].as<uint32_t>();
    g.hint = hint;
    g.off = static_cast<const Tw *>(o.hit_off);
    g.npat = npat;
    g.capacity = o.capacity;
    g.positions = static_cast<Tw *>(o.positions);
    g.piece_ids = static_cast<Tw *>(o.piece_ids);
    g.bigq = buf[B_BIGQ].as<uint32_t>();
    g.bign = buf[B_QN].as<unsigned long long>() + 4;
    g.req = count_work ? idx->d_work + 3 : nullptr;
    CUDA_TRY(cudaMemsetAsync(g.bign, 0, 8, st));
    const bool multi = idx->hdr.kind == FMX_KIND_MULTI;
    Tw *off_out = static_cast<Tw *>(o.hit_off);
    const ScanLast last = [&](unsigned ntiles, const uint64_t *prefix, const uint64_t *sums) -> int {
        if (multi) k_offsets_emit<FMX_KIND_MULTI_, Tw><<<ntiles, SCAN_THREADS, 0, st>>>(idx->dev, g, off_out, prefix, sums);
        else k_offsets_emit<FMX_KIND_FM_, Tw><<<ntiles, SCAN_THREADS, 0, st>>>(idx->dev, g, off_out, prefix, sums);
        LAUNCH_CHECK();
        return 0;
    };
    if ((rc = device_scan_load<LoadCount32, Tw, OpSum, true>(LoadCount32{g.rs, g.re}, npat, off_out, OpSum(), true, buf[B_TILES], st, &last)))
        return rc;
    phase_mark(idx, 5, st);
    uint64_t bblocks = (uint64_t)idx->sms * 8;
    if (bblocks > (npat + 7) / 8) bblocks = (npat + 7) / 8;
    if (multi) k_emit_big<FMX_KIND_MULTI_, Tw, Tw><<<(unsigned)bblocks, 256, 0, st>>>(idx->dev, g);
    else k_emit_big<FMX_KIND_FM_, Tw, Tw><<<(unsigned)bblocks, 256, 0, st>>>(idx->dev, g);
    LAUNCH_CHECK();
    return 0;
}

// The whole query on device buffers, asynchronous.  HBM-rich indexes: phased search (no inverse-suffix-array
// request unless the caller wants the rows) + positions from the hints / the resident suffix array.  Every other
// index: k_search + the LF-walk locate kernels.
static int query_device(const fmx_index *idx, DevBuf *buf, int mode, const PatSrc &ps, uint64_t npat, const QueryOut &o,
                        cudaStream_t st, bool count_work) {
    int rc;
    const bool want_hits = o.positions != nullptr || o.piece_ids != nullptr;
    const bool want_off = o.hit_off != nullptr;
    const bool want_rows = o.out_s != nullptr || o.out_e != nullptr;
    const int prefix_only = (mode == FMX_SEARCH_PREFIX || mode == FMX_SEARCH_EXACT) ? 1 : 0;
    if (o.width != 8 && o.width != 4) return fail(FMX_ERR_INVALID_ARG, "out_width must be 8 or 4");
    if (want_hits && !want_off) return fail(FMX_ERR_INVALID_ARG, "positions / piece_ids need hit_off");
    if (want_rows && (!o.out_s || !o.out_e)) return fail(FMX_ERR_INVALID_ARG, "out_s and out_e must both be given");
    if (want_hits && (rc = locate_args_ok(idx, o.piece_ids != nullptr))) return rc;
    if (npat >= 0xFFFFFFFFull) return fail(FMX_ERR_UNSUPPORTED, "at most 2^32 - 2 patterns per call");
    SearchArgs a;
    if ((rc = make_search_args(idx, mode, ps, npat, nullptr, nullptr, count_work, a))) return rc;
    if (a.work) CUDA_TRY(cudaMemsetAsync(idx->d_work, 0, 4 * sizeof(unsigned long long), st));
    if (npat == 0) {
        if (want_off) CUDA_TRY(cudaMemsetAsync(o.hit_off, 0, o.width, st));
        return 0;
    }
    phase_mark(idx, 0, st);
    const bool dense = locate_dense(idx) || (!want_hits && has_dense_sa(idx));
    const bool rich_out = dense && !prefix_only;   // counts / offsets / positions straight from (rs, re, hint)
    const uint32_t *hint = nullptr;
    uint64_t *d_s = o.out_s, *d_e = o.out_e;
    if (phased_ok(idx, a)) {
        const bool rows = want_rows || !rich_out;
        if ((rc = search_phased(idx, buf, a, rows, st))) return rc;
        hint = buf[B_HINT].as<uint32_t>();
        if (rows) {
            if (!d_s) {
                if ((rc = buf[B_S].ensure(npat * 8))) return rc;
                if ((rc = buf[B_E].ensure(npat * 8))) return rc;
                d_s = buf[B_S].as<uint64_t>();
                d_e = buf[B_E].as<uint64_t>();
            }
            k_ph_widen<<<grid_for(npat, 256), 256, 0, st>>>(buf[B_RS].as<uint32_t>(), buf[B_RE].as<uint32_t>(), npat, d_s, d_e);
            LAUNCH_CHECK();
        }
    } else {
        if (!d_s) {
            if ((rc = buf[B_S].ensure(npat * 8))) return rc;
            if ((rc = buf[B_E].ensure(npat * 8))) return rc;
            d_s = buf[B_S].as<uint64_t>();
            d_e = buf[B_E].as<uint64_t>();
        }
        a.out_s = d_s;
        a.out_e = d_e;
        if ((rc = dispatch_search(idx, a, st, true, buf))) return rc;
        phase_mark(idx, 1, st);
        if (rich_out) {  // (s, e) as u32 for the emit kernels
            if ((rc = buf[B_RS].ensure(npat * 4 + 16))) return rc;
            if ((rc = buf[B_RE].ensure(npat * 4 + 16))) return rc;
            k_convert<uint64_t, uint32_t><<<grid_for(npat, 256), 256, 0, st>>>(d_s, npat, 0, buf[B_RS].as<uint32_t>());
            k_convert<uint64_t, uint32_t><<<grid_for(npat, 256), 256, 0, st>>>(d_e, npat, 0, buf[B_RE].as<uint32_t>());
            g_launches.fetch_add(1, std::memory_order_relaxed);
            LAUNCH_CHECK();
        }
    }
    phase_mark(idx, 4, st);
    if (rich_out) {
        const uint32_t *rs = buf[B_RS].as<uint32_t>(), *re = buf[B_RE].as<uint32_t>();
        if (o.counts) {
            if (o.width == 8) k_range_counts<uint32_t, uint64_t><<<grid_for(npat, 256), 256, 0, st>>>(rs, re, npat, static_cast<uint64_t *>(o.counts));
            else k_range_counts<uint32_t, uint32_t><<<grid_for(npat, 256), 256, 0, st>>>(rs, re, npat, static_cast<uint32_t *>(o.counts));
            LAUNCH_CHECK();
        }
        if (!want_off) return 0;
        if (want_hits && o.capacity > 0 && idx->opt_emit_fused) {  // offsets and positions in one pass over the ranges
            if (a.work) CUDA_TRY(cudaMemsetAsync(idx->d_work + 1, 0, sizeof(unsigned long long), st));
            rc = o.width == 8 ? offsets_emit_fused<uint64_t>(idx, buf, npat, hint, o, st, a.work != nullptr)
                              : offsets_emit_fused<uint32_t>(idx, buf, npat, hint, o, st, a.work != nullptr);
            phase_mark(idx, 6, st);
            return rc;
        }
        if (o.width == 8) rc = device_scan_load<LoadCount32, uint64_t, OpSum, true>(LoadCount32{rs, re}, npat, static_cast<uint64_t *>(o.hit_off), OpSum(), true, buf[B_TILES], st);
        else rc = device_scan_load<LoadCount32, uint32_t, OpSum, true>(LoadCount32{rs, re}, npat, static_cast<uint32_t *>(o.hit_off), OpSum(), true, buf[B_TILES], st);
        if (rc) return rc;
        phase_mark(idx, 5, st);
        if (!want_hits || o.capacity == 0) return 0;
        if (a.work) CUDA_TRY(cudaMemsetAsync(idx->d_work + 1, 0, sizeof(unsigned long long), st));
        rc = o.width == 8 ? emit_dense<uint64_t>(idx, buf, npat, hint, o, st, a.work != nullptr)
                          : emit_dense<uint32_t>(idx, buf, npat, hint, o, st, a.work != nullptr);
        phase_mark(idx, 6, st);
        return rc;
    }
    // LF-walk locate (compact indexes, filtered modes): u64 internally, narrowed at the end for out_width 4
    if (o.counts) {
        if (o.width == 8) k_range_counts<uint64_t, uint64_t><<<grid_for(npat, 256), 256, 0, st>>>(d_s, d_e, npat, static_cast<uint64_t *>(o.counts));
        else k_range_counts<uint64_t, uint32_t><<<grid_for(npat, 256), 256, 0, st>>>(d_s, d_e, npat, static_cast<uint32_t *>(o.counts));
        LAUNCH_CHECK();
    }
    if (!want_off) return 0;
    if (o.width == 8) {
        rc = locate_async(idx, buf, prefix_only, d_s, d_e, npat, static_cast<uint64_t *>(o.hit_off),
                          static_cast<uint64_t *>(o.positions), static_cast<uint64_t *>(o.piece_ids), o.capacity, st, count_work);
        phase_mark(idx, 6, st);
        return rc;
    }
    if ((rc = buf[B_W64].ensure((npat + 1) * 8))) return rc;
    uint64_t *w_off = buf[B_W64].as<uint64_t>(), *w_pos = nullptr, *w_pid = nullptr;
    if (o.positions) {
        if ((rc = buf[B_POS].ensure(o.capacity * 8 + 8))) return rc;
        w_pos = buf[B_POS].as<uint64_t>();
    }
    if (o.piece_ids) {
        if ((rc = buf[B_PID].ensure(o.capacity * 8 + 8))) return rc;
        w_pid = buf[B_PID].as<uint64_t>();
    }
    if ((rc = locate_async(idx, buf, prefix_only, d_s, d_e, npat, w_off, w_pos, w_pid, o.capacity, st, count_work))) return rc;
    k_convert<uint64_t, uint32_t><<<grid_for(npat + 1, 256), 256, 0, st>>>(w_off, npat + 1, 0, static_cast<uint32_t *>(o.hit_off));
    if (w_pos) k_convert<uint64_t, uint32_t><<<grid_for(o.capacity, 256), 256, 0, st>>>(w_pos, o.capacity, 0, static_cast<uint32_t *>(o.positions));
    if (w_pid) k_convert<uint64_t, uint32_t><<<grid_for(o.capacity, 256), 256, 0, st>>>(w_pid, o.capacity, 0, static_cast<uint32_t *>(o.piece_ids));
    LAUNCH_CHECK();
    return 0;
}

static int query_args_ok(const fmx_index *idx, const fmx_query *q) {
    if (!idx || !q) return fail(FMX_ERR_INVALID_ARG, "null argument");
    if (q->npat && !q->patterns && (q->pat_off ? true : q->fixed_len != 0)) return fail(FMX_ERR_INVALID_ARG, "null pattern buffer");
    if (q->out_width != 8 && q->out_width != 4) return fail(FMX_ERR_INVALID_ARG, "out_width must be 8 or 4");
    if ((q->positions || q->piece_ids) && !q->hit_off) return fail(FMX_ERR_INVALID_ARG, "positions / piece_ids need hit_off");
    if (q->packed_bits && (q->pat_off || !q->fixed_len)) return fail(FMX_ERR_INVALID_ARG, "packed patterns need a fixed length");
    return 0;
}

extern "C" int fmx_query_batch_device(const fmx_index *idx, const fmx_query *q, void *stream) {
    int rc = query_args_ok(idx, q);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    PatSrc ps;
    if (q->packed_bits) ps.packed = static_cast<const uint64_t *>(q->patterns);
    else ps.pat = static_cast<const uint8_t *>(q->patterns);
    ps.packed_bits = q->packed_bits;
    ps.pat_off = q->pat_off;
    ps.fixed_len = q->fixed_len;
    QueryOut o;
    o.width = q->out_width;
    o.out_s = q->out_s;
    o.out_e = q->out_e;
    o.counts = q->counts;
    o.hit_off = q->hit_off;
    o.positions = q->positions;
    o.piece_ids = q->piece_ids;
    o.capacity = q->capacity;
    return query_device(idx, idx->buf, q->mode, ps, q->npat, o, pick_stream(idx, stream), true);
}

// ------------------------------------------------------------------ the same for host buffers: chunked pipeline

static int ensure_lanes(const fmx_index *idx) {
    if (idx->lanes_ready) return 0;
    for (auto &l : idx->lane) {
        CUDA_TRY(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&l.ev, cudaEventDisableTiming));
        CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&l.h_total), 64));
    }
    idx->lanes_ready = true;
    return 0;
}

// The batch is cut into chunks that flow through FMX_LANES lanes (stream + scratch each): stage 1 of a chunk
// (H2D patterns, the whole query_device with a guessed position capacity, hit total -> pinned word, event) is
// enqueued FMX_LANES - 1 chunks ahead of stage 2 (wait for the total, re-emit in the rare case the guess was too
// small, rebase the offsets, D2H of exactly the bytes produced).  The host waits once per chunk, on an event
// that is usually already complete.
extern "C" int fmx_query_batch(const fmx_index *idx, const fmx_query *q, uint64_t *total_hits) {
    int rc = query_args_ok(idx, q);
    if (rc) return rc;
    if (total_hits) *total_hits = 0;
    const uint32_t W = q->out_width;
    const uint64_t npat = q->npat;
    auto put = [&](void *base, uint64_t i, uint64_t v) {
        if (W == 8) static_cast<uint64_t *>(base)[i] = v; else static_cast<uint32_t *>(base)[i] = (uint32_t)v;
    };
    if (q->hit_off) put(q->hit_off, 0, 0);
    const bool want_hits = q->positions != nullptr || q->piece_ids != nullptr;
    if (want_hits && (rc = locate_args_ok(idx, q->piece_ids != nullptr))) return rc;
    if (npat == 0) return FMX_OK;
    const uint8_t *pat8 = static_cast<const uint8_t *>(q->patterns);
    const uint64_t wpp = q->packed_bits ? (q->fixed_len * q->packed_bits + 63) / 64 : 0;
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    if ((rc = ensure_lanes(idx))) return rc;
    // chunks: enough of them to overlap copies with kernels, big enough to fill the GPU
    uint64_t chunk = (npat + 3) / 4;
    if (chunk < (1ull << 16)) chunk = 1ull << 16;
    if (chunk > (1ull << 23)) chunk = 1ull << 23;
    if (idx->opt_pipeline_chunk) chunk = idx->opt_pipeline_chunk;
    const uint64_t nchunks = (npat + chunk - 1) / chunk;
    uint64_t running = 0;
    bool overflow = false;
    uint64_t guess_num = 2, guess_den = 1;  // position capacity of a chunk = n * num / den (refined from the chunks seen)

    auto chunk_src = [&](uint64_t lo, uint64_t n, Lane &L, PatSrc &ps, int &r) {
        r = 0;
        if (q->packed_bits) {
            const uint64_t nbytes = n * wpp * 8;
            if ((r = L.buf[B_PAT].ensure(nbytes + 16))) return;
            if (cudaMemcpyAsync(L.buf[B_PAT].p, pat8 + lo * wpp * 8, nbytes, cudaMemcpyHostToDevice, L.st) != cudaSuccess) {
                r = fail(FMX_ERR_CUDA, "H2D copy of the patterns failed");
                return;
            }
            ps.packed = L.buf[B_PAT].as<uint64_t>();
            ps.packed_bits = q->packed_bits;
            ps.fixed_len = q->fixed_len;
            return;
        }
        const uint64_t cw = char_width_of(idx);
        const uint64_t off0 = (q->pat_off ? q->pat_off[lo] : lo * q->fixed_len) * cw;
        const uint64_t nbytes = (q->pat_off ? q->pat_off[lo + n] : (lo + n) * q->fixed_len) * cw - off0;
        if ((r = L.buf[B_PAT].ensure(nbytes + 16))) return;
        if (nbytes && cudaMemcpyAsync(L.buf[B_PAT].p, pat8 + off0, nbytes, cudaMemcpyHostToDevice, L.st) != cudaSuccess) {
            r = fail(FMX_ERR_CUDA, "H2D copy of the patterns failed");
            return;
        }
        ps.pat = L.buf[B_PAT].as<uint8_t>();
        ps.fixed_len = q->fixed_len;
        if (q->pat_off) {
            if ((r = L.buf[B_OFF].ensure((n + 1) * 8))) return;
            if (cudaMemcpyAsync(L.buf[B_OFF].p, q->pat_off + lo, (n + 1) * 8, cudaMemcpyHostToDevice, L.st) != cudaSuccess) {
                r = fail(FMX_ERR_CUDA, "H2D copy of the pattern offsets failed");
                return;
            }
            ps.pat_off = L.buf[B_OFF].as<uint64_t>();
            ps.pat -= off0;  // the kernel adds the batch-global byte offsets
        }
    };
    auto lane_out = [&](Lane &L, uint64_t n, uint64_t cap, QueryOut &o) -> int {
        int r;
        o.width = W;
        if (q->out_s) {
            if ((r = L.buf[B_S].ensure(n * 8))) return r;
            if ((r = L.buf[B_E].ensure(n * 8))) return r;
            o.out_s = L.buf[B_S].as<uint64_t>();
            o.out_e = L.buf[B_E].as<uint64_t>();
        }
        if (q->counts) {
            if ((r = L.buf[B_CNT].ensure(n * 8))) return r;
            o.counts = L.buf[B_CNT].p;
        }
        if (q->hit_off) {
            if ((r = L.buf[B_HOFF].ensure((n + 1) * 8))) return r;
            o.hit_off = L.buf[B_HOFF].p;
        }
        if (q->positions) {
            if ((r = L.buf[B_OUT8].ensure(cap * W + 16))) return r;
            o.positions = L.buf[B_OUT8].p;
        }
        if (q->piece_ids) {
            if ((r = L.buf[B_OUT32].ensure(cap * W + 16))) return r;
            o.piece_ids = L.buf[B_OUT32].p;
        }
        o.capacity = cap;
        return 0;
    };
    auto stage1 = [&](uint64_t c) -> int {
        Lane &L = idx->lane[c % FMX_LANES];
        const uint64_t lo = c * chunk, n = (lo + chunk < npat ? chunk : npat - lo);
        PatSrc ps;
        int r;
        chunk_src(lo, n, L, ps, r);
        if (r) return r;
        uint64_t cap = want_hits ? n * guess_num / guess_den + 4096 : 0;
        QueryOut o;
        if ((r = lane_out(L, n, cap, o))) return r;
        L.pos_cap = cap;
        // work counters (option "count_work") are one set per index: only a batch that is a single chunk counts
        if ((r = query_device(idx, L.buf, q->mode, ps, n, o, L.st, nchunks == 1))) return r;
        if (q->out_s) {
            CUDA_TRY(cudaMemcpyAsync(q->out_s + lo, o.out_s, n * 8, cudaMemcpyDeviceToHost, L.st));
            CUDA_TRY(cudaMemcpyAsync(q->out_e + lo, o.out_e, n * 8, cudaMemcpyDeviceToHost, L.st));
        }
        if (q->counts) CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(q->counts) + lo * W, o.counts, n * W, cudaMemcpyDeviceToHost, L.st));
        if (q->hit_off) {
            *L.h_total = 0;
            CUDA_TRY(cudaMemcpyAsync(L.h_total, static_cast<uint8_t *>(o.hit_off) + n * W, W, cudaMemcpyDeviceToHost, L.st));
        }
        CUDA_TRY(cudaEventRecord(L.ev, L.st));
        return 0;
    };
    auto stage2 = [&](uint64_t c) -> int {
        Lane &L = idx->lane[c % FMX_LANES];
        const uint64_t lo = c * chunk, n = (lo + chunk < npat ? chunk : npat - lo);
        if (!q->hit_off) return 0;
        CUDA_TRY(cudaEventSynchronize(L.ev));
        const uint64_t hits = *L.h_total;
        int r;
        if (W == 4 && running + hits > 0xFFFFFFFFull) overflow = true;
        if (want_hits && hits > L.pos_cap && running + hits <= q->capacity && !overflow) {
            // the guess was too small for this chunk: emit again with room for every hit (the patterns and the
            // search results of the chunk are still in the lane's buffers)
            PatSrc ps;
            QueryOut o;
            if ((r = lane_out(L, n, hits, o))) return r;
            L.pos_cap = hits;
            if (q->packed_bits) {
                ps.packed = L.buf[B_PAT].as<uint64_t>();
                ps.packed_bits = q->packed_bits;
            } else {
                ps.pat = L.buf[B_PAT].as<uint8_t>();
                if (q->pat_off) {
                    ps.pat_off = L.buf[B_OFF].as<uint64_t>();
                    ps.pat -= q->pat_off[lo] * char_width_of(idx);
                }
            }
            ps.fixed_len = q->fixed_len;
            o.counts = nullptr;
            if ((r = query_device(idx, L.buf, q->mode, ps, n, o, L.st, false))) return r;
        }
        if (hits * 2 > n * guess_num / guess_den) {  // later chunks: twice what this one needed
            guess_num = (hits * 2 + n - 1) / n + 1;
            guess_den = 1;
        }
        // offsets of the chunk, rebased to the batch
        if ((r = L.buf[B_INS].ensure((n + 1) * 8))) return r;
        if (W == 8) k_convert<uint64_t, uint64_t><<<grid_for(n + 1, 256), 256, 0, L.st>>>(L.buf[B_HOFF].as<uint64_t>(), n + 1, running, L.buf[B_INS].as<uint64_t>());
        else k_convert<uint32_t, uint32_t><<<grid_for(n + 1, 256), 256, 0, L.st>>>(L.buf[B_HOFF].as<uint32_t>(), n + 1, running, L.buf[B_INS].as<uint32_t>());
        LAUNCH_CHECK();
        CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(q->hit_off) + lo * W, L.buf[B_INS].p, (n + 1) * W, cudaMemcpyDeviceToHost, L.st));
        if (want_hits && hits) {
            if (running + hits > q->capacity) {
                overflow = true;
            } else if (!overflow) {
                if (q->positions) CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(q->positions) + running * W, L.buf[B_OUT8].p, hits * W, cudaMemcpyDeviceToHost, L.st));
                if (q->piece_ids) CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(q->piece_ids) + running * W, L.buf[B_OUT32].p, hits * W, cudaMemcpyDeviceToHost, L.st));
            }
        }
        running += hits;
        return 0;
    };
    rc = 0;
    for (uint64_t c = 0; c < nchunks + FMX_LANES - 1 && rc == 0; c++) {
        if (c < nchunks) rc = stage1(c);
        if (rc == 0 && c >= FMX_LANES - 1 && c - (FMX_LANES - 1) < nchunks) rc = stage2(c - (FMX_LANES - 1));
    }
    for (auto &l : idx->lane) cudaStreamSynchronize(l.st);
    if (rc) return rc;
    if (total_hits) *total_hits = running;
    if ((rc = check_err_flag(idx, idx->lane[0].st))) return rc;
    if (overflow) return fail(FMX_ERR_CAPACITY, "output buffer too small (or more than 2^32 - 1 hits with out_width 4): total_hits holds the size needed");
    return FMX_OK;
}

extern "C" int fmx_search_locate_batch(const fmx_index *idx, int mode, const uint8_t *pat, const uint64_t *pat_off,
                                       uint64_t fixed_len, uint64_t npat, uint64_t *out_s, uint64_t *out_e,
                                       uint64_t *hit_off, uint64_t *positions, uint64_t *piece_ids, uint64_t capacity,
                                       uint64_t *total_hits) {
    if (!idx || !hit_off || !total_hits) return fail(FMX_ERR_INVALID_ARG, "null argument");
    if ((out_s == nullptr) != (out_e == nullptr)) return fail(FMX_ERR_INVALID_ARG, "out_s and out_e must both be given");
    fmx_query q;
    std::memset(&q, 0, sizeof(q));
    q.mode = mode;
    q.patterns = pat;
    q.pat_off = pat_off;
    q.fixed_len = fixed_len;
    q.npat = npat;
    q.out_width = 8;
    q.out_s = out_s;
    q.out_e = out_e;
    q.hit_off = hit_off;
    q.positions = positions;
    q.piece_ids = piece_ids;
    q.capacity = capacity;
    return fmx_query_batch(idx, &q, total_hits);
}

// ------------------------------------------------------------------ CSR merge (piece-partitioned MultiPieces)
// Each part (one per partition of the pieces) answered EVERY pattern on its own index: part r holds a CSR
// (hit_off_r, positions_r, piece_ids_r) in its local coordinates.  A pattern without \0 cannot span two pieces, so
// the match set of the whole text is the disjoint union of the parts' sets (multi_pieces.rs:188-223) and counts
// add.  One scan over the interleaved counts C[p * R + r] gives every (pattern, part) its slot in the merged CSR;
// one kernel moves the hits there, shifting positions and piece ids into the coordinates of the whole text.
// Hits of one pattern come part-major, inside a part in that part's SA-row order.

#define FMX_MAX_PARTS 16
struct MergeParts {
    const void *off[FMX_MAX_PARTS];
    const void *pos[FMX_MAX_PARTS];
    const void *pid[FMX_MAX_PARTS];
    uint64_t pos_base[FMX_MAX_PARTS];
    uint64_t pid_base[FMX_MAX_PARTS];
    int n;
};

template <class T>
__global__ void k_merge_counts(const __grid_constant__ MergeParts parts, uint64_t npat, uint32_t *cnt) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npat * parts.n) return;
    const uint64_t p = t / parts.n;
    const int r = (int)(t % parts.n);
    const T *off = static_cast<const T *>(parts.off[r]);
    cnt[t] = (uint32_t)(off[p + 1] - off[p]);
}

template <class T>
__global__ void k_merge_scatter(const __grid_constant__ MergeParts parts, uint64_t npat, const uint32_t *cnt, const uint64_t *slot,
                                uint64_t *hit_off, uint64_t *positions, uint64_t *piece_ids, uint64_t capacity) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = npat * parts.n;
    if (t > total) return;
    if (t == total) {
        hit_off[npat] = slot[total];
        return;
    }
    const uint64_t p = t / parts.n;
    const int r = (int)(t % parts.n);
    const uint64_t dst = slot[t];
    if (r == 0) hit_off[p] = dst;
    const uint32_t c = cnt[t];
    if (!c) return;
    const uint64_t src = static_cast<const T *>(parts.off[r])[p];
    const T *pos = static_cast<const T *>(parts.pos[r]);
    const T *pid = static_cast<const T *>(parts.pid[r]);
    for (uint32_t j = 0; j < c; j++) {
        if (dst + j >= capacity) break;
        if (positions && pos) positions[dst + j] = (uint64_t)pos[src + j] + parts.pos_base[r];
        if (piece_ids && pid) piece_ids[dst + j] = (uint64_t)pid[src + j] + parts.pid_base[r];
    }
}

static std::mutex g_merge_mu;
static DevBuf g_merge_buf[64][3];  // per device: counts, slots, scan tiles

extern "C" int fmx_csr_merge_device(int device, const fmx_csr_part *parts, int nparts, uint64_t npat, uint32_t width,
                                    uint64_t *d_hit_off, uint64_t *d_positions, uint64_t *d_piece_ids, uint64_t capacity,
                                    void *stream) {
    if (!parts || !d_hit_off) return fail(FMX_ERR_INVALID_ARG, "null argument");
    if (nparts < 1 || nparts > FMX_MAX_PARTS) return fail(FMX_ERR_INVALID_ARG, "1 .. 16 parts");
    if (width != 8 && width != 4) return fail(FMX_ERR_INVALID_ARG, "width must be 8 or 4");
    if (device < 0 || device >= 64) return fail(FMX_ERR_INVALID_ARG, "device ordinal out of range");
    if (npat * (uint64_t)nparts >= 0xFFFFFFFFull * 4) return fail(FMX_ERR_UNSUPPORTED, "too many (pattern, part) pairs");
    MergeParts mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.n = nparts;
    for (int r = 0; r < nparts; r++) {
        if (!parts[r].hit_off) return fail(FMX_ERR_INVALID_ARG, "part without hit offsets");
        mp.off[r] = parts[r].hit_off;
        mp.pos[r] = parts[r].positions;
        mp.pid[r] = parts[r].piece_ids;
        mp.pos_base[r] = parts[r].position_base;
        mp.pid_base[r] = parts[r].piece_base;
    }
    std::lock_guard<std::mutex> lk(g_merge_mu);
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const uint64_t total = npat * (uint64_t)nparts;
    DevBuf *b = g_merge_buf[device];
    int rc;
    if ((rc = b[0].ensure(total * 4 + 16))) return rc;
    if ((rc = b[1].ensure((total + 1) * 8))) return rc;
    uint32_t *d_cnt = b[0].as<uint32_t>();
    uint64_t *d_slot = b[1].as<uint64_t>();
    if (total) {
        if (width == 8) k_merge_counts<uint64_t><<<grid_for(total, 256), 256, 0, st>>>(mp, npat, d_cnt);
        else k_merge_counts<uint32_t><<<grid_for(total, 256), 256, 0, st>>>(mp, npat, d_cnt);
        LAUNCH_CHECK();
    }
    if ((rc = device_scan<uint32_t, uint64_t, OpSum, true>(d_cnt, total, d_slot, OpSum(), true, b[2], st))) return rc;
    if (width == 8) k_merge_scatter<uint64_t><<<grid_for(total + 1, 256), 256, 0, st>>>(mp, npat, d_cnt, d_slot, d_hit_off, d_positions, d_piece_ids, capacity);
    else k_merge_scatter<uint32_t><<<grid_for(total + 1, 256), 256, 0, st>>>(mp, npat, d_cnt, d_slot, d_hit_off, d_positions, d_piece_ids, capacity);
    LAUNCH_CHECK();
    return FMX_OK;
}

// ------------------------------------------------------------------ extraction / primitives

static int extract_device(const fmx_index *idx, const uint64_t *d_rows, uint64_t nrows, uint32_t k, int forward,
                          uint8_t *d_out, uint32_t *d_len, cudaStream_t st) {
    if (nrows == 0 || k == 0) return 0;
    if (idx->dev.text && has_dense_sa(idx) && idx->opt_extract_text && idx->dev.cw_shift == 0) {
        // HBM-rich indexes: the characters are read from the text at SA[row] (k_extract_text)
        const unsigned gt = grid_for(nrows * FMX_XT_LANES, 256);
        if (idx->hdr.kind == FMX_KIND_MULTI) k_extract_text<FMX_KIND_MULTI_><<<gt, 256, 0, st>>>(idx->dev, d_rows, nrows, k, forward, d_out, d_len);
        else k_extract_text<FMX_KIND_FM_><<<gt, 256, 0, st>>>(idx->dev, d_rows, nrows, k, forward, d_out, d_len);
        LAUNCH_CHECK();
        return 0;
    }
    unsigned g = grid_for(nrows, 256);
    dispatch(idx, [&](auto K, auto LY) { k_extract<K(), LY()><<<g, 256, 0, st>>>(idx->dev, d_rows, nrows, k, forward, d_out, d_len); });
    LAUNCH_CHECK();
    return 0;
}

extern "C" int fmx_extract_batch_device(const fmx_index *idx, const uint64_t *d_rows, uint64_t nrows, uint32_t k,
                                        int forward, uint8_t *d_out, uint32_t *d_out_len, void *stream) {
    if (!idx || (!d_rows && nrows) || (!d_out && nrows && k)) return fail(FMX_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(idx->device));
    return extract_device(idx, d_rows, nrows, k, forward, d_out, d_out_len, pick_stream(idx, stream));
}

extern "C" int fmx_extract_batch(const fmx_index *idx, const uint64_t *rows, uint64_t nrows, uint32_t k, int forward,
                                 uint8_t *out, uint32_t *out_len) {
    if (!idx || (!rows && nrows) || (!out && nrows && k)) return fail(FMX_ERR_INVALID_ARG, "null argument");
    if (nrows == 0) return FMX_OK;
    for (uint64_t r = 0; r < nrows; r++)
        if (rows[r] >= idx->hdr.n) return fail(FMX_ERR_INVALID_ARG, "row out of range");
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    int rc;
    if ((rc = idx->buf[B_S].ensure(nrows * 8))) return rc;
    const uint64_t out_bytes = nrows * (uint64_t)k * char_width_of(idx);  // characters leave in the text's width
    if ((rc = idx->buf[B_OUT8].ensure(out_bytes + 16))) return rc;
    if ((rc = idx->buf[B_OUT32].ensure(nrows * 4))) return rc;
    CUDA_TRY(cudaMemcpyAsync(idx->buf[B_S].p, rows, nrows * 8, cudaMemcpyHostToDevice, st));
    if ((rc = extract_device(idx, idx->buf[B_S].as<uint64_t>(), nrows, k, forward, idx->buf[B_OUT8].as<uint8_t>(),
                             idx->buf[B_OUT32].as<uint32_t>(), st)))
        return rc;
    if (k) CUDA_TRY(cudaMemcpyAsync(out, idx->buf[B_OUT8].p, out_bytes, cudaMemcpyDeviceToHost, st));
    if (out_len) {
        if (k) CUDA_TRY(cudaMemcpyAsync(out_len, idx->buf[B_OUT32].p, nrows * 4, cudaMemcpyDeviceToHost, st));
        else std::memset(out_len, 0, nrows * 4);
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return FMX_OK;
}

extern "C" int fmx_rows_op(const fmx_index *idx, int op, const uint64_t *rows, uint64_t nrows, uint64_t *out) {
    if (!idx || (!rows && nrows) || (!out && nrows)) return fail(FMX_ERR_INVALID_ARG, "null argument");
    if (op < 0 || op > 5) return fail(FMX_ERR_INVALID_ARG, "bad op");
    if (nrows == 0) return FMX_OK;
    if (op == 4 && !idx->hdr.has_locate) return fail(FMX_ERR_NO_LOCATE, "index has no sampled suffix array");
    if (op == 5 && idx->hdr.kind != FMX_KIND_MULTI) return fail(FMX_ERR_UNSUPPORTED, "piece_id needs a MultiPieces index");
    for (uint64_t r = 0; r < nrows; r++)
        if (rows[r] >= idx->hdr.n) return fail(FMX_ERR_INVALID_ARG, "row out of range");
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    int rc;
    if ((rc = idx->buf[B_S].ensure(nrows * 8))) return rc;
    if ((rc = idx->buf[B_E].ensure(nrows * 8))) return rc;
    CUDA_TRY(cudaMemcpyAsync(idx->buf[B_S].p, rows, nrows * 8, cudaMemcpyHostToDevice, st));
    unsigned g = grid_for(nrows, 256);
    const uint64_t *d_rows = idx->buf[B_S].as<uint64_t>();
    uint64_t *d_out = idx->buf[B_E].as<uint64_t>();
    dispatch(idx, [&](auto K, auto LY) { k_rows_op<K(), LY()><<<g, 256, 0, st>>>(idx->dev, op, d_rows, nrows, d_out); });
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(out, d_out, nrows * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return FMX_OK;
}

extern "C" int fmx_lf_map2_batch(const fmx_index *idx, const uint8_t *c, const uint64_t *i, uint64_t nrows, uint64_t *out) {
    if (!idx || ((!c || !i || !out) && nrows)) return fail(FMX_ERR_INVALID_ARG, "null argument");
    if (nrows == 0) return FMX_OK;
    // characters arrive in the index's character width; the kernel takes them as u32
    const uint64_t cw = char_width_of(idx);
    std::vector<uint32_t> c32(nrows);
    for (uint64_t r = 0; r < nrows; r++) {
        if (i[r] > idx->hdr.n) return fail(FMX_ERR_INVALID_ARG, "row out of range");
        const uint64_t v = cw == 1 ? c[r]
                         : cw == 2 ? reinterpret_cast<const uint16_t *>(c)[r]
                         : cw == 4 ? reinterpret_cast<const uint32_t *>(c)[r] : reinterpret_cast<const uint64_t *>(c)[r];
        if (v > idx->hdr.max_character) return fail(FMX_ERR_PATTERN_CHAR, "character larger than max_character");
        c32[r] = (uint32_t)v;
    }
    std::lock_guard<std::mutex> lk(idx->mu);
    CUDA_TRY(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    int rc;
    if ((rc = idx->buf[B_S].ensure(nrows * 8))) return rc;
    if ((rc = idx->buf[B_E].ensure(nrows * 8))) return rc;
    if ((rc = idx->buf[B_PAT].ensure(nrows * 4 + 16))) return rc;
    CUDA_TRY(cudaMemcpyAsync(idx->buf[B_S].p, i, nrows * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(idx->buf[B_PAT].p, c32.data(), nrows * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));  // c32 is a local
    unsigned g = grid_for(nrows, 256);
    const uint32_t *d_c = idx->buf[B_PAT].as<uint32_t>();
    const uint64_t *d_i = idx->buf[B_S].as<uint64_t>();
    uint64_t *d_out = idx->buf[B_E].as<uint64_t>();
    dispatch(idx, [&](auto K, auto LY) { k_lf_map2<K(), LY()><<<g, 256, 0, st>>>(idx->dev, d_c, d_i, nrows, d_out); });
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(out, d_out, nrows * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return FMX_OK;
}

// ------------------------------------------------------------------ measurement helpers

// Milliseconds of the phases of the last query with option "phase_timing" on: [0] seed (table lookups; the whole
// k_search on indexes without the phased kernels), [1] steps, [2] verify, [3] second steps pass (+ row widening),
// [4] counts + offsets scan, [5] emit / locate.  Phases that did not run are 0.  Synchronises the stream.
extern "C" int fmx_last_phase_ms(const fmx_index *idx, void *stream, float *ms, int n) {
    if (!idx || !ms) return fail(FMX_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(idx->device));
    CUDA_TRY(cudaStreamSynchronize(pick_stream(idx, stream)));
    for (int k = 0; k < n; k++) ms[k] = 0.f;
    int prev = (idx->pev_mask & 1u) ? 0 : -1;
    for (int k = 1; k <= FMX_NPHASE && prev >= 0; k++) {
        if (!(idx->pev_mask & (1u << k))) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, idx->pev[prev], idx->pev[k]) != cudaSuccess) {
            cudaGetLastError();
            t = 0.f;
        }
        if (k - 1 < n) ms[k - 1] = t;
        prev = k;
    }
    return FMX_OK;
}

extern "C" int fmx_last_work(const fmx_index *idx, void *stream, uint64_t *search_steps, uint64_t *lf_steps) {
    if (!idx) return fail(FMX_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(idx->device));
    cudaStream_t st = pick_stream(idx, stream);
    unsigned long long w[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(w, idx->d_work, sizeof(w), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (search_steps) *search_steps = w[0];
    if (lf_steps) *lf_steps = w[1];
    return FMX_OK;
}

// Index requests the phased kernels of the last query issued (option "count_work"): lane-level loads of table
// entries, rank blocks, suffix-array / inverse entries and text sectors by the search phases, and of suffix-array
// entries by the emit kernels.  An upper bound on the L2 requests of those kernels (neighbouring lanes share
// lines); 0 on indexes that do not run the phased kernels.  Synchronises the stream.
extern "C" int fmx_last_requests(const fmx_index *idx, void *stream, uint64_t *search_requests, uint64_t *emit_requests) {
    if (!idx) return fail(FMX_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(idx->device));
    cudaStream_t st = pick_stream(idx, stream);
    unsigned long long w[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(w, idx->d_work, sizeof(w), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (search_requests) *search_requests = w[2];
    if (emit_requests) *emit_requests = w[3];
    return FMX_OK;
}

extern "C" int fmx_random_gather_bench(int device, uint64_t bytes, uint64_t nloads, int iters, uint32_t load_bytes,
                                       double *sectors_per_s) {
    if (!sectors_per_s || bytes < 4096 || nloads == 0 || iters <= 0) return fail(FMX_ERR_INVALID_ARG, "bad argument");
    if (load_bytes != 32 && load_bytes != 64 && load_bytes != 128) return fail(FMX_ERR_INVALID_ARG, "load_bytes must be 32, 64 or 128");
    if (bytes / 32 >= 0xFFFFFFFFull) return fail(FMX_ERR_INVALID_ARG, "buffer too large (max 128 GiB)");
    CUDA_TRY(cudaSetDevice(device));
    void *buf = nullptr;
    uint32_t *sink = nullptr;
    CUDA_TRY(cudaMalloc(&buf, bytes));
    CUDA_TRY(cudaMalloc(&sink, 256));
    CUDA_TRY(cudaMemset(buf, 1, bytes));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    unsigned grid = (unsigned)sms * 8;
    uint64_t nunits = bytes / load_bytes;
    double best = 0;
    for (int it = 0; it < iters + 1; it++) {
        CUDA_TRY(cudaEventRecord(e0, 0));
        const uint4 *b4 = static_cast<const uint4 *>(buf);
        uint64_t sd = 0x1234567ull * (it + 1);
        if (load_bytes == 32) k_random_gather<32><<<grid, 256>>>(b4, nunits, nloads, sd, sink);
        else if (load_bytes == 64) k_random_gather<64><<<grid, 256>>>(b4, nunits, nloads, sd, sink);
        else k_random_gather<128><<<grid, 256>>>(b4, nunits, nloads, sd, sink);
        LAUNCH_CHECK();
        CUDA_TRY(cudaEventRecord(e1, 0));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double rate = (double)nloads * (load_bytes / 32) / (ms * 1e-3);
        if (it > 0 && rate > best) best = rate;  // first iteration is warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(sink);
    *sectors_per_s = best;
    return FMX_OK;
}
