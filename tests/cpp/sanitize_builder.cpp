// ASan + UBSan harness for the HOST builder (fm-index_b200/csrc/builder.cpp + sais.hpp): no device, no CUDA.
// tests/test_sanitizers.py compiles it with -fsanitize=address,undefined -fno-sanitize-recover and runs it: every kind x
// alphabet x layout x mode the host builder serves, the wide-character builder, the suffix sorters, the text validation
// errors, and check_blob on truncated / corrupted blobs (a loaded blob's header must not be trusted).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../fm-index_b200/csrc/builder.h"
#include "../../include/fmx.h"

namespace fmx {
// the one device-side symbol builder.cpp refers to (gpu_sa.cu); never reached with sa_device = -1
int gpu_suffix_array(const uint8_t *, uint64_t, uint32_t, int, uint32_t *, int *, std::string &err) {
    err = "no device in the sanitizer harness";
    return FMX_ERR_CUDA;
}
}  // namespace fmx

static int failures = 0;
#define EXPECT(cond)                                                        \
    do {                                                                    \
        if (!(cond)) {                                                      \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                     \
        }                                                                   \
    } while (0)

static std::vector<uint8_t> make_text(std::mt19937_64 &rng, size_t n, unsigned sigma, bool multi, bool repetitive) {
    std::vector<uint8_t> t(n + 1);
    std::vector<uint8_t> base(257);
    for (auto &b : base) b = (uint8_t)(1 + rng() % sigma);
    for (size_t i = 0; i < n; i++) t[i] = repetitive ? base[i % base.size()] : (uint8_t)(1 + rng() % sigma);
    if (repetitive)
        for (size_t i = 0; i < n; i += 97) t[i] = (uint8_t)(1 + rng() % sigma);
    if (multi)
        for (size_t k = 0; k < n / 300 + 1; k++) t[(2 + rng() % (n - 4)) / 2 * 2] = 0;   // never two zeros in a row
    t[n] = 0;
    return t;
}

static void build_and_check(const std::vector<uint8_t> &text, uint64_t mc, int kind, int level, int mode) {
    fmx::HostBlob blob;
    std::string err;
    const int rc = fmx::build_blob(text.data(), text.size(), mc, kind, level, blob, err, -1, mode);
    EXPECT(rc == 0);
    if (rc) {
        std::fprintf(stderr, "  build_blob(kind %d, mc %llu, level %d, mode %d, n %zu): %s\n", kind, (unsigned long long)mc, level, mode,
                     text.size(), err.c_str());
        return;
    }
    FmxBlobHeader hdr;
    EXPECT(fmx::check_blob(blob.p, blob.n, hdr, err) == 0);
    EXPECT(hdr.n == text.size() && hdr.kind == (uint32_t)kind);
    // truncated and corrupted copies must be refused, not read out of bounds
    std::vector<uint8_t> copy(blob.p, blob.p + blob.n);
    EXPECT(fmx::check_blob(copy.data(), copy.size() / 2, hdr, err) != 0);
    EXPECT(fmx::check_blob(copy.data(), sizeof(FmxBlobHeader) - 1, hdr, err) != 0);
    std::mt19937_64 rng(copy.size());
    for (int t = 0; t < 64; t++) {
        std::vector<uint8_t> bad = copy;
        const size_t at = 8 + rng() % (sizeof(FmxBlobHeader) - 8);   // somewhere in the header, past the magic
        bad[at] ^= (uint8_t)(1u << (rng() % 8));
        FmxBlobHeader h2;
        (void)fmx::check_blob(bad.data(), bad.size(), h2, err);       // either verdict is fine; it must not crash
    }
}

int main() {
    std::mt19937_64 rng(20260101);
    const int kinds[3] = {FMX_KIND_FM, FMX_KIND_RLFM, FMX_KIND_MULTI};
    const int modes[3] = {FMX_MODE_AUTO, FMX_MODE_COMPACT, FMX_MODE_RICH};
    // u8 texts: DNA-sized (Q4), small and byte alphabets (SYM), count-only and sampled, every mode
    for (int kind : kinds)
        for (unsigned sigma : {2u, 4u, 6u, 37u, 255u})
            for (int level : {-1, 0, 2, 5})
                for (int mode : modes) {
                    const size_t n = 500 + rng() % 5000;
                    auto text = make_text(rng, n, sigma, kind == FMX_KIND_MULTI, kind == FMX_KIND_RLFM);
                    build_and_check(text, sigma <= 4 ? 4 : (sigma == 255 ? 255 : sigma), kind, level, mode);
                }
    // the fallback layouts: binary wavelet matrix, quaternary wavelet matrix
    setenv("FMX_FORCE_WAVELET", "1", 1);
    for (int kind : kinds) build_and_check(make_text(rng, 3000, 6, kind == FMX_KIND_MULTI, false), 6, kind, 2, FMX_MODE_AUTO);
    unsetenv("FMX_FORCE_WAVELET");
    setenv("FMX_SYM_BUDGET_MB", "0", 1);
    for (int kind : kinds) build_and_check(make_text(rng, 3000, 37, kind == FMX_KIND_MULTI, false), 37, kind, 2, FMX_MODE_RICH);
    unsetenv("FMX_SYM_BUDGET_MB");
    // dense verify structures on a small text, and a single text with interior zeros (never gets them)
    setenv("FMX_VERIFY_MIN_RANK_MB", "0", 1);
    build_and_check(make_text(rng, 20000, 4, false, false), 4, FMX_KIND_FM, 2, FMX_MODE_AUTO);
    {
        auto t = make_text(rng, 6000, 4, true, false);
        build_and_check(t, 4, FMX_KIND_FM, 2, FMX_MODE_RICH);
    }
    unsetenv("FMX_VERIFY_MIN_RANK_MB");
    // tiny texts
    for (int kind : kinds) {
        build_and_check({1, 0}, 4, kind, 2, FMX_MODE_AUTO);
        if (kind != FMX_KIND_RLFM) build_and_check({0}, 4, kind, -1, FMX_MODE_AUTO);   // RLFM refuses it (unreachable!() at rlfmi.rs:62)
    }
    // wide characters: u16 / u32 / u64 over large alphabets (WIDE layout), and the wide suffix sorter
    for (uint32_t width : {2u, 4u, 8u})
        for (int kind : kinds) {
            const size_t n = 2000;
            const uint64_t mc = width == 2 ? 65535 : (width == 4 ? 5000011 : 70000);
            std::vector<uint64_t> sym(40);
            for (auto &v : sym) v = 1 + rng() % mc;
            std::vector<uint8_t> raw((n + 1) * width, 0);
            for (size_t i = 0; i < n; i++) {
                uint64_t c = sym[rng() % sym.size()];
                if (kind == FMX_KIND_MULTI && i % 211 == 210) c = 0;
                std::memcpy(raw.data() + i * width, &c, width);
            }
            fmx::HostBlob blob;
            std::string err;
            EXPECT(fmx::build_blob_wide(raw.data(), width, n + 1, mc, kind, 2, blob, err, FMX_MODE_AUTO) == 0);
            FmxBlobHeader hdr;
            EXPECT(blob.p && fmx::check_blob(blob.p, blob.n, hdr, err) == 0);
            std::vector<uint64_t> sa(n + 1);
            EXPECT(fmx::build_suffix_array_wide(raw.data(), width, n + 1, sa.data(), err) == 0);
        }
    // suffix sorter + validation errors (sais.rs:128-139)
    {
        auto t = make_text(rng, 10000, 4, false, true);
        std::vector<uint64_t> sa(t.size());
        std::string err;
        EXPECT(fmx::build_suffix_array(t.data(), t.size(), sa.data(), err) == 0);
        std::vector<uint8_t> lead = {0, 1, 2, 0}, twice = {1, 2, 0, 0}, open = {1, 2, 3};
        EXPECT(fmx::build_suffix_array(lead.data(), lead.size(), sa.data(), err) == FMX_ERR_INVALID_TEXT);
        EXPECT(fmx::build_suffix_array(twice.data(), twice.size(), sa.data(), err) == FMX_ERR_INVALID_TEXT);
        EXPECT(fmx::build_suffix_array(open.data(), open.size(), sa.data(), err) == FMX_ERR_INVALID_TEXT);
        fmx::HostBlob blob;
        EXPECT(fmx::build_blob(open.data(), open.size(), 4, FMX_KIND_FM, 2, blob, err) == FMX_ERR_INVALID_TEXT);
        std::vector<uint8_t> above = {1, 9, 0};
        EXPECT(fmx::build_blob(above.data(), above.size(), 4, FMX_KIND_FM, 2, blob, err) != 0);   // a character above max_character
    }
    std::printf("%s (%d failed expectations)\n", failures ? "FAILED" : "sanitize_builder ok", failures);
    return failures ? 1 : 0;
}
