// builder.h -- host-side index construction: text -> device-layout blob (fmx_layout.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "fmx_layout.h"

namespace fmx {

// The blob on the host: one malloc'd buffer (tens of GB for the SYM / verify layouts).  Zeroed by all cores at
// once -- first touch by one thread was the largest single cost of building the byte-alphabet index.
struct HostBlob {
    uint8_t *p = nullptr;
    uint64_t n = 0;
    HostBlob() = default;
    HostBlob(const HostBlob &) = delete;
    HostBlob &operator=(const HostBlob &) = delete;
    ~HostBlob();
    int alloc(uint64_t bytes, bool zero);  // 0 or FMX_ERR_OOM
    uint8_t *release() {
        uint8_t *q = p;
        p = nullptr;
        n = 0;
        return q;
    }
};

// Returns 0 or a negative fmx_status; on failure `err` holds the message
// (for invalid texts: the reference's Error::InvalidText strings, sais.rs:128-139).
// sa_device >= 0: build the suffix array on that GPU (gpu_sa.cu) when the text is large enough,
// otherwise with the host SA-IS.  The blob is byte-identical either way.
// mode: FMX_MODE_AUTO / _COMPACT / _RICH (include/fmx.h)
int build_blob(const uint8_t *text, uint64_t n, uint64_t max_character, int kind, int level,
               HostBlob &blob, std::string &err, int sa_device = -1, int mode = 0);

// Texts of u16 / u32 / u64 characters with max_character > 255: the WIDE layout (fmx_layout.h).  char_width = 2, 4 or 8.
int build_blob_wide(const void *text, uint32_t char_width, uint64_t n, uint64_t max_character, int kind, int level,
                    HostBlob &blob, std::string &err, int mode = 0);
int build_suffix_array_wide(const void *text, uint32_t char_width, uint64_t n, uint64_t *sa_out, std::string &err);

// pieces of build_blob shared with the GPU builder (gpu_build.cu)
struct VerifyPlan {
    bool verify = false, dense = false, dense_sa = false;
    uint32_t isa_level = 0;
};
// free bytes of the device being built for (0 = none): clamps the SYM and verify budgets (sized for 180 GB of HBM)
void set_device_memory_hint(uint64_t free_bytes);
int resolve_mode(int &mode, std::string &err);
VerifyPlan plan_verify(int kind, uint64_t n, int mode, int level, bool interior_zero, uint64_t rank_bytes, bool use_sym);
void layout_sections(FmxBlobHeader &hdr, const uint64_t bytes[SEC_COUNT]);
bool q4_forbidden_by_env();
bool sym_layout_chosen(uint64_t cs_len, uint64_t len);
uint64_t sym_layout_bytes(uint64_t cs_len, uint64_t len);

// gpu_build.cu: the whole blob built in device memory (Q4 and SYM layouts of FM / MultiPieces indexes).
// Returns 0 and a cudaMalloc'd blob + its header; FMX_ERR_UNSUPPORTED when this text / kind is not its case
// (the caller then takes the host builder).
int gpu_build_blob(const uint8_t *text, uint64_t n, uint64_t max_character, int kind, int level, int mode, int device,
                      void **d_blob_out, FmxBlobHeader *hdr_out, std::string &err);

int gpu_build_rlfm_blob(const uint8_t *text, uint64_t n, uint64_t max_character, int level, int mode, int device,
                        void **d_blob_out, FmxBlobHeader *hdr_out, std::string &err);

// gpu_sa.cu
int gpu_suffix_array_device(const uint8_t *d_text, uint64_t n, uint32_t bits, int device, uint32_t **d_sa_out, int *rounds_out,
                            std::string &err);
int gpu_suffix_array(const uint8_t *text, uint64_t n, uint32_t bits, int device, uint32_t *sa_out, int *rounds_out,
                     std::string &err);

// sais.rs:115-144 (validation + suffix array)
int build_suffix_array(const uint8_t *text, uint64_t n, uint64_t *sa_out, std::string &err);

// Validates a blob's header/section table; fills `hdr`.
int check_blob(const void *blob, uint64_t bytes, FmxBlobHeader &hdr, std::string &err);

}  // namespace fmx
