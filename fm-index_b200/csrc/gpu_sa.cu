// gpu_sa.cu -- suffix array construction on the GPU (prefix doubling over radix sorts).
//
// Construction is not the query hot path (the step BEFORE the path), but at
// GB scale the host SA-IS dominates everything else (143 s for 10^9 symbols on the GPU box's CPU),
// so the index builder uses the B200 for it: the suffix array is unique for a text, so the blob is
// byte-identical to the host-built one (tests/test_gpu_parity.py::test_gpu_suffix_array_*).
//
// Algorithm (Manber-Myers / Larsson-Sadakane doubling, all on device):
//   round 0 : key[i] = the first k symbols of suffix i packed MSB-first (k = 64 / bits-per-symbol),
//             radix-sort (key, i)
//   round r : rank[i] = 1 + index of the first suffix of i's group; key = (rank[i] << 32) | rank[i+h]
//             (0 when i + h is past the end: a suffix that ends is smaller than any continuation),
//             radix-sort again; h doubles.  Stop when every group is a single suffix.
// Result: plain lexicographic suffix order with \0 an ordinary smallest symbol, i.e. exactly what
// sais::build_suffix_array (reference src/suffix_array/sais.rs:115-144) returns.
// The radix sorts are CUB's (library code off the hot path); the scans are this repo's own (scan3.cuh).
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/fmx.h"
#include "scan3.cuh"

namespace fmx {

__global__ void k_sa_init(const uint8_t *text, uint64_t n, uint32_t bits, uint32_t k, uint64_t *key, uint32_t *sa) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t v = 0;
    for (uint32_t j = 0; j < k; j++) {
        uint64_t p = i + j;
        v = (v << bits) | (p < n ? (uint64_t)text[p] : 0ull);
    }
    key[i] = v;
    sa[i] = (uint32_t)i;
}

// head[j] = j + 1 where a new group starts, else 0; *nheads counts the groups
__global__ void k_sa_heads(const uint64_t *key, uint64_t n, uint32_t *head, unsigned long long *nheads) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool h = false;
    if (j < n) {
        h = j == 0 || key[j] != key[j - 1];
        head[j] = h ? (uint32_t)(j + 1) : 0u;
    }
    unsigned m = __ballot_sync(0xffffffffu, h);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(nheads, (unsigned long long)__popc(m));
}

__global__ void k_sa_scatter_rank(const uint32_t *sa, const uint32_t *grp, uint64_t n, uint32_t *rank) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) rank[sa[j]] = grp[j];
}

__global__ void k_sa_make_keys(const uint32_t *sa, const uint32_t *grp, const uint32_t *rank, uint64_t n, uint64_t h,
                               uint64_t *key) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t p = (uint64_t)sa[j] + h;
    uint32_t r2 = p < n ? rank[p] : 0u;
    key[j] = ((uint64_t)grp[j] << 32) | r2;
}

#define SA_TRY(expr)                                                                  \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            err = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
            if (_e == cudaErrorMemoryAllocation) rc = FMX_ERR_OOM;                    \
            goto fail;                                                                \
        }                                                                             \
    } while (0)

// d_text: DEVICE pointer, n symbols each < 2^bits.  *d_sa_out: the suffix array in device memory (cudaMalloc'd; the
// caller frees it).  Everything else is released before returning.
int gpu_suffix_array_device(const uint8_t *d_text, uint64_t n, uint32_t bits, int device, uint32_t **d_sa_out, int *rounds_out,
                            std::string &err) {
    *d_sa_out = nullptr;
    if (n == 0) return 0;
    if (n >= 0xFFFFFFFEull) {
        err = "text length must be below 2^32 - 2";
        return FMX_ERR_UNSUPPORTED;
    }
    uint64_t *d_key[2] = {nullptr, nullptr};
    uint32_t *d_sa[2] = {nullptr, nullptr}, *d_rank = nullptr, *d_grp = nullptr;
    unsigned long long *d_nheads = nullptr;
    void *d_temp = nullptr;
    size_t temp_bytes = 0;
    int rc = FMX_ERR_CUDA;
    int rounds = 0;
    const unsigned threads = 256;
    const unsigned grid = (unsigned)((n + threads - 1) / threads);
    const uint32_t k = 64 / bits;
    int nbits_n = 1;
    while ((1ull << nbits_n) <= n + 1 && nbits_n < 32) nbits_n++;

    SA_TRY(cudaSetDevice(device));
    SA_TRY(cudaMalloc(&d_key[0], n * 8));
    SA_TRY(cudaMalloc(&d_key[1], n * 8));
    SA_TRY(cudaMalloc(&d_sa[0], n * 4));
    SA_TRY(cudaMalloc(&d_sa[1], n * 4));
    SA_TRY(cudaMalloc(&d_rank, n * 4));
    SA_TRY(cudaMalloc(&d_grp, n * 4));
    SA_TRY(cudaMalloc(&d_nheads, 8));
    {
        cub::DoubleBuffer<uint64_t> keys(d_key[0], d_key[1]);
        cub::DoubleBuffer<uint32_t> vals(d_sa[0], d_sa[1]);
        SA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys, vals, (long long)n, 0, 64));
        SA_TRY(cudaMalloc(&d_temp, temp_bytes));

        k_sa_init<<<grid, threads>>>(d_text, n, bits, k, keys.Current(), vals.Current());
        SA_TRY(cudaGetLastError());
        size_t tb = temp_bytes;
        SA_TRY(cub::DeviceRadixSort::SortPairs(d_temp, tb, keys, vals, (long long)n, 0, (int)(k * bits)));
        uint64_t h = k;
        for (;;) {
            rounds++;
            unsigned long long nheads = 0;
            SA_TRY(cudaMemset(d_nheads, 0, 8));
            k_sa_heads<<<grid, threads>>>(keys.Current(), n, d_grp, d_nheads);
            SA_TRY(cudaGetLastError());
            SA_TRY(cudaMemcpy(&nheads, d_nheads, 8, cudaMemcpyDeviceToHost));
            if (nheads == n) break;  // every suffix is alone in its group: sorted
            if (h >= n) {
                err = "prefix doubling did not converge";
                rc = FMX_ERR_CUDA;
                goto fail;
            }
            // group id of every suffix = index of its group's first member + 1: a running maximum over the head marks
            SA_TRY((scan3::run<uint32_t, scan3::Max32, false>(d_grp, n, d_grp, scan3::Max32(), 0u, 0)));
            k_sa_scatter_rank<<<grid, threads>>>(vals.Current(), d_grp, n, d_rank);
            SA_TRY(cudaGetLastError());
            k_sa_make_keys<<<grid, threads>>>(vals.Current(), d_grp, d_rank, n, h, keys.Current());
            SA_TRY(cudaGetLastError());
            tb = temp_bytes;
            SA_TRY(cub::DeviceRadixSort::SortPairs(d_temp, tb, keys, vals, (long long)n, 0, 32 + nbits_n));
            h *= 2;
        }
        SA_TRY(cudaDeviceSynchronize());
        // hand the buffer holding the result to the caller
        const int cur = vals.Current() == d_sa[0] ? 0 : 1;
        *d_sa_out = d_sa[cur];
        d_sa[cur] = nullptr;
    }
    rc = 0;
fail:
    cudaFree(d_key[0]);
    cudaFree(d_key[1]);
    cudaFree(d_sa[0]);
    cudaFree(d_sa[1]);
    cudaFree(d_rank);
    cudaFree(d_grp);
    cudaFree(d_nheads);
    cudaFree(d_temp);
    if (rc) cudaGetLastError();
    if (rounds_out) *rounds_out = rounds;
    return rc;
}

// text: host pointer, n symbols each < 2^bits.  sa_out: host, n entries.
int gpu_suffix_array(const uint8_t *text, uint64_t n, uint32_t bits, int device, uint32_t *sa_out, int *rounds_out,
                     std::string &err) {
    if (n == 0) return 0;
    uint8_t *d_text = nullptr;
    uint32_t *d_sa = nullptr;
    int rc = FMX_ERR_CUDA;
    SA_TRY(cudaSetDevice(device));
    SA_TRY(cudaMalloc(&d_text, n));
    SA_TRY(cudaMemcpy(d_text, text, n, cudaMemcpyHostToDevice));
    rc = gpu_suffix_array_device(d_text, n, bits, device, &d_sa, rounds_out, err);
    if (rc == 0 && cudaMemcpy(sa_out, d_sa, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
        err = "suffix array: device to host copy failed";
        rc = FMX_ERR_CUDA;
    }
    cudaFree(d_text);
    cudaFree(d_sa);
    if (rc) cudaGetLastError();
    return rc;
fail:
    cudaFree(d_text);
    cudaGetLastError();
    return rc;
}

}  // namespace fmx
