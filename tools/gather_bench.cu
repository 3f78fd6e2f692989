// gather_bench.cu -- what bounds dependent-free random gathers on a B200?  (roofline denominator study)
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/gather_bench tools/gather_bench.cu
// run  : tools/gather_bench > gpurun_out/gather_bench.jsonl
//
// Every experiment issues independent random loads over a buffer and reports loads/s and 32-byte
// sectors/s.  Questions:
//   occupancy : does the rate depend on loads in flight per SM (Little's law) or is it a fixed service rate?
//   coop      : 32 lanes x own 32 B sector  vs  groups of G lanes reading ONE 32*G-byte block in one
//               instruction (G = 2, 4): is the limit per sector, or per distinct line/request?
//   pages     : same touched bytes spread over few or many 2 MiB pages (TLB reach 256 MB?)
//   bulk      : cp.async.bulk (TMA path, bypasses L1TEX miss tracking) of 32/64/128 B per request
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                             \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) {                                                          \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                      \
        }                                                                                 \
    } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z *= 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t ld32B(const void *p) {
    uint32_t w[8];
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
    return w[0] ^ w[7];
}

// G lanes share one block of 32*G bytes (G = 1: every lane its own sector).  ILP independent loads per
// thread per loop trip.  page_bytes/touch_bytes: addresses = page * page_bytes + (offset < touch_bytes).
template <int G, int ILP>
__global__ void __launch_bounds__(256) k_gather(const char *buf, uint64_t npages, uint64_t page_bytes, uint64_t touch_bytes,
                                                uint64_t nloads_per_thread, uint64_t seed, uint32_t *sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t grp = tid / G;
    const uint32_t sub = (uint32_t)(tid % G);
    const uint64_t blocks_per_page = touch_bytes / (32 * G);
    uint32_t acc = 0;
    for (uint64_t k = 0; k < nloads_per_thread; k += ILP) {
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            uint64_t z = mix(seed + grp * 0x100000001B3ull + (k + j));
            uint64_t page = __umul64hi(z, npages);
            uint64_t blk = __umul64hi(z * 0x9E3779B97F4A7C15ull, blocks_per_page);
            acc += ld32B(buf + page * page_bytes + blk * (32 * G) + sub * 32);
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// every lane of a warp picks its own random line, but all 32 inside ONE random 2 MiB page per instruction
template <int ILP>
__global__ void __launch_bounds__(256) k_gather_samepage(const char *buf, uint64_t npages, uint64_t page_bytes,
                                                         uint64_t nloads_per_thread, uint64_t seed, uint32_t *sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t wid = tid >> 5;
    uint32_t acc = 0;
    for (uint64_t k = 0; k < nloads_per_thread; k += ILP) {
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            uint64_t zp = mix(seed + wid * 0x100000001B3ull + (k + j));
            uint64_t zl = mix(seed * 3 + tid * 0x100000001B3ull + (k + j));
            uint64_t page = __umul64hi(zp, npages);
            uint64_t blk = __umul64hi(zl, page_bytes / 32);
            acc += ld32B(buf + page * page_bytes + blk * 32);
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// cp.async.bulk global -> shared, `bytes` per request, one elected lane per warp issues NREQ requests per
// phase and the warp waits on the CTA-local mbarrier.
template <int NREQ>
__global__ void __launch_bounds__(256) k_bulk(const char *buf, uint64_t nunits, uint32_t bytes, uint64_t phases, uint64_t seed,
                                              uint32_t *sink) {
    extern __shared__ __align__(128) char smem[];
    __shared__ __align__(8) uint64_t bar[8];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[warp]);
    char *dst = smem + (size_t)warp * NREQ * 128;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    }
    __syncwarp();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint32_t acc = 0, parity = 0;
    const uint64_t wid = (uint64_t)blockIdx.x * 8 + warp;
    for (uint64_t ph = 0; ph < phases; ph++) {
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes * NREQ) : "memory");
#pragma unroll 4
            for (int j = 0; j < NREQ; j++) {
                uint64_t z = mix(seed + wid * 0x100000001B3ull + ph * NREQ + j);
                const char *src = buf + __umul64hi(z, nunits) * bytes;
                uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + j * 128);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                             "l"(src), "r"(bytes), "r"(bar_a)
                             : "memory");
            }
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(bar_a), "r"(parity)
                : "memory");
        }
        parity ^= 1;
        acc += *reinterpret_cast<volatile uint32_t *>(dst + (lane % NREQ) * 128);
    }
    if (acc == 0x12345678u) *sink = acc;
}

static int g_sms = 148;

template <int G, int ILP>
static void run_gather(const char *name, const char *buf, uint64_t npages, uint64_t page_bytes, uint64_t touch_bytes, int blocks_per_sm,
                       uint32_t *sink) {
    const uint64_t per_thread = 2048;  // loads per thread
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    unsigned grid = (unsigned)(g_sms * blocks_per_sm);
    double best = 0;
    for (int it = 0; it < 4; it++) {
        CK(cudaEventRecord(e0));
        k_gather<G, ILP><<<grid, 256>>>(buf, npages, page_bytes, touch_bytes, per_thread, 0x1234567ull * (it + 1), sink);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double loads = (double)grid * 256 * per_thread;  // lane-level 32 B loads = sectors
        double rate = loads / (ms * 1e-3);
        if (it > 0 && rate > best) best = rate;
    }
    printf("{\"exp\": \"%s\", \"lanes_per_block\": %d, \"ilp\": %d, \"blocks_per_sm\": %d, \"pages\": %llu, \"page_bytes\": %llu, "
           "\"touch_bytes\": %llu, \"footprint_MiB\": %.1f, \"gsectors_per_s\": %.2f, \"gblocks_per_s\": %.2f}\n",
           name, G, ILP, blocks_per_sm, (unsigned long long)npages, (unsigned long long)page_bytes, (unsigned long long)touch_bytes,
           (double)npages * touch_bytes / 1048576.0, best / 1e9, best / G / 1e9);
    fflush(stdout);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

static void run_samepage(const char *buf, uint64_t npages, uint64_t page_bytes, uint32_t *sink) {
    const uint64_t per_thread = 2048;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    unsigned grid = (unsigned)(g_sms * 8);
    double best = 0;
    for (int it = 0; it < 4; it++) {
        CK(cudaEventRecord(e0));
        k_gather_samepage<4><<<grid, 256>>>(buf, npages, page_bytes, per_thread, 0x1234567ull * (it + 1), sink);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double rate = (double)grid * 256 * per_thread / (ms * 1e-3);
        if (it > 0 && rate > best) best = rate;
    }
    printf("{\"exp\": \"samepage_per_warp\", \"pages\": %llu, \"footprint_MiB\": %.1f, \"gsectors_per_s\": %.2f, "
           "\"gwarp_instr_per_s\": %.2f}\n",
           (unsigned long long)npages, (double)npages * page_bytes / 1048576.0, best / 1e9, best / 32 / 1e9);
    fflush(stdout);
}

#define CKD(x)                                                                    \
    do {                                                                          \
        CUresult r_ = (x);                                                        \
        if (r_ != CUDA_SUCCESS) {                                                 \
            const char *m_ = nullptr;                                             \
            cuGetErrorString(r_, &m_);                                            \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, m_ ? m_ : "?"); \
            return nullptr;                                                       \
        }                                                                         \
    } while (0)

// a buffer mapped through the virtual-memory-management API with a chosen physical chunk size and
// VA alignment: does the driver then map it with pages larger than 2 MiB?
static char *vmm_alloc(uint64_t total, uint64_t chunk, uint64_t align) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    size_t gmin = 0, grec = 0;
    CKD(cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    CKD(cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    printf("{\"exp\": \"vmm_granularity\", \"minimum\": %zu, \"recommended\": %zu}\n", gmin, grec);
    CUdeviceptr va = 0;
    CKD(cuMemAddressReserve(&va, total, align, 0, 0));
    for (uint64_t off = 0; off < total; off += chunk) {
        CUmemGenericAllocationHandle h;
        CKD(cuMemCreate(&h, chunk, &prop, 0));
        CKD(cuMemMap(va + off, chunk, 0, h, 0));
        CKD(cuMemRelease(h));
    }
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CKD(cuMemSetAccess(va, total, &acc, 1));
    return reinterpret_cast<char *>(va);
}

template <int NREQ>
static void run_bulk(const char *buf, uint64_t bytes_total, uint32_t bytes, int blocks_per_sm, uint32_t *sink) {
    const uint64_t phases = 512;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    unsigned grid = (unsigned)(g_sms * blocks_per_sm);
    size_t smem = (size_t)8 * NREQ * 128;
    CK(cudaFuncSetAttribute(k_bulk<NREQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double best = 0;
    for (int it = 0; it < 4; it++) {
        CK(cudaEventRecord(e0));
        k_bulk<NREQ><<<grid, 256, smem>>>(buf, bytes_total / bytes, bytes, phases, 0x777ull * (it + 1), sink);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double reqs = (double)grid * 8 * phases * NREQ;
        double rate = reqs / (ms * 1e-3);
        if (it > 0 && rate > best) best = rate;
    }
    printf("{\"exp\": \"bulk\", \"bytes_per_request\": %u, \"requests_in_flight_per_warp\": %d, \"blocks_per_sm\": %d, "
           "\"footprint_MiB\": %.1f, \"grequests_per_s\": %.2f, \"gsectors_per_s\": %.2f}\n",
           bytes, NREQ, blocks_per_sm, bytes_total / 1048576.0, best / 1e9, best * (bytes / 32) / 1e9);
    fflush(stdout);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
    const uint64_t GiB = 1ull << 30, MiB = 1ull << 20;
    const uint64_t total = 16 * GiB;
    char *buf;
    uint32_t *sink;
    CK(cudaMalloc(&buf, total));
    CK(cudaMalloc(&sink, 256));
    CK(cudaMemset(buf, 1, total));
    const uint64_t PG = 2 * MiB;

    // 1. occupancy / ILP at a 4 GiB footprint (everything misses L2)
    {
        uint64_t np = 4 * GiB / PG;
        run_gather<1, 1>("occupancy", buf, np, PG, PG, 1, sink);
        run_gather<1, 1>("occupancy", buf, np, PG, PG, 2, sink);
        run_gather<1, 1>("occupancy", buf, np, PG, PG, 4, sink);
        run_gather<1, 1>("occupancy", buf, np, PG, PG, 8, sink);
        run_gather<1, 2>("occupancy", buf, np, PG, PG, 8, sink);
        run_gather<1, 4>("occupancy", buf, np, PG, PG, 8, sink);
        run_gather<1, 8>("occupancy", buf, np, PG, PG, 8, sink);
        run_gather<1, 4>("occupancy", buf, np, PG, PG, 2, sink);
        run_gather<1, 8>("occupancy", buf, np, PG, PG, 1, sink);
    }
    // 2. cooperative blocks: G lanes read one 32*G-byte block in one instruction
    for (uint64_t fp : {64 * MiB, 512 * MiB, 2 * GiB, 8 * GiB}) {
        uint64_t np = fp / PG;
        run_gather<1, 4>("coop", buf, np, PG, PG, 8, sink);
        run_gather<2, 4>("coop", buf, np, PG, PG, 8, sink);
        run_gather<4, 4>("coop", buf, np, PG, PG, 8, sink);
        run_gather<8, 4>("coop", buf, np, PG, PG, 8, sink);
    }
    // 3. pages: 64 MiB of touched bytes (fits L2) over 32 .. 8192 pages; then 1 GiB touched over 512 .. 8192 pages
    for (uint64_t np : {32ull, 128ull, 512ull, 2048ull, 8192ull}) run_gather<1, 4>("pages64M", buf, np, PG, 64 * MiB / np, 8, sink);
    for (uint64_t np : {512ull, 2048ull, 8192ull}) run_gather<1, 4>("pages1G", buf, np, PG, GiB / np, 8, sink);
    // 3b. L2-miss-bound rate with every page in the TLB: 200 MiB over 100 pages
    run_gather<1, 4>("tlbfit200M", buf, 100, PG, PG, 8, sink);
    run_gather<4, 4>("tlbfit200M", buf, 100, PG, PG, 8, sink);
    run_gather<1, 4>("tlbfit240M", buf, 120, PG, PG, 8, sink);
    // 3c. one translation per warp instruction: all 32 lanes inside one random page
    run_samepage(buf, 2048, PG, sink);
    run_samepage(buf, 8192, PG, sink);
    // 3d. virtual-memory-management allocations: one 4 GiB physical chunk / 512 MiB chunks, 512 MiB-aligned VA
    for (uint64_t chunk : {4 * GiB, 512 * MiB, 32 * MiB}) {
        char *vb = vmm_alloc(4 * GiB, chunk, 512 * MiB);
        if (vb) {
            CK(cudaMemset(vb, 1, 4 * GiB));
            char nm[64];
            snprintf(nm, sizeof nm, "vmm_chunk%lluM", (unsigned long long)(chunk / MiB));
            run_gather<1, 4>(nm, vb, 4 * GiB / PG, PG, PG, 8, sink);
        }
    }
    // 4. TMA bulk copies
    for (uint32_t b : {32u, 64u, 128u}) {
        run_bulk<8>(buf, 4 * GiB, b, 4, sink);
        run_bulk<16>(buf, 4 * GiB, b, 8, sink);
        run_bulk<32>(buf, 4 * GiB, b, 8, sink);
    }
    run_bulk<32>(buf, 64 * MiB, 32, 8, sink);
    return 0;
}
