// fetch_gran.cu -- how many DRAM bytes does ONE random 32-byte read cost on a B200, per load flavour?
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/fetch_gran tools/fetch_gran.cu
// run  : ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,gpu__time_duration.sum \
//            --clock-control none --csv --log-file gpurun_out/fetch_gran.csv tools/fetch_gran [limit]
// (the optional argument sets cudaLimitMaxL2FetchGranularity before the first allocation)
// Each kernel issues N independent random 32-byte reads over a 4 GiB buffer (everything misses L2).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>

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

enum { M_NC_V8, M_NC_V4X2, M_CA_V4X2, M_CG_V4X2, M_CS_V4X2, M_LU_V4X2, M_CV_V4X2, M_NC_NOALLOC, M_NC_EVICT_FIRST, M_NC_V4_ONE,
       M_NC_U64_ONE, M_LDGSTS16, M_NC_V8_L2_64, M_NC_V8_L2_256, M_COUNT };

template <int MODE>
__device__ __forceinline__ uint32_t load32(const char *p, uint64_t pol, char *sm) {
    uint32_t a = 0, b = 0, c = 0, d = 0, e = 0, f = 0, g = 0, h = 0;
    if (MODE == M_NC_V8)
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    else if (MODE == M_NC_V8_L2_64)
        asm volatile("ld.global.nc.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    else if (MODE == M_NC_V8_L2_256)
        asm volatile("ld.global.nc.L2::256B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    else if (MODE == M_NC_V4X2) {
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 16));
    } else if (MODE == M_CA_V4X2) {
        asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 16));
    } else if (MODE == M_CG_V4X2) {
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 16));
    } else if (MODE == M_CS_V4X2) {
        asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 16));
    } else if (MODE == M_LU_V4X2) {
        asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 16));
    } else if (MODE == M_CV_V4X2) {
        asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 16));
    } else if (MODE == M_NC_NOALLOC) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 16));
    } else if (MODE == M_NC_EVICT_FIRST) {
        asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p), "l"(pol));
        asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p + 16), "l"(pol));
    } else if (MODE == M_NC_V4_ONE) {
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
    } else if (MODE == M_NC_U64_ONE) {
        asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
    } else if (MODE == M_LDGSTS16) {
        uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(p) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s + 16), "l"(p + 16) : "memory");
    }
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_fetch(const char *buf, uint64_t nblocks, uint64_t per_thread, uint32_t *sink) {
    __shared__ __align__(32) char sm[256 * 32];
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t pol = 0;
    if (MODE == M_NC_EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    uint32_t acc = 0;
    for (uint64_t k = 0; k < per_thread; k += 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint64_t z = mix(tid * 0x100000001B3ull + (k + j) + MODE * 77);
            acc += load32<MODE>(buf + __umul64hi(z, nblocks) * 32, pol, sm + threadIdx.x * 32);
        }
        if (MODE == M_LDGSTS16) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            acc += *reinterpret_cast<volatile uint32_t *>(sm + threadIdx.x * 32);
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int MODE>
static void run(const char *name, const char *buf, uint64_t bytes, uint32_t *sink) {
    const uint64_t per_thread = 512;
    unsigned grid = 148 * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k_fetch<MODE><<<grid, 256>>>(buf, bytes / 32, per_thread, sink);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double loads = (double)grid * 256 * per_thread;
    printf("{\"mode\": %d, \"name\": \"%s\", \"loads\": %.0f, \"ms\": %.3f, \"gloads_per_s\": %.2f}\n", MODE, name, loads, ms, loads / ms / 1e6);
    fflush(stdout);
}

// ---------------------------------------------------------------------------------------------
// design study for the byte-alphabet rank structure: one "LF step" per iteration, every step's
// address depending on the data the previous step loaded (as in backward search)
//   wm4    : 4 dependent random 32 B sectors, one per level array (quaternary wavelet matrix)
//   occ    : ONE random 1152-byte block [128 B of BWT | 256 x u32 counters]: the 4 B counter of a random
//            symbol + NSEC sectors of the BWT line, all independent of each other
//   occdep : the same, but the counter address depends on the loaded BWT byte (access then rank)
//   coop16 : 4 lanes per pattern, 2 dependent random 128 B blocks (16-ary wavelet, 2 levels)
__device__ __forceinline__ uint32_t ld32Bx(const char *p) {
    uint32_t w[8];
    asm volatile("ld.global.nc.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
    return w[0] ^ w[7];
}
__device__ __forceinline__ uint32_t ld4B(const char *p) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(256) k_wm4(const char *buf, uint64_t level_bytes, uint32_t steps, uint32_t *sink) {
    uint64_t st = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x100000001B3ull;
    const uint64_t nb = level_bytes / 32;
    uint32_t acc = 0;
    for (uint32_t k = 0; k < steps; k++) {
#pragma unroll
        for (int l = 0; l < 4; l++) {
            st = mix(st + acc + l);
            acc += ld32Bx(buf + (uint64_t)l * level_bytes + __umul64hi(st, nb) * 32);
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int NSEC, bool DEP>
__global__ void __launch_bounds__(256) k_occ(const char *buf, uint64_t nblocks, uint32_t steps, uint32_t *sink) {
    uint64_t st = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x100000001B3ull;
    uint32_t acc = 0;
    for (uint32_t k = 0; k < steps; k++) {
        st = mix(st + acc);
        const char *blk = buf + __umul64hi(st, nblocks) * 1152;
        uint32_t d = 0;
#pragma unroll
        for (int q = 0; q < NSEC; q++) d += ld32Bx(blk + 32 * q);
        uint32_t c = DEP ? ((d ^ (uint32_t)st) & 255u) : ((uint32_t)(st >> 40) & 255u);
        acc += d + ld4B(blk + 128 + 4 * c);
    }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void __launch_bounds__(256) k_coop16(const char *buf, uint64_t level_bytes, uint32_t steps, uint32_t *sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t st = (tid >> 2) * 0x100000001B3ull;
    const uint32_t sub = tid & 3;
    const uint64_t nb = level_bytes / 128;
    uint32_t acc = 0;
    for (uint32_t k = 0; k < steps; k++) {
#pragma unroll
        for (int l = 0; l < 2; l++) {
            st = mix(st + acc + l);
            uint32_t v = ld32Bx(buf + (uint64_t)l * level_bytes + __umul64hi(st, nb) * 128 + sub * 32);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            acc += v;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <class F>
static void time_it(const char *name, double steps_total, F &&launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double best = 1e30;
    for (int it = 0; it < 3; it++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    printf("{\"study\": \"%s\", \"ms\": %.3f, \"gsteps_per_s\": %.2f}\n", name, best, steps_total / best / 1e6);
    fflush(stdout);
}

static void design_study(uint32_t *sink) {
    const uint64_t GiB = 1ull << 30;
    char *buf;
    CK(cudaMalloc(&buf, 9 * GiB));
    CK(cudaMemset(buf, 1, 9 * GiB));
    const unsigned grid = 148 * 8;
    const uint32_t steps = 256;
    const double threads = (double)grid * 256;
    time_it("wm4: 4 dependent 32 B sectors / step (2 GiB)", threads * steps, [&] { k_wm4<<<grid, 256>>>(buf, GiB / 2, steps, sink); });
    time_it("occ: counter + 4 sectors of one 1152 B block / step (9 GiB)", threads * steps,
            [&] { k_occ<4, false><<<grid, 256>>>(buf, 9 * GiB / 1152, steps, sink); });
    time_it("occ: counter + 2 sectors / step (9 GiB)", threads * steps,
            [&] { k_occ<2, false><<<grid, 256>>>(buf, 9 * GiB / 1152, steps, sink); });
    time_it("occdep: 2 sectors then dependent counter / step (9 GiB)", threads * steps,
            [&] { k_occ<2, true><<<grid, 256>>>(buf, 9 * GiB / 1152, steps, sink); });
    time_it("coop16: 4 lanes, 2 dependent 128 B blocks / step (2 GiB)", threads / 4 * steps,
            [&] { k_coop16<<<grid, 256>>>(buf, GiB, steps, sink); });
    time_it("coop16 x4 grid", threads * steps, [&] { k_coop16<<<grid * 4, 256>>>(buf, GiB, steps, sink); });
    cudaFree(buf);
}

int main(int argc, char **argv) {
    CK(cudaSetDevice(0));
    if (argc > 1) {
        CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1])));
    }
    size_t lim = 0;
    CK(cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity));
    printf("{\"cudaLimitMaxL2FetchGranularity\": %zu}\n", lim);
    const uint64_t bytes = 4ull << 30;
    char *buf;
    uint32_t *sink;
    CK(cudaMalloc(&sink, 256));
    if (argc > 2) {
        design_study(sink);
        return 0;
    }
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    run<M_NC_V8>("ld.global.nc.v8.u32", buf, bytes, sink);
    run<M_NC_V4X2>("2 x ld.global.nc.v4.u32", buf, bytes, sink);
    run<M_CA_V4X2>("2 x ld.global.ca.v4.u32", buf, bytes, sink);
    run<M_CG_V4X2>("2 x ld.global.cg.v4.u32", buf, bytes, sink);
    run<M_CS_V4X2>("2 x ld.global.cs.v4.u32", buf, bytes, sink);
    run<M_LU_V4X2>("2 x ld.global.lu.v4.u32", buf, bytes, sink);
    run<M_CV_V4X2>("2 x ld.global.cv.v4.u32", buf, bytes, sink);
    run<M_NC_NOALLOC>("2 x ld.global.nc.L1::no_allocate.v4.u32", buf, bytes, sink);
    run<M_NC_EVICT_FIRST>("2 x ld.global.nc.L2::cache_hint(evict_first).v4.u32", buf, bytes, sink);
    run<M_NC_V4_ONE>("1 x ld.global.nc.v4.u32 (16 B)", buf, bytes, sink);
    run<M_NC_U64_ONE>("1 x ld.global.nc.v2.u32 (8 B)", buf, bytes, sink);
    run<M_LDGSTS16>("2 x cp.async.cg 16 B", buf, bytes, sink);
    run<M_NC_V8_L2_64>("ld.global.nc.L2::64B.v8.u32", buf, bytes, sink);
    run<M_NC_V8_L2_256>("ld.global.nc.L2::256B.v8.u32", buf, bytes, sink);
    return 0;
}
