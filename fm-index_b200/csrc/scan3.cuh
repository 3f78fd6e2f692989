// scan3.cuh -- a small three-phase device scan (tile reduce -> scan of the tile sums -> tile apply), header-only and
// templated so that the construction code (gpu_sa.cu, gpu_build.cu) owns its scans instead of calling a library.
// Not on the query path (that one has its own fused-load scans in kernels.cuh); sized for 10^9 elements.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace fmx {
namespace scan3 {

constexpr int THREADS = 256;
constexpr int ITEMS = 8;
constexpr int TILE = THREADS * ITEMS;

template <class V, class Op>
__device__ __forceinline__ V block_excl(V v, Op op, V ident, V &total, V *s_warp) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    V inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        V t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc = op(t, inc);
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        V w = lane < (THREADS / 32) ? s_warp[lane] : ident;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            V t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (uint32_t)o) w = op(t, w);
        }
        if (lane < (THREADS / 32)) s_warp[lane] = w;
    }
    __syncthreads();
    const V warp_prefix = wid ? s_warp[wid - 1] : ident;
    total = s_warp[THREADS / 32 - 1];
    V excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = ident;
    __syncthreads();
    return op(warp_prefix, excl);
}

template <class V, class Op>
__global__ void __launch_bounds__(THREADS) k_reduce(const V *in, uint64_t n, V *tile_sum, Op op, V ident) {
    __shared__ V s_warp[THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * TILE + (uint64_t)threadIdx.x * ITEMS;
    V acc = ident;
#pragma unroll
    for (int k = 0; k < ITEMS; k++)
        if (base + k < n) acc = op(acc, in[base + k]);
    V total;
    block_excl(acc, op, ident, total, s_warp);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// EXCL: out[i] excludes in[i].  tile_prefix (nullable) = exclusive scan of the tile sums.  In-place is fine.
template <class V, class Op, bool EXCL>
__global__ void __launch_bounds__(THREADS) k_apply(const V *in, uint64_t n, const V *tile_prefix, V *out, Op op, V ident) {
    __shared__ V s_warp[THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * TILE + (uint64_t)threadIdx.x * ITEMS;
    V v[ITEMS];
    V acc = ident;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        v[k] = base + k < n ? in[base + k] : ident;
        acc = op(acc, v[k]);
    }
    V total;
    V carry = op(tile_prefix ? tile_prefix[blockIdx.x] : ident, block_excl(acc, op, ident, total, s_warp));
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const V inc = op(carry, v[k]);
        if (base + k < n) out[base + k] = EXCL ? carry : inc;
        carry = inc;
    }
}

// out = scan(in) over n elements on stream `st`; scratch for the tile sums is allocated and released here
// (construction path).  Returns cudaSuccess or the first error.
template <class V, class Op, bool EXCL>
cudaError_t run(const V *in, uint64_t n, V *out, Op op, V ident, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    struct Level {
        V *sum;
        uint64_t n;
    };
    std::vector<Level> lv;
    uint64_t need = 0;
    for (uint64_t m = (n + TILE - 1) / TILE; ; m = (m + TILE - 1) / TILE) {
        lv.push_back({nullptr, m});
        need += m;
        if (m == 1) break;
    }
    V *scratch = nullptr;
    cudaError_t e = cudaMalloc(&scratch, need * sizeof(V));
    if (e != cudaSuccess) return e;
    V *cur = scratch;
    for (auto &l : lv) {
        l.sum = cur;
        cur += l.n;
    }
    // reduce up
    k_reduce<V, Op><<<(unsigned)lv[0].n, THREADS, 0, st>>>(in, n, lv[0].sum, op, ident);
    for (size_t k = 1; k < lv.size(); k++) k_reduce<V, Op><<<(unsigned)lv[k].n, THREADS, 0, st>>>(lv[k - 1].sum, lv[k - 1].n, lv[k].sum, op, ident);
    // scan down: the top level is one tile (exclusive, in place), every lower level applies with its parent's prefix
    for (size_t k = lv.size(); k-- > 0;) {
        const V *prefix = (k + 1 < lv.size()) ? lv[k + 1].sum : nullptr;
        k_apply<V, Op, true><<<(unsigned)((lv[k].n + TILE - 1) / TILE), THREADS, 0, st>>>(lv[k].sum, lv[k].n, prefix, lv[k].sum, op, ident);
    }
    k_apply<V, Op, EXCL><<<(unsigned)lv[0].n, THREADS, 0, st>>>(in, n, lv[0].sum, out, op, ident);
    e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(scratch);
    return e != cudaSuccess ? e : e2;
}

struct Sum32 {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a + b; }
};
struct Max32 {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

}  // namespace scan3
}  // namespace fmx
