// stream_patterns.cu -- microbenchmarks behind the PageRank kernel design:
// how fast can a B200 (a) stream 4-byte column ids in the access shapes the kernels use and
// (b) gather 8-byte values at random from an L2-resident vector.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_patterns stream_patterns.cu
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t ldnc(const uint32_t *p) { uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t mix(uint32_t h) { h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h; }

// A: grid-stride, 16 bytes per lane per load
__global__ void kA(const uint4 *col, uint64_t n16, unsigned long long *out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, st = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long s = 0;
    for (; i < n16; i += st) { uint4 v = col[i]; s += v.x + v.y + v.z + v.w; }
    if (s == 0x1234567) *out = s;
}
// B: one warp per 512-entry piece, 4-byte loads, unroll 8 (long-row path)
template <bool GATHER, bool NC>
__global__ void kB(const uint32_t *col, uint64_t m, const double *w, uint32_t wmask, double *out)
{
    const uint64_t nwarp = ((uint64_t)gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
    double s = 0;
    for (uint64_t u = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u * 512 < m; u += nwarp) {
        const uint64_t b = u * 512, e_end = b + 512 < m ? b + 512 : m;
#pragma unroll 8
        for (uint64_t e = b + lane; e < e_end; e += 32) {
            uint32_t c = NC ? ldnc(col + e) : col[e];
            s += GATHER ? w[c & wmask] : (double)c;
        }
    }
    if (s == 1.2345) *out = s;
}
// C: 8-lane group per row (short-row path): rowptr pair, then the row's entries
template <bool GATHER>
__global__ void kC(const uint64_t *rowptr, const uint32_t *col, uint64_t nrows, const double *w, uint32_t wmask, double *out)
{
    const unsigned sub = threadIdx.x & 7;
    uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const uint64_t ng = ((uint64_t)gridDim.x * blockDim.x) >> 3;
    double s = 0;
    for (; g < nrows; g += ng) {
        const uint64_t b = rowptr[g], e_end = rowptr[g + 1];
#pragma unroll 4
        for (uint64_t e = b + sub; e < e_end; e += 8) { uint32_t c = ldnc(col + e); s += GATHER ? w[c & wmask] : (double)c; }
    }
    if (s == 1.2345) *out = s;
}
// E: pure random gathers, indices from a hash (no index stream)
__global__ void kE(const double *w, uint32_t wmask, uint64_t per_thread, double *out)
{
    uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    double s = 0;
#pragma unroll 8
    for (uint64_t i = 0; i < per_thread; i++) { h = mix(h + (uint32_t)i); s += w[h & wmask]; }
    if (s == 1.2345) *out = s;
}

template <class F> float timeit(F f, int reps = 5)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main()
{
    const uint64_t m = 64ull << 20;          // entries (256 MB of column ids, > L2)
    const uint32_t wn = 1u << 21;            // 2 M doubles = 16 MB gather target (L2 resident)
    uint32_t *col; double *w, *outd; unsigned long long *outu; uint64_t *rowptr;
    CK(cudaMalloc(&col, m * 4)); CK(cudaMalloc(&w, (size_t)wn * 8)); CK(cudaMalloc(&outd, 8)); CK(cudaMalloc(&outu, 8));
    std::vector<uint32_t> hc(m); uint32_t x = 12345;
    for (uint64_t i = 0; i < m; i++) { x = x * 1664525u + 1013904223u; hc[i] = x >> 8; }
    CK(cudaMemcpy(col, hc.data(), m * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(w, 0, (size_t)wn * 8));
    // rows with a skewed short length distribution, mean ~12
    std::vector<uint64_t> rp; rp.push_back(0);
    while (rp.back() < m) { x = x * 1664525u + 1013904223u; uint32_t r = x >> 24; uint64_t len = r < 64 ? 0 : r < 160 ? 1 + (r & 7) : r < 240 ? 8 + (r & 15) : 32 + (r & 127); rp.push_back(rp.back() + len > m ? m : rp.back() + len); }
    const uint64_t nrows = rp.size() - 1;
    CK(cudaMalloc(&rowptr, rp.size() * 8)); CK(cudaMemcpy(rowptr, rp.data(), rp.size() * 8, cudaMemcpyHostToDevice));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto rep = [&](const char *name, float ms, double bytes, double entries) {
        printf("%-46s %8.1f us  %7.1f GB/s  %6.1f G entries/s\n", name, ms * 1e3, bytes / ms / 1e6, entries / ms / 1e6);
    };
    rep("A grid-stride uint4 stream", timeit([&] { kA<<<sms * 8, 256>>>((const uint4 *)col, m / 4, outu); }), m * 4.0, (double)m);
    rep("B warp/512 piece, u32 ld.nc, no gather", timeit([&] { kB<false, true><<<sms * 8, 256>>>(col, m, w, wn - 1, outd); }), m * 4.0, (double)m);
    rep("B warp/512 piece, plain ld, no gather", timeit([&] { kB<false, false><<<sms * 8, 256>>>(col, m, w, wn - 1, outd); }), m * 4.0, (double)m);
    rep("B warp/512 piece + gather (16 MB vector)", timeit([&] { kB<true, true><<<sms * 8, 256>>>(col, m, w, wn - 1, outd); }), m * 4.0, (double)m);
    rep("B same, 2 x 1024-thread CTAs per SM", timeit([&] { kB<true, true><<<sms * 2, 1024>>>(col, m, w, wn - 1, outd); }), m * 4.0, (double)m);
    rep("C 8-lane group per short row, no gather", timeit([&] { kC<false><<<sms * 8, 256>>>(rowptr, col, nrows, w, wn - 1, outd); }), m * 4.0, (double)m);
    rep("C 8-lane group per short row + gather", timeit([&] { kC<true><<<sms * 8, 256>>>(rowptr, col, nrows, w, wn - 1, outd); }), m * 4.0, (double)m);
    const uint64_t per_thread = 256; const uint64_t threads = (uint64_t)sms * 8 * 256;
    rep("E pure random 8 B gathers (16 MB vector)", timeit([&] { kE<<<sms * 8, 256>>>(w, wn - 1, per_thread, outd); }), 0, (double)threads * per_thread);
    rep("E pure random 8 B gathers (256 KB vector)", timeit([&] { kE<<<sms * 8, 256>>>(w, (1u << 15) - 1, per_thread, outd); }), 0, (double)threads * per_thread);
    printf("rows %llu, mean length %.1f\n", (unsigned long long)nrows, (double)m / nrows);
    return 0;
}
