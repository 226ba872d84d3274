// gather_paths.cu -- which SM datapaths can serve random 8-byte gathers from an L2-resident vector,
// and can they run concurrently?  
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t mix(uint32_t h) { h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h; }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MODE 0: plain LDG.64; 1: ld.global.nc.L1::no_allocate; 2: ld.global.cg; 3: ld.volatile (cv)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_ldg(const double *w, uint32_t wmask, int per_thread, double *out)
{
    uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    double s = 0;
    for (int i = 0; i < per_thread; i += 8) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            h = mix(h + (uint32_t)(i + j));
            const double *p = w + (h & wmask);
            if (MODE == 0) v[j] = *p;
            else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v[j]) : "l"(p));
            else if (MODE == 2) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v[j]) : "l"(p));
            else asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v[j]) : "l"(p));
        }
#pragma unroll
        for (int j = 0; j < 8; j++) s += v[j];
    }
    if (s == 1.2345) *out = s;
}

// texture path: tex1Dfetch<int2> on a linear texture over the same vector (TEX_PER of every 8 gathers per lane through
// the texture unit, the others LDG): does the texture pipe have a tag rate of its own?
template <int TEX_PER>
__global__ void __launch_bounds__(1024, 1) k_tex(cudaTextureObject_t tex, const double *w, uint32_t wmask, int per_thread, double *out)
{
    uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    double s = 0;
    for (int i = 0; i < per_thread; i += 8) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            h = mix(h + (uint32_t)(i + j));
            const uint32_t idx = h & wmask;
            if (j < TEX_PER) { const int2 t = tex1Dfetch<int2>(tex, (int)idx); v[j] = __hiloint2double(t.y, t.x); }
            else v[j] = w[idx];
        }
#pragma unroll
        for (int j = 0; j < 8; j++) s += v[j];
    }
    if (s == 1.2345) *out = s;
}

// bulk-copy gathers: every lane issues BATCH 16-byte cp.async.bulk copies into its own staging slots, the warp waits
// on its mbarrier, lanes read their values back from shared memory.  LSU_PER: additional LDG gathers per batch (mixed)
template <int BATCH, int LSU_PER>
__global__ void __launch_bounds__(1024, 1) k_bulk(const double *w, uint32_t wmask, int batches, double *out)
{
    extern __shared__ __align__(16) double2 stage_raw[];
    double2 (*stage)[BATCH][32] = (double2 (*)[BATCH][32])stage_raw;   // [warp][j][lane]
    __shared__ __align__(8) uint64_t bar[32];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t bar_a = smem_u32(&bar[wib]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    double s = 0;
    uint32_t phase = 0;
    for (int b = 0; b < batches; b++) {
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(32u * BATCH * 16u) : "memory");
        __syncwarp();
        uint32_t odd[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; j++) {
            h = mix(h + (uint32_t)(b * BATCH + j));
            const uint32_t idx = h & wmask;
            odd[j] = idx & 1u;
            const double *src = w + (idx & ~1u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                         ::"r"(smem_u32(&stage[wib][j][lane])), "l"(src), "r"(bar_a) : "memory");
        }
        double v2[LSU_PER > 0 ? LSU_PER : 1];
#pragma unroll
        for (int j = 0; j < LSU_PER; j++) { h = mix(h + 77u + (uint32_t)j); v2[j] = w[h & wmask]; }
        // wait
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
        phase ^= 1;
#pragma unroll
        for (int j = 0; j < BATCH; j++) { const double2 d = stage[wib][j][lane]; s += odd[j] ? d.y : d.x; }
#pragma unroll
        for (int j = 0; j < LSU_PER; j++) s += v2[j];
        __syncwarp();
    }
    if (s == 1.2345) *out = s;
}

// DSMEM: the vector lives in the shared memories of the CLUSTER's CTAs (CL x SLOTS doubles); random 8-byte reads
template <int CL>
__global__ void __launch_bounds__(1024, 1) k_dsmem(const double *w, int per_thread, uint32_t slots_pow2, double *out)
{
    extern __shared__ double sw[];
    cg::cluster_group cluster = cg::this_cluster();
    for (uint32_t i = threadIdx.x; i < slots_pow2; i += blockDim.x) sw[i] = w[i];
    cluster.sync();
    uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    const uint32_t base = smem_u32(sw);
    double s = 0;
    for (int i = 0; i < per_thread; i += 8) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            h = mix(h + (uint32_t)(i + j));
            const uint32_t r = (h >> 20) % CL, off = h & (slots_pow2 - 1);
            uint32_t ra;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(base + off * 8u), "r"(r));
            asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v[j]) : "r"(ra));
        }
#pragma unroll
        for (int j = 0; j < 8; j++) s += v[j];
    }
    if (s == 1.2345) *out = s;
    cluster.sync();
}

template <class F> float timeit(F f, int reps = 5)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

template <int CL> int run_dsmem(const double *w, double *outd, int sms)
{
    const uint32_t slots = 16384; // 128 KB per CTA
    auto kern = k_dsmem<CL>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, slots * 8));
    if (CL > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    const int grid = (sms / CL) * CL;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = slots * 8;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const int per_thread = 256;
    cudaError_t err = cudaSuccess;
    float ms = timeit([&] { cudaError_t e = cudaLaunchKernelEx(&cfg, kern, w, per_thread, slots, outd); if (e != cudaSuccess) err = e; });
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("DSMEM cluster %d: launch failed (%s)\n", CL, cudaGetErrorString(err)); return 0; }
    const double g = (double)grid * 1024 * per_thread;
    printf("DSMEM random 8 B reads, cluster %d (%d CTAs, %u KB each)   %8.1f us  %6.1f G gathers/s (incl. %u KB fill per CTA)\n", CL, grid, slots * 8 / 1024, ms * 1e3, g / ms / 1e6, slots * 8 / 1024);
    return 0;
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("SMs %d, clock %d MHz\n", sms, khz / 1000);
    const uint32_t wn = 1u << 21;   // 16 MB
    double *w, *outd;
    CK(cudaMalloc(&w, (size_t)wn * 8)); CK(cudaMalloc(&outd, 8));
    CK(cudaMemset(w, 0, (size_t)wn * 8));
    const int per_thread = 256;
    const double g = (double)sms * 1024 * per_thread;
    auto rep = [&](const char *name, float ms, double gathers) { printf("%-64s %8.1f us  %6.1f G gathers/s\n", name, ms * 1e3, gathers / ms / 1e6); };
    rep("LDG.64 plain, 16 MB vector", timeit([&] { k_ldg<0><<<sms, 1024>>>(w, wn - 1, per_thread, outd); }), g);
    rep("ld.global.nc.L1::no_allocate", timeit([&] { k_ldg<1><<<sms, 1024>>>(w, wn - 1, per_thread, outd); }), g);
    rep("ld.global.cg", timeit([&] { k_ldg<2><<<sms, 1024>>>(w, wn - 1, per_thread, outd); }), g);
    rep("ld.volatile.global", timeit([&] { k_ldg<3><<<sms, 1024>>>(w, wn - 1, per_thread, outd); }), g);
#define BULK(B, L, DIV, NAME) { auto kf = k_bulk<B, L>; const int sm = 32 * B * 32 * 16; \
      CK(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)); \
      rep(NAME, timeit([&] { kf<<<sms, 1024, sm>>>(w, wn - 1, per_thread / DIV, outd); }), g); CK(cudaGetLastError()); }
    BULK(8, 0, 8, "cp.async.bulk 16 B gathers, batch 8 per lane");
    BULK(4, 0, 4, "cp.async.bulk 16 B gathers, batch 4 per lane");
    BULK(4, 4, 8, "mixed: bulk batch 4 + 4 LDG per lane");
    BULK(2, 6, 8, "mixed: bulk batch 2 + 6 LDG per lane");
    BULK(1, 7, 8, "mixed: bulk batch 1 + 7 LDG per lane");
    {
        cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = w;
        rd.res.linear.desc = cudaCreateChannelDesc<int2>(); rd.res.linear.sizeInBytes = (size_t)wn * 8;
        cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        rep("tex1Dfetch<int2>, all 8 gathers per lane", timeit([&] { k_tex<8><<<sms, 1024>>>(tex, w, wn - 1, per_thread, outd); }), g);
        rep("mixed: 4 tex1Dfetch + 4 LDG per lane", timeit([&] { k_tex<4><<<sms, 1024>>>(tex, w, wn - 1, per_thread, outd); }), g);
        rep("mixed: 2 tex1Dfetch + 6 LDG per lane", timeit([&] { k_tex<2><<<sms, 1024>>>(tex, w, wn - 1, per_thread, outd); }), g);
        CK(cudaDeviceSynchronize());
        CK(cudaDestroyTextureObject(tex));
    }
    CK(cudaDeviceSynchronize());
    run_dsmem<1>(w, outd, sms);
    run_dsmem<2>(w, outd, sms);
    run_dsmem<4>(w, outd, sms);
    run_dsmem<8>(w, outd, sms);
    CK(cudaDeviceSynchronize());
    return 0;
}
