// rmat.cu -- synthetic Graph500 RMAT inputs built entirely in HBM
// (SURVEY.md 8(d)): (A,B,C,D) = (.57,.19,.19,.05), counter-based splitmix64
// keyed by (seed, edge index, level pair), scrambled ids, self-loops and
// duplicates removed, isolated ids dropped, dense id = rank of the original id.
// Bit-for-bit the same edge stream as the host generators
// (ldbc_graphalytics_platforms_graphblas_b200/rmat.py and the test-side C generator).
#include <cub/device/device_scan.cuh>

#include <cstdlib>
#include <vector>

#include "graph.cuh"

namespace gx {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint64_t scramble(uint64_t v, int scale, uint64_t k0, uint64_t k1)
{
    const uint64_t mask = (1ull << scale) - 1;
    const int sh = scale / 2 + 1;
    uint64_t x = v & mask;
    x = (x * 0x9E3779B97F4A7C15ull + k0) & mask;
    x ^= x >> sh;
    x = (x * 0xD1B54A32D192ED03ull + k1) & mask;
    x ^= x >> sh;
    return x;
}

__global__ void k_rmat_gen(int scale, uint64_t seed, uint64_t k0, uint64_t k1, uint64_t count, int directed,
                           uint64_t *__restrict__ keys)
{
    const uint32_t tA = (uint32_t)(0.57 * 4294967296.0);
    const uint32_t tAB = (uint32_t)(0.76 * 4294967296.0);
    const uint32_t tABC = (uint32_t)(0.95 * 4294967296.0);
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) {
        uint64_t s = 0, d = 0, h = 0;
        for (int l = 0; l < scale; l++) {
            if ((l & 1) == 0) h = splitmix64(seed + (i << 5) + (uint64_t)(l >> 1));
            const uint32_t r = (l & 1) ? (uint32_t)(h >> 32) : (uint32_t)h;
            const uint32_t q = (r < tA) ? 0u : (r < tAB) ? 1u : (r < tABC) ? 2u : 3u;
            s = (s << 1) | (q >> 1);
            d = (d << 1) | (q & 1);
        }
        s = scramble(s, scale, k0, k1);
        d = scramble(d, scale, k0, k1);
        const uint64_t loop = ~0ull;
        if (directed) keys[i] = (s == d) ? loop : ((s << 32) | d);
        else {
            keys[2 * i] = (s == d) ? loop : ((s << 32) | d);
            keys[2 * i + 1] = (s == d) ? loop : ((d << 32) | s);
        }
    }
}

__global__ void k_unique_heads(const uint64_t *__restrict__ keys, uint64_t cnt, uint8_t *__restrict__ head)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < cnt; i += stride) {
        const uint64_t k = keys[i];
        head[i] = (k != ~0ull && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
    }
}

__global__ void k_mark_present(const uint64_t *__restrict__ keys, uint64_t cnt, uint32_t *__restrict__ pres)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < cnt; i += stride) {
        const uint64_t k = keys[i];
        pres[k >> 32] = 1;
        pres[k & 0xFFFFFFFFull] = 1;
    }
}

__global__ void k_fill_mapping(const uint32_t *__restrict__ pres, const uint32_t *__restrict__ newid, uint64_t ids,
                               uint32_t *__restrict__ mapping)
{
    uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; x < ids; x += stride)
        if (pres[x]) mapping[newid[x]] = (uint32_t)x;
}

__device__ __forceinline__ double edge_weight(uint64_t a, uint64_t b, uint64_t kw)
{
    const uint64_t lo = a < b ? a : b, hi = a < b ? b : a;
    const uint64_t h = splitmix64(kw ^ (lo * 0x100000001B3ull + hi));
    return (double)((h >> 11) + 1) * (1.0 / 9007199254740992.0);
}

// relabel in place to dense ids (monotone, so the order survives) and emit weights
__global__ void k_relabel(uint64_t *__restrict__ keys, uint64_t cnt, const uint32_t *__restrict__ newid, uint64_t kw,
                          double *__restrict__ w)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < cnt; i += stride) {
        const uint64_t k = keys[i];
        const uint64_t s = k >> 32, d = k & 0xFFFFFFFFull;
        if (w) w[i] = edge_weight(s, d, kw);
        keys[i] = ((uint64_t)newid[s] << 32) | newid[d];
    }
}

__global__ void k_keys_low32(const uint64_t *__restrict__ keys, uint64_t m, uint32_t *__restrict__ out)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < m; e += stride) out[e] = (uint32_t)keys[e];
}

} // namespace gx

using namespace gx;

extern "C" int gx_rmat_create(gx_graph **out, int scale, int edgefactor, uint64_t seed, int directed, int weighted,
                              uint64_t **mapping_out)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(out != nullptr, "graph handle is NULL");
        GX_REQUIRE(scale >= 1 && scale <= 31, "scale must be in 1..31");
        GX_REQUIRE(edgefactor >= 1, "edgefactor must be positive");
        Context &c = ctx();
        c.timing = gx_timing{};
        PhaseTimer tb(&c.timing.build_ms);
        const uint64_t ids = 1ull << scale;
        const uint64_t gen = (uint64_t)edgefactor << scale;
        const uint64_t cnt = directed ? gen : 2 * gen;
        const uint64_t k0 = splitmix64(seed ^ 0xA5A5A5A5ull), k1 = splitmix64(seed ^ 0x5A5A5A5A5Aull);
        const uint64_t kw = splitmix64(seed ^ 0x57E1687ull);
        gx_graph *g = new gx_graph();
        try {
            DevBuf<uint64_t> ukeys;
            uint64_t m = 0;
            {
                DevBuf<uint64_t> keys(cnt);
                GX_LAUNCH(k_rmat_gen, grid_persistent(16), 256, 0, scale, seed, k0, k1, gen, directed, keys.p);
                sort_keys64(keys, cnt, 32 + scale);
                DevBuf<uint8_t> head(cnt);
                GX_LAUNCH(k_unique_heads, grid_persistent(8), 256, 0, keys.p, cnt, head.p);
                ukeys.alloc(cnt);
                m = select_flagged(keys.p, head.p, cnt, ukeys);
            }
            DevBuf<uint32_t> pres(ids), newid(ids);
            pres.zero();
            GX_LAUNCH(k_mark_present, grid_persistent(8), 256, 0, ukeys.p, m, pres.p);
            {
                size_t tbytes = 0;
                GX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tbytes, pres.p, newid.p, (int64_t)ids, c.stream));
                DevBuf<char> tmp(tbytes);
                GX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tbytes, pres.p, newid.p, (int64_t)ids, c.stream));
            }
            uint32_t tail[2];
            read_back(&tail[0], newid.p + (ids - 1), sizeof(uint32_t));
            read_back(&tail[1], pres.p + (ids - 1), sizeof(uint32_t));
            const uint64_t n = (uint64_t)tail[0] + tail[1];
            g->n = n;
            g->m = m;
            g->directed = directed != 0;
            g->weighted = weighted != 0;
            DevBuf<uint32_t> map32(n ? n : 1);
            GX_LAUNCH(k_fill_mapping, grid_persistent(8), 256, 0, pres.p, newid.p, ids, map32.p);
            if (weighted) g->out.w.alloc(m);
            GX_LAUNCH(k_relabel, grid_persistent(8), 256, 0, ukeys.p, m, newid.p, kw, weighted ? g->out.w.p : nullptr);
            g->out.rowptr.alloc(n + 1);
            g->out.col.alloc(m);
            rowptr_from_sorted_keys(ukeys.p, m, n, g->out.rowptr.p);
            GX_LAUNCH(k_keys_low32, grid_persistent(8), 256, 0, ukeys.p, m, g->out.col.p);
            if (mapping_out) {
                std::vector<uint32_t> h(n);
                GX_CUDA(cudaMemcpyAsync(h.data(), map32.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
                GX_CUDA(cudaStreamSynchronize(c.stream));
                uint64_t *mp = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
                if (!mp) throw Error(GX_ERR_OOM, "host allocation failed");
                for (uint64_t i = 0; i < n; i++) mp[i] = h[i];
                *mapping_out = mp;
            }
            GX_CUDA(cudaStreamSynchronize(c.stream));
        } catch (...) {
            delete g;
            throw;
        }
        *out = g;
    });
}
