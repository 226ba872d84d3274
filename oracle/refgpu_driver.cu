// refgpu_driver.cu -- TEST INFRASTRUCTURE ONLY: C entry point around the reference's cdlp_gpu
// (cdlp_kernel.cu:1144-1359, compiled unchanged into oracle/_ref/libcdlp_ref.so by oracle/Makefile) so that
// tools/cdlp_vs_reference.py can time it on the same B200 and the same graph as gx_cdlp.  The timed window is the
// reference's own (cdlp_cuda.cu:241-243): cudaMalloc + H2D + kernels + D2H + the n-step setElement loop.
#include <chrono>
#include <cstring>

#include "cdlp_kernel.cuh"

extern "C" int ref_cdlp_gpu(const uint64_t *Ap, const uint64_t *Aj, uint64_t N, uint64_t nnz, int symmetric, int itermax,
                            uint64_t *labels_out, double *window_ms)
{
    GrB_Vector out = nullptr;
    cudaFree(0); // context creation is not part of the reference's window either (the process already holds one)
    const auto t0 = std::chrono::steady_clock::now();
    cdlp_gpu((GrB_Index *)Ap, (N + 1) * sizeof(GrB_Index), (GrB_Index *)Aj, nnz * sizeof(GrB_Index), &out, N, nnz, symmetric != 0, itermax);
    const auto t1 = std::chrono::steady_clock::now();
    if (window_ms) *window_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (!out) return -1;
    std::memcpy(labels_out, out->v.data(), N * sizeof(uint64_t));
    delete out;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
