/* graphio.h -- SHIM, test infrastructure only.  Stands in for the reference's include/graphio.h (which pulls in
 * GraphBLAS.h / LAGraph.h, absent from this image) so that the reference's own CUDA CDLP
 * (/root/reference/src/main/c/src/algorithms/cdlp_kernel.cu, compiled UNCHANGED from where it lies) builds for
 * sm_100a as the GPU bar for CDLP (BASELINE.md "the kernel to beat").  Only the handful of GraphBLAS names that
 * file touches are provided: the index type and a vector that stores what cdlp_gpu hands back
 * (cdlp_kernel.cu:1350-1356).  Never linked into libgxb200.so. */
#pragma once
#include <stdint.h>
#include <vector>

typedef uint64_t GrB_Index;
struct RefVector { std::vector<uint64_t> v; };
typedef RefVector *GrB_Vector;
enum RefType { GrB_UINT64 = 0 };
static inline int GrB_Vector_new(GrB_Vector *out, RefType, GrB_Index n) { *out = new RefVector(); (*out)->v.assign(n, 0); return 0; }
static inline int GrB_Vector_setElement_UINT64(GrB_Vector v, uint64_t x, GrB_Index i) { v->v[i] = x; return 0; }
