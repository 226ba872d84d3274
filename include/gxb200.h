/*
 * gxb200.h -- C ABI of the B200-native execution path for the six LDBC
 * Graphalytics kernels (BFS, PR, WCC, CDLP, LCC, SSSP).
 *
 * Drop-in boundary.  The reference's wrappers hand a GraphBLAS matrix to one
 * LAGraph call per algorithm; its only array-level GPU precedent is
 *
 *   void cdlp_gpu(GrB_Index *Ap, GrB_Index Ap_size, GrB_Index *Aj, GrB_Index Aj_size,
 *                 GrB_Vector *CDLP_handle, GrB_Index N, GrB_Index nnz,
 *                 bool symmetric, int itermax);      (cdlp_kernel.cuh:22)
 *
 * i.e. host CSR arrays of GrB_Index (uint64_t) in, one dense result vector
 * out.  Every entry point below keeps that shape: borrowed host pointers and
 * sizes in, caller-allocated length-n host arrays out, `int` status back
 * (0 = GX_OK, negative = error, text via gx_last_error()).  No GraphBLAS, no
 * torch, no C++ types cross this boundary.  One context per process, not
 * thread-safe -- the reference runs one single-threaded process per job
 * (GraphblasJob.java:70-97).
 *
 * There is no CPU fallback: every compute entry point fails with
 * GX_ERR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef GXB200_H
#define GXB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GX_OK 0
#define GX_ERR_INVALID (-1)      /* bad argument (GrB_INVALID_VALUE / GrB_NULL_POINTER) */
#define GX_ERR_NO_DEVICE (-2)    /* no usable CUDA device / gx_init not called */
#define GX_ERR_CUDA (-3)         /* CUDA runtime or NCCL failure, see gx_last_error() */
#define GX_ERR_OOM (-4)          /* GrB_OUT_OF_MEMORY */
#define GX_ERR_IO (-5)           /* file could not be read / parsed */
#define GX_ERR_NOT_IMPLEMENTED (-6) /* GrB_NOT_IMPLEMENTED (cdlp_cuda.cu:190) */

#define GX_UNREACHED_LEVEL INT64_MAX /* bfs.cpp:59-63 prints 9223372036854775807 */

typedef struct gx_graph gx_graph; /* opaque: device-resident CSR (+ cached CSC, row blocks, ...) */

/* ---- context ---------------------------------------------------------------------------- */

/* Replaces LAGraph_Init + GxB_Global_Option_set(NTHREADS) (bfs.cpp:88-89): binds the process
 * to CUDA device `device` and creates the stream / memory pool.  Idempotent per device. */
int gx_init(int device);
int gx_finalize(void);
int gx_device_count(void);
/* Grows the library's device memory pool to `bytes` ahead of time (capped at 60 % of the free memory; best effort), so
 * that the algorithms' scratch buffers are not mapped on first use inside a timed window. */
int gx_reserve(uint64_t bytes);
const char *gx_last_error(void);

/* Multi-GPU: one process per GPU.  Rank 0 calls gx_comm_unique_id, ships the 128 bytes to the
 * other ranks by any means (torch.distributed / files / MPI), then every rank calls
 * gx_comm_init.  With a communicator set, graphs are 1-D row partitioned across ranks and the
 * algorithms exchange frontier / rank / label slices with NCCL. */
int gx_comm_unique_id(void *id128);
int gx_comm_init(int rank, int nranks, const void *id128);
int gx_comm_destroy(void);

/* ---- graph ------------------------------------------------------------------------------ */

/* Replaces GxB_Matrix_export_CSR + the H2D copies of cdlp_gpu (cdlp_cuda.cu:181,
 * cdlp_kernel.cu:1162-1196).  CSR by row over dense ids 0..n-1; row i lists the out-neighbours
 * of i; undirected graphs are stored symmetric (as LAGraph_MMRead expands them).  `weights`
 * NULL = structural (iso GrB_BOOL), else FP64 per entry.  Arrays are borrowed for the call
 * only; pinned host memory makes the upload faster but is not required.  Column ids inside a
 * row need not be sorted (they are sorted on the device). */
int gx_graph_create_csr(gx_graph **g, uint64_t n, uint64_t nnz, const uint64_t *rowptr,
                        const uint64_t *colidx, const double *weights, int directed);
/* Same with 32-bit column ids (halves the upload; n <= 2^32 - 2). */
int gx_graph_create_csr32(gx_graph **g, uint64_t n, uint64_t nnz, const uint64_t *rowptr,
                          const uint32_t *colidx, const double *weights, int directed);
/* gx_graph_create_csr32 + gx_graph_cache(what) in one call, for a caller that keeps the graph across several
 * algorithms (the analogue of LAGraph_Cached_AT, pr.cpp:58-59, folded into the upload).  The drop-in binaries pass
 * cache = 0 and let each algorithm build what it needs inside its timed window, as the reference does.  With GX_CACHE_AT on a directed graph on one
 * GPU the transposition is pipelined with the upload: the column ids travel in row-block chunks and
 * every chunk is validated, sorted by column and histogrammed while the next one is on the bus. */
int gx_graph_create_csr32_cached(gx_graph **g, uint64_t n, uint64_t nnz, const uint64_t *rowptr,
                                 const uint32_t *colidx, const double *weights, int directed, unsigned cache);
/* Replaces ReadMatrixMarket + ReadMapping (graphio.cpp:4-60): loads `dir`/graph.grb+graph.vtb
 * (binary != 0) or `dir`/graph.mtx+graph.vtx and uploads it.  *mapping (malloc'd uint64[n],
 * free with gx_free_host) receives the original vertex ids in dense order. */
int gx_graph_load(gx_graph **g, const char *dir, int binary, int directed, uint64_t **mapping,
                  uint64_t *n_out);
/* The text half of ReadMatrixMarket (graphio.cpp:10-24, LAGraph_MMRead of graph.mtx) with the body tokenised ON THE
 * DEVICE: the file's bytes are copied to HBM as they are, kernels find the entries, parse `row column [value]`
 * (FP64 values correctly rounded: the same doubles strtod returns) and build the CSR with the same cleaning as the
 * host parser (self-loops and repeated entries dropped, rows sorted, `symmetric` files mirrored).  `cache` as in
 * gx_graph_create_csr32_cached.  gx_graph_load uses it for text inputs unless GX_LOADER=host. */
int gx_graph_load_mtx(gx_graph **g, const char *path, int directed, unsigned cache);
void gx_free_host(void *p);
int gx_graph_free(gx_graph *g);
int gx_graph_info(const gx_graph *g, uint64_t *n, uint64_t *nnz, int *directed, int *weighted);
/* Build the cached structures an algorithm needs ahead of its first run -- the analogue of
 * LAGraph_Cached_AT / LAGraph_Cached_OutDegree (pr.cpp:58-59).  `what` is a bit set of GX_CACHE_*.
 * Algorithms build what is missing on first use. */
#define GX_CACHE_AT 1u      /* transposed adjacency (in-edges) */
#define GX_CACHE_LCC 2u     /* A v A' with multiplicities, degree-oriented */
int gx_graph_cache(gx_graph *g, unsigned what);
/* Download the device CSR back to host arrays (caller-allocated; colidx sorted per row). */
int gx_graph_download(const gx_graph *g, uint64_t *rowptr, uint32_t *colidx, double *weights);

/* ---- the six kernels ----------------------------------------------------------------------
 * Result arrays: host, length n, caller-allocated; NULL leaves the result on the device
 * (used to time the device path alone).  Vertex arguments are DENSE ids. */

/* LA_BFS (bfs.cpp:70-83) -> LAGr_BreadthFirstSearch(&level, NULL, G, src):
 * level[src] = 0, GX_UNREACHED_LEVEL where LAGraph leaves no entry. */
int gx_bfs(gx_graph *g, uint64_t src, int64_t *level);
/* LA_PR (pr.cpp:47-66) -> LAGr_PageRankGX(&r, &iters, G, (float)damping, itermax):
 * exactly `iters` iterations; damping is rounded through float like LAGraph's argument. */
int gx_pagerank(gx_graph *g, double damping, int iters, double *rank);
/* WeaklyConnectedComponents (wcc.cpp:39-66) -> A v A' + LAGr_ConnectedComponents (FastSV):
 * comp[v] = smallest dense id of v's component. */
int gx_wcc(gx_graph *g, uint64_t *comp);
/* MY_CDLP_GPU / LA_CDLP_CPU (cdlp.cpp:54-81) -> LAGraph_cdlp semantics (LAGraph_cdlp.c:241-333):
 * label[v] is a dense id (the caller maps it through mapping[], cdlp.cpp:48). */
int gx_cdlp(gx_graph *g, int itermax, uint64_t *label);
/* LA_LCC (lcc.cpp:61-71) -> LAGraph_lcc: 0.0 where LAGraph leaves no entry (degree < 2). */
int gx_lcc(gx_graph *g, double *lcc);
/* LA_SSSP (sssp.cpp:53-81) -> LAGr_SingleSourceShortestPath: +inf where unreached. */
int gx_sssp(gx_graph *g, uint64_t src, double *dist);

/* ---- measurement ------------------------------------------------------------------------ */

typedef struct gx_timing {
    double h2d_ms;              /* host->device copies of the last call */
    double build_ms;            /* device-side graph construction (sort, transpose, ...) */
    double kernel_ms;           /* the algorithm's kernels, CUDA events on the library stream */
    double comm_ms;             /* NCCL collectives (0 on one GPU) */
    double d2h_ms;              /* result download */
    uint64_t algorithmic_bytes; /* compulsory HBM bytes of the last run, DESIGN.md formulas */
    uint64_t edges_inspected;   /* adjacency entries actually read */
    uint32_t kernel_launches;   /* kernels this library launched in the last call */
    uint32_t iterations;        /* levels / iterations / sweeps executed */
} gx_timing;
int gx_last_timing(gx_timing *t);

/* CUDA-event stopwatch on the library's own stream (torch.cuda.Event cannot see it). */
int gx_timer_start(void);
int gx_timer_stop(double *elapsed_ms);
int gx_sync(void);
/* Overwrites a buffer larger than L2 (126 MB) so the next timed step starts cold. */
int gx_flush_l2(void);
/* Per-kernel timing session (CUDA event pair around every launch on the library stream).
 * gx_profile(1) clears and starts, gx_profile(0) stops; the report is one line per kernel,
 * "<name>\t<launches>\t<total ms>\n", sorted by total time. */
int gx_profile(int enable);
const char *gx_profile_report(void);
/* Pinned host memory for callers that want full-speed uploads. */
int gx_host_alloc(void **p, uint64_t bytes);
int gx_host_free(void *p);
/* Page-lock / release host arrays the caller already owns (the CSR arrays a loader filled); best effort. */
int gx_host_register(const void *p, uint64_t bytes);
int gx_host_unregister(const void *p);

/* ---- result files ------------------------------------------------------------------------
 * Replaces SerializeBFSResult / SerializePageRankResult / SerializeWCCResult / SerializeCDLPResult / SerializeLCCResult /
 * SerializeSSSPResult (bfs.cpp:11-68, pr.cpp:17-45, wcc.cpp:11-37, cdlp.cpp:21-52, lcc.cpp:17-59, sssp.cpp:11-51): one
 * line "<ids[i]> <value>" per vertex in array order; INT64 / UINT64 in decimal (UINT64: value_map[values[i]] when
 * value_map is given -- CDLP prints mapping[label], cdlp.cpp:48), FP64 as precision(16) << scientific with +inf written
 * as the literal `infinity` (sssp.cpp:41-46).  Formatted on all host threads; host-only, needs no device. */
#define GX_RESULT_INT64 0
#define GX_RESULT_UINT64 1
#define GX_RESULT_FP64 2
int gx_result_write(const char *path, int kind, const uint64_t *ids, const void *values, uint64_t n,
                    const uint64_t *value_map);

/* ---- load stage -------------------------------------------------------------------------
 * Replaces bin/py/relabel.py:8-79 (DuckDB joins) in bin/sh/load-graph.sh:49-60: X.v / X.e -> out_dir/graph.vtx (original id
 * of dense vertex k on line k, .v row order) and out_dir/graph.mtx (banner, `%%GraphBLAS GrB_BOOL|GrB_FP64`, `n n nnz`,
 * 1-based `src dst val` in .e order; weights kept as the text they came as).  Parsing, id lookup and formatting run on all
 * host threads; host-only.  bin/py/relabel.py of this repo is the same-flags front-end. */
int gx_relabel(const char *vertex_path, const char *edge_path, const char *out_dir, int weighted, int directed,
               uint64_t *n_out, uint64_t *nnz_out);

/* ---- synthetic inputs ----------------------------------------------------
 * Graph500 RMAT (0.57,0.19,0.19,0.05), `edgefactor` * 2^scale generated edges, counter-based
 * RNG keyed by `seed`, scrambled ids, self-loops and duplicates removed, isolated ids dropped.
 * Built entirely in HBM.  *mapping (optional, malloc'd) = original ids of the dense vertices. */
int gx_rmat_create(gx_graph **g, int scale, int edgefactor, uint64_t seed, int directed,
                   int weighted, uint64_t **mapping);
/* Max out-degree vertex, ties -> smallest dense id (the BFS/SSSP source convention). */
int gx_graph_max_degree_vertex(const gx_graph *g, uint64_t *v);

#ifdef __cplusplus
}
#endif
#endif /* GXB200_H */
