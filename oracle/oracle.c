/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped path.
 *
 * CPU restatement of the six LDBC Graphalytics kernels as the reference's
 * C++ wrappers call them through LAGraph (the reference's arithmetic lives
 * in SuiteSparse:GraphBLAS v7.4.4 + LAGraph `dev`, neither vendored under
 * /root/reference nor installed here, so this file restates the published
 * algorithms and is pinned against the reference's 24 golden output files in
 * example-data-sets/graphs/ -- see tests/test_oracle_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker or
 * the timed CPU baseline.  The product (libgxb200.so) never links it.
 *
 * Layout convention (matches the reference after GxB_Matrix_export_CSR,
 * cdlp_cuda.cu:181): CSR by row, rowptr uint64[n+1]; column ids are held as
 * uint32 here (n <= 2^26 in every config).  Row i lists the out-neighbours of
 * dense vertex i.  Undirected graphs are stored symmetric.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_UNREACHED INT64_MAX

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_threads(int t)
{
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/* ------------------------------------------------------------------------
 * Transpose (what LAGraph_Cached_AT does inside the reference's timed window,
 * pr.cpp:59).  Stable counting sort => rows of the result are sorted whenever
 * the input is processed in row order.
 * ---------------------------------------------------------------------- */
int oracle_transpose(uint64_t n, const uint64_t *rp, const uint32_t *ci, const double *w,
                     uint64_t *trp, uint32_t *tci, double *tw)
{
    /* Parallel and still stable (SuiteSparse's own transpose is an OpenMP bucket sort, so a serial
     * one would sandbag the CPU baseline): (1) column counts with atomic increments, (2) prefix sum,
     * (3) the columns are cut into one range per thread, balanced by entry count; every thread walks
     * ALL rows in order and scatters only the entries whose column lies in its range, so inside a
     * column the rows stay ascending and no two threads touch the same cursor. */
    const uint64_t m = rp[n];
    memset(trp, 0, (n + 1) * sizeof(uint64_t));
#pragma omp parallel for schedule(static)
    for (uint64_t e = 0; e < m; e++) __atomic_fetch_add(&trp[(uint64_t)ci[e] + 1], 1, __ATOMIC_RELAXED);
    for (uint64_t i = 0; i < n; i++) trp[i + 1] += trp[i];
    uint64_t *cur = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    if (!cur) return -1;
    memcpy(cur, trp, (n + 1) * sizeof(uint64_t));
    int nt = oracle_num_threads();
    if (nt < 1) nt = 1;
    if ((uint64_t)nt > n) nt = n ? (int)n : 1;
    uint64_t *cut = (uint64_t *)malloc(((size_t)nt + 1) * sizeof(uint64_t));
    if (!cut) { free(cur); return -1; }
    cut[0] = 0;
    for (int t = 1; t < nt; t++) {
        /* first column c with trp[c] >= t * m / nt */
        const uint64_t target = (uint64_t)(((__uint128_t)m * (uint64_t)t) / (uint64_t)nt);
        uint64_t lo = cut[t - 1], hi = n;
        while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (trp[mid] < target) lo = mid + 1; else hi = mid; }
        cut[t] = lo;
    }
    cut[nt] = n;
#pragma omp parallel for schedule(static, 1) num_threads(nt)
    for (int t = 0; t < nt; t++) {
        const uint64_t c0 = cut[t], c1 = cut[t + 1];
        if (c0 >= c1) continue;
        for (uint64_t i = 0; i < n; i++) {
            for (uint64_t e = rp[i]; e < rp[i + 1]; e++) {
                const uint64_t c = ci[e];
                if (c < c0 || c >= c1) continue;
                const uint64_t p = cur[c]++;
                tci[p] = (uint32_t)i;
                if (w && tw) tw[p] = w[e];
            }
        }
    }
    free(cut);
    free(cur);
    return 0;
}

/* ------------------------------------------------------------------------
 * BFS -- bfs.cpp:70-83 -> LAGr_BreadthFirstSearch(&level, NULL, G, src).
 * The wrapper caches neither AT nor the out-degree, so LAGraph runs its
 * push-only variant: level-synchronous frontier expansion, level(src) = 0,
 * unreached vertices have no entry (serialised as INT64_MAX, bfs.cpp:59-63).
 * ---------------------------------------------------------------------- */
int oracle_bfs(uint64_t n, const uint64_t *rp, const uint32_t *ci, uint64_t src, int64_t *level)
{
    if (src >= n) return -2;
    for (uint64_t i = 0; i < n; i++) level[i] = ORACLE_UNREACHED;
    uint32_t *q = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint32_t *q2 = (uint32_t *)malloc(n * sizeof(uint32_t));
    if (!q || !q2) { free(q); free(q2); return -1; }
    uint64_t qn = 1;
    q[0] = (uint32_t)src;
    level[src] = 0;
    int64_t depth = 0;
    while (qn) {
        depth++;
        uint64_t q2n = 0;
#pragma omp parallel
        {
            /* per-thread local buffer flushed into q2 */
            uint32_t buf[1024];
            int bn = 0;
#pragma omp for schedule(dynamic, 64) nowait
            for (uint64_t k = 0; k < qn; k++) {
                uint32_t u = q[k];
                for (uint64_t e = rp[u]; e < rp[u + 1]; e++) {
                    uint32_t v = ci[e];
                    if (__atomic_load_n(&level[v], __ATOMIC_RELAXED) == ORACLE_UNREACHED) {
                        int64_t expect = ORACLE_UNREACHED;
                        if (__atomic_compare_exchange_n(&level[v], &expect, depth, 0,
                                                        __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
                            buf[bn++] = v;
                            if (bn == 1024) {
                                uint64_t pos = __atomic_fetch_add(&q2n, (uint64_t)bn, __ATOMIC_RELAXED);
                                memcpy(q2 + pos, buf, bn * sizeof(uint32_t));
                                bn = 0;
                            }
                        }
                    }
                }
            }
            if (bn) {
                uint64_t pos = __atomic_fetch_add(&q2n, (uint64_t)bn, __ATOMIC_RELAXED);
                memcpy(q2 + pos, buf, bn * sizeof(uint32_t));
            }
        }
        uint32_t *t = q; q = q2; q2 = t;
        qn = q2n;
    }
    free(q); free(q2);
    return 0;
}

/* ------------------------------------------------------------------------
 * PageRank -- pr.cpp:47-66 -> LAGraph_Cached_OutDegree, LAGraph_Cached_AT,
 * LAGr_PageRankGX(&r, &iters, G, (float)damping, itermax): exactly `iters`
 * iterations, no tolerance.  r0 = 1/n;  d = outdeg / damping (prescaled);
 * each iteration: teleport' = (1-damping)/n + (damping/n) * sum_{sinks} r;
 * w = r ./ d;  r = teleport' + A' (plus.second) w.
 * damping is a `float` in LAGraph's signature, widened to double for FP64
 * arithmetic (pinned by test-pr-undirected-PR to 3e-16).
 * (trp,tci) is the transpose (in-edges); pass NULL to have it built here,
 * as the reference does inside its timed window.
 * ---------------------------------------------------------------------- */
int oracle_pagerank(uint64_t n, const uint64_t *rp, const uint32_t *ci,
                    const uint64_t *trp_in, const uint32_t *tci_in,
                    double damping_in, int iters, double *rank)
{
    if (n == 0) return 0;
    const double damping = damping_in; /* callers pass (double)(float)d, LAGraph's `float damping` */
    uint64_t m = rp[n];
    uint64_t *trp = NULL; uint32_t *tci = NULL;
    const uint64_t *tp = trp_in; const uint32_t *tc = tci_in;
    if (!tp || !tc) {
        trp = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
        tci = (uint32_t *)malloc((m ? m : 1) * sizeof(uint32_t));
        if (!trp || !tci) { free(trp); free(tci); return -1; }
        oracle_transpose(n, rp, ci, NULL, trp, tci, NULL);
        tp = trp; tc = tci;
    }
    double *d = (double *)malloc(n * sizeof(double));
    double *w = (double *)malloc(n * sizeof(double));
    double *t = (double *)malloc(n * sizeof(double));
    if (!d || !w || !t) { free(d); free(w); free(t); free(trp); free(tci); return -1; }
    const double teleport = (1.0 - damping) / (double)n;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < n; i++) {
        rank[i] = 1.0 / (double)n;
        uint64_t od = rp[i + 1] - rp[i];
        d[i] = (double)od / damping;
    }
    for (int it = 0; it < iters; it++) {
        double sink = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : sink)
        for (uint64_t i = 0; i < n; i++) {
            t[i] = rank[i];
            if (rp[i + 1] == rp[i]) { sink += t[i]; w[i] = 0.0; }
            else w[i] = t[i] / d[i];
        }
        const double tele = teleport + damping * sink / (double)n;
#pragma omp parallel for schedule(dynamic, 1024)
        for (uint64_t v = 0; v < n; v++) {
            double s = 0.0;
            for (uint64_t e = tp[v]; e < tp[v + 1]; e++) s += w[tc[e]];
            rank[v] = tele + s;
        }
    }
    free(d); free(w); free(t); free(trp); free(tci);
    return 0;
}

/* ------------------------------------------------------------------------
 * WCC -- wcc.cpp:39-66: directed => A = A v A' (GrB_LOR eWiseAdd with T1),
 * then LAGr_ConnectedComponents == FastSV.  Parent f, grandparent gp,
 * mngp[u] = min over neighbours of gp; hook f[f[u]] = min(., mngp[u]);
 * f = min(f, mngp, gp); shortcut gp = f[f]; stop when gp is stable.
 * Result: min dense index of each component (wcc.cpp:31-34 prints it raw).
 * (trp,tci) = in-edges, used only when `directed` != 0.
 * ---------------------------------------------------------------------- */
int oracle_wcc(uint64_t n, const uint64_t *rp, const uint32_t *ci,
               const uint64_t *trp, const uint32_t *tci, int directed, uint64_t *comp)
{
    uint32_t *f = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *gp = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *mngp = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *fold = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    if (!f || !gp || !mngp || !fold) { free(f); free(gp); free(mngp); free(fold); return -1; }
    for (uint64_t i = 0; i < n; i++) { f[i] = gp[i] = mngp[i] = (uint32_t)i; }
    int changed = 1;
    while (changed) {
        /* mngp = min(mngp, A min.second gp) */
#pragma omp parallel for schedule(dynamic, 1024)
        for (uint64_t u = 0; u < n; u++) {
            uint32_t mn = mngp[u];
            for (uint64_t e = rp[u]; e < rp[u + 1]; e++) { uint32_t g = gp[ci[e]]; if (g < mn) mn = g; }
            if (directed && trp)
                for (uint64_t e = trp[u]; e < trp[u + 1]; e++) { uint32_t g = gp[tci[e]]; if (g < mn) mn = g; }
            mngp[u] = mn;
        }
        /* hooking on the old parent vector: f[fold[u]] = min(f[fold[u]], mngp[u]) */
        memcpy(fold, f, n * sizeof(uint32_t));
        for (uint64_t u = 0; u < n; u++) {
            uint32_t p = fold[u];
            if (mngp[u] < f[p]) f[p] = mngp[u];
        }
        /* f = min(f, mngp, gp) */
#pragma omp parallel for schedule(static)
        for (uint64_t u = 0; u < n; u++) {
            uint32_t x = f[u];
            if (mngp[u] < x) x = mngp[u];
            if (gp[u] < x) x = gp[u];
            f[u] = x;
        }
        /* shortcut: gp' = f[f]; converged when gp' == gp */
        changed = 0;
#pragma omp parallel for schedule(static) reduction(| : changed)
        for (uint64_t u = 0; u < n; u++) {
            uint32_t g = f[f[u]];
            if (g != gp[u]) changed |= 1;
            fold[u] = g;
        }
        memcpy(gp, fold, n * sizeof(uint32_t));
    }
    for (uint64_t i = 0; i < n; i++) comp[i] = f[i];
    free(f); free(gp); free(mngp); free(fold);
    return 0;
}

/* ------------------------------------------------------------------------
 * CDLP -- cdlp.cpp:54-67 -> LAGraph_cdlp; semantic spec LAGraph_cdlp.c:241-333:
 * L0(v) = v.  Each iteration every stored edge (i,j) carries label(j)
 * (S = S min.second L, :272); for directed graphs the transposed matrix
 * contributes the in-neighbours as a second multiset (:280-283), so a
 * reciprocal pair counts twice.  The (row,label) tuples are sorted (:286) and
 * the first longest run per row wins (:293-323) == smallest most frequent
 * label.  Vertices with no neighbours keep their label (Graphalytics spec;
 * pinned by the goldens).  Stops at itermax or at a fix-point (:328-332).
 * ---------------------------------------------------------------------- */
static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

int oracle_cdlp(uint64_t n, const uint64_t *rp, const uint32_t *ci,
                const uint64_t *trp, const uint32_t *tci, int directed, int itermax,
                uint64_t *label_out)
{
    uint32_t *lab = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *nxt = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    if (!lab || !nxt) { free(lab); free(nxt); return -1; }
    for (uint64_t i = 0; i < n; i++) lab[i] = (uint32_t)i;
    int err = 0;
    for (int it = 0; it < itermax; it++) {
        int diff = 0;
#pragma omp parallel reduction(| : diff)
        {
            uint64_t cap = 64;
            uint32_t *buf = (uint32_t *)malloc(cap * sizeof(uint32_t));
#pragma omp for schedule(dynamic, 256)
            for (uint64_t v = 0; v < n; v++) {
                uint64_t deg = rp[v + 1] - rp[v];
                if (directed && trp) deg += trp[v + 1] - trp[v];
                if (deg == 0) { nxt[v] = lab[v]; continue; }
                if (deg > cap) {
                    cap = deg * 2;
                    free(buf);
                    buf = (uint32_t *)malloc(cap * sizeof(uint32_t));
                    if (!buf) { err = 1; cap = 0; nxt[v] = lab[v]; continue; }
                }
                uint64_t k = 0;
                for (uint64_t e = rp[v]; e < rp[v + 1]; e++) buf[k++] = lab[ci[e]];
                if (directed && trp)
                    for (uint64_t e = trp[v]; e < trp[v + 1]; e++) buf[k++] = lab[tci[e]];
                qsort(buf, k, sizeof(uint32_t), cmp_u32);
                uint32_t best = buf[0]; uint64_t bestlen = 0, run = 1;
                for (uint64_t j = 1; j <= k; j++) {
                    if (j == k || buf[j] != buf[j - 1]) {
                        if (run > bestlen) { bestlen = run; best = buf[j - 1]; }
                        run = 0;
                    }
                    run++;
                }
                nxt[v] = best;
                if (best != lab[v]) diff |= 1;
            }
            free(buf);
        }
        uint32_t *t = lab; lab = nxt; nxt = t;
        if (!diff) break;
    }
    for (uint64_t i = 0; i < n; i++) label_out[i] = lab[i];
    free(lab); free(nxt);
    return err ? -1 : 0;
}

/* ------------------------------------------------------------------------
 * LCC -- lcc.cpp:61-71 -> LAGraph_lcc(&d, A, symmetric=!directed, false,...).
 * N(v) = (in U out neighbours) \ {v}, d = |N(v)|;
 * lcc(v) = #{(a,b) in E : a in N(v), b in N(v)} / (d (d-1)),  0 if d < 2.
 * For a symmetric store each undirected neighbour pair counts twice, which is
 * the Graphalytics 2*tri/(d(d-1)).  Vertices with d < 2 have no entry in
 * LAGraph's result and are serialised as 0.0 (lcc.cpp:49-54).
 * `subset` (optional): only these `ns` vertices are evaluated, others get NaN
 * -- used for spot checks at full benchmark scale.
 * ---------------------------------------------------------------------- */
int oracle_lcc(uint64_t n, const uint64_t *rp, const uint32_t *ci,
               const uint64_t *trp, const uint32_t *tci, int directed,
               const uint64_t *subset, uint64_t ns, double *lcc)
{
    int err = 0;
    uint64_t total = subset ? ns : n;
    if (subset) for (uint64_t i = 0; i < n; i++) lcc[i] = NAN;
#pragma omp parallel
    {
        /* mark[x] == v+1  <=>  x in N(v) */
        uint64_t *mark = (uint64_t *)calloc(n ? n : 1, sizeof(uint64_t));
        uint64_t cap = 64;
        uint32_t *nb = (uint32_t *)malloc(cap * sizeof(uint32_t));
        if (!mark || !nb) err = 1;
        /* small chunks: one hub costs as much as thousands of ordinary vertices, and a subset may hold several */
#pragma omp for schedule(dynamic, 4)
        for (uint64_t k = 0; k < total; k++) {
            if (err) continue;
            uint64_t v = subset ? subset[k] : k;
            uint64_t maxd = rp[v + 1] - rp[v];
            if (directed && trp) maxd += trp[v + 1] - trp[v];
            if (maxd > cap) {
                cap = maxd * 2; free(nb);
                nb = (uint32_t *)malloc(cap * sizeof(uint32_t));
                if (!nb) { err = 1; continue; }
            }
            uint64_t d = 0;
            for (uint64_t e = rp[v]; e < rp[v + 1]; e++) {
                uint32_t x = ci[e];
                if (x != v && mark[x] != v + 1) { mark[x] = v + 1; nb[d++] = x; }
            }
            if (directed && trp)
                for (uint64_t e = trp[v]; e < trp[v + 1]; e++) {
                    uint32_t x = tci[e];
                    if (x != v && mark[x] != v + 1) { mark[x] = v + 1; nb[d++] = x; }
                }
            if (d < 2) { lcc[v] = 0.0; continue; }
            uint64_t cnt = 0;
            for (uint64_t i = 0; i < d; i++) {
                uint32_t a = nb[i];
                uint32_t prev = UINT32_MAX; /* rows are sorted: skip duplicate entries */
                for (uint64_t e = rp[a]; e < rp[a + 1]; e++) {
                    uint32_t b = ci[e];
                    if (b == prev) continue;
                    prev = b;
                    if (b != a && mark[b] == v + 1) cnt++;
                }
            }
            lcc[v] = (double)cnt / ((double)d * (double)(d - 1));
        }
        free(mark); free(nb);
    }
    return err ? -1 : 0;
}

/* ------------------------------------------------------------------------
 * SSSP -- sssp.cpp:53-81: zero diagonal, LAGr_SingleSourceShortestPath with
 * delta = 2.5 (delta-stepping over min.plus FP64).  The bucket width does not
 * affect the result: distances are the fix-point of
 * d(v) = min_u fl(d(u) + w(u,v)), d(src) = 0, which Dijkstra reaches as well
 * because rounded addition is monotone.  Unreached => +inf (serialised as the
 * literal `infinity`, sssp.cpp:41-46).
 * ---------------------------------------------------------------------- */
typedef struct { double d; uint32_t v; } heap_item;

static void heap_push(heap_item *h, uint64_t *hn, double d, uint32_t v)
{
    uint64_t i = (*hn)++;
    while (i > 0) {
        uint64_t p = (i - 1) >> 1;
        if (h[p].d <= d) break;
        h[i] = h[p]; i = p;
    }
    h[i].d = d; h[i].v = v;
}

static heap_item heap_pop(heap_item *h, uint64_t *hn)
{
    heap_item top = h[0];
    heap_item last = h[--(*hn)];
    uint64_t i = 0, nn = *hn;
    for (;;) {
        uint64_t c = 2 * i + 1;
        if (c >= nn) break;
        if (c + 1 < nn && h[c + 1].d < h[c].d) c++;
        if (h[c].d >= last.d) break;
        h[i] = h[c]; i = c;
    }
    if (nn) h[i] = last;
    return top;
}

int oracle_sssp(uint64_t n, const uint64_t *rp, const uint32_t *ci, const double *w,
                uint64_t src, double *dist)
{
    if (src >= n) return -2;
    uint64_t m = rp[n];
    for (uint64_t i = 0; i < n; i++) dist[i] = INFINITY;
    heap_item *h = (heap_item *)malloc((m + n + 1) * sizeof(heap_item));
    if (!h) return -1;
    uint64_t hn = 0;
    dist[src] = 0.0;
    heap_push(h, &hn, 0.0, (uint32_t)src);
    while (hn) {
        heap_item it = heap_pop(h, &hn);
        if (it.d > dist[it.v]) continue;
        uint32_t u = it.v;
        for (uint64_t e = rp[u]; e < rp[u + 1]; e++) {
            double nd = it.d + w[e];
            uint32_t v = ci[e];
            if (nd < dist[v]) { dist[v] = nd; heap_push(h, &hn, nd, v); }
        }
    }
    free(h);
    return 0;
}

/* ------------------------------------------------------------------------
 * Synthetic input generator shared bit-for-bit with the device generator
 * (csrc/rmat.cu): Graph500-style RMAT (A,B,C,D)=(.57,.19,.19,.05), counter
 * based (splitmix64 keyed by seed, edge index and level pair) so any slice of
 * the edge list can be produced independently.  SURVEY.md 8(d).
 * ---------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* bijection on [0, 2^scale): two rounds of odd-multiply + xorshift */
uint64_t oracle_scramble(uint64_t v, int scale, uint64_t seed)
{
    const uint64_t mask = (scale >= 64) ? ~0ull : ((1ull << scale) - 1);
    const int sh = scale / 2 + 1;
    uint64_t k0 = splitmix64(seed ^ 0xA5A5A5A5ull), k1 = splitmix64(seed ^ 0x5A5A5A5A5Aull);
    uint64_t x = v & mask;
    x = (x * 0x9E3779B97F4A7C15ull + k0) & mask;
    x ^= x >> sh;
    x = (x * 0xD1B54A32D192ED03ull + k1) & mask;
    x ^= x >> sh;
    return x;
}

void oracle_rmat_edges(int scale, uint64_t seed, uint64_t first, uint64_t count,
                       uint64_t *src, uint64_t *dst)
{
    const uint32_t tA = (uint32_t)(0.57 * 4294967296.0);
    const uint32_t tAB = (uint32_t)(0.76 * 4294967296.0);
    const uint32_t tABC = (uint32_t)(0.95 * 4294967296.0);
#pragma omp parallel for schedule(static)
    for (uint64_t k = 0; k < count; k++) {
        uint64_t i = first + k;
        uint64_t s = 0, d = 0, h = 0;
        for (int l = 0; l < scale; l++) {
            if ((l & 1) == 0) h = splitmix64(seed + (i << 5) + (uint64_t)(l >> 1));
            uint32_t r = (l & 1) ? (uint32_t)(h >> 32) : (uint32_t)h;
            uint32_t q = (r < tA) ? 0u : (r < tAB) ? 1u : (r < tABC) ? 2u : 3u;
            s = (s << 1) | (q >> 1);
            d = (d << 1) | (q & 1);
        }
        src[k] = oracle_scramble(s, scale, seed);
        dst[k] = oracle_scramble(d, scale, seed);
    }
}

/* weight in (0,1], symmetric in (a,b): hash of the ORIGINAL (scrambled) ids */
double oracle_edge_weight(uint64_t a, uint64_t b, uint64_t seed)
{
    uint64_t lo = a < b ? a : b, hi = a < b ? b : a;
    uint64_t h = splitmix64(splitmix64(seed ^ 0x57E1687ull) ^ (lo * 0x100000001B3ull + hi));
    return (double)((h >> 11) + 1) * (1.0 / 9007199254740992.0);
}

void oracle_edge_weights(uint64_t count, const uint64_t *a, const uint64_t *b, uint64_t seed, double *w)
{
#pragma omp parallel for schedule(static)
    for (uint64_t k = 0; k < count; k++) w[k] = oracle_edge_weight(a[k], b[k], seed);
}

/* ------------------------------------------------------------------------
 * Host-side construction of the benchmark graph for the CPU reference arm
 * (bench.py --impl reference), on all host threads: the same graph as the device
 * generator (csrc/rmat.cu) -- self-loops and duplicate edges dropped, isolated ids
 * removed, dense id = rank of the scrambled id, directed.  numpy's single-threaded sort
 * of the 2^(scale+4) keys took minutes from scale 24 on.
 * Caller-allocated: ids[2^scale], rowptr[2^scale + 1], colidx[ef * 2^scale],
 * scratch[2 * ef * 2^scale] (uint64).  Returns n (dense vertices), *m_out = entries; < 0 on error.
 * ---------------------------------------------------------------------- */
static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

int64_t oracle_rmat_csr(int scale, uint64_t seed, uint64_t nedges, uint64_t *ids, uint64_t *rowptr,
                        uint32_t *colidx, uint64_t *scratch, uint64_t *m_out)
{
    if (scale < 1 || scale > 31) return -2;
    const uint64_t N = 1ull << scale;
    uint64_t *keys = scratch, *tmp = scratch + nedges;
    uint8_t *present = (uint8_t *)calloc(N, 1);
    uint32_t *lut = (uint32_t *)malloc(N * sizeof(uint32_t));
    if (!present || !lut) { free(present); free(lut); return -1; }
    /* edges as (src << 32 | dst) of scrambled ids; tmp holds the dst half during generation */
    oracle_rmat_edges(scale, seed, 0, nedges, keys, tmp);
#pragma omp parallel for schedule(static)
    for (uint64_t k = 0; k < nedges; k++) {
        const uint64_t s = keys[k], d = tmp[k];
        if (s != d) { present[s] = 1; present[d] = 1; } /* benign: every writer stores 1 */
        keys[k] = (s << 32) | d;
    }
    uint64_t n = 0;
    for (uint64_t v = 0; v < N; v++)
        if (present[v]) { lut[v] = (uint32_t)n; ids[n++] = v; }
    free(present);
    /* dense keys; self-loops become the sentinel (sorts last) */
#pragma omp parallel for schedule(static)
    for (uint64_t k = 0; k < nedges; k++) {
        const uint64_t s = keys[k] >> 32, d = keys[k] & 0xFFFFFFFFull;
        keys[k] = (s == d) ? ~0ull : (((uint64_t)lut[s] << 32) | lut[d]);
    }
    free(lut);
    /* bucket sort by the top bits of the dense source id, then qsort inside the buckets */
    int nb_bits = 12;
    int id_bits = 1;
    while (id_bits < 32 && (1ull << id_bits) < n) id_bits++;
    if (nb_bits > id_bits) nb_bits = id_bits;
    const uint64_t NB = (1ull << nb_bits) + 1; /* + the sentinel bucket */
    const int shift = 32 + id_bits - nb_bits;
    int nt = oracle_num_threads();
    if (nt < 1) nt = 1;
    uint64_t *hist = (uint64_t *)calloc((size_t)nt * NB, sizeof(uint64_t));
    uint64_t *start = (uint64_t *)malloc((NB + 1) * sizeof(uint64_t));
    uint64_t *uniq = (uint64_t *)calloc(NB + 1, sizeof(uint64_t));
    if (!hist || !start || !uniq) { free(hist); free(start); free(uniq); return -1; }
#define BUCKET(key) ((key) == ~0ull ? NB - 1 : (uint64_t)((key) >> shift))
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num();
        const uint64_t a = nedges * (uint64_t)t / (uint64_t)nt, b = nedges * (uint64_t)(t + 1) / (uint64_t)nt;
        uint64_t *h = hist + (size_t)t * NB;
        for (uint64_t k = a; k < b; k++) h[BUCKET(keys[k])]++;
    }
    uint64_t acc = 0;
    for (uint64_t bkt = 0; bkt < NB; bkt++) {
        start[bkt] = acc;
        for (int t = 0; t < nt; t++) { const uint64_t c = hist[(size_t)t * NB + bkt]; hist[(size_t)t * NB + bkt] = acc; acc += c; }
    }
    start[NB] = acc;
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num();
        const uint64_t a = nedges * (uint64_t)t / (uint64_t)nt, b = nedges * (uint64_t)(t + 1) / (uint64_t)nt;
        uint64_t *h = hist + (size_t)t * NB;
        for (uint64_t k = a; k < b; k++) tmp[h[BUCKET(keys[k])]++] = keys[k];
    }
#undef BUCKET
    /* sort + dedupe every real bucket in place (the sentinel bucket is dropped) */
    const uint64_t NBR = NB - 1; /* real buckets */
#pragma omp parallel for schedule(dynamic, 1)
    for (uint64_t bkt = 0; bkt < NBR; bkt++) {
        uint64_t *p = tmp + start[bkt];
        const uint64_t c = start[bkt + 1] - start[bkt];
        if (!c) continue;
        qsort(p, c, sizeof(uint64_t), cmp_u64);
        uint64_t o = 1;
        for (uint64_t k = 1; k < c; k++)
            if (p[k] != p[k - 1]) p[o++] = p[k];
        uniq[bkt + 1] = o;
    }
    for (uint64_t bkt = 0; bkt + 1 < NB; bkt++) uniq[bkt + 1] += uniq[bkt];
    const uint64_t m = uniq[NB - 1];
    memset(rowptr, 0, (n + 1) * sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 1)
    for (uint64_t bkt = 0; bkt < NBR; bkt++) {
        const uint64_t *p = tmp + start[bkt];
        const uint64_t c = uniq[bkt + 1] - uniq[bkt], o = uniq[bkt];
        for (uint64_t k = 0; k < c; k++) {
            colidx[o + k] = (uint32_t)p[k];
            __atomic_fetch_add(&rowptr[(p[k] >> 32) + 1], 1, __ATOMIC_RELAXED); /* a row may span two buckets */
        }
    }
    for (uint64_t i = 0; i < n; i++) rowptr[i + 1] += rowptr[i];
    free(hist); free(start); free(uniq);
    *m_out = m;
    return (int64_t)n;
}
