/* Canonical fp32 decision rule of the B200 VQ kernels, restated in plain C.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/vq_oracle.py header).  Built by
 * oracle/Makefile into oracle/_build/libvq_canon.so and loaded with ctypes by
 * tests/ and bench.py's cpu_baseline leg; never by the product package.
 *
 * What it restates: the nearest-code decision of EuclideanCodebook.forward,
 * /root/reference/timevqvae/models/vq.py:210-222 —
 *     dist = -( sum(x^2) - (2x) @ e^T + sum(e^2) );  ind = argmax(dist)   (first index on ties)
 * The reference evaluates the three terms in fp32 with whatever summation
 * order its BLAS picks.  The CUDA kernels and this file pin the order-free
 * version of the same formula: each of the three reductions is accumulated in
 * fp64 (products of two fp32 numbers are exact in fp64) and rounded ONCE to
 * fp32, then combined with the reference's two fp32 roundings
 *     d_k = fl32( fl32( x2 - xe2_k ) + e2_k ),   ind = first k minimising d_k.
 * The fp64 accumulation follows the kernels' lane structure exactly (32
 * partial sums over 16-byte chunks, xor-butterfly), so the GPU result is
 * reproducible here bit for bit.  Against the reference's own fp32 sgemm the
 * indices can differ only on rows whose top-2 scores are within a couple of
 * ulps (tests/ report those rows separately).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* Minimal row-parallel driver (no OpenMP runtime in the image): rows are split
 * into contiguous slabs, one pthread per online core.  Every row is computed
 * independently, so the result does not depend on the thread count.          */
typedef void (*row_fn)(int64_t r0, int64_t r1, void *ctx);
typedef struct { row_fn fn; void *ctx; int64_t r0, r1; } slab_t;
static void *slab_main(void *p) { slab_t *s = (slab_t *)p; s->fn(s->r0, s->r1, s->ctx); return NULL; }
static int g_threads = 0;
void tvq_canon_set_threads(int t) { g_threads = t; }
int tvq_canon_get_threads(void)
{
    if (g_threads > 0) return g_threads;
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}
static void parallel_rows(int64_t n, row_fn fn, void *ctx)
{
    int t = tvq_canon_get_threads();
    if (t > 64) t = 64;
    if ((int64_t)t > n) t = n > 0 ? (int)n : 1;
    pthread_t th[64];
    slab_t slab[64];
    int64_t per = (n + t - 1) / t;
    for (int i = 0; i < t; ++i) {
        slab[i].fn = fn; slab[i].ctx = ctx;
        slab[i].r0 = i * per < n ? i * per : n;
        slab[i].r1 = (i + 1) * per < n ? (i + 1) * per : n;
        if (i == t - 1) slab_main(&slab[i]);
        else pthread_create(&th[i], NULL, slab_main, &slab[i]);
    }
    for (int i = 0; i < t - 1; ++i) pthread_join(th[i], NULL);
}

/* sum_i a[i]*b[i] in fp64 with the kernels' summation tree:
 * chunk c = elements [4c, 4c+4) belongs to lane c % 32; each lane adds its
 * elements in increasing index order; lanes are combined by an xor butterfly
 * with offsets 16, 8, 4, 2, 1 (lane 0's value is the result).               */
static double canon_dot(const float *a, const float *b, int d)
{
    double p[32], t[32];
    int nchunk = (d + 3) / 4;
    for (int l = 0; l < 32; ++l) {
        double acc = 0.0;
        for (int c = l; c < nchunk; c += 32) {
            int hi = 4 * c + 4 < d ? 4 * c + 4 : d;
            for (int i = 4 * c; i < hi; ++i)
                acc = fma((double)a[i], (double)b[i], acc);
        }
        p[l] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < 32; ++l) t[l] = p[l] + p[l ^ off];
        memcpy(p, t, sizeof p);
    }
    return p[0];
}

/* e2[k] = fl32(|e_k|^2), the per-code constant of the score. */
void tvq_canon_norms(const float *codebook, int k, int d, float *e2)
{
    for (int j = 0; j < k; ++j)
        e2[j] = (float)canon_dot(codebook + (size_t)j * d, codebook + (size_t)j * d, d);
}

/* Canonical score of one (latent, code) pair: what the reference calls -dist. */
static inline float canon_score(float x2, double xe, float e2)
{
    volatile float xe2 = (float)(2.0 * xe);      /* fl32 of (2x).e ; x2 is exact */
    volatile float t = x2 - xe2;                 /* volatile: keep the two fp32 roundings apart */
    volatile float dk = t + e2;
    return dk;
}

typedef struct { const float *x, *cb, *e2; int k, d; int64_t *idx; float *best, *second, *out; } assign_ctx;

static void assign_rows(int64_t r0, int64_t r1, void *p)
{
    assign_ctx *c = (assign_ctx *)p;
    for (int64_t r = r0; r < r1; ++r) {
        const float *xr = c->x + (size_t)r * c->d;
        float x2 = (float)canon_dot(xr, xr, c->d);
        float b1 = INFINITY, b2 = INFINITY;
        int64_t arg = 0;
        for (int j = 0; j < c->k; ++j) {
            float dk = canon_score(x2, canon_dot(xr, c->cb + (size_t)j * c->d, c->d), c->e2[j]);
            if (c->out) c->out[(size_t)r * c->k + j] = dk;
            if (dk < b1) { b2 = b1; b1 = dk; arg = j; }
            else if (dk < b2) { b2 = dk; }
        }
        if (c->idx) c->idx[r] = arg;
        if (c->best) c->best[r] = b1;
        if (c->second) c->second[r] = b2;
    }
}

/* Nearest code per latent.  best/second (optional) receive the smallest and
 * second smallest canonical score of each row (for margin accounting).       */
void tvq_canon_assign(const float *x, const float *codebook, int64_t n, int k, int d,
                      int64_t *idx, float *best, float *second)
{
    float *e2 = (float *)malloc(sizeof(float) * (size_t)k);
    tvq_canon_norms(codebook, k, d, e2);
    assign_ctx c = { x, codebook, e2, k, d, idx, best, second, NULL };
    parallel_rows(n, assign_rows, &c);
    free(e2);
}

/* Full N x K canonical score matrix (small cases only). */
void tvq_canon_scores(const float *x, const float *codebook, int64_t n, int k, int d, float *out)
{
    float *e2 = (float *)malloc(sizeof(float) * (size_t)k);
    tvq_canon_norms(codebook, k, d, e2);
    assign_ctx c = { x, codebook, e2, k, d, NULL, NULL, NULL, out };
    parallel_rows(n, assign_rows, &c);
    free(e2);
}

/* Straight-through output, commitment-loss sum and EMA batch statistics for
 * given indices (vq.py:225-234, :358-364): q_st = x + (e[idx] - x) as two
 * rounded fp32 ops, loss_sum = sum (q_st - x)^2 in fp64, counts and per-code
 * sums of x in fp64.  Single-threaded, order-independent up to fp64 rounding. */
void tvq_canon_apply(const float *x, const float *codebook, const int64_t *idx, int64_t n, int k, int d,
                     float *q_st, double *loss_sum, double *counts, double *embed_sum)
{
    double loss = 0.0;
    memset(counts, 0, sizeof(double) * (size_t)k);
    memset(embed_sum, 0, sizeof(double) * (size_t)k * d);
    for (int64_t r = 0; r < n; ++r) {
        const float *xr = x + (size_t)r * d;
        const float *er = codebook + (size_t)idx[r] * d;
        counts[idx[r]] += 1.0;
        for (int i = 0; i < d; ++i) {
            volatile float diff = er[i] - xr[i];
            volatile float q = xr[i] + diff;
            volatile float back = q - xr[i];
            if (q_st) q_st[(size_t)r * d + i] = q;
            loss += (double)back * (double)back;
            embed_sum[(size_t)idx[r] * d + i] += (double)xr[i];
        }
    }
    *loss_sum = loss;
}
