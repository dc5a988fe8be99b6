/* oracle.c -- C restatement of the reference's CPU path.          *** TEST INFRASTRUCTURE ***
 *
 * Same algorithm, same order of operations as the reference's Python (citations are paths under
 * the reference repo); used by tests/ as a second checker next to oracle/int4_oracle.py and by
 * bench.py as the multi-threaded CPU baseline ("kind": "port").  Never part of the product path.
 * Build: make -C oracle  ->  oracle/_build/liboracle.so   (gcc -O2 -pthread -ffp-contract=off:
 * no FMA contraction, so every step is the same IEEE fp32 operation torch / numpy perform).
 * Rows are split over all online cores with plain pthreads (this image's gcc has no libgomp).
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

typedef void (*row_fn)(int64_t n0, int64_t n1, void* ctx);
typedef struct { row_fn fn; int64_t n0, n1; void* ctx; } job_t;
static void* job_main(void* a) { job_t* j = (job_t*)a; j->fn(j->n0, j->n1, j->ctx); return NULL; }

int oracle_max_threads(void) {
    const char* e = getenv("ORACLE_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}

static void parallel_rows(row_fn fn, int64_t N, void* ctx) {
    int T = oracle_max_threads();
    if (T > N) T = (int)(N > 0 ? N : 1);
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < T; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].n0 = N * t / T; jobs[t].n1 = N * (t + 1) / T;
        if (t > 0) pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    job_main(&jobs[0]);
    for (int t = 1; t < T; ++t) pthread_join(th[t], NULL);
}

static float clampf_torch(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* python/quantize.py:38-124 */
typedef struct { const float* w; int64_t K; uint8_t* packed; float* scales; float* zps; } qctx_t;
static void quantize_rows(int64_t n0, int64_t n1, void* c) {
    const qctx_t* q = (const qctx_t*)c;
    const float* w = q->w; const int64_t K = q->K; uint8_t* packed = q->packed; float* scales = q->scales; float* zps = q->zps;
    for (int64_t n = n0; n < n1; ++n) {
        const float* r = w + n * K;
        float mn = r[0], mx = r[0];
        for (int64_t k = 1; k < K; ++k) {                       /* :73-74 */
            if (r[k] < mn) mn = r[k];
            if (r[k] > mx) mx = r[k];
        }
        float scale = (mx - mn) / 15.0f;                         /* :80 */
        if (mx == mn) {                                          /* :85-92 */
            float a = fabsf(mx);
            scale = (a < 1.0f ? 1.0f : a) / 15.0f;
        }
        if (scale < 1e-8f) scale = 1e-8f;                        /* :94 */
        float zp = clampf_torch(rintf(-mn / scale), 0.0f, 15.0f); /* :100-101 */
        scales[n] = scale;
        zps[n] = zp;
        for (int64_t b = 0; b < K / 2; ++b) {                    /* :106-122 */
            float q0 = clampf_torch(rintf(r[2 * b] / scale + zp), 0.0f, 15.0f);
            float q1 = clampf_torch(rintf(r[2 * b + 1] / scale + zp), 0.0f, 15.0f);
            packed[n * (K / 2) + b] = (uint8_t)(((unsigned)q1 << 4) | (unsigned)q0);
        }
    }
}
void oracle_quantize_weights(const float* w, int64_t N, int64_t K, uint8_t* packed, float* scales, float* zps) {
    qctx_t c = {w, K, packed, scales, zps};
    parallel_rows(quantize_rows, N, &c);
}

/* python/quantize.py:127-173: (q - zp) * scale */
typedef struct { const uint8_t* packed; const float* scales; const float* zps; int64_t K; float* out; } dctx_t;
static void dequantize_rows(int64_t n0, int64_t n1, void* c) {
    const dctx_t* d = (const dctx_t*)c;
    const uint8_t* packed = d->packed; const float* scales = d->scales; const float* zps = d->zps;
    const int64_t K = d->K; float* out = d->out;
    for (int64_t n = n0; n < n1; ++n) {
        const float s = scales[n], z = zps[n];
        const uint8_t* p = packed + n * (K / 2);
        float* o = out + n * K;
        for (int64_t b = 0; b < K / 2; ++b) {
            o[2 * b] = ((float)(p[b] & 0x0F) - z) * s;           /* :152, :162, :172 */
            o[2 * b + 1] = ((float)(p[b] >> 4) - z) * s;         /* :153, :163, :172 */
        }
    }
}
void oracle_dequantize_weights(const uint8_t* packed, const float* scales, const float* zps, int64_t N, int64_t K,
                               float* out) {
    dctx_t c = {packed, scales, zps, K, out};
    parallel_rows(dequantize_rows, N, &c);
}

/* python/quantize.py:176-202: materialise the fp32 weights (into w_scratch [N,K]), then
 * y[m,n] = sum_k x[m,k] * w[n,k] (F.linear), fp32 accumulation in k order. */
typedef struct { const float* x; int64_t M, N, K; const float* w; float* y; } lctx_t;
static void linear_rows(int64_t n0, int64_t n1, void* c) {
    const lctx_t* l = (const lctx_t*)c;
    const float* x = l->x; const int64_t M = l->M, N = l->N, K = l->K; const float* w_scratch = l->w; float* y = l->y;
    for (int64_t n = n0; n < n1; ++n) {
        const float* wr = w_scratch + n * K;
        for (int64_t m = 0; m < M; ++m) {
            const float* xr = x + m * K;
            /* 8 independent partial sums (what a vectorised BLAS dot does), combined at the end */
            float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int64_t k = 0;
            for (; k + 8 <= K; k += 8)
                for (int j = 0; j < 8; ++j) a[j] += xr[k + j] * wr[k + j];
            float t = 0.0f;
            for (; k < K; ++k) t += xr[k] * wr[k];
            y[m * N + n] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7])) + t;
        }
    }
}
void oracle_reference_quantized_linear(const float* x, int64_t M, const uint8_t* packed, const float* scales,
                                       const float* zps, int64_t N, int64_t K, float* w_scratch, float* y) {
    oracle_dequantize_weights(packed, scales, zps, N, K, w_scratch);
    lctx_t c = {x, M, N, K, w_scratch, y};
    parallel_rows(linear_rows, N, &c);
}
