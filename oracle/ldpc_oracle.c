/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU oracle: a plain-C restatement of the reference's flooding BP syndrome decoders
 * (/root/reference/src/qkd_ldpc_algorithm.cpp:3-1258) and of the three helpers they call
 * (array_and_matrix_operations.cpp:936-972). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path (qkd_ldpc_v_b200/csrc,
 * qkd_ldpc_v_b200/host) never does and fails loudly without the CUDA library.
 *
 * Parity pin: the reference ships no tests or golden vectors. The pin is the reference
 * itself: oracle/_ref/libqkdref.so is built from the unmodified sources (oracle/Makefile) and
 *   - tests/test_oracle_vs_ref.py (runs where /root/reference exists) demands bit-identical words,
 *     iteration counts and flags between this file's f64 flavour and the compiled reference;
 *   - tests/golden/ holds outputs of the compiled reference (made by tests/golden/make_golden.py)
 *     that this file must reproduce wherever the tests run, including the N=6 known answer of
 *     example/qkd_ldpc_example.cpp.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL double
#define REAL_MAX DBL_MAX
#define SFX(name) name##_f64
#include "ldpc_oracle_body.inc"
#undef REAL
#undef REAL_MAX
#undef SFX

#define REAL float
#define REAL_MAX FLT_MAX
#define SFX(name) name##_f32
#include "ldpc_oracle_body.inc"
#undef REAL
#undef REAL_MAX
#undef SFX

/* CSC (bit_nodes) from CSR (check_nodes) in ascending check order -- what every reference loader produces for
 * the shipped matrices ("sorted ascending, consistent and duplicate-free"). */
void oracle_csr_to_csc(int n, int m, const int *row_ptr, const int *col_idx, int *col_ptr, int *row_idx) {
    memset(col_ptr, 0, sizeof(int) * (size_t)(n + 1));
    for (int e = 0; e < row_ptr[m]; ++e) col_ptr[col_idx[e] + 1]++;
    for (int i = 0; i < n; ++i) col_ptr[i + 1] += col_ptr[i];
    int *cur = (int *)malloc(sizeof(int) * (size_t)n);
    memcpy(cur, col_ptr, sizeof(int) * (size_t)n);
    for (int j = 0; j < m; ++j)
        for (int e = row_ptr[j]; e < row_ptr[j + 1]; ++e) row_idx[cur[col_idx[e]]++] = j;
    free(cur);
}

typedef struct {
    int precision; /* 64 or 32 */
    int n, m;
    const int *row_ptr, *col_idx, *col_ptr, *row_idx;
    int alg, max_iter, enable_thr;
    double primary, secondary, thr;
    const uint8_t *alice, *bob; /* [frames][n], one byte per bit, EXTENDED frames */
    const double *qber;         /* per frame */
    const int *punct, *shortd;
    int n_punct, n_short;
    uint8_t *bits_out; /* [frames][n] or NULL */
    int32_t *iters;
    uint8_t *flags; /* bit0 syndromes_match, bit1 keys_match (arrays_equal(alice, bob_solution), :1087) */
    long lo, hi;
} job_t;

static void *worker(void *arg) {
    job_t *J = (job_t *)arg;
    const int n = J->n, m = J->m;
    int *a = (int *)malloc(sizeof(int) * (size_t)n), *b = (int *)malloc(sizeof(int) * (size_t)n);
    int *out = (int *)malloc(sizeof(int) * (size_t)n), *syn = (int *)malloc(sizeof(int) * (size_t)m);
    double *l64 = (double *)malloc(sizeof(double) * (size_t)n);
    float *l32 = (float *)malloc(sizeof(float) * (size_t)n);
    for (long f = J->lo; f < J->hi; ++f) {
        for (int i = 0; i < n; ++i) { a[i] = J->alice[(size_t)f * n + i]; b[i] = J->bob[(size_t)f * n + i]; }
        int match = 0, it;
        if (J->precision == 64) {
            oracle_frame_setup_f64(n, m, J->row_ptr, J->col_idx, a, b, J->qber[f], J->punct, J->n_punct, J->shortd,
                                   J->n_short, l64, syn);
            it = oracle_decode_f64(n, m, J->row_ptr, J->col_idx, J->col_ptr, J->row_idx, J->alg, l64, syn,
                                   J->max_iter, J->primary, J->secondary, J->enable_thr, J->thr, out, &match, NULL);
        } else {
            oracle_frame_setup_f32(n, m, J->row_ptr, J->col_idx, a, b, J->qber[f], J->punct, J->n_punct, J->shortd,
                                   J->n_short, l32, syn);
            it = oracle_decode_f32(n, m, J->row_ptr, J->col_idx, J->col_ptr, J->row_idx, J->alg, l32, syn,
                                   J->max_iter, J->primary, J->secondary, J->enable_thr, J->thr, out, &match, NULL);
        }
        int keys = 1;
        for (int i = 0; i < n; ++i) if (a[i] != out[i]) { keys = 0; break; }
        J->iters[f] = it;
        J->flags[f] = (uint8_t)((match ? 1 : 0) | (keys ? 2 : 0));
        if (J->bits_out) for (int i = 0; i < n; ++i) J->bits_out[(size_t)f * n + i] = (uint8_t)out[i];
    }
    free(a); free(b); free(out); free(syn); free(l64); free(l32);
    return NULL;
}

/* QKD_LDPC / QKD_LDPC_RATE_ADAPT (:1031-1258) for a batch of independent frames over `threads` host threads
 * (static block partition, like BS::thread_pool::detach_loop at simulation.cpp:740-746). */
int oracle_qkd_ldpc_batch(int precision, int n, int m, const int *row_ptr, const int *col_idx, const int *col_ptr,
                          const int *row_idx, int alg, int max_iter, double primary, double secondary, int enable_thr,
                          double thr, long n_frames, const uint8_t *alice, const uint8_t *bob, const double *qber,
                          const int *punct, int n_punct, const int *shortd, int n_short, uint8_t *bits_out,
                          int32_t *iters, uint8_t *flags, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    job_t jobs[256];
    pthread_t th[256];
    for (int t = 0; t < threads; ++t) {
        job_t J = {precision, n, m, row_ptr, col_idx, col_ptr, row_idx, alg, max_iter, enable_thr, primary,
                   secondary, thr, alice, bob, qber, punct, shortd, n_punct, n_short, bits_out, iters, flags,
                   n_frames * t / threads, n_frames * (t + 1) / threads};
        jobs[t] = J;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    return 0;
}
