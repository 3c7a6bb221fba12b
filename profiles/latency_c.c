/* Latency of the host-synchronous C-ABI calls as a C (or Julia ccall) caller sees them -- no Python / ctypes in the timed loop.
 *   gcc -O2 -std=c99 -Iinclude profiles/latency_c.c -Lstarformationhistories.jl_b200 -lsfhcuda -Wl,-rpath,$PWD/starformationhistories.jl_b200 -lm -o /tmp/latency_c
 * Prints one JSON line per case: median / mean / p90 microseconds per sfh_eval_fg (with and without the gradient) and per
 * sfh_eval_fg_hier (PowerLawMZR + GaussianDispersion). */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "sfhcuda.h"

#define CHECK(call)                                                                              \
    do {                                                                                         \
        int st_ = (call);                                                                        \
        if (st_ != SFH_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, st_, sfh_last_error()); exit(1); } \
    } while (0)

static double now_us(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return 1e6 * (double)t.tv_sec + 1e-3 * (double)t.tv_nsec;
}
static int cmp(const void *a, const void *b) { const double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }
static void stats(double *t, int n, double *med, double *mean, double *p90) {
    double s = 0.0;
    int i;
    for (i = 0; i < n; ++i) s += t[i];
    qsort(t, (size_t)n, sizeof(double), cmp);
    *med = t[n / 2]; *mean = s / n; *p90 = t[(int)(0.9 * n)];
}
static double urand(unsigned long long *s) { *s = *s * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(*s >> 11) / 9007199254740992.0; }

static void flat(const char *label, long long nb, long long nt, int n) {
    sfh_stack *s = NULL; sfh_ctx *c = NULL;
    double *x = malloc(sizeof(double) * (size_t)nt), *G = malloc(sizeof(double) * (size_t)nt), *t = malloc(sizeof(double) * (size_t)n);
    double nl = 0.0, m1, a1, p1, m2, a2, p2;
    unsigned long long seed = 12345;
    long long i;
    for (i = 0; i < nt; ++i) x[i] = 100.0 * urand(&seed);
    CHECK(sfh_stack_create_synthetic(&s, nb, nt, SFH_F64, 1, 1.0, x, NULL));
    CHECK(sfh_ctx_create(s, NULL, &c));
    for (i = 0; i < 30; ++i) CHECK(sfh_eval_fg(c, x, &nl, G, NULL));
    for (i = 0; i < n; ++i) { const double t0 = now_us(); CHECK(sfh_eval_fg(c, x, &nl, G, NULL)); t[i] = now_us() - t0; }
    stats(t, n, &m1, &a1, &p1);
    for (i = 0; i < 30; ++i) CHECK(sfh_eval_fg(c, x, &nl, NULL, NULL));
    for (i = 0; i < n; ++i) { const double t0 = now_us(); CHECK(sfh_eval_fg(c, x, &nl, NULL, NULL)); t[i] = now_us() - t0; }
    stats(t, n, &m2, &a2, &p2);
    printf("{\"case\": \"%s\", \"nb\": %lld, \"nt\": %lld, \"caller\": \"C\", \"calls\": %d, \"eval_fg_median_us\": %.2f, \"eval_fg_mean_us\": %.2f, \"eval_fg_p90_us\": %.2f, "
           "\"eval_f_only_median_us\": %.2f, \"eval_f_only_mean_us\": %.2f, \"neg_logL\": %.12g}\n", label, nb, nt, n, m1, a1, p1, m2, a2, nl);
    fflush(stdout);
    sfh_ctx_destroy(c); sfh_stack_destroy(s); free(x); free(G); free(t);
}

static void hier(const char *label, long long nb, int nj, int nk, int n) {
    const long long nt = (long long)nj * nk;
    sfh_stack *s = NULL; sfh_ctx *c = NULL;
    double *la = malloc(sizeof(double) * (size_t)nt), *mh = malloc(sizeof(double) * (size_t)nt), *x = malloc(sizeof(double) * (size_t)nt);
    double *v = malloc(sizeof(double) * (size_t)(nj + 3)), *G = malloc(sizeof(double) * (size_t)(nj + 3)), *t = malloc(sizeof(double) * (size_t)n);
    const double fixed[4] = {6.0, 0, 0, 0};
    const uint8_t mask[3] = {1, 1, 1};
    double nl = 0.0, m1, a1, p1;
    unsigned long long seed = 777;
    int64_t nages = 0;
    int j, k, i;
    sfh_opts o; memset(&o, 0, sizeof o); o.struct_size = (int32_t)sizeof o;
    for (j = 0; j < nj; ++j)
        for (k = 0; k < nk; ++k) { la[j * nk + k] = 10.1 - 3.5 * j / (nj - 1.0); mh[j * nk + k] = -2.5 + 2.5 * k / (nk - 1.0); }
    for (j = 0; j < nj; ++j) v[j] = 1e6 * urand(&seed);
    v[nj] = 1.0; v[nj + 1] = -2.0; v[nj + 2] = 0.2;
    /* truth coefficients from the library itself: a throw-away stack, then the real one whose data follow M * coeffs */
    for (i = 0; i < nt; ++i) x[i] = 1.0;
    CHECK(sfh_stack_create_synthetic(&s, 1024, nt, SFH_F64, 2, 1e-5, x, &o));
    CHECK(sfh_ctx_create(s, NULL, &c));
    CHECK(sfh_hier_bind(c, la, mh, &nages));
    CHECK(sfh_calculate_coeffs(c, SFH_MH_POWERLAW_MZR, fixed, SFH_DISP_GAUSSIAN, v, x));
    sfh_ctx_destroy(c); sfh_stack_destroy(s); s = NULL; c = NULL;
    CHECK(sfh_stack_create_synthetic(&s, nb, nt, SFH_F64, 2, 1e-5, x, &o));
    CHECK(sfh_ctx_create(s, NULL, &c));
    CHECK(sfh_hier_bind(c, la, mh, &nages));
    for (j = 0; j < nj + 3; ++j) v[j] *= 1.03;
    for (i = 0; i < 30; ++i) CHECK(sfh_eval_fg_hier(c, SFH_MH_POWERLAW_MZR, fixed, SFH_DISP_GAUSSIAN, v, mask, &nl, G));
    for (i = 0; i < n; ++i) { const double t0 = now_us(); CHECK(sfh_eval_fg_hier(c, SFH_MH_POWERLAW_MZR, fixed, SFH_DISP_GAUSSIAN, v, mask, &nl, G)); t[i] = now_us() - t0; }
    stats(t, n, &m1, &a1, &p1);
    printf("{\"case\": \"%s\", \"nb\": %lld, \"nt\": %lld, \"caller\": \"C\", \"calls\": %d, \"eval_fg_hier_median_us\": %.2f, \"eval_fg_hier_mean_us\": %.2f, \"eval_fg_hier_p90_us\": %.2f, \"neg_logL\": %.12g}\n",
           label, nb, nt, n, m1, a1, p1, nl);
    fflush(stdout);
    sfh_ctx_destroy(c); sfh_stack_destroy(s); free(la); free(mh); free(x); free(v); free(G); free(t);
}

int main(void) {
    flat("config1 100x100 bins x 100", 10000, 100, 3000);
    flat("config2 stack 200x200 x 500", 40000, 500, 2000);
    flat("config3 200x300 x 2400", 60000, 2400, 1500);
    hier("mzr_test 100x100 x 546", 10000, 21, 26, 2000);
    hier("config3 hier 60 x 40", 60000, 60, 40, 1500);
    return 0;
}
