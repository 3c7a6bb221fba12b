/* Plain-C consumer of include/sfhcuda.h: proves the header is valid C99 (no C++-isms), that a C program links against
 * libsfhcuda.so, and exercises the host-only entry points (checksum, container files, the BFGS engine through a C callback)
 * plus the "no device -> SFH_ERR_NO_DEVICE, never a CPU fallback" contract when run on a box without a GPU.
 * Built and run by tests/test_abi_cpu.py::test_plain_c_consumer.  Exit code 0 = all checks passed. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "sfhcuda.h"

#define CHECK(cond)                                                         \
    do {                                                                    \
        if (!(cond)) {                                                      \
            fprintf(stderr, "FAILED %s:%d: %s (last error: %s)\n", __FILE__, __LINE__, #cond, sfh_last_error()); \
            return 1;                                                       \
        }                                                                   \
    } while (0)

/* f(x) = sum_i (i+1) (x_i - i)^2 */
static int quad(void *user, const double *x, int64_t n, double *f, double *g) {
    int64_t i;
    int *calls = (int *)user;
    ++*calls;
    *f = 0.0;
    for (i = 0; i < n; ++i) {
        const double d = x[i] - (double)i;
        *f += (double)(i + 1) * d * d;
        g[i] = 2.0 * (double)(i + 1) * d;
    }
    return 0;
}

int main(int argc, char **argv) {
    int ndev = -1, calls = 0, kind = -1, narr = -1, idx = -2;
    uint64_t cs = 0;
    double x[4] = {5.0, 5.0, 5.0, 5.0}, invH[16];
    const double payload[6] = {1.0, 2.0, 3.0, 4.0, 5.0, 6.0};
    const void *ptrs[1];
    const void *data = NULL;
    int64_t attrs[8] = {42, 0, 0, 0, 0, 0, 0, -7}, back[8];
    sfh_array_desc d, got;
    sfh_bfgs_opts bo;
    sfh_bfgs_report rep;
    sfh_file *f = NULL;
    sfh_stack *stack = NULL;
    const char *path = argc > 1 ? argv[1] : "/tmp/abi_c_smoke.sfh";

    CHECK(sfh_version() == SFH_VERSION_MAJOR * 100 + SFH_VERSION_MINOR);
    CHECK(sfh_device_count(&ndev) == SFH_OK && ndev >= 0);

    /* BFGS engine through a C callback */
    memset(&bo, 0, sizeof bo);
    bo.struct_size = (int32_t)sizeof bo;
    bo.g_abstol = 1e-10;
    CHECK(sfh_minimize_bfgs(quad, &calls, 4, x, &bo, &rep, invH) == SFH_OK);
    CHECK(rep.converged == 1 && rep.status == 0 && rep.f_calls == calls);
    CHECK(fabs(x[0]) < 1e-8 && fabs(x[3] - 3.0) < 1e-8);
    /* true inverse Hessian = diag(1 / (2 (i+1))); the BFGS estimate after a handful of steps is only approximate */
    CHECK(invH[0] > 0.25 && invH[0] < 1.0 && invH[15] > 0.06 && invH[15] < 0.25 && fabs(invH[1] - invH[4]) < 1e-12);

    /* container: write, reopen, verify */
    CHECK(sfh_checksum64(payload, (int64_t)sizeof payload, &cs) == SFH_OK && cs != 0);
    memset(&d, 0, sizeof d);
    strcpy(d.name, "payload");
    d.dtype = SFH_F64; d.ndim = 2; d.dims[0] = 3; d.dims[1] = 2;
    ptrs[0] = payload;
    CHECK(sfh_file_write(path, SFH_FILE_GENERIC, attrs, 1, &d, ptrs) == SFH_OK);
    CHECK(sfh_file_open(path, &f) == SFH_OK);
    CHECK(sfh_file_info(f, &kind, &narr, back) == SFH_OK && kind == SFH_FILE_GENERIC && narr == 1 && back[0] == 42 && back[7] == -7);
    CHECK(sfh_file_find(f, "payload", &idx) == SFH_OK && idx == 0);
    CHECK(sfh_file_find(f, "absent", &idx) == SFH_OK && idx == -1);
    CHECK(sfh_file_array(f, 0, &got, &data) == SFH_OK && got.nbytes == 48 && got.checksum == cs && got.dims[0] == 3);
    CHECK(memcmp(data, payload, sizeof payload) == 0);
    CHECK(sfh_file_verify(f, -1) == SFH_OK);
    CHECK(sfh_file_close(f) == SFH_OK);
    CHECK(sfh_stack_create_from_file(&stack, path, 0, NULL) == SFH_ERR_IO); /* a generic file is not a stack file */
    remove(path);

    if (ndev == 0) { /* the product never computes on the CPU */
        CHECK(sfh_stack_create(&stack, payload, 3, 2, SFH_F64, payload, SFH_F64, NULL) == SFH_ERR_NO_DEVICE && stack == NULL);
        CHECK(strlen(sfh_last_error()) > 0);
    }
    printf("abi_c_smoke: ok (%d device(s))\n", ndev);
    return 0;
}
