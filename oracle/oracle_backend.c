/*
 * TEST INFRASTRUCTURE — CPU oracle for the Backend<'T> hot path.  NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libdiffsharp_cuda.so) never links, loads or calls it.
 *
 * PARITY UNPINNED: the reference's OpenBLAS backend is not in the reference tree and its tests hold
 * no Backend<'T> known-answer vectors.  The only in-tree expectation,
 * tests/DiffSharp.Tests/Script.fsx:7-16 (RepeatReshapeCopy_V_MRows), is checked in tests/.
 * What this oracle restates is the call-site semantics of src/DiffSharp/AD.Float32.fs /
 * AD.Float64.fs over GENUINE OpenBLAS — the same library family the reference deployment binds
 * (lib/OpenBLAS-v0.2.14/15) — resolved at run time from the wheel-bundled shared object
 * (scipy.libs/libscipy_openblas-*.so, 0.3.x, symbols prefixed "scipy_").
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC oracle_backend.c -ldl -lm).
 */
#include <dlfcn.h>
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- OpenBLAS entry points, resolved by oracle_init ---------------------------------------- */
typedef float (*sdot_t)(int, const float*, int, const float*, int);
typedef double (*ddot_t)(int, const double*, int, const double*, int);
typedef float (*sasum_t)(int, const float*, int);
typedef double (*dasum_t)(int, const double*, int);
typedef size_t (*isamax_t)(int, const float*, int);
typedef size_t (*idamax_t)(int, const double*, int);
typedef void (*sgemm_t)(int, int, int, int, int, int, float, const float*, int, const float*, int, float, float*, int);
typedef void (*dgemm_t)(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
typedef void (*sgemv_t)(int, int, int, int, float, const float*, int, const float*, int, float, float*, int);
typedef void (*dgemv_t)(int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
typedef void (*sger_t)(int, int, int, float, const float*, int, const float*, int, float*, int);
typedef void (*dger_t)(int, int, int, double, const double*, int, const double*, int, double*, int);
typedef void (*sgesv_t)(int*, int*, float*, int*, int*, float*, int*, int*);
typedef void (*dgesv_t)(int*, int*, double*, int*, int*, double*, int*, int*);
typedef void (*ssysv_t)(const char*, int*, int*, float*, int*, int*, float*, int*, float*, int*, int*);
typedef void (*dsysv_t)(const char*, int*, int*, double*, int*, int*, double*, int*, double*, int*, int*);
typedef void (*sgetrf_t)(int*, int*, float*, int*, int*, int*);
typedef void (*dgetrf_t)(int*, int*, double*, int*, int*, int*);
typedef void (*sgetri_t)(int*, float*, int*, int*, float*, int*, int*);
typedef void (*dgetri_t)(int*, double*, int*, int*, double*, int*, int*);

static struct {
    void* lib;
    sdot_t sdot; ddot_t ddot; sasum_t sasum; dasum_t dasum; sdot_t dummy;
    sasum_t snrm2; dasum_t dnrm2; isamax_t isamax; idamax_t idamax;
    sgemm_t sgemm; dgemm_t dgemm; sgemv_t sgemv; dgemv_t dgemv; sger_t sger; dger_t dger;
    sgesv_t sgesv; dgesv_t dgesv; ssysv_t ssysv; dsysv_t dsysv; sgetrf_t sgetrf; dgetrf_t dgetrf; sgetri_t sgetri; dgetri_t dgetri;
    void (*set_threads)(int);
    int (*get_threads)(void);
    char* (*get_config)(void);
} g;

static void* sym(const char* prefix, const char* name) {
    char buf[128];
    snprintf(buf, sizeof buf, "%s%s", prefix, name);
    void* p = dlsym(g.lib, buf);
    if (!p) fprintf(stderr, "oracle: OpenBLAS symbol %s not found\n", buf);
    return p;
}

/* path: OpenBLAS shared object; prefix: "scipy_" for the scipy wheel, "" for a stock build.
 * returns 0 on success. */
int oracle_init(const char* path, const char* prefix) {
    if (g.lib) return 0;
    g.lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!g.lib) { fprintf(stderr, "oracle: dlopen(%s): %s\n", path, dlerror()); return -1; }
#define S(field, name) *(void**)(&g.field) = sym(prefix, name); if (!g.field) return -2;
    S(sdot, "cblas_sdot") S(ddot, "cblas_ddot") S(sasum, "cblas_sasum") S(dasum, "cblas_dasum")
    S(snrm2, "cblas_snrm2") S(dnrm2, "cblas_dnrm2") S(isamax, "cblas_isamax") S(idamax, "cblas_idamax")
    S(sgemm, "cblas_sgemm") S(dgemm, "cblas_dgemm") S(sgemv, "cblas_sgemv") S(dgemv, "cblas_dgemv")
    S(sger, "cblas_sger") S(dger, "cblas_dger")
    S(sgesv, "sgesv_") S(dgesv, "dgesv_") S(ssysv, "ssysv_") S(dsysv, "dsysv_") S(sgetrf, "sgetrf_") S(dgetrf, "dgetrf_") S(sgetri, "sgetri_") S(dgetri, "dgetri_")
    S(set_threads, "openblas_set_num_threads") S(get_threads, "openblas_get_num_threads") S(get_config, "openblas_get_config")
#undef S
    return 0;
}
void oracle_set_num_threads(int n) { if (g.set_threads) g.set_threads(n); }
int oracle_get_num_threads(void) { return g.get_threads ? g.get_threads() : 0; }
const char* oracle_openblas_config(void) { return g.get_config ? g.get_config() : ""; }

#define T float
#define FN(name) oracle_s##name
#define BL(name) g.BLS_##name
#define BLS_dot sdot
#define BLS_asum sasum
#define BLS_nrm2 snrm2
#define BLS_iamax isamax
#define BLS_gemm sgemm
#define BLS_gemv sgemv
#define BLS_ger sger
#define BLS_gesv sgesv
#define BLS_sysv ssysv
#define BLS_getrf sgetrf
#define BLS_getri sgetri
#include "oracle_backend_impl.h"
#undef T
#undef FN
#undef BL

#define T double
#define FN(name) oracle_d##name
#define BL(name) g.BLD_##name
#define BLD_dot ddot
#define BLD_asum dasum
#define BLD_nrm2 dnrm2
#define BLD_iamax idamax
#define BLD_gemm dgemm
#define BLD_gemv dgemv
#define BLD_ger dger
#define BLD_gesv dgesv
#define BLD_sysv dsysv
#define BLD_getrf dgetrf
#define BLD_getri dgetri
#include "oracle_backend_impl.h"
