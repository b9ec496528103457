/*
 * TEST INFRASTRUCTURE — CPU oracle, included twice by oracle_backend.c (T = float / double).
 *
 * Restates, for one precision, the CPU backend that the reference deployment plugs in behind
 * Backend<'T> (reference: src/DiffSharp/Backend.fs:74-139).  That backend is NOT in the reference
 * tree, so the semantics are those implied by the call sites in
 * src/DiffSharp/AD.Float32.fs / AD.Float64.fs (cited per function) and by upstream DiffSharp 0.7's
 * OpenBLAS backend conventions: BLAS-backed members call genuine OpenBLAS (resolved at run time from
 * the wheel-bundled library, see oracle_backend.c), everything else is a sequential loop in 'T with
 * double-precision libm for the transcendental functions (System.Math in a .NET 4.5 deployment).
 *
 * Macros expected: T (element type), FN(name) (symbol decoration), BL(name) (BLAS fn pointer).
 */

/* ---- scalar valued: Backend.fs:81-88 ---------------------------------------------------- */
T FN(dot)(int n, const T* a, const T* b) { return n ? BL(dot)(n, a, 1, b, 1) : (T)0; }   /* Mul_Dot_V_V, AD.Float32.fs:1109 */
T FN(asum)(int n, const T* a) { return n ? BL(asum)(n, a, 1) : (T)0; }                  /* L1Norm_V, :1530 */
T FN(nrm2)(int n, const T* a) { return n ? BL(nrm2)(n, a, 1) : (T)0; }                  /* L2Norm_V, :1539,1549 */
T FN(amax)(int n, const T* a) {                                                         /* SupNorm_V (0 call sites) */
    if (!n) return (T)0;
    size_t i = BL(iamax)(n, a, 1);
    return a[i] < 0 ? -a[i] : a[i];
}
T FN(sum)(int n, const T* a) {                                                          /* Sum_V / Sum_M, :1557,2807: sequential in 'T */
    T s = (T)0;
    for (int i = 0; i < n; ++i) s += a[i];
    return s;
}
int FN(argmax)(int n, const T* a) {                                                     /* MaxIndex_V, :2975; host twin :1629-1638 */
    int mi = 0;
    T mv = a[0];
    for (int i = 0; i < n; ++i)
        if (a[i] > mv) { mi = i; mv = a[i]; }
    return mi;
}
int FN(argmin)(int n, const T* a) {                                                     /* MinIndex_V, :2983; host twin :1641-1650 */
    int mi = 0;
    T mv = a[0];
    for (int i = 0; i < n; ++i)
        if (a[i] < mv) { mi = i; mv = a[i]; }
    return mi;
}

/* ---- elementwise ------------------------------------------------------------------------- */
static double FN(sgn)(double x) { return x < 0 ? -1.0 : (x > 0 ? 1.0 : 0.0); }          /* signummod, Util.fs:88-91 */

/* Map_F_V / Map_F_M (Backend.fs:105,130): closure text per op-code at AD.Float32.fs:2631-2934.
 * System.Math semantics: computed in double, rounded to 'T; Math.Round is half-to-even; F# `max 0 v`
 * propagates NaN.  returns 0, or -6 for a non-unary op-code. */
int FN(map)(int op, int n, const T* x, T* y) {
    for (int i = 0; i < n; ++i) {
        double v = (double)x[i], r;
        switch (op) {
            case 8: r = log(v); break;
            case 9: r = log10(v); break;
            case 10: r = exp(v); break;
            case 11: r = sin(v); break;
            case 12: r = cos(v); break;
            case 13: r = tan(v); break;
            case 14: r = sqrt(v); break;
            case 15: r = sinh(v); break;
            case 16: r = cosh(v); break;
            case 17: r = tanh(v); break;
            case 18: r = asin(v); break;
            case 19: r = acos(v); break;
            case 20: r = atan(v); break;
            case 21: r = fabs(v); break;
            case 22: r = FN(sgn)(v); break;
            case 23: r = floor(v); break;
            case 24: r = ceil(v); break;
            case 25: r = nearbyint(v); break; /* banker's rounding under FE_TONEAREST */
            case 26: r = (v != v) ? v : (v > 0 ? v : 0.0); break;
            case 27: {  /* number1 / (number1 + exp -v) evaluated in 'T with a double exp (:2934) */
                T e = (T)exp(-(double)x[i]);
                y[i] = (T)1 / ((T)1 + e);
                continue;
            }
            default: return -6;
        }
        y[i] = (T)r;
    }
    return 0;
}

/* Map_F_S_V / Map_F_S_M (Backend.fs:106,131).  DIV: s / v (:2496); POW_AB: v ** s (:2556);
 * POW_BA: s ** v (:2568); ATAN2_AB: atan2 v s (:2580); ATAN2_BA: atan2 s v (:2591). */
int FN(map_s)(int op, T s, int n, const T* x, T* y) {
    for (int i = 0; i < n; ++i) {
        switch (op) {
            case 1: y[i] = s / x[i]; break;
            case 4: y[i] = (T)pow((double)x[i], (double)s); break;
            case 5: y[i] = (T)pow((double)s, (double)x[i]); break;
            case 6: y[i] = (T)atan2((double)x[i], (double)s); break;
            case 7: y[i] = (T)atan2((double)s, (double)x[i]); break;
            default: return -6;
        }
    }
    return 0;
}

/* Map2_F_V_V / Map2_F_M_M (Backend.fs:107,132): MUL (:1121), DIV (:1145,2430), POW_AB (:1157,2441),
 * ATAN2_AB (:1169,2452); ADD/SUB serve Add_V_V/Sub_V_V/Add_M_M/Sub_M_M (:1085,1097,2358,2370) and
 * MUL serves Mul_Had_M_M (:2394). */
int FN(map2)(int op, int n, const T* a, const T* b, T* y) {
    for (int i = 0; i < n; ++i) {
        switch (op) {
            case 0: y[i] = a[i] * b[i]; break;
            case 1: y[i] = a[i] / b[i]; break;
            case 2: y[i] = a[i] + b[i]; break;
            case 3: y[i] = a[i] - b[i]; break;
            case 4: y[i] = (T)pow((double)a[i], (double)b[i]); break;
            case 6: y[i] = (T)atan2((double)a[i], (double)b[i]); break;
            default: return -6;
        }
    }
    return 0;
}

/* Add_S / Sub_S / Sub_?_S / Mul_S (Backend.fs:93-97,114,117-120) */
int FN(scalar_op)(int kind, T s, int n, const T* x, T* y) {
    for (int i = 0; i < n; ++i) {
        switch (kind) {
            case 0: y[i] = x[i] + s; break;
            case 1: y[i] = s - x[i]; break;
            case 2: y[i] = x[i] - s; break;
            case 3: y[i] = s * x[i]; break;
            default: return -6;
        }
    }
    return 0;
}

/* Add_V_V_InPlace(src, so, dst, do, n) (Backend.fs:91, :2866) and Add_M_M_InPlace (Backend.fs:113, :4118) */
void FN(add_inplace_range)(const T* src, int so, T* dst, int dof, int n) {
    for (int i = 0; i < n; ++i) dst[dof + i] += src[so + i];
}

/* fused adjoint chains: the same roundings the unfused call sequence performs.
 *  0 tanh   : secha = 1/cosh(aP) [Map(Cosh), Map_F_S(Div,1)]; (dA*secha)*secha      (:4238-4240)
 *  1 sigmoid: (dA*dP)*(1-dP)                                                         (:4281)
 *  2 relu   : dA*((sign(aP)+1)*0.5)   [the host multiplies by 1/2: (x+1)/2 = Mul_S(0.5)] (:4280) */
int FN(fused_bwd)(int kind, int n, const T* dA, const T* p, T* y) {
    for (int i = 0; i < n; ++i) {
        switch (kind) {
            case 0: { T c = (T)cosh((double)p[i]); T s = (T)1 / c; y[i] = (dA[i] * s) * s; break; }
            case 1: y[i] = (dA[i] * p[i]) * ((T)1 - p[i]); break;
            case 2: y[i] = dA[i] * (((T)FN(sgn)((double)p[i]) + (T)1) * (T)0.5); break;
            default: return -6;
        }
    }
    return 0;
}

/* ---- BLAS-2/3 ------------------------------------------------------------------------------ */
/* Mul_M_M (Backend.fs:119, :2382) and its transposed-operand forms (what Transpose_M + Mul_M_M
 * computes at :4127-4143); bias adds v[i] to every column (Mul_M_M_Add_V_MCols, Backend.fs:121). */
void FN(gemm)(int ta, int tb, int m, int n, int k, const T* A, const T* B, const T* bias, T* C) {
    if (k == 0) {
        for (size_t i = 0; i < (size_t)m * n; ++i) C[i] = (T)0;
    } else {
        BL(gemm)(101 /*RowMajor*/, ta ? 112 : 111, tb ? 112 : 111, m, n, k, (T)1, A, ta ? m : k, B, tb ? k : n, (T)0, C, n);
    }
    if (bias)
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < n; ++j) C[(size_t)i * n + j] += bias[i];
}
/* Mul_M_V (Backend.fs:98, :2406), Mul_V_M (Backend.fs:101, :2418), Mul_M_V_Add_V (Backend.fs:99) */
void FN(gemv)(int trans, int m, int n, const T* A, const T* x, const T* y, T* out) {
    int olen = trans ? n : m;
    for (int i = 0; i < olen; ++i) out[i] = y ? y[i] : (T)0;
    if (m && n) BL(gemv)(101, trans ? 112 : 111, m, n, (T)1, A, n, x, 1, (T)1, out, 1);
}
/* Mul_Out_V_V (Backend.fs:111, :1133) */
void FN(ger)(int m, int n, const T* x, const T* y, T* out) {
    for (size_t i = 0; i < (size_t)m * n; ++i) out[i] = (T)0;
    if (m && n) BL(ger)(101, m, n, (T)1, x, 1, y, 1, out, n);
}

/* ---- layout -------------------------------------------------------------------------------- */
void FN(transpose)(int m, int n, const T* A, T* B) {                                    /* Transpose_M, Backend.fs:127, :2788 */
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) B[(size_t)j * m + i] = A[(size_t)i * n + j];
}
void FN(permute)(int rank, const long long* shape, const int* perm, const T* in, T* out) { /* Permute_M, Backend.fs:128, :2773 */
    long long in_stride[8], out_shape[8], st[8], total = 1;
    for (int d = rank - 1; d >= 0; --d) { in_stride[d] = total; total *= shape[d]; }
    for (int d = 0; d < rank; ++d) { out_shape[d] = shape[perm[d]]; st[d] = in_stride[perm[d]]; }
    for (long long i = 0; i < total; ++i) {
        long long rem = i, src = 0;
        for (int d = rank - 1; d >= 0; --d) { src += (rem % out_shape[d]) * st[d]; rem /= out_shape[d]; }
        out[i] = in[src];
    }
}
void FN(repeat_rows)(int m, int n, const T* v, T* out) {                                /* RepeatReshapeCopy_V_MRows, Backend.fs:135, :1862; Script.fsx:10-16 */
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) out[(size_t)i * n + j] = v[j];
}
void FN(repeat_cols)(int m, int n, const T* v, T* out) {                                /* RepeatReshapeCopy_V_MCols, Backend.fs:136 */
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) out[(size_t)i * n + j] = v[i];
}
void FN(colsum_inplace)(int m, int n, const T* M, T* v) {                               /* Add_M_Colwise_V_InPlace, Backend.fs:123, :4255-4256: a.A <- a.A + row, row by row */
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) v[j] += M[(size_t)i * n + j];
}
void FN(add_v_mcols)(int m, int n, const T* v, const T* M, T* out) {                    /* Add_V_MCols, Backend.fs:115 */
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) out[(size_t)i * n + j] = M[(size_t)i * n + j] + v[i];
}
void FN(diag)(int m, int n, const T* A, T* out) {                                       /* Diagonal_M, Backend.fs:104, :2796 */
    int k = m < n ? m : n;
    for (int i = 0; i < k; ++i) out[i] = A[(size_t)i * n + i];
}
void FN(gather2d)(int cols, int r0, int r1, int c0, int c1, const T* A, T* out) {       /* GetStackedValues, Util.fs:138,204-210 */
    int oc = c1 - c0 + 1;
    for (int r = r0; r <= r1; ++r)
        for (int c = c0; c <= c1; ++c) out[(size_t)(r - r0) * oc + (c - c0)] = A[(size_t)r * cols + c];
}

/* ---- LAPACK class: Solve_M_V / SolveSymmetric_M_V / Inverse_M / Det_M (Backend.fs:102-103,125-126) --- */
/* returns 0, or -7 when LAPACK reports a singular factor (managed side: None) */
int FN(gesv)(int n, const T* A, const T* b, T* x) {
    T* At = (T*)malloc(sizeof(T) * (size_t)n * n);
    int* ipiv = (int*)malloc(sizeof(int) * (size_t)n);
    FN(transpose)(n, n, A, At); /* row-major -> column-major */
    memcpy(x, b, sizeof(T) * (size_t)n);
    int nn = n, nrhs = 1, info = 0;
    BL(gesv)(&nn, &nrhs, At, &nn, ipiv, x, &nn, &info);
    free(At);
    free(ipiv);
    return info == 0 ? 0 : -7;
}
/* SolveSymmetric_M_V (Backend.fs:103, call site AD.Float32.fs:2832): LAPACK ?sysv (Bunch-Kaufman LDL^T) on the
 * upper triangle of the column-major matrix = the lower triangle of the row-major one. */
int FN(sysv)(int n, const T* A, const T* b, T* x) {
    T* At = (T*)malloc(sizeof(T) * (size_t)n * n);
    int* ipiv = (int*)malloc(sizeof(int) * (size_t)n);
    int lwork = n * 64 > 64 ? n * 64 : 64;
    T* work = (T*)malloc(sizeof(T) * (size_t)lwork);
    FN(transpose)(n, n, A, At);
    memcpy(x, b, sizeof(T) * (size_t)n);
    int nn = n, nrhs = 1, info = 0;
    BL(sysv)("U", &nn, &nrhs, At, &nn, ipiv, x, &nn, work, &lwork, &info);
    free(At);
    free(ipiv);
    free(work);
    return info == 0 ? 0 : -7;
}
int FN(getri)(int n, const T* A, T* Ainv) {
    T* At = (T*)malloc(sizeof(T) * (size_t)n * n);
    int* ipiv = (int*)malloc(sizeof(int) * (size_t)n);
    FN(transpose)(n, n, A, At);
    int nn = n, info = 0, lwork = n * 64;
    T* work = (T*)malloc(sizeof(T) * (size_t)lwork);
    BL(getrf)(&nn, &nn, At, &nn, ipiv, &info);
    if (info == 0) BL(getri)(&nn, At, &nn, ipiv, work, &lwork, &info);
    if (info == 0) FN(transpose)(n, n, At, Ainv);
    free(At);
    free(ipiv);
    free(work);
    return info == 0 ? 0 : -7;
}
int FN(det)(int n, const T* A, T* out) {
    T* At = (T*)malloc(sizeof(T) * (size_t)n * n);
    int* ipiv = (int*)malloc(sizeof(int) * (size_t)n);
    FN(transpose)(n, n, A, At);
    int nn = n, info = 0;
    BL(getrf)(&nn, &nn, At, &nn, ipiv, &info);
    T d = (T)1;
    for (int i = 0; i < n; ++i) {
        d *= At[(size_t)i * n + i];
        if (ipiv[i] != i + 1) d = -d;
    }
    *out = d;
    free(At);
    free(ipiv);
    return info == 0 ? 0 : -7;
}
