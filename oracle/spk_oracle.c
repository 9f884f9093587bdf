/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement ("oracle") of the numeric hot path of Sparspak.jl:
 *   _lufactor!  / _lulsolve! / _luusolve!   src/SparseMethod/SpkLUFactor.jl:60-255, 269-323, 325-377
 *   _ldltfactor! / _ldltsolve! / _pchole!   src/SparseSpdMethod/SpkLDLtFactor.jl:58-246, 266-293, 347-378
 *   _ldindx! _igathr! _assmb! _mmpyi! _luswap!  src/SparseSpdMethod/SpkSpdMMOps.jl:41-175
 * with the dense arithmetic in the loop order of the reference's generic kernels
 *   ggetrf! ggemm! ggemv! gtrsm! glaswp!    src/Utilities/GenericBlasLapackFragments.jl:56-495
 * (ggetrf!'s first-max pivot rule, :64-74, is the pivot-sequence specification).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's library.  The product path (libspkb200.so) never does.
 *
 * Parity pinning: LU path pinned by the reference's golden vectors
 * (test/test_sparse_method.jl:87-90,127-131,172-175,219-227) — see tests/test_oracle_golden.py.
 * SPD path: "parity unpinned" by the reference's own suite (its only SPD test is
 * disabled and the code as written is defective); this file
 * implements the INTENDED LDL^T and is cross-checked against the LU path.
 *
 * Same schedule and storage as the reference: left-looking, per-target linked
 * lists (LIFO), rectangular supernode blocks, 1-based int64 index arrays.
 *
 * For the CPU-baseline timing the four dense call sites (gemm / trsm / getrf /
 * gemv) can be redirected to a BLAS/LAPACK library (OpenBLAS from the SciPy
 * wheel) with spko_use_blas(): that is what the reference executes for Float64
 * (SpkSpdMMOps.jl:222-351).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <dlfcn.h>

typedef int64_t I;
#define API __attribute__((visibility("default")))

/* ---------------------------------------------------------------- BLAS hooks */
typedef void (*dgemm_t)(const char*, const char*, const int*, const int*, const int*, const double*,
                        const double*, const int*, const double*, const int*, const double*, double*, const int*);
typedef void (*dtrsm_t)(const char*, const char*, const char*, const char*, const int*, const int*,
                        const double*, const double*, const int*, double*, const int*);
typedef void (*dgetrf_t)(const int*, const int*, double*, const int*, int*, int*);
typedef void (*dgemv_t)(const char*, const int*, const int*, const double*, const double*, const int*,
                        const double*, const int*, const double*, double*, const int*);
static dgemm_t  p_dgemm  = 0;
static dtrsm_t  p_dtrsm  = 0;
static dgetrf_t p_dgetrf = 0;
static dgemv_t  p_dgemv  = 0;
static void* blas_handle = 0;

/* Load an LP64 BLAS/LAPACK (symbol prefix e.g. "scipy_" for the SciPy OpenBLAS). 0 on success. */
API int spko_use_blas(const char* path, const char* prefix) {
    char name[128];
    if (!path) { p_dgemm = 0; p_dtrsm = 0; p_dgetrf = 0; p_dgemv = 0; return 0; }
    blas_handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!blas_handle) return -1;
#define LOADSYM(var, type, base) \
    strcpy(name, prefix ? prefix : ""); strcat(name, base); var = (type)dlsym(blas_handle, name); if (!var) return -2;
    LOADSYM(p_dgemm, dgemm_t, "dgemm_")
    LOADSYM(p_dtrsm, dtrsm_t, "dtrsm_")
    LOADSYM(p_dgetrf, dgetrf_t, "dgetrf_")
    LOADSYM(p_dgemv, dgemv_t, "dgemv_")
    return 0;
}

/* ------------------------------------------------- generic dense fragments */
/* C := alpha*A*B^T + beta*C  (ggemm! 'n','t' branch, GenericBlasLapackFragments.jl:175-191) */
static void gemm_nt(I m, I n, I k, double alpha, const double* A, I lda, const double* B, I ldb,
                    double beta, double* C, I ldc) {
    if (m <= 0 || n <= 0) return;
    if (p_dgemm) {
        int mi = (int)m, ni = (int)n, ki = (int)k, la = (int)lda, lb = (int)ldb, lc = (int)ldc;
        if (la < 1) la = 1; if (lb < 1) lb = 1;
        p_dgemm("N", "T", &mi, &ni, &ki, &alpha, A, &la, B, &lb, &beta, C, &lc);
        return;
    }
    for (I j = 0; j < n; ++j) {
        double* c = C + j * ldc;
        if (beta == 0.0) { for (I i = 0; i < m; ++i) c[i] = 0.0; }
        else if (beta != 1.0) { for (I i = 0; i < m; ++i) c[i] = beta * c[i]; }
        for (I l = 0; l < k; ++l) {
            double t = alpha * B[j + l * ldb];
            const double* a = A + l * lda;
            for (I i = 0; i < m; ++i) c[i] += t * a[i];
        }
    }
}

/* LU with partial pivoting, first-max rule (ggetrf!, GenericBlasLapackFragments.jl:56-100).
 * ipiv gets 1-based block-local row numbers.  Returns LAPACK-style info (first zero pivot, 1-based; 0 if none). */
static I getrf(I m, I n, double* A, I lda, I* ipiv) {
    if (p_dgetrf && m > 0 && n > 0) {
        int mi = (int)m, ni = (int)n, la = (int)lda, info = 0;
        int* ip = (int*)malloc(sizeof(int) * (size_t)(m < n ? m : n));
        p_dgetrf(&mi, &ni, A, &la, ip, &info);
        for (I k = 0; k < (m < n ? m : n); ++k) ipiv[k] = ip[k];
        free(ip);
        return info;
    }
    I info = 0, mn = m < n ? m : n;
    for (I k = 0; k < mn; ++k) {
        I kp = k;
        if (k < m - 1) {
            double amax = fabs(A[k + k * lda]);
            for (I i = k + 1; i < m; ++i) {
                double v = fabs(A[i + k * lda]);
                if (v > amax) { kp = i; amax = v; }
            }
        }
        ipiv[k] = kp + 1;
        if (A[kp + k * lda] != 0.0) {
            if (kp != k)
                for (I j = 0; j < n; ++j) { double t = A[k + j * lda]; A[k + j * lda] = A[kp + j * lda]; A[kp + j * lda] = t; }
            double inv = 1.0 / A[k + k * lda];
            for (I i = k + 1; i < m; ++i) A[i + k * lda] *= inv;
        } else if (info == 0) info = k + 1;
        for (I j = k + 1; j < n; ++j) {
            double akj = A[k + j * lda];
            for (I i = k + 1; i < m; ++i) A[i + j * lda] -= A[i + k * lda] * akj;
        }
    }
    return info;
}

/* B := B * inv(U), U upper non-unit  ('r','u','n','n'; gtrsm!, ref-BLAS column order) */
static void trsm_runn(I m, I n, const double* A, I lda, double* B, I ldb) {
    if (m <= 0 || n <= 0) return;
    if (p_dtrsm) { int mi = (int)m, ni = (int)n, la = (int)lda, lb = (int)ldb; double one = 1.0;
        p_dtrsm("R", "U", "N", "N", &mi, &ni, &one, A, &la, B, &lb); return; }
    for (I j = 0; j < n; ++j) {
        for (I k = 0; k < j; ++k) {
            double akj = A[k + j * lda];
            if (akj != 0.0) for (I i = 0; i < m; ++i) B[i + j * ldb] -= akj * B[i + k * ldb];
        }
        double t = 1.0 / A[j + j * lda];
        for (I i = 0; i < m; ++i) B[i + j * ldb] = t * B[i + j * ldb];
    }
}

/* B := B * inv(L^T), L unit lower  ('r','l','t','u') */
static void trsm_rltu(I m, I n, const double* A, I lda, double* B, I ldb) {
    if (m <= 0 || n <= 0) return;
    if (p_dtrsm) { int mi = (int)m, ni = (int)n, la = (int)lda, lb = (int)ldb; double one = 1.0;
        p_dtrsm("R", "L", "T", "U", &mi, &ni, &one, A, &la, B, &lb); return; }
    /* X L^T = B  =>  column j of X: x_j = b_j - sum_{k<j} x_k * L[j,k] */
    for (I j = 0; j < n; ++j)
        for (I k = 0; k < j; ++k) {
            double ljk = A[j + k * lda];
            if (ljk != 0.0) for (I i = 0; i < m; ++i) B[i + j * ldb] -= ljk * B[i + k * ldb];
        }
}

/* x := inv(L) x, L unit lower ('l','l','n','u', one rhs) */
static void trsv_lnu(I n, const double* A, I lda, double* x) {
    if (p_dtrsm && n > 0) { int ni = (int)n, one_i = 1, la = (int)lda; double one = 1.0;
        p_dtrsm("L", "L", "N", "U", &ni, &one_i, &one, A, &la, x, &ni); return; }
    for (I k = 0; k < n; ++k) {
        double xk = x[k];
        if (xk != 0.0) for (I i = k + 1; i < n; ++i) x[i] -= xk * A[i + k * lda];
    }
}

/* x := inv(U) x, U upper non-unit ('l','u','n','n', one rhs) */
static void trsv_unn(I n, const double* A, I lda, double* x) {
    if (p_dtrsm && n > 0) { int ni = (int)n, one_i = 1, la = (int)lda; double one = 1.0;
        p_dtrsm("L", "U", "N", "N", &ni, &one_i, &one, A, &la, x, &ni); return; }
    for (I k = n - 1; k >= 0; --k) {
        if (x[k] != 0.0) {
            x[k] /= A[k + k * lda];
            double xk = x[k];
            for (I i = 0; i < k; ++i) x[i] -= xk * A[i + k * lda];
        }
    }
}

/* ------------------------------------------------------------ index helpers */
/* _ldindx!: map[row] = distance of row from the bottom of J's list (SpkSpdMMOps.jl:70-79) */
static void ldindx(I jlen, const I* list, I* map1) {
    I kk = jlen - 1;
    for (I t = 0; t < jlen; ++t) map1[list[t]] = kk--;
}

/* _mmpyi!: rank-1 indexed update z[col k] -= (y[k]*diag) * x  (SpkSpdMMOps.jl:125-143).
 * iz1 is the 1-based column pointer array (xlnz or xunz), z1 1-based values. */
static void mmpyi(I m, I q, const I* zrows, const I* zcols, const double* x, const double* y,
                  const I* iz1, double* z1, const I* map1, double diag) {
    for (I k = 0; k < q; ++k) {
        double t = y[k] * diag;
        I zlast = iz1[zcols[k] + 1] - 1;
        for (I j = 0; j < m; ++j) z1[zlast - map1[zrows[j]]] -= t * x[j];
    }
}

/* _assmb!: scatter-add a tlen x nq update block (SpkSpdMMOps.jl:41-49).
 * xz1f = column pointer array viewed from fj (1-based: xz1f[1] == xlnz[fj]). */
static void assmb(I tlen, I nq, const double* temp, const I* relcol, const I* relind,
                  const I* xz1f, double* z1, I jlen) {
    for (I j = 0; j < nq; ++j) {
        I lbot = xz1f[jlen - relcol[j] + 1] - 1;
        for (I k = 0; k < tlen; ++k) z1[lbot - relind[k]] += temp[j * tlen + k];
    }
}

/* ================================================================ LU factor */
API I spko_lufactor(I n, I nsuper, const I* xsuper0, const I* snode0, const I* xlindx0, const I* lindx0,
                    const I* xlnz0, double* lnz0, const I* xunz0, double* unz0, I* ipvt0) {
    const I *xsuper = xsuper0 - 1, *snode = snode0 - 1, *xlindx = xlindx0 - 1, *lindx = lindx0 - 1;
    const I *xlnz = xlnz0 - 1, *xunz = xunz0 - 1;
    double *lnz = lnz0 - 1, *unz = unz0 - 1;
    I* ipvt = ipvt0 - 1;
    I iflag = 0, tmpsiz = 0;
    I* link = (I*)calloc((size_t)nsuper + 1, sizeof(I));
    I* lngth = (I*)calloc((size_t)nsuper + 1, sizeof(I));
    for (I i = 1; i <= nsuper; ++i) {
        lngth[i] = xlindx[i + 1] - xlindx[i];
        I need = (xsuper[i + 1] - xsuper[i]) * (xlnz[xsuper[i] + 1] - xlnz[xsuper[i]]);
        if (need > tmpsiz) tmpsiz = need;
    }
    I* map = (I*)calloc((size_t)n + 1, sizeof(I));
    I* relind = (I*)calloc((size_t)n + 1, sizeof(I));
    double* temp = (double*)calloc((size_t)tmpsiz + 1, sizeof(double));

    for (I jsup = 1; jsup <= nsuper; ++jsup) {
        I fj = xsuper[jsup], lj = xsuper[jsup + 1] - 1, nj = lj - fj + 1;
        I jlen = xlnz[fj + 1] - xlnz[fj];
        I jxpnt = xlindx[jsup], jlpnt = xlnz[fj], jupnt = xunz[fj];
        ldindx(jlen, &lindx[jxpnt], map);
        for (;;) {
            I ksup = link[jsup];
            if (ksup == 0) break;
            link[jsup] = link[ksup]; link[ksup] = 0;
            I fk = xsuper[ksup], nk = xsuper[ksup + 1] - fk;
            I ksuplen = xlnz[fk + 1] - xlnz[fk];
            I klen = lngth[ksup];
            I kxpnt = xlindx[ksup + 1] - klen;
            I klpnt = xlnz[fk + 1] - klen, kupnt = xunz[fk + 1] - klen;
            I nups, nxt = 0;
            if (klen == jlen) {
                /* dense, same structure (SpkLUFactor.jl:150-159) */
                gemm_nt(jlen, nj, nk, -1.0, &lnz[klpnt], ksuplen, &unz[kupnt], ksuplen - nk, 1.0, &lnz[jlpnt], jlen);
                if (jlen > nj)
                    gemm_nt(jlen - nj, nj, nk, -1.0, &unz[kupnt + nj], ksuplen - nk, &lnz[klpnt], ksuplen, 1.0,
                            &unz[jupnt], jlen - nj);
                nups = nj;
                if (klen > nj) nxt = lindx[jxpnt + nj];
            } else {
                nups = klen;
                for (I i = 0; i < klen; ++i) {
                    nxt = lindx[kxpnt + i];
                    if (nxt > lj) { nups = i; break; }
                }
                if (nk == 1) {
                    /* rank-1 indexed updates (SpkLUFactor.jl:171-176) */
                    mmpyi(klen, nups, &lindx[kxpnt], &lindx[kxpnt], &lnz[klpnt], &unz[kupnt], xlnz, lnz, map, 1.0);
                    mmpyi(klen - nups, nups, &lindx[kxpnt + nups], &lindx[kxpnt], &unz[kupnt + nups], &lnz[klpnt],
                          xunz, unz, map, 1.0);
                } else {
                    I kfirst = lindx[kxpnt], klast = lindx[kxpnt + klen - 1];
                    I inddif = map[kfirst] - map[klast];
                    if (inddif < klen) {
                        /* dense contiguous (SpkLUFactor.jl:186-194) */
                        I ilpnt = xlnz[kfirst] + (kfirst - fj);
                        gemm_nt(klen, nups, nk, -1.0, &lnz[klpnt], ksuplen, &unz[kupnt], ksuplen - nk, 1.0, &lnz[ilpnt], jlen);
                        I iupnt = xunz[kfirst];
                        if (klen > nups)
                            gemm_nt(klen - nups, nups, nk, -1.0, &unz[kupnt + nups], ksuplen - nk, &lnz[klpnt], ksuplen,
                                    1.0, &unz[iupnt], jlen - nj);
                    } else {
                        /* general sparse: product into temp, then scatter-add (SpkLUFactor.jl:195-214) */
                        if (klen * nups > tmpsiz) iflag = -2;
                        for (I t = 0; t < klen; ++t) relind[1 + t] = map[lindx[kxpnt + t]];   /* _igathr! */
                        gemm_nt(klen, nups, nk, -1.0, &lnz[klpnt], ksuplen, &unz[kupnt], ksuplen - nk, 0.0, temp, klen);
                        assmb(klen, nups, temp, &relind[1], &relind[1], &xlnz[fj - 1], lnz, jlen);
                        if (klen > nups) {
                            gemm_nt(klen - nups, nups, nk, -1.0, &unz[kupnt + nups], ksuplen - nk, &lnz[klpnt], ksuplen,
                                    0.0, temp, klen - nups);
                            assmb(klen - nups, nups, temp, &relind[1], &relind[1 + nups], &xunz[fj - 1], unz, jlen);
                        }
                    }
                }
            }
            if (klen > nups) {
                I nxtsup = snode[nxt];
                link[ksup] = link[nxtsup]; link[nxtsup] = ksup;
                lngth[ksup] = klen - nups;
            } else lngth[ksup] = 0;
        }
        /* diagonal block + panels (SpkLUFactor.jl:230-241) */
        I info = getrf(nj, nj, &lnz[jlpnt], jlen, &ipvt[fj]);
        iflag = (info != 0) ? -1 : 0;        /* reference quirk: overwritten per supernode (:230-233) */
        trsm_runn(jlen - nj, nj, &lnz[jlpnt], jlen, &lnz[jlpnt + nj], jlen);
        if (jlen > nj) {
            I m = jlen - nj;                  /* _luswap!: column swaps of the U^T block, k = 1..nj */
            double* a = &unz[jupnt];
            for (I k = 1; k <= nj; ++k) {
                I i = ipvt[fj + k - 1];
                if (i != k) for (I r = 0; r < m; ++r) { double t = a[(i - 1) * m + r]; a[(i - 1) * m + r] = a[(k - 1) * m + r]; a[(k - 1) * m + r] = t; }
            }
            trsm_rltu(m, nj, &lnz[jlpnt], jlen, a, m);
            I nx = lindx[jxpnt + nj], nxtsup = snode[nx];
            link[jsup] = link[nxtsup]; link[nxtsup] = jsup;
            lngth[jsup] = m;
        } else lngth[jsup] = 0;
    }
    free(link); free(lngth); free(map); free(relind); free(temp);
    return iflag;
}

/* ============================================================== LU solves */
API I spko_lulsolve(I nsuper, const I* xsuper0, const I* xlindx0, const I* lindx0, const I* xlnz0,
                    const double* lnz0, const I* ipiv0, double* rhs0) {
    const I *xsuper = xsuper0 - 1, *xlindx = xlindx0 - 1, *lindx = lindx0 - 1, *xlnz = xlnz0 - 1, *ipiv = ipiv0 - 1;
    const double* lnz = lnz0 - 1; double* rhs = rhs0 - 1;
    if (nsuper <= 0) return 0;
    for (I jsup = 1; jsup <= nsuper; ++jsup) {
        I fj = xsuper[jsup], nj = xsuper[jsup + 1] - fj;
        I jlen = xlnz[fj + 1] - xlnz[fj], jxpnt = xlindx[jsup], jlpnt = xlnz[fj];
        for (I k = 1; k <= nj; ++k) {          /* laswp on the rhs block */
            I ip = ipiv[fj + k - 1];
            if (ip != k) { double t = rhs[fj + k - 1]; rhs[fj + k - 1] = rhs[fj + ip - 1]; rhs[fj + ip - 1] = t; }
        }
        trsv_lnu(nj, &lnz[jlpnt], jlen, &rhs[fj]);
        /* temp = -L21 * rhs_J (ggemv! 'n': per row the sum runs over columns in ascending
         * order), then scatter-add into the rows below (SpkLUFactor.jl:312-320) */
        for (I r = 0; r < jlen - nj; ++r) {
            double t = 0.0;
            for (I c = 0; c < nj; ++c) t += (-rhs[fj + c]) * lnz[jlpnt + nj + r + c * jlen];
            rhs[lindx[jxpnt + nj + r]] += t;
        }
    }
    return 1;
}

API I spko_luusolve(I n, I nsuper, const I* xsuper0, const I* xlindx0, const I* lindx0, const I* xlnz0,
                    const double* lnz0, const I* xunz0, const double* unz0, double* rhs0) {
    const I *xsuper = xsuper0 - 1, *xlindx = xlindx0 - 1, *lindx = lindx0 - 1, *xlnz = xlnz0 - 1, *xunz = xunz0 - 1;
    const double *lnz = lnz0 - 1, *unz = unz0 - 1; double* rhs = rhs0 - 1;
    (void)n;
    if (nsuper <= 0) return 0;
    for (I jsup = nsuper; jsup >= 1; --jsup) {
        I fj = xsuper[jsup], nj = xsuper[jsup + 1] - fj;
        I jlen = xlnz[fj + 1] - xlnz[fj], jxpnt = xlindx[jsup], jlpnt = xlnz[fj], jupnt = xunz[fj];
        I m = jlen - nj;
        for (I c = 0; c < nj; ++c) {           /* rhs_J -= (U12^T)^T * rhs[below]  (ggemv! 't') */
            double t = 0.0;
            for (I r = 0; r < m; ++r) t += unz[jupnt + r + c * m] * rhs[lindx[jxpnt + nj + r]];
            rhs[fj + c] += -t;
        }
        trsv_unn(nj, &lnz[jlpnt], jlen, &rhs[fj]);
    }
    return 1;
}

/* ========================================================= LDL^T (intended) */
/* In-block LDL^T + panel: the INTENDED _pchole! (SpkLDLtFactor.jl:347-378 with the division
 * hoisted out of the i-loop).  Returns number of zero pivots met. */
static I pchole(double* A, I nj, I lda) {
    I nzero = 0;
    for (I c = 0; c < nj; ++c) {
        for (I i = 0; i < c; ++i) {
            double f = A[c + i * lda] * A[i + i * lda];
            for (I r = c; r < nj; ++r) A[r + c * lda] -= f * A[r + i * lda];
        }
        double d = A[c + c * lda];
        if (d == 0.0) ++nzero;
        for (I r = c + 1; r < nj; ++r) A[r + c * lda] /= d;
    }
    trsm_rltu(lda - nj, nj, A, lda, A + nj, lda);
    for (I c = 0; c < nj; ++c) {
        double d = A[c + c * lda];
        for (I r = nj; r < lda; ++r) A[r + c * lda] /= d;
    }
    return nzero;
}

API I spko_ldltfactor(I n, I nsuper, const I* xsuper0, const I* snode0, const I* xlindx0, const I* lindx0,
                      const I* xlnz0, double* lnz0) {
    const I *xsuper = xsuper0 - 1, *snode = snode0 - 1, *xlindx = xlindx0 - 1, *lindx = lindx0 - 1, *xlnz = xlnz0 - 1;
    double* lnz = lnz0 - 1;
    I iflag = 0, tmpsiz = 0, maxwidth = 0, nzero = 0;
    I* link = (I*)calloc((size_t)nsuper + 1, sizeof(I));
    I* lngth = (I*)calloc((size_t)nsuper + 1, sizeof(I));
    for (I i = 1; i <= nsuper; ++i) {
        lngth[i] = xlindx[i + 1] - xlindx[i];
        I width = xsuper[i + 1] - xsuper[i];
        I need = width * (xlnz[xsuper[i] + 1] - xlnz[xsuper[i]]);
        if (need > tmpsiz) tmpsiz = need;
        if (width > maxwidth) maxwidth = width;
    }
    I* map = (I*)calloc((size_t)n + 1, sizeof(I));
    I* relind = (I*)calloc((size_t)n + 1, sizeof(I));
    double* diag = (double*)calloc((size_t)maxwidth + 1, sizeof(double));
    double* temp = (double*)calloc((size_t)tmpsiz + 1, sizeof(double));
    double* temp2 = (double*)calloc((size_t)(maxwidth * maxwidth) + 1, sizeof(double));

    for (I jsup = 1; jsup <= nsuper; ++jsup) {
        I fj = xsuper[jsup], lj = xsuper[jsup + 1] - 1, nj = lj - fj + 1;
        I jlen = xlnz[fj + 1] - xlnz[fj], jxpnt = xlindx[jsup], jlpnt = xlnz[fj];
        ldindx(jlen, &lindx[jxpnt], map);
        for (;;) {
            I ksup = link[jsup];
            if (ksup == 0) break;
            link[jsup] = link[ksup]; link[ksup] = 0;
            I fk = xsuper[ksup], nk = xsuper[ksup + 1] - fk;
            I ksuplen = xlnz[fk + 1] - xlnz[fk];
            I klen = lngth[ksup];
            I kxpnt = xlindx[ksup + 1] - klen, klpnt = xlnz[fk + 1] - klen;
            for (I c = 0; c < nk; ++c) diag[c] = lnz[xlnz[fk] + c + c * ksuplen];      /* _loaddiag! */
            I nups, nxt = 0;
            if (klen == jlen) {
                nups = nj;
                /* intended _matrixdiagmm!: temp2[r,c] = D_c * L_K^act[r,c], leading dim ksuplen (row S4) */
                for (I c = 0; c < nk; ++c) for (I r = 0; r < nups; ++r) temp2[r + c * nups] = diag[c] * lnz[klpnt + r + c * ksuplen];
                gemm_nt(jlen, nj, nk, -1.0, &lnz[klpnt], ksuplen, temp2, nups, 1.0, &lnz[jlpnt], jlen);
                if (klen > nj) nxt = lindx[jxpnt + nj];
            } else {
                nups = klen;
                for (I i = 0; i < klen; ++i) {
                    nxt = lindx[kxpnt + i];
                    if (nxt > lj) { nups = i; break; }
                }
                if (nk == 1) {
                    mmpyi(klen, nups, &lindx[kxpnt], &lindx[kxpnt], &lnz[klpnt], &lnz[klpnt], xlnz, lnz, map, lnz[xlnz[fk]]);
                } else {
                    I kfirst = lindx[kxpnt], klast = lindx[kxpnt + klen - 1];
                    I inddif = map[kfirst] - map[klast];
                    for (I c = 0; c < nk; ++c) for (I r = 0; r < nups; ++r) temp2[r + c * nups] = diag[c] * lnz[klpnt + r + c * ksuplen];
                    if (inddif < klen) {
                        I ilpnt = xlnz[kfirst] + (kfirst - fj);
                        gemm_nt(klen, nups, nk, -1.0, &lnz[klpnt], ksuplen, temp2, nups, 1.0, &lnz[ilpnt], jlen);
                    } else {
                        if (klen * nups > tmpsiz) iflag = -2;
                        for (I t = 0; t < klen; ++t) relind[1 + t] = map[lindx[kxpnt + t]];
                        gemm_nt(klen, nups, nk, -1.0, &lnz[klpnt], ksuplen, temp2, nups, 0.0, temp, klen);
                        assmb(klen, nups, temp, &relind[1], &relind[1], &xlnz[fj - 1], lnz, jlen);
                    }
                }
            }
            if (klen > nups) {
                I nxtsup = snode[nxt];
                link[ksup] = link[nxtsup]; link[nxtsup] = ksup;
                lngth[ksup] = klen - nups;
            } else lngth[ksup] = 0;
        }
        nzero += pchole(&lnz[jlpnt], nj, jlen);
        if (jlen > nj) {
            I nx = lindx[jxpnt + nj], nxtsup = snode[nx];
            link[jsup] = link[nxtsup]; link[nxtsup] = jsup;
            lngth[jsup] = jlen - nj;
        } else lngth[jsup] = 0;
    }
    free(link); free(lngth); free(map); free(relind); free(diag); free(temp); free(temp2);
    if (iflag == 0 && nzero > 0) iflag = -1;   /* stricter than the reference (never sets -1): documented superset */
    return iflag;
}

/* _ldltsolve! (SpkLDLtFactor.jl:266-293) */
API I spko_ldltsolve(I nsuper, const I* xsuper0, const I* xlindx0, const I* lindx0, const I* xlnz0,
                     const double* lnz0, double* rhs0) {
    const I *xsuper = xsuper0 - 1, *xlindx = xlindx0 - 1, *lindx = lindx0 - 1, *xlnz = xlnz0 - 1;
    const double* lnz = lnz0 - 1; double* rhs = rhs0 - 1;
    for (I jsup = 1; jsup <= nsuper; ++jsup) {
        I fjcol = xsuper[jsup], ljcol = xsuper[jsup + 1] - 1;
        I fsub = xlindx[jsup], lsub = xlindx[jsup + 1] - 1;
        for (I jcol = fjcol; jcol <= ljcol; ++jcol) {
            I ofst = jcol - fjcol;
            I nzf = xlnz[jcol] + 1 + ofst;
            ++fsub;
            double t = rhs[jcol];
            for (I s = fsub, q = nzf; s <= lsub; ++s, ++q) rhs[lindx[s]] -= t * lnz[q];
        }
    }
    for (I jsup = nsuper; jsup >= 1; --jsup) {
        I fjcol = xsuper[jsup], ljcol = xsuper[jsup + 1] - 1;
        I fsub = xlindx[jsup] + ljcol - fjcol;
        for (I jcol = ljcol; jcol >= fjcol; --jcol) {
            I ofst = jcol - fjcol;
            I nzf = xlnz[jcol] + ofst, m = xlnz[jcol + 1] - nzf - 1;
            double t = rhs[jcol] / lnz[nzf];
            for (I k = 1; k <= m; ++k) t -= lnz[nzf + k] * rhs[lindx[fsub + k]];
            rhs[jcol] = t; --fsub;
        }
    }
    return 1;
}

/* Number of cmod(J,K) operations and schedule flops the reference's linked-list schedule
 * executes (structure only) — used for reporting, never as a roofline numerator. */
API I spko_schedule_stats(I n, I nsuper, const I* xsuper0, const I* snode0, const I* xlindx0, const I* lindx0,
                          double* out /* [ncmod, nrank1, flops_lu, flops_spd] */) {
    const I *xsuper = xsuper0 - 1, *snode = snode0 - 1, *xlindx = xlindx0 - 1, *lindx = lindx0 - 1;
    (void)n;
    I* link = (I*)calloc((size_t)nsuper + 1, sizeof(I));
    I* lngth = (I*)calloc((size_t)nsuper + 1, sizeof(I));
    double ncmod = 0, nr1 = 0, flu = 0, fspd = 0;
    for (I i = 1; i <= nsuper; ++i) lngth[i] = xlindx[i + 1] - xlindx[i];
    for (I jsup = 1; jsup <= nsuper; ++jsup) {
        I fj = xsuper[jsup], lj = xsuper[jsup + 1] - 1, nj = lj - fj + 1;
        I jlen = xlindx[jsup + 1] - xlindx[jsup], jxpnt = xlindx[jsup];
        for (;;) {
            I ksup = link[jsup];
            if (ksup == 0) break;
            link[jsup] = link[ksup]; link[ksup] = 0;
            I nk = xsuper[ksup + 1] - xsuper[ksup];
            I klen = lngth[ksup], kxpnt = xlindx[ksup + 1] - klen;
            I nups = klen, nxt = 0;
            for (I i = 0; i < klen; ++i) { nxt = lindx[kxpnt + i]; if (nxt > lj) { nups = i; break; } }
            ncmod += 1; if (nk == 1) nr1 += 1;
            fspd += 2.0 * (double)klen * (double)nups * (double)nk;
            flu += 2.0 * (double)klen * (double)nups * (double)nk + 2.0 * (double)(klen - nups) * (double)nups * (double)nk;
            if (klen > nups) { I ns = snode[nxt]; link[ksup] = link[ns]; link[ns] = ksup; lngth[ksup] = klen - nups; }
            else lngth[ksup] = 0;
        }
        double m = (double)(jlen - nj), w = (double)nj;
        fspd += w * w * w / 3.0 + m * w * w;
        flu += 2.0 * w * w * w / 3.0 + 2.0 * m * w * w;
        if (jlen > nj) { I nx = lindx[jxpnt + nj], ns = snode[nx]; link[jsup] = link[ns]; link[ns] = jsup; lngth[jsup] = jlen - nj; }
        else lngth[jsup] = 0;
    }
    out[0] = ncmod; out[1] = nr1; out[2] = flu; out[3] = fspd;
    free(link); free(lngth);
    return 0;
}
