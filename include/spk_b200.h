/* spk_b200.h — C ABI of the B200-native numeric factor/solve engine for Sparspak.jl's
 * supernodal LU (SparseSolver) and LDL^T (SparseSpdSolver).
 *
 * Every pointer below is HOST memory owned by the caller (in the reference: Julia
 * `Vector`s of the `_SparseBase` / `_SparseSpdBase` structs).  Index arrays are the
 * reference's own 1-based Int64 arrays, passed unmodified.  Nothing is retained past a
 * call except inside an explicit plan handle.  No torch / CUDA types appear here.
 *
 * Array lengths (SpkSparseBase.jl:99-125):
 *   xsuper, xlindx : nsuper+1      snode, ipvt : n      xlnz, xunz : n+1
 *   lindx : xlindx[nsuper]-1 (nsub)  lnz : xlnz[n]-1    unz : xunz[n]-1
 *
 * Return value of the factor entry points = the reference's `iflag`
 * (SpkLUFactor.jl:29-34):  0 ok,  -1 zero pivot (ANY supernode — a documented superset of
 * the reference, which keeps only the last supernode's status, SpkLUFactor.jl:230-233),
 * -2 insufficient workspace (never produced here), and new codes
 * -100 - cudaError for device/runtime failures (see spk_last_error()).  The library never aborts
 * the calling process.
 */
#ifndef SPK_B200_H
#define SPK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SPK_API __attribute__((visibility("default")))
#else
#define SPK_API
#endif

/* ---- stateless drop-ins: one call = upload, compute on the GPU, download -------------
 * The plans behind these calls are CACHED per structure (and the factors stay resident: a solve that presents the
 * arrays the last factor call wrote back does not upload them again); see spk_cache_clear below. */

/* replaces _lufactor!(n,nsuper,xsuper,snode,xlindx,lindx,xlnz,lnz,xunz,unz,ipvt)
 * src/SparseMethod/SpkLUFactor.jl:60-255, called from _factor! SpkSparseBase.jl:384 */
SPK_API int64_t spk_lufactor_f64(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                                 const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz, double* lnz,
                                 const int64_t* xunz, double* unz, int64_t* ipvt);

/* replaces _lulsolve!(nsuper,xsuper,xlindx,lindx,xlnz,lnz,ipiv,rhs)  SpkLUFactor.jl:269-323
 * (called from _triangularsolve! SpkSparseBase.jl:409).  rhs: length n = xsuper[nsuper]-1, permuted
 * order, in place. */
SPK_API int64_t spk_lulsolve_f64(int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx,
                                 const int64_t* lindx, const int64_t* xlnz, const double* lnz,
                                 const int64_t* ipiv, double* rhs);

/* replaces _luusolve!(n,nsuper,xsuper,xlindx,lindx,xlnz,lnz,xunz,unz,rhs)  SpkLUFactor.jl:325-377
 * (SpkSparseBase.jl:411) */
SPK_API int64_t spk_luusolve_f64(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx,
                                 const int64_t* lindx, const int64_t* xlnz, const double* lnz,
                                 const int64_t* xunz, const double* unz, double* rhs);

/* replaces _ldltfactor!(n,nsuper,xsuper,snode,xlindx,lindx,xlnz,lnz)
 * src/SparseSpdMethod/SpkLDLtFactor.jl:58-246, called from _factor! SpkSparseSpdBase.jl:325.
 * Computes the INTENDED LDL^T (the reference code as written is defective: SURVEY.md §8a S3/S4). */
SPK_API int64_t spk_ldltfactor_f64(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                                   const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz, double* lnz);

/* replaces _ldltsolve!(nsuper,xsuper,xlindx,lindx,xlnz,lnz,rhs)  SpkLDLtFactor.jl:266-293
 * (SpkSparseSpdBase.jl:351) */
SPK_API int64_t spk_ldltsolve_f64(int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx,
                                  const int64_t* lindx, const int64_t* xlnz, const double* lnz, double* rhs);

/* Float32 twins of the five drop-ins (the reference dispatches Float32 to sgemm/sgetrf/strsm,
 * SpkSpdMMOps.jl:186-351; _factor!(s::_SparseBase{Int64,Float32}) binds these).  Values are widened on
 * entry and narrowed on exit; the arithmetic is the FP64 engine's. */
SPK_API int64_t spk_lufactor_f32(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                                 const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz, float* lnz,
                                 const int64_t* xunz, float* unz, int64_t* ipvt);
SPK_API int64_t spk_lulsolve_f32(int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx,
                                 const int64_t* lindx, const int64_t* xlnz, const float* lnz,
                                 const int64_t* ipiv, float* rhs);
SPK_API int64_t spk_luusolve_f32(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx,
                                 const int64_t* lindx, const int64_t* xlnz, const float* lnz,
                                 const int64_t* xunz, const float* unz, float* rhs);
SPK_API int64_t spk_ldltfactor_f32(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                                   const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz, float* lnz);
SPK_API int64_t spk_ldltsolve_f32(int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx,
                                  const int64_t* lindx, const int64_t* xlnz, const float* lnz, float* rhs);

/* ---- stateful plan: structure + factors stay resident in HBM --------------------------
 * One plan per symbolic factorisation (rebuild when symbolicfactor! reruns,
 * SpkSparseSolver.jl:163-175).  Re-entrant per handle; every call is synchronous at return. */
typedef struct spk_plan spk_plan;

#define SPK_LU   0   /* xunz != NULL: supernodal LU with in-supernode partial pivoting */
#define SPK_LDLT 1   /* xunz == NULL: supernodal LDL^T */

/* device: CUDA ordinal.  part/nparts: elimination-subtree partition for multi-GPU
 * (part 0 of 1 = whole matrix).  Returns NULL on failure (see spk_last_error()). */
SPK_API spk_plan* spk_plan_create(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                                  const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz,
                                  const int64_t* xunz_or_null, int32_t device, int32_t part, int32_t nparts);
SPK_API void spk_plan_destroy(spk_plan* p);

/* _inmatrix! on the device (SpkSparseBase.jl:302-372 / SpkSparseSpdBase.jl:234-311).  dest[k] names the slot
 * of the reference layout that nzval[k] is added to: lnz[dest[k]-1] (dest>0), unz[-dest[k]-1] (dest<0), none
 * (dest==0); duplicates accumulate.  The values are scattered STRAIGHT INTO THE ZEROED FRONTAL MATRICES the
 * factorisation works on; the device copies of lnz/unz (and spk_plan_get_factors) are only meaningful again
 * after the next spk_plan_factor.  The map is translated and uploaded once (pass NULL afterwards to reuse it;
 * nnz must then equal the nnz the map was built for). */
SPK_API int64_t spk_plan_inmatrix(spk_plan* p, int64_t nnz, const int64_t* dest_or_null, const double* nzval);

/* the same with the values already resident from the last spk_plan_inmatrix (no host traffic; asynchronous) */
SPK_API int64_t spk_plan_reassemble(spk_plan* p);

/* upload assembled lnz/unz (what _inmatrix! produced on the host) */
SPK_API int64_t spk_plan_set_values(spk_plan* p, const double* lnz, const double* unz_or_null);

/* numeric factorisation of the resident values; returns iflag */
SPK_API int64_t spk_plan_factor(spk_plan* p);

/* download factors in the reference layout (any pointer may be NULL to skip) */
SPK_API int64_t spk_plan_get_factors(spk_plan* p, double* lnz, double* unz, int64_t* ipvt);
/* upload factors computed elsewhere (stateless solve entry points use this) */
SPK_API int64_t spk_plan_set_factors(spk_plan* p, const double* lnz, const double* unz_or_null, const int64_t* ipvt_or_null);

/* forward+backward solve of nrhs right-hand sides held column-major (ld = ldrhs >= n) in the
 * PERMUTED order the reference's numeric routines use; in place.
 * which: 0 = both sweeps, 1 = forward only (_lulsolve!), 2 = backward only (_luusolve!). */
SPK_API int64_t spk_plan_solve(spk_plan* p, double* rhs, int64_t nrhs, int64_t ldrhs, int32_t which);

/* _triangularsolve! (SpkSparseBase.jl:400-416): x := P^T solve(P b) with the 1-based
 * permutations given once (rperm, rinvp: length n); b in original order, in place. */
SPK_API int64_t spk_plan_set_perm(spk_plan* p, const int64_t* rperm, const int64_t* rinvp);
SPK_API int64_t spk_plan_triangularsolve(spk_plan* p, double* b, int64_t nrhs, int64_t ldb);

/* ---- device-resident variants used by bench.py / multi-GPU drivers -------------------- */
/* Residual and iterative refinement on the device (computeresidual SpkProblem.jl:448-496, the
 * refinement the reference only keeps as commented-out Fortran, SpkSparseSpdSolver.jl:267-459).
 * spk_plan_set_matrix: A as SparseMatrixCSC (colptr[n+1], rowval[nnz], nzval[nnz], 1-based, ORIGINAL ordering).
 * spk_plan_residual:   res = b - A x (res may be NULL), relnorm[q] = ||res_q||_2 / ||b_q||_2.
 * spk_plan_refine:     x += inv(A)(b - A x) with the resident factors until max relnorm <= tol or maxit
 *                      corrections; returns the number of corrections (>= 0) or an error code (<= -100). */
SPK_API int64_t spk_plan_set_matrix(spk_plan* p, int64_t nnz, const int64_t* colptr, const int64_t* rowval, const double* nzval);
SPK_API int64_t spk_plan_residual(spk_plan* p, const double* b, const double* x, int64_t nrhs, int64_t ld, double* res_or_null, double* relnorm);
SPK_API int64_t spk_plan_refine(spk_plan* p, const double* b, double* x, int64_t nrhs, int64_t ld, int32_t maxit, double tol, double* relnorm);

/* 1-norm condition estimate cond_1(A) ~ ||A||_1 * est(||inv(A)||_1) with the resident factors (needs
 * spk_plan_set_matrix, spk_plan_set_perm and a factorisation).  LDL^T plans: Hager / Higham (LAPACK xLACON);
 * LU plans: a probing LOWER BOUND (no transposed sweeps in the engine), flagged in info2[1].
 * out2 = {||A||_1, estimate of ||inv(A)||_1} (may be NULL), info2 = {solves used, 1 if lower bound only}. */
SPK_API double spk_plan_condest(spk_plan* p, double* out2, int32_t* info2);

/* Float32 callers of a plan (the plan itself stays FP64) */
SPK_API int64_t spk_plan_inmatrix_f32(spk_plan* p, int64_t nnz, const int64_t* dest_or_null, const float* nzval);
SPK_API int64_t spk_plan_get_factors_f32(spk_plan* p, float* lnz, float* unz, int64_t* ipvt);
SPK_API int64_t spk_plan_triangularsolve_f32(spk_plan* p, float* b, int64_t nrhs, int64_t ldb);

SPK_API void*   spk_plan_device_ptr(spk_plan* p, int32_t what);   /* 0 lnz, 1 unz, 2 ipiv(int32), 5 frontal arena, 6 solve work vectors */
SPK_API int64_t spk_plan_device_len(spk_plan* p, int32_t what);   /* element counts of the above */
SPK_API int64_t spk_plan_factor_phase(spk_plan* p, int32_t phase); /* multi-GPU: 0 local subtrees, 1 top set + write-back */
SPK_API int64_t spk_plan_solve_phase(spk_plan* p, double* d_rhs, int64_t nrhs, int64_t ldrhs, int32_t phase); /* 0 fwd local, 1 top, 2 bwd local */
/* exchange lists of a multi-part plan (count when out == NULL).  what 0: subtree-root fronts
 * {owner, arena offset, arena length, w offset, w length, front}; what 1: owned storage ranges
 * {owner, lnz offset, lnz length, unz offset, unz length, first column, #columns} */
SPK_API int64_t spk_plan_xchg_info(spk_plan* p, int32_t what, int64_t i, int64_t* out);
SPK_API int64_t spk_plan_solve_device(spk_plan* p, double* d_rhs, int64_t nrhs, int64_t ldrhs, int32_t which);

/* ---- multi-GPU: one plan per GPU (part r of nparts on device d_r), elimination subtrees dealt to
 * the parts, NCCL over NVLink / NVSwitch for the exchanges.  The library owns the communicator:
 *   one process per GPU:  id from ONE call of spk_nccl_unique_id (128 bytes), carried to every process by the
 *                         host application (MPI / torch.distributed / a file), then spk_plan_comm_init on each.
 *   one process, N GPUs:  spk_multi_create below.
 * spk_plan_factor_multi = own subtrees -> exchange of the subtree roots' update matrices -> top set (LDL^T:
 * distributed by column blocks with one panel broadcast per outer block; LU: replicated) -> write-back, enqueued
 * in one go.  spk_plan_solve_multi: d_rhs on the device in permuted order, full solution on every part at return. */
SPK_API int64_t spk_nccl_unique_id(void* out128);
SPK_API int64_t spk_plan_comm_init(spk_plan* p, const void* id128);
SPK_API int64_t spk_plan_factor_multi(spk_plan* p);
SPK_API int64_t spk_plan_solve_multi(spk_plan* p, double* d_rhs, int64_t nrhs, int64_t ldrhs);

/* One process, N GPUs (devices 0..ngpus-1): a single ccall-able handle that owns the per-GPU plans and the NCCL
 * communicators; every call fans out to one host thread per GPU and joins before returning.  Same semantics as
 * the spk_plan_* calls of the same name (host pointers; lnz/unz/ipvt/b overwritten in place). */
typedef struct spk_multi spk_multi;
SPK_API spk_multi* spk_multi_create(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                                    const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz,
                                    const int64_t* xunz_or_null, int32_t ngpus);
SPK_API void      spk_multi_destroy(spk_multi* m);
SPK_API spk_plan* spk_multi_plan(spk_multi* m, int32_t r);          /* part r's plan (introspection) */
SPK_API int64_t   spk_multi_inmatrix(spk_multi* m, int64_t nnz, const int64_t* dest_or_null, const double* nzval);
SPK_API int64_t   spk_multi_set_values(spk_multi* m, const double* lnz, const double* unz_or_null);
SPK_API int64_t   spk_multi_factor(spk_multi* m);
SPK_API int64_t   spk_multi_get_factors(spk_multi* m, double* lnz, double* unz, int64_t* ipvt);
SPK_API int64_t   spk_multi_set_perm(spk_multi* m, const int64_t* rperm, const int64_t* rinvp);
SPK_API int64_t   spk_multi_triangularsolve(spk_multi* m, double* b, int64_t nrhs, int64_t ldb);

/* ---- introspection --------------------------------------------------------------------- */
/* what: 0 kernel launches of the last factor, 1 of the last solve, 2 #fronts, 3 #levels,
 *       4 device bytes held, 5 #big fronts, 6 update-matrix (S) doubles */
SPK_API int64_t spk_plan_stat(spk_plan* p, int32_t what);
/* what: 0 structural factor flops (sum cc^2 for LDLT, 2 sum cc^2 - sum cc for LU), 1 nnz(L)=sum cc,
 *       2 ms of the last factor (CUDA events), 3 ms of the last solve,
 *       4 flops executed by the DMMA trailing-update kernels in the last factor, 5 ms spent in them
 *       (5 and 10+kind / 30+kind = ms / launches per kernel kind need profiling mode: spk_plan_stat(p, 100)) */
SPK_API double  spk_plan_statf(spk_plan* p, int32_t what);
/* frees the plans cached behind the stateless drop-ins (see spk_b200.cu: plan cache; SPK_PLAN_CACHE=<k>, 0 = off) */
SPK_API void spk_cache_clear(void);
SPK_API const char* spk_last_error(void);
SPK_API int32_t spk_device_count(void);
SPK_API const char* spk_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SPK_B200_H */
