// kernels.cuh — sm_100a kernels of the multifrontal supernodal LU / LDL^T engine.
//
// All kernels are "task-list" kernels: one launch executes a list of independent tasks
// built at plan time (plan.hpp); a thread block finds its task by binary search in a
// block-prefix array.  Dense work happens in frontal matrices held in a device arena;
// the reference layout (lnz / unz, SpkSparseBase.jl:1-87) is gathered from / scattered to
// by k_load_chunks / k_store_chunks, and read directly by the triangular solves.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "plan.hpp"

namespace spk {

struct DFront {
    int64_t fofs, relofs, wofs, F0, pbofs;
    int32_t W, R, m, ld, parent, child0, nchild, c0, nch, ps0, nps;
    int32_t ownofs;                  // distributed top-set front: offset of its column-owner array in DevCtx::fown, else -1
};
struct DChunk {
    int64_t lofs, uofs, posofs, fofs;
    int32_t nj, jlen, o, ld;
};

struct DevCtx {
    double* F;                       // frontal-matrix arena
    double* lnz; double* unz; double* w; double* pb;   // pb: partial sums of the backward sweep
    int32_t* ipiv; int32_t* iflag;
    const DFront* fronts; const DChunk* chunks; const PStep* psteps; const int32_t* subw;
    const int32_t* childlist; const int32_t* rel; const int32_t* pos;
    const SolveTask* solvet;
    int64_t wlen, pblen;
    int32_t lu;
    int32_t me;                      // this part (multi-GPU)
    const int8_t* fown;              // column owners of the distributed top-set fronts
    const double* tinvf; const double* tinvb;   // explicit inverses of the panel steps' diagonal blocks (solve), or NULL
};

// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start
// while its predecessor in the stream is still running.  pdl_trigger() lets the NEXT kernel start launching;
// pdl_wait() blocks until the PREVIOUS kernel has completed and its writes are visible.  Everything that only
// depends on the factors (task records, the diagonal block, the panel rows) is fetched before pdl_wait(), so
// the launch gap and those round trips disappear from the chain of dependent panel steps.  Both are no-ops in
// a kernel launched the ordinary way.  Every block calls pdl_wait() before it exits, which keeps completion
// transitive along the chain.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ int find_task(const int32_t* __restrict__ pfx, int count, int b) {
    int lo = 0, hi = count;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (pfx[mid] <= b) lo = mid; else hi = mid; }
    return lo;
}

// ------------------------------------------------------------------------------------
// inmatrix on the device (SpkSparseBase.jl:302-372): scatter A's values
// straight into the frontal matrices through a destination map built once per pattern.
// Duplicate (i,j) entries of a CSC accumulate, as in the reference's `lnz[..] += nzval[k]`: atomicAdd (native
// FP64 red on sm_100a; with unique destinations — every sorted, summed CSC — the result is order-independent).
__global__ void k_scatter_values(int64_t nnz, const int64_t* __restrict__ dest, const double* __restrict__ v,
                                 double* __restrict__ F) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    int64_t d = dest[k];
    if (d >= 0) atomicAdd(F + d, v[k]);
}

// reference layout -> fronts (values assembled by the host _inmatrix!) and back (factors)
constexpr int CHUNK_EPB = 1024;      // entries per block
// FILLU (load of FACTORS into LDL^T fronts): also write U = D L^T into the upper triangle, which is where the
// factorisation leaves it and where the backward sweep reads it.
template <bool STORE, bool FILLU>
__device__ __forceinline__ void chunk_io(const DevCtx& c, const DChunk& ch, int lb) {
    if (ch.fofs < 0) return;                          // a front this part holds no storage for (multi-GPU)
    const int32_t* __restrict__ pos = c.pos + ch.posofs;
    double* __restrict__ F = c.F + ch.fofs;
    const int64_t nl = (int64_t)ch.jlen * ch.nj;
    const int64_t nu = c.lu ? (int64_t)(ch.jlen - ch.nj) * ch.nj : 0;
    const int ldu = ch.jlen - ch.nj;
    for (int it = 0; it < CHUNK_EPB / 256; ++it) {
        int64_t e = (int64_t)lb * CHUNK_EPB + it * 256 + threadIdx.x;
        if (e < nl) {
            int j = (int)(e / ch.jlen), i = (int)(e - (int64_t)j * ch.jlen);
            int64_t f = (int64_t)pos[i] + (int64_t)(ch.o + j) * ch.ld;
            if (STORE) c.lnz[ch.lofs + e] = F[f]; else F[f] = c.lnz[ch.lofs + e];
            if (FILLU && !STORE && i > j) F[(int64_t)(ch.o + j) + (int64_t)pos[i] * ch.ld] = c.lnz[ch.lofs + (int64_t)j * ch.jlen + j] * c.lnz[ch.lofs + e];
        } else if (e - nl < nu) {
            int64_t eu = e - nl;
            int j = (int)(eu / ldu), i = (int)(eu - (int64_t)j * ldu);
            int64_t f = (int64_t)(ch.o + j) + (int64_t)pos[ch.nj + i] * ch.ld;
            if (STORE) c.unz[ch.uofs + eu] = F[f]; else F[f] = c.unz[ch.uofs + eu];
        }
    }
}
template <bool STORE, bool FILLU = false>
__global__ void __launch_bounds__(256) k_chunks(DevCtx c, const int32_t* __restrict__ pfx, int count) {
    int t = find_task(pfx, count, blockIdx.x);
    chunk_io<STORE, FILLU>(c, c.chunks[t], blockIdx.x - pfx[t]);
}
// the chunks of ONE LEVEL of the front tree (list of chunk ids): the factors of a level are written back to
// lnz / unz on a third stream while the next levels are being factored (HBM-bound copy under tensor-bound GEMMs)
__global__ void __launch_bounds__(256) k_chunks_store_list(DevCtx c, const int32_t* __restrict__ list,
                                                           const int32_t* __restrict__ pfx, int count) {
    int t = find_task(pfx, count, blockIdx.x);
    chunk_io<true, false>(c, c.chunks[list[t]], blockIdx.x - pfx[t]);
}

// ------------------------------------------------------------------------------------
// Extend-add of one child's update matrix into its parent front (the reference's
// _assmb!/_mmpyi! scatter through relative indices, SpkSpdMMOps.jl:41-49,125-143,
// re-expressed child -> parent).  Entry (i,j) of S_child goes to exactly one place.
__device__ __forceinline__ void asm_entry(const DevCtx& c, const DFront& C, const DFront& P,
                                          const int32_t* __restrict__ rel, int64_t e) {
    const int32_t m = C.m;
    const int32_t j = (int32_t)(e / m), i = (int32_t)(e - (int64_t)j * m);
    if (!c.lu && i < j) return;                       // LDL^T: lower triangle only
    if (P.ownofs >= 0 && c.fown[P.ownofs + rel[j]] != c.me) return;   // distributed parent: only the columns this part owns
    const double v = c.F[C.fofs + (int64_t)(C.W + i) + (int64_t)(C.W + j) * C.ld];
    c.F[P.fofs + (int64_t)rel[i] + (int64_t)rel[j] * P.ld] += v;
}

// One block = ASM_TPB rows x ASM_COLS columns of the child's update matrix.  A thread owns one row: its ASM_COLS
// child entries and the ASM_COLS parent entries they go to are all loaded before the first store (no index
// division, 2 * ASM_COLS independent loads in flight per thread); rows are contiguous in the child and run-wise
// contiguous in the parent.  LDL^T: tiles strictly above the diagonal exit at once.
// FILTER (distributed top-set parents): only the parent columns this part owns.
template <bool FILTER>
__global__ void __launch_bounds__(ASM_TPB) k_assemble(DevCtx c, const AsmTask* __restrict__ tasks,
                                                     const int32_t* __restrict__ pfx, int count) {
    int t = find_task(pfx, count, blockIdx.x);
    int lb = blockIdx.x - pfx[t];
    AsmTask a = tasks[t];
    const DFront C = c.fronts[a.child], P = c.fronts[a.parent];
    const int32_t* __restrict__ rel = c.rel + C.relofs;
    const int32_t m = C.m;
    const int32_t nrt = (m + ASM_TPB - 1) / ASM_TPB;
    const int32_t rt = lb % nrt, ct = lb / nrt;
    const int32_t i = rt * ASM_TPB + threadIdx.x, j0 = ct * ASM_COLS;
    if (!c.lu && rt * ASM_TPB + ASM_TPB - 1 < j0) return;           // whole tile above the diagonal
    if (i >= m) return;
    const double* __restrict__ src = c.F + C.fofs + (int64_t)(C.W + i) + (int64_t)(C.W + j0) * C.ld;
    double* __restrict__ dst = c.F + P.fofs + (int64_t)rel[i];
    double v[ASM_COLS], old[ASM_COLS];
    int64_t pofs[ASM_COLS];
#pragma unroll
    for (int cc = 0; cc < ASM_COLS; ++cc) {
        const int32_t j = j0 + cc;
        bool on = j < m && (c.lu || i >= j);
        if (FILTER) { if (on && P.ownofs >= 0) on = c.fown[P.ownofs + rel[j]] == c.me; }
        pofs[cc] = on ? (int64_t)rel[j] * P.ld : -1;
        v[cc] = on ? __ldcs(src + (int64_t)cc * C.ld) : 0.0;
    }
#pragma unroll
    for (int cc = 0; cc < ASM_COLS; ++cc) old[cc] = pofs[cc] >= 0 ? dst[pofs[cc]] : 0.0;
#pragma unroll
    for (int cc = 0; cc < ASM_COLS; ++cc) if (pofs[cc] >= 0) dst[pofs[cc]] = old[cc] + v[cc];
}

// parents with more than ASM_ROUNDS children: one block walks the remaining children in order
__global__ void __launch_bounds__(256) k_assemble_tail(DevCtx c, const AsmTask* __restrict__ tasks, int count) {
    AsmTask a = tasks[blockIdx.x];
    const DFront P = c.fronts[a.parent];
    for (int r = ASM_ROUNDS; r < P.nchild; ++r) {
        const DFront C = c.fronts[c.childlist[P.child0 + r]];
        const int32_t* rel = c.rel + C.relofs;
        int64_t total = (int64_t)C.m * C.m;
        for (int64_t e = threadIdx.x; e < total; e += blockDim.x) asm_entry(c, C, P, rel, e);
        __syncthreads();
    }
}

// Distributed top set: U[ob0 + k, c] = D_k * L[c, ob0 + k] (k < e - ob0, e <= c < R) rebuilt from a panel that
// arrived by broadcast — the upper triangle is where the update kernel reads its B operand and where the
// backward sweep reads U.  32 x 32 tiles transposed through shared memory (coalesced on both sides).
__global__ void __launch_bounds__(256) k_fill_u(DevCtx c, const FillTask* __restrict__ tasks, const int32_t* __restrict__ pfx, int count) {
    __shared__ double tile[32][33];
    const int t = find_task(pfx, count, blockIdx.x);
    const int lb = blockIdx.x - pfx[t];
    const FillTask ft = tasks[t];
    const int nct = (ft.R - ft.e + 31) / 32;
    const int c0 = ft.e + (lb % nct) * 32, k0 = ft.ob0 + (lb / nct) * 32;
    double* __restrict__ Fm = c.F + ft.fofs;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int kk = ty; kk < 32; kk += 8) {                          // L[c0 + tx, k0 + kk] scaled by D
        const int cc = c0 + tx, k = k0 + kk;
        if (cc < ft.R && k < ft.e) tile[kk][tx] = Fm[(int64_t)k + (int64_t)k * ft.ld] * Fm[(int64_t)cc + (int64_t)k * ft.ld];
    }
    __syncthreads();
    for (int cl = ty; cl < 32; cl += 8) {                          // U[k0 + tx, c0 + cl]
        const int cc = c0 + cl, k = k0 + tx;
        if (cc < ft.R && k < ft.e) Fm[(int64_t)k + (int64_t)cc * ft.ld] = tile[tx][cl];
    }
}

// Cooperative copy of a w x w block between global memory (leading dimension ld) and shared memory
// (leading dimension lds), 8 independent loads in flight per thread (a plain load->store loop
// serialises one global round trip per element on the critical path of every panel step).
template <int NT>
__device__ __forceinline__ void block_g2s(double* __restrict__ S, int lds, const double* __restrict__ G, int ld, int w) {
    const int tot = w * w;
    for (int e0 = threadIdx.x; e0 < tot; e0 += 8 * NT) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { int e = e0 + u * NT; int j = e / w, i = e - j * w; v[u] = e < tot ? __ldcg(G + i + (size_t)j * ld) : 0.0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { int e = e0 + u * NT; int j = e / w, i = e - j * w; if (e < tot) S[i + j * lds] = v[u]; }
    }
}

// ------------------------------------------------------------------------------------
// Diagonal block of a panel step (w x w, one or more chunks).
// LU: partial pivoting restricted to each chunk's own rows, first maximum wins (ggetrf!,
// GenericBlasLapackFragments.jl:64-74 == LAPACK idamax); row interchanges are applied from the
// chunk's first column rightwards (never to L entries of earlier chunks), ipiv is chunk-local
// 1-based (SpkLUFactor.jl:230).  LDL^T: the intended _pchole! in-block part
// (SpkLDLtFactor.jl:351-362).
template <bool LU>
__device__ void diag_factor(double* A, int lda, int w, const int32_t* __restrict__ subw, int nsub,
                            int32_t* ipiv, int32_t* iflag, int32_t* s_piv, double* s_pv) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int tx = tid & 15, ty = tid >> 4;               // 16 x 16 layout for the rank-1 updates (blockDim.x == 256)
    if (LU) {
        int s0 = 0;
        for (int b = 0; b < nsub; ++b) {
            const int s1 = s0 + subw[b];
            for (int k = s0; k < s1; ++k) {
                if (tid < 32) {
                    double best = -1.0; int bi = 0x7fffffff;
                    for (int i = k + tid; i < s1; i += 32) {
                        double v = fabs(A[i + (size_t)k * lda]);
                        if (v > best) { best = v; bi = i; }
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        double ob = __shfl_xor_sync(0xffffffffu, best, off);
                        int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                    }
                    if (tid == 0) {
                        if (bi == 0x7fffffff) bi = k;                  // only NaNs: keep the diagonal
                        double pv0 = A[bi + (size_t)k * lda];
                        *s_piv = bi; *s_pv = pv0; ipiv[k] = bi - s0 + 1;
                        if (pv0 == 0.0) atomicExch(iflag, -1);
                    }
                }
                __syncthreads();
                const int kp = *s_piv;
                const double pv = *s_pv;
                if (kp != k && pv != 0.0)
                    for (int j = s0 + tid; j < w; j += nt) {
                        double t = A[k + (size_t)j * lda]; A[k + (size_t)j * lda] = A[kp + (size_t)j * lda]; A[kp + (size_t)j * lda] = t;
                    }
                __syncthreads();
                if (pv != 0.0) {
                    const double inv = 1.0 / A[k + (size_t)k * lda];
                    for (int i = k + 1 + tid; i < w; i += nt) A[i + (size_t)k * lda] *= inv;
                }
                __syncthreads();
                for (int j = k + 1 + ty; j < w; j += 16) {
                    const double akj = A[k + (size_t)j * lda];
                    for (int i = k + 1 + tx; i < w; i += 16) A[i + (size_t)j * lda] -= A[i + (size_t)k * lda] * akj;
                }
                __syncthreads();
            }
            s0 = s1;
        }
    } else {
        for (int k = 0; k < w; ++k) {
            const double d = A[k + (size_t)k * lda];          // final since the previous step's barrier; not written below
            if (tid == 0 && d == 0.0) atomicExch(iflag, -1);
            for (int i = k + 1 + tid; i < w; i += nt) A[i + (size_t)k * lda] /= d;
            __syncthreads();
            for (int s = k + 1 + ty; s < w; s += 16) {
                const double f = A[s + (size_t)k * lda] * d;
                for (int r = s + tx; r < w; r += 16) A[r + (size_t)s * lda] -= f * A[r + (size_t)k * lda];
            }
            __syncthreads();
        }
    }
}

template <bool LU>
__global__ void __launch_bounds__(256) k_diag(DevCtx c, const int32_t* __restrict__ pslist, int smem_w) {
    extern __shared__ double sm[];
    __shared__ int32_t s_piv;
    __shared__ double s_pv;
    const PStep ps = c.psteps[pslist[blockIdx.x]];
    double* G = c.F + ps.fofs + (int64_t)ps.o + (int64_t)ps.o * ps.ld;
    const int w = ps.w, ld = ps.ld;
    int32_t* ipiv = c.ipiv + ps.col0;
    const int32_t* subw = c.subw + ps.sub0;
    if (w <= smem_w) {
        const int lds = w | 1;                          // odd leading dimension: conflict-free columns
        block_g2s<256>(sm, lds, G, ld, w);
        __syncthreads();
        diag_factor<LU>(sm, lds, w, subw, ps.nsub, ipiv, c.iflag, &s_piv, &s_pv);
        for (int e = threadIdx.x; e < w * w; e += blockDim.x) { int j = e / w, i = e % w; G[i + (size_t)j * ld] = sm[i + j * lds]; }
    } else {
        diag_factor<LU>(G, ld, w, subw, ps.nsub, ipiv, c.iflag, &s_piv, &s_pv);
    }
}

// LDL^T of a w x w diagonal block held in REGISTERS: TG x TG threads, thread (tx,ty) owns the entries
// (tx + TG a, ty + TG b), a,b < NB (cyclic, so the shrinking active part stays balanced).  Per column k the
// TG owners of that column (consecutive lanes of one warp) get d_k by a shuffle, publish l_r = a_r / d_k
// (multiplication by the correctly rounded reciprocal) through a double-buffered shared vector; everybody
// then updates its own entries from registers: one barrier per column.  TG = 8 (two warps, 64 entries per
// thread) keeps that barrier cheap: the kernel sits on the critical path of every panel step.
template <int NB, int TG>
__global__ void __launch_bounds__(TG * TG) k_diag_ldlt_reg(DevCtx c, const int32_t* __restrict__ pslist) {
    __shared__ double col[2][TG * NB + 1];
    const PStep ps = c.psteps[pslist[blockIdx.x]];
    double* G = c.F + ps.fofs + (int64_t)ps.o + (int64_t)ps.o * ps.ld;
    const int w = ps.w, ld = ps.ld;
    const int tx = threadIdx.x % TG, ty = threadIdx.x / TG;
    double v[NB][NB];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int a = 0; a < NB; ++a) {
            const int r = tx + TG * a, s = ty + TG * b;
            v[a][b] = (r < w && s < w && r >= s) ? __ldcg(G + r + (size_t)s * ld) : 0.0;
        }
    for (int k = 0; k < w; ++k) {
        double* cb = col[k & 1];
        const int kb = k / TG, kt = k % TG;
        if (ty == kt) {                                     // the TG owners of column k: consecutive lanes of one warp
            double dloc = 0.0;
#pragma unroll
            for (int a = 0; a < NB; ++a) if (a == kb) dloc = v[a][a];          // meaningful in thread tx == kt only
            const int lane0 = (threadIdx.x & 31) - tx;                         // first lane of this owner group
            const unsigned grp = (TG == 32 ? 0xffffffffu : ((1u << TG) - 1u) << lane0);
            const double d0 = __shfl_sync(grp, dloc, lane0 + kt);
            const double rinv = 1.0 / d0;
            if (tx == kt) { cb[TG * NB] = d0; if (d0 == 0.0) atomicExch(c.iflag, -1); }
#pragma unroll
            for (int b = 0; b < NB; ++b) if (b == kb) {
#pragma unroll
                for (int a = 0; a < NB; ++a) {
                    const int r = tx + TG * a;
                    if (r > k && r < w) { v[a][b] *= rinv; cb[r] = v[a][b]; }
                }
            }
        }
        __syncthreads();
        const double d = cb[TG * NB];
        double lr[NB], fs[NB];
#pragma unroll
        for (int a = 0; a < NB; ++a) { const int r = tx + TG * a; lr[a] = (r > k && r < w) ? cb[r] : 0.0; }
#pragma unroll
        for (int b = 0; b < NB; ++b) { const int s = ty + TG * b; fs[b] = (s > k && s < w) ? cb[s] * d : 0.0; }
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int a = 0; a < NB; ++a) {
                const int r = tx + TG * a, s = ty + TG * b;
                if (r >= s && s > k) v[a][b] -= fs[b] * lr[a];
            }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int a = 0; a < NB; ++a) {
            const int r = tx + TG * a, s = ty + TG * b;
            if (r < w && s < w && r >= s) __stcg(G + r + (size_t)s * ld, v[a][b]);
        }
}

// 1/d without the IEEE slow path: hardware seed (>= 20 bits) + two Newton steps (relative error ~1e-16 for normal d).
// The division sits on the dependent chain of every column of a diagonal block.
__device__ __forceinline__ double fast_rcp(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0); x = fma(x, e, x);
    e = fma(-d, x, 1.0); x = fma(x, e, x);
    return x;
}

// LDL^T of a w x w diagonal block (w <= 64) held in registers, one row per DIAG_NS threads (thread NS*r+h owns
// the column pairs P = h (mod NS) of row r).  Per column k the owners publish their still unscaled entries
// a(r,k) (= d_k for r = k, = d_k l_rk below) through a double-buffered shared vector; after ONE barrier every
// thread forms l_rk = a(r,k) / d_k from the published values and updates a(r,j) -= l_rk * (d_k l_jk) for its
// own columns — the entry of column k+1 first, which is published at once; the global stores and the
// zero-pivot flag stay off that dependent chain.  The code is a runtime loop over blocks of 8 columns (the
// register row is shifted down after each block, so every register index is static): a fully unrolled
// version is instruction-fetch bound (ncu: stall_no_instruction 8 cycles/issue), one thread per row is issue
// bound, four threads per row pay for the wider barrier (tools/ubench_steps.cu: 17.9k / 25.4k / 27.6k clocks
// for NS = 2 / 4 / 1 at w = 57) — the kernel runs once, on one SM, on the critical path of every panel step.
// Entries right of the diagonal are computed without predicates and never stored (don't-care upper triangle).
constexpr int DIAG_NS = 2;
__global__ void __launch_bounds__(64 * DIAG_NS) k_diag_ldlt_row(DevCtx c, const int32_t* __restrict__ pslist) {
    constexpr int WP = 64, NS = DIAG_NS, NE = WP / NS, PB = 8 / NS;   // entries per thread, entries per 8-column block
    __shared__ __align__(16) double col[2][2 * WP];
    pdl_trigger();
    const PStep ps = c.psteps[pslist[blockIdx.x]];
    double* G = c.F + ps.fofs + (int64_t)ps.o + (int64_t)ps.o * ps.ld;
    const int w = ps.w, ld = ps.ld;
    pdl_wait();
    // slices by WARP: the coefficient loads below are then warp-uniform 128-bit broadcasts (2 clocks each; the
    // same loads with 2-4 distinct addresses per warp cost 4 and made shared memory the bottleneck)
    const int r = threadIdx.x & (WP - 1), h = threadIdx.x / WP;
    double a[NE];
#pragma unroll
    for (int li = 0; li < NE; ++li) {
        const int j = ((li >> 1) * NS + h) * 2 + (li & 1);
        a[li] = (j <= r && r < w) ? __ldcg(G + r + (size_t)j * ld) : 0.0;
    }
    for (int i = threadIdx.x; i < 2 * WP; i += 64 * NS) col[i >> 6][WP + (i & 63)] = 0.0;
    if (h == 0) col[0][r] = a[0];                       // column 0
    bool bad = false;
    double* gk = G + r;                                 // &G(r, k)
#pragma unroll 1
    for (int kb = 0; kb < w; kb += 8) {
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const int k = kb + cc;
            if (k < w) {                                // uniform
                const double* cb = col[k & 1];
                double* cn = col[(k + 1) & 1];
                __syncthreads();
                const double d = cb[k];
                const double ar = cb[r];
                const double* cj = cb + kb + h * 2;
                double u[NE];
#pragma unroll
                for (int li = 0; li < NE; ++li) u[li] = (kb + ((li >> 1) * NS + h) * 2 < w) ? cj[(li >> 1) * 2 * NS + (li & 1)] : 0.0;   // uniform: dead pairs cost no bandwidth
                const double l = ar * fast_rcp(d);
                const int cn1 = cc + 1;                 // column k+1: inside this block, or the first one of the next
                const int ho = (cn1 >> 1) % NS, lcn = ((cn1 >> 1) / NS) * 2 + (cn1 & 1);
                a[lcn] -= l * u[lcn];
                if (h == ho && r > k) cn[r] = a[lcn];
#pragma unroll
                for (int li = 0; li < NE; ++li) if (li != lcn) a[li] -= l * u[li];
                bad |= d == 0.0;
                if (h == ((cc >> 1) % NS) && r >= k && r < w) __stcg(gk, r == k ? d : l);
                gk += ld;
            }
        }
#pragma unroll
        for (int li = 0; li < NE - PB; ++li) a[li] = a[li + PB];
#pragma unroll
        for (int li = NE - PB; li < NE; ++li) a[li] = 0.0;
    }
    if (bad && threadIdx.x == 0) atomicExch(c.iflag, -1);
}

// LU with chunk-local partial pivoting of a w x w diagonal block (w <= 64) held in registers: thread (r, h) of
// 64 x LU_NS owns the column pairs P = h (mod LU_NS) of ONE PHYSICAL ROW; slices are per warp, so coefficient
// loads are warp-uniform.  Rows are never moved between threads: an interchange swaps the LOGICAL positions
// `pos` of two threads.  Final values go to a shared image S of the block in logical layout: the pivot row
// publishes its register slice in `prow` — the broadcast buffer of the update, copied to S(k, k..) off the
// dependent chain — and every row below writes its multiplier to S(pos, k).  Interchanges reach back only to
// the chunk's first column (_luswap!/dgetrf on the chunk's own rows, SpkLUFactor.jl:230-240): the already final
// multipliers in S(k, s0..k-1) and S(kp, s0..k-1) are swapped in the image.  Per column: keys (column k by
// logical row) -> barrier -> every warp finds the pivot redundantly (first maximum wins, ggetrf! / idamax;
// three warp reductions on the bit pattern of |x|) -> pivot row published -> barrier -> multipliers (correctly
// rounded reciprocal, as dgetf2) and update, next column's key first.  Predicated per-entry stores and loads
// made a first version instruction bound (440 instructions per column): slices are moved whole.
constexpr int LU_NS = 2;
constexpr int LU_SLD = 66;                              // row stride of the shared image (16-byte aligned rows)
__global__ void __launch_bounds__(64 * LU_NS) k_diag_lu_row(DevCtx c, const int32_t* __restrict__ pslist) {
    constexpr int WP = 64, NS = LU_NS, NE = WP / NS, PB = 8 / NS;
    __shared__ __align__(16) double S[WP * LU_SLD];
    __shared__ __align__(16) double prow[2 * WP];       // the pivot row of the current column, by absolute column
    __shared__ double keys[2][WP];
    pdl_trigger();
    const PStep ps = c.psteps[pslist[blockIdx.x]];
    double* G = c.F + ps.fofs + (int64_t)ps.o + (int64_t)ps.o * ps.ld;
    const int w = ps.w, ld = ps.ld;
    int32_t* ipiv = c.ipiv + ps.col0;
    const int32_t* subw = c.subw + ps.sub0;
    pdl_wait();
    const int r = threadIdx.x & (WP - 1), h = threadIdx.x / WP, lane = threadIdx.x & 31;
    double a[NE];
#pragma unroll
    for (int li = 0; li < NE; ++li) {
        const int j = ((li >> 1) * NS + h) * 2 + (li & 1);
        a[li] = (j < w && r < w) ? __ldcg(G + r + (size_t)j * ld) : 0.0;
    }
    int pos = r;                                        // logical row of this thread's physical row
    if (h == 0) { keys[0][r] = a[0]; prow[WP + r] = 0.0; }
    bool bad = false;
    int s0 = 0, s1 = subw[0], sb = 0;
#pragma unroll 1
    for (int kb = 0; kb < w; kb += 8) {
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const int k = kb + cc;
            if (k < w) {                                // uniform
                if (k == s1) { s0 = s1; s1 += subw[++sb]; }
                const double* kc = keys[k & 1];
                double* kn = keys[(k + 1) & 1];
                __syncthreads();                        // keys of column k (by logical row) visible
                // pivot search, every warp for itself: rows [k, s1).  |x| compares like its bit pattern, so
                // three warp reductions (high word, low word, index) replace a 5-round shuffle tournament.
                unsigned hi = 0, lo = 0; int bi = 0x7fffffff;
                {
                    const double v0 = fabs(kc[lane]), v1 = fabs(kc[lane + 32]);
                    const bool c0 = lane >= k && lane < s1 && v0 == v0, c1 = lane + 32 >= k && lane + 32 < s1 && v1 == v1;
                    const unsigned h0 = __double2hiint(v0), l0 = __double2loint(v0), h1 = __double2hiint(v1), l1 = __double2loint(v1);
                    const bool take1 = c1 && (!c0 || h1 > h0 || (h1 == h0 && l1 > l0));
                    if (c0 || c1) { hi = take1 ? h1 : h0; lo = take1 ? l1 : l0; bi = take1 ? lane + 32 : lane; }
                    const unsigned cand = bi != 0x7fffffff;
                    const unsigned mh = __reduce_max_sync(0xffffffffu, cand ? hi : 0u);
                    const bool inh = cand && hi == mh;
                    const unsigned ml = __reduce_max_sync(0xffffffffu, inh ? lo : 0u);
                    const bool inl = inh && lo == ml;
                    bi = (int)__reduce_min_sync(0xffffffffu, inl ? (unsigned)bi : 0x7fffffffu);
                    if (bi == 0x7fffffff) bi = k;       // only NaNs: keep the diagonal
                }
                const int kp = bi;
                const double pv = kc[kp];
                const bool ok = pv != 0.0;
                bad |= !ok;
                const double rinv = ok ? __drcp_rn(pv) : 1.0;
                const int oldpos = pos;
                if (ok && kp != k) { if (pos == kp) pos = k; else if (pos == k) pos = kp; }
                if (pos == k) {                         // the pivot row: publish its slice (entries left of k are dead values)
                    double2* pw = reinterpret_cast<double2*>(prow + kb + h * 2);
#pragma unroll
                    for (int p2 = 0; p2 < NE / 2; ++p2) pw[p2 * NS] = make_double2(a[2 * p2], a[2 * p2 + 1]);
                }
                __syncthreads();                        // pivot row visible
                if (threadIdx.x == 0) ipiv[k] = kp - s0 + 1;
                if (h == NS - 1) {
                    if (r >= k) S[k * LU_SLD + r] = prow[r];            // row k of U is final
                    // multipliers of this chunk already in the image travel with their rows (nobody reads them here)
                    else if (ok && kp != k && r >= s0) { const double t = S[k * LU_SLD + r]; S[k * LU_SLD + r] = S[kp * LU_SLD + r]; S[kp * LU_SLD + r] = t; }
                }
                const double l = kc[oldpos] * rinv;
                const bool below = pos > k;
                if (below && h == ((cc >> 1) % NS) && pos < w) S[pos * LU_SLD + k] = l;
                const double lz = below ? l : 0.0;      // finished rows: keep the registers finite
                const double* pj = prow + kb + h * 2;
                double u[NE];
#pragma unroll
                for (int li = 0; li < NE; ++li) u[li] = pj[(li >> 1) * 2 * NS + (li & 1)];
                const int cn1 = cc + 1;                 // column k+1: inside this block, or the first one of the next
                const int ho = (cn1 >> 1) % NS, lcn = ((cn1 >> 1) / NS) * 2 + (cn1 & 1);
                a[lcn] -= lz * u[lcn];
                if (h == ho) kn[pos] = a[lcn];
#pragma unroll
                for (int li = 0; li < NE; ++li) if (li != lcn) a[li] -= lz * u[li];
            }
        }
#pragma unroll
        for (int li = 0; li < NE - PB; ++li) a[li] = a[li + PB];
#pragma unroll
        for (int li = NE - PB; li < NE; ++li) a[li] = 0.0;
    }
    if (bad && threadIdx.x == 0) atomicExch(c.iflag, -1);
    __syncthreads();
    for (int e = threadIdx.x; e < w * w; e += 64 * NS) { const int j = e / w, i = e - j * w; __stcg(G + i + (size_t)j * ld, S[i * LU_SLD + j]); }
}

// ------------------------------------------------------------------------------------
// Panels of a panel step.  One block = PANEL_ROWS front rows below the block (L side) or PANEL_ROWS
// front columns to its right (U side); the factored w x w block T and the block's slice of the
// panel are staged in shared memory (coalesced both ways), one thread then owns one row / column:
//   LU  L-side: X = A21 * inv(U11)                     (dtrsm 'r','u','n','n', SpkLUFactor.jl:235)
//   LU  U-side: per chunk, apply its row interchanges then inv(L11)  (:238-240, _luswap! SpkSpdMMOps.jl:168-175)
//   LDLT:       X = A21 * inv(L11^T), then each column / D  (SpkLDLtFactor.jl:367-377)
// Shared layout: Ts[w*w] (column-major), Xs[w][PANEL_ROWS] (element k of thread t at Xs[k*PANEL_ROWS + t]).
// Triangular substitution on one thread's vector x (element k at x[k * PANEL_ROWS]) against the staged
// block Ts (w x w, column-major).  Columns are processed 8 at a time: the contributions of already final
// unknowns run as 8 independent FMA chains (throughput-bound), only the 8 x 8 in-block part is a dependent
// chain.  Per unknown the terms are still subtracted in ascending k, exactly as a plain column loop would.
//   UPPER: x_j = (a_j - sum_{k<j} x_k T[k,j]) / T[j,j]     (right solve with U11)
//   else : x_j =  a_j - sum_{k<j} x_k T[j,k]               (right solve with unit-lower L11^T / left solve with L11)
template <bool UPPER>
__device__ __forceinline__ double ts_coef(const double* __restrict__ Ts, int w, int k, int j) {
    return UPPER ? Ts[k + (size_t)j * w] : Ts[j + (size_t)k * w];      // w = leading dimension of the block
}
// x_j -= sum_{k in [k0,k1)} coef(k,j) x_k  for j in [j0,j1);  requires k1 <= j0
template <bool UPPER>
__device__ __forceinline__ void ts_accum(double* x, const double* __restrict__ Ts, int w, int j0, int j1, int k0, int k1) {
    for (int jb = j0; jb < j1; jb += 8) {
        const int nb = min(8, j1 - jb);
        double acc[8];
#pragma unroll
        for (int cidx = 0; cidx < 8; ++cidx) acc[cidx] = cidx < nb ? x[(jb + cidx) * PANEL_ROWS] : 0.0;
        for (int k = k0; k < k1; ++k) {
            const double xk = x[k * PANEL_ROWS];
#pragma unroll
            for (int cidx = 0; cidx < 8; ++cidx) if (cidx < nb) acc[cidx] -= ts_coef<UPPER>(Ts, w, k, jb + cidx) * xk;
        }
#pragma unroll
        for (int cidx = 0; cidx < 8; ++cidx) if (cidx < nb) x[(jb + cidx) * PANEL_ROWS] = acc[cidx];
    }
}
// triangular solve restricted to [j0,j1): unknowns before j0 contribute nothing here
template <bool UPPER>
__device__ __forceinline__ void ts_solve(double* x, const double* __restrict__ Ts, int w, int j0, int j1) {
    for (int jb = j0; jb < j1; jb += 8) {
        const int nb = min(8, j1 - jb);
        double acc[8];
#pragma unroll
        for (int cidx = 0; cidx < 8; ++cidx) acc[cidx] = cidx < nb ? x[(jb + cidx) * PANEL_ROWS] : 0.0;
        for (int k = j0; k < jb; ++k) {
            const double xk = x[k * PANEL_ROWS];
#pragma unroll
            for (int cidx = 0; cidx < 8; ++cidx) if (cidx < nb) acc[cidx] -= ts_coef<UPPER>(Ts, w, k, jb + cidx) * xk;
        }
#pragma unroll
        for (int cidx = 0; cidx < 8; ++cidx) {
            if (cidx < nb) {
#pragma unroll
                for (int c2 = 0; c2 < 8; ++c2) if (c2 < cidx) acc[cidx] -= ts_coef<UPPER>(Ts, w, jb + c2, jb + cidx) * acc[c2];
                if (UPPER) acc[cidx] = (1.0 / Ts[(jb + cidx) + (size_t)(jb + cidx) * w]) * acc[cidx];
                x[(jb + cidx) * PANEL_ROWS] = acc[cidx];
            }
        }
    }
}

// T is staged in shared memory next to the panel slice when both fit; wide steps keep T in global memory
// (every thread reads the same entries, so they are L1 broadcasts).
constexpr size_t PANEL_SMEM_LIMIT = 200 * 1024;
inline bool panel_stage_T(int w) { return ((size_t)w * w + (size_t)w * PANEL_ROWS) * sizeof(double) <= PANEL_SMEM_LIMIT; }
inline size_t panel_smem_bytes(int w) { return ((panel_stage_T(w) ? (size_t)w * w : 0) + (size_t)w * PANEL_ROWS) * sizeof(double); }

template <bool LU>
__global__ void __launch_bounds__(PANEL_ROWS) k_panel(DevCtx c, const int32_t* __restrict__ pslist,
                                                      const int32_t* __restrict__ pfx, int count, int skip_lside) {
    extern __shared__ double psm[];
    int t = find_task(pfx, count, blockIdx.x);
    int lb = blockIdx.x - pfx[t];
    const PStep ps = c.psteps[pslist[t]];
    const int w = ps.w, ld = ps.ld, e0 = ps.o + ps.w;
    const int below = ps.R - e0;
    const int nb = (below + PANEL_ROWS - 1) / PANEL_ROWS;
    const int tid = threadIdx.x;
    double* Fm = c.F + ps.fofs;
    const double* __restrict__ T = Fm + (int64_t)ps.o + (int64_t)ps.o * ld;   // factored w x w block
    const bool stage = ((size_t)w * w + (size_t)w * PANEL_ROWS) * sizeof(double) <= PANEL_SMEM_LIMIT;
    const double* Ts = stage ? psm : T;                 // Ts(i,j) = Ts[i + j * ldt]
    const int ldt = stage ? w : ld;
    double* Xs = psm + (stage ? w * w : 0);
    if (stage) block_g2s<PANEL_ROWS>(psm, w, T, ld, w);
    const bool lside = lb < nb;
    if (lside && skip_lside) return;                    // done by k_panel_reg
    const int i0 = (lside ? lb : lb - nb) * PANEL_ROWS;
    const int cnt = min(PANEL_ROWS, below - i0);
    if (lside) {
        const double* X0 = Fm + (int64_t)(e0 + i0) + (int64_t)ps.o * ld;       // rows contiguous
        if (tid < cnt)
            for (int k0 = 0; k0 < w; k0 += 8) {                                // 8 independent loads in flight
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = (k0 + u < w) ? __ldcg(X0 + tid + (size_t)(k0 + u) * ld) : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u) if (k0 + u < w) Xs[(k0 + u) * PANEL_ROWS + tid] = v[u];
            }
    } else {
        const double* Y0 = Fm + (int64_t)ps.o + (int64_t)(e0 + i0) * ld;       // each column: w contiguous entries
        const int tot = w * cnt;
        for (int e0i = tid; e0i < tot; e0i += 8 * PANEL_ROWS) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { int e = e0i + u * PANEL_ROWS; int col = e / w, k = e - col * w; v[u] = (e < tot) ? __ldcg(Y0 + k + (size_t)col * ld) : 0.0; }
#pragma unroll
            for (int u = 0; u < 8; ++u) { int e = e0i + u * PANEL_ROWS; int col = e / w, k = e - col * w; if (e < tot) Xs[k * PANEL_ROWS + col] = v[u]; }
        }
    }
    __syncthreads();
    if (tid < cnt) {
        double* x = Xs + tid;                                                  // x[k * PANEL_ROWS]
        if (lside) {
            if (LU) ts_solve<true>(x, Ts, ldt, 0, w);                          // X = A21 * inv(U11)
            else ts_solve<false>(x, Ts, ldt, 0, w);                            // X = A21 * inv(L11^T), kept UNSCALED here
        } else if (LU) {
            const int32_t* ipiv = c.ipiv + ps.col0;
            const int32_t* subw = c.subw + ps.sub0;
            int s0 = 0;
            for (int b = 0; b < ps.nsub; ++b) {
                const int s1 = s0 + subw[b];
                ts_accum<false>(x, Ts, ldt, s0, s1, 0, s0);                      // contributions of earlier chunks (unswapped rows)
                for (int k = s0; k < s1; ++k) {                                // this chunk's interchanges
                    int ip = s0 + ipiv[k] - 1;
                    if (ip != k) { double tmp = x[k * PANEL_ROWS]; x[k * PANEL_ROWS] = x[ip * PANEL_ROWS]; x[ip * PANEL_ROWS] = tmp; }
                }
                ts_solve<false>(x, Ts, ldt, s0, s1);                             // unit-lower solve inside the chunk
                s0 = s1;
            }
        }
    }
    __syncthreads();
    if (lside) {
        double* X0 = Fm + (int64_t)(e0 + i0) + (int64_t)ps.o * ld;
        if (LU) {
            for (int k = 0; k < w; ++k) if (tid < cnt) X0[tid + (size_t)k * ld] = Xs[k * PANEL_ROWS + tid];
        } else {
            // L21 = X / D (SpkLDLtFactor.jl:372-377); U12 = X^T = D * L21^T goes to the upper triangle of the
            // front, where the trailing-update kernel reads its B operand
            for (int k = 0; k < w; ++k) if (tid < cnt) X0[tid + (size_t)k * ld] = Xs[k * PANEL_ROWS + tid] / Ts[k + (size_t)k * ldt];
            double* Y0 = Fm + (int64_t)ps.o + (int64_t)(e0 + i0) * ld;
            for (int e = tid; e < w * cnt; e += PANEL_ROWS) { int col = e / w, k = e - col * w; Y0[k + (size_t)col * ld] = Xs[k * PANEL_ROWS + col]; }
        }
    } else if (LU) {
        double* Y0 = Fm + (int64_t)ps.o + (int64_t)(e0 + i0) * ld;
        for (int e = tid; e < w * cnt; e += PANEL_ROWS) { int col = e / w, k = e - col * w; Y0[k + (size_t)col * ld] = Xs[k * PANEL_ROWS + col]; }
    }
}

// L-side panel of a panel step with w <= 64, one thread per row, the ROW IN REGISTERS and the factored
// block T (transposed for LU, so that the coefficients of one elimination step are contiguous) zero-padded
// in shared memory together with the reciprocals of its diagonal.  All loads of a row are in flight at once;
// the substitution is right-looking (x_j -= t_jk x_k for all j > k: independent FMAs, no communication between
// threads, coefficients fetched as warp-uniform 128-bit broadcasts), written as a runtime loop over blocks of
// 8 columns with the register row shifted down by 8 after each block (a fully unrolled body is
// instruction-fetch bound); coefficient pairs beyond the block's width are skipped.
// Measured alternatives (tools/ubench_steps.cu): rows split over 4 lanes + shuffles cost 4 clocks per
// coefficient load (4 distinct addresses) and twice the instructions; 4 rows x 16 columns per thread leaves
// one warp per scheduler, issue bound.
//   LU  : X = A21 * inv(U11);   LDLT: X = A21 * inv(L11^T), L21 = X / D written below, U12 = X^T above.
constexpr int PANEL_REG_SPLIT = 1;                                    // launch blocks per PANEL_ROWS rows (more SMs on the few blocks of a top front)
constexpr int PANEL_REG_THREADS = PANEL_ROWS / PANEL_REG_SPLIT;
#ifndef PANEL_REG_MINB
#define PANEL_REG_MINB 1      /* 3 blocks per SM (168 registers, spills) measured slower: 26 vs 21 ms of panel time at 96^3 */
#endif
template <bool LU>
__global__ void __launch_bounds__(PANEL_REG_THREADS, PANEL_REG_MINB) k_panel_reg(DevCtx c, const int32_t* __restrict__ pslist,
                                                                 const int32_t* __restrict__ pfx, int count) {
    constexpr int WP = 64;
    __shared__ __align__(16) double Ts[WP * WP + WP];   // Ts[j + k*WP] = coefficient of x_k in unknown j (j > k)
    __shared__ double rd[WP];                           // 1 / diagonal (0 where the diagonal is 0, as the LU reference does)
    __shared__ int32_t perm[WP], s0r[WP];               // U side: original row of each final row, first row of its chunk
    pdl_trigger();
    const int bx = blockIdx.x / PANEL_REG_SPLIT, part = blockIdx.x % PANEL_REG_SPLIT;
    int t = find_task(pfx, count, bx);
    int lb = bx - pfx[t];
    const PStep ps = c.psteps[pslist[t]];
    const int w = ps.w, ld = ps.ld, e0 = ps.o + ps.w;
    const int below = ps.R - e0;
    const int nb = (below + PANEL_ROWS - 1) / PANEL_ROWS;
    pdl_wait();
    // LU, U side (blocks nb .. 2nb-1): U12 = inv(L11) * P * A12 with P the chunk-local interchanges.  The rows of
    // A12 are gathered through P up front and the multipliers a chunk inherits from earlier chunks (stored in
    // pre-interchange row order, SpkLUFactor.jl:230-240) are gathered the same way, which leaves one plain
    // unit-lower substitution — the same loop as the L side, one thread per column of U12.
    const bool lside = !LU || lb < nb;
    if (!lside) lb -= nb;
    const int tid = threadIdx.x;
    double* Fm = c.F + ps.fofs;
    const double* __restrict__ T = Fm + (int64_t)ps.o + (int64_t)ps.o * ld;
    const int i = lb * PANEL_ROWS + part * PANEL_REG_THREADS + tid;
    if (i - tid >= below) return;                       // whole block beyond the panel
    const bool active = i < below;
    double* xp = Fm + (int64_t)(e0 + (active ? i : 0)) + (int64_t)ps.o * ld;          // L side: &X(i, k)
    double* yp = Fm + (int64_t)ps.o + (int64_t)(e0 + (active ? i : 0)) * ld;          // LDLT: &U12(k, i); LU U side: &U12(k, i)
    double x[WP];
    if (lside) {
#pragma unroll
        for (int k = 0; k < WP; ++k) x[k] = (active && k < w) ? __ldcs(xp + (size_t)k * ld) : 0.0;
    } else {
        if (tid < WP) perm[tid] = tid;
        __syncthreads();
        if (tid < ps.nsub) {                            // one thread per chunk replays its interchanges
            const int32_t* subw = c.subw + ps.sub0;
            const int32_t* ipiv = c.ipiv + ps.col0;
            int s0 = 0;
            for (int b2 = 0; b2 < tid; ++b2) s0 += subw[b2];
            const int s1 = s0 + subw[tid];
            for (int k = s0; k < s1; ++k) {
                const int kp = s0 + ipiv[k] - 1;
                const int tp = perm[k]; perm[k] = perm[kp]; perm[kp] = tp;
                s0r[k] = s0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < WP; ++k) x[k] = (active && k < w) ? __ldcs(yp + perm[k]) : 0.0;
    }
    constexpr int TB = 16;                              // loads in flight per thread while staging T
    for (int e0i = tid; e0i < WP * WP; e0i += TB * PANEL_REG_THREADS) {
        double v[TB];
#pragma unroll
        for (int u = 0; u < TB; ++u) {
            const int e = e0i + u * PANEL_REG_THREADS; const int k = e / WP, j = e - k * WP;       // smem slot (j,k)
            const bool in = j < w && k < w && j > k;
            const double* src;
            if (!LU) src = T + j + (size_t)k * ld;
            else if (lside) src = T + k + (size_t)j * ld;
            else src = T + ((in && k < s0r[j]) ? perm[j] : j) + (size_t)k * ld;
            v[u] = in ? __ldcg(src) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < TB; ++u) Ts[e0i + u * PANEL_REG_THREADS] = v[u];
    }
    static_assert((WP * WP) % (TB * PANEL_REG_THREADS) == 0, "whole batches");
    if (tid < WP) {
        Ts[WP * WP + tid] = 0.0;
        const double dg = tid < w ? __ldcg(T + tid + (size_t)tid * ld) : 0.0;
        rd[tid] = lside ? (dg != 0.0 ? 1.0 / dg : 0.0) : 1.0;
    }
    __syncthreads();
#pragma unroll 1
    for (int kb = 0; kb < w; kb += 8) {
        const int live = w - kb;                        // unknowns kb .. w-1 are still open: local indices < live
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const int k = kb + cc;
            const double* tk = Ts + kb + k * WP;        // tk[jj] = coefficient for unknown kb + jj
            double xk = x[cc];
            if (LU) xk *= rd[k & (WP - 1)];
            if (active && k < w) {
                if (LU) { if (lside) *xp = xk; else *yp = xk; }
                else { *xp = xk * rd[k]; *yp = xk; }
            }
            xp += ld; ++yp;
#pragma unroll
            for (int q0 = 0; q0 < WP; q0 += 16) {
                if (q0 < live) {                        // uniform: dead quarters cost no shared-memory bandwidth
#pragma unroll
                    for (int jj = q0; jj < q0 + 16; jj += 2) {
                        if (jj + 1 > cc) {              // static
                            const double2 tt = *reinterpret_cast<const double2*>(tk + jj);
                            if (jj > cc) x[jj] -= tt.x * xk;
                            x[jj + 1] -= tt.y * xk;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < WP - 8; ++j) x[j] = x[j + 8];
#pragma unroll
        for (int j = WP - 8; j < WP; ++j) x[j] = 0.0;
    }
}

// C -= A * B for the tiny fronts at the bottom of the tree (every task of the launch has m, n <= GEMM_TINY, so
// one block per task): element-wise dot products straight from global memory, 128-thread blocks, many per SM —
// tens of thousands of such tasks per level are launch-slot bound in a tiled kernel.  Same summation order over k
// as the tiled kernel.
constexpr int GEMM_TINY = 48;
__global__ void __launch_bounds__(128) k_gemm_tiny(DevCtx c, const GemmTask* __restrict__ tasks, int count) {
    if ((int)blockIdx.x >= count) return;
    const GemmTask g = tasks[blockIdx.x];
    const double* __restrict__ A = c.F + g.a0;
    const double* __restrict__ B = c.F + g.b0;
    double* __restrict__ C = c.F + g.c0;
    const int ld = g.ld, tot = g.m * g.n;
    for (int e = threadIdx.x; e < tot; e += blockDim.x) {
        const int j = e / g.m, i = e - j * g.m;
        if (g.lower && i + g.roff < j) continue;
        double acc = 0.0;
        for (int k = 0; k < g.k; ++k) acc += A[(size_t)i + (size_t)k * ld] * B[(size_t)k + (size_t)j * ld];
        C[(size_t)i + (size_t)j * ld] -= acc;
    }
}

// ------------------------------------------------------------------------------------
// C -= A * B inside a frontal matrix (small-tile DFMA kernel; 64x64 tile, 4x4 per thread).
// The big-tile DMMA kernels live in gemm_dmma.cuh.
__global__ void __launch_bounds__(256) k_gemm_small(DevCtx c, const GemmTask* __restrict__ tasks,
                                                    const int32_t* __restrict__ pfx, int count) {
    constexpr int TM = GEMM_TM, TN = GEMM_TN, TK = 16;
    __shared__ double As[TK][TM + 4];
    __shared__ double Bs[TK][TN + 4];
    int t = find_task(pfx, count, blockIdx.x);
    int lb = blockIdx.x - pfx[t];
    const GemmTask g = tasks[t];
    const int mt = (g.m + TM - 1) / TM;
    const int row0 = (lb % mt) * TM, col0 = (lb / mt) * TN;
    if (g.lower && row0 + TM - 1 + g.roff < col0) return;   // tile strictly above the diagonal
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    const double* __restrict__ A = c.F + g.a0;
    const double* __restrict__ B = c.F + g.b0;
    const int ld = g.ld;
    for (int k0 = 0; k0 < g.k; k0 += TK) {
        {   // A tile: rows contiguous
            const int lr = tid & 63, lk = tid >> 6;
#pragma unroll
            for (int p = 0; p < TK / 4; ++p) {
                int kk = lk + p * 4, kg = k0 + kk;
                double a = 0.0;
                if (kg < g.k && row0 + lr < g.m) a = A[(size_t)(row0 + lr) + (size_t)kg * ld];
                As[kk][lr] = a;
            }
        }
        {   // B tile: k contiguous
            const int kk = tid & 15, ln = tid >> 4;
#pragma unroll
            for (int p = 0; p < TN / 16; ++p) {
                int nn = ln + p * 16, kg = k0 + kk;
                double b = 0.0;
                if (kg < g.k && col0 + nn < g.n) b = B[(size_t)kg + (size_t)(col0 + nn) * ld];
                Bs[kk][nn] = b;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][tx * 4 + i]; b[i] = Bs[kk][ty * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
    double* __restrict__ C = c.F + g.c0;
    double cv[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int cc = col0 + ty * 4 + j, r = row0 + tx * 4 + i;
            const bool ok = cc < g.n && r < g.m && !(g.lower && r + g.roff < cc);
            cv[i][j] = ok ? C[(size_t)r + (size_t)cc * ld] : 0.0;
        }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int cc = col0 + ty * 4 + j, r = row0 + tx * 4 + i;
            if (cc < g.n && r < g.m && !(g.lower && r + g.roff < cc)) C[(size_t)r + (size_t)cc * ld] = cv[i][j] - acc[i][j];
        }
}

// ------------------------------------------------------------------------------------
// Triangular solves (SpkLUFactor.jl:269-377, SpkLDLtFactor.jl:266-293), level-scheduled over
// the front tree, reading lnz / unz in the reference layout.  Each front owns a work vector
// w_f indexed by front row: [unknowns of the front (W) ; contributions / values of the rows below (m)].
// blockIdx.y = right-hand side.
__global__ void __launch_bounds__(256) k_fwd_gather(DevCtx c, const int32_t* __restrict__ flist,
                                                    const double* __restrict__ rhs, int64_t ldrhs) {
    const DFront F = c.fronts[flist[blockIdx.x]];
    double* w = c.w + (size_t)blockIdx.y * c.wlen + F.wofs;
    const double* b = rhs + (size_t)blockIdx.y * ldrhs + F.F0;
    for (int i = threadIdx.x; i < F.R; i += blockDim.x) w[i] = i < F.W ? b[i] : 0.0;
    __syncthreads();
    for (int r = 0; r < F.nchild; ++r) {
        const DFront C = c.fronts[c.childlist[F.child0 + r]];
        const double* wc = c.w + (size_t)blockIdx.y * c.wlen + C.wofs + C.W;
        const int32_t* rel = c.rel + C.relofs;
        for (int i = threadIdx.x; i < C.m; i += blockDim.x) w[rel[i]] += wc[i];
        __syncthreads();
    }
}

template <bool LU>
__global__ void __launch_bounds__(128) k_fwd_diag(DevCtx c, const int32_t* __restrict__ clist) {
    const SolveTask t = c.solvet[clist[blockIdx.x]];
    double* x = c.w + (size_t)blockIdx.y * c.wlen + t.wofs + t.o;
    const double* __restrict__ T = c.lnz + t.lofs;
    if (LU) {
        if (threadIdx.x == 0) {
            const int32_t* ipiv = c.ipiv + t.col0;
            for (int k = 0; k < t.nj; ++k) { int ip = ipiv[k] - 1; if (ip != k) { double tmp = x[k]; x[k] = x[ip]; x[ip] = tmp; } }
        }
        __syncthreads();
    }
    for (int k = 0; k < t.nj - 1; ++k) {
        double xk = x[k];
        for (int i = k + 1 + threadIdx.x; i < t.nj; i += blockDim.x) x[i] -= xk * T[i + (size_t)k * t.ld];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(UPD_ROWS) k_fwd_update(DevCtx c, const int32_t* __restrict__ clist,
                                                         const int32_t* __restrict__ pfx, int count) {
    int ti = find_task(pfx, count, blockIdx.x);
    int lb = blockIdx.x - pfx[ti];
    const SolveTask t = c.solvet[clist[ti]];
    int i = lb * UPD_ROWS + threadIdx.x;
    if (i >= t.m) return;
    double* wf = c.w + (size_t)blockIdx.y * c.wlen + t.wofs;
    const double* xk = wf + t.o;
    const double* __restrict__ L = c.lnz + t.lofs + t.nj + i;
    double acc = 0.0;
    for (int k = 0; k < t.nj; ++k) acc += (-xk[k]) * L[(size_t)k * t.ld];
    wf[c.pos[t.posofs + t.nj + i]] += acc;
}

__global__ void __launch_bounds__(256) k_bwd_gather(DevCtx c, const int32_t* __restrict__ flist,
                                                    const int32_t* __restrict__ pfx, int count) {
    int ti = find_task(pfx, count, blockIdx.x);
    int lb = blockIdx.x - pfx[ti];
    const DFront F = c.fronts[flist[ti]];
    int i = lb * 256 + threadIdx.x;
    if (i >= F.m) return;
    const DFront P = c.fronts[F.parent];
    const double* wp = c.w + (size_t)blockIdx.y * c.wlen + P.wofs;
    double* w = c.w + (size_t)blockIdx.y * c.wlen + F.wofs;
    w[F.W + i] = wp[c.rel[F.relofs + i]];
}

// one block per column k of the chunk: x_k -= sum_i B[i,k] * x_below[i]   (fixed-order block reduction)
template <bool LU>
__global__ void __launch_bounds__(256) k_bwd_update(DevCtx c, const int32_t* __restrict__ clist,
                                                    const int32_t* __restrict__ pfx, int count) {
    __shared__ double red[8];
    int ti = find_task(pfx, count, blockIdx.x);
    const int k = blockIdx.x - pfx[ti];
    const SolveTask t = c.solvet[clist[ti]];
    if (k >= t.nj) return;
    double* wf = c.w + (size_t)blockIdx.y * c.wlen + t.wofs;
    const double* __restrict__ B = LU ? c.unz + t.uofs + (size_t)k * t.ldu : c.lnz + t.lofs + t.nj + (size_t)k * t.ld;
    const int32_t* __restrict__ pos = c.pos + t.posofs + t.nj;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = threadIdx.x;
    for (; i + 768 < t.m; i += 1024) {
        s0 += B[i] * wf[pos[i]]; s1 += B[i + 256] * wf[pos[i + 256]];
        s2 += B[i + 512] * wf[pos[i + 512]]; s3 += B[i + 768] * wf[pos[i + 768]];
    }
    for (; i < t.m; i += 256) s0 += B[i] * wf[pos[i]];
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        s = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
        double* xk = wf + t.o + k;
        if (LU) *xk += -s;
        else *xk = *xk / c.lnz[t.lofs + k + (size_t)k * t.ld] - s;
    }
}

template <bool LU>
__global__ void __launch_bounds__(128) k_bwd_diag(DevCtx c, const int32_t* __restrict__ clist,
                                                  double* __restrict__ rhs, int64_t ldrhs) {
    const SolveTask t = c.solvet[clist[blockIdx.x]];
    double* x = c.w + (size_t)blockIdx.y * c.wlen + t.wofs + t.o;
    const double* __restrict__ T = c.lnz + t.lofs;
    for (int k = t.nj - 1; k >= 0; --k) {
        if (LU) {
            if (threadIdx.x == 0) x[k] /= T[k + (size_t)k * t.ld];
            __syncthreads();
            double xk = x[k];
            for (int i = threadIdx.x; i < k; i += blockDim.x) x[i] -= xk * T[i + (size_t)k * t.ld];
        } else {
            double xk = x[k];
            for (int i = threadIdx.x; i < k; i += blockDim.x) x[i] -= xk * T[k + (size_t)i * t.ld];
        }
        __syncthreads();
    }
    double* out = rhs + (size_t)blockIdx.y * ldrhs + t.col0;
    for (int k = threadIdx.x; k < t.nj; k += blockDim.x) out[k] = x[k];
}

// Small fronts: one block walks all chunks of the front (same arithmetic as the per-chunk kernels).
template <bool LU>
__global__ void __launch_bounds__(256) k_fwd_front(DevCtx c, const int32_t* __restrict__ flist) {
    const DFront F = c.fronts[flist[blockIdx.x]];
    double* wf = c.w + (size_t)blockIdx.y * c.wlen + F.wofs;
    for (int tc = 0; tc < F.nch; ++tc) {
        const SolveTask t = c.solvet[F.c0 + tc];
        double* x = wf + t.o;
        const double* __restrict__ T = c.lnz + t.lofs;
        if (LU) {
            if (threadIdx.x == 0) {
                const int32_t* ipiv = c.ipiv + t.col0;
                for (int k = 0; k < t.nj; ++k) { int ip = ipiv[k] - 1; if (ip != k) { double tmp = x[k]; x[k] = x[ip]; x[ip] = tmp; } }
            }
            __syncthreads();
        }
        for (int k = 0; k < t.nj - 1; ++k) {
            double xk = x[k];
            for (int i = k + 1 + threadIdx.x; i < t.nj; i += blockDim.x) x[i] -= xk * T[i + (size_t)k * t.ld];
            __syncthreads();
        }
        const int32_t* __restrict__ pos = c.pos + t.posofs + t.nj;
        for (int i = threadIdx.x; i < t.m; i += blockDim.x) {
            const double* __restrict__ L = T + t.nj + i;
            double acc = 0.0;
            for (int k = 0; k < t.nj; ++k) acc += (-x[k]) * L[(size_t)k * t.ld];
            wf[pos[i]] += acc;
        }
        __syncthreads();
    }
}

template <bool LU>
__global__ void __launch_bounds__(256) k_bwd_front(DevCtx c, const int32_t* __restrict__ flist,
                                                   double* __restrict__ rhs, int64_t ldrhs) {
    const DFront F = c.fronts[flist[blockIdx.x]];
    double* wf = c.w + (size_t)blockIdx.y * c.wlen + F.wofs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int tc = F.nch - 1; tc >= 0; --tc) {
        const SolveTask t = c.solvet[F.c0 + tc];
        double* x = wf + t.o;
        const double* __restrict__ T = c.lnz + t.lofs;
        const int32_t* __restrict__ pos = c.pos + t.posofs + t.nj;
        if (t.m > 0 || !LU) {
            for (int k = warp; k < t.nj; k += 8) {
                const double* __restrict__ B = LU ? c.unz + t.uofs + (size_t)k * t.ldu : T + t.nj + (size_t)k * t.ld;
                double s = 0.0;
                for (int i = lane; i < t.m; i += 32) s += B[i] * wf[pos[i]];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if (lane == 0) { if (LU) x[k] += -s; else x[k] = x[k] / T[k + (size_t)k * t.ld] - s; }
            }
            __syncthreads();
        }
        for (int k = t.nj - 1; k >= 0; --k) {
            if (LU) {
                if (threadIdx.x == 0) x[k] /= T[k + (size_t)k * t.ld];
                __syncthreads();
                double xk = x[k];
                for (int i = threadIdx.x; i < k; i += blockDim.x) x[i] -= xk * T[i + (size_t)k * t.ld];
            } else {
                double xk = x[k];
                for (int i = threadIdx.x; i < k; i += blockDim.x) x[i] -= xk * T[k + (size_t)i * t.ld];
            }
            __syncthreads();
        }
    }
    double* out = rhs + (size_t)blockIdx.y * ldrhs + F.F0;
    for (int k = threadIdx.x; k < F.W; k += blockDim.x) out[k] = wf[k];
}

// ------------------------------------------------------------------------------------
// Panel-step solves on the frontal matrices.  After the factorisation a front still holds L (below the
// diagonal, unit) and U (on/above; LDL^T: D on the diagonal) of its own W columns as dense, uniformly
// strided panels, so one solve step covers a whole panel step (<= ~64 columns, several reference chunks)
// with dense row-contiguous panels instead of one step per chunk through position maps.
// Same arithmetic as SpkLUFactor.jl:297-321,350-375 / SpkLDLtFactor.jl:268-291, regrouped.
//
// pf_diag: x := inv(L11) P x for the w unknowns of the step (Ts = staged w x w block, xs = staged x).
// Executed by ONE warp on shared memory (column sweep, __syncwarp between columns): no block barriers.
// Each lane keeps the unknowns lane, lane+32, lane+64 in registers; x_k is broadcast with one shuffle
// per column, so a column costs a shuffle + FMA instead of a shared-memory round trip (w <= 96).


// ------------------------------------------------------------------------------------
// FUSED small-front kernel (K_FRONT_SMALL): a front of at most FUSED_MAXR rows is handled by ONE thread block in
// shared memory — load (the values scattered by inmatrix), extend-add of all children in their fixed order, partial
// factorisation of its W columns (right-looking, column by column; LU: partial pivoting restricted to each chunk's
// own rows, first maximum wins, interchanges applied from the chunk's first column rightwards, block-local ipiv —
// the rules of k_diag / k_panel), store.  The bottom levels of the tree hold tens of thousands of such fronts; the
// kernel-per-operation path spends ~10 launches and as many global round trips per level on them for < 1 % of the
// flops.  LDL^T fronts leave U = D L^T above the diagonal as the other kernels do.
constexpr int FS_NT = 64;
inline size_t front_small_smem_bytes(int maxR) { return (size_t)maxR * (maxR | 1) * sizeof(double); }
template <bool LU>
__global__ void __launch_bounds__(FS_NT) k_front_small(DevCtx c, const int32_t* __restrict__ flist) {
    extern __shared__ double S[];
    __shared__ int s_piv;
    const DFront F = c.fronts[flist[blockIdx.x]];
    const int R = F.R, lds = R | 1, tid = threadIdx.x;
    double* __restrict__ G = c.F + F.fofs;
    for (int e = tid; e < R * R; e += FS_NT) { const int j = e / R, i = e - j * R; S[i + j * lds] = G[i + (int64_t)j * F.ld]; }
    __syncthreads();
    for (int ch = 0; ch < F.nchild; ++ch) {                 // children in plan order: the same sums as k_assemble
        const DFront C = c.fronts[c.childlist[F.child0 + ch]];
        const int32_t* __restrict__ rel = c.rel + C.relofs;
        const int m = C.m;
        const double* __restrict__ CS = c.F + C.fofs + C.W + (int64_t)C.W * C.ld;
        for (int e = tid; e < m * m; e += FS_NT) {
            const int j = e / m, i = e - j * m;
            if (!LU && i < j) continue;
            S[rel[i] + rel[j] * lds] += CS[i + (int64_t)j * C.ld];
        }
        __syncthreads();
    }
    for (int js = 0; js < F.nps; ++js) {
        const PStep ps = c.psteps[F.ps0 + js];
        int s0 = ps.o;
        for (int b = 0; b < ps.nsub; ++b) {
            const int s1 = s0 + c.subw[ps.sub0 + b];
            for (int k = s0; k < s1; ++k) {
                if (LU) {
                    if (tid < 32) {                         // first maximum of |S[k.., k]| over the chunk's rows
                        double best = -1.0; int bi = k;
                        for (int i = k + tid; i < s1; i += 32) { const double v = fabs(S[i + k * lds]); if (v > best) { best = v; bi = i; } }
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) {
                            const double ov = __shfl_xor_sync(0xffffffffu, best, off); const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                        }
                        if (tid == 0) { s_piv = bi; c.ipiv[F.F0 + k] = bi - s0 + 1; }
                    }
                    __syncthreads();
                    const int kp = s_piv;
                    const double pv = S[kp + k * lds];
                    if (pv == 0.0) { if (tid == 0) *c.iflag = -1; }
                    else {
                        if (kp != k) for (int j = s0 + tid; j < R; j += FS_NT) { const double t = S[k + j * lds]; S[k + j * lds] = S[kp + j * lds]; S[kp + j * lds] = t; }
                        __syncthreads();
                        const double inv = 1.0 / S[k + k * lds];
                        for (int i = k + 1 + tid; i < R; i += FS_NT) S[i + k * lds] *= inv;
                    }
                    __syncthreads();
                    for (int j = k + 1 + tid; j < R; j += FS_NT) {
                        const double uj = S[k + j * lds];
                        if (uj != 0.0) for (int i = k + 1; i < R; ++i) S[i + j * lds] -= S[i + k * lds] * uj;
                    }
                } else {
                    const double d = S[k + k * lds];
                    if (d == 0.0 && tid == 0) *c.iflag = -1;
                    for (int i = k + 1 + tid; i < R; i += FS_NT) { const double u = S[i + k * lds]; S[k + i * lds] = u; S[i + k * lds] = u / d; }   // U = D L^T above, L below
                    __syncthreads();
                    for (int j = k + 1 + tid; j < R; j += FS_NT) {
                        const double uj = S[k + j * lds];
                        if (uj != 0.0) for (int i = j; i < R; ++i) S[i + j * lds] -= S[i + k * lds] * uj;
                    }
                }
                __syncthreads();
            }
            s0 = s1;
        }
    }
    for (int e = tid; e < R * R; e += FS_NT) { const int j = e / R, i = e - j * R; G[i + (int64_t)j * F.ld] = S[i + j * lds]; }
}

// ---- explicit inverses of the diagonal blocks (solve) ---------------------------------------------------------------
// The in-block triangular solve is a chain of w dependent steps (~3.5 us of the ~11 us a panel step of a sweep costs).
// After a factorisation k_diag_inverse applies the forward / backward in-block operators of every panel step to the
// identity once; the sweeps then stage that w x w matrix instead of the diagonal block and apply it as a matrix-vector
// product (w independent dot products).  Forward: M_f = L11^{-1} P (LU: interchanges of all pivot sub-blocks folded
// in, so M_f is a general matrix; LDL^T: unit lower triangular).  Backward: M_b = U11^{-1} (LU) or L11^{-T} (LDL^T).
// LDL^T keeps D on the diagonal of both (the callers divide by it); the product treats that diagonal as 1.
__device__ __forceinline__ const double* diag_src(const DevCtx& c, const PStep& ps, bool fwd, int& ld) {
    if (c.tinvf) { ld = ps.w; return (fwd ? c.tinvf : c.tinvb) + ps.iofs; }
    ld = ps.ld;
    return c.F + ps.fofs + (int64_t)ps.o + (int64_t)ps.o * ps.ld;
}
// xs := M xs by one warp (M staged column-major with leading dimension w); UNIT: the diagonal of M counts as 1
template <bool UNIT>
__device__ __forceinline__ void inv_apply_warp(const double* __restrict__ Ms, double* xs, int w) {
    const int lane = threadIdx.x & 31;
    if (w <= 96) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        const int i0 = lane, i1 = lane + 32, i2 = lane + 64;
        for (int k = 0; k < w; ++k) {
            const double xk = xs[k];
            const double* __restrict__ col = Ms + k * w;
            if (i0 < w) a0 += ((UNIT && k == i0) ? 1.0 : col[i0]) * xk;
            if (i1 < w) a1 += ((UNIT && k == i1) ? 1.0 : col[i1]) * xk;
            if (i2 < w) a2 += ((UNIT && k == i2) ? 1.0 : col[i2]) * xk;
        }
        __syncwarp();
        if (i0 < w) xs[i0] = a0; if (i1 < w) xs[i1] = a1; if (i2 < w) xs[i2] = a2;
        __syncwarp();
        return;
    }
    double acc[8];                                          // w <= 256
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.0;
    for (int k = 0; k < w; ++k) {
        const double xk = xs[k];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = lane + 32 * u; if (i < w) acc[u] += ((UNIT && k == i) ? 1.0 : Ms[i + k * w]) * xk; }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = lane + 32 * u; if (i < w) xs[i] = acc[u]; }
    __syncwarp();
}
__device__ int g_solve_dbg = 0;                            // SPK_SOLVE_DBG (timing experiments only): 1 = skip the in-block solves
template <bool LU>
__device__ __forceinline__ void pf_diag_warp(const DevCtx& c, const PStep& ps, const double* Ts, double* xs) {
    const int lane = threadIdx.x & 31, w = ps.w;
    if (g_solve_dbg & 1) return;
    if (c.tinvf) { inv_apply_warp<!LU>(Ts, xs, w); return; }          // Ts holds M_f
    if (w > 96) {                                           // generic shared-memory sweep
        if (LU) {
            const int32_t* ipiv = c.ipiv + ps.col0; const int32_t* subw = c.subw + ps.sub0;
            int s0 = 0;
            for (int b = 0; b < ps.nsub; ++b) {
                const int s1 = s0 + subw[b];
                if (lane == 0) for (int k = s0; k < s1; ++k) { int ip = s0 + ipiv[k] - 1; if (ip != k) { double t = xs[k]; xs[k] = xs[ip]; xs[ip] = t; } }
                __syncwarp();
                for (int k = s0; k < s1; ++k) { const double xk = xs[k]; for (int i = k + 1 + lane; i < w; i += 32) xs[i] -= xk * Ts[i + k * w]; __syncwarp(); }
                s0 = s1;
            }
        } else {
            for (int k = 0; k < w - 1; ++k) { const double xk = xs[k]; for (int i = k + 1 + lane; i < w; i += 32) xs[i] -= xk * Ts[i + k * w]; __syncwarp(); }
        }
        return;
    }
    double x0 = lane < w ? xs[lane] : 0.0, x1 = lane + 32 < w ? xs[lane + 32] : 0.0, x2 = lane + 64 < w ? xs[lane + 64] : 0.0;
    auto sweep = [&](int k) {
        const int slot = k >> 5, src = k & 31;
        const double xk = __shfl_sync(0xffffffffu, slot == 0 ? x0 : (slot == 1 ? x1 : x2), src);
        const double* __restrict__ col = Ts + k * w;
        if (lane > k && lane < w) x0 -= xk * col[lane];
        if (lane + 32 > k && lane + 32 < w) x1 -= xk * col[lane + 32];
        if (lane + 64 > k && lane + 64 < w) x2 -= xk * col[lane + 64];
    };
    if (LU) {
        const int32_t* ipiv = c.ipiv + ps.col0; const int32_t* subw = c.subw + ps.sub0;
        int s0 = 0;
        for (int b = 0; b < ps.nsub; ++b) {
            const int s1 = s0 + subw[b];
            bool any = false;
            for (int k = s0; k < s1; ++k) any |= (s0 + ipiv[k] - 1 != k);
            if (any) {                                      // interchanges of this chunk: through shared memory
                if (lane < w) xs[lane] = x0; if (lane + 32 < w) xs[lane + 32] = x1; if (lane + 64 < w) xs[lane + 64] = x2;
                __syncwarp();
                if (lane == 0) for (int k = s0; k < s1; ++k) { int ip = s0 + ipiv[k] - 1; if (ip != k) { double t = xs[k]; xs[k] = xs[ip]; xs[ip] = t; } }
                __syncwarp();
                x0 = lane < w ? xs[lane] : 0.0; x1 = lane + 32 < w ? xs[lane + 32] : 0.0; x2 = lane + 64 < w ? xs[lane + 64] : 0.0;
            }
            for (int k = s0; k < s1; ++k) sweep(k);
            s0 = s1;
        }
    } else {
        for (int k = 0; k < w - 1; ++k) sweep(k);
    }
    if (lane < w) xs[lane] = x0; if (lane + 32 < w) xs[lane + 32] = x1; if (lane + 64 < w) xs[lane + 64] = x2;
    __syncwarp();
}
// pb_diag: backward in-block solve.  LU: x := inv(U11) y.  LDL^T: x := inv(L11^T) y (y already divided by D).
// Leading dimension of the staged diagonal block in the backward kernels.  The LDL^T in-block solve walks ROWS of
// L11 (L[k,i], i < k): with the compact leading dimension w = 64 all 32 lanes hit one bank (a 32-way conflict on
// every one of the w steps, ~3 us of the ~11 us a backward panel step costs); an odd leading dimension is
// conflict-free for rows and columns alike.  LU (column access) and the inverted blocks (M_b, compact) keep w.
template <bool LU>
__device__ __forceinline__ int pb_tl(const DevCtx& c, int w) { return (LU || c.tinvf) ? w : (w | 1); }
template <bool LU>
__device__ __forceinline__ void pb_diag_warp(const DevCtx& c, const PStep& ps, const double* Ts, double* xs) {
    const int lane = threadIdx.x & 31, w = ps.w, tl = pb_tl<LU>(c, w);
    if (g_solve_dbg & 1) return;
    if (c.tinvf) { inv_apply_warp<!LU>(Ts, xs, w); return; }          // Ts holds M_b
    if (w > 96) {
        for (int k = w - 1; k >= 0; --k) {
            if (LU) { if (lane == 0) xs[k] /= Ts[k + k * tl]; __syncwarp(); }
            const double xk = xs[k];
            if (LU) { for (int i = lane; i < k; i += 32) xs[i] -= xk * Ts[i + k * tl]; }
            else { for (int i = lane; i < k; i += 32) xs[i] -= xk * Ts[k + i * tl]; }
            __syncwarp();
        }
        return;
    }
    double x0 = lane < w ? xs[lane] : 0.0, x1 = lane + 32 < w ? xs[lane + 32] : 0.0, x2 = lane + 64 < w ? xs[lane + 64] : 0.0;
    for (int k = w - 1; k >= 0; --k) {
        const int slot = k >> 5, src = k & 31;
        double xk = __shfl_sync(0xffffffffu, slot == 0 ? x0 : (slot == 1 ? x1 : x2), src);
        if (LU) {
            xk /= Ts[k + k * tl];
            if (lane == src) { if (slot == 0) x0 = xk; else if (slot == 1) x1 = xk; else x2 = xk; }
            const double* __restrict__ col = Ts + k * tl;                    // U[i,k], i < k
            if (lane < k) x0 -= xk * col[lane];
            if (lane + 32 < k) x1 -= xk * col[lane + 32];
            if (lane + 64 < k) x2 -= xk * col[lane + 64];
        } else {
            const double* __restrict__ rowk = Ts + k;                        // L[k,i] = Ts[k + i*tl], i < k (tl odd: conflict-free)
            if (lane < k) x0 -= xk * rowk[lane * tl];
            if (lane + 32 < k) x1 -= xk * rowk[(lane + 32) * tl];
            if (lane + 64 < k) x2 -= xk * rowk[(lane + 64) * tl];
        }
    }
    if (lane < w) xs[lane] = x0; if (lane + 32 < w) xs[lane + 32] = x1; if (lane + 64 < w) xs[lane + 64] = x2;
    __syncwarp();
}

// rows [r0, r1) of the front below the step: wf[r] -= sum_k L[r, o+k] x[k]
// (16 independent loads in flight per thread: the panel is streamed once from HBM, so the sweep is
//  latency-bound unless every thread keeps many requests outstanding)
template <int B = 64>
__device__ __forceinline__ void pf_update_rows(const DevCtx& c, const PStep& ps, double* wf, const double* xs, int r0, int r1) {
    const double* __restrict__ Fm = c.F + ps.fofs + (int64_t)ps.o * ps.ld;
    const int w = ps.w;
    if (w <= 64) {                                          // B entries of the row in flight per pass (64: one round trip to memory)
        for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
            const double* __restrict__ row = Fm + r;
            double acc = 0.0;
#pragma unroll
            for (int h0 = 0; h0 < 64; h0 += B) {
                if (h0 < w) {
                    double v[B];
#pragma unroll
                    for (int u = 0; u < B; ++u) v[u] = __ldcs(row + (size_t)min(h0 + u, w - 1) * ps.ld);     // unconditional: all issued before the first use
#pragma unroll
                    for (int u = 0; u < B; ++u) if (h0 + u < w) acc += v[u] * xs[h0 + u];
                }
            }
            wf[r] -= acc;
        }
        return;
    }
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        double acc = 0.0;
        const double* __restrict__ row = Fm + r;
        for (int k0 = 0; k0 < w; k0 += 16) {
            double v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = (k0 + u < w) ? __ldcs(row + (size_t)(k0 + u) * ps.ld) : 0.0;
#pragma unroll
            for (int u = 0; u < 16; ++u) if (k0 + u < w) acc += v[u] * xs[k0 + u];
        }
        wf[r] -= acc;
    }
}
// asynchronous staging of a w x w block (column-major, leading dimension w in shared memory): issued early,
// waited for with stage_wait() once the block is needed
__device__ __forceinline__ void stage_block_async(double* S, const double* __restrict__ G, int ld, int w, int lds) {
    for (int e = threadIdx.x; e < w * w; e += blockDim.x) {
        const int j = e / w, i = e - j * w;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(S + i + j * lds);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(G + i + (size_t)j * ld));
    }
    asm volatile("cp.async.commit_group;");
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// partial sums over the front indices [r0, r1) beyond the step: out[k] = sum_r coef(r,k) * wf[r]
//   LU: coef = U[o+k, r] (column r of the U panel: w contiguous entries);  LDL^T: coef = L[r, o+k]
// red: shared scratch of (blockDim.x/32) * w doubles; the warps' sums are added in a fixed order.
template <bool LU>
__device__ __forceinline__ void pb_partial(const DevCtx& c, const PStep& ps, const double* wf, int r0, int r1, double* red, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, w = ps.w;
    const double* __restrict__ Fm = c.F + ps.fofs;
    for (int k = threadIdx.x; k < nw * w; k += blockDim.x) red[k] = 0.0;
    __syncthreads();
    if (w <= 64) {
        // U panel (LDL^T fronts hold U = D L^T above the diagonal; the caller divides by D): column r holds its
        // w entries contiguously, so lane k reads U[o+k, r] (coalesced) for 32 columns at a time — all 64 loads
        // in flight — and sums over r in registers: no shuffle reduction
        double acc0 = 0.0, acc1 = 0.0;
        const double* __restrict__ U0 = Fm + (int64_t)ps.o + min(lane, w - 1);
        const double* __restrict__ U1 = Fm + (int64_t)ps.o + min(lane + 32, w - 1);
        for (int rb = r0 + warp * 32; rb < r1; rb += nw * 32) {
            const double xl = rb + lane < r1 ? wf[rb + lane] : 0.0;
            double a0[32], a1[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { const size_t cofs = (size_t)min(rb + j, r1 - 1) * ps.ld; a0[j] = __ldcs(U0 + cofs); a1[j] = __ldcs(U1 + cofs); }
#pragma unroll
            for (int j = 0; j < 32; ++j) { const double xj = __shfl_sync(0xffffffffu, xl, j); acc0 += a0[j] * xj; acc1 += a1[j] * xj; }
        }
        if (lane < w) red[warp * w + lane] = acc0;
        if (lane + 32 < w) red[warp * w + lane + 32] = acc1;
    } else
    for (int rb = r0 + warp * 32; rb < r1; rb += nw * 32) {
        const int r = rb + lane;
        const double xr = r < r1 ? wf[r] : 0.0;
        const double* __restrict__ col = LU ? Fm + (int64_t)ps.o + (int64_t)min(r, r1 - 1) * ps.ld
                                           : Fm + (int64_t)min(r, r1 - 1) + (int64_t)ps.o * ps.ld;
        if (w <= 64) {                                      // the whole row / column in flight: one round trip to memory
            double a[64];
#pragma unroll
            for (int u = 0; u < 64; ++u) { const int uu = min(u, w - 1); a[u] = __ldcs(LU ? col + uu : col + (size_t)uu * ps.ld); }   // unconditional
#pragma unroll
            for (int k0 = 0; k0 < 64; k0 += 8) {
                if (k0 < w) {
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = k0 + u < w ? a[k0 + u] * xr : 0.0;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                        for (int u = 0; u < 8; ++u) v[u] += __shfl_xor_sync(0xffffffffu, v[u], off);
                    if (lane == 0)
#pragma unroll
                        for (int u = 0; u < 8; ++u) if (k0 + u < w) red[warp * w + k0 + u] += v[u];
                }
            }
            continue;
        }
        for (int k0 = 0; k0 < w; k0 += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (k0 + u < w) ? __ldcs(LU ? col + (k0 + u) : col + (size_t)(k0 + u) * ps.ld) * xr : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] += __shfl_xor_sync(0xffffffffu, v[u], off);
            if (lane == 0)
#pragma unroll
                for (int u = 0; u < 8; ++u) if (k0 + u < w) red[warp * w + k0 + u] += v[u];
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < w; k += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < nw; ++q) s += red[q * w + k];
        out[k] = s;
    }
}

// pb_partial for w <= 64 and one 32-row slab per warp (blockDim.x == SV_ROWS), split in two so that the loads
// of the factor entries can be issued before the previous kernel's result is awaited (PDL, see pdl_wait).
// Both factorisations read the U panel (LDL^T fronts hold U = D L^T above the diagonal): column r holds its w
// entries contiguously, lane k reads U[o+k, r] (coalesced) and sums over r in registers — no shuffle tree;
// a[j] / a[32+j] = U[o+lane, rb+j] / U[o+lane+32, rb+j].  LDL^T divides the sums by D afterwards.
template <bool LU>
__device__ __forceinline__ void pb_prefetch64(const DevCtx& c, const PStep& ps, int r0, int r1, double (&a)[64]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, w = ps.w;
    const double* __restrict__ Fm = c.F + ps.fofs;
    const int rb = r0 + warp * 32;
    const double* __restrict__ U0 = Fm + (int64_t)ps.o + min(lane, w - 1);
    const double* __restrict__ U1 = Fm + (int64_t)ps.o + min(lane + 32, w - 1);
#pragma unroll
    for (int j = 0; j < 32; ++j) { const size_t cofs = (size_t)max(min(rb + j, r1 - 1), r0) * ps.ld; a[j] = __ldcs(U0 + cofs); a[32 + j] = __ldcs(U1 + cofs); }
}
template <bool LU>
__device__ __forceinline__ void pb_apply64(const PStep& ps, const double* wf, int r0, int r1, const double (&a)[64], double* red, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, w = ps.w;
    const int rb = r0 + warp * 32;
    const double xl = rb + lane < r1 ? wf[rb + lane] : 0.0;
    {
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int j = 0; j < 32; ++j) { const double xj = __shfl_sync(0xffffffffu, xl, j); acc0 += a[j] * xj; acc1 += a[32 + j] * xj; }
        if (lane < w) red[warp * w + lane] = acc0;
        if (lane + 32 < w) red[warp * w + lane + 32] = acc1;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < w; k += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < nw; ++q) s += red[q * w + k];
        out[k] = s;
    }
}

inline size_t pstep_smem_bytes(int w) { return ((size_t)(w | 1) * w + (size_t)w + 8 * (size_t)w) * sizeof(double); }
// the fused step / front kernels with NR right-hand sides per block: T, NR x (padded) x, NR x next x or the
// (NR x) 8 warps x w partial sums
constexpr int SOLVE_NR = 8;                             // right-hand sides per block when nrhs > 1 (== warps per block)
inline size_t pstep_smem_bytes_mr(int w, int nr, bool front) {
    const size_t xst = w <= 64 ? 64 : (size_t)w;      // stride of one right-hand side's x (zero-padded to 64: unconditional FMAs)
    return ((size_t)(w | 1) * w + (size_t)(2 * nr) * xst + (size_t)(8 * (front ? nr : 1)) * (w + 8) + 8) * sizeof(double);
}

template <bool LU>
__global__ void __launch_bounds__(128) k_pf_diag(DevCtx c, const int32_t* __restrict__ plist) {
    extern __shared__ double ssm[];
    const PStep ps = c.psteps[plist[blockIdx.x]];
    const DFront F = c.fronts[ps.front];
    double* wf = c.w + (size_t)blockIdx.y * c.wlen + F.wofs;
    double* Ts = ssm; double* xs = ssm + ps.w * ps.w;
    { int sld; const double* src = diag_src(c, ps, true, sld); block_g2s<128>(Ts, ps.w, src, sld, ps.w); }
    for (int k = threadIdx.x; k < ps.w; k += blockDim.x) xs[k] = wf[ps.o + k];
    __syncthreads();
    if (threadIdx.x < 32) pf_diag_warp<LU>(c, ps, Ts, xs);
    __syncthreads();
    for (int k = threadIdx.x; k < ps.w; k += blockDim.x) wf[ps.o + k] = xs[k];
}

__global__ void __launch_bounds__(SV_ROWS) k_pf_update(DevCtx c, const int32_t* __restrict__ plist,
                                                       const int32_t* __restrict__ pfx, int count) {
    __shared__ double xs[128];
    int ti = find_task(pfx, count, blockIdx.x);
    int lb = blockIdx.x - pfx[ti];
    const PStep ps = c.psteps[plist[ti]];
    const DFront F = c.fronts[ps.front];
    double* wf = c.w + (size_t)blockIdx.y * c.wlen + F.wofs;
    for (int k = threadIdx.x; k < ps.w; k += blockDim.x) xs[k] = wf[ps.o + k];
    __syncthreads();
    const int e0 = ps.o + ps.w;
    pf_update_rows(c, ps, wf, xs, e0 + lb * SV_ROWS, min(ps.R, e0 + (lb + 1) * SV_ROWS));
}

template <bool LU>
__global__ void __launch_bounds__(SV_ROWS) k_pb_update(DevCtx c, const int32_t* __restrict__ plist,
                                                       const int32_t* __restrict__ pfx, int count, int maxpw) {
    extern __shared__ double ssm[];
    int ti = find_task(pfx, count, blockIdx.x);
    int lb = blockIdx.x - pfx[ti];
    const PStep ps = c.psteps[plist[ti]];
    const DFront F = c.fronts[ps.front];
    const double* wf = c.w + (size_t)blockIdx.y * c.wlen + F.wofs;
    double* out = c.pb + (size_t)blockIdx.y * c.pblen + F.pbofs + (size_t)lb * maxpw;
    const int e0 = ps.o + ps.w;
    pb_partial<LU>(c, ps, wf, e0 + lb * SV_ROWS, min(ps.R, e0 + (lb + 1) * SV_ROWS), ssm, out);
}

template <bool LU>
__global__ void __launch_bounds__(128) k_pb_diag(DevCtx c, const int32_t* __restrict__ plist,
                                                 double* __restrict__ rhs, int64_t ldrhs, int maxpw) {
    extern __shared__ double ssm[];
    const PStep ps = c.psteps[plist[blockIdx.x]];
    const DFront F = c.fronts[ps.front];
    double* wf = c.w + (size_t)blockIdx.y * c.wlen + F.wofs;
    const double* pb = c.pb + (size_t)blockIdx.y * c.pblen + F.pbofs;
    const int tl = pb_tl<LU>(c, ps.w);
    double* Ts = ssm; double* xs = ssm + (ps.w | 1) * ps.w;
    { int sld; const double* src = diag_src(c, ps, false, sld); block_g2s<128>(Ts, tl, src, sld, ps.w); }
    const int nblk = (ps.R - ps.o - ps.w + SV_ROWS - 1) / SV_ROWS;
    __syncthreads();
    for (int k = threadIdx.x; k < ps.w; k += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < nblk; ++q) s += pb[(size_t)q * maxpw + k];
        const double y = wf[ps.o + k];
        xs[k] = LU ? y - s : (ps.w <= 64 ? (y - s) / Ts[k + k * tl] : y / Ts[k + k * tl] - s);   // w <= 64: sums of U = D L^T entries
    }
    __syncthreads();
    if (threadIdx.x < 32) pb_diag_warp<LU>(c, ps, Ts, xs);
    __syncthreads();
    double* out = rhs + (size_t)blockIdx.y * ldrhs + ps.col0;
    for (int k = threadIdx.x; k < ps.w; k += blockDim.x) { wf[ps.o + k] = xs[k]; out[k] = xs[k]; }
}

// Fused forward step: every block updates its rows beyond step j with x_j; block 0 (whose rows contain all
// unknowns of step j+1) then also solves the diagonal block of step j+1, so a sweep needs one launch per
// panel step instead of two.
template <bool LU, int NR>
__global__ void __launch_bounds__(SV_ROWS) k_pf_step(DevCtx c, const int32_t* __restrict__ plist,
                                                     const int32_t* __restrict__ pfx, int count, int nrhs) {
    extern __shared__ double ssm[];
    pdl_trigger();
    int ti = find_task(pfx, count, blockIdx.x);
    int lb = blockIdx.x - pfx[ti];
    const int pid = plist[ti];
    const PStep ps = c.psteps[pid];
    const DFront F = c.fronts[ps.front];
    // NR right-hand sides per block (blockIdx.y = group): the panel row of a thread is read ONCE for all of them
    const int q0 = blockIdx.y * NR, nr = min(NR, nrhs - q0);
    double* wf0 = c.w + (size_t)q0 * c.wlen + F.wofs;           // right-hand side q: wf0 + q * c.wlen
    const int w = ps.w, wp = w <= 64 ? 64 : w;                 // x_j zero-padded to 64 entries: the FMAs below need no guard
    double* xs = ssm;                                           // x_j: NR x wp
    double* Ts = ssm + NR * wp;                                 // block 0: diagonal block of step j+1, staged while the update runs
    const bool next = lb == 0 && pid + 1 < F.ps0 + F.nps;
    PStep nx;
    if (next) {
        nx = c.psteps[pid + 1];
        { int sld; const double* src = diag_src(c, nx, true, sld); stage_block_async(Ts, src, sld, nx.w, nx.w); }
    }
    const int e0 = ps.o + ps.w;
    const int r0 = e0 + lb * SV_ROWS, r1 = min(ps.R, e0 + (lb + 1) * SV_ROWS);
    if (w <= 64) {
        // one row per thread; the row of the panel is fetched BEFORE the previous step's result is awaited
        const int r = r0 + threadIdx.x;
        const double* __restrict__ row = c.F + ps.fofs + (int64_t)ps.o * ps.ld + min(r, r1 - 1);
        double v[64];
#pragma unroll
        for (int u = 0; u < 64; ++u) v[u] = __ldcs(row + (size_t)min(u, w - 1) * ps.ld);
        pdl_wait();
        for (int e = threadIdx.x; e < nr * 64; e += blockDim.x) { const int q = e >> 6, k = e & 63; xs[e] = k < w ? wf0[(size_t)q * c.wlen + ps.o + k] : 0.0; }
        double old[NR];                                         // all right-hand sides' entries of this row in flight together
#pragma unroll
        for (int q = 0; q < NR; ++q) old[q] = (q < nr && r < r1) ? wf0[(size_t)q * c.wlen + r] : 0.0;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            if (q < nr) {
                double acc = 0.0;
#pragma unroll
                for (int u = 0; u < 64; ++u) acc += v[u] * xs[q * 64 + u];
                old[q] -= acc;
            }
        }
#pragma unroll
        for (int q = 0; q < NR; ++q) if (q < nr && r < r1) wf0[(size_t)q * c.wlen + r] = old[q];
    } else {
        pdl_wait();
        for (int q = 0; q < nr; ++q) {
            __syncthreads();
            for (int k = threadIdx.x; k < w; k += blockDim.x) xs[k] = wf0[(size_t)q * c.wlen + ps.o + k];
            __syncthreads();
            pf_update_rows(c, ps, wf0 + (size_t)q * c.wlen, xs, r0, r1);
        }
    }
    if (next) {
        stage_wait();
        __syncthreads();                                        // this block's rows (incl. step j+1's unknowns) are final, T staged
        double* xn = Ts + nx.w * nx.w;                          // NR x nx.w
        for (int e = threadIdx.x; e < nr * nx.w; e += blockDim.x) { const int q = e / nx.w, k = e - q * nx.w; xn[e] = wf0[(size_t)q * c.wlen + nx.o + k]; }
        __syncthreads();
        if ((threadIdx.x >> 5) < nr) pf_diag_warp<LU>(c, nx, Ts, xn + (threadIdx.x >> 5) * nx.w);   // one warp per right-hand side
        __syncthreads();
        for (int e = threadIdx.x; e < nr * nx.w; e += blockDim.x) { const int q = e / nx.w, k = e - q * nx.w; wf0[(size_t)q * c.wlen + nx.o + k] = xn[e]; }
    }
}

// Fused backward step: every block writes the partial sums of its rows/columns beyond step j; the block that
// arrives last (device-wide counter) adds the partials in a fixed order and solves the diagonal block of step j.
template <bool LU, int NR>
__global__ void __launch_bounds__(SV_ROWS, 1) k_pb_step(DevCtx c, const int32_t* __restrict__ plist,
                                                     const int32_t* __restrict__ pfx, int count,
                                                     double* __restrict__ rhs, int64_t ldrhs, int maxpw, int32_t* counters, int nrhs) {
    extern __shared__ double ssm[];
    __shared__ int s_last;
    pdl_trigger();
    int ti = find_task(pfx, count, blockIdx.x);
    int lb = blockIdx.x - pfx[ti];
    const PStep ps = c.psteps[plist[ti]];
    const DFront F = c.fronts[ps.front];
    const int q0 = blockIdx.y * NR, nr = min(NR, nrhs - q0);   // NR right-hand sides per block, factor entries read once
    double* wf0 = c.w + (size_t)q0 * c.wlen + F.wofs;
    double* pb0 = c.pb + (size_t)q0 * c.pblen + F.pbofs;
    const int e0 = ps.o + ps.w, below = ps.R - e0, w = ps.w;
    const int nblk = (below + SV_ROWS - 1) / SV_ROWS;
    const int tl = pb_tl<LU>(c, w);
    double* Ts = ssm; double* xs = ssm + (w | 1) * w;          // xs: NR x w, then 8 x w partial sums
    double* red = xs + NR * w;
    // every block stages the diagonal block while it forms its partial sums: the one that arrives last needs it at once
    { int sld; const double* src = diag_src(c, ps, false, sld); stage_block_async(Ts, src, sld, w, tl); }
    if (nblk > 0) {
        const int r0 = e0 + lb * SV_ROWS, r1 = min(ps.R, e0 + (lb + 1) * SV_ROWS);
        if (w <= 64) {
            double a[64];
            pb_prefetch64<LU>(c, ps, r0, r1, a);            // factor entries in flight before the previous step's x is awaited
            pdl_wait();
            for (int q = 0; q < nr; ++q) {
                pb_apply64<LU>(ps, wf0 + (size_t)q * c.wlen, r0, r1, a, red, pb0 + (size_t)q * c.pblen + (size_t)lb * maxpw);
                __syncthreads();
            }
        } else {
            pdl_wait();
            for (int q = 0; q < nr; ++q) {
                pb_partial<LU>(c, ps, wf0 + (size_t)q * c.wlen, r0, r1, red, pb0 + (size_t)q * c.pblen + (size_t)lb * maxpw);
                __syncthreads();
            }
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            int32_t* cnt = counters + (size_t)blockIdx.y * gridDim.x + pfx[ti];      // one counter per (rhs group, task)
            int prev = atomicAdd(cnt, 1);
            s_last = (prev == nblk - 1);
            if (s_last) *cnt = 0;                                                    // ready for the next sweep
        }
        __syncthreads();
        if (!s_last) { stage_wait(); return; }
        __threadfence();
    } else pdl_wait();
    stage_wait();
    __syncthreads();
    for (int e = threadIdx.x; e < nr * w; e += blockDim.x) {
        const int q = e / w, k = e - q * w;
        const double* pb = pb0 + (size_t)q * c.pblen;
        double sum = 0.0;
        for (int b2 = 0; b2 < nblk; ++b2) sum += __ldcg(pb + (size_t)b2 * maxpw + k);
        const double y = wf0[(size_t)q * c.wlen + ps.o + k];
        xs[e] = LU ? y - sum : (w <= 64 ? (y - sum) / Ts[k + k * tl] : y / Ts[k + k * tl] - sum);     // w <= 64: sums of U = D L^T entries
    }
    __syncthreads();
    if ((threadIdx.x >> 5) < nr) pb_diag_warp<LU>(c, ps, Ts, xs + (threadIdx.x >> 5) * w);              // one warp per right-hand side
    __syncthreads();
    for (int e = threadIdx.x; e < nr * w; e += blockDim.x) {
        const int q = e / w, k = e - q * w;
        wf0[(size_t)q * c.wlen + ps.o + k] = xs[e];
        rhs[(size_t)(q0 + q) * ldrhs + ps.col0 + k] = xs[e];
    }
}



// One block per panel step: applies the in-block forward and backward operators to the identity (thread c owns column
// c of the result, kept in shared memory with an odd leading dimension: conflict-free although the threads walk
// different columns), with the arithmetic of pf_diag_warp / pb_diag_warp.  Output: compact w x w, column-major.
inline size_t inverse_smem_bytes(int w) { return (size_t)2 * w * (w | 1) * sizeof(double); }
template <bool LU>
__global__ void __launch_bounds__(64) k_diag_inverse(DevCtx c, const int32_t* __restrict__ list, double* __restrict__ outf, double* __restrict__ outb) {
    extern __shared__ double ssm[];
    const PStep ps = c.psteps[list[blockIdx.x]];
    if (ps.fofs < 0) return;                               // a front this part holds no storage for
    const int w = ps.w, lp = w | 1;
    double* T = ssm; double* M = ssm + (size_t)w * lp;
    const double* __restrict__ G = c.F + ps.fofs + (int64_t)ps.o + (int64_t)ps.o * ps.ld;
    for (int e = threadIdx.x; e < w * w; e += blockDim.x) { const int i = e % w, j = e / w; T[i + j * lp] = G[i + (int64_t)j * ps.ld]; }
    __syncthreads();
    // ---- forward operator
    for (int col = threadIdx.x; col < w; col += blockDim.x) {
        double* x = M + (size_t)col * lp;
        for (int i = 0; i < w; ++i) x[i] = i == col ? 1.0 : 0.0;
        if (LU) {
            const int32_t* ipiv = c.ipiv + ps.col0; const int32_t* subw = c.subw + ps.sub0;
            int s0 = 0;
            for (int b = 0; b < ps.nsub; ++b) {
                const int s1 = s0 + subw[b];
                for (int k = s0; k < s1; ++k) { const int ip = s0 + ipiv[k] - 1; if (ip != k) { const double t = x[k]; x[k] = x[ip]; x[ip] = t; } }
                for (int k = s0; k < s1; ++k) { const double xk = x[k]; if (xk != 0.0) for (int i = k + 1; i < w; ++i) x[i] -= xk * T[i + k * lp]; }
                s0 = s1;
            }
        } else {
            for (int k = col; k < w - 1; ++k) { const double xk = x[k]; for (int i = k + 1; i < w; ++i) x[i] -= xk * T[i + k * lp]; }
            x[col] = T[col + col * lp];                      // D on the diagonal (the product treats it as 1)
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < w * w; e += blockDim.x) { const int i = e % w, j = e / w; outf[ps.iofs + e] = M[i + j * lp]; }
    __syncthreads();
    // ---- backward operator
    for (int col = threadIdx.x; col < w; col += blockDim.x) {
        double* x = M + (size_t)col * lp;
        for (int i = 0; i < w; ++i) x[i] = i == col ? 1.0 : 0.0;
        if (LU) {
            for (int k = col; k >= 0; --k) { x[k] /= T[k + k * lp]; const double xk = x[k]; for (int i = 0; i < k; ++i) x[i] -= xk * T[i + k * lp]; }
        } else {
            for (int k = col; k >= 1; --k) { const double xk = x[k]; for (int i = 0; i < k; ++i) x[i] -= xk * T[k + i * lp]; }
            x[col] = T[col + col * lp];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < w * w; e += blockDim.x) { const int i = e % w, j = e / w; outb[ps.iofs + e] = M[i + j * lp]; }
}

// ------------------------------------------------------------------------------------
// DATAFLOW sweeps on the fronts with many panel steps (K_PF_FLOW / K_PB_FLOW): one launch per level instead of one
// launch per panel step.  A thread block owns ROWS of a front for the whole sweep — the pivot rows of a few
// consecutive panel steps, or (forward) a slab of the rows below the front's columns — one row per thread, the
// row's running value in a register.  The block that owns step j solves its diagonal block as soon as its rows have
// received the updates of all earlier steps and writes x_j into a MAILBOX that was pre-filled with a sentinel (an
// all-ones NaN); every other block has the panel row of step j in flight already (it depends on the factors only),
// polls the w mailbox words until they are no longer the sentinel — data and "ready" signal in ONE L2 round trip,
// no fence, no flag — and applies the step to its rows.  The chain of dependent steps costs an in-block solve plus
// one L2 round trip per step instead of a kernel launch, and the panels stream from HBM (L2-prefetched two steps
// ahead) behind it.
// Blocks take their task by TICKET (atomic counter): tasks are listed in dependency order, so a block only ever waits
// for blocks that started before it — no co-residency assumption, no deadlock when the grid exceeds the machine.
constexpr int FLOW_NT = 128;
constexpr unsigned long long FLOW_EMPTY = 0xFFFFFFFFFFFFFFFFull;      // mailbox sentinel (cudaMemset 0xFF)
inline size_t flow_smem_bytes(int maxw, int nr) { const size_t wp = maxw <= 64 ? 64 : (size_t)maxw; return (2 * (size_t)(maxw | 1) * maxw + (size_t)nr * wp) * sizeof(double); }

__device__ int g_flow_backoff = 64;                        // ns between polls of a mailbox word (SPK_FLOW_BACKOFF)
__device__ __forceinline__ double flow_poll(const double* p) {
    unsigned long long v;
    const unsigned ns = (unsigned)g_flow_backoff;
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (v != FLOW_EMPTY) break;
        if (ns) __nanosleep(ns);                            // ~100 blocks wait for the same x_j: do not hammer its L2 slice
    }
    return __longlong_as_double((long long)v);
}
__device__ __forceinline__ void flow_post(double* p, double x) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(x) : "memory"); }

// val[u][q] -= sum_k M[r_u, col0 + k] * xs[q][k]  for the rows r_u, u >= U0, of this thread inside [lo, hi)
template <int NR, int RPT, int U0>
__device__ __forceinline__ void flow_apply(const double* __restrict__ Fm, int ld, int col0, int w, int wp, const double* xs,
                                           const int (&row)[RPT], int lo, int hi, double (&val)[RPT][NR], int nr) {
#pragma unroll
    for (int u = U0; u < RPT; ++u) {
        const int r = row[u];
        if (r < lo || r >= hi) continue;
        const double* __restrict__ src = Fm + r + (size_t)col0 * ld;
        for (int k0 = 0; k0 < w; k0 += 16) {
            double v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = __ldcs(src + (size_t)min(k0 + k, w - 1) * ld);
#pragma unroll
            for (int q = 0; q < NR; ++q) {
                if (q >= nr) continue;
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < 16; ++k) acc += (k0 + k < w) ? v[k] * xs[q * wp + k0 + k] : 0.0;
                val[u][q] -= acc;
            }
        }
    }
}
// pull the panel row segments this block will need for the step starting at column `col0` towards L2
__device__ __forceinline__ void flow_prefetch(const double* __restrict__ Fm, int ld, int col0, int w, int r, bool ok) {
    if (ok && (threadIdx.x & 15) == 0) for (int k = 0; k < w; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(Fm + r + (size_t)(col0 + k) * ld));
}

template <bool LU, int NR, int RPT>
__global__ void __launch_bounds__(FLOW_NT) k_pf_flow(DevCtx c, const FlowTask* __restrict__ tasks, int32_t* ticket, double* __restrict__ box,
                                                     int64_t box_stride, int nrhs, int maxw) {
    extern __shared__ double ssm[];
    __shared__ int s_t;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) s_t = atomicAdd(ticket + blockIdx.y, 1);
    __syncthreads();
    const FlowTask t = tasks[s_t];
    const DFront F = c.fronts[t.front];
    const int q0 = blockIdx.y * NR, nr = min(NR, nrhs - q0);
    double* wf0 = c.w + (size_t)q0 * c.wlen + F.wofs;
    double* bx = box + (size_t)q0 * box_stride + F.F0;           // x of front column k, right-hand side q: bx[q * box_stride + k]
    const double* __restrict__ Fm = c.F + F.fofs;
    const int ld = F.ld, wp = maxw <= 64 ? 64 : maxw;
    double* Tb[2] = {ssm, ssm + (size_t)maxw * maxw};
    double* xs = ssm + 2 * (size_t)maxw * maxw;
    const bool pivot = t.jb > t.ja;
    int R0, R1;
    if (pivot) { const PStep a = c.psteps[F.ps0 + t.ja], b = c.psteps[F.ps0 + t.jb - 1]; R0 = a.o; R1 = b.o + b.w; }
    else { R0 = t.r0; R1 = t.r1; }
    int row[RPT]; double val[RPT][NR];
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
        row[u] = R0 + tid + u * FLOW_NT;
        if (row[u] >= R1) row[u] = 1 << 30;
#pragma unroll
        for (int q = 0; q < NR; ++q) val[u][q] = (q < nr && row[u] < R1) ? wf0[(size_t)q * c.wlen + row[u]] : 0.0;
    }
    if (pivot) {                                            // diagonal blocks of my first two steps: staged while the updates arrive
        for (int j = t.ja; j < min(t.jb, t.ja + 2); ++j) { const PStep ps = c.psteps[F.ps0 + j]; int sld; const double* src = diag_src(c, ps, true, sld); stage_block_async(Tb[(j - t.ja) & 1], src, sld, ps.w, ps.w); }
    }
    const int jend = pivot ? t.jb : F.nps;
    PStep ps = c.psteps[F.ps0];
    for (int j = 0; j < jend; ++j) {
        const PStep nx = c.psteps[F.ps0 + min(j + 1, F.nps - 1)];      // next step's record: in flight during this step
        const int w = ps.w, e0 = ps.o + w;
        const bool mine = pivot && j >= t.ja;
        // the thread's panel row of this step (beyond the step's own rows): in flight before x_j is known
        const bool act = row[0] < R1 && row[0] >= e0;
        double v[64];
        if (w <= 64) {
            const double* __restrict__ src = Fm + (act ? row[0] : 0) + (size_t)ps.o * ld;
#pragma unroll
            for (int k = 0; k < 64; ++k) v[k] = act ? __ldcs(src + (size_t)min(k, w - 1) * ld) : 0.0;
        }
        if (j + 2 < jend) { const PStep p2 = c.psteps[F.ps0 + j + 2]; flow_prefetch(Fm, ld, p2.o, p2.w, row[0], row[0] < R1 && row[0] >= p2.o + p2.w); }
        if (mine) {
#pragma unroll
            for (int u = 0; u < RPT; ++u) if (row[u] >= ps.o && row[u] < e0)
#pragma unroll
                for (int q = 0; q < NR; ++q) if (q < nr) xs[q * wp + row[u] - ps.o] = val[u][q];
            stage_wait();
            __syncthreads();
            for (int q = warp; q < nr; q += FLOW_NT / 32) pf_diag_warp<LU>(c, ps, Tb[(j - t.ja) & 1], xs + q * wp);
            __syncthreads();
            for (int e = tid; e < nr * w; e += FLOW_NT) {
                const int q = e / w, k = e - q * w;
                flow_post(bx + (size_t)q * box_stride + ps.o + k, xs[q * wp + k]);
                wf0[(size_t)q * c.wlen + ps.o + k] = xs[q * wp + k];
            }
            if (j + 2 < t.jb) { const PStep p2 = c.psteps[F.ps0 + j + 2]; int sld; const double* src = diag_src(c, p2, true, sld); stage_block_async(Tb[(j - t.ja) & 1], src, sld, p2.w, p2.w); }
        } else {
            for (int e = tid; e < nr * wp; e += FLOW_NT) { const int q = e / wp, k = e - q * wp; xs[e] = k < w ? flow_poll(bx + (size_t)q * box_stride + ps.o + k) : 0.0; }
            __syncthreads();
        }
        if (w <= 64) {
            if (act) {
#pragma unroll
                for (int q = 0; q < NR; ++q) {
                    if (q >= nr) continue;
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < 64; ++k) acc += (k < w) ? v[k] * xs[q * wp + k] : 0.0;
                    val[0][q] -= acc;
                }
            }
            flow_apply<NR, RPT, 1>(Fm, ld, ps.o, w, wp, xs, row, e0, R1, val, nr);
        } else flow_apply<NR, RPT, 0>(Fm, ld, ps.o, w, wp, xs, row, e0, R1, val, nr);
        __syncthreads();                                    // xs is rewritten by the next step
        ps = nx;
    }
    if (!pivot) {
#pragma unroll
        for (int u = 0; u < RPT; ++u) if (row[u] < R1)
#pragma unroll
            for (int q = 0; q < NR; ++q) if (q < nr) wf0[(size_t)q * c.wlen + row[u]] = val[u][q];
    }
}

// Backward: a block owns the pivot rows of steps [ja, jb); steps are solved last to first.  Row r of the U panel of
// step j (LDL^T fronts hold U = D L^T above the diagonal) is F[r, o_j + k]: consecutive threads read consecutive rows.
template <bool LU, int NR, int RPT>
__global__ void __launch_bounds__(FLOW_NT) k_pb_flow(DevCtx c, const FlowTask* __restrict__ tasks, int32_t* ticket, double* __restrict__ box,
                                                     int64_t box_stride, double* __restrict__ rhs, int64_t ldrhs, int nrhs, int maxw) {
    extern __shared__ double ssm[];
    __shared__ int s_t;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) s_t = atomicAdd(ticket + blockIdx.y, 1);
    __syncthreads();
    const FlowTask t = tasks[s_t];
    const DFront F = c.fronts[t.front];
    const int q0 = blockIdx.y * NR, nr = min(NR, nrhs - q0);
    double* wf0 = c.w + (size_t)q0 * c.wlen + F.wofs;
    double* bx = box + (size_t)q0 * box_stride + F.F0;
    const double* __restrict__ Fm = c.F + F.fofs;
    const int ld = F.ld, wp = maxw <= 64 ? 64 : maxw;
    double* Tb[2] = {ssm, ssm + (size_t)(maxw | 1) * maxw};
    double* xs = ssm + 2 * (size_t)(maxw | 1) * maxw;
    const PStep pa = c.psteps[F.ps0 + t.ja], pz = c.psteps[F.ps0 + t.jb - 1];
    const int R0 = pa.o, R1 = pz.o + pz.w;
    int row[RPT]; double val[RPT][NR];
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
        row[u] = R0 + tid + u * FLOW_NT;
        if (row[u] >= R1) row[u] = 1 << 30;
#pragma unroll
        for (int q = 0; q < NR; ++q) val[u][q] = (q < nr && row[u] < R1) ? wf0[(size_t)q * c.wlen + row[u]] : 0.0;
    }
    for (int j = t.jb - 1; j >= max(t.ja, t.jb - 2); --j) { const PStep ps = c.psteps[F.ps0 + j]; int sld; const double* src = diag_src(c, ps, false, sld); stage_block_async(Tb[(t.jb - 1 - j) & 1], src, sld, ps.w, pb_tl<LU>(c, ps.w)); }
    // the unknowns below the front's columns are known (gathered from the parent): their contribution first
    for (int c0 = F.W; c0 < F.R; c0 += wp) {
        const int wc = min(wp, F.R - c0);
        __syncthreads();
        for (int e = tid; e < nr * wp; e += FLOW_NT) { const int q = e / wp, k = e - q * wp; xs[e] = k < wc ? wf0[(size_t)q * c.wlen + c0 + k] : 0.0; }
        __syncthreads();
        flow_apply<NR, RPT, 0>(Fm, ld, c0, wc, wp, xs, row, R0, R1, val, nr);
    }
    __syncthreads();
    PStep ps = c.psteps[F.ps0 + F.nps - 1];
    for (int j = F.nps - 1; j >= t.ja; --j) {
        const PStep nx = c.psteps[F.ps0 + max(j - 1, 0)];
        const int w = ps.w, e0 = ps.o + w;
        const bool mine = j < t.jb;
        const bool act = row[0] < ps.o;                         // my rows above the step (all of them when the step is not mine)
        double v[64];
        if (w <= 64) {
            const double* __restrict__ src = Fm + (act ? row[0] : 0) + (size_t)ps.o * ld;
#pragma unroll
            for (int k = 0; k < 64; ++k) v[k] = act ? __ldcs(src + (size_t)min(k, w - 1) * ld) : 0.0;
        }
        if (j - 2 >= t.ja) { const PStep p2 = c.psteps[F.ps0 + j - 2]; flow_prefetch(Fm, ld, p2.o, p2.w, row[0], row[0] < p2.o); }
        if (mine) {
            double* T = Tb[(t.jb - 1 - j) & 1];
            stage_wait();
            __syncthreads();
#pragma unroll
            for (int u = 0; u < RPT; ++u) if (row[u] >= ps.o && row[u] < e0) {
                const int k = row[u] - ps.o;
#pragma unroll
                for (int q = 0; q < NR; ++q) if (q < nr) xs[q * wp + k] = LU ? val[u][q] : val[u][q] / T[k + k * pb_tl<LU>(c, w)];
            }
            __syncthreads();
            for (int q = warp; q < nr; q += FLOW_NT / 32) pb_diag_warp<LU>(c, ps, T, xs + q * wp);
            __syncthreads();
            for (int e = tid; e < nr * w; e += FLOW_NT) {
                const int q = e / w, k = e - q * w;
                flow_post(bx + (size_t)q * box_stride + ps.o + k, xs[q * wp + k]);
                wf0[(size_t)q * c.wlen + ps.o + k] = xs[q * wp + k];
                rhs[(size_t)(q0 + q) * ldrhs + ps.col0 + k] = xs[q * wp + k];
            }
            if (j - 2 >= t.ja) { const PStep p2 = c.psteps[F.ps0 + j - 2]; int sld; const double* src = diag_src(c, p2, false, sld); stage_block_async(T, src, sld, p2.w, pb_tl<LU>(c, p2.w)); }
        } else {
            for (int e = tid; e < nr * wp; e += FLOW_NT) { const int q = e / wp, k = e - q * wp; xs[e] = k < w ? flow_poll(bx + (size_t)q * box_stride + ps.o + k) : 0.0; }
            __syncthreads();
        }
        if (w <= 64) {
            if (act) {
#pragma unroll
                for (int q = 0; q < NR; ++q) {
                    if (q >= nr) continue;
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < 64; ++k) acc += (k < w) ? v[k] * xs[q * wp + k] : 0.0;
                    val[0][q] -= acc;
                }
            }
            flow_apply<NR, RPT, 1>(Fm, ld, ps.o, w, wp, xs, row, R0, ps.o, val, nr);
        } else flow_apply<NR, RPT, 0>(Fm, ld, ps.o, w, wp, xs, row, R0, ps.o, val, nr);
        __syncthreads();
        ps = nx;
    }
}

// small fronts: one block walks all panel steps of the front, NR right-hand sides at a time (one warp per
// right-hand side in the diagonal solves; every factor entry is read once per block)
// NT = 256 threads, or 64 for the tiny fronts at the bottom of the tree: those blocks are pure latency (a chain of
// dependent loads for a few hundred flops), so eight of them per SM instead of two is what helps.
template <bool LU, int NR, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 8) k_pf_front(DevCtx c, const int32_t* __restrict__ flist, int nrhs) {
    extern __shared__ double ssm[];
    const DFront F = c.fronts[flist[blockIdx.x]];
    const int q0 = blockIdx.y * NR, nr = min(NR, nrhs - q0);
    double* wf0 = c.w + (size_t)q0 * c.wlen + F.wofs;
    const int warp = threadIdx.x >> 5;
    for (int j = 0; j < F.nps; ++j) {
        const PStep ps = c.psteps[F.ps0 + j];
        const int w = ps.w;
        const int xst = (NR > 1 && w <= 64) ? 64 : w;          // x of one right-hand side, zero-padded to 64 (unconditional FMAs)
        double* Ts = ssm; double* xs = ssm + w * w;             // xs: NR x xst
        { int sld; const double* src = diag_src(c, ps, true, sld); block_g2s<NT>(Ts, w, src, sld, w); }
        for (int e = threadIdx.x; e < nr * xst; e += blockDim.x) { const int q = e / xst, k = e - q * xst; xs[e] = k < w ? wf0[(size_t)q * c.wlen + ps.o + k] : 0.0; }
        __syncthreads();
        for (int q = warp; q < nr; q += NT / 32) pf_diag_warp<LU>(c, ps, Ts, xs + q * xst);
        __syncthreads();
        for (int e = threadIdx.x; e < nr * xst; e += blockDim.x) { const int q = e / xst, k = e - q * xst; if (k < w) wf0[(size_t)q * c.wlen + ps.o + k] = xs[e]; }
        if (NR == 1 || w > 64) {
            for (int q = 0; q < nr; ++q) pf_update_rows<32>(c, ps, wf0 + (size_t)q * c.wlen, xs + q * xst, ps.o + w, ps.R);
        } else {
            const double* __restrict__ Fm = c.F + ps.fofs + (int64_t)ps.o * ps.ld;
            for (int r = ps.o + w + threadIdx.x; r < ps.R; r += blockDim.x) {
                double old[NR], acc[NR];
#pragma unroll
                for (int q = 0; q < NR; ++q) { old[q] = q < nr ? wf0[(size_t)q * c.wlen + r] : 0.0; acc[q] = 0.0; }
#pragma unroll
                for (int half = 0; half < 64; half += 32) {     // two passes of 32 entries keep the register count of this
                    if (half < w) {                             // throughput-bound kernel low (several blocks per SM)
                        double v[32];
#pragma unroll
                        for (int u = 0; u < 32; ++u) v[u] = __ldcs(Fm + r + (size_t)min(half + u, w - 1) * ps.ld);
#pragma unroll
                        for (int q = 0; q < NR; ++q)
#pragma unroll
                            for (int u = 0; u < 32; ++u) acc[q] += v[u] * xs[q * 64 + half + u];
                    }
                }
#pragma unroll
                for (int q = 0; q < NR; ++q) if (q < nr) wf0[(size_t)q * c.wlen + r] = old[q] - acc[q];
            }
        }
        __syncthreads();
    }
}

template <bool LU, int NR, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 8) k_pb_front(DevCtx c, const int32_t* __restrict__ flist,
                                                  double* __restrict__ rhs, int64_t ldrhs, int nrhs) {
    extern __shared__ double ssm[];
    const DFront F = c.fronts[flist[blockIdx.x]];
    const int q0 = blockIdx.y * NR, nr = min(NR, nrhs - q0);
    double* wf0 = c.w + (size_t)q0 * c.wlen + F.wofs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = F.nps - 1; j >= 0; --j) {
        const PStep ps = c.psteps[F.ps0 + j];
        const int w = ps.w, e0 = ps.o + w;
        const int tl = pb_tl<LU>(c, w);
        double* Ts = ssm; double* xs = ssm + (w | 1) * w; double* red = xs + NR * w;  // red: NR x 8 warps x w (NR == 1: 8 x w)
        { int sld; const double* src = diag_src(c, ps, false, sld); block_g2s<NT>(Ts, tl, src, sld, w); }
        if (NR == 1 || w > 64) {
            for (int q = 0; q < nr; ++q) { pb_partial<LU>(c, ps, wf0 + (size_t)q * c.wlen, e0, ps.R, red, xs + q * w); __syncthreads(); }
        } else {
            // 32-column slabs of the U panel beyond the step: entries fetched once, applied to every right-hand side
            const double* __restrict__ Fm = c.F + ps.fofs;
            double acc0[NR], acc1[NR];                          // LU: running sums over the slabs (same order as the single-RHS path)
#pragma unroll
            for (int q = 0; q < NR; ++q) { acc0[q] = 0.0; acc1[q] = 0.0; }
            for (int rb = e0 + warp * 32; rb < ps.R; rb += NT) {
                double a[64];
                {
                    const double* __restrict__ U0 = Fm + (int64_t)ps.o + min(lane, w - 1);
                    const double* __restrict__ U1 = Fm + (int64_t)ps.o + min(lane + 32, w - 1);
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) { const size_t cofs = (size_t)min(rb + jj, ps.R - 1) * ps.ld; a[jj] = __ldcs(U0 + cofs); a[32 + jj] = __ldcs(U1 + cofs); }
                }
#pragma unroll
                for (int q = 0; q < NR; ++q) {
                    if (q >= nr) continue;
                    const double xl = rb + lane < ps.R ? wf0[(size_t)q * c.wlen + rb + lane] : 0.0;
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) { const double xj = __shfl_sync(0xffffffffu, xl, jj); acc0[q] += a[jj] * xj; acc1[q] += a[32 + jj] * xj; }
                }
            }
#pragma unroll
            for (int q = 0; q < NR; ++q) {
                if (q >= nr) continue;
                double* rq = red + (size_t)(q * (NT / 32) + warp) * w;
                if (lane < w) rq[lane] = acc0[q];
                if (lane + 32 < w) rq[lane + 32] = acc1[q];
            }
            __syncthreads();
            for (int e = threadIdx.x; e < nr * w; e += blockDim.x) {
                const int q = e / w, k = e - q * w;
                double sum = 0.0;
                for (int wq = 0; wq < NT / 32; ++wq) sum += red[(size_t)(q * (NT / 32) + wq) * w + k];
                xs[e] = sum;
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < nr * w; e += blockDim.x) {
            const int q = e / w, k = e - q * w;
            const double y = wf0[(size_t)q * c.wlen + ps.o + k];
            xs[e] = LU ? y - xs[e] : (w <= 64 ? (y - xs[e]) / Ts[k + k * tl] : y / Ts[k + k * tl] - xs[e]);   // w <= 64: sums of U = D L^T entries
        }
        __syncthreads();
        for (int q = warp; q < nr; q += NT / 32) pb_diag_warp<LU>(c, ps, Ts, xs + q * w);
        __syncthreads();
        for (int e = threadIdx.x; e < nr * w; e += blockDim.x) { const int q = e / w, k = e - q * w; wf0[(size_t)q * c.wlen + ps.o + k] = xs[e]; }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < nr * F.W; e += blockDim.x) { const int q = e / F.W, k = e - q * F.W; rhs[(size_t)(q0 + q) * ldrhs + F.F0 + k] = wf0[(size_t)q * c.wlen + k]; }
}

// forward-only result / backward-only input: copy between rhs and the front vectors
__global__ void k_copy_front_x(DevCtx c, int nfronts, double* __restrict__ rhs, int64_t ldrhs, int to_rhs) {
    int f = blockIdx.x;
    if (f >= nfronts) return;
    const DFront F = c.fronts[f];
    double* w = c.w + (size_t)blockIdx.y * c.wlen + F.wofs;
    double* b = rhs + (size_t)blockIdx.y * ldrhs + F.F0;
    for (int i = threadIdx.x; i < F.W; i += blockDim.x) { if (to_rhs) b[i] = w[i]; else w[i] = b[i]; }
}

// ------------------------------------------------------------------------------------
// Residual / iterative refinement on the device (computeresidual, SpkProblem.jl:448-496; the refinement of
// SpkSparseSpdSolver.jl:267-459 exists only as commented-out Fortran in the reference).
// r = b - A x for a CSR copy of A in the ORIGINAL ordering: one thread per row, terms subtracted in storage order.
__global__ void k_csr_residual(int64_t n, const int64_t* __restrict__ rp, const int32_t* __restrict__ ci,
                               const double* __restrict__ av, const double* __restrict__ b, const double* __restrict__ x,
                               double* __restrict__ r, int64_t ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* xq = x + (size_t)blockIdx.y * ld;
    double acc = b[(size_t)blockIdx.y * ld + i];
    for (int64_t k = rp[i]; k < rp[i + 1]; ++k) acc -= av[k] * xq[ci[k]];
    r[(size_t)blockIdx.y * ld + i] = acc;
}
// sums of squares, one partial per block, fixed summation order (the host adds the partials in order)
__global__ void __launch_bounds__(256) k_sumsq_partial(int64_t n, const double* __restrict__ v, int64_t ld, double* __restrict__ part) {
    __shared__ double sh[256];
    const double* vq = v + (size_t)blockIdx.y * ld;
    const int64_t i0 = (int64_t)blockIdx.x * 4096;
    double acc = 0.0;
    for (int k = 0; k < 16; ++k) { const int64_t i = i0 + threadIdx.x + 256 * k; if (i < n) acc += vq[i] * vq[i]; }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) { if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off]; __syncthreads(); }
    if (threadIdx.x == 0) part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = sh[0];
}
__global__ void k_add_inplace(int64_t n, double* __restrict__ x, const double* __restrict__ d, int64_t ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[(size_t)blockIdx.y * ld + i] += d[(size_t)blockIdx.y * ld + i];
}

// rhs permutation gathers of _triangularsolve! (SpkSparseBase.jl:406-413): out[i] = in[idx[i]-1]
__global__ void k_perm_gather(int64_t n, const int64_t* __restrict__ idx, const double* __restrict__ in,
                              double* __restrict__ out, int64_t ldin, int64_t ldout) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[(size_t)blockIdx.y * ldout + i] = in[(size_t)blockIdx.y * ldin + idx[i] - 1];
}

__global__ void k_ipiv_widen(int64_t n, const int32_t* __restrict__ in, int64_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
__global__ void k_ipiv_narrow(int64_t n, const int64_t* __restrict__ in, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)in[i];
}

} // namespace spk
