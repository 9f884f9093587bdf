// gemm_dmma.cuh — FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) trailing-update kernel.
//
//   C[m x n] -= A[m x k] * B        inside one frontal matrix (uniform leading dimension)
//     LU   : B = a row block of U           (B(k,n) contiguous along k)
//     LDLT : B = (L21' * diag(D))^T         (B(n,k) contiguous along n, scaled on the fly)
//
// This is the Schur-complement / in-front update of the supernodal factorisation (the
// reference's dgemm('n','t') call sites, SpkLUFactor.jl:152-209, SpkLDLtFactor.jl:145-205).
// One thread block owns a TM x TN tile of C and streams k through a multi-stage cp.async
// shared-memory ring, so C is read and written once per task.  tcgen05 has no FP64 kind; on
// sm_100a every f64 mma shape lowers to DMMA.8x8x4, which is what is issued here directly.
// Operand bases are only 8-byte aligned in general (arbitrary panel offsets inside a front),
// hence 8-byte cp.async.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "plan.hpp"
#include "kernels.cuh"

namespace spk {

constexpr int DM_TK = 16;            // k per pipeline stage
constexpr int DM_STAGES = 4;
constexpr int DM_THREADS = 256;
constexpr int DM_LDBK = DM_TK + 4;   // row stride of the k-contiguous B tile (conflict-free fragment loads)

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int TM, int TN>
struct DmmaCfg {
    static constexpr int LDA = TM + 4, LDB = TN + 4;               // +4 doubles: conflict-free fragment loads
    static constexpr int A_DOUBLES = DM_TK * LDA;
    static constexpr int B_DOUBLES = (DM_TK * LDB > TN * DM_LDBK) ? DM_TK * LDB : TN * DM_LDBK;
    static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES + DM_TK;   // A tile, B tile, D slice
    static constexpr size_t SMEM = (size_t)DM_STAGES * STAGE_DOUBLES * sizeof(double);
};

// WARPS_M x WARPS_N warps, each owning a (TM/WARPS_M) x (TN/WARPS_N) sub-tile.
// BK: B is k-contiguous (LU) instead of n-contiguous + D-scaled (LDL^T).
template <int TM, int TN, int WARPS_M, int WARPS_N, bool BK>
__global__ void __launch_bounds__(DM_THREADS, 1)
k_gemm_dmma(DevCtx c, const GemmTask* __restrict__ tasks, const int32_t* __restrict__ pfx, int count) {
    static_assert(WARPS_M * WARPS_N * 32 == DM_THREADS, "warp layout");
    using Cfg = DmmaCfg<TM, TN>;
    constexpr int WM = TM / WARPS_M, WN = TN / WARPS_N;
    constexpr int FM = WM / 8, FN = WN / 8;
    extern __shared__ double smem[];

    int t = find_task(pfx, count, blockIdx.x);
    int lb = blockIdx.x - pfx[t];
    const GemmTask g = tasks[t];
    const int mt = (g.m + TM - 1) / TM;
    const int row0 = (lb % mt) * TM, col0 = (lb / mt) * TN;
    if (g.lower && row0 + TM - 1 + g.roff < col0) return;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm0 = (warp % WARPS_M) * WM, wn0 = (warp / WARPS_M) * WN;
    const int lk = lane & 3, lr = lane >> 2;
    const int ld = g.ld;
    const double* __restrict__ A = c.F + g.a0;
    const double* __restrict__ B = c.F + g.b0;
    const double* __restrict__ D = c.F + g.d0;

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    const int ntiles = (g.k + DM_TK - 1) / DM_TK;

    auto issue = [&](int tile) {
        const int stage = tile % DM_STAGES, k0 = tile * DM_TK;
        double* As = smem + (size_t)stage * Cfg::STAGE_DOUBLES;
        double* Bs = As + Cfg::A_DOUBLES;
        double* Ds = Bs + Cfg::B_DOUBLES;
#pragma unroll
        for (int e = tid; e < TM * DM_TK; e += DM_THREADS) {       // consecutive threads -> consecutive rows
            int r = e % TM, kk = e / TM, kg = k0 + kk;
            bool ok = (row0 + r < g.m) && (kg < g.k);
            const double* src = ok ? A + (size_t)(row0 + r) + (size_t)kg * ld : A;
            cp_async8(As + kk * Cfg::LDA + r, src, ok);
        }
        if (BK) {
#pragma unroll
            for (int e = tid; e < TN * DM_TK; e += DM_THREADS) {   // consecutive threads -> consecutive k
                int kk = e % DM_TK, r = e / DM_TK, kg = k0 + kk;
                bool ok = (col0 + r < g.n) && (kg < g.k);
                const double* src = ok ? B + (size_t)kg + (size_t)(col0 + r) * ld : B;
                cp_async8(Bs + r * DM_LDBK + kk, src, ok);
            }
        } else {
#pragma unroll
            for (int e = tid; e < TN * DM_TK; e += DM_THREADS) {
                int r = e % TN, kk = e / TN, kg = k0 + kk;
                bool ok = (col0 + r < g.n) && (kg < g.k);
                const double* src = ok ? B + (size_t)(col0 + r) + (size_t)kg * ld : B;
                cp_async8(Bs + kk * Cfg::LDB + r, src, ok);
            }
            if (tid < DM_TK) {
                int kg = k0 + tid;
                bool ok = kg < g.k;
                const double* src = ok ? D + (size_t)kg * (ld + 1) : D;
                cp_async8(Ds + tid, src, ok);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < DM_STAGES - 1; ++s) {
        if (s < ntiles) issue(s);
        cp_async_commit();
    }
    for (int tile = 0; tile < ntiles; ++tile) {
        cp_async_wait<DM_STAGES - 2>();
        __syncthreads();
        if (tile + DM_STAGES - 1 < ntiles) issue(tile + DM_STAGES - 1);   // refill the slot freed last iteration
        cp_async_commit();

        const int stage = tile % DM_STAGES;
        const double* As = smem + (size_t)stage * Cfg::STAGE_DOUBLES;
        const double* Bs = As + Cfg::A_DOUBLES;
        const double* Ds = Bs + Cfg::B_DOUBLES;
#pragma unroll
        for (int k4 = 0; k4 < DM_TK / 4; ++k4) {
            const int kk = k4 * 4 + lk;
            double a[FM], b[FN];
#pragma unroll
            for (int i = 0; i < FM; ++i) a[i] = As[kk * Cfg::LDA + wm0 + i * 8 + lr];
            if (BK) {
#pragma unroll
                for (int j = 0; j < FN; ++j) b[j] = Bs[(wn0 + j * 8 + lr) * DM_LDBK + kk];
            } else {
                const double d = Ds[kk];
#pragma unroll
                for (int j = 0; j < FN; ++j) b[j] = Bs[kk * Cfg::LDB + wn0 + j * 8 + lr] * d;
            }
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();                                   // every warp is done reading the last stage

    // Epilogue: C -= acc, staged through shared memory (the pipeline ring is free now).
    //  1. accumulator fragments -> Cs[col][row]      (fragment (i,j): rows wm0+8i+lr, cols wn0+8j+2*lk+{0,1})
    //  2. coalesced read-modify-write of C: consecutive threads own consecutive rows of one column, with
    //     EPI_U independent loads in flight per thread.  (A direct `C[..] -= acc` from the fragments keeps
    //     ~64 dependent, 64-byte-granular global round trips per thread on the critical path: ncu showed the
    //     tensor pipe idle for longer than the whole k loop, all warps in long-scoreboard stalls on the DADDs.)
    constexpr int LDC = TM + 4;
    static_assert((size_t)TN * LDC * sizeof(double) <= Cfg::SMEM, "C tile must fit in the pipeline ring");
    double* Cs = smem;
#pragma unroll
    for (int j = 0; j < FN; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < FM; ++i)
                Cs[(wn0 + j * 8 + 2 * lk + h) * LDC + wm0 + i * 8 + lr] = acc[i][j][h];
    __syncthreads();

    double* __restrict__ C = c.F + g.c0;
    constexpr int COLS_PER_PASS = DM_THREADS / TM;     // 2 columns per pass with 256 threads
    constexpr int EPI_U = 8;
    const int er = tid % TM, ec = tid / TM;
    const int r = row0 + er;
    const bool rok = r < g.m;
#pragma unroll 1
    for (int p0 = 0; p0 < TN / COLS_PER_PASS; p0 += EPI_U) {
        double cv[EPI_U];
        bool ok[EPI_U];
#pragma unroll
        for (int u = 0; u < EPI_U; ++u) {
            const int cl = (p0 + u) * COLS_PER_PASS + ec, cc = col0 + cl;
            ok[u] = rok && cc < g.n && !(g.lower && r + g.roff < cc);
            cv[u] = ok[u] ? __ldcg(C + (size_t)r + (size_t)cc * ld) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < EPI_U; ++u) {
            const int cl = (p0 + u) * COLS_PER_PASS + ec, cc = col0 + cl;
            if (ok[u]) __stcg(C + (size_t)r + (size_t)cc * ld, cv[u] - Cs[cl * LDC + er]);
        }
    }
}

using GemmKernel = void (*)(DevCtx, const GemmTask*, const int32_t*, int);
inline GemmKernel gemm_dmma_kernel(int kind, bool lu) {
    if (kind == K_GEMM_B128) return lu ? k_gemm_dmma<BIG_TM, 128, 2, 4, true> : k_gemm_dmma<BIG_TM, 128, 2, 4, false>;
    return lu ? k_gemm_dmma<BIG_TM, 64, 4, 2, true> : k_gemm_dmma<BIG_TM, 64, 4, 2, false>;
}
inline size_t gemm_dmma_smem(int kind) {
    return kind == K_GEMM_B128 ? DmmaCfg<BIG_TM, 128>::SMEM : DmmaCfg<BIG_TM, 64>::SMEM;
}
inline cudaError_t gemm_dmma_init() {
    cudaError_t e;
    e = cudaFuncSetAttribute(k_gemm_dmma<BIG_TM, 128, 2, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DmmaCfg<BIG_TM, 128>::SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_gemm_dmma<BIG_TM, 128, 2, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DmmaCfg<BIG_TM, 128>::SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_gemm_dmma<BIG_TM, 64, 4, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DmmaCfg<BIG_TM, 64>::SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_gemm_dmma<BIG_TM, 64, 4, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DmmaCfg<BIG_TM, 64>::SMEM);
}

} // namespace spk
