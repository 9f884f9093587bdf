// gemm_dmma.cuh — FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) trailing-update kernel.
//
//   C[m x n] -= A[m x k] * B[k x n]        inside one frontal matrix (uniform leading dimension)
//     A = a column block of L (rows contiguous), B = a row block of U (k contiguous).
//     LDL^T fronts keep U = D * L^T in their upper triangle, so both factorisations use this form.
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

// Pipeline shape is a template parameter pair: TK = k per stage, STAGES = ring depth.

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int TM, int TN, int DM_TK = 16, int DM_STAGES = 4>
struct DmmaCfg {
    static constexpr int LDA = TM + 4;                              // +4 doubles: conflict-free fragment loads
    static constexpr int LDC = TM + 4;
    static constexpr int LDBK = DM_TK + 4;                          // row stride of the k-contiguous B tile
    static constexpr int A_DOUBLES = DM_TK * LDA;
    static constexpr int B_DOUBLES = TN * LDBK;
    static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
    static constexpr int RING_DOUBLES = DM_STAGES * STAGE_DOUBLES;
    static constexpr int C_DOUBLES = TN * LDC;
    static constexpr size_t SMEM = sizeof(double) * (size_t)(RING_DOUBLES > C_DOUBLES ? RING_DOUBLES : C_DOUBLES);
};

// WARPS_M x WARPS_N warps, each owning a (TM/WARPS_M) x (TN/WARPS_N) sub-tile of the block's C tile.
template <int TM, int TN, int WARPS_M, int WARPS_N, int MINB, int DM_TK = 16, int DM_STAGES = 4>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, MINB)
k_gemm_dmma(DevCtx c, const GemmTask* __restrict__ tasks, const int32_t* __restrict__ pfx, int count) {
    using Cfg = DmmaCfg<TM, TN, DM_TK, DM_STAGES>;
    constexpr int DM_LDBK = Cfg::LDBK;
    constexpr int NT = WARPS_M * WARPS_N * 32;
    constexpr int WM = TM / WARPS_M, WN = TN / WARPS_N;
    constexpr int FM = WM / 8, FN = WN / 8;
    static_assert(NT % TM == 0 && (TN * DM_TK) % NT == 0 && (TM * DM_TK) % NT == 0, "loader shapes");
    extern __shared__ double smem[];

    pdl_trigger();
    int t = find_task(pfx, count, blockIdx.x);
    int lb = blockIdx.x - pfx[t];
    const GemmTask g = tasks[t];
    const int mt = (g.m + TM - 1) / TM;
    const int row0 = (lb % mt) * TM, col0 = (lb / mt) * TN;
    pdl_wait();
    if (g.lower && row0 + TM - 1 + g.roff < col0) return;          // tile strictly above the diagonal

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm0 = (warp % WARPS_M) * WM, wn0 = (warp / WARPS_M) * WN;
    const int lk = lane & 3, lr = lane >> 2;
    const int ld = g.ld;
    const double* __restrict__ A = c.F + g.a0;
    const double* __restrict__ B = c.F + g.b0;
    double* __restrict__ C = c.F + g.c0;

    // pull the C tile towards L2 while the k loop runs (it is read once, in the epilogue)
    {
        constexpr int LINES_PER_COL = TM / 16;                      // 128-byte lines per tile column
        for (int e = tid; e < TN * LINES_PER_COL; e += NT) {
            const int cl = e / LINES_PER_COL, rl = (e % LINES_PER_COL) * 16;
            const int r = row0 + rl, cc = col0 + cl;
            if (r < g.m && cc < g.n && !(g.lower && r + 15 + g.roff < cc)) prefetch_l2(C + (size_t)r + (size_t)cc * ld);
        }
    }

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
#pragma unroll
        for (int e = tid; e < TM * DM_TK; e += NT) {                // consecutive threads -> consecutive rows
            int r = e % TM, kk = e / TM, kg = k0 + kk;
            bool ok = (row0 + r < g.m) && (kg < g.k);
            const double* src = ok ? A + (size_t)(row0 + r) + (size_t)kg * ld : A;
            cp_async8(As + kk * Cfg::LDA + r, src, ok);
        }
#pragma unroll
        for (int e = tid; e < TN * DM_TK; e += NT) {                // consecutive threads -> consecutive k
            int kk = e % DM_TK, r = e / DM_TK, kg = k0 + kk;
            bool ok = (col0 + r < g.n) && (kg < g.k);
            const double* src = ok ? B + (size_t)kg + (size_t)(col0 + r) * ld : B;
            cp_async8(Bs + r * DM_LDBK + kk, src, ok);
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
#pragma unroll
        for (int k4 = 0; k4 < DM_TK / 4; ++k4) {
            const int kk = k4 * 4 + lk;
            double a[FM], b[FN];
#pragma unroll
            for (int i = 0; i < FM; ++i) a[i] = As[kk * Cfg::LDA + wm0 + i * 8 + lr];
#pragma unroll
            for (int j = 0; j < FN; ++j) b[j] = Bs[(wn0 + j * 8 + lr) * DM_LDBK + kk];
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
    double* Cs = smem;
#pragma unroll
    for (int j = 0; j < FN; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < FM; ++i)
                Cs[(wn0 + j * 8 + 2 * lk + h) * Cfg::LDC + wm0 + i * 8 + lr] = acc[i][j][h];
    __syncthreads();

    constexpr int COLS_PER_PASS = NT / TM;
    constexpr int NPASS = TN / COLS_PER_PASS;
    constexpr int EPI_U = NPASS < 16 ? NPASS : 16;
    const int er = tid % TM, ec = tid / TM;
    const int r = row0 + er;
    const bool rok = r < g.m;
#pragma unroll 1
    for (int p0 = 0; p0 < NPASS; p0 += EPI_U) {
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
            if (ok[u]) __stcg(C + (size_t)r + (size_t)cc * ld, cv[u] - Cs[cl * Cfg::LDC + er]);
        }
    }
}

// Kernel variants (SPK_DMMA_VARIANT): 0 = 8 warps, one block per SM; 1 = 16 warps per block;
// 2 = 128x64 tiles only, two co-resident blocks per SM (one block's epilogue overlaps the other's k loop).
using GemmKernel = void (*)(DevCtx, const GemmTask*, const int32_t*, int);
struct GemmVariant { GemmKernel fn; int threads; size_t smem; };
inline GemmVariant gemm_dmma_variant(int kind, int variant) {
    if (kind == K_GEMM_B128) {
        if (variant == 1) return {k_gemm_dmma<BIG_TM, 128, 4, 4, 1>, 512, DmmaCfg<BIG_TM, 128>::SMEM};
        return {k_gemm_dmma<BIG_TM, 128, 2, 4, 1>, 256, DmmaCfg<BIG_TM, 128>::SMEM};
    }
    if (variant == 1) return {k_gemm_dmma<BIG_TM, 64, 4, 4, 1>, 512, DmmaCfg<BIG_TM, 64>::SMEM};
    if (variant == 2) return {k_gemm_dmma<BIG_TM, 64, 4, 2, 2>, 256, DmmaCfg<BIG_TM, 64>::SMEM};
    if (variant == 3) return {k_gemm_dmma<BIG_TM, 64, 4, 2, 2, 32, 2>, 256, DmmaCfg<BIG_TM, 64, 32, 2>::SMEM};
    if (variant == 4) return {k_gemm_dmma<BIG_TM, 64, 4, 2, 2, 16, 3>, 256, DmmaCfg<BIG_TM, 64, 16, 3>::SMEM};
    if (variant == 5) return {k_gemm_dmma<BIG_TM, 64, 4, 2, 2, 8, 6>, 256, DmmaCfg<BIG_TM, 64, 8, 6>::SMEM};
    if (variant == 6) return {k_gemm_dmma<BIG_TM, 64, 4, 2, 2, 16, 2>, 256, DmmaCfg<BIG_TM, 64, 16, 2>::SMEM};
    if (variant == 7) return {k_gemm_dmma<BIG_TM, 64, 4, 2, 2, 8, 4>, 256, DmmaCfg<BIG_TM, 64, 8, 4>::SMEM};
    if (variant == 8) return {k_gemm_dmma<BIG_TM, 64, 4, 2, 2, 8, 3>, 256, DmmaCfg<BIG_TM, 64, 8, 3>::SMEM};
    return {k_gemm_dmma<BIG_TM, 64, 4, 2, 1>, 256, DmmaCfg<BIG_TM, 64>::SMEM};
}
inline cudaError_t gemm_dmma_init() {
    for (int kind : {(int)K_GEMM_B64, (int)K_GEMM_B128})
        for (int variant = 0; variant < 9; ++variant) {
            GemmVariant v = gemm_dmma_variant(kind, variant);
            cudaError_t e = cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem);
            if (e != cudaSuccess) return e;
        }
    return cudaSuccess;
}

} // namespace spk
