// gemm_dmma.cuh — FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) trailing-update kernel.
//
//   C[m x n] -= A[m x k] * B[k x n]        inside one frontal matrix (uniform leading dimension)
//     A = a column block of L (rows contiguous), B = a row block of U (k contiguous).
//     LDL^T fronts keep U = D * L^T in their upper triangle, so both factorisations use this form.
//
// This is the Schur-complement / in-front update of the supernodal factorisation (the
// reference's dgemm('n','t') call sites, SpkLUFactor.jl:152-209, SpkLDLtFactor.jl:145-205).
// tcgen05 has no FP64 kind; on sm_100a every f64 mma shape lowers to DMMA.8x8x4, which is what
// is issued here directly.
//
// PERSISTENT kernel over a TILE LIST built at plan time (plan.hpp: dmma_tiles): the list holds
// only the tiles that have work (LDL^T updates skip everything strictly above the diagonal), a
// launch runs min(#tiles, grid cap) blocks and every block pulls tiles from an atomic counter —
// no empty blocks, no partial last wave per task, and the grid cap can leave SM slots free for
// the latency-critical diagonal / panel kernels that run beside a trailing update.
//
// Operands are fetched with 16-BYTE cp.async.  Panel offsets inside a front have arbitrary
// parity, so a task's origin is moved to the previous even row (sa) / even k (sb): the extra
// row is computed and never stored, the extra k-slice is zeroed in shared memory.  A TM x TN
// tile of C per block, k streamed through a multi-stage shared-memory ring; C is read and
// written once per tile through a shared-memory staged epilogue.
//
// Why not TMA for the operand tiles: the m8n8k4 fragment layout reads A as (8 rows) x (4 k) per
// instruction; with TMA's dense or 128B-swizzled box layouts those fragment loads are 2- to
// 4-way bank-conflicted, the +4-double row padding used here is conflict-free,
// and on the large updates the kernel is bound by the DMMA pipe, not by operand delivery.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "plan.hpp"
#include "kernels.cuh"

namespace spk {

// 16-byte asynchronous copy; `bytes` (0, 8 or 16) are read from global memory, the rest of the 16 is zero-filled
template <bool CA>
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    if (CA) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
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
k_gemm_dmma(DevCtx c, const GemmTask* __restrict__ tasks, const GemmTile* __restrict__ tiles, int ntiles, int32_t* __restrict__ counter, int flags) {
    using Cfg = DmmaCfg<TM, TN, DM_TK, DM_STAGES>;
    constexpr int DM_LDBK = Cfg::LDBK;
    constexpr int NT = WARPS_M * WARPS_N * 32;
    constexpr int WM = TM / WARPS_M, WN = TN / WARPS_N;
    constexpr int FM = WM / 8, FN = WN / 8;
    constexpr int A_CHUNKS = TM * DM_TK / 2, B_CHUNKS = TN * DM_TK / 2;     // 16-byte chunks per stage
    static_assert(A_CHUNKS % NT == 0 && B_CHUNKS % NT == 0 && NT % TM == 0, "loader shapes");
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_tile, s_iter;
    if (threadIdx.x == 0) s_iter = 0;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm0 = (warp % WARPS_M) * WM, wn0 = (warp / WARPS_M) * WN;
    const int lk = lane & 3, lr = lane >> 2;

    pdl_trigger();
    pdl_wait();
    const bool ca = flags & 1;                             // operand tiles through L1 (.ca) or L2 only (.cg)
    if ((flags >> 8) && (blockIdx.x & 1)) {                // de-phase the two co-resident blocks of an SM (flags >> 8 = delay in 256 ns units)
        const unsigned ns = (unsigned)(flags >> 8) * 256u;
        for (unsigned waited = 0; waited < ns; waited += 1000u) __nanosleep(1000u);
    }
    for (;;) {
        if (tid == 0) s_tile = (flags & 2) ? ((int)blockIdx.x + (int)gridDim.x * s_iter) : atomicAdd(counter, 1);
        __syncthreads();                                   // also: the previous tile's epilogue is done with the ring
        const int tile_id = s_tile;
        if (tile_id >= ntiles) break;
        if (tid == 0) ++s_iter;
        const GemmTile tl = tiles[tile_id];
        const GemmTask g = tasks[tl.task];
        // origin moved to the previous even row / even k (16-byte aligned operand chunks)
        const int sa = (int)(g.a0 & 1), sb = (int)(g.b0 & 1);
        const int mp = g.m + sa, kp = g.k + sb, roffp = g.roff - sa;
        const int row0 = (int)tl.ti * TM, col0 = (int)tl.tj * TN;
        const int ld = g.ld;
        const double* __restrict__ A = c.F + (g.a0 - sa) - (int64_t)sb * ld;
        const double* __restrict__ B = c.F + (g.b0 - sb);
        double* __restrict__ C = c.F + (g.c0 - sa);

        // pull the C tile towards L2 while the k loop runs (it is read once, in the epilogue)
        {
            constexpr int LINES_PER_COL = TM / 16;                      // 128-byte lines per tile column
            for (int e = tid; e < TN * LINES_PER_COL; e += NT) {
                const int cl = e / LINES_PER_COL, rl = (e % LINES_PER_COL) * 16;
                const int r = row0 + rl, cc = col0 + cl;
                if (r < mp && cc < g.n && !(g.lower && r + 15 + roffp < cc)) prefetch_l2(C + (size_t)r + (size_t)cc * ld);
            }
        }

        double acc[FM][FN][2];
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

        const int ntk = (kp + DM_TK - 1) / DM_TK;

        auto issue = [&](int kt) {
            const int stage = kt % DM_STAGES, k0 = kt * DM_TK;
            double* As = smem + (size_t)stage * Cfg::STAGE_DOUBLES;
            double* Bs = As + Cfg::A_DOUBLES;
#pragma unroll
            for (int e = tid; e < A_CHUNKS; e += NT) {              // consecutive threads -> consecutive row pairs
                const int r = (e % (TM / 2)) * 2, kk = e / (TM / 2), kg = k0 + kk;
                int valid = mp - (row0 + r); valid = valid < 0 ? 0 : (valid > 2 ? 2 : valid);
                if (kg >= kp || kg < sb) valid = 0;                 // past the end, or the zeroed k-slice in front of an odd origin
                const double* src = valid ? A + (size_t)(row0 + r) + (size_t)kg * ld : A;
                if (ca) cp_async16<true>(As + kk * Cfg::LDA + r, src, valid * 8); else cp_async16<false>(As + kk * Cfg::LDA + r, src, valid * 8);
            }
#pragma unroll
            for (int e = tid; e < B_CHUNKS; e += NT) {              // consecutive threads -> consecutive k pairs
                const int kk = (e % (DM_TK / 2)) * 2, r = e / (DM_TK / 2), kg = k0 + kk;
                int valid = kp - kg; valid = valid < 0 ? 0 : (valid > 2 ? 2 : valid);
                if (col0 + r >= g.n) valid = 0;
                const double* src = valid ? B + (size_t)kg + (size_t)(col0 + r) * ld : B;
                if (ca) cp_async16<true>(Bs + r * DM_LDBK + kk, src, valid * 8); else cp_async16<false>(Bs + r * DM_LDBK + kk, src, valid * 8);
            }
        };

#pragma unroll
        for (int s = 0; s < DM_STAGES - 1; ++s) {
            if (s < ntk) issue(s);
            cp_async_commit();
        }
        for (int kt = 0; kt < ntk; ++kt) {
            cp_async_wait<DM_STAGES - 2>();
            __syncthreads();
            if (kt + DM_STAGES - 1 < ntk) issue(kt + DM_STAGES - 1);   // refill the slot freed last iteration
            cp_async_commit();

            const int stage = kt % DM_STAGES;
            double* As = smem + (size_t)stage * Cfg::STAGE_DOUBLES;
            double* Bs = As + Cfg::A_DOUBLES;
            if (kt == 0 && sb) {                                    // odd k origin: B's slice k' = 0 is not part of the product
                if (tid < TN) Bs[tid * DM_LDBK] = 0.0;              // (A's is zero-filled by the loader; 0 * 0, never 0 * junk)
                __syncthreads();
            }
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
        const bool rok = r < mp && r >= sa;                // the row in front of an odd origin belongs to somebody else
#pragma unroll 1
        for (int p0 = 0; p0 < NPASS; p0 += EPI_U) {
            double cv[EPI_U];
            bool ok[EPI_U];
#pragma unroll
            for (int u = 0; u < EPI_U; ++u) {
                const int cl = (p0 + u) * COLS_PER_PASS + ec, cc = col0 + cl;
                ok[u] = rok && cc < g.n && !(g.lower && r + roffp < cc);
                cv[u] = ok[u] ? __ldcg(C + (size_t)r + (size_t)cc * ld) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < EPI_U; ++u) {
                const int cl = (p0 + u) * COLS_PER_PASS + ec, cc = col0 + cl;
                if (ok[u]) __stcg(C + (size_t)r + (size_t)cc * ld, cv[u] - Cs[cl * Cfg::LDC + er]);
            }
        }
    }
}

// Tile shapes.  K_GEMM_T64 = 64 x 64 tiles, 4 warps, four co-resident blocks per SM, k streamed 8 at a time through a
// 4-stage ring: the DEFAULT for every DMMA launch.  Measured on the delayed updates' own shapes (tools/ubench_dmma.cu,
// lower-triangular m = n = 9000, k = 456): 31.1 TFLOP/s against 27.7 for the 128 x 64 / 8-warp / 2-per-SM shape, 26.5
// against 23.4 at m = n = 4000 — four small independent blocks keep the DMMA pipe fed through each other's barriers,
// prologues and epilogues better than two large ones, and short k-steps shorten the fill of a k = 456 tile.  Larger
// block or warp tiles (128 x 128, 256 x 64, 64 x 32 per warp) were all slower.  K_GEMM_B64 = 128 x 64 tiles, 8 warps,
// two blocks per SM (SPK_DMMA_BIG=1 selects it for launches with enough tiles; SPK_DMMA_VARIANT = its pipeline shape).
// Launches with fewer than Plan::dmma_narrow 64 x 64 tiles (the in-block updates of the top fronts: n = 64, ~200 tiles on 148
// SMs) run 64 x 32 tiles, 4 warps of 32 x 16: twice the blocks, 37 vs 43 us at 13000 x 64 x 400.
using GemmKernel = void (*)(DevCtx, const GemmTask*, const GemmTile*, int, int32_t*, int);
struct GemmVariant { GemmKernel fn; int threads; size_t smem; int blocks_per_sm; };
inline GemmVariant gemm_dmma_variant(int kind, int variant, int variant64 = 8, int tile_n = 64) {
    if (kind == K_GEMM_T64 && tile_n == 32) return {k_gemm_dmma<64, 32, 2, 2, 4, 8, 4>, 128, DmmaCfg<64, 32, 8, 4>::SMEM, 4};   // under-filled launches
    if (kind == K_GEMM_T64) {
        if (variant64 == 4) return {k_gemm_dmma<64, 64, 2, 2, 4, 16, 3>, 128, DmmaCfg<64, 64, 16, 3>::SMEM, 4};
        return {k_gemm_dmma<64, 64, 2, 2, 4, 8, 4>, 128, DmmaCfg<64, 64, 8, 4>::SMEM, 4};
    }
    if (variant == 3) return {k_gemm_dmma<BIG_TM, 64, 4, 2, 2, 32, 2>, 256, DmmaCfg<BIG_TM, 64, 32, 2>::SMEM, 2};
    if (variant == 5) return {k_gemm_dmma<BIG_TM, 64, 4, 2, 2, 8, 6>, 256, DmmaCfg<BIG_TM, 64, 8, 6>::SMEM, 2};
    if (variant == 6) return {k_gemm_dmma<BIG_TM, 64, 4, 2, 2, 16, 4>, 256, DmmaCfg<BIG_TM, 64, 16, 4>::SMEM, 2};
    return {k_gemm_dmma<BIG_TM, 64, 4, 2, 2, 16, 3>, 256, DmmaCfg<BIG_TM, 64, 16, 3>::SMEM, 2};
}
inline cudaError_t gemm_dmma_init() {
    for (int kind : {(int)K_GEMM_B64, (int)K_GEMM_T64})
        for (int variant : {3, 4, 5, 6})
            for (int variant64 : {4, 8, 32}) {
                GemmVariant v = gemm_dmma_variant(kind, variant, variant64 == 32 ? 8 : variant64, variant64 == 32 ? 32 : 64);
                cudaError_t e = cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem);
                if (e != cudaSuccess) return e;
            }
    return cudaSuccess;
}

} // namespace spk
