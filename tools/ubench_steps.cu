// Micro-benchmark of the critical-path kernels of one panel step (diag block, panel) on a synthetic front.
// Development tool only: timings guide kernel design; never a bench number.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I sparspak.jl_b200/csrc tools/ubench_steps.cu -o tools/ubench_steps
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "plan.hpp"
#include "kernels.cuh"
using namespace spk;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

// experimental variants of the row kernel: what does each part of the dependent chain cost?
// NS = threads per row (1, 2, 4): thread NS*r+h owns the column pairs P = h (mod NS)
template <int VAR, int NS>
__global__ void __launch_bounds__(64 * NS) k_diag_var(DevCtx c, const int32_t* __restrict__ pslist, long long* clk) {
    constexpr int WP = 64, NE = WP / NS, PB = 8 / NS;       // entries per thread, entries per 8-column block
    __shared__ __align__(16) double col[2][2 * WP];
    const PStep ps = c.psteps[pslist[blockIdx.x]];
    double* G = c.F + ps.fofs + (int64_t)ps.o + (int64_t)ps.o * ps.ld;
    const int w = ps.w, ld = ps.ld;
    const int r = threadIdx.x / NS, h = threadIdx.x % NS;
    long long t0 = clock64();
    double a[NE];
#pragma unroll
    for (int li = 0; li < NE; ++li) {
        const int j = ((li >> 1) * NS + h) * 2 + (li & 1);
        a[li] = (j <= r && r < w) ? __ldcg(G + r + (size_t)j * ld) : 0.0;
    }
    for (int i = threadIdx.x; i < 2 * WP; i += 64 * NS) col[i >> 6][WP + (i & 63)] = 0.0;
    long long t1 = 0;
#pragma unroll 1
    for (int kb = 0; kb < w; kb += 8) {
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const int k = kb + cc;
            if (k < w) {
                double* cb = col[k & 1];
                const bool owner = h == ((cc >> 1) % NS);
                const int lc = ((cc >> 1) / NS) * 2 + (cc & 1);         // local index of column cc
                if (owner && r >= k) cb[r] = a[lc];
                if (VAR != 2) __syncthreads();
                if (kb == 0 && cc == 0) t1 = clock64();
                const double d = cb[k];
                double l;
                if (VAR == 1) l = cb[r] * (d * 0.5);
                else if (VAR == 4) l = cb[r] / d;
                else l = cb[r] * fast_rcp(d);
                if (owner && r >= k && r < w) {
                    if (r == k) { __stcg(G + r + (size_t)k * ld, d); if (d == 0.0) atomicExch(c.iflag, -1); }
                    else __stcg(G + r + (size_t)k * ld, l);
                }
                const double* cj = cb + kb + h * 2;
                if (VAR != 3) {
#pragma unroll
                    for (int li = 0; li < NE; ++li) a[li] -= l * cj[(li >> 1) * 2 * NS + (li & 1)];
                } else a[(cc + 1) & 1] -= l * cj[(cc + 1) & 1];
            }
        }
#pragma unroll
        for (int li = 0; li < NE - PB; ++li) a[li] = a[li + PB];
#pragma unroll
        for (int li = NE - PB; li < NE; ++li) a[li] = 0.0;
    }
    long long t2 = clock64();
    if (threadIdx.x == 0 && clk) { clk[0] = t1 - t0; clk[1] = t2 - t1; }
}

// v2: stores and the zero-pivot flag off the dependent chain, coefficient loads issued right after the barrier,
// the next column's entry updated and published first
template <int NS>
__global__ void __launch_bounds__(64 * NS) k_diag_v2(DevCtx c, const int32_t* __restrict__ pslist, long long* clk) {
    constexpr int WP = 64, NE = WP / NS, PB = 8 / NS;
    __shared__ __align__(16) double col[2][2 * WP];
    const PStep ps = c.psteps[pslist[blockIdx.x]];
    double* G = c.F + ps.fofs + (int64_t)ps.o + (int64_t)ps.o * ps.ld;
    const int w = ps.w, ld = ps.ld;
    const int r = threadIdx.x / NS, h = threadIdx.x % NS;
    long long t0 = clock64();
    double a[NE];
#pragma unroll
    for (int li = 0; li < NE; ++li) {
        const int j = ((li >> 1) * NS + h) * 2 + (li & 1);
        a[li] = (j <= r && r < w) ? __ldcg(G + r + (size_t)j * ld) : 0.0;
    }
    for (int i = threadIdx.x; i < 2 * WP; i += 64 * NS) col[i >> 6][WP + (i & 63)] = 0.0;
    if (h == 0) col[0][r] = a[0];                       // column 0
    bool bad = false;
    double* gk = G + r;                                  // &G(r, k)
    long long t1 = clock64();
#pragma unroll 1
    for (int kb = 0; kb < w; kb += 8) {
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const int k = kb + cc;
            if (k < w) {
                const double* cb = col[k & 1];
                double* cn = col[(k + 1) & 1];
                long long ta = 0, tb = 0, tc = 0, td = 0, te = 0, tf = 0;
                if (k == 16) ta = clock64();
                __syncthreads();
                if (k == 16) tb = clock64();
                const double d = cb[k];
                const double ar = cb[r];
                const double* cj = cb + kb + h * 2;
                double u[NE];
#pragma unroll
                for (int li = 0; li < NE; ++li) u[li] = cj[(li >> 1) * 2 * NS + (li & 1)];
                const double l = ar * fast_rcp(d);
                // next column first: update, publish
                constexpr int dummy = 0; (void)dummy;
                const int cn1 = cc + 1;                                   // column k+1 inside this block (or first of the next)
                const int ho = ((cn1 >> 1) % NS), lcn = ((cn1 >> 1) / NS) * 2 + (cn1 & 1);
                if (k == 16) tc = clock64();
                a[lcn] -= l * u[lcn];
                if (k == 16) td = clock64();
                if (h == ho && r > k) cn[r] = a[lcn];
                if (k == 16) te = clock64();
#pragma unroll
                for (int li = 0; li < NE; ++li) if (li != lcn) a[li] -= l * u[li];
                bad |= d == 0.0;
                if (h == ((cc >> 1) % NS) && r >= k && r < w) __stcg(gk, r == k ? d : l);
                gk += ld;
                if (k == 16) { tf = clock64(); if (threadIdx.x == 0 && clk) { clk[2] = tb - ta; clk[3] = tc - tb; clk[4] = td - tc; clk[5] = te - td; clk[6] = tf - te; } }
            }
        }
#pragma unroll
        for (int li = 0; li < NE - PB; ++li) a[li] = a[li + PB];
#pragma unroll
        for (int li = NE - PB; li < NE; ++li) a[li] = 0.0;
    }
    if (bad && threadIdx.x == 0) atomicExch(c.iflag, -1);
    long long t2 = clock64();
    if (threadIdx.x == 0 && clk) { clk[0] = t1 - t0; clk[1] = t2 - t1; }
}

// PANELCLK-BEGIN
template <bool LU>
__global__ void __launch_bounds__(PANEL_REG_THREADS) k_panel_clk(DevCtx c, long long* clk, const int32_t* __restrict__ pslist,
                                                                 const int32_t* __restrict__ pfx, int count) {
    constexpr int WP = 64;
    __shared__ __align__(16) double Ts[WP * WP + WP];   // Ts[j + k*WP] = coefficient of x_k in unknown j (j > k)
    __shared__ double rd[WP];                           // 1 / diagonal (0 where the diagonal is 0, as the LU reference does)
    long long t0 = clock64();
    const int bx = blockIdx.x / PANEL_REG_SPLIT, part = blockIdx.x % PANEL_REG_SPLIT;
    int t = find_task(pfx, count, bx);
    int lb = bx - pfx[t];
    const PStep ps = c.psteps[pslist[t]];
    const int w = ps.w, ld = ps.ld, e0 = ps.o + ps.w;
    const int below = ps.R - e0;
    const int nb = (below + PANEL_ROWS - 1) / PANEL_ROWS;
    if (lb >= nb) return;                               // U-side blocks (LU) belong to k_panel
    const int tid = threadIdx.x;
    long long t1 = clock64();
    double* Fm = c.F + ps.fofs;
    const double* __restrict__ T = Fm + (int64_t)ps.o + (int64_t)ps.o * ld;
    const int i = lb * PANEL_ROWS + part * PANEL_REG_THREADS + tid;
    if (i - tid >= below) return;                       // whole block beyond the panel
    const bool active = i < below;
    double* xp = Fm + (int64_t)(e0 + (active ? i : 0)) + (int64_t)ps.o * ld;          // &X(i, k)
    double* yp = Fm + (int64_t)ps.o + (int64_t)(e0 + (active ? i : 0)) * ld;          // LDLT: &U12(k, i)
    double x[WP];
#pragma unroll
    for (int k = 0; k < WP; ++k) x[k] = (active && k < w) ? __ldcs(xp + (size_t)k * ld) : 0.0;
    constexpr int TB = 16;                              // loads in flight per thread while staging T
    for (int e0i = tid; e0i < WP * WP; e0i += TB * PANEL_REG_THREADS) {
        double v[TB];
#pragma unroll
        for (int u = 0; u < TB; ++u) {
            const int e = e0i + u * PANEL_REG_THREADS; const int k = e / WP, j = e - k * WP;       // smem slot (j,k)
            const bool in = j < w && k < w && j > k;
            v[u] = in ? __ldcg(LU ? T + k + (size_t)j * ld : T + j + (size_t)k * ld) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < TB; ++u) Ts[e0i + u * PANEL_REG_THREADS] = v[u];
    }
    static_assert((WP * WP) % (TB * PANEL_REG_THREADS) == 0, "whole batches");
    if (tid < WP) {
        Ts[WP * WP + tid] = 0.0;
        const double dg = tid < w ? __ldcg(T + tid + (size_t)tid * ld) : 0.0; rd[tid] = dg != 0.0 ? 1.0 / dg : 0.0;
    }
    long long t2 = clock64();
    __syncthreads();
    long long t3 = clock64();
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
                if (LU) *xp = xk;
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
    long long t4 = clock64();
    if (blockIdx.x == 0 && tid == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; }
}
// PANELCLK-END
// LUCLK-BEGIN
__global__ void __launch_bounds__(64 * LU_NS) k_diag_lu_clk(DevCtx c, const int32_t* __restrict__ pslist, long long* clk) {
    constexpr int WP = 64, NS = LU_NS, NE = WP / NS, PB = 8 / NS;
    __shared__ __align__(16) double S[WP * LU_SLD + WP];
    __shared__ double keys[2][WP];
    const PStep ps = c.psteps[pslist[blockIdx.x]];
    double* G = c.F + ps.fofs + (int64_t)ps.o + (int64_t)ps.o * ps.ld;
    const int w = ps.w, ld = ps.ld;
    int32_t* ipiv = c.ipiv + ps.col0;
    const int32_t* subw = c.subw + ps.sub0;
    const int r = threadIdx.x & (WP - 1), h = threadIdx.x / WP, lane = threadIdx.x & 31;
    double a[NE];
#pragma unroll
    for (int li = 0; li < NE; ++li) {
        const int j = ((li >> 1) * NS + h) * 2 + (li & 1);
        a[li] = (j < w && r < w) ? __ldcg(G + r + (size_t)j * ld) : 0.0;
    }
    long long T0 = clock64(), ta = 0, tb = 0, tc = 0, td = 0, te = 0, tf = 0;
    int pos = r;                                        // logical row of this thread's physical row
    if (h == 0) keys[0][r] = a[0];
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
                if (k == 16) ta = clock64();
                __syncthreads();                        // keys of column k (by logical row) visible
                if (k == 16) tb = clock64();
                // pivot search, every warp for itself: rows [k, s1).  |x| compares like its bit pattern, so
                // three warp reductions (high word, low word, index) replace a 5-round shuffle tournament.
                unsigned hi = 0, lo = 0; int bi = 0x7fffffff;
                {
                    const int i0 = lane, i1 = lane + 32;
                    if (i0 >= k && i0 < s1) { const double v = fabs(kc[i0]); if (v == v) { hi = __double2hiint(v); lo = __double2loint(v); bi = i0; } }
                    if (i1 >= k && i1 < s1) {
                        const double v = fabs(kc[i1]);
                        if (v == v) {
                            const unsigned h1 = __double2hiint(v), l1 = __double2loint(v);
                            if (bi == 0x7fffffff || h1 > hi || (h1 == hi && l1 > lo)) { hi = h1; lo = l1; bi = i1; }
                        }
                    }
                    const unsigned cand = bi != 0x7fffffff;
                    const unsigned mh = __reduce_max_sync(0xffffffffu, cand ? hi : 0u);
                    const bool inh = cand && hi == mh;
                    const unsigned ml = __reduce_max_sync(0xffffffffu, inh ? lo : 0u);
                    const bool inl = inh && lo == ml;
                    bi = (int)__reduce_min_sync(0xffffffffu, inl ? (unsigned)bi : 0x7fffffffu);
                    if (bi == 0x7fffffff) bi = k;       // only NaNs: keep the diagonal
                }
                if (k == 16) tc = clock64();
                const int kp = bi;
                const double pv = kc[kp];
                const bool ok = pv != 0.0;
                bad |= !ok;
                const double rinv = ok ? __drcp_rn(pv) : 1.0;
                const int oldpos = pos;
                if (ok && kp != k) { if (pos == kp) pos = k; else if (pos == k) pos = kp; }
                if (pos == k) {                         // the pivot row: publish what is left of it
#pragma unroll
                    for (int li = 0; li < NE; ++li) {
                        const int j = kb + ((li >> 1) * NS + h) * 2 + (li & 1);
                        if (j >= k && j < WP) S[k * LU_SLD + j] = a[li];
                    }
                }
                if (k == 16) td = clock64();
                __syncthreads();                        // pivot row visible
                if (k == 16) te = clock64();
                if (threadIdx.x == 0) ipiv[k] = kp - s0 + 1;
                // multipliers of this chunk already in the image travel with their rows (nobody reads them here)
                if (ok && kp != k && h == NS - 1 && r >= s0 && r < k) { const double t = S[k * LU_SLD + r]; S[k * LU_SLD + r] = S[kp * LU_SLD + r]; S[kp * LU_SLD + r] = t; }
                const double l = kc[oldpos] * rinv;
                const bool below = pos > k;
                if (below && h == ((cc >> 1) % NS) && pos < w) S[pos * LU_SLD + k] = l;
                const double lz = below ? l : 0.0;      // finished rows: keep the registers finite
                const double* pj = S + k * LU_SLD + kb + h * 2;
                double u[NE];
#pragma unroll
                for (int li = 0; li < NE; ++li) u[li] = (kb + ((li >> 1) * NS + h) * 2 < w) ? pj[(li >> 1) * 2 * NS + (li & 1)] : 0.0;
                const int cn1 = cc + 1;                 // column k+1: inside this block, or the first one of the next
                const int ho = (cn1 >> 1) % NS, lcn = ((cn1 >> 1) / NS) * 2 + (cn1 & 1);
                a[lcn] -= lz * u[lcn];
                if (h == ho) kn[pos] = a[lcn];
#pragma unroll
                for (int li = 0; li < NE; ++li) if (li != lcn) a[li] -= lz * u[li];
                if (k == 16) tf = clock64();
            }
        }
#pragma unroll
        for (int li = 0; li < NE - PB; ++li) a[li] = a[li + PB];
#pragma unroll
        for (int li = NE - PB; li < NE; ++li) a[li] = 0.0;
    }
    long long T1 = clock64();
    if (bad && threadIdx.x == 0) atomicExch(c.iflag, -1);
    __syncthreads();
    for (int e = threadIdx.x; e < w * w; e += 64 * NS) { const int j = e / w, i = e - j * w; __stcg(G + i + (size_t)j * ld, S[i * LU_SLD + j]); }
    if (threadIdx.x == 0 && clk) { clk[0] = 0; clk[1] = T1 - T0; clk[2] = tb - ta; clk[3] = tc - tb; clk[4] = td - tc; clk[5] = te - td; clk[6] = tf - te; }
}
// LUCLK-END
// instruction latencies seen by ONE warp (dependent chains), in clocks per operation
__global__ void k_lat(double* out, long long* clk, double seed) {
    __shared__ double sh[64];
    double x = seed + threadIdx.x, y = seed * 0.5;
    long long t[8];
    t[0] = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = fma(x, y, 1.0);
    t[1] = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = fast_rcp(x);
    t[2] = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    t[3] = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) { sh[threadIdx.x & 63] = x; __syncwarp(); x = sh[(threadIdx.x + 1) & 63]; __syncwarp(); }
    t[4] = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) { sh[threadIdx.x & 63] = x; __syncthreads(); x = sh[(threadIdx.x + 1) & 63]; }
    t[5] = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = 1.0 / x;
    t[6] = clock64();
    float f = (float)x;
#pragma unroll
    for (int i = 0; i < 256; ++i) f = fmaf(f, 1.0001f, 1.0f);
    t[7] = clock64();
    out[threadIdx.x] = x + f;
    if (threadIdx.x == 0) for (int i = 0; i < 7; ++i) clk[i] = t[i + 1] - t[i];
}

// throughput seen by one SM: independent DFMA chains, and the LDS.128 + 2 DFMA pattern of the panel kernel
__global__ void k_thr(double* out, long long* clk, double seed) {
    __shared__ __align__(16) double sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = seed * i;
    __syncthreads();
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = seed + i + threadIdx.x;
    const double y = seed * 0.5;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], y, 1.0);
    }
    long long t1 = clock64();
    const int h = threadIdx.x & 3;
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        const double2* tj = reinterpret_cast<const double2*>(sh + (it & 7) * 64 + h * 2);
#pragma unroll
        for (int p = 0; p < 8; ++p) { const double2 t = tj[p * 4]; x[2 * p] -= t.x * y; x[2 * p + 1] -= t.y * y; }
    }
    long long t2 = clock64();
    const int hu = (threadIdx.x >> 5) & 3;               // warp-uniform slice
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        const double2* tj = reinterpret_cast<const double2*>(sh + (it & 7) * 64 + hu * 2);
#pragma unroll
        for (int p = 0; p < 8; ++p) { const double2 t = tj[p * 4]; x[2 * p] -= t.x * y; x[2 * p + 1] -= t.y * y; }
    }
    long long t3 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        const double* tj = sh + (it & 7) * 64 + hu * 2;
#pragma unroll
        for (int p = 0; p < 16; ++p) { x[p] -= tj[p * 4] * y; }
    }
    long long t4 = clock64();
    double sum = 0; for (int i = 0; i < 16; ++i) sum += x[i];
    out[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; }
}

int main(int argc, char** argv) {
    const int R = argc > 1 ? atoi(argv[1]) : 9000, w = argc > 2 ? atoi(argv[2]) : 57, reps = 200;
    const int ld = R + (R & 1);
    std::vector<double> hF((size_t)ld * 64 + (size_t)ld * 0, 0.0);
    double* F; CK(cudaMalloc(&F, (size_t)ld * R * sizeof(double)));
    CK(cudaMemset(F, 0, (size_t)ld * R * sizeof(double)));
    // first 64 columns: diagonally dominant block + small panel entries
    srand(1);
    for (int j = 0; j < 64; ++j) for (int i = 0; i < R; ++i) hF[i + (size_t)j * ld] = (i == j ? 100.0 : 0.01 * ((rand() % 200) - 100) / 100.0);
    PStep ps{}; ps.fofs = 0; ps.col0 = 0; ps.ld = ld; ps.R = R; ps.o = 0; ps.w = w; ps.ob_end = w; ps.sub0 = 0; ps.nsub = 2; ps.front = 0;
    PStep* dps; CK(cudaMalloc(&dps, sizeof(PStep))); CK(cudaMemcpy(dps, &ps, sizeof(PStep), cudaMemcpyHostToDevice));
    int32_t zero = 0, *dlist, *dpfx, *dflag, *dsubw; int32_t pf[2] = {0, (R - w + PANEL_ROWS - 1) / PANEL_ROWS};
    CK(cudaMalloc(&dlist, 4)); CK(cudaMemcpy(dlist, &zero, 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dpfx, 8)); CK(cudaMemcpy(dpfx, pf, 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dflag, 4)); CK(cudaMemset(dflag, 0, 4));
    int32_t hsub[2] = {30, w - 30}; CK(cudaMalloc(&dsubw, 8)); CK(cudaMemcpy(dsubw, hsub, 8, cudaMemcpyHostToDevice));
    int32_t* dipiv; CK(cudaMalloc(&dipiv, 4 * 64));
    long long* dclk; CK(cudaMalloc(&dclk, 64)); CK(cudaMemset(dclk, 0, 64));
    DevCtx c{}; c.F = F; c.psteps = dps; c.iflag = dflag; c.subw = dsubw; c.ipiv = dipiv;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto reset = [&] { CK(cudaMemcpy(F, hF.data(), (size_t)ld * 64 * sizeof(double), cudaMemcpyHostToDevice)); };
    auto timeit = [&](const char* name, auto launch) {
        reset(); launch(); CK(cudaDeviceSynchronize());
        float best = 1e9f, tot = 0;
        for (int it = 0; it < 20; ++it) {
            reset(); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; tot += ms;
        }
        CK(cudaGetLastError());
        long long hc[8] = {0}; CK(cudaMemcpy(hc, dclk, 64, cudaMemcpyDeviceToHost));
        printf("%-34s best %7.2f us  avg %7.2f us   clk load %lld loop %lld | bar %lld lds+rcp %lld fma %lld sts %lld rest %lld\n", name, best * 1e3, tot / 20 * 1e3, hc[0], hc[1], hc[2], hc[3], hc[4], hc[5], hc[6]);
        CK(cudaMemset(dclk, 0, 64));
    };
    (void)reps;
    const int nbp = pf[1];
    {
        long long* dl; double* dout; CK(cudaMalloc(&dl, 64)); CK(cudaMalloc(&dout, 8 * 1024));
        for (int nt : {32, 64, 128, 256}) {
            k_lat<<<1, nt>>>(dout, dl, 1.25); CK(cudaDeviceSynchronize());
            long long hl[7]; CK(cudaMemcpy(hl, dl, 56, cudaMemcpyDeviceToHost));
            printf("latency (%3d threads): dfma %.1f  fast_rcp %.1f  shfl %.1f  sts+syncwarp+lds %.1f  sts+bar+lds %.1f  ieee div %.1f  ffma %.1f clk\n", nt,
                   hl[0] / 256.0, hl[1] / 64.0, hl[2] / 64.0, hl[3] / 64.0, hl[4] / 64.0, hl[5] / 64.0, hl[6] / 256.0);
        }
    }
    {
        long long* dl; double* dout; CK(cudaMalloc(&dl, 64)); CK(cudaMalloc(&dout, 8 * 1024));
        for (int nt : {32, 128, 256, 512, 1024}) {
            k_thr<<<1, nt>>>(dout, dl, 1.25); CK(cudaDeviceSynchronize());
            long long hl[4]; CK(cudaMemcpy(hl, dl, 32, cudaMemcpyDeviceToHost));
            printf("throughput (%4d threads): dfma %.1f lanes/clk/SM   lds128(4 addr)+2dfma: %.1f   lds128(uniform)+2dfma: %.1f   lds64(uniform)+dfma: %.1f\n", nt,
                   64.0 * 16 * nt / hl[0], 64.0 * 16 * nt / hl[1], 64.0 * 16 * nt / hl[2], 64.0 * 16 * nt / hl[3]);
        }
    }
    printf("R=%d w=%d panel blocks=%d\n", R, w, nbp);
    timeit("empty-ish (k_perm_gather n=0)", [&] { k_ipiv_widen<<<1, 32>>>(0, dipiv, (int64_t*)dclk); });
    timeit("diag_ldlt_row (product)", [&] { k_diag_ldlt_row<<<1, 64 * DIAG_NS>>>(c, dlist); });
    timeit("diag_ldlt_reg<4,16>", [&] { k_diag_ldlt_reg<4, 16><<<1, 256>>>(c, dlist); });
    timeit("diag LU row (bar1 | argmax | publish | bar2 | update)", [&] { k_diag_lu_clk<<<1, 64 * LU_NS>>>(c, dlist, dclk); });
    timeit("diag LU row product", [&] { k_diag_lu_row<<<1, 64 * LU_NS>>>(c, dlist); });
    timeit("diag v2 NS=4", [&] { k_diag_v2<4><<<1, 256>>>(c, dlist, dclk); });
    timeit("diag v2 NS=2", [&] { k_diag_v2<2><<<1, 128>>>(c, dlist, dclk); });
    timeit("diag v2 NS=1", [&] { k_diag_v2<1><<<1, 64>>>(c, dlist, dclk); });
    timeit("diag NS=4 var0", [&] { k_diag_var<0, 4><<<1, 256>>>(c, dlist, dclk); });
    timeit("diag NS=2 var0", [&] { k_diag_var<0, 2><<<1, 128>>>(c, dlist, dclk); });
    timeit("diag NS=2 var2 (no barrier)", [&] { k_diag_var<2, 2><<<1, 128>>>(c, dlist, dclk); });
    timeit("diag NS=2 var3 (no update)", [&] { k_diag_var<3, 2><<<1, 128>>>(c, dlist, dclk); });
    timeit("diag NS=1 var0", [&] { k_diag_var<0, 1><<<1, 64>>>(c, dlist, dclk); });
    timeit("diag NS=1 var2 (no barrier)", [&] { k_diag_var<2, 1><<<1, 64>>>(c, dlist, dclk); });
    timeit("diag NS=1 var3 (no update)", [&] { k_diag_var<3, 1><<<1, 64>>>(c, dlist, dclk); });
    timeit("panel clk (ps | issue loads | wait | compute)", [&] { k_panel_clk<false><<<nbp * PANEL_REG_SPLIT, PANEL_REG_THREADS>>>(c, dclk, dlist, dpfx, 1); });
    timeit("panel_reg<ldlt>", [&] { k_panel_reg<false><<<nbp * PANEL_REG_SPLIT, PANEL_REG_THREADS>>>(c, dlist, dpfx, 1); });
    {
        size_t sm = 0; for (int ww = 1; ww <= w; ++ww) sm = std::max(sm, panel_smem_bytes(ww));
        CK(cudaFuncSetAttribute(k_panel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        timeit("panel smem<ldlt>", [&] { k_panel<false><<<nbp, PANEL_ROWS, sm>>>(c, dlist, dpfx, 1, 0); });
    }
    return 0;
}
