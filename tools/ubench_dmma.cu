// Micro-benchmark of the DMMA trailing-update kernel (gemm_dmma.cuh) on the shapes of the delayed updates:
// one task  C[m x n] -= A[m x k] * B[k x n]  inside a synthetic front (full, or lower-triangular like the LDL^T
// updates), tile list built by the plan's own dmma_tiles().  Prints TFLOP/s per kernel variant and checks every
// variant against a plain FP64 kernel on a sample of entries.  Development tool only; never a bench number.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr \
//        -I sparspak.jl_b200/csrc -I include tools/ubench_dmma.cu -o tools/ubench_dmma
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "plan.hpp"
#include "kernels.cuh"
#include "gemm_dmma.cuh"
using namespace spk;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void k_fill(double* p, size_t n, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned x = (unsigned)(i * 2654435761u) ^ seed; x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15;
        p[i] = ((double)(x & 0xFFFFF) / 1048576.0 - 0.5);
    }
}
// reference entries: out[s] = sum_k A(i,k) B(k,j) for sampled (i,j)
__global__ void k_ref(const double* F, GemmTask g, const int* si, const int* sj, int ns, double* out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    const double* A = F + g.a0 + si[s];
    const double* B = F + g.b0 + (size_t)sj[s] * g.ld;
    double acc = 0.0;
    for (int k = 0; k < g.k; ++k) acc += A[(size_t)k * g.ld] * B[k];
    out[s] = acc;
}

struct Variant { const char* name; GemmKernel fn; int threads; size_t smem; int tm, tn; };

int main(int argc, char** argv) {
    int m = argc > 1 ? atoi(argv[1]) : 9000, k = argc > 2 ? atoi(argv[2]) : 456, lower = argc > 3 ? atoi(argv[3]) : 1;
    int odd = argc > 4 ? atoi(argv[4]) : 0; int flags = argc > 5 ? atoi(argv[5]) : 1;
    int n = argc > 6 ? atoi(argv[6]) : m;                       // columns of C (default square; a narrow n = the in-block left-looking updates)                     // 1: odd panel offsets (exercises the shifted-origin path)
    int reps = 5;
    CK(cudaSetDevice(0));
    CK(gemm_dmma_init());
    // synthetic front: R = k + m (+ odd), panel columns [o, o+k), trailing block [o+k, R)
    const int o = odd ? 1 : 0;
    const int R = o + k + m;
    const int ld = (R + 1) & ~1;
    const size_t nel = (size_t)ld * R;
    double *F = nullptr, *F0 = nullptr;
    CK(cudaMalloc(&F, nel * sizeof(double))); CK(cudaMalloc(&F0, nel * sizeof(double)));
    k_fill<<<1024, 256>>>(F0, nel, 12345u); CK(cudaDeviceSynchronize());
    GemmTask g{};
    g.ld = ld; g.m = m; g.n = n; g.k = k;
    const int e = o + k;
    g.a0 = (int64_t)e + (int64_t)o * ld; g.c0 = (int64_t)e + (int64_t)e * ld; g.b0 = (int64_t)o + (int64_t)e * ld;
    g.lower = (uint8_t)lower; g.roff = 0;
    double flops = 2.0 * m * (double)n * k; if (lower) flops -= (double)n * n * k;
    GemmTask* d_task; CK(cudaMalloc(&d_task, sizeof(GemmTask))); CK(cudaMemcpy(d_task, &g, sizeof(g), cudaMemcpyHostToDevice));
    int32_t* d_ctr; CK(cudaMalloc(&d_ctr, 64 * sizeof(int32_t)));
    // samples
    const int ns = 4096;
    std::vector<int> si(ns), sj(ns);
    unsigned s = 777;
    for (int i = 0; i < ns; ++i) {
        s = s * 1664525u + 1013904223u; int a = (s >> 8) % m; s = s * 1664525u + 1013904223u; int b = (s >> 8) % n;
        if (lower && a < b) std::swap(a, b);
        si[i] = a; sj[i] = b;
    }
    int *d_si, *d_sj; double* d_ref;
    CK(cudaMalloc(&d_si, ns * sizeof(int))); CK(cudaMalloc(&d_sj, ns * sizeof(int))); CK(cudaMalloc(&d_ref, ns * sizeof(double)));
    CK(cudaMemcpy(d_si, si.data(), ns * sizeof(int), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_sj, sj.data(), ns * sizeof(int), cudaMemcpyHostToDevice));
    k_ref<<<(ns + 127) / 128, 128>>>(F0, g, d_si, d_sj, ns, d_ref); CK(cudaDeviceSynchronize());
    std::vector<double> ref(ns), c0(ns), c1(ns);
    CK(cudaMemcpy(ref.data(), d_ref, ns * sizeof(double), cudaMemcpyDeviceToHost));

    std::vector<Variant> vars;
    auto addv = [&](const char* name, int kind, int variant, int tm) {
        GemmVariant v = gemm_dmma_variant(kind, variant);
        vars.push_back({name, v.fn, v.threads, v.smem, tm, 64});
    };
    addv("128x64 tk16 s3 (default)", K_GEMM_B64, 4, BIG_TM);
    addv("128x64 tk32 s2", K_GEMM_B64, 3, BIG_TM);
    addv("128x64 tk8 s6", K_GEMM_B64, 5, BIG_TM);
    addv("128x64 tk16 s4", K_GEMM_B64, 6, BIG_TM);
    addv("64x64 tk16 s3", K_GEMM_T64, 4, 64);
#define ADDK(name, TM, TN, WM, WN, MINB, TK, ST) vars.push_back({name, k_gemm_dmma<TM, TN, WM, WN, MINB, TK, ST>, WM * WN * 32, DmmaCfg<TM, TN, TK, ST>::SMEM, TM, TN})
    ADDK("64x64 tk8 s4", 64, 64, 2, 2, 4, 8, 4);
    ADDK("64x64 8wp(32x16) tk8 s4", 64, 64, 2, 4, 2, 8, 4);
    ADDK("64x64 8wp(16x32) tk8 s4", 64, 64, 4, 2, 2, 8, 4);
    ADDK("64x64 16wp(16x16) tk16 s3", 64, 64, 4, 4, 1, 16, 3);
    ADDK("32x64 2wp tk8 s4", 32, 64, 1, 2, 8, 8, 4);
    ADDK("32x64 4wp(16x32) tk8 s4", 32, 64, 2, 2, 4, 8, 4);
    ADDK("64x32 2wp tk8 s4", 64, 32, 2, 1, 8, 8, 4);
    ADDK("64x32 4wp(32x16) tk8 s4", 64, 32, 2, 2, 4, 8, 4);
    ADDK("32x32 4wp(16x16) tk8 s4", 32, 32, 2, 2, 4, 8, 4);
    ADDK("32x32 1wp tk8 s4", 32, 32, 1, 1, 8, 8, 4);
    DevCtx c{}; c.F = F;
    printf("m=%d n=%d k=%d lower=%d odd=%d ld=%d  flops=%.3e\n", m, n, k, lower, odd, ld, flops);
    const char* only = getenv("UB_ONLY");                     // run one variant (by substring of its name): ncu captures
    if (const char* r = getenv("UB_REPS")) reps = atoi(r);
    for (const Variant& v : vars) {
        if (only && !strstr(v.name, only)) continue;
        std::vector<GemmTile> tiles;
        dmma_tiles(g, v.tm, v.tn, 0, &tiles, nullptr);
        GemmTile* d_tiles; CK(cudaMalloc(&d_tiles, tiles.size() * sizeof(GemmTile)));
        CK(cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(GemmTile), cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int r = 0; r < reps + 1; ++r) {
            CK(cudaMemcpy(F, F0, nel * sizeof(double), cudaMemcpyDeviceToDevice));
            CK(cudaMemset(d_ctr, 0, 64 * sizeof(int32_t)));
            CK(cudaEventRecord(e0));
            v.fn<<<(int)tiles.size(), v.threads, v.smem>>>(c, d_task, d_tiles, (int)tiles.size(), d_ctr, flags);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0 && ms < best) best = ms;
        }
        // check: C_new = C_old - ref on the samples
        double maxerr = 0.0;
        for (int i = 0; i < ns; ++i) {
            double cn, co;
            const size_t off = (size_t)g.c0 + si[i] + (size_t)sj[i] * ld;
            CK(cudaMemcpy(&cn, F + off, sizeof(double), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&co, F0 + off, sizeof(double), cudaMemcpyDeviceToHost));
            maxerr = std::max(maxerr, std::fabs((co - cn) - ref[i]) / (1.0 + std::fabs(ref[i])));
            if (i >= 255) break;                               // 256 samples per variant keep the tool quick
        }
        printf("  %-28s tiles %6zu  %8.3f ms  %6.2f TFLOP/s   max rel err %.2e %s\n", v.name, tiles.size(), best, flops / best / 1e9, maxerr, maxerr < 1e-11 ? "" : "  <-- MISMATCH");
        CK(cudaFree(d_tiles));
    }
    return 0;
}
