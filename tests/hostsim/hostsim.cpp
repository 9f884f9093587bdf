// TEST INFRASTRUCTURE — sequential host executor of the device schedule built by
// sparspak.jl_b200/csrc/plan.hpp.  It runs the SAME task lists the CUDA kernels run
// (load / assemble / diag / panel / GEMM / store / solve steps) with plain loops, so the
// schedule (fronts, relative indices, position maps, blocking) can be validated against the
// oracle on a machine without a GPU.  Never linked into the product library.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "plan.hpp"
using namespace spk;
#define API extern "C" __attribute__((visibility("default")))

struct Sim {
    Plan P;
    std::vector<double> F, lnz, unz, w;
    std::vector<int32_t> ipiv;
    int iflag = 0;
};

static void chunks_io(Sim& s, bool store) {
    Plan& P = s.P;
    for (const Chunk& c : P.chunks) {
        if (c.fofs < 0) continue;                           // a front this part holds no storage for
        const int32_t* pos = P.pos.data() + c.posofs;
        double* F = s.F.data() + c.fofs;
        for (int j = 0; j < c.nj; ++j)
            for (int i = 0; i < c.jlen; ++i) {
                double& f = F[(int64_t)pos[i] + (int64_t)(c.o + j) * c.ld];
                double& l = s.lnz[c.lofs + i + (int64_t)j * c.jlen];
                if (store) l = f; else f = l;
            }
        if (P.lu) {
            int ldu = c.jlen - c.nj;
            for (int j = 0; j < c.nj; ++j)
                for (int i = 0; i < ldu; ++i) {
                    double& f = F[(int64_t)(c.o + j) + (int64_t)pos[c.nj + i] * c.ld];
                    double& u = s.unz[c.uofs + i + (int64_t)j * ldu];
                    if (store) u = f; else f = u;
                }
        }
    }
}

static void assemble(Sim& s, const Front& C, const Front& Pa) {
    Plan& P = s.P;
    const int32_t* rel = P.rel.data() + C.relofs;
    const int32_t pf = (int32_t)(&Pa - P.fronts.data());
    const int8_t* own = (!P.ownofs.empty() && P.ownofs[pf] >= 0) ? P.fown.data() + P.ownofs[pf] : nullptr;
    for (int j = 0; j < C.m; ++j)
        for (int i = 0; i < C.m; ++i) {
            if (!P.lu && i < j) continue;
            if (own && own[rel[j]] != P.part) continue;     // distributed parent: only the columns this part owns
            s.F[Pa.fofs + (int64_t)rel[i] + (int64_t)rel[j] * Pa.ld] += s.F[C.fofs + (int64_t)(C.W + i) + (int64_t)(C.W + j) * C.ld];
        }
}

static void diag(Sim& s, const PStep& ps) {
    Plan& P = s.P;
    double* A = s.F.data() + ps.fofs + (int64_t)ps.o + (int64_t)ps.o * ps.ld;
    const int ld = ps.ld, w = ps.w;
    if (P.lu) {
        int32_t* ip = s.ipiv.data() + ps.col0;
        int s0 = 0;
        for (int b = 0; b < ps.nsub; ++b) {
            int s1 = s0 + P.subw[ps.sub0 + b];
            for (int k = s0; k < s1; ++k) {
                int kp = k; double best = std::fabs(A[k + (size_t)k * ld]);
                for (int i = k + 1; i < s1; ++i) { double v = std::fabs(A[i + (size_t)k * ld]); if (v > best) { best = v; kp = i; } }
                ip[k] = kp - s0 + 1;
                double pv = A[kp + (size_t)k * ld];
                if (pv == 0.0) s.iflag = -1;
                if (pv != 0.0) {
                    if (kp != k) for (int j = s0; j < w; ++j) std::swap(A[k + (size_t)j * ld], A[kp + (size_t)j * ld]);
                    double inv = 1.0 / A[k + (size_t)k * ld];
                    for (int i = k + 1; i < w; ++i) A[i + (size_t)k * ld] *= inv;
                }
                for (int j = k + 1; j < w; ++j) for (int i = k + 1; i < w; ++i) A[i + (size_t)j * ld] -= A[i + (size_t)k * ld] * A[k + (size_t)j * ld];
            }
            s0 = s1;
        }
    } else {
        for (int k = 0; k < w; ++k) {
            double d = A[k + (size_t)k * ld];
            if (d == 0.0) s.iflag = -1;
            for (int i = k + 1; i < w; ++i) A[i + (size_t)k * ld] /= d;
            for (int c = k + 1; c < w; ++c) for (int r = c; r < w; ++r) A[r + (size_t)c * ld] -= (A[c + (size_t)k * ld] * d) * A[r + (size_t)k * ld];
        }
    }
}

static void panel(Sim& s, const PStep& ps) {
    Plan& P = s.P;
    const int w = ps.w, ld = ps.ld, e0 = ps.o + ps.w, below = ps.R - e0;
    double* Fm = s.F.data() + ps.fofs;
    const double* T = Fm + (int64_t)ps.o + (int64_t)ps.o * ld;
    for (int i = 0; i < below; ++i) {
        double* X = Fm + (int64_t)(e0 + i) + (int64_t)ps.o * ld;
        if (P.lu) {
            for (int j = 0; j < w; ++j) {
                double acc = X[(size_t)j * ld];
                for (int k = 0; k < j; ++k) acc -= T[k + (size_t)j * ld] * X[(size_t)k * ld];
                X[(size_t)j * ld] = (1.0 / T[j + (size_t)j * ld]) * acc;
            }
            double* Y = Fm + (int64_t)ps.o + (int64_t)(e0 + i) * ld;
            const int32_t* ip = s.ipiv.data() + ps.col0;
            int s0 = 0;
            for (int b = 0; b < ps.nsub; ++b) {
                int s1 = s0 + P.subw[ps.sub0 + b];
                for (int j = s0; j < s1; ++j) { double acc = Y[j]; for (int k = 0; k < s0; ++k) acc -= T[j + (size_t)k * ld] * Y[k]; Y[j] = acc; }
                for (int k = s0; k < s1; ++k) { int q = s0 + ip[k] - 1; if (q != k) std::swap(Y[k], Y[q]); }
                for (int j = s0; j < s1; ++j) { double acc = Y[j]; for (int k = s0; k < j; ++k) acc -= T[j + (size_t)k * ld] * Y[k]; Y[j] = acc; }
                s0 = s1;
            }
        } else {
            for (int j = 0; j < w; ++j) {
                double acc = X[(size_t)j * ld];
                for (int k = 0; k < j; ++k) acc -= T[j + (size_t)k * ld] * X[(size_t)k * ld];
                X[(size_t)j * ld] = acc;
            }
            double* Y = Fm + (int64_t)ps.o + (int64_t)(e0 + i) * ld;       // U12 = X^T = D * L21^T in the upper triangle
            for (int j = 0; j < w; ++j) { Y[j] = X[(size_t)j * ld]; X[(size_t)j * ld] /= T[j + (size_t)j * ld]; }
        }
    }
}

static void gemm(Sim& s, const GemmTask& g) {
    std::vector<double> acc((size_t)g.m * g.n, 0.0);
    const double* A = s.F.data() + g.a0; const double* B = s.F.data() + g.b0;
    for (int k = 0; k < g.k; ++k)
        for (int j = 0; j < g.n; ++j) {
            double b = B[(size_t)k + (size_t)j * g.ld];
            for (int i = 0; i < g.m; ++i) acc[i + (size_t)j * g.m] += A[i + (size_t)k * g.ld] * b;
        }
    double* C = s.F.data() + g.c0;
    for (int j = 0; j < g.n; ++j) for (int i = 0; i < g.m; ++i) if (!(g.lower && i + g.roff < j)) C[i + (size_t)j * g.ld] -= acc[i + (size_t)j * g.m];
}

// one C tile of a DMMA launch, with the kernel's conventions: origin moved to the previous even row (sa),
// rows in front of the origin and entries above the diagonal (lower) masked
static void gemm_tile(Sim& s, const GemmTask& g, const GemmTile& tl, int tm, int tn) {
    const int sa = (int)(g.a0 & 1);
    const int mp = g.m + sa, roffp = g.roff - sa;
    const double* A = s.F.data() + (g.a0 - sa); const double* B = s.F.data() + g.b0;
    double* C = s.F.data() + (g.c0 - sa);
    for (int j = tl.tj * tn; j < std::min<int>(g.n, (tl.tj + 1) * tn); ++j)
        for (int i = std::max<int>(tl.ti * tm, sa); i < std::min<int>(mp, (tl.ti + 1) * tm); ++i) {
            if (g.lower && i + roffp < j) continue;
            double acc = 0.0;
            for (int k = 0; k < g.k; ++k) acc += A[i + (size_t)k * g.ld] * B[(size_t)k + (size_t)j * g.ld];
            C[i + (size_t)j * g.ld] -= acc;
        }
}

API void* sim_create2(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode, const int64_t* xlindx,
                      const int64_t* lindx, const int64_t* xlnz, const int64_t* xunz, int use_dmma_buckets,
                      int relax_abs, double relax_frac, int alloc, int part, int nparts);
API void* sim_create(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode, const int64_t* xlindx,
                     const int64_t* lindx, const int64_t* xlnz, const int64_t* xunz, int use_dmma_buckets,
                     int relax_abs, double relax_frac, int alloc) {
    return sim_create2(n, nsuper, xsuper, snode, xlindx, lindx, xlnz, xunz, use_dmma_buckets, relax_abs, relax_frac, alloc, 0, 1);
}
API void* sim_create2(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode, const int64_t* xlindx,
                      const int64_t* lindx, const int64_t* xlnz, const int64_t* xunz, int use_dmma_buckets,
                      int relax_abs, double relax_frac, int alloc, int part, int nparts) {
    Sim* s = new Sim();
    s->P.part = part; s->P.nparts = nparts;
    plan_env_overrides(s->P);
    s->P.relax_abs = relax_abs; s->P.relax_frac = relax_frac;
    if (!analyze(s->P, n, nsuper, xsuper, snode, xlindx, lindx, xlnz, xunz)) { fprintf(stderr, "analyze: %s\n", s->P.error.c_str()); delete s; return nullptr; }
    s->P.use_dmma = use_dmma_buckets != 0;
    if (const char* e = getenv("SPK_SOLVE_SMALL")) s->P.solve_small = atoll(e);
    build_schedule(s->P);
    if (alloc) {
        s->F.assign(std::max<int64_t>(s->P.arena, 1), 0.0);
        s->w.assign(std::max<int64_t>(s->P.wlen, 1), 0.0);
        s->ipiv.assign(n, 0);
    }
    return s;
}
API void sim_destroy(void* h) { delete (Sim*)h; }
API int64_t sim_stat(void* h, int what) {
    Sim* s = (Sim*)h; Plan& P = s->P;
    switch (what) {
    case 0: return (int64_t)P.fronts.size(); case 1: return P.nlevels; case 2: return P.arena; case 3: return P.wlen;
    case 4: return (int64_t)P.factor_launches.size(); case 5: return (int64_t)(P.fwd_launches.size() + P.bwd_launches.size());
    case 6: return (int64_t)P.gemmt.size(); case 7: return P.maxpw; case 8: return P.maxR;
    case 9: { int64_t k = 0; for (auto& f : P.fronts) if (f.nch > 1) ++k; return k; }
    case 10: { int64_t k = 0; for (auto& L : P.factor_launches) k += L.nblocks; return k; }
    case 11: return (int64_t)P.psteps.size();
    case 12: { int64_t k = 0; for (auto& f : P.fronts) k += (int64_t)f.m * f.m; return k; }
    case 13: return P.dist_top ? 1 : 0;
    case 20: { uint64_t h = 1469598103934665603ull; for (int32_t v : P.pos) { h ^= (uint32_t)v; h *= 1099511628211ull; } return (int64_t)(h >> 1); }   // position-map fingerprint
    case 21: return (int64_t)P.pos.size();
    case 14: { int64_t k = 0; for (int32_t o : P.owner) if (o == -1) ++k; return k; }
    default: return 0; }
}
API double sim_statf(void* h, int what) {
    Sim* s = (Sim*)h; Plan& P = s->P;
    if (what == 0) return P.flops_struct; if (what == 1) return P.nnzL;
    if (what == 2) { double f = 0; for (auto& L : P.factor_launches) f += L.flops; return f; }
    if (what == 3) { double f = 0; for (auto& L : P.factor_launches) if (L.kind == K_GEMM_B64 || L.kind == K_GEMM_T64) f += L.flops; return f; }
    return 0;
}

static int64_t run_factor_list(Sim* s, const std::vector<Launch>& Ls) {
    Plan& P = s->P;
    for (const Launch& L : Ls) {
        switch (L.kind) {
        case K_ASM:
            for (int t = 0; t < L.count; ++t) { AsmTask a = P.asmt[L.first + t]; assemble(*s, P.fronts[a.child], P.fronts[a.parent]); } break;
        case K_ASM_TAIL:
            for (int t = 0; t < L.count; ++t) {
                const Front& Pa = P.fronts[P.asmt[L.first + t].parent];
                for (int r = ASM_ROUNDS; r < Pa.nchild; ++r) assemble(*s, P.fronts[P.childlist[Pa.child0 + r]], Pa);
            } break;
        case K_DIAG: for (int t = 0; t < L.count; ++t) diag(*s, P.psteps[P.pslist[L.first + t]]); break;
        case K_PANEL: for (int t = 0; t < L.count; ++t) panel(*s, P.psteps[P.pslist[L.first + t]]); break;
        case K_GEMM: for (int t = 0; t < L.count; ++t) gemm(*s, P.gemmt[L.first + t]); break;
        case K_GEMM_B64: case K_GEMM_T64:                       // tile by tile, as the persistent kernel walks its list
            for (int64_t q = L.tile0; q < L.tile0 + L.ntiles; ++q) gemm_tile(*s, P.gemmt[L.first + P.tiles[q].task], P.tiles[q], L.tile_m, L.tile_n);
            break;
        case K_FRONT_SMALL:                                      // fused small fronts: extend-add, then step by step with right-looking updates
            for (int t = 0; t < L.count; ++t) {
                const Front& F = P.fronts[P.pslist[L.first + t]];
                for (int r = 0; r < F.nchild; ++r) assemble(*s, P.fronts[P.childlist[F.child0 + r]], F);
                for (int j = 0; j < F.nps; ++j) {
                    const PStep& ps = P.psteps[F.ps0 + j];
                    diag(*s, ps); panel(*s, ps);
                    const int e = ps.o + ps.w;
                    if (e < F.R) gemm(*s, front_gemm(P, F, e, F.R - e, e, F.R - e, ps.o, ps.w));
                }
            } break;
        case K_FILLU:
            for (int t = 0; t < L.count; ++t) {
                const FillTask& ft = P.fillt[L.first + t];
                double* Fm = s->F.data() + ft.fofs;
                for (int c = ft.e; c < ft.R; ++c) for (int k = ft.ob0; k < ft.e; ++k) Fm[(int64_t)k + (int64_t)c * ft.ld] = Fm[(int64_t)k + (int64_t)k * ft.ld] * Fm[(int64_t)c + (int64_t)k * ft.ld];
            } break;
        case K_BCAST: return -101;                          // the caller drives the exchanges (sim_top_next)
        default: return -100;
        }
    }
    return 0;
}

API void sim_load(void* h, const double* lnz, const double* unz) {
    Sim* s = (Sim*)h; Plan& P = s->P;
    s->lnz.assign(lnz, lnz + P.nlnz);
    if (P.lu) s->unz.assign(unz, unz + P.nunz);
    std::fill(s->F.begin(), s->F.end(), 0.0);
    s->iflag = 0;
    chunks_io(*s, false);
}
API void sim_store(void* h) { chunks_io(*(Sim*)h, true); }
// which: 0 all, 1 this part's subtrees, 2 top set
API int64_t sim_factor_list(void* h, int which) {
    Sim* s = (Sim*)h; Plan& P = s->P;
    int64_t rc = run_factor_list(s, which == 0 ? P.factor_launches : (which == 1 ? P.factor_local : P.factor_top));
    return rc ? rc : s->iflag;
}
// Top-set list of a multi-part plan: runs the launches [from, ...) up to the next K_BCAST launch and returns its
// index (the caller performs its broadcasts, then continues at index + 1), or -1 when the list is done.
API int64_t sim_top_next(void* h, int64_t from) {
    Sim* s = (Sim*)h; Plan& P = s->P;
    for (int64_t i = from; i < (int64_t)P.factor_top.size(); ++i) {
        if (P.factor_top[i].kind == K_BCAST) return i;
        std::vector<Launch> one(1, P.factor_top[i]);
        if (run_factor_list(s, one)) return -100;
    }
    return -1;
}
// broadcasts of launch `li` of the top-set list: count when out == NULL, else out = {arena offset, length, root}
API int64_t sim_bcast_info(void* h, int64_t li, int64_t i, int64_t* out) {
    Sim* s = (Sim*)h; Plan& P = s->P;
    const Launch& L = P.factor_top[li];
    if (!out) return L.count;
    const Bcast& b = P.bcasts[L.first + i];
    out[0] = b.ofs; out[1] = b.len; out[2] = b.root;
    return 0;
}
API int64_t sim_iflag(void* h) { return ((Sim*)h)->iflag; }
API void* sim_ptr(void* h, int what) {
    Sim* s = (Sim*)h;
    switch (what) { case 0: return s->lnz.data(); case 1: return s->unz.data(); case 2: return s->ipiv.data(); case 5: return s->F.data(); case 6: return s->w.data(); default: return nullptr; }
}
API int64_t sim_len(void* h, int what) {
    Sim* s = (Sim*)h;
    switch (what) { case 0: return (int64_t)s->lnz.size(); case 1: return (int64_t)s->unz.size(); case 2: return (int64_t)s->ipiv.size(); case 5: return (int64_t)s->F.size(); case 6: return (int64_t)s->w.size(); default: return 0; }
}
API int64_t sim_xchg_info(void* h, int what, int64_t i, int64_t* out) {
    Sim* s = (Sim*)h; const Plan& P = s->P;
    if (what == 0) {
        if (!out) return (int64_t)P.xchg.size();
        const Front& F = P.fronts[P.xchg[i]];
        out[0] = P.owner[P.xchg[i]]; out[1] = F.fofs; out[2] = (int64_t)F.ld * F.R; out[3] = F.wofs; out[4] = F.R; out[5] = P.xchg[i];
        return 0;
    }
    if (!out) return (int64_t)P.ranges.size();
    const Plan::Range& g = P.ranges[i];
    out[0] = g.owner; out[1] = g.lnz0; out[2] = g.lnz1 - g.lnz0; out[3] = g.unz0; out[4] = g.unz1 - g.unz0; out[5] = g.col0; out[6] = g.col1 - g.col0;
    return 0;
}

API int64_t sim_factor(void* h, double* lnz, double* unz, int64_t* ipvt) {
    Sim* s = (Sim*)h; Plan& P = s->P;
    sim_load(h, lnz, unz);
    if (!P.factor_pipe.empty()) {                       // the product's default order: tree pipelines, then the top set
        for (auto& Lp : P.factor_pipe) if (run_factor_list(s, Lp)) return -100;
        if (run_factor_list(s, P.factor_ptop)) return -100;
    } else if (run_factor_list(s, P.factor_launches)) return -100;
    chunks_io(*s, true);
    std::copy(s->lnz.begin(), s->lnz.end(), lnz);
    if (P.lu) { std::copy(s->unz.begin(), s->unz.end(), unz); for (int64_t i = 0; i < P.n; ++i) ipvt[i] = s->ipiv[i]; }
    return s->iflag;
}

static void fwd_diag(Sim& s, const SolveTask& t) {
    double* x = s.w.data() + t.wofs + t.o; const double* T = s.lnz.data() + t.lofs;
    if (s.P.lu) for (int k = 0; k < t.nj; ++k) { int q = s.ipiv[t.col0 + k] - 1; if (q != k) std::swap(x[k], x[q]); }
    for (int k = 0; k < t.nj; ++k) for (int i = k + 1; i < t.nj; ++i) x[i] -= x[k] * T[i + (size_t)k * t.ld];
}
static void fwd_update(Sim& s, const SolveTask& t) {
    double* wf = s.w.data() + t.wofs;
    for (int i = 0; i < t.m; ++i) { double acc = 0; for (int k = 0; k < t.nj; ++k) acc += (-wf[t.o + k]) * s.lnz[t.lofs + t.nj + i + (size_t)k * t.ld]; wf[s.P.pos[t.posofs + t.nj + i]] += acc; }
}
static void bwd_update(Sim& s, const SolveTask& t) {
    double* wf = s.w.data() + t.wofs; const bool lu = s.P.lu;
    if (t.m == 0 && lu) return;
    for (int k = 0; k < t.nj; ++k) {
        double sum = 0;
        for (int i = 0; i < t.m; ++i) sum += (lu ? s.unz[t.uofs + i + (size_t)k * t.ldu] : s.lnz[t.lofs + t.nj + i + (size_t)k * t.ld]) * wf[s.P.pos[t.posofs + t.nj + i]];
        if (lu) wf[t.o + k] += -sum; else wf[t.o + k] = wf[t.o + k] / s.lnz[t.lofs + k + (size_t)k * t.ld] - sum;
    }
}
static void bwd_diag(Sim& s, const SolveTask& t, double* rhs) {
    double* x = s.w.data() + t.wofs + t.o; const double* T = s.lnz.data() + t.lofs; const bool lu = s.P.lu;
    for (int k = t.nj - 1; k >= 0; --k) {
        if (lu) { x[k] /= T[k + (size_t)k * t.ld]; for (int i = 0; i < k; ++i) x[i] -= x[k] * T[i + (size_t)k * t.ld]; }
        else for (int i = 0; i < k; ++i) x[i] -= x[k] * T[k + (size_t)i * t.ld];
    }
    for (int k = 0; k < t.nj; ++k) rhs[t.col0 + k] = x[k];
}

// ---- panel-step solves on the frontal matrices (same semantics as the k_pf_* / k_pb_* kernels)
static void pf_diag(Sim& s, const PStep& ps) {
    Plan& P = s.P; const Front& F = P.fronts[ps.front];
    double* x = s.w.data() + F.wofs + ps.o;
    const double* T = s.F.data() + ps.fofs + (int64_t)ps.o + (int64_t)ps.o * ps.ld;
    const int w = ps.w, ld = ps.ld;
    if (P.lu) {
        int s0 = 0;
        for (int b = 0; b < ps.nsub; ++b) {
            int s1 = s0 + P.subw[ps.sub0 + b];
            for (int k = s0; k < s1; ++k) { int q = s0 + s.ipiv[ps.col0 + k] - 1; if (q != k) std::swap(x[k], x[q]); }
            for (int k = s0; k < s1; ++k) for (int i = k + 1; i < w; ++i) x[i] -= x[k] * T[i + (size_t)k * ld];
            s0 = s1;
        }
    } else {
        for (int k = 0; k < w; ++k) for (int i = k + 1; i < w; ++i) x[i] -= x[k] * T[i + (size_t)k * ld];
    }
}
static void pf_update(Sim& s, const PStep& ps) {
    const Front& F = s.P.fronts[ps.front];
    double* wf = s.w.data() + F.wofs;
    const double* Fm = s.F.data() + ps.fofs;
    for (int r = ps.o + ps.w; r < ps.R; ++r) {
        double acc = 0;
        for (int k = 0; k < ps.w; ++k) acc += Fm[(size_t)r + (size_t)(ps.o + k) * ps.ld] * wf[ps.o + k];
        wf[r] -= acc;
    }
}
static void pb_step(Sim& s, const PStep& ps, double* rhs) {
    Plan& P = s.P; const Front& F = P.fronts[ps.front]; const bool lu = P.lu;
    double* wf = s.w.data() + F.wofs; double* x = wf + ps.o;
    const double* Fm = s.F.data() + ps.fofs;
    const double* T = Fm + (int64_t)ps.o + (int64_t)ps.o * ps.ld;
    const int w = ps.w, ld = ps.ld;
    for (int k = 0; k < w; ++k) {
        double sum = 0;
        for (int r = ps.o + w; r < ps.R; ++r) sum += (lu ? Fm[(size_t)(ps.o + k) + (size_t)r * ld] : Fm[(size_t)r + (size_t)(ps.o + k) * ld]) * wf[r];
        x[k] = lu ? x[k] - sum : x[k] / T[k + (size_t)k * ld] - sum;
    }
    for (int k = w - 1; k >= 0; --k) {
        if (lu) { x[k] /= T[k + (size_t)k * ld]; for (int i = 0; i < k; ++i) x[i] -= x[k] * T[i + (size_t)k * ld]; }
        else for (int i = 0; i < k; ++i) x[i] -= x[k] * T[k + (size_t)i * ld];
    }
    for (int k = 0; k < w; ++k) rhs[ps.col0 + k] = x[k];
}

static int64_t run_solve_list(Sim* s, const std::vector<Launch>& Ls, double* rhs) {
    Plan& P = s->P; double* w = s->w.data();
    for (const Launch& L : Ls) {
        const int32_t* list = P.gathert.data() + L.first;
        for (int ti = 0; ti < L.count; ++ti) {
            switch (L.kind) {
            case K_FWD_GATHER: {
                const Front& F = P.fronts[list[ti]]; double* wf = w + F.wofs;
                for (int i = 0; i < F.R; ++i) wf[i] = i < F.W ? rhs[F.F0 + i] : 0.0;
                for (int r = 0; r < F.nchild; ++r) { const Front& C = P.fronts[P.childlist[F.child0 + r]]; for (int i = 0; i < C.m; ++i) wf[P.rel[C.relofs + i]] += w[C.wofs + C.W + i]; }
                break; }
            case K_FWD_DIAG: fwd_diag(*s, P.solvet[list[ti]]); break;
            case K_FWD_UPDATE: fwd_update(*s, P.solvet[list[ti]]); break;
            case K_FWD_FRONT: { const Front& F = P.fronts[list[ti]]; for (int tc = 0; tc < F.nch; ++tc) { fwd_diag(*s, P.solvet[F.c0 + tc]); fwd_update(*s, P.solvet[F.c0 + tc]); } break; }
            case K_BWD_GATHER: { const Front& F = P.fronts[list[ti]]; const Front& Pa = P.fronts[F.parent]; for (int i = 0; i < F.m; ++i) w[F.wofs + F.W + i] = w[Pa.wofs + P.rel[F.relofs + i]]; break; }
            case K_BWD_UPDATE: bwd_update(*s, P.solvet[list[ti]]); break;
            case K_BWD_DIAG: bwd_diag(*s, P.solvet[list[ti]], rhs); break;
            case K_BWD_FRONT: { const Front& F = P.fronts[list[ti]]; for (int tc = F.nch - 1; tc >= 0; --tc) { bwd_update(*s, P.solvet[F.c0 + tc]); bwd_diag(*s, P.solvet[F.c0 + tc], rhs); } break; }
            case K_PF_FRONT: { const Front& F = P.fronts[list[ti]]; for (int j = 0; j < F.nps; ++j) { pf_diag(*s, P.psteps[F.ps0 + j]); pf_update(*s, P.psteps[F.ps0 + j]); } break; }
            case K_PF_DIAG: pf_diag(*s, P.psteps[list[ti]]); break;
            case K_PF_UPDATE: pf_update(*s, P.psteps[list[ti]]); break;
            case K_PB_FRONT: { const Front& F = P.fronts[list[ti]]; for (int j = F.nps - 1; j >= 0; --j) pb_step(*s, P.psteps[F.ps0 + j], rhs); break; }
            case K_PB_UPDATE: break;                          // folded into pb_step at the K_PB_DIAG launch
            case K_PB_DIAG: pb_step(*s, P.psteps[list[ti]], rhs); break;
            case K_PF_FLOW: {                                 // blocks of one front are consecutive: run the front once, in step order
                const FlowTask& ft = P.flowt[L.first + ti];
                if (ft.jb > ft.ja && ft.ja == 0) { const Front& F = P.fronts[ft.front]; for (int j = 0; j < F.nps; ++j) { pf_diag(*s, P.psteps[F.ps0 + j]); pf_update(*s, P.psteps[F.ps0 + j]); } }
                break; }
            case K_PB_FLOW: {
                const FlowTask& ft = P.flowt[L.first + ti];
                const Front& F = P.fronts[ft.front];
                if (ft.jb == F.nps) for (int j = F.nps - 1; j >= 0; --j) pb_step(*s, P.psteps[F.ps0 + j], rhs);
                break; }
            case K_PF_STEP: { const PStep& ps = P.psteps[list[ti]]; pf_update(*s, ps); const Front& F = P.fronts[ps.front]; if (list[ti] + 1 < F.ps0 + F.nps) pf_diag(*s, P.psteps[list[ti] + 1]); break; }
            case K_PB_STEP: pb_step(*s, P.psteps[list[ti]], rhs); break;
            default: return -100;
            }
        }
    }
    return 0;
}
// which: 0 fwd local, 1 top (fwd + bwd), 2 bwd local   (multi-part plans)
API int64_t sim_solve_phase(void* h, double* rhs, int which) {
    Sim* s = (Sim*)h; Plan& P = s->P;
    if (which == 0) return run_solve_list(s, P.fwd_local, rhs);
    if (which == 1) { int64_t rc = run_solve_list(s, P.fwd_top, rhs); return rc ? rc : run_solve_list(s, P.bwd_top, rhs); }
    return run_solve_list(s, P.bwd_local, rhs);
}
// rhs in permuted order, in place; factors as left by sim_factor
API int64_t sim_solve(void* h, double* rhs) {
    Sim* s = (Sim*)h; Plan& P = s->P;
    int64_t rc = run_solve_list(s, P.fwd_launches, rhs);
    return rc ? rc : run_solve_list(s, P.bwd_launches, rhs);
}
