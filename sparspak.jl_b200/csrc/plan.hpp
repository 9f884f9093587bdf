// plan.hpp — host-side static analysis of a symbolic factorisation into a device schedule.
//
// Input: the reference's flat 1-based arrays (xsuper, snode, xlindx, lindx, xlnz, xunz;
// SpkSparseBase.jl:99-125).  Output:
//   * FRONTS: chains of consecutive reference supernodes ("chunks") whose row structures nest
//     (exactly = the fundamental supernodes the reference split at ~maxblocksize,
//     SpkSymFct.jl:415-464; or up to a few extra rows = relaxed chains, which absorb the long
//     runs of one-column supernodes a 7-point nested dissection produces).  Each front is a dense
//     R x R frontal matrix (column-major, padded ld) in a private device arena:
//         [ F11 (W x W)   F12 = U12 ]      W = columns of the front,
//         [ F21 = L21     S         ]      S = update matrix handed to the parent front.
//   * the front tree with relative-index maps (child below-rows -> parent rows),
//   * per-chunk position maps (stored row of the reference layout -> front row),
//   * per-level LAUNCH lists of independent TASKS.
// The reference's own layout (lnz / unz) is what goes in and what comes out: values are
// gathered into the fronts and the factors are scattered back, so the caller sees exactly
// the arrays `_lufactor!` / `_ldltfactor!` would have produced.
//
// Pure C++ (no CUDA) so tests can execute the same task lists on the host
// (tests/hostsim) to validate the schedule independently of the kernels.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <memory>
#include <utility>
#include <thread>
#include <atomic>
#include <algorithm>

namespace spk {

// std::vector allocator whose default construction leaves trivially constructible elements uninitialised
// (resize() without the zero fill: the position maps are 0.5 GB at 96^3 and every entry is written right after)
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    NoInitAlloc() = default;
    template <class U> NoInitAlloc(const NoInitAlloc<U>&) {}
    template <class U, class... A> void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U; else ::new ((void*)p) U(std::forward<A>(a)...);
    }
};

struct Chunk {            // one reference supernode
    int64_t fj;           // first column (0-based)
    int64_t lofs, uofs;   // start of its block in lnz / unz (0-based)
    int64_t posofs;       // pos[posofs + i] = front row of stored row i (i < jlen)
    int64_t fofs;         // owning front's matrix
    int32_t nj, jlen;     // width, rows
    int32_t front, o;     // owning front, column offset inside the front
    int32_t ld, pad;      // owning front's leading dimension
};

struct Front {
    int64_t F0;           // first column (0-based)
    int64_t fofs;         // frontal matrix in the arena (doubles), ld below
    int64_t relofs;       // rel[relofs + i] = row of below-row i in the PARENT front
    int64_t wofs;         // solve work vector (R entries)
    int32_t c0, nch;      // chunk range
    int32_t W, R, m, ld;  // columns, rows, m = R - W, leading dimension (R padded)
    int32_t parent, level;
    int32_t child0, nchild;
    int32_t ps0, nps;     // panel steps
    int64_t pbofs;        // partial-sum scratch of the backward sweep (large fronts)
};

// One panel step of the dense partial factorisation of a front: columns [o, o+w), made of
// `nsub` consecutive chunks (pivoting is restricted to each chunk's own diagonal block,
// SpkLUFactor.jl:230).  ob_end = first column after the outer block this step belongs to.
struct PStep {
    int64_t fofs, col0;   // front matrix; global first column (ipiv index)
    int32_t ld, R, o, w, ob_end, sub0, nsub, front;
    int64_t iofs;         // its w x w slot in the arrays of inverted diagonal blocks (solve)
};

struct GemmTask {         // C -= A * B   inside one frontal matrix (offsets relative to the arena)
    int64_t a0, b0, c0;   // element offsets of A(0,0) [A(i,k) at a0 + i + k*ld], B(0,0) [B(k,n) at b0 + k + n*ld], C(0,0)
    int32_t ld, m, n, k;
    int32_t roff;         // row index of C(0,0) minus its column index inside the front (for `lower`)
    uint8_t lower;        // only entries on/below the front's diagonal are needed / written (LDL^T)
    uint8_t pad0, pad1, pad2;
};
struct GemmTile { int32_t task; uint16_t ti, tj; };   // one C tile of a DMMA launch: task (relative to the launch), tile row / column
struct AsmTask { int32_t child, parent; };
struct SolveTask {        // one chunk of a triangular sweep (values read from lnz / unz)
    int64_t lofs, uofs, col0, wofs, posofs;
    int32_t ld, ldu, nj, m, o, front;
};

enum Kind : int32_t {
    K_ASM = 0, K_ASM_TAIL, K_DIAG, K_PANEL, K_GEMM, K_GEMM_B64 /* DMMA, 128 x 64 tiles */, K_GEMM_T64 /* DMMA, 64 x 64 tiles */,
    K_FWD_GATHER, K_FWD_DIAG, K_FWD_UPDATE, K_BWD_GATHER, K_BWD_UPDATE, K_BWD_DIAG, K_FWD_FRONT, K_BWD_FRONT,
    // solve on the frontal matrices, one step per PANEL STEP (dense, uniformly strided panels)
    K_PF_FRONT, K_PF_DIAG, K_PF_UPDATE, K_PB_FRONT, K_PB_UPDATE, K_PB_DIAG,
    K_PF_STEP, K_PB_STEP,     // fused: update + next diagonal block (forward), partial sums + diagonal block (backward)
    // distributed top set (multi-GPU LDL^T): broadcast of a column slab from its owner; U = D L^T rebuilt from a received panel
    K_BCAST, K_FILLU,
    // solve sweeps of the big fronts of a level as ONE flag-synchronised dataflow launch (k_pf_flow / k_pb_flow)
    K_PF_FLOW, K_PB_FLOW,
    // a whole small front (R <= FUSED_MAXR) in one thread block: load, extend-add, partial factorisation, store
    K_FRONT_SMALL
};
struct Launch {
    int32_t kind;
    int32_t first, count;     // task range in the kind's task array
    int32_t nblocks;          // total thread blocks
    int64_t pfx;              // block prefix: blkpfx[pfx .. pfx+count] (count+1 entries)
    double  flops;            // executed flops (GEMM launches)
    int32_t level, step;
    int32_t maxw;             // widest panel step in a DIAG / PANEL launch (sizes its shared memory)
    // look-ahead: stream 0 = panel stream (diag / panel / in-block updates / next-block strip),
    // stream 1 = trailing-update stream (the bulk GEMMs).  Listed order is always a valid serial order.
    // DMMA launches (persistent kernel over a tile list): tiles [tile0, tile0 + ntiles) of Plan::tiles, atomic tile
    // counter `ctr`, and `reserve`: leave that many block slots of the machine free (trailing updates that run beside
    // the latency-critical diagonal / panel chain of the next outer block)
    int64_t tile0; int32_t ntiles, ctr; int32_t reserve;
    int32_t tile_m = 0, tile_n = 0;   // C tile of a DMMA launch (128 x 64, 64 x 64, or 64 x 32 for launches that do not fill the machine)
    uint8_t stream, wait_other, record, wait_mask;   // wait_other: wait for the other stream's last record first
    // distributed lists use three streams (0 panel, 1 trailing update, 2 communication) and wait_mask: bit s = wait
    // for the last record of stream s before launching
};
// one thread block of a dataflow solve launch: the pivot rows of panel steps [ja, jb) of a front (jb > ja), or a slab
// [r0, r1) of the rows below its columns (forward sweep only)
struct FlowTask { int32_t front, ja, jb, r0, r1; };
struct Bcast { int64_t ofs, len; int32_t root, front; };          // in-place broadcast of F[ofs, ofs+len) from part `root`
struct FillTask { int64_t fofs; int32_t ld, R, ob0, e; };          // U[ob0+k, c] = D_k * L[c, ob0+k] for k < e-ob0, e <= c < R

constexpr int ASM_ROUNDS = 8;      // children handled by per-round launches; the rest by a tail kernel
constexpr int ASM_TPB = 256, ASM_EPT = 4, ASM_COLS = 8;   // extend-add: rows per block, (tail kernel) entries per thread, columns per block
constexpr int PANEL_ROWS = 128;    // rows (L side) / columns (U side) per panel block
constexpr int GEMM_TM = 64, GEMM_TN = 64;   // C tile of the small-tile kernel
constexpr int BIG_TM = 128;                 // C tile rows of the big DMMA kernel (128 x 64 tiles; the small one: 64 x 64)
constexpr int DMMA_FILL = 296;              // a launch with fewer 128-row tiles than this (2 blocks x 148 SMs) uses 64-row tiles
constexpr int UPD_ROWS = 256;      // rows per block in the forward-solve update
constexpr int SV_ROWS = 256;       // rows / columns per block in the panel-step solve kernels
constexpr int BWD_COLS = 1;        // columns per block in the backward-solve update (one block reduces one column)
constexpr int PS_WIDTH = 64;       // target panel-step width (a wider single chunk stays alone)
constexpr int OB_WIDTH = 512;      // target outer-block width (delayed trailing update)
constexpr int OB_STEPS = 8;        // panel steps per outer block (aligned across the fronts of a level for the look-ahead)
constexpr int FUSED_MAXR = 64;     // fronts of at most this many rows are factored by the fused one-block kernel (k_front_small)
constexpr int FLOW_MIN_STEPS = 8;  // fronts with at least this many panel steps take the dataflow solve kernels
constexpr int FLOW_ROWS = 128;     // rows per thread block of the dataflow solve kernels (one per thread; wider single steps: two)
constexpr int64_t SOLVE_SMALL = 65536;  // fronts with at most this many stored L entries are solved by one block
constexpr int RELAX_ABS = 4;       // a chunk joins the chain if it adds at most this many rows ...
constexpr double RELAX_FRAC = 0.02;//   ... or this fraction of its rows

struct Plan {
    bool lu = false;
    int64_t n = 0, nsuper = 0, nsub = 0, nlnz = 0, nunz = 0;
    std::vector<Chunk> chunks;
    std::vector<Front> fronts;
    std::vector<PStep> psteps;
    std::vector<int32_t> subw;                // pivot sub-block widths of all panel steps
    std::vector<int32_t> childlist;
    std::vector<int32_t> rel;                 // relative indices, all fronts
    std::vector<int32_t, NoInitAlloc<int32_t>> pos;   // per-chunk stored-row -> front-row maps (every entry written by analyze(): not zero-filled first)
    std::vector<int32_t> col2chunk;
    int64_t arena = 0, wlen = 0, pblen = 0, tinv_len = 0;
    bool solve_on_fronts = true;              // solve sweeps read the frontal matrices (panel steps) instead of lnz/unz (chunks)
    int32_t nlevels = 0, maxnj = 0, maxR = 0, maxpw = 0;
    double flops_struct = 0, nnzL = 0;        // sum cc^2 (or 2 sum cc^2 - sum cc), sum cc
    bool use_dmma = true;
    int64_t solve_small = SOLVE_SMALL;
    int ob_width = OB_WIDTH, ps_width = PS_WIDTH, ob_steps = OB_STEPS;
    bool lookahead = true;
    bool split_rest = false;                   // SPK_SPLIT_REST=1: the delayed trailing update of an outer block in two launches; the next strips wait for the first only (measured: no gain, the stream is never idle)
    int32_t dmma_narrow = 296;                 // SPK_DMMA_NARROW: DMMA launches with fewer 64 x 64 tiles than this (2 per SM) run 64 x 32 tiles (0 = never)
    bool dmma_big = false;                     // SPK_DMMA_BIG=1: 128 x 64 DMMA tiles for launches with >= DMMA_FILL of them
    bool left_inblock = true;                  // SPK_LL=0: right-looking rank-w updates inside an outer block (LU always)
    int relax_abs = RELAX_ABS; double relax_frac = RELAX_FRAC;
    // schedules
    std::vector<AsmTask> asmt;
    std::vector<GemmTask> gemmt;
    std::vector<GemmTile> tiles;              // tile lists of the DMMA launches
    int32_t nctr = 0;                         // atomic tile counters (one per DMMA launch)
    int32_t gemm_reserve = 32;                // SPK_GEMM_RESERVE
    std::vector<int32_t> pslist;              // panel-step ids, grouped per DIAG/PANEL launch
    std::vector<SolveTask> solvet;            // one per chunk
    std::vector<int32_t> gathert;
    std::vector<FlowTask> flowt;              // thread blocks of the dataflow solve launches
    int32_t nflowctr = 0;                     // ticket counters (one per dataflow launch)
    bool levels_by_depth = true;              // SPK_LEVELS=height: levels by height above the leaves
    int32_t fused_maxr = FUSED_MAXR;          // SPK_FUSED_MAXR (0 = off)
    int32_t flow_min_steps = FLOW_MIN_STEPS;  // SPK_FLOW_MIN_STEPS
    // SPK_SOLVE_FLOW=1: dataflow sweeps (one flag-free, mailbox-synchronised launch per level for the fronts with many
    // panel steps).  MEASURED SLOWER than one launch per panel step (96^3: 20.4 vs 13.6 ms): a step costs ~11 us either
    // way, of which ~3.5 us is the in-block triangular solve — the launch is not what bounds the chain.  Off by default.
    bool solve_flow = false;
    std::vector<int32_t> blkpfx;
    std::vector<Launch> factor_launches, fwd_launches, bwd_launches;
    // multi-GPU (elimination-subtree partition): owner[f] = part that factors front f, or -1 for the TOP SET
    // (the ancestors of all subtree roots), which every part factors redundantly after the exchange.
    int32_t part = 0, nparts = 1;
    int32_t force_splits = -1;                 // SPK_TOP_SPLITS=k: exactly k fronts split off the top of the tree (experiments)
    int32_t max_subtrees = 1 << 30;            // SPK_MAX_SUBTREES: cap on the number of subtrees (tests: parts left without one)
    std::vector<int32_t> owner;
    std::vector<int32_t> xchg;                 // subtree-root fronts (owner >= 0, parent in the top set)
    struct Range { int32_t owner; int32_t f0, f1; int64_t lnz0, lnz1, unz0, unz1, col0, col1; };
    std::vector<Range> ranges;                 // contiguous front / storage ranges owned by one part
    std::vector<Launch> factor_local, factor_top, fwd_local, fwd_top, bwd_top, bwd_local;
    // DISTRIBUTED TOP SET (LDL^T, nparts > 1): the columns of every top-set front are dealt to the parts by OUTER
    // BLOCK (block q of front f -> part (q + shift_f) mod nparts); the ownership of a front's update-matrix columns
    // follows the ancestor column they are added to, so extend-adds between top-set fronts stay local.  Only the
    // owner factors a block's panel; the factored column slab is broadcast in place (every part keeps a full copy
    // of the top-set FACTORS, so the solve and the write-back stay as they are) and every part applies the
    // delayed rank-(block width) update to the column blocks it owns.
    bool dist_top = false; int dist_top_env = -1;      // SPK_DIST_TOP=0/1 overrides the default (on for LDL^T with nparts > 1)
    std::vector<int8_t> fown;                  // owner of every front column, all top-set fronts (DFront::ownofs)
    std::vector<int32_t> ownofs;               // per front: offset into fown, or -1 (not a distributed front)
    std::vector<Bcast> bcasts;
    std::vector<FillTask> fillt;
    std::vector<uint8_t> held;                 // fronts this part keeps storage for (its subtrees, the top set, the exchanged subtree roots)
    // single GPU, TREE PIPELINES: the front tree is cut like the multi-GPU partition into `pipes` sets of
    // subtrees plus a top set; the sets are factored CONCURRENTLY, each on its own (panel, update) stream
    // pair, then the top set.  Within one pipeline a level ends in a latency-bound tail (the last outer
    // blocks of its fronts: a chain of small dependent kernels with the update stream idle); the other
    // pipelines' trailing updates fill it.  Same arithmetic per front, so the factors do not change.
    // MEASURED (96^3, SPK_PIPES=2..4): no gain — a trace of the two-stream run (SPK_TRACE) shows the GPU busy
    // with >= 148-block kernels for 214 of 232 ms, i.e. the factorisation is throughput-, not schedule-bound.
    // Off by default (SPK_PIPES=1).
    int32_t pipes = 1;
    std::vector<std::vector<Launch>> factor_pipe;
    std::vector<Launch> factor_ptop;
    std::string error;
};

inline int32_t cdiv(int64_t a, int64_t b) { return (int32_t)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------
inline bool analyze(Plan& P, int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                    const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz, const int64_t* xunz) {
    (void)snode;
    P.lu = (xunz != nullptr);
    P.n = n; P.nsuper = nsuper;
    if (n <= 0 || nsuper <= 0) { P.error = "empty problem"; return false; }
    P.nsub = xlindx[nsuper] - 1;
    P.nlnz = xlnz[n] - 1;
    P.nunz = P.lu ? xunz[n] - 1 : 0;

    P.chunks.resize(nsuper);
    int64_t posofs = 0;
    for (int64_t s = 0; s < nsuper; ++s) {
        Chunk& c = P.chunks[s];
        c.fj = xsuper[s] - 1;
        c.nj = (int32_t)(xsuper[s + 1] - xsuper[s]);
        c.jlen = (int32_t)(xlindx[s + 1] - xlindx[s]);
        if ((int64_t)c.jlen != xlnz[c.fj + 1] - xlnz[c.fj] || c.nj <= 0 || c.jlen < c.nj) { P.error = "inconsistent supernode storage"; return false; }
        c.lofs = xlnz[c.fj] - 1;
        c.uofs = P.lu ? xunz[c.fj] - 1 : 0;
        c.posofs = posofs; posofs += c.jlen;
        P.maxnj = std::max(P.maxnj, c.nj);
    }
    P.pos.resize(posofs);
    P.col2chunk.assign(n, 0);
    for (int64_t s = 0; s < nsuper; ++s) for (int32_t j = 0; j < P.chunks[s].nj; ++j) P.col2chunk[P.chunks[s].fj + j] = (int32_t)s;

    // fronts: chunk e+1 joins the chain of chunk e iff it is e's parent (first below-row of e is
    // its first column) and it adds at most a few rows to what e hands down.
    for (int64_t s = 0; s < nsuper;) {
        Front f{};
        f.c0 = (int32_t)s; f.F0 = P.chunks[s].fj;
        int32_t o = 0; int64_t e = s;
        for (;;) {
            Chunk& c = P.chunks[e];
            c.front = (int32_t)P.fronts.size(); c.o = o; o += c.nj;
            bool more = false;
            if (e + 1 < nsuper && c.jlen > c.nj && lindx[xlindx[e] - 1 + c.nj] == P.chunks[e + 1].fj + 1) {
                const Chunk& d = P.chunks[e + 1];
                int32_t extra = d.jlen - (c.jlen - c.nj);
                int32_t lim = std::max<int32_t>(P.relax_abs, (int32_t)(P.relax_frac * d.jlen));
                if (extra <= lim) more = true;
            }
            ++e;
            if (!more) break;
        }
        f.nch = (int32_t)(e - s); f.W = o;
        const Chunk& last = P.chunks[e - 1];
        f.m = last.jlen - last.nj; f.R = f.W + f.m;
        f.ld = (f.R + 1) & ~1;                     // even leading dimension (16-byte aligned columns)
        f.parent = -1;
        P.maxR = std::max(P.maxR, f.R);
        P.fronts.push_back(f);
        s = e;
    }
    const int32_t nf = (int32_t)P.fronts.size();
    // per-chunk position maps: own columns map to o..o+nj-1; below rows are located in the front's
    // row list = [columns of the chain] ++ [below rows of the last chunk].  One entry per stored row of every chunk
    // (96^3: 1.3e8): fronts are independent and write disjoint ranges, so the loop is dealt to host threads in
    // contiguous front ranges of about equal size (the result does not depend on the thread count).
    {
        std::atomic<int> bad{0};                        // 1 = row above supernode, 2 = chain rows do not nest
        auto work = [&](int32_t f0, int32_t f1) {
            for (int32_t f = f0; f < f1 && !bad.load(std::memory_order_relaxed); ++f) {
                const Front& F = P.fronts[f];
                const Chunk& last = P.chunks[F.c0 + F.nch - 1];
                const int64_t* below = lindx + (xlindx[F.c0 + F.nch - 1] - 1) + last.nj;   // m entries, 1-based, sorted
                for (int32_t t = 0; t < F.nch; ++t) {
                    const Chunk& c = P.chunks[F.c0 + t];
                    int32_t* pos = P.pos.data() + c.posofs;
                    const int64_t* rows = lindx + (xlindx[F.c0 + t] - 1);
                    int32_t q = 0;
                    for (int32_t i = 0; i < c.jlen; ++i) {
                        int64_t r = rows[i] - 1;            // 0-based global row
                        if (r < F.F0 + F.W) {               // a column of the chain
                            if (r < c.fj) { bad.store(1); return; }
                            pos[i] = (int32_t)(r - F.F0);
                        } else {
                            while (q < F.m && below[q] - 1 < r) ++q;
                            if (q >= F.m || below[q] - 1 != r) { bad.store(2); return; }
                            pos[i] = F.W + q;
                        }
                    }
                }
            }
        };
        int nth = 1;
        if (posofs > (int64_t)1 << 22) {
            nth = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
            if (const char* e = getenv("SPK_HOST_THREADS")) nth = std::max(1, atoi(e));
        }
        if (nth <= 1) work(0, nf);
        else {
            std::vector<std::thread> th;
            int32_t f0 = 0;
            for (int t = 0; t < nth && f0 < nf; ++t) {
                const int64_t target = posofs / nth * (t + 1);
                int32_t f1 = f0;
                while (f1 < nf && (t == nth - 1 || P.chunks[P.fronts[f1].c0].posofs < target)) ++f1;
                if (f1 > f0) { try { th.emplace_back(work, f0, f1); } catch (...) { work(f0, f1); } }   // no thread to be had: this range inline
                f0 = f1;
            }
            for (auto& x : th) x.join();
        }
        if (bad.load() == 1) { P.error = "row above supernode"; return false; }
        if (bad.load() == 2) { P.error = "chain rows do not nest"; return false; }
    }
    // front tree, relative indices, arena
    int64_t relofs = 0, fofs = 0, wofs = 0;
    std::vector<int32_t> nchild(nf, 0);
    for (int32_t f = 0; f < nf; ++f) {
        Front& F = P.fronts[f];
        F.relofs = relofs; F.fofs = fofs; F.wofs = wofs;
        relofs += F.m; fofs += (int64_t)F.ld * F.R; fofs = (fofs + 1) & ~(int64_t)1; wofs += F.R;
        for (int32_t t = 0; t < F.nch; ++t) { P.chunks[F.c0 + t].fofs = F.fofs; P.chunks[F.c0 + t].ld = F.ld; }
        if (F.m > 0) {
            const Chunk& last = P.chunks[F.c0 + F.nch - 1];
            int64_t prow = lindx[(xlindx[F.c0 + F.nch - 1] - 1) + last.nj] - 1;
            F.parent = P.chunks[P.col2chunk[prow]].front;
            if (F.parent <= f) { P.error = "front tree not topologically ordered"; return false; }
            ++nchild[F.parent];
        }
    }
    P.arena = fofs; P.wlen = wofs;
    P.rel.assign(relofs, 0);
    for (int32_t f = 0; f < nf; ++f) {
        const Front& F = P.fronts[f];
        if (F.m == 0) continue;
        const Front& Pa = P.fronts[F.parent];
        const Chunk& last = P.chunks[F.c0 + F.nch - 1];
        const int64_t* fr = lindx + (xlindx[F.c0 + F.nch - 1] - 1) + last.nj;
        const Chunk& plast = P.chunks[Pa.c0 + Pa.nch - 1];
        const int64_t* pbelow = lindx + (xlindx[Pa.c0 + Pa.nch - 1] - 1) + plast.nj;
        int32_t q = 0;
        for (int32_t i = 0; i < F.m; ++i) {
            int64_t r = fr[i] - 1;
            if (r < Pa.F0 + Pa.W) {
                if (r < Pa.F0) { P.error = "child row before parent"; return false; }
                P.rel[F.relofs + i] = (int32_t)(r - Pa.F0);
            } else {
                while (q < Pa.m && pbelow[q] - 1 < r) ++q;
                if (q >= Pa.m || pbelow[q] - 1 != r) { P.error = "child rows not contained in parent rows"; return false; }
                P.rel[F.relofs + i] = Pa.W + q;
            }
        }
    }
    int32_t acc = 0;
    for (int32_t f = 0; f < nf; ++f) { P.fronts[f].child0 = acc; acc += nchild[f]; P.fronts[f].nchild = 0; P.fronts[f].level = 0; }
    P.childlist.assign(acc, 0);
    for (int32_t f = 0; f < nf; ++f) {
        int32_t p = P.fronts[f].parent;
        if (p >= 0) {
            Front& Pa = P.fronts[p];
            P.childlist[Pa.child0 + Pa.nchild++] = f;
            Pa.level = std::max(Pa.level, P.fronts[f].level + 1);
        }
    }
    P.nlevels = 0;
    for (int32_t f = 0; f < nf; ++f) P.nlevels = std::max(P.nlevels, P.fronts[f].level + 1);
    // Levels by DEPTH below the root (as late as possible) instead of height above the leaves: siblings whose subtrees
    // differ in height by one would otherwise sit alone in consecutive levels, and a level with a single big front is
    // bound by that front's chain of dependent panel steps (96^3: the four second-level separators end up in one
    // level instead of three).  SPK_LEVELS=height restores the height levels.
    if (P.levels_by_depth) {
        for (int32_t f = nf - 1; f >= 0; --f) {
            Front& F = P.fronts[f];
            F.level = F.parent < 0 ? P.nlevels - 1 : P.fronts[F.parent].level - 1;
        }
    }
    // panel steps: greedy groups of consecutive chunks up to PS_WIDTH columns; outer blocks up to OB_WIDTH
    for (int32_t f = 0; f < nf; ++f) {
        Front& F = P.fronts[f];
        F.ps0 = (int32_t)P.psteps.size();
        int32_t t = 0;
        while (t < F.nch) {
            PStep ps{}; ps.fofs = F.fofs; ps.ld = F.ld; ps.R = F.R; ps.front = f;
            ps.o = P.chunks[F.c0 + t].o; ps.col0 = P.chunks[F.c0 + t].fj; ps.sub0 = (int32_t)P.subw.size();
            int32_t w = 0, ns = 0;
            while (t < F.nch && (ns == 0 || w + P.chunks[F.c0 + t].nj <= P.ps_width)) { w += P.chunks[F.c0 + t].nj; P.subw.push_back(P.chunks[F.c0 + t].nj); ++ns; ++t; }
            ps.w = w; ps.nsub = ns;
            ps.iofs = P.tinv_len; P.tinv_len += (int64_t)w * w;
            P.maxpw = std::max(P.maxpw, w);
            P.psteps.push_back(ps);
        }
        F.nps = (int32_t)P.psteps.size() - F.ps0;
        int32_t j = 0;
        while (j < F.nps) {
            int32_t j0 = j, w = 0;
            while (j < F.nps && (j - j0) < P.ob_steps) { w += P.psteps[F.ps0 + j].w; ++j; }   // j0 is a multiple of ob_steps
            int32_t ob_end = P.psteps[F.ps0 + j0].o + w;
            for (int32_t q = j0; q < j; ++q) P.psteps[F.ps0 + q].ob_end = ob_end;
        }
    }
    // structural work: cc_j for column j of a chunk = jlen - j
    double s1 = 0, s2 = 0;
    for (int64_t s = 0; s < nsuper; ++s) {
        const Chunk& c = P.chunks[s];
        for (int32_t j = 0; j < c.nj; ++j) { double cc = c.jlen - j; s1 += cc; s2 += cc * cc; }
    }
    P.nnzL = s1;
    P.flops_struct = P.lu ? 2.0 * s2 - s1 : s2;
    return true;
}

// ---------------------------------------------------------------------------------------
struct LaunchBuilder {
    Plan& P; std::vector<Launch>& out;
    Launch cur{};
    LaunchBuilder(Plan& p, std::vector<Launch>& o) : P(p), out(o) {}
    void begin(int32_t kind, int32_t first, int32_t level, int32_t step, int stream = 0, int wait_other = 0, int record = 0) {
        cur = Launch{}; cur.kind = kind; cur.first = first; cur.count = 0; cur.nblocks = 0;
        cur.stream = (uint8_t)stream; cur.wait_other = (uint8_t)wait_other; cur.record = (uint8_t)record;
        cur.pfx = (int64_t)P.blkpfx.size(); cur.flops = 0; cur.level = level; cur.step = step;
        P.blkpfx.push_back(0);
    }
    void add(int32_t nblocks, double flops = 0, int32_t w = 0) {
        cur.count++; cur.nblocks += nblocks; cur.flops += flops; P.blkpfx.push_back(cur.nblocks);
        cur.maxw = std::max(cur.maxw, w);
    }
    void end() {
        if (cur.count > 0 && cur.nblocks > 0) out.push_back(cur);
        else P.blkpfx.resize(cur.pfx);
    }
};

inline int32_t gemm_blocks(const GemmTask& t, int tm, int tn) { return cdiv(t.m, tm) * cdiv(t.n, tn); }

// Tiles of a DMMA task that have work.  The kernel moves a task's origin to the previous even row (sa): tile rows
// are counted from there.  LDL^T updates skip the tiles strictly above the diagonal.
inline void dmma_tiles(const GemmTask& t, int tm, int tn, int32_t task_rel, std::vector<GemmTile>* out, int64_t* count) {
    const int sa = (int)(t.a0 & 1);
    const int mp = t.m + sa, roffp = t.roff - sa;
    const int mt = cdiv(mp, tm), nt = cdiv(t.n, tn);
    for (int tj = 0; tj < nt; ++tj)
        for (int ti = 0; ti < mt; ++ti) {
            if (t.lower && ti * tm + tm - 1 + roffp < tj * tn) continue;
            if (out) out->push_back(GemmTile{task_rel, (uint16_t)ti, (uint16_t)tj});
            if (count) ++*count;
        }
}

struct GemmBatch {
    struct Item { GemmTask t; double flops; };
    std::vector<Item> small, dmma;
    void add(const Plan& P, const GemmTask& t, double flops) {
        if (t.m <= 0 || t.n <= 0 || t.k <= 0) return;
        if (P.use_dmma && t.m >= 64 && t.n > 16) dmma.push_back({t, flops});
        else small.push_back({t, flops});
    }
    bool empty() const { return small.empty() && dmma.empty(); }
    // every launch of the batch waits for the other stream (cheap) and records its own completion
    void emit(Plan& P, LaunchBuilder& fb, int32_t lev, int32_t step, int stream = 0, int wait_other = 0, int record = 0, int reserve = 0) {
        if (!small.empty()) {
            fb.begin(K_GEMM, (int32_t)P.gemmt.size(), lev, step, stream, wait_other, record);
            for (const Item& it : small) { P.gemmt.push_back(it.t); fb.add(gemm_blocks(it.t, GEMM_TM, GEMM_TN), it.flops, std::max(it.t.m, it.t.n)); }   // maxw = largest C dimension of the launch
            fb.end();
            small.clear();
        }
        if (!dmma.empty()) {
            int64_t big = 0;
            for (const Item& it : dmma) dmma_tiles(it.t, BIG_TM, 64, 0, nullptr, &big);
            const bool t64 = !P.dmma_big || big < DMMA_FILL;        // default: 64 x 64 tiles everywhere (measured faster); SPK_DMMA_BIG=1: 128-row tiles when they fill the machine
            const int tm = t64 ? 64 : BIG_TM;
            // launches that would leave SM slots empty even with 64 x 64 tiles (the in-block updates of the top fronts, on the
            // critical path of every panel step): 64 x 32 tiles, twice the blocks (tools/ubench_dmma.cu: 37 vs 43 us at
            // 13000 x 64 x 400)
            int64_t n64 = 0;
            if (t64 && P.dmma_narrow > 0) for (const Item& it : dmma) dmma_tiles(it.t, 64, 64, 0, nullptr, &n64);
            const int tn = (t64 && P.dmma_narrow > 0 && n64 < P.dmma_narrow) ? 32 : 64;
            fb.begin(t64 ? K_GEMM_T64 : K_GEMM_B64, (int32_t)P.gemmt.size(), lev, step, stream, wait_other, record);
            fb.cur.tile0 = (int64_t)P.tiles.size(); fb.cur.ctr = P.nctr++; fb.cur.reserve = reserve;
            fb.cur.tile_m = tm; fb.cur.tile_n = tn;
            int32_t rel = 0;
            for (const Item& it : dmma) {
                P.gemmt.push_back(it.t);
                const size_t n0 = P.tiles.size();
                dmma_tiles(it.t, tm, tn, rel++, &P.tiles, nullptr);
                fb.add((int32_t)(P.tiles.size() - n0), it.flops, std::max(it.t.m, it.t.n));
            }
            fb.cur.ntiles = (int32_t)(P.tiles.size() - (size_t)fb.cur.tile0);
            fb.end();
            dmma.clear();
        }
    }
};

// C(r0.., c0..) [m x n] -= A(r0.., k0..k0+k) * B(k0.., c0..)   inside front F
inline GemmTask front_gemm(const Plan& P, const Front& F, int32_t r0, int32_t m, int32_t c0, int32_t n, int32_t k0, int32_t k) {
    GemmTask g{};
    g.ld = F.ld; g.m = m; g.n = n; g.k = k;
    g.a0 = F.fofs + (int64_t)r0 + (int64_t)k0 * F.ld;
    g.c0 = F.fofs + (int64_t)r0 + (int64_t)c0 * F.ld;
    // B is a row block of U.  LDL^T fronts keep U = D * L^T in their (otherwise unused) upper triangle
    // (written by the panel kernel), so both factorisations run the same kernel with no scaling in the k loop.
    g.b0 = F.fofs + (int64_t)k0 + (int64_t)c0 * F.ld;
    g.lower = P.lu ? 0 : 1;
    g.roff = r0 - c0;
    return g;
}
inline double gemm_flops(const GemmTask& g) {
    double f = 2.0 * g.m * g.n * g.k;
    if (g.lower && g.roff < g.n) {                 // entries above the diagonal are skipped
        double t = (double)(g.n - std::max(g.roff, 0));
        f -= t * t * g.k;
    }
    return f;
}

inline void plan_env_overrides(Plan& P) {        // test / tuning knobs
    if (const char* e = getenv("SPK_NO_DMMA")) P.use_dmma = !(e[0] == '1');
    if (const char* e = getenv("SPK_SOLVE_SMALL")) P.solve_small = atoll(e);
    if (const char* e = getenv("SPK_OB_STEPS")) P.ob_steps = std::max(1, atoi(e));
    if (const char* e = getenv("SPK_LOOKAHEAD")) P.lookahead = e[0] != '0';
    if (const char* e = getenv("SPK_LL")) P.left_inblock = e[0] != '0';
    if (const char* e = getenv("SPK_MAX_SUBTREES")) P.max_subtrees = std::max(2, atoi(e));
    if (const char* e = getenv("SPK_PIPES")) P.pipes = std::min(4, std::max(1, atoi(e)));
    if (const char* e = getenv("SPK_SOLVE_LNZ")) P.solve_on_fronts = e[0] != '1';
    if (const char* e = getenv("SPK_LEVELS")) P.levels_by_depth = e[0] != 'h';
    if (const char* e = getenv("SPK_FUSED_MAXR")) P.fused_maxr = std::min(FUSED_MAXR, std::max(0, atoi(e)));
    if (const char* e = getenv("SPK_SOLVE_FLOW")) P.solve_flow = e[0] != '0';
    if (const char* e = getenv("SPK_FLOW_MIN_STEPS")) P.flow_min_steps = std::max(1, atoi(e));
    if (const char* e = getenv("SPK_PS_WIDTH")) P.ps_width = std::max(1, atoi(e));
    if (const char* e = getenv("SPK_TOP_SPLITS")) P.force_splits = atoi(e);
    if (const char* e = getenv("SPK_GEMM_RESERVE")) P.gemm_reserve = std::max(0, atoi(e));
    if (const char* e = getenv("SPK_DIST_TOP")) P.dist_top_env = e[0] != '0';
    if (const char* e = getenv("SPK_SPLIT_REST")) P.split_rest = e[0] != '0';
    if (const char* e = getenv("SPK_DMMA_BIG")) P.dmma_big = e[0] != '0';
    if (const char* e = getenv("SPK_DMMA_NARROW")) P.dmma_narrow = std::max(0, atoi(e));
}

// Launch lists for the fronts selected by `sel` (all of them, one part's subtrees, or the top set).
inline void build_lists(Plan& P, const std::vector<uint8_t>& sel, std::vector<Launch>& factor_out,
                        std::vector<Launch>& fwd_out, std::vector<Launch>& bwd_out) {
    const int32_t nf = (int32_t)P.fronts.size();
    std::vector<std::vector<int32_t>> bylevel(P.nlevels);
    for (int32_t f = 0; f < nf; ++f) if (sel[f]) bylevel[P.fronts[f].level].push_back(f);
    const bool lu = P.lu;

    LaunchBuilder fb(P, factor_out);
    for (int32_t lev = 0; lev < P.nlevels; ++lev) {
        // ---- small fronts: one fused launch (load, extend-add, partial factorisation, store in one block per front)
        std::vector<int32_t> fr;                                  // the fronts that take the kernel-per-operation path
        fb.begin(K_FRONT_SMALL, (int32_t)P.pslist.size(), lev, 0, 0, 1, 1);
        for (int32_t f : bylevel[lev]) {
            if (P.fronts[f].R <= P.fused_maxr) { P.pslist.push_back(f); fb.add(1, 0, P.fronts[f].R); }
            else fr.push_back(f);
        }
        fb.end();
        int32_t maxch = 0, maxnps = 0;
        for (int32_t f : fr) { maxch = std::max(maxch, P.fronts[f].nchild); maxnps = std::max(maxnps, P.fronts[f].nps); }
        // ---- extend-add of the children's update matrices
        for (int32_t r = 0; r < std::min(maxch, ASM_ROUNDS); ++r) {
            fb.begin(K_ASM, (int32_t)P.asmt.size(), lev, r, 0, 1, 0);
            for (int32_t f : fr) {
                const Front& F = P.fronts[f];
                if (F.nchild <= r) continue;
                int32_t c = P.childlist[F.child0 + r];
                int64_t mc = P.fronts[c].m;
                P.asmt.push_back(AsmTask{c, f});
                fb.add((int32_t)(cdiv(mc, ASM_TPB) * cdiv(mc, ASM_COLS)));
            }
            fb.end();
        }
        if (maxch > ASM_ROUNDS) {
            fb.begin(K_ASM_TAIL, (int32_t)P.asmt.size(), lev, ASM_ROUNDS, 0, 1, 0);
            for (int32_t f : fr) if (P.fronts[f].nchild > ASM_ROUNDS) { P.asmt.push_back(AsmTask{-1, f}); fb.add(1); }
            fb.end();
        }
        // ---- dense partial factorisation, panel step by panel step
        for (int32_t j = 0; j < maxnps; ++j) {
            // LDL^T, inside an outer block: LEFT-looking — just before panel step j is factored its columns get the
            // updates of all earlier steps of the block in ONE GEMM (k = 57..400) instead of one rank-57 update
            // after every step (those ran at 11 TFLOP/s: 3.5 k-tiles between prologue and epilogue).
            const bool ll = P.left_inblock && !lu;
            if (ll && (j % P.ob_steps) != 0) {
                GemmBatch gl;
                for (int32_t f : fr) if (P.fronts[f].nps > j) {
                    const Front& F = P.fronts[f];
                    const PStep& ps = P.psteps[F.ps0 + j];
                    const int32_t ob0 = P.psteps[F.ps0 + j - (j % P.ob_steps)].o;
                    GemmTask g = front_gemm(P, F, ps.o, F.R - ps.o, ps.o, ps.w, ob0, ps.o - ob0);
                    gl.add(P, g, gemm_flops(g));
                }
                gl.emit(P, fb, lev, j, 0, 0, 0);
            }
            fb.begin(K_DIAG, (int32_t)P.pslist.size(), lev, j, 0, j == 0 ? 1 : 0, 1);
            for (int32_t f : fr) if (P.fronts[f].nps > j) { P.pslist.push_back(P.fronts[f].ps0 + j); fb.add(1, 0, P.psteps[P.fronts[f].ps0 + j].w); }
            fb.end();
            fb.begin(K_PANEL, (int32_t)P.pslist.size(), lev, j, 0, 0, 1);
            for (int32_t f : fr) if (P.fronts[f].nps > j) {
                const PStep& ps = P.psteps[P.fronts[f].ps0 + j];
                int32_t below = ps.R - ps.o - ps.w;
                if (below <= 0) continue;
                P.pslist.push_back(P.fronts[f].ps0 + j);
                fb.add(cdiv(below, PANEL_ROWS) * (lu ? 2 : 1), 0, ps.w);
            }
            fb.end();
            // Trailing updates.  Within an outer block (ob_steps panel steps) only the block's own remaining
            // columns (and rows, LU) are updated eagerly (panel stream).  At a block boundary the delayed
            // rank-(block width) update is split: the STRIP that the next block's panel steps need goes to the
            // panel stream, the REST (the bulk of the flops) to the trailing-update stream, where it overlaps
            // the diagonal / panel kernels of the next block (look-ahead).  A front's last step sends its whole
            // remaining update (the update matrix S) to the trailing-update stream.
            GemmBatch gp, ga, gg;
            const bool boundary = ((j + 1) % P.ob_steps) == 0;
            for (int32_t f : fr) if (P.fronts[f].nps > j) {
                const Front& F = P.fronts[f];
                const PStep& ps = P.psteps[F.ps0 + j];
                const int32_t e = ps.o + ps.w;             // first column after the panel
                if (e >= F.R) continue;
                const bool last = (F.nps == j + 1);
                if (!boundary && !last) {
                    if (ll) continue;                      // left-looking inside the block: nothing to do after the panel
                    GemmTask g = front_gemm(P, F, e, F.R - e, e, ps.ob_end - e, ps.o, ps.w);
                    gp.add(P, g, gemm_flops(g));
                    if (lu && ps.ob_end < F.R) {
                        GemmTask h = front_gemm(P, F, e, ps.ob_end - e, ps.ob_end, F.R - ps.ob_end, ps.o, ps.w);
                        gp.add(P, h, gemm_flops(h));
                    }
                    continue;
                }
                int32_t ob0 = ps.o;                         // first column of the (possibly partial) block that ends here
                for (int32_t q = j; q >= 0 && P.psteps[F.ps0 + q].ob_end == ps.ob_end; --q) ob0 = P.psteps[F.ps0 + q].o;
                const int32_t kb = e - ob0;
                if (last) {
                    GemmTask g = front_gemm(P, F, e, F.R - e, e, F.R - e, ob0, kb);
                    gg.add(P, g, gemm_flops(g));
                } else {
                    const int32_t e2 = P.psteps[F.ps0 + j + 1].ob_end;     // end of the next block's columns
                    GemmTask g = front_gemm(P, F, e, F.R - e, e, e2 - e, ob0, kb);            // column strip
                    gp.add(P, g, gemm_flops(g));
                    if (e2 < F.R) {
                        if (lu) {
                            GemmTask h = front_gemm(P, F, e, e2 - e, e2, F.R - e2, ob0, kb);  // row strip
                            gp.add(P, h, gemm_flops(h));
                        }
                        // The rest.  SPLIT: the part of it that the NEXT boundary's strips update again (the columns —
                        // LU: and rows — of the block after the next one) goes first, as a launch of its own that
                        // records the "early" event; the next strips wait for that event only and then run beside
                        // the remainder instead of between two bulk updates (0.3 ms of an idle trailing-update
                        // stream per outer block at the top of the tree).
                        const int32_t jn = j + 1 + P.ob_steps;                                // first step of the block after the next
                        const int32_t e3 = (P.split_rest && jn < F.nps) ? P.psteps[F.ps0 + jn].ob_end : F.R;
                        if (e3 < F.R) {
                            GemmTask ra = front_gemm(P, F, e2, F.R - e2, e2, e3 - e2, ob0, kb);
                            ga.add(P, ra, gemm_flops(ra));
                            if (lu) {
                                GemmTask rh = front_gemm(P, F, e2, e3 - e2, e3, F.R - e3, ob0, kb);
                                ga.add(P, rh, gemm_flops(rh));
                            }
                            GemmTask rb = front_gemm(P, F, e3, F.R - e3, e3, F.R - e3, ob0, kb);
                            gg.add(P, rb, gemm_flops(rb));
                        } else {
                            GemmTask r = front_gemm(P, F, e2, F.R - e2, e2, F.R - e2, ob0, kb);
                            (P.split_rest ? ga : gg).add(P, r, gemm_flops(r));                // nothing behind it: all of it is "early"
                        }
                    }
                }
            }
            // strips wait for the EARLY part of the previous block's rest (same target region: record bit 2 / wait bit 2;
            // without the split: for all of it); rests wait for this step's panels.  Every launch of the trailing-update
            // stream records the ordinary event too (the next level waits for the last of them).
            gp.emit(P, fb, lev, j, 0, boundary ? (P.split_rest ? 2 : 1) : 0, 1);
            ga.emit(P, fb, lev, j, 1, 1, 3);
            // the bulk update runs beside the next block's diagonal / panel chain: leave that chain some block slots
            gg.emit(P, fb, lev, j, 1, 1, 1, (j + 1 < maxnps) ? P.gemm_reserve : 0);
        }
    }

    // ---- solves on the frontal matrices: one step per panel step.  Small fronts: one block walks the whole
    // front.  Large fronts: a (diag, update) launch pair per panel step.
    if (P.solve_on_fronts) {
        // three classes: small fronts (one block walks the front), fronts with many panel steps (dataflow launch),
        // the rest (one launch per panel step); `stepf` = the last class
        std::vector<uint8_t> smallf(nf), flowf(nf), stepf(nf);
        for (int32_t f = 0; f < nf; ++f) {
            smallf[f] = (int64_t)P.fronts[f].R * P.fronts[f].W <= P.solve_small;
            flowf[f] = !smallf[f] && P.solve_flow && P.fronts[f].nps >= P.flow_min_steps;
            stepf[f] = !smallf[f] && !flowf[f];
        }
        LaunchBuilder sf(P, fwd_out);
        for (int32_t lev = 0; lev < P.nlevels; ++lev) {
            const std::vector<int32_t>& fr = bylevel[lev];
            int32_t maxnps = 0;
            sf.begin(K_FWD_GATHER, (int32_t)P.gathert.size(), lev, 0);
            bool anyflow = false;
            for (int32_t f : fr) { P.gathert.push_back(f); sf.add(1); if (stepf[f]) maxnps = std::max(maxnps, P.fronts[f].nps); anyflow |= flowf[f] != 0; }
            sf.end();
            sf.begin(K_PF_FRONT, (int32_t)P.gathert.size(), lev, 0);
            for (int32_t f : fr) if (smallf[f]) { P.gathert.push_back(f); int32_t mw = 0; for (int32_t q = 0; q < P.fronts[f].nps; ++q) mw = std::max(mw, P.psteps[P.fronts[f].ps0 + q].w); sf.add(1, 0, mw); }
            sf.end();
            if (anyflow) {
                // dataflow sweep: per big front, blocks owning the pivot rows of a few consecutive panel steps (in step
                // order: a block only ever waits for blocks listed before it), then blocks owning slabs of the rows below
                sf.begin(K_PF_FLOW, (int32_t)P.flowt.size(), lev, 0);
                sf.cur.ctr = P.nflowctr++;
                for (int32_t f : fr) if (flowf[f]) {
                    const Front& F = P.fronts[f];
                    for (int32_t j = 0; j < F.nps;) {
                        int32_t j1 = j, rows = 0;
                        while (j1 < F.nps && (j1 == j || rows + P.psteps[F.ps0 + j1].w <= FLOW_ROWS)) { rows += P.psteps[F.ps0 + j1].w; ++j1; }
                        int32_t mw = 0; for (int32_t q = j; q < j1; ++q) mw = std::max(mw, P.psteps[F.ps0 + q].w);
                        P.flowt.push_back(FlowTask{f, j, j1, 0, 0}); sf.add(1, 0, mw);
                        j = j1;
                    }
                    for (int32_t r = F.W; r < F.R; r += FLOW_ROWS) { P.flowt.push_back(FlowTask{f, 0, 0, r, std::min(F.R, r + FLOW_ROWS)}); sf.add(1, 0, 0); }
                }
                sf.end();
            }
            for (int32_t j = 0; j < maxnps; ++j) {
                if (j == 0) {                       // later diagonal blocks are solved inside the previous fused step
                    sf.begin(K_PF_DIAG, (int32_t)P.gathert.size(), lev, j);
                    for (int32_t f : fr) if (stepf[f] && P.fronts[f].nps > j) { P.gathert.push_back(P.fronts[f].ps0 + j); sf.add(1, 0, P.psteps[P.fronts[f].ps0 + j].w); }
                    sf.end();
                }
                sf.begin(K_PF_STEP, (int32_t)P.gathert.size(), lev, j);
                for (int32_t f : fr) if (stepf[f] && P.fronts[f].nps > j) {
                    const PStep& ps = P.psteps[P.fronts[f].ps0 + j];
                    int32_t below = ps.R - ps.o - ps.w;
                    int32_t wn = P.fronts[f].nps > j + 1 ? P.psteps[P.fronts[f].ps0 + j + 1].w : 0;
                    if (below > 0) { P.gathert.push_back(P.fronts[f].ps0 + j); sf.add(cdiv(below, SV_ROWS), 0, std::max(ps.w, wn)); }
                }
                sf.end();
            }
        }
        LaunchBuilder sb(P, bwd_out);
        for (int32_t lev = P.nlevels - 1; lev >= 0; --lev) {
            const std::vector<int32_t>& fr = bylevel[lev];
            int32_t maxnps = 0;
            sb.begin(K_BWD_GATHER, (int32_t)P.gathert.size(), lev, 0);
            bool anyflow = false;
            for (int32_t f : fr) {
                if (stepf[f]) maxnps = std::max(maxnps, P.fronts[f].nps);
                anyflow |= flowf[f] != 0;
                if (P.fronts[f].m > 0) { P.gathert.push_back(f); sb.add(cdiv(P.fronts[f].m, 256)); }
            }
            sb.end();
            sb.begin(K_PB_FRONT, (int32_t)P.gathert.size(), lev, 0);
            for (int32_t f : fr) if (smallf[f]) { P.gathert.push_back(f); int32_t mw = 0; for (int32_t q = 0; q < P.fronts[f].nps; ++q) mw = std::max(mw, P.psteps[P.fronts[f].ps0 + q].w); sb.add(1, 0, mw); }
            sb.end();
            if (anyflow) {
                sb.begin(K_PB_FLOW, (int32_t)P.flowt.size(), lev, 0);
                sb.cur.ctr = P.nflowctr++;
                for (int32_t f : fr) if (flowf[f]) {
                    const Front& F = P.fronts[f];
                    for (int32_t j1 = F.nps; j1 > 0;) {                       // last steps first: the order of the backward sweep
                        int32_t j = j1, rows = 0;
                        while (j > 0 && (j == j1 || rows + P.psteps[F.ps0 + j - 1].w <= FLOW_ROWS)) { rows += P.psteps[F.ps0 + j - 1].w; --j; }
                        int32_t mw = 0; for (int32_t q = j; q < j1; ++q) mw = std::max(mw, P.psteps[F.ps0 + q].w);
                        P.flowt.push_back(FlowTask{f, j, j1, 0, 0}); sb.add(1, 0, mw);
                        j1 = j;
                    }
                }
                sb.end();
            }
            for (int32_t j = maxnps - 1; j >= 0; --j) {
                sb.begin(K_PB_STEP, (int32_t)P.gathert.size(), lev, j);
                for (int32_t f : fr) if (stepf[f] && P.fronts[f].nps > j) {
                    const PStep& ps = P.psteps[P.fronts[f].ps0 + j];
                    int32_t below = ps.R - ps.o - ps.w;
                    P.gathert.push_back(P.fronts[f].ps0 + j); sb.add(std::max(1, cdiv(below, SV_ROWS)), 0, ps.w);
                }
                sb.end();
            }
        }
        return;
    }

    // ---- solves (on lnz / unz in the reference layout).  Small fronts: one block walks all chunks of the
    // front; large fronts: one launch pair per chunk step so that many blocks share the panel.
    auto front_entries = [&](const Front& F) { int64_t e = 0; for (int32_t t = 0; t < F.nch; ++t) { const Chunk& c = P.chunks[F.c0 + t]; e += (int64_t)c.jlen * c.nj; } return e; };
    std::vector<uint8_t> small(nf);
    for (int32_t f = 0; f < nf; ++f) small[f] = front_entries(P.fronts[f]) <= P.solve_small;
    LaunchBuilder sf(P, fwd_out);
    for (int32_t lev = 0; lev < P.nlevels; ++lev) {
        const std::vector<int32_t>& fr = bylevel[lev];
        int32_t maxnch = 0;
        sf.begin(K_FWD_GATHER, (int32_t)P.gathert.size(), lev, 0);
        for (int32_t f : fr) { P.gathert.push_back(f); sf.add(1); if (!small[f]) maxnch = std::max(maxnch, P.fronts[f].nch); }
        sf.end();
        sf.begin(K_FWD_FRONT, (int32_t)P.gathert.size(), lev, 0);
        for (int32_t f : fr) if (small[f]) { P.gathert.push_back(f); sf.add(1); }
        sf.end();
        for (int32_t t = 0; t < maxnch; ++t) {
            sf.begin(K_FWD_DIAG, (int32_t)P.gathert.size(), lev, t);
            for (int32_t f : fr) if (!small[f] && P.fronts[f].nch > t) { P.gathert.push_back(P.fronts[f].c0 + t); sf.add(1); }
            sf.end();
            sf.begin(K_FWD_UPDATE, (int32_t)P.gathert.size(), lev, t);
            for (int32_t f : fr) if (!small[f] && P.fronts[f].nch > t) {
                const Chunk& c = P.chunks[P.fronts[f].c0 + t];
                if (c.jlen > c.nj) { P.gathert.push_back(P.fronts[f].c0 + t); sf.add(cdiv(c.jlen - c.nj, UPD_ROWS)); }
            }
            sf.end();
        }
    }
    LaunchBuilder sb(P, bwd_out);
    for (int32_t lev = P.nlevels - 1; lev >= 0; --lev) {
        const std::vector<int32_t>& fr = bylevel[lev];
        int32_t maxnch = 0;
        sb.begin(K_BWD_GATHER, (int32_t)P.gathert.size(), lev, 0);
        for (int32_t f : fr) {
            if (!small[f]) maxnch = std::max(maxnch, P.fronts[f].nch);
            if (P.fronts[f].m > 0) { P.gathert.push_back(f); sb.add(cdiv(P.fronts[f].m, 256)); }
        }
        sb.end();
        sb.begin(K_BWD_FRONT, (int32_t)P.gathert.size(), lev, 0);
        for (int32_t f : fr) if (small[f]) { P.gathert.push_back(f); sb.add(1); }
        sb.end();
        for (int32_t t = maxnch - 1; t >= 0; --t) {
            sb.begin(K_BWD_UPDATE, (int32_t)P.gathert.size(), lev, t);
            for (int32_t f : fr) if (!small[f] && P.fronts[f].nch > t) {
                const Chunk& c = P.chunks[P.fronts[f].c0 + t];
                // LDL^T: the D^-1 scaling of the block's unknowns happens here, so every chunk gets a task
                if (c.jlen > c.nj || !lu) { P.gathert.push_back(P.fronts[f].c0 + t); sb.add(cdiv(c.nj, BWD_COLS)); }
            }
            sb.end();
            sb.begin(K_BWD_DIAG, (int32_t)P.gathert.size(), lev, t);
            for (int32_t f : fr) if (!small[f] && P.fronts[f].nch > t) { P.gathert.push_back(P.fronts[f].c0 + t); sb.add(1); }
            sb.end();
        }
    }
}


// ---------------------------------------------------------------------------------------
// Distributed top set: ownership of the front columns, storage held by this part, and the factor launch list.
inline void assign_ownership(Plan& P) {
    const int32_t nf = (int32_t)P.fronts.size();
    P.ownofs.assign(nf, -1); P.fown.clear();
    if (!P.dist_top) return;
    std::vector<int32_t> shift(nf, 0), cnt(P.nlevels, 0);
    for (int32_t f = 0; f < nf; ++f) if (P.owner[f] == -1) shift[f] = cnt[P.fronts[f].level]++;
    for (int32_t f = nf - 1; f >= 0; --f) {                    // parents before children
        if (P.owner[f] != -1) continue;
        const Front& F = P.fronts[f];
        P.ownofs[f] = (int32_t)P.fown.size();
        P.fown.resize(P.fown.size() + F.R, 0);
        int8_t* own = P.fown.data() + P.ownofs[f];
        for (int32_t j = 0; j < F.nps; ++j) {
            const PStep& ps = P.psteps[F.ps0 + j];
            const int8_t o = (int8_t)((j / P.ob_steps + shift[f]) % P.nparts);
            for (int32_t c = ps.o; c < ps.o + ps.w; ++c) own[c] = o;
        }
        if (F.m > 0) {
            const int8_t* pown = P.fown.data() + P.ownofs[F.parent];     // the parent of a top-set front is in the top set
            for (int32_t i = 0; i < F.m; ++i) own[F.W + i] = pown[P.rel[F.relofs + i]];
        }
    }
}

// Storage: a part of a multi-part plan keeps frontal matrices only for its own subtrees, the top set and the
// subtree roots whose update matrices it receives; everything else gets no arena space (fofs = -1).
inline void assign_storage(Plan& P) {
    const int32_t nf = (int32_t)P.fronts.size();
    P.held.assign(nf, 1);
    if (P.nparts <= 1) return;
    for (int32_t f = 0; f < nf; ++f) P.held[f] = (P.owner[f] == P.part || P.owner[f] == -1);
    for (int32_t f : P.xchg) P.held[f] = 1;
    int64_t fofs = 0;
    for (int32_t f = 0; f < nf; ++f) {
        Front& F = P.fronts[f];
        if (P.held[f]) { F.fofs = fofs; fofs += (int64_t)F.ld * F.R; fofs = (fofs + 1) & ~(int64_t)1; }
        else F.fofs = -1;
        for (int32_t t = 0; t < F.nch; ++t) P.chunks[F.c0 + t].fofs = F.fofs;
        for (int32_t j = 0; j < F.nps; ++j) P.psteps[F.ps0 + j].fofs = F.fofs;
    }
    P.arena = fofs;
}

inline void build_lists_dist(Plan& P, std::vector<Launch>& out) {
    const int32_t nf = (int32_t)P.fronts.size(), me = P.part;
    std::vector<std::vector<int32_t>> bylevel(P.nlevels);
    for (int32_t f = 0; f < nf; ++f) if (P.owner[f] == -1) bylevel[P.fronts[f].level].push_back(f);
    LaunchBuilder fb(P, out);
    auto begin = [&](int32_t kind, int32_t first, int32_t lev, int32_t step, int stream, int wait_mask) {
        fb.begin(kind, first, lev, step, stream, 0, 1); fb.cur.wait_mask = (uint8_t)wait_mask;
    };
    auto emit = [&](GemmBatch& g, int32_t lev, int32_t step, int stream, int wait_mask, int reserve = 0) {
        const size_t n0 = out.size();
        g.emit(P, fb, lev, step, stream, 0, 1, reserve);
        for (size_t i = n0; i < out.size(); ++i) out[i].wait_mask = (uint8_t)wait_mask;
    };
    // the one exchange: update-matrix column slabs of the subtree roots, from their owners to everybody
    begin(K_BCAST, (int32_t)P.bcasts.size(), 0, 0, 2, 1 | 2);
    for (int32_t f : P.xchg) {
        const Front& F = P.fronts[f];
        if (F.m <= 0) continue;
        P.bcasts.push_back(Bcast{F.fofs + (int64_t)F.W * F.ld, (int64_t)F.ld * F.m, P.owner[f], f});
        fb.add(1);
    }
    fb.end();
    for (int32_t lev = 0; lev < P.nlevels; ++lev) {
        const std::vector<int32_t>& fr = bylevel[lev];
        if (fr.empty()) continue;
        int32_t maxch = 0, maxnps = 0;
        for (int32_t f : fr) { maxch = std::max(maxch, P.fronts[f].nchild); maxnps = std::max(maxnps, P.fronts[f].nps); }
        // extend-add into the columns this part owns (k_assemble filters by DFront::ownofs)
        for (int32_t r = 0; r < std::min(maxch, ASM_ROUNDS); ++r) {
            begin(K_ASM, (int32_t)P.asmt.size(), lev, r, 0, 2 | 4);
            for (int32_t f : fr) {
                const Front& F = P.fronts[f];
                if (F.nchild <= r) continue;
                int32_t c = P.childlist[F.child0 + r];
                int64_t mc = P.fronts[c].m;
                P.asmt.push_back(AsmTask{c, f});
                fb.add((int32_t)(cdiv(mc, ASM_TPB) * cdiv(mc, ASM_COLS)));
            }
            fb.end();
        }
        if (maxch > ASM_ROUNDS) {
            begin(K_ASM_TAIL, (int32_t)P.asmt.size(), lev, ASM_ROUNDS, 0, 2 | 4);
            for (int32_t f : fr) if (P.fronts[f].nchild > ASM_ROUNDS) { P.asmt.push_back(AsmTask{-1, f}); fb.add(1); }
            fb.end();
        }
        for (int32_t j = 0; j < maxnps; ++j) {
            auto mine = [&](int32_t f, int32_t step) {
                const Front& F = P.fronts[f];
                return P.fown[P.ownofs[f] + P.psteps[F.ps0 + step].o] == me;
            };
            // ---- the owner's chain: left-looking in-block update, diagonal block, panel
            if ((j % P.ob_steps) != 0) {
                GemmBatch gl;
                for (int32_t f : fr) if (P.fronts[f].nps > j && mine(f, j)) {
                    const Front& F = P.fronts[f];
                    const PStep& ps = P.psteps[F.ps0 + j];
                    const int32_t ob0 = P.psteps[F.ps0 + j - (j % P.ob_steps)].o;
                    GemmTask g = front_gemm(P, F, ps.o, F.R - ps.o, ps.o, ps.w, ob0, ps.o - ob0);
                    gl.add(P, g, gemm_flops(g));
                }
                emit(gl, lev, j, 0, 0);
            }
            begin(K_DIAG, (int32_t)P.pslist.size(), lev, j, 0, j == 0 ? 2 : 0);
            for (int32_t f : fr) if (P.fronts[f].nps > j && mine(f, j)) { P.pslist.push_back(P.fronts[f].ps0 + j); fb.add(1, 0, P.psteps[P.fronts[f].ps0 + j].w); }
            fb.end();
            begin(K_PANEL, (int32_t)P.pslist.size(), lev, j, 0, 0);
            for (int32_t f : fr) if (P.fronts[f].nps > j && mine(f, j)) {
                const PStep& ps = P.psteps[P.fronts[f].ps0 + j];
                int32_t below = ps.R - ps.o - ps.w;
                if (below <= 0) continue;
                P.pslist.push_back(P.fronts[f].ps0 + j);
                fb.add(cdiv(below, PANEL_ROWS), 0, ps.w);
            }
            fb.end();
            // ---- the factored columns of THIS step travel at once (in place, from the owner): the broadcast of a block's
            // columns overlaps the owner's chain over the block's remaining steps instead of following it.  A step's
            // column slab is final after its panel kernel (stream 0); on the receivers only the U rebuild of earlier
            // blocks (stream 0) touches these columns — the trailing updates (stream 1) never do.
            begin(K_BCAST, (int32_t)P.bcasts.size(), lev, j, 2, 1);
            for (int32_t f : fr) if (P.fronts[f].nps > j) {
                const Front& F = P.fronts[f];
                const PStep& ps = P.psteps[F.ps0 + j];
                P.bcasts.push_back(Bcast{F.fofs + (int64_t)ps.o * F.ld, (int64_t)F.ld * ps.w, P.fown[P.ownofs[f] + ps.o], f});
                fb.add(1);
            }
            fb.end();
            // ---- end of an outer block: rebuild U = D L^T from the received columns, delayed updates
            const bool boundary = ((j + 1) % P.ob_steps) == 0;
            struct End { int32_t f, ob0, e, e2; bool last; };
            std::vector<End> ends;
            for (int32_t f : fr) if (P.fronts[f].nps > j) {
                const Front& F = P.fronts[f];
                const PStep& ps = P.psteps[F.ps0 + j];
                const bool last = (F.nps == j + 1);
                if (!boundary && !last) continue;
                int32_t ob0 = ps.o;
                for (int32_t q = j; q >= 0 && P.psteps[F.ps0 + q].ob_end == ps.ob_end; --q) ob0 = P.psteps[F.ps0 + q].o;
                const int32_t e = ps.o + ps.w;
                ends.push_back(End{f, ob0, e, last ? e : P.psteps[F.ps0 + j + 1].ob_end, last});
            }
            if (ends.empty()) continue;
            bool all_mine = true;
            for (const End& E : ends) all_mine = all_mine && P.fown[P.ownofs[E.f] + E.ob0] == me;
            begin(K_FILLU, (int32_t)P.fillt.size(), lev, j, 0, all_mine ? 0 : 4);      // the owner already holds the panel
            for (const End& E : ends) {
                const Front& F = P.fronts[E.f];
                if (E.e >= F.R) continue;
                P.fillt.push_back(FillTask{F.fofs, F.ld, F.R, E.ob0, E.e});
                fb.add(cdiv(F.R - E.e, 32) * cdiv(E.e - E.ob0, 32));
            }
            fb.end();
            GemmBatch gp, gg;
            for (const End& E : ends) {
                const Front& F = P.fronts[E.f];
                if (E.e >= F.R) continue;
                const int8_t* own = P.fown.data() + P.ownofs[E.f];
                const int32_t kb = E.e - E.ob0;
                if (!E.last && own[E.e] == me) {             // the strip the next block's owner needs first
                    GemmTask g = front_gemm(P, F, E.e, F.R - E.e, E.e, E.e2 - E.e, E.ob0, kb);
                    gp.add(P, g, gemm_flops(g));
                }
                for (int32_t c0 = E.e2; c0 < F.R;) {         // the rest, by runs of columns this part owns
                    if (own[c0] != me) { ++c0; continue; }
                    int32_t c1 = c0;
                    while (c1 < F.R && own[c1] == me) ++c1;
                    GemmTask r = front_gemm(P, F, c0, F.R - c0, c0, c1 - c0, E.ob0, kb);
                    gg.add(P, r, gemm_flops(r));
                    c0 = c1;
                }
            }
            emit(gp, lev, j, 0, 2);                           // after the previous block's rest (same target region)
            emit(gg, lev, j, 1, 1, (j + 1 < maxnps) ? P.gemm_reserve : 0);   // after this block's U is in place; beside the next chain
        }
    }
}

// Elimination-subtree partition: split the front tree from the root until there are at
// least `nparts` subtrees, biggest first; the fronts split off form the top set; subtrees are dealt to
// parts by decreasing work (LPT).  Subtrees are contiguous front ranges because the reference post-orders
// the elimination tree (SpkETree.jl:106-139).
inline void partition(Plan& P) {
    const int32_t nf = (int32_t)P.fronts.size();
    P.owner.assign(nf, P.nparts > 1 ? -2 : 0);
    P.xchg.clear(); P.ranges.clear();
    if (P.nparts <= 1) return;
    std::vector<double> work(nf, 0.0);
    std::vector<int32_t> first(nf);                            // first front of the subtree rooted at f
    for (int32_t f = 0; f < nf; ++f) {
        const Front& F = P.fronts[f];
        double W = F.W, m = F.m;
        work[f] += W * W * W / 3.0 + W * W * m + W * m * m + 1.0;
        first[f] = f;
    }
    for (int32_t f = 0; f < nf; ++f) {                          // children precede parents
        for (int32_t q = 0; q < P.fronts[f].nchild; ++q) {
            int32_t ch = P.childlist[P.fronts[f].child0 + q];
            work[f] += work[ch]; first[f] = std::min(first[f], first[ch]);
        }
    }
    std::vector<int32_t> roots;
    for (int32_t f = 0; f < nf; ++f) if (P.fronts[f].parent < 0) roots.push_back(f);
    // Estimated time of a configuration = (replicated) top-set time + the slowest part under LPT.  A front costs
    // its flops at the measured update rate, but never less than its chain of dependent panel steps (measured
    // ~75 us per step): the top of the tree is latency-, not flop-bound, so splitting further stops paying even
    // when flops still balance better (96^3 on 8 GPUs: 8 subtrees + 6 replicated top fronts ran slower than
    // 4 subtrees + 2).  Keep splitting the largest splittable subtree (it joins the top set) and remember the
    // best configuration seen; parts may stay without a subtree if that is faster.
    // `own` below = the LDL^T flops of a front (LU: twice that); measured update rate ~24 TFLOP/s
    const double rate = (P.lu ? 0.5 : 1.0) * 24e12;
    const double tstep = 75e-6;
    // distributed top set: + one broadcast, one U rebuild and one strip update per outer block on the chain
    const double tstep_top = P.dist_top ? 150e-6 : tstep;
    auto own = [&](int32_t f) { double W = P.fronts[f].W, m = P.fronts[f].m; return W * W * W / 3.0 + W * W * m + W * m * m + 1.0; };
    std::vector<double> chain(nf, 0.0);                          // dependent panel steps below and including f, in seconds
    for (int32_t f = 0; f < nf; ++f) {
        double below = 0.0;
        for (int32_t q = 0; q < P.fronts[f].nchild; ++q) below = std::max(below, chain[P.childlist[P.fronts[f].child0 + q]]);
        chain[f] = below + P.fronts[f].nps * tstep;
    }
    auto sub_time = [&](int32_t r) { return std::max(work[r] / rate, chain[r]); };
    auto lpt_max = [&](std::vector<int32_t> rs) {
        std::sort(rs.begin(), rs.end(), [&](int32_t a, int32_t b) { return sub_time(a) != sub_time(b) ? sub_time(a) > sub_time(b) : a < b; });
        std::vector<double> ld(P.nparts, 0.0);
        for (int32_t r : rs) { int32_t t = 0; for (int32_t q = 1; q < P.nparts; ++q) if (ld[q] < ld[t]) t = q; ld[t] += sub_time(r); }
        return *std::max_element(ld.begin(), ld.end());
    };
    auto top_time = [&](const std::vector<int32_t>& tp) {       // fronts of one level share their launches
        std::vector<double> fl(P.nlevels, 0.0), st(P.nlevels, 0.0);
        for (int32_t f : tp) { int32_t l = P.fronts[f].level; fl[l] += own(f) / rate; st[l] = std::max(st[l], P.fronts[f].nps * tstep_top); }
        double t = 0.0;
        // distributed top set: the trailing updates of a level are shared by all parts, the chain of panel steps is not
        for (int32_t l = 0; l < P.nlevels; ++l) t += std::max(P.dist_top ? fl[l] / P.nparts : fl[l], st[l]);
        return t;
    };
    std::vector<int32_t> top, best_roots = roots, best_top;
    double best_cost = 1e300;
    for (int iter = 0; iter < 64; ++iter) {
        if (P.force_splits >= 0 && iter == P.force_splits && roots.size() >= 2) { best_roots = roots; best_top = top; best_cost = 0.0; break; }   // SPK_TOP_SPLITS
        if ((roots.size() >= 2 && (int32_t)roots.size() <= P.max_subtrees) || P.nparts == 1) {
            double cost = top_time(top) + lpt_max(roots);
            if (getenv("SPK_PARTITION_DEBUG")) fprintf(stderr, "[partition] parts=%d splits=%d subtrees=%zu top=%.1f ms subtrees(LPT max)=%.1f ms total=%.1f ms\n", P.nparts, iter, roots.size(), top_time(top) * 1e3, lpt_max(roots) * 1e3, cost * 1e3);
            if (cost < 0.98 * best_cost) { best_cost = cost; best_roots = roots; best_top = top; }   // ties go to the shallower cut
        }
        int32_t pick = -1;
        for (size_t i = 0; i < roots.size(); ++i)
            if (P.fronts[roots[i]].nchild > 0 && (pick < 0 || work[roots[i]] > work[roots[pick]])) pick = (int32_t)i;
        if (pick < 0) break;
        int32_t f = roots[pick];
        roots.erase(roots.begin() + pick);
        top.push_back(f);
        for (int32_t q = 0; q < P.fronts[f].nchild; ++q) roots.push_back(P.childlist[P.fronts[f].child0 + q]);
        if ((int32_t)roots.size() >= 4 * P.nparts && top_time(top) > best_cost) break;   // cannot improve any more
    }
    if (best_cost >= 1e300) { best_roots = roots; best_top = top; }        // tree too thin to split
    roots = best_roots;
    for (int32_t f : best_top) P.owner[f] = -1;
    std::sort(roots.begin(), roots.end(), [&](int32_t a, int32_t b) { return work[a] != work[b] ? work[a] > work[b] : a < b; });
    std::vector<double> load(P.nparts, 0.0);
    for (int32_t r : roots) {
        int32_t tgt = 0;
        for (int32_t q = 1; q < P.nparts; ++q) if (load[q] < load[tgt]) tgt = q;
        load[tgt] += work[r];
        for (int32_t f = first[r]; f <= r; ++f) P.owner[f] = tgt;
        if (P.fronts[r].parent >= 0) P.xchg.push_back(r);
        Plan::Range g{};
        g.owner = tgt; g.f0 = first[r]; g.f1 = r + 1;
        const Chunk& c0 = P.chunks[P.fronts[first[r]].c0];
        const Front& Fr = P.fronts[r];
        const Chunk& c1 = P.chunks[Fr.c0 + Fr.nch - 1];
        g.lnz0 = c0.lofs; g.lnz1 = c1.lofs + (int64_t)c1.jlen * c1.nj;
        g.unz0 = c0.uofs; g.unz1 = c1.uofs + (P.lu ? (int64_t)(c1.jlen - c1.nj) * c1.nj : 0);
        g.col0 = c0.fj; g.col1 = c1.fj + c1.nj;
        P.ranges.push_back(g);
    }
    std::sort(P.xchg.begin(), P.xchg.end());
    std::sort(P.ranges.begin(), P.ranges.end(), [](const Plan::Range& a, const Plan::Range& b) { return a.f0 < b.f0; });
}

inline void build_schedule(Plan& P) {
    const int32_t nf = (int32_t)P.fronts.size();
    P.solvet.resize(P.chunks.size());
    for (size_t s = 0; s < P.chunks.size(); ++s) {
        const Chunk& c = P.chunks[s]; const Front& F = P.fronts[c.front];
        SolveTask t{}; t.lofs = c.lofs; t.uofs = c.uofs; t.col0 = c.fj; t.wofs = F.wofs; t.posofs = c.posofs;
        t.ld = c.jlen; t.ldu = c.jlen - c.nj; t.nj = c.nj; t.m = c.jlen - c.nj; t.o = c.o; t.front = c.front;
        P.solvet[s] = t;
    }
    P.pblen = 0;
    for (int32_t f = 0; f < nf; ++f) {
        Front& F = P.fronts[f];
        F.pbofs = P.pblen;
        if ((int64_t)F.R * F.W > P.solve_small) P.pblen += (int64_t)cdiv(F.R, SV_ROWS) * P.maxpw;
    }
    P.dist_top = P.nparts > 1 && !P.lu && P.dist_top_env != 0;
    partition(P);
    std::vector<uint8_t> sel(nf, 1);
    if (P.nparts <= 1) {
        build_lists(P, sel, P.factor_launches, P.fwd_launches, P.bwd_launches);
        P.factor_pipe.clear(); P.factor_ptop.clear();
        if (P.pipes > 1 && P.lookahead && nf > 1) {
            P.nparts = P.pipes;
            partition(P);
            const std::vector<int32_t> vowner = P.owner;
            P.nparts = 1; P.owner.assign(nf, 0); P.xchg.clear(); P.ranges.clear();
            bool any_top = false;
            for (int32_t f = 0; f < nf; ++f) any_top |= vowner[f] == -1;
            if (any_top) {
                std::vector<Launch> unused_fwd, unused_bwd;
                const bool sof = P.solve_on_fronts;
                P.factor_pipe.resize(P.pipes);
                for (int32_t v = 0; v < P.pipes; ++v) {
                    for (int32_t f = 0; f < nf; ++f) sel[f] = vowner[f] == v;
                    build_lists(P, sel, P.factor_pipe[v], unused_fwd, unused_bwd);
                }
                for (int32_t f = 0; f < nf; ++f) sel[f] = vowner[f] == -1;
                build_lists(P, sel, P.factor_ptop, unused_fwd, unused_bwd);
                P.solve_on_fronts = sof;
            }
        }
        return;
    }
    assign_ownership(P);
    assign_storage(P);
    for (int32_t f = 0; f < nf; ++f) sel[f] = P.owner[f] == P.part;
    build_lists(P, sel, P.factor_local, P.fwd_local, P.bwd_local);
    for (int32_t f = 0; f < nf; ++f) sel[f] = P.owner[f] == -1;
    build_lists(P, sel, P.factor_top, P.fwd_top, P.bwd_top);
    if (P.dist_top) { P.factor_top.clear(); build_lists_dist(P, P.factor_top); }
    else {
        // replicated top set: the exchange = the whole subtree-root fronts, then the ordinary two-stream lists
        std::vector<Launch> top;
        LaunchBuilder fb(P, top);
        fb.begin(K_BCAST, (int32_t)P.bcasts.size(), 0, 0, 2, 0, 1); fb.cur.wait_mask = 1 | 2;
        for (int32_t f : P.xchg) { const Front& F = P.fronts[f]; P.bcasts.push_back(Bcast{F.fofs, (int64_t)F.ld * F.R, P.owner[f], f}); fb.add(1); }
        fb.end();
        top.insert(top.end(), P.factor_top.begin(), P.factor_top.end());
        P.factor_top.swap(top);
    }
}

} // namespace spk
