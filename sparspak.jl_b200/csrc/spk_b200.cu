// spk_b200.cu — C ABI (include/spk_b200.h) and plan runtime of the B200 numeric engine.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include "spk_b200.h"
#include "plan.hpp"
#include "kernels.cuh"
#include "gemm_dmma.cuh"

using namespace spk;

static thread_local std::string g_err;
static void set_err(const std::string& s) { g_err = s; }

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            set_err(std::string(#call) + ": " + cudaGetErrorString(e_));                      \
            return -100 - (int64_t)e_;                                                        \
        }                                                                                     \
    } while (0)

template <class T, class A>
static cudaError_t upload(T** d, const std::vector<T, A>& h) {
    *d = nullptr;
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void**)d, bytes);
    if (e != cudaSuccess) return e;
    if (!h.empty()) e = cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

struct spk_plan {
    Plan P;
    int device = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;   // panel stream (high priority), trailing-update stream
    cudaEvent_t evs0 = nullptr, evs1 = nullptr;          // last record of each stream
    cudaStream_t stream_c = nullptr; cudaEvent_t evsc = nullptr;   // communication stream of a multi-part plan (NCCL broadcasts)
    ncclComm_t comm = nullptr; bool comm_owned = false;  // communicator over the parts (spk_plan_comm_init / spk_multi_create)
    int8_t* d_fown = nullptr; FillTask* d_fillt = nullptr;
    GemmTile* d_tiles = nullptr; int32_t* d_tilectr = nullptr; int num_sms = 148;
    FlowTask* d_flowt = nullptr; int32_t* d_flow = nullptr; int64_t flow_ints = 0;   // dataflow solve: tasks; ticket counters (zeroed per solve)
    double *d_tinvf = nullptr, *d_tinvb = nullptr; bool use_inv = true, inv_ready = false;   // inverted diagonal blocks (SPK_SOLVE_INV=0: off)
    int32_t *d_invlist = nullptr; int32_t inv_nsmall = 0, inv_nlarge = 0, inv_maxw_small = 32;
    // SPK_SOLVE_INV bits: 1 = step / flow kernels (big fronts), 2 = one-block-per-front kernels.  Default: LU 3, LDL^T 0
    // (measured at 80^3 LU: 11.1 -> 6.1 ms, the in-block solve with its interchanges is the expensive part of a step;
    //  at 96^3 LDL^T: 13.9 -> 17.1 ms, the register / shuffle sweep over a unit triangle is already cheap).
    int inv_mask = -1;
    double* d_box = nullptr;            // dataflow solve: x mailboxes, [forward | backward] x (rhs of a batch) x n, sentinel-filled per solve
    double ms_xchg = 0;
    // tree pipelines (Plan::pipes): stream / event pair per pipeline; pair 0 = (stream, stream2, evs0, evs1)
    cudaStream_t pst[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    cudaEvent_t pev[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    cudaEvent_t pev2[4] = {nullptr, nullptr, nullptr, nullptr};      // "early" event of a pair's trailing-update stream (split rest updates)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evg0 = nullptr, evg1 = nullptr;
    // device state
    double *d_pb = nullptr, *d_F = nullptr, *d_lnz = nullptr, *d_unz = nullptr, *d_w = nullptr, *d_rhs = nullptr, *d_tmp = nullptr;
    int32_t *d_ipiv = nullptr, *d_iflag = nullptr, *d_counters = nullptr;
    DFront* d_fronts = nullptr; DChunk* d_chunks = nullptr; PStep* d_psteps = nullptr;
    int32_t *d_subw = nullptr, *d_childlist = nullptr, *d_rel = nullptr, *d_pos = nullptr, *d_blkpfx = nullptr,
            *d_gathert = nullptr, *d_pslist = nullptr, *d_chunkpfx = nullptr;
    int64_t *d_dest = nullptr, *d_rperm = nullptr, *d_rinvp = nullptr;
    double* d_nzval = nullptr; int64_t nzcap = 0, nz_last = 0, map_nnz = -1;   // map_nnz: the nnz the destination map was built for
    AsmTask* d_asmt = nullptr; GemmTask* d_gemmt = nullptr; SolveTask* d_solvet = nullptr;
    int32_t chunk_blocks = 0;
    // CSR copy of A in the original ordering + work vectors (spk_plan_set_matrix / residual / refine)
    int64_t* d_arp = nullptr; int32_t* d_aci = nullptr; double* d_av = nullptr; int64_t annz = -1;
    double *d_rb = nullptr, *d_rx = nullptr, *d_rr = nullptr, *d_rpart = nullptr; int64_t rcap = 0;
    int32_t *d_stlist = nullptr, *d_stpfx = nullptr;    // chunk ids / block prefixes by level (overlapped factor write-back)
    std::vector<int32_t> st_list0[2], st_pfx0[2], st_count[2], st_blocks[2];   // [0] phase-0 / all fronts, [1] top set
    cudaStream_t stream3 = nullptr;                     // write-back stream (lowest priority)
    cudaEvent_t ev3a = nullptr, ev3b = nullptr, ev3c = nullptr;
    bool store_overlap = true;                          // SPK_STORE_OVERLAP=0: one write-back kernel after the factorisation
    bool solve_graphs = true;           // SPK_SOLVE_GRAPH=0 disables CUDA-graph replay of the solve sweeps
    cudaGraphExec_t sg_exec = nullptr; double* sg_rhs = nullptr; double* sg_w = nullptr;
    int64_t sg_nrhs = 0, sg_ld = 0, sg_launches = 0; int32_t sg_which = -1;
    int diag_tg = 1;                    // SPK_DIAG_TG: 1 = row kernel (4 threads per row), 8 / 16 = thread grid of the cyclic register kernel
    bool pdl = true;                    // SPK_PDL=0: launch the solve steps without programmatic dependent launch
    bool use_pipes = true;              // SPK_PIPES=1 builds no pipelines; this switches them off at run time (tests)
    bool pdl_gemm = false;              // SPK_PDL_GEMM=1: also the DMMA kernels (measured slower: early blocks hold SM resources)
    bool pdl_factor = false;            // SPK_PDL_FACTOR=1: same for the diagonal / panel kernels of the factorisation (measured: no gain)
    int panel_reg_minw = 32;            // SPK_PANEL_REG_MINW
    bool panel_smem_only = false;       // SPK_PANEL_SMEM=1: always use the shared-memory panel kernel
    bool diag_smem_only = false;        // SPK_DIAG_SMEM=1: always use the shared-memory diagonal kernel
    int dmma_variant = 4, dmma_variant64 = 8;   // SPK_DMMA_VARIANT, SPK_DMMA_VARIANT64 (see gemm_dmma.cuh)
    bool dmma_persist = false; int dmma_flags = 1;     // SPK_DMMA_PERSIST, SPK_DMMA_CA (bit 0), SPK_DMMA_STATIC (bit 1), SPK_DMMA_DEPHASE (us, bits 8..)
    bool values_in_fronts = false;      // inmatrix scattered straight into the fronts
    int64_t w_nrhs = 0, rhs_cap = 0;
    size_t dev_bytes = 0;
    int diag_smem_nj = 0; size_t diag_smem_bytes = 0, panel_smem = 0;
    bool have_perm = false, factored = false;
    bool phase0_async = false;          // phase 0 returns without synchronising (the top set is enqueued right behind it)
    bool ev0_armed = false;             // spk_plan_reassemble recorded ev0: the next factorisation's time starts there
    // stats
    int64_t launches_factor = 0, launches_solve = 0;
    double ms_factor = 0, ms_solve = 0, gemm_flops = 0, gemm_ms = 0, ms_phase0 = 0, ms_phase1 = 0;
    std::vector<float> launch_ms;       // optional per-launch timing (profiling mode)
    double kind_ms[16] = {0}; int64_t kind_n[16] = {0};
    bool profile = false;
};

// Launch with (pdl) or without the programmatic-stream-serialization attribute (see pdl_wait in kernels.cuh).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static DevCtx make_ctx(spk_plan* p) {
    DevCtx c{};
    c.F = p->d_F; c.lnz = p->d_lnz; c.unz = p->d_unz; c.w = p->d_w; c.pb = p->d_pb; c.pblen = p->P.pblen;
    c.ipiv = p->d_ipiv; c.iflag = p->d_iflag;
    c.fronts = p->d_fronts; c.chunks = p->d_chunks; c.psteps = p->d_psteps; c.subw = p->d_subw;
    c.childlist = p->d_childlist; c.rel = p->d_rel; c.pos = p->d_pos;
    c.solvet = p->d_solvet; c.wlen = p->P.wlen; c.lu = p->P.lu ? 1 : 0;
    c.me = p->P.part; c.fown = p->d_fown;
    c.tinvf = p->inv_ready ? p->d_tinvf : nullptr; c.tinvb = p->inv_ready ? p->d_tinvb : nullptr;
    return c;
}

// Dynamic shared-memory limits are a property of (function, device), not of a plan: every kernel that may need
// more than 48 KB gets the device's opt-in maximum (minus its static shared memory) ONCE per device, so plans of
// different panel widths can live side by side (a per-plan value would lower the limit under an older plan).
template <class K>
static cudaError_t set_max_dyn_smem(K kern, int optin) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}
static int64_t init_kernel_attributes(int device) {
    static std::mutex mu;
    static bool done[64] = {false};
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0 || device >= 64) { set_err("device index out of range"); return -100; }
    if (done[device]) return 0;
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
#define SPK_SMEM(k) CK(set_max_dyn_smem(k, optin))
    SPK_SMEM(k_diag<true>); SPK_SMEM(k_diag<false>); SPK_SMEM(k_panel<true>); SPK_SMEM(k_panel<false>);
    SPK_SMEM((k_pf_front<true, 1, 256>)); SPK_SMEM((k_pf_front<true, 1, 64>)); SPK_SMEM((k_pf_front<false, 1, 256>)); SPK_SMEM((k_pf_front<false, 1, 64>));
    SPK_SMEM((k_pb_front<true, 1, 256>)); SPK_SMEM((k_pb_front<true, 1, 64>)); SPK_SMEM((k_pb_front<false, 1, 256>)); SPK_SMEM((k_pb_front<false, 1, 64>));
    SPK_SMEM((k_pf_front<true, SOLVE_NR, 256>)); SPK_SMEM((k_pf_front<true, SOLVE_NR, 64>)); SPK_SMEM((k_pf_front<false, SOLVE_NR, 256>)); SPK_SMEM((k_pf_front<false, SOLVE_NR, 64>));
    SPK_SMEM((k_pb_front<true, SOLVE_NR, 256>)); SPK_SMEM((k_pb_front<true, SOLVE_NR, 64>)); SPK_SMEM((k_pb_front<false, SOLVE_NR, 256>)); SPK_SMEM((k_pb_front<false, SOLVE_NR, 64>));
    SPK_SMEM((k_pf_step<true, 1>)); SPK_SMEM((k_pf_step<false, 1>)); SPK_SMEM((k_pb_step<true, 1>)); SPK_SMEM((k_pb_step<false, 1>));
    SPK_SMEM((k_pf_step<true, SOLVE_NR>)); SPK_SMEM((k_pf_step<false, SOLVE_NR>)); SPK_SMEM((k_pb_step<true, SOLVE_NR>)); SPK_SMEM((k_pb_step<false, SOLVE_NR>));
    SPK_SMEM(k_pf_diag<true>); SPK_SMEM(k_pf_diag<false>); SPK_SMEM(k_pb_diag<true>); SPK_SMEM(k_pb_diag<false>);
    SPK_SMEM(k_pb_update<true>); SPK_SMEM(k_pb_update<false>);
    SPK_SMEM(k_diag_inverse<true>); SPK_SMEM(k_diag_inverse<false>);
    SPK_SMEM((k_pf_flow<true, 1, 1>)); SPK_SMEM((k_pf_flow<true, 1, 2>)); SPK_SMEM((k_pf_flow<true, SOLVE_NR, 1>)); SPK_SMEM((k_pf_flow<true, SOLVE_NR, 2>));
    SPK_SMEM((k_pf_flow<false, 1, 1>)); SPK_SMEM((k_pf_flow<false, 1, 2>)); SPK_SMEM((k_pf_flow<false, SOLVE_NR, 1>)); SPK_SMEM((k_pf_flow<false, SOLVE_NR, 2>));
    SPK_SMEM((k_pb_flow<true, 1, 1>)); SPK_SMEM((k_pb_flow<true, 1, 2>)); SPK_SMEM((k_pb_flow<true, SOLVE_NR, 1>)); SPK_SMEM((k_pb_flow<true, SOLVE_NR, 2>));
    SPK_SMEM((k_pb_flow<false, 1, 1>)); SPK_SMEM((k_pb_flow<false, 1, 2>)); SPK_SMEM((k_pb_flow<false, SOLVE_NR, 1>)); SPK_SMEM((k_pb_flow<false, SOLVE_NR, 2>));
#undef SPK_SMEM
    if (const char* e = getenv("SPK_SOLVE_DBG")) { int v = atoi(e); CK(cudaMemcpyToSymbol(g_solve_dbg, &v, sizeof(int))); }
    if (const char* e = getenv("SPK_FLOW_BACKOFF")) { int ns = atoi(e); CK(cudaMemcpyToSymbol(g_flow_backoff, &ns, sizeof(int))); }
    CK(gemm_dmma_init());
    done[device] = true;
    return 0;
}

// ---- NCCL, resolved at run time (the library has no link-time dependency on it; a process that already
// loaded libnccl.so.2 — e.g. through torch.distributed — shares that copy) -------------------------------
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi* nccl_api() {
    static NcclApi api; static std::once_flag once; static bool ok = false;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { api.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (api.h) break; }
        if (!api.h) return;
        auto sym = [&](const char* n) { return dlsym(api.h, n); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.Broadcast && api.GroupStart && api.GroupEnd && api.CommDestroy && api.GetErrorString;
    });
    return ok ? &api : nullptr;
}
#define NK(call)                                                                              \
    do {                                                                                      \
        ncclResult_t r_ = (call);                                                             \
        if (r_ != ncclSuccess) {                                                              \
            set_err(std::string(#call) + ": " + nccl_api()->GetErrorString(r_));              \
            return -300 - (int64_t)r_;                                                        \
        }                                                                                     \
    } while (0)

extern "C" {

SPK_API const char* spk_last_error(void) { return g_err.c_str(); }
SPK_API const char* spk_version(void) { return "sparspak.jl_b200 0.1.0 (sm_100a)"; }
SPK_API int32_t spk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

SPK_API void spk_plan_destroy(spk_plan* p) {
    if (!p) return;
    if (p->device >= 0) {
        cudaSetDevice(p->device);
        void* ptrs[] = {p->d_counters, p->d_pb, p->d_F, p->d_lnz, p->d_unz, p->d_w, p->d_rhs, p->d_tmp, p->d_ipiv, p->d_iflag, p->d_fronts,
                        p->d_chunks, p->d_psteps, p->d_subw, p->d_childlist, p->d_rel, p->d_pos, p->d_blkpfx,
                        p->d_gathert, p->d_pslist, p->d_chunkpfx, p->d_dest, p->d_rperm, p->d_rinvp, p->d_nzval,
                        p->d_asmt, p->d_gemmt, p->d_solvet};
        for (void* q : ptrs) if (q) cudaFree(q);
        if (p->ev0) cudaEventDestroy(p->ev0);
        if (p->ev1) cudaEventDestroy(p->ev1);
        if (p->evg0) cudaEventDestroy(p->evg0);
        if (p->evg1) cudaEventDestroy(p->evg1);
        if (p->sg_exec) cudaGraphExecDestroy(p->sg_exec);
        if (p->evs0) cudaEventDestroy(p->evs0);
        if (p->evs1) cudaEventDestroy(p->evs1);
        if (p->comm && p->comm_owned && nccl_api()) nccl_api()->CommDestroy(p->comm);
        if (p->d_flowt) cudaFree(p->d_flowt);
        if (p->d_flow) cudaFree(p->d_flow);
        if (p->d_box) cudaFree(p->d_box);
        if (p->d_tinvf) cudaFree(p->d_tinvf);
        if (p->d_tinvb) cudaFree(p->d_tinvb);
        if (p->d_invlist) cudaFree(p->d_invlist);
        if (p->d_tiles) cudaFree(p->d_tiles);
        if (p->d_tilectr) cudaFree(p->d_tilectr);
        if (p->d_fown) cudaFree(p->d_fown);
        if (p->d_fillt) cudaFree(p->d_fillt);
        if (p->evsc) cudaEventDestroy(p->evsc);
        if (p->stream_c) cudaStreamDestroy(p->stream_c);
        if (p->stream2) cudaStreamDestroy(p->stream2);
        if (p->stream3) cudaStreamDestroy(p->stream3);
        if (p->ev3a) cudaEventDestroy(p->ev3a);
        if (p->ev3b) cudaEventDestroy(p->ev3b);
        if (p->ev3c) cudaEventDestroy(p->ev3c);
        for (void* q : {(void*)p->d_arp, (void*)p->d_aci, (void*)p->d_av, (void*)p->d_rb, (void*)p->d_rx, (void*)p->d_rr, (void*)p->d_rpart}) if (q) cudaFree(q);
        if (p->d_stlist) cudaFree(p->d_stlist);
        if (p->d_stpfx) cudaFree(p->d_stpfx);
        for (int v = 1; v < 4; ++v) for (int q = 0; q < 2; ++q) { if (p->pst[v][q]) cudaStreamDestroy(p->pst[v][q]); if (p->pev[v][q]) cudaEventDestroy(p->pev[v][q]); }
        for (int v = 0; v < 4; ++v) if (p->pev2[v]) cudaEventDestroy(p->pev2[v]);
        if (p->stream) cudaStreamDestroy(p->stream);
    }
    delete p;
}

static int64_t plan_upload(spk_plan* p) {
    Plan& P = p->P;
    CK(cudaSetDevice(p->device));
    {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&p->stream, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithPriority(&p->stream2, cudaStreamNonBlocking, lo));
        CK(cudaEventCreateWithFlags(&p->evs0, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p->evs1, cudaEventDisableTiming));
        p->pst[0][0] = p->stream; p->pst[0][1] = p->stream2; p->pev[0][0] = p->evs0; p->pev[0][1] = p->evs1;
        for (int v = 0; v < 4; ++v) CK(cudaEventCreateWithFlags(&p->pev2[v], cudaEventDisableTiming));
        CK(cudaStreamCreateWithPriority(&p->stream3, cudaStreamNonBlocking, lo));
        CK(cudaStreamCreateWithPriority(&p->stream_c, cudaStreamNonBlocking, hi));
        CK(cudaEventCreateWithFlags(&p->evsc, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p->ev3a, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p->ev3b, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p->ev3c, cudaEventDisableTiming));
        for (int v = 1; v < 4; ++v) {                 // panel streams share the top priority; update streams rank by pipeline
            CK(cudaStreamCreateWithPriority(&p->pst[v][0], cudaStreamNonBlocking, hi));
            CK(cudaStreamCreateWithPriority(&p->pst[v][1], cudaStreamNonBlocking, std::min(lo, hi + 1 + v)));
            CK(cudaEventCreateWithFlags(&p->pev[v][0], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&p->pev[v][1], cudaEventDisableTiming));
        }
    }
    CK(cudaEventCreate(&p->ev0)); CK(cudaEventCreate(&p->ev1));
    CK(cudaEventCreate(&p->evg0)); CK(cudaEventCreate(&p->evg1));
    std::vector<DFront> df(P.fronts.size());
    for (size_t i = 0; i < df.size(); ++i) {
        const Front& F = P.fronts[i];
        df[i] = DFront{F.fofs, F.relofs, F.wofs, F.F0, F.pbofs, F.W, F.R, F.m, F.ld, F.parent, F.child0, F.nchild, F.c0, F.nch, F.ps0, F.nps, P.ownofs.empty() ? -1 : P.ownofs[i]};
    }
    std::vector<DChunk> dc(P.chunks.size());
    std::vector<int32_t> cpfx(P.chunks.size() + 1, 0);
    for (size_t i = 0; i < dc.size(); ++i) {
        const Chunk& C = P.chunks[i];
        dc[i] = DChunk{C.lofs, C.uofs, C.posofs, C.fofs, C.nj, C.jlen, C.o, C.ld};
        int64_t ne = (int64_t)C.jlen * C.nj + (P.lu ? (int64_t)(C.jlen - C.nj) * C.nj : 0);
        cpfx[i + 1] = cpfx[i] + cdiv(ne, CHUNK_EPB);
    }
    p->chunk_blocks = cpfx.back();
    {   // chunk lists by level of the front tree (overlapped write-back of the factors, see k_chunks_store_list).
        // Class 0 = the fronts this part factors in phase 0 (all fronts of a single-part plan), class 1 = the top set.
        const int nl = std::max(P.nlevels, 1);
        std::vector<int32_t> slist, spfx;
        for (int cls = 0; cls < 2; ++cls) {
            std::vector<std::vector<int32_t>> byl(nl);
            for (size_t i = 0; i < dc.size(); ++i) {
                const int32_t f = P.chunks[i].front;
                const int32_t ow = P.nparts > 1 ? P.owner[f] : 0;
                const bool in = cls == 0 ? (P.nparts > 1 ? ow == P.part : true) : (P.nparts > 1 && ow == -1);
                if (in) byl[P.fronts[f].level].push_back((int32_t)i);
            }
            p->st_list0[cls].assign(nl, 0); p->st_pfx0[cls].assign(nl, 0); p->st_count[cls].assign(nl, 0); p->st_blocks[cls].assign(nl, 0);
            for (int l = 0; l < nl; ++l) {
                p->st_list0[cls][l] = (int32_t)slist.size(); p->st_pfx0[cls][l] = (int32_t)spfx.size(); p->st_count[cls][l] = (int32_t)byl[l].size();
                int32_t acc = 0; spfx.push_back(0);
                for (int32_t ci : byl[l]) { slist.push_back(ci); acc += cpfx[ci + 1] - cpfx[ci]; spfx.push_back(acc); }
                p->st_blocks[cls][l] = acc;
            }
        }
        CK(upload(&p->d_stlist, slist));
        CK(upload(&p->d_stpfx, spfx));
    }
    CK(upload(&p->d_fronts, df));
    CK(upload(&p->d_chunks, dc));
    CK(upload(&p->d_chunkpfx, cpfx));
    CK(upload(&p->d_psteps, P.psteps));
    CK(upload(&p->d_subw, P.subw));
    CK(upload(&p->d_childlist, P.childlist));
    CK(upload(&p->d_rel, P.rel));
    CK(upload(&p->d_pos, P.pos));
    CK(upload(&p->d_blkpfx, P.blkpfx));
    CK(upload(&p->d_gathert, P.gathert));
    CK(upload(&p->d_pslist, P.pslist));
    CK(upload(&p->d_asmt, P.asmt));
    CK(upload(&p->d_gemmt, P.gemmt));
    CK(upload(&p->d_solvet, P.solvet));
    if (p->inv_mask < 0) p->inv_mask = P.lu ? 3 : 0;
    p->use_inv = p->inv_mask != 0;
    if (p->use_inv && P.maxpw <= 110 && P.solve_on_fronts) {       // two w x (w|1) blocks of the widest step must fit in shared memory
        CK(cudaMalloc((void**)&p->d_tinvf, (size_t)std::max<int64_t>(P.tinv_len, 1) * sizeof(double)));
        CK(cudaMalloc((void**)&p->d_tinvb, (size_t)std::max<int64_t>(P.tinv_len, 1) * sizeof(double)));
        std::vector<int32_t> small, large;
        for (size_t i = 0; i < P.psteps.size(); ++i) if (P.psteps[i].fofs >= 0) (P.psteps[i].w <= p->inv_maxw_small ? small : large).push_back((int32_t)i);
        p->inv_nsmall = (int32_t)small.size(); p->inv_nlarge = (int32_t)large.size();
        small.insert(small.end(), large.begin(), large.end());
        CK(upload(&p->d_invlist, small));
    } else p->use_inv = false;
    CK(upload(&p->d_flowt, P.flowt));
    p->flow_ints = (int64_t)P.nflowctr * 4;                                      // 4 right-hand-side groups per batch of 32
    CK(cudaMalloc((void**)&p->d_flow, (size_t)std::max<int64_t>(p->flow_ints, 1) * sizeof(int32_t)));
    CK(upload(&p->d_tiles, P.tiles));
    CK(cudaMalloc((void**)&p->d_tilectr, (size_t)std::max(P.nctr, 1) * sizeof(int32_t)));
    CK(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device));
    CK(upload(&p->d_fown, P.fown));
    CK(upload(&p->d_fillt, P.fillt));
    CK(cudaMalloc((void**)&p->d_F, std::max<int64_t>(P.arena, 1) * sizeof(double)));
    CK(cudaMalloc((void**)&p->d_lnz, std::max<int64_t>(P.nlnz, 1) * sizeof(double)));
    CK(cudaMalloc((void**)&p->d_unz, std::max<int64_t>(P.nunz, 1) * sizeof(double)));
    CK(cudaMalloc((void**)&p->d_ipiv, P.n * sizeof(int32_t)));
    CK(cudaMalloc((void**)&p->d_iflag, sizeof(int32_t)));
    {   // arrival counters of the fused backward solve step: one per (rhs, block slot) of the largest launch
        int64_t mx = 1;
        for (const auto* Ls : {&P.bwd_launches, &P.bwd_local, &P.bwd_top}) for (const Launch& L : *Ls) mx = std::max<int64_t>(mx, L.nblocks);
        CK(cudaMalloc((void**)&p->d_counters, mx * 32 * sizeof(int32_t)));
        CK(cudaMemset(p->d_counters, 0, mx * 32 * sizeof(int32_t)));
    }
    CK(cudaMemset(p->d_ipiv, 0, P.n * sizeof(int32_t)));
    p->dev_bytes = (size_t)(P.nlnz + P.nunz + P.arena) * 8 + (size_t)P.n * 4 + (P.rel.size() + P.pos.size()) * 4 +
                   P.blkpfx.size() * 4 + P.gemmt.size() * sizeof(GemmTask) + P.solvet.size() * sizeof(SolveTask) +
                   dc.size() * sizeof(DChunk) + df.size() * sizeof(DFront);
    // diagonal-block kernel: shared-memory copy of the block when it fits
    int maxnj = P.maxpw;
    size_t need = (size_t)maxnj * (maxnj | 1) * sizeof(double);
    size_t cap = 200 * 1024;
    if (need > cap) { int q = 1; while ((size_t)(q + 1) * ((q + 1) | 1) * 8 <= cap) ++q; p->diag_smem_nj = q; need = (size_t)q * (q | 1) * 8; }
    else p->diag_smem_nj = maxnj;
    p->diag_smem_bytes = need;
    p->panel_smem = 0;
    for (int w = 1; w <= P.maxpw; ++w) p->panel_smem = std::max(p->panel_smem, panel_smem_bytes(w));   // not monotone in w (staging of T)
    if (p->panel_smem > 220 * 1024) { set_err("panel step too wide for shared memory (reduce maxblocksize)"); return -100; }
    {
        int64_t rc = init_kernel_attributes(p->device);
        if (rc) return rc;
        int optin = 0;
        CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device));
        const size_t worst = std::max({need, p->panel_smem, pstep_smem_bytes(P.maxpw), pstep_smem_bytes_mr(P.maxpw, SOLVE_NR, true),
                                       pstep_smem_bytes_mr(P.maxpw, SOLVE_NR, false)});
        if (worst + 1024 > (size_t)optin) { set_err("panel step too wide for shared memory (reduce maxblocksize)"); return -100; }
    }
    return 0;
}

SPK_API spk_plan* spk_plan_create(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                                  const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz,
                                  const int64_t* xunz_or_null, int32_t device, int32_t part, int32_t nparts) {
    spk_plan* p = new spk_plan();
    p->device = device;
    p->P.part = part; p->P.nparts = nparts < 1 ? 1 : nparts;
    if (part < 0 || part >= p->P.nparts) { set_err("bad part / nparts"); delete p; return nullptr; }
    plan_env_overrides(p->P);
    if (!analyze(p->P, n, nsuper, xsuper, snode, xlindx, lindx, xlnz, xunz_or_null)) {
        set_err("analyze: " + p->P.error); delete p; return nullptr;
    }
    plan_env_overrides(p->P);
    if (const char* e = getenv("SPK_DMMA_VARIANT")) p->dmma_variant = atoi(e);
    if (const char* e = getenv("SPK_DMMA_VARIANT64")) p->dmma_variant64 = atoi(e);
    // persistent blocks hold their SM slots for the whole launch, which defeats the stream priorities the look-ahead
    // relies on (measured: 243.6 ms persistent with 32 reserved slots vs 236.6 ms one block per tile); off by default
    if (const char* e = getenv("SPK_SOLVE_INV")) p->inv_mask = atoi(e) & 3;
    if (const char* e = getenv("SPK_DMMA_PERSIST")) p->dmma_persist = e[0] != '0';
    if (const char* e = getenv("SPK_DMMA_CA")) p->dmma_flags = (p->dmma_flags & ~1) | (e[0] != '0' ? 1 : 0);
    if (const char* e = getenv("SPK_DMMA_STATIC")) p->dmma_flags = (p->dmma_flags & ~2) | (e[0] != '0' ? 2 : 0);
    if (const char* e = getenv("SPK_DMMA_DEPHASE")) p->dmma_flags = (p->dmma_flags & 255) | ((atoi(e) * 1000 / 256) << 8);
    if (!p->dmma_persist) p->dmma_flags |= 2;         // one block per tile: tile = blockIdx.x
    if (const char* e = getenv("SPK_DIAG_SMEM")) p->diag_smem_only = e[0] == '1';
    if (const char* e = getenv("SPK_SOLVE_GRAPH")) p->solve_graphs = e[0] != '0';
    if (const char* e = getenv("SPK_DIAG_TG")) p->diag_tg = atoi(e);
    if (const char* e = getenv("SPK_PANEL_SMEM")) p->panel_smem_only = e[0] == '1';
    if (const char* e = getenv("SPK_PANEL_REG_MINW")) p->panel_reg_minw = atoi(e);
    if (const char* e = getenv("SPK_STORE_OVERLAP")) p->store_overlap = e[0] != '0';
    if (const char* e = getenv("SPK_PDL")) p->pdl = e[0] != '0';
    if (const char* e = getenv("SPK_PDL_FACTOR")) p->pdl_factor = e[0] != '0';
    if (const char* e = getenv("SPK_PDL_GEMM")) p->pdl_gemm = e[0] == '1';
    build_schedule(p->P);
    if (device < 0) return p;                      // host-only plan: structure statistics without a GPU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) {
        cudaGetLastError();
        set_err("no CUDA device " + std::to_string(device) + " (the numeric path has no CPU fallback)");
        delete p; return nullptr;
    }
    if (plan_upload(p) != 0) { spk_plan_destroy(p); return nullptr; }
    return p;
}

#define NEED_DEV(p) do { if (!(p) || (p)->device < 0) { set_err("plan has no device"); return -100; } CK(cudaSetDevice((p)->device)); } while (0)

static int64_t build_inverses(spk_plan* p, cudaStream_t st);

SPK_API int64_t spk_plan_set_values(spk_plan* p, const double* lnz, const double* unz) {
    NEED_DEV(p);
    CK(cudaMemcpyAsync(p->d_lnz, lnz, p->P.nlnz * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    if (p->P.lu && p->P.nunz > 0) {
        if (!unz) { set_err("unz required for LU"); return -100; }
        CK(cudaMemcpyAsync(p->d_unz, unz, p->P.nunz * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    }
    CK(cudaStreamSynchronize(p->stream));
    p->factored = false; p->values_in_fronts = false;
    return 0;
}

SPK_API int64_t spk_plan_set_factors(spk_plan* p, const double* lnz, const double* unz, const int64_t* ipvt) {
    int64_t rc = spk_plan_set_values(p, lnz, unz);
    if (rc) return rc;
    if (p->P.lu && ipvt) {
        int64_t* tmp = nullptr;
        CK(cudaMalloc((void**)&tmp, p->P.n * sizeof(int64_t)));
        cudaError_t e = cudaMemcpy(tmp, ipvt, p->P.n * sizeof(int64_t), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) { k_ipiv_narrow<<<cdiv(p->P.n, 256), 256, 0, p->stream>>>(p->P.n, tmp, p->d_ipiv); e = cudaStreamSynchronize(p->stream); }
        cudaFree(tmp);
        CK(e);
    }
    {   // the solve sweeps read the frontal matrices: rebuild them from the uploaded factors
        DevCtx c = make_ctx(p);
        CK(cudaMemsetAsync(p->d_F, 0, p->P.arena * sizeof(double), p->stream));
        if (p->P.lu) k_chunks<false><<<p->chunk_blocks, 256, 0, p->stream>>>(c, p->d_chunkpfx, (int)p->P.chunks.size());
        else k_chunks<false, true><<<p->chunk_blocks, 256, 0, p->stream>>>(c, p->d_chunkpfx, (int)p->P.chunks.size());   // + U = D L^T
        int64_t rci = build_inverses(p, p->stream); if (rci) return rci;
        CK(cudaStreamSynchronize(p->stream));
    }
    p->factored = true;
    return 0;
}

// destination slot in the reference layout (1-based lnz slot, or -(1-based unz slot)) -> arena element
static int64_t slot_to_arena(const Plan& P, int64_t d) {
    if (d == 0) return -1;
    if (d > 0) {
        int64_t slot = d - 1;
        size_t lo = 0, hi = P.chunks.size();
        while (hi - lo > 1) { size_t mid = (lo + hi) >> 1; if (P.chunks[mid].lofs <= slot) lo = mid; else hi = mid; }
        const Chunk& c = P.chunks[lo];
        int64_t e = slot - c.lofs;
        int64_t j = e / c.jlen, i = e - j * c.jlen;
        if (j >= c.nj) return -2;
        if (c.fofs < 0) return -1;                      // a front this part holds no storage for
        return c.fofs + P.pos[c.posofs + i] + (int64_t)(c.o + j) * c.ld;
    }
    int64_t slot = -d - 1;
    size_t lo = 0, hi = P.chunks.size();
    while (hi - lo > 1) { size_t mid = (lo + hi) >> 1; if (P.chunks[mid].uofs <= slot) lo = mid; else hi = mid; }
    const Chunk& c = P.chunks[lo];
    int64_t ldu = c.jlen - c.nj, e = slot - c.uofs;
    if (ldu <= 0) return -2;
    int64_t j = e / ldu, i = e - j * ldu;
    if (j >= c.nj) return -2;
    if (c.fofs < 0) return -1;
    return c.fofs + (int64_t)(c.o + j) + (int64_t)P.pos[c.posofs + c.nj + i] * c.ld;
}

SPK_API int64_t spk_plan_inmatrix(spk_plan* p, int64_t nnz, const int64_t* dest, const double* nzval) {
    NEED_DEV(p);
    if (nnz < 0) { set_err("inmatrix: nnz < 0"); return -100; }
    if (!dest && (!p->d_dest || nnz != p->map_nnz)) {
        set_err("inmatrix: no destination map for this nnz (pass dest on the first call and whenever the pattern changes)");
        return -100;
    }
    if (dest) {
        std::vector<int64_t> ad((size_t)nnz);
        for (int64_t k = 0; k < nnz; ++k) {
            ad[k] = slot_to_arena(p->P, dest[k]);
            if (ad[k] == -2) { set_err("inmatrix: destination outside the factor structure"); return -100; }   // the old map (if any) stays valid
        }
        if (nnz > p->nzcap || !p->d_dest) {
            if (p->d_nzval) cudaFree(p->d_nzval);
            if (p->d_dest) cudaFree(p->d_dest);
            p->d_nzval = nullptr; p->d_dest = nullptr; p->nzcap = 0; p->map_nnz = -1;
            CK(cudaMalloc((void**)&p->d_nzval, std::max<int64_t>(nnz, 1) * sizeof(double)));
            CK(cudaMalloc((void**)&p->d_dest, std::max<int64_t>(nnz, 1) * sizeof(int64_t)));
            p->nzcap = nnz;
        }
        p->map_nnz = -1;
        CK(cudaMemcpy(p->d_dest, ad.data(), nnz * sizeof(int64_t), cudaMemcpyHostToDevice));
        p->map_nnz = nnz;
    }
    // the arena clear (HBM-bound, ms) runs on the second stream beside the upload of A's values (PCIe-bound)
    CK(cudaEventRecord(p->ev3a, p->stream)); CK(cudaStreamWaitEvent(p->stream2, p->ev3a, 0));
    CK(cudaMemsetAsync(p->d_F, 0, p->P.arena * sizeof(double), p->stream2));
    CK(cudaEventRecord(p->ev3b, p->stream2));
    CK(cudaMemcpyAsync(p->d_nzval, nzval, nnz * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamWaitEvent(p->stream, p->ev3b, 0));
    if (nnz > 0) k_scatter_values<<<cdiv(nnz, 256), 256, 0, p->stream>>>(nnz, p->d_dest, p->d_nzval, p->d_F);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(p->stream));
    p->factored = false; p->values_in_fronts = true; p->nz_last = nnz;
    return 0;
}

// zero the fronts and re-scatter the values uploaded by the last spk_plan_inmatrix (no host traffic)
SPK_API int64_t spk_plan_reassemble(spk_plan* p) {
    NEED_DEV(p);
    if (!p->d_dest || p->nz_last <= 0) { set_err("spk_plan_inmatrix has not been called"); return -100; }
    CK(cudaEventRecord(p->ev0, p->stream));            // the factor time reported by the next spk_plan_factor includes this clear + scatter
    p->ev0_armed = true;
    CK(cudaMemsetAsync(p->d_F, 0, p->P.arena * sizeof(double), p->stream));
    k_scatter_values<<<cdiv(p->nz_last, 256), 256, 0, p->stream>>>(p->nz_last, p->d_dest, p->d_nzval, p->d_F);
    CK(cudaGetLastError());
    p->factored = false; p->values_in_fronts = true;
    return 0;
}

// explicit inverses of the diagonal blocks of every panel step this part holds (solve, see k_diag_inverse)
static int64_t build_inverses(spk_plan* p, cudaStream_t st) {
    if (!p->use_inv) return 0;
    p->inv_ready = false;
    DevCtx c = make_ctx(p);
    const bool lu = p->P.lu;
    if (p->inv_nsmall > 0) {
        const size_t sm = inverse_smem_bytes(p->inv_maxw_small);
        if (lu) k_diag_inverse<true><<<p->inv_nsmall, 64, sm, st>>>(c, p->d_invlist, p->d_tinvf, p->d_tinvb);
        else k_diag_inverse<false><<<p->inv_nsmall, 64, sm, st>>>(c, p->d_invlist, p->d_tinvf, p->d_tinvb);
    }
    if (p->inv_nlarge > 0) {
        const size_t sm = inverse_smem_bytes(p->P.maxpw);
        if (lu) k_diag_inverse<true><<<p->inv_nlarge, 64, sm, st>>>(c, p->d_invlist + p->inv_nsmall, p->d_tinvf, p->d_tinvb);
        else k_diag_inverse<false><<<p->inv_nlarge, 64, sm, st>>>(c, p->d_invlist + p->inv_nsmall, p->d_tinvf, p->d_tinvb);
    }
    CK(cudaGetLastError());
    p->launches_factor += 2;
    p->inv_ready = true;
    return 0;
}

// ---- factor ---------------------------------------------------------------------------
static int64_t run_factor_launch(spk_plan* p, const DevCtx& c, const Launch& L, bool two_streams, int pair = 0, cudaStream_t force = nullptr) {
    const int32_t* pfx = p->d_blkpfx + L.pfx;
    cudaStream_t st = p->stream;
    if (force) { st = force; two_streams = false; }     // the caller handles waits / records (lists with broadcasts)
    else if (two_streams) {
        st = p->pst[pair][L.stream ? 1 : 0];
        if (L.wait_other & 1) CK(cudaStreamWaitEvent(st, p->pev[pair][L.stream ? 0 : 1], 0));
        if (L.wait_other & 2) CK(cudaStreamWaitEvent(st, p->pev2[pair], 0));          // the early part of the previous trailing update only
    }
    const bool lu = p->P.lu;
    switch (L.kind) {
    case K_ASM:
        if (force && p->P.dist_top) k_assemble<true><<<L.nblocks, ASM_TPB, 0, st>>>(c, p->d_asmt + L.first, pfx, L.count);
        else k_assemble<false><<<L.nblocks, ASM_TPB, 0, st>>>(c, p->d_asmt + L.first, pfx, L.count);
        break;
    case K_ASM_TAIL:
        k_assemble_tail<<<L.count, 256, 0, st>>>(c, p->d_asmt + L.first, L.count); break;
    case K_DIAG: {
        // shared memory sized for the widest block of THIS launch (tiny fronts keep high occupancy)
        int wl = std::min(L.maxw, p->diag_smem_nj);
        size_t sm = (size_t)wl * (wl | 1) * sizeof(double);
        if (lu && L.maxw <= 64 && !p->diag_smem_only) CK(launch_pdl(k_diag_lu_row, dim3(L.count), dim3(64 * LU_NS), 0, st, p->pdl_factor, c, (const int32_t*)(p->d_pslist + L.first)));
        else if (lu) k_diag<true><<<L.count, 256, sm, st>>>(c, p->d_pslist + L.first, wl);
        else if (L.maxw <= 64 && !p->diag_smem_only && p->diag_tg == 1) CK(launch_pdl(k_diag_ldlt_row, dim3(L.count), dim3(64 * DIAG_NS), 0, st, p->pdl_factor, c, (const int32_t*)(p->d_pslist + L.first)));
        else if (L.maxw <= 64 && !p->diag_smem_only && p->diag_tg == 8) k_diag_ldlt_reg<8, 8><<<L.count, 64, 0, st>>>(c, p->d_pslist + L.first);
        else if (L.maxw <= 64 && !p->diag_smem_only) k_diag_ldlt_reg<4, 16><<<L.count, 256, 0, st>>>(c, p->d_pslist + L.first);
        else if (L.maxw <= 96 && !p->diag_smem_only) k_diag_ldlt_reg<6, 16><<<L.count, 256, 0, st>>>(c, p->d_pslist + L.first);
        else if (L.maxw <= 128 && !p->diag_smem_only) k_diag_ldlt_reg<8, 16><<<L.count, 256, 0, st>>>(c, p->d_pslist + L.first);
        else k_diag<false><<<L.count, 256, sm, st>>>(c, p->d_pslist + L.first, wl);
        break;
    }
    case K_PANEL: {
        size_t sm = 0;                                   // tasks narrower than maxw may stage T: size for the worst case
        for (int w = 1; w <= L.maxw; ++w) sm = std::max(sm, panel_smem_bytes(w));
        // register kernel for 32 < w <= 64; narrower steps (the tiny fronts at the bottom of the tree: tens of thousands
        // of blocks of a few rows) keep the shared-memory kernel, whose footprint scales with w (measured per level)
        const bool fast = L.maxw <= 64 && L.maxw > p->panel_reg_minw && !p->panel_smem_only;
        if (fast) {
            if (lu) CK(launch_pdl(k_panel_reg<true>, dim3(L.nblocks * PANEL_REG_SPLIT), dim3(PANEL_REG_THREADS), 0, st, p->pdl_factor, c, (const int32_t*)(p->d_pslist + L.first), pfx, (int)L.count));
            else CK(launch_pdl(k_panel_reg<false>, dim3(L.nblocks * PANEL_REG_SPLIT), dim3(PANEL_REG_THREADS), 0, st, p->pdl_factor, c, (const int32_t*)(p->d_pslist + L.first), pfx, (int)L.count));
        }
        else if (lu) k_panel<true><<<L.nblocks, PANEL_ROWS, sm, st>>>(c, p->d_pslist + L.first, pfx, L.count, 0);
        else k_panel<false><<<L.nblocks, PANEL_ROWS, sm, st>>>(c, p->d_pslist + L.first, pfx, L.count, 0);
        break;
    }
    case K_GEMM:
        if (L.maxw > 0 && L.maxw <= GEMM_TINY && L.nblocks == L.count) k_gemm_tiny<<<L.count, 128, 0, st>>>(c, p->d_gemmt + L.first, L.count);
        else k_gemm_small<<<L.nblocks, 256, 0, st>>>(c, p->d_gemmt + L.first, pfx, L.count);
        break;
    case K_GEMM_B64:
    case K_GEMM_T64: {
        GemmVariant v = gemm_dmma_variant(L.kind, p->dmma_variant, p->dmma_variant64, L.tile_n);
        int cap = v.blocks_per_sm * p->num_sms - (L.reserve > 0 ? L.reserve * v.blocks_per_sm / 2 : 0);
        if (cap < p->num_sms) cap = p->num_sms;
        const int grid = p->dmma_persist ? std::min<int>(L.ntiles, cap) : L.ntiles;      // SPK_DMMA_PERSIST=0: one block per tile
        CK(launch_pdl(v.fn, dim3(grid), dim3(v.threads), v.smem, st, p->pdl_factor && p->pdl_gemm, c, (const GemmTask*)(p->d_gemmt + L.first),
                      (const GemmTile*)(p->d_tiles + L.tile0), (int)L.ntiles, (int32_t*)(p->d_tilectr + L.ctr), (int)p->dmma_flags));
        break;
    }
    case K_FRONT_SMALL: {
        const size_t sm = front_small_smem_bytes(std::max(L.maxw, 1));
        if (lu) k_front_small<true><<<L.count, FS_NT, sm, st>>>(c, p->d_pslist + L.first);
        else k_front_small<false><<<L.count, FS_NT, sm, st>>>(c, p->d_pslist + L.first);
        break;
    }
    case K_FILLU:
        k_fill_u<<<L.nblocks, 256, 0, st>>>(c, p->d_fillt + L.first, pfx, L.count); break;
    default:
        set_err("bad launch kind"); return -100;
    }
    if (two_streams && (L.record & 1)) CK(cudaEventRecord(p->pev[pair][L.stream ? 1 : 0], st));
    if (two_streams && (L.record & 2)) CK(cudaEventRecord(p->pev2[pair], st));
    return 0;
}

// Launch list of the top set of a multi-part plan: three streams (0 panel, 1 trailing update, 2 communication),
// K_BCAST launches = in-place NCCL broadcasts of column slabs on the communication stream.  Every part enqueues
// the broadcasts in the same (static) order, so no host-side coordination is needed while the list runs.
static int64_t run_top_list(spk_plan* p, const DevCtx& c, const std::vector<Launch>& Ls) {
    Plan& P = p->P;
    NcclApi* N = nccl_api();
    cudaStream_t S[3] = {p->stream, p->stream2, p->stream_c};
    cudaEvent_t E[3] = {p->evs0, p->evs1, p->evsc};
    for (int q = 0; q < 3; ++q) CK(cudaEventRecord(E[q], S[q]));
    const char* trace_path = getenv("SPK_TRACE_TOP");   // completion time of every launch of the top-set list (profiling only)
    std::vector<cudaEvent_t> tev;
    if (trace_path) { tev.resize(Ls.size() + 1); for (auto& e : tev) cudaEventCreate(&e); cudaEventRecord(tev[0], S[0]); }
    size_t li = 0;
    for (const Launch& L : Ls) {
        const int sid = L.stream < 3 ? L.stream : 0;
        cudaStream_t st = S[sid];
        if ((L.wait_other & 1) && sid < 2) CK(cudaStreamWaitEvent(st, E[1 - sid], 0));        // replicated lists (two-stream look-ahead)
        if ((L.wait_other & 2) && sid < 2) CK(cudaStreamWaitEvent(st, p->pev2[0], 0));
        for (int q = 0; q < 3; ++q) if (((L.wait_mask >> q) & 1) && q != sid) CK(cudaStreamWaitEvent(st, E[q], 0));
        if (L.kind == K_BCAST) {
            if (!p->comm || !N) { set_err("multi-part plan without a communicator (spk_plan_comm_init)"); return -100; }
            NK(N->GroupStart());
            for (int32_t i = 0; i < L.count; ++i) {
                const Bcast& b = P.bcasts[L.first + i];
                NK(N->Broadcast(p->d_F + b.ofs, p->d_F + b.ofs, (size_t)b.len, ncclDouble, b.root, p->comm, st));
            }
            NK(N->GroupEnd());
            CK(cudaEventRecord(E[sid], st));
            if (!P.dist_top) { CK(cudaStreamWaitEvent(S[0], E[2], 0)); CK(cudaStreamWaitEvent(S[1], E[2], 0)); }   // replicated top set: everything waits for the exchange
        } else {
            int64_t rc = run_factor_launch(p, c, L, false, 0, st);
            if (rc) return rc;
            if (L.record & 1) CK(cudaEventRecord(E[sid], st));
            if (L.record & 2) CK(cudaEventRecord(p->pev2[0], st));
            if (!p->profile && (L.kind == K_GEMM_B64 || L.kind == K_GEMM_T64)) p->gemm_flops += L.flops;   // profiling mode times phase 0 only
        }
        ++p->launches_factor;
        if (trace_path) cudaEventRecord(tev[++li], st);
    }
    // join on the main stream
    CK(cudaEventRecord(E[1], S[1])); CK(cudaStreamWaitEvent(S[0], E[1], 0));
    CK(cudaEventRecord(E[2], S[2])); CK(cudaStreamWaitEvent(S[0], E[2], 0));
    if (trace_path) {
        CK(cudaStreamSynchronize(S[0]));
        std::string path = std::string(trace_path) + "." + std::to_string(P.part);
        if (FILE* f = fopen(path.c_str(), "w")) {
            fprintf(f, "idx,kind,level,step,stream,tasks,blocks,flops,wait_mask,t_end_ms\n");
            for (size_t i = 0; i < Ls.size(); ++i) {
                float t = 0; cudaEventElapsedTime(&t, tev[0], tev[i + 1]);
                fprintf(f, "%zu,%d,%d,%d,%d,%d,%d,%.6g,%d,%.6f\n", i, Ls[i].kind, Ls[i].level, Ls[i].step, (int)Ls[i].stream, Ls[i].count, Ls[i].nblocks, Ls[i].flops, (int)Ls[i].wait_mask, t);
            }
            fclose(f);
        }
        for (auto& e : tev) cudaEventDestroy(e);
    }
    return 0;
}

// phase -1: everything (single part).  Multi-part plans: phase 0 = this part's subtrees, phase 1 = the
// top set (after the caller exchanged the subtree-root fronts, see spk_plan_xchg_info) + write-back.
SPK_API int64_t spk_plan_factor_phase(spk_plan* p, int32_t phase) {
    NEED_DEV(p);
    Plan& P = p->P;
    if (P.nparts > 1 && phase < 0) { set_err("multi-part plan: call phases 0 and 1 with the exchange in between"); return -100; }
    if (P.nparts <= 1) phase = -1;
    const std::vector<Launch>& Ls = phase < 0 ? P.factor_launches : (phase == 0 ? P.factor_local : P.factor_top);
    DevCtx c = make_ctx(p);
    cudaStream_t st = p->stream;
    if (phase == 1) CK(cudaEventRecord(p->evg0, st));
    else if (!p->ev0_armed) CK(cudaEventRecord(p->ev0, st));
    p->ev0_armed = false;
    if (phase == 1) {
        // top set of a multi-part plan: exchange + (distributed or replicated) factorisation, then the write-back
        int64_t rc = run_top_list(p, c, Ls);
        if (rc) return rc;
        for (int lv = 0; lv < (int)p->st_count[1].size(); ++lv)
            if (p->st_blocks[1][lv] > 0) {
                k_chunks_store_list<<<p->st_blocks[1][lv], 256, 0, st>>>(c, p->d_stlist + p->st_list0[1][lv], p->d_stpfx + p->st_pfx0[1][lv], p->st_count[1][lv]);
                ++p->launches_factor;
            }
        { int64_t rci = build_inverses(p, st); if (rci) return rci; }
        CK(cudaGetLastError());
        CK(cudaEventRecord(p->ev1, st));
        CK(cudaStreamSynchronize(st));
        p->values_in_fronts = false;
        float ms1 = 0; CK(cudaEventElapsedTime(&ms1, p->evg0, p->ev1)); p->ms_phase1 = ms1;
        if (p->phase0_async) { float ms0 = 0; CK(cudaEventElapsedTime(&ms0, p->ev0, p->evg1)); p->ms_phase0 = ms0; p->phase0_async = false; }
        p->ms_factor = p->ms_phase0 + ms1;
        int32_t flag1 = 0;
        CK(cudaMemcpy(&flag1, p->d_iflag, sizeof(int32_t), cudaMemcpyDeviceToHost));
        p->factored = true;
        return flag1;
    }
    if (phase <= 0) {
        CK(cudaMemsetAsync(p->d_iflag, 0, sizeof(int32_t), st));
        CK(cudaMemsetAsync(p->d_tilectr, 0, (size_t)std::max(P.nctr, 1) * sizeof(int32_t), st));    // tile counters of the persistent DMMA launches
        p->launches_factor = 0; p->gemm_flops = 0; p->gemm_ms = 0;
        p->inv_ready = false;
        if (!p->values_in_fronts) {
            // gather the assembled matrix (reference layout) into the zeroed frontal matrices
            CK(cudaMemsetAsync(p->d_F, 0, P.arena * sizeof(double), st));
            k_chunks<false><<<p->chunk_blocks, 256, 0, st>>>(c, p->d_chunkpfx, (int)P.chunks.size());
            ++p->launches_factor;
        }
    }
    std::vector<cudaEvent_t> evs;
    if (p->profile) { evs.resize(Ls.size() + 1); for (auto& e : evs) cudaEventCreate(&e); cudaEventRecord(evs[0], st); }
    size_t li = 0;
    const bool two = P.lookahead && !p->profile;      // per-launch timing needs the serial order on one stream
    if (two) { CK(cudaEventRecord(p->evs0, st)); CK(cudaEventRecord(p->evs1, p->stream2)); }
    const char* trace_path = getenv("SPK_TRACE");     // completion time of every launch in the real two-stream run
    const bool trace = trace_path && two && phase < 0;
    std::vector<cudaEvent_t> tev;
    if (trace) { tev.resize(Ls.size() + 1); for (auto& e : tev) cudaEventCreate(&e); cudaEventRecord(tev[0], st); }
    const bool piped = phase < 0 && two && p->use_pipes && !P.factor_pipe.empty() && !trace;
    if (piped) {
        // tree pipelines: every pipeline's streams start after the load on the main stream; launches are
        // enqueued round-robin so that no pipeline waits for the CPU; the top set follows on pair 0
        const int V = (int)P.factor_pipe.size();
        for (int v = 0; v < V; ++v) for (int q = 0; q < 2; ++q) {
            if (v == 0 && q == 0) continue;
            CK(cudaStreamWaitEvent(p->pst[v][q], p->evs0, 0));
            CK(cudaEventRecord(p->pev[v][q], p->pst[v][q]));
        }
        std::vector<size_t> at(V, 0);
        for (bool more = true; more;) {
            more = false;
            for (int v = 0; v < V; ++v) {
                if (at[v] >= P.factor_pipe[v].size()) continue;
                const Launch& L = P.factor_pipe[v][at[v]++];
                int64_t rc = run_factor_launch(p, c, L, true, v);
                if (rc) return rc;
                ++p->launches_factor;
                if (L.kind == K_GEMM_B64 || L.kind == K_GEMM_T64) p->gemm_flops += L.flops;
                more = true;
            }
        }
        for (int v = 0; v < V; ++v) for (int q = 0; q < 2; ++q) {     // join on the main stream
            if (v == 0 && q == 0) continue;
            CK(cudaEventRecord(p->pev[v][q], p->pst[v][q]));
            CK(cudaStreamWaitEvent(st, p->pev[v][q], 0));
        }
        CK(cudaEventRecord(p->evs0, st));
        CK(cudaStreamWaitEvent(p->stream2, p->evs0, 0));
        CK(cudaEventRecord(p->evs1, p->stream2));
    }
    // write-back of a finished level on the third stream, under the factorisation of the next levels
    // (multi-part plans: phase 0 writes back the part's own subtrees, phase 1 the top set — not the other parts' chunks)
    const bool ovl = two && !piped && !trace && p->store_overlap && !Ls.empty();
    const int scls = phase == 1 ? 1 : 0;
    auto store_level = [&](int lev, cudaStream_t s3) -> int64_t {
        if (lev < 0 || lev >= (int)p->st_count[scls].size() || p->st_blocks[scls][lev] == 0) return 0;
        k_chunks_store_list<<<p->st_blocks[scls][lev], 256, 0, s3>>>(c, p->d_stlist + p->st_list0[scls][lev], p->d_stpfx + p->st_pfx0[scls][lev], p->st_count[scls][lev]);
        ++p->launches_factor;
        return 0;
    };
    int cur_level = Ls.empty() ? 0 : Ls[0].level;
    for (const Launch& L : (piped ? P.factor_ptop : Ls)) {
        if (ovl && L.level != cur_level) {
            CK(cudaEventRecord(p->ev3a, p->stream)); CK(cudaEventRecord(p->ev3b, p->stream2));
            CK(cudaStreamWaitEvent(p->stream3, p->ev3a, 0)); CK(cudaStreamWaitEvent(p->stream3, p->ev3b, 0));
            for (int lv = cur_level; lv < L.level; ++lv) store_level(lv, p->stream3);
            cur_level = L.level;
        }
        int64_t rc = run_factor_launch(p, c, L, two);
        if (rc) return rc;
        ++p->launches_factor;
        if (L.kind == K_GEMM_B64 || L.kind == K_GEMM_T64) p->gemm_flops += L.flops;
        if (p->profile) cudaEventRecord(evs[++li], st);
        if (trace) cudaEventRecord(tev[++li], p->pst[0][L.stream ? 1 : 0]);
    }
    if (two) { CK(cudaEventRecord(p->evs1, p->stream2)); CK(cudaStreamWaitEvent(st, p->evs1, 0)); }
    if (ovl) {
        for (int lv = cur_level; lv < (int)p->st_count[scls].size(); ++lv) store_level(lv, st);     // the last level(s), then join
        CK(cudaEventRecord(p->ev3c, p->stream3));
        CK(cudaStreamWaitEvent(st, p->ev3c, 0));
    } else if (phase != 0) {
        // scatter the factors back into the reference layout (lnz / unz)
        k_chunks<true><<<p->chunk_blocks, 256, 0, st>>>(c, p->d_chunkpfx, (int)P.chunks.size());
        ++p->launches_factor;
    }
    if (phase < 0) { int64_t rci = build_inverses(p, st); if (rci) return rci; }
    CK(cudaGetLastError());
    CK(cudaEventRecord(phase == 0 ? p->evg1 : p->ev1, st));
    if (phase == 0 && p->phase0_async) return 0;        // spk_plan_factor_multi: phase 1 follows without a host synchronisation
    CK(cudaStreamSynchronize(st));
    if (phase != 0) p->values_in_fronts = false;
    if (trace) {
        if (FILE* f = fopen(trace_path, "w")) {
            fprintf(f, "idx,kind,level,step,stream,blocks,flops,wait_other,t_end_ms\n");
            for (size_t i = 0; i < Ls.size(); ++i) {
                float t = 0; cudaEventElapsedTime(&t, tev[0], tev[i + 1]);
                fprintf(f, "%zu,%d,%d,%d,%d,%d,%.6g,%d,%.6f\n", i, Ls[i].kind, Ls[i].level, Ls[i].step, (int)Ls[i].stream, Ls[i].nblocks, Ls[i].flops, (int)Ls[i].wait_other, t);
            }
            fclose(f);
        }
        for (auto& e : tev) cudaEventDestroy(e);
    }
    float ms = 0;
    if (phase < 0) { CK(cudaEventElapsedTime(&ms, p->ev0, p->ev1)); p->ms_factor = ms; }
    else if (phase == 0) { CK(cudaEventElapsedTime(&ms, p->ev0, p->evg1)); p->ms_factor = ms; p->ms_phase0 = ms; }
    else { CK(cudaEventElapsedTime(&ms, p->evg0, p->ev1)); p->ms_phase1 = ms; p->ms_factor = p->ms_phase0 + ms; }
    if (p->profile) {
        p->launch_ms.assign(Ls.size(), 0.f);
        for (int k = 0; k < 16; ++k) { p->kind_ms[k] = 0; p->kind_n[k] = 0; }
        for (size_t i = 0; i < Ls.size(); ++i) {
            cudaEventElapsedTime(&p->launch_ms[i], evs[i], evs[i + 1]);
            p->kind_ms[Ls[i].kind & 15] += p->launch_ms[i]; p->kind_n[Ls[i].kind & 15]++;
            if (Ls[i].kind == K_GEMM_B64 || Ls[i].kind == K_GEMM_T64) p->gemm_ms += p->launch_ms[i];
        }
        for (auto& e : evs) cudaEventDestroy(e);
        if (const char* path = getenv("SPK_DUMP_LAUNCHES")) {       // per-launch CSV for profiles/
            if (FILE* f = fopen(path, "w")) {
                fprintf(f, "idx,kind,level,step,tasks,blocks,maxw,flops,ms,stream,wait_other,record\n");
                for (size_t i = 0; i < Ls.size(); ++i) {
                    const Launch& L = Ls[i];
                    fprintf(f, "%zu,%d,%d,%d,%d,%d,%d,%.6g,%.6f,%d,%d,%d\n", i, L.kind, L.level, L.step, L.count, L.nblocks, L.maxw, L.flops, p->launch_ms[i], (int)L.stream, (int)L.wait_other, (int)L.record);
                }
                fclose(f);
            }
        }
    }
    int32_t flag = 0;
    CK(cudaMemcpy(&flag, p->d_iflag, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (phase != 0) p->factored = true;
    return flag;
}

SPK_API int64_t spk_plan_factor(spk_plan* p) { return spk_plan_factor_phase(p, -1); }

// ---- multi-part plans: communicator + whole-factorisation / whole-solve entry points --------------------------
SPK_API int64_t spk_nccl_unique_id(void* out128) {
    NcclApi* N = nccl_api();
    if (!N) { set_err("libnccl.so.2 not found"); return -100; }
    ncclUniqueId id;
    NK(N->GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(out128, &id, 128);
    return 0;
}
// every part calls this with the id produced by ONE call of spk_nccl_unique_id (collective, blocks until all joined)
SPK_API int64_t spk_plan_comm_init(spk_plan* p, const void* id128) {
    NEED_DEV(p);
    NcclApi* N = nccl_api();
    if (!N) { set_err("libnccl.so.2 not found"); return -100; }
    if (p->comm) { set_err("communicator already set"); return -100; }
    ncclUniqueId id; memcpy(&id, id128, 128);
    NK(N->CommInitRank(&p->comm, p->P.nparts, id, p->P.part));
    p->comm_owned = true;
    return 0;
}
// phase 0 (own subtrees), exchange, top set, write-back: one call, no host synchronisation in between
SPK_API int64_t spk_plan_factor_multi(spk_plan* p) {
    NEED_DEV(p);
    if (p->P.nparts <= 1) return spk_plan_factor_phase(p, -1);
    if (!p->comm) { set_err("multi-part plan without a communicator (spk_plan_comm_init)"); return -100; }
    p->phase0_async = !p->profile;                  // per-launch timing reads its events at the end of phase 0
    int64_t rc = spk_plan_factor_phase(p, 0);
    if (rc < 0) { p->phase0_async = false; return rc; }
    return spk_plan_factor_phase(p, 1);
}

SPK_API int64_t spk_plan_get_factors(spk_plan* p, double* lnz, double* unz, int64_t* ipvt) {
    NEED_DEV(p);
    if (lnz) CK(cudaMemcpyAsync(lnz, p->d_lnz, p->P.nlnz * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (unz && p->P.nunz > 0) CK(cudaMemcpyAsync(unz, p->d_unz, p->P.nunz * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (ipvt && p->P.lu) {
        int64_t* tmp = nullptr;
        CK(cudaMalloc((void**)&tmp, p->P.n * sizeof(int64_t)));
        k_ipiv_widen<<<cdiv(p->P.n, 256), 256, 0, p->stream>>>(p->P.n, p->d_ipiv, tmp);
        cudaError_t e = cudaMemcpyAsync(ipvt, tmp, p->P.n * sizeof(int64_t), cudaMemcpyDeviceToHost, p->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        cudaFree(tmp);
        CK(e);
    }
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

// ---- solve ----------------------------------------------------------------------------
static int64_t ensure_w(spk_plan* p, int64_t nrhs) {
    if (nrhs > p->w_nrhs) {
        if (p->d_w) cudaFree(p->d_w);
        p->d_w = nullptr; p->w_nrhs = 0;
        CK(cudaMalloc((void**)&p->d_w, (size_t)p->P.wlen * nrhs * sizeof(double)));
        if (p->d_pb) cudaFree(p->d_pb);
        p->d_pb = nullptr;
        CK(cudaMalloc((void**)&p->d_pb, std::max<size_t>((size_t)p->P.pblen * nrhs, 1) * sizeof(double)));
        if (p->d_box) cudaFree(p->d_box);
        p->d_box = nullptr;
        if (p->P.nflowctr > 0) CK(cudaMalloc((void**)&p->d_box, (size_t)2 * p->P.n * nrhs * sizeof(double)));
        p->w_nrhs = nrhs;
    }
    return 0;
}

static int64_t run_solve_launches(spk_plan* p, const DevCtx& c0, const std::vector<Launch>& Ls, double* d_rhs,
                                  int64_t nrhs, int64_t ldrhs) {
    cudaStream_t st = p->stream;
    const bool lu = p->P.lu;
    DevCtx c_noinv = c0; c_noinv.tinvf = nullptr; c_noinv.tinvb = nullptr;
    for (const Launch& L : Ls) {
        const bool frontk = L.kind == K_PF_FRONT || L.kind == K_PB_FRONT;
        const DevCtx& c = (p->inv_mask & (frontk ? 2 : 1)) ? c0 : c_noinv;      // inverted diagonal blocks per kernel class
        const int32_t* pfx = p->d_blkpfx + L.pfx;
        const int32_t* list = p->d_gathert + L.first;
        dim3 grid(L.nblocks, (unsigned)nrhs);
        const unsigned ngrp = (unsigned)((nrhs + SOLVE_NR - 1) / SOLVE_NR);    // right-hand-side groups of the fused kernels
        switch (L.kind) {
        case K_FWD_GATHER: k_fwd_gather<<<dim3(L.count, (unsigned)nrhs), 256, 0, st>>>(c, list, d_rhs, ldrhs); break;
        case K_FWD_DIAG:
            if (lu) k_fwd_diag<true><<<dim3(L.count, (unsigned)nrhs), 128, 0, st>>>(c, list);
            else k_fwd_diag<false><<<dim3(L.count, (unsigned)nrhs), 128, 0, st>>>(c, list);
            break;
        case K_FWD_UPDATE: k_fwd_update<<<grid, UPD_ROWS, 0, st>>>(c, list, pfx, L.count); break;
        case K_FWD_FRONT:
            if (lu) k_fwd_front<true><<<dim3(L.count, (unsigned)nrhs), 256, 0, st>>>(c, list);
            else k_fwd_front<false><<<dim3(L.count, (unsigned)nrhs), 256, 0, st>>>(c, list);
            break;
        case K_BWD_FRONT:
            if (lu) k_bwd_front<true><<<dim3(L.count, (unsigned)nrhs), 256, 0, st>>>(c, list, d_rhs, ldrhs);
            else k_bwd_front<false><<<dim3(L.count, (unsigned)nrhs), 256, 0, st>>>(c, list, d_rhs, ldrhs);
            break;
        case K_PF_FRONT: {
            const bool tiny = L.maxw <= 32;                  // bottom of the tree: 64-thread blocks, 8 per SM
            const size_t sm = pstep_smem_bytes_mr(L.maxw, nrhs == 1 ? 1 : SOLVE_NR, true);
            const dim3 g(L.count, nrhs == 1 ? 1 : ngrp);
            if (nrhs == 1) {
                if (tiny) { if (lu) k_pf_front<true, 1, 64><<<g, 64, sm, st>>>(c, list, (int)nrhs); else k_pf_front<false, 1, 64><<<g, 64, sm, st>>>(c, list, (int)nrhs); }
                else { if (lu) k_pf_front<true, 1, 256><<<g, 256, sm, st>>>(c, list, (int)nrhs); else k_pf_front<false, 1, 256><<<g, 256, sm, st>>>(c, list, (int)nrhs); }
            } else {
                if (tiny) { if (lu) k_pf_front<true, SOLVE_NR, 64><<<g, 64, sm, st>>>(c, list, (int)nrhs); else k_pf_front<false, SOLVE_NR, 64><<<g, 64, sm, st>>>(c, list, (int)nrhs); }
                else { if (lu) k_pf_front<true, SOLVE_NR, 256><<<g, 256, sm, st>>>(c, list, (int)nrhs); else k_pf_front<false, SOLVE_NR, 256><<<g, 256, sm, st>>>(c, list, (int)nrhs); }
            }
            break;
        }
        case K_PB_FRONT: {
            const bool tiny = L.maxw <= 32;
            const size_t sm = pstep_smem_bytes_mr(L.maxw, nrhs == 1 ? 1 : SOLVE_NR, true);
            const dim3 g(L.count, nrhs == 1 ? 1 : ngrp);
            if (nrhs == 1) {
                if (tiny) { if (lu) k_pb_front<true, 1, 64><<<g, 64, sm, st>>>(c, list, d_rhs, ldrhs, (int)nrhs); else k_pb_front<false, 1, 64><<<g, 64, sm, st>>>(c, list, d_rhs, ldrhs, (int)nrhs); }
                else { if (lu) k_pb_front<true, 1, 256><<<g, 256, sm, st>>>(c, list, d_rhs, ldrhs, (int)nrhs); else k_pb_front<false, 1, 256><<<g, 256, sm, st>>>(c, list, d_rhs, ldrhs, (int)nrhs); }
            } else {
                if (tiny) { if (lu) k_pb_front<true, SOLVE_NR, 64><<<g, 64, sm, st>>>(c, list, d_rhs, ldrhs, (int)nrhs); else k_pb_front<false, SOLVE_NR, 64><<<g, 64, sm, st>>>(c, list, d_rhs, ldrhs, (int)nrhs); }
                else { if (lu) k_pb_front<true, SOLVE_NR, 256><<<g, 256, sm, st>>>(c, list, d_rhs, ldrhs, (int)nrhs); else k_pb_front<false, SOLVE_NR, 256><<<g, 256, sm, st>>>(c, list, d_rhs, ldrhs, (int)nrhs); }
            }
            break;
        }
        case K_PF_FLOW:
        case K_PB_FLOW: {
            const bool fwd = L.kind == K_PF_FLOW;
            const int nr = nrhs == 1 ? 1 : SOLVE_NR;
            const int mw = std::max(L.maxw, 1);
            const size_t sm = flow_smem_bytes(mw, nr);
            const dim3 g(L.count, nrhs == 1 ? 1 : ngrp);
            const FlowTask* tk = p->d_flowt + L.first;
            int32_t* ticket = p->d_flow + (size_t)L.ctr * 4;
            const int64_t bstride = p->P.n;
            double* box = p->d_box + (fwd ? 0 : (size_t)p->w_nrhs * p->P.n);
            const bool wide = mw > FLOW_NT;                   // a single panel step wider than one row per thread
#define SPK_FLOW(LUv, NRv, RPTv)                                                                                                      \
            do {                                                                                                                     \
                if (fwd) k_pf_flow<LUv, NRv, RPTv><<<g, FLOW_NT, sm, st>>>(c, tk, ticket, box, bstride, (int)nrhs, mw);               \
                else k_pb_flow<LUv, NRv, RPTv><<<g, FLOW_NT, sm, st>>>(c, tk, ticket, box, bstride, d_rhs, (int64_t)ldrhs, (int)nrhs, mw); \
            } while (0)
            if (lu) { if (nr == 1) { if (wide) SPK_FLOW(true, 1, 2); else SPK_FLOW(true, 1, 1); } else { if (wide) SPK_FLOW(true, SOLVE_NR, 2); else SPK_FLOW(true, SOLVE_NR, 1); } }
            else { if (nr == 1) { if (wide) SPK_FLOW(false, 1, 2); else SPK_FLOW(false, 1, 1); } else { if (wide) SPK_FLOW(false, SOLVE_NR, 2); else SPK_FLOW(false, SOLVE_NR, 1); } }
#undef SPK_FLOW
            break;
        }
        case K_PF_DIAG: {
            size_t sm = pstep_smem_bytes(L.maxw);
            if (lu) k_pf_diag<true><<<dim3(L.count, (unsigned)nrhs), 128, sm, st>>>(c, list);
            else k_pf_diag<false><<<dim3(L.count, (unsigned)nrhs), 128, sm, st>>>(c, list);
            break;
        }
        case K_PF_UPDATE: k_pf_update<<<grid, SV_ROWS, 0, st>>>(c, list, pfx, L.count); break;
        case K_PF_STEP: {
            if (nrhs == 1) {
                size_t sm = pstep_smem_bytes_mr(L.maxw, 1, false);
                if (lu) CK(launch_pdl(k_pf_step<true, 1>, dim3(L.nblocks, 1), dim3(SV_ROWS), sm, st, p->pdl, c, list, pfx, (int)L.count, (int)nrhs));
                else CK(launch_pdl(k_pf_step<false, 1>, dim3(L.nblocks, 1), dim3(SV_ROWS), sm, st, p->pdl, c, list, pfx, (int)L.count, (int)nrhs));
            } else {
                size_t sm = pstep_smem_bytes_mr(L.maxw, SOLVE_NR, false);
                if (lu) CK(launch_pdl(k_pf_step<true, SOLVE_NR>, dim3(L.nblocks, ngrp), dim3(SV_ROWS), sm, st, p->pdl, c, list, pfx, (int)L.count, (int)nrhs));
                else CK(launch_pdl(k_pf_step<false, SOLVE_NR>, dim3(L.nblocks, ngrp), dim3(SV_ROWS), sm, st, p->pdl, c, list, pfx, (int)L.count, (int)nrhs));
            }
            break;
        }
        case K_PB_STEP: {
            if (nrhs == 1) {
                size_t sm = pstep_smem_bytes_mr(L.maxw, 1, false);
                if (lu) CK(launch_pdl(k_pb_step<true, 1>, dim3(L.nblocks, 1), dim3(SV_ROWS), sm, st, p->pdl, c, list, pfx, (int)L.count, d_rhs, (int64_t)ldrhs, (int)p->P.maxpw, p->d_counters, (int)nrhs));
                else CK(launch_pdl(k_pb_step<false, 1>, dim3(L.nblocks, 1), dim3(SV_ROWS), sm, st, p->pdl, c, list, pfx, (int)L.count, d_rhs, (int64_t)ldrhs, (int)p->P.maxpw, p->d_counters, (int)nrhs));
            } else {
                size_t sm = pstep_smem_bytes_mr(L.maxw, SOLVE_NR, false);
                if (lu) CK(launch_pdl(k_pb_step<true, SOLVE_NR>, dim3(L.nblocks, ngrp), dim3(SV_ROWS), sm, st, p->pdl, c, list, pfx, (int)L.count, d_rhs, (int64_t)ldrhs, (int)p->P.maxpw, p->d_counters, (int)nrhs));
                else CK(launch_pdl(k_pb_step<false, SOLVE_NR>, dim3(L.nblocks, ngrp), dim3(SV_ROWS), sm, st, p->pdl, c, list, pfx, (int)L.count, d_rhs, (int64_t)ldrhs, (int)p->P.maxpw, p->d_counters, (int)nrhs));
            }
            break;
        }
        case K_PB_UPDATE: {
            size_t sm = (size_t)(SV_ROWS / 32) * L.maxw * sizeof(double);
            if (lu) k_pb_update<true><<<grid, SV_ROWS, sm, st>>>(c, list, pfx, L.count, p->P.maxpw);
            else k_pb_update<false><<<grid, SV_ROWS, sm, st>>>(c, list, pfx, L.count, p->P.maxpw);
            break;
        }
        case K_PB_DIAG: {
            size_t sm = pstep_smem_bytes(L.maxw);
            if (lu) k_pb_diag<true><<<dim3(L.count, (unsigned)nrhs), 128, sm, st>>>(c, list, d_rhs, ldrhs, p->P.maxpw);
            else k_pb_diag<false><<<dim3(L.count, (unsigned)nrhs), 128, sm, st>>>(c, list, d_rhs, ldrhs, p->P.maxpw);
            break;
        }
        case K_BWD_GATHER: k_bwd_gather<<<grid, 256, 0, st>>>(c, list, pfx, L.count); break;
        case K_BWD_UPDATE:
            if (lu) k_bwd_update<true><<<grid, 256, 0, st>>>(c, list, pfx, L.count);
            else k_bwd_update<false><<<grid, 256, 0, st>>>(c, list, pfx, L.count);
            break;
        case K_BWD_DIAG:
            if (lu) k_bwd_diag<true><<<dim3(L.count, (unsigned)nrhs), 128, 0, st>>>(c, list, d_rhs, ldrhs);
            else k_bwd_diag<false><<<dim3(L.count, (unsigned)nrhs), 128, 0, st>>>(c, list, d_rhs, ldrhs);
            break;
        default: set_err("bad solve launch kind"); return -100;
        }
        ++p->launches_solve;
    }
    return 0;
}

// d_rhs: device, permuted order, column-major ld = ldrhs
static int64_t enqueue_solve(spk_plan* p, double* d_rhs, int64_t nrhs, int64_t ldrhs, int32_t which) {
    const int64_t maxbatch = 32;                       // bounds the work-vector arena
    cudaStream_t st = p->stream;
    for (int64_t r0 = 0; r0 < nrhs; r0 += maxbatch) {
        int64_t nb = std::min(maxbatch, nrhs - r0);
        DevCtx c = make_ctx(p);
        double* b = d_rhs + (size_t)r0 * ldrhs;
        const int nf = (int)p->P.fronts.size();
        int64_t rc = 0;
        if (p->flow_ints > 0) { CK(cudaMemsetAsync(p->d_flow, 0, (size_t)p->flow_ints * sizeof(int32_t), st)); CK(cudaMemsetAsync(p->d_box, 0xFF, (size_t)2 * p->P.n * p->w_nrhs * sizeof(double), st)); }   // tickets; mailboxes := sentinel
        if (which == 2) { k_copy_front_x<<<dim3(nf, (unsigned)nb), 64, 0, st>>>(c, nf, b, ldrhs, 0); ++p->launches_solve; }
        if (which == 0 || which == 1) { rc = run_solve_launches(p, c, p->P.fwd_launches, b, nb, ldrhs); if (rc) return rc; }
        if (which == 1) { k_copy_front_x<<<dim3(nf, (unsigned)nb), 64, 0, st>>>(c, nf, b, ldrhs, 1); ++p->launches_solve; }
        if (which == 0 || which == 2) { rc = run_solve_launches(p, c, p->P.bwd_launches, b, nb, ldrhs); if (rc) return rc; }
    }
    return 0;
}

// d_rhs: device, permuted order, column-major ld = ldrhs.  The launch sequence of a sweep is a pure
// function of the plan, so it is captured once into a CUDA graph per (buffer, nrhs, ld, which) and
// replayed: thousands of small dependent launches no longer pay the CPU launch path.
SPK_API int64_t spk_plan_solve_device(spk_plan* p, double* d_rhs, int64_t nrhs, int64_t ldrhs, int32_t which) {
    NEED_DEV(p);
    if (nrhs <= 0) return 0;
    if (!p->factored) { set_err("solve: the plan holds no factors (call spk_plan_factor or spk_plan_set_factors first)"); return -100; }
    cudaStream_t st = p->stream;
    int64_t rc = ensure_w(p, std::min<int64_t>(nrhs, 32)); if (rc) return rc;
    const bool use_graph = p->solve_graphs;
    if (use_graph && !(p->sg_exec && p->sg_rhs == d_rhs && p->sg_nrhs == nrhs && p->sg_ld == ldrhs && p->sg_which == which && p->sg_w == p->d_w)) {
        if (p->sg_exec) { cudaGraphExecDestroy(p->sg_exec); p->sg_exec = nullptr; }
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        p->launches_solve = 0;
        rc = enqueue_solve(p, d_rhs, nrhs, ldrhs, which);
        cudaError_t ce = cudaStreamEndCapture(st, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        CK(ce);
        CK(cudaGraphInstantiate(&p->sg_exec, g, 0));
        CK(cudaGraphDestroy(g));
        p->sg_rhs = d_rhs; p->sg_nrhs = nrhs; p->sg_ld = ldrhs; p->sg_which = which; p->sg_w = p->d_w;
        p->sg_launches = p->launches_solve;
    }
    CK(cudaEventRecord(p->ev0, st));
    if (use_graph) { CK(cudaGraphLaunch(p->sg_exec, st)); p->launches_solve = p->sg_launches; }
    else { p->launches_solve = 0; rc = enqueue_solve(p, d_rhs, nrhs, ldrhs, which); if (rc) return rc; }
    CK(cudaGetLastError());
    CK(cudaEventRecord(p->ev1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0; CK(cudaEventElapsedTime(&ms, p->ev0, p->ev1)); p->ms_solve = ms;
    return 0;
}

static int64_t ensure_rhs(spk_plan* p, int64_t nrhs) {
    int64_t need = p->P.n * nrhs;
    if (need > p->rhs_cap) {
        if (p->d_rhs) cudaFree(p->d_rhs);
        if (p->d_tmp) cudaFree(p->d_tmp);
        p->d_rhs = p->d_tmp = nullptr; p->rhs_cap = 0;
        CK(cudaMalloc((void**)&p->d_rhs, need * sizeof(double)));
        CK(cudaMalloc((void**)&p->d_tmp, need * sizeof(double)));
        p->rhs_cap = need;
    }
    return 0;
}

SPK_API int64_t spk_plan_solve(spk_plan* p, double* rhs, int64_t nrhs, int64_t ldrhs, int32_t which) {
    NEED_DEV(p);
    if (nrhs <= 0) return 0;
    const int64_t n = p->P.n;
    int64_t rc = ensure_rhs(p, nrhs); if (rc) return rc;
    CK(cudaMemcpy2DAsync(p->d_rhs, n * sizeof(double), rhs, ldrhs * sizeof(double), n * sizeof(double), nrhs,
                         cudaMemcpyHostToDevice, p->stream));
    rc = spk_plan_solve_device(p, p->d_rhs, nrhs, n, which); if (rc) return rc;
    CK(cudaMemcpy2DAsync(rhs, ldrhs * sizeof(double), p->d_rhs, n * sizeof(double), n * sizeof(double), nrhs,
                         cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}

SPK_API int64_t spk_plan_set_perm(spk_plan* p, const int64_t* rperm, const int64_t* rinvp) {
    NEED_DEV(p);
    const int64_t n = p->P.n;
    if (!p->d_rperm) { CK(cudaMalloc((void**)&p->d_rperm, n * sizeof(int64_t))); CK(cudaMalloc((void**)&p->d_rinvp, n * sizeof(int64_t))); }
    CK(cudaMemcpy(p->d_rperm, rperm, n * sizeof(int64_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->d_rinvp, rinvp, n * sizeof(int64_t), cudaMemcpyHostToDevice));
    p->have_perm = true;
    return 0;
}

SPK_API int64_t spk_plan_triangularsolve(spk_plan* p, double* b, int64_t nrhs, int64_t ldb) {
    NEED_DEV(p);
    if (!p->have_perm) { set_err("spk_plan_set_perm not called"); return -100; }
    if (nrhs <= 0) return 0;
    const int64_t n = p->P.n;
    int64_t rc = ensure_rhs(p, nrhs); if (rc) return rc;
    cudaStream_t st = p->stream;
    CK(cudaMemcpy2DAsync(p->d_tmp, n * sizeof(double), b, ldb * sizeof(double), n * sizeof(double), nrhs,
                         cudaMemcpyHostToDevice, st));
    k_perm_gather<<<dim3(cdiv(n, 256), (unsigned)nrhs), 256, 0, st>>>(n, p->d_rperm, p->d_tmp, p->d_rhs, n, n);
    rc = spk_plan_solve_device(p, p->d_rhs, nrhs, n, 0); if (rc) return rc;
    k_perm_gather<<<dim3(cdiv(n, 256), (unsigned)nrhs), 256, 0, st>>>(n, p->d_rinvp, p->d_rhs, p->d_tmp, n, n);
    CK(cudaMemcpy2DAsync(b, ldb * sizeof(double), p->d_tmp, n * sizeof(double), n * sizeof(double), nrhs,
                         cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

// Multi-part solve: phase 0 = forward sweep over this part's subtrees, phase 1 = forward + backward over the
// top set (after the caller exchanged the subtree-root work vectors), phase 2 = backward over the subtrees.
SPK_API int64_t spk_plan_solve_phase(spk_plan* p, double* d_rhs, int64_t nrhs, int64_t ldrhs, int32_t phase) {
    NEED_DEV(p);
    if (p->P.nparts <= 1) { set_err("single-part plan"); return -100; }
    if (!p->factored) { set_err("solve: the plan holds no factors"); return -100; }
    if (nrhs > 32) { set_err("multi-part solve: at most 32 right-hand sides per call"); return -100; }
    int64_t rc = ensure_w(p, 32); if (rc) return rc;
    DevCtx c = make_ctx(p);
    cudaStream_t st = p->stream;
    CK(cudaEventRecord(p->ev0, st));
    if (phase == 0 && p->flow_ints > 0) { CK(cudaMemsetAsync(p->d_flow, 0, (size_t)p->flow_ints * sizeof(int32_t), st)); CK(cudaMemsetAsync(p->d_box, 0xFF, (size_t)2 * p->P.n * p->w_nrhs * sizeof(double), st)); }
    if (phase == 0) { p->launches_solve = 0; rc = run_solve_launches(p, c, p->P.fwd_local, d_rhs, nrhs, ldrhs); }
    else if (phase == 1) { rc = run_solve_launches(p, c, p->P.fwd_top, d_rhs, nrhs, ldrhs); if (!rc) rc = run_solve_launches(p, c, p->P.bwd_top, d_rhs, nrhs, ldrhs); }
    else rc = run_solve_launches(p, c, p->P.bwd_local, d_rhs, nrhs, ldrhs);
    if (rc) return rc;
    CK(cudaGetLastError());
    CK(cudaEventRecord(p->ev1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0; CK(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
    p->ms_solve = (phase == 0 ? 0.0 : p->ms_solve) + ms;
    return 0;
}

// whole multi-part solve in one call: forward over the own subtrees, broadcast of the subtree-root work vectors,
// top set (replicated), backward over the own subtrees, broadcast of the owned pieces of x.  d_rhs: device,
// permuted order; on return every part holds the full solution.
SPK_API int64_t spk_plan_solve_multi(spk_plan* p, double* d_rhs, int64_t nrhs, int64_t ldrhs) {
    NEED_DEV(p);
    Plan& P = p->P;
    if (P.nparts <= 1) return spk_plan_solve_device(p, d_rhs, nrhs, ldrhs, 0);
    if (!p->comm) { set_err("multi-part plan without a communicator (spk_plan_comm_init)"); return -100; }
    if (!p->factored) { set_err("solve: the plan holds no factors"); return -100; }
    NcclApi* N = nccl_api();
    cudaStream_t st = p->stream;
    CK(cudaEventRecord(p->ev0, st));
    p->launches_solve = 0;
    for (int64_t r0 = 0; r0 < nrhs; r0 += 32) {
        const int64_t nb = std::min<int64_t>(32, nrhs - r0);
        int64_t rc = ensure_w(p, 32); if (rc) return rc;
        DevCtx c = make_ctx(p);
        double* b = d_rhs + (size_t)r0 * ldrhs;
        if (p->flow_ints > 0) { CK(cudaMemsetAsync(p->d_flow, 0, (size_t)p->flow_ints * sizeof(int32_t), st)); CK(cudaMemsetAsync(p->d_box, 0xFF, (size_t)2 * p->P.n * p->w_nrhs * sizeof(double), st)); }
        rc = run_solve_launches(p, c, P.fwd_local, b, nb, ldrhs); if (rc) return rc;
        NK(N->GroupStart());
        for (int32_t f : P.xchg) {
            const Front& F = P.fronts[f];
            for (int64_t q = 0; q < nb; ++q) { double* w = p->d_w + (size_t)q * P.wlen + F.wofs; NK(N->Broadcast(w, w, (size_t)F.R, ncclDouble, P.owner[f], p->comm, st)); }
        }
        NK(N->GroupEnd());
        rc = run_solve_launches(p, c, P.fwd_top, b, nb, ldrhs); if (rc) return rc;
        rc = run_solve_launches(p, c, P.bwd_top, b, nb, ldrhs); if (rc) return rc;
        rc = run_solve_launches(p, c, P.bwd_local, b, nb, ldrhs); if (rc) return rc;
        NK(N->GroupStart());
        for (const Plan::Range& g : P.ranges)
            for (int64_t q = 0; q < nb; ++q) { double* x = b + (size_t)q * ldrhs + g.col0; NK(N->Broadcast(x, x, (size_t)(g.col1 - g.col0), ncclDouble, g.owner, p->comm, st)); }
        NK(N->GroupEnd());
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(p->ev1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0; CK(cudaEventElapsedTime(&ms, p->ev0, p->ev1)); p->ms_solve = ms;
    return 0;
}

// exchange lists of a multi-part plan.  what = 0: subtree-root fronts, out = {owner, F offset, F length,
// w offset, w length (per right-hand side), front id};  what = 1: storage ranges owned by one part,
// out = {owner, lnz offset, lnz length, unz offset, unz length, first column, #columns}.  Returns the count
// when out == NULL.
SPK_API int64_t spk_plan_xchg_info(spk_plan* p, int32_t what, int64_t i, int64_t* out) {
    if (!p) return 0;
    const Plan& P = p->P;
    if (what == 0) {
        if (!out) return (int64_t)P.xchg.size();
        if (i < 0 || i >= (int64_t)P.xchg.size()) return -1;
        const Front& F = P.fronts[P.xchg[i]];
        out[0] = P.owner[P.xchg[i]]; out[1] = F.fofs; out[2] = (int64_t)F.ld * F.R; out[3] = F.wofs; out[4] = F.R; out[5] = P.xchg[i];
        return 0;
    }
    if (!out) return (int64_t)P.ranges.size();
    if (i < 0 || i >= (int64_t)P.ranges.size()) return -1;
    const Plan::Range& g = P.ranges[i];
    out[0] = g.owner; out[1] = g.lnz0; out[2] = g.lnz1 - g.lnz0; out[3] = g.unz0; out[4] = g.unz1 - g.unz0; out[5] = g.col0; out[6] = g.col1 - g.col0;
    return 0;
}


// ---- one process, N GPUs -----------------------------------------------------------------------------------
// spk_multi owns one plan per GPU (part r on device r) and the NCCL communicators (ncclCommInitAll); every call
// fans out to one host thread per GPU (the parts enqueue their launch lists and their side of the broadcasts
// concurrently) and joins before it returns, so a single ccall from the host application scales over the box.
struct spk_multi {
    std::vector<spk_plan*> plans;
    std::vector<ncclComm_t> comms;
    std::vector<double*> d_b;           // per device: right-hand sides (original order)
    int64_t bcap = 0;
};
} // extern "C"
#include <thread>
template <class Fn>
static int64_t multi_run(spk_multi* m, Fn fn) {
    const int N = (int)m->plans.size();
    std::vector<int64_t> rc(N, 0); std::vector<std::string> err(N);
    std::vector<std::thread> th;
    for (int r = 0; r < N; ++r) th.emplace_back([&, r] { rc[r] = fn(r, m->plans[r]); if (rc[r] <= -100) err[r] = g_err; });
    for (auto& t : th) t.join();
    int64_t worst = 0;
    for (int r = 0; r < N; ++r) if (rc[r] <= -100) { set_err("device " + std::to_string(r) + ": " + err[r]); return rc[r]; } else worst = std::min(worst, rc[r]);
    return worst;
}
extern "C" {
SPK_API void spk_multi_destroy(spk_multi* m) {
    if (!m) return;
    for (size_t r = 0; r < m->plans.size(); ++r) {
        if (m->plans[r]) { cudaSetDevice(m->plans[r]->device); if (r < m->d_b.size() && m->d_b[r]) cudaFree(m->d_b[r]); }
        if (r < m->comms.size() && m->comms[r] && nccl_api()) nccl_api()->CommDestroy(m->comms[r]);
        if (m->plans[r]) { m->plans[r]->comm = nullptr; spk_plan_destroy(m->plans[r]); }
    }
    delete m;
}
SPK_API spk_multi* spk_multi_create(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                                    const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz,
                                    const int64_t* xunz_or_null, int32_t ngpus) {
    if (ngpus < 1 || ngpus > spk_device_count()) { set_err("spk_multi_create: ngpus out of range"); return nullptr; }
    spk_multi* m = new spk_multi();
    m->plans.assign(ngpus, nullptr); m->comms.assign(ngpus, nullptr); m->d_b.assign(ngpus, nullptr);
    int64_t rc = multi_run(m, [&](int r, spk_plan*) -> int64_t {
        m->plans[r] = spk_plan_create(n, nsuper, xsuper, snode, xlindx, lindx, xlnz, xunz_or_null, r, r, ngpus);
        return m->plans[r] ? 0 : -100;
    });
    if (rc) { spk_multi_destroy(m); return nullptr; }
    if (ngpus > 1) {
        NcclApi* N = nccl_api();
        std::vector<int> devs(ngpus);
        for (int r = 0; r < ngpus; ++r) devs[r] = r;
        if (!N) { set_err("libnccl.so.2 not found"); spk_multi_destroy(m); return nullptr; }
        ncclResult_t e = N->CommInitAll(m->comms.data(), ngpus, devs.data());
        if (e != ncclSuccess) { set_err(std::string("ncclCommInitAll: ") + N->GetErrorString(e)); spk_multi_destroy(m); return nullptr; }
        for (int r = 0; r < ngpus; ++r) { m->plans[r]->comm = m->comms[r]; m->plans[r]->comm_owned = false; }
    }
    return m;
}
SPK_API spk_plan* spk_multi_plan(spk_multi* m, int32_t r) { return (m && r >= 0 && r < (int)m->plans.size()) ? m->plans[r] : nullptr; }
SPK_API int64_t spk_multi_inmatrix(spk_multi* m, int64_t nnz, const int64_t* dest_or_null, const double* nzval) {
    return multi_run(m, [&](int, spk_plan* p) { return spk_plan_inmatrix(p, nnz, dest_or_null, nzval); });
}
SPK_API int64_t spk_multi_set_values(spk_multi* m, const double* lnz, const double* unz_or_null) {
    return multi_run(m, [&](int, spk_plan* p) { return spk_plan_set_values(p, lnz, unz_or_null); });
}
SPK_API int64_t spk_multi_factor(spk_multi* m) {
    return multi_run(m, [&](int, spk_plan* p) { return spk_plan_factor_multi(p); });
}
// factors in the reference layout: every subtree's storage range from its owner, the top set from part 0
SPK_API int64_t spk_multi_get_factors(spk_multi* m, double* lnz, double* unz, int64_t* ipvt) {
    spk_plan* p0 = m->plans[0];
    if (m->plans.size() == 1) return spk_plan_get_factors(p0, lnz, unz, ipvt);
    int64_t rc = spk_plan_get_factors(p0, lnz, unz, ipvt);            // top set + part 0's subtrees (+ stale ranges, overwritten below)
    if (rc) return rc;
    const Plan& P = p0->P;
    std::vector<int64_t> ip;
    for (const Plan::Range& g : P.ranges) {
        if (g.owner == 0) continue;
        spk_plan* q = m->plans[g.owner];
        CK(cudaSetDevice(q->device));
        if (lnz) CK(cudaMemcpy(lnz + g.lnz0, q->d_lnz + g.lnz0, (g.lnz1 - g.lnz0) * sizeof(double), cudaMemcpyDeviceToHost));
        if (unz && P.lu && g.unz1 > g.unz0) CK(cudaMemcpy(unz + g.unz0, q->d_unz + g.unz0, (g.unz1 - g.unz0) * sizeof(double), cudaMemcpyDeviceToHost));
        if (ipvt && P.lu) {
            std::vector<int32_t> t((size_t)(g.col1 - g.col0));
            CK(cudaMemcpy(t.data(), q->d_ipiv + g.col0, t.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < t.size(); ++i) ipvt[g.col0 + (int64_t)i] = t[i];
        }
    }
    return 0;
}
SPK_API int64_t spk_multi_set_perm(spk_multi* m, const int64_t* rperm, const int64_t* rinvp) {
    return multi_run(m, [&](int, spk_plan* p) { return spk_plan_set_perm(p, rperm, rinvp); });
}
// _triangularsolve! over all GPUs: b (host, original order, ld = ldb) in place
SPK_API int64_t spk_multi_triangularsolve(spk_multi* m, double* b, int64_t nrhs, int64_t ldb) {
    if (m->plans.size() == 1) return spk_plan_triangularsolve(m->plans[0], b, nrhs, ldb);
    if (nrhs <= 0) return 0;
    return multi_run(m, [&](int r, spk_plan* p) -> int64_t {
        NEED_DEV(p);
        if (!p->have_perm) { set_err("spk_multi_set_perm not called"); return -100; }
        const int64_t n = p->P.n;
        int64_t rc = ensure_rhs(p, nrhs); if (rc) return rc;
        cudaStream_t st = p->stream;
        CK(cudaMemcpy2DAsync(p->d_tmp, n * sizeof(double), b, ldb * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyHostToDevice, st));
        k_perm_gather<<<dim3(cdiv(n, 256), (unsigned)nrhs), 256, 0, st>>>(n, p->d_rperm, p->d_tmp, p->d_rhs, n, n);
        rc = spk_plan_solve_multi(p, p->d_rhs, nrhs, n); if (rc) return rc;
        if (r == 0) {
            k_perm_gather<<<dim3(cdiv(n, 256), (unsigned)nrhs), 256, 0, st>>>(n, p->d_rinvp, p->d_rhs, p->d_tmp, n, n);
            CK(cudaMemcpy2DAsync(b, ldb * sizeof(double), p->d_tmp, n * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        return 0;
    });
}

// ---- introspection ----------------------------------------------------------------------
SPK_API void* spk_plan_device_ptr(spk_plan* p, int32_t what) {
    if (!p) return nullptr;
    switch (what) { case 0: return p->d_lnz; case 1: return p->d_unz; case 2: return p->d_ipiv; case 5: return p->d_F; case 6: return p->d_w; default: return nullptr; }
}
SPK_API int64_t spk_plan_device_len(spk_plan* p, int32_t what) {
    if (!p) return 0;
    switch (what) { case 0: return p->P.nlnz; case 1: return p->P.nunz; case 2: return p->P.n; case 5: return p->P.arena; case 6: return p->P.wlen * p->w_nrhs; default: return 0; }
}
SPK_API int64_t spk_plan_stat(spk_plan* p, int32_t what) {
    if (!p) return 0;
    switch (what) {
    case 0: return p->launches_factor;
    case 1: return p->launches_solve;
    case 2: return (int64_t)p->P.fronts.size();
    case 3: return p->P.nlevels;
    case 4: return (int64_t)p->dev_bytes;
    case 5: { int64_t k = 0; for (const Front& f : p->P.fronts) if (f.nch > 1) ++k; return k; }
    case 6: return p->P.arena;
    case 7: return (int64_t)p->P.factor_launches.size();
    case 8: return (int64_t)(p->P.fwd_launches.size() + p->P.bwd_launches.size());
    case 9: return p->P.wlen;
    case 10: return p->P.maxpw;
    case 11: return p->P.maxR;
    case 12: return p->P.nparts;
    case 13: { int64_t k = 0; for (int32_t o : p->P.owner) if (o == -1) ++k; return k; }
    case 15: return (int64_t)p->P.ranges.size();                 // elimination subtrees dealt to the parts
    case 14: { int64_t k = 0; for (const Front& f : p->P.fronts) k = std::max<int64_t>(k, (f.nps + p->P.ob_steps - 1) / p->P.ob_steps); return k; }   // outer blocks of the widest front
    case 100: p->profile = true; return 0;
    case 101: p->profile = false; return 0;
    default: return 0;
    }
}
SPK_API double spk_plan_statf(spk_plan* p, int32_t what) {
    if (!p) return 0;
    switch (what) {
    case 0: return p->P.flops_struct;
    case 1: return p->P.nnzL;
    case 2: return p->ms_factor;
    case 3: return p->ms_solve;
    case 4: return p->gemm_flops;
    case 5: return p->gemm_ms;
    case 6: return p->ms_phase0;
    case 7: return p->ms_phase1;
    default:
        if (what >= 10 && what < 26) return p->kind_ms[what - 10];
        if (what >= 30 && what < 46) return (double)p->kind_n[what - 30];
        return 0;
    }
}

// ---- stateless drop-ins -------------------------------------------------------------------
// The reference's own call sites pass flat arrays only (SpkSparseBase.jl:384,409-411; SpkSparseSpdBase.jl:325,351),
// so these entry points have no handle to keep.  Behind them sits a small PLAN CACHE keyed on the structure
// (n, nsuper, LU / LDL^T, hash of xsuper / xlindx / lindx / xlnz): `_triangularsolve!` calls `_lulsolve!` then
// `_luusolve!` for every right-hand side, and neither should re-analyse the structure, reallocate and zero tens
// of GB of frontal storage, or — when the arrays are untouched since the factorisation that produced them —
// upload the factors again.  A cached plan remembers the host address and a sampled fingerprint of the lnz / unz /
// ipiv it wrote back; a solve that presents the same arrays uses the resident factors, anything else is uploaded.
// SPK_PLAN_CACHE=0 disables the cache (every call builds and destroys its plan), SPK_PLAN_CACHE=<k> keeps k plans
// (default 2); spk_cache_clear() frees them.
static int64_t n_from_xsuper(int64_t nsuper, const int64_t* xsuper) { return xsuper[nsuper] - 1; }

struct CacheEntry {
    uint64_t key = 0; int64_t n = 0, nsuper = 0; bool lu = false;
    spk_plan* plan = nullptr;
    const void* h_lnz = nullptr; const void* h_unz = nullptr; const void* h_ipiv = nullptr;
    uint64_t fp[3] = {0, 0, 0};                     // sampled fingerprints of lnz / unz / ipiv as written back
    bool resident = false; bool f32 = false;
    uint64_t stamp = 0;
};
static std::mutex g_cache_mu;
static std::vector<CacheEntry> g_cache;
static uint64_t g_cache_clock = 0;

static uint64_t fnv(const void* data, size_t bytes, uint64_t h) {
    const uint64_t* w = (const uint64_t*)data;
    for (size_t i = 0; i < bytes / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    return h;
}
static uint64_t structure_key(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz, bool lu) {
    uint64_t h = 1469598103934665603ull ^ (uint64_t)n * 31 ^ (uint64_t)nsuper * 131 ^ (lu ? 0x9e3779b97f4a7c15ull : 0);
    h = fnv(xsuper, (size_t)(nsuper + 1) * 8, h);
    h = fnv(xlindx, (size_t)(nsuper + 1) * 8, h);
    h = fnv(lindx, (size_t)(xlindx[nsuper] - 1) * 8, h);
    h = fnv(xlnz, (size_t)(n + 1) * 8, h);
    return h;
}
// sampled fingerprint of one factor array (4096 evenly spaced entries + the last one); es = bytes per value
static uint64_t array_fingerprint(const void* a, int64_t len, int es) {
    uint64_t h = 1469598103934665603ull;
    if (!a || len <= 0) return h;
    const int64_t step = std::max<int64_t>(1, len / 4096);
    for (int64_t i = 0; i < len; i += step) { uint64_t v = 0; memcpy(&v, (const char*)a + i * es, es); h ^= v + (uint64_t)i; h *= 1099511628211ull; }
    uint64_t v = 0; memcpy(&v, (const char*)a + (len - 1) * es, es); h ^= v; h *= 1099511628211ull;
    return h;
}
static int cache_capacity() {
    static int cap = -1;
    if (cap < 0) { const char* e = getenv("SPK_PLAN_CACHE"); cap = e ? std::max(0, atoi(e)) : 2; }
    return cap;
}
// returns the cached plan for this structure (building it on a miss); *slot = its cache entry or nullptr when uncached
static spk_plan* cached_plan(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode, const int64_t* xlindx,
                             const int64_t* lindx, const int64_t* xlnz, const int64_t* xunz, bool lu, CacheEntry** slot) {
    *slot = nullptr;
    std::vector<int64_t> sn, xu;
    if (!snode) { sn.resize(n); for (int64_t s = 0; s < nsuper; ++s) for (int64_t j = xsuper[s]; j < xsuper[s + 1]; ++j) sn[j - 1] = s + 1; snode = sn.data(); }
    if (lu && !xunz) {                                   // `_lulsolve!` gets no xunz: rebuild it from the structure
        xu.resize(n + 1);
        int64_t up = 1;
        for (int64_t s = 0; s < nsuper; ++s) {
            int64_t w = xsuper[s + 1] - xsuper[s], len = xlindx[s + 1] - xlindx[s];
            for (int64_t j = xsuper[s]; j < xsuper[s + 1]; ++j) { xu[j - 1] = up; up += len - w; }
        }
        xu[n] = up; xunz = xu.data();
    }
    if (cache_capacity() == 0) return spk_plan_create(n, nsuper, xsuper, snode, xlindx, lindx, xlnz, lu ? xunz : nullptr, 0, 0, 1);
    const uint64_t key = structure_key(n, nsuper, xsuper, xlindx, lindx, xlnz, lu);
    for (CacheEntry& e : g_cache)
        if (e.key == key && e.n == n && e.nsuper == nsuper && e.lu == lu) { e.stamp = ++g_cache_clock; *slot = &e; return e.plan; }
    if ((int)g_cache.size() >= cache_capacity()) {       // evict the least recently used plan BEFORE allocating the new arena
        size_t old = 0;
        for (size_t i = 1; i < g_cache.size(); ++i) if (g_cache[i].stamp < g_cache[old].stamp) old = i;
        spk_plan_destroy(g_cache[old].plan);
        g_cache.erase(g_cache.begin() + old);
    }
    spk_plan* p = spk_plan_create(n, nsuper, xsuper, snode, xlindx, lindx, xlnz, lu ? xunz : nullptr, 0, 0, 1);
    if (!p) return nullptr;
    CacheEntry e; e.key = key; e.n = n; e.nsuper = nsuper; e.lu = lu; e.plan = p; e.stamp = ++g_cache_clock;
    g_cache.push_back(e);
    *slot = &g_cache.back();
    return p;
}
static void release_plan(spk_plan* p, CacheEntry* slot) { if (!slot && p) spk_plan_destroy(p); }

SPK_API void spk_cache_clear(void) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (CacheEntry& e : g_cache) spk_plan_destroy(e.plan);
    g_cache.clear();
}

// device-side Float32 <-> Float64 conversion of the value arrays (the _f32 twins move half the bytes over the bus)
__global__ void k_widen_f32(int64_t len, const float* __restrict__ in, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}
__global__ void k_narrow_f32(int64_t len, const double* __restrict__ in, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) out[i] = (float)in[i];
}
static int64_t upload_values(spk_plan* p, double* d_dst, const void* h_src, int64_t len, bool f32) {
    if (len <= 0) return 0;
    if (!f32) { CK(cudaMemcpyAsync(d_dst, h_src, len * sizeof(double), cudaMemcpyHostToDevice, p->stream)); return 0; }
    float* tmp = nullptr;
    CK(cudaMalloc((void**)&tmp, len * sizeof(float)));
    cudaError_t e = cudaMemcpyAsync(tmp, h_src, len * sizeof(float), cudaMemcpyHostToDevice, p->stream);
    if (e == cudaSuccess) { k_widen_f32<<<1184, 256, 0, p->stream>>>(len, tmp, d_dst); e = cudaStreamSynchronize(p->stream); }
    cudaFree(tmp);
    CK(e);
    return 0;
}
static int64_t download_values(spk_plan* p, void* h_dst, const double* d_src, int64_t len, bool f32) {
    if (len <= 0) return 0;
    if (!f32) { CK(cudaMemcpyAsync(h_dst, d_src, len * sizeof(double), cudaMemcpyDeviceToHost, p->stream)); CK(cudaStreamSynchronize(p->stream)); return 0; }
    float* tmp = nullptr;
    CK(cudaMalloc((void**)&tmp, len * sizeof(float)));
    k_narrow_f32<<<1184, 256, 0, p->stream>>>(len, d_src, tmp);
    cudaError_t e = cudaMemcpyAsync(h_dst, tmp, len * sizeof(float), cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    cudaFree(tmp);
    CK(e);
    return 0;
}

// factor: values up (FP64 or FP32), factorisation, factors back in place; the plan stays cached with the factors resident
static int64_t dropin_factor(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode, const int64_t* xlindx,
                             const int64_t* lindx, const int64_t* xlnz, void* lnz, const int64_t* xunz, void* unz, int64_t* ipvt, bool lu, bool f32) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    CacheEntry* slot = nullptr;
    spk_plan* p = cached_plan(n, nsuper, xsuper, snode, xlindx, lindx, xlnz, xunz, lu, &slot);
    if (!p) return -100;
    if (slot) slot->resident = false;
    int64_t rc = cudaSetDevice(p->device) == cudaSuccess ? 0 : -100;
    if (!rc) rc = upload_values(p, p->d_lnz, lnz, p->P.nlnz, f32);
    if (!rc && lu) rc = upload_values(p, p->d_unz, unz, p->P.nunz, f32);
    if (!rc && cudaStreamSynchronize(p->stream) != cudaSuccess) rc = -100;
    int64_t flag = 0;
    if (!rc) { p->factored = false; p->values_in_fronts = false; flag = spk_plan_factor(p); if (flag < -1) rc = flag; }
    if (!rc) rc = download_values(p, lnz, p->d_lnz, p->P.nlnz, f32);
    if (!rc && lu) rc = download_values(p, unz, p->d_unz, p->P.nunz, f32);
    if (!rc && lu) rc = spk_plan_get_factors(p, nullptr, nullptr, ipvt);
    if (!rc && slot) {
        slot->h_lnz = lnz; slot->h_unz = lu ? unz : nullptr; slot->h_ipiv = lu ? ipvt : nullptr; slot->f32 = f32;
        const int es = f32 ? 4 : 8;
        slot->fp[0] = array_fingerprint(lnz, p->P.nlnz, es); slot->fp[1] = array_fingerprint(slot->h_unz, p->P.nunz, es); slot->fp[2] = array_fingerprint(slot->h_ipiv, n, 8);
        slot->resident = true;
    }
    release_plan(p, slot);
    return rc ? rc : flag;
}

// solve: which = 0 both sweeps (LDL^T), 1 forward (_lulsolve!), 2 backward (_luusolve!)
static int64_t dropin_solve(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx, const int64_t* lindx,
                            const int64_t* xlnz, const void* lnz, const int64_t* xunz, const void* unz, const int64_t* ipiv,
                            void* rhs, bool lu, int which, bool f32) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    CacheEntry* slot = nullptr;
    spk_plan* p = cached_plan(n, nsuper, xsuper, nullptr, xlindx, lindx, xlnz, xunz, lu, &slot);
    if (!p) return -100;
    int64_t rc = cudaSetDevice(p->device) == cudaSuccess ? 0 : -100;
    // the forward sweep needs lnz + ipiv, the backward sweep lnz + unz; arrays not passed by the caller count as unchanged
    bool resident = slot && slot->resident && p->factored && slot->f32 == f32 && slot->h_lnz == lnz &&
                    (!unz || slot->h_unz == unz) && (!ipiv || slot->h_ipiv == ipiv);
    const int es = f32 ? 4 : 8;                          // only the arrays handed to THIS call are looked at (the others may be gone)
    if (resident) resident = slot->fp[0] == array_fingerprint(lnz, p->P.nlnz, es) && (!unz || slot->fp[1] == array_fingerprint(unz, p->P.nunz, es)) &&
                             (!ipiv || slot->fp[2] == array_fingerprint(ipiv, n, 8));
    if (!rc && !resident) {
        if (slot) slot->resident = false;
        rc = upload_values(p, p->d_lnz, lnz, p->P.nlnz, f32);
        if (!rc && lu) {
            if (unz) rc = upload_values(p, p->d_unz, unz, p->P.nunz, f32);
            else if (cudaMemsetAsync(p->d_unz, 0, std::max<int64_t>(p->P.nunz, 1) * sizeof(double), p->stream) != cudaSuccess) rc = -100;
        }
        if (!rc && lu && ipiv) {
            std::vector<int32_t> ip((size_t)n);
            for (int64_t i = 0; i < n; ++i) ip[i] = (int32_t)ipiv[i];
            if (cudaMemcpy(p->d_ipiv, ip.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) rc = -100;
        }
        if (!rc) {   // the sweeps read the frontal matrices: rebuild them from the uploaded factors
            DevCtx c = make_ctx(p);
            if (cudaMemsetAsync(p->d_F, 0, p->P.arena * sizeof(double), p->stream) != cudaSuccess) rc = -100;
            if (!rc) {
                if (lu) k_chunks<false><<<p->chunk_blocks, 256, 0, p->stream>>>(c, p->d_chunkpfx, (int)p->P.chunks.size());
                else k_chunks<false, true><<<p->chunk_blocks, 256, 0, p->stream>>>(c, p->d_chunkpfx, (int)p->P.chunks.size());   // + U = D L^T
                if (build_inverses(p, p->stream) != 0) rc = -100;
                if (cudaStreamSynchronize(p->stream) != cudaSuccess) rc = -100;
            }
        }
        if (!rc) {
            p->factored = true;
            // a forward-only call leaves unz / a backward-only call leaves ipiv unset: not a complete resident copy
            if (slot && (!lu || (unz && ipiv))) {
                slot->h_lnz = lnz; slot->h_unz = unz; slot->h_ipiv = ipiv; slot->f32 = f32;
                slot->fp[0] = array_fingerprint(lnz, p->P.nlnz, es); slot->fp[1] = array_fingerprint(unz, p->P.nunz, es); slot->fp[2] = array_fingerprint(ipiv, n, 8);
                slot->resident = true;
            }
        }
    }
    if (!rc) {
        if (!f32) rc = spk_plan_solve(p, (double*)rhs, 1, n, which);
        else {
            std::vector<double> r((size_t)n);
            for (int64_t i = 0; i < n; ++i) r[i] = ((float*)rhs)[i];
            rc = spk_plan_solve(p, r.data(), 1, n, which);
            if (!rc) for (int64_t i = 0; i < n; ++i) ((float*)rhs)[i] = (float)r[i];
        }
    }
    release_plan(p, slot);
    return rc ? rc : 1;
}

SPK_API int64_t spk_lufactor_f64(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                                 const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz, double* lnz,
                                 const int64_t* xunz, double* unz, int64_t* ipvt) {
    return dropin_factor(n, nsuper, xsuper, snode, xlindx, lindx, xlnz, lnz, xunz, unz, ipvt, true, false);
}
SPK_API int64_t spk_ldltfactor_f64(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                                   const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz, double* lnz) {
    return dropin_factor(n, nsuper, xsuper, snode, xlindx, lindx, xlnz, lnz, nullptr, nullptr, nullptr, false, false);
}
SPK_API int64_t spk_lulsolve_f64(int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx, const int64_t* lindx,
                                 const int64_t* xlnz, const double* lnz, const int64_t* ipiv, double* rhs) {
    return dropin_solve(n_from_xsuper(nsuper, xsuper), nsuper, xsuper, xlindx, lindx, xlnz, lnz, nullptr, nullptr, ipiv, rhs, true, 1, false);
}
SPK_API int64_t spk_luusolve_f64(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx,
                                 const int64_t* lindx, const int64_t* xlnz, const double* lnz, const int64_t* xunz,
                                 const double* unz, double* rhs) {
    return dropin_solve(n, nsuper, xsuper, xlindx, lindx, xlnz, lnz, xunz, unz, nullptr, rhs, true, 2, false);
}
SPK_API int64_t spk_ldltsolve_f64(int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx, const int64_t* lindx,
                                  const int64_t* xlnz, const double* lnz, double* rhs) {
    return dropin_solve(n_from_xsuper(nsuper, xsuper), nsuper, xsuper, xlindx, lindx, xlnz, lnz, nullptr, nullptr, nullptr, rhs, false, 0, false);
}

// ---- residual and iterative refinement on the device --------------------
// A as the reference holds it: SparseMatrixCSC colptr / rowval / nzval, 1-based, ORIGINAL ordering.
SPK_API int64_t spk_plan_set_matrix(spk_plan* p, int64_t nnz, const int64_t* colptr, const int64_t* rowval, const double* nzval) {
    NEED_DEV(p);
    const int64_t n = p->P.n;
    if (nnz < 0 || colptr[n] - 1 != nnz) { set_err("spk_plan_set_matrix: colptr[n+1]-1 != nnz"); return -100; }
    std::vector<int64_t> rp(n + 1, 0);
    for (int64_t k = 0; k < nnz; ++k) { const int64_t i = rowval[k] - 1; if (i < 0 || i >= n) { set_err("spk_plan_set_matrix: row index out of range"); return -100; } rp[i + 1]++; }
    for (int64_t i = 0; i < n; ++i) rp[i + 1] += rp[i];
    std::vector<int32_t> ci((size_t)nnz); std::vector<double> av((size_t)nnz);
    std::vector<int64_t> at(rp.begin(), rp.end() - 1);
    for (int64_t j = 0; j < n; ++j)                       // columns ascending => every row's entries end up sorted by column
        for (int64_t k = colptr[j] - 1; k < colptr[j + 1] - 1; ++k) { const int64_t q = at[rowval[k] - 1]++; ci[q] = (int32_t)j; av[q] = nzval[k]; }
    for (void* q : {(void*)p->d_arp, (void*)p->d_aci, (void*)p->d_av}) if (q) cudaFree(q);
    p->d_arp = nullptr; p->d_aci = nullptr; p->d_av = nullptr; p->annz = -1;
    CK(upload(&p->d_arp, rp)); CK(upload(&p->d_aci, ci)); CK(upload(&p->d_av, av));
    p->annz = nnz;
    return 0;
}

static int64_t ensure_refine(spk_plan* p, int64_t nrhs) {
    const int64_t need = p->P.n * nrhs;
    if (need > p->rcap) {
        for (void* q : {(void*)p->d_rb, (void*)p->d_rx, (void*)p->d_rr, (void*)p->d_rpart}) if (q) cudaFree(q);
        p->d_rb = p->d_rx = p->d_rr = p->d_rpart = nullptr; p->rcap = 0;
        CK(cudaMalloc((void**)&p->d_rb, need * sizeof(double)));
        CK(cudaMalloc((void**)&p->d_rx, need * sizeof(double)));
        CK(cudaMalloc((void**)&p->d_rr, need * sizeof(double)));
        CK(cudaMalloc((void**)&p->d_rpart, (size_t)(cdiv(p->P.n, 4096) + 1) * nrhs * sizeof(double)));
        p->rcap = need;
    }
    return 0;
}
// ||v_q||_2 for q < nrhs (device vector, leading dimension n)
static int64_t dev_norms(spk_plan* p, const double* d_v, int64_t nrhs, double* out) {
    const int64_t n = p->P.n; const int nb = (int)cdiv(n, 4096);
    k_sumsq_partial<<<dim3(nb, (unsigned)nrhs), 256, 0, p->stream>>>(n, d_v, n, p->d_rpart);
    std::vector<double> part((size_t)nb * nrhs);
    CK(cudaMemcpyAsync(part.data(), p->d_rpart, part.size() * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    for (int64_t q = 0; q < nrhs; ++q) { double s2 = 0.0; for (int b2 = 0; b2 < nb; ++b2) s2 += part[(size_t)q * nb + b2]; out[q] = std::sqrt(s2); }
    return 0;
}
static int64_t refine_chunk(spk_plan* p, const double* b, double* x, int64_t nrhs, int64_t ld, int32_t maxit, double tol,
                            double* res_or_null, double* relnorm, bool correct) {
    const int64_t n = p->P.n;
    cudaStream_t st = p->stream;
    int64_t rc = ensure_refine(p, nrhs); if (rc) return rc;
    CK(cudaMemcpy2DAsync(p->d_rb, n * sizeof(double), b, ld * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpy2DAsync(p->d_rx, n * sizeof(double), x, ld * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyHostToDevice, st));
    std::vector<double> nb(nrhs), nr(nrhs);
    rc = dev_norms(p, p->d_rb, nrhs, nb.data()); if (rc) return rc;
    int64_t steps = 0;
    for (int32_t it = 0;; ++it) {
        k_csr_residual<<<dim3((unsigned)cdiv(n, 256), (unsigned)nrhs), 256, 0, st>>>(n, p->d_arp, p->d_aci, p->d_av, p->d_rb, p->d_rx, p->d_rr, n);
        rc = dev_norms(p, p->d_rr, nrhs, nr.data()); if (rc) return rc;
        double worst = 0.0;
        for (int64_t q = 0; q < nrhs; ++q) { relnorm[q] = nb[q] > 0.0 ? nr[q] / nb[q] : nr[q]; worst = std::max(worst, relnorm[q]); }
        if (!correct || it >= maxit || worst <= tol) break;
        // d = A^{-1} r with the resident factors (permute, sweeps, un-permute), x += d
        rc = ensure_rhs(p, nrhs); if (rc) return rc;
        k_perm_gather<<<dim3(cdiv(n, 256), (unsigned)nrhs), 256, 0, st>>>(n, p->d_rperm, p->d_rr, p->d_rhs, n, n);
        rc = spk_plan_solve_device(p, p->d_rhs, nrhs, n, 0); if (rc) return rc;
        k_perm_gather<<<dim3(cdiv(n, 256), (unsigned)nrhs), 256, 0, st>>>(n, p->d_rinvp, p->d_rhs, p->d_tmp, n, n);
        k_add_inplace<<<dim3((unsigned)cdiv(n, 256), (unsigned)nrhs), 256, 0, st>>>(n, p->d_rx, p->d_tmp, n);
        ++steps;
    }
    if (res_or_null) CK(cudaMemcpy2DAsync(res_or_null, ld * sizeof(double), p->d_rr, n * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyDeviceToHost, st));
    if (correct) CK(cudaMemcpy2DAsync(x, ld * sizeof(double), p->d_rx, n * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return steps;
}
// res = b - A x (res may be NULL), relnorm[q] = ||res_q|| / ||b_q||
SPK_API int64_t spk_plan_residual(spk_plan* p, const double* b, const double* x, int64_t nrhs, int64_t ld, double* res_or_null, double* relnorm) {
    NEED_DEV(p);
    if (p->annz < 0) { set_err("spk_plan_set_matrix not called"); return -100; }
    for (int64_t q0 = 0; q0 < nrhs; q0 += 32) {
        const int64_t nq = std::min<int64_t>(32, nrhs - q0);
        int64_t rc = refine_chunk(p, b + (size_t)q0 * ld, const_cast<double*>(x) + (size_t)q0 * ld, nq, ld, 0, 0.0,
                                  res_or_null ? res_or_null + (size_t)q0 * ld : nullptr, relnorm + q0, false);
        if (rc < 0) return rc;
    }
    return 0;
}
// iterative refinement in FP64: x += A^{-1}(b - A x) until max_q relnorm[q] <= tol or maxit corrections; returns
// the largest number of corrections applied to a block of right-hand sides (>= 0) or an error code (<= -100)
SPK_API int64_t spk_plan_refine(spk_plan* p, const double* b, double* x, int64_t nrhs, int64_t ld, int32_t maxit, double tol, double* relnorm) {
    NEED_DEV(p);
    if (p->annz < 0) { set_err("spk_plan_set_matrix not called"); return -100; }
    if (!p->have_perm) { set_err("spk_plan_set_perm not called"); return -100; }
    if (!p->factored) { set_err("spk_plan_refine: no factors"); return -100; }
    int64_t most = 0;
    for (int64_t q0 = 0; q0 < nrhs; q0 += 32) {
        const int64_t nq = std::min<int64_t>(32, nrhs - q0);
        int64_t rc = refine_chunk(p, b + (size_t)q0 * ld, x + (size_t)q0 * ld, nq, ld, maxit, tol, nullptr, relnorm + q0, true);
        if (rc < 0) return rc;
        most = std::max(most, rc);
    }
    return most;
}


// ---- 1-norm condition estimate with the resident factors ---------------------------------
// The reference keeps its estimator only as commented-out Fortran (SpkSparseSpdSolver.jl:267-459).  This is the
// Hager / Higham estimator (the algorithm of LAPACK's xLACON) of ||inv(A)||_1, driven by triangular solves with the
// factors held on the device, times ||A||_1 computed from the device copy of A (spk_plan_set_matrix).
//   LDL^T plans: A is symmetric, so the inv(A)^T products Hager's iteration needs are inv(A) products: the full
//                algorithm, with Higham's alternating-sign safeguard vector.
//   LU plans:    the engine has no transposed sweeps; the estimate is the largest ||inv(A) v||_1 / ||v||_1 over the
//                safeguard vector, the uniform vector and four fixed random sign vectors — a LOWER BOUND
//                (info[1] = 1 says so).
// Returns cond_1 estimate (>= 1) or a negative error code; out[0] = ||A||_1, out[1] = estimate of ||inv(A)||_1.
__global__ void k_csr_colabs(int64_t n, const int64_t* __restrict__ rp, const int32_t* __restrict__ ci, const double* __restrict__ av, double* __restrict__ colsum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int64_t k = rp[i]; k < rp[i + 1]; ++k) atomicAdd(colsum + ci[k], fabs(av[k]));
}
SPK_API double spk_plan_condest(spk_plan* p, double* out2, int32_t* info2) {
    if (!p || p->device < 0) { set_err("plan has no device"); return -100.0; }
    if (cudaSetDevice(p->device) != cudaSuccess) return -100.0;
    if (p->annz < 0) { set_err("spk_plan_condest: spk_plan_set_matrix not called"); return -100.0; }
    if (!p->have_perm) { set_err("spk_plan_condest: spk_plan_set_perm not called"); return -100.0; }
    if (!p->factored) { set_err("spk_plan_condest: no factors"); return -100.0; }
    const int64_t n = p->P.n;
    // ||A||_1 = max column sum of |a_ij|
    double anorm = 0.0;
    {
        double* d_cs = nullptr;
        if (cudaMalloc((void**)&d_cs, (size_t)n * sizeof(double)) != cudaSuccess) { set_err("cudaMalloc"); return -100.0; }
        cudaMemsetAsync(d_cs, 0, (size_t)n * sizeof(double), p->stream);
        k_csr_colabs<<<cdiv(n, 256), 256, 0, p->stream>>>(n, p->d_arp, p->d_aci, p->d_av, d_cs);
        std::vector<double> cs((size_t)n);
        cudaError_t e = cudaMemcpyAsync(cs.data(), d_cs, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, p->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        cudaFree(d_cs);
        if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return -100.0; }
        for (double v : cs) anorm = std::max(anorm, v);
    }
    auto solve = [&](std::vector<double>& v) -> int64_t { return spk_plan_triangularsolve(p, v.data(), 1, n); };
    auto norm1 = [&](const std::vector<double>& v) { double s1 = 0.0; for (double x : v) s1 += std::fabs(x); return s1; };
    std::vector<double> x((size_t)n), y, z;
    double est = 0.0; int iters = 0, lower_bound = 0;
    if (!p->P.lu) {
        for (int64_t i = 0; i < n; ++i) x[i] = 1.0 / (double)n;
        int64_t jlast = -1;
        for (int it = 0; it < 5; ++it) {
            y = x; if (solve(y) < 0) return -100.0; ++iters;
            const double e1 = norm1(y);
            if (it > 0 && e1 <= est) break;                        // no increase: converged
            est = e1;
            z.resize((size_t)n);
            for (int64_t i = 0; i < n; ++i) z[i] = y[i] >= 0.0 ? 1.0 : -1.0;
            if (solve(z) < 0) return -100.0; ++iters;              // inv(A)^T sign(y) = inv(A) sign(y)
            int64_t j = 0; double zmax = 0.0, ztx = 0.0;
            for (int64_t i = 0; i < n; ++i) { if (std::fabs(z[i]) > zmax) { zmax = std::fabs(z[i]); j = i; } ztx += z[i] * x[i]; }
            if (zmax <= ztx || j == jlast) break;
            jlast = j;
            std::fill(x.begin(), x.end(), 0.0); x[j] = 1.0;
        }
        // Higham's safeguard: x_i = (-1)^i (1 + i/(n-1)),  estimate 2 ||inv(A) x||_1 / (3 n)
        for (int64_t i = 0; i < n; ++i) x[i] = ((i & 1) ? -1.0 : 1.0) * (1.0 + (n > 1 ? (double)i / (double)(n - 1) : 0.0));
        y = x; if (solve(y) < 0) return -100.0; ++iters;
        est = std::max(est, 2.0 * norm1(y) / (3.0 * (double)n));
    } else {
        lower_bound = 1;
        uint64_t rng = 0x9876ull;
        for (int probe = 0; probe < 6; ++probe) {
            for (int64_t i = 0; i < n; ++i) {
                if (probe == 0) x[i] = 1.0 / (double)n;
                else if (probe == 1) x[i] = ((i & 1) ? -1.0 : 1.0) * (1.0 + (n > 1 ? (double)i / (double)(n - 1) : 0.0));
                else { rng = rng * 6364136223846793005ull + 1442695040888963407ull; x[i] = ((rng >> 33) & 1) ? 1.0 : -1.0; }
            }
            const double xn = norm1(x);
            y = x; if (solve(y) < 0) return -100.0; ++iters;
            est = std::max(est, norm1(y) / xn);
        }
    }
    if (out2) { out2[0] = anorm; out2[1] = est; }
    if (info2) { info2[0] = iters; info2[1] = lower_bound; }
    return anorm * est;
}

// ---- Float32 twins ---------------------------------------------------------------------------
// The reference routes Float32 problems to sgemm/sgetrf/strsm (SpkSpdMMOps.jl:186-351).  On B200 the FP64
// tensor pipe is the fastest pipe this path can use (tcgen05 has no FP32-accumulate-FP32-input kind short of
// TF32 rounding), so the _f32 entry points widen on entry, run the FP64 engine and narrow on exit: results are
// at least as accurate as an FP32 computation; the pivot sequence is the FP64 one.
static void narrow(const std::vector<double>& v, float* a) { for (size_t i = 0; i < v.size(); ++i) a[i] = (float)v[i]; }
static std::vector<double> widen(const float* a, int64_t len) { std::vector<double> v((size_t)std::max<int64_t>(len, 0)); for (int64_t i = 0; i < len; ++i) v[i] = a[i]; return v; }

// The value arrays cross the bus as Float32 and are widened / narrowed ON THE DEVICE (k_widen_f32 / k_narrow_f32).
SPK_API int64_t spk_lufactor_f32(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                                 const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz, float* lnz,
                                 const int64_t* xunz, float* unz, int64_t* ipvt) {
    return dropin_factor(n, nsuper, xsuper, snode, xlindx, lindx, xlnz, lnz, xunz, unz, ipvt, true, true);
}
SPK_API int64_t spk_ldltfactor_f32(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* snode,
                                   const int64_t* xlindx, const int64_t* lindx, const int64_t* xlnz, float* lnz) {
    return dropin_factor(n, nsuper, xsuper, snode, xlindx, lindx, xlnz, lnz, nullptr, nullptr, nullptr, false, true);
}
SPK_API int64_t spk_lulsolve_f32(int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx, const int64_t* lindx,
                                 const int64_t* xlnz, const float* lnz, const int64_t* ipiv, float* rhs) {
    return dropin_solve(n_from_xsuper(nsuper, xsuper), nsuper, xsuper, xlindx, lindx, xlnz, lnz, nullptr, nullptr, ipiv, rhs, true, 1, true);
}
SPK_API int64_t spk_luusolve_f32(int64_t n, int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx,
                                 const int64_t* lindx, const int64_t* xlnz, const float* lnz, const int64_t* xunz,
                                 const float* unz, float* rhs) {
    return dropin_solve(n, nsuper, xsuper, xlindx, lindx, xlnz, lnz, xunz, unz, nullptr, rhs, true, 2, true);
}
SPK_API int64_t spk_ldltsolve_f32(int64_t nsuper, const int64_t* xsuper, const int64_t* xlindx, const int64_t* lindx,
                                  const int64_t* xlnz, const float* lnz, float* rhs) {
    return dropin_solve(n_from_xsuper(nsuper, xsuper), nsuper, xsuper, xlindx, lindx, xlnz, lnz, nullptr, nullptr, nullptr, rhs, false, 0, true);
}
// plan twins: values / right-hand sides cross the boundary as Float32, the plan stays FP64
SPK_API int64_t spk_plan_inmatrix_f32(spk_plan* p, int64_t nnz, const int64_t* dest_or_null, const float* nzval) {
    std::vector<double> v = widen(nzval, nnz);
    return spk_plan_inmatrix(p, nnz, dest_or_null, v.data());
}
SPK_API int64_t spk_plan_get_factors_f32(spk_plan* p, float* lnz, float* unz, int64_t* ipvt) {
    NEED_DEV(p);
    int64_t rc = 0;
    if (lnz) rc = download_values(p, lnz, p->d_lnz, p->P.nlnz, true);          // narrowed on the device
    if (!rc && unz && p->P.nunz > 0) rc = download_values(p, unz, p->d_unz, p->P.nunz, true);
    if (!rc && ipvt) rc = spk_plan_get_factors(p, nullptr, nullptr, ipvt);
    return rc;
}
SPK_API int64_t spk_plan_triangularsolve_f32(spk_plan* p, float* b, int64_t nrhs, int64_t ldb) {
    NEED_DEV(p);
    const int64_t n = p->P.n;
    std::vector<double> r((size_t)n * std::max<int64_t>(nrhs, 0));
    for (int64_t q = 0; q < nrhs; ++q) for (int64_t i = 0; i < n; ++i) r[(size_t)q * n + i] = b[(size_t)q * ldb + i];
    int64_t rc = spk_plan_triangularsolve(p, r.data(), nrhs, n);
    if (!rc) for (int64_t q = 0; q < nrhs; ++q) for (int64_t i = 0; i < n; ++i) b[(size_t)q * ldb + i] = (float)r[(size_t)q * n + i];
    return rc;
}

} // extern "C"
