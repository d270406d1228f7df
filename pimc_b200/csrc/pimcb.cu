// pimcb.cu -- host side of libpimc_b200.so: context, staging, launch logic and the C ABI declared in
// include/pimc_b200.h.  Single backend: CUDA, sm_100a.  There is no CPU fallback anywhere in this file;
// every entry point fails with PIMCB_ECUDA when no device is usable.
#include "../../include/pimc_b200.h"
#include "kernels.cuh"
#include "kernels_ext.cuh"
#include "kernels_pair.cuh"
#include "table_codec.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <array>
#include <cstring>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges are no-ops unless a profiler injects itself (nsys / ncu --nvtx)
#include <new>
#include <string>
#include <thread>
#include <vector>

using namespace pimcb;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(PIMCB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int kSlots = 4;
constexpr int kKernels = 8;
enum { K_RHO = 0, K_CORR = 1, K_DIRECT = 2, K_BINS = 3, K_PAIR = 4, K_TRANSPOSE = 5, K_VARIANT = 6, K_VIRIAL = 7 };

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return fail(PIMCB_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;   // recorded after the last H2D that read this buffer
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e != cudaSuccess) return fail(PIMCB_ENOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Slot {
    DevBuf pos;            // pos[b][t][d][Npad]
    DevBuf aos;            // landing buffer of the zero-bounce path: the reference AoS array as DMA'd
    int B = 0, M = 0, N = 0, Npad = 0, Next = 0;
    bool staged = false;
    unsigned long gen = 0;            // staging generation (identifies the configuration batch held by the slot)
    bool needs_transpose = false;     // aos holds the data; the first consumer transposes it on the compute stream
    cudaEvent_t ready = nullptr;      // H2D (+ transpose) complete
    cudaEvent_t consumed = nullptr;   // last kernel reading this slot has been enqueued before this event
};

}  // namespace

struct pimcb_ctx {
    int device = 0, ndim = 3, sm_count = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    bool have_box = false;
    double side[3] = {0, 0, 0};
    unsigned periodic[3] = {1, 1, 1};
    BoxDev box{};
    // wave-vectors
    int nq = 0, ncomm = 0, nsel = 0;
    std::vector<double> q_host;            // AoS
    std::vector<unsigned char> comm;
    std::vector<int> qn;                   // lattice indices [nq][ndim] (valid when commensurate)
    int nmax[3] = {0, 0, 0};
    int ngroups = 0;                       // sign-symmetry groups of the lattice path (0 = path unavailable)
    double max_phase = 0.0;
    DevBuf d_q, d_comm, d_qn, d_qidx, d_plan;
    size_t plan_off[7] = {0, 0, 0, 0, 0, 0, 0};   // int offsets of gout / ent / tasks / warp_first / gdesc / lmap / rmap in d_plan
    int mma_nL = 0, mma_nR = 0;            // L rows / R cols of the DMMA formulation (0 = not available)
    int last_rho_path = -1, last_ML = 0, last_NR = 0;   // what launch_rho used last: 0 generic, 1 DMMA lattice, 2 CUDA-core lattice
    std::vector<int> mma_lmap, mma_rmap, mma_gout, mma_gdesc;   // host copies of the DMMA plan tables
    int unfold_NR = -1;                    // N-tile count the device unfold table (d_unfold) was built for
    DevBuf d_unfold;
    int lattice_J = 0;                     // 0 = choose from N; else forced (PIMCB_LATTICE_J)
    int lattice_warps = kLatticeWarps;     // warps per CTA of the lattice kernel (PIMCB_LATTICE_WARPS may lower it to 2)
    int rho_mode = 1;
    int corr_mode = 1;                     // DMMA tau-correlation when M <= 510: 1 = one CTA per item, 2 = persistent CTAs with the
                                           // next pair prefetched behind the tensor work; 0 = CUDA-core kernel (PIMCB_CORR_MODE)
    // beads
    Slot slots[kSlots];
    int cur = -1;
    PinBuf pin[2];
    int pin_next = 0;
    // work buffers
    DevBuf d_rho, d_cfg, d_bins, d_partial;
    long n_acc = 0;
    PinBuf h_cnt;                          // pimcb_reduce_bins without num_total: the global count lands here, read at the next sync
    bool n_acc_pending = false;
    // pipelined exchange (pimcb_reduce_bins_begin / _end): a snapshot of the bin travels on its own stream while the next
    // bin accumulates
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t xchg_ready = nullptr, xchg_done = nullptr;
    DevBuf d_xchg;                         // [bins_len] doubles + {my count, total count}
    PinBuf h_xchg;                         // root: the reduced bin; every rank: {my count, total count} behind it
    bool xchg_inflight = false;
    int xchg_root = 0, xchg_nq = 0, xchg_M = 0;
    size_t xchg_len = 0;
    size_t bins_len = 0;
    int bins_M = 0;                        // time slices of the current bin layout
    PinBuf h_out;
    // pair potential
    DevBuf d_V, d_dV, d_vint, d_f2, d_hist;
    DevBuf d_VD, d_DD;                     // (V, dV/dr) and (dV/dr, d2V/dr2) packed four entries per sector (table_codec.h)
    bool vd_ok = false, dd_ok = false;     // the packed tables were built AND verified bit for bit on the device
    long vd_raw = 0, dd_raw = 0;           // sectors left verbatim (zero crossings, core, switch of the damping function)
    std::vector<double> h_dV;              // host copy of dV/dr for packing (dV/dr, d2V/dr2) when the third table arrives
    int tab_len = 0;
    bool have_dV = false;
    double dr = 0, extV[2] = {0, 0}, extdV[2] = {0, 0};
    // profiling: (kernel id, start, stop) event records, resolved lazily by pimcb_kernel_times
    unsigned profiling = 0;                 // bit k set: launches of kernel id k are bracketed by events
    int prof_stride = 1;                    // ... every prof_stride-th launch of that kernel
    long prof_seen[kKernels] = {};
    cudaEvent_t ev0[kKernels] = {}, ev1[kKernels] = {};      // scratch events (fences, fp64 peak)
    struct Rec { int k; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> ev_pool;
    double k_ms[kKernels] = {};
    long k_count[kKernels] = {};
    long launches = 0;
    DevBuf d_scratch;
    DevBuf d_binrows;                      // persistent quad rows [rows][nq][M/2+1] of the measure path (folded into d_bins on read)
    int binrows_n = 0;                     // rows in use since the last fold (0 = nothing pending)
    int binrows_cap = 0;                   // rows the buffer holds for the current (nq, M)
    // per-configuration results cache: d_cfg holds S(q)/F(q,tau) of slot cfg_slot at staging generation cfg_gen
    unsigned long gen_counter = 0, cfg_gen = 0;
    int cfg_slot = -1;
    // scattering variants (elastic, cylinder S(q)) and virial slice sums
    DevBuf d_var, d_inside, d_d2V, d_delta_aos, d_delta, d_vir, d_gext, d_g2ext;
    unsigned long g2ext_gen = 0;            // staging generation the uploaded external-potential Laplacian belongs to (0 = none)
    unsigned long gext_gen = 0;             // staging generation the uploaded external-potential gradient belongs to (0 = none)
    bool have_d2V = false;
    double extd2V[2] = {0, 0};
    // single-walker fast path: the fixed sequence H2D -> transpose -> rho_q -> tau-correlation -> D2H of
    // pimcb_ssf_isf_beads as ONE CUDA graph launch (captured on the second call with the same source / shape / q-set)
    struct FusedGraph {
        cudaGraphExec_t exec = nullptr;
        const double* src = nullptr;
        int M = 0, N = 0, Next = 0, slot = -1, rho_mode = -1, corr_mode = -1, primed = 0, failed = 0;
        unsigned long qgen = 0;
        long launches = 0;
        const void* baked[8] = {};          // device / pinned addresses baked into the graph; any change drops it
    } fused;
    unsigned long qgen = 0;                 // bumped by pimcb_set_qvecs
    int use_graph = 1;                      // PIMCB_GRAPH=0 disables the graph path
    // multi-GPU: NCCL communicator (library resolved with dlopen at pimcb_comm_init; nothing links against NCCL)
    void* nccl_comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    DevBuf d_gather, d_count;
    DevBuf d_sched;                        // ticket counter + retire counter of the persistent-warp rho kernel (self re-arming)
};

namespace {

// ---- NCCL, resolved at run time ---------------------------------------------------------------------------------------
// Only the handful of entry points the path's single exchange step needs (one reduce / all-gather per bin).
// Prototypes follow nccl.h (2.x ABI): ncclUniqueId is 128 opaque bytes passed by value, ncclDouble = 8, ncclSum = 0,
// ncclInt64 = 4.
struct NcclId { char internal[128]; };
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return 0;
    // An NCCL already in the process (e.g. the one a host framework brought along) is used as is: loading a second copy
    // under the same soname would shadow it for everything loaded later.  Otherwise PIMCB_NCCL_LIB, then the default
    // sonames.  RTLD_LOCAL: our symbols are taken from the handle, nothing is injected into the global namespace.
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    const char* names[] = {std::getenv("PIMCB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (h) break;
        if (n && *n) h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    }
    if (!h) return fail(PIMCB_ESTATE, "NCCL library not found (set PIMCB_NCCL_LIB): %s", dlerror());
    NcclApi a;
    a.handle = h;
#define NCCL_SYM(field, name) *reinterpret_cast<void**>(&a.field) = dlsym(h, name); \
    if (!a.field) return fail(PIMCB_ESTATE, "NCCL symbol %s missing", name)
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    NCCL_SYM(CommInitRank, "ncclCommInitRank");
    NCCL_SYM(CommDestroy, "ncclCommDestroy");
    NCCL_SYM(Reduce, "ncclReduce");
    NCCL_SYM(AllGather, "ncclAllGather");
    NCCL_SYM(GroupStart, "ncclGroupStart");
    NCCL_SYM(GroupEnd, "ncclGroupEnd");
    NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
    g_nccl = a;
    return 0;
}
#define NCCLCHK(call)                                                                                   \
    do {                                                                                                \
        int r_ = (call);                                                                                \
        if (r_ != 0) return fail(PIMCB_ECUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_));       \
    } while (0)

cudaEvent_t pool_event(pimcb_ctx* c) {
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// NVTX range names per kernel id (the reference has no tracing; these make the stage / rho / corr /
// pair / reduce phases visible on a profiler's timeline).
const char* const kRangeNames[kKernels] = {"pimcb:rho_q", "pimcb:tau_correlation", "pimcb:ssf_direct", "pimcb:bins", "pimcb:pair_sums",
                                           "pimcb:aos_to_soa", "pimcb:variant", "pimcb:virial_sums"};
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// Brackets one kernel launch on `stream` with CUDA events when profiling is on; always counts the launch.
struct KTimer {
    pimcb_ctx* c; int k; cudaStream_t st; cudaEvent_t a = nullptr;
    KTimer(pimcb_ctx* c_, int k_, cudaStream_t st_ = nullptr) : c(c_), k(k_), st(st_ ? st_ : c_->stream) {
        nvtxRangePushA(kRangeNames[k]);
        if (((c->profiling >> k) & 1u) && (c->prof_seen[k]++ % c->prof_stride) == 0 && (a = pool_event(c))) cudaEventRecord(a, st);
    }
    ~KTimer() {
        nvtxRangePop();
        if (a) {
            cudaEvent_t b = pool_event(c);
            if (b) { cudaEventRecord(b, st); c->recs.push_back({k, a, b}); }
        }
        c->launches++;
    }
};

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Choose the particle split P (chunks per q) for the rho kernels: maximise lane utilisation of a
// 256-thread CTA subject to the partial-sum shared-memory budget.
void choose_split(int nitems, int N, int threads, size_t smem_fixed, size_t part_bytes_per_item, size_t smem_limit,
                  int* P_out, int* chunk_out) {
    int bestP = 1;
    double best = -1.0;
    for (int P = 1; P <= 64 && P <= N; ++P) {
        const int chunk = (N + P - 1) / P;
        if (static_cast<long>(chunk) * (P - 1) >= N) continue;                       // empty trailing chunk
        const size_t smem = smem_fixed + part_bytes_per_item * P * nitems;
        if (P > 1 && smem > smem_limit) break;
        const long items = static_cast<long>(nitems) * P;
        const long passes = (items + threads - 1) / threads;
        double eff = static_cast<double>(nitems) * N / (static_cast<double>(passes) * threads * chunk);
        if (items < threads) eff *= 0.999;                                           // prefer full CTAs on ties
        eff -= 1e-4 * P;                                                             // prefer fewer partials on ties
        if (eff > best) { best = eff; bestP = P; }
    }
    *P_out = bestP;
    *chunk_out = (N + bestP - 1) / bestP;
}

// One CTA per (configuration, slice): the hardware CTA scheduler hands slices out dynamically, so SMs never
// idle behind a static partition (a fixed grid of k CTAs/SM with fewer than k resident serialises the excess).
int grid_for(const pimcb_ctx*, int nslices, int /*ctas_per_sm*/) {
    return std::max(1, nslices);
}

template <class K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    return 0;
}

int need_cur(pimcb_ctx* c, Slot** s) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    if (c->cur < 0 || !c->slots[c->cur].staged) return fail(PIMCB_ESTATE, "no beads staged");
    *s = &c->slots[c->cur];
    return 0;
}

// Scheduler words of the persistent-warp rho kernel: [0] ticket, [1] retired warps, [2 + slice] parts done (split > 1).
// All of them are left at zero by the kernel itself; growing the buffer re-zeroes it on the compute stream.
int ensure_sched(pimcb_ctx* c, size_t words) {
    if (sizeof(unsigned) * words <= c->d_sched.cap) return 0;
    CU(cudaStreamSynchronize(c->stream));
    int rc = c->d_sched.ensure(sizeof(unsigned) * std::max<size_t>(words, 4096));
    if (rc) return rc;
    CU(cudaMemsetAsync(c->d_sched.p, 0, c->d_sched.cap, c->stream));
    return 0;
}

// Unfold table of the DMMA rho kernel: for every output value k = 2 q + {re, im} the two entries of the staged C tiles
// it is made of and their signs (kernels.cuh, MmaPlan::unfold).  With K0..K3 = C(rr,cr), C(ri,ci), C(rr,ci), C(ri,cr):
//   rho_re = K0 -+ sg K1,   rho_im = +-(sg K3 +- K2)
// where (rr, ri) are the rows of X^a Y^b or X^a conj(Y^b) (3-D, chosen by the relative sign of the second component),
// sg = -1 when the stored row is the conjugate of the wanted one, and the remaining signs follow the sign pattern.
int build_unfold(pimcb_ctx* c, int NR) {
    if (c->unfold_NR == NR) return 0;
    const int nd = c->ndim, npat = 1 << nd, G = c->ngroups, nq = c->nq;
    const int zrow = c->mma_nL - 1;
    auto off = [NR](int row, int col) {
        return static_cast<unsigned>(((row >> 3) * NR + (col >> 3)) * 64 + ((row & 7) * 4 + ((col & 7) >> 1)) * 2 + (col & 1));
    };
    auto pack = [](unsigned o0, bool n0, unsigned o1, bool n1) { return o0 | (o1 << 12) | (n0 ? 1u << 24 : 0u) | (n1 ? 1u << 25 : 0u); };
    std::vector<unsigned> tab(2 * static_cast<size_t>(nq), 0u);
    for (int g = 0; g < G; ++g) {
        const int* e = &c->mma_gdesc[static_cast<size_t>(g) * 8];
        for (int pat = 0; pat < npat; ++pat) {
            const int iq = c->mma_gout[static_cast<size_t>(g) * npat + pat];
            if (iq < 0) continue;
            const int sa = pat & 1;                                  // the pattern with all signs flipped is the conjugate
            const int sb = nd > 1 ? (((pat >> 1) & 1) ^ sa) : 0;
            const int sc = nd > 2 ? (((pat >> 2) & 1) ^ sa) : 0;
            const int cr = e[5], ci = e[6];
            if (nd == 1) {
                tab[2 * iq + 0] = pack(off(e[0], cr), false, off(zrow, 0), false);
                tab[2 * iq + 1] = pack(off(e[0], ci), sa != 0, off(zrow, 0), false);
                continue;
            }
            const int side = nd == 3 ? sb : 0;
            const int rr = side ? e[2] : e[0], ri = side ? e[3] : e[1];
            const bool sgneg = side && e[4] < 0;
            const int sl = nd == 3 ? sc : sb;                        // sign of the last multiplied factor
            // re = K0 + (sl ? +sg : -sg) K1
            tab[2 * iq + 0] = pack(off(rr, cr), false, off(ri, ci), sl ? sgneg : !sgneg);
            // im = (sa ? -1 : 1) [ sg K3 + (sl ? -1 : 1) K2 ]
            tab[2 * iq + 1] = pack(off(ri, cr), sgneg != (sa != 0), off(rr, ci), (sl != 0) != (sa != 0));
        }
    }
    int rc = c->d_unfold.ensure(sizeof(unsigned) * tab.size());
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->d_unfold.p, tab.data(), sizeof(unsigned) * tab.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));        // `tab` is a local
    c->unfold_NR = NR;
    return 0;
}

// ---- rho_q build + correlation + direct S(q) into d_cfg (per configuration results) ------------
// Tile shape and shared memory of the DMMA formulation for the current q-set; false when it does not apply.
static bool rho_mma_shape(const pimcb_ctx* c, int* ML_out, int* NR_out, size_t* smem_out) {
    const int nd = c->ndim;
    auto up = [](int v) { return v <= 1 ? 1 : (v <= 2 ? 2 : (v <= 4 ? 4 : (v <= 8 ? 8 : 16))); };
    const int mlx = (c->mma_nL + 7) / 8, nrx = (c->mma_nR + 7) / 8;        // exact tile counts
    const int NR = c->mma_nR > 0 ? (nrx <= 4 ? nrx : up(nrx)) : 0;         // NT = 1..4 compiled exactly
    // MT = 1..8 compiled exactly for NT <= 2, MT = 1..4 for NT = 3, 4 (C3's 2-D grid of 17 x 17 q: 3 x 3 tiles; rounded up
    // to the 4 x 4 shape it ran 16 DMMAs per k-step instead of 9)
    const int ML = c->mma_nL > 0 ? ((NR <= 2 && mlx <= 8) || (NR > 2 && mlx <= 4) ? mlx : up(mlx)) : 0;
    const bool mma_fits = ML > 0 && NR > 0 && c->mma_nL <= 128 && NR <= 4 && ML * NR <= 16 &&
                          c->mma_lmap.size() <= 81 && c->mma_rmap.size() <= 17 && (nd < 3 || c->nmax[1] <= 8);
    const size_t mma_smem = sizeof(double) * kMmaWarps * (8 * static_cast<size_t>(ML + NR) * kMmaStride + static_cast<size_t>(ML) * NR * 64) +
                            sizeof(unsigned) * 2 * static_cast<size_t>(c->nq);
                            // per-warp operand planes + C staging, CTA copy of the unfold table
    if (ML_out) *ML_out = ML;
    if (NR_out) *NR_out = NR;
    if (smem_out) *smem_out = mma_smem;
    return c->rho_mode == 1 && c->ngroups > 0 && mma_fits && mma_smem <= 160 * 1024;
}

// aos_src != nullptr (single-walker graph, DMMA formulation only -- check rho_mma_shape first): the kernel reads the
// reference's beads array double[M][Next][nd] at aos_src (device-visible: the page-locked host array itself) instead of
// s.pos, and writes s.pos on the way.
int launch_rho(pimcb_ctx* c, const Slot& s, const double* aos_src = nullptr) {
    const int nd = c->ndim, nq = c->nq;
    const int nsl = s.B * s.M;
    const size_t limit = 200 * 1024;
    int rows = 0;
    for (int d = 0; d < nd; ++d) rows += c->nmax[d] + 1;
    const int J = c->lattice_J ? c->lattice_J : (s.N > 128 ? 8 : (s.N > 64 ? 4 : (s.N > 32 ? 2 : 1)));
    const int JJ = J >= 8 ? 8 : (J >= 4 ? 4 : (J >= 2 ? 2 : 1));
    const size_t lattice_fixed = sizeof(double) * (2 * static_cast<size_t>(rows) * lattice_stride(s.N, JJ) + static_cast<size_t>(nd) * s.Npad);
    // lattice path only when every q is commensurate and the phase-power table leaves room for >= 2 CTAs per SM
    const bool lattice = c->rho_mode >= 1 && c->ngroups > 0 && lattice_fixed <= 100 * 1024;
    int rc = c->d_rho.ensure(sizeof(double) * 8 * static_cast<size_t>(s.B) * rho_tblocks(s.M) * nq);   // rho[b][t/4][q][cs][t%4]
    if (rc) return rc;
    int P = 1, chunk = 0;
    KTimer kt(c, K_RHO);
    const int grid = grid_for(c, nsl, 8);
    // DMMA formulation: tile counts rounded up to a compiled accumulator shape MT x NT (MT*NT <= 16)
    int ML = 0, NR = 0;
    size_t mma_smem = 0;
    const bool use_mma = rho_mma_shape(c, &ML, &NR, &mma_smem);
    if (aos_src && !use_mma) return fail(PIMCB_ESTATE, "the direct beads-array form needs the DMMA rho_q kernel");
    if (use_mma) {
        const int3 nmax = make_int3(c->nmax[0], c->nmax[1], c->nmax[2]);
        const double twopi = 2.0 * M_PI;
        const double3 kph = make_double3(twopi / c->side[0], nd > 1 ? twopi / c->side[1] : 0.0, nd > 2 ? twopi / c->side[2] : 0.0);
        const int* pl = c->d_plan.as<int>();
        MmaPlan plan{};
        (void)pl;
        if ((rc = build_unfold(c, NR))) return rc;
        plan.unfold = c->d_unfold.as<unsigned>();
        plan.G = c->ngroups; plan.nL = c->mma_nL; plan.nR = c->mma_nR;
        for (short& v : plan.lmap) v = -1;      // the unrolled kernels probe entries beyond this q-set's nmax
        for (short& v : plan.rmap) v = -1;
        for (size_t k = 0; k < c->mma_lmap.size(); ++k) plan.lmap[k] = static_cast<short>(c->mma_lmap[k]);
        for (size_t k = 0; k < c->mma_rmap.size(); ++k) plan.rmap[k] = static_cast<short>(c->mma_rmap[k]);
        c->last_rho_path = 1; c->last_ML = ML; c->last_NR = NR;
        int pgrid = 0;
        const int nchunk = (s.N + kMmaChunk - 1) / kMmaChunk;      // 32-particle blocks per slice
        const int nm3 = std::max(c->nmax[0], std::max(c->nmax[1], c->nmax[2]));
#define LAUNCH_MMA_NM2(ND, MT, NT, NM, DIRECT)                                                                             \
        { rc = set_smem(rho_lattice_mma_kernel<ND, MT, NT, NM, DIRECT>, mma_smem); if (rc) return rc;                      \
        int occ = 1;                                                                                               \
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rho_lattice_mma_kernel<ND, MT, NT, NM, DIRECT>, 128, mma_smem)); \
        if (const char* e = std::getenv("PIMCB_RHO_OCC")) occ = std::max(1, std::min(occ, std::atoi(e)));          \
        const int wmax = c->sm_count * std::max(1, occ) * kMmaWarps;          /* resident warps of a full grid */   \
        int split = 1;                                                         /* warps sharing one slice */         \
        while (split < 8 && 2 * split <= nchunk && nsl * split < wmax) split *= 2;                                   \
        /* few waves of work items: a ragged last wave costs every warp a whole item (C4, 8 walkers per GPU: 2560 slices \
           on 1776 resident warps = 1.44 waves, 72 % busy); split further while that buys >= 10 % */                     \
        { auto eff = [&](int sp) { const double w = static_cast<double>(nsl) * sp / wmax; return w / std::ceil(w); };  \
          while (split < 8 && 2 * split <= nchunk && static_cast<double>(nsl) * split / wmax < 4.0 &&                    \
                 eff(2 * split) > eff(split) + 0.10) split *= 2; }                                                       \
        if (const char* e = std::getenv("PIMCB_RHO_SPLIT")) split = std::max(1, std::min(nchunk, std::atoi(e)));   \
        if (split > 1) {                                                                                           \
            rc = c->d_partial.ensure(sizeof(double) * static_cast<size_t>(nsl) * split * (MT * NT * 64)); if (rc) return rc; \
            rc = ensure_sched(c, 2 + static_cast<size_t>(nsl)); if (rc) return rc;                                 \
        }                                                                                                          \
        const int items = nsl * split;                                                                             \
        pgrid = std::max(1, std::min((items + kMmaWarps - 1) / kMmaWarps, c->sm_count * std::max(1, occ)));        \
        rho_lattice_mma_kernel<ND, MT, NT, NM, DIRECT><<<pgrid, 128, mma_smem, c->stream>>>(DIRECT ? aos_src : s.pos.as<double>(), plan,         \
                                                                                    c->d_rho.as<double>(),                   \
                                                                                    nsl, s.N, s.Npad, nq, nmax, kph,         \
                                                                                    c->d_sched.as<unsigned>(), 0, split,     \
                                                                                    c->d_partial.as<double>(), s.M, 0xffffffffu, \
                                                                                    make_int3(s.Next * ND, ND, 1),           \
                                                                                    DIRECT ? s.pos.as<double>() : nullptr); }
#define LAUNCH_MMA_NM(ND, MT, NT, NM) { if (aos_src) LAUNCH_MMA_NM2(ND, MT, NT, NM, true) else LAUNCH_MMA_NM2(ND, MT, NT, NM, false) }
        // 3-D with every |n_d| <= 2 (or 3): phase A fully unrolled; the R columns then fit one N tile
#define LAUNCH_MMA(ND, MT, NT)                                                                                    \
        if (ND == 3 && NT == 1 && nm3 <= 2) LAUNCH_MMA_NM(ND, MT, NT, (ND == 3 && NT == 1 ? 2 : 0))                \
        else if (ND == 3 && NT == 1 && nm3 <= 3) LAUNCH_MMA_NM(ND, MT, NT, (ND == 3 && NT == 1 ? 3 : 0))           \
        else LAUNCH_MMA_NM(ND, MT, NT, 0)
#define LAUNCH_MMA_M8(ND, NT)                                                                                     \
        switch (ML) { case 1: LAUNCH_MMA(ND, 1, NT) break; case 2: LAUNCH_MMA(ND, 2, NT) break;                    \
                      case 3: LAUNCH_MMA(ND, 3, NT) break; case 4: LAUNCH_MMA(ND, 4, NT) break;                    \
                      case 5: LAUNCH_MMA(ND, 5, NT) break; case 6: LAUNCH_MMA(ND, 6, NT) break;                    \
                      case 7: LAUNCH_MMA(ND, 7, NT) break; default: LAUNCH_MMA(ND, 8, NT) break; }
#define LAUNCH_MMA_SHAPE(ND)                                                                                      \
        if (NR == 1) { if (ML == 16) LAUNCH_MMA(ND, 16, 1) else LAUNCH_MMA_M8(ND, 1) }                             \
        else if (NR == 2) { LAUNCH_MMA_M8(ND, 2) }                                                                 \
        else if (NR == 3) { switch (ML) { case 1: LAUNCH_MMA(ND, 1, 3) break; case 2: LAUNCH_MMA(ND, 2, 3) break;  \
                                          case 3: LAUNCH_MMA(ND, 3, 3) break; default: LAUNCH_MMA(ND, 4, 3) break; } } \
        else { switch (ML) { case 1: LAUNCH_MMA(ND, 1, 4) break; case 2: LAUNCH_MMA(ND, 2, 4) break;               \
                             case 3: LAUNCH_MMA(ND, 3, 4) break; default: LAUNCH_MMA(ND, 4, 4) break; } }
        if (nd == 1) { LAUNCH_MMA_SHAPE(1) } else if (nd == 2) { LAUNCH_MMA_SHAPE(2) } else { LAUNCH_MMA_SHAPE(3) }
#undef LAUNCH_MMA_SHAPE
#undef LAUNCH_MMA_NM2
#undef LAUNCH_MMA_M8
#undef LAUNCH_MMA
#undef LAUNCH_MMA_NM
        CU(cudaGetLastError());
        return 0;
    }
    c->last_rho_path = lattice ? 2 : 0;
    if (!lattice) {
        const size_t fixed = sizeof(double) * nd * s.Npad;
        choose_split(nq, s.N, 256, fixed, sizeof(double) * 2, limit, &P, &chunk);
        const size_t smem = fixed + (P > 1 ? sizeof(double) * 2 * P * nq : 0);
#define LAUNCH_GENERIC(ND)                                                                                        \
        rc = set_smem(rho_generic_kernel<ND>, smem); if (rc) return rc;                                            \
        rho_generic_kernel<ND><<<grid, 256, smem, c->stream>>>(s.pos.as<double>(), c->d_q.as<double>(),           \
                                                               c->d_rho.as<double>(), nsl, s.N, s.Npad, nq, P, chunk, s.M)
        if (nd == 1) { LAUNCH_GENERIC(1); } else if (nd == 2) { LAUNCH_GENERIC(2); } else { LAUNCH_GENERIC(3); }
#undef LAUNCH_GENERIC
    } else {
        const int G = c->ngroups;
        const int nk = nd == 3 ? 8 : (nd == 2 ? 4 : 2);
        const size_t smem = lattice_fixed + sizeof(double) * nk * G;
        const int3 nmax = make_int3(c->nmax[0], c->nmax[1], c->nmax[2]);
        const double twopi = 2.0 * M_PI;
        const double3 kph = make_double3(twopi / c->side[0], nd > 1 ? twopi / c->side[1] : 0.0, nd > 2 ? twopi / c->side[2] : 0.0);
        const int* pl = c->d_plan.as<int>();
        const LatticePlan plan{pl + c->plan_off[0], pl + c->plan_off[1], pl + c->plan_off[2], pl + c->plan_off[3], G};
        (void)P; (void)chunk;
#define LAUNCH_LATTICE(ND, JJ)                                                                                    \
        rc = set_smem(rho_lattice_kernel<ND, JJ>, smem); if (rc) return rc;                                        \
        rho_lattice_kernel<ND, JJ><<<grid, 32 * c->lattice_warps, smem, c->stream>>>(s.pos.as<double>(), plan, c->d_rho.as<double>(), nsl, \
                                                                   s.N, s.Npad, nq, nmax, kph, s.M)
#define LAUNCH_LATTICE_J(ND)                                                                                      \
        if (J >= 8) { LAUNCH_LATTICE(ND, 8); } else if (J >= 4) { LAUNCH_LATTICE(ND, 4); }                         \
        else if (J >= 2) { LAUNCH_LATTICE(ND, 2); } else { LAUNCH_LATTICE(ND, 1); }
        if (nd == 1) { LAUNCH_LATTICE_J(1) } else if (nd == 2) { LAUNCH_LATTICE_J(2) } else { LAUNCH_LATTICE_J(3) }
#undef LAUNCH_LATTICE_J
#undef LAUNCH_LATTICE
    }
    CU(cudaGetLastError());
    return 0;
}

// Folds the pending quad rows of the measure path into d_bins (one small launch per bin read-out).
int fold_binrows(pimcb_ctx* c, int M) {
    if (c->binrows_n == 0) return 0;
    const int total = c->nq * (M / 2 + 1);
    bins_fold_kernel<<<(total + 127) / 128, 128, 0, c->stream>>>(c->d_binrows.as<double>(), c->d_bins.as<double>(), c->binrows_n, c->nq, M,
                                                                  c->d_comm.as<unsigned char>());
    CU(cudaGetLastError());
    c->launches++;
    c->binrows_n = 0;
    return 0;
}

// partial_rows != nullptr asks for the quad-summed form (one row per four configurations, folded into the bin by the
// kernel itself) when the DMMA kernel can provide it; *partial_rows returns the number of rows of d_cfg that still have
// to be accumulated into the bin (0 after the fused form, B after a per-configuration form).
int launch_corr(pimcb_ctx* c, const Slot& s, int* partial_rows = nullptr) {
    KTimer kt(c, K_CORR);
    if (partial_rows) *partial_rows = s.B;
    const int mtc = (s.M / 2 + 1 + 63) / 64;                      // 64-tau accumulator tiles of the DMMA formulation
    if (c->corr_mode >= 1 && mtc <= 4) {
        // 2: persistent CTAs, next pair prefetched behind the DMMAs (at most two periodic images per value: M >= 64 mtc + 11)
        const bool pipe = c->corr_mode == 2 && s.M >= 64 * mtc + 11;
        const int off = 64 * mtc, mpad = (s.M + 3) & ~3, ext = off + mpad + 8;
        const int plen = ext + 4 * (ext >> 3) + 4;
        size_t smem = sizeof(double) * 2 * static_cast<size_t>(plen) * 4;
        if (pipe) {
            size_t park = partial_rows ? 8 * static_cast<size_t>(s.M / 2 + 1) : 0;   // double-buffered parking area (doubles)
            smem += sizeof(double) * park;
        }
        // PIMCB_CORR_OCC=n: pad the shared-memory request so that at most n CTAs are resident per SM (occupancy A/B)
        static const int corr_occ = std::getenv("PIMCB_CORR_OCC") ? std::atoi(std::getenv("PIMCB_CORR_OCC")) : 0;
        if (corr_occ > 0) smem = std::max(smem, static_cast<size_t>(226 * 1024) / corr_occ - 1024);
        const int quads = (s.B + 3) / 4;
        const bool partial = partial_rows != nullptr;
        int rc = 0;
        if (partial) {
            if (quads > c->binrows_cap) {                                    // grow the persistent rows (pending sums folded first)
                if ((rc = fold_binrows(c, s.M))) return rc;
                const size_t bytes = sizeof(double) * quads * c->nq * (s.M / 2 + 1);
                if ((rc = c->d_binrows.ensure(bytes))) return rc;
                CU(cudaMemsetAsync(c->d_binrows.p, 0, c->d_binrows.cap, c->stream));
                c->binrows_cap = quads;
            }
            c->binrows_n = std::max(c->binrows_n, quads);
        }
#define LAUNCH_CORR_MMA2(MTC, PART)                                                                                \
        rc = set_smem(isf_corr_mma_kernel<MTC, PART>, smem); if (rc) return rc;                                    \
        isf_corr_mma_kernel<MTC, PART><<<quads * c->nq, 128, smem, c->stream>>>(c->d_rho.as<double>(),                          \
                                                                                PART ? c->d_binrows.as<double>() : c->d_cfg.as<double>(), \
                                                                                s.M, c->nq, s.B, 1.0 / s.N, c->d_comm.as<unsigned char>())
#define LAUNCH_CORR_PIPE2(MTC, PART)                                                                               \
        { rc = set_smem(isf_corr_mma_pipe_kernel<MTC, PART>, smem); if (rc) return rc;                             \
          int occ = 1;                                                                                             \
          CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, isf_corr_mma_pipe_kernel<MTC, PART>, 128, smem)); \
          const int nitems = quads * c->nq;                                                                        \
          const int pgrid = std::max(1, std::min(nitems, c->sm_count * std::max(1, occ)));                         \
          isf_corr_mma_pipe_kernel<MTC, PART><<<pgrid, 128, smem, c->stream>>>(c->d_rho.as<double>(),              \
                                                                                PART ? c->d_binrows.as<double>() : c->d_cfg.as<double>(), \
                                                                                s.M, c->nq, s.B, 1.0 / s.N, c->d_comm.as<unsigned char>(), nitems); }
#define LAUNCH_CORR_MMA(MTC) if (pipe) { if (partial) LAUNCH_CORR_PIPE2(MTC, true) else LAUNCH_CORR_PIPE2(MTC, false) }       \
                             else if (partial) { LAUNCH_CORR_MMA2(MTC, true); } else { LAUNCH_CORR_MMA2(MTC, false); }
        if (mtc == 1) { LAUNCH_CORR_MMA(1) } else if (mtc == 2) { LAUNCH_CORR_MMA(2) } else if (mtc == 3) { LAUNCH_CORR_MMA(3) } else { LAUNCH_CORR_MMA(4) }
#undef LAUNCH_CORR_MMA
#undef LAUNCH_CORR_MMA2
#undef LAUNCH_CORR_PIPE2
        CU(cudaGetLastError());
        if (partial) *partial_rows = 0;         // accumulated into the persistent quad rows by the kernel
        return 0;
    }
    const int nblk = (s.M / 2 + 1 + 7) / 8;                       // tau blocks of 8 per (config, q) pair
    const int lpq = nblk <= 8 ? 8 : (nblk <= 16 ? 16 : 32);       // lanes owning the tau blocks of one pair
    const int tsplit = lpq <= 16 ? 2 : 1;                         // lane groups splitting the t0 range of the pair
    const int ppc = (32 / (lpq * tsplit)) * 4;                    // pairs per 128-thread CTA
    const int len = 2 * s.M + 16;
    const int plen = len + 2 * (len >> 3) + 2;
    const size_t smem = sizeof(double) * 2 * static_cast<size_t>(plen) * ppc;
    if (smem > 200 * 1024) return fail(PIMCB_EINVAL, "M = %d too large for the correlation kernel's shared-memory staging", s.M);
    int rc = set_smem(isf_corr_kernel, smem);
    if (rc) return rc;
    const int npairs = s.B * c->nq;
    isf_corr_kernel<<<(npairs + ppc - 1) / ppc, 128, smem, c->stream>>>(c->d_rho.as<double>(), c->d_cfg.as<double>(), s.M, c->nq,
                                                                          npairs, lpq, tsplit, 1.0 / s.N, c->d_comm.as<unsigned char>());
    CU(cudaGetLastError());
    return 0;
}

int launch_direct(pimcb_ctx* c, const Slot& s) {
    if (c->nsel == 0) return 0;
    const int nd = c->ndim, nsl = s.B * s.M;
    int rc = c->d_partial.ensure(sizeof(double) * static_cast<size_t>(nsl) * c->nsel);
    if (rc) return rc;
    KTimer kt(c, K_DIRECT);
    const size_t smem = sizeof(double) * nd * s.Npad;
    const int grid = grid_for(c, nsl, 8);
#define LAUNCH_DIRECT(ND)                                                                                          \
    rc = set_smem(ssf_direct_kernel<ND, 4>, smem); if (rc) return rc;                                              \
    ssf_direct_kernel<ND, 4><<<grid, 256, smem, c->stream>>>(s.pos.as<double>(), c->d_q.as<double>(), c->d_qidx.as<int>(), \
                                                             c->nsel, c->nq, c->d_partial.as<double>(), nsl, s.N, s.Npad, c->box)
    if (nd == 1) { LAUNCH_DIRECT(1); } else if (nd == 2) { LAUNCH_DIRECT(2); } else { LAUNCH_DIRECT(3); }
#undef LAUNCH_DIRECT
    CU(cudaGetLastError());
    const int tot = s.B * c->nsel;
    ssf_direct_finalize_kernel<<<(tot + 127) / 128, 128, 0, c->stream>>>(c->d_partial.as<double>(), c->d_qidx.as<int>(), c->nsel,
                                                                         c->d_cfg.as<double>(), s.B, s.M, s.N, c->nq);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

int materialize(pimcb_ctx* c, Slot& s);

// rows != nullptr (pimcb_measure): only the sum over configurations is wanted, so the correlation may write
// quad-summed rows; *rows = number of rows of d_cfg to accumulate into the bin.
int run_estimators(pimcb_ctx* c, Slot** sp, int* rows = nullptr) {
    Slot* s;
    int rc = need_cur(c, &s);
    if (rc) return rc;
    if (c->nq <= 0) return fail(PIMCB_ESTATE, "no q-vectors set");
    if (!c->have_box) return fail(PIMCB_ESTATE, "no box set");
    if (c->max_phase > 1.0e5) return fail(PIMCB_EINVAL, "max |q.r| = %g exceeds the sincos validity range 1e5", c->max_phase);
    const size_t len = static_cast<size_t>(c->nq) * (1 + s->M);
    if (c->bins_len != len) {
        rc = c->d_bins.ensure(sizeof(double) * len);
        if (rc) return rc;
        CU(cudaMemsetAsync(c->d_bins.p, 0, sizeof(double) * len, c->stream));
        c->bins_len = len;
        c->bins_M = s->M;
        c->n_acc = 0;
        c->binrows_n = 0;                  // pending quad rows belonged to the old layout
        c->binrows_cap = 0;                // re-zeroed and re-sized on the next measurement
    }
    rc = c->d_cfg.ensure(sizeof(double) * len * s->B);
    if (rc) return rc;
    CU(cudaStreamWaitEvent(c->stream, s->ready, 0));
    if ((rc = materialize(c, *s))) return rc;
    if ((rc = launch_rho(c, *s))) return rc;
    // quad-summed rows only when every S(q) comes from the correlation (no direct min-image q writes per-configuration rows)
    const bool partial = rows != nullptr && c->nsel == 0;
    if (rows) *rows = s->B;
    if ((rc = launch_corr(c, *s, partial ? rows : nullptr))) return rc;
    if ((rc = launch_direct(c, *s))) return rc;
    CU(cudaEventRecord(s->consumed, c->stream));
    *sp = s;
    // d_cfg holds complete per-configuration rows only on the per-configuration path (no quad-summed rows)
    c->cfg_slot = (rows == nullptr) ? c->cur : -1;
    c->cfg_gen = s->gen;
    return 0;
}

int copy_out(pimcb_ctx* c, const Slot& s, double* ssf_out, double* isf_out) {
    const size_t len = static_cast<size_t>(c->nq) * (1 + s.M);
    const size_t bytes = sizeof(double) * len * s.B;
    int rc = c->h_out.ensure(bytes);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->h_out.p, c->d_cfg.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const double* h = static_cast<const double*>(c->h_out.p);
    for (int b = 0; b < s.B; ++b) {
        if (ssf_out) std::memcpy(ssf_out + static_cast<size_t>(b) * c->nq, h + b * len, sizeof(double) * c->nq);
        if (isf_out) std::memcpy(isf_out + static_cast<size_t>(b) * c->nq * s.M, h + b * len + c->nq, sizeof(double) * c->nq * s.M);
    }
    return 0;
}

// Makes sure slot s is in the kernels' SoA layout: a slot staged through the zero-bounce path is transposed here, on
// the compute stream, so that the copy stream carries nothing but back-to-back DMAs.
int materialize(pimcb_ctx* c, Slot& s) {
    if (!s.needs_transpose) return 0;
    const int nd = c->ndim;
    const size_t nsl = static_cast<size_t>(s.B) * s.M;
    const size_t smem = sizeof(double) * s.N * nd;
    const int grid = static_cast<int>(std::min<size_t>(nsl, static_cast<size_t>(c->sm_count) * 8));
    int rc = 0;
#define LAUNCH_T(ND)                                                                                  \
    rc = set_smem(aos_to_soa_kernel<ND>, smem); if (rc) return rc;                                     \
    aos_to_soa_kernel<ND><<<grid, 256, smem, c->stream>>>(s.aos.as<double>(), s.pos.as<double>(), static_cast<int>(nsl), s.N, s.Next, s.Npad)
    {
        KTimer kt(c, K_TRANSPOSE);
        if (nd == 1) { LAUNCH_T(1); } else if (nd == 2) { LAUNCH_T(2); } else { LAUNCH_T(3); }
    }
#undef LAUNCH_T
    CU(cudaGetLastError());
    s.needs_transpose = false;
    return 0;
}

// wait = false (pimcb_stage_batch_async): a page-locked source is NOT waited for; the caller keeps it untouched until
// pimcb_stage_wait or any later synchronising call on this slot's data.
int stage_into(pimcb_ctx* c, int slot, const double* beads, int B, int M, int N, int Next, bool wait = true) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    if (!beads || B < 1 || M < 1 || N < 1 || Next < N) return fail(PIMCB_EINVAL, "bad staging arguments (B=%d M=%d N=%d N_ext=%d)", B, M, N, Next);
    if (slot < 0 || slot >= kSlots) return fail(PIMCB_EINVAL, "slot %d out of range", slot);
    CU(cudaSetDevice(c->device));
    NvtxRange range("pimcb:stage");
    const int nd = c->ndim;
    Slot& s = c->slots[slot];
    const int Npad = round_up(N, 16);
    const size_t nsl = static_cast<size_t>(B) * M;
    const size_t soa_bytes = sizeof(double) * nsl * nd * Npad;
    // kernels already enqueued on the compute stream may still read THIS slot: order the copy after them
    // (other slots are untouched, so H2D into slot k+1 overlaps the kernels working on slot k)
    CU(cudaStreamWaitEvent(c->copy_stream, s.consumed, 0));
    int rc = s.pos.ensure(soa_bytes);
    if (rc) return rc;
    s.B = B; s.M = M; s.N = N; s.Npad = Npad; s.Next = Next;

    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, beads) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
        // zero-bounce path: DMA the reference AoS array as-is into the slot's landing buffer; the transpose runs later
        // on the compute stream (materialize), so consecutive stagings are back-to-back DMAs on the copy stream
        const size_t aos_bytes = sizeof(double) * nsl * Next * nd;
        rc = s.aos.ensure(aos_bytes);
        if (rc) return rc;
        CU(cudaMemcpyAsync(s.aos.p, beads, aos_bytes, cudaMemcpyHostToDevice, c->copy_stream));
        s.Next = Next;
        s.needs_transpose = true;
        CU(cudaEventRecord(s.ready, c->copy_stream));
        // the caller may mutate its buffer when we return: the DMA must have consumed it
        if (wait) CU(cudaEventSynchronize(s.ready));
    } else {
        // pageable source: pack AoS -> SoA into the next pinned bounce buffer, then one H2D
        PinBuf& pb = c->pin[c->pin_next];
        c->pin_next ^= 1;
        CU(cudaEventSynchronize(pb.done));
        rc = pb.ensure(soa_bytes);
        if (rc) return rc;
        double* dst = static_cast<double*>(pb.p);
        for (size_t sl = 0; sl < nsl; ++sl) {
            const double* src = beads + sl * Next * nd;
            double* row = dst + sl * nd * Npad;
            for (int d = 0; d < nd; ++d) {
                double* o = row + static_cast<size_t>(d) * Npad;
                for (int i = 0; i < N; ++i) o[i] = src[static_cast<size_t>(i) * nd + d];
                for (int i = N; i < Npad; ++i) o[i] = 0.0;
            }
        }
        CU(cudaMemcpyAsync(s.pos.p, pb.p, soa_bytes, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(pb.done, c->copy_stream));
        CU(cudaEventRecord(s.ready, c->copy_stream));
        s.needs_transpose = false;
    }
    s.staged = true;
    s.gen = ++c->gen_counter;
    if (c->cfg_slot == slot) c->cfg_slot = -1;
    return 0;
}

// Decodes every entry of a packed table pair and compares it with the verbatim tables, bit for bit; mismatches[0] counts
// the entries that differ (must be 0), mismatches[1] the RAW sectors.
__global__ void codec_verify_kernel(const TableSector* __restrict__ sec, const double* __restrict__ F, const double* __restrict__ G,
                                    int len, CodecSteps st, unsigned long long* __restrict__ mismatches) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= len) return;
    const TableSector s = sec[k >> 2];
    if (sector_is_raw(s)) {
        if ((k & 3) == 0) atomicAdd(mismatches + 1, 1ull);
        return;
    }
    double f, g;
    const int j = k & 3;
    sector_decode<true>(s, j, st.x[j], st.xh[j], st.x3[j], f, g);
    if (__double_as_longlong(f) != __double_as_longlong(F[k]) || __double_as_longlong(g) != __double_as_longlong(G[k]))
        atomicAdd(mismatches, 1ull);
}

// Packs (F, G) on the host (several threads), uploads the sectors and verifies them on the device against the verbatim
// device tables dF / dG.  *ok = the packed table may be used; any mismatch or > 5 % RAW sectors leaves it unused.
int build_packed_table(pimcb_ctx* c, const double* F, const double* G, int len, double dr, const double* dF, const double* dG,
                       DevBuf& dst, bool* ok, long* nraw) {
    *ok = false;
    *nraw = 0;
    static const bool codec_on = !(std::getenv("PIMCB_TABLE_CODEC") && std::atoi(std::getenv("PIMCB_TABLE_CODEC")) == 0);
    if (!codec_on || len < 8) return 0;
    const int ns = (len + 3) / 4;
    std::vector<TableSector> sec(ns);
    const CodecSteps st = codec_steps(dr);
    const int nth = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (int w = 0; w < nth; ++w)
        pool.emplace_back([&, w]() {
            const int s0 = static_cast<int>(static_cast<long long>(ns) * w / nth), s1 = static_cast<int>(static_cast<long long>(ns) * (w + 1) / nth);
            for (int s = s0; s < s1; ++s) sector_encode(F, G, len, 4 * s, dr, st, sec[s]);
        });
    for (auto& t : pool) t.join();
    int rc = dst.ensure(sizeof(TableSector) * static_cast<size_t>(ns));
    if (rc) return rc;
    if ((rc = c->d_count.ensure(2 * sizeof(unsigned long long)))) return rc;
    CU(cudaMemcpyAsync(dst.p, sec.data(), sizeof(TableSector) * static_cast<size_t>(ns), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_count.p, 0, 2 * sizeof(unsigned long long), c->stream));
    codec_verify_kernel<<<(len + 255) / 256, 256, 0, c->stream>>>(dst.as<TableSector>(), dF, dG, len, st, c->d_count.as<unsigned long long>());
    CU(cudaGetLastError());
    unsigned long long res[2] = {1, 0};
    CU(cudaMemcpyAsync(res, c->d_count.p, sizeof res, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->launches++;
    *nraw = static_cast<long>(res[1]);
    *ok = res[0] == 0 && res[1] * 20 <= static_cast<unsigned long long>(ns);
    return 0;
}

// Everything the tile kernels need to turn a separation into a table index and a histogram bin without divisions
// (kernels_pair.cuh).  Returns false when the table is too long for the bit trick (never for int lengths below 2^30).
bool make_index_params(const pimcb_ctx* c, double dSep, bool want_hist, TileIndexParams* ix) {
    int ebits = 24;
    while ((1ll << (ebits - 1)) <= static_cast<long long>(c->tab_len) + 2) ++ebits;
    if (ebits > 31) return false;
    ix->len = c->tab_len; ix->dr = c->dr; ix->inv_dr = 1.0 / c->dr;
    ix->dSep = dSep; ix->want_hist = want_hist ? 1 : 0;
    ix->magic = std::ldexp(1.5, ebits); ix->fb = 52 - ebits;
    // bin = (k * round(2^S dr/dSep)) >> S with the largest S <= 56 that keeps k * mul below 2^64
    ix->hmul_lo = ix->hmul_hi = 0; ix->hshift = 32; ix->hslop = 0; ix->hspan = 0;      // hspan = 0: never safe (exact path)
    if (want_hist && dSep > 0.0) {
        const double cbin = c->dr / dSep;
        int S = 56;
        while (S >= 32 && std::ldexp(cbin, S) * std::ldexp(1.0, ebits - 1) >= 1.8e19) --S;
        const double width = std::ldexp(cbin, 32) * 1.01 + 4096.0;      // width of the interval k pins r/dSep to + rounding slop, in 2^-32 bins
        if (S >= 32 && width < 1.0e9) {
            const unsigned long long mul = static_cast<unsigned long long>(std::llround(std::ldexp(cbin, S)));
            ix->hmul_lo = static_cast<unsigned>(mul & 0xffffffffull);
            ix->hmul_hi = static_cast<unsigned>(mul >> 32);
            ix->hshift = S;
            ix->hslop = 4096u;
            ix->hspan = static_cast<unsigned>(4294967296.0 - width - 4096.0);
        }
    }
    const CodecSteps st = codec_steps(c->dr);
    for (int j = 0; j < 4; ++j) { ix->x[j] = st.x[j]; ix->xh[j] = st.xh[j]; ix->x3[j] = st.x3[j]; }
    return true;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* pimcb_last_error(void) { return g_err.c_str(); }
int pimcb_version(void) { return 100; }

int pimcb_create(pimcb_ctx** out, int device, int ndim) {
    if (!out) return fail(PIMCB_EINVAL, "null out pointer");
    *out = nullptr;
    if (ndim < 1 || ndim > 3) return fail(PIMCB_EINVAL, "ndim must be 1, 2 or 3 (got %d)", ndim);
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PIMCB_ECUDA, "device %d not available (%d CUDA devices visible)", device, ndev);
    CU(cudaSetDevice(device));
    pimcb_ctx* c = new (std::nothrow) pimcb_ctx();
    if (!c) return fail(PIMCB_ENOMEM, "out of host memory");
    c->device = device;
    c->ndim = ndim;
    if (const char* e = std::getenv("PIMCB_LATTICE_J")) c->lattice_J = std::atoi(e);
    if (const char* e = std::getenv("PIMCB_CORR_MODE")) c->corr_mode = std::max(0, std::min(2, std::atoi(e)));
    if (const char* e = std::getenv("PIMCB_GRAPH")) c->use_graph = std::atoi(e) ? 1 : 0;
    if (const char* e = std::getenv("PIMCB_LATTICE_WARPS")) {
        const int w = std::atoi(e);
        if (w == 2 || w == 4) c->lattice_warps = w;
    }
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < kKernels; ++k) {
        CU(cudaEventCreate(&c->ev0[k]));
        CU(cudaEventCreate(&c->ev1[k]));
    }
    for (auto& s : c->slots) {
        CU(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
    }
    for (auto& p : c->pin) {
        CU(cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming));
        CU(cudaEventRecord(p.done, c->copy_stream));
    }
    if (int rc = ensure_sched(c, 2)) { pimcb_destroy(c); return rc; }
    *out = c;
    return 0;
}

int pimcb_destroy(pimcb_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    for (auto& s : c->slots) {
        s.pos.release();
        s.aos.release();
        if (s.ready) cudaEventDestroy(s.ready);
        if (s.consumed) cudaEventDestroy(s.consumed);
    }
    for (auto& p : c->pin) { p.release(); if (p.done) cudaEventDestroy(p.done); }
    c->h_out.release();
    c->h_cnt.release();
    if (c->comm_stream) { cudaStreamSynchronize(c->comm_stream); cudaStreamDestroy(c->comm_stream); c->comm_stream = nullptr; }
    if (c->xchg_ready) cudaEventDestroy(c->xchg_ready);
    if (c->xchg_done) cudaEventDestroy(c->xchg_done);
    c->xchg_ready = c->xchg_done = nullptr;
    c->d_xchg.release();
    c->h_xchg.release();
    for (DevBuf* b : {&c->d_q, &c->d_comm, &c->d_qn, &c->d_qidx, &c->d_plan, &c->d_rho, &c->d_cfg, &c->d_bins, &c->d_partial,
                      &c->d_V, &c->d_dV, &c->d_VD, &c->d_DD, &c->d_vint, &c->d_f2, &c->d_hist, &c->d_scratch, &c->d_sched, &c->d_binrows, &c->d_unfold,
                      &c->d_var, &c->d_inside, &c->d_d2V, &c->d_delta_aos, &c->d_delta, &c->d_vir, &c->d_gext, &c->d_g2ext, &c->d_gather, &c->d_count})
        b->release();
    for (int k = 0; k < kKernels; ++k) { cudaEventDestroy(c->ev0[k]); cudaEventDestroy(c->ev1[k]); }
    for (auto& r : c->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->fused.exec) cudaGraphExecDestroy(c->fused.exec);
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->copy_stream);
    delete c;
    return 0;
}

int pimcb_set_box(pimcb_ctx* c, const double* side, const unsigned* periodic) {
    if (!c || !side) return fail(PIMCB_EINVAL, "null argument");
    for (int d = 0; d < c->ndim; ++d) {
        if (!(side[d] > 0.0)) return fail(PIMCB_EINVAL, "side[%d] = %g must be positive", d, side[d]);
        c->side[d] = side[d];
        c->periodic[d] = periodic ? periodic[d] : 1u;
        c->box.sideInv[d] = 1.0 / side[d];                  // src/container.cpp:122
        c->box.pSide[d] = c->periodic[d] * side[d];         // src/container.cpp:129
    }
    c->have_box = true;
    if (c->nq > 0) {   // re-classify against the new box
        std::vector<double> q = c->q_host;
        return pimcb_set_qvecs(c, q.data(), c->nq);
    }
    return 0;
}

int pimcb_set_qvecs(pimcb_ctx* c, const double* q, int nq) {
    if (!c || !q || nq < 1) return fail(PIMCB_EINVAL, "bad q-vector arguments");
    if (!c->have_box) return fail(PIMCB_ESTATE, "pimcb_set_box must precede pimcb_set_qvecs");
    CU(cudaSetDevice(c->device));
    const int nd = c->ndim;
    c->nq = nq;
    c->q_host.assign(q, q + static_cast<size_t>(nq) * nd);
    c->comm.assign(nq, 0);
    c->qn.assign(static_cast<size_t>(nq) * nd, 0);
    std::vector<int> sel;
    std::vector<double> qsoa(static_cast<size_t>(nq) * nd);
    c->ncomm = 0;
    c->max_phase = 0.0;
    for (int d = 0; d < 3; ++d) c->nmax[d] = 0;
    for (int k = 0; k < nq; ++k) {
        bool ok = true;
        double phase = 0.0;
        for (int d = 0; d < nd; ++d) {
            const double v = q[static_cast<size_t>(k) * nd + d];
            if (!std::isfinite(v)) return fail(PIMCB_EINVAL, "q[%d][%d] is not finite", k, d);
            qsoa[static_cast<size_t>(d) * nq + k] = v;
            phase += std::fabs(v) * c->side[d];             // |x_d| may reach a few box lengths for unwrapped beads
            if (c->periodic[d]) {
                const double n = v * c->side[d] / (2.0 * M_PI);
                const double rn = std::nearbyint(n);
                if (std::fabs(n - rn) > 1e-9 * std::max(1.0, std::fabs(rn)) || std::fabs(rn) > 4096.0) ok = false;
                c->qn[static_cast<size_t>(k) * nd + d] = static_cast<int>(rn);
            } else if (v != 0.0) {
                ok = false;
            }
        }
        c->max_phase = std::max(c->max_phase, 4.0 * phase);
        c->comm[k] = ok;
        if (ok) c->ncomm++; else sel.push_back(k);
    }
    if (c->ncomm == nq)
        for (int k = 0; k < nq; ++k)
            for (int d = 0; d < nd; ++d) c->nmax[d] = std::max(c->nmax[d], std::abs(c->qn[static_cast<size_t>(k) * nd + d]));
    c->nsel = static_cast<int>(sel.size());
    int rc;
    // Lattice-path plan: sign-symmetry groups (key = (|n_0|,..,|n_{nd-1}|), member slot = sign pattern), stored
    // column by column (column = leading nd-1 key components), cut into tasks and scheduled on the CTA's warps (LPT).
    c->ngroups = 0;
    c->mma_nL = c->mma_nR = 0;
    if (c->ncomm == nq) {
        const int npat = 1 << nd;
        struct Grp { int key[3]; std::vector<int> out; };
        std::vector<Grp> groups;
        for (int k = 0; k < nq; ++k) {
            int key[3] = {0, 0, 0}, pat = 0;
            for (int d = 0; d < nd; ++d) {
                const int n = c->qn[static_cast<size_t>(k) * nd + d];
                key[d] = std::abs(n);
                if (n < 0) pat |= 1 << d;
            }
            Grp* hit = nullptr;
            for (Grp& g : groups)
                if (g.key[0] == key[0] && g.key[1] == key[1] && g.key[2] == key[2] && g.out[pat] < 0) { hit = &g; break; }
            if (!hit) {                                   // a repeated q opens a new group with the same key
                groups.push_back(Grp{{key[0], key[1], key[2]}, std::vector<int>(npat, -1)});
                hit = &groups.back();
            }
            hit->out[pat] = k;
        }
        std::stable_sort(groups.begin(), groups.end(), [](const Grp& a, const Grp& b) {
            return std::lexicographical_compare(a.key, a.key + 3, b.key, b.key + 3);
        });
        const int G = static_cast<int>(groups.size());
        auto col_of = [&](const Grp& g, int which) { return which < nd - 1 ? g.key[which] : 0; };
        auto same_col = [&](const Grp& a, const Grp& b) { return col_of(a, 0) == col_of(b, 0) && col_of(a, 1) == col_of(b, 1); };
        const int last = nd - 1;
        // candidate run lengths: pick the one whose LPT makespan is smallest
        int maxlen = 1;
        for (int g0 = 0; g0 < G;) {
            int g1 = g0 + 1;
            while (g1 < G && same_col(groups[g0], groups[g1])) ++g1;
            maxlen = std::max(maxlen, g1 - g0);
            g0 = g1;
        }
        std::vector<int> best_tasks, best_first;
        double best_span = 1e300;
        for (int emax = 1; emax <= maxlen; ++emax) {
            std::vector<std::array<int, 4>> tasks;
            std::vector<double> cost;
            for (int g0 = 0; g0 < G;) {
                int g1 = g0 + 1;
                while (g1 < G && same_col(groups[g0], groups[g1])) ++g1;
                for (int a = g0; a < g1; a += emax) {
                    const int b = std::min(g1, a + emax);
                    double cst = nd == 3 ? 0.8 : (nd == 2 ? 0.3 : 0.0);
                    for (int g = a; g < b; ++g) cst += (groups[g].key[last] == 0 && nd > 1) ? 0.6 : 1.0;
                    tasks.push_back({col_of(groups[g0], 0), col_of(groups[g0], 1), a, b});
                    cost.push_back(cst);
                }
                g0 = g1;
            }
            std::vector<int> order(tasks.size());
            for (size_t i = 0; i < order.size(); ++i) order[i] = static_cast<int>(i);
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
            const int W = c->lattice_warps;
            double load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            std::vector<int> owner(tasks.size());
            for (int t : order) {
                int w = 0;
                for (int k = 1; k < W; ++k) if (load[k] < load[w]) w = k;
                load[w] += cost[t];
                owner[t] = w;
            }
            const double span = *std::max_element(load, load + W);
            if (span < best_span - 1e-9) {
                best_span = span;
                best_tasks.clear();
                best_first.assign(9, 0);
                for (int w = 0; w < 8; ++w) {
                    best_first[w] = static_cast<int>(best_tasks.size() / 4);
                    for (size_t t = 0; t < tasks.size(); ++t)
                        if (owner[t] == w && w < W) best_tasks.insert(best_tasks.end(), tasks[t].begin(), tasks[t].end());
                }
                best_first[8] = static_cast<int>(best_tasks.size() / 4);
            }
        }
        std::vector<int> plan;
        auto align4 = [&plan]() { while (plan.size() % 4) plan.push_back(0); };
        c->plan_off[0] = plan.size();
        for (const Grp& g : groups) plan.insert(plan.end(), g.out.begin(), g.out.end());
        align4();
        c->plan_off[1] = plan.size();
        for (const Grp& g : groups) plan.push_back(g.key[last]);
        align4();
        c->plan_off[2] = plan.size();
        plan.insert(plan.end(), best_tasks.begin(), best_tasks.end());
        align4();
        c->plan_off[3] = plan.size();
        plan.insert(plan.end(), best_first.begin(), best_first.end());
        // DMMA formulation: L rows per (leading-key) column, R columns per last-key value; coinciding factors are
        // stored once (see kernels.cuh, MmaPlan) and one all-zero row / column is reserved at the end
        // The 3-D kernel indexes lmap with the fixed stride 9 (|n_x|, |n_y| <= 8): q-sets beyond that have no DMMA plan
        // (mma_nL = mma_nR = 0) and run on the CUDA-core lattice / generic kernels.
        bool dmma_plan = !(nd == 3 && (c->nmax[0] > 8 || c->nmax[1] > 8)) && c->nmax[last] <= 16;
        // ... and every group key is checked against the tables it will index (3-D: lmap[|n_x| * 9 + |n_y|] of (nmax_x + 1) * 9
        // entries, rmap[|n_last|] of nmax_last + 1): a key outside them drops the DMMA plan instead of writing past a vector
        // (round-1 advisor finding: |n_y| >= 9 used to corrupt the heap here)
        for (const Grp& g : groups) {
            if (g.key[last] > c->nmax[last] || (nd > 1 && g.key[0] > c->nmax[0]) || (nd == 3 && g.key[1] > 8)) dmma_plan = false;
        }
        c->mma_lmap.clear(); c->mma_rmap.clear(); c->mma_gdesc.clear(); c->mma_gout.clear();
        c->unfold_NR = -1;
        if (dmma_plan) {
            const int n0 = c->nmax[0] + 1, n1 = nd == 3 ? 9 : 1;   // 3-D lmap has the fixed stride 9 the kernel indexes with
            std::vector<int> lmap(nd == 1 ? 1 : static_cast<size_t>(n0) * n1, -1), rmap(c->nmax[last] + 1, -1);
            int nL = nd == 1 ? 1 : 0, nR = 0;
            if (nd == 1) lmap[0] = 0;
            for (const Grp& g : groups) {
                const size_t ri = static_cast<size_t>(g.key[last]);
                const size_t li = nd == 3 ? static_cast<size_t>(g.key[0]) * n1 + g.key[1] : static_cast<size_t>(nd > 1 ? g.key[0] : 0);
                int& r = rmap[ri];
                if (r < 0) { r = nR; nR += (g.key[last] > 0 || nd == 1) ? 2 : 1; }
                if (nd > 1) {
                    int& l = lmap[li];
                    if (l < 0) {
                        l = nL;
                        if (nd == 3) nL += (g.key[0] > 0 && g.key[1] > 0) ? 4 : ((g.key[0] > 0 || g.key[1] > 0) ? 2 : 1);
                        else nL += g.key[0] > 0 ? 2 : 1;
                    }
                }
            }
            const int zrow = nL++, zcol = nR++;              // reserved zero planes
            std::vector<int> gdesc;
            for (const Grp& g : groups) {
                const int c0 = rmap[g.key[last]];
                const bool czero = g.key[last] == 0 && nd > 1;
                int rrp = 0, rip = zrow, rrm = 0, rim = zrow, sgn = 1;
                if (nd == 3) {
                    const int a = g.key[0], b = g.key[1], l = lmap[a * n1 + b];
                    if (a > 0 && b > 0) { rrp = l; rip = l + 1; rrm = l + 2; rim = l + 3; }
                    else if (a > 0) { rrp = rrm = l; rip = rim = l + 1; }
                    else if (b > 0) { rrp = rrm = l; rip = rim = l + 1; sgn = -1; }       // X conj(Y) = conj(Y^b)
                    else { rrp = rrm = l; rip = rim = zrow; }
                } else if (nd == 2) {
                    const int a = g.key[0], l = lmap[a];
                    rrp = rrm = l;
                    rip = rim = a > 0 ? l + 1 : zrow;
                } else {
                    rrp = rrm = 0;                            // the row of ones
                }
                const int e[8] = {rrp, rip, rrm, rim, sgn, c0, czero ? zcol : c0 + 1, 0};
                gdesc.insert(gdesc.end(), e, e + 8);
            }
            c->mma_nL = nL;
            c->mma_nR = nR;
            c->mma_lmap = lmap;
            c->mma_rmap = rmap;
            c->mma_gdesc = gdesc;
            c->mma_gout.clear();
            for (const Grp& g : groups) c->mma_gout.insert(c->mma_gout.end(), g.out.begin(), g.out.end());
            c->unfold_NR = -1;
            align4();
            c->plan_off[4] = plan.size();
            plan.insert(plan.end(), gdesc.begin(), gdesc.end());
        }
        c->ngroups = G;
        if ((rc = c->d_plan.ensure(sizeof(int) * plan.size()))) return rc;
        CU(cudaMemcpyAsync(c->d_plan.p, plan.data(), sizeof(int) * plan.size(), cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));   // `plan` is a local
    }
    if ((rc = c->d_q.ensure(sizeof(double) * qsoa.size()))) return rc;
    if ((rc = c->d_comm.ensure(nq))) return rc;
    if ((rc = c->d_qn.ensure(sizeof(int) * c->qn.size()))) return rc;
    if ((rc = c->d_qidx.ensure(sizeof(int) * std::max<size_t>(1, sel.size())))) return rc;
    CU(cudaMemcpyAsync(c->d_q.p, qsoa.data(), sizeof(double) * qsoa.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_comm.p, c->comm.data(), nq, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_qn.p, c->qn.data(), sizeof(int) * c->qn.size(), cudaMemcpyHostToDevice, c->stream));
    if (!sel.empty()) CU(cudaMemcpyAsync(c->d_qidx.p, sel.data(), sizeof(int) * sel.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->bins_len = 0;   // layout changed: bins are re-created on the next measurement
    c->cfg_slot = -1;
    c->qgen++;
    return 0;
}

int pimcb_num_commensurate(const pimcb_ctx* c) { return c ? c->ncomm : PIMCB_EINVAL; }

int pimcb_set_rho_mode(pimcb_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 2) return fail(PIMCB_EINVAL, "rho mode must be 0, 1 or 2");
    c->rho_mode = mode;
    c->cfg_slot = -1;
    return 0;
}

int pimcb_set_corr_mode(pimcb_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 2) return fail(PIMCB_EINVAL, "corr mode must be 0, 1 or 2");
    c->corr_mode = mode;
    c->cfg_slot = -1;
    return 0;
}

int pimcb_num_slots(const pimcb_ctx*) { return kSlots; }

int pimcb_stage_batch_slot(pimcb_ctx* c, int slot, const double* beads, int B, int M, int N, int Next) {
    return stage_into(c, slot, beads, B, M, N, Next);
}

int pimcb_select_slot(pimcb_ctx* c, int slot) {
    if (!c || slot < 0 || slot >= kSlots) return fail(PIMCB_EINVAL, "slot out of range");
    if (!c->slots[slot].staged) return fail(PIMCB_ESTATE, "slot %d has no staged beads", slot);
    c->cur = slot;
    return 0;
}

int pimcb_stage_batch(pimcb_ctx* c, const double* beads, int B, int M, int N, int Next) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    const int slot = (c->cur + 1 + kSlots) % kSlots;
    int rc = stage_into(c, slot, beads, B, M, N, Next);
    if (rc) return rc;
    c->cur = slot;
    return 0;
}

int pimcb_stage_batch_async(pimcb_ctx* c, const double* beads, int B, int M, int N, int Next) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    const int slot = (c->cur + 1 + kSlots) % kSlots;
    int rc = stage_into(c, slot, beads, B, M, N, Next, false);
    if (rc) return rc;
    c->cur = slot;
    return 0;
}

int pimcb_stage_wait(pimcb_ctx* c) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->copy_stream));
    return 0;
}

int pimcb_stage_beads(pimcb_ctx* c, const double* beads, int M, int N, int Next) {
    return pimcb_stage_batch(c, beads, 1, M, N, Next);
}

int pimcb_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(PIMCB_EINVAL, "null pointer");
    CU(cudaMallocHost(ptr, bytes));
    return 0;
}
int pimcb_host_free(void* ptr) { CU(cudaFreeHost(ptr)); return 0; }
int pimcb_host_register(void* ptr, size_t bytes) { CU(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault)); return 0; }
int pimcb_host_unregister(void* ptr) { CU(cudaHostUnregister(ptr)); return 0; }

int pimcb_ssf_isf(pimcb_ctx* c, double* ssf_out, double* isf_out) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    CU(cudaSetDevice(c->device));
    Slot* s;
    int rc = run_estimators(c, &s);
    if (rc) return rc;
    return copy_out(c, *s, ssf_out, isf_out);
}
// Stage + evaluate + read back with ONE synchronisation: the DMA of a page-locked source is not waited for on its own,
// the final stream synchronisation of the read-back covers it (the compute stream waits for the slot's ready event).
//
// Graph path.  An estimator's accumulate() calls this once per measurement with the same page-locked source
// (Path::beads), the same shape and the same q-set, so after one ordinary call (which sizes every buffer) the fixed
// sequence  [H2D(AoS) -> aos_to_soa ->] rho_q build -> tau-correlation -> D2H  (the bracketed nodes disappear when the rho
// kernel can read the page-locked array in place, see fused_capture)  is captured from the compute stream into
// a CUDA graph and replayed: one cudaGraphLaunch + one synchronisation per measurement instead of five enqueues on two
// streams with an event between them.  Every address baked into the graph is re-checked before a replay; anything that
// changes (source pointer, shape, q-set, kernel mode, a buffer that had to grow) drops the graph and the ordinary path
// takes over until the next capture.  Not used with profiling on, with non-commensurate q (an extra kernel pair) or with
// pageable sources (the host-side repacking is not stream work).
namespace {

void fused_addresses(pimcb_ctx* c, const Slot& s, const void** a) {
    a[0] = s.pos.p; a[1] = s.aos.p; a[2] = c->d_rho.p; a[3] = c->d_cfg.p; a[4] = c->h_out.p; a[5] = c->d_unfold.p;
    a[6] = c->d_sched.p; a[7] = c->d_partial.p;
}

void fused_drop(pimcb_ctx* c) {
    if (c->fused.exec) cudaGraphExecDestroy(c->fused.exec);
    const int failed = c->fused.failed;
    c->fused = pimcb_ctx::FusedGraph{};
    c->fused.failed = failed;
}

int fused_finish(pimcb_ctx* c, Slot& s, int slot, double* ssf_out, double* isf_out) {
    CU(cudaStreamSynchronize(c->stream));
    const size_t len = static_cast<size_t>(c->nq) * (1 + s.M);
    const double* h = static_cast<const double*>(c->h_out.p);
    if (ssf_out) std::memcpy(ssf_out, h, sizeof(double) * c->nq);
    if (isf_out) std::memcpy(isf_out, h + c->nq, sizeof(double) * (len - c->nq));
    // the slot now holds exactly the graph's configuration: shape fields included (the rotating stagers may have put
    // another shape with the same buffers into this slot in between)
    s.B = 1; s.M = c->fused.M; s.N = c->fused.N; s.Npad = round_up(c->fused.N, 16); s.Next = c->fused.Next;
    s.staged = true;
    s.needs_transpose = false;
    s.gen = ++c->gen_counter;
    c->cur = slot;
    c->cfg_slot = slot;
    c->cfg_gen = s.gen;
    return 0;
}

// Captures and instantiates the graph for (beads, M, N, Next) into slot kSlots-1, then launches it once.
// Returns 1 when the capture could not be made (the caller falls back to the ordinary path), 0 on success, < 0 on error.
int fused_capture(pimcb_ctx* c, const double* beads, int M, int N, int Next, double* ssf_out, double* isf_out) {
    const int slot = kSlots - 1, nd = c->ndim;
    Slot& s = c->slots[slot];
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->stream));
    const int Npad = round_up(N, 16);
    const size_t aos_bytes = sizeof(double) * static_cast<size_t>(M) * Next * nd;
    const size_t len = static_cast<size_t>(c->nq) * (1 + M);
    int rc;
    if ((rc = s.pos.ensure(sizeof(double) * static_cast<size_t>(M) * nd * Npad))) return rc;
    if ((rc = s.aos.ensure(aos_bytes))) return rc;
    if ((rc = c->d_cfg.ensure(sizeof(double) * len))) return rc;
    if ((rc = c->h_out.ensure(sizeof(double) * len))) return rc;
    s.B = 1; s.M = M; s.N = N; s.Npad = Npad; s.Next = Next;
    const long launches0 = c->launches;
    bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
        // Small configurations: the transpose kernel reads the page-locked source over the host link itself (zero copy)
        // instead of a DMA into the landing buffer followed by the transpose -- one node and one dependency less.
        // Measured per call (tools/latency_ab.py, profiles/r01zc_zerocopy.txt): C1 (55 KB) 48 -> 37.5 us, C2 (1.03 MB)
        // 79.6 -> 74.7 us, C4 (7.7 MB) 320 -> 337 us: SM-issued reads do not reach the DMA engine's link rate on large
        // arrays, so the zero-copy form is used up to 2 MB.  PIMCB_ZEROCOPY=0 / 1 forces it off / on.
        static const int zc_env = std::getenv("PIMCB_ZEROCOPY") ? std::atoi(std::getenv("PIMCB_ZEROCOPY")) : -1;
        const bool zerocopy = zc_env < 0 ? aos_bytes <= (2u << 20) : zc_env != 0;
        // ... and with the DMMA rho_q kernel not even the transpose kernel is needed: the rho kernel reads the beads array
        // where it lies and writes the transposed copy (for the pair / virial kernels that may follow) on the way.
        // PIMCB_RHO_DIRECT=0 keeps the transpose node (A/B).
        static const int direct_env = std::getenv("PIMCB_RHO_DIRECT") ? std::atoi(std::getenv("PIMCB_RHO_DIRECT")) : 1;
        void* mapped = nullptr;
        const double* rho_src = nullptr;
        // (measured, C2 / C3 / C4: 78.4 -> 70.1, 118.8 -> 104.6, 318 -> 284 us per call: the reads over the link overlap the
        // kernel's own arithmetic, so this form also wins beyond the 2 MB where the transpose-only zero-copy form stops.
        // The mirror image on the output side -- the correlation kernel writing its rows straight into the page-locked
        // read-back buffer instead of a D2H copy node -- was measured and LOSES: 75.3 vs 70.1 us for C2, 171 vs 105 us for
        // C3; scattered 8-byte posted writes over the link are far slower than one 87 KB copy.)
        const bool direct_ok = zc_env < 0 ? true : zc_env != 0;
        if (direct_ok && direct_env != 0 && rho_mma_shape(c, nullptr, nullptr, nullptr) &&
            cudaHostGetDevicePointer(&mapped, const_cast<double*>(beads), 0) == cudaSuccess && mapped) {
            rho_src = static_cast<const double*>(mapped);
            s.needs_transpose = false;
        } else if (zerocopy && cudaHostGetDevicePointer(&mapped, const_cast<double*>(beads), 0) == cudaSuccess && mapped) {
            const size_t tsmem = sizeof(double) * N * nd;
            const int nslc = M;
            const int tgrid = std::min(nslc, c->sm_count * 8);
            const double* src = static_cast<const double*>(mapped);
            if (nd == 1) { set_smem(aos_to_soa_kernel<1>, tsmem); aos_to_soa_kernel<1><<<tgrid, 256, tsmem, c->stream>>>(src, s.pos.as<double>(), nslc, N, Next, Npad); }
            else if (nd == 2) { set_smem(aos_to_soa_kernel<2>, tsmem); aos_to_soa_kernel<2><<<tgrid, 256, tsmem, c->stream>>>(src, s.pos.as<double>(), nslc, N, Next, Npad); }
            else { set_smem(aos_to_soa_kernel<3>, tsmem); aos_to_soa_kernel<3><<<tgrid, 256, tsmem, c->stream>>>(src, s.pos.as<double>(), nslc, N, Next, Npad); }
            ok = cudaGetLastError() == cudaSuccess;
            c->launches++;
            s.needs_transpose = false;
        } else {
            cudaGetLastError();
            ok = cudaMemcpyAsync(s.aos.p, beads, aos_bytes, cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
            s.needs_transpose = true;
        }
        ok = ok && materialize(c, s) == 0;
        ok = ok && launch_rho(c, s, rho_src) == 0;
        ok = ok && launch_corr(c, s) == 0;
        ok = ok && cudaMemcpyAsync(c->h_out.p, c->d_cfg.p, sizeof(double) * len, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess;
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        ok = ok && e == cudaSuccess && graph != nullptr;
        if (ok) ok = cudaGraphInstantiate(&c->fused.exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
    }
    cudaGetLastError();
    if (!ok) {
        c->launches = launches0;
        fused_drop(c);
        c->fused.failed++;
        s.staged = false;
        return 1;
    }
    pimcb_ctx::FusedGraph& g = c->fused;
    g.src = beads; g.M = M; g.N = N; g.Next = Next; g.slot = slot; g.qgen = c->qgen; g.rho_mode = c->rho_mode; g.corr_mode = c->corr_mode;
    g.launches = c->launches - launches0;           // kernels per replay ([transpose,] rho_q, tau-correlation)
    fused_addresses(c, s, g.baked);
    CU(cudaGraphLaunch(g.exec, c->stream));
    return fused_finish(c, s, slot, ssf_out, isf_out);
}

}  // namespace

int pimcb_ssf_isf_beads(pimcb_ctx* c, const double* beads, int M, int N, int Next, double* ssf_out, double* isf_out) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    if (c->use_graph && !c->profiling && c->fused.failed < 3 && c->nq > 0 && c->nsel == 0 && c->have_box && beads && M > 0 && N > 0 &&
        Next >= N && c->max_phase <= 1.0e5) {
        CU(cudaSetDevice(c->device));
        pimcb_ctx::FusedGraph& g = c->fused;
        const bool same = g.src == beads && g.M == M && g.N == N && g.Next == Next && g.qgen == c->qgen && g.rho_mode == c->rho_mode &&
                          g.corr_mode == c->corr_mode;
        if (same && g.exec) {
            const void* now[8];
            fused_addresses(c, c->slots[g.slot], now);
            if (std::memcmp(now, g.baked, sizeof now) == 0) {
                // a pimcb_stage_batch_async into the graph's slot may still be copying on the copy stream
                CU(cudaStreamWaitEvent(c->stream, c->slots[g.slot].ready, 0));
                CU(cudaGraphLaunch(g.exec, c->stream));
                c->launches += g.launches;
                return fused_finish(c, c->slots[g.slot], g.slot, ssf_out, isf_out);
            }
            fused_drop(c);
        } else if (same && g.primed) {
            cudaPointerAttributes attr{};
            const bool pinned = cudaPointerGetAttributes(&attr, beads) == cudaSuccess && attr.type == cudaMemoryTypeHost;
            cudaGetLastError();
            if (pinned) {
                const int r = fused_capture(c, beads, M, N, Next, ssf_out, isf_out);
                if (r <= 0) return r;
            } else {
                g.primed = 0;
            }
        } else {
            fused_drop(c);                       // a new key: the ordinary call below sizes every buffer, the next call captures
            g.src = beads; g.M = M; g.N = N; g.Next = Next; g.qgen = c->qgen; g.rho_mode = c->rho_mode; g.corr_mode = c->corr_mode;
            g.primed = 1;
        }
    }
    const int slot = (c->cur + 1 + kSlots) % kSlots;
    int rc = stage_into(c, slot, beads, 1, M, N, Next, false);
    if (rc) return rc;
    c->cur = slot;
    return pimcb_ssf_isf(c, ssf_out, isf_out);
}
int pimcb_ssf(pimcb_ctx* c, double* out) { return pimcb_ssf_isf(c, out, nullptr); }
int pimcb_isf(pimcb_ctx* c, double* out) { return pimcb_ssf_isf(c, nullptr, out); }

// The global configuration count of a pimcb_reduce_bins that did not wait for it: valid once the ctx stream has passed the
// copy.  `synced` = the caller has just synchronised the stream.
static int settle_count(pimcb_ctx* c, bool synced) {
    if (!c->n_acc_pending) return 0;
    if (!synced) CU(cudaStreamSynchronize(c->stream));
    c->n_acc = static_cast<long>(*static_cast<const long long*>(c->h_cnt.p));
    c->n_acc_pending = false;
    return 0;
}

int pimcb_measure(pimcb_ctx* c) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    CU(cudaSetDevice(c->device));
    Slot* s;
    int rows = 0;
    int rc = settle_count(c, false);        // accumulating on top of a reduced bin: its count first
    if (rc) return rc;
    rc = run_estimators(c, &s, &rows);
    if (rc) return rc;
    if (rows > 0) {
        KTimer kt(c, K_BINS);
        const size_t len = c->bins_len;
        bins_accumulate_kernel<<<static_cast<unsigned>((kBinLanes * len + 255) / 256), 256, 0, c->stream>>>(c->d_cfg.as<double>(), c->d_bins.as<double>(), rows, len);
        CU(cudaGetLastError());
    }
    c->n_acc += s->B;
    return 0;
}

// Creates the (zeroed) bin for M time slices before anything was measured, so that a rank whose share of a walker batch is
// empty still takes part in pimcb_reduce_bins / pimcb_gather_bins_q with a zero contribution instead of leaving its peers
// waiting inside the collective.
int pimcb_init_bins(pimcb_ctx* c, int M) {
    if (!c || M < 1) return fail(PIMCB_EINVAL, "bad arguments");
    if (c->nq <= 0) return fail(PIMCB_ESTATE, "no q-vectors set");
    CU(cudaSetDevice(c->device));
    const size_t len = static_cast<size_t>(c->nq) * (1 + M);
    if (c->bins_len == len && c->bins_M == M) return 0;             // already laid out (possibly holding sums)
    int rc = c->d_bins.ensure(sizeof(double) * len);
    if (rc) return rc;
    CU(cudaMemsetAsync(c->d_bins.p, 0, sizeof(double) * len, c->stream));
    c->bins_len = len;
    c->bins_M = M;
    c->n_acc = 0;
    c->binrows_n = 0;
    c->binrows_cap = 0;
    return 0;
}

int pimcb_reset_bins(pimcb_ctx* c) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    if (c->bins_len) CU(cudaMemsetAsync(c->d_bins.p, 0, sizeof(double) * c->bins_len, c->stream));
    if (c->binrows_n > 0) {
        CU(cudaMemsetAsync(c->d_binrows.p, 0, c->d_binrows.cap, c->stream));
        c->binrows_n = 0;
    }
    c->n_acc = 0;
    c->n_acc_pending = false;
    return 0;
}

int pimcb_read_bins(pimcb_ctx* c, double* ssf, double* isf, long* num_acc) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    if (!c->bins_len) return fail(PIMCB_ESTATE, "no measurement accumulated yet");
    CU(cudaSetDevice(c->device));
    const size_t bytes = sizeof(double) * c->bins_len;
    int rc = c->h_out.ensure(bytes);
    if (rc) return rc;
    if ((rc = fold_binrows(c, c->bins_M))) return rc;
    CU(cudaMemcpyAsync(c->h_out.p, c->d_bins.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const double* h = static_cast<const double*>(c->h_out.p);
    if (ssf) std::memcpy(ssf, h, sizeof(double) * c->nq);
    if (isf) std::memcpy(isf, h + c->nq, bytes - sizeof(double) * c->nq);
    if ((rc = settle_count(c, true))) return rc;
    if (num_acc) *num_acc = c->n_acc;
    return 0;
}

int pimcb_bins_device_ptr(pimcb_ctx* c, void** dptr, size_t* count) {
    if (!c || !dptr) return fail(PIMCB_EINVAL, "null argument");
    if (!c->bins_len) return fail(PIMCB_ESTATE, "no measurement accumulated yet");
    CU(cudaSetDevice(c->device));
    if (int rc = fold_binrows(c, c->bins_M)) return rc;      // enqueued on the ctx stream, like every consumer of the bin
    *dptr = c->d_bins.p;
    if (count) *count = c->bins_len;
    return 0;
}

int pimcb_sync(pimcb_ctx* c) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int pimcb_stream(pimcb_ctx* c, void** stream) {
    if (!c || !stream) return fail(PIMCB_EINVAL, "null argument");
    *stream = c->stream;
    return 0;
}

int pimcb_set_pair_table(pimcb_ctx* c, const double* V, const double* dVdr, int len, double dr, const double* extV, const double* extdVdr) {
    if (!c || !V || len < 1 || !(dr > 0.0)) return fail(PIMCB_EINVAL, "bad pair-table arguments");
    CU(cudaSetDevice(c->device));
    int rc;
    c->tab_len = len;
    c->dr = dr;
    c->extV[0] = extV ? extV[0] : 0.0; c->extV[1] = extV ? extV[1] : 0.0;
    c->extdV[0] = extdVdr ? extdVdr[0] : 0.0; c->extdV[1] = extdVdr ? extdVdr[1] : 0.0;
    // Device tables have len + 1 entries: entry 0 holds ext[0] and entry `len` holds ext[1].  TabulatedPotential::direct
    // (include/potential.h:249-260) returns ext[0] for k <= 0 and ext[1] for k >= len and never reads entry 0, so
    // table[min(max(k, 0), len)] is the same function without a branch; entries 1 .. len-1 are the caller's, verbatim.
    std::vector<double> hV(V, V + len);
    hV.push_back(c->extV[1]);
    hV[0] = c->extV[0];
    const size_t bytes = sizeof(double) * hV.size();
    if ((rc = c->d_V.ensure(bytes))) return rc;
    CU(cudaMemcpyAsync(c->d_V.p, hV.data(), bytes, cudaMemcpyHostToDevice, c->stream));
    c->have_dV = dVdr != nullptr;
    c->h_dV.clear();
    if (dVdr) {
        c->h_dV.assign(dVdr, dVdr + len);
        c->h_dV.push_back(c->extdV[1]);
        c->h_dV[0] = c->extdV[0];
        if ((rc = c->d_dV.ensure(bytes))) return rc;
        CU(cudaMemcpyAsync(c->d_dV.p, c->h_dV.data(), bytes, cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    // one-sector-per-pair form of (V, dV/dr): packed on the host, verified bit for bit on the device (table_codec.h)
    c->vd_ok = c->dd_ok = false;
    c->have_d2V = false;
    if (dVdr && (rc = build_packed_table(c, hV.data(), c->h_dV.data(), len + 1, dr, c->d_V.as<double>(), c->d_dV.as<double>(), c->d_VD,
                                         &c->vd_ok, &c->vd_raw)))
        return rc;
    return 0;
}

int pimcb_table_codec_info(const pimcb_ctx* c, long* info) {
    if (!c || !info) return fail(PIMCB_EINVAL, "null argument");
    info[0] = c->vd_ok ? 1 : 0; info[1] = c->vd_raw; info[2] = c->dd_ok ? 1 : 0; info[3] = c->dd_raw;
    info[4] = (c->tab_len + 1 + 3) / 4;
    return 0;
}

// Uploads per-bead vectors given in the staged beads' own AoS shape ([B][M][N_ext][ndim]) into slice rows like pos.
static int upload_bead_vectors(pimcb_ctx* c, Slot* s, const double* aos, DevBuf& dst) {
    const int nd = c->ndim, nsl = s->B * s->M;
    const int Next = s->Next > 0 ? s->Next : s->N;
    const size_t aos_bytes = sizeof(double) * static_cast<size_t>(nsl) * Next * nd;
    int rc;
    if ((rc = c->d_delta_aos.ensure(aos_bytes))) return rc;
    if ((rc = dst.ensure(sizeof(double) * static_cast<size_t>(nsl) * nd * s->Npad))) return rc;
    CU(cudaMemcpyAsync(c->d_delta_aos.p, aos, aos_bytes, cudaMemcpyHostToDevice, c->stream));
    const size_t tsmem = sizeof(double) * s->N * nd;
    const int tgrid = static_cast<int>(std::min<size_t>(nsl, static_cast<size_t>(c->sm_count) * 8));
#define LAUNCH_TV(ND)                                                                                              \
    rc = set_smem(aos_to_soa_kernel<ND>, tsmem); if (rc) return rc;                                                 \
    aos_to_soa_kernel<ND><<<tgrid, 256, tsmem, c->stream>>>(c->d_delta_aos.as<double>(), dst.as<double>(), nsl, s->N, Next, s->Npad)
    {
        KTimer kt(c, K_TRANSPOSE);
        if (nd == 1) { LAUNCH_TV(1); } else if (nd == 2) { LAUNCH_TV(2); } else { LAUNCH_TV(3); }
    }
#undef LAUNCH_TV
    CU(cudaGetLastError());
    return 0;
}

int pimcb_set_external_gradient(pimcb_ctx* c, const double* gext_aos) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    if (!gext_aos) { c->gext_gen = 0; return 0; }
    CU(cudaSetDevice(c->device));
    Slot* s;
    int rc = need_cur(c, &s);
    if (rc) return rc;
    CU(cudaStreamWaitEvent(c->stream, s->ready, 0));
    if ((rc = upload_bead_vectors(c, s, gext_aos, c->d_gext))) return rc;
    CU(cudaStreamSynchronize(c->stream));          // the caller's array may go away
    c->gext_gen = s->gen;
    return 0;
}

int pimcb_set_external_laplacian(pimcb_ctx* c, const double* g2ext_aos) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    if (!g2ext_aos) { c->g2ext_gen = 0; return 0; }
    CU(cudaSetDevice(c->device));
    Slot* s;
    int rc = need_cur(c, &s);
    if (rc) return rc;
    const int nsl = s->B * s->M, Next = s->Next > 0 ? s->Next : s->N;
    // [B][M][N_ext] -> slice rows [sl][Npad] (zero padded); O(N M) values, packed on the host
    std::vector<double> rows(static_cast<size_t>(nsl) * s->Npad, 0.0);
    for (int sl = 0; sl < nsl; ++sl)
        std::memcpy(&rows[static_cast<size_t>(sl) * s->Npad], g2ext_aos + static_cast<size_t>(sl) * Next, sizeof(double) * s->N);
    if ((rc = c->d_g2ext.ensure(sizeof(double) * rows.size()))) return rc;
    CU(cudaMemcpyAsync(c->d_g2ext.p, rows.data(), sizeof(double) * rows.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));          // `rows` is a local
    c->g2ext_gen = s->gen;
    return 0;
}

int pimcb_pair_sums(pimcb_ctx* c, double* vint, double* f2, int* sephist, double dSep, int f2_parity) {
    if (!c || !vint) return fail(PIMCB_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    Slot* s;
    int rc = need_cur(c, &s);
    if (rc) return rc;
    if (!c->have_box) return fail(PIMCB_ESTATE, "no box set");
    if (!c->tab_len) return fail(PIMCB_ESTATE, "no pair table set");
    if (f2 && !c->have_dV) return fail(PIMCB_ESTATE, "gradVSquared requested but no dV/dr table set");
    if (sephist && !(dSep > 0.0)) return fail(PIMCB_EINVAL, "dSep must be positive");
    if (f2_parity < -1 || f2_parity > 1) return fail(PIMCB_EINVAL, "f2_parity must be -1, 0 or 1");
    const int nsl = s->B * s->M, nd = c->ndim;
    if ((rc = c->d_vint.ensure(sizeof(double) * nsl))) return rc;
    if (f2 && (rc = c->d_f2.ensure(sizeof(double) * nsl))) return rc;
    if (sephist && (rc = c->d_hist.ensure(sizeof(int) * static_cast<size_t>(nsl) * kNPCFSEP))) return rc;
    PairParams pp{c->d_V.as<double>(), c->d_dV.as<double>(), c->tab_len, c->dr, {c->extV[0], c->extV[1]}, {c->extdV[0], c->extdV[1]},
                  sephist ? dSep : 1.0, sephist ? 1 : 0, f2_parity, s->M,
                  (c->gext_gen != 0 && c->gext_gen == s->gen) ? c->d_gext.as<double>() : nullptr};
    CU(cudaStreamWaitEvent(c->stream, s->ready, 0));
    if ((rc = materialize(c, *s))) return rc;
    {
        KTimer kt(c, K_PAIR);
        const size_t smem = sizeof(double) * nd * s->Npad;
        const int grid = grid_for(c, nsl, 8);
#define LAUNCH_PAIR(ND, F2)                                                                                       \
        rc = set_smem(pair_kernel<ND, F2>, smem); if (rc) return rc;                                               \
        pair_kernel<ND, F2><<<grid, 256, smem, c->stream>>>(s->pos.as<double>(), nsl, s->N, s->Npad, c->box, pp,  \
                                                            c->d_vint.as<double>(), c->d_f2.as<double>(), c->d_hist.as<int>())
        // force slices: every pair once (pair_sym_kernel) when the slice fits its particles-per-thread variants;
        // PIMCB_PAIR_SYM=0 forces the both-ends kernel (A/B)
        static const bool sym_on = !(std::getenv("PIMCB_PAIR_SYM") && std::atoi(std::getenv("PIMCB_PAIR_SYM")) == 0);
        // second-generation kernel (kernels_pair.cuh): 32 x 32 tiles, partner forces by warp shuffle, division-free table
        // index with an exact fallback; PIMCB_PAIR_TILE=0 selects the first-generation kernels (A/B)
        static const bool tile_on = !(std::getenv("PIMCB_PAIR_TILE") && std::atoi(std::getenv("PIMCB_PAIR_TILE")) == 0);
        const int G = (s->N + 31) / 32;
        const int spc = std::max(1, kPairWarps / G);
        const size_t smem_tile = sizeof(double) * (static_cast<size_t>(spc) * nd * 32 * G * (2 + kPairRound) + spc * G + 1) +
                                 sizeof(int) * spc * kNPCFSEP;
        const size_t smem_sym = sizeof(double) * 2 * nd * s->Npad;
        TileIndexParams ixp{};
        if (tile_on && smem_tile <= 200 * 1024 && make_index_params(c, pp.dSep, pp.want_hist != 0, &ixp)) {
            // Packed (V, dV/dr) sectors when the call has force slices (one sector request per pair instead of two reads
            // from 106 MB of tables); V-only calls read the verbatim V table, which fits L2 on its own (53 MB for C2) and
            // needs no decoding.  PIMCB_PAIR_PACKED=0 / 1 forces one or the other (A/B).
            static const int packed_env = std::getenv("PIMCB_PAIR_PACKED") ? std::atoi(std::getenv("PIMCB_PAIR_PACKED")) : -1;
            const bool use_packed = c->vd_ok && (packed_env < 0 ? f2 != nullptr : packed_env != 0);
            PairTileParams tp{c->d_V.as<double>(), c->d_dV.as<double>(), ixp, f2_parity, s->M, pp.gext, G, spc,
                              use_packed ? c->d_VD.as<TableSector>() : nullptr, -1, 0};
            // the V-only kernel needs neither force accumulators nor partner slots
            const size_t smem_v = sizeof(double) * (static_cast<size_t>(spc) * nd * 32 * G + spc * G + 1) + sizeof(int) * spc * kNPCFSEP;
#define LAUNCH_PTILE2(ND, CODEC, FK, NSEL, SMEM)                                                                   \
            { rc = set_smem(pair_tile_kernel<ND, CODEC, FK>, SMEM); if (rc) return rc;                              \
            pair_tile_kernel<ND, CODEC, FK><<<((NSEL) + spc - 1) / spc, 32 * kPairWarps, SMEM, c->stream>>>(s->pos.as<double>(), NSEL, s->N, s->Npad, c->box, tp, \
                                                                                  c->d_vint.as<double>(), f2 ? c->d_f2.as<double>() : nullptr, \
                                                                                  c->d_hist.as<int>()); }
#define LAUNCH_PTILE(ND, FK, NSEL, SMEM) if (tp.VD) LAUNCH_PTILE2(ND, true, FK, NSEL, SMEM) else LAUNCH_PTILE2(ND, false, FK, NSEL, SMEM)
#define LAUNCH_PTILE_ND(FK, NSEL, SMEM) { if (nd == 1) { LAUNCH_PTILE(1, FK, NSEL, SMEM) } else if (nd == 2) { LAUNCH_PTILE(2, FK, NSEL, SMEM) } else { LAUNCH_PTILE(3, FK, NSEL, SMEM) } }
            // PIMCB_PAIR_SPLIT=0: one launch of the force-capable kernel over all slices, as before (A/B)
            static const bool split_on = !(std::getenv("PIMCB_PAIR_SPLIT") && std::atoi(std::getenv("PIMCB_PAIR_SPLIT")) == 0);
            if (!f2) {
                LAUNCH_PTILE_ND(false, nsl, smem_v)                        // V-only call
            } else if (split_on && (f2_parity == 0 || f2_parity == 1) && s->M >= 2) {
                // gsf-type call: the slices that carry gradVSquared go through the force kernel, the others through the
                // V-only kernel (64 registers, four CTAs per SM; verbatim V table unless PIMCB_PAIR_PACKED=1) -- two launches
                // over disjoint slices and disjoint outputs
                const int cnt_f = (s->M - f2_parity + 1) / 2, cnt_v = s->M - cnt_f;
                tp.sel_p = f2_parity; tp.sel_cnt = cnt_f;
                LAUNCH_PTILE_ND(true, s->B * cnt_f, smem_tile)
                c->launches++;
                tp.sel_p = 1 - f2_parity; tp.sel_cnt = cnt_v;
                if (packed_env <= 0) tp.VD = nullptr;
                if (cnt_v > 0) LAUNCH_PTILE_ND(false, s->B * cnt_v, smem_v)
            } else {
                LAUNCH_PTILE_ND(true, nsl, smem_tile)                      // forces on every slice (or none selected)
            }
#undef LAUNCH_PTILE_ND
#undef LAUNCH_PTILE
#undef LAUNCH_PTILE2
        } else if (f2 && sym_on && s->N <= 1024 && smem_sym <= 200 * 1024) {
#define LAUNCH_PSYM(ND, PPT)                                                                                       \
            rc = set_smem(pair_sym_kernel<ND, PPT>, smem_sym); if (rc) return rc;                                   \
            pair_sym_kernel<ND, PPT><<<grid, 256, smem_sym, c->stream>>>(s->pos.as<double>(), nsl, s->N, s->Npad, c->box, pp, \
                                                                         c->d_vint.as<double>(), c->d_f2.as<double>(), c->d_hist.as<int>())
#define LAUNCH_PSYM_P(ND) if (s->N <= 256) { LAUNCH_PSYM(ND, 1); } else if (s->N <= 512) { LAUNCH_PSYM(ND, 2); } else { LAUNCH_PSYM(ND, 4); }
            if (nd == 1) { LAUNCH_PSYM_P(1) } else if (nd == 2) { LAUNCH_PSYM_P(2) } else { LAUNCH_PSYM_P(3) }
#undef LAUNCH_PSYM_P
#undef LAUNCH_PSYM
        } else if (f2) {
            if (nd == 1) { LAUNCH_PAIR(1, true); } else if (nd == 2) { LAUNCH_PAIR(2, true); } else { LAUNCH_PAIR(3, true); }
        } else {
            if (nd == 1) { LAUNCH_PAIR(1, false); } else if (nd == 2) { LAUNCH_PAIR(2, false); } else { LAUNCH_PAIR(3, false); }
        }
#undef LAUNCH_PAIR
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(s->consumed, c->stream));
    CU(cudaMemcpyAsync(vint, c->d_vint.p, sizeof(double) * nsl, cudaMemcpyDeviceToHost, c->stream));
    if (f2) CU(cudaMemcpyAsync(f2, c->d_f2.p, sizeof(double) * nsl, cudaMemcpyDeviceToHost, c->stream));
    if (sephist) CU(cudaMemcpyAsync(sephist, c->d_hist.p, sizeof(int) * static_cast<size_t>(nsl) * kNPCFSEP, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- scattering variants ------------------------------------------------------------------------------------------
int pimcb_elastic(pimcb_ctx* c, double* out) {
    if (!c || !out) return fail(PIMCB_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    Slot* s;
    int rc = need_cur(c, &s);
    if (rc) return rc;
    // S(q)/F(q,tau) of this very configuration batch may already be in d_cfg (pimcb_ssf_isf just before): reuse it
    if (!(c->cfg_slot == c->cur && c->cfg_gen == s->gen) && (rc = run_estimators(c, &s))) return rc;
    const int pairs = s->B * c->nq;
    if ((rc = c->d_var.ensure(sizeof(double) * pairs))) return rc;
    {
        KTimer kt(c, K_VARIANT);
        elastic_kernel<<<(pairs + 3) / 4, 128, 0, c->stream>>>(c->d_cfg.as<double>(), c->d_var.as<double>(), s->B, c->nq, s->M);
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(out, c->d_var.p, sizeof(double) * pairs, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int pimcb_ssf_cyl(pimcb_ctx* c, double maxR, double* out, int* n_inside) {
    if (!c || !out) return fail(PIMCB_EINVAL, "null argument");
    if (c->ndim < 2) return fail(PIMCB_EINVAL, "the cylinder cut-off needs at least two spatial dimensions");
    if (!(maxR > 0.0)) return fail(PIMCB_EINVAL, "maxR must be positive");
    CU(cudaSetDevice(c->device));
    Slot* s;
    int rc = need_cur(c, &s);
    if (rc) return rc;
    if (c->nq <= 0) return fail(PIMCB_ESTATE, "no q-vectors set");
    if (!c->have_box) return fail(PIMCB_ESTATE, "no box set");
    if (c->max_phase > 1.0e5) return fail(PIMCB_EINVAL, "max |q.r| = %g exceeds the sincos validity range 1e5", c->max_phase);
    const int nd = c->ndim, nsl = s->B * s->M;
    if ((rc = c->d_partial.ensure(sizeof(double) * static_cast<size_t>(nsl) * c->nq))) return rc;
    if ((rc = c->d_var.ensure(sizeof(double) * s->B * c->nq))) return rc;
    if ((rc = c->d_inside.ensure(sizeof(int) * s->B))) return rc;
    CU(cudaStreamWaitEvent(c->stream, s->ready, 0));
    if ((rc = materialize(c, *s))) return rc;
    {
        KTimer kt(c, K_VARIANT);
        const size_t smem = sizeof(double) * nd * s->Npad + static_cast<size_t>(s->Npad);
#define LAUNCH_CYL(ND)                                                                                             \
        rc = set_smem(ssf_cyl_kernel<ND>, smem); if (rc) return rc;                                                 \
        ssf_cyl_kernel<ND><<<nsl, 256, smem, c->stream>>>(s->pos.as<double>(), c->d_q.as<double>(), c->d_comm.as<unsigned char>(), \
                                                          c->nq, maxR, c->d_partial.as<double>(), c->d_inside.as<int>(), nsl, s->M, \
                                                          s->N, s->Npad, c->box)
        if (nd == 2) { LAUNCH_CYL(2); } else { LAUNCH_CYL(3); }
#undef LAUNCH_CYL
        CU(cudaGetLastError());
    }
    const int tot = s->B * c->nq;
    ssf_cyl_finalize_kernel<<<(tot + 127) / 128, 128, 0, c->stream>>>(c->d_partial.as<double>(), c->d_var.as<double>(), s->B, s->M, c->nq);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(s->consumed, c->stream));
    CU(cudaMemcpyAsync(out, c->d_var.p, sizeof(double) * tot, cudaMemcpyDeviceToHost, c->stream));
    if (n_inside) CU(cudaMemcpyAsync(n_inside, c->d_inside.p, sizeof(int) * s->B, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- virial slice sums ----------------------------------------------------------------------------------------------
int pimcb_set_pair_table_d2(pimcb_ctx* c, const double* d2Vdr2, int len, const double* extd2Vdr2) {
    if (!c || !d2Vdr2) return fail(PIMCB_EINVAL, "null argument");
    if (!c->tab_len || !c->have_dV) return fail(PIMCB_ESTATE, "pimcb_set_pair_table (with dV/dr) must precede pimcb_set_pair_table_d2");
    if (len != c->tab_len) return fail(PIMCB_EINVAL, "d2V/dr2 table length %d differs from the V table length %d", len, c->tab_len);
    CU(cudaSetDevice(c->device));
    int rc;
    c->extd2V[0] = extd2Vdr2 ? extd2Vdr2[0] : 0.0;
    c->extd2V[1] = extd2Vdr2 ? extd2Vdr2[1] : 0.0;
    std::vector<double> h2(d2Vdr2, d2Vdr2 + len);      // len + 1 entries like the other tables (entry 0 / len = the extremal values)
    h2.push_back(c->extd2V[1]);
    h2[0] = c->extd2V[0];
    if ((rc = c->d_d2V.ensure(sizeof(double) * h2.size()))) return rc;
    CU(cudaMemcpyAsync(c->d_d2V.p, h2.data(), sizeof(double) * h2.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->have_d2V = true;
    c->dd_ok = false;
    if (static_cast<int>(c->h_dV.size()) == len + 1 &&
        (rc = build_packed_table(c, c->h_dV.data(), h2.data(), len + 1, c->dr, c->d_dV.as<double>(), c->d_d2V.as<double>(), c->d_DD,
                                 &c->dd_ok, &c->dd_raw)))
        return rc;
    return 0;
}

int pimcb_virial_sums(pimcb_ctx* c, const double* delta_aos, int t2_parity, double* out) {
    if (!c || !out) return fail(PIMCB_EINVAL, "null argument");
    if (t2_parity < -2 || t2_parity > 1) return fail(PIMCB_EINVAL, "t2_parity must be -2, -1, 0 or 1");
    CU(cudaSetDevice(c->device));
    Slot* s;
    int rc = need_cur(c, &s);
    if (rc) return rc;
    if (!c->have_box) return fail(PIMCB_ESTATE, "no box set");
    if (!c->tab_len || !c->have_dV) return fail(PIMCB_ESTATE, "no dV/dr table set");
    if (t2_parity != -2 && !c->have_d2V) return fail(PIMCB_ESTATE, "T-matrix terms requested but no d2V/dr2 table set");
    const int nd = c->ndim, nsl = s->B * s->M;
    if ((rc = c->d_vir.ensure(sizeof(double) * 4 * nsl))) return rc;
    CU(cudaStreamWaitEvent(c->stream, s->ready, 0));
    if ((rc = materialize(c, *s))) return rc;
    const double* d_delta = nullptr;
    if (delta_aos) {
        // same AoS shape as the staged beads ([B][M][N_ext][ndim]); transposed on the device into the slice-row layout
        if ((rc = upload_bead_vectors(c, s, delta_aos, c->d_delta))) return rc;
        d_delta = c->d_delta.as<double>();
    }
    VirialParams vp{c->d_dV.as<double>(), c->d_d2V.as<double>(), c->tab_len, c->dr, {c->extdV[0], c->extdV[1]},
                    {c->extd2V[0], c->extd2V[1]}, t2_parity, s->M};
    {
        KTimer kt(c, K_VIRIAL);
        // symmetric kernel (every pair once, partner side accumulated in shared memory) when the slice fits its
        // particles-per-thread variants; PIMCB_VIRIAL_SYM=0 forces the both-ends kernel (A/B)
        static const bool sym_on = !(std::getenv("PIMCB_VIRIAL_SYM") && std::atoi(std::getenv("PIMCB_VIRIAL_SYM")) == 0);
        const int nc = nd + nd * (nd + 1) / 2;
        const size_t smem_sym = sizeof(double) * (2 * nd + nc) * s->Npad;
        // a non-free external potential (gradient / Laplacian uploaded for THIS configuration) couples to every term:
        // both-ends kernel with the per-bead external arrays
        const double* d_gext = (c->gext_gen != 0 && c->gext_gen == s->gen) ? c->d_gext.as<double>() : nullptr;
        const double* d_g2ext = (c->g2ext_gen != 0 && c->g2ext_gen == s->gen) ? c->d_g2ext.as<double>() : nullptr;
        const bool ext = d_gext != nullptr || d_g2ext != nullptr;
        // second-generation kernel: 32 x 32 tiles, partner sums by warp shuffle, packed (dV/dr, d2V/dr2) sectors;
        // PIMCB_VIRIAL_TILE=0 selects the first-generation kernels (A/B)
        static const bool vtile_on = !(std::getenv("PIMCB_VIRIAL_TILE") && std::atoi(std::getenv("PIMCB_VIRIAL_TILE")) == 0);
        const int G = (s->N + 31) / 32;
        const int spc = std::max(1, kPairWarps / G);
        auto vtile_smem = [&](int R) {
            return sizeof(double) * (static_cast<size_t>(spc) * 32 * G * (nd + nc * (1 + R)) + static_cast<size_t>(spc) * 4 * kPairWarps);
        };
        int rounds = kPairRound;
        while (rounds > 1 && vtile_smem(rounds) > 100 * 1024) --rounds;      // two CTAs per SM when it fits
        TileIndexParams ixp{};
        if (!ext && vtile_on && vtile_smem(rounds) <= 200 * 1024 && make_index_params(c, 1.0, false, &ixp)) {
            const bool packed = c->dd_ok && (t2_parity == -2 || c->have_d2V);
            VirialTileParams tp{c->d_dV.as<double>(), c->d_d2V.as<double>(), ixp, t2_parity, s->M, G, spc, rounds,
                                packed ? c->d_DD.as<TableSector>() : nullptr, -1, 0};
            const size_t smem_t = vtile_smem(rounds);
            // the gV-only kernel keeps nd instead of nc components per particle
            const size_t smem_g = sizeof(double) * (static_cast<size_t>(spc) * 32 * G * (nd + nd * (1 + rounds)) + static_cast<size_t>(spc) * 4 * kPairWarps);
#define LAUNCH_VTILE2(ND, CODEC, T2K, NSEL, SMEM)                                                                  \
            { rc = set_smem(virial_tile_kernel<ND, CODEC, T2K>, SMEM); if (rc) return rc;                           \
            virial_tile_kernel<ND, CODEC, T2K><<<((NSEL) + spc - 1) / spc, 32 * kPairWarps, SMEM, c->stream>>>(s->pos.as<double>(), d_delta, NSEL, s->N, s->Npad, \
                                                                                         c->box, tp, c->d_vir.as<double>()); }
#define LAUNCH_VTILE(ND, T2K, NSEL, SMEM) if (packed) LAUNCH_VTILE2(ND, true, T2K, NSEL, SMEM) else LAUNCH_VTILE2(ND, false, T2K, NSEL, SMEM)
#define LAUNCH_VTILE_ND(T2K, NSEL, SMEM) { if (nd == 1) { LAUNCH_VTILE(1, T2K, NSEL, SMEM) } else if (nd == 2) { LAUNCH_VTILE(2, T2K, NSEL, SMEM) } else { LAUNCH_VTILE(3, T2K, NSEL, SMEM) } }
            // PIMCB_VIRIAL_SPLIT=0: one launch of the T-matrix-capable kernel over all slices, as before (A/B)
            static const bool vsplit_on = !(std::getenv("PIMCB_VIRIAL_SPLIT") && std::atoi(std::getenv("PIMCB_VIRIAL_SPLIT")) == 0);
            if (t2_parity == -2) {
                LAUNCH_VTILE_ND(false, nsl, smem_g)                        // no slice carries the T-matrix terms
            } else if (vsplit_on && (t2_parity == 0 || t2_parity == 1) && s->M >= 2) {
                // gsf-type call: two launches over disjoint slices and disjoint output rows
                const int cnt_t = (s->M - t2_parity + 1) / 2, cnt_g = s->M - cnt_t;
                tp.sel_p = t2_parity; tp.sel_cnt = cnt_t;
                LAUNCH_VTILE_ND(true, s->B * cnt_t, smem_t)
                c->launches++;
                tp.sel_p = 1 - t2_parity; tp.sel_cnt = cnt_g;
                if (cnt_g > 0) LAUNCH_VTILE_ND(false, s->B * cnt_g, smem_g)
            } else {
                LAUNCH_VTILE_ND(true, nsl, smem_t)
            }
#undef LAUNCH_VTILE_ND
#undef LAUNCH_VTILE
#undef LAUNCH_VTILE2
        } else if (!ext && sym_on && s->N <= 1024 && smem_sym <= 200 * 1024) {
#define LAUNCH_VSYM(ND, PPT)                                                                                       \
            rc = set_smem(virial_sym_kernel<ND, PPT>, smem_sym); if (rc) return rc;                                 \
            virial_sym_kernel<ND, PPT><<<nsl, 256, smem_sym, c->stream>>>(s->pos.as<double>(), d_delta, nsl, s->N, s->Npad, c->box, vp, \
                                                                          c->d_vir.as<double>())
#define LAUNCH_VSYM_P(ND) if (s->N <= 256) { LAUNCH_VSYM(ND, 1); } else if (s->N <= 512) { LAUNCH_VSYM(ND, 2); } else { LAUNCH_VSYM(ND, 4); }
            if (nd == 1) { LAUNCH_VSYM_P(1) } else if (nd == 2) { LAUNCH_VSYM_P(2) } else { LAUNCH_VSYM_P(3) }
#undef LAUNCH_VSYM_P
#undef LAUNCH_VSYM
        } else {
        const size_t smem = sizeof(double) * 2 * nd * s->Npad;
#define LAUNCH_VIR2(ND, EXT)                                                                                       \
        rc = set_smem(virial_kernel<ND, EXT>, smem); if (rc) return rc;                                             \
        virial_kernel<ND, EXT><<<nsl, 256, smem, c->stream>>>(s->pos.as<double>(), d_delta, nsl, s->N, s->Npad, c->box, vp, \
                                                              c->d_vir.as<double>(), d_gext, d_g2ext)
#define LAUNCH_VIR(ND) if (ext) { LAUNCH_VIR2(ND, true); } else { LAUNCH_VIR2(ND, false); }
        if (nd == 1) { LAUNCH_VIR(1) } else if (nd == 2) { LAUNCH_VIR(2) } else { LAUNCH_VIR(3) }
#undef LAUNCH_VIR
#undef LAUNCH_VIR2
        }
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(s->consumed, c->stream));
    CU(cudaMemcpyAsync(out, c->d_vir.p, sizeof(double) * 4 * nsl, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- multi-GPU exchange step ------------------------------------------------------------------------------------------
int pimcb_comm_unique_id(void* id_out) {
    if (!id_out) return fail(PIMCB_EINVAL, "null argument");
    if (int rc = load_nccl()) return rc;
    NcclId id;
    NCCLCHK(g_nccl.GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof id);
    return 0;
}

int pimcb_comm_init(pimcb_ctx* c, int nranks, int rank, const void* unique_id) {
    if (!c || !unique_id || nranks < 1 || rank < 0 || rank >= nranks) return fail(PIMCB_EINVAL, "bad communicator arguments");
    if (c->nccl_comm) return fail(PIMCB_ESTATE, "communicator already initialised");
    if (int rc = load_nccl()) return rc;
    CU(cudaSetDevice(c->device));
    NcclId id;
    std::memcpy(&id, unique_id, sizeof id);
    NCCLCHK(g_nccl.CommInitRank(&c->nccl_comm, nranks, id, rank));
    c->comm_rank = rank;
    c->comm_size = nranks;
    return 0;
}

int pimcb_comm_destroy(pimcb_ctx* c) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    if (c->nccl_comm) {
        CU(cudaStreamSynchronize(c->stream));
        if (c->comm_stream) CU(cudaStreamSynchronize(c->comm_stream));
        c->xchg_inflight = false;
        NCCLCHK(g_nccl.CommDestroy(c->nccl_comm));
        c->nccl_comm = nullptr;
    }
    c->comm_rank = 0;
    c->comm_size = 1;
    return 0;
}

int pimcb_reduce_bins(pimcb_ctx* c, int root, long* num_total) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    if (!c->nccl_comm) return fail(PIMCB_ESTATE, "pimcb_comm_init has not been called");
    if (root < 0 || root >= c->comm_size) return fail(PIMCB_EINVAL, "root %d out of range", root);
    if (!c->bins_len) return fail(PIMCB_ESTATE, "no measurement accumulated yet");
    CU(cudaSetDevice(c->device));
    NvtxRange range("pimcb:reduce_bins");
    int rc;
    if ((rc = fold_binrows(c, c->bins_M))) return rc;
    if ((rc = c->d_count.ensure(2 * sizeof(long long)))) return rc;
    long long* cnt = c->d_count.as<long long>();
    const long long mine = c->n_acc;
    CU(cudaMemcpyAsync(cnt, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
    // one group: the bin (in place on the root) and the number of configurations in it
    NCCLCHK(g_nccl.GroupStart());
    NCCLCHK(g_nccl.Reduce(c->d_bins.p, c->d_bins.p, c->bins_len, /*ncclDouble*/ 8, /*ncclSum*/ 0, root, c->nccl_comm, c->stream));
    NCCLCHK(g_nccl.Reduce(cnt, cnt + 1, 1, /*ncclInt64*/ 4, /*ncclSum*/ 0, root, c->nccl_comm, c->stream));
    NCCLCHK(g_nccl.GroupEnd());
    if (c->comm_rank == root) {
        // the root's bin now holds every rank's configurations.  With num_total the call waits for the count; without, it
        // returns with everything enqueued and the count is picked up by the next pimcb_read_bins (which synchronises
        // anyway): no host synchronisation per bin, so the next bin's launches queue up behind the collective.
        if ((rc = c->h_cnt.ensure(sizeof(long long)))) return rc;
        CU(cudaMemcpyAsync(c->h_cnt.p, cnt + 1, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        c->n_acc_pending = true;
        if (num_total) {
            if ((rc = settle_count(c, false))) return rc;
            *num_total = c->n_acc;
        }
    } else if (num_total) {
        *num_total = 0;
    }
    return 0;
}

// The same exchange, pipelined: _begin snapshots the folded bin and its count and starts the reduce on the library's
// communication stream; the caller resets the bin and goes on measuring; _end (one bin later, or whenever the row is
// written) waits for the snapshot's reduce and hands the global bin to the root.  One exchange in flight per ctx.
int pimcb_reduce_bins_begin(pimcb_ctx* c, int root) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    if (!c->nccl_comm) return fail(PIMCB_ESTATE, "pimcb_comm_init has not been called");
    if (root < 0 || root >= c->comm_size) return fail(PIMCB_EINVAL, "root %d out of range", root);
    if (!c->bins_len) return fail(PIMCB_ESTATE, "no measurement accumulated yet");
    if (c->xchg_inflight) return fail(PIMCB_ESTATE, "the previous exchange has not been collected (pimcb_reduce_bins_end)");
    CU(cudaSetDevice(c->device));
    NvtxRange range("pimcb:reduce_bins_begin");
    int rc;
    if ((rc = settle_count(c, false))) return rc;
    if (!c->comm_stream) {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));          // hi = numerically lowest = highest priority
        CU(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
        CU(cudaEventCreateWithFlags(&c->xchg_ready, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->xchg_done, cudaEventDisableTiming));
    }
    if ((rc = fold_binrows(c, c->bins_M))) return rc;
    const size_t len = c->bins_len, bytes = sizeof(double) * len;
    if ((rc = c->d_xchg.ensure(bytes + 2 * sizeof(long long)))) return rc;
    if ((rc = c->h_xchg.ensure(bytes + 2 * sizeof(long long)))) return rc;
    long long* hcnt = reinterpret_cast<long long*>(static_cast<char*>(c->h_xchg.p) + bytes);
    long long* dcnt = reinterpret_cast<long long*>(static_cast<char*>(c->d_xchg.p) + bytes);
    hcnt[0] = c->n_acc;
    hcnt[1] = 0;
    CU(cudaMemcpyAsync(c->d_xchg.p, c->d_bins.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaEventRecord(c->xchg_ready, c->stream));
    CU(cudaStreamWaitEvent(c->comm_stream, c->xchg_ready, 0));
    CU(cudaMemcpyAsync(dcnt, hcnt, sizeof(long long), cudaMemcpyHostToDevice, c->comm_stream));
    NCCLCHK(g_nccl.GroupStart());
    NCCLCHK(g_nccl.Reduce(c->d_xchg.p, c->d_xchg.p, len, /*ncclDouble*/ 8, /*ncclSum*/ 0, root, c->nccl_comm, c->comm_stream));
    NCCLCHK(g_nccl.Reduce(dcnt, dcnt + 1, 1, /*ncclInt64*/ 4, /*ncclSum*/ 0, root, c->nccl_comm, c->comm_stream));
    NCCLCHK(g_nccl.GroupEnd());
    if (c->comm_rank == root) {
        CU(cudaMemcpyAsync(c->h_xchg.p, c->d_xchg.p, bytes, cudaMemcpyDeviceToHost, c->comm_stream));
        CU(cudaMemcpyAsync(hcnt + 1, dcnt + 1, sizeof(long long), cudaMemcpyDeviceToHost, c->comm_stream));
    }
    CU(cudaEventRecord(c->xchg_done, c->comm_stream));
    c->xchg_inflight = true;
    c->xchg_root = root;
    c->xchg_nq = c->nq;
    c->xchg_M = c->bins_M;
    c->xchg_len = len;
    return 0;
}

int pimcb_reduce_bins_end(pimcb_ctx* c, double* ssf, double* isf, long* num_total) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    if (!c->xchg_inflight) return fail(PIMCB_ESTATE, "no exchange in flight (pimcb_reduce_bins_begin)");
    CU(cudaSetDevice(c->device));
    NvtxRange range("pimcb:reduce_bins_end");
    CU(cudaEventSynchronize(c->xchg_done));
    c->xchg_inflight = false;
    if (c->comm_rank == c->xchg_root) {
        const double* h = static_cast<const double*>(c->h_xchg.p);
        if (ssf) std::memcpy(ssf, h, sizeof(double) * c->xchg_nq);
        if (isf) std::memcpy(isf, h + c->xchg_nq, sizeof(double) * (c->xchg_len - c->xchg_nq));
        if (num_total) *num_total = static_cast<long>(reinterpret_cast<const long long*>(h + c->xchg_len)[1]);
    } else if (num_total) {
        *num_total = 0;
    }
    return 0;
}

int pimcb_gather_bins_q(pimcb_ctx* c, const int* nq_per_rank, double* ssf, double* isf) {
    if (!c || !nq_per_rank) return fail(PIMCB_EINVAL, "null argument");
    if (!c->nccl_comm) return fail(PIMCB_ESTATE, "pimcb_comm_init has not been called");
    if (!c->bins_len) return fail(PIMCB_ESTATE, "no measurement accumulated yet");
    if (nq_per_rank[c->comm_rank] != c->nq) return fail(PIMCB_EINVAL, "nq_per_rank[%d] = %d but this rank holds %d q-vectors",
                                                       c->comm_rank, nq_per_rank[c->comm_rank], c->nq);
    CU(cudaSetDevice(c->device));
    NvtxRange range("pimcb:gather_bins_q");
    int rc;
    if ((rc = fold_binrows(c, c->bins_M))) return rc;
    const int M = c->bins_M;
    int width = 0, total = 0;
    for (int r = 0; r < c->comm_size; ++r) { width = std::max(width, nq_per_rank[r]); total += nq_per_rank[r]; }
    const size_t slot = static_cast<size_t>(width) * (1 + M);           // padded shard: [width] S then [width][M] F
    if ((rc = c->d_gather.ensure(sizeof(double) * slot * (c->comm_size + 1)))) return rc;
    double* send = c->d_gather.as<double>() + slot * c->comm_size;
    CU(cudaMemsetAsync(send, 0, sizeof(double) * slot, c->stream));
    CU(cudaMemcpyAsync(send, c->d_bins.p, sizeof(double) * c->nq, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(send + width, c->d_bins.as<double>() + c->nq, sizeof(double) * c->nq * M, cudaMemcpyDeviceToDevice, c->stream));
    NCCLCHK(g_nccl.AllGather(send, c->d_gather.p, slot, /*ncclDouble*/ 8, c->nccl_comm, c->stream));
    if ((rc = c->h_out.ensure(sizeof(double) * slot * c->comm_size))) return rc;
    CU(cudaMemcpyAsync(c->h_out.p, c->d_gather.p, sizeof(double) * slot * c->comm_size, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const double* h = static_cast<const double*>(c->h_out.p);
    int q0 = 0;
    for (int r = 0; r < c->comm_size; ++r) {                              // rank order = q order
        const int n = nq_per_rank[r];
        if (ssf) std::memcpy(ssf + q0, h + slot * r, sizeof(double) * n);
        if (isf) std::memcpy(isf + static_cast<size_t>(q0) * M, h + slot * r + width, sizeof(double) * n * M);
        q0 += n;
    }
    (void)total;
    return 0;
}

int pimcb_measure_fp64_peak(pimcb_ctx* c, double* tflops, double seconds_target) {
    if (!c || !tflops) return fail(PIMCB_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    int rc = c->d_scratch.ensure(64);
    if (rc) return rc;
    const int grid = c->sm_count * 8, threads = 256, iters = 4096;
    const double flop_per_launch = 2.0 * 32.0 * iters * static_cast<double>(grid) * threads;
    cudaEvent_t e0 = c->ev0[kKernels - 2], e1 = c->ev1[kKernels - 2];
    for (int w = 0; w < 3; ++w) fp64_peak_kernel<<<grid, threads, 0, c->stream>>>(c->d_scratch.as<double>(), iters, 1.0000001, 1e-9);
    CU(cudaStreamSynchronize(c->stream));
    double best = 0.0, elapsed = 0.0;
    int reps = 0;
    while (elapsed < seconds_target || reps < 3) {
        CU(cudaEventRecord(e0, c->stream));
        for (int k = 0; k < 4; ++k) fp64_peak_kernel<<<grid, threads, 0, c->stream>>>(c->d_scratch.as<double>(), iters, 1.0000001, 1e-9);
        CU(cudaEventRecord(e1, c->stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        best = std::max(best, 4.0 * flop_per_launch / (ms * 1e-3) / 1e12);
        elapsed += ms * 1e-3;
        c->launches += 4;
        if (++reps > 10000) break;
    }
    CU(cudaGetLastError());
    *tflops = best;
    return 0;
}

// Bare host-to-device rate of THIS process's link: `reps` back-to-back cudaMemcpyAsync of a page-locked buffer on the copy
// stream, nothing else in flight -- the ceiling the staged (e2e) path is measured against in the same run.
int pimcb_measure_h2d_peak(pimcb_ctx* c, const void* pinned_src, size_t bytes, int reps, double* gbs) {
    if (!c || !pinned_src || !gbs || bytes == 0 || reps < 1) return fail(PIMCB_EINVAL, "bad arguments");
    CU(cudaSetDevice(c->device));
    int rc = c->d_scratch.ensure(bytes);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    cudaEvent_t e0 = c->ev0[kKernels - 1], e1 = c->ev1[kKernels - 1];
    CU(cudaMemcpyAsync(c->d_scratch.p, pinned_src, bytes, cudaMemcpyHostToDevice, c->copy_stream));     // warm-up
    CU(cudaEventRecord(e0, c->copy_stream));
    for (int r = 0; r < reps; ++r) CU(cudaMemcpyAsync(c->d_scratch.p, pinned_src, bytes, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaEventRecord(e1, c->copy_stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    *gbs = static_cast<double>(bytes) * reps / (ms * 1e-3) / 1e9;
    return 0;
}

int pimcb_set_profiling_stride(pimcb_ctx* c, int stride) {
    if (!c || stride < 1) return fail(PIMCB_EINVAL, "profiling stride must be >= 1");
    c->prof_stride = stride;
    for (long& v : c->prof_seen) v = 0;
    return 0;
}

int pimcb_set_profiling(pimcb_ctx* c, int on) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    c->profiling = on == 1 ? ~0u : static_cast<unsigned>(on) >> 1;   // 0 off, 1 all, else mask with bit (id + 1) per kernel
    return 0;
}

int pimcb_kernel_times(pimcb_ctx* c, double* ms_total, long* count, int reset) {
    if (!c) return fail(PIMCB_EINVAL, "null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->stream));
    for (auto& r : c->recs) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, r.a, r.b));
        c->k_ms[r.k] += ms;
        c->k_count[r.k]++;
        c->ev_pool.push_back(r.a);
        c->ev_pool.push_back(r.b);
    }
    c->recs.clear();
    for (int k = 0; k < kKernels; ++k) {
        if (ms_total) ms_total[k] = c->k_ms[k];
        if (count) count[k] = c->k_count[k];
        if (reset) { c->k_ms[k] = 0.0; c->k_count[k] = 0; }
    }
    return 0;
}

long pimcb_launch_count(const pimcb_ctx* c) { return c ? c->launches : 0; }

int pimcb_rho_plan_info(const pimcb_ctx* c, int* info) {
    if (!c || !info) return fail(PIMCB_EINVAL, "null argument");
    const int v[12] = {c->last_rho_path, c->ngroups, c->mma_nL, c->mma_nR, c->last_ML, c->last_NR,
                       c->nmax[0], c->nmax[1], c->nmax[2], c->ncomm, c->nsel, c->nq};
    for (int k = 0; k < 12; ++k) info[k] = v[k];
    return 0;
}

}  // extern "C"
