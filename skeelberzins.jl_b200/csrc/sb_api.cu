// sb_api.cu -- C ABI of the B200-native batched Skeel-Berzins solve path (include/skeelberzins_b200.h).
//
// Host responsibilities kept here (all O(Nx) or O(B) and once per problem): problem_init
// (src/utils.jl:609-663) incl. the quadrature points/weights (src/utils.jl:683-718) and the
// interpolation weights (src/utils.jl:18-48), choice of the thread decomposition (R rows per thread,
// T threads per system), device allocation, layout transforms, the Newton software pipeline.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/skeelberzins_b200.h"
#include "sb_longmesh.cuh"
#include "sb_registry.h"

using namespace sb;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(SB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int kSlots = 128;  // ring of per-iteration "still active" counters

// designs whose per-iteration kernel is the persistent k_newton_fused_p (the resident design adds the whole-solve kernel
// for the loops driven from C)
inline bool is_tma(int design) { return design == SB_DESIGN_FUSED_TMA || design == SB_DESIGN_RESIDENT; }

std::atomic<long long> g_launches{0};  // kernels launched by this library (all contexts; contexts may run on different host threads)

// Kernel handles of one problem: host function pointers of the statically compiled registry kernels, or
// cudaKernel_t handles of kernels compiled at run time (NVRTC) for a user functor.  Both launch through
// cudaLaunchKernel.
struct KernelSet {
    const void *mass = nullptr, *assemble_jac = nullptr, *assemble_val = nullptr, *fused = nullptr, *fused_p = nullptr, *resident = nullptr, *solve = nullptr;
    const void* resident_pipe = nullptr;   // k_solve_resident<..., PIPE = true> (sb_solve_from_host)
    bool pipe_configured = false;
    const void *long_chunk = nullptr, *long_back = nullptr, *long_assemble_jac = nullptr, *long_assemble_val = nullptr;
    size_t fused_p_smem = 0, long_smem = 0, resident_smem = 0;
    bool long_configured = false, small_configured = false;
    cudaLibrary_t lib = nullptr;  // owner of the run-time compiled kernels (user functors)
};

template <class... Args>
cudaError_t launch_any(const void* func, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    void* p[] = {(void*)&args...};
    ++g_launches;
    return cudaLaunchKernel(func, grid, block, p, smem, st);
}

// Launch with programmatic stream serialization: the kernel may begin (up to its griddepcontrol.wait) while the
// previous kernel in the stream drains.  Used for the persistent Newton-iteration kernel, whose prologue (barrier
// set-up, mesh weights -> shared memory) does not depend on the previous iteration.  SB_NO_PDL=1 disables it.
template <class... Args>
cudaError_t launch_pdl(const void* func, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    static const bool off = getenv("SB_NO_PDL") != nullptr;
    if (off) return launch_any(func, grid, block, smem, st, args...);
    void* p[] = {(void*)&args...};
    ++g_launches;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelExC(&cfg, func, p);
}

struct UserFunctor {
    std::string source, name;
    int npde, nparams;
};
// Registered user functors.  Contexts are independent across host threads, the table is shared: every access goes
// through these helpers (a deque never moves its elements, so the pointer a problem keeps stays valid).
std::mutex g_uf_mu;
std::deque<UserFunctor>& user_functors_unlocked() {
    static std::deque<UserFunctor> v;
    return v;
}
const UserFunctor* user_functor(int functor_id) {
    std::lock_guard<std::mutex> lk(g_uf_mu);
    std::deque<UserFunctor>& v = user_functors_unlocked();
    const int i = functor_id - SB_USER_FUNCTOR_BASE;
    return (functor_id >= SB_USER_FUNCTOR_BASE && i < (int)v.size()) ? &v[i] : nullptr;
}
int add_user_functor(const UserFunctor& uf) {
    std::lock_guard<std::mutex> lk(g_uf_mu);
    std::deque<UserFunctor>& v = user_functors_unlocked();
    for (size_t i = 0; i < v.size(); ++i)   // registering the same source again returns the same id (and its compiled kernels)
        if (v[i].npde == uf.npde && v[i].nparams == uf.nparams && v[i].name == uf.name && v[i].source == uf.source)
            return SB_USER_FUNCTOR_BASE + (int)i;
    user_functors_unlocked().push_back(uf);
    return SB_USER_FUNCTOR_BASE + (int)user_functors_unlocked().size() - 1;
}

// K1 on the device for m == 0 (get_quad_points_weights, src/utils.jl:683-718, and the interpolation / volume weights
// of src/utils.jl:18-48, src/assembler.jl:36,68-69,98): every weight of slot s from the mesh points around it.  The
// expressions are those of the host loop in sb_problem_create (divisions are IEEE on both sides, nothing contracts to
// an FMA), so the arrays agree bit for bit with what the host computes for m >= 1 meshes -- without a 9*L staging
// vector and a pass over it on the host (2^22 nodes: 300 MB).  Layout: xi[L] | awbw[2L] | gwwf[2L] | frac[2L] | gw[L]
// | ivol[L] | xs[L]; xs is the input.
__global__ void k_mesh_weights_m0(double* __restrict__ mesh, int Nx, int R, int T, size_t L) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= L) return;
    const int Lt = R * T;
    auto slot_of = [&](int i) { const int tile = i / Lt, il = i - tile * Lt; return (size_t)tile * Lt + (size_t)(il % R) * T + il / R; };
    const int tile = (int)(s / Lt), sl = (int)(s - (size_t)tile * Lt);
    const int i = tile * Lt + (sl % T) * R + sl / T;
    const double* xs = mesh + 9 * L;
    double xi = 0.0, aw = 0.0, bw = 0.0, gw = 0.0, wf = 0.0, fR = 0.0, fL = 0.0, ivol = 0.0;
    if (i < Nx) {
        const double xc = xs[s];
        if (i < Nx - 1) {
            const double xr = xs[slot_of(i + 1)], q = (xc + xr) / 2, h = xr - xc;
            xi = q; aw = (xr - q) / h; bw = (q - xc) / h; gw = 1.0 / h; wf = 1.0;
            fR = q - xc;
        }
        if (i >= 1) fL = xc - (xs[slot_of(i - 1)] + xc) / 2;
        if (i >= 1 && i < Nx - 1) ivol = 1.0 / (fR + fL);
    }
    mesh[s] = xi;
    mesh[L + 2 * s] = aw; mesh[L + 2 * s + 1] = bw;
    mesh[3 * L + 2 * s] = gw; mesh[3 * L + 2 * s + 1] = wf;
    mesh[5 * L + 2 * s] = fR; mesh[5 * L + 2 * s + 1] = fL;
    mesh[7 * L + s] = gw;
    mesh[8 * L + s] = ivol;
}

// reference layout [B][Nx][W] <-> device layout [B][tile][R][T][W]   (W doubles per node, tiles of Lt = R*T nodes,
// L = ntiles*Lt padded nodes per instance).  identity != 0: padding nodes receive the W = P*P identity block (Jd)
__global__ void k_to_device_layout(const double* __restrict__ src, double* __restrict__ dst, int64_t B, int Nx, int R,
                                   int T, int L, int W, int identity_P) {
    const int Lt = R * T;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*L*W device elements
    if (idx >= (size_t)B * L * W) return;
    const int w = (int)(idx % W);
    const size_t ns = idx / W;
    const int64_t k = ns / L;
    const int s = (int)(ns - (size_t)k * L);
    const int tile = s / Lt, sl = s - tile * Lt;
    const int r = sl / T, t = sl - r * T;
    const int i = tile * Lt + t * R + r;
    double pad = 0.0;
    if (identity_P > 0 && (w / identity_P) == (w % identity_P)) pad = 1.0;
    dst[idx] = (i < Nx) ? src[((size_t)k * Nx + i) * W + w] : pad;
}

__global__ void k_from_device_layout(const double* __restrict__ src, double* __restrict__ dst, int64_t B, int Nx, int R,
                                     int T, int L, int W) {
    const int Lt = R * T;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*Nx*W reference elements
    if (idx >= (size_t)B * Nx * W) return;
    const int w = (int)(idx % W);
    const size_t ni = idx / W;
    const int64_t k = ni / Nx;
    const int i = (int)(ni - (size_t)k * Nx);
    const int tile = i / Lt, il = i - tile * Lt;
    dst[idx] = src[((size_t)k * L + (size_t)tile * Lt + (size_t)(il % R) * T + il / R) * W + w];
}

// pdeval / interpolate_sol_space (src/utils.jl:396-438, :460-559): one thread per (instance, evaluation point);
// bisection on the shared mesh, then the interpolation of src/utils.jl:18-48 at the point itself
__global__ void k_pdeval(const double* __restrict__ u, const double* __restrict__ x, int Nx, int R, int T, int L, int P,
                         int m, int singular, int64_t B, int nq, const double* __restrict__ xq, double* __restrict__ uo,
                         double* __restrict__ dxo) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * nq) return;
    const int64_t k = idx / nq;
    const int q = (int)(idx - (size_t)k * nq);
    const double pt = xq[q];
    int il;
    if (pt == x[0] || fabs(pt - x[0]) <= 1.0e-10 * fabs(x[1] - x[0])) {
        il = 0;
    } else {
        int lo = 0, hi = Nx;             // first index with x[i] >= pt (searchsortedfirst)
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (x[mid] < pt) lo = mid + 1; else hi = mid; }
        il = lo - 1;                     // the host has checked 1 <= lo <= Nx-1
    }
    const int Lt = R * T;
    auto slot = [&](int i) { const int tile = i / Lt, r = i - tile * Lt; return (size_t)tile * Lt + (size_t)(r % R) * T + r / R; };
    const double xl = x[il], xr = x[il + 1];
    double aw, bw, gw;                   // U = aw*ul + bw*ur, Ux = gw*(ur - ul)
    if (singular) { const double h = xr * xr - xl * xl, th = (pt * pt - xl * xl) / h; aw = 1 - th; bw = th; gw = (2 * pt) / h; }
    else if (m == 0) { const double h = xr - xl; aw = (xr - pt) / h; bw = (pt - xl) / h; gw = 1.0 / h; }
    else if (m == 1) { const double lg = log(xr / xl), th = log(pt / xl) / lg; aw = 1 - th; bw = th; gw = 1.0 / (pt * lg); }
    else { const double h = xr - xl, th = (xr * (pt - xl)) / (pt * h); aw = 1 - th; bw = th; gw = (xr * xl) / (pt * pt * h); }
    const double* ub = u + (size_t)k * L * P;
    for (int j = 0; j < P; ++j) {
        const double ul = ub[slot(il) * P + j], ur = ub[slot(il + 1) * P + j];
        uo[idx * P + j] = aw * ul + bw * ur;
        dxo[idx * P + j] = gw * (ur - ul);
    }
}

const FunctorEntry* registry() {
    static FunctorEntry table[SB_NUM_BUILTIN_FUNCTORS];
    static std::once_flag once;
    std::call_once(once, [] {
        table[SB_LIN_DIFF] = entry_lin_diff();
        table[SB_NONLIN_DIFF] = entry_nonlin_diff();
        table[SB_LIN_DIFF_CYL] = entry_lin_diff_cyl();
        table[SB_POISSON] = entry_poisson();
        table[SB_STAT_NONLIN] = entry_stat_nonlin();
        table[SB_REACT_DIFF] = entry_react_diff();
        table[SB_LIN_DIFF_SPH] = entry_lin_diff_sph();
        table[SB_CONV_DIFF] = entry_conv_diff();
        table[SB_LIN_REACT] = entry_lin_react();
        table[SB_GENERIC_SCALAR] = entry_generic_scalar();
    });
    return table;
}

SolveKernel solve_kernel(int P, int dec) {   // registry functors: npde <= 2 (wider systems come through the plugin, NVRTC)
    static SolveKernel t1[kNumDecomp], t2[kNumDecomp];
    static std::once_flag once;
    std::call_once(once, [] {
        fill_solve<1>(t1, std::make_integer_sequence<int, kNumDecomp>{});
        fill_solve<2>(t2, std::make_integer_sequence<int, kNumDecomp>{});
    });
    return P == 1 ? t1[dec] : P == 2 ? t2[dec] : nullptr;
}

inline double ipow(double x, int n) {
    double r = 1.0;
    for (int k = 0; k < n; ++k) r *= x;
    return r;
}

// ---- guard mode (SB_GUARD=1): every device array of a problem gets 4 KiB of canary bytes on both sides; they are
// checked when the array is freed and violations are counted (sb_debug_guard_violations).  compute-sanitizer is not
// available on every pool; this catches out-of-bounds WRITES of any kernel over a whole test run.
constexpr size_t kGuardBytes = 4096;
constexpr int kGuardByte = 0xFF;   // all-ones: a double read from a guard region is a NaN, so out-of-bounds READS that matter poison the results
std::atomic<long long> g_guard_violations{0}, g_guard_arrays{0};
struct GuardRec { char* base; size_t bytes; };
std::mutex g_guard_mu;
std::unordered_map<void*, GuardRec> g_guard;
bool guard_on() { static const bool on = getenv("SB_GUARD") != nullptr; return on; }

cudaError_t guarded_malloc(void** p, size_t bytes) {
    if (!guard_on()) return cudaMalloc(p, bytes);
    char* base = nullptr;
    cudaError_t e = cudaMalloc((void**)&base, bytes + 2 * kGuardBytes);
    if (e != cudaSuccess) return e;
    cudaMemset(base, kGuardByte, kGuardBytes);
    cudaMemset(base + kGuardBytes + bytes, kGuardByte, kGuardBytes);
    *p = base + kGuardBytes;
    static std::atomic<bool> selftest{getenv("SB_GUARD_SELFTEST") != nullptr};   // positive control: spoil one canary byte once
    if (selftest.exchange(false)) cudaMemset(base + kGuardBytes + bytes, 0, 1);
    std::lock_guard<std::mutex> lk(g_guard_mu);
    g_guard[*p] = GuardRec{base, bytes};
    return cudaSuccess;
}

cudaError_t guarded_free(void* p) {
    if (!p) return cudaSuccess;
    if (guard_on()) {
        GuardRec r{nullptr, 0};
        {
            std::lock_guard<std::mutex> lk(g_guard_mu);
            auto it = g_guard.find(p);
            if (it != g_guard.end()) { r = it->second; g_guard.erase(it); }
        }
        if (r.base) {
            std::vector<unsigned char> h(2 * kGuardBytes);
            cudaDeviceSynchronize();
            cudaMemcpy(h.data(), r.base, kGuardBytes, cudaMemcpyDeviceToHost);
            cudaMemcpy(h.data() + kGuardBytes, r.base + kGuardBytes + r.bytes, kGuardBytes, cudaMemcpyDeviceToHost);
            bool bad = false;
            for (unsigned char c : h) bad |= (c != kGuardByte);
            if (bad) ++g_guard_violations;
            ++g_guard_arrays;
            return cudaFree(r.base);
        }
    }
    return cudaFree(p);
}

template <class T>
int dev_alloc(T** p, size_t n) {
    CU(guarded_malloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    return SB_OK;
}

}  // namespace

struct sb_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    int refs;   // the creator's reference + one per live problem: a context outlives its problems whatever the order
                // in which a garbage-collected host language finalises the two handles
};

static void ctx_unref(sb_ctx* ctx) {
    if (--ctx->refs > 0) return;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

struct sb_problem {
    sb_ctx* ctx;
    const FunctorEntry* fe;  // registry entry (nullptr for user functors)
    KernelSet ks;
    int nparams = 0;
    ProbDev pd;
    int P, dec, sym0, warp, design, design_pending;  // dec: index into kDecomp
    int* d_list = nullptr;          // 2*B compacted active lists (fused_tma design)
    // long-mesh path (K6): levels of the recursive partition
    bool longmesh = false;
    std::vector<LevelDev> lv;
    int top_R = 1;
    bool top_configured = false;
    double *d_lm = nullptr, *d_partial = nullptr;
    double* h_prm = nullptr;   // pinned staging of the parameter blocks (sb_problem_set_params)
    // pipelined upload (sb_solve_from_host): boundary nodes of every instance (pinned / device), chunk-arrival word and the
    // pinned values that are copied into it, epoch of the words (they only ever grow, so a stale word never satisfies a wait)
    double *h_bnd = nullptr, *d_bnd = nullptr;
    unsigned int *d_ready = nullptr, *h_ready_vals = nullptr;
    unsigned int ready_epoch = 0;
    int* d_auto = nullptr;              // device-side stepping: {time step, iteration} (sb_solve_steps)
    double* d_tgrid = nullptr;          // its time grid
    int tgrid_cap = 0;
    int p_grid = 0;                 // persistent grid of the fused_tma kernel (0: not configured yet)
    size_t p_smem = 0;
    std::vector<double> xmesh, xi, zeta;
    // device allocations
    double *d_mesh = nullptr, *d_prm = nullptr, *d_u = nullptr, *d_b = nullptr, *d_F = nullptr, *d_J = nullptr,
           *d_mass = nullptr, *d_norm = nullptr, *d_hist = nullptr, *d_tmp = nullptr;
    size_t tmp_elems = 0;
    int *d_active = nullptr, *d_iters = nullptr, *d_status = nullptr;
    long long* d_total = nullptr;
    unsigned int* d_count = nullptr;
    unsigned int* h_count = nullptr;  // mapped pinned: the kernels publish the active count of each launch here
    // step state
    double t_next = 0.0, inv_tau = 0.0;
    long launched = 0, confirmed = 0;   // Newton-iteration launches / delivered counts of the current step
    long slot_base = 0;                 // ring slot of the step's iteration 0 (the ring runs on across steps)
    long total_launched = 0;            // launches through the ring since the problem was created
    long predicted_iters = 0;           // Newton iterations the previous solve needed (0: unknown) -- speculation limit
    int slot_of(long it) const { return (int)((slot_base + it) % kSlots); }
    int64_t last_active = 0;
    bool mass_ready = false;
    // begin_step fusion (fused_tma): b = M.*u_{n+1} is written by the converging iteration of the previous step
    bool b_ready = false;        // b already holds M.*u for the next step
    bool lazy_pending = false;   // the current step started without k_begin_step and no iteration has run yet
    bool step_all_tma = false;   // every Newton iteration of the current step so far was a k_newton_fused_p launch: only that
                                 // kernel writes b = M.*u_{n+1} when an instance converges and maintains the compacted list
    // whole-solve kernel (sb_resident.cuh)
    int r_grid = 0;              // its grid (0: not configured yet)
    double* d_itgrid = nullptr;  // per-instance time grids [B][nsteps+1]
    size_t itgrid_cap = 0;
    int* d_step_iters = nullptr; // [B][nsteps]
    size_t step_iters_cap = 0;
    double* d_snap = nullptr;    // [nsteps+1][B][Nx*P] snapshots, reference layout
    size_t snap_cap = 0;
    unsigned int* d_rmisc = nullptr;      // fail_step | max_iters | sum of iterations (2 words) | round counters [kMaxRounds]
    unsigned int* h_rounds = nullptr;     // mapped pinned: round-complete flags [kMaxRounds]
    cudaStream_t copy_stream = nullptr;   // D2H of finished snapshot rounds while later rounds compute
    double* d_piv = nullptr;     // workspace pool of the pivoted fallback solve (band arrays) ...
    int* d_piv_lock = nullptr;   // ... and its slot locks
    double* d_uin = nullptr;     // scratch state in device layout for sb_assemble / sb_assemble_device / sb_pdeval (kept)
    double* d_eval = nullptr;    // scratch of sb_pdeval: mesh | points | results (kept, grown on demand)
    size_t eval_elems = 0;
    int step_design = -1;
    double* d_ckpt = nullptr;  // sb_state_checkpoint copy of u (device layout)
    // profiling (sb_profile_begin/end): timed events around every Newton-iteration launch
    bool prof = false, prof_ev_ok = false;
    cudaEvent_t evA[kSlots], evM[kSlots], evB[kSlots];
    long prof_harvested = 0;
    int64_t prof_prev_active = 0;
    double prof_ms_assemble = 0.0, prof_ms_solve = 0.0;
    int64_t prof_launches = 0, prof_noop = 0, prof_inst_iters = 0;
};

extern "C" {
static int materialize_begin(sb_problem* pb);
static int run_begin_step_kernel(sb_problem* pb);
}

namespace {

size_t node_elems(const sb_problem* pb) { return (size_t)pb->pd.B * pb->pd.L * pb->P; }

int ensure_tmp(sb_problem* pb, size_t elems) {
    if (pb->tmp_elems >= elems) return SB_OK;
    if (pb->d_tmp) guarded_free(pb->d_tmp);
    pb->d_tmp = nullptr; pb->tmp_elems = 0;
    CU(guarded_malloc((void**)&pb->d_tmp, elems * sizeof(double)));
    pb->tmp_elems = elems;
    return SB_OK;
}

int ensure_jacobian(sb_problem* pb) {
    if (pb->d_J) return SB_OK;
    const size_t n = node_elems(pb) * pb->P;
    CU(guarded_malloc((void**)&pb->d_J, 3 * n * sizeof(double)));
    pb->pd.Jl = pb->d_J; pb->pd.Jd = pb->d_J + n; pb->pd.Ju = pb->d_J + 2 * n;
    return SB_OK;
}

// the one-CTA-per-system kernels of wide systems (npde >= 3) exchange more than the default 48 KB of shared memory
int configure_small_kernels(sb_problem* pb) {
    if (pb->ks.small_configured) return SB_OK;
    const size_t smem = pb->warp ? 0 : solve_smem_bytes(pb->P, pb->dec);
    if (smem > 48 * 1024) {
        if (pb->ks.solve) CU(cudaFuncSetAttribute(pb->ks.solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (pb->ks.fused) CU(cudaFuncSetAttribute(pb->ks.fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    pb->ks.small_configured = true;
    return SB_OK;
}

void launch_geometry(const sb_problem* pb, dim3& grid, dim3& block, size_t& smem) {
    if (pb->warp) {
        block = dim3(kWarpBlockThreads);
        grid = dim3((unsigned)((pb->pd.B + (kWarpBlockThreads / 32) - 1) / (kWarpBlockThreads / 32)));
    } else {
        block = dim3(pb->pd.T);
        grid = dim3((unsigned)pb->pd.B);
    }
    smem = solve_smem_bytes(pb->P, pb->dec);
}

int to_device(sb_problem* pb, const double* src_ref_dev, double* dst, int W, int identity_P) {
    const size_t n = (size_t)pb->pd.B * pb->pd.L * W;
    const int bs = 256;
    k_to_device_layout<<<(unsigned)((n + bs - 1) / bs), bs, 0, pb->ctx->stream>>>(src_ref_dev, dst, pb->pd.B, pb->pd.Nx, pb->pd.R,
                                                                               pb->pd.T, pb->pd.L, W, identity_P);
    ++g_launches;
    CU(cudaGetLastError());
    return SB_OK;
}

int from_device(sb_problem* pb, const double* src, double* dst_ref_dev, int W) {
    const size_t n = (size_t)pb->pd.B * pb->pd.Nx * W;
    const int bs = 256;
    k_from_device_layout<<<(unsigned)((n + bs - 1) / bs), bs, 0, pb->ctx->stream>>>(src, dst_ref_dev, pb->pd.B, pb->pd.Nx, pb->pd.R,
                                                                                 pb->pd.T, pb->pd.L, W);
    ++g_launches;
    CU(cudaGetLastError());
    return SB_OK;
}

// host array (reference layout, W per node) -> device array in device layout, through d_tmp
int upload_nodes(sb_problem* pb, const double* host, double* dst, int W, int identity_P) {
    const size_t n = (size_t)pb->pd.B * pb->pd.Nx * W;
    int rc = ensure_tmp(pb, n);
    if (rc) return rc;
    CU(cudaMemcpyAsync(pb->d_tmp, host, n * sizeof(double), cudaMemcpyHostToDevice, pb->ctx->stream));
    return to_device(pb, pb->d_tmp, dst, W, identity_P);
}

int download_nodes(sb_problem* pb, const double* src, double* host, int W) {
    const size_t n = (size_t)pb->pd.B * pb->pd.Nx * W;
    int rc = ensure_tmp(pb, n);
    if (rc) return rc;
    rc = from_device(pb, src, pb->d_tmp, W);
    if (rc) return rc;
    CU(cudaMemcpyAsync(host, pb->d_tmp, n * sizeof(double), cudaMemcpyDeviceToHost, pb->ctx->stream));
    CU(cudaStreamSynchronize(pb->ctx->stream));
    return SB_OK;
}

// Choose the thread decomposition (index into kDecomp).  Default: the smallest warp-per-system
// decomposition that covers the mesh with R <= Rpref rows per thread, else the CTA-per-system one with
// the preferred R and the fewest threads.  SB_ROWS_PER_THREAD / SB_THREADS_PER_SYSTEM override (tuning).
int pick_decomposition(int Nx, int P) {
    // measured: (8, 32) beats (4, 64) for npde = 2 at Nx = 256; (8, 128) is best for npde = 1 at Nx = 1024; npde = 3, 4: what
    // the register file allows (max_rows_per_thread)
    const int Rpref = std::min(8, max_rows_per_thread(P));
    int want_R = 0, want_T = 0;
    if (const char* e = getenv("SB_ROWS_PER_THREAD")) want_R = atoi(e);
    if (const char* e = getenv("SB_THREADS_PER_SYSTEM")) want_T = atoi(e);
    if (want_R || want_T) {
        int best = -1;
        for (int d = 0; d < kNumDecomp; ++d) {
            if (!decomp_ok(P, d) || kDecomp[d].R * kDecomp[d].TT < Nx) continue;
            if (want_R && kDecomp[d].R != want_R) continue;
            if (want_T && kDecomp[d].TT != want_T) continue;
            if (best < 0 || kDecomp[d].R * kDecomp[d].TT < kDecomp[best].R * kDecomp[best].TT) best = d;
        }
        if (best >= 0) return best;
    }
    int best = -1;
    long best_cost = 0;
    for (int d = 0; d < kNumDecomp; ++d) {
        if (!decomp_ok(P, d) || kDecomp[d].R * kDecomp[d].TT < Nx) continue;
        const int R = kDecomp[d].R, TT = kDecomp[d].TT;
        // cost: padded size first, then distance from the preferred R (warp mode only up to Rpref)
        long cost = (long)R * TT * 64 + (R > Rpref ? 32 * (R / Rpref) : (Rpref / R)) + (TT == 32 ? 0 : 1);
        if (TT == 32 && R > Rpref) cost += (long)R * TT * 64;   // prefer a CTA decomposition over R > Rpref warps
        if (best < 0 || cost < best_cost) { best = d; best_cost = cost; }
    }
    return best;
}

// ---- user functor plugin (NVRTC) ------------------------------------------------------------------
#ifndef SB_TMA_STAGES
#define SB_TMA_STAGES 1
#endif
#ifndef SB_PERSISTENT_WARPS_PER_SM
#define SB_PERSISTENT_WARPS_PER_SM 16
#endif

std::string csrc_dir() {
    if (const char* e = getenv("SB_CSRC_DIR")) return e;
    Dl_info info;
    if (dladdr((const void*)&sb_version, &info) && info.dli_fname) {
        std::string so = info.dli_fname;
        const size_t slash = so.find_last_of('/');
        return (slash == std::string::npos ? std::string(".") : so.substr(0, slash)) + "/csrc";
    }
    return "csrc";
}

// NVRTC stage (needs no GPU): source of the user functor -> CUBIN of the kernels of one problem shape
// (decomposition R x TT, coordinate symmetry, long mesh or not) + their lowered names
int nvrtc_user_kernels(const UserFunctor& uf, int R, int TT, bool sym0, bool longmesh, std::vector<char>& cubin,
                       std::vector<std::string>& lowered) {
    const char* sym = sym0 ? "true" : "false";
    char buf[256];
    std::string src = "#include \"sb_longmesh.cuh\"\nusing namespace sb;\n" + uf.source + "\n";
    snprintf(buf, sizeof buf, "__device__ unsigned long long sb_user_fused_p_smem = sb::FusedPLayout<%s, %d, %d, %s>::bytes;\n",
             uf.name.c_str(), R, TT, sym);
    src += buf;
    snprintf(buf, sizeof buf, "__device__ unsigned long long sb_user_resident_smem = sb::ResidentLayout<%s, %d, %d, %s>::bytes;\n",
             uf.name.c_str(), R, TT, sym);
    src += buf;
    if (longmesh) {
        if (uf.npde == 1)
            snprintf(buf, sizeof buf, "__device__ unsigned long long sb_user_long_smem = sb::LongStage<%s, %d, %d, %s>::bytes;\n",
                     uf.name.c_str(), R, TT, sym);
        else
            snprintf(buf, sizeof buf, "__device__ unsigned long long sb_user_long_smem = sb::long_sys_smem_bytes(%d);\n", uf.npde);
        src += buf;
    }
    snprintf(buf, sizeof buf, "static_assert(%s::npde == %d && %s::nparams == %d, \"npde/nparams differ from the registration\");\n",
             uf.name.c_str(), uf.npde, uf.name.c_str(), uf.nparams);
    src += buf;
    std::vector<std::string> names;
    auto name = [&](const char* fmt) { snprintf(buf, sizeof buf, fmt, uf.name.c_str(), R, TT, sym); names.push_back(buf); };
    snprintf(buf, sizeof buf, "sb::k_mass_flags<%s>", uf.name.c_str()); names.push_back(buf);
    name("sb::k_assemble<%s, %d, %d, %s, true>");
    name("sb::k_assemble<%s, %d, %d, %s, false>");
    name("sb::k_newton_fused<%s, %d, %d, %s>");
    name("sb::k_newton_fused_p<%s, %d, %d, %s>");
    name("sb::k_solve_resident<%s, %d, %d, %s>");
    snprintf(buf, sizeof buf, "sb::k_solve<%d, %d, %d>", uf.npde, R, TT); names.push_back(buf);
    name("sb::k_solve_resident_pipe<%s, %d, %d, %s>");
    if (longmesh && uf.npde == 1) {
        name("sb::k6_chunk_asm<%s, %d, %d, %s>");
        name("sb::k6_back_asm<%s, %d, %d, %s>");
        name("sb::k6_assemble<%s, %d, %d, %s, true>");
        name("sb::k6_assemble<%s, %d, %d, %s, false>");
    } else if (longmesh) {
        name("sb::k6s_chunk_asm<%s, %d, %d, %s>");
        name("sb::k6s_back_asm<%s, %d, %d, %s>");
        name("sb::k6s_assemble<%s, %d, %d, %s, true>");
        name("sb::k6s_assemble<%s, %d, %d, %s, false>");
    }
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, src.c_str(), "sb_user_functor.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
        return fail(SB_ERR_PLUGIN, "nvrtcCreateProgram failed");
    for (const std::string& n : names) nvrtcAddNameExpression(prog, n.c_str());
    const std::string inc = "-I" + csrc_dir();
    char d1[64], d2[64], d3[64], d4[64];
    snprintf(d1, sizeof d1, "-DSB_TMA_STAGES=%d", (int)SB_TMA_STAGES);
    snprintf(d2, sizeof d2, "-DSB_PERSISTENT_WARPS_PER_SM=%d", (int)SB_PERSISTENT_WARPS_PER_SM);
    snprintf(d3, sizeof d3, "-DSB_PERSISTENT_WARPS_P2=%d", (int)SB_PERSISTENT_WARPS_P2);
    snprintf(d4, sizeof d4, "-DSB_LONG_T0=%d", (int)SB_LONG_T0);
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", inc.c_str(), d1, d2, d3, d4};
    const nvrtcResult r = nvrtcCompileProgram(prog, 8, opts);
    if (r != NVRTC_SUCCESS) {
        size_t ls = 0;
        nvrtcGetProgramLogSize(prog, &ls);
        std::string log(ls + 1, '\0');
        nvrtcGetProgramLog(prog, &log[0]);
        nvrtcDestroyProgram(&prog);
        return fail(SB_ERR_PLUGIN, "NVRTC compilation of functor %s failed: %s\n%.700s", uf.name.c_str(), nvrtcGetErrorString(r), log.c_str());
    }
    size_t cs = 0;
    nvrtcGetCUBINSize(prog, &cs);
    cubin.resize(cs);
    nvrtcGetCUBIN(prog, cubin.data());
    lowered.clear();
    for (size_t i = 0; i < names.size(); ++i) {
        const char* low = nullptr;
        nvrtcGetLoweredName(prog, names[i].c_str(), &low);
        lowered.push_back(low ? low : "");
    }
    nvrtcDestroyProgram(&prog);
    return SB_OK;
}

// Compile the kernels this problem needs (its decomposition, its coordinate symmetry) for a user functor and load them.
struct CompiledShape { std::vector<char> cubin; std::vector<std::string> lowered; };
std::mutex g_cubin_mu;
std::map<std::string, CompiledShape>& cubin_cache() { static std::map<std::string, CompiledShape> m; return m; }

int compile_user_kernels(sb_problem* pb, const UserFunctor& uf) {
    // one NVRTC run per (functor, problem shape); later problems of the same shape load the cached CUBIN
    char key[512];
    snprintf(key, sizeof key, "%p|%s|%d|%d|%d|%d", (const void*)&uf, uf.name.c_str(), pb->pd.R, pb->pd.T, pb->sym0, (int)pb->longmesh);
    std::vector<char> cubin;
    std::vector<std::string> lowered;
    bool hit = false;
    {
        std::lock_guard<std::mutex> lk(g_cubin_mu);
        auto it = cubin_cache().find(key);
        if (it != cubin_cache().end()) { cubin = it->second.cubin; lowered = it->second.lowered; hit = true; }
    }
    if (!hit) {
        int rc = nvrtc_user_kernels(uf, pb->pd.R, pb->pd.T, pb->sym0 != 0, pb->longmesh, cubin, lowered);
        if (rc != SB_OK) return rc;
        std::lock_guard<std::mutex> lk(g_cubin_mu);
        cubin_cache()[key] = CompiledShape{cubin, lowered};
    }
    KernelSet& ks = pb->ks;
    cudaError_t e = cudaLibraryLoadData(&ks.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) return fail(SB_ERR_PLUGIN, "cudaLibraryLoadData: %s", cudaGetErrorString(e));
    const void** slots[] = {&ks.mass, &ks.assemble_jac, &ks.assemble_val, &ks.fused, &ks.fused_p, &ks.resident, &ks.solve, &ks.resident_pipe,
                            &ks.long_chunk, &ks.long_back, &ks.long_assemble_jac, &ks.long_assemble_val};
    for (size_t i = 0; i < lowered.size(); ++i) {
        cudaKernel_t kern;
        e = cudaLibraryGetKernel(&kern, ks.lib, lowered[i].c_str());
        if (e != cudaSuccess) return fail(SB_ERR_PLUGIN, "cudaLibraryGetKernel(%s): %s", lowered[i].c_str(), cudaGetErrorString(e));
        *slots[i] = (const void*)kern;
    }
    void* dptr = nullptr; size_t bytes = 0;
    e = cudaLibraryGetGlobal(&dptr, &bytes, ks.lib, "sb_user_fused_p_smem");
    unsigned long long smem = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&smem, dptr, sizeof smem, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(SB_ERR_PLUGIN, "reading the shared-memory size of the compiled kernel: %s", cudaGetErrorString(e));
    ks.fused_p_smem = (size_t)smem;
    e = cudaLibraryGetGlobal(&dptr, &bytes, ks.lib, "sb_user_resident_smem");
    if (e == cudaSuccess) e = cudaMemcpy(&smem, dptr, sizeof smem, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(SB_ERR_PLUGIN, "reading the shared-memory size of the compiled resident kernel: %s", cudaGetErrorString(e));
    ks.resident_smem = (size_t)smem;
    if (pb->longmesh) {
        e = cudaLibraryGetGlobal(&dptr, &bytes, ks.lib, "sb_user_long_smem");
        if (e == cudaSuccess) e = cudaMemcpy(&smem, dptr, sizeof smem, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return fail(SB_ERR_PLUGIN, "reading the shared-memory size of the compiled long-mesh kernels: %s", cudaGetErrorString(e));
        ks.long_smem = (size_t)smem;
    }
    return SB_OK;
}

// ---- long-mesh path (K6, sb_longmesh.cuh) -------------------------------------------------------
int longmesh_setup(sb_problem* pb) {
    const int64_t B = pb->pd.B;
    const int P = pb->P;
    std::vector<LevelDev> lv;
    int n = pb->pd.Nx;
    const int max_top_rows = long_rows_top(P) * kLongTT;   // the top level is solved by one CTA of kLongTT threads
    int top_rows = max_top_rows;
    if (const char* e = getenv("SB_LONG_TOP_ROWS")) {   // testing: stored levels on meshes of moderate length
        top_rows = std::max(kLongTT, std::min(max_top_rows, atoi(e)));
    }
    for (int l = 0;; ++l) {
        LevelDev L;
        memset(&L, 0, sizeof L);
        L.n = n;
        const bool top = (l > 0 && n <= top_rows);
        if (top) { L.R = 1; while (L.R * kLongTT < n) L.R *= 2; }
        else L.R = (l == 0) ? pb->pd.R : long_rows_stored(P);
        const int tt = (l == 0) ? kLongT0 : kLongTT;
        L.ntiles = (n + L.R * tt - 1) / (L.R * tt);
        L.stride = (size_t)L.ntiles * L.R * tt;
        // rows of the next level: scalar PDEs reduce every level-0 tile to one row inside its CTA; a stored level, and
        // level 0 of a system, keeps one row per thread chunk (padding chunks included)
        L.nchunks = (l == 0 && P == 1) ? L.ntiles : L.ntiles * tt;
        lv.push_back(L);
        if (top) break;
        n = L.nchunks;
        if (l > 24) return fail(SB_ERR_UNSUPPORTED, "long-mesh recursion too deep");
    }
    if (lv[0].ntiles != pb->pd.ntiles) return fail(SB_ERR_INVALID, "long-mesh tiling mismatch");
    // doubles per stored row (a, c, d), per solution entry, per summary record
    const size_t rowd = (size_t)(2 * P * P + P + 1) / 2 * 2, sumd = (size_t)(5 * P * P + 2 * P + 7) / 8 * 8;
    size_t total = 0;
    for (size_t l = 0; l < lv.size(); ++l) {
        if (l > 0) total += ((l + 1 < lv.size() ? rowd : 0) + P) * lv[l].stride * B;   // stored rows (not at the top) + x
        if (l + 1 < lv.size()) { lv[l].up_stride = lv[l + 1].stride; lv[l].up_R = lv[l + 1].R; total += sumd * lv[l].up_stride * B; }
    }
    CU(guarded_malloc((void**)&pb->d_lm, total * sizeof(double)));
    CU(cudaMemset(pb->d_lm, 0, total * sizeof(double)));
    double* p = pb->d_lm;
    for (size_t l = 0; l < lv.size(); ++l) {
        if (l > 0) {
            if (l + 1 < lv.size()) { lv[l].acd = reinterpret_cast<double2*>(p); p += rowd * lv[l].stride * B; }
            lv[l].x = p; p += (size_t)P * lv[l].stride * B;
        }
        if (l + 1 < lv.size()) { lv[l].sum = reinterpret_cast<double2*>(p); p += sumd * lv[l].up_stride * B; }
    }
    // per-tile sums | edge copies [2*ntiles + 2][P] | (scalar PDEs) the chunks' separator coefficients
    const size_t nt = (size_t)pb->pd.ntiles;
    const size_t edge_d = (2 * nt + 2) * P, coef_d = (P == 1) ? 3 * (size_t)kLongT0 * nt : 0;
    CU(guarded_malloc((void**)&pb->d_partial, (size_t)B * (nt + edge_d + coef_d) * sizeof(double)));
    lv[0].edge = lv[1].edge = pb->d_partial + (size_t)B * nt;
    lv[0].coef = lv[1].coef = pb->d_partial + (size_t)B * (nt + edge_d);
    pb->lv = lv;
    pb->top_R = lv.back().R;
    return SB_OK;
}

template <int P, int RT>
cudaError_t launch_top(sb_problem* pb, const LevelDev& top, const LevelDev& below) {
    const size_t smem = (size_t)ChunkThomas<P, RT>::template smem_doubles<kLongTT>() * sizeof(double);
    const void* kern = nullptr;
    if constexpr (P == 1) kern = (const void*)k6_top<RT, kLongTT>;
    else kern = (const void*)k6s_top<P, RT, kLongTT>;
    if (smem > 48 * 1024 && !pb->top_configured) {     // (a problem has one top kernel: its P and its top_R)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        pb->top_configured = true;
    }
    return launch_pdl(kern, dim3((unsigned)pb->pd.B), dim3(kLongTT), smem, pb->ctx->stream, top, below, (const int*)pb->d_active);
}

// the levels between the two level-0 sweeps: stored levels up, the top, stored levels down
template <int P>
int longmesh_middle(sb_problem* pb) {
    cudaStream_t st = pb->ctx->stream;
    const unsigned B = (unsigned)pb->pd.B;
    std::vector<LevelDev>& lv = pb->lv;
    const int top = (int)lv.size() - 1;
    constexpr int RS = long_rows_stored(P);
    for (int l = 1; l < top; ++l) {
        if constexpr (P == 1)
            CU(launch_pdl((const void*)k6_chunk_sys<RS, kLongTT>, dim3(lv[l].ntiles, B), dim3(kLongTT), 0, st, lv[l], lv[l - 1], (const int*)pb->d_active));
        else
            CU(launch_pdl((const void*)k6s_chunk_sys<P, RS, kLongTT>, dim3(lv[l].ntiles, B), dim3(kLongTT), 0, st, lv[l], lv[l - 1], (const int*)pb->d_active));
    }
    const int RT = pb->top_R;
    if (RT <= 1) CU((launch_top<P, 1>(pb, lv[top], lv[top - 1])));
    else if (RT == 2) CU((launch_top<P, 2>(pb, lv[top], lv[top - 1])));
    else if (RT == 4) { if constexpr (long_rows_top(P) >= 4) CU((launch_top<P, 4>(pb, lv[top], lv[top - 1]))); }
    else if (RT == 8) { if constexpr (long_rows_top(P) >= 8) CU((launch_top<P, 8>(pb, lv[top], lv[top - 1]))); }
    else { if constexpr (long_rows_top(P) >= 16) CU((launch_top<P, 16>(pb, lv[top], lv[top - 1]))); }
    for (int l = top - 1; l >= 1; --l) {
        if constexpr (P == 1)
            CU(launch_pdl((const void*)k6_back_sys<RS, kLongTT>, dim3(lv[l].ntiles, B), dim3(kLongTT), 0, st, lv[l], lv[l + 1], (const int*)pb->d_active));
        else
            CU(launch_pdl((const void*)k6s_back_sys<P, RS, kLongTT>, dim3(lv[l].ntiles, B), dim3(kLongTT), 0, st, lv[l], lv[l + 1], (const int*)pb->d_active));
    }
    return SB_OK;
}

// one Newton iteration of the long-mesh path; publishes the active count like the batched kernels
int longmesh_iteration(sb_problem* pb, double tol, int slot) {
    cudaStream_t st = pb->ctx->stream;
    const unsigned B = (unsigned)pb->pd.B;
    std::vector<LevelDev>& lv = pb->lv;
    const size_t smem = pb->ks.long_smem;
    if (!pb->ks.long_configured) {  // one-time opt-in to the staged tiles' shared memory (static + dynamic may exceed 48 KB)
        CU(cudaFuncSetAttribute(pb->ks.long_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 64)));
        CU(cudaFuncSetAttribute(pb->ks.long_back, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 64)));
        pb->ks.long_configured = true;
    }
    // every kernel of the sequence is launched with programmatic stream serialization: launch latency and the
    // dependence-free prologues (barrier set-up, mesh points of the level-0 tiles) overlap the predecessor's tail
    CU(launch_pdl(pb->ks.long_chunk, dim3(lv[0].ntiles, B), dim3(kLongT0), pb->P == 1 ? smem : 0, st, pb->pd, pb->t_next, pb->inv_tau, lv[0]));
    int rc = SB_OK;
    switch (pb->P) {
        case 1: rc = longmesh_middle<1>(pb); break;
        case 2: rc = longmesh_middle<2>(pb); break;
        case 3: rc = longmesh_middle<3>(pb); break;
        default: rc = longmesh_middle<4>(pb); break;
    }
    if (rc) return rc;
    CU(launch_pdl(pb->ks.long_back, dim3(lv[0].ntiles, B), dim3(kLongT0), smem, st, pb->pd, pb->t_next, pb->inv_tau, lv[1], pb->d_partial));
    CU(launch_pdl((const void*)k6_finalize<0>, dim3(B), dim3(256), 0, st, pb->pd, (const double*)pb->d_partial, tol, slot));
    CU(cudaGetLastError());
    return SB_OK;
}

}  // namespace

extern "C" {

const char* sb_version(void) { return "skeelberzins_b200 0.2.0 (sm_100a)"; }
#ifndef SB_SOURCE_DIGEST
#define SB_SOURCE_DIGEST "unknown"
#endif
const char* sb_build_digest(void) { return SB_SOURCE_DIGEST; }
const char* sb_last_error(void) { return g_err.c_str(); }

int sb_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(SB_ERR_INVALID, "out is NULL");
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return SB_OK;
}

int sb_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return SB_OK;
}

int sb_debug_guard_violations(int64_t* violations, int64_t* arrays_checked) {
    if (!violations) return fail(SB_ERR_INVALID, "violations is NULL");
    *violations = g_guard_violations;
    if (arrays_checked) *arrays_checked = g_guard_arrays;
    return SB_OK;
}

int sb_kernel_launches(int64_t* count) {
    if (!count) return fail(SB_ERR_INVALID, "count is NULL");
    *count = g_launches;
    return SB_OK;
}

int sb_device_count(int* count) {
    if (!count) return fail(SB_ERR_INVALID, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return fail(SB_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    return SB_OK;
}

int sb_ctx_create(int device, void* stream, sb_ctx** out) {
    if (!out) return fail(SB_ERR_INVALID, "out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(SB_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(SB_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(SB_ERR_CUDA, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
    sb_ctx* c = new sb_ctx;
    c->device = device;
    c->refs = 1;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else {
        cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e2 != cudaSuccess) { delete c; return fail(SB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e2)); }
        c->own_stream = true;
    }
    *out = c;
    return SB_OK;
}

int sb_ctx_destroy(sb_ctx* ctx) {
    if (!ctx) return SB_OK;
    ctx_unref(ctx);   // destroyed once the last problem created on it is gone
    return SB_OK;
}

int sb_ctx_synchronize(sb_ctx* ctx) {
    if (!ctx) return fail(SB_ERR_INVALID, "ctx is NULL");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

int sb_user_functor_compile_check(int functor_id, int Nx, int m, size_t* cubin_bytes) {
    const UserFunctor* ufp = user_functor(functor_id);
    if (!ufp) return fail(SB_ERR_INVALID, "functor id %d is not a registered user functor", functor_id);
    const UserFunctor& uf = *ufp;
    int dec = pick_decomposition(Nx, uf.npde);
    bool longmesh = false;
    int R, TT;
    if (dec < 0) { longmesh = true; R = long_rows_level0(uf.npde); TT = kLongT0; }
    else { R = kDecomp[dec].R; TT = kDecomp[dec].TT; }
    std::vector<char> cubin;
    std::vector<std::string> lowered;
    int rc = nvrtc_user_kernels(uf, R, TT, m == 0, longmesh, cubin, lowered);
    if (rc == SB_OK && cubin_bytes) *cubin_bytes = cubin.size();
    return rc;
}

int sb_functor_info(int functor_id, int* npde, int* nparams, const char** name) {
    if (const UserFunctor* up = user_functor(functor_id)) {
        const UserFunctor& u = *up;
        if (npde) *npde = u.npde;
        if (nparams) *nparams = u.nparams;
        if (name) *name = u.name.c_str();
        return SB_OK;
    }
    if (functor_id < 0 || functor_id >= SB_NUM_BUILTIN_FUNCTORS) return fail(SB_ERR_INVALID, "unknown functor id %d", functor_id);
    const FunctorEntry& e = registry()[functor_id];
    if (npde) *npde = e.npde;
    if (nparams) *nparams = e.nparams;
    if (name) *name = e.name;
    return SB_OK;
}

int sb_register_user_functor(sb_ctx* ctx, const char* cuda_source, const char* struct_name, int npde, int nparams,
                             int* functor_id) {
    (void)ctx;
    if (!cuda_source || !struct_name || !functor_id) return fail(SB_ERR_INVALID, "NULL argument");
    if (npde < 1 || npde > SB_MAX_NPDE) return fail(SB_ERR_UNSUPPORTED, "npde = %d: kernels are compiled for npde <= %d", npde, SB_MAX_NPDE);
    if (nparams < 1) return fail(SB_ERR_INVALID, "a functor needs at least one per-instance parameter");
    UserFunctor uf;
    uf.source = cuda_source; uf.name = struct_name; uf.npde = npde; uf.nparams = nparams;
    *functor_id = add_user_functor(uf);
    return SB_OK;   // kernels are compiled when a problem is created (they depend on the mesh decomposition)
}

int sb_problem_create(sb_ctx* ctx, int m, int Nx, int64_t B, const double* xmesh, int functor_id, const double* params,
                      int nparams, sb_problem** out) {
    if (!ctx || !out || !xmesh) return fail(SB_ERR_INVALID, "NULL argument");
    // asserts of src/pdepe.jl:55-57
    if (!(m == 0 || m == 1 || m == 2)) return fail(SB_ERR_INVALID, "Parameter m invalid");
    if (m > 0 && !(xmesh[0] >= 0)) return fail(SB_ERR_INVALID, "Non-conforming mesh");
    if (Nx < 3) return fail(SB_ERR_INVALID, "Nx = %d: need at least 3 mesh points", Nx);
    if (B < 1) return fail(SB_ERR_INVALID, "B = %lld: need at least one instance", (long long)B);
    for (int i = 0; i + 1 < Nx; ++i)
        if (!(xmesh[i + 1] > xmesh[i])) return fail(SB_ERR_INVALID, "xmesh must be strictly increasing (at index %d)", i);
    const FunctorEntry* fe = nullptr;
    const UserFunctor* uf = nullptr;
    int f_npde, f_nparams;
    const char* f_name;
    if (functor_id >= 0 && functor_id < SB_NUM_BUILTIN_FUNCTORS) {
        fe = &registry()[functor_id];
        f_npde = fe->npde; f_nparams = fe->nparams; f_name = fe->name;
    } else if ((uf = user_functor(functor_id)) != nullptr) {
        f_npde = uf->npde; f_nparams = uf->nparams; f_name = uf->name.c_str();
    } else {
        return fail(SB_ERR_INVALID, "unknown functor id %d", functor_id);
    }
    if (nparams != f_nparams)
        return fail(SB_ERR_INVALID, "functor %s takes %d parameters per instance, got %d", f_name, f_nparams, nparams);
    if (f_nparams > 0 && !params) return fail(SB_ERR_INVALID, "params is NULL");
    CU(cudaSetDevice(ctx->device));

    sb_problem* pb = new sb_problem;
    ctx->refs++;
    pb->ctx = ctx; pb->fe = fe; pb->P = f_npde; pb->nparams = f_nparams; pb->design = pb->design_pending = SB_DESIGN_RESIDENT;
    pb->dec = pick_decomposition(Nx, pb->P);
    if (getenv("SB_FORCE_LONGMESH")) pb->dec = -1;   // testing: long-mesh path on short meshes
    int R, T, ntiles = 1;
    if (pb->dec < 0) {
        pb->longmesh = true;
        R = long_rows_level0(pb->P); T = kLongT0;
        ntiles = (Nx + R * T - 1) / (R * T);
        for (int d = 0; d < kNumDecomp; ++d) if (kDecomp[d].R == R && kDecomp[d].TT == T) pb->dec = d;
    } else {
        R = kDecomp[pb->dec].R; T = kDecomp[pb->dec].TT;
    }
    pb->warp = (T == 32);
    pb->sym0 = (m == 0);
    ProbDev& pd = pb->pd;
    memset(&pd, 0, sizeof pd);
    pd.Nx = Nx; pd.R = R; pd.T = T; pd.Lt = R * T; pd.ntiles = ntiles; pd.L = ntiles * R * T; pd.m = m; pd.nprm = f_nparams; pd.nprm_s = (f_nparams + 1) / 2 * 2; pd.iter_no = 0; pd.B = B;
    pd.singular = (m >= 1) && (xmesh[0] == 0);  // src/utils.jl:616
    pb->xmesh.assign(xmesh, xmesh + Nx);

    // quadrature points / weights, src/utils.jl:683-718
    const int NI = Nx - 1, mp = m + 1;
    pb->xi.resize(NI); pb->zeta.resize(NI);
    for (int i = 0; i < NI; ++i) {
        const double a = xmesh[i], b = xmesh[i + 1], g = (a + b) / 2;
        double xi, ze;
        if (m == 0) { xi = g; ze = g; }
        else if (m == 1) {
            if (pd.singular) { xi = (2.0 / 3.0) * (a + b - ((a * b) / (a + b))); ze = std::pow((b * b - a * a) / (2 * std::log(b / a)), 0.5); }
            else { xi = (b - a) / std::log(b / a); ze = std::pow(xi * g, 0.5); }
        } else {
            if (pd.singular) xi = (2.0 / 3.0) * (a + b - ((a * b) / (a + b)));
            else xi = (a * b * std::log(b / a)) / (b - a);
            ze = std::pow(a * b * g, 1.0 / 3.0);
        }
        pb->xi[i] = xi; pb->zeta[i] = ze;
    }
    // per-interval interpolation weights (src/utils.jl:18-48) and flux weights, per-node volume weights
    // (src/assembler.jl:36,68-69,74,84,98), written straight in the [r][t] device layout
    const size_t L = pd.L;
    // device image: xi[L] | awbw[2L] | gwwf[2L] | frac[2L] | gw[L] | ivol[L] | xs[L].  m == 0: only xs is staged, the
    // device computes the rest (k_mesh_weights_m0); m >= 1 (logarithms, cube roots: libm on the host, as the oracle)
    // fills the whole image here
    const bool device_weights = (m == 0);
    std::vector<double> mesh((device_weights ? 1 : 10) * L, 0.0);
    double* h_xs = &mesh[device_weights ? 0 : 9 * L];
    for (size_t i = 0; i < L; ++i) {   // the mesh points themselves (padding: the mesh continued with its last spacing)
        const size_t s = (i / (R * T)) * (R * T) + ((i % (R * T)) % R) * T + (i % (R * T)) / R;
        h_xs[s] = (i < (size_t)Nx) ? xmesh[i] : xmesh[Nx - 1] + (double)(i - (Nx - 1)) * (xmesh[Nx - 1] - xmesh[Nx - 2]);
    }
    if (!device_weights) {
        double *h_xi = &mesh[0], *h_awbw = &mesh[L], *h_gwwf = &mesh[3 * L], *h_frac = &mesh[5 * L], *h_gw = &mesh[7 * L],
               *h_ivol = &mesh[8 * L];
        for (int i = 0; i < Nx; ++i) {
            const size_t s = (size_t)(i / (R * T)) * (R * T) + (size_t)((i % (R * T)) % R) * T + (i % (R * T)) / R;
            double fR = 0.0, fL = 0.0;
            if (i < NI) {
                const double xl = xmesh[i], xr = xmesh[i + 1], q = pb->xi[i];
                double aw, bw, gw;
                if (pd.singular) { const double h = xr * xr - xl * xl, th = (q * q - xl * xl) / h; aw = 1 - th; bw = th; gw = (2 * q) / h; }
                else if (m == 1) { const double lg = std::log(xr / xl), th = std::log(q / xl) / lg; aw = 1 - th; bw = th; gw = 1.0 / (q * lg); }
                else { const double h = xr - xl, th = (xr * (q - xl)) / (q * h); aw = 1 - th; bw = th; gw = (xr * xl) / (q * q * h); }
                h_xi[s] = q; h_awbw[2 * s] = aw; h_awbw[2 * s + 1] = bw; h_gwwf[2 * s] = gw; h_gw[s] = gw;
                h_gwwf[2 * s + 1] = pd.singular ? ipow(pb->zeta[i], mp) / q : ipow(q, m);
                fR = (ipow(pb->zeta[i], mp) - ipow(xmesh[i], mp)) / mp;
            }
            if (i >= 1) fL = (ipow(xmesh[i], mp) - ipow(pb->zeta[i - 1], mp)) / mp;
            h_frac[2 * s] = fR; h_frac[2 * s + 1] = fL;
            if (i >= 1 && i < NI) h_ivol[s] = 1.0 / (fR + fL);
        }
    }
    pd.x0 = xmesh[0]; pd.xN = xmesh[Nx - 1]; pd.x0m = ipow(xmesh[0], m); pd.xNm = ipow(xmesh[Nx - 1], m); pd.xi0 = pb->xi[0];

    int rc = SB_OK;
    const size_t n = (size_t)B * L * pb->P;
#define TRY(x) do { if (rc == SB_OK) rc = (x); } while (0)
    TRY(dev_alloc(&pb->d_mesh, 10 * L));
    TRY(dev_alloc(&pb->d_prm, (size_t)B * std::max(pd.nprm_s, 2)));
    TRY(dev_alloc(&pb->d_u, n));
    TRY(dev_alloc(&pb->d_b, n));
    TRY(dev_alloc(&pb->d_F, n));
    TRY(dev_alloc(&pb->d_mass, (size_t)B * 4 * pb->P));
    TRY(dev_alloc(&pb->d_norm, (size_t)B));
    TRY(dev_alloc(&pb->d_active, (size_t)B));
    TRY(dev_alloc(&pb->d_iters, (size_t)B));
    TRY(dev_alloc(&pb->d_status, (size_t)B));
    TRY(dev_alloc(&pb->d_total, (size_t)B));
    TRY(dev_alloc(&pb->d_count, (size_t)2 * kSlots));  // active counts [kSlots] + done tickets [kSlots]
    TRY(dev_alloc(&pb->d_list, (size_t)2 * B));
    constexpr int kPivSlots = 64;   // concurrent pivoted fallback solves (a group that finds none free waits for one)
    const size_t piv_stride = (size_t)Nx * pb->P * ((size_t)(6 * pb->P - 2) + 1);   // BandShape<P>::ldab = 3*(2P-1) + 1 band rows + the rhs
    if (!pb->longmesh) {
        TRY(dev_alloc(&pb->d_piv, kPivSlots * piv_stride));
        TRY(dev_alloc(&pb->d_piv_lock, (size_t)kPivSlots));
    }
#undef TRY
    if (rc != SB_OK) { sb_problem_destroy(pb); return rc; }
    cudaError_t e = cudaHostAlloc((void**)&pb->h_count, 3 * kSlots * sizeof(unsigned int), cudaHostAllocMapped);  // count | step | iteration
    if (e != cudaSuccess) { sb_problem_destroy(pb); return fail(SB_ERR_CUDA, "cudaHostAlloc: %s", cudaGetErrorString(e)); }
    unsigned int* d_hcount = nullptr;
    e = cudaHostGetDevicePointer((void**)&d_hcount, pb->h_count, 0);
    if (e != cudaSuccess) { sb_problem_destroy(pb); return fail(SB_ERR_CUDA, "cudaHostGetDevicePointer: %s", cudaGetErrorString(e)); }
    for (int i = 0; i < 3 * kSlots; ++i) pb->h_count[i] = 0;
    cudaStream_t st = ctx->stream;
    e = cudaMemcpyAsync(pb->d_mesh + (device_weights ? 9 * L : 0), mesh.data(), mesh.size() * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && device_weights) {
        k_mesh_weights_m0<<<(unsigned)((L + 255) / 256), 256, 0, st>>>(pb->d_mesh, Nx, R, T, L);
        ++g_launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && f_nparams > 0)
        e = cudaMemcpy2DAsync(pb->d_prm, pd.nprm_s * sizeof(double), params, f_nparams * sizeof(double),
                              f_nparams * sizeof(double), (size_t)B, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(pb->d_u, 0, n * sizeof(double), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(pb->d_b, 0, n * sizeof(double), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(pb->d_F, 0, n * sizeof(double), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(pb->d_active, 0, B * sizeof(int), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(pb->d_iters, 0, B * sizeof(int), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(pb->d_status, 0, B * sizeof(int), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(pb->d_total, 0, B * sizeof(long long), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(pb->d_norm, 0, B * sizeof(double), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(pb->d_count, 0, 2 * kSlots * sizeof(unsigned int), st);
    if (e == cudaSuccess && pb->d_piv_lock) e = cudaMemsetAsync(pb->d_piv_lock, 0, kPivSlots * sizeof(int), st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // host vectors go out of scope
    if (e != cudaSuccess) { sb_problem_destroy(pb); return fail(SB_ERR_CUDA, "problem upload: %s", cudaGetErrorString(e)); }
    // all-ones mass flags until sb_mass_flags is called
    {
        std::vector<double> ones((size_t)B * 4 * pb->P, 1.0);
        e = cudaMemcpy(pb->d_mass, ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { sb_problem_destroy(pb); return fail(SB_ERR_CUDA, "mass init: %s", cudaGetErrorString(e)); }
    }
    pd.mesh.xi = pb->d_mesh;
    pd.mesh.awbw = reinterpret_cast<const double2*>(pb->d_mesh + L);
    pd.mesh.gwwf = reinterpret_cast<const double2*>(pb->d_mesh + 3 * L);
    pd.mesh.frac = reinterpret_cast<const double2*>(pb->d_mesh + 5 * L);
    pd.mesh.gw = pb->d_mesh + 7 * L;
    pd.mesh.ivol = pb->d_mesh + 8 * L;
    pd.mesh.xs = pb->d_mesh + 9 * L;
    pd.prm = pb->d_prm; pd.u = pb->d_u; pd.b = pb->d_b; pd.F = pb->d_F; pd.mass = pb->d_mass;
    pd.active = pb->d_active; pd.iters = pb->d_iters; pd.total_iters = pb->d_total; pd.norm = pb->d_norm;
    pd.status = pb->d_status; pd.hist = nullptr; pd.hist_cap = 0; pd.active_count = pb->d_count;
    pd.list[0] = pb->d_list; pd.list[1] = pb->d_list + B;
    pd.done_count = pb->d_count + kSlots; pd.host_count = d_hcount;
    pd.piv_band = pb->d_piv; pd.piv_lock = pb->d_piv_lock; pd.piv_slots = pb->d_piv ? kPivSlots : 0; pd.piv_stride = piv_stride;
    pb->last_active = B;
    if (fe) {
        KernelSet& ks = pb->ks;
        ks.mass = (const void*)fe->mass;
        ks.assemble_jac = (const void*)fe->assemble_jac[pb->dec][pb->sym0];
        ks.assemble_val = (const void*)fe->assemble_val[pb->dec][pb->sym0];
        ks.fused = (const void*)fe->fused[pb->dec][pb->sym0];
        ks.fused_p = (const void*)fe->fused_p[pb->dec][pb->sym0];
        ks.fused_p_smem = fe->fused_p_smem[pb->dec][pb->sym0];
        ks.resident = (const void*)fe->resident[pb->dec][pb->sym0];
        ks.resident_pipe = (const void*)fe->resident_pipe[pb->dec][pb->sym0];
        ks.solve = (const void*)solve_kernel(pb->P, pb->dec);
        ks.resident_smem = fe->resident_smem[pb->dec][pb->sym0];
        ks.long_chunk = (const void*)fe->long_chunk[pb->sym0];
        ks.long_back = (const void*)fe->long_back[pb->sym0];
        ks.long_smem = fe->long_smem[pb->sym0];
        ks.long_assemble_jac = (const void*)fe->long_assemble_jac[pb->sym0];
        ks.long_assemble_val = (const void*)fe->long_assemble_val[pb->sym0];
    } else {
        int rcu = compile_user_kernels(pb, *uf);
        if (rcu != SB_OK) { sb_problem_destroy(pb); return rcu; }
    }
    if (pb->longmesh) {
        int rc2 = longmesh_setup(pb);
        if (rc2 != SB_OK) { sb_problem_destroy(pb); return rc2; }
    }
    *out = pb;
    return SB_OK;
}

int sb_problem_destroy(sb_problem* pb) {
    if (!pb) return SB_OK;
    cudaSetDevice(pb->ctx->device);
    cudaStreamSynchronize(pb->ctx->stream);
    guarded_free(pb->d_mesh); guarded_free(pb->d_prm); guarded_free(pb->d_u); guarded_free(pb->d_b); guarded_free(pb->d_F); guarded_free(pb->d_J);
    guarded_free(pb->d_mass); guarded_free(pb->d_norm); guarded_free(pb->d_hist); guarded_free(pb->d_tmp); guarded_free(pb->d_active);
    guarded_free(pb->d_iters); guarded_free(pb->d_status); guarded_free(pb->d_total); guarded_free(pb->d_count); guarded_free(pb->d_ckpt); guarded_free(pb->d_list); guarded_free(pb->d_lm); guarded_free(pb->d_partial); guarded_free(pb->d_auto); guarded_free(pb->d_tgrid); guarded_free(pb->d_uin); guarded_free(pb->d_eval); guarded_free(pb->d_piv); guarded_free(pb->d_piv_lock); guarded_free(pb->d_itgrid); guarded_free(pb->d_step_iters); guarded_free(pb->d_snap); guarded_free(pb->d_rmisc);
    if (pb->h_rounds) cudaFreeHost(pb->h_rounds);
    if (pb->h_prm) cudaFreeHost(pb->h_prm);
    if (pb->h_bnd) cudaFreeHost(pb->h_bnd);
    if (pb->h_ready_vals) cudaFreeHost(pb->h_ready_vals);
    guarded_free(pb->d_bnd); guarded_free(pb->d_ready);
    if (pb->copy_stream) cudaStreamDestroy(pb->copy_stream);
    if (pb->prof_ev_ok) for (int i = 0; i < kSlots; ++i) { cudaEventDestroy(pb->evA[i]); cudaEventDestroy(pb->evM[i]); cudaEventDestroy(pb->evB[i]); }
    if (pb->h_count) cudaFreeHost(pb->h_count);
    if (pb->ks.lib) cudaLibraryUnload(pb->ks.lib);
    ctx_unref(pb->ctx);
    delete pb;
    return SB_OK;
}

int sb_problem_info(sb_problem* pb, int* npde, int* Nx, int64_t* B, int* singular) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    if (npde) *npde = pb->P;
    if (Nx) *Nx = pb->pd.Nx;
    if (B) *B = pb->pd.B;
    if (singular) *singular = pb->pd.singular;
    return SB_OK;
}

int sb_problem_layout(sb_problem* pb, int* R, int* T, int* warp) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    if (R) *R = pb->pd.R;
    if (T) *T = pb->pd.T;
    if (warp) *warp = pb->warp;
    return SB_OK;
}

int sb_problem_quadrature(sb_problem* pb, double* xi, double* zeta) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    if (xi) memcpy(xi, pb->xi.data(), pb->xi.size() * sizeof(double));
    if (zeta) memcpy(zeta, pb->zeta.data(), pb->zeta.size() * sizeof(double));
    return SB_OK;
}

int sb_problem_set_design(sb_problem* pb, int design) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    if (design != SB_DESIGN_SPLIT && design != SB_DESIGN_FUSED && design != SB_DESIGN_FUSED_TMA && design != SB_DESIGN_RESIDENT)
        return fail(SB_ERR_INVALID, "unknown design %d", design);
    pb->design_pending = design;  // takes effect at the next sb_begin_step (a step is never mixed)
    if (pb->launched == 0) {
        if (!is_tma(design) && pb->lazy_pending) {   // only the fused_tma kernel can start from b alone
            CU(cudaSetDevice(pb->ctx->device));
            int rc = materialize_begin(pb);
            if (rc) return rc;
        }
        pb->design = design;
    }
    return SB_OK;
}

int sb_problem_set_params(sb_problem* pb, const double* params_host) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    if (pb->nparams == 0) return SB_OK;
    if (!params_host) return fail(SB_ERR_INVALID, "params is NULL");
    CU(cudaSetDevice(pb->ctx->device));
    // packed into the padded device stride in a pinned staging buffer, then ONE contiguous copy: a strided 2-D copy of
    // B rows of a few doubles from pageable memory costs 0.2 us per row (cfg4: 2.7 ms of a 4 ms call)
    const size_t n = (size_t)pb->pd.B * pb->pd.nprm_s;
    if (!pb->h_prm) CU(cudaHostAlloc((void**)&pb->h_prm, n * sizeof(double), cudaHostAllocDefault));
    const int np = pb->nparams, ns = pb->pd.nprm_s;
    for (int64_t k = 0; k < pb->pd.B; ++k) {
        for (int i = 0; i < np; ++i) pb->h_prm[(size_t)k * ns + i] = params_host[(size_t)k * np + i];
        for (int i = np; i < ns; ++i) pb->h_prm[(size_t)k * ns + i] = 0.0;
    }
    CU(cudaMemcpyAsync(pb->d_prm, pb->h_prm, n * sizeof(double), cudaMemcpyHostToDevice, pb->ctx->stream));
    CU(cudaStreamSynchronize(pb->ctx->stream));
    return SB_OK;
}

int sb_state_upload(sb_problem* pb, const double* u_host) {
    if (!pb || !u_host) return fail(SB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(pb->ctx->device));
    pb->b_ready = false; pb->lazy_pending = false; pb->pd.lazy_b = 0;
    int rc = upload_nodes(pb, u_host, pb->d_u, pb->P, 0);
    if (rc) return rc;
    CU(cudaMemsetAsync(pb->d_total, 0, pb->pd.B * sizeof(long long), pb->ctx->stream));
    CU(cudaMemsetAsync(pb->d_status, 0, pb->pd.B * sizeof(int), pb->ctx->stream));
    CU(cudaStreamSynchronize(pb->ctx->stream));
    return SB_OK;
}

int sb_state_download(sb_problem* pb, double* u_host) {
    if (!pb || !u_host) return fail(SB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(pb->ctx->device));
    { int rc = materialize_begin(pb); if (rc) return rc; }
    return download_nodes(pb, pb->d_u, u_host, pb->P);
}

int sb_mass_flags(sb_problem* pb, double t0, int stationary) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    CU(cudaSetDevice(pb->ctx->device));
    pb->b_ready = false;
    { int rc = materialize_begin(pb); if (rc) return rc; }
    const int bs = 128;
    CU(launch_any(pb->ks.mass, dim3((unsigned)((pb->pd.B + bs - 1) / bs)), dim3(bs), 0, pb->ctx->stream, pb->pd, t0, pb->d_mass, stationary, (const double*)nullptr));
    CU(cudaGetLastError());
    pb->mass_ready = true;
    return SB_OK;
}

int sb_mass_download(sb_problem* pb, double* mass_host) {
    if (!pb || !mass_host) return fail(SB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(pb->ctx->device));
    const int P = pb->P, Nx = pb->pd.Nx;
    std::vector<double> f((size_t)pb->pd.B * 4 * P);
    CU(cudaStreamSynchronize(pb->ctx->stream));
    CU(cudaMemcpy(f.data(), pb->d_mass, f.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t k = 0; k < pb->pd.B; ++k)
        for (int i = 0; i < Nx; ++i) {
            const int cls = (i == 0) ? 0 : ((i == Nx - 1) ? 2 : 1);
            for (int p = 0; p < P; ++p) mass_host[((size_t)k * Nx + i) * P + p] = f[(size_t)k * 4 * P + cls * P + p];
        }
    return SB_OK;
}

// persistent scratch state (device layout) for entry points that take a state argument
static int ensure_uin(sb_problem* pb) {
    if (pb->d_uin) return SB_OK;
    CU(guarded_malloc((void**)&pb->d_uin, node_elems(pb) * sizeof(double)));
    return SB_OK;
}

// assemble! (src/assembler.jl:19-125) of the state in d_state (device layout) into d_F; asynchronous on the stream
static int launch_assemble_val(sb_problem* pb, const double* d_state, double t) {
    ProbDev pd = pb->pd;
    pd.u = const_cast<double*>(d_state);
    dim3 grid, block; size_t smem;
    launch_geometry(pb, grid, block, smem);
    cudaError_t e;
    if (pb->longmesh) e = launch_any(pb->ks.long_assemble_val, dim3(pd.ntiles, (unsigned)pd.B), dim3(kLongT0), 0, pb->ctx->stream, pd, t, 0.0, 1);
    else e = launch_any(pb->ks.assemble_val, grid, block, 0, pb->ctx->stream, pd, t, 0.0, 1);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SB_ERR_CUDA, "k_assemble launch: %s", cudaGetErrorString(e));
    return SB_OK;
}

int sb_assemble(sb_problem* pb, const double* u_host, double t, double* du_host) {
    if (!pb || !du_host) return fail(SB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(pb->ctx->device));
    int rc;
    const double* d_state = pb->d_u;
    if (u_host) {
        if ((rc = ensure_uin(pb))) return rc;
        if ((rc = upload_nodes(pb, u_host, pb->d_uin, pb->P, 0))) return rc;
        d_state = pb->d_uin;
    } else if ((rc = materialize_begin(pb))) return rc;
    if ((rc = launch_assemble_val(pb, d_state, t))) return rc;
    return download_nodes(pb, pb->d_F, du_host, pb->P);
}

// The same operator on DEVICE memory, for an ODE/DAE integrator that keeps its vectors on the GPU (src/diffEq.jl:18-27:
// f(du, u, p, t) = assemble!(du, u, pb, t)).  d_u, d_du: device pointers, reference layout [B][Nx][npde]; d_u NULL: the
// resident state.  Enqueued on the context's stream; no allocation after the first call, no synchronisation.
int sb_assemble_device(sb_problem* pb, const double* d_u, double t, double* d_du) {
    if (!pb || !d_du) return fail(SB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(pb->ctx->device));
    int rc;
    const double* d_state = pb->d_u;
    if (d_u) {
        if ((rc = ensure_uin(pb))) return rc;
        if ((rc = to_device(pb, d_u, pb->d_uin, pb->P, 0))) return rc;
        d_state = pb->d_uin;
    } else if ((rc = materialize_begin(pb))) return rc;
    if ((rc = launch_assemble_val(pb, d_state, t))) return rc;
    return from_device(pb, pb->d_F, d_du, pb->P);
}

static int run_begin_step_kernel(sb_problem* pb) {
    const size_t total = (size_t)pb->pd.B * pb->pd.L;
    const int bs = 256;
    cudaStream_t st = pb->ctx->stream;
    k_begin_step<0><<<(unsigned)((total + bs - 1) / bs), bs, 0, st>>>(pb->pd, pb->P);
    ++g_launches;
    CU(cudaGetLastError());
    return SB_OK;
}

// a step that started lazily (no k_begin_step) is about to be observed or continued by something other than the
// fused_tma kernel: run the kernel now (u still holds the true u_n, so the result is the same)
static int materialize_begin(sb_problem* pb) {
    if (!pb->lazy_pending) return SB_OK;
    pb->lazy_pending = false;
    pb->pd.lazy_b = 0;
    return run_begin_step_kernel(pb);
}

// profiling: fold the timed events of iterations [prof_harvested, launched) of the current step into the totals
static int prof_harvest(sb_problem* pb) {
    if (!pb->prof || pb->prof_harvested >= pb->launched) return SB_OK;
    CU(cudaStreamSynchronize(pb->ctx->stream));
    for (long j = pb->prof_harvested; j < pb->launched; ++j) {
        const int s = pb->slot_of(j);
        const int64_t before = (j == 0) ? pb->pd.B : pb->prof_prev_active;
        float ms_a = 0.f, ms_s = 0.f;
        if (pb->design == SB_DESIGN_SPLIT) {
            CU(cudaEventElapsedTime(&ms_a, pb->evA[s], pb->evM[s]));
            CU(cudaEventElapsedTime(&ms_s, pb->evM[s], pb->evB[s]));
        } else {
            CU(cudaEventElapsedTime(&ms_s, pb->evA[s], pb->evB[s]));
        }
        if (before > 0) {
            pb->prof_ms_assemble += ms_a; pb->prof_ms_solve += ms_s;
            pb->prof_launches += 1; pb->prof_inst_iters += before;
        } else {
            pb->prof_noop += 1;
        }
        pb->prof_prev_active = pb->h_count[s];
    }
    pb->prof_harvested = pb->launched;
    return SB_OK;
}

int sb_pdeval(sb_problem* pb, const double* u_host, int nq, const double* xq_host, double* u_out, double* dudx_out) {
    if (!pb || !xq_host || !u_out || !dudx_out || nq < 1) return fail(SB_ERR_INVALID, "NULL argument or nq < 1");
    CU(cudaSetDevice(pb->ctx->device));
    const int Nx = pb->pd.Nx, P = pb->P;
    const std::vector<double>& x = pb->xmesh;
    for (int q = 0; q < nq; ++q) {       // src/utils.jl:403-414
        const double pt = xq_host[q];
        if (pt == x[0] || std::fabs(pt - x[0]) <= 1.0e-10 * std::fabs(x[1] - x[0])) continue;
        const long idx = (std::lower_bound(x.begin(), x.end(), pt) - x.begin()) + 1;
        if (idx == 1 || idx > Nx) return fail(SB_ERR_INVALID, "Evaluation point oustide of the spatial discretization.");
    }
    int rc = SB_OK;
    const double* d_u = pb->d_u;
    if (u_host) {
        if ((rc = ensure_uin(pb))) return rc;
        if ((rc = upload_nodes(pb, u_host, pb->d_uin, P, 0))) return rc;
        d_u = pb->d_uin;
    } else if ((rc = materialize_begin(pb))) return rc;
    const size_t no = (size_t)pb->pd.B * nq * P;
    const size_t need = (size_t)Nx + nq + 2 * no;              // mesh | points | u | dudx
    if (pb->eval_elems < need) {
        if (pb->d_eval) { CU(cudaStreamSynchronize(pb->ctx->stream)); guarded_free(pb->d_eval); pb->d_eval = nullptr; pb->eval_elems = 0; }
        CU(guarded_malloc((void**)&pb->d_eval, need * sizeof(double)));
        pb->eval_elems = need;
    }
    double *d_x = pb->d_eval, *d_q = d_x + Nx, *d_o = d_q + nq;
    cudaStream_t st = pb->ctx->stream;
    CU(cudaMemcpyAsync(d_x, x.data(), Nx * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_q, xq_host, nq * sizeof(double), cudaMemcpyHostToDevice, st));
    const size_t n = (size_t)pb->pd.B * nq;
    k_pdeval<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_u, d_x, Nx, pb->pd.R, pb->pd.T, pb->pd.L, P, pb->pd.m, pb->pd.singular,
                                                         pb->pd.B, nq, d_q, d_o, d_o + no);
    ++g_launches;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(u_out, d_o, no * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(dudx_out, d_o + no, no * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SB_OK;
}

int sb_begin_step(sb_problem* pb, double t_next, double inv_tau) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    CU(cudaSetDevice(pb->ctx->device));
    if (pb->prof) { int rc = prof_harvest(pb); if (rc) return rc; pb->prof_harvested = 0; }
    pb->t_next = t_next; pb->inv_tau = inv_tau;
    pb->design = pb->design_pending;
    const bool lazy = pb->b_ready && is_tma(pb->design) && !pb->longmesh && !getenv("SB_NO_LAZY_BEGIN");
    pb->b_ready = false;
    pb->lazy_pending = false;
    pb->pd.lazy_b = 0;
    if (lazy) {
        pb->lazy_pending = true;   // the first fused_tma iteration takes b as the iterate and resets the per-step flags
        pb->pd.lazy_b = 1;
    } else {
        int rc = run_begin_step_kernel(pb);
        if (rc) return rc;
    }
    pb->slot_base = pb->slot_of(pb->launched);
    pb->launched = 0; pb->confirmed = 0; pb->last_active = pb->pd.B;
    pb->step_all_tma = true;
    return SB_OK;
}

constexpr unsigned int kPending = 0xFFFFFFFFu;  // host_count value of a launch that has not reported yet

// spin until the launch that uses ring slot s has published its active count (mapped host memory)
static int wait_slot(sb_problem* pb, int s, unsigned int* value) {
    volatile unsigned int* p = pb->h_count + s;
    unsigned long spins = 0;
    unsigned int v;
    while ((v = *p) == kPending) {
        if ((++spins & 0x3FFF) == 0) {  // every so often make sure the stream is still alive
            cudaError_t e = cudaStreamQuery(pb->ctx->stream);
            if (e == cudaSuccess) {
                if ((v = *p) != kPending) break;
                return fail(SB_ERR_CUDA, "Newton-iteration launch finished without reporting its active count");
            }
            if (e != cudaErrorNotReady) return fail(SB_ERR_CUDA, "stream error while waiting: %s", cudaGetErrorString(e));
        }
    }
    *value = v;
    if (v == 0 && pb->step_all_tma && !pb->longmesh) pb->b_ready = true;  // every instance wrote its next b when it converged
    return SB_OK;
}

static int prepare_slot(sb_problem* pb, int* slot) {
    const int s = pb->slot_of(pb->launched);
    pb->pd.iter_no = (int)pb->launched;  // Newton iteration number inside the step (lock-step: same for all active instances)
    if (pb->total_launched >= kSlots) {  // the slot's previous user (kSlots launches ago) must have delivered its count
        if (pb->prof && pb->launched >= kSlots) { int rc = prof_harvest(pb); if (rc) return rc; }
        volatile unsigned int* p = pb->h_count + s;
        unsigned long spins = 0;
        while (*p == kPending) {
            if ((++spins & 0x3FFF) == 0) {
                cudaError_t e = cudaStreamQuery(pb->ctx->stream);
                if (e != cudaSuccess && e != cudaErrorNotReady) return fail(SB_ERR_CUDA, "stream error while waiting: %s", cudaGetErrorString(e));
                if (e == cudaSuccess && *p == kPending) return fail(SB_ERR_CUDA, "Newton-iteration launch finished without reporting its active count");
            }
        }
        if (pb->launched >= kSlots && pb->confirmed < pb->launched - kSlots + 1) pb->confirmed = pb->launched - kSlots + 1;
    }
    *((volatile unsigned int*)pb->h_count + s) = kPending;
    *slot = s;
    return SB_OK;
}

static int finish_slot(sb_problem* pb, int slot) {
    (void)slot;
    pb->launched++;
    pb->total_launched++;
    return SB_OK;
}

int sb_residual_jacobian(sb_problem* pb) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    CU(cudaSetDevice(pb->ctx->device));
    int rc = ensure_jacobian(pb);
    if (rc) return rc;
    if ((rc = materialize_begin(pb))) return rc;
    dim3 grid, block; size_t smem;
    launch_geometry(pb, grid, block, smem);
    const int ps = pb->slot_of(pb->launched);
    if (pb->prof) CU(cudaEventRecord(pb->evA[ps], pb->ctx->stream));
    if (pb->longmesh)
        CU(launch_any(pb->ks.long_assemble_jac, dim3(pb->pd.ntiles, (unsigned)pb->pd.B), dim3(kLongT0), 0, pb->ctx->stream, pb->pd, pb->t_next, pb->inv_tau, 0));
    else
        CU(launch_any(pb->ks.assemble_jac, grid, block, 0, pb->ctx->stream, pb->pd, pb->t_next, pb->inv_tau, 0));
    CU(cudaGetLastError());
    if (pb->prof) CU(cudaEventRecord(pb->evM[ps], pb->ctx->stream));
    return SB_OK;
}

int sb_solve_update_norm(sb_problem* pb, double tol) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    if (!pb->d_J) return fail(SB_ERR_INVALID, "sb_solve_update_norm called before sb_residual_jacobian");
    if (pb->longmesh)
        return fail(SB_ERR_UNSUPPORTED, "the long-mesh path solves from on-the-fly assembly only: use sb_newton_iteration");
    CU(cudaSetDevice(pb->ctx->device));
    int slot;
    int rc = materialize_begin(pb);
    if (rc) return rc;
    pb->step_all_tma = false;   // k_solve neither writes the next step's b nor the compacted list
    if ((rc = prepare_slot(pb, &slot))) return rc;
    dim3 grid, block; size_t smem;
    launch_geometry(pb, grid, block, smem);
    if ((rc = configure_small_kernels(pb))) return rc;
    CU(launch_any(pb->ks.solve, grid, block, smem, pb->ctx->stream, pb->pd, tol, slot, 1));
    CU(cudaGetLastError());
    if (pb->prof) CU(cudaEventRecord(pb->evB[slot], pb->ctx->stream));
    return finish_slot(pb, slot);
}

int sb_newton_iteration(sb_problem* pb, double tol) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    if (pb->longmesh) {
        CU(cudaSetDevice(pb->ctx->device));
        int slot;
        int rc = prepare_slot(pb, &slot);
        if (rc) return rc;
        pb->step_all_tma = false;
        if (pb->prof) CU(cudaEventRecord(pb->evA[slot], pb->ctx->stream));
        rc = longmesh_iteration(pb, tol, slot);
        if (rc) return rc;
        if (pb->prof) { CU(cudaEventRecord(pb->evM[slot], pb->ctx->stream)); CU(cudaEventRecord(pb->evB[slot], pb->ctx->stream)); }
        return finish_slot(pb, slot);
    }
    if (pb->design == SB_DESIGN_SPLIT) {
        int rc = sb_residual_jacobian(pb);
        if (rc) return rc;
        return sb_solve_update_norm(pb, tol);
    }
    CU(cudaSetDevice(pb->ctx->device));
    if (is_tma(pb->design) && pb->launched > 0 && !pb->step_all_tma)
        return fail(SB_ERR_INVALID, "a fused_tma iteration cannot follow a split / fused / two-call iteration inside the same "
                    "step: only fused_tma launches maintain the compacted list of active instances it reads");
    int slot;
    int rc = SB_OK;
    if (!is_tma(pb->design) && (rc = materialize_begin(pb))) return rc;   // k_newton_fused reads u and active[]
    if ((rc = prepare_slot(pb, &slot))) return rc;
    dim3 grid, block; size_t smem;
    launch_geometry(pb, grid, block, smem);
    if (is_tma(pb->design)) {
        const void* kern = pb->ks.fused_p;
        if (pb->p_grid == 0) {  // one-time configuration: opt in to the shared memory, size the persistent grid
            pb->p_smem = pb->ks.fused_p_smem;
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb->p_smem));
            int occ = 0, sms = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (int)block.x, pb->p_smem));
            CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pb->ctx->device));
            if (occ < 1) return fail(SB_ERR_UNSUPPORTED, "fused_tma kernel does not fit on an SM (%zu bytes of shared memory)", pb->p_smem);
            if (const char* e = getenv("SB_PERSISTENT_CTAS_PER_SM")) { int v = atoi(e); if (v >= 1 && v < occ) occ = v; }
            pb->p_grid = (int)std::min<int64_t>((int64_t)occ * sms, (int64_t)grid.x);
        }
        const int prev_slot = (pb->launched == 0) ? -1 : pb->slot_of(pb->launched - 1);
        const int parity = (int)(pb->launched & 1);
        if (pb->prof) CU(cudaEventRecord(pb->evA[slot], pb->ctx->stream));
        CU(launch_pdl(kern, dim3(pb->p_grid), block, pb->p_smem, pb->ctx->stream, pb->pd, pb->t_next, pb->inv_tau, tol, slot, prev_slot, parity));
        CU(cudaGetLastError());
        pb->lazy_pending = false;   // pd.lazy_b stays set for this step: only iter_no == 0 reads b as the iterate
        if (pb->prof) CU(cudaEventRecord(pb->evB[slot], pb->ctx->stream));
        return finish_slot(pb, slot);
    }
    pb->step_all_tma = false;
    if ((rc = configure_small_kernels(pb))) return rc;
    if (pb->prof) CU(cudaEventRecord(pb->evA[slot], pb->ctx->stream));
    CU(launch_any(pb->ks.fused, grid, block, smem, pb->ctx->stream, pb->pd, pb->t_next, pb->inv_tau, tol, slot));
    CU(cudaGetLastError());
    if (pb->prof) CU(cudaEventRecord(pb->evB[slot], pb->ctx->stream));
    return finish_slot(pb, slot);
}

int sb_poll_active(sb_problem* pb, int blocking, int64_t* n_active) {
    if (!pb || !n_active) return fail(SB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(pb->ctx->device));
    if (pb->launched == 0) { *n_active = pb->last_active; return SB_OK; }
    if (blocking) {
        unsigned int v;
        int rc = wait_slot(pb, pb->slot_of(pb->launched - 1), &v);
        if (rc) return rc;
        pb->confirmed = pb->launched;
        pb->last_active = v;
    } else {
        while (pb->confirmed < pb->launched) {
            const unsigned int v = *((volatile unsigned int*)pb->h_count + pb->slot_of(pb->confirmed));
            if (v == kPending) break;
            if (v == 0 && pb->step_all_tma && !pb->longmesh) pb->b_ready = true;
            pb->last_active = v;
            pb->confirmed++;
        }
    }
    *n_active = pb->last_active;
    return SB_OK;
}

int sb_wait_iteration(sb_problem* pb, int it, int64_t* n_active) {
    if (!pb || !n_active) return fail(SB_ERR_INVALID, "NULL argument");
    if (it < 0 || it >= pb->launched) return fail(SB_ERR_INVALID, "iteration %d has not been launched in this step", it);
    if (it < pb->launched - kSlots) return fail(SB_ERR_INVALID, "iteration %d is older than the %d-slot counter ring", it, kSlots);
    CU(cudaSetDevice(pb->ctx->device));
    unsigned int v;
    int rc = wait_slot(pb, pb->slot_of(it), &v);
    if (rc) return rc;
    *n_active = v;
    if (pb->confirmed < it + 1) { pb->confirmed = it + 1; pb->last_active = v; }
    return SB_OK;
}

}  // extern "C"

// ---- whole solve in one launch (sb_resident.cuh) ---------------------------------------------------------------------
constexpr int kMaxRounds = 4096;   // rounds of the persistent grid over the batch that snapshot streaming can track

static bool resident_enabled(const sb_problem* pb) {
    static const bool off = getenv("SB_NO_RESIDENT") != nullptr;
    return !off && !pb->longmesh && pb->design_pending == SB_DESIGN_RESIDENT && pb->ks.resident != nullptr;
}

template <class T>
static int grow(T** p, size_t* cap, size_t need, cudaStream_t st) {
    if (*cap >= need) return SB_OK;
    if (*p) { CU(cudaStreamSynchronize(st)); guarded_free(*p); *p = nullptr; *cap = 0; }
    CU(guarded_malloc((void**)p, need * sizeof(T)));
    *cap = need;
    return SB_OK;
}

// nsteps implicit-Euler steps (tgrid != NULL) or one Newton solve at (t_single, inv_tau_single) on the b that
// sb_begin_step prepared (tgrid == NULL), every instance resident on its SM from the first to the last iteration.
// snap_host: every state [nsteps+1][B][row], or -- final_only -- u(t_end) alone [B][row]; either way finished rounds of
// instances leave for the host while later rounds compute
static int solve_resident(sb_problem* pb, int nsteps, const double* tgrid, const double* inst_tgrid, double tau, double t_single,
                          double inv_tau_single, double tol, int maxit, int32_t* step_iters_host, double* snap_host,
                          int* steps_done, int* max_iters_out, bool final_only = false, const double* u0_host = nullptr) {
    cudaStream_t st = pb->ctx->stream;
    const ProbDev& pd = pb->pd;
    const int64_t B = pd.B;
    const bool b_given = (tgrid == nullptr && inst_tgrid == nullptr);
    const int G = resident_groups(pd.T);
    const unsigned threads = (unsigned)(G * pd.T);
    if (pb->r_grid == 0) {
        CU(cudaFuncSetAttribute(pb->ks.resident, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb->ks.resident_smem));
        int occ = 0, sms = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pb->ks.resident, (int)threads, pb->ks.resident_smem));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pb->ctx->device));
        if (occ < 1) return fail(SB_ERR_UNSUPPORTED, "resident kernel does not fit on an SM (%zu bytes of shared memory)", pb->ks.resident_smem);
        if (const char* e = getenv("SB_PERSISTENT_CTAS_PER_SM")) { int v = atoi(e); if (v >= 1 && v < occ) occ = v; }
        pb->r_grid = (int)std::min<int64_t>((int64_t)occ * sms, (B + G - 1) / G);
    }
    const int64_t stride = (int64_t)pb->r_grid * G;
    const int rounds = (int)((B + stride - 1) / stride);
    if (!pb->d_rmisc) {
        CU(guarded_malloc((void**)&pb->d_rmisc, (4 + kMaxRounds) * sizeof(unsigned int)));
        CU(cudaHostAlloc((void**)&pb->h_rounds, kMaxRounds * sizeof(unsigned int), cudaHostAllocMapped));
    }
    ResidentArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.nsteps = nsteps; ra.maxit = maxit; ra.b_given = b_given ? 1 : 0; ra.tol = tol; ra.tau = tau;
    ra.t_single = t_single; ra.inv_tau_single = inv_tau_single;
    if (tgrid) {
        if (pb->tgrid_cap < nsteps + 1) {
            if (pb->d_tgrid) { CU(cudaStreamSynchronize(st)); CU(guarded_free(pb->d_tgrid)); pb->d_tgrid = nullptr; }
            CU(guarded_malloc((void**)&pb->d_tgrid, (size_t)(nsteps + 1) * sizeof(double)));
            pb->tgrid_cap = nsteps + 1;
        }
        CU(cudaMemcpyAsync(pb->d_tgrid, tgrid, (size_t)(nsteps + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
        ra.tgrid = pb->d_tgrid;
    }
    if (inst_tgrid) {
        const size_t n = (size_t)B * (nsteps + 1);
        int rc = grow(&pb->d_itgrid, &pb->itgrid_cap, n, st);
        if (rc) return rc;
        CU(cudaMemcpyAsync(pb->d_itgrid, inst_tgrid, n * sizeof(double), cudaMemcpyHostToDevice, st));
        ra.inst_tgrid = pb->d_itgrid;
    }
    if (step_iters_host) {
        int rc = grow(&pb->d_step_iters, &pb->step_iters_cap, (size_t)B * nsteps, st);
        if (rc) return rc;
        ra.step_iters = pb->d_step_iters;
    }
    const size_t row = (size_t)pd.Nx * pb->P;              // doubles per instance and snapshot
    const bool stream_snaps = snap_host != nullptr && rounds <= kMaxRounds;
    if (snap_host) {
        int rc = grow(&pb->d_snap, &pb->snap_cap, (size_t)(final_only ? 1 : nsteps + 1) * B * row, st);
        if (rc) return rc;
        ra.snap = pb->d_snap;
        ra.snap_last = final_only ? 1 : 0;
        if (!pb->copy_stream) CU(cudaStreamCreateWithFlags(&pb->copy_stream, cudaStreamNonBlocking));
    }
    unsigned int init[4] = {0x7fffffffu, 0u, 0u, 0u};
    CU(cudaMemcpyAsync(pb->d_rmisc, init, sizeof init, cudaMemcpyHostToDevice, st));
    ra.fail_step = reinterpret_cast<int*>(pb->d_rmisc);
    ra.max_iters = reinterpret_cast<int*>(pb->d_rmisc + 1);
    ra.sum_iters = reinterpret_cast<unsigned long long*>(pb->d_rmisc + 2);
    if (stream_snaps) {
        CU(cudaMemsetAsync(pb->d_rmisc + 4, 0, (size_t)rounds * sizeof(unsigned int), st));
        for (int r = 0; r < rounds; ++r) pb->h_rounds[r] = 0u;
        unsigned int* d_h = nullptr;
        CU(cudaHostGetDevicePointer((void**)&d_h, pb->h_rounds, 0));
        ra.round_done = pb->d_rmisc + 4;
        ra.host_rounds = d_h;
    }
    constexpr int kMaxChunks = 4096;
    ResidentPipe rp;
    memset(&rp, 0, sizeof rp);
    if (u0_host) {
        // The initial states travel chunk by chunk on the copy stream, each chunk followed by its arrival word, and ALL of
        // it is enqueued BEFORE the launch: once the kernel runs (it may hold every SM while it waits for a chunk) nothing
        // it waits for depends on this or any other host thread making progress -- copy engines need neither an SM nor a
        // driver lock that a synchronising call of another thread (cudaFree, cudaMalloc ...) could be holding.
        int64_t chunk = std::max<int64_t>(1, (int64_t)((2u << 20) / (row * sizeof(double))));          // about 2 MB per chunk: every chunk is two enqueue calls before the launch
        if ((B + chunk - 1) / chunk > kMaxChunks) chunk = (B + kMaxChunks - 1) / kMaxChunks;
        const int nchunks = (int)((B + chunk - 1) / chunk);
        if (!pb->d_ready) {
            CU(guarded_malloc((void**)&pb->d_ready, 4 * sizeof(unsigned int)));
            CU(cudaHostAlloc((void**)&pb->h_ready_vals, kMaxChunks * sizeof(unsigned int), cudaHostAllocDefault));
            CU(cudaMemsetAsync(pb->d_ready, 0, 4 * sizeof(unsigned int), st));
            CU(cudaStreamSynchronize(st));
        }
        if (!pb->copy_stream) CU(cudaStreamCreateWithFlags(&pb->copy_stream, cudaStreamNonBlocking));
        int rc = ensure_tmp(pb, (size_t)B * row);
        if (rc) return rc;
        rp.u_ref = pb->d_tmp; rp.ready = pb->d_ready; rp.ready_base = pb->ready_epoch; rp.chunk = (int)chunk;
        for (int c = 0; c < nchunks; ++c) {
            const int64_t k0 = (int64_t)c * chunk, nk = std::min<int64_t>(chunk, B - k0);
            pb->h_ready_vals[c] = pb->ready_epoch + (unsigned int)c + 1u;
            CU(cudaMemcpyAsync(pb->d_tmp + (size_t)k0 * row, u0_host + (size_t)k0 * row, (size_t)nk * row * sizeof(double), cudaMemcpyHostToDevice, pb->copy_stream));
            CU(cudaMemcpyAsync(pb->d_ready, pb->h_ready_vals + c, sizeof(unsigned int), cudaMemcpyHostToDevice, pb->copy_stream));
        }
        pb->ready_epoch += (unsigned int)nchunks + 1u;
    }
    if (pb->prof) CU(cudaEventRecord(pb->evA[0], st));
    const void* kern = pb->ks.resident;
    if (u0_host) {
        kern = pb->ks.resident_pipe;
        if (!pb->ks.pipe_configured) {
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb->ks.resident_smem));
            pb->ks.pipe_configured = true;
        }
    }
    if (u0_host) CU(launch_any(kern, dim3((unsigned)pb->r_grid), dim3(threads), pb->ks.resident_smem, st, pb->pd, ra, rp));
    else CU(launch_any(kern, dim3((unsigned)pb->r_grid), dim3(threads), pb->ks.resident_smem, st, pb->pd, ra));
    CU(cudaGetLastError());
    if (pb->prof) CU(cudaEventRecord(pb->evB[0], st));
    // from here on the state is what the kernel leaves: u = u(t_end), b unspecified
    pb->b_ready = false; pb->lazy_pending = false; pb->pd.lazy_b = 0; pb->step_all_tma = false;
    // the per-iteration launches of an earlier step keep their ring slots (a slot is cleaned only kSlots/2 launches
    // later): the ring position moves past them before the counter of the step is cleared, exactly as sb_begin_step does
    // -- a pooled problem that went host loop -> whole-solve kernel -> host loop reused uncleaned slots otherwise
    pb->slot_base = pb->slot_of(pb->launched);
    pb->launched = 0; pb->confirmed = 0;

    if (snap_host) {
        const int s0 = (b_given && !final_only) ? 1 : 0, s1 = final_only ? 0 : nsteps;
        if (stream_snaps) {
            // finished rounds leave for the host while later rounds compute (src/pdepe.jl:99-107 stores every step)
            for (int r = 0; r < rounds; ++r) {
                volatile unsigned int* flag = pb->h_rounds + r;
                unsigned long spins = 0;
                while (*flag == 0u) {
                    if ((++spins & 0x3FFF) == 0) {
                        cudaError_t e = cudaStreamQuery(st);
                        if (e == cudaSuccess) { if (*flag == 0u) return fail(SB_ERR_CUDA, "resident kernel ended without finishing round %d", r); break; }
                        if (e != cudaErrorNotReady) return fail(SB_ERR_CUDA, "stream error while waiting: %s", cudaGetErrorString(e));
                    }
                }
                const int64_t k0 = (int64_t)r * stride, nk = std::min<int64_t>(stride, B - k0);
                for (int s = s0; s <= s1; ++s) {
                    const size_t off = ((size_t)s * B + k0) * row;
                    CU(cudaMemcpyAsync(snap_host + off, pb->d_snap + off, (size_t)nk * row * sizeof(double), cudaMemcpyDeviceToHost, pb->copy_stream));
                }
            }
            CU(cudaStreamSynchronize(pb->copy_stream));
        } else {
            CU(cudaStreamSynchronize(st));
            const size_t off = (size_t)s0 * B * row;
            CU(cudaMemcpy(snap_host + off, pb->d_snap + off, ((size_t)(s1 + 1 - s0) * B * row) * sizeof(double), cudaMemcpyDeviceToHost));
        }
    }
    if (u0_host) CU(cudaStreamSynchronize(pb->copy_stream));   // (the pinned words and the caller's buffer are free again)
    unsigned int res[4];
    CU(cudaMemcpyAsync(res, pb->d_rmisc, sizeof res, cudaMemcpyDeviceToHost, st));
    if (step_iters_host)
        CU(cudaMemcpyAsync(step_iters_host, pb->d_step_iters, (size_t)B * nsteps * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int fail_step = (int)res[0];
    unsigned long long sum_iters;
    memcpy(&sum_iters, res + 2, sizeof sum_iters);
    if (pb->prof) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, pb->evA[0], pb->evB[0]));
        pb->prof_ms_solve += ms; pb->prof_launches += 1; pb->prof_inst_iters += (int64_t)sum_iters;
    }
    if (max_iters_out) *max_iters_out = (int)res[1];
    const bool failed = fail_step != 0x7fffffff;
    pb->last_active = failed ? 1 : 0;
    pb->predicted_iters = 0;
    if (steps_done) *steps_done = failed ? fail_step : nsteps;
    if (failed) return fail(SB_ERR_CONVERGENCE, "convergence failed");   // src/utils.jl:336
    return SB_OK;
}

extern "C" {

int sb_newton_solve(sb_problem* pb, double tol, int maxit, int* iters_done) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    if (maxit < 1) return fail(SB_ERR_INVALID, "maxit must be >= 1");
    CU(cudaSetDevice(pb->ctx->device));
    if (resident_enabled(pb) && pb->design == SB_DESIGN_RESIDENT && pb->launched == 0) {
        // the whole Newton loop of this step in one launch, on the b that sb_begin_step prepared
        int mx = 0;
        int rc = solve_resident(pb, 1, nullptr, nullptr, -1.0, pb->t_next, pb->inv_tau, tol, maxit, nullptr, nullptr, nullptr, &mx);
        if (iters_done) *iters_done = mx;
        return rc;
    }
    const long depth = 2;  // iterations kept in flight beyond the last one whose count has been read
    const long base = pb->launched;
    // Speculation limit: launches are enqueued ahead of the host's knowledge only up to the number of iterations
    // the previous solve needed (consecutive time steps need the same or one fewer); past that, the next iteration
    // is launched only once the previous one has reported active instances.  A step that behaves like its
    // predecessor then ends without a surplus no-op launch; one that needs more pays a host round trip.
    const long limit = (pb->predicted_iters > 0 && !getenv("SB_NO_PREDICT")) ? pb->predicted_iters : (long)maxit;
    long done = 0;          // iterations whose active count has been read
    int64_t n_act = pb->pd.B;
    while (true) {
        while (pb->launched - base < maxit && (pb->launched - base) - done < depth &&
               ((pb->launched - base) < limit || (pb->launched - base) == done)) {
            int rc = sb_newton_iteration(pb, tol);
            if (rc) return rc;
        }
        unsigned int v;
        int rc = wait_slot(pb, pb->slot_of(base + done), &v);
        if (rc) return rc;
        n_act = v;
        done++;
        pb->confirmed = base + done;
        pb->last_active = n_act;
        if (n_act == 0) break;
        if (done >= maxit) break;
    }
    pb->predicted_iters = done;
    if (iters_done) *iters_done = (int)(pb->launched - base);
    if (n_act != 0) return fail(SB_ERR_CONVERGENCE, "convergence failed");  // src/utils.jl:336
    return SB_OK;
}

// The time loop with device-side stepping (fused_tma): the host launches Newton iterations exactly as in
// sb_newton_solve -- two in flight beyond the one it waits for -- but a launch that finds every instance converged
// starts the next time step itself (ProbDev::auto_*), so the look-ahead launches at the end of a step are the first
// iterations of the next one instead of no-ops and the GPU never waits for the host's "step finished" round trip.
// The host follows through the mapped (count, step, iteration) triple of every launch and enforces maxit.
static int solve_steps_auto(sb_problem* pb, int nsteps, const double* tgrid, double tau, double tol, int maxit, int* steps_done) {
    cudaStream_t st = pb->ctx->stream;
    if (pb->tgrid_cap < nsteps + 1) {
        if (pb->d_tgrid) CU(guarded_free(pb->d_tgrid));
        CU(guarded_malloc((void**)&pb->d_tgrid, (size_t)(nsteps + 1) * sizeof(double)));
        pb->tgrid_cap = nsteps + 1;
    }
    if (!pb->d_auto) CU(guarded_malloc((void**)&pb->d_auto, 2 * sizeof(int)));
    CU(cudaMemcpyAsync(pb->d_tgrid, tgrid, (size_t)(nsteps + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(pb->d_auto, 0, 2 * sizeof(int), st));
    const double tau0 = (tau > 0) ? tau : tgrid[1] - tgrid[0];
    int rc = sb_begin_step(pb, tgrid[1], std::isinf(tau0) ? 0.0 : 1.0 / tau0);   // b = M.*u_0 (or already there)
    if (rc) return rc;
    pb->pd.auto_on = 1; pb->pd.auto_nsteps = nsteps; pb->pd.auto_state = pb->d_auto;
    pb->pd.auto_tgrid = pb->d_tgrid; pb->pd.auto_tau = (tau > 0) ? tau : -1.0;
    const long depth = 2;
    long read = 0;
    int fail_step = -1, last_step = 0;
    bool finished = false;
    while (!finished) {
        while (pb->launched - read <= depth) {
            rc = sb_newton_iteration(pb, tol);
            if (rc) { pb->pd.auto_on = 0; return rc; }
        }
        unsigned int count;
        rc = wait_slot(pb, pb->slot_of(read), &count);
        if (rc) { pb->pd.auto_on = 0; return rc; }
        const int s = pb->slot_of(read);
        const int step = (int)*((volatile unsigned int*)pb->h_count + kSlots + s);
        const int iter = (int)*((volatile unsigned int*)pb->h_count + 2 * kSlots + s);
        ++read;
        last_step = step;
        if (step >= nsteps || (count == 0 && step == nsteps - 1)) finished = true;
        else if (count != 0 && iter + 1 >= maxit) { fail_step = step; finished = true; }   // src/utils.jl:336
    }
    // launches still in flight find nothing to do (or, after a failure, iterate on); what follows in the stream waits for them
    pb->pd.auto_on = 0;
    pb->confirmed = pb->launched;
    pb->last_active = (fail_step >= 0) ? 1 : 0;
    pb->b_ready = (fail_step < 0);
    pb->predicted_iters = 0;
    if (steps_done) *steps_done = (fail_step >= 0) ? fail_step : std::min(last_step + 1, nsteps);
    if (fail_step >= 0) return fail(SB_ERR_CONVERGENCE, "convergence failed");
    return SB_OK;
}

int sb_solve_steps(sb_problem* pb, int nsteps, const double* tgrid, double tau, double tol, int maxit, int* steps_done) {
    if (!pb || !tgrid || nsteps < 1) return fail(SB_ERR_INVALID, "NULL argument or nsteps < 1");
    if (steps_done) *steps_done = 0;
    if (resident_enabled(pb) && (!pb->pd.hist || nsteps == 1) && maxit >= 1) {
        CU(cudaSetDevice(pb->ctx->device));
        pb->design = pb->design_pending;
        return solve_resident(pb, nsteps, tgrid, nullptr, tau, 0.0, 0.0, tol, maxit, nullptr, nullptr, steps_done, nullptr);
    }
    if (is_tma(pb->design_pending) && !pb->longmesh && !pb->prof && !pb->pd.hist && nsteps > 1 && maxit >= 2 &&
        !getenv("SB_NO_DEVICE_STEPPING")) {
        CU(cudaSetDevice(pb->ctx->device));
        return solve_steps_auto(pb, nsteps, tgrid, tau, tol, maxit, steps_done);
    }
    for (int i = 0; i < nsteps; ++i) {                                    // src/pdepe.jl:90
        const double tau_i = (tau > 0) ? tau : tgrid[i + 1] - tgrid[i];  // :92 (scalar tstep | diff of the vector grid)
        int rc = sb_begin_step(pb, tgrid[i + 1], std::isinf(tau_i) ? 0.0 : 1.0 / tau_i);
        if (rc) return rc;
        rc = sb_newton_solve(pb, tol, maxit, nullptr);
        if (rc) return rc;
        if (steps_done) *steps_done = i + 1;
    }
    return SB_OK;
}

int sb_solve_steps_opt(sb_problem* pb, int nsteps, const double* tgrid, const double* tgrid_per_instance, double tau, double tol,
                       int maxit, int32_t* step_iters, double* snapshots, int* steps_done) {
    if (!pb || (!tgrid && !tgrid_per_instance) || nsteps < 1) return fail(SB_ERR_INVALID, "NULL argument or nsteps < 1");
    if (steps_done) *steps_done = 0;
    if (maxit < 1) return fail(SB_ERR_INVALID, "maxit must be >= 1");
    if (pb->longmesh) return fail(SB_ERR_UNSUPPORTED, "sb_solve_steps_opt: the long-mesh path has no whole-solve kernel; use sb_solve_steps");
    if (!resident_enabled(pb)) return fail(SB_ERR_UNSUPPORTED, "sb_solve_steps_opt needs the resident design (whole-solve kernel)");
    if (pb->pd.hist && nsteps > 1) return fail(SB_ERR_UNSUPPORTED, "norm histories are kept per step: drive the steps one by one with hist enabled");
    CU(cudaSetDevice(pb->ctx->device));
    pb->design = pb->design_pending;
    return solve_resident(pb, nsteps, tgrid_per_instance ? nullptr : tgrid, tgrid_per_instance, tgrid_per_instance ? -1.0 : tau, 0.0, 0.0,
                          tol, maxit, step_iters, snapshots, steps_done, nullptr);
}

int sb_solve_steps_final(sb_problem* pb, int nsteps, const double* tgrid, double tau, double tol, int maxit, double* final_state,
                         int* steps_done) {
    if (!pb || !tgrid || !final_state || nsteps < 1) return fail(SB_ERR_INVALID, "NULL argument or nsteps < 1");
    if (steps_done) *steps_done = 0;
    if (maxit < 1) return fail(SB_ERR_INVALID, "maxit must be >= 1");
    if (pb->longmesh) return fail(SB_ERR_UNSUPPORTED, "sb_solve_steps_final: the long-mesh path has no whole-solve kernel; use sb_solve_steps and sb_state_download");
    if (!resident_enabled(pb)) return fail(SB_ERR_UNSUPPORTED, "sb_solve_steps_final needs the resident design (whole-solve kernel)");
    if (pb->pd.hist && nsteps > 1) return fail(SB_ERR_UNSUPPORTED, "norm histories are kept per step: drive the steps one by one with hist enabled");
    CU(cudaSetDevice(pb->ctx->device));
    pb->design = pb->design_pending;
    return solve_resident(pb, nsteps, tgrid, nullptr, tau, 0.0, 0.0, tol, maxit, nullptr, final_state, steps_done, nullptr, true);
}

int sb_solve_from_host(sb_problem* pb, const double* u0_host, double t0, int nsteps, const double* tgrid, double tau, double tol,
                       int maxit, double* final_state, int* steps_done) {
    if (!pb || !u0_host || !tgrid || !final_state || nsteps < 1) return fail(SB_ERR_INVALID, "NULL argument or nsteps < 1");
    if (steps_done) *steps_done = 0;
    if (maxit < 1) return fail(SB_ERR_INVALID, "maxit must be >= 1");
    if (pb->longmesh) return fail(SB_ERR_UNSUPPORTED, "sb_solve_from_host: the long-mesh path has no whole-solve kernel");
    if (!resident_enabled(pb)) return fail(SB_ERR_UNSUPPORTED, "sb_solve_from_host needs the resident design (whole-solve kernel)");
    if (pb->pd.hist && nsteps > 1) return fail(SB_ERR_UNSUPPORTED, "norm histories are kept per step: drive the steps one by one with hist enabled");
    CU(cudaSetDevice(pb->ctx->device));
    pb->design = pb->design_pending;
    cudaStream_t st = pb->ctx->stream;
    const int P = pb->P, Nx = pb->pd.Nx;
    const int64_t B = pb->pd.B;
    // what sb_state_upload resets
    pb->b_ready = false; pb->lazy_pending = false; pb->pd.lazy_b = 0;
    CU(cudaMemsetAsync(pb->d_total, 0, B * sizeof(long long), st));
    CU(cudaMemsetAsync(pb->d_status, 0, B * sizeof(int), st));
    // mass flags (src/pdepe.jl:78-79) need nodes 0, 1 and Nx-1 of every instance only: those go ahead of the states
    const size_t nb = (size_t)B * 3 * P;
    if (!pb->h_bnd) {
        CU(cudaHostAlloc((void**)&pb->h_bnd, nb * sizeof(double), cudaHostAllocDefault));
        CU(guarded_malloc((void**)&pb->d_bnd, nb * sizeof(double)));
    }
    const size_t row = (size_t)Nx * P;
    for (int64_t k = 0; k < B; ++k)
        for (int p = 0; p < P; ++p) {
            pb->h_bnd[((size_t)k * 3 + 0) * P + p] = u0_host[(size_t)k * row + p];
            pb->h_bnd[((size_t)k * 3 + 1) * P + p] = u0_host[(size_t)k * row + P + p];
            pb->h_bnd[((size_t)k * 3 + 2) * P + p] = u0_host[(size_t)k * row + (size_t)(Nx - 1) * P + p];
        }
    CU(cudaMemcpyAsync(pb->d_bnd, pb->h_bnd, nb * sizeof(double), cudaMemcpyHostToDevice, st));
    const int bs = 128;
    CU(launch_any(pb->ks.mass, dim3((unsigned)((B + bs - 1) / bs)), dim3(bs), 0, st, pb->pd, t0, pb->d_mass, 0, (const double*)pb->d_bnd));
    CU(cudaGetLastError());
    pb->mass_ready = true;
    return solve_resident(pb, nsteps, tgrid, nullptr, tau, 0.0, 0.0, tol, maxit, nullptr, final_state, steps_done, nullptr, true, u0_host);
}

#define SB_GET(NAME, TYPE, DEVPTR)                                                                       \
    int NAME(sb_problem* pb, TYPE* host) {                                                               \
        if (!pb || !host) return fail(SB_ERR_INVALID, "NULL argument");                                  \
        CU(cudaSetDevice(pb->ctx->device));                                                              \
        CU(cudaStreamSynchronize(pb->ctx->stream));                                                      \
        CU(cudaMemcpy(host, pb->DEVPTR, (size_t)pb->pd.B * sizeof(TYPE), cudaMemcpyDeviceToHost));       \
        return SB_OK;                                                                                    \
    }
SB_GET(sb_get_step_iters, int32_t, d_iters)
SB_GET(sb_get_last_norm, double, d_norm)
SB_GET(sb_get_status, int32_t, d_status)
#undef SB_GET

int sb_get_total_iters(sb_problem* pb, int64_t* host) {
    if (!pb || !host) return fail(SB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(pb->ctx->device));
    CU(cudaStreamSynchronize(pb->ctx->stream));
    CU(cudaMemcpy(host, pb->d_total, (size_t)pb->pd.B * sizeof(long long), cudaMemcpyDeviceToHost));
    return SB_OK;
}

int sb_total_active_iterations(sb_problem* pb, int64_t* total) {
    if (!pb || !total) return fail(SB_ERR_INVALID, "NULL argument");
    std::vector<int64_t> h((size_t)pb->pd.B);
    int rc = sb_get_total_iters(pb, h.data());
    if (rc) return rc;
    int64_t s = 0;
    for (int64_t v : h) s += v;
    *total = s;
    return SB_OK;
}

int sb_state_checkpoint(sb_problem* pb) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    CU(cudaSetDevice(pb->ctx->device));
    const size_t bytes = node_elems(pb) * sizeof(double);
    { int rc = materialize_begin(pb); if (rc) return rc; }
    if (!pb->d_ckpt) CU(guarded_malloc((void**)&pb->d_ckpt, bytes));
    CU(cudaMemcpyAsync(pb->d_ckpt, pb->d_u, bytes, cudaMemcpyDeviceToDevice, pb->ctx->stream));
    return SB_OK;
}

int sb_state_rollback(sb_problem* pb) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    if (!pb->d_ckpt) return fail(SB_ERR_INVALID, "no checkpoint taken");
    CU(cudaSetDevice(pb->ctx->device));
    cudaStream_t st = pb->ctx->stream;
    pb->b_ready = false; pb->lazy_pending = false; pb->pd.lazy_b = 0;
    CU(cudaMemcpyAsync(pb->d_u, pb->d_ckpt, node_elems(pb) * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CU(cudaMemsetAsync(pb->d_total, 0, pb->pd.B * sizeof(long long), st));
    CU(cudaMemsetAsync(pb->d_status, 0, pb->pd.B * sizeof(int), st));
    return SB_OK;
}

int sb_profile_begin(sb_problem* pb) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    CU(cudaSetDevice(pb->ctx->device));
    if (!pb->prof_ev_ok) {
        for (int i = 0; i < kSlots; ++i) {
            CU(cudaEventCreate(&pb->evA[i])); CU(cudaEventCreate(&pb->evM[i])); CU(cudaEventCreate(&pb->evB[i]));
        }
        pb->prof_ev_ok = true;
    }
    CU(cudaStreamSynchronize(pb->ctx->stream));
    pb->prof = true; pb->prof_harvested = pb->launched; pb->prof_prev_active = pb->last_active;
    pb->prof_ms_assemble = pb->prof_ms_solve = 0.0;
    pb->prof_launches = pb->prof_noop = pb->prof_inst_iters = 0;
    return SB_OK;
}

int sb_profile_end(sb_problem* pb, double* ms_assemble, double* ms_solve, int64_t* launches, int64_t* noop_launches,
                   int64_t* instance_iterations) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    if (!pb->prof) return fail(SB_ERR_INVALID, "profiling is not active");
    CU(cudaSetDevice(pb->ctx->device));
    int rc = prof_harvest(pb);
    if (rc) return rc;
    pb->prof = false;
    if (ms_assemble) *ms_assemble = pb->prof_ms_assemble;
    if (ms_solve) *ms_solve = pb->prof_ms_solve;
    if (launches) *launches = pb->prof_launches;
    if (noop_launches) *noop_launches = pb->prof_noop;
    if (instance_iterations) *instance_iterations = pb->prof_inst_iters;
    return SB_OK;
}

int sb_enable_history(sb_problem* pb, int cap) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    CU(cudaSetDevice(pb->ctx->device));
    CU(cudaStreamSynchronize(pb->ctx->stream));
    if (pb->d_hist) { guarded_free(pb->d_hist); pb->d_hist = nullptr; }
    pb->pd.hist = nullptr; pb->pd.hist_cap = 0;
    if (cap > 0) {
        CU(guarded_malloc((void**)&pb->d_hist, (size_t)pb->pd.B * cap * sizeof(double)));
        CU(cudaMemset(pb->d_hist, 0, (size_t)pb->pd.B * cap * sizeof(double)));
        pb->pd.hist = pb->d_hist; pb->pd.hist_cap = cap;
    }
    return SB_OK;
}

int sb_get_history(sb_problem* pb, double* hist_host, int cap) {
    if (!pb || !hist_host) return fail(SB_ERR_INVALID, "NULL argument");
    if (!pb->d_hist || cap != pb->pd.hist_cap) return fail(SB_ERR_INVALID, "history not enabled with cap %d", cap);
    CU(cudaSetDevice(pb->ctx->device));
    CU(cudaStreamSynchronize(pb->ctx->stream));
    CU(cudaMemcpy(hist_host, pb->d_hist, (size_t)pb->pd.B * cap * sizeof(double), cudaMemcpyDeviceToHost));
    return SB_OK;
}

int sb_debug_download_system(sb_problem* pb, double* F, double* Jl, double* Jd, double* Ju) {
    if (!pb) return fail(SB_ERR_INVALID, "problem is NULL");
    if (!pb->d_J) return fail(SB_ERR_INVALID, "no Jacobian has been assembled");
    CU(cudaSetDevice(pb->ctx->device));
    int rc = SB_OK;
    if (F) rc = download_nodes(pb, pb->d_F, F, pb->P);
    if (!rc && Jl) rc = download_nodes(pb, pb->pd.Jl, Jl, pb->P * pb->P);
    if (!rc && Jd) rc = download_nodes(pb, pb->pd.Jd, Jd, pb->P * pb->P);
    if (!rc && Ju) rc = download_nodes(pb, pb->pd.Ju, Ju, pb->P * pb->P);
    return rc;
}

int sb_debug_solve(sb_problem* pb, const double* Jl, const double* Jd, const double* Ju, const double* F, double* x) {
    if (!pb || !Jl || !Jd || !Ju || !F || !x) return fail(SB_ERR_INVALID, "NULL argument");
    if (pb->longmesh) return fail(SB_ERR_UNSUPPORTED, "sb_debug_solve is not available on the long-mesh path");
    CU(cudaSetDevice(pb->ctx->device));
    int rc = ensure_jacobian(pb);
    if (rc) return rc;
    const int PP = pb->P * pb->P;
    if ((rc = upload_nodes(pb, Jl, pb->pd.Jl, PP, 0))) return rc;
    if ((rc = upload_nodes(pb, Jd, pb->pd.Jd, PP, pb->P))) return rc;
    if ((rc = upload_nodes(pb, Ju, pb->pd.Ju, PP, 0))) return rc;
    if ((rc = upload_nodes(pb, F, pb->d_F, pb->P, 0))) return rc;
    dim3 grid, block; size_t smem;
    launch_geometry(pb, grid, block, smem);
    if ((rc = configure_small_kernels(pb))) return rc;
    CU(launch_any(pb->ks.solve, grid, block, smem, pb->ctx->stream, pb->pd, 0.0, 0, 0));
    CU(cudaGetLastError());
    return download_nodes(pb, pb->d_F, x, pb->P);
}

}  // extern "C"
