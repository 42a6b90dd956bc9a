// sb_kernels.cuh -- sm_100a kernels of the batched Skeel-Berzins Newton iteration.
//
// Data layout in HBM ("chunk-transposed SoA").  A system (one PDE instance, Nx nodes, P = npde
// components) is owned by TT threads; thread t owns the R consecutive nodes i = t*R + r, r < R
// (padded to L = R*TT nodes).  Every per-node array is stored as [instance][r][t][P] so that for a
// fixed r the TT threads of a system touch consecutive addresses: all global loads/stores of the hot
// kernels are fully coalesced while each thread still owns a contiguous piece of the mesh, which is
// what the partitioned (block-)Thomas elimination needs.  Mesh-dependent weights use the same
// [r][t] layout and are shared by the whole batch (L1/L2 resident).
//
// Kernels (template parameters: functor Fn, rows per thread R, threads per system TT (32 = one warp
// per system, shuffles only; 64/128/256 = one CTA per system), SYM0 = coordinate symmetry m == 0)
//   k_assemble<Fn,R,TT,SYM0,JAC>  K3: residual F and block-tridiagonal Jacobian (JAC) or assemble! (!JAC)
//   k_solve<P,R,TT>               K4/K5+K7: partitioned block-Thomas, update, ||h||_2, stop test
//   k_newton_fused<Fn,R,TT,SYM0>  K3+K4+K7 in one kernel: J and F never leave the SM
//   k_begin_step, k_mass_flags
#pragma once
#include "sb_banded.cuh"
#include "sb_dual.cuh"
#include "sb_functors.cuh"

namespace sb {

constexpr int kWarpBlockThreads = 128;  // TT == 32: 4 independent systems per CTA
constexpr int kMaxBlockT = 256;         // TT > 32: TT threads = one system per CTA

struct MeshDev {           // [R][TT] each; interval j = (node j, node j+1) is stored at node j's slot
    const double* xi;      // quadrature point of the interval (x argument of pdefun)
    const double2* awbw;   // U  = aw*ul + bw*ur                      (src/utils.jl:18-48)
    const double2* gwwf;   // Ux = gw*(ur - ul) ; flux weight wf = xi^m (regular, src/assembler.jl:84)
                           //                                      | zeta^(m+1)/xi (singular, :74)
    const double2* frac;   // .x = (zeta_i^(m+1) - x_i^(m+1))/(m+1)      frac1, src/assembler.jl:68 (:36 for i=0)
                           // .y = (x_i^(m+1) - zeta_{i-1}^(m+1))/(m+1)  frac2, src/assembler.jl:69 (:98 for i=N-1)
    const double* gw;      // copy of gwwf.x (the only interval weight needed when m == 0)
    const double* ivol;    // 1/(frac1 + frac2): reciprocal control-volume weight of interior nodes (c == 1 functors)
    const double* xs;      // the mesh points themselves in the same [r][t] layout (long-mesh path, m == 0: every weight
                           // above is a few flops away from x, so the level-0 sweeps read 8 instead of 32 bytes per node)
};

// weight loads: through the read-only data path when the arrays are in global memory (GLOB), plain loads
// when a kernel has staged them in shared memory
template <bool GLOB> SB_D double ldw(const double* p, int i) { if constexpr (GLOB) return __ldg(p + i); else return p[i]; }
template <bool GLOB> SB_D double2 ldw(const double2* p, int i) { if constexpr (GLOB) return __ldg(p + i); else return p[i]; }

struct ProbDev {
    int Nx, R, T, L;       // L = padded nodes per instance = ntiles * R*T (ntiles = 1 on the batched path)
    int Lt, ntiles;        // tile length R*T and number of tiles of an instance (long-mesh path: ntiles > 1)
    int m, singular, nprm;
    int nprm_s;            // stride of the per-instance parameter block (even, so every block is 16-byte aligned)
    int iter_no;           // Newton iteration number inside the current step (same for every active instance)
    int lazy_b;            // fused_tma: this step started without k_begin_step: at iter_no == 0 the iterate is b itself
    int64_t B;
    double x0, xN, x0m, xNm, xi0;  // first/last mesh point, x^m of both, first quadrature point
    MeshDev mesh;
    const double* prm;     // [B][nprm_s]
    double* u;             // Newton iterate          [B][R][T][P]
    double* b;             // M .* u_n                [B][R][T][P]
    double* F;             // residual / rhs          [B][R][T][P]
    double* Jl;            // sub-diagonal blocks     [B][R][T][P*P]
    double* Jd;
    double* Ju;
    const double* mass;    // [B][4][P]: left node, interior, right node (0/1), padding (16-byte aligned blocks)
    int* active;           // [B]
    int* iters;            // [B] Newton iterations this step
    long long* total_iters;  // [B]
    double* norm;          // [B] last ||h||_2
    int* status;           // [B]
    double* hist;          // [B][hist_cap] or null
    int hist_cap;
    unsigned int* active_count;  // [slots]
    int* list[2];          // compacted indices of the instances still active (ping-pong between iterations)
    unsigned int* done_count;             // [slots] groups (systems' thread groups / CTAs) that finished the launch
    volatile unsigned int* host_count;    // [3][slots] mapped pinned host memory: active count of a finished launch,
                                          // and (device-side stepping) the time step / iteration it belonged to
    // Device-side time stepping (sb_solve_steps): a launch that finds every instance converged starts the next time
    // step itself, so the host's look-ahead launches are useful work instead of no-ops; the host follows through
    // host_count.  Off (auto_on == 0) for every other entry point.
    int auto_on, auto_nsteps;
    int* auto_state;                      // [2] time step, Newton iteration inside it: what the next launch continues with
    const double* auto_tgrid;             // [auto_nsteps + 1]
    double auto_tau;                      // constant time step, or <= 0: tgrid[s+1] - tgrid[s]
    // Workspace of the pivoted fallback solve (sb_banded.cuh): piv_slots band arrays of piv_stride doubles each (band
    // followed by the right-hand side), handed out through piv_lock[slot] (0 free, 1 taken)
    double* piv_band;
    int* piv_lock;
    int piv_slots;
    unsigned long long piv_stride;
};

// Completion notice of a Newton-iteration launch without a copy node in the stream: every thread group takes a
// ticket when it is done; the last one publishes the number of still-active instances straight into mapped
// host memory, where the host loop polls it (src/utils.jl:329-333 is the decision it feeds).
// The counters live in a ring of kRingSlots launches that cleans itself: the launch that closes slot s zeroes slot
// s + kRingSlots/2, which no launch in flight can be using, so no memset ever sits between two Newton-iteration kernels.
constexpr int kRingSlots = 128;
SB_D void group_done(const ProbDev& pd, int slot, unsigned int total_groups, int auto_step = -1, int auto_iter = 0) {
    __threadfence();
    const unsigned int ticket = atomicAdd(pd.done_count + slot, 1u);
    if (ticket == total_groups - 1) {
        __threadfence();
        const unsigned int c = atomicAdd(pd.active_count + slot, 0u);
        if (auto_step >= 0) {  // device-side stepping: where the next launch continues, and what this one was
            pd.auto_state[0] = auto_step;
            pd.auto_state[1] = auto_iter + 1;
            pd.host_count[kRingSlots + slot] = (unsigned int)auto_step;
            pd.host_count[2 * kRingSlots + slot] = (unsigned int)auto_iter;
            __threadfence_system();
        }
        pd.host_count[slot] = c;
        const int far = (slot + kRingSlots / 2) % kRingSlots;
        pd.active_count[far] = 0u;
        pd.done_count[far] = 0u;
        __threadfence_system();
    }
}

// reciprocal without the special-case slow path of 1.0/x: MUFU.RCP64H seed + two Newton steps
// (full double precision for normal, non-zero arguments; a zero pivot turns into NaN and is
// reported through the non-finite status instead of being patched up)
SB_D double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#ifdef SB_RCP_SHORT   // tuning: r0*(1+e)*(1+e^2), a dependency chain of 3 instead of 4 (last bit not self-corrected)
    const double e = fma(-x, r, 1.0);
    const double p = fma(r, e, r), e2 = e * e;
    return fma(p, e2, p);
#else
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
#endif
}

template <int P> SB_D Mat<P> inverse_fast(const Mat<P>& A) {
    if constexpr (P == 1) { Mat<P> r; r.a[0] = fast_rcp(A.a[0]); return r; }
    else if constexpr (P == 2) {
        const bool sw = fabs(A.a[2]) > fabs(A.a[0]);  // partial pivoting inside the block
        const double a = sw ? A.a[2] : A.a[0], b = sw ? A.a[3] : A.a[1];
        const double c = sw ? A.a[0] : A.a[2], d = sw ? A.a[1] : A.a[3];
        const double ia = fast_rcp(a);
        const double l = c * ia;
        const double is = fast_rcp(d - l * b);
        const double bia = b * ia;
        const double i11 = ia + bia * l * is, i12 = -bia * is, i21 = -l * is, i22 = is;
        Mat<P> r;
        r.a[0] = sw ? i12 : i11; r.a[1] = sw ? i11 : i12;
        r.a[2] = sw ? i22 : i21; r.a[3] = sw ? i21 : i22;
        return r;
    } else {
        return inverse(A);
    }
}

// Barrier of the TT threads that own a system: a warp (TT == 32), the whole CTA (bar == 0), or -- when a CTA holds
// several systems of TT > 32 threads each (the whole-solve kernel at TT == 64) -- the named barrier `bar` (1 .. 15)
// with TT participants.
template <int TT> SB_D void group_sync(int bar = 0) {
    if constexpr (TT == 32) __syncwarp();
    else if (bar == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(TT) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Partitioned block-Thomas ("SPIKE"-type) elimination of one block-tridiagonal system.
// Rows arrive one by one (row(r, A, B, C, d): A x_{i-1} + B x_i + C x_{i+1} = d).  The thread
// eliminates its R-1 interior rows against its last row (the separator); the TT separators form a
// reduced block-tridiagonal system.  TT == 32: the reduced system is solved by parallel cyclic
// reduction with warp shuffles.  TT > 32: it is handed to warp 0, whose lanes treat TT/32
// consecutive reduced rows each with the same elimination (second level) and finish with the
// 32-row shuffle PCR.  Interior unknowns follow by back substitution.  No pivoting between block
// rows (the Newton matrices of this discretisation are block diagonally dominant); pivoting
// inside P x P blocks.
// ---------------------------------------------------------------------------------------------
template <int P, int R>
struct ChunkThomas {
    static constexpr int RI = (R > 1) ? R - 1 : 1;
    static constexpr int NW = 2 * P * P + P;
    Mat<P> ct[RI], vt[RI];
    Vec<P> dt[RI];
    Mat<P> As, Bs, Cs;
    Vec<P> ds;

    // Pivot-growth monitor.  Without row interchanges the elimination is only as good as its normalised factors are
    // small: on (block) diagonally dominant matrices every entry of ct, inv*A and of the cyclic-reduction rows stays
    // below 1 in magnitude.  mon = largest exponent word seen; above kGrowthHi (2^14; NaN and Inf included) the caller
    // re-solves the instance with the pivoted banded LU of sb_banded.cuh.  Integer max on the high word: two ALU
    // instructions per watched value, nothing on the fp64 pipe.
    static constexpr unsigned kGrowthHi = (1023u + 14u) << 20;
    unsigned mon = 0;
    SB_D void watch(double v) {
#ifndef SB_NO_PIVOT_MONITOR
        mon = max(mon, (unsigned)__double2hiint(v) & 0x7fffffffu);
#endif
    }
    SB_D void watch(const Mat<P>& M) {
#pragma unroll
        for (int i = 0; i < P * P; ++i) watch(M.a[i]);
    }

    SB_D void row(int r, const Mat<P>& A, const Mat<P>& Bm, const Mat<P>& C, const Vec<P>& d) {
        if (r == R - 1) { As = A; Bs = Bm; Cs = C; ds = d; return; }
        if (r == 0) {
            const Mat<P> inv = inverse_fast(Bm);
            ct[0] = mul(inv, C); vt[0] = mul(inv, A); dt[0] = mul(inv, d);
            watch(ct[0]); watch(vt[0]);
        } else {
            const Mat<P> inv = inverse_fast(sub_mul(Bm, A, ct[r - 1]));
            const Mat<P> la = mul(inv, A);                    // normalised sub-diagonal block of the row
            ct[r] = mul(inv, C);
            dt[r] = sub_mul(mul(inv, d), la, dt[r - 1]);
            vt[r] = neg(mul(la, vt[r - 1]));
            watch(ct[r]); watch(la);
        }
    }

    static SB_D void pack(double (&w)[NW], const Mat<P>& a, const Mat<P>& c, const Vec<P>& d) {
#pragma unroll
        for (int i = 0; i < P * P; ++i) { w[i] = a.a[i]; w[P * P + i] = c.a[i]; }
#pragma unroll
        for (int i = 0; i < P; ++i) w[2 * P * P + i] = d.a[i];
    }
    static SB_D void unpack(const double (&w)[NW], Mat<P>& a, Mat<P>& c, Vec<P>& d) {
#pragma unroll
        for (int i = 0; i < P * P; ++i) { a.a[i] = w[i]; c.a[i] = w[P * P + i]; }
#pragma unroll
        for (int i = 0; i < P; ++i) d.a[i] = w[2 * P * P + i];
    }

    // shared memory of finish<TT>() in doubles: the exchange region (two equations per warp) ...
#ifdef SB_FINISH_PARALLEL   // tuning builds: every warp reduces its own separators (see finish_parallel)
    template <int TT> static __host__ __device__ constexpr int exch_doubles() { return TT == 32 ? 0 : (2 * NW + 1) * (TT / 32); }
#else                       // default: warp 0 solves the TT separators (finish_serial): warps' first rows | reduced rows | unknowns | monitor words
    template <int TT> static __host__ __device__ constexpr int exch_doubles() { return TT == 32 ? 0 : (NW + P) * TT + (NW + 1) * (TT / 32) + 1; }
#endif
    // the all-warps variant reuses its exchange region while slow warps may still read it: two buffers; the serial one
    // has a barrier between every write and the reads it could disturb: one
#ifdef SB_FINISH_PARALLEL
    static constexpr int exch_buffers = 2;
#else
    static constexpr int exch_buffers = 1;
#endif
    // ... followed by one partial sum per warp for system_sum()
    template <int TT> static __host__ __device__ constexpr int smem_doubles() { return TT == 32 ? 0 : exch_doubles<TT>() + TT / 32; }

    // representation of the chunk's first row: x_first = y0 - v0*x_prevsep - w0*x_sep
    // and of its last interior row:             x_last  = yL - vL*x_prevsep - wL*x_sep
    // (from the forward factors, which stay intact: the long-mesh path recomputes them on the way back)
    SB_D void summary(Vec<P>& y0, Mat<P>& v0, Mat<P>& w0, Vec<P>& yL, Mat<P>& vL, Mat<P>& wL) const {
        if constexpr (R == 1) {
            y0 = vec_zero<P>(); v0 = mat_zero<P>(); w0 = neg(mat_identity<P>());
            yL = vec_zero<P>(); vL = neg(mat_identity<P>()); wL = mat_zero<P>();
        } else {
            yL = dt[R - 2]; vL = vt[R - 2]; wL = ct[R - 2];
            y0 = yL; v0 = vL; w0 = wL;
#pragma unroll
            for (int r = R - 3; r >= 0; --r) {
                y0 = sub_mul(dt[r], ct[r], y0);
                v0 = sub_mul(vt[r], ct[r], v0);
                w0 = neg(mul(ct[r], w0));
            }
        }
    }
    // The same recurrence kept for EVERY interior row, in place: afterwards x_r = dt[r] - vt[r]*x_prevsep - ct[r]*x_sep,
    // so the back substitution has no chain between rows (explicit()).
    SB_D void to_explicit() {
#pragma unroll
        for (int r = R - 3; r >= 0; --r) {
            dt[r] = sub_mul(dt[r], ct[r], dt[r + 1]);
            vt[r] = sub_mul(vt[r], ct[r], vt[r + 1]);
            ct[r] = neg(mul(ct[r], ct[r + 1]));
        }
    }
    SB_D void explicit_rows(Vec<P> (&x)[R], const Vec<P>& xp, const Vec<P>& xs) const {
        x[R - 1] = xs;
#pragma unroll
        for (int r = 0; r < R - 1; ++r) x[r] = sub_mul(sub_mul(dt[r], vt[r], xp), ct[r], xs);
    }
    // reduced row of the chunk's separator, normalised to unit diagonal, given the first-row representation
    // (v0n, w0n, y0n) of the next chunk
    static SB_D void reduced_row(const Mat<P>& As_, const Mat<P>& Bs_, const Mat<P>& Cs_, const Vec<P>& ds_, const Vec<P>& yL,
                                 const Mat<P>& vL, const Mat<P>& wL, const Mat<P>& v0n, const Mat<P>& w0n, const Vec<P>& y0n,
                                 Mat<P>& an, Mat<P>& cn, Vec<P>& dn) {
        const Mat<P> ar = neg(mul(As_, vL));
        const Mat<P> br = sub_mul(sub_mul(Bs_, As_, wL), Cs_, v0n);
        const Mat<P> cr = neg(mul(Cs_, w0n));
        const Vec<P> dr = sub_mul(sub_mul(ds_, As_, yL), Cs_, y0n);
        const Mat<P> binv = inverse_fast(br);
        an = mul(binv, ar); cn = mul(binv, cr); dn = mul(binv, dr);
    }
    // unknowns of the chunk's rows from its own (xs) and the previous (xp) separator unknowns (forward factors)
    SB_D void back(Vec<P> (&x)[R], const Vec<P>& xp, const Vec<P>& xs) const {
        x[R - 1] = xs;
#pragma unroll
        for (int r = R - 2; r >= 0; --r) x[r] = sub_mul(sub_mul(dt[r], vt[r], xp), ct[r], x[r + 1]);
    }

    struct NoHook { SB_D void operator()() const {} };

    template <int TT>
    SB_D bool finish(Vec<P> (&x)[R], double* sm, int t, int bar = 0) { return finish<TT>(x, sm, t, NoHook(), bar); }

    // Solve.  x[r] = unknowns of the thread's rows, t = thread index inside the system, sm = exch_doubles<TT>() doubles.
    // hook(): called by every thread right after the first synchronisation point (TT == 32: after the first shuffle),
    // when all threads of the system have finished reading their inputs, e.g. to start the prefetch of the next instance.
    // Returns true (the same value in every thread of the system) when the pivot-growth monitor tripped anywhere in
    // the system: x is then not to be trusted and the caller solves the instance again with row interchanges.
    //
    // Level 1 (thread): the R-1 interior rows were eliminated against the separator by row(); to_explicit() turns the
    // factors into x_r = y_r - v_r*x_prevsep - w_r*x_sep.  The TT separators form a reduced block-tridiagonal system.
    // TT == 32: parallel cyclic reduction with shuffles (finish_parallel).  TT > 32, two variants, both measured:
    //  * finish_serial (default): the reduced rows go to shared memory and warp 0 solves them alone -- TT/32 rows per
    //    lane by the same partition, then the 32-row shuffle PCR -- while the other warps wait at a barrier.  Fewest
    //    instructions (739 per warp and system at (8,128)); the idle warps' issue slots go to the other CTAs of the SM.
    //  * finish_parallel (-DSB_FINISH_PARALLEL): every warp runs the shuffle PCR on its own 32 separators with the
    //    couplings to the neighbouring warps carried along as spikes, then one exchange of two equations per warp and
    //    a redundant elimination through the TT/32 warps: one barrier instead of three, no idle warp, but 914
    //    instructions.  Per-iteration launches (k_newton_fused_p): same speed (35.5 vs 35.1 us per cfg2 launch, barrier
    //    stall 1.4 vs 3.7 per issue).  Whole-solve kernel: 13.5 ms vs 11.8 ms per cfg2 solve -- instructions win.
    // bar: the barrier of the system's threads (group_sync)
    template <int TT, class Hook>
    SB_D bool finish(Vec<P> (&x)[R], double* sm, int t, Hook hook, int bar = 0) {
#ifndef SB_FINISH_PARALLEL
        if constexpr (TT > 32) return finish_serial<TT>(x, sm, t, hook, bar);
        else
#endif
        return finish_parallel<TT>(x, sm, t, hook, bar);
    }

    template <int TT, class Hook>
    SB_D bool finish_serial(Vec<P> (&x)[R], double* sm, int t, Hook hook, int bar = 0) {
        constexpr int W = TT / 32;
        const int lane = t & 31, wi = t >> 5;
        Vec<P> y0, yL; Mat<P> v0, w0, vL, wL;
        if constexpr (R == 1) summary(y0, v0, w0, yL, vL, wL);
        else { to_explicit(); y0 = dt[0]; v0 = vt[0]; w0 = ct[0]; yL = dt[R - 2]; vL = vt[R - 2]; wL = ct[R - 2]; }
        // first-row representation of thread t+1: by shuffle inside the warp.  At a warp's last lane the first row of the
        // next warp stays in the reduced row as an unknown (coefficient c) -- warp 0 eliminates it below with the
        // representation that warp's lane 0 leaves in shared memory -- so no barrier is needed before the reduced rows.
        double mine[NW], nb[NW];
        pack(mine, v0, w0, y0);
        double* frow = sm;                       // [W][NW] first-row representations of the warps' first threads
        double* red = sm + NW * W;               // [NW][TT] reduced rows
        double* xsm = red + NW * TT;             // [P][TT] separator unknowns
        unsigned* monw = reinterpret_cast<unsigned*>(xsm + P * TT);   // [W] per-warp monitors, [2W] the verdict
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const double v = __shfl_down_sync(0xffffffffu, mine[w], 1);
            const bool wdiag = (w >= P * P) && (w < 2 * P * P) && ((w - P * P) / P == (w - P * P) % P);
            nb[w] = (lane < 31) ? v : ((wdiag && t != TT - 1) ? -1.0 : 0.0);
            if (lane == 0) frow[wi * NW + w] = mine[w];
        }
        Mat<P> v0n, w0n; Vec<P> y0n;
        unpack(nb, v0n, w0n, y0n);
        Mat<P> an, cn; Vec<P> dn;
        reduced_row(As, Bs, Cs, ds, yL, vL, wL, v0n, w0n, y0n, an, cn, dn);
        watch(an); watch(cn);
        constexpr int Q = TT / 32;  // reduced rows per lane of warp 0
        pack(mine, an, cn, dn);
#pragma unroll
        for (int w = 0; w < NW; ++w) red[w * TT + t] = mine[w];
#ifndef SB_NO_PIVOT_MONITOR
        {
            const unsigned g = __reduce_max_sync(0xffffffffu, mon);
            if (lane == 0) monw[wi] = g;
        }
#endif
        group_sync<TT>(bar);
        hook();
        if (t < 32) {
            ChunkThomas<P, Q> lvl2;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int j = t * Q + q;                       // reduced row = separator of thread j
                double rw[NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) rw[w] = red[w * TT + j];
                Mat<P> a, c; Vec<P> d;
                unpack(rw, a, c, d);
                if ((j & 31) == 31 && j != TT - 1) {           // last thread of a warp: x_j + a x_{j-1} + c F = d with
                    double fw[NW];                             // F = y0' - v0' x_j - w0' x_{j+1} the next warp's first row
#pragma unroll
                    for (int w = 0; w < NW; ++w) fw[w] = frow[((j >> 5) + 1) * NW + w];
                    Mat<P> v0f, w0f; Vec<P> y0f;
                    unpack(fw, v0f, w0f, y0f);
                    const Mat<P> inv = inverse_fast(sub_mul(mat_identity<P>(), c, v0f));
                    d = mul(inv, sub_mul(d, c, y0f));
                    a = mul(inv, a);
                    c = neg(mul(inv, mul(c, w0f)));
                    lvl2.watch(a); lvl2.watch(c);
                }
                lvl2.row(q, a, mat_identity<P>(), c, d);
            }
            Vec<P> y[Q];
            const bool bad2 = lvl2.template finish_parallel<32>(y, nullptr, t, NoHook());
#pragma unroll
            for (int q = 0; q < Q; ++q)
#pragma unroll
                for (int i = 0; i < P; ++i) xsm[i * TT + t * Q + q] = y[q].a[i];
            if (t == 0) {
                unsigned g = bad2 ? 0x7ff00000u : 0u;
#pragma unroll
                for (int w = 0; w < W; ++w) g = max(g, monw[w]);
                monw[2 * W] = g;
            }
        }
        group_sync<TT>(bar);
        Vec<P> xs, xp;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            xs.a[i] = xsm[i * TT + t];
            xp.a[i] = (t > 0) ? xsm[i * TT + t - 1] : 0.0;
        }
        if constexpr (R == 1) x[0] = xs;
        else explicit_rows(x, xp, xs);
#ifndef SB_NO_PIVOT_MONITOR
        return monw[2 * W] > kGrowthHi;
#else
        return false;
#endif
    }

    // The all-warps variant (and the solver of one warp's worth of rows: TT == 32, and level 2 of finish_serial):
    //  warp: the 32 separators of a warp form a tridiagonal system whose two outer couplings reach into the
    //     neighbouring warps: lane 0 to X_prev (the previous warp's last separator), lane 31 to X_next (the FIRST
    //     ROW of the next warp's first chunk, kept as an unknown instead of being eliminated, so nothing has to cross
    //     the warp boundary beforehand).  Parallel cyclic reduction with shuffles; an elimination step that would
    //     reach outside the warp is skipped, which leaves the coefficient of X_prev / X_next in the a / c slot
    //     ("spikes" at no cost).  After 5 steps  x_sep(lane) = d - a*X_prev - c*X_next.
    //  CTA: per warp two equations in (X_prev, own first row F, own last separator L, X_next) go to shared memory;
    //     after the barrier every thread eliminates through the TT/32 warps (a few dozen flops) and keeps the two
    //     values its own warp needs.
    template <int TT, class Hook>
    SB_D bool finish_parallel(Vec<P> (&x)[R], double* sm, int t, Hook hook, int bar = 0) {
        constexpr int W = TT / 32;
        const int lane = t & 31;
        Vec<P> y0, yL; Mat<P> v0, w0, vL, wL;
        if constexpr (R == 1) {
            summary(y0, v0, w0, yL, vL, wL);
        } else {
            to_explicit();
            y0 = dt[0]; v0 = vt[0]; w0 = ct[0];
            yL = dt[R - 2]; vL = vt[R - 2]; wL = ct[R - 2];
        }
        // (v0, w0, y0) of thread t+1; lane 31 keeps the next warp's first row as the unknown X_next: x_first = X_next
        double mine[NW], nb[NW];
        pack(mine, v0, w0, y0);
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const double v = __shfl_down_sync(0xffffffffu, mine[w], 1);
            const bool wdiag = (w >= P * P) && (w < 2 * P * P) && ((w - P * P) / P == (w - P * P) % P);
            nb[w] = (lane < 31) ? v : ((wdiag && t != TT - 1) ? -1.0 : 0.0);   // (the system's last row has no successor)
        }
        if constexpr (W == 1) hook();
        Mat<P> v0n, w0n; Vec<P> y0n;
        unpack(nb, v0n, w0n, y0n);
        Mat<P> an, cn; Vec<P> dn;
        reduced_row(As, Bs, Cs, ds, yL, vL, wL, v0n, w0n, y0n, an, cn, dn);
        watch(an); watch(cn);

#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {  // parallel cyclic reduction across the warp
            double om[NW], op[NW];
            pack(mine, an, cn, dn);
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                om[w] = __shfl_up_sync(0xffffffffu, mine[w], s);
                op[w] = __shfl_down_sync(0xffffffffu, mine[w], s);
            }
            const bool hu = lane >= s, hd = lane + s < 32;   // the partner row exists inside the warp
            Mat<P> am, cm, ap, cp; Vec<P> dm, dp;
            unpack(om, am, cm, dm);
            unpack(op, ap, cp, dp);
            Mat<P> anm = an, cnp = cn;                        // multipliers of the elimination (0: step skipped)
#pragma unroll
            for (int i = 0; i < P * P; ++i) { anm.a[i] = hu ? an.a[i] : 0.0; cnp.a[i] = hd ? cn.a[i] : 0.0; }
            const Mat<P> b2 = sub_mul(sub_mul(mat_identity<P>(), anm, cm), cnp, ap);
            const Vec<P> d2 = sub_mul(sub_mul(dn, anm, dm), cnp, dp);
            Mat<P> a2 = neg(mul(an, am)), c2 = neg(mul(cn, cp));
            if constexpr (W > 1) {                            // skipped step: the coefficient stays (it multiplies X_prev / X_next)
#pragma unroll
                for (int i = 0; i < P * P; ++i) { a2.a[i] = hu ? a2.a[i] : an.a[i]; c2.a[i] = hd ? c2.a[i] : cn.a[i]; }
            } else {                                          // one warp = the whole system: nothing outside
#pragma unroll
                for (int i = 0; i < P * P; ++i) { a2.a[i] = hu ? a2.a[i] : 0.0; c2.a[i] = hd ? c2.a[i] : 0.0; }
            }
            const Mat<P> inv = inverse_fast(b2);
            an = mul(inv, a2); cn = mul(inv, c2); dn = mul(inv, d2);
            watch(an); watch(cn);
        }
#ifndef SB_NO_PIVOT_MONITOR
        unsigned grow = __reduce_max_sync(0xffffffffu, mon);   // largest factor seen anywhere in the warp
#else
        unsigned grow = 0;
#endif

        Vec<P> xs = dn, xp;
        Vec<P> Xp = vec_zero<P>();
        if constexpr (W > 1) {
            const int wi = t >> 5;
            // warp wi:  F + fa*L(wi-1) + fc*F(wi+1) = fd   (its first row, lane 0)
            //           L + la*L(wi-1) + lc*F(wi+1) = ld   (its last separator, lane 31)
            double* eq = sm + (size_t)wi * 2 * NW;
            if (lane == 0) {
                pack(mine, sub_mul(v0, w0, an), neg(mul(w0, cn)), sub_mul(y0, w0, dn));
#pragma unroll
                for (int w = 0; w < NW; ++w) eq[w] = mine[w];
            }
            if (lane == 31) {
                pack(mine, an, cn, dn);
#pragma unroll
                for (int w = 0; w < NW; ++w) eq[NW + w] = mine[w];
                reinterpret_cast<unsigned*>(sm + 2 * NW * W)[2 * wi] = grow;
            }
            group_sync<TT>(bar);
            hook();
#pragma unroll
            for (int w = 0; w < W; ++w) grow = max(grow, reinterpret_cast<const unsigned*>(sm + 2 * NW * W)[2 * w]);
            mon = grow;
            // forward: L(w) = al[w] - be[w]*F(w+1),  F(w) = ga[w] - de[w]*F(w+1)
            Vec<P> al[W], ga[W]; Mat<P> be[W], de[W];
            {
                Mat<P> la;
                double rw[NW];
#pragma unroll
                for (int w = 0; w < NW; ++w) rw[w] = sm[NW + w];
                unpack(rw, la, be[0], al[0]);                 // warp 0 has no predecessor: la = 0
            }
#pragma unroll
            for (int w = 1; w < W; ++w) {
                double rw[NW];
                Mat<P> fa, fc; Vec<P> fd;
#pragma unroll
                for (int q = 0; q < NW; ++q) rw[q] = sm[(size_t)w * 2 * NW + q];
                unpack(rw, fa, fc, fd);
                const Mat<P> G = inverse_fast(sub_mul(mat_identity<P>(), fa, be[w - 1]));
                ga[w] = mul(G, sub_mul(fd, fa, al[w - 1]));
                de[w] = mul(G, fc);
                watch(de[w]);
                if (w < W - 1) {
                    Mat<P> la, lc; Vec<P> ld;
#pragma unroll
                    for (int q = 0; q < NW; ++q) rw[q] = sm[(size_t)w * 2 * NW + NW + q];
                    unpack(rw, la, lc, ld);
                    const Mat<P> lab = mul(la, be[w - 1]);
                    al[w] = sub_mul(sub_mul(ld, la, al[w - 1]), neg(lab), ga[w]);
                    be[w] = sub_mul(lc, neg(lab), de[w]);
                    watch(be[w]);
                }
            }
            grow = mon;                                       // the same in every thread: identical arithmetic on shared data
            // backward: every warp picks up X_prev = L(wi-1) and X_next = F(wi+1)
            Vec<P> Fn = vec_zero<P>(), Xn = vec_zero<P>();
#pragma unroll
            for (int w = W - 1; w >= 1; --w) {
                const Vec<P> Fw = sub_mul(ga[w], de[w], Fn);
                const Vec<P> Lp = sub_mul(al[w - 1], be[w - 1], Fw);
                if (wi == w) { Xp = Lp; Xn = Fn; }
                Fn = Fw;
            }
            if (wi == 0) Xn = Fn;
            xs = sub_mul(sub_mul(dn, an, Xp), cn, Xn);
        }
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const double v = __shfl_up_sync(0xffffffffu, xs.a[i], 1);
            xp.a[i] = (lane > 0) ? v : Xp.a[i];
        }
        if constexpr (R == 1) x[0] = xs;
        else explicit_rows(x, xp, xs);
        return grow > kGrowthHi;
    }
};

// sum over the TT threads of a system, result valid in thread t == 0.  sm: TT/32 doubles that no thread of the
// system reads any more (the partial-sum slots behind the exchange region of finish())
template <int TT>
SB_D double system_sum(double v, double* sm, int t) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if constexpr (TT > 32) {
        if ((t & 31) == 0) sm[t >> 5] = v;
        __syncthreads();
        if (t == 0) {
#pragma unroll
            for (int w = 1; w < TT / 32; ++w) v += sm[w];
        }
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// Assembly of the rows of one chunk (assemble!, src/assembler.jl:19-125, plus implicitEuler!,
// src/utils.jl:245-257, and the Jacobian the reference obtains by forward-mode AD, :312-314).
// Functor traits: Fn::c_unit (c == 1 identically), Fn::s_zero (s == 0 identically) let the
// compiler drop the corresponding dual arithmetic.
// ---------------------------------------------------------------------------------------------
// optional functor trait: `static constexpr bool s_const = true;` -- the source s does not depend on u or dudx (its partials
// are structural zeros the compiler cannot fold: 0 * x is not 0 for a NaN x), so the Jacobian rows skip its terms
template <class Fn, class = void> struct fn_s_const { static constexpr bool value = false; };
template <class Fn> struct fn_s_const<Fn, decltype((void)Fn::s_const)> { static constexpr bool value = Fn::s_const; };

template <int P, int ND>
struct IntervalRes {  // (c,f,s) of one interval; partials w.r.t. (u_left[0..P), u_right[0..P))
    Dual<ND> c[P], f[P], s[P];
    double wf;
};

// XW (m == 0 only): the interval's weights come from its two mesh points xl, xr instead of the weight arrays -- the
// same expressions as the set-up (k_mesh_weights_m0); the reciprocals are fast_rcp's (within an ulp of the set-up's
// divisions; __drcp_rn would agree bit for bit at about twice the instructions and a branch per reciprocal).
template <class Fn, int ND, bool SYM0, bool GLOB = true, bool XW = false>
SB_D void eval_interval(const MeshDev& mesh, int slot, const double* ul, const double* ur, double tt, const double* prm,
                        IntervalRes<Fn::npde, ND>& out, double xl = 0.0, double xr = 0.0) {
    constexpr int P = Fn::npde;
    static_assert(!XW || SYM0, "weights from x: m == 0 only");
    double xi;
    if constexpr (XW) xi = (xl + xr) / 2;
    else xi = __ldg(mesh.xi + slot);  // always global; dead (and removed) when the functor ignores x
    double aw, bw, gw;
    if constexpr (XW) {
        aw = 0.5; bw = 0.5; gw = fast_rcp(xr - xl); out.wf = 1.0;
    } else if constexpr (SYM0) {  // m == 0: xi is the midpoint, linear interpolation, unit flux weight
        aw = 0.5; bw = 0.5; gw = ldw<GLOB>(mesh.gw, slot); out.wf = 1.0;
    } else {
        const double2 ab = ldw<GLOB>(mesh.awbw, slot), gf = ldw<GLOB>(mesh.gwwf, slot);
        aw = ab.x; bw = ab.y; gw = gf.x; out.wf = gf.y;
    }
    Dual<ND> U[P], Ux[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        U[p] = Dual<ND>(SYM0 ? 0.5 * (ul[p] + ur[p]) : aw * ul[p] + bw * ur[p]);
        Ux[p] = Dual<ND>(gw * (ur[p] - ul[p]));
        if constexpr (ND > 0) {
            U[p].d[p] = aw; U[p].d[P + p] = bw;
            Ux[p].d[p] = -gw; Ux[p].d[P + p] = gw;
        }
    }
    Fn::template pde<Dual<ND>>(xi, tt, U, Ux, prm, out.c, out.f, out.s);
}

template <int P>
struct RowOut { Mat<P> A, Bm, C; Vec<P> d; };

// Boundary rows (src/assembler.jl:36-57 left, :98-122 right).  side = 0: node 0 with I = interval 0;
// side = 1: node Nx-1 with I = interval Nx-2.  ui/bi: u and b at the boundary node.
// JAC: out = (Jl, Jd, Ju, F); !JAC: out.d = du.
// XW: the node's volume weight (frac1 at the left end, frac2 at the right end) is handed in as xfrac.
template <class Fn, bool JAC, int ND, bool GFRAC = true, bool XW = false>
SB_D void boundary_row(const ProbDev& pd, const MeshDev& mesh, const double* u0p, const double* uNp, const double* mass,
                       int side, double tt, double inv_tau, const double* prm, const IntervalRes<Fn::npde, ND>& I,
                       const double* ui, const double* bi, int slot, RowOut<Fn::npde>& out, double xfrac = 0.0) {
    constexpr int P = Fn::npde;
    constexpr int PO = JAC ? P : 0;
    Dual<PO> ul_[P], ur_[P], pl[P], pr[P];
    double ql[P], qr[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        ul_[p] = Dual<PO>(u0p[p]); ur_[p] = Dual<PO>(uNp[p]);
        if constexpr (JAC) { ul_[p].d[p] = 1.0; ur_[p].d[p] = 1.0; }
    }
    Fn::template bc<Dual<PO>>(pd.x0, ul_, pd.xN, ur_, tt, prm, pl, ql, pr, qr);  // src/assembler.jl:26/29

    double du[P], dm[P][P], dc[P][P], dp[P][P], mss[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        du[j] = 0.0;
#pragma unroll
        for (int q = 0; q < P; ++q) { dm[j][q] = 0.0; dc[j][q] = 0.0; dp[j][q] = 0.0; }
    }
    if (side == 0) {
        double frac;
        if constexpr (XW) frac = xfrac; else frac = ldw<GFRAC>(mesh.frac, slot).x;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            mss[j] = mass[j];
            const Dual<ND>&c = I.c[j], &f = I.f[j], &s = I.s[j];
            if (pd.singular) {  // src/assembler.jl:38-45
                const double mp = (double)(pd.m + 1) / pd.xi0;
                const double num = mp * f.v + s.v;
                const bool hc = c.v != 0.0;
                const double id = hc ? 1.0 / c.v : 1.0;
                du[j] = num * id;
                if constexpr (JAC) {
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        const double n0 = mp * f.d[q] + s.d[q], n1 = mp * f.d[PO + q] + s.d[PO + q];
                        dc[j][q] = hc ? (n0 - du[j] * c.d[q]) * id : n0;
                        dp[j][q] = hc ? (n1 - du[j] * c.d[PO + q]) * id : n1;
                    }
                }
            } else if (ql[j] != 0.0) {  // src/assembler.jl:48-52
                const double qx = ql[j] / pd.x0m;
                const double num = pl[j].v + qx * (I.wf * f.v + frac * s.v);
                const bool hc = c.v != 0.0;
                const double den = qx * frac * c.v;
                const double id = hc ? 1.0 / den : 1.0;
                du[j] = num * id;
                if constexpr (JAC) {
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        const double n0 = pl[j].d[q] + qx * (I.wf * f.d[q] + frac * s.d[q]);
                        const double n1 = qx * (I.wf * f.d[PO + q] + frac * s.d[PO + q]);
                        dc[j][q] = hc ? (n0 - du[j] * (qx * frac * c.d[q])) * id : n0;
                        dp[j][q] = hc ? (n1 - du[j] * (qx * frac * c.d[PO + q])) * id : n1;
                    }
                }
            } else {  // Dirichlet, src/assembler.jl:54
                du[j] = pl[j].v;
                if constexpr (JAC) {
#pragma unroll
                    for (int q = 0; q < P; ++q) dc[j][q] = pl[j].d[q];
                }
            }
        }
    } else {  // src/assembler.jl:98-122
        double frac;
        if constexpr (XW) frac = xfrac; else frac = ldw<GFRAC>(mesh.frac, slot).y;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            mss[j] = mass[2 * P + j];
            const Dual<ND>&c = I.c[j], &f = I.f[j], &s = I.s[j];
            if (qr[j] != 0.0) {
                const double qx = qr[j] / pd.xNm;
                const double num = pr[j].v + qx * (I.wf * f.v - frac * s.v);
                const bool hc = c.v != 0.0;
                const double den = -qx * frac * c.v;
                const double id = hc ? 1.0 / den : 1.0;
                du[j] = num * id;
                if constexpr (JAC) {
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        const double n0 = qx * (I.wf * f.d[q] - frac * s.d[q]);
                        const double n1 = pr[j].d[q] + qx * (I.wf * f.d[PO + q] - frac * s.d[PO + q]);
                        dm[j][q] = hc ? (n0 - du[j] * (-qx * frac * c.d[q])) * id : n0;
                        dc[j][q] = hc ? (n1 - du[j] * (-qx * frac * c.d[PO + q])) * id : n1;
                    }
                }
            } else {
                du[j] = pr[j].v;
                if constexpr (JAC) {
#pragma unroll
                    for (int q = 0; q < P; ++q) dc[j][q] = pr[j].d[q];
                }
            }
        }
    }
    if constexpr (JAC) {
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const double mt = inv_tau * mss[j];
            out.d.a[j] = (mt * ui[j] - du[j]) - inv_tau * bi[j];  // src/utils.jl:251-256, :314
#pragma unroll
            for (int q = 0; q < P; ++q) {
                out.A.a[j * P + q] = -dm[j][q];
                out.Bm.a[j * P + q] = -dc[j][q] + ((j == q) ? mt : 0.0);
                out.C.a[j * P + q] = -dp[j][q];
            }
        }
    } else {
        out.A = mat_zero<P>(); out.Bm = mat_zero<P>(); out.C = mat_zero<P>();
#pragma unroll
        for (int j = 0; j < P; ++j) out.d.a[j] = du[j];
    }
}

// Sink concept: row(r, Jl, Jd, Ju, F) when JAC, value(r, du) otherwise.
// mesh: weight arrays (global or shared memory); ub/bb: this instance's u and b in device layout (global or
// shared memory); mass: its 3*P mass flags; prm: its functor parameters.
// XW (long-mesh path, m == 0): mesh.xs is the tile of mesh points (slots -1 .. R*TT valid where the mesh goes on) and
// replaces gw, ivol and frac.
template <class Fn, int R, int TT, bool SYM0, bool JAC, bool GLOB = true, bool XW = false, class Sink>
SB_D void assemble_chunk(const ProbDev& pd, const MeshDev& mesh, const double* ub, const double* bb, const double* mass,
                         int t, double tt, double inv_tau, const double* prm, const double (&uc)[R][Fn::npde],
                         const double (&bc)[R][Fn::npde], Sink& sink, int i_base = 0, const double* u0p = nullptr,
                         const double* uNp = nullptr) {
    // Long meshes are stored as tiles of R*TT nodes: i_base = global index of the tile's first node, ub/bb/mesh
    // point at the tile, the neighbouring tiles are adjacent in memory (slot -1 and slot R*TT), and u0p/uNp point
    // at the first / last mesh node of the instance.  Batched path: one tile, i_base = 0.
    constexpr int P = Fn::npde;
    constexpr int ND = JAC ? 2 * P : 0;
    constexpr int PO = JAC ? P : 0;
    const int Nx = pd.Nx;
    const int i0 = i_base + t * R;
    if (!u0p) { u0p = ub; uNp = ub + (size_t)(((Nx - 1) % R) * TT + (Nx - 1) / R) * P; }
    double mI[P];  // interior mass flags * inv_tau
#pragma unroll
    for (int p = 0; p < P; ++p) mI[p] = inv_tau * mass[P + p];

    // The right boundary row is prepared once, by the thread that owns node Nx-1 (one inlined copy of the
    // boundary code instead of one per unrolled row); the left one in the r == 0 copy of the loop below.
    RowOut<P> rowN;
    rowN.A = mat_zero<P>(); rowN.Bm = mat_identity<P>(); rowN.C = mat_zero<P>(); rowN.d = vec_zero<P>();
    if (Nx - 1 >= i0 && Nx - 1 < i0 + R) {
        const int rN = Nx - 1 - i0;
        const int sN = rN * TT + t;
        const int sM = (rN >= 1) ? (rN - 1) * TT + t : ((t > 0) ? (R - 1) * TT + t - 1 : -1);
        double uM[P], uN[P], bN[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            uM[p] = ub[(ptrdiff_t)sM * P + p]; uN[p] = ub[(size_t)sN * P + p];
            bN[p] = JAC ? bb[(size_t)sN * P + p] : 0.0;
        }
        IntervalRes<P, ND> IN;
        double xM = 0.0, xN = 0.0;
        if constexpr (XW) { xM = mesh.xs[sM]; xN = mesh.xs[sN]; }
        eval_interval<Fn, ND, SYM0, GLOB, XW>(mesh, sM, uM, uN, tt, prm, IN, xM, xN);
        boundary_row<Fn, JAC, ND, (GLOB || (Fn::c_unit && Fn::s_zero)), XW>(pd, mesh, u0p, uNp, mass, 1, tt, inv_tau, prm, IN, uN, bN, sN, rowN,
                                                                            xN - (xM + xN) / 2);
    }

    double un[P];  // first node of the next chunk
    IntervalRes<P, ND> Lr, Rr;
    double x_prev = 0.0, x_cur = 0.0, x_next = 0.0;  // XW: mesh points around the current row
    if constexpr (XW) x_cur = mesh.xs[t];
    if (i0 >= 1 && i0 - 1 <= Nx - 2) {
        double ul[P];
        const int slot = (t > 0) ? (R - 1) * TT + t - 1 : -1;  // t == 0: last node of the previous tile
#pragma unroll
        for (int p = 0; p < P; ++p) ul[p] = ub[(ptrdiff_t)slot * P + p];
        if constexpr (XW) x_prev = mesh.xs[slot];
        eval_interval<Fn, ND, SYM0, GLOB, XW>(mesh, slot, ul, uc[0], tt, prm, Lr, x_prev, x_cur);
    } else {
#pragma unroll
        for (int p = 0; p < P; ++p) { Lr.c[p] = Dual<ND>(1.0); Lr.f[p] = Dual<ND>(0.0); Lr.s[p] = Dual<ND>(0.0); }
        Lr.wf = 0.0;
    }
    Rr = Lr;
#pragma unroll
    for (int p = 0; p < P; ++p)  // t == TT-1: first node of the next tile (if the mesh goes on)
        un[p] = (t + 1 < TT) ? ub[(size_t)(t + 1) * P + p] : ((i0 + R <= Nx - 1) ? ub[(size_t)R * TT * P + p] : 0.0);

#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = i0 + r;
        const int slot = r * TT + t;
        // evaluated for every row: for i > Nx-2 the slot is padding (zeros) and the result is never used
        if constexpr (XW) x_next = mesh.xs[(r + 1 < R) ? (r + 1) * TT + t : ((t + 1 < TT) ? t + 1 : R * TT)];
        eval_interval<Fn, ND, SYM0, GLOB, XW>(mesh, slot, uc[r], (r + 1 < R) ? uc[r + 1] : un, tt, prm, Rr, x_cur, x_next);
        const double wR = SYM0 ? 1.0 : Rr.wf, wL = SYM0 ? 1.0 : Lr.wf;  // m == 0: unit flux weights, folded away

        // interior node, src/assembler.jl:60-95 (computed for every row; boundary / padding rows are replaced below)
        double frR = 0.0, frL = 0.0;
        double ivol = 0.0;
        if constexpr (XW) {   // zeta = interval midpoints: frac1 = zeta_i - x_i, frac2 = x_i - zeta_{i-1} (m + 1 = 1)
            frR = (x_cur + x_next) / 2 - x_cur; frL = x_cur - (x_prev + x_cur) / 2;
            if constexpr (Fn::c_unit) ivol = fast_rcp(frR + frL);
        } else {
            if constexpr (!(Fn::c_unit && Fn::s_zero)) { const double2 fr = ldw<GLOB>(mesh.frac, slot); frR = fr.x; frL = fr.y; }
            if constexpr (Fn::c_unit) ivol = ldw<GLOB>(mesh.ivol, slot);
        }
        RowOut<P> row;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            double num = wR * Rr.f[j].v - wL * Lr.f[j].v;
            if constexpr (!Fn::s_zero) num += frR * Rr.s[j].v + frL * Lr.s[j].v;
            bool hc = true;
            double id;
            if constexpr (Fn::c_unit) id = ivol;
            else {
                hc = (Lr.c[j].v != 0.0) || (Rr.c[j].v != 0.0);
                id = hc ? fast_rcp(frR * Rr.c[j].v + frL * Lr.c[j].v) : 1.0;
            }
            const double du = num * id;
            if constexpr (JAC) {
                row.d.a[j] = (mI[j] * uc[r][j] - du) - inv_tau * bc[r][j];  // src/utils.jl:251-256, :314
#pragma unroll
                for (int q = 0; q < P; ++q) {
                    double nm = -(wL * Lr.f[j].d[q]);
                    double nc = wR * Rr.f[j].d[q] - wL * Lr.f[j].d[PO + q];
                    double np = wR * Rr.f[j].d[PO + q];
                    if constexpr (!Fn::s_zero && !fn_s_const<Fn>::value) {
                        nm += frL * Lr.s[j].d[q];
                        nc += frR * Rr.s[j].d[q] + frL * Lr.s[j].d[PO + q];
                        np += frR * Rr.s[j].d[PO + q];
                    }
                    double ddm, ddc, ddp;
                    if constexpr (Fn::c_unit) { ddm = nm * id; ddc = nc * id; ddp = np * id; }
                    else {
                        ddm = hc ? (nm - du * (frL * Lr.c[j].d[q])) * id : nm;
                        ddc = hc ? (nc - du * (frR * Rr.c[j].d[q] + frL * Lr.c[j].d[PO + q])) * id : nc;
                        ddp = hc ? (np - du * (frR * Rr.c[j].d[PO + q])) * id : np;
                    }
                    row.A.a[j * P + q] = -ddm;
                    row.Bm.a[j * P + q] = ((j == q) ? mI[j] : 0.0) - ddc;
                    row.C.a[j * P + q] = -ddp;
                }
            } else {
                row.d.a[j] = du;
            }
        }
        if (r == 0 && i0 == 0) {  // left boundary node: only thread 0 (of tile 0), only this unrolled copy
            double b0[P];
#pragma unroll
            for (int p = 0; p < P; ++p) b0[p] = JAC ? bc[0][p] : 0.0;
            boundary_row<Fn, JAC, ND, (GLOB || (Fn::c_unit && Fn::s_zero)), XW>(pd, mesh, u0p, uNp, mass, 0, tt, inv_tau, prm, Rr, uc[0], b0, slot, row,
                                                                                frR);
        }
        if (i >= Nx - 1) {  // right boundary node (prepared above) or padding (identity row)
            if (i == Nx - 1) row = rowN;
            else {
                row.A = mat_zero<P>(); row.C = mat_zero<P>(); row.d = vec_zero<P>();
                row.Bm = JAC ? mat_identity<P>() : mat_zero<P>();
            }
        }
        if constexpr (JAC) sink.row(r, row.A, row.Bm, row.C, row.d);
        else sink.value(r, row.d);
        Lr = Rr;
        if constexpr (XW) { x_prev = x_cur; x_cur = x_next; }
    }
}

// ---------------------------------------------------------------------------------------------
// thread/system bookkeeping
// ---------------------------------------------------------------------------------------------
template <int TT>
SB_D bool locate(const ProbDev& pd, int64_t& k, int& t) {
    if constexpr (TT == 32) {
        k = (int64_t)blockIdx.x * (kWarpBlockThreads / 32) + (threadIdx.x >> 5);
        t = threadIdx.x & 31;
    } else {
        k = blockIdx.x;
        t = threadIdx.x;
    }
    return k < pd.B;
}

template <int P, int R, int TT>
SB_D void load_chunk(const double* base, int t, double (&v)[R][P]) {
    if constexpr (P == 2) {
        const double2* b2 = reinterpret_cast<const double2*>(base) + t;
#pragma unroll
        for (int r = 0; r < R; ++r) { const double2 w = b2[r * TT]; v[r][0] = w.x; v[r][1] = w.y; }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int p = 0; p < P; ++p) v[r][p] = base[(size_t)(r * TT + t) * P + p];
    }
}

// bookkeeping of one instance after its Newton update (thread t == 0 of the system; src/utils.jl:323-333).
// No global loads: the iteration number is the same for every active instance of a step and comes from the
// host; the total counter is a fire-and-forget reduction.  Returns the position in the compacted list of
// still-active instances (-1 when the instance converged); the caller stores list_out[pos] = k later, so
// the latency of the returning atomic is not exposed.
SB_D int finalize_system(const ProbDev& pd, int64_t k, double ss, double tol, int slot, int it) {
    const double nm = ::sqrt(ss);
    pd.norm[k] = nm;
    if (pd.hist && it < pd.hist_cap) pd.hist[(size_t)k * pd.hist_cap + it] = nm;
    pd.iters[k] = it + 1;
    atomicAdd((unsigned long long*)(pd.total_iters + k), 1ull);
    if (!(nm == nm) || nm > 1.0e300) pd.status[k] = 2;
    if (nm < tol) { pd.active[k] = 0; return -1; }
    return (int)atomicAdd(pd.active_count + slot, 1u);
}

// update u, ||h||_2, iteration counter and stop test (src/utils.jl:321-333)
template <int P, int R, int TT>
SB_D void update_and_test(const ProbDev& pd, int64_t k, int t, const double (&uc)[R][P], const Vec<P> (&h)[R],
                          double tol, int slot, double* sm, int* list_out = nullptr) {
    double* ub = pd.u + (size_t)k * pd.L * P;
    double ss = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (t * R + r < pd.Nx) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                ub[(size_t)(r * TT + t) * P + p] = uc[r][p] - h[r].a[p];
                ss += h[r].a[p] * h[r].a[p];
            }
        }
    }
    ss = system_sum<TT>(ss, sm + ChunkThomas<P, R>::template exch_doubles<TT>(), t);
    if (t == 0) {
        const int pos = finalize_system(pd, k, ss, tol, slot, pd.iter_no);
        if (pos >= 0 && list_out) list_out[pos] = (int)k;
        group_done(pd, slot, gridDim.x * (TT == 32 ? kWarpBlockThreads / 32 : 1));
    }
}

// ---------------------------------------------------------------------------------------------
// Pivoted fallback (the reference's linear solve, src/utils.jl:317-319) for instances whose pivot-growth monitor
// tripped in the parallel elimination: the thread group writes the block-tridiagonal system into a band array in
// global memory, thread 0 factorises and solves it with row interchanges (sb_banded.cuh, one thread: slow, rare,
// right), every thread picks up the unknowns of its rows.  status[k] = 3 records that the fallback ran (2: the matrix
// is singular even with pivoting).
// ---------------------------------------------------------------------------------------------

SB_D int piv_acquire(const ProbDev& pd, int start) {
    int s = start % pd.piv_slots;
    while (atomicCAS(pd.piv_lock + s, 0, 1) != 0) {     // holders never wait for anybody, so this ends
        s = (s + 1 == pd.piv_slots) ? 0 : s + 1;
        __nanosleep(200);
    }
    __threadfence();
    return s;
}
SB_D void piv_release(const ProbDev& pd, int s) {
    __threadfence();
    atomicExch(pd.piv_lock + s, 0);
}

// slot for this thread group (taken by its thread 0, made known to all) and a zeroed band
template <int P, int TT>
SB_D int piv_begin(const ProbDev& pd, int t, int group, volatile int* sh_slot, int bar = 0) {
    int s = 0;
    if (t == 0) s = piv_acquire(pd, group);
    if constexpr (TT == 32) s = __shfl_sync(0xffffffffu, s, 0);
    else {
        if (t == 0) *sh_slot = s;
        group_sync<TT>(bar);
        s = *sh_slot;
    }
    double* ab = pd.piv_band + (size_t)s * pd.piv_stride;
    const int n = pd.Nx * P * BandShape<P>::ldab;
    for (int i = t; i < n; i += TT) ab[i] = 0.0;
    group_sync<TT>(bar);
    return s;
}

template <int P>
struct BandSink {   // row sink of assemble_chunk: block rows into LAPACK band storage, F into the right-hand side
    double* ab; double* rhs; int i0, Nx;
    SB_D void row(int r, const Mat<P>& A, const Mat<P>& Bm, const Mat<P>& C, const Vec<P>& d) {
        const int i = i0 + r;
        if (i >= Nx) return;                               // padding rows are not part of the system
#pragma unroll
        for (int rr = 0; rr < P; ++rr) {
#pragma unroll
            for (int c = 0; c < P; ++c) {
                if (i > 0) ab[band_index<P>(i, rr, -1, c)] = A.a[rr * P + c];
                ab[band_index<P>(i, rr, 0, c)] = Bm.a[rr * P + c];
                if (i + 1 < Nx) ab[band_index<P>(i, rr, 1, c)] = C.a[rr * P + c];
            }
            rhs[i * P + rr] = d.a[rr];
        }
    }
};

// band complete -> solve (thread 0) -> h of this thread's rows; releases the slot
template <int P, int R, int TT>
SB_D bool piv_solve(const ProbDev& pd, int64_t k, int t, int s, double (&h)[R][P], int bar = 0) {
    double* ab = pd.piv_band + (size_t)s * pd.piv_stride;
    double* rhs = ab + (size_t)pd.Nx * P * BandShape<P>::ldab;
    group_sync<TT>(bar);
    if (t == 0) {
        const int info = banded_lu_solve<BandShape<P>::kl, BandShape<P>::ku, volatile double*>(ab, pd.Nx * P, rhs);
        if (info) rhs[0] = __longlong_as_double(0x7ff8000000000000LL);   // singular even with pivoting: a NaN everybody sees
        if (info || pd.status[k] != 2) pd.status[k] = info ? 2 : 3;
    }
    group_sync<TT>(bar);
    const bool singular = !(__ldcg(rhs) == __ldcg(rhs));
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int i = t * R + r;
            h[r][p] = (i < pd.Nx) ? __ldcg(rhs + (size_t)i * P + p) : 0.0;
        }
    group_sync<TT>(bar);
    if (t == 0) piv_release(pd, s);
    return singular;
}

// the whole fallback for the fused kernels: assemble the rows again (from global memory: the staged copies may have
// been refilled already), this time into the band.  Not inlined: a cold path with its own register allocation.
template <class Fn, int R, int TT, bool SYM0, bool GLOB>
__device__ __noinline__ bool pivoted_newton_rows(ProbDev pd, MeshDev mesh, long long k, const double* ub, const double* bb,
                                                 int t, int group, double tt, double inv_tau, volatile int* sh_slot,
                                                 double (&h)[R][Fn::npde], int bar = 0) {
    constexpr int P = Fn::npde;
    const int s = piv_begin<P, TT>(pd, t, group, sh_slot, bar);
    double uc[R][P], bc[R][P];
    load_chunk<P, R, TT>(ub, t, uc);
    load_chunk<P, R, TT>(bb, t, bc);
    double* ab = pd.piv_band + (size_t)s * pd.piv_stride;
    BandSink<P> sink{ab, ab + (size_t)pd.Nx * P * BandShape<P>::ldab, t * R, pd.Nx};
    assemble_chunk<Fn, R, TT, SYM0, true, GLOB>(pd, mesh, ub, bb, pd.mass + (size_t)k * 4 * P, t, tt, inv_tau,
                                               pd.prm + (size_t)k * pd.nprm_s, uc, bc, sink);
    return piv_solve<P, R, TT>(pd, k, t, s, h, bar);
}

// the fallback for the split design: the rows are already in HBM (Jl, Jd, Ju, F)
template <int P, int R, int TT>
__device__ __noinline__ bool pivoted_stored_rows(ProbDev pd, long long k, int t, volatile int* sh_slot, double (&h)[R][P]) {
    const int s = piv_begin<P, TT>(pd, t, (int)k, sh_slot);
    double* ab = pd.piv_band + (size_t)s * pd.piv_stride;
    BandSink<P> sink{ab, ab + (size_t)pd.Nx * P * BandShape<P>::ldab, t * R, pd.Nx};
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const size_t o = (size_t)k * pd.L + (size_t)r * TT + t;
        Mat<P> A, Bm, C; Vec<P> d;
#pragma unroll
        for (int i = 0; i < P * P; ++i) { A.a[i] = pd.Jl[o * P * P + i]; Bm.a[i] = pd.Jd[o * P * P + i]; C.a[i] = pd.Ju[o * P * P + i]; }
#pragma unroll
        for (int i = 0; i < P; ++i) d.a[i] = pd.F[o * P + i];
        sink.row(r, A, Bm, C, d);
    }
    return piv_solve<P, R, TT>(pd, k, t, s, h);
}

// ---------------------------------------------------------------------------------------------
// K3: residual + Jacobian to HBM (JAC) / assemble! value into F (!JAC)
// ---------------------------------------------------------------------------------------------
template <int P, int TT>
struct StoreSink {
    const ProbDev& pd; int64_t k; int t;
    SB_D void row(int r, const Mat<P>& A, const Mat<P>& Bm, const Mat<P>& C, const Vec<P>& d) {
        const size_t o = (size_t)k * pd.L + (size_t)r * TT + t;
#pragma unroll
        for (int i = 0; i < P * P; ++i) { pd.Jl[o * P * P + i] = A.a[i]; pd.Jd[o * P * P + i] = Bm.a[i]; pd.Ju[o * P * P + i] = C.a[i]; }
#pragma unroll
        for (int i = 0; i < P; ++i) pd.F[o * P + i] = d.a[i];
    }
    SB_D void value(int r, const Vec<P>& d) {
        const size_t o = (size_t)k * pd.L + (size_t)r * TT + t;
#pragma unroll
        for (int i = 0; i < P; ++i) pd.F[o * P + i] = d.a[i];
    }
};

#ifndef SB_MIN_BLOCKS
#define SB_MIN_BLOCKS 1
#endif
#define SB_LAUNCH_BOUNDS(TT) __launch_bounds__((TT) == 32 ? kWarpBlockThreads : (TT), ((TT) <= 128 ? SB_MIN_BLOCKS : 1))

template <class Fn, int R, int TT, bool SYM0, bool JAC>
__global__ void SB_LAUNCH_BOUNDS(TT) k_assemble(ProbDev pd, double tt, double inv_tau, int all_instances) {
    constexpr int P = Fn::npde;
    int64_t k; int t;
    if (!locate<TT>(pd, k, t)) return;
    if (!all_instances && !pd.active[k]) return;
    double uc[R][P], bc[R][P];
    load_chunk<P, R, TT>(pd.u + (size_t)k * pd.L * P, t, uc);
    if constexpr (JAC) load_chunk<P, R, TT>(pd.b + (size_t)k * pd.L * P, t, bc);
    StoreSink<P, TT> sink{pd, k, t};
    assemble_chunk<Fn, R, TT, SYM0, JAC>(pd, pd.mesh, pd.u + (size_t)k * pd.L * P, pd.b + (size_t)k * pd.L * P,
                                         pd.mass + (size_t)k * 4 * P, t, tt, inv_tau, pd.prm + (size_t)k * pd.nprm_s, uc, bc, sink);
}

// ---------------------------------------------------------------------------------------------
// K4/K5 + K7: solve J h = F from HBM, update, norm, stop test
// ---------------------------------------------------------------------------------------------
template <int P, int R, int TT>
__global__ void SB_LAUNCH_BOUNDS(TT) k_solve(ProbDev pd, double tol, int slot, int update) {
    extern __shared__ double sm[];
    int64_t k; int t;
    const bool in_range = locate<TT>(pd, k, t);
    if (!in_range || (update && !pd.active[k])) {
        if (update && t == 0) group_done(pd, slot, gridDim.x * (TT == 32 ? kWarpBlockThreads / 32 : 1));
        return;
    }
    ChunkThomas<P, R> th;
    double uc[R][P];
    if constexpr (P == 1 && R <= 8) {
        // scalar systems: issue all loads of the chunk (rows and the iterate) before the dependent elimination
        // chain starts, so the memory latency is paid once per thread instead of once per row
        double ra[R], rb[R], rc[R], rd[R];
        const size_t o0 = (size_t)k * pd.L + t;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            ra[r] = __ldcs(pd.Jl + o0 + (size_t)r * TT); rb[r] = __ldcs(pd.Jd + o0 + (size_t)r * TT);
            rc[r] = __ldcs(pd.Ju + o0 + (size_t)r * TT); rd[r] = __ldcs(pd.F + o0 + (size_t)r * TT);
        }
        if (update) load_chunk<P, R, TT>(pd.u + (size_t)k * pd.L * P, t, uc);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            Mat<P> A, Bm, C; Vec<P> d;
            A.a[0] = ra[r]; Bm.a[0] = rb[r]; C.a[0] = rc[r]; d.a[0] = rd[r];
            th.row(r, A, Bm, C, d);
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const size_t o = (size_t)k * pd.L + (size_t)r * TT + t;
            Mat<P> A, Bm, C; Vec<P> d;
#pragma unroll
            for (int i = 0; i < P * P; ++i) { A.a[i] = pd.Jl[o * P * P + i]; Bm.a[i] = pd.Jd[o * P * P + i]; C.a[i] = pd.Ju[o * P * P + i]; }
#pragma unroll
            for (int i = 0; i < P; ++i) d.a[i] = pd.F[o * P + i];
            th.row(r, A, Bm, C, d);
        }
        if (update) load_chunk<P, R, TT>(pd.u + (size_t)k * pd.L * P, t, uc);
    }
    Vec<P> h[R];
    if (th.template finish<TT>(h, sm, t)) {   // pivot growth: the same system again, with row interchanges
        double hp[R][P];
        pivoted_stored_rows<P, R, TT>(pd, k, t, reinterpret_cast<volatile int*>(sm + ChunkThomas<P, R>::template smem_doubles<TT>()), hp);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int p = 0; p < P; ++p) h[r].a[p] = hp[r][p];
    }
    if (update) {
        update_and_test<P, R, TT>(pd, k, t, uc, h, tol, slot, sm);
    } else {  // debug: leave x in F
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < P; ++i) pd.F[((size_t)k * pd.L + (size_t)r * TT + t) * P + i] = h[r].a[i];
    }
}

// ---------------------------------------------------------------------------------------------
// fused Newton iteration: assemble rows straight into the elimination
// ---------------------------------------------------------------------------------------------
template <class Fn, int R, int TT, bool SYM0>
__global__ void SB_LAUNCH_BOUNDS(TT) k_newton_fused(ProbDev pd, double tt, double inv_tau, double tol, int slot) {
    constexpr int P = Fn::npde;
    extern __shared__ double sm[];
    int64_t k; int t;
    const bool in_range = locate<TT>(pd, k, t);
    if (!in_range || !pd.active[k]) {
        if (t == 0) group_done(pd, slot, gridDim.x * (TT == 32 ? kWarpBlockThreads / 32 : 1));
        return;
    }
    double uc[R][P], bc[R][P];
    load_chunk<P, R, TT>(pd.u + (size_t)k * pd.L * P, t, uc);
    load_chunk<P, R, TT>(pd.b + (size_t)k * pd.L * P, t, bc);
    ChunkThomas<P, R> th;
    assemble_chunk<Fn, R, TT, SYM0, true>(pd, pd.mesh, pd.u + (size_t)k * pd.L * P, pd.b + (size_t)k * pd.L * P,
                                          pd.mass + (size_t)k * 4 * P, t, tt, inv_tau, pd.prm + (size_t)k * pd.nprm_s, uc, bc, th);
    Vec<P> h[R];
    if (th.template finish<TT>(h, sm, t)) {   // pivot growth: assemble again into a band, solve with row interchanges
        double hp[R][P];
        pivoted_newton_rows<Fn, R, TT, SYM0, true>(pd, pd.mesh, k, pd.u + (size_t)k * pd.L * P, pd.b + (size_t)k * pd.L * P, t, (int)k, tt,
                                                   inv_tau, reinterpret_cast<volatile int*>(sm + ChunkThomas<P, R>::template smem_doubles<TT>()), hp);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int p = 0; p < P; ++p) h[r].a[p] = hp[r][p];
    }
    update_and_test<P, R, TT>(pd, k, t, uc, h, tol, slot, sm);
}

// ---------------------------------------------------------------------------------------------
// Persistent fused Newton iteration with TMA prefetch.
// One CTA per SM slot loops over the compacted list of still-active instances.  While instance j is
// being assembled/eliminated out of shared memory, the bulk-copy engine (cp.async.bulk, completion on
// an mbarrier) already streams u and b of instance j+stride into the other stage, so DRAM latency is off
// the critical path; the mesh weights are copied to shared memory once per CTA.  TT == 32: every warp
// of the CTA runs its own pipeline (own stages, own mbarriers, warp-level syncs only).
// ---------------------------------------------------------------------------------------------
SB_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Programmatic dependent launch: a kernel launched with programmatic stream serialization may start while its
// predecessor in the stream is still draining; pdl_wait() returns once the predecessor has completed and its
// writes are visible, pdl_release() lets the successor's CTAs take SM slots as they free up.  Both are no-ops in
// a kernel launched the ordinary way.
SB_D void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
SB_D void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

SB_D void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SB_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SB_D void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
SB_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

#ifndef SB_TMA_STAGES
#define SB_TMA_STAGES 1   // 1: the next instance is prefetched into the same buffers as soon as every thread of the
#endif                    //    group has consumed the current one (less smem -> more CTAs/SM); 2: double buffering
#ifndef SB_PERSISTENT_WARPS_PER_SM
#define SB_PERSISTENT_WARPS_PER_SM 16
#endif
#ifndef SB_PERSISTENT_WARPS_P2   // npde = 2: twice the registers per thread
#define SB_PERSISTENT_WARPS_P2 (SB_PERSISTENT_WARPS_PER_SM / 2)
#endif

template <class Fn, int R, int TT, bool SYM0>
struct FusedPLayout {  // shared-memory carve-up (offsets in doubles)
    static constexpr int P = Fn::npde, L = R * TT, G = (TT == 32) ? kWarpBlockThreads / 32 : 1, S = SB_TMA_STAGES;
    static constexpr int NPS = (Fn::nparams + 1) / 2 * 2;  // parameter block stride (even)
    static constexpr bool kFrac = !(Fn::c_unit && Fn::s_zero);
    static constexpr int nGw = SYM0 ? L : 0, nAB = SYM0 ? 0 : 4 * L, nIvol = Fn::c_unit ? L : 0, nFrac = kFrac ? 2 * L : 0;
    static constexpr int oGw = 0, oAB = oGw + nGw, oIvol = oAB + nAB, oFrac = oIvol + nIvol, mesh_doubles = oFrac + nFrac;
    static constexpr int exch = ChunkThomas<P, R>::template exch_doubles<TT>(), nexch = ChunkThomas<P, R>::exch_buffers;
    static constexpr int scratch = nexch * exch + (nexch * exch) % 2;   // exchange region(s) of finish(), padded to 16 bytes
    // per group: u[S][L*P], b[S][L*P], prm[S][NPS], mass[S][4P], scratch, partial sums [2][8], mass flags kept [4P], instance index [2]
    static constexpr int oU = 0, oB = oU + S * L * P, oPrm = oB + S * L * P, oMass = oPrm + S * NPS, oScr = oMass + S * 4 * P,
                         oSum = oScr + scratch, oKeep = oSum + 16, oK = oKeep + 4 * P, group_doubles = oK + 4;   // oK: instance index [2], fallback slot [2 ints], pad
    static_assert(group_doubles % 2 == 0 && oB % 2 == 0 && oPrm % 2 == 0 && oMass % 2 == 0, "bulk copies need 16-byte aligned destinations");
    static constexpr int bar_doubles = 2 * G + 2;  // mbarriers (8 bytes): S per group (2 reserved) + one for the mesh weights
    static constexpr size_t bytes = sizeof(double) * (size_t)(bar_doubles + mesh_doubles + G * group_doubles);
};

// resident warps per SM the persistent kernel is compiled for (register cap): scalar PDEs keep ~100 registers per
// thread; 2x2 block arithmetic needs about twice that, so it is given half the warps instead of spilling
__host__ __device__ constexpr int persistent_min_blocks(int TT, int P) {
    const int threads = (TT == 32) ? kWarpBlockThreads : TT;
    const int warps = (P == 1) ? SB_PERSISTENT_WARPS_PER_SM : SB_PERSISTENT_WARPS_P2;   // P > 2: as P = 2 (255 registers)
    return (32 * warps) / threads > 0 ? (32 * warps) / threads : 1;
}
template <class Fn, int R, int TT, bool SYM0>
__global__ void __launch_bounds__((TT) == 32 ? kWarpBlockThreads : (TT), persistent_min_blocks(TT, Fn::npde))
k_newton_fused_p(ProbDev pd, double tt, double inv_tau, double tol, int slot,
                                                          int prev_slot, int parity) {
    using LY = FusedPLayout<Fn, R, TT, SYM0>;
    constexpr int P = LY::P, L = LY::L, G = LY::G, NPS = LY::NPS, S = LY::S;
    extern __shared__ __align__(16) double smp[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smp);
    double* mesh_sm = smp + LY::bar_doubles;
    const int g = (TT == 32) ? (threadIdx.x >> 5) : 0;
    const int t = (TT == 32) ? (threadIdx.x & 31) : threadIdx.x;
    double* grp = mesh_sm + LY::mesh_doubles + (size_t)g * LY::group_doubles;
    double* ubuf = grp + LY::oU;      // [S][L*P]
    double* bbuf = grp + LY::oB;      // [S][L*P]
    double* pbuf = grp + LY::oPrm;    // [S][NPS]
    double* mbuf = grp + LY::oMass;   // [S][4P]
    double* scratch = grp + LY::oScr;
    double* sums = grp + LY::oSum;    // [2][8] per-warp sums of h^2 of the instance in flight / the one before
    double* keep = grp + LY::oKeep;   // [4P] TT == 32: mass flags of the instance in flight
    long long* kslot = reinterpret_cast<long long*>(grp + LY::oK);  // [2] instance index per stage / parity
    uint64_t* bar = bars + 2 * g;

    // Prologue that does not depend on the previous Newton-iteration launch (it overlaps that launch's tail when
    // this one was started with programmatic stream serialization): barriers, mesh weights -> shared memory, once
    // per CTA, by the bulk-copy engine (waited for right before first use).
    pdl_release();
    MeshDev mesh = pd.mesh;
    uint64_t* mesh_bar = bars + 2 * G;
    if constexpr (SYM0) mesh.gw = mesh_sm + LY::oGw;
    else {
        mesh.awbw = reinterpret_cast<const double2*>(mesh_sm + LY::oAB);
        mesh.gwwf = reinterpret_cast<const double2*>(mesh_sm + LY::oAB + 2 * L);
    }
    if constexpr (Fn::c_unit) mesh.ivol = mesh_sm + LY::oIvol;
    if constexpr (LY::kFrac) mesh.frac = reinterpret_cast<const double2*>(mesh_sm + LY::oFrac);
    if (threadIdx.x == 0) mbar_init(mesh_bar, 1);
    if (t == 0) {
        mbar_init(bar + 0, 1);
        mbar_init(bar + 1, 1);
    }
    if (threadIdx.x == 0 || t == 0) {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        constexpr uint32_t LB = (uint32_t)(L * sizeof(double));
        mbar_expect_tx(mesh_bar, (uint32_t)(LY::mesh_doubles * sizeof(double)));
        if constexpr (SYM0) bulk_g2s(mesh_sm + LY::oGw, pd.mesh.gw, LB, mesh_bar);
        else bulk_g2s(mesh_sm + LY::oAB, pd.mesh.awbw, 4 * LB, mesh_bar);  // awbw[2L] and gwwf[2L] are contiguous
        if constexpr (Fn::c_unit) bulk_g2s(mesh_sm + LY::oIvol, pd.mesh.ivol, LB, mesh_bar);
        if constexpr (LY::kFrac) bulk_g2s(mesh_sm + LY::oFrac, pd.mesh.frac, 2 * LB, mesh_bar);
    }

    pdl_wait();  // everything below reads what the previous launch wrote (active count, list, u, b)
    int n_in = (prev_slot < 0) ? (int)pd.B : (int)pd.active_count[prev_slot];
    int it_no = pd.iter_no;                       // Newton iteration number inside the step
    // first iteration of a step that skipped k_begin_step: the Newton iterate is b = M.*u_n itself (src/utils.jl:309)
    bool from_b = pd.lazy_b && it_no == 0;
    int auto_step = -1;
    if (pd.auto_on) {  // device-side stepping: step / iteration come from the previous launch, not from the host
        const int s0 = pd.auto_state[0], i0 = pd.auto_state[1];
        const bool advance = (i0 > 0 && n_in == 0);          // the previous launch converged the last instances
        auto_step = s0 + (advance ? 1 : 0);
        it_no = advance ? 0 : i0;
        if (auto_step >= pd.auto_nsteps) {                    // past the last time step: nothing left to do
            if (threadIdx.x == 0) mbar_wait(mesh_bar, 0);
            __syncthreads();
            if (t == 0) group_done(pd, slot, gridDim.x * G, auto_step, it_no);
            return;
        }
        tt = pd.auto_tgrid[auto_step + 1];
        const double tau_s = (pd.auto_tau > 0.0) ? pd.auto_tau : tt - pd.auto_tgrid[auto_step];
        inv_tau = (tau_s > 1.0e300) ? 0.0 : 1.0 / tau_s;
        from_b = (it_no == 0);                                // every instance left b = M.*u_{n+1} when it converged
        if (from_b) n_in = (int)pd.B;
        parity = it_no & 1;
    }
    const int* list_in = (from_b || prev_slot < 0) ? nullptr : pd.list[parity];
    int* list_out = pd.list[parity ^ 1];
    const int stride = gridDim.x * G;
    int j = blockIdx.x * G + g;
    if (blockIdx.x * G >= n_in) {  // whole CTA has nothing to do
        if (threadIdx.x == 0) mbar_wait(mesh_bar, 0);  // do not exit with the bulk copy in flight
        __syncthreads();
        if (t == 0) group_done(pd, slot, gridDim.x * G, auto_step, it_no);
        return;
    }

    constexpr uint32_t kBytes = (uint32_t)(L * P * sizeof(double));
    constexpr uint32_t kPrmBytes = (uint32_t)(NPS * sizeof(double)), kMassBytes = (uint32_t)(4 * P * sizeof(double));
    // thread t == 0 of the group is the producer: it knows the instance indices two steps ahead
    long long k_next = -1, k_next2 = -1;
    if (t == 0) {
        if (j < n_in) k_next = list_in ? list_in[j] : j;
        if (j + stride < n_in) k_next2 = list_in ? list_in[j + stride] : j + stride;
    }
    // fetch instance kk into buffer `stage`; its index is published in kslot[seq & 1] (seq = position in this group's sequence)
    auto issue = [&](long long kk, int stage, int seq) {
        kslot[seq & 1] = kk;
        mbar_expect_tx(bar + stage, (from_b ? 1 : 2) * kBytes + kPrmBytes + kMassBytes);
        if (!from_b) bulk_g2s(ubuf + (size_t)stage * L * P, pd.u + (size_t)kk * L * P, kBytes, bar + stage);
        bulk_g2s(bbuf + (size_t)stage * L * P, pd.b + (size_t)kk * L * P, kBytes, bar + stage);
        bulk_g2s(pbuf + stage * NPS, pd.prm + (size_t)kk * NPS, kPrmBytes, bar + stage);
        bulk_g2s(mbuf + stage * 4 * P, pd.mass + (size_t)kk * 4 * P, kMassBytes, bar + stage);
    };
    if (t == 0 && k_next >= 0) issue(k_next, 0, 0);
    if constexpr (TT == 32) __syncwarp();
    else __syncthreads();  // kslot[0] visible
    mbar_wait(mesh_bar, 0);

    int pend_pos = -1, pend_k = 0;  // deferred list write of the previous instance (thread t == 0)
    // The outcome of an instance (converged or not) needs the sum of h^2 over all warps of its CTA.  Instead of a
    // second CTA barrier per instance, the per-warp sums are parked in shared memory and the instance is closed
    // ("settled") behind the one barrier of the NEXT instance's solve (or a final one after the loop).
    int64_t prev_k = -1;
    // close an instance: un = its new iterate (this thread's rows), ss = ||h||^2, mk = its mass flags
    auto settle = [&](int64_t kk, double ss, const double (&un)[R][P], const double* mk) {
        if (::sqrt(ss) < tol) {
            // converged: this is u_{n+1}.  Write b = M .* u_{n+1} for the next time step right away (src/pdepe.jl:94),
            // which then needs no separate pass over the state.
            double* bg = pd.b + (size_t)kk * L * P;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = t * R + r;
                if (i < pd.Nx) {
                    const int cls = (i == 0) ? 0 : ((i == pd.Nx - 1) ? 2 : 1);
#pragma unroll
                    for (int p = 0; p < P; ++p) bg[(size_t)(r * TT + t) * P + p] = mk[cls * P + p] * un[r][p];
                }
            }
        }
        if (t == 0) {
            if (pend_pos >= 0) list_out[pend_pos] = pend_k;   // the entry of the instance settled before this one
            if (from_b) pd.active[kk] = 1;
            pend_pos = finalize_system(pd, kk, ss, tol, slot, it_no);
            pend_k = (int)kk;
        }
    };
    // TT > 32: call with every warp past its store into sums[it_prev & 1] (i.e. behind a CTA barrier)
    auto settle_prev = [&](int it_prev) {
        const double* sl = sums + (it_prev & 1) * 8;
        double ss = sl[0];                    // every thread forms the same total in the same order
#pragma unroll
        for (int w = 1; w < TT / 32; ++w) ss += sl[w];
        double un[R][P];
        if (::sqrt(ss) < tol) {               // u is read back from L2, where this very thread wrote it a moment ago
            const double* ug = pd.u + (size_t)prev_k * L * P;
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int p = 0; p < P; ++p) un[r][p] = __ldcg(ug + (size_t)(r * TT + t) * P + p);
        }
        settle(prev_k, ss, un, pd.mass + (size_t)prev_k * 4 * P);
    };
    int it = 0;
    for (; j < n_in; j += stride, ++it) {
        const int stage = (S == 2) ? (it & 1) : 0;
        // the producer refills: S == 2 -> the other stage, now (released by the previous iteration's barrier);
        //                       S == 1 -> the same buffers, as soon as the whole group has consumed them (hook below)
        auto prefetch = [&]() {
            if (t == 0) {
                if (k_next2 >= 0) issue(k_next2, (S == 2) ? (stage ^ 1) : 0, it + 1);
                k_next = k_next2;
                const int j3 = j + 2 * stride;
                k_next2 = (j3 < n_in) ? (list_in ? (long long)list_in[j3] : (long long)j3) : -1;
            }
        };
        if constexpr (S == 2) prefetch();
        if (t == 0 && pend_pos >= 0) { list_out[pend_pos] = pend_k; pend_pos = -1; }
        const double* bs = bbuf + (size_t)stage * L * P;
        const double* us = from_b ? bs : ubuf + (size_t)stage * L * P;
        const double* prm = pbuf + stage * NPS;
        const double* mass = mbuf + stage * 4 * P;
        mbar_wait(bar + stage, (S == 2) ? ((it >> 1) & 1) : (it & 1));
        const int64_t k = kslot[it & 1];
        if constexpr (TT == 32) {   // mass flags of the instance in flight (S == 1: the stage is refilled before they are used)
            if (t < 3 * P) keep[t] = mass[t];
        }

        double uc[R][P], bc[R][P];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int p = 0; p < P; ++p) { uc[r][p] = us[(r * TT + t) * P + p]; bc[r][p] = bs[(r * TT + t) * P + p]; }
        ChunkThomas<P, R> th;
        assemble_chunk<Fn, R, TT, SYM0, true, false>(pd, mesh, us, bs, mass, t, tt, inv_tau, prm, uc, bc, th);
        Vec<P> h[R];
        double* exch = scratch + (it % LY::nexch) * LY::exch;
        bool suspect;
        if constexpr (TT == 32) {
            __syncwarp();  // every lane is done with the staged inputs
            if constexpr (S == 2) suspect = th.template finish<TT>(h, exch, t);
            else suspect = th.template finish<TT>(h, exch, t, prefetch);
        } else {
            // behind the barrier inside finish(): every thread has consumed the staged inputs (S == 1: refill them) and
            // every warp has parked its partial sum of the previous instance (settle it)
            auto hook = [&]() {
                if constexpr (S == 1) prefetch();
                if (it > 0) settle_prev(it - 1);
            };
            suspect = th.template finish<TT>(h, exch, t, hook);
        }
        if (suspect) {   // pivot growth (same verdict in every thread of the group): solve again with row interchanges
            const double* bg = pd.b + (size_t)k * L * P;
            double hp[R][P];
            pivoted_newton_rows<Fn, R, TT, SYM0, false>(pd, mesh, k, from_b ? bg : pd.u + (size_t)k * L * P, bg, t, blockIdx.x * G + g, tt,
                                                        inv_tau, reinterpret_cast<volatile int*>(kslot) + 4 + (it & 1), hp);
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int p = 0; p < P; ++p) h[r].a[p] = hp[r][p];
        }

        // update, ||h||^2 of the system
        double* ug = pd.u + (size_t)k * L * P;
        double ss = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int p = 0; p < P; ++p) uc[r][p] -= h[r].a[p];
            if (t * R + r < pd.Nx) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    ug[(size_t)(r * TT + t) * P + p] = uc[r][p];
                    ss += h[r].a[p] * h[r].a[p];
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_down_sync(0xffffffffu, ss, o);
        if constexpr (TT > 32) {
            if ((t & 31) == 0) sums[(it & 1) * 8 + (t >> 5)] = ss;
            prev_k = k;
        } else {
            ss = __shfl_sync(0xffffffffu, ss, 0);
            settle(k, ss, uc, keep);
            __syncwarp();   // keep[] is rewritten by the next instance
        }
    }
    if constexpr (TT > 32) {
        __syncthreads();
        if (it > 0) settle_prev(it - 1);
    }
    if (t == 0) {
        if (pend_pos >= 0) list_out[pend_pos] = pend_k;
        group_done(pd, slot, gridDim.x * G, auto_step, it_no);
    }
}

// ---------------------------------------------------------------------------------------------
// K2: mass flags (mass_matrix, src/assembler.jl:138-174), one thread per instance
// ---------------------------------------------------------------------------------------------
template <class Fn>
__global__ void k_mass_flags(ProbDev pd, double t0, double* mass, int stationary, const double* bnd) {
    constexpr int P = Fn::npde;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pd.B) return;
    double* mk = mass + (size_t)k * 4 * P;
    if (stationary) {
        for (int i = 0; i < 3 * P; ++i) mk[i] = 1.0;
        return;
    }
    const double* ub = pd.u + (size_t)k * pd.L * P;
    const double* prm = pd.prm + (size_t)k * pd.nprm_s;
    const int R = pd.R, T = pd.T, Nx = pd.Nx;
    const int Lt = pd.Lt;
    auto slot_of = [&](int i) { const int tile = i / Lt, il = i - tile * Lt; return tile * Lt + (il % R) * T + il / R; };
    double u0[P], u1[P], uN[P];
    for (int p = 0; p < P; ++p) {
        if (bnd) {   // [B][3][P]: nodes 0, 1 and Nx-1 of every instance, gathered by the host (the state itself is still on its way)
            u0[p] = bnd[((size_t)k * 3 + 0) * P + p]; u1[p] = bnd[((size_t)k * 3 + 1) * P + p]; uN[p] = bnd[((size_t)k * 3 + 2) * P + p];
        } else {
            u0[p] = ub[(size_t)slot_of(0) * P + p];
            u1[p] = ub[(size_t)slot_of(1) * P + p];
            uN[p] = ub[(size_t)slot_of(Nx - 1) * P + p];
        }
    }
    Dual<0> ul_[P], ur_[P], pl[P], pr[P];
    double ql[P], qr[P];
    for (int p = 0; p < P; ++p) { ul_[p] = Dual<0>(u0[p]); ur_[p] = Dual<0>(uN[p]); }
    Fn::template bc<Dual<0>>(pd.x0, ul_, pd.xN, ur_, t0, prm, pl, ql, pr, qr);
    IntervalRes<P, 0> res;
    eval_interval<Fn, 0, false>(pd.mesh, slot_of(0), u0, u1, t0, prm, res);
    for (int i = 0; i < P; ++i) {
        mk[i] = (ql[i] == 0.0 && !pd.singular) ? 0.0 : 1.0;   // :162-165
        mk[P + i] = (res.c[i].v == 0.0) ? 0.0 : 1.0;           // :157-160
        mk[2 * P + i] = (qr[i] == 0.0) ? 0.0 : 1.0;            // :167-170
    }
}

// b = M .* u_n ; u <- b ; active, counters (src/pdepe.jl:94, src/utils.jl:309)
template <int UNUSED = 0>   // (a template only so that the header can be included from several translation units)
__global__ void k_begin_step(ProbDev pd, int P) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*L node slots
    const size_t total = (size_t)pd.B * pd.L;
    if (idx >= total) return;
    const int64_t k = idx / pd.L;
    const int s = (int)(idx - (size_t)k * pd.L);
    const int tile = s / pd.Lt, sl = s - tile * pd.Lt;
    const int r = sl / pd.T, t = sl - r * pd.T;
    const int i = tile * pd.Lt + t * pd.R + r;
    if (s == 0) { pd.active[k] = 1; pd.iters[k] = 0; }
    const double* mass = pd.mass + (size_t)k * 4 * P;
    const int cls = (i == 0) ? 0 : ((i == pd.Nx - 1) ? 2 : 1);
    for (int p = 0; p < P; ++p) {
        const double mv = mass[cls * P + p];
        const double v = (i < pd.Nx) ? mv * pd.u[idx * P + p] : 0.0;
        pd.b[idx * P + p] = v;
        if (mv == 0.0 || i >= pd.Nx) pd.u[idx * P + p] = v;  // u <- b differs from u only where M = 0
    }
}

}  // namespace sb
