// sb_resident.cuh -- the whole batched solve in ONE launch: every thread group keeps its PDE instance on the SM
// through all Newton iterations of all time steps.
//
// The instances of a batch are independent (src/pdepe.jl:90-108 and src/utils.jl:311-336 run per instance in the
// reference), so nothing forces them through the Newton loop in lock step.  Here a thread group (one warp for Nx <= 256,
// else one CTA) loads the state of an instance once, then runs
//     for every time step:  b = M .* u_n ; u <- b                                  (src/pdepe.jl:94, src/utils.jl:309)
//         repeat:  assemble F, J by duals -> eliminate -> u -= h -> ||h||_2 < tol ?   (src/utils.jl:312-333)
// with u in registers and a shared-memory copy for the neighbours' end values, b in shared memory, and writes
// u(t_end) back.  Per Newton iteration nothing crosses HBM; what used to be a launch per iteration (launch ramp and
// tail, lock-step waiting for the slowest instance, compacted lists, prefetch pipelines) is gone.  The arithmetic per
// iteration is the same code as k_newton_fused_p (assemble_chunk, ChunkThomas, pivoted fallback), so states and
// iteration counts are bit-identical to the host-owned loops.
//
// Per-instance extras that are natural here: per-instance time grids (vector tstep per instance), the iteration count
// of every instance in every step, and every intermediate state (snapshots, src/pdepe.jl:99-107) written to HBM in
// the reference layout, round by round, so that the host can stream finished rounds out while later rounds compute.
#pragma once
#include "sb_kernels.cuh"

namespace sb {

struct ResidentArgs {
    int nsteps, maxit;
    int b_given;                    // first step: b (and the iterate) come from pd.b as sb_begin_step left it
    double tol;
    double tau;                     // > 0: constant time step (src/pdepe.jl:92 scalar tstep); <= 0: diff of the grid
    const double* tgrid;            // [nsteps + 1] grid shared by the batch (null with t_single)
    const double* inst_tgrid;       // [B][nsteps + 1] per-instance grids, or null
    double t_single, inv_tau_single;  // nsteps == 1 without a grid: one Newton solve at (t, 1/tau)  (sb_newton_solve)
    int* step_iters;                // [B][nsteps] Newton iterations of every instance in every step, or null
    double* snap;                   // [nsteps + 1][B][Nx*P] every state, reference layout, or null
    int snap_last;                  // snap holds ONE state per instance, [B][Nx*P]: u(t_end) only (streamed out like the snapshots)
    unsigned int* round_done;       // [rounds] instances finished per round of the grid (snapshot streaming), or null
    volatile unsigned int* host_rounds;   // [rounds] mapped host flags: round complete
    int* fail_step;                 // [1] smallest step index in which an instance failed to converge (init INT_MAX)
    int* max_iters;                 // [1] largest iteration count of a step (sb_newton_solve reports it)
    unsigned long long* sum_iters;  // [1] Newton iterations summed over instances and steps (the work counter)
};

// Systems per CTA of the whole-solve kernel: four warps (TT == 32); two 64-thread groups with a named barrier each
// (TT == 64: the CTA's copy of the mesh weights is the larger part of its shared memory, sharing it between two
// systems lifts the resident warps per SM from 11 to 16 on cfg3); one otherwise.
__host__ __device__ constexpr int resident_groups(int TT) { return TT == 32 ? kWarpBlockThreads / 32 : (TT == 64 ? 2 : 1); }
__host__ __device__ constexpr int resident_min_blocks(int TT, int P) {
    const int warps = (P == 1) ? SB_PERSISTENT_WARPS_PER_SM : SB_PERSISTENT_WARPS_P2;
    const int threads = resident_groups(TT) * TT;
    return (32 * warps) / threads > 0 ? (32 * warps) / threads : 1;
}

// pipelined upload (sb_solve_from_host): the initial states arrive in u_ref, reference layout [B][Nx*P], chunk by chunk of
// `chunk` instances while the kernel already runs; *ready >= ready_base + c + 1 once chunk c has landed.  A kernel
// argument of its own (and a kernel of its own, k_solve_resident_pipe): adding these fields and the branch that uses them
// to the headline kernel cost it 1.7 % (12.40 -> 12.61 ms per cfg2 solve) although the code runs once per instance --
// the register allocation of everything after it changed.
struct ResidentPipe {
    const double* u_ref;
    const volatile unsigned int* ready;
    unsigned int ready_base;
    int chunk;
};

template <class Fn, int R, int TT, bool SYM0>
struct ResidentLayout {  // shared-memory carve-up (offsets in doubles)
    static constexpr int P = Fn::npde, L = R * TT, G = resident_groups(TT);
    static constexpr int NPS = (Fn::nparams + 1) / 2 * 2;
    static constexpr bool kFrac = !(Fn::c_unit && Fn::s_zero);
    static constexpr int nGw = SYM0 ? L : 0, nAB = SYM0 ? 0 : 4 * L, nIvol = Fn::c_unit ? L : 0, nFrac = kFrac ? 2 * L : 0;
    static constexpr int oGw = 0, oAB = oGw + nGw, oIvol = oAB + nAB, oFrac = oIvol + nIvol, mesh_doubles = oFrac + nFrac;
    static constexpr int exch = ChunkThomas<P, R>::template exch_doubles<TT>(), nexch = ChunkThomas<P, R>::exch_buffers;
    static constexpr int scr = nexch * exch + (nexch * exch) % 2;
    // per group: u[L*P] (+2 so that slot -1 and slot L exist), b[L*P], prm[NPS], mass[4P], exchange region(s), sums [2][8], slot word
    static constexpr int oU = 2, oB = oU + L * P + 2, oPrm = oB + L * P, oMass = oPrm + NPS, oScr = oMass + 4 * P,
                         oSum = oScr + scr, oSlot = oSum + 16, group_doubles = oSlot + 2;
    static_assert(group_doubles % 2 == 0 && mesh_doubles % 2 == 0, "double2 accesses need 16-byte aligned groups");
    static constexpr size_t bytes = sizeof(double) * (size_t)(mesh_doubles + G * group_doubles);
};

template <class Fn, int R, int TT, bool SYM0, bool PIPE>
SB_D void resident_body(const ProbDev& pd, const ResidentArgs& ra, const ResidentPipe& rp) {
    using LY = ResidentLayout<Fn, R, TT, SYM0>;
    constexpr int P = LY::P, L = LY::L, G = LY::G, NPS = LY::NPS;
    extern __shared__ __align__(16) double smr[];
    double* mesh_sm = smr;
    const int g = (G > 1) ? (int)threadIdx.x / TT : 0;
    const int t = (G > 1) ? (int)threadIdx.x % TT : (int)threadIdx.x;
    const int bar = (G > 1 && TT > 32) ? 1 + g : 0;   // barrier of this system's threads (group_sync)
    double* grp = mesh_sm + LY::mesh_doubles + (size_t)g * LY::group_doubles;
    double* utile = grp + LY::oU;     // [L*P] current iterate, device-layout slots (r*TT + t)
    double* btile = grp + LY::oB;     // [L*P] b = M .* u_n of the step
    double* pbuf = grp + LY::oPrm;    // functor parameters of the instance
    double* mbuf = grp + LY::oMass;   // its mass flags [4P]
    double* scratch = grp + LY::oScr;
    double* sums = grp + LY::oSum;
    volatile int* sh_slot = reinterpret_cast<volatile int*>(grp + LY::oSlot);

    // mesh weights -> shared memory, once per CTA
    MeshDev mesh = pd.mesh;
    {
        const int nthr = G * TT;
        if constexpr (SYM0) {
            for (int i = threadIdx.x; i < L; i += nthr) mesh_sm[LY::oGw + i] = pd.mesh.gw[i];
            mesh.gw = mesh_sm + LY::oGw;
        } else {
            const double* src = reinterpret_cast<const double*>(pd.mesh.awbw);   // awbw[2L] and gwwf[2L] are contiguous
            for (int i = threadIdx.x; i < 4 * L; i += nthr) mesh_sm[LY::oAB + i] = src[i];
            mesh.awbw = reinterpret_cast<const double2*>(mesh_sm + LY::oAB);
            mesh.gwwf = reinterpret_cast<const double2*>(mesh_sm + LY::oAB + 2 * L);
        }
        if constexpr (Fn::c_unit) {
            for (int i = threadIdx.x; i < L; i += nthr) mesh_sm[LY::oIvol + i] = pd.mesh.ivol[i];
            mesh.ivol = mesh_sm + LY::oIvol;
        }
        if constexpr (LY::kFrac) {
            const double* src = reinterpret_cast<const double*>(pd.mesh.frac);
            for (int i = threadIdx.x; i < 2 * L; i += nthr) mesh_sm[LY::oFrac + i] = src[i];
            mesh.frac = reinterpret_cast<const double2*>(mesh_sm + LY::oFrac);
        }
        __syncthreads();
    }

    const int Nx = pd.Nx, nsteps = ra.nsteps;
    const double inv_tau_const = (ra.tau > 1.0e300) ? 0.0 : 1.0 / ra.tau;
    const double tol2 = ra.tol * ra.tol, tol2_lo = tol2 * (1.0 - 0x1p-48), tol2_hi = tol2 * (1.0 + 0x1p-48);
    const int64_t stride = (int64_t)gridDim.x * G;
    int round = 0;
    for (int64_t k = (int64_t)blockIdx.x * G + g; k < pd.B; k += stride, ++round) {
        // ---- the instance comes on chip --------------------------------------------------------------------------
        if (t < NPS) pbuf[t] = pd.prm[(size_t)k * NPS + t];
        if (t < 4 * P) mbuf[t] = pd.mass[(size_t)k * 4 * P + t];
        if constexpr (PIPE) {
            // pipelined upload: wait for the instance's chunk, then bring its rows from the reference layout into the
            // device layout (every thread moves the rows it owns, so the load below reads what this thread wrote).  Kept
            // apart from the load so that the code after it is the same with and without the pipeline.
            if (t == 0) {
                const unsigned int need = rp.ready_base + (unsigned int)(k / rp.chunk) + 1u;
                while ((int)(*rp.ready - need) < 0) __nanosleep(200);     // copy engines need no SM: this always ends
                __threadfence();
            }
            group_sync<TT>(bar);
            const double* src = rp.u_ref + (size_t)k * Nx * P;
            double* dst = pd.u + (size_t)k * L * P;
#pragma unroll 1
            for (int r = 0; r < R; ++r) {
                const int i = t * R + r;
                for (int p = 0; p < P; ++p) dst[(size_t)(r * TT + t) * P + p] = (i < Nx) ? __ldcg(src + (size_t)i * P + p) : 0.0;
            }
        }
        double uc[R][P];
        load_chunk<P, R, TT>((ra.b_given ? pd.b : pd.u) + (size_t)k * L * P, t, uc);
        const double* tg = ra.inst_tgrid ? ra.inst_tgrid + (size_t)k * (nsteps + 1) : ra.tgrid;
        auto snapshot = [&](int s) {   // reference layout: node-major, components fastest; a thread's rows are contiguous
            double* dst = ra.snap + ((size_t)s * pd.B + k) * Nx * P;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = t * R + r;
                if (i < Nx) {
#pragma unroll
                    for (int p = 0; p < P; ++p) __stcs(dst + (size_t)i * P + p, uc[r][p]);
                }
            }
        };
        group_sync<TT>(bar);               // parameters and mass flags visible
        if (ra.snap && !ra.snap_last && !ra.b_given) snapshot(0);

        long long total = 0;
        int status = 0, it_last = 0, failed_at = -1;
        double ss_last = 0.0;       // sum of h^2 of the last finished iteration (its root is taken when somebody asks for it)
        for (int s = 0; s < nsteps; ++s) {
            double tt, inv_tau;
            if (tg) {
                tt = tg[s + 1];
                if (ra.tau > 0.0) inv_tau = inv_tau_const;                  // (the division is hoisted out of the time loop)
                else { const double tau_s = tt - tg[s]; inv_tau = (tau_s > 1.0e300) ? 0.0 : 1.0 / tau_s; }
            } else { tt = ra.t_single; inv_tau = ra.inv_tau_single; }
            // ---- b = M .* u_n ; the Newton iterate starts from b (src/pdepe.jl:94, src/utils.jl:309) -------------
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = t * R + r;
                const int cls = (i == 0) ? 0 : ((i == Nx - 1) ? 2 : 1);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    double v = uc[r][p];
                    if (!(s == 0 && ra.b_given)) v = (i < Nx) ? mbuf[cls * P + p] * v : 0.0;
                    uc[r][p] = v;
                    btile[(r * TT + t) * P + p] = v;
                    utile[(r * TT + t) * P + p] = v;
                }
            }
            int it = 0;
            bool conv = false;
            double ss = 0.0;
            while (true) {
                group_sync<TT>(bar);       // (A) the iterate of every thread and the partial sums of h^2 are visible
                if (it > 0) {
                    if constexpr (TT > 32) {   // every thread forms the same total in the same order
                        const double* sl = sums + ((it - 1) & 1) * 8;
                        ss = sl[0];
#pragma unroll
                        for (int w = 1; w < TT / 32; ++w) ss += sl[w];
                    }
                    ss_last = ss;
                    if (t == 0 && pd.hist && it - 1 < pd.hist_cap) pd.hist[(size_t)k * pd.hist_cap + it - 1] = ::sqrt(ss);
                    // ||h||_2 > 1e300 or NaN <=> the sum of squares is not a finite number (a finite sum has a root below 1.4e154)
                    if (!(ss < __longlong_as_double(0x7ff0000000000000LL))) status = 2;     // goes on to maxit like the reference
                    // sqrt(ss) < tol (src/utils.jl:329-333), decided on ss wherever the rounding of the root cannot matter
                    const bool below = (ss < tol2_lo) || (!(ss > tol2_hi) && ::sqrt(ss) < ra.tol);
                    if (below) { conv = true; break; }
                    if (it >= ra.maxit) break;                                          // src/utils.jl:336
                }
                double bc[R][P];
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int p = 0; p < P; ++p) bc[r][p] = btile[(r * TT + t) * P + p];
                ChunkThomas<P, R> th;
                assemble_chunk<Fn, R, TT, SYM0, true, false>(pd, mesh, utile, btile, mbuf, t, tt, inv_tau, pbuf, uc, bc, th);
                Vec<P> h[R];
                const bool suspect = th.template finish<TT>(h, scratch + (it % LY::nexch) * LY::exch, t, bar);   // (B) inside
                if (suspect) {          // pivot growth: the same rows into a band, solved with row interchanges
                    double hp[R][P];
                    const bool singular = pivoted_newton_rows<Fn, R, TT, SYM0, false>(pd, mesh, k, utile, btile, t, blockIdx.x * G + g, tt,
                                                                                      inv_tau, sh_slot, hp, bar);
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int p = 0; p < P; ++p) h[r].a[p] = hp[r][p];
                    if (status != 2) status = singular ? 2 : 3;
                }
                // ---- u -= h (src/utils.jl:321); every thread is past its reads of the old iterate (barrier B) ----
                ss = 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        uc[r][p] -= h[r].a[p];
                        utile[(r * TT + t) * P + p] = uc[r][p];
                    }
                    if (t * R + r < Nx) {
#pragma unroll
                        for (int p = 0; p < P; ++p) ss += h[r].a[p] * h[r].a[p];
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ss += __shfl_down_sync(0xffffffffu, ss, o);
                if constexpr (TT > 32) {
                    if ((t & 31) == 0) sums[(it & 1) * 8 + (t >> 5)] = ss;
                } else {
                    ss = __shfl_sync(0xffffffffu, ss, 0);
                }
                ++it;
            }
            total += it;
            it_last = it;
            if (t == 0) {
                if (ra.step_iters) ra.step_iters[(size_t)k * nsteps + s] = it;
                if (ra.max_iters) atomicMax(ra.max_iters, it);
            }
            if (!conv) { failed_at = s; if (status == 0 || status == 3) status = 1; break; }
            if (ra.snap && (!ra.snap_last || s + 1 == nsteps)) snapshot(ra.snap_last ? 0 : s + 1);
        }
        // ---- the instance goes back to HBM -------------------------------------------------------------------------
        double* ug = pd.u + (size_t)k * L * P;
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int p = 0; p < P; ++p) ug[(size_t)(r * TT + t) * P + p] = uc[r][p];
        if (t == 0) {
            pd.iters[k] = it_last;
            atomicAdd((unsigned long long*)(pd.total_iters + k), (unsigned long long)total);
            atomicAdd(ra.sum_iters, (unsigned long long)total);
            pd.norm[k] = ::sqrt(ss_last);
            if (status) pd.status[k] = status;
            pd.active[k] = (failed_at >= 0) ? 1 : 0;
            if (failed_at >= 0) atomicMin(ra.fail_step, failed_at);
            if (ra.round_done) {        // snapshot streaming: the last instance of a round tells the host
                __threadfence();
                const int64_t first = (int64_t)round * stride;
                const unsigned int in_round = (unsigned int)((pd.B - first < stride) ? pd.B - first : stride);
                if (atomicAdd(ra.round_done + round, 1u) == in_round - 1) {
                    __threadfence_system();
                    ra.host_rounds[round] = 1u;
                }
            }
        }
        group_sync<TT>(bar);               // the tiles, parameters and flags are rewritten by the next instance
    }
}

template <class Fn, int R, int TT, bool SYM0>
__global__ void __launch_bounds__(resident_groups(TT) * TT, resident_min_blocks(TT, Fn::npde))
k_solve_resident(ProbDev pd, ResidentArgs ra) {
    ResidentPipe none;
    none.u_ref = nullptr; none.ready = nullptr; none.ready_base = 0u; none.chunk = 1;
    resident_body<Fn, R, TT, SYM0, false>(pd, ra, none);
}

// the variant launched by sb_solve_from_host: every instance waits for its chunk of the pipelined upload
template <class Fn, int R, int TT, bool SYM0>
__global__ void __launch_bounds__(resident_groups(TT) * TT, resident_min_blocks(TT, Fn::npde))
k_solve_resident_pipe(ProbDev pd, ResidentArgs ra, ResidentPipe rp) {
    resident_body<Fn, R, TT, SYM0, true>(pd, ra, rp);
}

}  // namespace sb
