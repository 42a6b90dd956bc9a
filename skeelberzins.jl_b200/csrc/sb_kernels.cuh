// sb_kernels.cuh -- sm_100a kernels of the batched Skeel-Berzins Newton iteration.
//
// Data layout in HBM ("chunk-transposed SoA").  A system (one PDE instance, Nx nodes, P = npde
// components) is owned by T threads; thread t owns the R consecutive nodes i = t*R + r, r < R
// (padded to L = R*T nodes).  Every per-node array is stored as [instance][r][t][P] so that for a
// fixed r the T threads of a system touch consecutive addresses: all global loads/stores of the hot
// kernels are fully coalesced while each thread still owns a contiguous piece of the mesh, which is
// what the partitioned (block-)Thomas elimination needs.  Mesh-dependent weights use the same
// [r][t] layout and are shared by the whole batch (L1/L2 resident).
//
// Kernels
//   k_assemble<Fn,R,WARP,JAC>  K3: residual F and block-tridiagonal Jacobian (JAC) or assemble! (!JAC)
//   k_solve<P,R,WARP>          K4/K5+K7: partitioned block-Thomas + PCR, update, ||h||_2, stop test
//   k_newton_fused<Fn,R,WARP>  K3+K4+K7 in one kernel: J and F never leave the SM
//   k_begin_step, k_mass_flags, k_to_device_layout, k_from_device_layout
#pragma once
#include <stdint.h>

#include "sb_dual.cuh"
#include "sb_functors.cuh"

namespace sb {

constexpr int kWarpBlockThreads = 128;  // WARP mode: 4 independent systems per CTA
constexpr int kMaxBlockT = 256;         // block mode: T threads = one system per CTA

struct MeshDev {          // [R][T] each; interval j = (node j, node j+1) is stored at node j's slot
    const double* xi;     // quadrature point of the interval (x argument of pdefun)
    const double* aw;     // U  = aw*ul + bw*ur           (src/utils.jl:18-48)
    const double* bw;
    const double* gw;     // Ux = gw*(ur - ul)
    const double* wf;     // flux weight: xi^m (regular, src/assembler.jl:84) | zeta^(m+1)/xi (singular, :74)
    const double* fracR;  // (zeta_i^(m+1) - x_i^(m+1))/(m+1)       frac1, src/assembler.jl:68 (and :36 for i=0)
    const double* fracL;  // (x_i^(m+1) - zeta_{i-1}^(m+1))/(m+1)   frac2, src/assembler.jl:69 (and :98 for i=N-1)
};

struct ProbDev {
    int Nx, R, T, L;      // L = R*T padded nodes
    int m, singular, nprm;
    int64_t B;
    double x0, xN, x0m, xNm, xi0;  // first/last mesh point, x^m of both, first quadrature point
    MeshDev mesh;
    const double* prm;    // [B][nprm]
    double* u;            // Newton iterate          [B][R][T][P]
    double* b;            // M .* u_n                [B][R][T][P]
    double* F;            // residual / rhs          [B][R][T][P]
    double* Jl;           // sub-diagonal blocks     [B][R][T][P*P]
    double* Jd;
    double* Ju;
    const double* mass;   // [B][3][P]: left node, interior, right node (0/1)
    int* active;          // [B]
    int* iters;           // [B] Newton iterations this step
    long long* total_iters;  // [B]
    double* norm;         // [B] last ||h||_2
    int* status;          // [B]
    double* hist;         // [B][hist_cap] or null
    int hist_cap;
    unsigned int* active_count;  // [slots]
};

// ---------------------------------------------------------------------------------------------
// neighbour exchange inside a system: shuffles (T == 32, one warp) or shared memory (T > 32)
// ---------------------------------------------------------------------------------------------
template <bool WARP, int NW>
SB_D void fetch2(const double (&mine)[NW], int tm, int tp, double (&outm)[NW], double (&outp)[NW], double* sm,
                 int t, int T) {
    if (WARP) {
        const bool vm = tm >= 0 && tm < T, vp = tp >= 0 && tp < T;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const double a = __shfl_sync(0xffffffffu, mine[w], tm & 31);
            const double c = __shfl_sync(0xffffffffu, mine[w], tp & 31);
            outm[w] = vm ? a : 0.0;
            outp[w] = vp ? c : 0.0;
        }
    } else {
#pragma unroll
        for (int w = 0; w < NW; ++w) sm[w * T + t] = mine[w];
        __syncthreads();
        const bool vm = tm >= 0 && tm < T, vp = tp >= 0 && tp < T;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            outm[w] = vm ? sm[w * T + tm] : 0.0;
            outp[w] = vp ? sm[w * T + tp] : 0.0;
        }
        __syncthreads();
    }
}

// sum over the T threads of a system, result valid in thread t == 0
template <bool WARP>
SB_D double system_sum(double v, double* sm, int t, int T) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (!WARP) {
        const int nw = T >> 5;
        if ((t & 31) == 0) sm[t >> 5] = v;
        __syncthreads();
        if (t == 0) {
            for (int w = 1; w < nw; ++w) v += sm[w];
        }
        __syncthreads();
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// Partitioned block-Thomas ("SPIKE"-type) elimination of one block-tridiagonal system.
// Rows arrive one by one (row(r, A, B, C, d): A x_{i-1} + B x_i + C x_{i+1} = d).  The thread
// eliminates its R-1 interior rows against its last row (the separator), the T separators form a
// reduced block-tridiagonal system solved by parallel cyclic reduction across the threads, then
// the interior unknowns follow by back substitution.  No pivoting between block rows (the Newton
// matrices of this discretisation are block diagonally dominant); pivoting inside P x P blocks.
// ---------------------------------------------------------------------------------------------
template <int P, int R>
struct ChunkThomas {
    static constexpr int RI = (R > 1) ? R - 1 : 1;
    Mat<P> ct[RI], vt[RI];
    Vec<P> dt[RI];
    Mat<P> As, Bs, Cs;
    Vec<P> ds;

    SB_D void row(int r, const Mat<P>& A, const Mat<P>& Bm, const Mat<P>& C, const Vec<P>& d) {
        if (r == R - 1) { As = A; Bs = Bm; Cs = C; ds = d; return; }
        if (r == 0) {
            const Mat<P> inv = inverse(Bm);
            ct[0] = mul(inv, C); vt[0] = mul(inv, A); dt[0] = mul(inv, d);
        } else {
            const Mat<P> inv = inverse(sub_mul(Bm, A, ct[r - 1]));
            ct[r] = mul(inv, C);
            dt[r] = mul(inv, sub_mul(d, A, dt[r - 1]));
            vt[r] = neg(mul(inv, mul(A, vt[r - 1])));
        }
    }

    // solve; x[r] = unknowns of the thread's rows
    template <bool WARP>
    SB_D void finish(Vec<P> (&x)[R], double* sm, int t, int T) {
        constexpr int NW = 2 * P * P + P;
        // representation of the first row: x_first = y0 - v0*x_prevsep - w0*x_sep
        // and of the last interior row:     x_last  = yL - vL*x_prevsep - wL*x_sep
        Vec<P> y0, yL; Mat<P> v0, w0, vL, wL;
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
        double mine[NW], om[NW], op[NW];
#pragma unroll
        for (int i = 0; i < P * P; ++i) { mine[i] = v0.a[i]; mine[P * P + i] = w0.a[i]; }
#pragma unroll
        for (int i = 0; i < P; ++i) mine[2 * P * P + i] = y0.a[i];
        fetch2<WARP, NW>(mine, -1, t + 1, om, op, sm, t, T);  // only the (t+1) side is used
        Mat<P> v0n, w0n; Vec<P> y0n;
#pragma unroll
        for (int i = 0; i < P * P; ++i) { v0n.a[i] = op[i]; w0n.a[i] = op[P * P + i]; }
#pragma unroll
        for (int i = 0; i < P; ++i) y0n.a[i] = op[2 * P * P + i];
        // reduced row of this thread's separator
        Mat<P> ar = neg(mul(As, vL));
        Mat<P> br = sub_mul(sub_mul(Bs, As, wL), Cs, v0n);
        Mat<P> cr = neg(mul(Cs, w0n));
        Vec<P> dr = sub_mul(sub_mul(ds, As, yL), Cs, y0n);
        // PCR on normalised rows
        Mat<P> inv = inverse(br);
        Mat<P> an = mul(inv, ar), cn = mul(inv, cr);
        Vec<P> dn = mul(inv, dr);
        for (int s = 1; s < T; s <<= 1) {
#pragma unroll
            for (int i = 0; i < P * P; ++i) { mine[i] = an.a[i]; mine[P * P + i] = cn.a[i]; }
#pragma unroll
            for (int i = 0; i < P; ++i) mine[2 * P * P + i] = dn.a[i];
            fetch2<WARP, NW>(mine, t - s, t + s, om, op, sm, t, T);
            Mat<P> am, cm, ap, cp; Vec<P> dm, dp;
#pragma unroll
            for (int i = 0; i < P * P; ++i) { am.a[i] = om[i]; cm.a[i] = om[P * P + i]; ap.a[i] = op[i]; cp.a[i] = op[P * P + i]; }
#pragma unroll
            for (int i = 0; i < P; ++i) { dm.a[i] = om[2 * P * P + i]; dp.a[i] = op[2 * P * P + i]; }
            const Mat<P> a2 = neg(mul(an, am));
            const Mat<P> c2 = neg(mul(cn, cp));
            const Mat<P> b2 = sub_mul(sub_mul(mat_identity<P>(), an, cm), cn, ap);
            const Vec<P> d2 = sub_mul(sub_mul(dn, an, dm), cn, dp);
            inv = inverse(b2);
            an = mul(inv, a2); cn = mul(inv, c2); dn = mul(inv, d2);
        }
        const Vec<P> xs = dn;  // separator unknown
        // previous separator
        double xm[P], xo[P], xq[P];
#pragma unroll
        for (int i = 0; i < P; ++i) xm[i] = xs.a[i];
        fetch2<WARP, P>(xm, t - 1, -1, xo, xq, sm, t, T);
        Vec<P> xp;
#pragma unroll
        for (int i = 0; i < P; ++i) xp.a[i] = xo[i];
        x[R - 1] = xs;
#pragma unroll
        for (int r = R - 2; r >= 0; --r) x[r] = sub_mul(sub_mul(dt[r], vt[r], xp), ct[r], x[r + 1]);
    }
};

// ---------------------------------------------------------------------------------------------
// Assembly of the rows of one chunk (assemble!, src/assembler.jl:19-125, plus implicitEuler!,
// src/utils.jl:245-257, and the Jacobian the reference obtains by forward-mode AD, :312-314).
// ---------------------------------------------------------------------------------------------
template <int P, int ND>
struct IntervalRes {  // (c,f,s) of one interval; partials w.r.t. (u_left[0..P), u_right[0..P))
    Dual<ND> c[P], f[P], s[P];
    double wf;
};

template <class Fn, int ND>
SB_D void eval_interval(const ProbDev& pd, int slot, const double* ul, const double* ur, double tt, const double* prm,
                        IntervalRes<Fn::npde, ND>& out) {
    constexpr int P = Fn::npde;
    const double xi = __ldg(pd.mesh.xi + slot), aw = __ldg(pd.mesh.aw + slot), bw = __ldg(pd.mesh.bw + slot),
                 gw = __ldg(pd.mesh.gw + slot);
    out.wf = __ldg(pd.mesh.wf + slot);
    Dual<ND> U[P], Ux[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        U[p] = Dual<ND>(aw * ul[p] + bw * ur[p]);
        Ux[p] = Dual<ND>(gw * (ur[p] - ul[p]));
        if constexpr (ND > 0) {
            U[p].d[p] = aw; U[p].d[P + p] = bw;
            Ux[p].d[p] = -gw; Ux[p].d[P + p] = gw;
        }
    }
    Fn::template pde<Dual<ND>>(xi, tt, U, Ux, prm, out.c, out.f, out.s);
}

// Sink concept: row(r, Jl, Jd, Ju, F) when JAC, value(r, du) otherwise.
template <class Fn, int R, bool JAC, class Sink>
SB_D void assemble_chunk(const ProbDev& pd, int64_t k, int t, double tt, double inv_tau, const double* prm,
                         const double (&uc)[R][Fn::npde], const double (&bc)[R][Fn::npde], Sink& sink) {
    constexpr int P = Fn::npde;
    constexpr int ND = JAC ? 2 * P : 0;
    constexpr int PO = JAC ? P : 0;  // offset of the right-end partials
    const int T = pd.T, Nx = pd.Nx;
    const int i0 = t * R;
    const double* ub = pd.u + (size_t)k * pd.L * P;
    const double* mass = pd.mass + (size_t)k * 3 * P;

    double un[P];  // first node of the next chunk
    IntervalRes<P, ND> Lr, Rr;
    if (i0 >= 1 && i0 - 1 <= Nx - 2) {
        double ul[P];
        const int slot = (R - 1) * T + t - 1;
#pragma unroll
        for (int p = 0; p < P; ++p) ul[p] = ub[(size_t)slot * P + p];
        eval_interval<Fn, ND>(pd, slot, ul, uc[0], tt, prm, Lr);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) un[p] = (t + 1 < T) ? ub[(size_t)(t + 1) * P + p] : 0.0;

#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = i0 + r;
        const int slot = r * T + t;
        if (i <= Nx - 2) eval_interval<Fn, ND>(pd, slot, uc[r], (r + 1 < R) ? uc[r + 1] : un, tt, prm, Rr);

        double du[P];
        double dm[P][P], dc[P][P], dp[P][P];  // d du_j / d u_{i-1,k}, u_{i,k}, u_{i+1,k}
        double mss[P];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            du[j] = 0.0; mss[j] = 0.0;
#pragma unroll
            for (int q = 0; q < P; ++q) { dm[j][q] = 0.0; dc[j][q] = 0.0; dp[j][q] = 0.0; }
        }
        bool pad = false;
        if (i >= Nx) {
            pad = true;
        } else if (i == 0 || i == Nx - 1) {
            // boundary rows: bdfun sees both end values (src/assembler.jl:26/29)
            double u0[P], uN[P];
            const int sN = ((Nx - 1) % R) * T + (Nx - 1) / R;
#pragma unroll
            for (int p = 0; p < P; ++p) { u0[p] = ub[p]; uN[p] = ub[(size_t)sN * P + p]; }
            Dual<PO> ul_[P], ur_[P], pl[P], pr[P];
            double ql[P], qr[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                ul_[p] = Dual<PO>(u0[p]); ur_[p] = Dual<PO>(uN[p]);
                if constexpr (JAC) { ul_[p].d[p] = 1.0; ur_[p].d[p] = 1.0; }
            }
            Fn::template bc<Dual<PO>>(pd.x0, ul_, pd.xN, ur_, tt, prm, pl, ql, pr, qr);
            if (i == 0) {
                const double frac = __ldg(pd.mesh.fracR + slot);
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    mss[j] = mass[j];
                    const Dual<ND>&c = Rr.c[j], &f = Rr.f[j], &s = Rr.s[j];
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
                        const double num = pl[j].v + qx * (Rr.wf * f.v + frac * s.v);
                        const bool hc = c.v != 0.0;
                        const double den = qx * frac * c.v;
                        const double id = hc ? 1.0 / den : 1.0;
                        du[j] = num * id;
                        if constexpr (JAC) {
#pragma unroll
                            for (int q = 0; q < P; ++q) {
                                const double n0 = pl[j].d[q] + qx * (Rr.wf * f.d[q] + frac * s.d[q]);
                                const double n1 = qx * (Rr.wf * f.d[PO + q] + frac * s.d[PO + q]);
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
            } else {  // right boundary, src/assembler.jl:98-122
                const double frac = __ldg(pd.mesh.fracL + slot);
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    mss[j] = mass[2 * P + j];
                    const Dual<ND>&c = Lr.c[j], &f = Lr.f[j], &s = Lr.s[j];
                    if (qr[j] != 0.0) {
                        const double qx = qr[j] / pd.xNm;
                        const double num = pr[j].v + qx * (Lr.wf * f.v - frac * s.v);
                        const bool hc = c.v != 0.0;
                        const double den = -qx * frac * c.v;
                        const double id = hc ? 1.0 / den : 1.0;
                        du[j] = num * id;
                        if constexpr (JAC) {
#pragma unroll
                            for (int q = 0; q < P; ++q) {
                                const double n0 = qx * (Lr.wf * f.d[q] - frac * s.d[q]);
                                const double n1 = pr[j].d[q] + qx * (Lr.wf * f.d[PO + q] - frac * s.d[PO + q]);
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
        } else {  // interior node, src/assembler.jl:60-95
            const double frR = __ldg(pd.mesh.fracR + slot), frL = __ldg(pd.mesh.fracL + slot);
#pragma unroll
            for (int j = 0; j < P; ++j) {
                mss[j] = mass[P + j];
                const double num = Rr.wf * Rr.f[j].v - Lr.wf * Lr.f[j].v + frR * Rr.s[j].v + frL * Lr.s[j].v;
                const bool hc = (Lr.c[j].v != 0.0) || (Rr.c[j].v != 0.0);
                const double den = frR * Rr.c[j].v + frL * Lr.c[j].v;
                const double id = hc ? 1.0 / den : 1.0;
                du[j] = num * id;
                if constexpr (JAC) {
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        const double nm = frL * Lr.s[j].d[q] - Lr.wf * Lr.f[j].d[q];
                        const double nc = Rr.wf * Rr.f[j].d[q] - Lr.wf * Lr.f[j].d[PO + q] + frR * Rr.s[j].d[q] +
                                          frL * Lr.s[j].d[PO + q];
                        const double np = Rr.wf * Rr.f[j].d[PO + q] + frR * Rr.s[j].d[PO + q];
                        dm[j][q] = hc ? (nm - du[j] * (frL * Lr.c[j].d[q])) * id : nm;
                        dc[j][q] = hc ? (nc - du[j] * (frR * Rr.c[j].d[q] + frL * Lr.c[j].d[PO + q])) * id : nc;
                        dp[j][q] = hc ? (np - du[j] * (frR * Rr.c[j].d[PO + q])) * id : np;
                    }
                }
            }
        }

        if constexpr (JAC) {
            Mat<P> A, Bm, C; Vec<P> d;
            if (pad) {
                A = mat_zero<P>(); C = mat_zero<P>(); Bm = mat_identity<P>(); d = vec_zero<P>();
            } else {
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const double mt = inv_tau * mss[j];
                    d.a[j] = (mt * uc[r][j] - du[j]) - inv_tau * bc[r][j];  // src/utils.jl:251-256, :314
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        A.a[j * P + q] = -dm[j][q];
                        Bm.a[j * P + q] = -dc[j][q] + ((j == q) ? mt : 0.0);
                        C.a[j * P + q] = -dp[j][q];
                    }
                }
            }
            sink.row(r, A, Bm, C, d);
        } else {
            Vec<P> d;
#pragma unroll
            for (int j = 0; j < P; ++j) d.a[j] = du[j];
            sink.value(r, d);
        }
        Lr = Rr;
    }
}

// ---------------------------------------------------------------------------------------------
// thread/system bookkeeping
// ---------------------------------------------------------------------------------------------
template <bool WARP>
SB_D bool locate(const ProbDev& pd, int64_t& k, int& t) {
    if (WARP) {
        k = (int64_t)blockIdx.x * (kWarpBlockThreads / 32) + (threadIdx.x >> 5);
        t = threadIdx.x & 31;
    } else {
        k = blockIdx.x;
        t = threadIdx.x;
    }
    return k < pd.B;
}

template <int P, int R>
SB_D void load_chunk(const double* base, int t, int T, double (&v)[R][P]) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int p = 0; p < P; ++p) v[r][p] = base[(size_t)(r * T + t) * P + p];
}

// update u, ||h||_2, iteration counter and stop test (src/utils.jl:321-333)
template <int P, int R, bool WARP>
SB_D void update_and_test(const ProbDev& pd, int64_t k, int t, const double (&uc)[R][P], const Vec<P> (&h)[R],
                          double tol, int slot, double* sm) {
    const int T = pd.T;
    double* ub = pd.u + (size_t)k * pd.L * P;
    double ss = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (t * R + r < pd.Nx) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                ub[(size_t)(r * T + t) * P + p] = uc[r][p] - h[r].a[p];
                ss += h[r].a[p] * h[r].a[p];
            }
        }
    }
    ss = system_sum<WARP>(ss, sm, t, T);
    if (t == 0) {
        const double nm = ::sqrt(ss);
        const int it = pd.iters[k];
        pd.norm[k] = nm;
        if (pd.hist && it < pd.hist_cap) pd.hist[(size_t)k * pd.hist_cap + it] = nm;
        pd.iters[k] = it + 1;
        pd.total_iters[k] += 1;
        const bool still = !(nm < tol);
        if (!still) pd.active[k] = 0;
        else atomicAdd(pd.active_count + slot, 1u);
        if (!(nm == nm) || nm > 1.0e300) pd.status[k] = 2;
    }
}

// ---------------------------------------------------------------------------------------------
// K3: residual + Jacobian to HBM (JAC) / assemble! value into F (!JAC)
// ---------------------------------------------------------------------------------------------
template <int P, int R>
struct StoreSink {
    const ProbDev& pd; int64_t k; int t;
    SB_D void row(int r, const Mat<P>& A, const Mat<P>& Bm, const Mat<P>& C, const Vec<P>& d) {
        const size_t o = (size_t)k * pd.L + (size_t)r * pd.T + t;
#pragma unroll
        for (int i = 0; i < P * P; ++i) { pd.Jl[o * P * P + i] = A.a[i]; pd.Jd[o * P * P + i] = Bm.a[i]; pd.Ju[o * P * P + i] = C.a[i]; }
#pragma unroll
        for (int i = 0; i < P; ++i) pd.F[o * P + i] = d.a[i];
    }
    SB_D void value(int r, const Vec<P>& d) {
        const size_t o = (size_t)k * pd.L + (size_t)r * pd.T + t;
#pragma unroll
        for (int i = 0; i < P; ++i) pd.F[o * P + i] = d.a[i];
    }
};

template <class Fn, int R, bool WARP, bool JAC>
__global__ void __launch_bounds__(WARP ? kWarpBlockThreads : kMaxBlockT)
k_assemble(ProbDev pd, double tt, double inv_tau, int all_instances) {
    constexpr int P = Fn::npde;
    int64_t k; int t;
    if (!locate<WARP>(pd, k, t)) return;
    if (!all_instances && !pd.active[k]) return;
    double uc[R][P], bc[R][P];
    load_chunk<P, R>(pd.u + (size_t)k * pd.L * P, t, pd.T, uc);
    if constexpr (JAC) load_chunk<P, R>(pd.b + (size_t)k * pd.L * P, t, pd.T, bc);
    StoreSink<P, R> sink{pd, k, t};
    assemble_chunk<Fn, R, JAC>(pd, k, t, tt, inv_tau, pd.prm + (size_t)k * pd.nprm, uc, bc, sink);
}

// ---------------------------------------------------------------------------------------------
// K4/K5 + K7: solve J h = F from HBM, update, norm, stop test
// ---------------------------------------------------------------------------------------------
template <int P, int R, bool WARP>
__global__ void __launch_bounds__(WARP ? kWarpBlockThreads : kMaxBlockT)
k_solve(ProbDev pd, double tol, int slot, int update) {
    extern __shared__ double sm[];
    int64_t k; int t;
    if (!locate<WARP>(pd, k, t)) return;
    if (update && !pd.active[k]) return;
    const int T = pd.T;
    ChunkThomas<P, R> th;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const size_t o = (size_t)k * pd.L + (size_t)r * T + t;
        Mat<P> A, Bm, C; Vec<P> d;
#pragma unroll
        for (int i = 0; i < P * P; ++i) { A.a[i] = pd.Jl[o * P * P + i]; Bm.a[i] = pd.Jd[o * P * P + i]; C.a[i] = pd.Ju[o * P * P + i]; }
#pragma unroll
        for (int i = 0; i < P; ++i) d.a[i] = pd.F[o * P + i];
        th.row(r, A, Bm, C, d);
    }
    Vec<P> h[R];
    th.template finish<WARP>(h, sm, t, T);
    if (update) {
        double uc[R][P];
        load_chunk<P, R>(pd.u + (size_t)k * pd.L * P, t, T, uc);
        update_and_test<P, R, WARP>(pd, k, t, uc, h, tol, slot, sm);
    } else {  // debug: leave x in F
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < P; ++i) pd.F[((size_t)k * pd.L + (size_t)r * T + t) * P + i] = h[r].a[i];
    }
}

// ---------------------------------------------------------------------------------------------
// fused Newton iteration: assemble rows straight into the elimination
// ---------------------------------------------------------------------------------------------
template <class Fn, int R, bool WARP>
__global__ void __launch_bounds__(WARP ? kWarpBlockThreads : kMaxBlockT)
k_newton_fused(ProbDev pd, double tt, double inv_tau, double tol, int slot) {
    constexpr int P = Fn::npde;
    extern __shared__ double sm[];
    int64_t k; int t;
    if (!locate<WARP>(pd, k, t)) return;
    if (!pd.active[k]) return;
    double uc[R][P], bc[R][P];
    load_chunk<P, R>(pd.u + (size_t)k * pd.L * P, t, pd.T, uc);
    load_chunk<P, R>(pd.b + (size_t)k * pd.L * P, t, pd.T, bc);
    ChunkThomas<P, R> th;
    assemble_chunk<Fn, R, true>(pd, k, t, tt, inv_tau, pd.prm + (size_t)k * pd.nprm, uc, bc, th);
    Vec<P> h[R];
    th.template finish<WARP>(h, sm, t, pd.T);
    update_and_test<P, R, WARP>(pd, k, t, uc, h, tol, slot, sm);
}

// ---------------------------------------------------------------------------------------------
// K2: mass flags (mass_matrix, src/assembler.jl:138-174), one thread per instance
// ---------------------------------------------------------------------------------------------
template <class Fn>
__global__ void k_mass_flags(ProbDev pd, double t0, double* mass, int stationary) {
    constexpr int P = Fn::npde;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pd.B) return;
    double* mk = mass + (size_t)k * 3 * P;
    if (stationary) {
        for (int i = 0; i < 3 * P; ++i) mk[i] = 1.0;
        return;
    }
    const double* ub = pd.u + (size_t)k * pd.L * P;
    const double* prm = pd.prm + (size_t)k * pd.nprm;
    const int R = pd.R, T = pd.T, Nx = pd.Nx;
    auto slot_of = [&](int i) { return (i % R) * T + i / R; };
    double u0[P], u1[P], uN[P];
    for (int p = 0; p < P; ++p) {
        u0[p] = ub[(size_t)slot_of(0) * P + p];
        u1[p] = ub[(size_t)slot_of(1) * P + p];
        uN[p] = ub[(size_t)slot_of(Nx - 1) * P + p];
    }
    Dual<0> ul_[P], ur_[P], pl[P], pr[P];
    double ql[P], qr[P];
    for (int p = 0; p < P; ++p) { ul_[p] = Dual<0>(u0[p]); ur_[p] = Dual<0>(uN[p]); }
    Fn::template bc<Dual<0>>(pd.x0, ul_, pd.xN, ur_, t0, prm, pl, ql, pr, qr);
    IntervalRes<P, 0> res;
    eval_interval<Fn, 0>(pd, slot_of(0), u0, u1, t0, prm, res);
    for (int i = 0; i < P; ++i) {
        mk[i] = (ql[i] == 0.0 && !pd.singular) ? 0.0 : 1.0;   // :162-165
        mk[P + i] = (res.c[i].v == 0.0) ? 0.0 : 1.0;           // :157-160
        mk[2 * P + i] = (qr[i] == 0.0) ? 0.0 : 1.0;            // :167-170
    }
}

// b = M .* u_n ; u <- b ; active, counters (src/pdepe.jl:94, src/utils.jl:309)
template <int P>
__global__ void k_begin_step(ProbDev pd) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*L node slots
    const size_t total = (size_t)pd.B * pd.L;
    if (idx >= total) return;
    const int64_t k = idx / pd.L;
    const int s = (int)(idx - (size_t)k * pd.L);
    const int r = s / pd.T, t = s - r * pd.T;
    const int i = t * pd.R + r;
    if (s == 0) { pd.active[k] = 1; pd.iters[k] = 0; }
    const double* mass = pd.mass + (size_t)k * 3 * P;
    const int cls = (i == 0) ? 0 : ((i == pd.Nx - 1) ? 2 : 1);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const double mv = mass[cls * P + p];
        const double v = (i < pd.Nx) ? mv * pd.u[idx * P + p] : 0.0;
        pd.b[idx * P + p] = v;
        if (mv == 0.0 || i >= pd.Nx) pd.u[idx * P + p] = v;  // u <- b differs from u only where M = 0
    }
}

}  // namespace sb
