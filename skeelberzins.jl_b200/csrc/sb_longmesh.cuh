// sb_longmesh.cuh -- K6: Newton iteration on meshes too long for one CTA (e.g. 2^22 nodes, one instance).
//
// The mesh is stored as tiles of R*TT nodes, each tile in the chunk-transposed layout of the batched path,
// so an instance looks like a batch of coupled tiles.  One Newton iteration is a recursive partition
// ("SPIKE") solve over levels:
//
//   level 0   k6_chunk_asm   every thread assembles its R rows on the fly (residual + Jacobian by duals, nothing
//                            but u, b and the mesh weights is read) and eliminates them -> 10 numbers per chunk
//             k6_form        reduced row of every chunk separator = row of the level-1 tridiagonal system
//   level l   k6_chunk_sys   the same elimination on the stored rows of level l        -> level l+1 ...
//   top       k6_top         a single CTA solves the last level (<= 4096 rows) with the batched-path solver
//   level l   k6_back_sys    recompute the elimination, back-substitute with the separator values from level l+1
//   level 0   k6_back_asm    recompute assembly + elimination, back-substitute, u -= h in place, per-tile sum of h^2
//             k6_finalize    ||h||_2 per instance in a fixed summation order, stop test, counters, host notice
//
// Recomputing instead of storing the block rows keeps the traffic at level 0 to two reads of (u, b, weights) and
// one write of u per Newton iteration.  Scalar PDEs (npde = 1) only.
#pragma once
#include "sb_kernels.cuh"

namespace sb {

constexpr int kLongR = 8, kLongTT = 256;  // decomposition of every level of the long-mesh path

struct LevelDev {      // one level of the recursive partition (level 0 has no stored rows: they are assembled)
    int n;             // rows
    int ntiles;        // tiles of kLongR*kLongTT rows
    size_t stride;     // doubles per instance in a, c, d, x  (= ntiles * kLongR * kLongTT)
    size_t cstride;    // chunks per instance (= ntiles * kLongTT)
    double *a, *c, *d; // normalised rows (unit diagonal) in tiled layout
    double* x;         // solution of the level in tiled layout
    double* chunk;     // [B][10][cstride]: y0 v0 w0 yL vL wL As Bs Cs ds of every chunk
    double* edge;      // levels 0 and 1 only: [B][2*ntiles0 + 2] iterate at the first / last node of every level-0 tile
                       // and at mesh nodes 0 and Nx-1, saved by the forward sweep so that the backward sweep can
                       // update u in place (a tile's neighbours may already hold the new iterate)
};

SB_D size_t tiled_pos(int j, int R, int TT) {
    const int Lt = R * TT;
    const int tile = j / Lt, il = j - tile * Lt;
    return (size_t)tile * Lt + (size_t)(il % R) * TT + il / R;
}

template <int R>
SB_D void store_chunk(const ChunkThomas<1, R>& th, double* chunk, size_t cstride, size_t c) {
    Vec<1> y0, yL; Mat<1> v0, w0, vL, wL;
    th.summary(y0, v0, w0, yL, vL, wL);
    chunk[0 * cstride + c] = y0.a[0]; chunk[1 * cstride + c] = v0.a[0]; chunk[2 * cstride + c] = w0.a[0];
    chunk[3 * cstride + c] = yL.a[0]; chunk[4 * cstride + c] = vL.a[0]; chunk[5 * cstride + c] = wL.a[0];
    chunk[6 * cstride + c] = th.As.a[0]; chunk[7 * cstride + c] = th.Bs.a[0]; chunk[8 * cstride + c] = th.Cs.a[0];
    chunk[9 * cstride + c] = th.ds.a[0];
}

// mesh weight pointers of one tile
SB_D MeshDev mesh_of_tile(const MeshDev& m, size_t off) {
    MeshDev r = m;
    r.xi += off; r.awbw += off; r.gwwf += off; r.frac += off; r.gw += off; r.ivol += off;
    return r;
}

// Shared-memory staging of one tile for the m == 0 kernels: u (with the edge nodes of the neighbouring tiles), b and
// the mesh weights arrive by bulk copies behind one mbarrier, so a sweep pays the DRAM latency once instead of once
// per row.  (m >= 1 needs 9 instead of 6 arrays per tile, which would leave one CTA per SM: those kernels keep
// reading global memory directly.)
template <class Fn, int R, int TT, bool SYM0>
struct LongStage {
    static constexpr bool on = SYM0;
    static constexpr int L = R * TT;
    static constexpr bool kFrac = !(Fn::c_unit && Fn::s_zero);
    static constexpr int oU = 2;                                  // [L + 4]: node slot s at oU + 2 + s, s = -2 .. L+1
    static constexpr int oB = oU + L + 4;                         // [L]
    static constexpr int oGw = oB + L;                            // [L + 2]: interval slot s at oGw + 2 + s, s = -2 .. L-1
    static constexpr int oIvol = oGw + L + 2;                     // [L]   (c == 1 functors)
    static constexpr int oFrac = oIvol + (Fn::c_unit ? L : 0);    // [2L]  (functors with c or s)
    static constexpr int doubles = oFrac + (kFrac ? 2 * L : 0);
    static constexpr size_t bytes = (on ? doubles : oB) * sizeof(double);   // m >= 1: the backward sweep stages u only
};

template <class Fn, int R, int TT>
SB_D void long_stage_tile(const ProbDev& pd, int64_t k, int tile, bool use_b, bool halo_u, double* sm, MeshDev& mesh,
                          const double*& ub, const double*& bb) {
    using LS = LongStage<Fn, R, TT, true>;
    constexpr int L = LS::L;
    constexpr uint32_t LB = (uint32_t)(L * sizeof(double));
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
    const size_t toff = (size_t)tile * L;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int lo = (tile > 0) ? 2 : 0, hi = (tile + 1 < pd.ntiles) ? 2 : 0;  // halo towards existing neighbours
        const int ulo = halo_u ? lo : 0, uhi = halo_u ? hi : 0;
        uint32_t bytes = (uint32_t)((L + ulo + uhi) * sizeof(double)) + (uint32_t)((L + lo) * sizeof(double));
        if (use_b) bytes += LB;
        if (Fn::c_unit) bytes += LB;
        if (LS::kFrac) bytes += 2 * LB;
        mbar_expect_tx(bar, bytes);
        bulk_g2s(sm + LS::oU + 2 - ulo, pd.u + (size_t)k * pd.L + toff - ulo, (uint32_t)((L + ulo + uhi) * sizeof(double)), bar);
        if (use_b) bulk_g2s(sm + LS::oB, pd.b + (size_t)k * pd.L + toff, LB, bar);
        bulk_g2s(sm + LS::oGw + 2 - lo, pd.mesh.gw + toff - lo, (uint32_t)((L + lo) * sizeof(double)), bar);
        if constexpr (Fn::c_unit) bulk_g2s(sm + LS::oIvol, pd.mesh.ivol + toff, LB, bar);
        if constexpr (LS::kFrac) bulk_g2s(sm + LS::oFrac, pd.mesh.frac + toff, 2 * LB, bar);
    }
    mesh = mesh_of_tile(pd.mesh, toff);  // xi (read only by functors that use x) stays in global memory
    mesh.gw = sm + LS::oGw + 2;
    if constexpr (Fn::c_unit) mesh.ivol = sm + LS::oIvol;
    if constexpr (LS::kFrac) mesh.frac = reinterpret_cast<const double2*>(sm + LS::oFrac);
    ub = sm + LS::oU + 2;
    bb = use_b ? sm + LS::oB : ub;       // tau = Inf: b is multiplied by 1/tau = 0 wherever it is still named
    mbar_wait(bar, 0);
}

// level 0, forward: assemble + eliminate the chunk, store its summary
template <class Fn, int R, int TT, bool SYM0>
__global__ void __launch_bounds__(TT) k6_chunk_asm(ProbDev pd, double tt, double inv_tau, LevelDev lv) {
    static_assert(Fn::npde == 1, "long-mesh path: scalar PDEs only");
    extern __shared__ __align__(16) double k6sm[];
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    if (!pd.active[k]) return;
    const size_t toff = (size_t)tile * R * TT;
    const double* ui = pd.u + (size_t)k * pd.L;
    const double* ub = ui + toff;
    const double* bb = pd.b + (size_t)k * pd.L + toff;
    const bool use_b = inv_tau != 0.0;  // stationary solves (tau = Inf) never look at b: rhs - b/tau = rhs
    MeshDev mesh;
    if constexpr (SYM0) long_stage_tile<Fn, R, TT>(pd, k, tile, use_b, true, k6sm, mesh, ub, bb);
    else mesh = mesh_of_tile(pd.mesh, toff);
    double uc[R][1], bc[R][1];
#pragma unroll
    for (int r = 0; r < R; ++r) { uc[r][0] = ub[r * TT + t]; bc[r][0] = use_b ? bb[r * TT + t] : 0.0; }
    {   // save the iterate at the tile edges and the mesh ends for the in-place backward sweep
        double* e = lv.edge + (size_t)k * (2 * pd.ntiles + 2);
        if (t == 0) e[2 * tile] = uc[0][0];
        if (t == TT - 1) e[2 * tile + 1] = uc[R - 1][0];
        if (tile == 0 && t == 0) e[2 * pd.ntiles] = uc[0][0];
        const int i0 = (int)toff + t * R;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (i0 + r == pd.Nx - 1) e[2 * pd.ntiles + 1] = uc[r][0];
    }
    ChunkThomas<1, R> th;
    assemble_chunk<Fn, R, TT, SYM0, true, !SYM0>(pd, mesh, ub, bb, pd.mass + (size_t)k * 4, t, tt, inv_tau,
                                                 pd.prm + (size_t)k * pd.nprm_s, uc, bc, th, (int)toff, ui,
                                                 ui + tiled_pos(pd.Nx - 1, R, TT));
    store_chunk<R>(th, lv.chunk + (size_t)k * 10 * lv.cstride, lv.cstride, (size_t)tile * TT + t);
}

// reduced separator rows of level `lv` -> rows of the next level `up` (tiled layout of the next level)
template <int RN, int TTN>
__global__ void k6_form(LevelDev lv, LevelDev up, const int* active) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k = blockIdx.y;
    if (c >= lv.cstride || !active[k]) return;
    const double* ch = lv.chunk + (size_t)k * 10 * lv.cstride;
    const size_t cs = lv.cstride;
    Vec<1> yL, y0n, ds, dn; Mat<1> vL, wL, v0n, w0n, As, Bs, Cs, an, cn;
    yL.a[0] = ch[3 * cs + c]; vL.a[0] = ch[4 * cs + c]; wL.a[0] = ch[5 * cs + c];
    As.a[0] = ch[6 * cs + c]; Bs.a[0] = ch[7 * cs + c]; Cs.a[0] = ch[8 * cs + c]; ds.a[0] = ch[9 * cs + c];
    const bool nx = c + 1 < cs;
    y0n.a[0] = nx ? ch[0 * cs + c + 1] : 0.0; v0n.a[0] = nx ? ch[1 * cs + c + 1] : 0.0; w0n.a[0] = nx ? ch[2 * cs + c + 1] : 0.0;
    ChunkThomas<1, 1>::reduced_row(As, Bs, Cs, ds, yL, vL, wL, v0n, w0n, y0n, an, cn, dn);
    const size_t pos = (size_t)k * up.stride + tiled_pos((int)c, RN, TTN);
    up.a[pos] = an.a[0]; up.c[pos] = cn.a[0]; up.d[pos] = dn.a[0];
}

template <int R, int TT>
SB_D void load_sys_rows(const LevelDev& lv, int64_t k, int tile, int t, ChunkThomas<1, R>& th) {
    const size_t base = (size_t)k * lv.stride + (size_t)tile * R * TT;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const size_t o = base + (size_t)r * TT + t;
        Mat<1> A, Bm, C; Vec<1> d;
        A.a[0] = lv.a[o]; Bm.a[0] = 1.0; C.a[0] = lv.c[o]; d.a[0] = lv.d[o];
        th.row(r, A, Bm, C, d);
    }
}

// level l >= 1, forward
template <int R, int TT>
__global__ void __launch_bounds__(TT) k6_chunk_sys(LevelDev lv, const int* active) {
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    if (!active[k]) return;
    ChunkThomas<1, R> th;
    load_sys_rows<R, TT>(lv, k, tile, t, th);
    store_chunk<R>(th, lv.chunk + (size_t)k * 10 * lv.cstride, lv.cstride, (size_t)tile * TT + t);
}

// top level: one CTA per instance, rows fit one tile of RT*TT rows
template <int RT, int TT>
__global__ void __launch_bounds__(TT) k6_top(LevelDev lv, const int* active) {
    extern __shared__ double sm[];
    const int t = threadIdx.x;
    const int64_t k = blockIdx.x;
    if (!active[k]) return;
    ChunkThomas<1, RT> th;
    const size_t base = (size_t)k * lv.stride;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        // the level is stored with the (kLongR, kLongTT) tiling; the top solver owns rows t*RT .. t*RT+RT-1
        const int j = t * RT + r;
        const size_t o = base + tiled_pos(j, kLongR, kLongTT);
        Mat<1> A, Bm, C; Vec<1> d;
        const bool in = j < lv.n;
        A.a[0] = in ? lv.a[o] : 0.0; Bm.a[0] = 1.0; C.a[0] = in ? lv.c[o] : 0.0; d.a[0] = in ? lv.d[o] : 0.0;
        th.row(r, A, Bm, C, d);
    }
    Vec<1> x[RT];
    th.template finish<TT>(x, sm, t);
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const int j = t * RT + r;
        if (j < lv.n) lv.x[base + tiled_pos(j, kLongR, kLongTT)] = x[r].a[0];
    }
}

// separator unknowns of chunk c (own and previous) from the solution of the next level
SB_D void separators(const LevelDev& up, int64_t k, size_t c, Vec<1>& xp, Vec<1>& xs) {
    const double* xu = up.x + (size_t)k * up.stride;
    xs.a[0] = xu[tiled_pos((int)c, kLongR, kLongTT)];
    xp.a[0] = (c > 0) ? xu[tiled_pos((int)c - 1, kLongR, kLongTT)] : 0.0;
}

// level l >= 1, backward
template <int R, int TT>
__global__ void __launch_bounds__(TT) k6_back_sys(LevelDev lv, LevelDev up, const int* active) {
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    if (!active[k]) return;
    ChunkThomas<1, R> th;
    load_sys_rows<R, TT>(lv, k, tile, t, th);
    Vec<1> xp, xs, x[R];
    separators(up, k, (size_t)tile * TT + t, xp, xs);
    th.back(x, xp, xs);
    const size_t base = (size_t)k * lv.stride + (size_t)tile * R * TT;
#pragma unroll
    for (int r = 0; r < R; ++r) lv.x[base + (size_t)r * TT + t] = x[r].a[0];
}

// level 0, backward: recompute, back-substitute, update u, per-tile sum of h^2
template <class Fn, int R, int TT, bool SYM0>
__global__ void __launch_bounds__(TT) k6_back_asm(ProbDev pd, double tt, double inv_tau, LevelDev up, double* partial) {
    __shared__ double red[TT / 32];
    extern __shared__ __align__(16) double k6sm[];
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    if (!pd.active[k]) return;
    const size_t toff = (size_t)tile * R * TT;
    double* ug = pd.u + (size_t)k * pd.L + toff;
    const double* ub = ug;
    const double* bb = pd.b + (size_t)k * pd.L + toff;
    const bool use_b = inv_tau != 0.0;  // stationary solves (tau = Inf) never look at b: rhs - b/tau = rhs
    // u is updated in place while other tiles are still being read: everything this CTA needs from outside its own
    // tile (the neighbours' edge nodes, both mesh ends for the boundary conditions) comes from the copy the
    // forward sweep saved in `edge`.
    const double* e = up.edge + (size_t)k * (2 * pd.ntiles + 2);
    using LS = LongStage<Fn, R, TT, SYM0>;
    MeshDev mesh;
    double uc[R][1], bc[R][1];
    if constexpr (SYM0) {
        long_stage_tile<Fn, R, TT>(pd, k, tile, use_b, false, k6sm, mesh, ub, bb);
    } else {
        mesh = mesh_of_tile(pd.mesh, toff);
#pragma unroll
        for (int r = 0; r < R; ++r) k6sm[LS::oU + 2 + r * TT + t] = ug[r * TT + t];
        ub = k6sm + LS::oU + 2;
    }
    if (t == 0 && tile > 0) k6sm[LS::oU + 1] = e[2 * (tile - 1) + 1];
    if (t == TT - 1 && tile + 1 < pd.ntiles) k6sm[LS::oU + 2 + R * TT] = e[2 * (tile + 1)];
    if constexpr (!SYM0) __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) { uc[r][0] = ub[r * TT + t]; bc[r][0] = use_b ? bb[r * TT + t] : 0.0; }
    ChunkThomas<1, R> th;
    assemble_chunk<Fn, R, TT, SYM0, true, !SYM0>(pd, mesh, ub, bb, pd.mass + (size_t)k * 4, t, tt, inv_tau,
                                                 pd.prm + (size_t)k * pd.nprm_s, uc, bc, th, (int)toff,
                                                 e + 2 * pd.ntiles, e + 2 * pd.ntiles + 1);
    Vec<1> xp, xs, h[R];
    separators(up, k, (size_t)tile * TT + t, xp, xs);
    th.back(h, xp, xs);
    double ss = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool in = (int)toff + t * R + r < pd.Nx;
        ug[r * TT + t] = in ? uc[r][0] - h[r].a[0] : 0.0;
        if (in) ss += h[r].a[0] * h[r].a[0];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_down_sync(0xffffffffu, ss, o);
    if ((t & 31) == 0) red[t >> 5] = ss;
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int w = 1; w < TT / 32; ++w) ss += red[w];
        partial[(size_t)k * pd.ntiles + tile] = ss;
    }
}

// residual-only / residual+Jacobian of a long mesh into F (and Jl, Jd, Ju): assemble! and the split-design API
template <class Fn, int R, int TT, bool SYM0, bool JAC>
__global__ void __launch_bounds__(TT) k6_assemble(ProbDev pd, double tt, double inv_tau, int all_instances) {
    static_assert(Fn::npde == 1, "long-mesh path: scalar PDEs only");
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    if (!all_instances && !pd.active[k]) return;
    const size_t toff = (size_t)tile * R * TT;
    const double* ui = pd.u + (size_t)k * pd.L;
    const double* ub = ui + toff;
    const double* bb = pd.b + (size_t)k * pd.L + toff;
    double uc[R][1], bc[R][1];
#pragma unroll
    for (int r = 0; r < R; ++r) { uc[r][0] = ub[r * TT + t]; bc[r][0] = JAC ? bb[r * TT + t] : 0.0; }
    ProbDev pt = pd;  // StoreSink addresses k*L + r*TT + t: shift the output arrays to the tile
    pt.F = pd.F + toff;
    if (JAC) { pt.Jl = pd.Jl + toff; pt.Jd = pd.Jd + toff; pt.Ju = pd.Ju + toff; }
    StoreSink<1, TT> sink{pt, k, t};
    const MeshDev mesh = mesh_of_tile(pd.mesh, toff);
    assemble_chunk<Fn, R, TT, SYM0, JAC, true>(pd, mesh, ub, bb, pd.mass + (size_t)k * 4, t, tt, inv_tau,
                                               pd.prm + (size_t)k * pd.nprm_s, uc, bc, sink, (int)toff, ui,
                                               ui + tiled_pos(pd.Nx - 1, R, TT));
}

// ||h||_2 per instance (fixed summation order: thread, lane tree, warps), stop test, counters; one CTA per instance,
// the last one publishes the active count
template <int UNUSED = 0>
__global__ void __launch_bounds__(256) k6_finalize(ProbDev pd, const double* partial, double tol, int slot) {
    __shared__ double red[8];
    const int64_t k = blockIdx.x;
    const int t = threadIdx.x;
    if (pd.active[k]) {
        const double* pk = partial + (size_t)k * pd.ntiles;
        double ss = 0.0;
        for (int i = t; i < pd.ntiles; i += 256) ss += pk[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_down_sync(0xffffffffu, ss, o);
        if ((t & 31) == 0) red[t >> 5] = ss;
        __syncthreads();
        if (t == 0) {
#pragma unroll
            for (int w = 1; w < 8; ++w) ss += red[w];
            finalize_system(pd, k, ss, tol, slot);
        }
    }
    if (t == 0) group_done(pd, slot, gridDim.x);
}

}  // namespace sb
