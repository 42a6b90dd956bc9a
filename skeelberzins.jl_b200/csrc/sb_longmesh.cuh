// sb_longmesh.cuh -- K6: Newton iteration on meshes too long for one CTA (e.g. 2^22 nodes, one instance).
//
// The mesh is stored as tiles of R*TT nodes, each tile in the chunk-transposed layout of the batched path,
// so an instance looks like a batch of coupled tiles.  One Newton iteration is a recursive partition
// ("SPIKE") solve over levels:
//
//   level 0   k6_chunk_asm   every thread assembles its R rows on the fly (residual + Jacobian by duals, nothing
//                            but u, b and the mesh weights is read) and eliminates them -> 7 numbers per chunk
//   level l   k6_chunk_sys   forms its rows (one per chunk of level l-1) from those numbers, keeps them for the way
//                            back, eliminates them                                        -> level l+1 ...
//   top       k6_top         a single CTA forms and solves the last level (<= 4096 rows) with the batched-path solver
//   level l   k6_back_sys    recompute the elimination, back-substitute with the separator values from level l+1
//   level 0   k6_back_asm    recompute assembly + elimination, back-substitute, u -= h in place, per-tile sum of h^2
//             k6_finalize    ||h||_2 per instance in a fixed summation order, stop test, counters, host notice
//
// Recomputing instead of storing the block rows keeps the traffic at level 0 to two reads of (u, b, weights) and
// one write of u per Newton iteration.  Scalar PDEs (npde = 1) only.
#pragma once
#include "sb_kernels.cuh"
#include "sb_resident.cuh"

namespace sb {

constexpr int kLongR = 8, kLongT0 = 64;   // level 0: rows per thread / threads per tile (tile = 512 mesh nodes; measured: 128 -> 137 us, 64 -> 134 us, 32 -> 136 us per cfg5 iteration)
constexpr int kLongTT = 256;              // stored levels and the top level: threads per tile
constexpr int kLongRS = 16;               // stored levels (1 .. top-1): rows per thread = reduction factor per level

struct LevelDev {      // one level of the recursive partition (level 0 has no stored rows: they are assembled)
    int n;             // rows
    int R;             // rows per thread in this level's tiled layout (level 0: kLongR, top: 1..16, else kLongRS)
    int ntiles;        // tiles of R * (kLongT0 | kLongTT) rows
    int nchunks;       // chunks of the level = rows of the next one (padding chunks included)
    size_t stride;     // padded rows per instance (= ntiles * R * kLongTT)
    double2* acd;      // levels 1 .. top-1: normalised rows (unit diagonal) [B][stride][(a, c), (d, -)], kept for the
                       // backward sweep
    double* x;         // levels >= 1: solution of the level [B][stride]
    double2* sum;      // levels < top: [B][up_stride][(y0, v0), (w0, ar), (bp, Cs), (dp, -)] chunk summaries, each
                       // stored at the position of the row it becomes in the next level's tiled layout.  One 64-byte
                       // record per chunk (two full sectors), so the scattered positions cost no write amplification
    size_t up_stride;  // stride of the next level
    int up_R;          // R of the next level
    double* edge;      // levels 0 and 1 only: [B][2*ntiles0 + 2] iterate at the first / last node of every level-0 tile
                       // and at mesh nodes 0 and Nx-1, saved by the forward sweep so that the backward sweep can
                       // update u in place (a tile's neighbours may already hold the new iterate)
};

SB_D size_t tiled_pos(int j, int R, int TT) {
    const int Lt = R * TT;
    const int tile = j / Lt, il = j - tile * Lt;
    return (size_t)tile * Lt + (size_t)(il % R) * TT + il / R;
}

// summary of chunk c of level `lv`: the first-row representation (y0, v0, w0) and the separator row with the
// chunk's own interior eliminated (ar, bp, Cs, dp); the next level combines it with the next chunk into a reduced row
template <int R>
SB_D void store_summary(const ChunkThomas<1, R>& th, const LevelDev& lv, int64_t k, int c) {
    Vec<1> y0, yL; Mat<1> v0, w0, vL, wL;
    th.summary(y0, v0, w0, yL, vL, wL);
    const double As = th.As.a[0];
    double2* S = lv.sum + ((size_t)k * lv.up_stride + tiled_pos(c, lv.up_R, kLongTT)) * 4;
    S[0] = make_double2(y0.a[0], v0.a[0]);
    S[1] = make_double2(w0.a[0], -(As * vL.a[0]));
    S[2] = make_double2(fma(-As, wL.a[0], th.Bs.a[0]), th.Cs.a[0]);
    S[3] = make_double2(fma(-As, yL.a[0], th.ds.a[0]), 0.0);
}

// rows j0 .. j0+R-1 of the level above `below` (thread t of tile `tile` in that level's (R, TT) layout) from the
// summaries of chunks j and j+1 of `below`; rows beyond the chunks of `below` are identity rows.  Branch-free: the
// records of padding rows lie inside the zero-filled array and their results are discarded.
template <int R, int TT>
SB_D void form_rows(const LevelDev& below, int64_t k, int tile, int t, double (&ra)[R], double (&rc)[R], double (&rd)[R]) {
    const int n = below.nchunks;
    const double2* S = below.sum + (size_t)k * below.up_stride * 4;
    const int j0 = (tile * TT + t) * R;
    const size_t p0 = (size_t)tile * R * TT + t;
    // record after the thread's last one: only its first-row representation is needed
    const size_t pl = (t + 1 < TT) ? (size_t)tile * R * TT + (t + 1) : (size_t)(tile + 1) * R * TT;
    double2 q[R + 1][2];
#pragma unroll
    for (int r = 0; r <= R; ++r) {
        const bool in = j0 + r < n;
        const size_t p = (r < R) ? p0 + (size_t)r * TT : (in ? pl : p0);
        q[r][0] = __ldcg(S + p * 4 + 0); q[r][1] = __ldcg(S + p * 4 + 1);   // rewritten every iteration: L2, not the read-only path
        if (!in) { q[r][0] = make_double2(0.0, 0.0); q[r][1].x = 0.0; }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const size_t p = p0 + (size_t)r * TT;
        const double2 bc = __ldcg(S + p * 4 + 2), dd = __ldcg(S + p * 4 + 3);
        const double ar = q[r][1].y, bp = bc.x, Cs = bc.y, dp = dd.x;
        const double y0n = q[r + 1][0].x, v0n = q[r + 1][0].y, w0n = q[r + 1][1].x;
        const double binv = fast_rcp(fma(-Cs, v0n, bp));
        const bool in = j0 + r < n;
        ra[r] = in ? binv * ar : 0.0;
        rc[r] = in ? -(binv * (Cs * w0n)) : 0.0;
        rd[r] = in ? binv * fma(-Cs, y0n, dp) : 0.0;
    }
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

// Part 1, before pdl_wait(): barriers and the mesh weights, which no kernel ever writes -- this overlaps the tail of
// the previous kernel in the stream.  Part 2, after it: u and b of the tile.
template <class Fn, int R, int TT>
SB_D void long_stage_weights(const ProbDev& pd, int tile, double* sm, MeshDev& mesh) {
    using LS = LongStage<Fn, R, TT, true>;
    constexpr int L = LS::L;
    constexpr uint32_t LB = (uint32_t)(L * sizeof(double));
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm);   // bar[0]: weights, bar[1]: state
    const size_t toff = (size_t)tile * L;
    if (threadIdx.x == 0) {
        mbar_init(bar + 0, 1);
        mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int lo = (tile > 0) ? 2 : 0;  // halo towards an existing neighbour
        uint32_t bytes = (uint32_t)((L + lo) * sizeof(double));
        if (Fn::c_unit) bytes += LB;
        if (LS::kFrac) bytes += 2 * LB;
        mbar_expect_tx(bar, bytes);
        bulk_g2s(sm + LS::oGw + 2 - lo, pd.mesh.gw + toff - lo, (uint32_t)((L + lo) * sizeof(double)), bar);
        if constexpr (Fn::c_unit) bulk_g2s(sm + LS::oIvol, pd.mesh.ivol + toff, LB, bar);
        if constexpr (LS::kFrac) bulk_g2s(sm + LS::oFrac, pd.mesh.frac + toff, 2 * LB, bar);
    }
    mesh = mesh_of_tile(pd.mesh, toff);  // xi (read only by functors that use x) stays in global memory
    mesh.gw = sm + LS::oGw + 2;
    if constexpr (Fn::c_unit) mesh.ivol = sm + LS::oIvol;
    if constexpr (LS::kFrac) mesh.frac = reinterpret_cast<const double2*>(sm + LS::oFrac);
}

// a CTA that turns out to have nothing to do must not exit with the weight copies in flight
SB_D void long_stage_drain(double* sm) {
    if (threadIdx.x == 0) mbar_wait(reinterpret_cast<uint64_t*>(sm), 0);
}

template <class Fn, int R, int TT>
SB_D void long_stage_state(const ProbDev& pd, int64_t k, int tile, bool use_b, bool halo_u, double* sm,
                           const double*& ub, const double*& bb) {
    using LS = LongStage<Fn, R, TT, true>;
    constexpr int L = LS::L;
    constexpr uint32_t LB = (uint32_t)(L * sizeof(double));
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
    const size_t toff = (size_t)tile * L;
    if (threadIdx.x == 0) {
        const int ulo = (halo_u && tile > 0) ? 2 : 0, uhi = (halo_u && tile + 1 < pd.ntiles) ? 2 : 0;
        const uint32_t ubytes = (uint32_t)((L + ulo + uhi) * sizeof(double));
        mbar_expect_tx(bar + 1, ubytes + (use_b ? LB : 0u));
        bulk_g2s(sm + LS::oU + 2 - ulo, pd.u + (size_t)k * pd.L + toff - ulo, ubytes, bar + 1);
        if (use_b) bulk_g2s(sm + LS::oB, pd.b + (size_t)k * pd.L + toff, LB, bar + 1);
    }
    ub = sm + LS::oU + 2;
    bb = use_b ? sm + LS::oB : ub;       // tau = Inf: b is multiplied by 1/tau = 0 wherever it is still named
    mbar_wait(bar + 1, 0);
    mbar_wait(bar + 0, 0);
}

// level 0, forward: assemble + eliminate the chunk, store its summary
template <class Fn, int R, int TT, bool SYM0>
__global__ void __launch_bounds__(TT) k6_chunk_asm(ProbDev pd, double tt, double inv_tau, LevelDev lv) {
    static_assert(Fn::npde == 1, "long-mesh path: scalar PDEs only");
    extern __shared__ __align__(16) double k6sm[];
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    const size_t toff = (size_t)tile * R * TT;
    MeshDev mesh;
    pdl_release();
    if constexpr (SYM0) long_stage_weights<Fn, R, TT>(pd, tile, k6sm, mesh);
    else mesh = mesh_of_tile(pd.mesh, toff);
    pdl_wait();
    if (!pd.active[k]) {
        if constexpr (SYM0) long_stage_drain(k6sm);
        return;
    }
    const double* ui = pd.u + (size_t)k * pd.L;
    const double* ub = ui + toff;
    const double* bb = pd.b + (size_t)k * pd.L + toff;
    const bool use_b = inv_tau != 0.0;  // stationary solves (tau = Inf) never look at b: rhs - b/tau = rhs
    if constexpr (SYM0) long_stage_state<Fn, R, TT>(pd, k, tile, use_b, true, k6sm, ub, bb);
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
    store_summary<R>(th, lv, k, tile * TT + t);
}

// level l >= 1, forward: form the level's rows, keep them, eliminate, store the chunk summary
template <int R, int TT>
__global__ void __launch_bounds__(TT) k6_chunk_sys(LevelDev lv, LevelDev below, const int* active) {
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    pdl_release();
    pdl_wait();
    if (!active[k]) return;
    double ra[R], rc[R], rd[R];
    form_rows<R, TT>(below, k, tile, t, ra, rc, rd);
    ChunkThomas<1, R> th;
    double2* rows = lv.acd + ((size_t)k * lv.stride + (size_t)tile * R * TT + t) * 2;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        rows[(size_t)r * TT * 2 + 0] = make_double2(ra[r], rc[r]);
        rows[(size_t)r * TT * 2 + 1] = make_double2(rd[r], 0.0);
        Mat<1> A, Bm, C; Vec<1> d;
        A.a[0] = ra[r]; Bm.a[0] = 1.0; C.a[0] = rc[r]; d.a[0] = rd[r];
        th.row(r, A, Bm, C, d);
    }
    store_summary<R>(th, lv, k, tile * TT + t);
}

// top level: one CTA per instance forms its rows (RT per thread, one tile) and solves them
template <int RT, int TT>
__global__ void __launch_bounds__(TT) k6_top(LevelDev lv, LevelDev below, const int* active) {
    extern __shared__ double sm[];
    const int t = threadIdx.x;
    const int64_t k = blockIdx.x;
    pdl_release();
    pdl_wait();
    if (!active[k]) return;
    double ra[RT], rc[RT], rd[RT];
    form_rows<RT, TT>(below, k, 0, t, ra, rc, rd);
    ChunkThomas<1, RT> th;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        Mat<1> A, Bm, C; Vec<1> d;
        A.a[0] = ra[r]; Bm.a[0] = 1.0; C.a[0] = rc[r]; d.a[0] = rd[r];
        th.row(r, A, Bm, C, d);
    }
    Vec<1> x[RT];
    th.template finish<TT>(x, sm, t);
    double* xo = lv.x + (size_t)k * lv.stride;
#pragma unroll
    for (int r = 0; r < RT; ++r) xo[(size_t)r * TT + t] = x[r].a[0];
}

// separator unknowns of chunk c (own and previous) from the solution of the next level
SB_D void separators(const LevelDev& up, int64_t k, int c, Vec<1>& xp, Vec<1>& xs) {
    const double* xu = up.x + (size_t)k * up.stride;
    xs.a[0] = __ldcg(xu + tiled_pos(c, up.R, kLongTT));
    xp.a[0] = (c > 0) ? __ldcg(xu + tiled_pos(c - 1, up.R, kLongTT)) : 0.0;
}

// level l >= 1, backward
template <int R, int TT>
__global__ void __launch_bounds__(TT) k6_back_sys(LevelDev lv, LevelDev up, const int* active) {
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    pdl_release();
    pdl_wait();
    if (!active[k]) return;
    ChunkThomas<1, R> th;
    const size_t base = (size_t)k * lv.stride + (size_t)tile * R * TT;
    const double2* rows = lv.acd + (base + t) * 2;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double2 ac = __ldcg(rows + (size_t)r * TT * 2), dd = __ldcg(rows + (size_t)r * TT * 2 + 1);   // written by k6_chunk_sys of this iteration
        Mat<1> A, Bm, C; Vec<1> d;
        A.a[0] = ac.x; Bm.a[0] = 1.0; C.a[0] = ac.y; d.a[0] = dd.x;
        th.row(r, A, Bm, C, d);
    }
    Vec<1> xp, xs, x[R];
    separators(up, k, tile * TT + t, xp, xs);
    th.back(x, xp, xs);
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
    const size_t toff = (size_t)tile * R * TT;
    using LS = LongStage<Fn, R, TT, SYM0>;
    MeshDev mesh;
    pdl_release();
    if constexpr (SYM0) long_stage_weights<Fn, R, TT>(pd, tile, k6sm, mesh);
    else mesh = mesh_of_tile(pd.mesh, toff);
    pdl_wait();
    if (!pd.active[k]) {
        if constexpr (SYM0) long_stage_drain(k6sm);
        return;
    }
    double* ug = pd.u + (size_t)k * pd.L + toff;
    const double* ub = ug;
    const double* bb = pd.b + (size_t)k * pd.L + toff;
    const bool use_b = inv_tau != 0.0;  // stationary solves (tau = Inf) never look at b: rhs - b/tau = rhs
    // u is updated in place while other tiles are still being read: everything this CTA needs from outside its own
    // tile (the neighbours' edge nodes, both mesh ends for the boundary conditions) comes from the copy the
    // forward sweep saved in `edge`.
    const double* e = up.edge + (size_t)k * (2 * pd.ntiles + 2);
    double uc[R][1], bc[R][1];
    if constexpr (SYM0) {
        long_stage_state<Fn, R, TT>(pd, k, tile, use_b, false, k6sm, ub, bb);
    } else {
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
    separators(up, k, tile * TT + t, xp, xs);
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
// the last one publishes the active count.  (Folding this into k6_back_asm with a per-instance ticket was measured:
// the fence + returning atomic at the end of every tile's CTA costs more than this launch.)
template <int UNUSED = 0>
__global__ void __launch_bounds__(256) k6_finalize(ProbDev pd, const double* partial, double tol, int slot) {
    __shared__ double red[8];
    const int64_t k = blockIdx.x;
    const int t = threadIdx.x;
    pdl_release();
    pdl_wait();
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
            finalize_system(pd, k, ss, tol, slot, pd.iter_no);
        }
    }
    if (t == 0) group_done(pd, slot, gridDim.x);
}

}  // namespace sb
