// sb_longmesh.cuh -- K6: Newton iteration on meshes too long for one CTA (e.g. 2^22 nodes, one instance).
//
// The mesh is stored as tiles of R*TT nodes, each tile in the chunk-transposed layout of the batched path,
// so an instance looks like a batch of coupled tiles.  One Newton iteration is a recursive partition
// ("SPIKE") solve over levels.  Scalar PDEs (the tuned path, four launches per iteration for up to 4096 tiles):
//
//   level 0   k6_chunk_asm   every thread assembles its R rows on the fly (residual + Jacobian by duals; m = 0: nothing
//                            but u, b and the mesh points x is read, every weight is computed from x) and eliminates
//                            them; the TT chunks of the tile are reduced inside the CTA (warp 0: TT/32 reduced rows per
//                            lane, then an open cyclic reduction across its lanes) -> ONE row of the level above per
//                            tile, and per chunk the three numbers that give its separator once the tile's two outer
//                            separators are known
//   top       k6_top         a single CTA forms and solves the last level (<= 4096 rows) with the batched-path solver
//   level 0   k6_back_asm    recompute assembly + elimination, separators from those three numbers, back-substitute,
//                            u -= h in place, per-tile sum of h^2 (no reduction, no barrier before the update)
//             k6_finalize    ||h||_2 per instance in a fixed summation order, stop test, counters, host notice
//
// More tiles than the top level takes go through stored levels in between (k6_chunk_sys forms its rows, one per chunk of
// the level below, keeps them for the way back and eliminates them; k6_back_sys recomputes the elimination and
// back-substitutes).  Recomputing instead of storing the block rows keeps the traffic at level 0 to two reads of (u, b, x)
// and one write of u per Newton iteration.  Systems (npde = 2, 3, 4): the block version at the end of this file (k6s_*),
// one row of the level above per THREAD chunk, weights from the arrays.
#pragma once
#include "sb_kernels.cuh"
#include "sb_resident.cuh"

namespace sb {

#ifndef SB_LONG_T0
#define SB_LONG_T0 128
#endif
constexpr int kLongR = 8, kLongT0 = SB_LONG_T0;   // level 0: rows per thread / threads per tile (tile = 1024 mesh nodes = one row of the level above; measured: 256 threads 87 us, 128 threads 80 us per cfg5 iteration)
// resident warps per SM the two level-0 kernels are compiled for (register caps 96 / 128): measured on cfg5 at 16 / 20 / 24
// warps -- forward sweep 35.2 / 33.9 / 35.2 us, backward sweep 32.2 / 33.6 / 39.2 us
#ifndef SB_LONG_WARPS_FWD
#define SB_LONG_WARPS_FWD 20
#endif
#ifndef SB_LONG_WARPS_BWD
#define SB_LONG_WARPS_BWD 16
#endif
constexpr int kLongCtasFwd = 32 * SB_LONG_WARPS_FWD / kLongT0, kLongCtasBwd = 32 * SB_LONG_WARPS_BWD / kLongT0;
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
    double* coef;      // level 0 only: [B][ntiles][3][T0] every thread chunk's separator as a function of its tile's two
                       // outer separators, s_j = D_j - A_j*S_prev - C_j*S_tile, left by the forward sweep so that the
                       // backward sweep needs no reduction (and no barrier) of its own
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
    r.xi += off; r.awbw += off; r.gwwf += off; r.frac += off; r.gw += off; r.ivol += off; r.xs += off;
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// Open reduction of a group of N <= 32 consecutive chunks held one per lane (every lane of the warp calls; lanes >= N
// carry nothing).  A chunk is known by its first-row representation and its separator row with the interior eliminated:
//     x_first(l) = y0 - v0*s_{l-1} - w0*s_l            ar*s_{l-1} + bp*s_l + cs*x_first(l+1) = dp
// (s_l = unknown of the chunk's last row; s_{-1} = S_prev, the separator of whatever precedes the group).  The group's
// last separator S_last = s_{N-1} stays an unknown: rows 0 .. N-2 are reduced by parallel cyclic reduction in which an
// elimination step that would reach outside the group is skipped, so the coefficient that is left multiplies S_prev /
// S_last ("spikes" at no cost).  Afterwards
//     s_l = dn - an*S_prev - cn*S_last                                                      (lanes l <= N-2)
// and the group as a whole is a chunk of the same form -- (Y0, V0, W0) valid in lane 0, (AR, BP, DP) in lane N-1, cs
// unchanged -- which is what the next level up consumes.  The backward sweep runs the same reduction again and turns
// (S_prev, S_last) into the s_l.  Scalar rows (npde = 1).
// ---------------------------------------------------------------------------------------------------------------------
struct OpenRows { double an, cn, dn; };

template <int N>
SB_D OpenRows open_reduce(int lane, double y0, double v0, double w0, double ar, double bp, double cs, double dp,
                          double& Y0, double& V0, double& W0, double& AR, double& BP, double& DP) {
    constexpr unsigned FULL = 0xffffffffu;
    double an = 0.0, cn = 0.0, dn = 0.0;
    if constexpr (N > 1) {
        // reduced row of lane l < N-1: the first row of chunk l+1 is eliminated with that chunk's representation
        const double y0n = __shfl_down_sync(FULL, y0, 1), v0n = __shfl_down_sync(FULL, v0, 1), w0n = __shfl_down_sync(FULL, w0, 1);
        if (lane < N - 1) {
            const double binv = fast_rcp(fma(-cs, v0n, bp));
            an = binv * ar; cn = -(binv * (cs * w0n)); dn = binv * fma(-cs, y0n, dp);
        }
#pragma unroll
        for (int s = 1; s < N - 1; s <<= 1) {
            const double am = __shfl_up_sync(FULL, an, s), cm = __shfl_up_sync(FULL, cn, s), dm = __shfl_up_sync(FULL, dn, s);
            const double ap = __shfl_down_sync(FULL, an, s), cp = __shfl_down_sync(FULL, cn, s), dq = __shfl_down_sync(FULL, dn, s);
            const bool hu = lane >= s, hd = lane + s <= N - 2;   // the partner row exists inside the group
            const double anm = hu ? an : 0.0, cnp = hd ? cn : 0.0;
            const double b2 = fma(-cnp, ap, fma(-anm, cm, 1.0));
            const double d2 = fma(-cnp, dq, fma(-anm, dm, dn));
            const double a2 = hu ? -(an * am) : an, c2 = hd ? -(cn * cp) : cn;
            const double inv = fast_rcp(b2);
            an = inv * a2; cn = inv * c2; dn = inv * d2;     // (lanes >= N-1 compute on, nobody reads them)
        }
    }
    // the group as one chunk
    Y0 = fma(-w0, dn, y0); V0 = fma(-w0, an, v0); W0 = -(w0 * cn);
    if constexpr (N > 1) {
        const double al = __shfl_up_sync(FULL, an, 1), cl = __shfl_up_sync(FULL, cn, 1), dl = __shfl_up_sync(FULL, dn, 1);
        AR = -(ar * al); BP = fma(-ar, cl, bp); DP = fma(-ar, dl, dp);
    } else {
        Y0 = y0; V0 = v0; W0 = w0; AR = ar; BP = bp; DP = dp;
    }
    OpenRows o; o.an = an; o.cn = cn; o.dn = dn;
    return o;
}

// separator of lane l (and of lane l-1) from the group's outer values
template <int N>
SB_D void open_back(int lane, const OpenRows& o, double S_prev, double S_last, double& xp, double& xs) {
    xs = (lane < N - 1) ? fma(-o.cn, S_last, fma(-o.an, S_prev, o.dn)) : S_last;
    const double up = __shfl_up_sync(0xffffffffu, xs, 1);
    xp = (lane > 0) ? up : S_prev;
}

// Shared memory of the level-0 kernels.  m == 0: u (with the edge nodes of the neighbouring tiles), b and the mesh
// points x of the tile arrive by bulk copies behind mbarriers, so a sweep pays the DRAM latency once instead of once per
// row, and gw / ivol / frac are computed from x (8 instead of 32 bytes per node and sweep).  m >= 1 needs logarithms and
// cube roots for that: those kernels keep reading the weight arrays from global memory and stage u only.
template <class Fn, int R, int TT, bool SYM0>
struct LongStage {
    static constexpr bool on = SYM0;
    static constexpr int L = R * TT, W = TT / 32;
    static constexpr int oBar = 0;                                // two mbarriers (weights | state)
    static constexpr int oRed = 2;                                // [W][8] warp summaries / [W + 2] separators on the way back
    static constexpr int oU = oRed + 8 * W;                       // [L + 4]: node slot s at oU + 2 + s, s = -2 .. L+1
    static constexpr int oB = oU + L + 4;                         // [L]
    static constexpr int oX = oB + L;                             // [L + 4]: mesh point of slot s at oX + 2 + s
    static constexpr int doubles = oX + L + 4;
    static_assert(oU % 2 == 0 && oB % 2 == 0 && oX % 2 == 0, "bulk copies need 16-byte aligned destinations");
    static constexpr size_t bytes = (on ? doubles : oB) * sizeof(double);   // m >= 1: u only
};

// Part 1, before pdl_wait(): barriers and the mesh points, which no kernel ever writes -- this overlaps the tail of
// the previous kernel in the stream.  Part 2, after it: u and b of the tile.
template <class Fn, int R, int TT>
SB_D void long_stage_mesh(const ProbDev& pd, int tile, double* sm, MeshDev& mesh) {
    using LS = LongStage<Fn, R, TT, true>;
    constexpr int L = LS::L;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + LS::oBar);   // bar[0]: mesh points, bar[1]: state
    const size_t toff = (size_t)tile * L;
    if (threadIdx.x == 0) {
        mbar_init(bar + 0, 1);
        mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int lo = (tile > 0) ? 2 : 0, hi = (tile + 1 < pd.ntiles) ? 2 : 0;  // halo towards an existing neighbour
        const uint32_t bytes = (uint32_t)((L + lo + hi) * sizeof(double));
        mbar_expect_tx(bar, bytes);
        bulk_g2s(sm + LS::oX + 2 - lo, pd.mesh.xs + toff - lo, bytes, bar);
    }
    mesh = mesh_of_tile(pd.mesh, toff);  // (the weight arrays are not read by the m == 0 kernels)
    mesh.xs = sm + LS::oX + 2;
}

// a CTA that turns out to have nothing to do must not exit with its copies in flight
SB_D void long_stage_drain(double* sm) {
    if (threadIdx.x == 0) {
        mbar_wait(reinterpret_cast<uint64_t*>(sm), 0);
        mbar_wait(reinterpret_cast<uint64_t*>(sm) + 1, 0);
    }
}

// u and b of the tile: issued right after pdl_wait(), before anything that depends on a global load, so that the
// latencies of the copy and of the per-instance loads (active flag, parameters, mass flags) overlap
template <class Fn, int R, int TT>
SB_D void long_stage_state_issue(const ProbDev& pd, int64_t k, int tile, bool use_b, bool halo_u, double* sm) {
    using LS = LongStage<Fn, R, TT, true>;
    constexpr int L = LS::L;
    constexpr uint32_t LB = (uint32_t)(L * sizeof(double));
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + LS::oBar);
    const size_t toff = (size_t)tile * L;
    if (threadIdx.x == 0) {
        const int ulo = (halo_u && tile > 0) ? 2 : 0, uhi = (halo_u && tile + 1 < pd.ntiles) ? 2 : 0;
        const uint32_t ubytes = (uint32_t)((L + ulo + uhi) * sizeof(double));
        mbar_expect_tx(bar + 1, ubytes + (use_b ? LB : 0u));
        bulk_g2s(sm + LS::oU + 2 - ulo, pd.u + (size_t)k * pd.L + toff - ulo, ubytes, bar + 1);
        if (use_b) bulk_g2s(sm + LS::oB, pd.b + (size_t)k * pd.L + toff, LB, bar + 1);
    }
}
template <class Fn, int R, int TT>
SB_D void long_stage_state_wait(bool use_b, double* sm, const double*& ub, const double*& bb) {
    using LS = LongStage<Fn, R, TT, true>;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + LS::oBar);
    ub = sm + LS::oU + 2;
    bb = use_b ? sm + LS::oB : ub;       // tau = Inf: b is multiplied by 1/tau = 0 wherever it is still named
    mbar_wait(bar + 1, 0);
    mbar_wait(bar + 0, 0);
}

// parameters and mass flags of the instance -> registers (asked for early: see long_stage_state_issue)
template <class Fn>
struct InstanceRegs {
    static constexpr int NP = Fn::nparams > 0 ? Fn::nparams : 1;
    double prm[NP], mass[4];
    int active;
    SB_D void load(const ProbDev& pd, int64_t k) {
        active = pd.active[k];
#pragma unroll
        for (int i = 0; i < Fn::nparams; ++i) prm[i] = pd.prm[(size_t)k * pd.nprm_s + i];
#pragma unroll
        for (int i = 0; i < 4; ++i) mass[i] = pd.mass[(size_t)k * 4 + i];
    }
};

// thread level -> the chunk's 7 numbers (what store_summary writes for the stored levels)
template <int R>
SB_D void chunk_numbers(const ChunkThomas<1, R>& th, double& y0, double& v0, double& w0, double& ar, double& bp, double& cs, double& dp) {
    Vec<1> y0v, yL; Mat<1> v0m, w0m, vL, wL;
    th.summary(y0v, v0m, w0m, yL, vL, wL);
    const double As = th.As.a[0];
    y0 = y0v.a[0]; v0 = v0m.a[0]; w0 = w0m.a[0];
    ar = -(As * vL.a[0]); bp = fma(-As, wL.a[0], th.Bs.a[0]); cs = th.Cs.a[0]; dp = fma(-As, yL.a[0], th.ds.a[0]);
}

// Reduction of the TT chunks of a tile inside its CTA.  Every thread leaves its chunk's 7 numbers in shared memory; warp 0
// alone turns them into reduced rows (as form_rows does for a stored level), TT/32 per lane, eliminates those with the
// same partition (lvl2) and finishes with the open reduction across its 32 lanes.  The other warps wait at the barrier
// -- their issue slots go to the other CTAs of the SM (an all-warps variant, open_reduce in every warp and then across
// the warps, was measured slower: more instructions, the same waiting).
template <int TT>
struct TileScratch {
    static constexpr int Q = TT / 32;
    static constexpr int S = TT + TT / 8 + 2;      // record j at pos(j) = j + j/8: lane q reads 8q+i -> stride 9, no bank conflicts
    double rec[7][S];                              // y0, v0, w0, ar, bp, cs, dp
    double sep[S];                                 // backward: [0] the separator before the tile, [pos(j) + 1] thread j's
    static SB_D int pos(int j) { return j + (j >> 3); }
};

template <int TT>
SB_D OpenRows tile_reduce_warp0(const TileScratch<TT>& ts, int lane, ChunkThomas<1, TT / 32>& lvl2, double& Y0, double& V0,
                                double& W0, double& AR, double& BP, double& CS, double& DP) {
    using TS = TileScratch<TT>;
    constexpr int Q = TS::Q;
    static_assert(Q >= 2, "tiles of at least 64 threads");
    const int j0 = lane * Q;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const int p = TS::pos(j0 + i), pn = TS::pos(j0 + i + 1);
        const double ar = ts.rec[3][p], bp = ts.rec[4][p], cs = ts.rec[5][p], dp = ts.rec[6][p];
        const double y0n = ts.rec[0][pn], v0n = ts.rec[1][pn], w0n = ts.rec[2][pn];   // chunk j+1: its first row is eliminated
        const double binv = fast_rcp(fma(-cs, v0n, bp));
        Mat<1> A, Bm, C; Vec<1> d;
        A.a[0] = binv * ar; Bm.a[0] = 1.0; C.a[0] = -(binv * (cs * w0n)); d.a[0] = binv * fma(-cs, y0n, dp);
        if (i == Q - 1 && lane == 31) {            // the tile's last chunk: its successor lives in the next tile
            A.a[0] = ar; Bm.a[0] = bp; C.a[0] = cs; d.a[0] = dp;
        }
        lvl2.row(i, A, Bm, C, d);
    }
    double y0, v0, w0, ar, bp, cs, dp, Y1, V1, W1;
    chunk_numbers<Q>(lvl2, y0, v0, w0, ar, bp, cs, dp);
    const OpenRows rows = open_reduce<32>(lane, y0, v0, w0, ar, bp, cs, dp, Y1, V1, W1, AR, BP, DP);
    CS = cs;
    // lane 0: (Y1, V1, W1) represent the tile's first separator; the tile's first mesh row follows with thread 0's numbers
    const double ty0 = ts.rec[0][0], tv0 = ts.rec[1][0], tw0 = ts.rec[2][0];
    Y0 = fma(-tw0, Y1, ty0); V0 = fma(-tw0, V1, tv0); W0 = -(tw0 * W1);
    return rows;
}

// Every chunk's separator in terms of the tile's outer separators: s_j = D_j - A_j*S_prev - C_j*S_tile, j = lane*Q + i.
// (lvl2 in explicit form gives x_i = y_i - v_i*xp - w_i*xs inside the lane's chunk; xp, xs are the open rows of lanes
// lane-1 and lane.)  cf: the tile's [3][TT] block of LevelDev::coef.
template <int TT>
SB_D void tile_coefficients(ChunkThomas<1, TT / 32>& lvl2, const OpenRows& rows, int lane, double* cf) {
    constexpr int Q = TT / 32;
    constexpr unsigned FULL = 0xffffffffu;
    // xs = Xd - Xa*S_prev - Xc*S_tile (lane 31: S_tile itself), xp likewise from the lane before (lane 0: S_prev itself)
    const double Xd = (lane < 31) ? rows.dn : 0.0, Xa = (lane < 31) ? rows.an : 0.0, Xc = (lane < 31) ? rows.cn : -1.0;
    const double ud = __shfl_up_sync(FULL, Xd, 1), ua = __shfl_up_sync(FULL, Xa, 1), uc = __shfl_up_sync(FULL, Xc, 1);
    const double Pd = (lane > 0) ? ud : 0.0, Pa = (lane > 0) ? ua : -1.0, Pc = (lane > 0) ? uc : 0.0;
    lvl2.to_explicit();
    const int j0 = lane * Q;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        double D = Xd, A = Xa, C = Xc;
        if (i < Q - 1) {
            const double y = lvl2.dt[i].a[0], v = lvl2.vt[i].a[0], w = lvl2.ct[i].a[0];
            D = fma(-w, Xd, fma(-v, Pd, y));
            A = -fma(w, Xa, v * Pa);
            C = -fma(w, Xc, v * Pc);
        }
        cf[j0 + i] = D; cf[TT + j0 + i] = A; cf[2 * TT + j0 + i] = C;
    }
}

// level 0, forward: assemble + eliminate the chunk (thread), reduce the 32 chunks of a warp and the TT/32 warps of the
// CTA to ONE summary per tile of R*TT mesh nodes: the level above has Nx / (R*TT) rows
template <class Fn, int R, int TT, bool SYM0>
__global__ void __launch_bounds__(TT, kLongCtasFwd) k6_chunk_asm(ProbDev pd, double tt, double inv_tau, LevelDev lv) {
    static_assert(Fn::npde == 1, "long-mesh path: scalar PDEs only");
    using LS = LongStage<Fn, R, TT, SYM0>;
    using TS = TileScratch<TT>;
    extern __shared__ __align__(16) double k6sm[];
    __shared__ TS ts;
    const int tile = blockIdx.x, t = threadIdx.x, lane = t & 31, wi = t >> 5;
    const int64_t k = blockIdx.y;
    const size_t toff = (size_t)tile * R * TT;
    MeshDev mesh;
    pdl_release();
    if constexpr (SYM0) long_stage_mesh<Fn, R, TT>(pd, tile, k6sm, mesh);
    else mesh = mesh_of_tile(pd.mesh, toff);
    pdl_wait();
    const bool use_b = inv_tau != 0.0;  // stationary solves (tau = Inf) never look at b: rhs - b/tau = rhs
    if constexpr (SYM0) long_stage_state_issue<Fn, R, TT>(pd, k, tile, use_b, true, k6sm);
    InstanceRegs<Fn> in;
    in.load(pd, k);
    if (!in.active) {
        if constexpr (SYM0) long_stage_drain(k6sm);
        return;
    }
    const double* ui = pd.u + (size_t)k * pd.L;
    const double* ub = ui + toff;
    const double* bb = pd.b + (size_t)k * pd.L + toff;
    if constexpr (SYM0) long_stage_state_wait<Fn, R, TT>(use_b, k6sm, ub, bb);
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
    assemble_chunk<Fn, R, TT, SYM0, true, !SYM0, SYM0>(pd, mesh, ub, bb, in.mass, t, tt, inv_tau, in.prm, uc, bc, th, (int)toff, ui,
                                                       ui + tiled_pos(pd.Nx - 1, R, TT));
    double y0, v0, w0, ar, bp, cs, dp;
    chunk_numbers<R>(th, y0, v0, w0, ar, bp, cs, dp);
    {
        const int p = TS::pos(t);
        ts.rec[0][p] = y0; ts.rec[1][p] = v0; ts.rec[2][p] = w0; ts.rec[3][p] = ar; ts.rec[4][p] = bp; ts.rec[5][p] = cs; ts.rec[6][p] = dp;
    }
    __syncthreads();
    if (wi == 0) {
        ChunkThomas<1, TS::Q> lvl2;
        double Y0, V0, W0, AR, BP, CS, DP;
        const OpenRows rows = tile_reduce_warp0<TT>(ts, lane, lvl2, Y0, V0, W0, AR, BP, CS, DP);
        // one 64-byte record per tile, at the position of the row it becomes in the level above
        double* S = reinterpret_cast<double*>(lv.sum + ((size_t)k * lv.up_stride + tiled_pos(tile, lv.up_R, kLongTT)) * 4);
        if (lane == 0) { S[0] = Y0; S[1] = V0; S[2] = W0; }
        if (lane == 31) { S[3] = AR; S[4] = BP; S[5] = CS; S[6] = DP; S[7] = 0.0; }
        // and, for the backward sweep, what every chunk's separator will be once the tile's outer separators are known
        tile_coefficients<TT>(lvl2, rows, lane, lv.coef + ((size_t)k * pd.ntiles + tile) * 3 * TT);
    }
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

// level 0, backward: recompute assembly and elimination of the chunk, take the chunk's separators from the coefficients
// the forward sweep left and the tile's two outer separators from the level above, back-substitute, u -= h in place,
// per-tile sum of h^2.  No reduction and no barrier between the assembly and the update.
template <class Fn, int R, int TT, bool SYM0>
__global__ void __launch_bounds__(TT, kLongCtasBwd) k6_back_asm(ProbDev pd, double tt, double inv_tau, LevelDev up, double* partial) {
    using LS = LongStage<Fn, R, TT, SYM0>;
    constexpr int W = TT / 32;
    __shared__ double red[W];
    __shared__ double outer[2];          // the separator before the tile, the tile's own
    extern __shared__ __align__(16) double k6sm[];
    const int tile = blockIdx.x, t = threadIdx.x, lane = t & 31, wi = t >> 5;
    const int64_t k = blockIdx.y;
    const size_t toff = (size_t)tile * R * TT;
    MeshDev mesh;
    pdl_release();
    if constexpr (SYM0) long_stage_mesh<Fn, R, TT>(pd, tile, k6sm, mesh);
    else mesh = mesh_of_tile(pd.mesh, toff);
    pdl_wait();
    const bool use_b = inv_tau != 0.0;  // stationary solves (tau = Inf) never look at b: rhs - b/tau = rhs
    if constexpr (SYM0) long_stage_state_issue<Fn, R, TT>(pd, k, tile, use_b, false, k6sm);
    InstanceRegs<Fn> in;
    in.load(pd, k);
    if (!in.active) {
        if constexpr (SYM0) long_stage_drain(k6sm);
        return;
    }
    double* ug = pd.u + (size_t)k * pd.L + toff;
    const double* ub = ug;
    const double* bb = pd.b + (size_t)k * pd.L + toff;
    // u is updated in place while other tiles are still being read: everything this CTA needs from outside its own
    // tile (the neighbours' edge nodes, both mesh ends for the boundary conditions) comes from the copy the
    // forward sweep saved in `edge`.
    const double* e = up.edge + (size_t)k * (2 * pd.ntiles + 2);
    if (t == 32 % TT) {                  // (a thread that does not issue the copies)
        Vec<1> Sp, Ss;
        separators(up, k, tile, Sp, Ss);
        outer[0] = Sp.a[0]; outer[1] = Ss.a[0];
    }
    // chunk t and chunk t-1: s = D - A*S_prev - C*S_tile (written by this iteration's forward sweep: L2, not the read-only path)
    const double* cf = up.coef + ((size_t)k * pd.ntiles + tile) * 3 * TT;
    const int tm = (t > 0) ? t - 1 : 0;
    const double cD = __ldcg(cf + t), cA = __ldcg(cf + TT + t), cC = __ldcg(cf + 2 * TT + t);
    const double pD = __ldcg(cf + tm), pA = __ldcg(cf + TT + tm), pC = __ldcg(cf + 2 * TT + tm);
    double uc[R][1], bc[R][1];
    if constexpr (SYM0) {
        long_stage_state_wait<Fn, R, TT>(use_b, k6sm, ub, bb);
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) k6sm[LS::oU + 2 + r * TT + t] = ug[r * TT + t];
        ub = k6sm + LS::oU + 2;
    }
    if (t == 0 && tile > 0) k6sm[LS::oU + 1] = e[2 * (tile - 1) + 1];
    if (t == TT - 1 && tile + 1 < pd.ntiles) k6sm[LS::oU + 2 + R * TT] = e[2 * (tile + 1)];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) { uc[r][0] = ub[r * TT + t]; bc[r][0] = use_b ? bb[r * TT + t] : 0.0; }
    ChunkThomas<1, R> th;
    assemble_chunk<Fn, R, TT, SYM0, true, !SYM0, SYM0>(pd, mesh, ub, bb, in.mass, t, tt, inv_tau, in.prm, uc, bc, th, (int)toff,
                                                       e + 2 * pd.ntiles, e + 2 * pd.ntiles + 1);
    const double Sp = outer[0], St = outer[1];
    Vec<1> xp, xs, h[R];
    xs.a[0] = fma(-cC, St, fma(-cA, Sp, cD));
    xp.a[0] = (t > 0) ? fma(-pC, St, fma(-pA, Sp, pD)) : Sp;
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
    if (lane == 0) red[wi] = ss;
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int w = 1; w < W; ++w) ss += red[w];
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
    // m == 0: weights from the mesh points, as in the Newton sweeps (the same arithmetic in both entry points)
    assemble_chunk<Fn, R, TT, SYM0, JAC, true, SYM0>(pd, mesh, ub, bb, pd.mass + (size_t)k * 4, t, tt, inv_tau,
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


// =====================================================================================================================
// Systems (npde >= 2) on long meshes.  The same recursive partition with P x P blocks, in round 1's structure: every
// THREAD chunk of level 0 becomes a row of the level above (no reduction inside the CTA, no weights from x), stored
// levels reduce kLongRSys<P>:1 until one CTA can solve the top.  Generic in P; tuned for nothing but correctness and
// coalesced traffic -- the batched path covers systems up to 2048 / 1024 / 512 nodes (npde = 2 / 3 / 4), this covers
// the rest of the reference's range.  Records (doubles): summary = y0[P] v0[PP] w0[PP] ar[PP] bp[PP] Cs[PP] dp[P], padded
// to a multiple of 8; stored row = a[PP] c[PP] d[P], padded to an even number.
// =====================================================================================================================
template <int P> struct LongSys {
    static constexpr int PP = P * P;
    static constexpr int sum_doubles = (5 * PP + 2 * P + 7) / 8 * 8;
    static constexpr int row_doubles = (2 * PP + P + 1) / 2 * 2;
    static constexpr int oY0 = 0, oV0 = P, oW0 = P + PP, oAr = P + 2 * PP, oBp = P + 3 * PP, oCs = P + 4 * PP, oDp = P + 5 * PP;
};
__host__ __device__ constexpr int long_rows_level0(int P) { return P == 1 ? kLongR : (P == 2 ? 8 : (P == 3 ? 4 : 2)); }   // as the batched path
__host__ __device__ constexpr int long_rows_stored(int P) { return P == 1 ? kLongRS : (P == 2 ? 8 : (P == 3 ? 4 : 2)); }
__host__ __device__ constexpr int long_rows_top(int P) { return P == 1 ? 16 : (P == 2 ? 8 : (P == 3 ? 4 : 2)); }        // x kLongTT rows at most
// dynamic shared memory of k6s_back_asm: the tile's u with one node of halo on either side
__host__ __device__ constexpr size_t long_sys_smem_bytes(int P) { return (size_t)(long_rows_level0(P) * kLongT0 + 2) * P * sizeof(double); }

template <int P> SB_D void ld_block(const double* p, Mat<P>& m) {
#pragma unroll
    for (int i = 0; i < P * P; ++i) m.a[i] = __ldcg(p + i);
}
template <int P> SB_D void ld_vec(const double* p, Vec<P>& v) {
#pragma unroll
    for (int i = 0; i < P; ++i) v.a[i] = __ldcg(p + i);
}
template <int P> SB_D void st_block(double* p, const Mat<P>& m) {
#pragma unroll
    for (int i = 0; i < P * P; ++i) p[i] = m.a[i];
}
template <int P> SB_D void st_vec(double* p, const Vec<P>& v) {
#pragma unroll
    for (int i = 0; i < P; ++i) p[i] = v.a[i];
}

template <int P, int R>
SB_D void store_summary_sys(const ChunkThomas<P, R>& th, const LevelDev& lv, int64_t k, int c) {
    using LS = LongSys<P>;
    Vec<P> y0, yL; Mat<P> v0, w0, vL, wL;
    th.summary(y0, v0, w0, yL, vL, wL);
    double* S = reinterpret_cast<double*>(lv.sum) + ((size_t)k * lv.up_stride + tiled_pos(c, lv.up_R, kLongTT)) * LS::sum_doubles;
    st_vec<P>(S + LS::oY0, y0); st_block<P>(S + LS::oV0, v0); st_block<P>(S + LS::oW0, w0);
    st_block<P>(S + LS::oAr, neg(mul(th.As, vL)));
    st_block<P>(S + LS::oBp, sub_mul(th.Bs, th.As, wL));
    st_block<P>(S + LS::oCs, th.Cs);
    st_vec<P>(S + LS::oDp, sub_mul(th.ds, th.As, yL));
}

// rows of the level above `below` from the summaries of chunks j and j+1 (see form_rows)
template <int P, int R, int TT>
SB_D void form_rows_sys(const LevelDev& below, int64_t k, int tile, int t, Mat<P> (&ra)[R], Mat<P> (&rc)[R], Vec<P> (&rd)[R]) {
    using LS = LongSys<P>;
    const int n = below.nchunks;
    const double* S = reinterpret_cast<const double*>(below.sum) + (size_t)k * below.up_stride * LS::sum_doubles;
    const int j0 = (tile * TT + t) * R;
    const size_t p0 = (size_t)tile * R * TT + t;
    const size_t pl = (t + 1 < TT) ? (size_t)tile * R * TT + (t + 1) : (size_t)(tile + 1) * R * TT;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool in = j0 + r < n, in_next = j0 + r + 1 < n;
        const double* rec = S + (p0 + (size_t)r * TT) * LS::sum_doubles;
        const double* nxt = S + ((r + 1 < R) ? p0 + (size_t)(r + 1) * TT : (in_next ? pl : p0)) * LS::sum_doubles;
        Mat<P> ar, bp, Cs, v0n, w0n; Vec<P> dp, y0n;
        ld_block<P>(rec + LS::oAr, ar); ld_block<P>(rec + LS::oBp, bp); ld_block<P>(rec + LS::oCs, Cs); ld_vec<P>(rec + LS::oDp, dp);
        ld_vec<P>(nxt + LS::oY0, y0n); ld_block<P>(nxt + LS::oV0, v0n); ld_block<P>(nxt + LS::oW0, w0n);
        if (!in_next) { y0n = vec_zero<P>(); v0n = mat_zero<P>(); w0n = mat_zero<P>(); }
        if (!in) { ar = mat_zero<P>(); bp = mat_identity<P>(); Cs = mat_zero<P>(); dp = vec_zero<P>(); }   // identity row
        const Mat<P> binv = inverse_fast(sub_mul(bp, Cs, v0n));
        ra[r] = mul(binv, ar);
        rc[r] = neg(mul(binv, mul(Cs, w0n)));
        rd[r] = mul(binv, sub_mul(dp, Cs, y0n));
    }
}

template <int P>
SB_D void separators_sys(const LevelDev& up, int64_t k, int c, Vec<P>& xp, Vec<P>& xs) {
    const double* xu = up.x + (size_t)k * up.stride * P;
    ld_vec<P>(xu + tiled_pos(c, up.R, kLongTT) * P, xs);
    if (c > 0) ld_vec<P>(xu + tiled_pos(c - 1, up.R, kLongTT) * P, xp);
    else xp = vec_zero<P>();
}

// level 0, forward
template <class Fn, int R, int TT, bool SYM0>
__global__ void __launch_bounds__(TT) k6s_chunk_asm(ProbDev pd, double tt, double inv_tau, LevelDev lv) {
    constexpr int P = Fn::npde;
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    const size_t toff = (size_t)tile * R * TT;
    pdl_release();
    pdl_wait();
    if (!pd.active[k]) return;
    const double* ui = pd.u + (size_t)k * pd.L * P;
    const double* ub = ui + toff * P;
    const double* bb = pd.b + ((size_t)k * pd.L + toff) * P;
    double uc[R][P], bc[R][P];
    load_chunk<P, R, TT>(ub, t, uc);
    load_chunk<P, R, TT>(bb, t, bc);
    {   // the iterate at the tile edges and the mesh ends, for the in-place backward sweep
        double* e = lv.edge + (size_t)k * (2 * pd.ntiles + 2) * P;
        const int i0 = (int)toff + t * R;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            if (t == 0) e[(2 * tile) * P + p] = uc[0][p];
            if (t == TT - 1) e[(2 * tile + 1) * P + p] = uc[R - 1][p];
            if (tile == 0 && t == 0) e[(2 * pd.ntiles) * P + p] = uc[0][p];
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (i0 + r == pd.Nx - 1) e[(2 * pd.ntiles + 1) * P + p] = uc[r][p];
        }
    }
    ChunkThomas<P, R> th;
    assemble_chunk<Fn, R, TT, SYM0, true, true>(pd, mesh_of_tile(pd.mesh, toff), ub, bb, pd.mass + (size_t)k * 4 * P, t, tt, inv_tau,
                                                pd.prm + (size_t)k * pd.nprm_s, uc, bc, th, (int)toff, ui, ui + tiled_pos(pd.Nx - 1, R, TT) * P);
    store_summary_sys<P, R>(th, lv, k, tile * TT + t);
}

// level l >= 1, forward: form the rows, keep them, eliminate, store the chunk summary
template <int P, int R, int TT>
__global__ void __launch_bounds__(TT) k6s_chunk_sys(LevelDev lv, LevelDev below, const int* active) {
    using LS = LongSys<P>;
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    pdl_release();
    pdl_wait();
    if (!active[k]) return;
    Mat<P> ra[R], rc[R]; Vec<P> rd[R];
    form_rows_sys<P, R, TT>(below, k, tile, t, ra, rc, rd);
    ChunkThomas<P, R> th;
    double* rows = reinterpret_cast<double*>(lv.acd) + ((size_t)k * lv.stride + (size_t)tile * R * TT + t) * LS::row_doubles;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        double* row = rows + (size_t)r * TT * LS::row_doubles;
        st_block<P>(row, ra[r]); st_block<P>(row + LS::PP, rc[r]); st_vec<P>(row + 2 * LS::PP, rd[r]);
        th.row(r, ra[r], mat_identity<P>(), rc[r], rd[r]);
    }
    store_summary_sys<P, R>(th, lv, k, tile * TT + t);
}

// top level: one CTA per instance
template <int P, int RT, int TT>
__global__ void __launch_bounds__(TT) k6s_top(LevelDev lv, LevelDev below, const int* active) {
    extern __shared__ double sm[];
    const int t = threadIdx.x;
    const int64_t k = blockIdx.x;
    pdl_release();
    pdl_wait();
    if (!active[k]) return;
    Mat<P> ra[RT], rc[RT]; Vec<P> rd[RT];
    form_rows_sys<P, RT, TT>(below, k, 0, t, ra, rc, rd);
    ChunkThomas<P, RT> th;
#pragma unroll
    for (int r = 0; r < RT; ++r) th.row(r, ra[r], mat_identity<P>(), rc[r], rd[r]);
    Vec<P> x[RT];
    th.template finish<TT>(x, sm, t);
    double* xo = lv.x + (size_t)k * lv.stride * P;
#pragma unroll
    for (int r = 0; r < RT; ++r) st_vec<P>(xo + ((size_t)r * TT + t) * P, x[r]);
}

// level l >= 1, backward
template <int P, int R, int TT>
__global__ void __launch_bounds__(TT) k6s_back_sys(LevelDev lv, LevelDev up, const int* active) {
    using LS = LongSys<P>;
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    pdl_release();
    pdl_wait();
    if (!active[k]) return;
    ChunkThomas<P, R> th;
    const size_t base = (size_t)k * lv.stride + (size_t)tile * R * TT;
    const double* rows = reinterpret_cast<const double*>(lv.acd) + (base + t) * LS::row_doubles;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double* row = rows + (size_t)r * TT * LS::row_doubles;
        Mat<P> A, C; Vec<P> d;
        ld_block<P>(row, A); ld_block<P>(row + LS::PP, C); ld_vec<P>(row + 2 * LS::PP, d);
        th.row(r, A, mat_identity<P>(), C, d);
    }
    Vec<P> xp, xs, x[R];
    separators_sys<P>(up, k, tile * TT + t, xp, xs);
    th.back(x, xp, xs);
#pragma unroll
    for (int r = 0; r < R; ++r) st_vec<P>(lv.x + (base + (size_t)r * TT + t) * P, x[r]);
}

// level 0, backward: recompute, back-substitute, update u in place, per-tile sum of h^2
template <class Fn, int R, int TT, bool SYM0>
__global__ void __launch_bounds__(TT) k6s_back_asm(ProbDev pd, double tt, double inv_tau, LevelDev up, double* partial) {
    constexpr int P = Fn::npde;
    __shared__ double red[TT / 32];
    extern __shared__ __align__(16) double k6sm[];      // [(R*TT + 2) * P]: the tile's u with one node of halo on either side
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    const size_t toff = (size_t)tile * R * TT;
    pdl_release();
    pdl_wait();
    if (!pd.active[k]) return;
    double* ug = pd.u + ((size_t)k * pd.L + toff) * P;
    const double* bb = pd.b + ((size_t)k * pd.L + toff) * P;
    const double* e = up.edge + (size_t)k * (2 * pd.ntiles + 2) * P;
    double* um = k6sm + P;
    double uc[R][P], bc[R][P];
    load_chunk<P, R, TT>(ug, t, uc);
    load_chunk<P, R, TT>(bb, t, bc);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int p = 0; p < P; ++p) um[(r * TT + t) * P + p] = uc[r][p];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        if (t == 0 && tile > 0) um[-P + p] = e[(2 * (tile - 1) + 1) * P + p];
        if (t == TT - 1 && tile + 1 < pd.ntiles) um[R * TT * P + p] = e[(2 * (tile + 1)) * P + p];
    }
    __syncthreads();
    ChunkThomas<P, R> th;
    assemble_chunk<Fn, R, TT, SYM0, true, true>(pd, mesh_of_tile(pd.mesh, toff), um, bb, pd.mass + (size_t)k * 4 * P, t, tt, inv_tau,
                                                pd.prm + (size_t)k * pd.nprm_s, uc, bc, th, (int)toff, e + (2 * pd.ntiles) * P,
                                                e + (2 * pd.ntiles + 1) * P);
    Vec<P> xp, xs, h[R];
    separators_sys<P>(up, k, tile * TT + t, xp, xs);
    th.back(h, xp, xs);
    double ss = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool in = (int)toff + t * R + r < pd.Nx;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            ug[(size_t)(r * TT + t) * P + p] = in ? uc[r][p] - h[r].a[p] : 0.0;
            if (in) ss += h[r].a[p] * h[r].a[p];
        }
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

// residual-only / residual+Jacobian of a long mesh of a system into F (and Jl, Jd, Ju)
template <class Fn, int R, int TT, bool SYM0, bool JAC>
__global__ void __launch_bounds__(TT) k6s_assemble(ProbDev pd, double tt, double inv_tau, int all_instances) {
    constexpr int P = Fn::npde;
    const int tile = blockIdx.x, t = threadIdx.x;
    const int64_t k = blockIdx.y;
    if (!all_instances && !pd.active[k]) return;
    const size_t toff = (size_t)tile * R * TT;
    const double* ui = pd.u + (size_t)k * pd.L * P;
    const double* ub = ui + toff * P;
    const double* bb = pd.b + ((size_t)k * pd.L + toff) * P;
    double uc[R][P], bc[R][P];
    load_chunk<P, R, TT>(ub, t, uc);
    if constexpr (JAC) load_chunk<P, R, TT>(bb, t, bc);
    else {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int p = 0; p < P; ++p) bc[r][p] = 0.0;
    }
    ProbDev pt = pd;  // StoreSink addresses (k*L + r*TT + t): shift the output arrays to the tile
    pt.F = pd.F + toff * P;
    if (JAC) { pt.Jl = pd.Jl + toff * P * P; pt.Jd = pd.Jd + toff * P * P; pt.Ju = pd.Ju + toff * P * P; }
    StoreSink<P, TT> sink{pt, k, t};
    assemble_chunk<Fn, R, TT, SYM0, JAC, true>(pd, mesh_of_tile(pd.mesh, toff), ub, bb, pd.mass + (size_t)k * 4 * P, t, tt, inv_tau,
                                               pd.prm + (size_t)k * pd.nprm_s, uc, bc, sink, (int)toff, ui, ui + tiled_pos(pd.Nx - 1, R, TT) * P);
}

}  // namespace sb
