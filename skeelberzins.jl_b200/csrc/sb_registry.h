// sb_registry.h -- host-side table of kernel entry points per registry functor.
// One translation unit per functor (sb_inst.cu compiled with -DSB_FN=...) fills one entry, so the
// instantiations build in parallel.
#pragma once
#include <utility>

#include "sb_longmesh.cuh"

namespace sb {

// Thread decompositions the kernels are compiled for: R rows per thread x TT threads per system.
// TT == 32: one warp per system (shuffles only); TT > 32: one CTA per system.
struct Decomp { int R, TT; };
constexpr int kNumDecomp = 16;
constexpr Decomp kDecomp[kNumDecomp] = {
    {1, 32}, {2, 32}, {4, 32}, {8, 32}, {16, 32},   // warp per system: Nx <= 32 .. 512
    {4, 64}, {8, 64}, {16, 64},                     // CTA per system
    {4, 128}, {8, 128},
    {4, 256}, {8, 256}, {16, 256},
    {2, 64}, {2, 128}, {2, 256},                    // npde = 4 only (two rows per thread is all its registers hold)
};
// rows per thread the register file allows: the factors of a row are 2*P*P + P doubles (R = 16 only for scalar PDEs)
constexpr int max_rows_per_thread(int P) { return P == 1 ? 16 : P == 2 ? 8 : P == 3 ? 4 : 2; }
constexpr bool decomp_ok(int P, int d) {
    return kDecomp[d].R <= max_rows_per_thread(P) && (P == 4 || !(kDecomp[d].R == 2 && kDecomp[d].TT > 32));
}

typedef void (*AssembleKernel)(ProbDev, double, double, int);
typedef void (*FusedKernel)(ProbDev, double, double, double, int);
typedef void (*FusedPKernel)(ProbDev, double, double, double, int, int, int);
typedef void (*SolveKernel)(ProbDev, double, int, int);
typedef void (*ResidentKernel)(ProbDev, ResidentArgs);
typedef void (*ResidentPipeKernel)(ProbDev, ResidentArgs, ResidentPipe);
typedef void (*MassKernel)(ProbDev, double, double*, int, const double*);
typedef void (*LongChunkKernel)(ProbDev, double, double, LevelDev);
typedef void (*LongBackKernel)(ProbDev, double, double, LevelDev, double*);

struct FunctorEntry {
    const char* name;
    int npde, nparams;
    AssembleKernel assemble_jac[kNumDecomp][2];  // [decomposition][SYM0]
    AssembleKernel assemble_val[kNumDecomp][2];
    FusedKernel fused[kNumDecomp][2];
    FusedPKernel fused_p[kNumDecomp][2];      // persistent TMA-prefetch variant
    size_t fused_p_smem[kNumDecomp][2];       // its dynamic shared memory (bytes)
    ResidentKernel resident[kNumDecomp][2];   // whole solve in one launch (sb_resident.cuh)
    ResidentPipeKernel resident_pipe[kNumDecomp][2];   // ... whose instances wait for the pipelined upload (sb_solve_from_host)
    size_t resident_smem[kNumDecomp][2];
    MassKernel mass;
    // long-mesh path (scalar PDEs): [SYM0]
    LongChunkKernel long_chunk[2];
    LongBackKernel long_back[2];
    size_t long_smem[2];                      // dynamic shared memory of long_chunk / long_back (bytes)
    AssembleKernel long_assemble_jac[2], long_assemble_val[2];
};

template <class Fn, int D>
void fill_one(FunctorEntry& e) {
    if constexpr (decomp_ok(Fn::npde, D)) {
        constexpr int R = kDecomp[D].R, TT = kDecomp[D].TT;
        e.assemble_jac[D][0] = k_assemble<Fn, R, TT, false, true>;
        e.assemble_jac[D][1] = k_assemble<Fn, R, TT, true, true>;
        e.assemble_val[D][0] = k_assemble<Fn, R, TT, false, false>;
        e.assemble_val[D][1] = k_assemble<Fn, R, TT, true, false>;
        e.fused[D][0] = k_newton_fused<Fn, R, TT, false>;
        e.fused[D][1] = k_newton_fused<Fn, R, TT, true>;
        e.fused_p[D][0] = k_newton_fused_p<Fn, R, TT, false>;
        e.fused_p[D][1] = k_newton_fused_p<Fn, R, TT, true>;
        e.fused_p_smem[D][0] = FusedPLayout<Fn, R, TT, false>::bytes;
        e.fused_p_smem[D][1] = FusedPLayout<Fn, R, TT, true>::bytes;
        e.resident[D][0] = k_solve_resident<Fn, R, TT, false>;
        e.resident[D][1] = k_solve_resident<Fn, R, TT, true>;
        e.resident_pipe[D][0] = k_solve_resident_pipe<Fn, R, TT, false>;
        e.resident_pipe[D][1] = k_solve_resident_pipe<Fn, R, TT, true>;
        e.resident_smem[D][0] = ResidentLayout<Fn, R, TT, false>::bytes;
        e.resident_smem[D][1] = ResidentLayout<Fn, R, TT, true>::bytes;
    } else {
        for (int s = 0; s < 2; ++s) {
            e.assemble_jac[D][s] = nullptr; e.assemble_val[D][s] = nullptr; e.fused[D][s] = nullptr;
            e.fused_p[D][s] = nullptr; e.fused_p_smem[D][s] = 0; e.resident[D][s] = nullptr; e.resident_pipe[D][s] = nullptr; e.resident_smem[D][s] = 0;
        }
    }
}

template <class Fn, int... D>
void fill_all(FunctorEntry& e, std::integer_sequence<int, D...>) { (fill_one<Fn, D>(e), ...); }

template <class Fn>
FunctorEntry make_entry(const char* name) {
    FunctorEntry e;
    e.name = name; e.npde = Fn::npde; e.nparams = Fn::nparams;
    fill_all<Fn>(e, std::make_integer_sequence<int, kNumDecomp>{});
    e.mass = k_mass_flags<Fn>;
    if constexpr (Fn::npde == 1) {
        e.long_chunk[0] = k6_chunk_asm<Fn, kLongR, kLongT0, false>; e.long_chunk[1] = k6_chunk_asm<Fn, kLongR, kLongT0, true>;
        e.long_back[0] = k6_back_asm<Fn, kLongR, kLongT0, false>; e.long_back[1] = k6_back_asm<Fn, kLongR, kLongT0, true>;
        e.long_smem[0] = LongStage<Fn, kLongR, kLongT0, false>::bytes; e.long_smem[1] = LongStage<Fn, kLongR, kLongT0, true>::bytes;
        e.long_assemble_jac[0] = k6_assemble<Fn, kLongR, kLongT0, false, true>;
        e.long_assemble_jac[1] = k6_assemble<Fn, kLongR, kLongT0, true, true>;
        e.long_assemble_val[0] = k6_assemble<Fn, kLongR, kLongT0, false, false>;
        e.long_assemble_val[1] = k6_assemble<Fn, kLongR, kLongT0, true, false>;
    } else {   // systems: the generic-P kernels (sb_longmesh.cuh, "Systems on long meshes")
        constexpr int R0 = long_rows_level0(Fn::npde);
        e.long_chunk[0] = k6s_chunk_asm<Fn, R0, kLongT0, false>; e.long_chunk[1] = k6s_chunk_asm<Fn, R0, kLongT0, true>;
        e.long_back[0] = k6s_back_asm<Fn, R0, kLongT0, false>; e.long_back[1] = k6s_back_asm<Fn, R0, kLongT0, true>;
        e.long_smem[0] = e.long_smem[1] = long_sys_smem_bytes(Fn::npde);
        e.long_assemble_jac[0] = k6s_assemble<Fn, R0, kLongT0, false, true>;
        e.long_assemble_jac[1] = k6s_assemble<Fn, R0, kLongT0, true, true>;
        e.long_assemble_val[0] = k6s_assemble<Fn, R0, kLongT0, false, false>;
        e.long_assemble_val[1] = k6s_assemble<Fn, R0, kLongT0, true, false>;
    }
    return e;
}

template <int P, int D>
void fill_solve_one(SolveKernel (&tab)[kNumDecomp]) {
    if constexpr (decomp_ok(P, D)) tab[D] = k_solve<P, kDecomp[D].R, kDecomp[D].TT>;
    else tab[D] = nullptr;
}
template <int P, int... D>
void fill_solve(SolveKernel (&tab)[kNumDecomp], std::integer_sequence<int, D...>) { (fill_solve_one<P, D>(tab), ...); }

// shared memory (bytes) of the solve / fused kernels for a decomposition
inline size_t solve_smem_bytes(int P, int d) {
    const int TT = kDecomp[d].TT;
    if (TT == 32) return 0;
    const int NW = 2 * P * P + P, W = TT / 32;
#ifdef SB_FINISH_PARALLEL
    const int exch = (2 * NW + 1) * W;
#else
    const int exch = (NW + P) * TT + (NW + 1) * W + 1;
#endif
    return (size_t)(exch + W /* partial sums */ + 1 /* the fallback's slot word */) * sizeof(double);   // ChunkThomas::smem_doubles<TT>() + 1
}

// defined in the per-functor translation units
FunctorEntry entry_lin_diff();
FunctorEntry entry_nonlin_diff();
FunctorEntry entry_lin_diff_cyl();
FunctorEntry entry_poisson();
FunctorEntry entry_stat_nonlin();
FunctorEntry entry_react_diff();
FunctorEntry entry_lin_diff_sph();
FunctorEntry entry_conv_diff();
FunctorEntry entry_lin_react();
FunctorEntry entry_generic_scalar();

}  // namespace sb
