// sb_registry.h -- host-side table of kernel entry points per registry functor.
// One translation unit per functor (sb_inst.cu compiled with -DSB_FN=...) fills one entry, so the
// instantiations build in parallel.
#pragma once
#include "sb_kernels.cuh"

namespace sb {

constexpr int kNumR = 5;  // rows per thread R = 1, 2, 4, 8, 16  (index = log2 R)

typedef void (*AssembleKernel)(ProbDev, double, double, int);
typedef void (*FusedKernel)(ProbDev, double, double, double, int);
typedef void (*SolveKernel)(ProbDev, double, int, int);
typedef void (*MassKernel)(ProbDev, double, double*, int);

struct FunctorEntry {
    const char* name;
    int npde, nparams;
    AssembleKernel assemble_jac[kNumR][2];  // [log2 R][WARP]
    AssembleKernel assemble_val[kNumR][2];
    FusedKernel fused[kNumR][2];
    MassKernel mass;
};

template <class Fn, int R>
void fill_r(FunctorEntry& e, int ri) {
    e.assemble_jac[ri][0] = k_assemble<Fn, R, false, true>;
    e.assemble_jac[ri][1] = k_assemble<Fn, R, true, true>;
    e.assemble_val[ri][0] = k_assemble<Fn, R, false, false>;
    e.assemble_val[ri][1] = k_assemble<Fn, R, true, false>;
    e.fused[ri][0] = k_newton_fused<Fn, R, false>;
    e.fused[ri][1] = k_newton_fused<Fn, R, true>;
}

template <class Fn>
FunctorEntry make_entry(const char* name) {
    FunctorEntry e;
    e.name = name; e.npde = Fn::npde; e.nparams = Fn::nparams;
    fill_r<Fn, 1>(e, 0);
    fill_r<Fn, 2>(e, 1);
    fill_r<Fn, 4>(e, 2);
    fill_r<Fn, 8>(e, 3);
    if constexpr (Fn::npde == 1) fill_r<Fn, 16>(e, 4);
    else { for (int w = 0; w < 2; ++w) { e.assemble_jac[4][w] = nullptr; e.assemble_val[4][w] = nullptr; e.fused[4][w] = nullptr; } }
    e.mass = k_mass_flags<Fn>;
    return e;
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
