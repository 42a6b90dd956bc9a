// sb_functors.cuh -- registry of device-compiled coefficient functors.
//
// Each functor restates one of the reference's example callbacks
//     pdefun(x,t,u,dudx) -> (c,f,s)          bdfun(xl,ul,xr,ur,t) -> (pl,ql,pr,qr)
// (docs/src/solvers.md:9-47) generically in the scalar type S, so the same code runs with plain
// doubles (assemble!) and with dual numbers (Newton Jacobian).  prm[] is the per-instance
// parameter vector of the sweep.  ql/qr are value-only (plain doubles): in every registry entry
// they are constants, exactly as in the reference examples.
// Ids and parameter order: include/skeelberzins_b200.h (enum sb_functor_id).
// Traits: c_unit = the capacity c is identically 1, s_zero = the source s is identically 0, s_const (optional) = s
// does not depend on u or dudx; they let the assembler drop dual arithmetic on structural zeros (the values are
// still what pde() returns).
#pragma once
#include "sb_dual.cuh"

namespace sb {

// f = D * ux family with Neumann ends -- examples/Example101_LinearDiffusion.jl:21-38
struct FnLinDiff {
    static constexpr int npde = 1, nparams = 1;
    static constexpr bool c_unit = true, s_zero = true;
    template <class S> static SB_D void pde(double, double, const S*, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(1.0); f[0] = p[0] * ux[0]; s[0] = S(0.0);
    }
    template <class S> static SB_D void bc(double, const S*, double, const S*, double, const double*, S* pl, double* ql, S* pr, double* qr) {
        pl[0] = S(0.0); ql[0] = 1.0; pr[0] = S(0.0); qr[0] = 1.0;
    }
};

// examples/Example102_NonlinearDiffusion.jl:44-65 : f = a*u*ux
struct FnNonlinDiff {
    static constexpr int npde = 1, nparams = 1;
    static constexpr bool c_unit = true, s_zero = true;
    template <class S> static SB_D void pde(double, double, const S* u, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(1.0); f[0] = p[0] * u[0] * ux[0]; s[0] = S(0.0);
    }
    template <class S> static SB_D void bc(double, const S*, double, const S*, double, const double*, S* pl, double* ql, S* pr, double* qr) {
        pl[0] = S(0.0); ql[0] = 1.0; pr[0] = S(0.0); qr[0] = 1.0;
    }
};

// examples/Example103_LinearDiffusionCylindrical.jl:35-58 : right Dirichlet g*exp(-D n^2 t)
struct FnLinDiffCyl {
    static constexpr int npde = 1, nparams = 3;
    static constexpr bool c_unit = true, s_zero = true;
    template <class S> static SB_D void pde(double, double, const S*, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(1.0); f[0] = p[0] * ux[0]; s[0] = S(0.0);
    }
    template <class S> static SB_D void bc(double, const S*, double, const S* ur, double t, const double* p, S* pl, double* ql, S* pr, double* qr) {
        pl[0] = S(0.0); ql[0] = 0.0;
        pr[0] = ur[0] - p[2] * ::exp(-p[0] * (p[1] * p[1]) * t); qr[0] = 0.0;
    }
};

// examples/Example104_Poisson.jl:20-41 ; test/test_solveStationary.jl:6-23 (c0 = 0)
struct FnPoisson {
    static constexpr int npde = 1, nparams = 3;
    static constexpr bool c_unit = false, s_zero = false, s_const = true;
    template <class S> static SB_D void pde(double, double, const S*, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(p[0]); f[0] = ux[0]; s[0] = S(p[1]);
    }
    template <class S> static SB_D void bc(double, const S* ul, double, const S* ur, double, const double* p, S* pl, double* ql, S* pr, double* qr) {
        pl[0] = ul[0] - p[2]; ql[0] = 0.0; pr[0] = ur[0] - p[2]; qr[0] = 0.0;
    }
};

// examples/Example105_StationaryNonlinearDiffusion.jl:30-51
struct FnStatNonlin {
    static constexpr int npde = 1, nparams = 3;
    static constexpr bool c_unit = true, s_zero = false, s_const = true;
    template <class S> static SB_D void pde(double, double, const S* u, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(1.0); f[0] = p[0] * u[0] * ux[0]; s[0] = S(p[1]);
    }
    template <class S> static SB_D void bc(double, const S* ul, double, const S* ur, double, const double* p, S* pl, double* ql, S* pr, double* qr) {
        pl[0] = ul[0] - p[2]; ql[0] = 0.0; pr[0] = ur[0] - p[2]; qr[0] = 0.0;
    }
};

// examples/Example106_SystemReactionDiffusion.jl:42-64 (npde = 2)
struct FnReactDiff {
    static constexpr int npde = 2, nparams = 3;
    static constexpr bool c_unit = true, s_zero = false;
    template <class S> static SB_D void pde(double, double, const S* u, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(1.0); c[1] = S(1.0);
        f[0] = p[0] * ux[0]; f[1] = p[1] * ux[1];
        const S y = u[0] - u[1];
        s[0] = -(p[2] * y); s[1] = p[2] * y;
    }
    template <class S> static SB_D void bc(double, const S* ul, double, const S* ur, double, const double*, S* pl, double* ql, S* pr, double* qr) {
        pl[0] = ul[0] - 1.0; pl[1] = S(0.0); ql[0] = 0.0; ql[1] = 1.0;
        pr[0] = S(0.0); pr[1] = ur[1]; qr[0] = 1.0; qr[1] = 0.0;
    }
};

// examples/Example107_LinearDiffusionSpherical.jl:34-55 : right Dirichlet 1 + 6 D t
struct FnLinDiffSph {
    static constexpr int npde = 1, nparams = 1;
    static constexpr bool c_unit = true, s_zero = true;
    template <class S> static SB_D void pde(double, double, const S*, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(1.0); f[0] = p[0] * ux[0]; s[0] = S(0.0);
    }
    template <class S> static SB_D void bc(double, const S*, double, const S* ur, double t, const double* p, S* pl, double* ql, S* pr, double* qr) {
        pl[0] = S(0.0); ql[0] = 0.0;
        pr[0] = ur[0] - (1.0 + 6.0 * p[0] * t); qr[0] = 0.0;
    }
};

// examples/Example201_PartialDerivativeApprox.jl:34-51
struct FnConvDiff {
    static constexpr int npde = 1, nparams = 3;
    static constexpr bool c_unit = true, s_zero = false;
    template <class S> static SB_D void pde(double, double, const S*, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(1.0); f[0] = p[0] * ux[0]; s[0] = -(p[0] * p[1] / p[2]) * ux[0];
    }
    template <class S> static SB_D void bc(double, const S* ul, double, const S* ur, double, const double*, S* pl, double* ql, S* pr, double* qr) {
        pl[0] = ul[0]; ql[0] = 0.0; pr[0] = ur[0]; qr[0] = 0.0;
    }
};

// examples/Example301_PDEOptimSciMLSensitivity.jl:31-48
struct FnLinReact {
    static constexpr int npde = 1, nparams = 2;
    static constexpr bool c_unit = true, s_zero = false;
    template <class S> static SB_D void pde(double, double, const S* u, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(1.0); f[0] = p[1] * ux[0]; s[0] = 2.0 * p[0] * u[0];
    }
    template <class S> static SB_D void bc(double, const S* ul, double, const S* ur, double, const double*, S* pl, double* ql, S* pr, double* qr) {
        pl[0] = ul[0]; ql[0] = 0.0; pr[0] = ur[0]; qr[0] = 0.0;
    }
};

// quasi-linear scalar family (parameter order in include/skeelberzins_b200.h)
struct FnGenericScalar {
    static constexpr int npde = 1, nparams = 19;
    static constexpr bool c_unit = false, s_zero = false;
    template <class S> static SB_D void pde(double, double, const S* u, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = p[0] + p[1] * u[0];
        f[0] = (p[2] + p[3] * u[0]) * ux[0];
        s[0] = p[4] + p[5] * u[0] + p[6] * ux[0];
    }
    template <class S> static SB_D void bc(double, const S* ul, double, const S* ur, double t, const double* p, S* pl, double* ql, S* pr, double* qr) {
        pl[0] = p[7] * ul[0] - (p[8] + p[9] * t + p[10] * ::exp(-p[11] * t)); ql[0] = p[12];
        pr[0] = p[13] * ur[0] - (p[14] + p[15] * t + p[16] * ::exp(-p[17] * t)); qr[0] = p[18];
    }
};

}  // namespace sb
