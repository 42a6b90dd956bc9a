// sb_oracle.cpp -- CPU ORACLE for the implicit-Euler / stationary-Newton path of SkeelBerzins.jl.
//
// TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement of the reference algorithm used as
// the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs.  Nothing in the product package (skeelberzins.jl_b200/) may import, link or call it.
//
// Why a restatement: the reference is pure Julia and `julia` is not installed in this image, and
// its Jacobian/linear-solve arithmetic lives in un-vendored third-party packages
// (SparseDiffTools "1.19, 2" + ForwardDiff, LinearSolve "2" -> KLU; Project.toml:18-27, no
// Manifest).  Their published contract is restated here: exact forward-mode Jacobian of the
// residual on the block-tridiagonal pattern, h = J\F by LU with partial pivoting, ||h||_2.
// Parity is PINNED against the reference's own golden values (tests/test_oracle_goldens.py checks
// all of examples/Example101..107,201 and test/test_solve*.jl literals).
//
// Each function cites the reference lines it follows (paths relative to the reference tree).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

// --------------------------------------------------------------------------------------------
// ForwardDiff.Dual restated: value + N partials, generic arithmetic
// --------------------------------------------------------------------------------------------
template <int N>
struct Dual {
    double v;
    double d[N];
    Dual() : v(0.0) { for (int k = 0; k < N; ++k) d[k] = 0.0; }
    Dual(double x) : v(x) { for (int k = 0; k < N; ++k) d[k] = 0.0; }
};
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int k = 0; k < N; ++k) r.d[k] = a.d[k] + b.d[k]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int k = 0; k < N; ++k) r.d[k] = a.d[k] - b.d[k]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int k = 0; k < N; ++k) r.d[k] = -a.d[k]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k]; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v / b.v; for (int k = 0; k < N; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) / b.v; return r; }
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double a, const Dual<N>& b) { return b + a; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double a, const Dual<N>& b) { Dual<N> r = -b; r.v += a; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v * b; for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * b; return r; }
template <int N> inline Dual<N> operator*(double a, const Dual<N>& b) { return b * a; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v / b; for (int k = 0; k < N; ++k) r.d[k] = a.d[k] / b; return r; }
template <int N> inline Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }
template <int N> inline Dual<N> exp(const Dual<N>& a) { Dual<N> r; r.v = std::exp(a.v); for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * r.v; return r; }
inline double value(double x) { return x; }
template <int N> inline double value(const Dual<N>& x) { return x.v; }

// --------------------------------------------------------------------------------------------
// Functor registry: the example callbacks restated generically in the scalar type T
// (ids documented in include/skeelberzins_b200.h; sources cited there and below)
// --------------------------------------------------------------------------------------------
enum {
    LIN_DIFF = 0, NONLIN_DIFF = 1, LIN_DIFF_CYL = 2, POISSON = 3, STAT_NONLIN = 4, REACT_DIFF = 5,
    LIN_DIFF_SPH = 6, CONV_DIFF = 7, LIN_REACT = 8, GENERIC_SCALAR = 9,
    EXP_DIFF = 10,  // not in the product registry: checker for run-time compiled user functors (tests/test_gpu_plugin.py)
    CHAIN3 = 11,    // likewise: three species, A -> B -> C with a quadratic first step (npde = 3)
    CHAIN4 = 12,    // four species, A -> B -> C -> D, second species with a u-dependent capacity (npde = 4)
    NUM_FUNCTORS = 13
};
static const int k_npde[NUM_FUNCTORS] = {1, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1, 3, 4};
static const int k_nprm[NUM_FUNCTORS] = {1, 1, 3, 3, 3, 3, 1, 3, 2, 19, 3, 5, 7};

// pdefun(x,t,u,dudx) -> (c,f,s)     docs/src/solvers.md:9-24
template <class T>
static void pdefun(int id, const double* p, double x, double t, const T* u, const T* ux, T* c, T* f, T* s) {
    (void)t;
    switch (id) {
    case LIN_DIFF:      // examples/Example101_LinearDiffusion.jl:21-27
    case LIN_DIFF_CYL:  // examples/Example103_LinearDiffusionCylindrical.jl:35-41
    case LIN_DIFF_SPH:  // examples/Example107_LinearDiffusionSpherical.jl:34-40
        c[0] = T(1.0); f[0] = p[0] * ux[0]; s[0] = T(0.0); break;
    case NONLIN_DIFF:   // examples/Example102_NonlinearDiffusion.jl:44-50  (a = 2)
        c[0] = T(1.0); f[0] = p[0] * u[0] * ux[0]; s[0] = T(0.0); break;
    case POISSON:       // examples/Example104_Poisson.jl:20-26 ; test/test_solveStationary.jl:6-12 (c=0)
        c[0] = T(p[0]); f[0] = ux[0]; s[0] = T(p[1]); break;
    case STAT_NONLIN:   // examples/Example105_StationaryNonlinearDiffusion.jl:30-36
        c[0] = T(1.0); f[0] = p[0] * u[0] * ux[0]; s[0] = T(p[1]); break;
    case REACT_DIFF: {  // examples/Example106_SystemReactionDiffusion.jl:42-50
        c[0] = T(1.0); c[1] = T(1.0);
        f[0] = p[0] * ux[0]; f[1] = p[1] * ux[1];
        T y = u[0] - u[1];
        s[0] = -(p[2] * y); s[1] = p[2] * y; break; }
    case CONV_DIFF:     // examples/Example201_PartialDerivativeApprox.jl:34-40
        c[0] = T(1.0); f[0] = p[0] * ux[0]; s[0] = -(p[0] * p[1] / p[2]) * ux[0]; break;
    case LIN_REACT:     // examples/Example301_PDEOptimSciMLSensitivity.jl:31-37
        c[0] = T(1.0); f[0] = p[1] * ux[0]; s[0] = 2.0 * p[0] * u[0]; break;
    case GENERIC_SCALAR:
        c[0] = p[0] + p[1] * u[0];
        f[0] = (p[2] + p[3] * u[0]) * ux[0];
        s[0] = p[4] + p[5] * u[0] + p[6] * ux[0]; break;
    case EXP_DIFF:      // c = 1, f = p0*exp(u)*ux, s = -p1*u^2 + x
        c[0] = T(1.0); f[0] = p[0] * exp(u[0]) * ux[0]; s[0] = -(p[1] * u[0] * u[0]) + x; break;
    case CHAIN3: {      // prm = {D1, D2, D3, k1, k2}: A -> B (rate k1*A^2), B -> C (rate k2*B)
        c[0] = T(1.0); c[1] = T(1.0); c[2] = T(1.0);
        f[0] = p[0] * ux[0]; f[1] = p[1] * ux[1]; f[2] = p[2] * ux[2];
        T r1 = p[3] * u[0] * u[0], r2 = p[4] * u[1];
        s[0] = -r1; s[1] = r1 - r2; s[2] = r2; break; }
    case CHAIN4: {      // prm = {D1..D4, k1, k2, k3}: linear chain; species 2 has capacity 1 + u2 and its flux feels u1
        c[0] = T(1.0); c[1] = 1.0 + u[1]; c[2] = T(1.0); c[3] = T(1.0);
        f[0] = p[0] * ux[0]; f[1] = p[1] * (1.0 + u[0]) * ux[1]; f[2] = p[2] * ux[2]; f[3] = p[3] * ux[3];
        T r1 = p[4] * u[0], r2 = p[5] * u[1], r3 = p[6] * u[2];
        s[0] = -r1; s[1] = r1 - r2; s[2] = r2 - r3; s[3] = r3; break; }
    }
}

// bdfun(xl,ul,xr,ur,t) -> (pl,ql,pr,qr)     docs/src/solvers.md:26-47
template <class T>
static void bdfun(int id, const double* p, double xl, const T* ul, double xr, const T* ur, double t,
                  T* pl, T* ql, T* pr, T* qr) {
    (void)xl; (void)xr;
    switch (id) {
    case LIN_DIFF:      // examples/Example101_LinearDiffusion.jl:31-38
    case NONLIN_DIFF:   // examples/Example102_NonlinearDiffusion.jl:57-65
        pl[0] = T(0.0); ql[0] = T(1.0); pr[0] = T(0.0); qr[0] = T(1.0); break;
    case LIN_DIFF_CYL:  // examples/Example103_LinearDiffusionCylindrical.jl:49-58
        pl[0] = T(0.0); ql[0] = T(0.0);
        pr[0] = ur[0] - p[2] * std::exp(-p[0] * (p[1] * p[1]) * t); qr[0] = T(0.0); break;
    case POISSON:       // examples/Example104_Poisson.jl:34-41
        pl[0] = ul[0] - p[2]; ql[0] = T(0.0); pr[0] = ur[0] - p[2]; qr[0] = T(0.0); break;
    case STAT_NONLIN:   // examples/Example105_StationaryNonlinearDiffusion.jl:44-51
        pl[0] = ul[0] - p[2]; ql[0] = T(0.0); pr[0] = ur[0] - p[2]; qr[0] = T(0.0); break;
    case REACT_DIFF:    // examples/Example106_SystemReactionDiffusion.jl:57-64
        pl[0] = ul[0] - 1.0; pl[1] = T(0.0); ql[0] = T(0.0); ql[1] = T(1.0);
        pr[0] = T(0.0); pr[1] = ur[1]; qr[0] = T(1.0); qr[1] = T(0.0); break;
    case LIN_DIFF_SPH:  // examples/Example107_LinearDiffusionSpherical.jl:47-55
        pl[0] = T(0.0); ql[0] = T(0.0);
        pr[0] = ur[0] - (1.0 + 6.0 * p[0] * t); qr[0] = T(0.0); break;
    case CONV_DIFF:     // examples/Example201_PartialDerivativeApprox.jl:44-51
    case LIN_REACT:     // examples/Example301_PDEOptimSciMLSensitivity.jl:41-48
        pl[0] = ul[0]; ql[0] = T(0.0); pr[0] = ur[0]; qr[0] = T(0.0); break;
    case GENERIC_SCALAR:
        pl[0] = p[7] * ul[0] - (p[8] + p[9] * t + p[10] * std::exp(-p[11] * t)); ql[0] = T(p[12]);
        pr[0] = p[13] * ur[0] - (p[14] + p[15] * t + p[16] * std::exp(-p[17] * t)); qr[0] = T(p[18]); break;
    case EXP_DIFF:      // left Dirichlet u = p2*(1+t), right homogeneous Neumann
        pl[0] = ul[0] - p[2] * (1.0 + t); ql[0] = T(0.0); pr[0] = T(0.0); qr[0] = T(1.0); break;
    case CHAIN3:        // A held at 1 on the left, everything else zero flux; C absorbed (Dirichlet 0) on the right
        pl[0] = ul[0] - 1.0; ql[0] = T(0.0); pl[1] = T(0.0); ql[1] = T(1.0); pl[2] = T(0.0); ql[2] = T(1.0);
        pr[0] = T(0.0); qr[0] = T(1.0); pr[1] = T(0.0); qr[1] = T(1.0); pr[2] = ur[2]; qr[2] = T(0.0); break;
    case CHAIN4:        // A held at 1 + t/10 on the left; D absorbed on the right; the rest zero flux
        pl[0] = ul[0] - (1.0 + 0.1 * t); ql[0] = T(0.0);
        for (int j = 1; j < 4; ++j) { pl[j] = T(0.0); ql[j] = T(1.0); }
        for (int j = 0; j < 3; ++j) { pr[j] = T(0.0); qr[j] = T(1.0); }
        pr[3] = ur[3]; qr[3] = T(0.0); break;
    }
}

// --------------------------------------------------------------------------------------------
// Problem record (ProblemDefinition, src/utils.jl:113-182) and problem_init (src/utils.jl:609-663)
// --------------------------------------------------------------------------------------------
struct Problem {
    int m, npde, Nx, functor;
    bool singular;
    const double* prm;
    std::vector<double> x, xi, zeta;
};

static inline double ipow(double x, int n) {  // Julia x^n for small Int n (x^0 == 1 even for x == 0)
    double r = 1.0;
    for (int k = 0; k < n; ++k) r *= x;
    return r;
}

// get_quad_points_weights, src/utils.jl:683-718
static void quad_points_weights(int m, bool singular, const std::vector<double>& x, std::vector<double>& xi,
                                std::vector<double>& zeta) {
    const int n = (int)x.size() - 1;
    xi.resize(n); zeta.resize(n);
    for (int i = 0; i < n; ++i) {
        const double a = x[i], b = x[i + 1], g = (a + b) / 2;  // alpha, beta, gamma  :618-620
        if (m == 0) {                                          // :686-690
            xi[i] = g; zeta[i] = g;
        } else if (m == 1) {                                   // :693-702
            if (singular) {
                xi[i] = (2.0 / 3.0) * (a + b - ((a * b) / (a + b)));
                zeta[i] = std::pow((b * b - a * a) / (2 * std::log(b / a)), 0.5);
            } else {
                xi[i] = (b - a) / std::log(b / a);
                zeta[i] = std::pow(xi[i] * g, 0.5);
            }
        } else {                                               // :705-715
            if (singular) xi[i] = (2.0 / 3.0) * (a + b - ((a * b) / (a + b)));
            else          xi[i] = (a * b * std::log(b / a)) / (b - a);
            zeta[i] = std::pow(a * b * g, 1.0 / 3.0);
        }
    }
}

static void problem_init(Problem& pb, int m, int Nx, const double* xmesh, int functor, const double* prm) {
    pb.m = m; pb.Nx = Nx; pb.functor = functor; pb.prm = prm; pb.npde = k_npde[functor];
    pb.x.assign(xmesh, xmesh + Nx);
    pb.singular = (m >= 1) && (xmesh[0] == 0);  // src/utils.jl:616
    quad_points_weights(m, pb.singular, pb.x, pb.xi, pb.zeta);
}

// interpolation, src/utils.jl:18-48 (scalar) and :67-104 (SVector): same formulas per component
template <class T>
static inline void interpolation(const Problem& pb, double xl, const T& ul, double xr, const T& ur, double q, T& u, T& dudx) {
    if (pb.singular) {                       // :19-24
        double h = xr * xr - xl * xl;
        double tmp = q * q - xl * xl;
        u = ul * (1 - (tmp / h)) + ur * (tmp / h);
        dudx = (2 * q * (ur - ul)) / h;
    } else if (pb.m == 0) {                  // :26-30
        double h = xr - xl;
        u = (ul * (xr - q) + ur * (q - xl)) / h;
        dudx = (ur - ul) / h;
    } else if (pb.m == 1) {                  // :32-37
        double tmp = std::log(xr / xl);
        double test_fct = std::log(q / xl) / tmp;
        u = ul * (1 - test_fct) + ur * test_fct;
        dudx = (ur - ul) / (q * tmp);
    } else {                                 // :39-44
        double h = xr - xl;
        double test_fct = (xr * (q - xl)) / (q * h);
        u = ul * (1 - test_fct) + ur * test_fct;
        dudx = (ur - ul) * ((xr * xl) / (q * q * h));
    }
}

// assemble!(du,u,pb,t), src/assembler.jl:19-125.  u, du flat with component fastest (:21-22).
template <class T>
static void assemble(T* du, const T* u, const Problem& pb, double t) {
    const int P = pb.npde, N = pb.Nx, m = pb.m;
    const double* x = pb.x.data();
    const double* xi = pb.xi.data();
    const double* ze = pb.zeta.data();
    T pl[4], ql[4], pr[4], qr[4], U[4], Ux[4], cl[4], fl[4], sl[4], cr[4], fr[4], sr[4];

    bdfun<T>(pb.functor, pb.prm, x[0], &u[0], x[N - 1], &u[(size_t)(N - 1) * P], t, pl, ql, pr, qr);   // :26/:29
    for (int j = 0; j < P; ++j) interpolation(pb, x[0], u[j], x[1], u[P + j], xi[0], U[j], Ux[j]);     // :27/:30
    pdefun<T>(pb.functor, pb.prm, xi[0], t, U, Ux, cl, fl, sl);                                       // :33

    double frac = (ipow(ze[0], m + 1) - ipow(x[0], m + 1)) / (m + 1);                                 // :36
    if (pb.singular) {                                                                                // :38-45
        for (int i = 0; i < P; ++i) {
            if (value(cl[i]) != 0) du[i] = ((double)(m + 1) * fl[i] / xi[0] + sl[i]) / cl[i];
            else                   du[i] = ((double)(m + 1) * fl[i] / xi[0] + sl[i]);
        }
    } else {                                                                                          // :46-57
        for (int i = 0; i < P; ++i) {
            if (value(ql[i]) != 0 && value(cl[i]) != 0)
                du[i] = (pl[i] + ql[i] / ipow(x[0], m) * (ipow(xi[0], m) * fl[i] + frac * sl[i])) /
                        (ql[i] / ipow(x[0], m) * frac * cl[i]);
            else if (value(ql[i]) != 0 && value(cl[i]) == 0)
                du[i] = (pl[i] + ql[i] / ipow(x[0], m) * (ipow(xi[0], m) * fl[i] + frac * sl[i]));
            else
                du[i] = pl[i];
        }
    }

    for (int i = 1; i < N - 1; ++i) {                                                                 // :60-95
        for (int j = 0; j < P; ++j)
            interpolation(pb, x[i], u[(size_t)i * P + j], x[i + 1], u[(size_t)(i + 1) * P + j], xi[i], U[j], Ux[j]);
        pdefun<T>(pb.functor, pb.prm, xi[i], t, U, Ux, cr, fr, sr);                                   // :66
        double frac1 = (ipow(ze[i], m + 1) - ipow(x[i], m + 1)) / (m + 1);                            // :68
        double frac2 = (ipow(x[i], m + 1) - ipow(ze[i - 1], m + 1)) / (m + 1);                        // :69
        for (int j = 0; j < P; ++j) {
            T num;
            if (pb.singular)                                                                          // :74
                num = ipow(ze[i], m + 1) / xi[i] * fr[j] - ipow(ze[i - 1], m + 1) / xi[i - 1] * fl[j] +
                      frac1 * sr[j] + frac2 * sl[j];
            else                                                                                      // :84
                num = ipow(xi[i], m) * fr[j] - ipow(xi[i - 1], m) * fl[j] + frac1 * sr[j] + frac2 * sl[j];
            if (value(cl[j]) != 0 || value(cr[j]) != 0) du[(size_t)i * P + j] = num / (frac1 * cr[j] + frac2 * cl[j]);
            else                                        du[(size_t)i * P + j] = num;
        }
        for (int j = 0; j < P; ++j) { cl[j] = cr[j]; fl[j] = fr[j]; sl[j] = sr[j]; }                  // :92-94
    }

    frac = (ipow(x[N - 1], m + 1) - ipow(ze[N - 2], m + 1)) / (m + 1);                                // :98
    T* dN = &du[(size_t)(N - 1) * P];
    for (int i = 0; i < P; ++i) {                                                                     // :100-122
        double w = pb.singular ? ipow(ze[N - 2], m + 1) / xi[N - 2] : ipow(xi[N - 2], m);
        if (value(qr[i]) != 0 && value(cl[i]) != 0)
            dN[i] = (pr[i] + qr[i] / ipow(x[N - 1], m) * (w * fl[i] - frac * sl[i])) /
                    (-qr[i] / ipow(x[N - 1], m) * frac * cl[i]);
        else if (value(qr[i]) != 0 && value(cl[i]) == 0)
            dN[i] = (pr[i] + qr[i] / ipow(x[N - 1], m) * (w * fl[i] - frac * sl[i]));
        else
            dN[i] = pr[i];
    }
}

// mass_matrix, src/assembler.jl:138-174: diagonal of M from the initial value at tspan[1]
static void mass_matrix(const Problem& pb, const double* inival, double t0, double* M) {
    const int P = pb.npde, N = pb.Nx;
    for (size_t k = 0; k < (size_t)P * N; ++k) M[k] = 1.0;                                            // :141
    double pl[4], ql[4], pr[4], qr[4], U[4], Ux[4], c[4], f[4], s[4];
    bdfun<double>(pb.functor, pb.prm, pb.x[0], &inival[0], pb.x[N - 1], &inival[(size_t)(N - 1) * P], t0, pl, ql, pr, qr);
    for (int j = 0; j < P; ++j) interpolation(pb, pb.x[0], inival[j], pb.x[1], inival[P + j], pb.xi[0], U[j], Ux[j]);
    pdefun<double>(pb.functor, pb.prm, pb.xi[0], t0, U, Ux, c, f, s);                                 // :154
    for (int i = 0; i < P; ++i) {
        if (c[i] == 0) for (int n = 1; n < N - 1; ++n) M[(size_t)n * P + i] = 0;                      // :157-160
        if (ql[i] == 0 && !pb.singular) M[i] = 0;                                                     // :162-165
        if (qr[i] == 0) M[(size_t)(N - 1) * P + i] = 0;                                               // :167-170
    }
}

// implicitEuler!, src/utils.jl:245-257 (mass != nullptr) and implicitEuler_stat!, :264-276 (mass == nullptr)
template <class T>
static void implicit_euler(T* y, const T* u, const Problem& pb, double tau, const double* mass, double t) {
    assemble<T>(y, u, pb, t);
    const size_t n = (size_t)pb.npde * pb.Nx;
    for (size_t k = 0; k < n; ++k) {
        y[k] = y[k] * -1.0;
        if (mass) y[k] = y[k] + ((1 / tau) * mass[k]) * u[k];
        else      y[k] = y[k] + (1 / tau) * u[k];
    }
}

// --------------------------------------------------------------------------------------------
// Banded LU with partial pivoting (LAPACK dgbtf2/dgbtrs restated; what :banded + LUFactorization
// runs, test/test_solveTransient.jl:54-58, and numerically what KLU does on these matrices)
// A is n x n with kl = ku = bw; storage ab[(2*bw+bw+1) x n], column-major LAPACK band layout.
// --------------------------------------------------------------------------------------------
struct Banded {
    int n, bw, ld;
    std::vector<double> ab;
    std::vector<int> piv;
    void init(int n_, int bw_) { n = n_; bw = bw_; ld = 3 * bw + 1; ab.assign((size_t)ld * n, 0.0); piv.resize(n); }
    inline double& at(int i, int j) { return ab[(size_t)j * ld + (2 * bw + i - j)]; }  // A(i,j), |i-j| within band(+fill)
    void clear() { std::fill(ab.begin(), ab.end(), 0.0); }
    int factor() {
        const int kv = 2 * bw;
        for (int j = 0; j < n; ++j) {
            int km = std::min(bw, n - 1 - j);
            int jp = 0; double amax = std::fabs(at(j, j));
            for (int k = 1; k <= km; ++k) { double a = std::fabs(at(j + k, j)); if (a > amax) { amax = a; jp = k; } }
            piv[j] = j + jp;
            if (at(j + jp, j) == 0.0) return j + 1;
            int ju = std::min(j + kv, n - 1);
            if (jp != 0) for (int c = j; c <= ju; ++c) std::swap(at(j, c), at(j + jp, c));
            double r = 1.0 / at(j, j);
            for (int k = 1; k <= km; ++k) at(j + k, j) *= r;
            for (int c = j + 1; c <= ju; ++c) {
                double ajc = at(j, c);
                if (ajc != 0.0) for (int k = 1; k <= km; ++k) at(j + k, c) -= at(j + k, j) * ajc;
            }
        }
        return 0;
    }
    void solve(double* b) {
        for (int j = 0; j < n; ++j) {
            int km = std::min(bw, n - 1 - j);
            if (piv[j] != j) std::swap(b[j], b[piv[j]]);
            for (int k = 1; k <= km; ++k) b[j + k] -= at(j + k, j) * b[j];
        }
        const int kv = 2 * bw;
        for (int j = n - 1; j >= 0; --j) {
            b[j] /= at(j, j);
            int lo = std::max(0, j - kv);
            for (int i = lo; i < j; ++i) b[i] -= at(i, j) * b[j];
        }
    }
};

// --------------------------------------------------------------------------------------------
// Residual + Jacobian by one dual sweep with 3*npde colours
// (forwarddiff_color_jacobian! + value! + rhs update, src/utils.jl:312-314 ; colours src/pdepe.jl:66)
// Output in block-tridiagonal form: Jl/Jd/Ju[i][r][c] = d y_{i,r} / d u_{i-1|i|i+1, c}
// --------------------------------------------------------------------------------------------
template <int P>
static void residual_jacobian(const Problem& pb, const double* u, const double* b, double tau, const double* mass,
                              double t, double* F, double* Jl, double* Jd, double* Ju) {
    constexpr int NC = 3 * P;
    const int N = pb.Nx;
    const size_t n = (size_t)P * N;
    std::vector<Dual<NC>> ud(n), yd(n);
    for (size_t k = 0; k < n; ++k) { ud[k] = Dual<NC>(u[k]); ud[k].d[k % NC] = 1.0; }
    implicit_euler<Dual<NC>>(yd.data(), ud.data(), pb, tau, mass, t);
    for (size_t k = 0; k < n; ++k) F[k] = yd[k].v - (1. / tau) * b[k];                               // :313-314
    for (int i = 0; i < N; ++i)
        for (int r = 0; r < P; ++r) {
            const Dual<NC>& y = yd[(size_t)i * P + r];
            for (int c = 0; c < P; ++c) {
                size_t o = ((size_t)i * P + r) * P + c;
                Jl[o] = (i > 0) ? y.d[((size_t)(i - 1) * P + c) % NC] : 0.0;
                Jd[o] = y.d[((size_t)i * P + c) % NC];
                Ju[o] = (i < N - 1) ? y.d[((size_t)(i + 1) * P + c) % NC] : 0.0;
            }
        }
}

struct NewtonWork {
    std::vector<double> F, Jl, Jd, Ju;
    Banded A;
};

// newton, src/utils.jl:303-337 / newton_stat, :344-378 (mass == nullptr).
// returns number of iterations (>0) or -1 on "convergence failed" (:336).  hist (may be null) gets ||h||_2 of
// the first hist_cap iterations.
template <int P>
static int newton(const Problem& pb, const double* b, double tau, double t, const double* mass, double tol, int maxit,
                  double* unP1, double* hist, NewtonWork& w, int hist_cap = 1 << 30) {
    const int N = pb.Nx;
    const size_t n = (size_t)P * N;
    w.F.resize(n); w.Jl.resize(n * P); w.Jd.resize(n * P); w.Ju.resize(n * P);
    if (w.A.n != (int)n || w.A.bw != 2 * P - 1) w.A.init((int)n, 2 * P - 1);
    std::memcpy(unP1, b, n * sizeof(double));                                                         // :309
    for (int it = 0; it < maxit; ++it) {
        residual_jacobian<P>(pb, unP1, b, tau, mass, t, w.F.data(), w.Jl.data(), w.Jd.data(), w.Ju.data());
        w.A.clear();
        for (int i = 0; i < N; ++i)
            for (int r = 0; r < P; ++r)
                for (int c = 0; c < P; ++c) {
                    size_t o = ((size_t)i * P + r) * P + c;
                    int row = i * P + r;
                    if (i > 0) w.A.at(row, (i - 1) * P + c) = w.Jl[o];
                    w.A.at(row, i * P + c) = w.Jd[o];
                    if (i < N - 1) w.A.at(row, (i + 1) * P + c) = w.Ju[o];
                }
        if (w.A.factor() != 0) return -2;
        w.A.solve(w.F.data());                                                                        // :317-319  h
        double ss = 0.0;
        for (size_t k = 0; k < n; ++k) { unP1[k] -= w.F[k]; ss += w.F[k] * w.F[k]; }                  // :321
        double nm = std::sqrt(ss);                                                                    // :323
        if (hist && it < hist_cap) hist[it] = nm;
        if (nm < tol) return it + 1;                                                                  // :329-333
    }
    return -1;                                                                                        // :336
}

// transient loop, src/pdepe.jl:78-108.  tgrid[nt] is the time grid; tau_scalar > 0: constant tstep (:92),
// else tau_i = diff(tgrid).  u (in: initial value, out: final state).  snapshots (may be null): nt*n TRUE states.
template <int P>
static int solve_transient(const Problem& pb, double* u, int nt, const double* tgrid, double tau_scalar, double tol,
                           int maxit, int* iters, double* hist, int hist_cap, double* snapshots) {
    const size_t n = (size_t)P * pb.Nx;
    std::vector<double> mass(n), un(u, u + n), unP1(n);
    mass_matrix(pb, u, tgrid[0], mass.data());                                                        // :78-79
    NewtonWork w; w.A.n = -1; w.A.bw = -1;
    if (snapshots) std::memcpy(snapshots, u, n * sizeof(double));
    for (int i = 0; i + 1 < nt; ++i) {                                                                // :90
        double tval = tgrid[i + 1];
        double tau = tau_scalar > 0 ? tau_scalar : tgrid[i + 1] - tgrid[i];                           // :92
        for (size_t k = 0; k < n; ++k) un[k] = mass[k] * un[k];                                       // :94 lmul!
        int it = newton<P>(pb, un.data(), tau, tval, mass.data(), tol, maxit, unP1.data(),
                           hist ? hist + (size_t)i * hist_cap : nullptr, w, hist_cap);
        if (iters) iters[i] = it;
        if (it < 0) return it;
        un = unP1;                                                                                    // :106
        if (snapshots) std::memcpy(snapshots + (size_t)(i + 1) * n, un.data(), n * sizeof(double));
    }
    std::memcpy(u, un.data(), n * sizeof(double));
    return 0;
}

// stationary, src/pdepe.jl:156-181: newton_stat(inival, Inf, 1, ...)
template <int P>
static int solve_stationary(const Problem& pb, double* u, double tol, int maxit, int* iters, double* hist) {
    const size_t n = (size_t)P * pb.Nx;
    std::vector<double> b(u, u + n), unP1(n);
    NewtonWork w; w.A.n = -1; w.A.bw = -1;
    int it = newton<P>(pb, b.data(), std::numeric_limits<double>::infinity(), 1.0, nullptr, tol, maxit, unP1.data(), hist, w);
    if (iters) *iters = it;
    if (it < 0) return it;
    std::memcpy(u, unP1.data(), n * sizeof(double));
    return 0;
}

// --------------------------------------------------------------------------------------------
// C entry points (ctypes).  All arrays in reference layout: [instance][node][component].
// --------------------------------------------------------------------------------------------
extern "C" {

int sbo_functor_info(int id, int* npde, int* nprm) {
    if (id < 0 || id >= NUM_FUNCTORS) return -1;
    *npde = k_npde[id]; *nprm = k_nprm[id];
    return 0;
}

int sbo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int sbo_quadrature(int m, int Nx, const double* xmesh, double* xi, double* zeta, int* singular) {
    Problem pb; problem_init(pb, m, Nx, xmesh, 0, nullptr);
    std::memcpy(xi, pb.xi.data(), sizeof(double) * (Nx - 1));
    std::memcpy(zeta, pb.zeta.data(), sizeof(double) * (Nx - 1));
    *singular = pb.singular;
    return 0;
}

// assemble! on B instances (prm[B][nprm])
int sbo_assemble(int functor, int m, int Nx, const double* xmesh, int64_t B, const double* prm, const double* u,
                 double t, double* du) {
    if (functor < 0 || functor >= NUM_FUNCTORS) return -1;
    const size_t n = (size_t)k_npde[functor] * Nx;
    for (int64_t k = 0; k < B; ++k) {
        Problem pb; problem_init(pb, m, Nx, xmesh, functor, prm + k * k_nprm[functor]);
        assemble<double>(du + k * n, u + k * n, pb, t);
    }
    return 0;
}

int sbo_mass(int functor, int m, int Nx, const double* xmesh, int64_t B, const double* prm, const double* u0,
             double t0, double* mass) {
    if (functor < 0 || functor >= NUM_FUNCTORS) return -1;
    const size_t n = (size_t)k_npde[functor] * Nx;
    for (int64_t k = 0; k < B; ++k) {
        Problem pb; problem_init(pb, m, Nx, xmesh, functor, prm + k * k_nprm[functor]);
        mass_matrix(pb, u0 + k * n, t0, mass + k * n);
    }
    return 0;
}

// F and block-tridiagonal J of one Newton iteration; mass == NULL selects the stationary residual
int sbo_residual_jacobian(int functor, int m, int Nx, const double* xmesh, int64_t B, const double* prm,
                          const double* u, const double* b, double tau, const double* mass, double t,
                          double* F, double* Jl, double* Jd, double* Ju) {
    if (functor < 0 || functor >= NUM_FUNCTORS) return -1;
    const int P = k_npde[functor];
    const size_t n = (size_t)P * Nx;
    for (int64_t k = 0; k < B; ++k) {
        Problem pb; problem_init(pb, m, Nx, xmesh, functor, prm + k * k_nprm[functor]);
        const double* mk = mass ? mass + k * n : nullptr;
        switch (P) {
        case 1: residual_jacobian<1>(pb, u + k * n, b + k * n, tau, mk, t, F + k * n, Jl + k * n * P, Jd + k * n * P, Ju + k * n * P); break;
        case 2: residual_jacobian<2>(pb, u + k * n, b + k * n, tau, mk, t, F + k * n, Jl + k * n * P, Jd + k * n * P, Ju + k * n * P); break;
        case 3: residual_jacobian<3>(pb, u + k * n, b + k * n, tau, mk, t, F + k * n, Jl + k * n * P, Jd + k * n * P, Ju + k * n * P); break;
        default: residual_jacobian<4>(pb, u + k * n, b + k * n, tau, mk, t, F + k * n, Jl + k * n * P, Jd + k * n * P, Ju + k * n * P); break;
        }
    }
    return 0;
}

// banded LU solve of block-tridiagonal systems given in the layout above (checker for the GPU solvers)
int sbo_block_tridiag_solve(int P, int Nx, int64_t B, const double* Jl, const double* Jd, const double* Ju,
                            const double* F, double* x) {
    const size_t n = (size_t)P * Nx;
    int rc = 0;
#pragma omp parallel for schedule(dynamic)
    for (int64_t k = 0; k < B; ++k) {
        Banded A; A.init((int)n, 2 * P - 1);
        const double *jl = Jl + k * n * P, *jd = Jd + k * n * P, *ju = Ju + k * n * P;
        for (int i = 0; i < Nx; ++i)
            for (int r = 0; r < P; ++r)
                for (int c = 0; c < P; ++c) {
                    size_t o = ((size_t)i * P + r) * P + c;
                    if (i > 0) A.at(i * P + r, (i - 1) * P + c) = jl[o];
                    A.at(i * P + r, i * P + c) = jd[o];
                    if (i < Nx - 1) A.at(i * P + r, (i + 1) * P + c) = ju[o];
                }
        std::memcpy(x + k * n, F + k * n, n * sizeof(double));
        if (A.factor() != 0) { rc = -2; continue; }
        A.solve(x + k * n);
    }
    return rc;
}

// pdeval(m, xmesh, u_approx, x_eval, pb), src/utils.jl:396-438, for B instances and nq evaluation points:
// uo/dxo[B][nq][P].  Returns -3 for a point outside the mesh ("Evaluation point oustide of the spatial discretization.").
int sbo_pdeval(int functor, int m, int Nx, const double* xmesh, int64_t B, const double* u, int nq, const double* xq,
               double* uo, double* dxo) {
    if (functor < 0 || functor >= NUM_FUNCTORS) return -1;
    const int P = k_npde[functor];
    Problem pb; problem_init(pb, m, Nx, xmesh, functor, nullptr);
    for (int64_t k = 0; k < B; ++k)
        for (int q = 0; q < nq; ++q) {
            const double pt = xq[q];
            int il;
            if (pt == xmesh[0] || std::fabs(pt - xmesh[0]) <= 1.0e-10 * std::fabs(xmesh[1] - xmesh[0])) {   // :403 isapprox(...; atol)
                il = 0;
            } else {
                const int idx = (int)(std::lower_bound(xmesh, xmesh + Nx, pt) - xmesh) + 1;               // :410 searchsortedfirst (1-based)
                if (idx == 1 || idx > Nx) return -3;                                                        // :412-414
                il = idx - 2;                                                                                // interval [idx-1, idx] (1-based)
            }
            for (int j = 0; j < P; ++j) {
                double U, Ux;
                interpolation(pb, xmesh[il], u[((size_t)k * Nx + il) * P + j], xmesh[il + 1], u[((size_t)k * Nx + il + 1) * P + j], pt, U, Ux);
                uo[((size_t)k * nq + q) * P + j] = U; dxo[((size_t)k * nq + q) * P + j] = Ux;
            }
        }
    return 0;
}

// Batched transient solve: the reference pdepe (src/pdepe.jl:42-128) looped over B instances (OpenMP over instances).
// u[B][n] in: initial values, out: final states.  iters[B][nt-1], hist[B][nt-1][hist_cap] (nullable),
// snapshots[B][nt][n] (nullable).  status[B]: 0 ok, -1 convergence failed, -2 singular Jacobian.
int sbo_pdepe_transient(int functor, int m, int Nx, const double* xmesh, int64_t B, const double* prm, double* u,
                        int nt, const double* tgrid, double tau_scalar, double tol, int maxit, int* iters,
                        double* hist, int hist_cap, double* snapshots, int* status, int nthreads) {
    if (functor < 0 || functor >= NUM_FUNCTORS) return -1;
    const int P = k_npde[functor];
    const size_t n = (size_t)P * Nx;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    (void)nthreads;
#pragma omp parallel for schedule(dynamic)
    for (int64_t k = 0; k < B; ++k) {
        Problem pb; problem_init(pb, m, Nx, xmesh, functor, prm + k * k_nprm[functor]);
        int* itk = iters ? iters + k * (nt - 1) : nullptr;
        double* hk = hist ? hist + (size_t)k * (nt - 1) * hist_cap : nullptr;
        double* sk = snapshots ? snapshots + (size_t)k * nt * n : nullptr;
        int rc = (P == 1) ? solve_transient<1>(pb, u + k * n, nt, tgrid, tau_scalar, tol, maxit, itk, hk, hist_cap, sk)
               : (P == 2) ? solve_transient<2>(pb, u + k * n, nt, tgrid, tau_scalar, tol, maxit, itk, hk, hist_cap, sk)
               : (P == 3) ? solve_transient<3>(pb, u + k * n, nt, tgrid, tau_scalar, tol, maxit, itk, hk, hist_cap, sk)
                          : solve_transient<4>(pb, u + k * n, nt, tgrid, tau_scalar, tol, maxit, itk, hk, hist_cap, sk);
        if (status) status[k] = rc;
    }
    return 0;
}

// Batched stationary solve (src/pdepe.jl:156-192).  iters[B], hist[B][maxit] (nullable)
int sbo_pdepe_stationary(int functor, int m, int Nx, const double* xmesh, int64_t B, const double* prm, double* u,
                         double tol, int maxit, int* iters, double* hist, int* status, int nthreads) {
    if (functor < 0 || functor >= NUM_FUNCTORS) return -1;
    const int P = k_npde[functor];
    const size_t n = (size_t)P * Nx;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    (void)nthreads;
#pragma omp parallel for schedule(dynamic)
    for (int64_t k = 0; k < B; ++k) {
        Problem pb; problem_init(pb, m, Nx, xmesh, functor, prm + k * k_nprm[functor]);
        int rc = (P == 1) ? solve_stationary<1>(pb, u + k * n, tol, maxit, iters ? iters + k : nullptr, hist ? hist + (size_t)k * maxit : nullptr)
               : (P == 2) ? solve_stationary<2>(pb, u + k * n, tol, maxit, iters ? iters + k : nullptr, hist ? hist + (size_t)k * maxit : nullptr)
               : (P == 3) ? solve_stationary<3>(pb, u + k * n, tol, maxit, iters ? iters + k : nullptr, hist ? hist + (size_t)k * maxit : nullptr)
                          : solve_stationary<4>(pb, u + k * n, tol, maxit, iters ? iters + k : nullptr, hist ? hist + (size_t)k * maxit : nullptr);
        if (status) status[k] = rc;
    }
    return 0;
}

}  // extern "C"
