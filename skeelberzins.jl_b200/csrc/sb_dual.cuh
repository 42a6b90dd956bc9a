// sb_dual.cuh -- forward-mode dual numbers and tiny dense matrices for the sm_100a kernels.
//
// Dual<N> plays the role ForwardDiff.Dual plays in the reference (the residual is evaluated once
// with dual arguments, src/utils.jl:312, so value and Jacobian come out of one sweep).  Here the
// seeds are the two end values of a mesh interval (2*npde partials), see sb_assemble.cuh.
// Dual<0> degenerates to a plain double and is used by the residual-only kernel (assemble!).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#else  // NVRTC (user-functor plugin): no system headers
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned int uint32_t;
typedef long ptrdiff_t;
#endif

#define SB_HD __host__ __device__ __forceinline__
#define SB_D __device__ __forceinline__

namespace sb {

template <int N>
struct Dual {
    double v;
    double d[N > 0 ? N : 1];
    SB_HD Dual() {}
    SB_HD Dual(double x) : v(x) {
#pragma unroll
        for (int k = 0; k < N; ++k) d[k] = 0.0;
    }
};

template <int N> SB_HD Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v + b.v;
#pragma unroll
    for (int k = 0; k < N; ++k) r.d[k] = a.d[k] + b.d[k];
    return r;
}
template <int N> SB_HD Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v - b.v;
#pragma unroll
    for (int k = 0; k < N; ++k) r.d[k] = a.d[k] - b.d[k];
    return r;
}
template <int N> SB_HD Dual<N> operator-(const Dual<N>& a) {
    Dual<N> r; r.v = -a.v;
#pragma unroll
    for (int k = 0; k < N; ++k) r.d[k] = -a.d[k];
    return r;
}
template <int N> SB_HD Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v * b.v;
#pragma unroll
    for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
    return r;
}
template <int N> SB_HD Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; const double ib = 1.0 / b.v; r.v = a.v * ib;
#pragma unroll
    for (int k = 0; k < N; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * ib;
    return r;
}
template <int N> SB_HD Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> SB_HD Dual<N> operator+(double a, const Dual<N>& b) { Dual<N> r = b; r.v += a; return r; }
template <int N> SB_HD Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> SB_HD Dual<N> operator-(double a, const Dual<N>& b) { Dual<N> r = -b; r.v += a; return r; }
template <int N> SB_HD Dual<N> operator*(const Dual<N>& a, double b) {
    Dual<N> r; r.v = a.v * b;
#pragma unroll
    for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * b;
    return r;
}
template <int N> SB_HD Dual<N> operator*(double a, const Dual<N>& b) { return b * a; }
template <int N> SB_HD Dual<N> operator/(const Dual<N>& a, double b) { return a * (1.0 / b); }
template <int N> SB_HD Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }
template <int N> SB_HD Dual<N> exp(const Dual<N>& a) {
    Dual<N> r; r.v = ::exp(a.v);
#pragma unroll
    for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * r.v;
    return r;
}
template <int N> SB_HD Dual<N> log(const Dual<N>& a) {
    Dual<N> r; r.v = ::log(a.v); const double ia = 1.0 / a.v;
#pragma unroll
    for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * ia;
    return r;
}
template <int N> SB_HD Dual<N> sqrt(const Dual<N>& a) {
    Dual<N> r; r.v = ::sqrt(a.v); const double h = 0.5 / r.v;
#pragma unroll
    for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * h;
    return r;
}

// ---------------------------------------------------------------------------------------------
// P x P blocks (row-major) and P-vectors; P = 1 collapses to scalars after unrolling
// ---------------------------------------------------------------------------------------------
template <int P> struct Mat { double a[P * P]; };
template <int P> struct Vec { double a[P]; };

template <int P> SB_HD Mat<P> mat_zero() { Mat<P> r;
#pragma unroll
    for (int i = 0; i < P * P; ++i) r.a[i] = 0.0;
    return r; }
template <int P> SB_HD Mat<P> mat_identity() { Mat<P> r = mat_zero<P>();
#pragma unroll
    for (int i = 0; i < P; ++i) r.a[i * P + i] = 1.0;
    return r; }
template <int P> SB_HD Vec<P> vec_zero() { Vec<P> r;
#pragma unroll
    for (int i = 0; i < P; ++i) r.a[i] = 0.0;
    return r; }

// C = A*B
template <int P> SB_HD Mat<P> mul(const Mat<P>& A, const Mat<P>& B) { Mat<P> C;
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int j = 0; j < P; ++j) { double s = A.a[i * P] * B.a[j];
#pragma unroll
            for (int k = 1; k < P; ++k) s += A.a[i * P + k] * B.a[k * P + j];
            C.a[i * P + j] = s; }
    return C; }
template <int P> SB_HD Vec<P> mul(const Mat<P>& A, const Vec<P>& x) { Vec<P> y;
#pragma unroll
    for (int i = 0; i < P; ++i) { double s = A.a[i * P] * x.a[0];
#pragma unroll
        for (int k = 1; k < P; ++k) s += A.a[i * P + k] * x.a[k];
        y.a[i] = s; }
    return y; }
// C = A - B*C2
template <int P> SB_HD Mat<P> sub_mul(const Mat<P>& A, const Mat<P>& B, const Mat<P>& C2) { Mat<P> C;
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int j = 0; j < P; ++j) { double s = A.a[i * P + j];
#pragma unroll
            for (int k = 0; k < P; ++k) s -= B.a[i * P + k] * C2.a[k * P + j];
            C.a[i * P + j] = s; }
    return C; }
template <int P> SB_HD Vec<P> sub_mul(const Vec<P>& a, const Mat<P>& B, const Vec<P>& x) { Vec<P> y;
#pragma unroll
    for (int i = 0; i < P; ++i) { double s = a.a[i];
#pragma unroll
        for (int k = 0; k < P; ++k) s -= B.a[i * P + k] * x.a[k];
        y.a[i] = s; }
    return y; }
template <int P> SB_HD Mat<P> neg(const Mat<P>& A) { Mat<P> C;
#pragma unroll
    for (int i = 0; i < P * P; ++i) C.a[i] = -A.a[i];
    return C; }

// inverse by Gauss-Jordan with partial pivoting (P <= 4, fully unrolled; P = 1: reciprocal)
template <int P> SB_HD Mat<P> inverse(const Mat<P>& A) {
    if constexpr (P == 1) { Mat<P> r; r.a[0] = 1.0 / A.a[0]; return r; }
    else if constexpr (P == 2) {
        // pivot on the larger entry of the first column
        const bool sw = fabs(A.a[2]) > fabs(A.a[0]);
        const double a = sw ? A.a[2] : A.a[0], b = sw ? A.a[3] : A.a[1];
        const double c = sw ? A.a[0] : A.a[2], d = sw ? A.a[1] : A.a[3];
        const double ia = 1.0 / a;
        const double l = c * ia;           // multiplier
        const double s = d - l * b;        // Schur complement
        const double is = 1.0 / s;
        // inverse of the (row-permuted) matrix [[a,b],[c,d]]: columns then un-permute
        const double i11 = ia + b * ia * l * is, i12 = -b * ia * is, i21 = -l * is, i22 = is;
        Mat<P> r;
        r.a[0] = sw ? i12 : i11; r.a[1] = sw ? i11 : i12;
        r.a[2] = sw ? i22 : i21; r.a[3] = sw ? i21 : i22;
        return r;
    } else {
    double M[P][2 * P];
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int j = 0; j < P; ++j) { M[i][j] = A.a[i * P + j]; M[i][P + j] = (i == j) ? 1.0 : 0.0; }
#pragma unroll
    for (int c = 0; c < P; ++c) {
        int p = c; double best = fabs(M[c][c]);
#pragma unroll
        for (int i = c + 1; i < P; ++i) { const double v = fabs(M[i][c]); if (v > best) { best = v; p = i; } }
#pragma unroll
        for (int i = c + 1; i < P; ++i)
            if (p == i) {
#pragma unroll
                for (int j = 0; j < 2 * P; ++j) { const double tmp = M[c][j]; M[c][j] = M[i][j]; M[i][j] = tmp; }
            }
        const double ip = 1.0 / M[c][c];
#pragma unroll
        for (int j = 0; j < 2 * P; ++j) M[c][j] *= ip;
#pragma unroll
        for (int i = 0; i < P; ++i)
            if (i != c) { const double f = M[i][c];
#pragma unroll
                for (int j = 0; j < 2 * P; ++j) M[i][j] -= f * M[c][j]; }
    }
    Mat<P> r;
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int j = 0; j < P; ++j) r.a[i * P + j] = M[i][P + j];
    return r;
    }
}

}  // namespace sb
