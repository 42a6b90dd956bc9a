// sb_banded.cuh -- sequential banded LU with partial pivoting (the reference's linear solve).
//
// The Newton kernels eliminate without pivoting between block rows, which is exact enough on the (block) diagonally
// dominant matrices this discretisation normally produces and is what makes them parallel.  The reference solves with
// pivoted LU (KLU or, with sparsity = :banded, LAPACK's banded LU: src/utils.jl:317-319).  When the pivot-growth
// monitor of the fast path trips, the instance is solved again by this routine -- one thread, LAPACK dgbtf2 / dgbtrs
// restated (unblocked right-looking banded LU with row interchanges, forward and back substitution folded in) -- so
// ill-conditioned or non-dominant systems get the reference's algorithm instead of lost digits.
//
// Storage: column-major band, LAPACK layout: AB[(kl + ku + i - j) + j*ldab] = A(i, j), ldab = 2*kl + ku + 1; the first
// kl rows of every column are fill-in space.  __host__ __device__ so that the CPU tests can exercise it without a GPU.
#pragma once
#include "sb_dual.cuh"

namespace sb {

// rows a block-tridiagonal matrix with P x P blocks needs in band storage, and its bandwidths
template <int P> struct BandShape { static constexpr int kl = 2 * P - 1, ku = 2 * P - 1, ldab = 2 * kl + ku + 1; };

// Solve A x = rhs in place (rhs <- x).  Returns 0, or j+1 if the pivot of column j is exactly zero (singular matrix:
// rhs is then unspecified).  ab is overwritten by the factors.
// Ptr: double* on the host; volatile double* in the kernels (the band was written by other threads of the CTA)
template <int KL, int KU, class Ptr>
__host__ __device__ inline int banded_lu_solve(Ptr ab, int n, Ptr rhs) {
    constexpr int KV = KL + KU, LD = 2 * KL + KU + 1;
    int ju = 0, info = 0;
    for (int j = 0; j < n; ++j) {
        Ptr col = ab + (size_t)j * LD;
        const int km = (KL < n - 1 - j) ? KL : n - 1 - j;
        int jp = 0;                                             // pivot search in A(j..j+km, j)
        double best = col[KV] < 0 ? -col[KV] : col[KV];
        for (int p = 1; p <= km; ++p) {
            const double v = col[KV + p] < 0 ? -col[KV + p] : col[KV + p];
            if (v > best) { best = v; jp = p; }
        }
        if (col[KV + jp] == 0.0) { if (!info) info = j + 1; continue; }
        { const int c = j + KU + jp; const int cm = c < n - 1 ? c : n - 1; if (cm > ju) ju = cm; }
        if (jp != 0) {                                          // interchange rows j and j+jp over columns j..ju, and the rhs
            for (int c = j; c <= ju; ++c) {
                Ptr cc = ab + (size_t)c * LD;
                const int r0 = KV - (c - j);
                const double tmp = cc[r0 + jp]; cc[r0 + jp] = cc[r0]; cc[r0] = tmp;
            }
            const double tr = rhs[j + jp]; rhs[j + jp] = rhs[j]; rhs[j] = tr;
        }
        if (km > 0) {
            const double ip = 1.0 / col[KV];
            for (int p = 1; p <= km; ++p) col[KV + p] = col[KV + p] * ip;    // multipliers
            for (int c = j + 1; c <= ju; ++c) {                 // rank-1 update of the trailing band
                Ptr cc = ab + (size_t)c * LD;
                const int r0 = KV - (c - j);
                const double ujc = cc[r0];
                if (ujc != 0.0)
                    for (int p = 1; p <= km; ++p) cc[r0 + p] = cc[r0 + p] - col[KV + p] * ujc;
            }
            const double rj = rhs[j];                           // L^{-1} applied to the rhs on the fly
            for (int p = 1; p <= km; ++p) rhs[j + p] = rhs[j + p] - col[KV + p] * rj;
        }
    }
    if (info) return info;
    for (int j = n - 1; j >= 0; --j) {                          // U x = y
        Ptr col = ab + (size_t)j * LD;
        const double xj = rhs[j] / col[KV];
        rhs[j] = xj;
        const int lm = (KV < j) ? KV : j;
        for (int p = 1; p <= lm; ++p) rhs[j - p] = rhs[j - p] - col[KV - p] * xj;
    }
    return 0;
}

// position of block-tridiagonal entry (block row i, row r of the block) x (block column i + dj, column c), dj in
// {-1, 0, 1}, in the band array of BandShape<P>
template <int P>
__host__ __device__ inline size_t band_index(int i, int r, int dj, int c) {
    constexpr int kl = BandShape<P>::kl, ku = BandShape<P>::ku, ld = BandShape<P>::ldab;
    const int row = i * P + r, col = (i + dj) * P + c;
    return (size_t)col * ld + (size_t)(kl + ku + row - col);
}

}  // namespace sb
