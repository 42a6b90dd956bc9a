// Host-side exercise of csrc/sb_banded.cuh (the pivoted fallback solver is __host__ __device__): random block-tridiagonal
// systems WITHOUT diagonal dominance, including exactly singular leading blocks that only row interchanges get past.
// Prints "P n worst_relative_residual worst_error_vs_dense" per case; tests/test_banded_host.py asserts on them.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../skeelberzins.jl_b200/csrc/sb_banded.cuh"

using namespace sb;

static double urand() { return 2.0 * rand() / RAND_MAX - 1.0; }

// dense Gaussian elimination with partial pivoting (reference answer)
static void dense_solve(std::vector<double> A, std::vector<double> b, int n, std::vector<double>& x) {
    for (int j = 0; j < n; ++j) {
        int p = j;
        for (int i = j + 1; i < n; ++i) if (std::fabs(A[(size_t)i * n + j]) > std::fabs(A[(size_t)p * n + j])) p = i;
        if (p != j) { for (int c = 0; c < n; ++c) std::swap(A[(size_t)p * n + c], A[(size_t)j * n + c]); std::swap(b[p], b[j]); }
        for (int i = j + 1; i < n; ++i) {
            const double f = A[(size_t)i * n + j] / A[(size_t)j * n + j];
            if (f == 0.0) continue;
            for (int c = j; c < n; ++c) A[(size_t)i * n + c] -= f * A[(size_t)j * n + c];
            b[i] -= f * b[j];
        }
    }
    x.assign(n, 0.0);
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int c = i + 1; c < n; ++c) s -= A[(size_t)i * n + c] * x[c];
        x[i] = s / A[(size_t)i * n + i];
    }
}

template <int P>
static void run(int N, int flavour) {
    const int n = N * P;
    constexpr int ld = BandShape<P>::ldab;
    std::vector<double> A((size_t)n * n, 0.0), b(n), ab((size_t)n * ld, 0.0);
    for (int i = 0; i < N; ++i)
        for (int dj = -1; dj <= 1; ++dj) {
            if (i + dj < 0 || i + dj >= N) continue;
            for (int r = 0; r < P; ++r)
                for (int c = 0; c < P; ++c) {
                    double v = urand();
                    if (flavour == 1 && dj == 0) v *= 1e-3;                 // off-diagonal blocks dominate
                    if (flavour == 2 && dj == 0 && r == c && i % 3 == 0) v = 0.0;   // exact zeros on the diagonal
                    A[(size_t)(i * P + r) * n + (i + dj) * P + c] = v;
                    ab[band_index<P>(i, r, dj, c)] = v;
                }
        }
    for (int i = 0; i < n; ++i) b[i] = urand();
    std::vector<double> x = b, xd;
    const int info = banded_lu_solve<BandShape<P>::kl, BandShape<P>::ku, double*>(ab.data(), n, x.data());
    dense_solve(A, b, n, xd);
    double res = 0.0, scale = 0.0, err = 0.0, xs = 0.0;
    for (int i = 0; i < n; ++i) {
        double s = -b[i], a = std::fabs(b[i]);
        for (int c = 0; c < n; ++c) { s += A[(size_t)i * n + c] * x[c]; a += std::fabs(A[(size_t)i * n + c] * x[c]); }
        res = std::fmax(res, std::fabs(s)); scale = std::fmax(scale, a);
        err = std::fmax(err, std::fabs(x[i] - xd[i])); xs = std::fmax(xs, std::fabs(xd[i]));
    }
    std::printf("%d %d %d %d %.3e %.3e\n", P, N, flavour, info, res / scale, err / xs);
}

int main() {
    srand(12345);
    for (int flavour = 0; flavour < 3; ++flavour) {
        for (int N : {3, 4, 17, 64, 300}) { run<1>(N, flavour); run<2>(N, flavour); run<3>(N, flavour); }
    }
    return 0;
}
