// hpfw_b200/csrc/eigh_host.h — small dense symmetric eigen-solver on the host (double): Householder tridiagonalisation +
// implicit-shift QL, the classic EISPACK tred2 / tql2 pair restated. Used by hpfw_calc_filters (learn.cu) for the 128 x 128
// Rayleigh-Ritz problems of the block subspace iteration: O(p^3) with small constants (~5 ms at p = 128) where the cyclic
// Jacobi it replaces took 100-200 ms per call (strided rotations, ~10 sweeps to its 1e-30 off-diagonal target).
#pragma once

#include <cmath>
#include <vector>

namespace hpfw_b200 {

// H: symmetric p x p, row-major (destroyed). On return w[j] = eigenvalue j and column j of Q (Q[k * p + j]) its unit
// eigenvector. Returns false if the QL iteration did not converge (the caller falls back to Jacobi).
inline bool sym_eigh_ql(std::vector<double> &H, int p, std::vector<double> &w, std::vector<double> &Q) {
    const int n = p;
    std::vector<double> &a = H;
    std::vector<double> d(n, 0.0), e(n, 0.0);
    auto A = [&](int i, int j) -> double & { return a[(size_t)i * n + j]; };
    // ---- Householder reduction to tridiagonal form; `a` becomes the accumulated orthogonal transformation
    for (int i = n - 1; i > 0; --i) {
        const int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; ++k) scale += std::fabs(A(i, k));
            if (scale == 0.0) {
                e[i] = A(i, l);
            } else {
                for (int k = 0; k <= l; ++k) {
                    A(i, k) /= scale;
                    h += A(i, k) * A(i, k);
                }
                double f = A(i, l);
                double g = f >= 0.0 ? -std::sqrt(h) : std::sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                A(i, l) = f - g;
                f = 0.0;
                for (int j = 0; j <= l; ++j) {
                    A(j, i) = A(i, j) / h;
                    g = 0.0;
                    for (int k = 0; k <= j; ++k) g += A(j, k) * A(i, k);
                    for (int k = j + 1; k <= l; ++k) g += A(k, j) * A(i, k);
                    e[j] = g / h;
                    f += e[j] * A(i, j);
                }
                const double hh = f / (h + h);
                for (int j = 0; j <= l; ++j) {
                    f = A(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; ++k) A(j, k) -= f * e[k] + g * A(i, k);
                }
            }
        } else {
            e[i] = A(i, l);
        }
        d[i] = h;
    }
    d[0] = 0.0;
    e[0] = 0.0;
    for (int i = 0; i < n; ++i) {
        const int l = i - 1;
        if (d[i] != 0.0) {
            for (int j = 0; j <= l; ++j) {
                double g = 0.0;
                for (int k = 0; k <= l; ++k) g += A(i, k) * A(k, j);
                for (int k = 0; k <= l; ++k) A(k, j) -= g * A(k, i);
            }
        }
        d[i] = A(i, i);
        A(i, i) = 1.0;
        for (int j = 0; j <= l; ++j) A(j, i) = A(i, j) = 0.0;
    }
    // ---- implicit QL on the tridiagonal matrix, rotations accumulated into `a`
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 80) return false;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, pp = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    e[i + 1] = (r = std::hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= pp;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - pp;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (pp = s * r);
                    g = c * r - b;
                    for (int k = 0; k < n; ++k) {
                        f = A(k, i + 1);
                        A(k, i + 1) = s * A(k, i) + c * f;
                        A(k, i) = c * A(k, i) - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= pp;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    w = d;
    Q = a;
    return true;
}

}  // namespace hpfw_b200
