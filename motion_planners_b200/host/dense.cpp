#include "dense.hpp"

#include <cmath>
#include <numeric>
#include <utility>

namespace stomp_b200 {
namespace host {

bool invert_full_pivot(const Dense& a, Dense& inverse)
{
    const int n = a.rows();
    if (n != a.cols()) return false;
    Dense w = a;
    // row_of[i] / col_of[j]: which original row / column currently sits at position i / j
    std::vector<int> row_of(n), col_of(n);
    std::iota(row_of.begin(), row_of.end(), 0);
    std::iota(col_of.begin(), col_of.end(), 0);
    for (int step = 0; step < n; ++step) {
        int best_i = step, best_j = step;
        double best = 0.0;
        for (int i = step; i < n; ++i) {
            const double* row = w.data() + (size_t)i * n;
            for (int j = step; j < n; ++j) {
                const double m = std::fabs(row[j]);
                if (m > best) { best = m; best_i = i; best_j = j; }
            }
        }
        if (best == 0.0) return false;
        if (best_i != step) {
            for (int j = 0; j < n; ++j) std::swap(w.at(step, j), w.at(best_i, j));
            std::swap(row_of[step], row_of[best_i]);
        }
        if (best_j != step) {
            for (int i = 0; i < n; ++i) std::swap(w.at(i, step), w.at(i, best_j));
            std::swap(col_of[step], col_of[best_j]);
        }
        const double pivot = w.at(step, step);
        for (int i = step + 1; i < n; ++i) w.at(i, step) /= pivot;
        for (int i = step + 1; i < n; ++i) {
            const double f = w.at(i, step);
            if (f == 0.0) continue;
            double* ri = w.data() + (size_t)i * n;
            const double* rs = w.data() + (size_t)step * n;
            for (int j = step + 1; j < n; ++j) ri[j] -= f * rs[j];
        }
    }
    // P a Q = L U  =>  a^-1 = Q U^-1 L^-1 P ; solve column by column
    inverse = Dense(n, n);
    std::vector<double> y(n);
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) y[i] = (row_of[i] == c) ? 1.0 : 0.0;
        for (int i = 1; i < n; ++i) {
            double s = y[i];
            const double* ri = w.data() + (size_t)i * n;
            for (int j = 0; j < i; ++j) s -= ri[j] * y[j];
            y[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            const double* ri = w.data() + (size_t)i * n;
            for (int j = i + 1; j < n; ++j) s -= ri[j] * y[j];
            y[i] = s / ri[i];
        }
        for (int i = 0; i < n; ++i) inverse.at(col_of[i], c) = y[i];
    }
    return true;
}

bool cholesky_lower(const Dense& a, Dense& lower)
{
    const int n = a.rows();
    if (n != a.cols()) return false;
    lower = Dense(n, n);
    for (int j = 0; j < n; ++j) {
        double diag = a.at(j, j);
        for (int k = 0; k < j; ++k) diag -= lower.at(j, k) * lower.at(j, k);
        if (!(diag > 0.0)) return false;
        const double root = std::sqrt(diag);
        lower.at(j, j) = root;
        for (int i = j + 1; i < n; ++i) {
            double s = a.at(i, j);
            for (int k = 0; k < j; ++k) s -= lower.at(i, k) * lower.at(j, k);
            lower.at(i, j) = s / root;
        }
    }
    return true;
}

}  // namespace host
}  // namespace stomp_b200
