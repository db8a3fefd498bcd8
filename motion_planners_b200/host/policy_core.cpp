#include "policy_core.hpp"

#include <algorithm>
#include <cmath>

namespace stomp_b200 {
namespace host {

const double kDiffRules[kNumDiffRules][kDiffRuleLength] = {
    {0, 0, 0, 1, 0, 0, 0},                                                   // position
    {0, 0, -1, 1, 0, 0, 0},                                                  // velocity
    {0, -1 / 12.0, 16 / 12.0, -30 / 12.0, 16 / 12.0, -1 / 12.0, 0},           // acceleration
    {0, 1 / 12.0, -17 / 12.0, 46 / 12.0, -46 / 12.0, 17 / 12.0, -1 / 12.0}};  // jerk

DiffBand differentiation_band(int n, int order, double dt)
{
    DiffBand b;
    b.n = n;
    b.c.assign((size_t)n * 7, 0.0);
    const double scale = 1.0 / std::pow(dt, order);
    for (int i = 0; i < n; ++i)
        for (int tap = 0; tap < kDiffRuleLength; ++tap) {
            int col = std::min(std::max(i + tap - 3, 0), n - 1);   // index clamping at both ends
            b.c[(size_t)i * 7 + (col - i + 3)] += scale * kDiffRules[order][tap];
        }
    return b;
}

bool PolicyCore::initialize(int num_time_steps, int num_dimensions, double movement_duration,
                            const double derivative_weights[kNumDiffRules], const double* initial_all)
{
    T = num_time_steps;
    D = num_dimensions;
    N = T + 2 * kPadding;
    duration = movement_duration;
    dt = duration / (T + 1);
    for (int r = 0; r < kNumDiffRules; ++r) {
        weights[r] = derivative_weights[r];
        diff[r] = differentiation_band(N, r, dt);
    }
    params_all.assign(initial_all, initial_all + (size_t)D * N);

    // R_all = sum_r dt * D_r^T diag(w_r) D_r ; only rows k with |k-i|<=3 and |k-j|<=3 contribute
    R_all = Dense(N, N);
    for (int r = 0; r < kNumDiffRules; ++r) {
        if (weights[r] == 0.0) continue;
        for (int i = 0; i < N; ++i)
            for (int j = std::max(0, i - 6); j <= std::min(N - 1, i + 6); ++j) {
                double s = 0.0;
                for (int k = std::max(0, std::max(i, j) - 3); k <= std::min(N - 1, std::min(i, j) + 3); ++k)
                    s += diff[r].entry(k, i) * weights[r] * diff[r].entry(k, j);
                R_all.at(i, j) += dt * s;
            }
    }
    R = Dense(T, T);
    for (int i = 0; i < T; ++i)
        for (int j = 0; j < T; ++j) R.at(i, j) = R_all.at(kPadding + i, kPadding + j);
    if (!invert_full_pivot(R, Rinv)) return false;
    if (!cholesky_lower(Rinv, L)) return false;
    computeLinearControlCosts();
    mincc.assign((size_t)D * T, 0.0);
    return true;
}

void PolicyCore::computeLinearControlCosts()
{
    linear.assign((size_t)D * T, 0.0);
    for (int d = 0; d < D; ++d) {
        const double* x = params_all.data() + (size_t)d * N;
        double* lin = linear.data() + (size_t)d * T;
        for (int j = 0; j < T; ++j) {
            double head = 0.0, tail = 0.0;
            for (int i = 0; i < kPadding; ++i) head += x[i] * R_all.at(i, kPadding + j);
            for (int i = 0; i < kPadding; ++i) tail += x[kPadding + T + i] * R_all.at(kPadding + T + i, kPadding + j);
            lin[j] = (head + tail) * 2.0;
            lin[j] += -dt * 2.0 * (x[kPadding + j] * weights[0]);
        }
    }
}

void PolicyCore::setToMinControlCost()
{
    for (int d = 0; d < D; ++d) {
        const double* lin = linear.data() + (size_t)d * T;
        double* x = params_all.data() + (size_t)d * N;
        for (int i = 0; i < T; ++i) {
            double s = 0.0;
            for (int j = 0; j < T; ++j) s += Rinv.at(i, j) * lin[j];
            x[kPadding + i] = -0.5 * s;
        }
    }
    updateMinControlCostParameters(params_all.data());
}

void PolicyCore::updateMinControlCostParameters(const double* params_all_in)
{
    mincc.resize((size_t)D * T);
    for (int d = 0; d < D; ++d)
        std::copy(params_all_in + (size_t)d * N + kPadding, params_all_in + (size_t)d * N + kPadding + T,
                  mincc.begin() + (size_t)d * T);
}

void linear_initial_trajectory(int T, int D, const double* start, const double* goal, double* initial_all)
{
    const int N = T + 2 * kPadding;
    for (int d = 0; d < D; ++d) {
        double* x = initial_all + (size_t)d * N;
        for (int i = 0; i < kPadding; ++i) {
            x[i] = start[d];
            x[kPadding + T + i] = goal[d];
        }
        const double increment = (goal[d] - start[d]) / (T - 1);
        for (int i = 0; i < T; ++i) x[kPadding + i] = start[d] + (i * increment);
    }
}

}  // namespace host
}  // namespace stomp_b200
