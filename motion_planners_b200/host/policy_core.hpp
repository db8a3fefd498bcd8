// Host-side, one-time products of the STOMP policy: what the reference's
// stomp::CovariantMovementPrimitive::initialize computes with Eigen
// (reference src/planners/stomp/src/CovariantMovementPrimitive.cpp:57-74,136-301) and
// stomp::getDifferentiationMatrix (src/planners/stomp/src/StompUtils.cpp:6-23).
// Used by the C ABI (stomp_b200_host_policy) and by the C++ stomp::CovariantMovementPrimitive mirror.
#pragma once
#include <vector>

#include "dense.hpp"

namespace stomp_b200 {
namespace host {

constexpr int kDiffRuleLength = 7;                  // StompUtils.hpp:56
constexpr int kPadding = kDiffRuleLength - 1;       // TRAJECTORY_PADDING, StompUtils.hpp:57
constexpr int kNumDiffRules = 4;                    // StompUtils.hpp:58
extern const double kDiffRules[kNumDiffRules][kDiffRuleLength];   // StompUtils.hpp:60-66

// Banded form of one differentiation matrix: row i holds the 7 entries of columns i-3 .. i+3; entries
// whose column falls outside [0, n) are zero and the coefficients the reference adds onto the clamped
// first / last column (StompUtils.cpp:14-19) are already folded into that column.
struct DiffBand {
    int n = 0;
    std::vector<double> c;   // [n][7]
    double entry(int i, int j) const { int o = j - i + 3; return (o < 0 || o > 6) ? 0.0 : c[(size_t)i * 7 + o]; }
};
DiffBand differentiation_band(int n, int order, double dt);

struct PolicyCore {
    int T = 0, D = 0, N = 0;
    double duration = 0, dt = 0;
    double weights[kNumDiffRules] = {0, 0, 0, 0};
    DiffBand diff[kNumDiffRules];
    Dense R_all;                       // control_costs_all_ [N][N]
    Dense R, Rinv, L;                  // control_costs_, inv_control_costs_, chol(Rinv)   [T][T]
    std::vector<double> params_all;    // parameters_all_ [D][N]
    std::vector<double> linear;        // linear_control_costs_ [D][T]
    std::vector<double> mincc;         // min_control_cost_parameters_free_ [D][T]

    // CovariantMovementPrimitive::initialize (:57-74) for derivative costs that are the same for every
    // joint and time step (what OptimizationTask::stompInitialize sets, OptimizationTask.cpp:22-44)
    bool initialize(int num_time_steps, int num_dimensions, double movement_duration,
                    const double derivative_weights[kNumDiffRules], const double* initial_all);
    void computeLinearControlCosts();                               // :136-172
    void setToMinControlCost();                                     // :128-132,174-189
    void updateMinControlCostParameters(const double* params_all_in);   // :191-200
};

// OptimizationTask::updateTrajectory (OptimizationTask.cpp:46-66)
void linear_initial_trajectory(int T, int D, const double* start, const double* goal, double* initial_all);

}  // namespace host
}  // namespace stomp_b200
