// stomp::Stomp over the C ABI (include/stomp_b200.h).  Mirrors the call structure of the reference's
// Stomp.cpp:56-94 (initialize) and :274-301 (runSingleIteration); the work happens on the GPU.
#include <stomp/Stomp.hpp>

#include <cstring>

#include <base-logging/Logging.hpp>
#include <wrapper/stomp/OptimizationTask.hpp>

#include "../../include/stomp_b200.h"

namespace stomp {

Stomp::Stomp() {}
Stomp::~Stomp() {}

bool Stomp::initialize(const StompConfig& config, std::shared_ptr<StompTask> task)
{
    stomp_config_ = config;
    stomp_task_ = task;
    optimization_task_ = std::dynamic_pointer_cast<motion_planners::OptimizationTask>(task);
    if (!optimization_task_) {
        LOG_ERROR_S << "[Stomp]: the CUDA loop runs motion_planners::OptimizationTask; other StompTask subclasses are not supported";
        return false;
    }
    if (!stomp_task_->getPolicy(policy_) || !policy_) {
        LOG_ERROR_S << "[Stomp]: the task has no policy (call setStartGoalTrajectory first)";
        return false;
    }
    engine_ = optimization_task_->engine();
    if (!engine_) return false;
    // upload the host-computed policy products (CovariantMovementPrimitive::initialize)
    std::vector<double> params_all, mincc;
    policy_->flatten(params_all, mincc);
    int rc = stomp_b200_set_control_cost_matrices(engine_, policy_->R(), policy_->Rinv(), policy_->L());
    if (!rc) rc = stomp_b200_set_policy(engine_, 0, params_all.data(), mincc.data());
    if (!rc) rc = stomp_b200_begin_solve(engine_);
    if (rc) {
        LOG_ERROR_S << "[Stomp]: " << stomp_b200_status_string(rc) << ": " << stomp_b200_last_error(engine_);
        return false;
    }
    noiseless_total_cost_ = 0.0;
    last_noiseless_rollout_valid_ = false;
    stop_ = false;
    return (initialized_ = true);
}

bool Stomp::runSingleIteration(int iteration_number)
{
    if (!initialized_) return false;
    double cost = 0.0;
    uint8_t valid = 0;
    int32_t stop = 0;
    const int rc = stomp_b200_iterate(engine_, iteration_number, nullptr, nullptr, &cost, &valid, &stop);
    if (rc) {
        LOG_ERROR_S << "[Stomp]: " << stomp_b200_status_string(rc) << ": " << stomp_b200_last_error(engine_);
        return false;
    }
    noiseless_total_cost_ = cost;
    last_noiseless_rollout_valid_ = valid != 0;
    stop_ = stop != 0;
    return true;
}

bool Stomp::runIterations(int first_iteration, int num_iterations, bool honour_stop, int& iterations_used)
{
    if (!initialized_) return false;
    int rc = stomp_b200_run(engine_, first_iteration, num_iterations, honour_stop ? 1 : 0);
    if (rc) {
        LOG_ERROR_S << "[Stomp]: " << stomp_b200_status_string(rc) << ": " << stomp_b200_last_error(engine_);
        return false;
    }
    if (!syncPolicyFromDevice()) return false;
    iterations_used = last_iterations_used_;
    return true;
}

// The loop of StompPlanner::solve (reference StompPlanner.cpp:96-141) queued on the device: no host round trip per
// iteration; the stop rule (:117) is evaluated on the device after every noise-less rollout and freezes the query.
bool Stomp::solveOnDevice(int max_iterations, int& iterations_used)
{
    if (!initialized_) return false;
    int32_t queued = 0;
    const int rc = stomp_b200_solve(engine_, max_iterations, 0, &queued);
    if (rc) {
        LOG_ERROR_S << "[Stomp]: " << stomp_b200_status_string(rc) << ": " << stomp_b200_last_error(engine_);
        return false;
    }
    if (!syncPolicyFromDevice()) return false;
    iterations_used = last_iterations_used_;
    return true;
}

bool Stomp::syncPolicyFromDevice()
{
    if (!initialized_) return false;
    int T = 0, D = 0;
    policy_->getNumTimeSteps(T);
    policy_->getNumDimensions(D);
    std::vector<double> solution((size_t)D * T);
    int32_t status = 0, iters = 0;
    double cost = 0.0;
    const int rc = stomp_b200_finish_solve(engine_, solution.data(), &status, &iters, &cost);
    if (rc) {
        LOG_ERROR_S << "[Stomp]: " << stomp_b200_status_string(rc) << ": " << stomp_b200_last_error(engine_);
        return false;
    }
    for (int d = 0; d < D; ++d)
        for (int t = 0; t < T; ++t) policy_->parameters_all_[d](TRAJECTORY_PADDING + t) = solution[(size_t)d * T + t];
    noiseless_total_cost_ = cost;
    last_iterations_used_ = iters;
    path_found_ = status != 0;
    initialized_ = false;   // a new solve needs initialize() again, like `new stomp::Stomp` per solve in the reference
    return true;
}

void Stomp::getAllRollouts(std::vector<Rollout>& rollouts)
{
    rollouts.clear();
    if (!engine_) return;
    int32_t n = 0, g = 0;
    stomp_b200_num_rollouts(engine_, &n, &g);
    int T = 0, D = 0;
    policy_->getNumTimeSteps(T);
    policy_->getNumDimensions(D);
    std::vector<double> noisy((size_t)n * D * T), state((size_t)n * T), total(n);
    if (stomp_b200_get_tensor(engine_, STOMP_B200_ROLLOUTS, noisy.data(), noisy.size() * sizeof(double))) return;
    if (stomp_b200_get_tensor(engine_, STOMP_B200_STATE_COSTS, state.data(), state.size() * sizeof(double))) return;
    if (stomp_b200_get_tensor(engine_, STOMP_B200_TOTAL_COST, total.data(), total.size() * sizeof(double))) return;
    rollouts.resize(n);
    for (int r = 0; r < n; ++r) {
        rollouts[r].parameters_noise_.assign(D, base::VectorXd::Zero(T));
        rollouts[r].state_costs_ = base::VectorXd::Zero(T);
        for (int d = 0; d < D; ++d)
            for (int t = 0; t < T; ++t) rollouts[r].parameters_noise_[d](t) = noisy[((size_t)r * D + d) * T + t];
        for (int t = 0; t < T; ++t) rollouts[r].state_costs_(t) = state[(size_t)r * T + t];
        rollouts[r].total_cost_ = total[r];
    }
}

void Stomp::getAdaptedStddevs(std::vector<double>& stddevs)
{
    int D = 0;
    policy_->getNumDimensions(D);
    stddevs.assign(D, 0.0);
    if (engine_) stomp_b200_get_tensor(engine_, STOMP_B200_STDDEVS, stddevs.data(), sizeof(double) * D);
}

bool Stomp::getParameters(std::vector<base::VectorXd>& parameters)
{
    if (!engine_ || !policy_) return false;
    int T = 0, D = 0;
    policy_->getNumTimeSteps(T);
    policy_->getNumDimensions(D);
    std::vector<double> flat((size_t)D * T);
    if (stomp_b200_get_tensor(engine_, STOMP_B200_PARAMETERS, flat.data(), flat.size() * sizeof(double))) return false;
    parameters.assign(D, base::VectorXd::Zero(T));
    for (int d = 0; d < D; ++d)
        for (int t = 0; t < T; ++t) parameters[d](t) = flat[(size_t)d * T + t];
    return true;
}

// reference Stomp.cpp:325-351
bool Stomp::runUntilValid(int max_iterations, int iterations_after_collision_free)
{
    int collision_free_iterations = 0;
    bool success = false;
    for (int i = 0; i < max_iterations; ++i) {
        if (!runSingleIteration(i)) return false;
        stomp_task_->onEveryIteration();
        if (last_noiseless_rollout_valid_) {
            success = true;
            collision_free_iterations++;
        }
        if (collision_free_iterations >= iterations_after_collision_free) break;
    }
    return success;
}

void Stomp::setCostCumulation(bool use_cumulative_costs)
{
    if (!engine_) return;
    const int rc = stomp_b200_set_cost_cumulation(engine_, use_cumulative_costs ? 1 : 0);
    if (rc) LOG_ERROR_S << "[Stomp]: setCostCumulation: " << stomp_b200_status_string(rc) << ": " << stomp_b200_last_error(engine_);
}

}  // namespace stomp
