// stomp::Stomp — the loop driver (reference src/planners/stomp/include/stomp/Stomp.hpp:47-131), here a thin
// host object over the C ABI: initialize() uploads the policy and resets the solve
// (stomp_b200_set_policy + stomp_b200_begin_solve), runSingleIteration() is one stomp_b200_iterate.
#pragma once
#include <memory>
#include <vector>

#include <boost/shared_ptr.hpp>
#include <base/Eigen.hpp>
#include <stomp/CovariantMovementPrimitive.hpp>
#include <stomp/StompConfig.hpp>
#include <stomp/StompTask.hpp>

struct stomp_b200_engine;
namespace motion_planners { class OptimizationTask; }

namespace stomp {

struct Rollout {   // the fields of reference PolicyImprovement.hpp:49-68 that callers read
    std::vector<base::VectorXd> parameters_noise_;
    base::VectorXd state_costs_;
    double total_cost_ = 0.0;
};

class Stomp {
public:
    Stomp();
    virtual ~Stomp();

    // task must already be initialized (policy created) at this point
    bool initialize(const StompConfig& config, std::shared_ptr<StompTask> task);
    bool runSingleIteration(int iteration_number);
    // num_iterations iterations queued on the device without host round trips; frozen at the wrapper's stop rule
    bool runIterations(int first_iteration, int num_iterations, bool honour_stop, int& iterations_used);
    // the whole solve loop on the device (stomp_b200_solve): up to max_iterations iterations, stop rule honoured on the
    // device, then the policy is synchronised back; pathFound() / getNoiselessRolloutTotalCost() / iterationsUsed() hold
    // the outcome
    bool solveOnDevice(int max_iterations, int& iterations_used);
    void getAllRollouts(std::vector<Rollout>& rollouts);
    double getNoiselessRolloutTotalCost() { return noiseless_total_cost_; }
    bool getLastNoiselessRolloutValid() const { return last_noiseless_rollout_valid_; }
    void getAdaptedStddevs(std::vector<double>& stddevs);
    // reference Stomp.cpp:356-359; false = per-time-step costs and probabilities (one GPU, no rollout reuse)
    void setCostCumulation(bool use_cumulative_costs);
    bool getParameters(std::vector<base::VectorXd>& parameters);   // current policy parameters (free part) from the device
    bool runUntilValid(int max_iterations, int iterations_after_collision_free);
    bool stopRuleFired() const { return stop_; }
    bool pathFound() const { return path_found_; }
    int iterationsUsed() const { return last_iterations_used_; }
    // copies the device parameters back into the task's policy (parameters_all_)
    bool syncPolicyFromDevice();

private:
    bool initialized_ = false;
    StompConfig stomp_config_;
    std::shared_ptr<StompTask> stomp_task_;
    std::shared_ptr<motion_planners::OptimizationTask> optimization_task_;
    boost::shared_ptr<CovariantMovementPrimitive> policy_;
    stomp_b200_engine* engine_ = nullptr;   // owned by the task (kept across solves: the SDF stays resident)
    double noiseless_total_cost_ = 0.0;
    bool last_noiseless_rollout_valid_ = false;
    bool stop_ = false;
    bool path_found_ = false;
    int last_iterations_used_ = 0;
};

}  // namespace stomp
