// stomp::CovariantMovementPrimitive — the policy: padded trajectory, quadratic control cost R, R^-1,
// the minimum-control-cost trajectory.  Same public surface as the reference for the members the
// wrapper touches (reference src/planners/stomp/include/stomp/CovariantMovementPrimitive.hpp, used at
// OptimizationTask.cpp:108-135 and StompPlanner.cpp:86-163); one-time host math, see
// motion_planners_b200/host/policy_core.hpp.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include <base/Eigen.hpp>
#include <stomp/StompUtils.hpp>

namespace stomp_b200 { namespace host { struct PolicyCore; } }

namespace stomp {

class CovariantMovementPrimitive {
public:
    CovariantMovementPrimitive();
    ~CovariantMovementPrimitive();
    CovariantMovementPrimitive(const CovariantMovementPrimitive& other);
    CovariantMovementPrimitive& operator=(const CovariantMovementPrimitive& other);

    // derivative_costs: [num_dimensions] (num_time_steps + 2*TRAJECTORY_PADDING) x NUM_DIFF_RULES.  This build
    // needs them to be the same for every joint and time step (as OptimizationTask sets them); returns false otherwise.
    bool initialize(const int num_time_steps, const int num_dimensions, const double movement_duration,
                    const std::vector<base::MatrixXd>& derivative_costs, const std::vector<base::VectorXd>& initial_trajectory);
    bool setToMinControlCost();
    bool updateMinControlCostParameters(const std::vector<base::VectorXd>& parameters_all);

    bool getParametersAll(std::vector<base::VectorXd>& parameters) const { parameters = parameters_all_; return true; }
    bool getParameters(std::vector<base::VectorXd>& parameters);
    bool setParameters(const std::vector<base::VectorXd>& parameters);
    bool setParametersAll(const std::vector<base::VectorXd>& parameters_all) { parameters_all_ = parameters_all; return true; }
    const std::vector<base::VectorXd>& getMinControlCostParameters() const { return min_control_cost_parameters_free_; }
    bool getNumTimeSteps(int& n) const { n = num_time_steps_; return true; }
    bool getNumDimensions(int& n) const { n = num_dimensions_; return true; }
    bool getControlCosts(std::vector<base::MatrixXd>& control_costs) const;
    bool getInvControlCosts(std::vector<base::MatrixXd>& inv_control_costs) const;
    bool writeToFile(const std::string abs_file_name);
    double getMovementDuration() const { return movement_duration_; }
    double getMovementDt() const { return movement_dt_; }
    base::MatrixXd getDifferentiationMatrix(int derivative_number) const;

    // flat views for the C ABI (stomp_b200_set_control_cost_matrices / stomp_b200_set_policy)
    const double* R() const;
    const double* Rinv() const;
    const double* L() const;
    void flatten(std::vector<double>& parameters_all, std::vector<double>& min_control_cost) const;
    const double* derivativeWeights() const { return derivative_weights_; }

    std::vector<base::VectorXd> parameters_all_;   // [num_dimensions] num_time_steps + 2*TRAJECTORY_PADDING (public in the reference too)

private:
    int num_time_steps_ = 0, num_dimensions_ = 0, num_vars_all_ = 0, free_vars_start_index_ = TRAJECTORY_PADDING;
    double movement_duration_ = 0.0, movement_dt_ = 0.0;
    double derivative_weights_[NUM_DIFF_RULES] = {0, 0, 0, 0};
    std::vector<base::VectorXd> min_control_cost_parameters_free_;
    std::shared_ptr<stomp_b200::host::PolicyCore> core_;
};

}  // namespace stomp
