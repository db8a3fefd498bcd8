// stomp::CovariantMovementPrimitive over the host policy math (policy_core.hpp).
#include <stomp/CovariantMovementPrimitive.hpp>

#include <cstdio>

#include "policy_core.hpp"

namespace stomp {

using stomp_b200::host::PolicyCore;

CovariantMovementPrimitive::CovariantMovementPrimitive() {}
CovariantMovementPrimitive::~CovariantMovementPrimitive() {}

CovariantMovementPrimitive::CovariantMovementPrimitive(const CovariantMovementPrimitive& o) { *this = o; }

CovariantMovementPrimitive& CovariantMovementPrimitive::operator=(const CovariantMovementPrimitive& o)
{
    if (this == &o) return *this;
    parameters_all_ = o.parameters_all_;
    num_time_steps_ = o.num_time_steps_; num_dimensions_ = o.num_dimensions_; num_vars_all_ = o.num_vars_all_;
    free_vars_start_index_ = o.free_vars_start_index_;
    movement_duration_ = o.movement_duration_; movement_dt_ = o.movement_dt_;
    for (int i = 0; i < NUM_DIFF_RULES; ++i) derivative_weights_[i] = o.derivative_weights_[i];
    min_control_cost_parameters_free_ = o.min_control_cost_parameters_free_;
    core_ = o.core_ ? std::make_shared<PolicyCore>(*o.core_) : nullptr;
    return *this;
}

bool CovariantMovementPrimitive::initialize(const int num_time_steps, const int num_dimensions, const double movement_duration,
                                            const std::vector<base::MatrixXd>& derivative_costs,
                                            const std::vector<base::VectorXd>& initial_trajectory)
{
    num_time_steps_ = num_time_steps;
    num_dimensions_ = num_dimensions;
    movement_duration_ = movement_duration;
    num_vars_all_ = num_time_steps + 2 * TRAJECTORY_PADDING;
    if ((int)derivative_costs.size() != num_dimensions || (int)initial_trajectory.size() != num_dimensions) return false;
    // the device path shares one R / L between all joints and time steps
    for (int r = 0; r < NUM_DIFF_RULES; ++r) derivative_weights_[r] = derivative_costs[0](0, r);
    for (int d = 0; d < num_dimensions; ++d) {
        if (derivative_costs[d].rows() != num_vars_all_ || derivative_costs[d].cols() != NUM_DIFF_RULES) return false;
        if (initial_trajectory[d].size() != num_vars_all_) return false;
        for (int i = 0; i < num_vars_all_; ++i)
            for (int r = 0; r < NUM_DIFF_RULES; ++r)
                if (derivative_costs[d](i, r) != derivative_weights_[r]) return false;
    }
    std::vector<double> flat((size_t)num_dimensions * num_vars_all_);
    for (int d = 0; d < num_dimensions; ++d)
        for (int i = 0; i < num_vars_all_; ++i) flat[(size_t)d * num_vars_all_ + i] = initial_trajectory[d](i);
    core_ = std::make_shared<PolicyCore>();
    if (!core_->initialize(num_time_steps, num_dimensions, movement_duration, derivative_weights_, flat.data())) return false;
    movement_dt_ = core_->dt;
    parameters_all_ = initial_trajectory;
    return true;
}

bool CovariantMovementPrimitive::setToMinControlCost()
{
    if (!core_) return false;
    // parameters_all_ may have been edited through the public member: hand the padding back to the core
    for (int d = 0; d < num_dimensions_; ++d)
        for (int i = 0; i < num_vars_all_; ++i) core_->params_all[(size_t)d * num_vars_all_ + i] = parameters_all_[d](i);
    core_->computeLinearControlCosts();
    core_->setToMinControlCost();
    min_control_cost_parameters_free_.assign(num_dimensions_, base::VectorXd::Zero(num_time_steps_));
    for (int d = 0; d < num_dimensions_; ++d) {
        for (int i = 0; i < num_vars_all_; ++i) parameters_all_[d](i) = core_->params_all[(size_t)d * num_vars_all_ + i];
        for (int t = 0; t < num_time_steps_; ++t) min_control_cost_parameters_free_[d](t) = core_->mincc[(size_t)d * num_time_steps_ + t];
    }
    return true;
}

bool CovariantMovementPrimitive::updateMinControlCostParameters(const std::vector<base::VectorXd>& parameters_all)
{
    min_control_cost_parameters_free_.assign(num_dimensions_, base::VectorXd::Zero(num_time_steps_));
    for (int d = 0; d < num_dimensions_; ++d)
        for (int t = 0; t < num_time_steps_; ++t) min_control_cost_parameters_free_[d](t) = parameters_all[d](free_vars_start_index_ + t);
    return true;
}

bool CovariantMovementPrimitive::getParameters(std::vector<base::VectorXd>& parameters)
{
    if ((int)parameters.size() != num_dimensions_) parameters.assign(num_dimensions_, base::VectorXd::Zero(num_time_steps_));
    for (int d = 0; d < num_dimensions_; ++d) {
        parameters[d].resize(num_time_steps_);
        for (int t = 0; t < num_time_steps_; ++t) parameters[d](t) = parameters_all_[d](free_vars_start_index_ + t);
    }
    return true;
}

bool CovariantMovementPrimitive::setParameters(const std::vector<base::VectorXd>& parameters)
{
    if ((int)parameters.size() != num_dimensions_) return false;
    for (int d = 0; d < num_dimensions_; ++d)
        for (int t = 0; t < num_time_steps_; ++t) parameters_all_[d](free_vars_start_index_ + t) = parameters[d](t);
    return true;
}

static base::MatrixXd to_matrix(const stomp_b200::host::Dense& m)
{
    base::MatrixXd out(m.rows(), m.cols());
    for (int i = 0; i < m.rows(); ++i)
        for (int j = 0; j < m.cols(); ++j) out(i, j) = m.at(i, j);
    return out;
}

bool CovariantMovementPrimitive::getControlCosts(std::vector<base::MatrixXd>& control_costs) const
{
    if (!core_) return false;
    control_costs.assign(num_dimensions_, to_matrix(core_->R));
    return true;
}

bool CovariantMovementPrimitive::getInvControlCosts(std::vector<base::MatrixXd>& inv_control_costs) const
{
    if (!core_) return false;
    inv_control_costs.assign(num_dimensions_, to_matrix(core_->Rinv));
    return true;
}

// rows free_start-1 .. free_end+1, one tab-separated "%f" per joint (reference CovariantMovementPrimitive.cpp:528-546)
bool CovariantMovementPrimitive::writeToFile(const std::string abs_file_name)
{
    FILE* f = fopen(abs_file_name.c_str(), "w");
    if (!f) return false;
    for (int i = free_vars_start_index_ - 1; i <= free_vars_start_index_ + num_time_steps_; ++i) {
        for (int d = 0; d < num_dimensions_; ++d) fprintf(f, "%f\t", parameters_all_[d](i));
        fprintf(f, "\n");
    }
    fclose(f);
    return true;
}

base::MatrixXd CovariantMovementPrimitive::getDifferentiationMatrix(int derivative_number) const
{
    base::MatrixXd m;
    stomp::getDifferentiationMatrix(num_vars_all_, (CostComponents)derivative_number, movement_dt_, m);
    return m;
}

const double* CovariantMovementPrimitive::R() const { return core_ ? core_->R.data() : nullptr; }
const double* CovariantMovementPrimitive::Rinv() const { return core_ ? core_->Rinv.data() : nullptr; }
const double* CovariantMovementPrimitive::L() const { return core_ ? core_->L.data() : nullptr; }

void CovariantMovementPrimitive::flatten(std::vector<double>& parameters_all, std::vector<double>& min_control_cost) const
{
    parameters_all.resize((size_t)num_dimensions_ * num_vars_all_);
    min_control_cost.resize((size_t)num_dimensions_ * num_time_steps_);
    for (int d = 0; d < num_dimensions_; ++d) {
        for (int i = 0; i < num_vars_all_; ++i) parameters_all[(size_t)d * num_vars_all_ + i] = parameters_all_[d](i);
        for (int t = 0; t < num_time_steps_; ++t)
            min_control_cost[(size_t)d * num_time_steps_ + t] = min_control_cost_parameters_free_[d](t);
    }
}

}  // namespace stomp
