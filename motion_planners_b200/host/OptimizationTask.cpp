// motion_planners::OptimizationTask (reference src/planners/src/wrappers/stomp/OptimizationTask.cpp).
#include <wrapper/stomp/OptimizationTask.hpp>

#include <algorithm>

#include "../../include/stomp_b200.h"

namespace motion_planners {

OptimizationTask::OptimizationTask(stomp::StompConfig config, std::shared_ptr<robot_model::RobotModel>& robot_model)
    : stomp_config_(config), robot_model_(robot_model)
{
    planning_group_name_ = robot_model_->getPlanningGroupName();
    robot_model_->getPlanningGroupJointsName(planning_group_name_, planning_group_joints_names_);
    if (!robot_model->getJointLimits(lower_limits_, upper_limits_)) LOG_FATAL_S << "[OptimizationTask]: Cannot get joint limits";
}

OptimizationTask::~OptimizationTask()
{
    policy_.reset();
    if (engine_) stomp_b200_destroy(engine_);
}

// reference :22-44: only the acceleration term carries weight
bool OptimizationTask::stompInitialize(int, int)
{
    const int N = stomp_config_.num_time_steps_ + 2 * stomp::TRAJECTORY_PADDING;
    derivative_costs_.assign(stomp_config_.num_dimensions_, base::MatrixXd::Zero(N, stomp::NUM_DIFF_RULES));
    initial_trajectory_.assign(stomp_config_.num_dimensions_, base::VectorXd::Zero(N));
    for (int d = 0; d < stomp_config_.num_dimensions_; ++d)
        for (int i = 0; i < N; ++i) derivative_costs_[d](i, stomp::STOMP_ACCELERATION) = 1.0;
    return true;
}

// reference :46-66: linear interpolation between start and goal, both repeated over the padding
void OptimizationTask::updateTrajectory(const base::samples::Joints& start, const base::samples::Joints& goal)
{
    const int T = stomp_config_.num_time_steps_, P = stomp::TRAJECTORY_PADDING;
    for (int d = 0; d < stomp_config_.num_dimensions_; ++d) {
        const double s = start.elements.at(d).position, g = goal.elements.at(d).position;
        for (int i = 0; i < P; ++i) {
            initial_trajectory_[d](i) = s;
            initial_trajectory_[d](P + T + i) = g;
        }
        const double increment = (g - s) / (T - 1);
        for (int i = 0; i < T; i++) initial_trajectory_[d](P + i) = s + (i * increment);
    }
}

bool OptimizationTask::getPolicy(boost::shared_ptr<stomp::CovariantMovementPrimitive>& policy)
{
    policy = policy_;
    return true;
}

bool OptimizationTask::setPolicy(const boost::shared_ptr<stomp::CovariantMovementPrimitive> policy)
{
    policy_ = policy;
    return true;
}

double OptimizationTask::getControlCostWeight() { return stomp_config_.control_cost_weight_; }

// reference :85-106 (the CUDA loop applies the same clamp to every generated rollout)
bool OptimizationTask::filter(std::vector<base::VectorXd>& parameters, int, int)
{
    bool filtered = false;
    for (unsigned int d = 0; d < parameters.size(); ++d)
        for (int t = 0; t < stomp_config_.num_time_steps_; ++t) {
            if (parameters[d](t) < lower_limits_.at(d)) { parameters[d](t) = lower_limits_.at(d); filtered = true; }
            if (parameters[d](t) > upper_limits_.at(d)) { parameters[d](t) = upper_limits_.at(d); filtered = true; }
        }
    return filtered;
}

// reference :108-119
void OptimizationTask::createPolicy()
{
    policy_.reset(new stomp::CovariantMovementPrimitive());
    policy_->initialize(stomp_config_.num_time_steps_, stomp_config_.num_dimensions_, stomp_config_.movement_duration_,
                        derivative_costs_, initial_trajectory_);
    policy_->setToMinControlCost();
    policy_->getParametersAll(initial_trajectory_);
    movement_dt_ = policy_->getMovementDt();
}

// reference :121-135 (warm start: the given trajectory is also the minimum-control-cost reference)
void OptimizationTask::updatePolicy()
{
    policy_.reset(new stomp::CovariantMovementPrimitive());
    policy_->initialize(stomp_config_.num_time_steps_, stomp_config_.num_dimensions_, stomp_config_.movement_duration_,
                        derivative_costs_, initial_trajectory_);
    policy_->updateMinControlCostParameters(initial_trajectory_);
    movement_dt_ = policy_->getMovementDt();
}

// alternative state costs -> the engine (stomp_b200_set_cost_extras); the shipped configuration sends all-off
bool OptimizationTask::applyCostSwitches()
{
    costs_dirty_ = false;
    const int D = stomp_config_.num_dimensions_;
    const bool jc = use_joints_constraint_cost_ && constraints_.use_constraint == motion_planners::JOINTS_CONSTRAINT &&
                    (int)constraints_.joint_constraint.value.size() == D && (int)constraints_.joint_constraint.tolerance.size() == D;
    std::vector<double> value(D, 0.0), tolerance(D, 0.0);
    for (int d = 0; d < D && jc; ++d) { value[d] = constraints_.joint_constraint.value(d); tolerance[d] = constraints_.joint_constraint.tolerance(d); }
    const int rc = stomp_b200_set_cost_extras(engine_, use_smooth_cost_ ? 1 : 0, smooth_margin_, smooth_weight_, jc ? 1 : 0,
                                              value.data(), tolerance.data(), joints_constraint_weight_);
    if (rc) LOG_ERROR_S << "[OptimizationTask]: " << stomp_b200_status_string(rc) << ": " << stomp_b200_last_error(engine_);
    return rc == 0;
}

stomp_b200_engine* OptimizationTask::engine()
{
    if (engine_) {
        if (costs_dirty_ && !applyCostSwitches()) return nullptr;
        // a world object added / removed / moved since the engine was configured (reference: handleCollisionObjectInWorld,
        // updateOctomap take effect at once): rebuild the distance field on the device before the next use
        if (robot_revision_ != robot_model_->robotRevision()) {       // a grasp object was attached / removed: new sphere list (+ scene)
            const int rc = robot_model_->configureEngine(engine_);
            if (rc) {
                LOG_ERROR_S << "[OptimizationTask]: " << stomp_b200_status_string(rc) << ": " << stomp_b200_last_error(engine_);
                return nullptr;
            }
            robot_revision_ = robot_model_->robotRevision();
            scene_revision_ = robot_model_->sceneRevision();
        }
        if (scene_revision_ != robot_model_->sceneRevision()) {
            const int rc = robot_model_->configureScene(engine_);
            if (rc) {
                LOG_ERROR_S << "[OptimizationTask]: " << stomp_b200_status_string(rc) << ": " << stomp_b200_last_error(engine_);
                return nullptr;
            }
            scene_revision_ = robot_model_->sceneRevision();
        }
        return engine_;
    }
    stomp_b200_config cfg;
    stomp_b200_default_config(&cfg);
    cfg.num_time_steps = stomp_config_.num_time_steps_;
    cfg.num_dimensions = stomp_config_.num_dimensions_;
    cfg.min_rollouts = stomp_config_.min_rollouts_;
    cfg.max_rollouts = stomp_config_.max_rollouts_;
    cfg.num_rollouts_per_iteration = stomp_config_.num_rollouts_per_iteration_;
    cfg.movement_duration = stomp_config_.movement_duration_;
    cfg.control_cost_weight = stomp_config_.control_cost_weight_;
    cfg.min_cost_improvement = stomp_config_.min_cost_improvement_;
    for (int d = 0; d < stomp_config_.num_dimensions_ && d < STOMP_B200_MAX_DIMS; ++d) {
        cfg.noise_stddev[d] = stomp_config_.noise_stddev_.at(d);
        cfg.noise_decay[d] = stomp_config_.noise_decay_.at(d);
        cfg.noise_min_stddev[d] = stomp_config_.noise_min_stddev_.at(d);
    }
    cfg.use_noise_adaptation = stomp_config_.use_noise_adaptation_ ? 1 : 0;
    cfg.device = stomp_config_.device_;
    cfg.seed = stomp_config_.seed_;
    int rc = stomp_b200_create(&cfg, &engine_);
    if (rc) {
        LOG_ERROR_S << "[OptimizationTask]: stomp_b200_create: " << stomp_b200_status_string(rc);
        engine_ = nullptr;
        return nullptr;
    }
    rc = robot_model_->configureEngine(engine_);
    if (rc) {
        LOG_ERROR_S << "[OptimizationTask]: " << stomp_b200_status_string(rc) << ": " << stomp_b200_last_error(engine_);
        stomp_b200_destroy(engine_);
        engine_ = nullptr;
        return nullptr;
    }
    scene_revision_ = robot_model_->sceneRevision();
    robot_revision_ = robot_model_->robotRevision();
    if ((costs_dirty_ || use_smooth_cost_ || use_joints_constraint_cost_) && !applyCostSwitches()) return nullptr;
    return engine_;
}

// reference :137-204 for one trajectory: cost 1.0 / 0.0 per timestep, validity = last timestep
bool OptimizationTask::execute(std::vector<base::VectorXd>& parameters, std::vector<base::VectorXd>&, base::VectorXd& costs,
                               base::MatrixXd&, const int, const int, int, bool, std::vector<base::VectorXd>&, bool& validity)
{
    const int T = stomp_config_.num_time_steps_, D = stomp_config_.num_dimensions_;
    costs = base::VectorXd::Zero(T);
    validity = true;
    stomp_b200_engine* e = engine();
    if (!e) return false;
    std::vector<double> theta((size_t)D * T);
    for (int d = 0; d < D; ++d)
        for (int t = 0; t < T; ++t) theta[(size_t)d * T + t] = parameters[d](t);
    uint8_t valid = 1;
    const int rc = stomp_b200_evaluate_states(e, theta.data(), 1, T, costs.data(), nullptr, &valid);
    if (rc) {
        LOG_ERROR_S << "[OptimizationTask]: " << stomp_b200_last_error(e);
        return false;
    }
    validity = valid != 0;
    return true;
}

}  // namespace motion_planners
