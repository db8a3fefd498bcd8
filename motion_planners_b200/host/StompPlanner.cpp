// motion_planners::StompPlanner (reference src/planners/src/wrappers/stomp/StompPlanner.cpp): same control
// flow, the iterations run on the GPU through stomp::Stomp -> include/stomp_b200.h.
#include <wrapper/stomp/StompPlanner.hpp>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <sstream>

namespace motion_planners {

StompPlanner::StompPlanner() : num_iterations_(0) {}
StompPlanner::~StompPlanner() {}

bool StompPlanner::initializePlanner(std::shared_ptr<robot_model::RobotModel>& robot_model, std::string config_file_path)
{
    YAML::Node input_config;
    motion_planners::loadConfigFile(config_file_path, input_config);
    const YAML::Node stomp_node = input_config["stomp"];
    const YAML::Node debug_node = input_config["debug"];
    if (!stomp_node) {
        LOG_ERROR_S << "[StompPlanner]: no 'stomp' node in " << config_file_path;
        return false;
    }
    return initializePlanner(robot_model, handle_stomp_config::getStompConfig(stomp_node),
                             debug_node ? handle_stomp_config::getDebugConfig(debug_node) : stomp::DebugConfig());
}

bool StompPlanner::initializePlanner(std::shared_ptr<robot_model::RobotModel>& robot_model, const stomp::StompConfig& config,
                                     const stomp::DebugConfig& debug)
{
    stomp_config_ = config;
    debug_config_ = debug;
    if (!assignPlanningJointInformation(robot_model)) return false;
    if ((int)planning_group_joints_name_.size() != stomp_config_.num_dimensions_) {
        // the reference asserts here (StompPlanner.cpp:31)
        LOG_ERROR_S << "[StompPlanner]: planning group has " << planning_group_joints_name_.size() << " joints, num_dimensions_ is "
                    << stomp_config_.num_dimensions_;
        return false;
    }
    optimization_task_.reset(new OptimizationTask(stomp_config_, robot_model_));
    optimization_task_->stompInitialize(1, 1);
    num_iterations_ = 0;
    return true;
}

bool StompPlanner::reInitializePlanner()
{
    if (!optimization_task_ || !robot_model_) {
        LOG_DEBUG_S << "[reInitializePlanner] The stomp planner and robot model were not initialised before. This function should be called only if the planner was initialised before";
        return false;
    }
    optimization_task_.reset(new OptimizationTask(stomp_config_, robot_model_));
    optimization_task_->stompInitialize(1, 1);
    num_iterations_ = 0;
    return true;
}

bool StompPlanner::reInitializeTimeSteps(const int& num_time_steps)
{
    stomp_config_.num_time_steps_ = num_time_steps;
    return reInitializePlanner();
}

bool StompPlanner::solve(base::JointsTrajectory& solution, PlannerStatus& planner_status)
{
    optimization_task_->setOptimizationConstraints(constraints_);
    stomp_.reset(new stomp::Stomp());
    if (!stomp_->initialize(stomp_config_, optimization_task_)) {
        planner_status.statuscode = motion_planners::PlannerStatus::PLANNER_INITIALISATION_FAILED;
        stomp_.reset();
        return false;
    }

    if ((debug_config_.save_noiseless_trajectories_) || (debug_config_.save_noisy_trajectories_)) mkdir(debug_config_.output_dir_.c_str(), 0755);
    FILE* num_rollouts_file = NULL;
    if (debug_config_.save_noisy_trajectories_) {
        std::stringstream name;
        name << debug_config_.output_dir_ << "/num_rollouts.txt";
        num_rollouts_file = fopen(name.str().c_str(), "w");
    }
    if (debug_config_.save_noiseless_trajectories_) {
        std::stringstream sss;
        sss << debug_config_.output_dir_ << "/noiseless_0.txt";
        optimization_task_->policy_->writeToFile(sss.str());
    }
    tmp_policy = *optimization_task_->policy_;

    double old_cost = 0.0;
    double cost_improvement = 0.0;
    double current_trajectory_totalcost = 0.0;
    num_iterations_ = 0;

    for (int i = 0; i < stomp_config_.num_iterations_; i++) {
        num_iterations_++;
        if (!stomp_->runSingleIteration(i)) {
            planner_status.statuscode = motion_planners::PlannerStatus::CRASH;
            stomp_.reset();
            return false;
        }
        current_trajectory_totalcost = stomp_->getNoiselessRolloutTotalCost();
        cost_improvement = current_trajectory_totalcost - old_cost;
        old_cost = current_trajectory_totalcost;
        LOG_DEBUG_S << "Iteration = " << i << ". Total Cost = " << current_trajectory_totalcost << " . Cost improvement = " << cost_improvement;

        // Stop criterion: a total cost below 1 means no timestep is in collision (each costs 1)
        if ((current_trajectory_totalcost < 1) && (fabs(cost_improvement) < stomp_config_.min_cost_improvement_)) break;

        if (debug_config_.save_noisy_trajectories_ && num_rollouts_file) {
            std::vector<stomp::Rollout> rollouts;
            stomp_->getAllRollouts(rollouts);
            fprintf(num_rollouts_file, "%d\n", int(rollouts.size()));
            for (unsigned int j = 0; j < rollouts.size(); ++j) {
                std::stringstream ss2;
                ss2 << debug_config_.output_dir_ << "/noisy_" << i + 1 << "_" << j << ".txt";
                tmp_policy.setParameters(rollouts[j].parameters_noise_);
                tmp_policy.writeToFile(ss2.str());
            }
        }
        if (debug_config_.save_noiseless_trajectories_) {
            // the parameters live on the device during the solve: fetch them for the dump
            std::vector<base::VectorXd> p;
            std::stringstream ss;
            ss << debug_config_.output_dir_ << "/noiseless_" << i + 1 << ".txt";
            if (stomp_->getParameters(p)) {
                tmp_policy.setParameters(p);
                tmp_policy.writeToFile(ss.str());
            }
        }
    }
    if (num_rollouts_file) fclose(num_rollouts_file);

    // parameters_all_ <- device (the reference's policy object is updated in place by updateParameters)
    const bool synced = stomp_->syncPolicyFromDevice();
    stomp_.reset();
    if (!synced) {
        planner_status.statuscode = motion_planners::PlannerStatus::CRASH;
        return false;
    }

    const int start = stomp::DIFF_RULE_LENGTH - 1;
    solution.names.resize(planning_group_joints_name_.size());
    solution.elements.resize(planning_group_joints_name_.size());
    for (int d = 0; d < stomp_config_.num_dimensions_; d++) {
        solution.names.at(d) = planning_group_joints_name_.at(d);
        solution.elements.at(d).resize(stomp_config_.num_time_steps_);
        for (int i = 0; i < stomp_config_.num_time_steps_; i++)
            solution.elements.at(d).at(i).position = optimization_task_->policy_->parameters_all_[d](i + start);
    }

    if ((current_trajectory_totalcost < 1) && (fabs(cost_improvement) <= stomp_config_.min_cost_improvement_)) {
        planner_status.statuscode = motion_planners::PlannerStatus::PATH_FOUND;
        return true;
    } else
        planner_status.statuscode = motion_planners::PlannerStatus::NO_PATH_FOUND;
    return false;
}

void StompPlanner::setStartGoalTrajectory(const base::samples::Joints& start, const base::samples::Joints& goal)
{
    optimization_task_->updateTrajectory(start, goal);
    optimization_task_->input_initial_trajectory_ = optimization_task_->initial_trajectory_;
    optimization_task_->createPolicy();
}

bool StompPlanner::updateInitialTrajectory(const base::JointsTrajectory& trajectory)
{
    if (trajectory.empty()) return false;
    const int P = stomp::TRAJECTORY_PADDING, T = stomp_config_.num_time_steps_;
    for (int d = 0; d < stomp_config_.num_dimensions_; ++d) {
        for (int i = 0; i < P; ++i) {
            optimization_task_->initial_trajectory_[d](i) = trajectory.elements.at(d).front().position;
            optimization_task_->initial_trajectory_[d](P + T + i) = trajectory.elements.at(d).back().position;
        }
        for (int i = 0; i < T; i++) optimization_task_->initial_trajectory_[d](P + i) = trajectory.elements.at(d).at(i).position;
    }
    optimization_task_->input_initial_trajectory_ = optimization_task_->initial_trajectory_;
    optimization_task_->updatePolicy();
    return true;
}

base::JointsTrajectory StompPlanner::getInitialTrajectory()
{
    const int start = stomp::DIFF_RULE_LENGTH - 1;
    base::JointsTrajectory trajectory;
    trajectory.names.resize(planning_group_joints_name_.size());
    trajectory.elements.resize(planning_group_joints_name_.size());
    for (int d = 0; d < stomp_config_.num_dimensions_; d++) {
        trajectory.names.at(d) = planning_group_joints_name_.at(d);
        trajectory.elements.at(d).resize(stomp_config_.num_time_steps_);
        for (int i = 0; i < stomp_config_.num_time_steps_; i++)
            trajectory.elements.at(d).at(i).position = optimization_task_->input_initial_trajectory_[d](i + start);
    }
    return trajectory;
}

double StompPlanner::getMovementDeltaTime()
{
    if (optimization_task_ && optimization_task_->policy_) return optimization_task_->policy_->getMovementDt();
    return 0.0;
}

}  // namespace motion_planners
