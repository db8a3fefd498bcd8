// motion_planners::StompPlanner (reference src/planners/src/wrappers/stomp/StompPlanner.cpp): same control
// flow, the iterations run on the GPU through stomp::Stomp -> include/stomp_b200.h.
#include <wrapper/stomp/StompPlanner.hpp>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <sstream>
#include <string>

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

// Writes the planner's current parameters_all_ (free part) into `solution` (reference StompPlanner.cpp:148-163).
void StompPlanner::fillSolution(base::JointsTrajectory& solution) const
{
    const int first_free = stomp::DIFF_RULE_LENGTH - 1;
    const size_t D = planning_group_joints_name_.size();
    solution.names.resize(D);
    solution.elements.resize(D);
    for (int d = 0; d < stomp_config_.num_dimensions_; d++) {
        solution.names.at(d) = planning_group_joints_name_.at(d);
        solution.elements.at(d).resize(stomp_config_.num_time_steps_);
        for (int t = 0; t < stomp_config_.num_time_steps_; t++)
            solution.elements.at(d).at(t).position = optimization_task_->policy_->parameters_all_[d](t + first_free);
    }
}

// reference StompPlanner.cpp:65-174.  Two drivers of the same device loop:
//   * the production path queues the whole loop on the device (stomp::Stomp::solveOnDevice -> stomp_b200_solve): the stop
//     rule of :117 is evaluated there after every noise-less rollout, the host only polls a flag every few iterations;
//   * with the debug dumps of :122-141 switched on, the host needs the rollouts of every iteration, so it drives one
//     iteration at a time (runSingleIteration) and applies the stop rule itself, as the reference does.
// Both end with the LAST parameters as the solution and PATH_FOUND iff the last noise-less cost is below 1 and the last
// improvement within min_cost_improvement (:165-173).
bool StompPlanner::solve(base::JointsTrajectory& solution, PlannerStatus& planner_status)
{
    optimization_task_->setOptimizationConstraints(constraints_);
    stomp_.reset(new stomp::Stomp());
    if (!stomp_->initialize(stomp_config_, optimization_task_)) {
        planner_status.statuscode = motion_planners::PlannerStatus::PLANNER_INITIALISATION_FAILED;
        stomp_.reset();
        return false;
    }
    num_iterations_ = 0;
    const bool dumps = debug_config_.save_noiseless_trajectories_ || debug_config_.save_noisy_trajectories_;
    bool path_found = false;
    if (!dumps) {
        int used = 0;
        if (!stomp_->solveOnDevice(stomp_config_.num_iterations_, used)) {
            planner_status.statuscode = motion_planners::PlannerStatus::CRASH;
            stomp_.reset();
            return false;
        }
        num_iterations_ = used;
        path_found = stomp_->pathFound();
        LOG_DEBUG_S << "Iterations = " << used << ". Total Cost = " << stomp_->getNoiselessRolloutTotalCost();
        stomp_.reset();
    } else {
        if (!solveWithDumps(planner_status, path_found)) return false;
    }
    fillSolution(solution);
    planner_status.statuscode = path_found ? motion_planners::PlannerStatus::PATH_FOUND : motion_planners::PlannerStatus::NO_PATH_FOUND;
    return path_found;
}

// the host-driven loop, one device iteration per step, with the reference's per-iteration files
// (noiseless_<i>.txt, noisy_<i>_<j>.txt, num_rollouts.txt; StompPlanner.cpp:79-93,122-141)
bool StompPlanner::solveWithDumps(PlannerStatus& planner_status, bool& path_found)
{
    const std::string& dir = debug_config_.output_dir_;
    mkdir(dir.c_str(), 0755);
    FILE* counts = debug_config_.save_noisy_trajectories_ ? fopen((dir + "/num_rollouts.txt").c_str(), "w") : NULL;
    if (debug_config_.save_noiseless_trajectories_) optimization_task_->policy_->writeToFile(dir + "/noiseless_0.txt");
    tmp_policy = *optimization_task_->policy_;
    double previous = 0.0, improvement = 0.0, cost = 0.0;
    for (int it = 0; it < stomp_config_.num_iterations_; it++) {
        num_iterations_++;
        if (!stomp_->runSingleIteration(it)) {
            planner_status.statuscode = motion_planners::PlannerStatus::CRASH;
            if (counts) fclose(counts);
            stomp_.reset();
            return false;
        }
        cost = stomp_->getNoiselessRolloutTotalCost();
        improvement = cost - previous;
        previous = cost;
        LOG_DEBUG_S << "Iteration = " << it << ". Total Cost = " << cost << " . Cost improvement = " << improvement;
        // a total cost below 1 means no time step is in collision (each costs 1)
        if ((cost < 1) && (fabs(improvement) < stomp_config_.min_cost_improvement_)) break;
        if (counts) {
            std::vector<stomp::Rollout> rollouts;
            stomp_->getAllRollouts(rollouts);
            fprintf(counts, "%d\n", int(rollouts.size()));
            for (size_t j = 0; j < rollouts.size(); ++j) {
                tmp_policy.setParameters(rollouts[j].parameters_noise_);
                tmp_policy.writeToFile(dir + "/noisy_" + std::to_string(it + 1) + "_" + std::to_string(j) + ".txt");
            }
        }
        if (debug_config_.save_noiseless_trajectories_) {
            std::vector<base::VectorXd> current;      // the parameters live on the device during the solve
            if (stomp_->getParameters(current)) {
                tmp_policy.setParameters(current);
                tmp_policy.writeToFile(dir + "/noiseless_" + std::to_string(it + 1) + ".txt");
            }
        }
    }
    if (counts) fclose(counts);
    const bool synced = stomp_->syncPolicyFromDevice();     // parameters_all_ <- device
    stomp_.reset();
    if (!synced) {
        planner_status.statuscode = motion_planners::PlannerStatus::CRASH;
        return false;
    }
    path_found = (cost < 1) && (fabs(improvement) <= stomp_config_.min_cost_improvement_);
    return true;
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
