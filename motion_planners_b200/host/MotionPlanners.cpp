// motion_planners::MotionPlanners facade, reduced to the calls on the STOMP path
// (reference src/MotionPlanners.cpp:16-60,91-160,175-219,497-515).
#include <motion_planners/MotionPlanners.hpp>

#include <chrono>

namespace motion_planners {

MotionPlanners::MotionPlanners(Config config) : config_(config) {}
MotionPlanners::~MotionPlanners() {}

bool MotionPlanners::initialize(PlannerStatus& planner_status)
{
    robot_model_.reset(new robot_model::RobotModel(config_.planner_config.robot_model_config));
    if (!robot_model_->initialization()) {
        planner_status.statuscode = PlannerStatus::ROBOTMODEL_INITIALISATION_FAILED;
        return false;
    }
    PlannerFactory planner_factory;
    planner_ = planner_factory.getPlannerTask(config_.planner_config.planner);
    if (!planner_ || !planner_->initializePlanner(robot_model_, config_.planner_config.planner_specific_config)) {
        planner_status.statuscode = PlannerStatus::PLANNER_INITIALISATION_FAILED;
        return false;
    }
    return true;
}

bool MotionPlanners::reInitializePlanner() { return planner_ && planner_->reInitializePlanner(); }
bool MotionPlanners::reInitializePlanner(const int& num_time_steps) { return planner_ && planner_->reInitializeTimeSteps(num_time_steps); }

// reference :91-128: the start state must be collision free
bool MotionPlanners::checkStartState(const base::samples::Joints& current_robot_status, PlannerStatus& planner_status)
{
    if (current_robot_status.empty()) {
        planner_status.statuscode = PlannerStatus::START_JOINTANGLES_NOT_AVAILABLE;
        return false;
    }
    double collision_cost = 0.0;
    robot_model_->updateJointGroup(current_robot_status);
    if (!robot_model_->isStateValid(collision_cost)) {
        planner_status.statuscode = PlannerStatus::START_STATE_IN_COLLISION;
        return false;
    }
    initial_joint_status_ = current_robot_status;
    return true;
}

// reference :130-160
bool MotionPlanners::checkGoalState(const base::samples::Joints& goal, PlannerStatus& planner_status)
{
    if (goal.empty()) {
        planner_status.statuscode = PlannerStatus::GOAL_JOINTANGLES_NOT_AVAILABLE;
        return false;
    }
    double collision_cost = 0.0;
    robot_model_->updateJointGroup(goal);
    if (!robot_model_->isStateValid(collision_cost)) {
        planner_status.statuscode = PlannerStatus::GOAL_STATE_IN_COLLISION;
        return false;
    }
    goal_joint_status_ = goal;
    return true;
}

// reference :188-219 (joint-space target)
bool MotionPlanners::assignPlanningRequest(const base::samples::Joints& start_jointvalues, const base::samples::Joints& target_jointvalues,
                                           PlannerStatus& planner_status)
{
    if (!checkStartState(start_jointvalues, planner_status)) return false;
    if (!checkGoalState(target_jointvalues, planner_status)) return false;
    planner_status.statuscode = PlannerStatus::PLANNING_REQUEST_SUCCESS;
    return true;
}

// reference :175-186
bool MotionPlanners::usePredictedTrajectory(base::JointsTrajectory& input_trajectory, PlannerStatus& planner_status)
{
    if (!planner_->updateInitialTrajectory(input_trajectory)) {
        planner_status.statuscode = PlannerStatus::INVALID;
        return false;
    }
    return true;
}

void MotionPlanners::setStartAndGoal()
{
    planner_->setConstraints(constrainted_target_);
    planner_->setStartGoalTrajectory(initial_joint_status_, goal_joint_status_);
}

bool MotionPlanners::solve(base::JointsTrajectory& solution, PlannerStatus& planner_status, double& time_taken)
{
    auto start_time = std::chrono::high_resolution_clock::now();
    bool res = planner_->solve(solution, planner_status);
    auto finish_time = std::chrono::high_resolution_clock::now();
    std::chrono::duration<double> elapsed = finish_time - start_time;
    time_taken = elapsed.count();
    return res;
}

}  // namespace motion_planners
