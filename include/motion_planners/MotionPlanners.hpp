// motion_planners::MotionPlanners — the facade test_motion_planners drives (reference
// include/motion_planners/MotionPlanners.hpp:24-224), reduced to the calls on the STOMP path:
// initialize, assignPlanningRequest (joint-space start / goal), setStartAndGoal, solve (wall-clock timed,
// reference src/MotionPlanners.cpp:503-515), usePredictedTrajectory, reInitializePlanner.
#pragma once
#include <memory>
#include <string>

#include <base/JointsTrajectory.hpp>
#include <base/samples/Joints.hpp>
#include <motion_planners/Config.hpp>
#include <abstract/AbstractPlanner.hpp>
#include <PlannerFactory.hpp>
#include <robot_model/RobotModel.hpp>

namespace motion_planners {

class MotionPlanners {
public:
    explicit MotionPlanners(Config config);
    ~MotionPlanners();
    bool initialize(PlannerStatus& planner_status);
    bool reInitializePlanner();
    bool reInitializePlanner(const int& num_time_steps);
    bool assignPlanningRequest(const base::samples::Joints& start_jointvalues, const base::samples::Joints& target_jointvalues,
                               PlannerStatus& planner_status);
    bool usePredictedTrajectory(base::JointsTrajectory& input_trajectory, PlannerStatus& planner_status);
    void setStartAndGoal();
    bool solve(base::JointsTrajectory& solution, PlannerStatus& planner_status, double& time_taken);
    // world and grasp objects, octomap (reference src/MotionPlanners.cpp:162-173,416-495): they change the scene the distance
    // field is built from (on the device, before the next validity check or solve) resp. the sphere list of the arm
    bool handleCollisionObjectInWorld(const ModelObject& known_object);
    bool handleGraspObject(const ModelObject& known_object);
    void updateOctomap(const OccupiedLeaves& octomap);
    void assignOctomapPlanningScene(const OccupiedLeaves& octomap);
    std::shared_ptr<robot_model::RobotModel> getRobotModel() { return robot_model_; }

    AbstractPlannerPtr planner_;

private:
    bool checkStartState(const base::samples::Joints& current_robot_status, PlannerStatus& planner_status);
    bool checkGoalState(const base::samples::Joints& goal, PlannerStatus& planner_status);
    bool checkNaN(const base::samples::Joints& joint_value) const;
    bool selectPlanningGroupJoints(const base::samples::Joints& given, const char* what, base::samples::Joints& out,
                                   bool& within_limits) const;
    Config config_;
    std::shared_ptr<robot_model::RobotModel> robot_model_;
    base::samples::Joints initial_joint_status_, goal_joint_status_;
    ConstraintPlanning constrainted_target_;
};

}  // namespace motion_planners
