// Base class of the planner plug-ins: the part of motion_planners::AbstractPlanner that is not inline
// (interface: reference src/planners/include/abstract/AbstractPlanner.hpp:83-156; behaviour to match:
// src/planners/src/abstract/AbstractPlanner.cpp:6-29 — frame names and the planning group's joints are taken from
// the robot model once, a missing joint chain is reported and fails the initialisation).
#include <abstract/AbstractPlanner.hpp>

namespace motion_planners {

AbstractPlanner::AbstractPlanner() : root_name_(), base_name_(), tip_name_() {}

bool AbstractPlanner::assignPlanningJointInformation(std::shared_ptr<robot_model::RobotModel> robot_model)
{
    robot_model_ = robot_model;
    const robot_model::RobotModel& model = *robot_model_;

    root_name_ = model.getWorldFrameName();
    base_name_ = model.getBaseFrameName();
    tip_name_ = model.getTipFrameName();
    planning_group_name_ = model.getPlanningGroupName();

    const bool have_chain =
        robot_model_->getPlanningGroupJointInformation(planning_group_name_, planning_group_joints_, planning_group_joints_name_);
    if (!have_chain)
        LOG_ERROR_S << "[AbstractPlanner]: planning group '" << planning_group_name_ << "' has no joint chain between '" << base_name_
                    << "' and '" << tip_name_ << "'";
    return have_chain;
}

}  // namespace motion_planners
