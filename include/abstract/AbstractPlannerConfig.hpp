// Planner status / constraint types with the reference's names and values
// (reference src/planners/include/abstract/AbstractPlannerConfig.hpp:12-97).
#pragma once
#include <string>
#include <vector>

#include <base/Eigen.hpp>
#include <base/samples/Joints.hpp>
#include <kinematics_library/KinematicsConfig.hpp>

namespace motion_planners {

enum Constraint { POSITION_CONSTRAINT, ORIENTATION_CONSTRAINT, POSE_CONSTRAINT, JOINTS_CONSTRAINT, NO_CONSTRAINT };

struct ConstraintValues {
    base::VectorXd value;
    base::VectorXd tolerance;
};

struct ConstraintPlanning {
    ConstraintPlanning() : use_constraint(motion_planners::NO_CONSTRAINT) {}
    Constraint use_constraint;
    ConstraintValues orientation_constraint;
    ConstraintValues position_constraint;
    ConstraintValues joint_constraint;
    base::samples::Joints target_joints_value;
};

struct PlannerStatus {
    enum StatusCode {
        PATH_FOUND, NO_PATH_FOUND, START_STATE_IN_COLLISION, GOAL_STATE_IN_COLLISION, START_JOINTANGLES_NOT_AVAILABLE,
        GOAL_JOINTANGLES_NOT_AVAILABLE, CONSTRAINED_POSE_NOT_WITHIN_BOUNDS, PLANNING_REQUEST_SUCCESS, TIMEOUT,
        INVALID_START_STATE, INVALID_GOAL_STATE, UNRECOGNIZED_GOAL_TYPE, APPROXIMATE_SOLUTION, EXACT_SOLUTION,
        ROBOTMODEL_INITIALISATION_FAILED, PLANNER_INITIALISATION_FAILED, NO_CONSTRAINT_AVAILABLE,
        JOINT_CONSTRAINT_SIZE_ERROR, CRASH, KINEMATIC_ERROR, INVALID
    } statuscode;
    kinematics_library::KinematicsStatus kinematic_status;
    PlannerStatus() : statuscode(INVALID) {}
};

}  // namespace motion_planners
