// motion_planners::Config (reference include/motion_planners/Config.hpp:11-62), reduced to the fields the
// STOMP path reads; collision_detection / octomap configuration is out of scope.
#pragma once
#include <string>
#include <kinematics_library/KinematicsConfig.hpp>
#include <robot_model/RobotModelConfig.hpp>

namespace motion_planners {

enum PlannerLibrary { STOMP, OMPL, TRAJOPT };

struct PlannerConfig {
    kinematics_library::KinematicsConfig kinematics_config;
    robot_model::RobotModelConfig robot_model_config;
    std::string planner_specific_config;
    enum PlannerLibrary planner;
};

struct EnvironmentConfig {
    std::string env_frame;
    std::string env_object_name;
};

struct Config {
    PlannerConfig planner_config;
    EnvironmentConfig env_config;
};

}  // namespace motion_planners
