// motion_planners::Config (reference include/motion_planners/Config.hpp:11-62), reduced to the fields the
// STOMP path reads, plus ModelObject (world / grasp objects) and the octomap's leaves.
#pragma once
#include <string>
#include <vector>
#include <kinematics_library/KinematicsConfig.hpp>
#include <robot_model/RobotModelConfig.hpp>

#include <base/Eigen.hpp>

// the slice of collision_detection's types that ModelObject uses (reference collision_detection/CollisionConfig.hpp;
// the package is not vendored — same names and meaning)
namespace collision_detection {
enum Operation { RESET, ADD, REMOVE };
enum ModelTypes { UNDEFINED, PRIMITIVES, MESH, OCTREE };
enum PrimitiveObjectTypes { UNDEFINED_PRIMITIVES, BOX, CYLINDER, SPHERE };
struct PrimitiveObject {
    PrimitiveObject() : primitive_type(UNDEFINED_PRIMITIVES), dimensions(base::Vector3d::Zero()), radius(0.0), height(0.0) {}
    PrimitiveObjectTypes primitive_type;
    base::Vector3d dimensions;   // box: full edge lengths
    double radius, height;       // sphere / cylinder
};
}  // namespace collision_detection

namespace motion_planners {

// reference include/motion_planners/Config.hpp:14-35.  relative_pose: only the position is used — the distance-field
// builders take axis-aligned primitives and meshes as given (an object with a rotated pose is rejected, not mis-placed).
struct ObjectPose {
    ObjectPose() : position(base::Vector3d::Zero()), orientation_is_identity(true) {}
    base::Vector3d position;
    bool orientation_is_identity;
};
struct ModelObject {
    ModelObject() : operation(collision_detection::RESET), model_type(collision_detection::UNDEFINED), object_path(""), object_name("") {}
    collision_detection::Operation operation;
    collision_detection::ModelTypes model_type;
    collision_detection::PrimitiveObject primitive_object;
    std::string object_path;        // mesh file (STL)
    std::string object_name;
    std::string attach_link_name;   // world objects: the frame they are given in must be the world frame ("" or the base link)
    ObjectPose relative_pose;
};

// an octomap as the planner sees it: its occupied leaves (centre, edge length), in the world frame
struct OccupiedLeaves {
    std::vector<double> centres;    // [m][3]
    std::vector<double> sizes;      // [m]
};

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
