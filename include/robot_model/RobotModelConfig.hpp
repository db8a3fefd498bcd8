// Shim for robot_model/RobotModelConfig.hpp (planning/robot_model, not vendored by the reference).
// urdf_file / srdf_file / planning_group_name are the reference's fields
// (reference test/test_motion_planners.cpp:38-50); the rest describes what replaces the collision
// meshes + FCL in this build: link spheres and a signed distance field of the environment.
#pragma once
#include <string>

namespace robot_model {

struct RobotModelConfig {
    std::string urdf_file;
    std::string srdf_file;            // its disable_collisions entries are honoured when self_collision is on
    std::string planning_group_name;
    std::string base_link;            // chain root; empty = the URDF's root link
    std::string tip_link;             // chain tip; empty = follow the movable joints to the end
    std::string spheres_file;         // YAML: spheres: { <link name>: [x, y, z, r, x, y, z, r, ...] }
    std::string environment_file;     // YAML: sdf: { resolution, lower, upper }, obstacles: { name: [sphere|box, ...] }
    bool self_collision = false;      // also check link spheres against each other (all link pairs except joined links
                                      // and the SRDF's disabled pairs); off: link spheres vs. environment only
    int device = 0;                   // CUDA device used for state validity queries
};

}  // namespace robot_model
