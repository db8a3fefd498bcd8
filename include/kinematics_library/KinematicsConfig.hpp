// Shim for kinematics_library/KinematicsConfig.hpp: only the status type PlannerStatus embeds
// (reference src/planners/include/abstract/AbstractPlannerConfig.hpp:93).  Inverse kinematics is outside
// the STOMP rollout path.
#pragma once
#include <string>
namespace kinematics_library {
enum KinematicSolver { IKFAST, SRS, IK7DOF, KDL, TRACIK, OPT };
struct KinematicsStatus {
    enum StatusCode { KDL_TREE_FAILED, KDL_CHAIN_FAILED, URDF_FAILED, NO_KINEMATIC_SOLVER_FOUND, IK_FOUND, NO_IK_SOLUTION,
                      NO_FK_SOLUTION, IK_TIMEOUT, IK_JOINTLIMITS_VIOLATED, NO_CONFIG_FILE, CONFIG_READ_ERROR, INVALID_STATE,
                      APPROX_IK_SOLUTION } statuscode;
    KinematicsStatus() : statuscode(INVALID_STATE) {}
};
struct KinematicsConfig {
    std::string config_name, base_name, tip_name, urdf_file, solver_config_abs_path, solver_config_filename;
    KinematicSolver kinematic_solver = KDL;
};
}  // namespace kinematics_library
