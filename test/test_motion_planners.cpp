// test_motion_planners — the reference's demo driver (reference test/test_motion_planners.cpp) re-targeted to
// the STOMP planner of this build: same start / goal (reference :212-215), same call sequence
// (initialize -> assignPlanningRequest -> setStartAndGoal -> solve), planner = motion_planners::STOMP with
// config/stomp.yml.  Needs a CUDA device.
//
//   ./test_motion_planners <absolute path to this test folder>
#include <cassert>
#include <iostream>

#include <motion_planners/MotionPlanners.hpp>

using namespace motion_planners;

static robot_model::RobotModelConfig getRobotModelConfig(const std::string& test_folder_path)
{
    robot_model::RobotModelConfig config;
    config.urdf_file = test_folder_path + "/data/iiwa_chain.urdf";
    config.srdf_file = "";
    config.planning_group_name = "manipulator";
    config.base_link = "base_link";
    config.tip_link = "link_7";
    config.spheres_file = test_folder_path + "/data/iiwa_spheres.yml";
    config.environment_file = test_folder_path + "/data/environment.yml";
    return config;
}

static motion_planners::Config getMotionPlannerConfig(const std::string& test_folder_path)
{
    motion_planners::Config config;
    config.planner_config.robot_model_config = getRobotModelConfig(test_folder_path);
    config.planner_config.planner_specific_config = test_folder_path + "/config/stomp.yml";
    config.planner_config.planner = motion_planners::STOMP;
    config.env_config.env_frame = "base_link";
    config.env_config.env_object_name = "environment";
    return config;
}

static base::samples::Joints convertToBaseJoints(const std::vector<double>& data)
{
    base::samples::Joints joint_values;
    joint_values.names = {"joint_a1", "joint_a2", "joint_a3", "joint_a4", "joint_a5", "joint_a6", "joint_a7"};
    joint_values.elements.resize(7);
    assert(joint_values.size() == data.size());
    for (size_t i = 0; i < data.size(); i++) joint_values.elements[i].position = data[i];
    return joint_values;
}

static void printTrajectory(const base::JointsTrajectory& traj)
{
    std::cout << "Number of timestep :" << traj.getTimeSteps() << ". Number of joints = " << traj.getNumberOfJoints() << std::endl;
    for (size_t i = 0; i < traj.getTimeSteps(); i++) {
        for (size_t j = 0; j < traj.elements.size(); j++) std::cout << traj.elements[j][i].position << "  ";
        std::cout << std::endl;
    }
}

static void printPlannerStatus(motion_planners::PlannerStatus& planner_status)
{
    switch (planner_status.statuscode) {
        case PlannerStatus::PATH_FOUND: std::cout << "PATH_FOUND" << std::endl; break;
        case PlannerStatus::NO_PATH_FOUND: std::cout << "NO_PATH_FOUND" << std::endl; break;
        case PlannerStatus::START_STATE_IN_COLLISION: std::cout << "START_STATE_IN_COLLISION" << std::endl; break;
        case PlannerStatus::GOAL_STATE_IN_COLLISION: std::cout << "GOAL_STATE_IN_COLLISION" << std::endl; break;
        case PlannerStatus::START_JOINTANGLES_NOT_AVAILABLE: std::cout << "START_JOINTANGLES_NOT_AVAILABLE" << std::endl; break;
        case PlannerStatus::GOAL_JOINTANGLES_NOT_AVAILABLE: std::cout << "GOAL_JOINTANGLES_NOT_AVAILABLE" << std::endl; break;
        case PlannerStatus::PLANNING_REQUEST_SUCCESS: std::cout << "PLANNING_REQUEST_SUCCESS" << std::endl; break;
        case PlannerStatus::ROBOTMODEL_INITIALISATION_FAILED: std::cout << "ROBOTMODEL_INITIALISATION_FAILED" << std::endl; break;
        case PlannerStatus::PLANNER_INITIALISATION_FAILED: std::cout << "PLANNER_INITIALISATION_FAILED" << std::endl; break;
        case PlannerStatus::CRASH: std::cout << "CRASH" << std::endl; break;
        default: std::cout << "UNKNOWN_STATE" << std::endl; break;
    }
}

int main(int argc, char* argv[])
{
    if (argc != 2) {
        std::cout << "usage: test_motion_planners <absolute path to the test folder>" << std::endl;
        return 0;
    }
    const std::string test_folder_path = argv[1];
    motion_planners::Config config = getMotionPlannerConfig(test_folder_path);
    motion_planners::MotionPlanners planner(config);
    PlannerStatus planner_status;
    if (!planner.initialize(planner_status)) {
        std::cout << "Motion planner failed at initialization. Refer to planner status to get the error information" << std::endl;
        printPlannerStatus(planner_status);
        return 2;
    }
    // reference test/test_motion_planners.cpp:212-215
    std::vector<double> start_vec_values = {0.5, 0.5, 0.5, -1.5, 0.5, 0.5, 0.5};
    base::samples::Joints start_joint_values = convertToBaseJoints(start_vec_values);
    std::vector<double> target_vec_values = {-1.5, -1.5, -1.5, 1.5, -1.5, -1.5, -0.5};
    base::samples::Joints target_joint_values = convertToBaseJoints(target_vec_values);

    int rc = 1;
    if (planner.assignPlanningRequest(start_joint_values, target_joint_values, planner_status)) {
        planner.setStartAndGoal();
        double solving_time = 0.0;
        base::JointsTrajectory solution;
        if (planner.solve(solution, planner_status, solving_time)) {
            std::cout << "Path Found" << std::endl;
            printTrajectory(solution);
            rc = 0;
        } else {
            std::cout << "No Path Found. Refer to planner status to get the error information" << std::endl;
            printPlannerStatus(planner_status);
        }
        std::cout << "iterations used: " << planner.planner_->getNumOfIterationsUsed() << ", solve time: " << solving_time << " s" << std::endl;
    } else {
        std::cout << "Assigning planning request failed. Refer to planner status to get the error information" << std::endl;
        printPlannerStatus(planner_status);
    }
    return rc;
}
