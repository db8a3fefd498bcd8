// GPU test of the facade around the hot path (reference src/MotionPlanners.cpp:91-219,416-460): planning requests are
// assembled by joint NAME, NaN / missing / out-of-limit values are refused, and a world object added after
// initialize() takes effect at once — in the validity checks and in the next solve (the engines rebuild their distance
// field on the device when the scene revision changes).
//   ./scene_and_request_test <absolute path to the test folder>
#include <cmath>
#include <cstdio>
#include <iostream>

#include <motion_planners/MotionPlanners.hpp>

using namespace motion_planners;

#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

static base::samples::Joints joints(const std::vector<std::string>& names, const std::vector<double>& values)
{
    base::samples::Joints j;
    j.names = names;
    j.elements.resize(values.size());
    for (size_t i = 0; i < values.size(); ++i) j.elements[i].position = values[i];
    return j;
}

int main(int argc, char** argv)
{
    if (argc != 2) return 2;
    const std::string dir = argv[1];
    Config config;
    config.planner_config.robot_model_config.urdf_file = dir + "/data/iiwa_chain.urdf";
    config.planner_config.robot_model_config.planning_group_name = "manipulator";
    config.planner_config.robot_model_config.base_link = "base_link";
    config.planner_config.robot_model_config.tip_link = "link_7";
    config.planner_config.robot_model_config.spheres_file = dir + "/data/iiwa_spheres.yml";
    config.planner_config.robot_model_config.environment_file = dir + "/data/environment.yml";
    config.planner_config.planner_specific_config = dir + "/config/stomp.yml";
    config.planner_config.planner = STOMP;
    MotionPlanners planner(config);
    PlannerStatus status;
    CHECK(planner.initialize(status));

    const std::vector<std::string> names = {"joint_a1", "joint_a2", "joint_a3", "joint_a4", "joint_a5", "joint_a6", "joint_a7"};
    const std::vector<double> start = {0.5, 0.5, 0.5, -1.5, 0.5, 0.5, 0.5}, goal = {-1.5, -1.5, -1.5, 1.5, -1.5, -1.5, -0.5};

    // ---- by name: reversed order plus a joint that is not in the planning group ----
    std::vector<std::string> rnames(names.rbegin(), names.rend());
    std::vector<double> rstart(start.rbegin(), start.rend()), rgoal(goal.rbegin(), goal.rend());
    rnames.push_back("gripper_finger"); rstart.push_back(0.01); rgoal.push_back(0.02);
    CHECK(planner.assignPlanningRequest(joints(rnames, rstart), joints(rnames, rgoal), status));
    CHECK(status.statuscode == PlannerStatus::PLANNING_REQUEST_SUCCESS);
    planner.setStartAndGoal();
    base::JointsTrajectory initial = planner.planner_->getInitialTrajectory();
    for (int d = 0; d < 7; ++d) {
        CHECK(initial.names[d] == names[d]);
        CHECK(initial.elements[d].front().position == start[d]);
        CHECK(std::fabs(initial.elements[d].back().position - goal[d]) < 1e-12);
    }
    std::puts("ok request_by_name");

    // ---- refused requests ----
    {
        std::vector<double> bad = start;
        bad[2] = NAN;
        CHECK(!planner.assignPlanningRequest(joints(names, bad), joints(names, goal), status));
        CHECK(status.statuscode == PlannerStatus::START_JOINTANGLES_NOT_AVAILABLE);
        bad = goal; bad[5] = NAN;
        CHECK(!planner.assignPlanningRequest(joints(names, start), joints(names, bad), status));
        CHECK(status.statuscode == PlannerStatus::GOAL_JOINTANGLES_NOT_AVAILABLE);
        std::vector<std::string> six(names.begin(), names.end() - 1);
        CHECK(!planner.assignPlanningRequest(joints(six, std::vector<double>(start.begin(), start.end() - 1)), joints(names, goal), status));
        CHECK(status.statuscode == PlannerStatus::START_JOINTANGLES_NOT_AVAILABLE);
        bad = start; bad[1] = 2.5;      // limit of joint_a2 is 2.0942
        CHECK(!planner.assignPlanningRequest(joints(names, bad), joints(names, goal), status));
        CHECK(status.statuscode == PlannerStatus::INVALID_START_STATE);
        bad = goal; bad[6] = -3.2;
        CHECK(!planner.assignPlanningRequest(joints(names, start), joints(names, bad), status));
        CHECK(status.statuscode == PlannerStatus::INVALID_GOAL_STATE);
        CHECK(!planner.assignPlanningRequest(base::samples::Joints(), joints(names, goal), status));
        CHECK(status.statuscode == PlannerStatus::START_JOINTANGLES_NOT_AVAILABLE);
    }
    std::puts("ok refused_requests");

    // ---- solve, then change the world ----
    CHECK(planner.assignPlanningRequest(joints(names, start), joints(names, goal), status));
    planner.setStartAndGoal();
    base::JointsTrajectory solution;
    double seconds = 0.0;
    CHECK(planner.solve(solution, status, seconds) && status.statuscode == PlannerStatus::PATH_FOUND);
    std::shared_ptr<robot_model::RobotModel> robot = planner.getRobotModel();
    const unsigned long rev0 = robot->sceneRevision();
    // a ball around the tool at the goal configuration: the goal, free a moment ago, is now in collision
    robot_model::Obstacle ball;
    ball.kind = 0; ball.name = "late_ball";
    {
        // tool position at the goal from the solution's last state: put the ball on a state the robot provably occupies
        double cost = 0.0;
        robot->updateJointGroup(joints(names, goal));
        CHECK(robot->isStateValid(cost));
    }
    // the blocker of environment.yml sits at (-0.17, 0.03, 1.31) r 0.15 on the straight line; a large ball at the same
    // place reaches the arm at the goal configuration as well
    ball.centre[0] = -0.17; ball.centre[1] = 0.03; ball.centre[2] = 1.31; ball.size[0] = ball.size[1] = ball.size[2] = 1.2;
    robot->addObstacle(ball);
    CHECK(robot->sceneRevision() == rev0 + 1);
    CHECK(!planner.assignPlanningRequest(joints(names, start), joints(names, goal), status));
    CHECK(status.statuscode == PlannerStatus::START_STATE_IN_COLLISION || status.statuscode == PlannerStatus::GOAL_STATE_IN_COLLISION);
    CHECK(robot->removeObstacle("late_ball") && !robot->removeObstacle("late_ball"));
    CHECK(planner.assignPlanningRequest(joints(names, start), joints(names, goal), status));
    std::puts("ok scene_change_reaches_the_validity_checks");

    // a wall across the solved path: every state of the OLD solution that it cuts must be gone from the new one
    robot_model::Obstacle wall;
    wall.kind = 1; wall.name = "late_wall";
    wall.centre[0] = 0.0; wall.centre[1] = 0.75; wall.centre[2] = 0.9; wall.size[0] = 0.6; wall.size[1] = 0.08; wall.size[2] = 0.25;
    robot->addObstacle(wall);
    auto colliding_states = [&](const base::JointsTrajectory& traj) {
        int n = 0;
        for (size_t t = 0; t < traj.getTimeSteps(); ++t) {
            std::vector<double> q(7);
            for (int d = 0; d < 7; ++d) q[d] = traj.elements[d][t].position;
            double cost = 0.0;
            robot->updateJointGroup(joints(names, q));
            if (!robot->isStateValid(cost)) ++n;
        }
        return n;
    };
    const int old_hits = colliding_states(solution);
    if (planner.assignPlanningRequest(joints(names, start), joints(names, goal), status)) {
        planner.setStartAndGoal();
        base::JointsTrajectory second;
        const bool found = planner.solve(second, status, seconds);
        const int new_hits = colliding_states(second);
        std::printf("old solution: %d states inside the late wall; new solve: found=%d, %d states in collision\n", old_hits, (int)found, new_hits);
        // the planner's own verdict and the validity engine agree about the new world
        CHECK(found == (new_hits == 0) || !found);
        if (found) CHECK(new_hits == 0);
    }
    std::puts("ok scene_change_reaches_the_next_solve");
    return 0;
}
