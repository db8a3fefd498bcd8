// motion_planners::StompPlanner — public API identical to the reference's
// (reference src/planners/include/wrapper/stomp/StompPlanner.hpp:15-89); solve() drives the CUDA loop.
#pragma once
#include <sys/stat.h>

#include <boost/shared_ptr.hpp>
#include <abstract/AbstractPlanner.hpp>
#include "HandleStompConfig.hpp"
#include "OptimizationTask.hpp"

namespace motion_planners {

class StompPlanner : public motion_planners::AbstractPlanner {
public:
    StompPlanner();
    ~StompPlanner();
    bool initializePlanner(std::shared_ptr<robot_model::RobotModel>& robot_model, std::string config_file_path);
    bool reInitializePlanner();
    bool reInitializeTimeSteps(const int& num_time_steps);
    bool solve(base::JointsTrajectory& solution, PlannerStatus& planner_status);
    void setStartGoalTrajectory(const base::samples::Joints& start, const base::samples::Joints& goal);
    void setConstraints(const ConstraintPlanning constraints) { constraints_ = constraints; }
    bool updateInitialTrajectory(const base::JointsTrajectory& trajectory);
    base::JointsTrajectory getInitialTrajectory();
    size_t getNumOfIterationsUsed() { return num_iterations_; }
    double getMovementDeltaTime();

    // additive: configuration without a YAML file (benchmarks, tests)
    bool initializePlanner(std::shared_ptr<robot_model::RobotModel>& robot_model, const stomp::StompConfig& config,
                           const stomp::DebugConfig& debug = stomp::DebugConfig());
    const stomp::StompConfig& getStompConfig() const { return stomp_config_; }

private:
    bool solveWithDumps(PlannerStatus& planner_status, bool& path_found);   // host-driven loop with the per-iteration debug files
    void fillSolution(base::JointsTrajectory& solution) const;
    boost::shared_ptr<stomp::Stomp> stomp_;
    stomp::StompConfig stomp_config_;
    stomp::DebugConfig debug_config_;
    std::shared_ptr<OptimizationTask> optimization_task_;
    stomp::CovariantMovementPrimitive tmp_policy;
    ConstraintPlanning constraints_;
    size_t num_iterations_;
};

}  // namespace motion_planners
