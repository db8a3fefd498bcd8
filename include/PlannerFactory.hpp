// motion_planners::PlannerFactory (reference src/planners/include/PlannerFactory.hpp:18-27).  STOMP is the
// planner of this build; OMPL and TrajOpt are other algorithms outside the rollout path and return NULL.
#pragma once
#include <motion_planners/Config.hpp>
#include <abstract/AbstractPlanner.hpp>
#include <wrapper/stomp/StompPlanner.hpp>

namespace motion_planners {

class PlannerFactory {
public:
    PlannerFactory();
    ~PlannerFactory();
    AbstractPlannerPtr getPlannerTask(motion_planners::PlannerLibrary library);
};

}  // namespace motion_planners
