// reference src/planners/src/PlannerFactory.cpp:13-46
#include <PlannerFactory.hpp>

namespace motion_planners {

PlannerFactory::PlannerFactory() {}
PlannerFactory::~PlannerFactory() {}

AbstractPlannerPtr PlannerFactory::getPlannerTask(motion_planners::PlannerLibrary library)
{
    AbstractPlannerPtr planner = NULL;
    switch (library) {
        case STOMP: {
            planner = std::shared_ptr<StompPlanner>(new StompPlanner());
            break;
        }
        case TRAJOPT: {
            LOG_FATAL_S << "[PlannerFactory]: TrajOpt is not part of this build (sequential SQP planner, outside the STOMP rollout path)";
            return NULL;
        }
        case OMPL: {
            LOG_FATAL_S << "[PlannerFactory]: OMPL is not installed. Please select an another Planner !";
            return NULL;
        }
        default: {
            std::cout << "No planner library selected" << std::endl;
            return NULL;
        }
    }
    return planner;
}

}  // namespace motion_planners
