// stomp::StompTask — the reference's cost plug-in seam (reference src/planners/stomp/include/stomp/StompTask.hpp:48-116).
// Kept for source compatibility.  The CUDA loop evaluates the sphere-vs-SDF task of
// motion_planners::OptimizationTask on the device; execute() of that task remains callable (it runs
// the same CUDA verdict kernel for one trajectory).  Arbitrary user subclasses cannot be run by the
// device loop and are rejected by stomp::Stomp::initialize.
#pragma once
#include <memory>
#include <vector>

#include <boost/shared_ptr.hpp>
#include <base/Eigen.hpp>
#include <stomp/CovariantMovementPrimitive.hpp>

namespace stomp {

class StompTask {
public:
    StompTask() {}
    virtual ~StompTask() {}
    virtual bool stompInitialize(int num_threads, int num_rollouts) = 0;
    virtual bool execute(std::vector<base::VectorXd>& parameters, std::vector<base::VectorXd>& projected_parameters,
                         base::VectorXd& costs, base::MatrixXd& weighted_feature_values, const int iteration_number,
                         const int rollout_number, int thread_id, bool compute_gradients,
                         std::vector<base::VectorXd>& gradients, bool& validity) = 0;
    virtual bool filter(std::vector<base::VectorXd>& parameters, int rollout_id, int thread_id) = 0;
    virtual bool getPolicy(boost::shared_ptr<stomp::CovariantMovementPrimitive>& policy) = 0;
    virtual bool setPolicy(const boost::shared_ptr<stomp::CovariantMovementPrimitive> policy) = 0;
    virtual double getControlCostWeight() = 0;
    virtual void onEveryIteration() {}
};

}  // namespace stomp
