// motion_planners::OptimizationTask — the STOMP task of the wrapper (reference
// src/planners/include/wrapper/stomp/OptimizationTask.hpp:14-121).  Same members; the per-rollout work
// (execute for K rollouts, filter) happens inside the CUDA loop, which this object configures: it owns
// the stomp_b200_engine so that the scene (chain, spheres, SDF) stays resident in HBM across solves.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include <boost/shared_ptr.hpp>
#include <base/samples/Joints.hpp>
#include <robot_model/RobotModel.hpp>
#include <stomp/Stomp.hpp>
#include <stomp/StompTask.hpp>
#include <abstract/AbstractPlanner.hpp>

struct stomp_b200_engine;

namespace motion_planners {

class OptimizationTask : public stomp::StompTask, public boost::enable_shared_from_this<motion_planners::OptimizationTask> {
public:
    OptimizationTask(stomp::StompConfig config, std::shared_ptr<robot_model::RobotModel>& robot_model);
    virtual ~OptimizationTask();

    bool stompInitialize(int num_threads, int num_rollouts);
    void updateTrajectory(const base::samples::Joints& start, const base::samples::Joints& goal);
    void createPolicy();
    void updatePolicy();

    // one trajectory through the CUDA verdict kernel (the loop itself costs all rollouts on the device)
    virtual bool execute(std::vector<base::VectorXd>& parameters, std::vector<base::VectorXd>& projected_parameters,
                         base::VectorXd& costs, base::MatrixXd& weighted_feature_values, const int iteration_number,
                         const int rollout_number, int thread_id, bool compute_gradients,
                         std::vector<base::VectorXd>& gradients, bool& validity);
    virtual bool filter(std::vector<base::VectorXd>& parameters, int rollout_id, int thread_id);
    virtual bool getPolicy(boost::shared_ptr<stomp::CovariantMovementPrimitive>& policy);
    virtual bool setPolicy(const boost::shared_ptr<stomp::CovariantMovementPrimitive> policy);
    virtual double getControlCostWeight();
    void setOptimizationConstraints(ConstraintPlanning constraints) { constraints_ = constraints; costs_dirty_ = true; }
    // The reference computes the joint-constraint cost in computeJointsConstraintCost (OptimizationTask.cpp:206-216) but keeps
    // its call in execute() commented out (:169-172).  Off by default here too; when switched on, a JOINTS_CONSTRAINT set
    // with setOptimizationConstraints adds weight * sum_d max(0, |value_d - q_d| - tolerance_d) to every state cost.
    void useJointsConstraintCost(bool on, double weight = 1.0) { use_joints_constraint_cost_ = on; joints_constraint_weight_ = weight; costs_dirty_ = true; }
    // Smooth obstacle cost in place of the 0 / 1 collision cost (see stomp_b200_set_cost_extras); off by default.
    void useSmoothObstacleCost(bool on, double margin = 0.05, double weight = 1.0) { use_smooth_cost_ = on; smooth_margin_ = margin; smooth_weight_ = weight; costs_dirty_ = true; }

    // the device engine for this task's configuration, created on first use
    stomp_b200_engine* engine();
    const stomp::StompConfig& config() const { return stomp_config_; }

    boost::shared_ptr<stomp::CovariantMovementPrimitive> policy_;
    std::vector<base::VectorXd> initial_trajectory_, input_initial_trajectory_;

private:
    stomp::StompConfig stomp_config_;
    double movement_dt_ = 0.0;
    std::vector<base::MatrixXd> derivative_costs_;
    std::shared_ptr<robot_model::RobotModel> robot_model_;
    std::string planning_group_name_;
    std::vector<std::string> planning_group_joints_names_;
    std::vector<double> lower_limits_, upper_limits_;
    ConstraintPlanning constraints_;
    bool use_joints_constraint_cost_ = false, use_smooth_cost_ = false, costs_dirty_ = false;
    double joints_constraint_weight_ = 1.0, smooth_margin_ = 0.05, smooth_weight_ = 1.0;
    bool applyCostSwitches();
    stomp_b200_engine* engine_ = nullptr;
    unsigned long robot_revision_ = 0;    // robot_model_->robotRevision() the engine's sphere list was uploaded at (grasp objects change it)
    unsigned long scene_revision_ = 0;    // robot_model_->sceneRevision() the engine's distance field was built at
};

}  // namespace motion_planners
