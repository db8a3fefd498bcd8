// motion_planners::MotionPlanners facade, reduced to the calls on the STOMP path
// (reference src/MotionPlanners.cpp:16-60,91-160,175-219,497-515).
#include <motion_planners/MotionPlanners.hpp>
#include <robot_model/MeshTools.hpp>

#include <chrono>
#include <cmath>
#include <stdexcept>

namespace motion_planners {

MotionPlanners::MotionPlanners(Config config) : config_(config) {}
MotionPlanners::~MotionPlanners() {}

bool MotionPlanners::initialize(PlannerStatus& planner_status)
{
    robot_model_.reset(new robot_model::RobotModel(config_.planner_config.robot_model_config));
    if (!robot_model_->initialization()) {
        planner_status.statuscode = PlannerStatus::ROBOTMODEL_INITIALISATION_FAILED;
        return false;
    }
    PlannerFactory planner_factory;
    planner_ = planner_factory.getPlannerTask(config_.planner_config.planner);
    if (!planner_ || !planner_->initializePlanner(robot_model_, config_.planner_config.planner_specific_config)) {
        planner_status.statuscode = PlannerStatus::PLANNER_INITIALISATION_FAILED;
        return false;
    }
    return true;
}

bool MotionPlanners::reInitializePlanner() { return planner_ && planner_->reInitializePlanner(); }
bool MotionPlanners::reInitializePlanner(const int& num_time_steps) { return planner_ && planner_->reInitializeTimeSteps(num_time_steps); }

// reference :563-571
bool MotionPlanners::checkNaN(const base::samples::Joints& joint_value) const
{
    for (size_t i = 0; i < joint_value.elements.size(); i++)
        if (std::isnan(joint_value.elements.at(i).position)) return false;
    return true;
}

// Picks the planning-group joints out of `given` BY NAME, in the planning group's order (reference :108-125,196-211): a
// full-robot joint state, or the same joints in another order, gives the same request.  Beyond the reference: a value
// outside the joint's limits is refused here — the planner clamps every sample to the limits (OptimizationTask::filter),
// so such a start / goal can never be the end of a sampled trajectory.
bool MotionPlanners::selectPlanningGroupJoints(const base::samples::Joints& given, const char* what, base::samples::Joints& out,
                                               bool& within_limits) const
{
    std::vector<std::pair<std::string, urdf::Joint> > group;
    if (!robot_model_->getPlanningGroupJointInformation(robot_model_->getPlanningGroupName(), group)) return false;
    out.clear();
    out.resize(group.size());
    within_limits = true;
    for (size_t i = 0; i < group.size(); i++) {
        try {
            out.names.at(i) = group.at(i).first;
            out.elements.at(i) = given.getElementByName(group.at(i).first);
        } catch (const std::exception&) {
            LOG_ERROR_S << "[MotionPlanners]: Joint " << group.at(i).first << " is given in planning group but is not available in the given " << what << " value";
            return false;
        }
        const double q = out.elements.at(i).position;
        const urdf::Joint& j = group.at(i).second;
        if (j.type != urdf::Joint::CONTINUOUS && (q < j.lower || q > j.upper)) {
            LOG_ERROR_S << "[MotionPlanners]: " << what << " value " << q << " of joint " << group.at(i).first << " is outside its limits [" << j.lower << ", " << j.upper << "]";
            within_limits = false;
        }
    }
    return true;
}

// reference :91-128: no NaN, collision free, then the start taken joint by joint by name
bool MotionPlanners::checkStartState(const base::samples::Joints& current_robot_status, PlannerStatus& planner_status)
{
    if (current_robot_status.empty() || !checkNaN(current_robot_status)) {
        planner_status.statuscode = PlannerStatus::START_JOINTANGLES_NOT_AVAILABLE;
        return false;
    }
    base::samples::Joints start;
    bool within_limits = true;
    if (!selectPlanningGroupJoints(current_robot_status, "start", start, within_limits)) {
        planner_status.statuscode = PlannerStatus::START_JOINTANGLES_NOT_AVAILABLE;
        return false;
    }
    if (!within_limits) {
        planner_status.statuscode = PlannerStatus::INVALID_START_STATE;
        return false;
    }
    double collision_cost = 0.0;
    robot_model_->updateJointGroup(current_robot_status);
    if (!robot_model_->isStateValid(collision_cost)) {
        planner_status.statuscode = PlannerStatus::START_STATE_IN_COLLISION;
        return false;
    }
    initial_joint_status_ = start;
    return true;
}

// reference :130-160
bool MotionPlanners::checkGoalState(const base::samples::Joints& goal, PlannerStatus& planner_status)
{
    if (goal.empty() || !checkNaN(goal)) {
        planner_status.statuscode = PlannerStatus::GOAL_JOINTANGLES_NOT_AVAILABLE;
        return false;
    }
    double collision_cost = 0.0;
    robot_model_->updateJointGroup(goal);
    if (!robot_model_->isStateValid(collision_cost)) {
        planner_status.statuscode = PlannerStatus::GOAL_STATE_IN_COLLISION;
        return false;
    }
    planner_status.statuscode = PlannerStatus::PLANNING_REQUEST_SUCCESS;
    return true;
}

// reference :188-219 (joint-space target): the goal is assembled by name before it is checked
bool MotionPlanners::assignPlanningRequest(const base::samples::Joints& start_jointvalues, const base::samples::Joints& target_jointvalues,
                                           PlannerStatus& planner_status)
{
    if (!checkStartState(start_jointvalues, planner_status)) return false;
    if (target_jointvalues.empty()) {
        planner_status.statuscode = PlannerStatus::GOAL_JOINTANGLES_NOT_AVAILABLE;
        return false;
    }
    base::samples::Joints goal;
    bool within_limits = true;
    if (!selectPlanningGroupJoints(target_jointvalues, "target", goal, within_limits)) {
        planner_status.statuscode = PlannerStatus::GOAL_JOINTANGLES_NOT_AVAILABLE;
        return false;
    }
    if (!checkNaN(goal)) {
        planner_status.statuscode = PlannerStatus::GOAL_JOINTANGLES_NOT_AVAILABLE;
        return false;
    }
    if (!within_limits) {
        planner_status.statuscode = PlannerStatus::INVALID_GOAL_STATE;
        return false;
    }
    if (!checkGoalState(goal, planner_status)) return false;
    goal_joint_status_ = goal;
    constrainted_target_.use_constraint = motion_planners::NO_CONSTRAINT;
    planner_status.statuscode = PlannerStatus::PLANNING_REQUEST_SUCCESS;
    return true;
}

// reference :175-186
bool MotionPlanners::usePredictedTrajectory(base::JointsTrajectory& input_trajectory, PlannerStatus& planner_status)
{
    if (!planner_->updateInitialTrajectory(input_trajectory)) {
        planner_status.statuscode = PlannerStatus::INVALID;
        return false;
    }
    return true;
}

void MotionPlanners::setStartAndGoal()
{
    planner_->setConstraints(constrainted_target_);
    planner_->setStartGoalTrajectory(initial_joint_status_, goal_joint_status_);
}

bool MotionPlanners::solve(base::JointsTrajectory& solution, PlannerStatus& planner_status, double& time_taken)
{
    auto start_time = std::chrono::high_resolution_clock::now();
    bool res = planner_->solve(solution, planner_status);
    auto finish_time = std::chrono::high_resolution_clock::now();
    std::chrono::duration<double> elapsed = finish_time - start_time;
    time_taken = elapsed.count();
    return res;
}

// ---- world / grasp objects (reference src/MotionPlanners.cpp:416-495) ----
namespace {

// the object's triangles in the frame it is given in (primitives are meshed for the sphere fit of grasp objects)
bool object_triangles(const ModelObject& o, std::vector<double>& tris)
{
    const double c[3] = {o.relative_pose.position.x(), o.relative_pose.position.y(), o.relative_pose.position.z()};
    if (o.model_type == collision_detection::MESH) return robot_model::loadStl(o.object_path, tris, nullptr, c);
    if (o.model_type != collision_detection::PRIMITIVES) return false;
    const collision_detection::PrimitiveObject& p = o.primitive_object;
    if (p.primitive_type == collision_detection::BOX) {
        const double half[3] = {0.5 * p.dimensions.x(), 0.5 * p.dimensions.y(), 0.5 * p.dimensions.z()};
        robot_model::appendBoxMesh(c, half, tris);
    } else if (p.primitive_type == collision_detection::CYLINDER) robot_model::appendCylinderMesh(c, p.radius, 0.5 * p.height, tris);
    else if (p.primitive_type == collision_detection::SPHERE) robot_model::appendSphereMesh(c, p.radius, tris);
    else return false;
    return true;
}

}  // namespace

bool MotionPlanners::handleCollisionObjectInWorld(const ModelObject& known_object)
{
    if (known_object.operation == collision_detection::RESET) { LOG_INFO_S << "[MotionPlanners]: Received known object with RESET"; return false; }
    if (known_object.model_type == collision_detection::UNDEFINED) { LOG_INFO_S << "[MotionPlanners]: object " << known_object.object_name << " is of UNDEFINED type"; return false; }
    if (known_object.operation == collision_detection::REMOVE) {
        if (known_object.model_type == collision_detection::OCTREE) { LOG_WARN_S << "[MotionPlanners]: removing a region from the octomap: hand in the updated leaves with updateOctomap"; return false; }
        return robot_model_->removeObstacle(known_object.object_name);
    }
    if (known_object.operation != collision_detection::ADD) return false;
    if (!known_object.relative_pose.orientation_is_identity) { LOG_ERROR_S << "[MotionPlanners]: world objects must be axis aligned in the world frame"; return false; }
    const base::Vector3d& c = known_object.relative_pose.position;
    if (known_object.model_type == collision_detection::PRIMITIVES) {
        const collision_detection::PrimitiveObject& p = known_object.primitive_object;
        robot_model::Obstacle o;
        o.name = known_object.object_name;
        o.centre[0] = c.x(); o.centre[1] = c.y(); o.centre[2] = c.z();
        if (p.primitive_type == collision_detection::BOX) { o.kind = 1; o.size[0] = 0.5 * p.dimensions.x(); o.size[1] = 0.5 * p.dimensions.y(); o.size[2] = 0.5 * p.dimensions.z(); }
        else if (p.primitive_type == collision_detection::CYLINDER) { o.kind = 2; o.size[0] = p.radius; o.size[1] = 0.5 * p.height; o.size[2] = 0.0; }
        else if (p.primitive_type == collision_detection::SPHERE) { o.kind = 0; o.size[0] = o.size[1] = o.size[2] = p.radius; }
        else return false;
        robot_model_->removeObstacle(o.name);
        robot_model_->addObstacle(o);
        return true;
    }
    if (known_object.model_type == collision_detection::MESH) {
        const double pos[3] = {c.x(), c.y(), c.z()};
        robot_model_->removeObstacle(known_object.object_name);
        return robot_model_->addMeshObstacleFromStl(known_object.object_name, known_object.object_path, pos);
    }
    return false;
}

bool MotionPlanners::handleGraspObject(const ModelObject& known_object)
{
    if (known_object.operation == collision_detection::RESET) { LOG_INFO_S << "[MotionPlanners]: Received grasp object with RESET"; return false; }
    if (known_object.operation == collision_detection::REMOVE) return robot_model_->removeGraspObject(known_object.object_name);
    if (known_object.operation != collision_detection::ADD) return false;
    std::vector<double> tris;
    if (!object_triangles(known_object, tris)) return false;
    robot_model::GraspObject g;
    g.name = known_object.object_name;
    for (const auto& f : robot_model::fitSpheres(tris, 8, 0.0)) {
        robot_model::CollisionSphere s;
        s.link = 0;
        for (int i = 0; i < 3; ++i) s.xyz[i] = f.xyz[i];
        s.radius = f.radius;
        g.spheres.push_back(s);
    }
    return robot_model_->addGraspObject(g, known_object.attach_link_name);
}

void MotionPlanners::updateOctomap(const OccupiedLeaves& octomap) { robot_model_->setOctomapLeaves(octomap.centres, octomap.sizes); }
void MotionPlanners::assignOctomapPlanningScene(const OccupiedLeaves& octomap) { robot_model_->setOctomapLeaves(octomap.centres, octomap.sizes); }

}  // namespace motion_planners
