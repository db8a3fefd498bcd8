// Shim for robot_model/RobotModel.hpp (planning/robot_model is an un-vendored dependency of the
// reference: manifest.xml:10-20).  It keeps the members the STOMP path calls —
// getPlanningGroupName / getPlanningGroupJointsName / getPlanningGroupJointInformation / getJointLimits /
// getWorldFrameName / getBaseFrameName / getTipFrameName / updateJointGroup / isStateValid
// (reference OptimizationTask.cpp:11-14,190-192, AbstractPlanner.cpp:13-27) — and replaces KDL + FCL by
// what the CUDA path consumes: the joint chain, link spheres and a signed distance field.
// updateJointGroup + isStateValid are answered by the same CUDA verdict kernel the planner uses
// (stomp_b200_evaluate_states): there is no CPU collision checker in this build.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include <base/Eigen.hpp>
#include <base/samples/Joints.hpp>
#include <robot_model/RobotModelConfig.hpp>

struct stomp_b200_engine;

namespace urdf {
struct Joint {
    enum Type { UNKNOWN, REVOLUTE, CONTINUOUS, PRISMATIC, FLOATING, PLANAR, FIXED };
    std::string name, parent_link_name, child_link_name;
    int type = UNKNOWN;
    double origin_xyz[3] = {0, 0, 0}, origin_rpy[3] = {0, 0, 0}, axis[3] = {1, 0, 0};
    double lower = 0, upper = 0, velocity = 0, effort = 0;
};
}  // namespace urdf

namespace robot_model {

struct CollisionSphere { int link; double xyz[3]; double radius; };      // link = index of the chain joint
struct Obstacle { int kind; double centre[3]; double size[3]; std::string name; };   // kind 0 sphere (size[0] = r), 1 box (half extents), 2 cylinder along z (radius, half height)
struct MeshObstacle { std::string name; std::vector<double> triangles; bool solid = true; };   // triangles [n][3][3] in the world frame
struct GraspObject { std::string name; std::vector<CollisionSphere> spheres; };            // spheres in the frame of the link they ride on

struct SignedDistanceField {
    int dims[3] = {0, 0, 0};
    double origin[3] = {0, 0, 0};
    double voxel = 0;
    std::vector<float> grid;   // x fastest
};

class RobotModel {
public:
    explicit RobotModel(const RobotModelConfig& config);
    ~RobotModel();
    bool initialization();    // loads URDF chain, spheres, environment; builds the SDF

    // ---- the reference's interface -------------------------------------------------------------
    std::string getPlanningGroupName() const { return config_.planning_group_name; }
    void getPlanningGroupJointsName(const std::string& group, std::vector<std::string>& names) const;
    bool getPlanningGroupJointInformation(const std::string& group, std::vector<std::pair<std::string, urdf::Joint> >& joints,
                                          std::vector<std::string>& names) const;
    bool getPlanningGroupJointInformation(const std::string& group, std::vector<std::pair<std::string, urdf::Joint> >& joints) const;
    bool getJointLimits(std::vector<double>& lower, std::vector<double>& upper) const;
    std::string getWorldFrameName() const { return world_frame_; }
    std::string getBaseFrameName() const { return base_link_; }
    std::string getTipFrameName() const { return tip_link_; }
    void updateJointGroup(const std::vector<std::string>& names, const base::VectorXd& positions);
    void updateJointGroup(const base::samples::Joints& joints);
    bool isStateValid(double& collision_cost);

    // ---- what the CUDA path consumes -------------------------------------------------------------
    const std::vector<urdf::Joint>& chain() const { return chain_; }
    const std::vector<CollisionSphere>& spheres() const { return spheres_; }
    // host copy of the distance field, for inspection: (re)built on the host when the scene changed since.  The engines
    // never read it — they build their field on the device from the same scene (configureScene)
    const SignedDistanceField& sdf();
    const std::vector<Obstacle>& obstacles() const { return obstacles_; }
    // programmatic setup (instead of files)
    void setChain(const std::vector<urdf::Joint>& chain, const std::string& base_link, const std::string& tip_link);
    void setSpheres(const std::vector<CollisionSphere>& spheres);
    // self collision: sphere pairs checked against each other.  Derived from the links the spheres sit on: every pair
    // of different links, except links joined by one joint and the pairs an SRDF disables (link names).
    void disableCollisions(const std::string& link1, const std::string& link2);
    void setSelfCollision(bool on) { config_.self_collision = on; }
    std::vector<std::pair<int, int> > selfCollisionPairs() const;
    // ---- the scene: world objects of the reference (MotionPlanners::handleCollisionObjectInWorld / updateOctomap,
    // src/MotionPlanners.cpp:162-173,416-460) as what the distance field is built from.  Every change bumps the scene
    // revision; engines compare it with the revision they were configured at and rebuild their field (on the device)
    // before the next query / solve, so a change takes effect immediately, as in the reference.
    void addObstacle(const Obstacle& o) { obstacles_.push_back(o); sceneChanged(); }
    bool removeObstacle(const std::string& name);
    void clearObstacles() { obstacles_.clear(); meshes_.clear(); leaf_centres_.clear(); leaf_sizes_.clear(); sceneChanged(); }
    // mesh world objects (the reference's MESH model objects): voxelised on the device, conservatively, interior filled when
    // `solid`; and an octomap as its occupied leaves (assignOctomapPlanningScene / updateOctomap): centres [m][3], edge
    // lengths [m].  With either present the field is the exact distance transform of the voxelised scene
    // (stomp_b200_build_sdf_scene) on the grid of setSdfGrid (or of setOccupancy); primitives are voxelised into it too.
    void addMeshObstacle(const MeshObstacle& m) { meshes_.push_back(m); sceneChanged(); }
    bool addMeshObstacleFromStl(const std::string& name, const std::string& path, const double position[3], const double scale[3] = nullptr, bool solid = true);
    void setOctomapLeaves(const std::vector<double>& centres, const std::vector<double>& sizes) { leaf_centres_ = centres; leaf_sizes_ = sizes; sceneChanged(); }
    const std::vector<MeshObstacle>& meshObstacles() const { return meshes_; }
    // grasped objects (MotionPlanners::handleGraspObject -> robot_model addGraspObject / removeGraspObject): the object's
    // spheres ride on the chain link `link_name` (the tip link when empty) from now on; engines pick the new sphere list up
    // before their next use, like a scene change
    bool addGraspObject(const GraspObject& object, const std::string& link_name);
    bool removeGraspObject(const std::string& name);
    // spheres for every chain link that has none, fitted to the link's collision mesh / primitive of the URDF
    // (MeshTools.hpp: fitSpheres); returns the number of links that received spheres
    int fitSpheresFromUrdfGeometry(int max_spheres_per_link = 6, double padding = 0.0);
    unsigned long robotRevision() const { return robot_revision_; }
    // occupancy world [nz][ny][nx] (a voxelised mesh, or an octomap's leaves at the grid's resolution): the field is then
    // its exact Euclidean distance transform; primitives present at the same time are voxelised into it
    void setOccupancy(const int dims[3], const double origin[3], double voxel, const std::vector<unsigned char>& occupied);
    void clearOccupancy() { occupancy_.clear(); sceneChanged(); }
    void setSdfGrid(int resolution, const double lower[3], const double upper[3]);
    void setSdf(const SignedDistanceField& sdf);   // an explicit field (wins over obstacles / occupancy until the scene changes)
    bool buildSdf();   // host statement of the primitive field (same formulas as the device builder); fills sdf()
    unsigned long sceneRevision() const { return scene_revision_; }
    // push chain / spheres / self-collision pairs / scene into an engine (used by the planner and by isStateValid)
    int configureEngine(stomp_b200_engine* engine) const;
    // the scene alone: distance field built on the device (primitives: exact union distance; occupancy: exact EDT)
    int configureScene(stomp_b200_engine* engine) const;
    int device() const { return config_.device; }

private:
    bool loadUrdf(const std::string& path);
    bool loadSpheres(const std::string& path);
    bool loadSrdf(const std::string& path);
    bool loadEnvironment(const std::string& path);
    bool ensureValidityEngine();
    void sceneChanged() { sdf_dirty_ = true; sdf_explicit_ = false; ++scene_revision_; }
    void gridGeometry(int dims[3], double origin[3], double& voxel) const;

    RobotModelConfig config_;
    std::string world_frame_, base_link_, tip_link_;
    std::vector<urdf::Joint> all_joints_, chain_;
    std::vector<CollisionSphere> spheres_;
    std::vector<std::pair<std::string, std::string> > disabled_link_pairs_;
    std::vector<Obstacle> obstacles_;
    std::vector<MeshObstacle> meshes_;
    std::vector<double> leaf_centres_, leaf_sizes_;
    std::vector<CollisionSphere> link_spheres_;           // the robot's own spheres (spheres_ = these + the grasp objects')
    std::vector<std::pair<GraspObject, int> > grasp_objects_;   // object, chain link index
    struct LinkGeometry { std::string link; int kind = -1; double size[3] = {0, 0, 0}; double origin[3] = {0, 0, 0}; std::string mesh_file; double mesh_scale[3] = {1, 1, 1}; };
    std::vector<LinkGeometry> link_geometry_;             // collision geometry of the URDF links, for sphere fitting
    unsigned long robot_revision_ = 1;                    // bumped when the sphere list changes
    void rebuildSphereList();
    SignedDistanceField sdf_;
    int sdf_resolution_ = 64;
    double sdf_lower_[3] = {-1.5, -1.5, -1.5}, sdf_upper_[3] = {1.5, 1.5, 1.5};
    bool sdf_dirty_ = true;
    bool sdf_explicit_ = false;             // sdf_ was handed in by setSdf and is what the engines get
    unsigned long scene_revision_ = 1;
    unsigned long validity_robot_revision_ = 0;
    unsigned long validity_revision_ = 0;   // scene revision validity_engine_ was configured at
    std::vector<unsigned char> occupancy_;  // [nz][ny][nx], empty: primitive world
    int occ_dims_[3] = {0, 0, 0};
    double occ_origin_[3] = {0, 0, 0}, occ_voxel_ = 0;
    std::vector<double> joint_state_;
    stomp_b200_engine* validity_engine_ = nullptr;
};

}  // namespace robot_model
