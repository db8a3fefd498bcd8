// robot_model::RobotModel shim: URDF joint chain + link spheres + environment SDF for the CUDA path.
#include <robot_model/RobotModel.hpp>
#include <robot_model/MeshTools.hpp>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>

#include <base-logging/Logging.hpp>
#include <yaml-cpp/yaml.h>

#include "../../include/stomp_b200.h"

namespace robot_model {

namespace {

// ---- a very small XML reader: enough for URDF <joint> / <link> elements -------------------------
struct XmlTag {
    std::string name;
    std::map<std::string, std::string> attr;
    bool closing = false, self_closing = false;
};

bool next_tag(const std::string& s, size_t& pos, XmlTag& tag)
{
    while (true) {
        size_t lt = s.find('<', pos);
        if (lt == std::string::npos) return false;
        if (s.compare(lt, 4, "<!--") == 0) { size_t e = s.find("-->", lt); if (e == std::string::npos) return false; pos = e + 3; continue; }
        if (s.compare(lt, 2, "<?") == 0) { size_t e = s.find("?>", lt); if (e == std::string::npos) return false; pos = e + 2; continue; }
        size_t gt = s.find('>', lt);
        if (gt == std::string::npos) return false;
        std::string body = s.substr(lt + 1, gt - lt - 1);
        pos = gt + 1;
        tag = XmlTag();
        if (!body.empty() && body[0] == '/') { tag.closing = true; body = body.substr(1); }
        if (!body.empty() && body.back() == '/') { tag.self_closing = true; body.pop_back(); }
        size_t i = 0;
        while (i < body.size() && !isspace((unsigned char)body[i])) ++i;
        tag.name = body.substr(0, i);
        while (i < body.size()) {
            while (i < body.size() && isspace((unsigned char)body[i])) ++i;
            size_t eq = body.find('=', i);
            if (eq == std::string::npos) break;
            std::string key = body.substr(i, eq - i);
            while (!key.empty() && isspace((unsigned char)key.back())) key.pop_back();
            size_t q1 = body.find_first_of("\"'", eq);
            if (q1 == std::string::npos) break;
            size_t q2 = body.find(body[q1], q1 + 1);
            if (q2 == std::string::npos) break;
            tag.attr[key] = body.substr(q1 + 1, q2 - q1 - 1);
            i = q2 + 1;
        }
        return true;
    }
}

void parse_triplet(const std::string& text, double out[3])
{
    std::stringstream ss(text);
    for (int i = 0; i < 3; ++i) ss >> out[i];
}

struct UrdfLinkGeometry { bool has = false; int kind = 0; double size[3] = {0, 0, 0}; double origin[3] = {0, 0, 0}; std::string mesh_file; double mesh_scale[3] = {1, 1, 1}; };

// signed distance to one primitive: the formulas of the device builder (csrc/sdf_builder.cuh), operation for operation
double obstacle_distance(const Obstacle& o, double x, double y, double z)
{
    const double dx = x - o.centre[0], dy = y - o.centre[1], dz = z - o.centre[2];
    if (o.kind == 0) return std::sqrt((dx * dx + dy * dy) + dz * dz) - o.size[0];
    if (o.kind == 2) {
        const double qr = std::sqrt(dx * dx + dy * dy) - o.size[0], qh = std::fabs(dz) - o.size[1];
        const double orr = std::fmax(qr, 0.0), oh = std::fmax(qh, 0.0);
        return std::sqrt(orr * orr + oh * oh) + std::fmin(std::fmax(qr, qh), 0.0);
    }
    const double qx = std::fabs(dx) - o.size[0], qy = std::fabs(dy) - o.size[1], qz = std::fabs(dz) - o.size[2];
    const double ox = std::fmax(qx, 0.0), oy = std::fmax(qy, 0.0), oz = std::fmax(qz, 0.0);
    const double outside = std::sqrt((ox * ox + oy * oy) + oz * oz);
    const double inside = std::fmin(std::fmax(std::fmax(qx, qy), qz), 0.0);
    return outside + inside;
}

}  // namespace

RobotModel::RobotModel(const RobotModelConfig& config) : config_(config) {}

RobotModel::~RobotModel()
{
    if (validity_engine_) stomp_b200_destroy(validity_engine_);
}

bool RobotModel::initialization()
{
    if (!config_.urdf_file.empty() && !loadUrdf(config_.urdf_file)) return false;
    if (chain_.empty()) { LOG_ERROR_S << "[RobotModel]: no kinematic chain"; return false; }
    if (!config_.spheres_file.empty() && !loadSpheres(config_.spheres_file)) return false;
    if (config_.self_collision && !config_.srdf_file.empty() && !loadSrdf(config_.srdf_file)) return false;
    if (!config_.environment_file.empty() && !loadEnvironment(config_.environment_file)) return false;
    if (spheres_.empty()) { LOG_ERROR_S << "[RobotModel]: no collision spheres"; return false; }
    // the distance field itself is built on the device when an engine is configured (configureScene)
    joint_state_.assign(chain_.size(), 0.0);
    return true;
}

bool RobotModel::loadUrdf(const std::string& path)
{
    std::ifstream f(path.c_str());
    if (!f) { LOG_ERROR_S << "[RobotModel]: cannot open URDF " << path; return false; }
    std::stringstream buf;
    buf << f.rdbuf();
    const std::string xml = buf.str();

    std::map<std::string, UrdfLinkGeometry> link_geometry;
    std::vector<std::string> link_names;
    all_joints_.clear();
    size_t pos = 0;
    XmlTag tag;
    std::string current_link;
    bool in_collision = false;
    urdf::Joint joint;
    bool in_joint = false;
    while (next_tag(xml, pos, tag)) {
        if (tag.name == "link" && !tag.closing && !in_joint) {
            current_link = tag.attr["name"];
            link_names.push_back(current_link);
            if (tag.self_closing) current_link.clear();
        } else if (tag.name == "link" && tag.closing) {
            current_link.clear();
        } else if (!current_link.empty() && tag.name == "collision") {
            in_collision = !tag.closing;
        } else if (!current_link.empty() && in_collision && tag.name == "origin" && !tag.closing) {
            parse_triplet(tag.attr.count("xyz") ? tag.attr["xyz"] : "0 0 0", link_geometry[current_link].origin);
        } else if (!current_link.empty() && in_collision && tag.name == "box" && !tag.closing) {
            UrdfLinkGeometry& g = link_geometry[current_link];
            g.has = true; g.kind = 1;
            parse_triplet(tag.attr["size"], g.size);
            for (double& s : g.size) s *= 0.5;
        } else if (!current_link.empty() && in_collision && tag.name == "sphere" && !tag.closing) {
            UrdfLinkGeometry& g = link_geometry[current_link];
            g.has = true; g.kind = 0;
            g.size[0] = std::atof(tag.attr["radius"].c_str());
        } else if (!current_link.empty() && in_collision && tag.name == "cylinder" && !tag.closing) {
            UrdfLinkGeometry& g = link_geometry[current_link];
            g.has = true; g.kind = 2;
            g.size[0] = std::atof(tag.attr["radius"].c_str());
            g.size[1] = 0.5 * std::atof(tag.attr["length"].c_str());
        } else if (!current_link.empty() && in_collision && tag.name == "mesh" && !tag.closing) {
            // <mesh filename="package://.../collision/link_1.stl" scale="..."/> (reference test/data/kuka_iiwa.urdf): kept for
            // sphere fitting; the path is taken relative to the URDF's directory when it is not absolute
            UrdfLinkGeometry& g = link_geometry[current_link];
            g.has = true; g.kind = 3;
            g.mesh_file = tag.attr["filename"];
            if (tag.attr.count("scale")) parse_triplet(tag.attr["scale"], g.mesh_scale);
        } else if (tag.name == "joint" && !tag.closing && current_link.empty() && tag.attr.count("type")) {
            joint = urdf::Joint();
            joint.name = tag.attr["name"];
            const std::string& t = tag.attr["type"];
            joint.type = t == "revolute" ? urdf::Joint::REVOLUTE : t == "continuous" ? urdf::Joint::CONTINUOUS
                       : t == "prismatic" ? urdf::Joint::PRISMATIC : t == "fixed" ? urdf::Joint::FIXED : urdf::Joint::UNKNOWN;
            in_joint = !tag.self_closing;
        } else if (in_joint && tag.name == "joint" && tag.closing) {
            all_joints_.push_back(joint);
            in_joint = false;
        } else if (in_joint && !tag.closing) {
            if (tag.name == "origin") {
                if (tag.attr.count("xyz")) parse_triplet(tag.attr["xyz"], joint.origin_xyz);
                if (tag.attr.count("rpy")) parse_triplet(tag.attr["rpy"], joint.origin_rpy);
            } else if (tag.name == "parent") joint.parent_link_name = tag.attr["link"];
            else if (tag.name == "child") joint.child_link_name = tag.attr["link"];
            else if (tag.name == "axis") parse_triplet(tag.attr["xyz"], joint.axis);
            else if (tag.name == "limit") {
                joint.lower = std::atof(tag.attr["lower"].c_str());
                joint.upper = std::atof(tag.attr["upper"].c_str());
                joint.velocity = std::atof(tag.attr["velocity"].c_str());
                joint.effort = std::atof(tag.attr["effort"].c_str());
            }
        }
    }
    // root link = a link that is nobody's child
    std::set<std::string> children;
    for (const auto& j : all_joints_) children.insert(j.child_link_name);
    std::string root;
    for (const auto& l : link_names)
        if (!children.count(l)) { root = l; break; }
    world_frame_ = root;
    base_link_ = config_.base_link.empty() ? root : config_.base_link;
    // follow the movable joints from the base link
    chain_.clear();
    std::string link = base_link_;
    tip_link_ = base_link_;
    while (true) {
        const urdf::Joint* next = nullptr;
        for (const auto& j : all_joints_)
            if (j.parent_link_name == link && j.type != urdf::Joint::FIXED && j.type != urdf::Joint::UNKNOWN) { next = &j; break; }
        if (!next) break;
        chain_.push_back(*next);
        link = next->child_link_name;
        tip_link_ = link;
        if (!config_.tip_link.empty() && link == config_.tip_link) break;
    }
    // collision geometry of the chain links: what fitSpheresFromUrdfGeometry fits spheres to
    link_geometry_.clear();
    const std::string urdf_dir = path.find('/') == std::string::npos ? std::string(".") : path.substr(0, path.rfind('/'));
    for (const auto& j : chain_) {
        auto it = link_geometry.find(j.child_link_name);
        if (it == link_geometry.end() || !it->second.has) continue;
        LinkGeometry g;
        g.link = j.child_link_name; g.kind = it->second.kind;
        for (int i = 0; i < 3; ++i) { g.size[i] = it->second.size[i]; g.origin[i] = it->second.origin[i]; g.mesh_scale[i] = it->second.mesh_scale[i]; }
        std::string file = it->second.mesh_file;
        const std::string pkg = "package://";
        if (file.compare(0, pkg.size(), pkg) == 0) { file = file.substr(pkg.size()); const size_t slash = file.find('/'); if (slash != std::string::npos) file = file.substr(slash + 1); }
        if (!file.empty() && file[0] != '/') file = urdf_dir + "/" + file;
        g.mesh_file = file;
        link_geometry_.push_back(g);
    }
    // links with a primitive collision body fixed to the base link are obstacles of the environment
    // (box_1 in reference test/data/kuka_iiwa.urdf:29-56)
    for (const auto& j : all_joints_) {
        if (j.type != urdf::Joint::FIXED || j.parent_link_name != base_link_) continue;
        auto it = link_geometry.find(j.child_link_name);
        if (it == link_geometry.end() || !it->second.has || it->second.kind == 3) continue;
        Obstacle o;
        o.kind = it->second.kind;
        o.name = j.child_link_name;
        for (int i = 0; i < 3; ++i) { o.centre[i] = j.origin_xyz[i] + it->second.origin[i]; o.size[i] = it->second.size[i]; }
        obstacles_.push_back(o);
        sceneChanged();
    }
    return !chain_.empty();
}

// the SRDF's <disable_collisions link1=".." link2=".."/> entries (reference test/data/kuka_iiwa.srdf:46-70)
bool RobotModel::loadSrdf(const std::string& path)
{
    std::ifstream f(path.c_str());
    if (!f) { LOG_ERROR_S << "[RobotModel]: cannot open SRDF " << path; return false; }
    std::stringstream buf;
    buf << f.rdbuf();
    const std::string xml = buf.str();
    size_t pos = 0;
    XmlTag tag;
    while (next_tag(xml, pos, tag))
        if (tag.name == "disable_collisions" && !tag.closing && tag.attr.count("link1") && tag.attr.count("link2"))
            disableCollisions(tag.attr["link1"], tag.attr["link2"]);
    return true;
}

void RobotModel::disableCollisions(const std::string& link1, const std::string& link2)
{
    disabled_link_pairs_.push_back(std::make_pair(link1, link2));
}

std::vector<std::pair<int, int> > RobotModel::selfCollisionPairs() const
{
    std::vector<std::pair<int, int> > pairs;
    if (!config_.self_collision) return pairs;
    auto disabled = [&](int a, int b) {
        const std::string& la = chain_[a].child_link_name;
        const std::string& lb = chain_[b].child_link_name;
        for (const auto& d : disabled_link_pairs_)
            if ((d.first == la && d.second == lb) || (d.first == lb && d.second == la)) return true;
        return false;
    };
    for (size_t i = 0; i < spheres_.size(); ++i)
        for (size_t j = i + 1; j < spheres_.size(); ++j) {
            const int a = spheres_[i].link, b = spheres_[j].link;    // sorted by link: a <= b
            if (a == b || b == a + 1 || disabled(a, b)) continue;     // same link, joined by one joint, SRDF
            pairs.push_back(std::make_pair((int)i, (int)j));
        }
    return pairs;
}

bool RobotModel::loadSpheres(const std::string& path)
{
    YAML::Node root;
    try { root = YAML::LoadFile(path); } catch (const YAML::Exception& e) { LOG_ERROR_S << "[RobotModel]: " << e.what(); return false; }
    const YAML::Node sp = root["spheres"];
    if (!sp) { LOG_ERROR_S << "[RobotModel]: no 'spheres' node in " << path; return false; }
    std::vector<CollisionSphere> out;
    for (size_t d = 0; d < chain_.size(); ++d) {
        const YAML::Node list = sp[chain_[d].child_link_name];
        if (!list) continue;
        if (list.size() % 4 != 0) { LOG_ERROR_S << "[RobotModel]: spheres of " << chain_[d].child_link_name << " must be x y z r groups"; return false; }
        for (size_t i = 0; i < list.size(); i += 4) {
            CollisionSphere s;
            s.link = (int)d;
            for (int k = 0; k < 3; ++k) s.xyz[k] = list[i + k].as<double>();
            s.radius = list[i + 3].as<double>();
            out.push_back(s);
        }
    }
    setSpheres(out);
    return !spheres_.empty();
}

bool RobotModel::loadEnvironment(const std::string& path)
{
    YAML::Node root;
    try { root = YAML::LoadFile(path); } catch (const YAML::Exception& e) { LOG_ERROR_S << "[RobotModel]: " << e.what(); return false; }
    if (const YAML::Node g = root["sdf"]) {
        int res = sdf_resolution_;
        double lo[3], hi[3];
        std::copy(sdf_lower_, sdf_lower_ + 3, lo);
        std::copy(sdf_upper_, sdf_upper_ + 3, hi);
        if (g["resolution"]) res = (int)g["resolution"].as<double>();
        if (g["lower"]) for (int i = 0; i < 3; ++i) lo[i] = g["lower"][i].as<double>();
        if (g["upper"]) for (int i = 0; i < 3; ++i) hi[i] = g["upper"][i].as<double>();
        setSdfGrid(res, lo, hi);
    }
    if (const YAML::Node obs = root["obstacles"]) {
        // obstacles: { name: [sphere, x, y, z, r] | [box, x, y, z, hx, hy, hz] } — names are not needed, walk by key
        // order of the underlying map (alphabetical): the union's distance does not depend on the order
        std::ifstream f(path.c_str());
        std::string line;
        bool in_obs = false;
        while (std::getline(f, line)) {
            const std::string t = YAML::detail::trim(YAML::detail::strip_comment(line));
            if (t == "obstacles:") { in_obs = true; continue; }
            if (!in_obs || t.empty()) continue;
            if (line[0] != ' ') { in_obs = false; continue; }
            const size_t colon = t.find(':');
            if (colon == std::string::npos) continue;
            const std::string key = YAML::detail::trim(t.substr(0, colon));
            const YAML::Node v = obs[key];
            if (!v || v.size() < 5) continue;
            Obstacle o;
            o.name = key;
            const std::string kind = v[0].as<std::string>();
            o.kind = kind == "sphere" ? 0 : (kind == "cylinder" ? 2 : 1);
            for (int i = 0; i < 3; ++i) o.centre[i] = v[1 + i].as<double>();
            if (o.kind == 0) { o.size[0] = o.size[1] = o.size[2] = v[4].as<double>(); }
            else if (o.kind == 2) { if (v.size() < 6) continue; o.size[0] = v[4].as<double>(); o.size[1] = v[5].as<double>(); o.size[2] = 0.0; }
            else { if (v.size() < 7) continue; for (int i = 0; i < 3; ++i) o.size[i] = v[4 + i].as<double>(); }
            addObstacle(o);
        }
    }
    return true;
}

void RobotModel::setChain(const std::vector<urdf::Joint>& chain, const std::string& base_link, const std::string& tip_link)
{
    chain_ = chain;
    base_link_ = base_link;
    tip_link_ = tip_link;
    if (world_frame_.empty()) world_frame_ = base_link;
    joint_state_.assign(chain_.size(), 0.0);
}

void RobotModel::setSpheres(const std::vector<CollisionSphere>& spheres)
{
    link_spheres_ = spheres;
    rebuildSphereList();
}

// spheres_ = the robot's own spheres + the spheres of the grasped objects, sorted by link (what the engines take)
void RobotModel::rebuildSphereList()
{
    spheres_ = link_spheres_;
    for (const auto& g : grasp_objects_)
        for (CollisionSphere s : g.first.spheres) { s.link = g.second; spheres_.push_back(s); }
    std::stable_sort(spheres_.begin(), spheres_.end(), [](const CollisionSphere& a, const CollisionSphere& b) { return a.link < b.link; });
    ++robot_revision_;
}

bool RobotModel::addGraspObject(const GraspObject& object, const std::string& link_name)
{
    int link = (int)chain_.size() - 1;
    if (!link_name.empty()) {
        link = -1;
        for (size_t d = 0; d < chain_.size(); ++d) if (chain_[d].child_link_name == link_name) link = (int)d;
    }
    if (link < 0 || object.spheres.empty()) return false;
    removeGraspObject(object.name);
    grasp_objects_.push_back(std::make_pair(object, link));
    rebuildSphereList();
    return true;
}

bool RobotModel::removeGraspObject(const std::string& name)
{
    for (size_t i = 0; i < grasp_objects_.size(); ++i)
        if (grasp_objects_[i].first.name == name) { grasp_objects_.erase(grasp_objects_.begin() + i); rebuildSphereList(); return true; }
    return false;
}

bool RobotModel::addMeshObstacleFromStl(const std::string& name, const std::string& path, const double position[3], const double scale[3], bool solid)
{
    MeshObstacle m;
    m.name = name; m.solid = solid;
    if (!loadStl(path, m.triangles, scale, position)) { LOG_ERROR_S << "[RobotModel]: cannot read STL " << path; return false; }
    addMeshObstacle(m);
    return true;
}

int RobotModel::fitSpheresFromUrdfGeometry(int max_spheres_per_link, double padding)
{
    int fitted = 0;
    std::vector<CollisionSphere> all = link_spheres_;
    for (size_t d = 0; d < chain_.size(); ++d) {
        bool has = false;
        for (const auto& s : link_spheres_) has = has || s.link == (int)d;
        if (has) continue;
        const LinkGeometry* g = nullptr;
        for (const auto& lg : link_geometry_) if (lg.link == chain_[d].child_link_name) g = &lg;
        if (!g) continue;
        std::vector<double> tris;
        if (g->kind == 3) { if (!loadStl(g->mesh_file, tris, g->mesh_scale, g->origin)) continue; }
        else if (g->kind == 1) appendBoxMesh(g->origin, g->size, tris);
        else if (g->kind == 2) appendCylinderMesh(g->origin, g->size[0], g->size[1], tris);
        else appendSphereMesh(g->origin, g->size[0], tris);
        const std::vector<FittedSphere> fit = fitSpheres(tris, max_spheres_per_link, padding);
        for (const auto& f : fit) { CollisionSphere s; s.link = (int)d; for (int i = 0; i < 3; ++i) s.xyz[i] = f.xyz[i]; s.radius = f.radius; all.push_back(s); }
        if (!fit.empty()) ++fitted;
    }
    if (fitted) setSpheres(all);
    return fitted;
}

void RobotModel::setSdfGrid(int resolution, const double lower[3], const double upper[3])
{
    sdf_resolution_ = resolution;
    std::copy(lower, lower + 3, sdf_lower_);
    std::copy(upper, upper + 3, sdf_upper_);
    sceneChanged();
}

void RobotModel::setSdf(const SignedDistanceField& sdf)
{
    sdf_ = sdf;
    sdf_dirty_ = false;
    sdf_explicit_ = true;
    ++scene_revision_;
}

bool RobotModel::removeObstacle(const std::string& name)
{
    for (size_t i = 0; i < meshes_.size(); ++i)
        if (meshes_[i].name == name) { meshes_.erase(meshes_.begin() + i); sceneChanged(); return true; }
    for (size_t i = 0; i < obstacles_.size(); ++i)
        if (obstacles_[i].name == name) {
            obstacles_.erase(obstacles_.begin() + i);
            sceneChanged();
            return true;
        }
    return false;
}

void RobotModel::setOccupancy(const int dims[3], const double origin[3], double voxel, const std::vector<unsigned char>& occupied)
{
    std::copy(dims, dims + 3, occ_dims_);
    std::copy(origin, origin + 3, occ_origin_);
    occ_voxel_ = voxel;
    occupancy_ = occupied;
    sceneChanged();
}

// cubic voxels sized by the x extent of the configured box
void RobotModel::gridGeometry(int dims[3], double origin[3], double& voxel) const
{
    const int n = std::max(2, sdf_resolution_);
    voxel = (sdf_upper_[0] - sdf_lower_[0]) / n;
    for (int i = 0; i < 3; ++i) {
        origin[i] = sdf_lower_[i];
        dims[i] = std::max(1, (int)std::lround((sdf_upper_[i] - sdf_lower_[i]) / voxel));
    }
}

bool RobotModel::buildSdf()
{
    if (sdf_resolution_ < 2) return false;
    gridGeometry(sdf_.dims, sdf_.origin, sdf_.voxel);
    const double h = sdf_.voxel;
    sdf_.grid.assign((size_t)sdf_.dims[0] * sdf_.dims[1] * sdf_.dims[2], 0.0f);
    for (int z = 0; z < sdf_.dims[2]; ++z)
        for (int y = 0; y < sdf_.dims[1]; ++y)
            for (int x = 0; x < sdf_.dims[0]; ++x) {
                const double px = sdf_.origin[0] + ((double)x + 0.5) * h, py = sdf_.origin[1] + ((double)y + 0.5) * h, pz = sdf_.origin[2] + ((double)z + 0.5) * h;
                double d = INFINITY;
                for (const auto& o : obstacles_) d = std::fmin(d, obstacle_distance(o, px, py, pz));
                sdf_.grid[((size_t)z * sdf_.dims[1] + y) * sdf_.dims[0] + x] = (float)d;
            }
    sdf_dirty_ = false;
    return true;
}

const SignedDistanceField& RobotModel::sdf()
{
    if (sdf_dirty_ && !sdf_explicit_) buildSdf();
    return sdf_;
}

void RobotModel::getPlanningGroupJointsName(const std::string&, std::vector<std::string>& names) const
{
    names.clear();
    for (const auto& j : chain_) names.push_back(j.name);
}

bool RobotModel::getPlanningGroupJointInformation(const std::string& group, std::vector<std::pair<std::string, urdf::Joint> >& joints,
                                                  std::vector<std::string>& names) const
{
    if (!getPlanningGroupJointInformation(group, joints)) return false;
    getPlanningGroupJointsName(group, names);
    return true;
}

bool RobotModel::getPlanningGroupJointInformation(const std::string&, std::vector<std::pair<std::string, urdf::Joint> >& joints) const
{
    joints.clear();
    for (const auto& j : chain_) joints.push_back(std::make_pair(j.name, j));
    return !joints.empty();
}

bool RobotModel::getJointLimits(std::vector<double>& lower, std::vector<double>& upper) const
{
    lower.clear(); upper.clear();
    for (const auto& j : chain_) { lower.push_back(j.lower); upper.push_back(j.upper); }
    return !chain_.empty();
}

void RobotModel::updateJointGroup(const std::vector<std::string>& names, const base::VectorXd& positions)
{
    joint_state_.resize(chain_.size(), 0.0);
    for (size_t i = 0; i < names.size() && (int)i < positions.size(); ++i)
        for (size_t d = 0; d < chain_.size(); ++d)
            if (chain_[d].name == names[i]) joint_state_[d] = positions((int)i);
}

void RobotModel::updateJointGroup(const base::samples::Joints& joints)
{
    joint_state_.resize(chain_.size(), 0.0);
    for (size_t i = 0; i < joints.names.size(); ++i)
        for (size_t d = 0; d < chain_.size(); ++d)
            if (chain_[d].name == joints.names[i]) joint_state_[d] = joints.elements[i].position;
}

int RobotModel::configureEngine(stomp_b200_engine* engine) const
{
    const int D = (int)chain_.size(), S = (int)spheres_.size();
    std::vector<double> xyz(3 * D), rpy(3 * D), axis(3 * D), lower(D), upper(D), sxyz(3 * S), rad(S);
    std::vector<int32_t> parent(D), prismatic(D), link(S);
    for (int d = 0; d < D; ++d) {
        for (int i = 0; i < 3; ++i) { xyz[3 * d + i] = chain_[d].origin_xyz[i]; rpy[3 * d + i] = chain_[d].origin_rpy[i]; axis[3 * d + i] = chain_[d].axis[i]; }
        parent[d] = d - 1;
        prismatic[d] = chain_[d].type == urdf::Joint::PRISMATIC ? 1 : 0;
        lower[d] = chain_[d].lower; upper[d] = chain_[d].upper;
    }
    for (int s = 0; s < S; ++s) {
        link[s] = spheres_[s].link;
        for (int i = 0; i < 3; ++i) sxyz[3 * s + i] = spheres_[s].xyz[i];
        rad[s] = spheres_[s].radius;
    }
    int rc = stomp_b200_set_chain(engine, D, xyz.data(), rpy.data(), axis.data(), parent.data(), prismatic.data(), lower.data(), upper.data());
    if (rc) return rc;
    rc = stomp_b200_set_spheres(engine, S, link.data(), sxyz.data(), rad.data());
    if (rc) return rc;
    const std::vector<std::pair<int, int> > pairs = selfCollisionPairs();
    if (!pairs.empty()) {
        std::vector<int32_t> flat;
        for (const auto& p : pairs) { flat.push_back(p.first); flat.push_back(p.second); }
        rc = stomp_b200_set_self_collision(engine, (int32_t)pairs.size(), flat.data());
        if (rc) return rc;
    }
    return configureScene(engine);
}

int RobotModel::configureScene(stomp_b200_engine* engine) const
{
    if (sdf_explicit_) return stomp_b200_set_sdf(engine, sdf_.dims, sdf_.origin, sdf_.voxel, sdf_.grid.data());
    if (!meshes_.empty() || !leaf_sizes_.empty()) {
        // meshes and / or octomap leaves: everything becomes occupancy on one grid — the occupancy world's grid when there
        // is one, else the grid of setSdfGrid; primitives are voxelised into it on the host (voxel centre inside), meshes and
        // leaves on the device
        int dims[3];
        double origin[3], voxel;
        if (!occupancy_.empty()) { std::copy(occ_dims_, occ_dims_ + 3, dims); std::copy(occ_origin_, occ_origin_ + 3, origin); voxel = occ_voxel_; }
        else gridGeometry(dims, origin, voxel);
        std::vector<unsigned char> occ;
        if (!occupancy_.empty()) occ = occupancy_;
        if (!obstacles_.empty()) {
            if (occ.empty()) occ.assign((size_t)dims[0] * dims[1] * dims[2], 0);
            for (int z = 0; z < dims[2]; ++z)
                for (int y = 0; y < dims[1]; ++y)
                    for (int x = 0; x < dims[0]; ++x) {
                        unsigned char& v = occ[((size_t)z * dims[1] + y) * dims[0] + x];
                        if (v) continue;
                        const double px = origin[0] + ((double)x + 0.5) * voxel, py = origin[1] + ((double)y + 0.5) * voxel, pz = origin[2] + ((double)z + 0.5) * voxel;
                        for (const auto& o : obstacles_)
                            if (obstacle_distance(o, px, py, pz) < 0.0) { v = 1; break; }
                    }
        }
        // one call per solidity class: solid meshes first (the interior fill sees only them), then the rest on top
        std::vector<double> solid_tris, shell_tris;
        for (const auto& m : meshes_) (m.solid ? solid_tris : shell_tris).insert((m.solid ? solid_tris : shell_tris).end(), m.triangles.begin(), m.triangles.end());
        if (!shell_tris.empty() && !solid_tris.empty()) {
            LOG_WARN_S << "[RobotModel]: shell and solid meshes in one scene: all treated as solid";
            solid_tris.insert(solid_tris.end(), shell_tris.begin(), shell_tris.end());
            shell_tris.clear();
        }
        const std::vector<double>& tris = solid_tris.empty() ? shell_tris : solid_tris;
        return stomp_b200_build_sdf_scene(engine, dims, origin, voxel, (int32_t)(tris.size() / 9), tris.empty() ? nullptr : tris.data(),
                                          solid_tris.empty() ? 0 : 1, (int32_t)leaf_sizes_.size(), leaf_sizes_.empty() ? nullptr : leaf_centres_.data(),
                                          leaf_sizes_.empty() ? nullptr : leaf_sizes_.data(), occ.empty() ? nullptr : occ.data());
    }
    if (!occupancy_.empty()) {
        // primitives present next to an occupancy world are voxelised into it (centre of the voxel inside the primitive)
        std::vector<unsigned char> occ = occupancy_;
        if (!obstacles_.empty())
            for (int z = 0; z < occ_dims_[2]; ++z)
                for (int y = 0; y < occ_dims_[1]; ++y)
                    for (int x = 0; x < occ_dims_[0]; ++x) {
                        unsigned char& v = occ[((size_t)z * occ_dims_[1] + y) * occ_dims_[0] + x];
                        if (v) continue;
                        const double px = occ_origin_[0] + ((double)x + 0.5) * occ_voxel_, py = occ_origin_[1] + ((double)y + 0.5) * occ_voxel_,
                                     pz = occ_origin_[2] + ((double)z + 0.5) * occ_voxel_;
                        for (const auto& o : obstacles_)
                            if (obstacle_distance(o, px, py, pz) < 0.0) { v = 1; break; }
                    }
        return stomp_b200_build_sdf_occupancy(engine, occ_dims_, occ_origin_, occ_voxel_, occ.data());
    }
    int dims[3];
    double origin[3], voxel;
    gridGeometry(dims, origin, voxel);
    const int n = (int)obstacles_.size();
    std::vector<int32_t> kind((size_t)std::max(n, 1));
    std::vector<double> centre(3 * (size_t)std::max(n, 1)), size(3 * (size_t)std::max(n, 1));
    for (int i = 0; i < n; ++i) {
        kind[i] = obstacles_[i].kind;
        for (int a = 0; a < 3; ++a) { centre[3 * i + a] = obstacles_[i].centre[a]; size[3 * i + a] = obstacles_[i].size[a]; }
    }
    return stomp_b200_build_sdf_primitives(engine, dims, origin, voxel, n, kind.data(), centre.data(), size.data());
}

bool RobotModel::ensureValidityEngine()
{
    if (validity_engine_) {
        if (validity_robot_revision_ != robot_revision_) {      // the sphere list changed (grasp object): chain, spheres, scene again
            const int rc = configureEngine(validity_engine_);
            if (rc) { LOG_ERROR_S << "[RobotModel]: " << stomp_b200_last_error(validity_engine_); return false; }
            validity_robot_revision_ = robot_revision_;
            validity_robot_revision_ = robot_revision_;
    validity_revision_ = scene_revision_;
        }
        if (validity_revision_ != scene_revision_) {      // the scene changed since: rebuild the field before answering
            const int rc = configureScene(validity_engine_);
            if (rc) { LOG_ERROR_S << "[RobotModel]: " << stomp_b200_last_error(validity_engine_); return false; }
            validity_revision_ = scene_revision_;
        }
        return true;
    }
    stomp_b200_config cfg;
    stomp_b200_default_config(&cfg);
    cfg.num_time_steps = 2;
    cfg.num_dimensions = (int)chain_.size();
    cfg.min_rollouts = cfg.max_rollouts = cfg.num_rollouts_per_iteration = 1;
    cfg.device = config_.device;
    int rc = stomp_b200_create(&cfg, &validity_engine_);
    if (rc) { LOG_ERROR_S << "[RobotModel]: stomp_b200_create failed: " << stomp_b200_status_string(rc); validity_engine_ = nullptr; return false; }
    rc = configureEngine(validity_engine_);
    if (rc) { LOG_ERROR_S << "[RobotModel]: " << stomp_b200_last_error(validity_engine_); return false; }
    validity_revision_ = scene_revision_;
    return true;
}

// robot_model's isStateValid returns false when the state is in collision; the cost argument is the
// "collision cost" the reference overwrites anyway (OptimizationTask.cpp:192-202)
bool RobotModel::isStateValid(double& collision_cost)
{
    collision_cost = 0.0;
    if (!ensureValidityEngine()) return false;
    uint8_t verdict = 1;
    double cost = 0.0;
    const int rc = stomp_b200_evaluate_states(validity_engine_, joint_state_.data(), 1, 1, &cost, &verdict, nullptr);
    if (rc) { LOG_ERROR_S << "[RobotModel]: " << stomp_b200_last_error(validity_engine_); return false; }
    collision_cost = cost;
    return verdict == 0;
}

}  // namespace robot_model
