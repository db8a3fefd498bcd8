// ORACLE — TEST INFRASTRUCTURE ONLY.  extern "C" surface of the CPU restatement, bound with ctypes by
// tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference).  Never linked into
// or called by the product library.  Parity status: see stomp_oracle.hpp.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "stomp_oracle.hpp"

using namespace oracle;

extern "C" {

struct oracle_config {
    int32_t num_time_steps, num_dimensions;
    int32_t min_rollouts, max_rollouts, num_rollouts_per_iteration, num_iterations;
    double movement_duration, control_cost_weight, min_cost_improvement;
    double noise_stddev[32], noise_decay[32], noise_min_stddev[32];
    int32_t use_noise_adaptation, use_openmp;
    int32_t use_cumulative_costs;   // reference default 1 (PolicyImprovement.cpp:56); 0 per time step; 2 forward cumulation (:473-477)
    int32_t use_projection;         // reference default 0 (PolicyImprovement.cpp:57)
    int32_t per_timestep_minmax;    // 0 = shipped behaviour; 1 = variant commented out at :518-528
    int32_t dense_control_costs;    // 1 = keep the reference's O(N^2) evaluation forms (CPU baseline)
    uint64_t seed;
};

}  // extern "C"

namespace {

struct OracleHandle {
    oracle_config cfg;
    StompConfig sc;
    std::shared_ptr<SphereSdfTask> task;
    std::unique_ptr<Stomp> stomp;
    std::vector<PrimitiveSpec> prims;   // analytic distance field (oracle_set_sdf_primitives)
    // StompPlanner::solve loop state (StompPlanner.cpp:96-141)
    double old_cost = 0, cost_improvement = 0, current_cost = 0;
    int num_iterations = 0;
};

StompConfig to_stomp_config(const oracle_config& c)
{
    StompConfig s;
    s.num_time_steps_ = c.num_time_steps;
    s.num_dimensions_ = c.num_dimensions;
    s.min_rollouts_ = c.min_rollouts;
    s.max_rollouts_ = c.max_rollouts;
    s.num_rollouts_per_iteration_ = c.num_rollouts_per_iteration;
    s.num_iterations_ = c.num_iterations;
    s.movement_duration_ = c.movement_duration;
    s.control_cost_weight_ = c.control_cost_weight;
    s.min_cost_improvement_ = c.min_cost_improvement;
    s.noise_stddev_.assign(c.noise_stddev, c.noise_stddev + c.num_dimensions);
    s.noise_decay_.assign(c.noise_decay, c.noise_decay + c.num_dimensions);
    s.noise_min_stddev_.assign(c.noise_min_stddev, c.noise_min_stddev + c.num_dimensions);
    s.use_noise_adaptation_ = c.use_noise_adaptation != 0;
    s.use_openmp_ = c.use_openmp != 0;
    return s;
}

void apply_switches(OracleHandle* h)
{
    PolicyImprovement& pi = h->stomp->policy_improvement_;
    pi.dense_control_costs_ = h->cfg.dense_control_costs != 0;
    pi.per_timestep_minmax_ = h->cfg.per_timestep_minmax != 0;
    pi.setCostCumulation(h->cfg.use_cumulative_costs != 0);
    pi.forward_cumulation_ = h->cfg.use_cumulative_costs == 2;      // 2 = forward cumulation (cost-to-go)
    if ((h->cfg.use_projection != 0) != pi.use_projection_) {
        pi.use_projection_ = h->cfg.use_projection != 0;
        pi.preComputeProjectionMatrices();
    }
}

}  // namespace

extern "C" {

void* oracle_create(const oracle_config* cfg)
{
    if (!cfg || cfg->num_dimensions <= 0 || cfg->num_dimensions > 32 || cfg->num_time_steps <= 1) return nullptr;
    OracleHandle* h = new OracleHandle();
    h->cfg = *cfg;
    h->sc = to_stomp_config(*cfg);
    h->task.reset(new SphereSdfTask(h->sc));
    h->task->stompInitialize();
    return h;
}

void oracle_destroy(void* hp) { delete static_cast<OracleHandle*>(hp); }

int oracle_set_chain(void* hp, int D, const double* origin_xyz, const double* origin_rpy, const double* axis,
                     const int32_t* parent, const int32_t* prismatic, const double* lower, const double* upper)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (D != h->cfg.num_dimensions) return -1;
    h->task->joints_.assign(D, JointSpec());
    h->task->lower_limits_.assign(lower, lower + D);
    h->task->upper_limits_.assign(upper, upper + D);
    for (int d = 0; d < D; ++d) {
        JointSpec& j = h->task->joints_[d];
        j.parent = parent ? parent[d] : (d == 0 ? -1 : d - 1);
        if (!(j.parent == -1 || j.parent == d - 1)) return -2;
        j.prismatic = prismatic ? prismatic[d] : 0;
        for (int i = 0; i < 3; ++i) { j.o[i] = origin_xyz[3 * d + i]; j.axis[i] = axis[3 * d + i]; }
        const double* rpy = origin_rpy + 3 * d;
        j.fixed_rot_identity = (rpy[0] == 0.0 && rpy[1] == 0.0 && rpy[2] == 0.0) ? 1 : 0;
        rpy_to_matrix(rpy, j.A);
        j.axis_kind = classify_axis(j.axis);
        j.lower = lower[d]; j.upper = upper[d];
    }
    return 0;
}

int oracle_set_spheres(void* hp, int S, const int32_t* link, const double* xyz, const double* radius)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    h->task->spheres_.assign(S, SphereSpec());
    for (int s = 0; s < S; ++s) {
        if (link[s] < 0 || link[s] >= h->cfg.num_dimensions) return -1;
        if (s > 0 && link[s] < link[s - 1]) return -2;   // must be sorted by link
        SphereSpec& sp = h->task->spheres_[s];
        sp.link = link[s];
        for (int i = 0; i < 3; ++i) sp.l[i] = xyz[3 * s + i];
        sp.r = radius[s];
    }
    return 0;
}

// alternative state costs (stomp_oracle.hpp: SphereSdfTask): smooth obstacle cost and / or joint-constraint cost;
// value / tolerance [D] may be NULL when use_joint_constraint == 0
int oracle_set_cost_extras(void* hp, int use_smooth, double smooth_margin, double smooth_weight, int use_joint_constraint,
                           const double* value, const double* tolerance, double jc_weight)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (!h || !h->task) return -1;
    h->task->smooth_cost_ = use_smooth != 0;
    h->task->smooth_margin_ = smooth_margin;
    h->task->smooth_weight_ = smooth_weight;
    h->task->joint_constraint_ = use_joint_constraint != 0;
    h->task->jc_weight_ = jc_weight;
    const int D = h->task->stomp_config_.num_dimensions_;
    if (use_joint_constraint) {
        if (!value || !tolerance) return -1;
        h->task->jc_value_.assign(value, value + D);
        h->task->jc_tolerance_.assign(tolerance, tolerance + D);
    }
    return 0;
}

// sphere pairs checked against each other (self collision); n == 0 switches the check off
int oracle_set_self_collision(void* hp, int n, const int32_t* pairs /*[n][2]*/)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    const int S = (int)h->task->spheres_.size();
    h->task->self_pairs_.clear();
    for (int p = 0; p < n; ++p) {
        int i = pairs[2 * p], j = pairs[2 * p + 1];
        if (i > j) std::swap(i, j);
        if (i < 0 || j >= S || i == j) return -1;
        SelfPairSpec pr;
        pr.i = i; pr.j = j;
        pr.limit2 = self_pair_limit2(h->task->spheres_[i].r, h->task->spheres_[j].r);
        h->task->self_pairs_.push_back(pr);
    }
    return 0;
}

// the grid is NOT copied: the caller keeps it alive for the lifetime of the handle
int oracle_set_sdf(void* hp, const int32_t* dims, const double* origin, double voxel, const float* grid)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    SdfSpec& g = h->task->sdf_;
    g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
    g.ox = origin[0]; g.oy = origin[1]; g.oz = origin[2];
    g.inv_h = 1.0 / voxel;
    g.offx = -(g.ox * g.inv_h); g.offy = -(g.oy * g.inv_h); g.offz = -(g.oz * g.inv_h);
    g.grid = grid;
    g.h = voxel; g.prims = nullptr; g.num_prims = 0;
    return 0;
}

// the distance field of a primitive world, evaluated lazily at the voxel a lookup hits (no grid in memory: the
// >= 2^31-voxel index test); same numbers as oracle_build_sdf_primitives writes
int oracle_set_sdf_primitives(void* hp, const int32_t* dims, const double* origin, double voxel, int n, const int32_t* kind,
                              const double* centre, const double* size)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    h->prims.assign(n, PrimitiveSpec());
    for (int i = 0; i < n; ++i) {
        h->prims[i].kind = kind[i];
        for (int a = 0; a < 3; ++a) { h->prims[i].c[a] = centre[3 * i + a]; h->prims[i].s[a] = size[3 * i + a]; }
    }
    SdfSpec& g = h->task->sdf_;
    g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
    g.ox = origin[0]; g.oy = origin[1]; g.oz = origin[2];
    g.inv_h = 1.0 / voxel;
    g.offx = -(g.ox * g.inv_h); g.offy = -(g.oy * g.inv_h); g.offz = -(g.oz * g.inv_h);
    g.grid = nullptr;
    g.h = voxel;
    g.prims = h->prims.data();
    g.num_prims = n;
    return 0;
}

// exact signed distance of a union of primitives at every voxel centre (kinematics_spec.hpp: primitive_field_value);
// OpenMP over z slabs
int oracle_build_sdf_primitives(const int32_t* dims, const double* origin, double voxel, int n, const int32_t* kind,
                                const double* centre, const double* size, float* out)
{
    std::vector<PrimitiveSpec> prims(n);
    for (int i = 0; i < n; ++i) {
        prims[i].kind = kind[i];
        for (int a = 0; a < 3; ++a) { prims[i].c[a] = centre[3 * i + a]; prims[i].s[a] = size[3 * i + a]; }
    }
    const int nx = dims[0], ny = dims[1], nz = dims[2];
#pragma omp parallel for schedule(dynamic, 1)
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
                out[((size_t)z * ny + y) * nx + x] = primitive_field_value(prims.data(), n, origin[0], origin[1], origin[2], voxel, x, y, z);
    return 0;
}

// ---- meshes and octomap leaves -> occupancy (restates csrc/sdf_builder.cuh: voxelise_triangles_kernel,
// voxelise_leaves_kernel, the boundary flood) ----
static bool triangle_overlaps_box(const double* tri, double cx, double cy, double cz, double hh)
{
    const double v0x = tri[0] - cx, v0y = tri[1] - cy, v0z = tri[2] - cz;
    const double v1x = tri[3] - cx, v1y = tri[4] - cy, v1z = tri[5] - cz;
    const double v2x = tri[6] - cx, v2y = tri[7] - cy, v2z = tri[8] - cz;
    using std::fmin; using std::fmax; using std::fabs;
    if (fmin(fmin(v0x, v1x), v2x) > hh || fmax(fmax(v0x, v1x), v2x) < -hh) return false;
    if (fmin(fmin(v0y, v1y), v2y) > hh || fmax(fmax(v0y, v1y), v2y) < -hh) return false;
    if (fmin(fmin(v0z, v1z), v2z) > hh || fmax(fmax(v0z, v1z), v2z) < -hh) return false;
    const double e0x = v1x - v0x, e0y = v1y - v0y, e0z = v1z - v0z;
    const double e1x = v2x - v1x, e1y = v2y - v1y, e1z = v2z - v1z;
    const double e2x = v0x - v2x, e2y = v0y - v2y, e2z = v0z - v2z;
    const double nx = e0y * e1z - e0z * e1y, ny = e0z * e1x - e0x * e1z, nz = e0x * e1y - e0y * e1x;
    const double dist = (nx * v0x + ny * v0y) + nz * v0z;
    const double rad = hh * ((fabs(nx) + fabs(ny)) + fabs(nz));
    if (dist > rad || dist < -rad) return false;
    auto separated = [&](double ax, double ay, double az) {
        const double p0 = (ax * v0x + ay * v0y) + az * v0z;
        const double p1 = (ax * v1x + ay * v1y) + az * v1z;
        const double p2 = (ax * v2x + ay * v2y) + az * v2z;
        const double r = hh * ((fabs(ax) + fabs(ay)) + fabs(az));
        return fmin(fmin(p0, p1), p2) > r || fmax(fmax(p0, p1), p2) < -r;
    };
    if (separated(0.0, -e0z, e0y) || separated(e0z, 0.0, -e0x) || separated(-e0y, e0x, 0.0)) return false;
    if (separated(0.0, -e1z, e1y) || separated(e1z, 0.0, -e1x) || separated(-e1y, e1x, 0.0)) return false;
    if (separated(0.0, -e2z, e2y) || separated(e2z, 0.0, -e2x) || separated(-e2y, e2x, 0.0)) return false;
    return true;
}

static void voxel_range(double a, double b, double origin, double inv_h, int n, int& lo, int& hi)
{
    const double fa = std::floor((a - origin) * inv_h), fb = std::floor((b - origin) * inv_h);
    lo = (int)std::fmax(fa - 1.0, 0.0);
    hi = (int)std::fmin(fb + 1.0, (double)(n - 1));
    if (!(fb + 1.0 >= 0.0) || !(fa - 1.0 <= (double)(n - 1))) { lo = 1; hi = 0; }
}

// occupancy [nz][ny][nx] of a scene: mesh (conservative, optionally solid), octomap leaves, given occupancy (or NULL)
int oracle_voxelise_scene(const int32_t* dims, const double* origin, double h, int num_triangles, const double* triangles, int solid,
                          int num_leaves, const double* leaf_centres, const double* leaf_sizes, const uint8_t* occupied, uint8_t* occ)
{
    const int nx = dims[0], ny = dims[1], nz = dims[2];
    const size_t count = (size_t)nx * ny * nz;
    const double inv_h = 1.0 / h, hh = (0.5 * h) * 1.000000001;     // inflated by 1e-9: see voxelise_triangles_kernel
    for (size_t i = 0; i < count; ++i) occ[i] = occupied ? occupied[i] : 0;
    for (int tr = 0; tr < num_triangles; ++tr) {
        const double* t = triangles + (size_t)tr * 9;
        int x0, x1, y0, y1, z0, z1;
        voxel_range(std::fmin(std::fmin(t[0], t[3]), t[6]), std::fmax(std::fmax(t[0], t[3]), t[6]), origin[0], inv_h, nx, x0, x1);
        voxel_range(std::fmin(std::fmin(t[1], t[4]), t[7]), std::fmax(std::fmax(t[1], t[4]), t[7]), origin[1], inv_h, ny, y0, y1);
        voxel_range(std::fmin(std::fmin(t[2], t[5]), t[8]), std::fmax(std::fmax(t[2], t[5]), t[8]), origin[2], inv_h, nz, z0, z1);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y)
                for (int x = x0; x <= x1; ++x) {
                    const double cx = origin[0] + ((double)x + 0.5) * h, cy = origin[1] + ((double)y + 0.5) * h, cz = origin[2] + ((double)z + 0.5) * h;
                    if (triangle_overlaps_box(t, cx, cy, cz, hh)) occ[((size_t)z * ny + y) * nx + x] = 1;
                }
    }
    if (num_triangles > 0 && solid) {       // breadth-first flood of the free voxels from the grid's boundary
        std::vector<uint8_t> outside(count, 0);
        std::vector<size_t> queue;
        auto push = [&](int x, int y, int z) {
            const size_t i = ((size_t)z * ny + y) * nx + x;
            if (!occ[i] && !outside[i]) { outside[i] = 1; queue.push_back(i); }
        };
        for (int z = 0; z < nz; ++z)
            for (int y = 0; y < ny; ++y)
                for (int x = 0; x < nx; ++x)
                    if (x == 0 || y == 0 || z == 0 || x == nx - 1 || y == ny - 1 || z == nz - 1) push(x, y, z);
        for (size_t head = 0; head < queue.size(); ++head) {
            const size_t i = queue[head];
            const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((size_t)nx * ny));
            if (x > 0) push(x - 1, y, z);
            if (x < nx - 1) push(x + 1, y, z);
            if (y > 0) push(x, y - 1, z);
            if (y < ny - 1) push(x, y + 1, z);
            if (z > 0) push(x, y, z - 1);
            if (z < nz - 1) push(x, y, z + 1);
        }
        for (size_t i = 0; i < count; ++i) if (!outside[i]) occ[i] = 1;
    }
    for (int lf = 0; lf < num_leaves; ++lf) {
        const double half = 0.5 * leaf_sizes[lf];
        int lo[3], hi[3];
        const int nn[3] = {nx, ny, nz};
        bool empty = false;
        for (int a = 0; a < 3; ++a) {
            const double c = leaf_centres[(size_t)lf * 3 + a];
            const double fl = std::ceil((c - half - origin[a]) * inv_h - 0.5), fh = std::floor((c + half - origin[a]) * inv_h - 0.5);
            lo[a] = (int)std::fmax(fl, 0.0);
            hi[a] = (int)std::fmin(fh, (double)(nn[a] - 1));
            if (!(fh >= 0.0) || !(fl <= (double)(nn[a] - 1))) empty = true;
        }
        if (empty) continue;
        for (int z = lo[2]; z <= hi[2]; ++z)
            for (int y = lo[1]; y <= hi[1]; ++y)
                for (int x = lo[0]; x <= hi[0]; ++x) occ[((size_t)z * ny + y) * nx + x] = 1;
    }
    return 0;
}

// exact Euclidean distance transform of an occupancy grid [nz][ny][nx], signed (positive outside, negative inside),
// centre to centre, in metres: h * sqrt(min squared voxel distance), rounded to binary32.  Brute force per line with the
// same three separable min-plus passes over integer squared distances as the CUDA builder (exact, so the order of the
// candidates does not matter).
int oracle_build_sdf_occupancy(const int32_t* dims, double voxel, const uint8_t* occ, float* out)
{
    const int nx = dims[0], ny = dims[1], nz = dims[2];
    const size_t count = (size_t)nx * ny * nz;
    const int INF = 0x3fffffff;
    auto transform = [&](std::vector<int32_t>& f) {
        std::vector<int32_t> g(count);
        auto pass = [&](const std::vector<int32_t>& in, std::vector<int32_t>& o, int n_line, size_t stride_line, int n_a, size_t stride_a, int n_b, size_t stride_b) {
#pragma omp parallel for collapse(2) schedule(static)
            for (int b = 0; b < n_b; ++b)
                for (int a = 0; a < n_a; ++a) {
                    const size_t base = (size_t)a * stride_a + (size_t)b * stride_b;
                    for (int i = 0; i < n_line; ++i) {
                        int best = INF;
                        for (int j = 0; j < n_line; ++j) {
                            const int v = in[base + (size_t)j * stride_line];
                            if (v >= INF) continue;
                            const int cand = v + (i - j) * (i - j);
                            if (cand < best) best = cand;
                        }
                        o[base + (size_t)i * stride_line] = best;
                    }
                }
        };
        pass(f, g, nx, 1, ny, nx, nz, (size_t)nx * ny);
        pass(g, f, ny, nx, nx, 1, nz, (size_t)nx * ny);
        pass(f, g, nz, (size_t)nx * ny, nx, 1, ny, nx);
        f.swap(g);
    };
    std::vector<int32_t> to_occ(count), to_free(count);
    for (size_t i = 0; i < count; ++i) { to_occ[i] = occ[i] ? 0 : INF; to_free[i] = occ[i] ? INF : 0; }
    transform(to_occ);
    transform(to_free);
    for (size_t i = 0; i < count; ++i) {
        const double d2 = (double)(occ[i] ? to_free[i] : to_occ[i]);
        const double d = voxel * std::sqrt(d2);
        out[i] = (float)(occ[i] ? -d : d);
    }
    return 0;
}

// StompPlanner::setStartGoalTrajectory (StompPlanner.cpp:177-184)
int oracle_set_start_goal(void* hp, const double* start, const double* goal)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    const int D = h->cfg.num_dimensions;
    h->task->updateTrajectory(Vec(start, start + D), Vec(goal, goal + D));
    h->task->input_initial_trajectory_ = h->task->initial_trajectory_;
    h->task->createPolicy();
    return 0;
}

// StompPlanner::updateInitialTrajectory (StompPlanner.cpp:186-208); trajectory is [D][T]
int oracle_set_initial_trajectory(void* hp, const double* traj)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    const int D = h->cfg.num_dimensions, T = h->cfg.num_time_steps, P = TRAJECTORY_PADDING;
    for (int d = 0; d < D; ++d) {
        for (int i = 0; i < P; ++i) {
            h->task->initial_trajectory_[d][i] = traj[(size_t)d * T] * 1.0;
            h->task->initial_trajectory_[d][P + T + i] = traj[(size_t)d * T + T - 1] * 1.0;
        }
        for (int i = 0; i < T; ++i) h->task->initial_trajectory_[d][P + i] = traj[(size_t)d * T + i];
    }
    h->task->input_initial_trajectory_ = h->task->initial_trajectory_;
    h->task->updatePolicy();
    return 0;
}

// host-side policy products (a14): any output may be null.  R, Rinv, L: [T][T]; params_all: [D][N];
// mincc: [D][T]; linear: [D][T]
int oracle_get_policy(void* hp, double* R, double* Rinv, double* L, double* params_all, double* mincc, double* linear)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (!h->task->policy_) return -1;
    const CovariantMovementPrimitive& p = *h->task->policy_;
    const int T = p.num_vars_free_, N = p.num_vars_all_, D = p.num_dimensions_;
    if (R) std::memcpy(R, p.control_costs_[0].a.data(), sizeof(double) * T * T);
    if (Rinv) std::memcpy(Rinv, p.inv_control_costs_[0].a.data(), sizeof(double) * T * T);
    if (L) { Mat Lm = llt_lower(p.inv_control_costs_[0]); std::memcpy(L, Lm.a.data(), sizeof(double) * T * T); }
    for (int d = 0; d < D; ++d) {
        if (params_all) std::memcpy(params_all + (size_t)d * N, p.parameters_all_[d].data(), sizeof(double) * N);
        if (mincc) std::memcpy(mincc + (size_t)d * T, p.min_control_cost_parameters_free_[d].data(), sizeof(double) * T);
        if (linear) std::memcpy(linear + (size_t)d * T, p.linear_control_costs_[d].data(), sizeof(double) * T);
    }
    return 0;
}

double oracle_get_movement_dt(void* hp)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    return h->task->policy_ ? h->task->policy_->movement_dt_ : 0.0;
}

// start of StompPlanner::solve (StompPlanner.cpp:65-73,96-99): a new Stomp per solve
int oracle_begin_solve(void* hp)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (!h->task->policy_) return -1;
    h->stomp.reset(new Stomp());
    h->stomp->initialize(h->sc, h->task, h->cfg.seed);
    apply_switches(h);
    h->old_cost = 0; h->cost_improvement = 0; h->current_cost = 0; h->num_iterations = 0;
    return 0;
}

// replace the Cholesky factor used for sampling (parity tests share one L between oracle and GPU)
int oracle_set_cholesky(void* hp, const double* L)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (!h->stomp) return -1;
    const int T = h->cfg.num_time_steps;
    for (auto& g : h->stomp->policy_improvement_.noise_generators_)
        std::memcpy(g.covariance_cholesky_.a.data(), L, sizeof(double) * T * T);
    return 0;
}

// one pass of the loop body of StompPlanner::solve (StompPlanner.cpp:101-118).  Returns 1 when the stop
// criterion fires, 0 otherwise.  noise / epsilon: [num_rollouts_gen][D][T] or null.
int oracle_iterate(void* hp, int iteration, const double* injected_noise, const double* epsilon)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (!h->stomp) return -1;
    NoiseSource src; src.injected = injected_noise; src.epsilon = epsilon;
    h->num_iterations++;
    h->stomp->runSingleIteration(iteration, src);
    h->current_cost = h->stomp->getNoiselessRolloutTotalCost();
    h->cost_improvement = h->current_cost - h->old_cost;
    h->old_cost = h->current_cost;
    if ((h->current_cost < 1) && (std::fabs(h->cost_improvement) < h->sc.min_cost_improvement_)) return 1;
    return 0;
}

// end of StompPlanner::solve (:148-173): solution = last parameters; returns 1 = PATH_FOUND, 0 = NO_PATH_FOUND
int oracle_finish_solve(void* hp, double* solution /*[D][T]*/, int32_t* iterations_used)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    const int D = h->cfg.num_dimensions, T = h->cfg.num_time_steps;
    const CovariantMovementPrimitive& p = *h->task->policy_;
    if (solution)
        for (int d = 0; d < D; ++d)
            for (int i = 0; i < T; ++i) solution[(size_t)d * T + i] = p.parameters_all_[d][i + (DIFF_RULE_LENGTH - 1)];
    if (iterations_used) *iterations_used = h->num_iterations;
    h->stomp.reset();
    return ((h->current_cost < 1) && (std::fabs(h->cost_improvement) <= h->sc.min_cost_improvement_)) ? 1 : 0;
}

// whole StompPlanner::solve with the internal generator; returns status as oracle_finish_solve.
// seconds_out = wall time of the iteration loop only (where the reference times: MotionPlanners.cpp:506-512
// brackets solve(); the one-time Stomp::initialize is reported separately in setup_seconds_out)
int oracle_solve(void* hp, int max_iterations, int honour_stop, double* solution, int32_t* iterations_used,
                 double* seconds_out, double* setup_seconds_out)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    auto t0 = std::chrono::steady_clock::now();
    if (oracle_begin_solve(hp) != 0) return -1;
    auto t1 = std::chrono::steady_clock::now();
    for (int i = 0; i < max_iterations; ++i) {
        int stop = oracle_iterate(hp, i, nullptr, nullptr);
        if (stop && honour_stop) break;
    }
    auto t2 = std::chrono::steady_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t2 - t1).count();
    if (setup_seconds_out) *setup_seconds_out = std::chrono::duration<double>(t1 - t0).count();
    (void)h;
    return oracle_finish_solve(hp, solution, iterations_used);
}

// the sphere / SDF task of a handle, for oracle/ref/ref_driver.cpp (which evaluates the reference's own loop on it)
void* oracle_task(void* hp) { return static_cast<OracleHandle*>(hp)->task.get(); }
const void* oracle_config_of(void* hp) { return &static_cast<OracleHandle*>(hp)->cfg; }

int oracle_num_rollouts(void* hp, int32_t* num_rollouts, int32_t* num_rollouts_gen)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (!h->stomp) return -1;
    *num_rollouts = h->stomp->policy_improvement_.num_rollouts_;
    *num_rollouts_gen = h->stomp->policy_improvement_.num_rollouts_gen_;
    return 0;
}

// field ids: 0 parameters_noise [n][D][T]; 1 noise [n][D][T]; 2 control_costs [n][D][T];
// 3 probabilities [n][D][T]; 4 cumulative_costs [n][D][T]; 5 total_costs [n][D][T];
// 6 state_costs [n][T]; 7 full_probabilities [n][D]; 8 full_costs [n][D]; 9 total_cost [n];
// 10 parameters_noise_projected [n][D][T]; 11 noise_projected [n][D][T]
int oracle_get_rollout_field(void* hp, int field, double* out)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (!h->stomp) return -1;
    const PolicyImprovement& pi = h->stomp->policy_improvement_;
    const int n = pi.num_rollouts_, D = pi.num_dimensions_, T = pi.num_time_steps_;
    for (int r = 0; r < n; ++r) {
        const Rollout& ro = pi.rollouts_[r];
        const std::vector<Vec>* f = nullptr;
        switch (field) {
            case 0: f = &ro.parameters_noise_; break;
            case 1: f = &ro.noise_; break;
            case 2: f = &ro.control_costs_; break;
            case 3: f = &ro.probabilities_; break;
            case 4: f = &ro.cumulative_costs_; break;
            case 5: f = &ro.total_costs_; break;
            case 10: f = &ro.parameters_noise_projected_; break;
            case 11: f = &ro.noise_projected_; break;
            case 6: std::memcpy(out + (size_t)r * T, ro.state_costs_.data(), sizeof(double) * T); continue;
            case 7: std::memcpy(out + (size_t)r * D, ro.full_probabilities_.data(), sizeof(double) * D); continue;
            case 8: std::memcpy(out + (size_t)r * D, ro.full_costs_.data(), sizeof(double) * D); continue;
            case 9: out[r] = ro.total_cost_; continue;
            default: return -2;
        }
        for (int d = 0; d < D; ++d) std::memcpy(out + ((size_t)r * D + d) * T, (*f)[d].data(), sizeof(double) * T);
    }
    return 0;
}

int oracle_get_rollout_validity(void* hp, uint8_t* out /*[num_rollouts_gen]*/)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (!h->stomp) return -1;
    std::memcpy(out, h->stomp->rollout_validity_.data(), h->stomp->rollout_validity_.size());
    return 0;
}

int oracle_get_updates(void* hp, double* out /*[D][T]*/)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (!h->stomp) return -1;
    const PolicyImprovement& pi = h->stomp->policy_improvement_;
    for (int d = 0; d < pi.num_dimensions_; ++d)
        for (int t = 0; t < pi.num_time_steps_; ++t) out[(size_t)d * pi.num_time_steps_ + t] = pi.parameter_updates_[d](0, t);
    return 0;
}

int oracle_get_parameters(void* hp, double* out /*[D][T]*/)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    std::vector<Vec> p;
    h->task->policy_->getParameters(p);
    for (size_t d = 0; d < p.size(); ++d) std::memcpy(out + d * p[d].size(), p[d].data(), sizeof(double) * p[d].size());
    return 0;
}

int oracle_get_stddevs(void* hp, double* out /*[D]*/)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (!h->stomp) return -1;
    const Vec& s = h->stomp->policy_improvement_.adapted_stddevs_;
    std::memcpy(out, s.data(), sizeof(double) * s.size());
    return 0;
}

int oracle_get_noiseless(void* hp, double* total_cost, int32_t* valid, double* state_costs /*[T] or null*/,
                         double* control_costs /*[D][T] or null*/, double* best_cost /* or null */)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (!h->stomp) return -1;
    const PolicyImprovement& pi = h->stomp->policy_improvement_;
    *total_cost = pi.noiseless_rollout_.total_cost_;
    *valid = h->stomp->last_noiseless_rollout_valid_ ? 1 : 0;
    const int T = pi.num_time_steps_;
    if (state_costs) std::memcpy(state_costs, pi.noiseless_rollout_.state_costs_.data(), sizeof(double) * T);
    if (control_costs)
        for (int d = 0; d < pi.num_dimensions_; ++d)
            std::memcpy(control_costs + (size_t)d * T, pi.noiseless_rollout_.control_costs_[d].data(), sizeof(double) * T);
    if (best_cost) *best_cost = h->stomp->best_noiseless_cost_;
    return 0;
}

// ---- kernel-level checkers -----------------------------------------------------------------------

void oracle_sincos(double x, double* s, double* c) { det_sincos(x, s, c); }

int oracle_sphere_centres(void* hp, const double* q, double* centres /*[S][3]*/)
{
    static_cast<OracleHandle*>(hp)->task->sphereCentres(q, centres);
    return 0;
}

// theta: [K][D][T]; costs: [K][T]; verdict: [K][T] (1 = in collision); validity: [K]
int oracle_state_costs(void* hp, const double* theta, int K, double* costs, uint8_t* verdict, uint8_t* validity, int threads)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    const int D = h->cfg.num_dimensions, T = h->cfg.num_time_steps;
    (void)threads;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int k = 0; k < K; ++k) {
        std::vector<double> q(D);
        for (int t = 0; t < T; ++t) {
            for (int d = 0; d < D; ++d) q[d] = theta[((size_t)k * D + d) * T + t];
            bool hit = h->task->stateCollides(q.data());
            if (costs) {      // SphereSdfTask::execute's cost of the state, alternative costs included
                double c = hit ? 1.0 : 0.0;
                if (h->task->smooth_cost_) c = h->task->statePenetration(q.data());
                if (h->task->joint_constraint_) c += h->task->jointConstraintCost(q.data());
                costs[(size_t)k * T + t] = c;
            }
            if (verdict) verdict[(size_t)k * T + t] = hit ? 1 : 0;
            if (validity && t == T - 1) validity[k] = hit ? 0 : 1;
        }
    }
    return 0;
}

// control costs of trajectories x = parameters + noise_projected, [K][D][T] each -> out [K][D][T]
int oracle_control_costs(void* hp, const double* parameters /*[D][T]*/, const double* noise /*[K][D][T]*/, int K,
                         double weight, double* out, int dense_form)
{
    OracleHandle* h = static_cast<OracleHandle*>(hp);
    if (!h->task->policy_) return -1;
    const int D = h->cfg.num_dimensions, T = h->cfg.num_time_steps;
    std::vector<Vec> p(D), n(D), cc;
    for (int d = 0; d < D; ++d) p[d].assign(parameters + (size_t)d * T, parameters + (size_t)(d + 1) * T);
    for (int k = 0; k < K; ++k) {
        for (int d = 0; d < D; ++d) n[d].assign(noise + ((size_t)k * D + d) * T, noise + ((size_t)k * D + d + 1) * T);
        h->task->policy_->computeControlCosts(p, n, weight, cc, dense_form != 0);
        for (int d = 0; d < D; ++d) std::memcpy(out + ((size_t)k * D + d) * T, cc[d].data(), sizeof(double) * T);
    }
    return 0;
}

// standalone factorisations (host-logic tests of the product's own LU / LLT)
int oracle_full_piv_lu_inverse(const double* A, int n, double* out)
{
    Mat m(n, n);
    std::memcpy(m.a.data(), A, sizeof(double) * n * n);
    Mat inv = full_piv_lu_inverse(m);
    std::memcpy(out, inv.a.data(), sizeof(double) * n * n);
    return 0;
}

int oracle_llt_lower(const double* A, int n, double* out)
{
    Mat m(n, n);
    std::memcpy(m.a.data(), A, sizeof(double) * n * n);
    Mat L = llt_lower(m);
    std::memcpy(out, L.a.data(), sizeof(double) * n * n);
    return 0;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// thread count of every later parallel region of this process (torchrun exports OMP_NUM_THREADS=1 to its workers; the
// CPU arm of bench.py sets the count it reports explicitly); returns the count in effect
int oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

}  // extern "C"
