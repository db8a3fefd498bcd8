// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the
// product path (motion_planners_b200/, include/); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it, and only as the checker / reported CPU baseline.
//
// PARITY STATUS OF THIS FILE — the one part of the path that stays UNPINNED: the forward kinematics and the
// collision verdict of the reference live in un-vendored, un-pinned third-party code (robot_model -> KDL / FCL;
// call sites src/planners/src/wrappers/stomp/OptimizationTask.cpp:190,192), absent from this image, and the
// reference ships no golden vectors for them.  (Everything else — the STOMP loop itself — is pinned against the
// reference's own compiled code: oracle/ref/.)  This header therefore *defines* the link-sphere-vs-SDF task named by
// BASELINE.json's north_star; what is inherited from the reference is the semantics of the result:
// cost 1.0 / 0.0 per timestep and validity == verdict of the LAST timestep
// (OptimizationTask.cpp:183-204).
//
// The arithmetic below is a *specification*: every operation is one IEEE-754 binary64 operation
// (add, mul, fma, rint, compare) in a fixed order, so that a second implementation that issues the
// same operations (the CUDA kernels, compiled with -fmad=false and explicit fma()) produces
// bit-identical sphere centres and therefore bit-identical verdicts.  Compile with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstddef>

namespace oracle {

// ---------------------------------------------------------------------------------------------
// Deterministic sin/cos.  3-term Cody-Waite reduction by pi/2 followed by the classic degree-13 /
// degree-14 minimax polynomials on [-pi/4, pi/4] (coefficients: Sun fdlibm k_sin.c / k_cos.c, which
// are mathematical constants).  Max error ~1 ulp for |x| < 1e5, ample for the 1e-9 budget.
// ---------------------------------------------------------------------------------------------
static inline double spec_fma(double a, double b, double c) { return __builtin_fma(a, b, c); }

static inline void det_sincos(double x, double* s_out, double* c_out)
{
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    const double PIO2_1 = 1.57079632673412561417e+00;  // first 33 bits of pi/2
    const double PIO2_2 = 6.07710050630396597660e-11;  // next 33 bits
    const double PIO2_3 = 2.02226624871116645580e-21;  // next 33 bits
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;

    double k = std::rint(x * TWO_OVER_PI);
    double nk = -k;
    double r = spec_fma(nk, PIO2_1, x);
    r = spec_fma(nk, PIO2_2, r);
    r = spec_fma(nk, PIO2_3, r);
    double z = r * r;

    double ps = spec_fma(z, S6, S5);
    ps = spec_fma(z, ps, S4);
    ps = spec_fma(z, ps, S3);
    ps = spec_fma(z, ps, S2);
    ps = spec_fma(z, ps, S1);
    double rz = r * z;
    double sr = spec_fma(rz, ps, r);

    double pc = spec_fma(z, C6, C5);
    pc = spec_fma(z, pc, C4);
    pc = spec_fma(z, pc, C3);
    pc = spec_fma(z, pc, C2);
    pc = spec_fma(z, pc, C1);
    double zz = z * z;
    double half = spec_fma(z, -0.5, 1.0);
    double cr = spec_fma(zz, pc, half);

    long long q = (long long)k;
    switch (q & 3) {
        case 0: *s_out = sr;  *c_out = cr;  break;
        case 1: *s_out = cr;  *c_out = -sr; break;
        case 2: *s_out = -sr; *c_out = -cr; break;
        default: *s_out = -cr; *c_out = sr; break;
    }
}

// ---------------------------------------------------------------------------------------------
// Kinematic chain description (what robot_model::RobotModel holds for the planning group; URDF
// joint elements: reference test/data/kuka_iiwa.urdf:162-210).
// ---------------------------------------------------------------------------------------------
enum AxisKind : int32_t {
    AXIS_X = 0, AXIS_Y = 1, AXIS_Z = 2, AXIS_NEG_X = 3, AXIS_NEG_Y = 4, AXIS_NEG_Z = 5, AXIS_GENERAL = 6
};

struct JointSpec {
    int32_t parent;      // -1: this joint hangs off the fixed base frame (chain restart), else d-1
    int32_t axis_kind;   // AxisKind
    int32_t fixed_rot_identity;  // 1 when rpy == 0 (skip R*A)
    int32_t prismatic;   // 0 revolute, 1 prismatic
    double o[3];         // origin xyz in the parent link frame
    double A[9];         // fixed rotation from rpy, row major: Rz(yaw)*Ry(pitch)*Rx(roll)
    double axis[3];      // unit axis (used by AXIS_GENERAL and prismatic)
    double lower, upper; // joint limits
};

struct SphereSpec {
    int32_t link;        // index of the joint whose child link carries the sphere
    double l[3];         // centre in the link frame
    double r;            // radius
};

// one world primitive (what the reference hands to FCL through handleCollisionObjectInWorld,
// src/MotionPlanners.cpp:416-460): kind 0 sphere (s[0] = radius), 1 box (s = half extents)
struct PrimitiveSpec {
    int32_t kind;
    double c[3];
    double s[3];
};

struct SdfSpec {
    int32_t nx, ny, nz;  // x fastest
    double ox, oy, oz;   // world position of the min corner of voxel (0,0,0)
    double inv_h;        // 1 / voxel size
    double offx, offy, offz;   // -(origin * inv_h): voxel coordinate = fma(c, inv_h, off)
    const float* grid;   // null: the field is evaluated from the primitives at the voxel centre on every lookup
    double h;                    // voxel size (analytic mode)
    const PrimitiveSpec* prims;  // analytic mode
    int32_t num_prims;
};

// SPEC of the primitive distance field (op order normative; the CUDA builder issues the same operations):
// signed distance from p to one primitive, negative inside
static inline double primitive_distance(const PrimitiveSpec& pr, double px, double py, double pz)
{
    const double dx = px - pr.c[0], dy = py - pr.c[1], dz = pz - pr.c[2];
    if (pr.kind == 0) return std::sqrt((dx * dx + dy * dy) + dz * dz) - pr.s[0];
    if (pr.kind == 2) {   // cylinder along z: s[0] radius, s[1] half height
        const double qr = std::sqrt(dx * dx + dy * dy) - pr.s[0], qh = std::fabs(dz) - pr.s[1];
        const double orr = std::fmax(qr, 0.0), oh = std::fmax(qh, 0.0);
        return std::sqrt(orr * orr + oh * oh) + std::fmin(std::fmax(qr, qh), 0.0);
    }
    const double qx = std::fabs(dx) - pr.s[0], qy = std::fabs(dy) - pr.s[1], qz = std::fabs(dz) - pr.s[2];
    const double ox = std::fmax(qx, 0.0), oy = std::fmax(qy, 0.0), oz = std::fmax(qz, 0.0);
    const double outside = std::sqrt((ox * ox + oy * oy) + oz * oz);
    const double inside = std::fmin(std::fmax(std::fmax(qx, qy), qz), 0.0);
    return outside + inside;
}

// value of voxel (ix, iy, iz): (float) min_i d_i(origin + (i + 0.5) * h), primitives in list order
static inline float primitive_field_value(const PrimitiveSpec* prims, int n, double ox, double oy, double oz, double h,
                                          int ix, int iy, int iz)
{
    const double px = ox + ((double)ix + 0.5) * h, py = oy + ((double)iy + 0.5) * h, pz = oz + ((double)iz + 0.5) * h;
    double d = INFINITY;
    for (int i = 0; i < n; ++i) d = std::fmin(d, primitive_distance(prims[i], px, py, pz));
    return (float)d;
}

struct Frame { double R[9]; double p[3]; };

static inline void frame_identity(Frame& f)
{
    f.R[0] = 1; f.R[1] = 0; f.R[2] = 0;
    f.R[3] = 0; f.R[4] = 1; f.R[5] = 0;
    f.R[6] = 0; f.R[7] = 0; f.R[8] = 1;
    f.p[0] = f.p[1] = f.p[2] = 0;
}

// classify a URDF axis vector; fixed rotation matrix from rpy (host-side setup, libm is fine here
// because both the oracle and the product receive the *resulting doubles* through their APIs).
static inline int32_t classify_axis(const double a[3])
{
    const double e = 0.0;
    if (a[1] == e && a[2] == e && a[0] == 1.0) return AXIS_X;
    if (a[0] == e && a[2] == e && a[1] == 1.0) return AXIS_Y;
    if (a[0] == e && a[1] == e && a[2] == 1.0) return AXIS_Z;
    if (a[1] == e && a[2] == e && a[0] == -1.0) return AXIS_NEG_X;
    if (a[0] == e && a[2] == e && a[1] == -1.0) return AXIS_NEG_Y;
    if (a[0] == e && a[1] == e && a[2] == -1.0) return AXIS_NEG_Z;
    return AXIS_GENERAL;
}

static inline void rpy_to_matrix(const double rpy[3], double A[9])
{
    double sr, cr, sp, cp, sy, cy;
    det_sincos(rpy[0], &sr, &cr);
    det_sincos(rpy[1], &sp, &cp);
    det_sincos(rpy[2], &sy, &cy);
    A[0] = cy * cp; A[1] = cy * sp * sr - sy * cr; A[2] = cy * sp * cr + sy * sr;
    A[3] = sy * cp; A[4] = sy * sp * sr + cy * cr; A[5] = sy * sp * cr - cy * sr;
    A[6] = -sp;     A[7] = cp * sr;                A[8] = cp * cr;
}

// advance the running frame through joint j at joint value q  (SPEC: op order is normative)
static inline void apply_joint(Frame& f, const JointSpec& j, double q)
{
    if (j.parent < 0) frame_identity(f);
    double* R = f.R; double* p = f.p;
    // p += R * o
    p[0] = spec_fma(R[2], j.o[2], spec_fma(R[1], j.o[1], spec_fma(R[0], j.o[0], p[0])));
    p[1] = spec_fma(R[5], j.o[2], spec_fma(R[4], j.o[1], spec_fma(R[3], j.o[0], p[1])));
    p[2] = spec_fma(R[8], j.o[2], spec_fma(R[7], j.o[1], spec_fma(R[6], j.o[0], p[2])));
    // R = R * A
    if (!j.fixed_rot_identity) {
        double N[9];
        for (int i = 0; i < 3; ++i)
            for (int c = 0; c < 3; ++c)
                N[3 * i + c] = spec_fma(R[3 * i + 2], j.A[6 + c],
                               spec_fma(R[3 * i + 1], j.A[3 + c], R[3 * i] * j.A[c]));
        for (int i = 0; i < 9; ++i) R[i] = N[i];
    }
    if (j.prismatic) {
        // p += q * (R * axis)
        for (int i = 0; i < 3; ++i) {
            double d = spec_fma(R[3 * i + 2], j.axis[2], spec_fma(R[3 * i + 1], j.axis[1], R[3 * i] * j.axis[0]));
            p[i] = spec_fma(q, d, p[i]);
        }
        return;
    }
    double s, c;
    det_sincos(q, &s, &c);
    int kind = j.axis_kind;
    if (kind == AXIS_NEG_X || kind == AXIS_NEG_Y || kind == AXIS_NEG_Z) { s = -s; kind -= 3; }
    double ns = -s;
    if (kind == AXIS_Z) {            // col0' = c*col0 + s*col1 ; col1' = c*col1 - s*col0
        for (int i = 0; i < 3; ++i) {
            double a = R[3 * i], b = R[3 * i + 1];
            R[3 * i]     = spec_fma(s, b, c * a);
            R[3 * i + 1] = spec_fma(ns, a, c * b);
        }
    } else if (kind == AXIS_Y) {     // col0' = c*col0 - s*col2 ; col2' = c*col2 + s*col0
        for (int i = 0; i < 3; ++i) {
            double a = R[3 * i], b = R[3 * i + 2];
            R[3 * i]     = spec_fma(ns, b, c * a);
            R[3 * i + 2] = spec_fma(s, a, c * b);
        }
    } else if (kind == AXIS_X) {     // col1' = c*col1 + s*col2 ; col2' = c*col2 - s*col1
        for (int i = 0; i < 3; ++i) {
            double a = R[3 * i + 1], b = R[3 * i + 2];
            R[3 * i + 1] = spec_fma(s, b, c * a);
            R[3 * i + 2] = spec_fma(ns, a, c * b);
        }
    } else {                          // Rodrigues: Q = c*I + s*[a]x + (1-c)*a a^T ; R = R*Q
        const double ax = j.axis[0], ay = j.axis[1], az = j.axis[2];
        double v = 1.0 - c;
        double Q[9];
        Q[0] = spec_fma(v * ax, ax, c);        Q[1] = spec_fma(v * ax, ay, -(s * az)); Q[2] = spec_fma(v * ax, az, s * ay);
        Q[3] = spec_fma(v * ay, ax, s * az);   Q[4] = spec_fma(v * ay, ay, c);         Q[5] = spec_fma(v * ay, az, -(s * ax));
        Q[6] = spec_fma(v * az, ax, -(s * ay)); Q[7] = spec_fma(v * az, ay, s * ax);   Q[8] = spec_fma(v * az, az, c);
        double N[9];
        for (int i = 0; i < 3; ++i)
            for (int cc = 0; cc < 3; ++cc)
                N[3 * i + cc] = spec_fma(R[3 * i + 2], Q[6 + cc],
                                spec_fma(R[3 * i + 1], Q[3 + cc], R[3 * i] * Q[cc]));
        for (int i = 0; i < 9; ++i) R[i] = N[i];
    }
}

static inline void sphere_centre(const Frame& f, const SphereSpec& sp, double c[3])
{
    const double* R = f.R;
    c[0] = spec_fma(R[2], sp.l[2], spec_fma(R[1], sp.l[1], spec_fma(R[0], sp.l[0], f.p[0])));
    c[1] = spec_fma(R[5], sp.l[2], spec_fma(R[4], sp.l[1], spec_fma(R[3], sp.l[0], f.p[1])));
    c[2] = spec_fma(R[8], sp.l[2], spec_fma(R[7], sp.l[1], spec_fma(R[6], sp.l[0], f.p[2])));
}

// nearest-voxel SDF lookup; coordinates clamped to the grid
static inline size_t sdf_index(const SdfSpec& g, const double c[3])
{
    double fx = spec_fma(c[0], g.inv_h, g.offx);
    double fy = spec_fma(c[1], g.inv_h, g.offy);
    double fz = spec_fma(c[2], g.inv_h, g.offz);
    fx = std::fmin(std::fmax(fx, 0.0), (double)(g.nx - 1));
    fy = std::fmin(std::fmax(fy, 0.0), (double)(g.ny - 1));
    fz = std::fmin(std::fmax(fz, 0.0), (double)(g.nz - 1));
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    return ((size_t)iz * (size_t)g.ny + (size_t)iy) * (size_t)g.nx + (size_t)ix;
}

static inline bool sphere_collides(const SdfSpec& g, const double c[3], double radius)
{
    const size_t idx = sdf_index(g, c);
    double d;
    if (g.grid) {
        d = (double)g.grid[idx];
    } else {   // analytic mode: the number the built grid holds at this voxel
        const size_t plane = (size_t)g.nx * (size_t)g.ny;
        const int iz = (int)(idx / plane), iy = (int)((idx - (size_t)iz * plane) / (size_t)g.nx);
        const int ix = (int)(idx - (size_t)iz * plane - (size_t)iy * (size_t)g.nx);
        d = (double)primitive_field_value(g.prims, g.num_prims, g.ox, g.oy, g.oz, g.h, ix, iy, iz);
    }
    return (d - radius) < 0.0;
}

// distance the grid holds under a centre, as a double (the number sphere_collides compares with the radius)
static inline double sphere_distance(const SdfSpec& g, const double c[3])
{
    const size_t idx = sdf_index(g, c);
    if (g.grid) return (double)g.grid[idx];
    const size_t plane = (size_t)g.nx * (size_t)g.ny;
    const int iz = (int)(idx / plane), iy = (int)((idx - (size_t)iz * plane) / (size_t)g.nx);
    const int ix = (int)(idx - (size_t)iz * plane - (size_t)iy * (size_t)g.nx);
    return (double)primitive_field_value(g.prims, g.num_prims, g.ox, g.oy, g.oz, g.h, ix, iy, iz);
}

// ---------------------------------------------------------------------------------------------
// Self collision (the "self" half of robot_model's isStateValid; the reference's SRDF lists the link
// pairs that are NOT checked: test/data/kuka_iiwa.srdf:46-70).  A pair of link spheres (i, j) collides
// when the squared distance of the centres is below (r_i + r_j)^2.  SPEC: op order is normative.
// ---------------------------------------------------------------------------------------------
struct SelfPairSpec {
    int32_t i, j;        // sphere indices, i < j
    double limit2;       // (r_i + r_j)^2, computed once on the host as s = r_i + r_j; s * s
};

static inline double self_pair_limit2(double ri, double rj)
{
    double s = ri + rj;
    return s * s;
}

static inline bool self_pair_collides(const double* centres /*[S][3]*/, const SelfPairSpec& pr)
{
    const double* a = centres + 3 * pr.i;
    const double* b = centres + 3 * pr.j;
    double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    double d2 = spec_fma(dz, dz, spec_fma(dy, dy, dx * dx));
    return d2 < pr.limit2;
}

}  // namespace oracle
