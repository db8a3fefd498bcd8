// STL loading and sphere fitting (include/robot_model/MeshTools.hpp).
#include <robot_model/MeshTools.hpp>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>

namespace robot_model {

namespace {

void push_vertex(std::vector<double>& t, double x, double y, double z, const double* scale, const double* tr)
{
    const double v[3] = {x, y, z};
    for (int a = 0; a < 3; ++a) t.push_back(v[a] * (scale ? scale[a] : 1.0) + (tr ? tr[a] : 0.0));
}

}  // namespace

bool loadStl(const std::string& path, std::vector<double>& triangles, const double scale[3], const double translation[3])
{
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    std::vector<char> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (data.size() < 15) return false;
    // binary: 80-byte header, uint32 count, 50 bytes per facet — recognised by its exact size (an ASCII file can also
    // start with "solid", and so do many binary ones)
    if (data.size() >= 84) {
        uint32_t n = 0;
        std::memcpy(&n, data.data() + 80, 4);
        if (data.size() == 84 + (size_t)n * 50) {
            for (uint32_t i = 0; i < n; ++i) {
                float v[12];
                std::memcpy(v, data.data() + 84 + (size_t)i * 50, 48);
                for (int k = 0; k < 3; ++k) push_vertex(triangles, v[3 + 3 * k], v[4 + 3 * k], v[5 + 3 * k], scale, translation);
            }
            return n > 0;
        }
    }
    std::istringstream in(std::string(data.begin(), data.end()));
    std::string word;
    size_t before = triangles.size();
    while (in >> word)
        if (word == "vertex") {
            double x, y, z;
            if (!(in >> x >> y >> z)) return false;
            push_vertex(triangles, x, y, z, scale, translation);
        }
    const size_t added = triangles.size() - before;
    if (added == 0 || added % 9 != 0) { triangles.resize(before); return false; }
    return true;
}

std::vector<FittedSphere> fitSpheres(const std::vector<double>& triangles, int max_spheres, double padding)
{
    std::vector<FittedSphere> out;
    const size_t nv = triangles.size() / 3;
    if (nv == 0 || max_spheres < 1) return out;
    double lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = std::numeric_limits<double>::max(); hi[a] = -lo[a]; }
    for (size_t v = 0; v < nv; ++v)
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], triangles[3 * v + a]); hi[a] = std::max(hi[a], triangles[3 * v + a]); }
    int axis = 0;
    for (int a = 1; a < 3; ++a) if (hi[a] - lo[a] > hi[axis] - lo[axis]) axis = a;
    const int b = (axis + 1) % 3, c = (axis + 2) % 3;
    const double length = hi[axis] - lo[axis];
    const double cross = std::sqrt((hi[b] - lo[b]) * (hi[b] - lo[b]) + (hi[c] - lo[c]) * (hi[c] - lo[c]));   // cross-section diameter
    int slabs = cross > 0.0 ? (int)std::ceil(length / cross) : max_spheres;
    slabs = std::max(1, std::min(slabs, max_spheres));
    const double step = slabs > 0 ? length / slabs : 0.0;
    // points the spheres have to hold: the vertices and, for triangles longer than half a slab, a barycentric lattice on
    // the triangle (a box meshed with twelve triangles has no vertex in its middle slabs)
    std::vector<double> pts(triangles.begin(), triangles.begin() + 3 * nv);
    const double spacing = std::max(0.5 * step, 1e-6);
    for (size_t tr = 0; tr + 8 < triangles.size(); tr += 9) {
        const double* p0 = &triangles[tr]; const double* p1 = p0 + 3; const double* p2 = p0 + 6;
        double longest = 0.0;
        const double* e[3][2] = {{p0, p1}, {p1, p2}, {p2, p0}};
        for (const auto& ed : e) { double d2 = 0; for (int a = 0; a < 3; ++a) d2 += (ed[0][a] - ed[1][a]) * (ed[0][a] - ed[1][a]); longest = std::max(longest, std::sqrt(d2)); }
        const int n = (int)std::min(64.0, std::ceil(longest / spacing));
        for (int i = 0; i <= n && n > 1; ++i)
            for (int j = 0; i + j <= n; ++j) {
                const double u = (double)i / n, v = (double)j / n, w = 1.0 - u - v;
                for (int a = 0; a < 3; ++a) pts.push_back(u * p0[a] + v * p1[a] + w * p2[a]);
            }
    }
    const size_t np = pts.size() / 3;
    for (int s = 0; s < slabs; ++s) {
        const double a0 = lo[axis] + s * step, a1 = (s == slabs - 1) ? hi[axis] : a0 + step;
        double slo[3], shi[3];
        for (int a = 0; a < 3; ++a) { slo[a] = std::numeric_limits<double>::max(); shi[a] = -slo[a]; }
        bool any = false;
        for (size_t v = 0; v < np; ++v) {
            const double t = pts[3 * v + axis];
            if (t < a0 || t > a1) continue;
            any = true;
            for (int a = 0; a < 3; ++a) { slo[a] = std::min(slo[a], pts[3 * v + a]); shi[a] = std::max(shi[a], pts[3 * v + a]); }
        }
        if (!any) continue;
        FittedSphere sp;
        for (int a = 0; a < 3; ++a) sp.xyz[a] = 0.5 * (slo[a] + shi[a]);
        double r2 = 0.0;
        for (size_t v = 0; v < np; ++v) {
            const double t = pts[3 * v + axis];
            if (t < a0 || t > a1) continue;
            double d2 = 0.0;
            for (int a = 0; a < 3; ++a) { const double d = pts[3 * v + a] - sp.xyz[a]; d2 += d * d; }
            r2 = std::max(r2, d2);
        }
        sp.radius = std::sqrt(r2) * (1.0 + 1e-12) + padding;
        out.push_back(sp);
    }
    return out;
}

void appendBoxMesh(const double centre[3], const double half[3], std::vector<double>& t)
{
    double c[8][3];
    for (int i = 0; i < 8; ++i)
        for (int a = 0; a < 3; ++a) c[i][a] = centre[a] + ((i >> a) & 1 ? half[a] : -half[a]);
    const int quads[6][4] = {{0, 1, 3, 2}, {4, 6, 7, 5}, {0, 4, 5, 1}, {2, 3, 7, 6}, {0, 2, 6, 4}, {1, 5, 7, 3}};
    for (const auto& q : quads) {
        const int tri[2][3] = {{q[0], q[1], q[2]}, {q[0], q[2], q[3]}};
        for (const auto& tr : tri)
            for (int k = 0; k < 3; ++k)
                for (int a = 0; a < 3; ++a) t.push_back(c[tr[k]][a]);
    }
}

void appendCylinderMesh(const double centre[3], double radius, double half_height, std::vector<double>& t, int segments)
{
    const double kPi = 3.14159265358979323846;
    auto put = [&](double x, double y, double z) { t.push_back(centre[0] + x); t.push_back(centre[1] + y); t.push_back(centre[2] + z); };
    for (int s = 0; s < segments; ++s) {
        const double a0 = 2 * kPi * s / segments, a1 = 2 * kPi * (s + 1) / segments;
        const double x0 = radius * std::cos(a0), y0 = radius * std::sin(a0), x1 = radius * std::cos(a1), y1 = radius * std::sin(a1);
        put(x0, y0, -half_height); put(x1, y1, -half_height); put(x1, y1, half_height);
        put(x0, y0, -half_height); put(x1, y1, half_height); put(x0, y0, half_height);
        put(0, 0, -half_height); put(x1, y1, -half_height); put(x0, y0, -half_height);
        put(0, 0, half_height); put(x0, y0, half_height); put(x1, y1, half_height);
    }
}

void appendSphereMesh(const double centre[3], double radius, std::vector<double>& t, int rings, int segments)
{
    const double kPi = 3.14159265358979323846;
    auto pt = [&](int r, int s, double* p) {
        const double th = kPi * r / rings, ph = 2 * kPi * s / segments;
        p[0] = centre[0] + radius * std::sin(th) * std::cos(ph);
        p[1] = centre[1] + radius * std::sin(th) * std::sin(ph);
        p[2] = centre[2] + radius * std::cos(th);
    };
    for (int r = 0; r < rings; ++r)
        for (int s = 0; s < segments; ++s) {
            double a[3], b[3], c[3], d[3];
            pt(r, s, a); pt(r + 1, s, b); pt(r + 1, s + 1, c); pt(r, s + 1, d);
            const double* tri[2][3] = {{a, b, c}, {a, c, d}};
            for (const auto& tr : tri)
                for (int k = 0; k < 3; ++k)
                    for (int i = 0; i < 3; ++i) t.push_back(tr[k][i]);
        }
}

}  // namespace robot_model
