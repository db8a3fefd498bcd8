// Environment -> signed distance field, on the device.
//
// The reference keeps its world as collision objects handed to FCL — primitives, meshes and an octomap
// (src/MotionPlanners.cpp:162-173 assignOctomapPlanningScene / updateOctomap, :416-495 handleCollisionObjectInWorld /
// handleGraspObject, include/motion_planners/Config.hpp:14-35) — and asks FCL per state.  Here the world is a distance
// field the state kernel gathers from, so the world objects are turned into that field once per scene change:
//   * primitives (sphere / box): exact signed distance of the union at every voxel centre, one thread per voxel
//     (build_sdf_primitives_kernel);
//   * occupancy (a voxelised mesh or an octomap's leaf grid): exact Euclidean distance transform, three separable
//     min-plus passes over squared integer distances (edt_pass_kernel), signed by running it on the occupancy and on its
//     complement (finish_edt_kernel).
// Arithmetic contract (the oracle's oracle_build_sdf_* issue the same operations; bit-identical grids): FP64, one IEEE
// operation per source operation (-fmad=false), voxel centre = origin + (i + 0.5) * h, distances rounded to binary32.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace stomp_b200 {

constexpr int kMaxPrimitives = 256;

struct PrimitiveList {
    int32_t n;
    int32_t kind[kMaxPrimitives];        // 0 sphere (size[0] = radius), 1 box (size = half extents)
    double centre[kMaxPrimitives][3];
    double size[kMaxPrimitives][3];
};

// signed distance from p to one primitive (negative inside)
__device__ __forceinline__ double primitive_distance(int kind, const double* c, const double* s, double px, double py, double pz)
{
    const double dx = px - c[0], dy = py - c[1], dz = pz - c[2];
    if (kind == 0) return sqrt((dx * dx + dy * dy) + dz * dz) - s[0];
    const double qx = fabs(dx) - s[0], qy = fabs(dy) - s[1], qz = fabs(dz) - s[2];
    const double ox = fmax(qx, 0.0), oy = fmax(qy, 0.0), oz = fmax(qz, 0.0);
    const double outside = sqrt((ox * ox + oy * oy) + oz * oz);
    const double inside = fmin(fmax(fmax(qx, qy), qz), 0.0);
    return outside + inside;
}

// grid[(z * ny + y) * nx + x] = (float) min_i d_i(centre of voxel (x, y, z)); the list lives in global memory (one
// broadcast load per primitive and warp).  grid (ceil(nx / 128), ny, nz) — x along the thread index: coalesced stores.
__global__ void __launch_bounds__(128)
build_sdf_primitives_kernel(float* __restrict__ grid, int nx, int ny, int nz, double ox, double oy, double oz, double h,
                            const PrimitiveList* __restrict__ list)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, z = blockIdx.z;
    if (x >= nx) return;
    const double px = ox + ((double)x + 0.5) * h, py = oy + ((double)y + 0.5) * h, pz = oz + ((double)z + 0.5) * h;
    double d = __longlong_as_double(0x7ff0000000000000ll);       // +inf
    const int n = list->n;
    for (int i = 0; i < n; ++i) d = fmin(d, primitive_distance(list->kind[i], list->centre[i], list->size[i], px, py, pz));
    grid[((size_t)z * ny + y) * nx + x] = (float)d;
}

// ---- exact Euclidean distance transform of an occupancy grid ---------------------------------------------------
// Squared distances in voxel units are integers: three passes of  out(i) = min_j (i - j)^2 + in(j)  along x, y, z
// (min-plus with a parabola; exact in int32 for grids up to 1024^3: 3 * 1023^2 < 2^31).  O(n) per voxel and pass — a
// 256^3 grid is 13 G integer operations, milliseconds on a B200, once per scene change; the O(1) lower-envelope sweep
// of Felzenszwalb & Huttenlocher is sequential along a line and does not map to a warp.
// One CTA per line: the line is staged in shared memory, thread i scans it.
constexpr int kEdtInf = 0x3fffffff;

__global__ void __launch_bounds__(256)
edt_pass_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int n_line, long long stride_line,
                int n_a, long long stride_a, long long stride_b)
{
    extern __shared__ int32_t line[];
    const size_t base = (size_t)blockIdx.x * stride_a + (size_t)blockIdx.y * stride_b;
    (void)n_a;
    for (int i = threadIdx.x; i < n_line; i += blockDim.x) line[i] = in[base + (size_t)i * stride_line];
    __syncthreads();
    for (int i = threadIdx.x; i < n_line; i += blockDim.x) {
        int best = kEdtInf;
        for (int j = 0; j < n_line; ++j) {
            const int v = line[j];
            const int dj = i - j;
            const int cand = (v >= kEdtInf) ? kEdtInf : v + dj * dj;
            best = min(best, cand);
        }
        out[base + (size_t)i * stride_line] = best;
    }
}

// seeds of the two transforms: distance to the nearest occupied voxel (0 on occupied voxels) and to the nearest free one
__global__ void edt_seed_kernel(const uint8_t* __restrict__ occupied, int32_t* __restrict__ to_occupied,
                                int32_t* __restrict__ to_free, size_t count)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const bool occ = occupied[i] != 0;
    to_occupied[i] = occ ? 0 : kEdtInf;
    to_free[i] = occ ? kEdtInf : 0;
}

// signed distance of the voxel centre to the occupied set's boundary voxels, in metres: +h * sqrt(d2 to the nearest
// occupied voxel) outside, -h * sqrt(d2 to the nearest free voxel) inside (centre-to-centre; a scene without occupied
// voxels is +inf-like: h * sqrt(kEdtInf))
__global__ void finish_edt_kernel(const uint8_t* __restrict__ occupied, const int32_t* __restrict__ to_occupied,
                                  const int32_t* __restrict__ to_free, float* __restrict__ grid, double h, size_t count)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const bool occ = occupied[i] != 0;
    const double d2 = (double)(occ ? to_free[i] : to_occupied[i]);
    const double d = h * sqrt(d2);
    grid[i] = (float)(occ ? -d : d);
}

// copy of the grid in 4 x 4 x 2-voxel bricks (SdfParams::bricks); one thread per brick voxel, padding voxels repeat the
// nearest grid voxel
__global__ void brick_sdf_kernel(const float* __restrict__ grid, float* __restrict__ bricks, int nx, int ny, int nz, int nbx, int nby, size_t count)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned in = (unsigned)(i & 31);
    size_t b = i >> 5;
    const int bx = (int)(b % (size_t)nbx); b /= (size_t)nbx;
    const int by = (int)(b % (size_t)nby);
    const int bz = (int)(b / (size_t)nby);
    const int x = min(bx * 4 + (int)(in & 3), nx - 1), y = min(by * 4 + (int)((in >> 2) & 3), ny - 1), z = min(bz * 2 + (int)(in >> 4), nz - 1);
    bricks[i] = grid[((size_t)z * ny + y) * nx + x];
}

}  // namespace stomp_b200
