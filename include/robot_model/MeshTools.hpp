// Mesh input for the scene and the robot: STL loading and automatic sphere fitting.
//
// The reference hands meshes to FCL as they are — MESH model objects of the world (include/motion_planners/Config.hpp:14-35,
// src/MotionPlanners.cpp:416-495) and the robot's own collision meshes (test/data/meshes/**/collision/*.stl).  The CUDA
// path needs a distance field for the world (stomp_b200_build_sdf_scene voxelises the triangles on the device) and
// spheres for the robot's links and grasped objects: fitSpheres covers a mesh with a few spheres along its longest axis.
#pragma once
#include <string>
#include <vector>

namespace robot_model {

struct FittedSphere { double xyz[3]; double radius; };

// binary or ASCII STL -> triangles [n][3 vertices][xyz], appended to `triangles`; vertices scaled, then translated
bool loadStl(const std::string& path, std::vector<double>& triangles, const double scale[3] = nullptr, const double translation[3] = nullptr);

// Covers the vertices of a mesh with spheres: the bounding box is cut into slabs along its longest axis (slab length about
// the box's cross-section diameter, at most max_spheres slabs); each non-empty slab gets the sphere around the centre of its
// own bounding box that holds all its vertices, inflated by `padding`.  Every vertex lies in at least one sphere.
std::vector<FittedSphere> fitSpheres(const std::vector<double>& triangles, int max_spheres = 8, double padding = 0.0);

// triangles of an axis-aligned box / a z-cylinder / a sphere (for grasp objects given as primitives)
void appendBoxMesh(const double centre[3], const double half[3], std::vector<double>& triangles);
void appendCylinderMesh(const double centre[3], double radius, double half_height, std::vector<double>& triangles, int segments = 24);
void appendSphereMesh(const double centre[3], double radius, std::vector<double>& triangles, int rings = 8, int segments = 16);

}  // namespace robot_model
