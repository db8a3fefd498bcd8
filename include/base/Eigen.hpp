// Shim for base/Eigen.hpp of Rock's base-types (absent from this image, as is Eigen itself): the small
// part of base::VectorXd / base::MatrixXd the STOMP path touches through its public API
// (reference usages: StompPlanner.cpp:148-163,186-229, OptimizationTask.cpp:46-66, StompTask.hpp:70-79).
#pragma once
#include <cstddef>
#include <vector>

namespace base {

class VectorXd {
public:
    VectorXd() {}
    explicit VectorXd(int n) : v_((size_t)n, 0.0) {}
    static VectorXd Zero(int n) { return VectorXd(n); }
    static VectorXd Ones(int n) { VectorXd r(n); for (auto& x : r.v_) x = 1.0; return r; }
    static VectorXd Constant(int n, double c) { VectorXd r(n); for (auto& x : r.v_) x = c; return r; }
    int size() const { return (int)v_.size(); }
    int rows() const { return (int)v_.size(); }
    void resize(int n) { v_.resize((size_t)n, 0.0); }
    void setZero() { for (auto& x : v_) x = 0.0; }
    double& operator()(int i) { return v_[(size_t)i]; }
    double operator()(int i) const { return v_[(size_t)i]; }
    double& operator[](int i) { return v_[(size_t)i]; }
    double operator[](int i) const { return v_[(size_t)i]; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }
    double sum() const { double s = 0.0; for (double x : v_) s += x; return s; }
private:
    std::vector<double> v_;
};

class Vector3d {   // positions / dimensions of world objects (base::Vector3d of base-types)
public:
    Vector3d() : v_{0.0, 0.0, 0.0} {}
    Vector3d(double x, double y, double z) : v_{x, y, z} {}
    static Vector3d Zero() { return Vector3d(); }
    double& x() { return v_[0]; } double& y() { return v_[1]; } double& z() { return v_[2]; }
    double x() const { return v_[0]; } double y() const { return v_[1]; } double z() const { return v_[2]; }
    double& operator()(int i) { return v_[i]; }
    double operator()(int i) const { return v_[i]; }
    double& operator[](int i) { return v_[i]; }
    double operator[](int i) const { return v_[i]; }
private:
    double v_[3];
};

class MatrixXd {   // row major
public:
    MatrixXd() : r_(0), c_(0) {}
    MatrixXd(int r, int c) : r_(r), c_(c), v_((size_t)r * c, 0.0) {}
    static MatrixXd Zero(int r, int c) { return MatrixXd(r, c); }
    static MatrixXd Identity(int r, int c) { MatrixXd m(r, c); for (int i = 0; i < r && i < c; ++i) m(i, i) = 1.0; return m; }
    int rows() const { return r_; }
    int cols() const { return c_; }
    void resize(int r, int c) { r_ = r; c_ = c; v_.assign((size_t)r * c, 0.0); }
    double& operator()(int i, int j) { return v_[(size_t)i * c_ + j]; }
    double operator()(int i, int j) const { return v_[(size_t)i * c_ + j]; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }
private:
    int r_, c_;
    std::vector<double> v_;
};

}  // namespace base
