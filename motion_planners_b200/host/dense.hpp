// Small dense FP64 helpers for the one-time host setup of the STOMP policy (the reference uses Eigen
// for these: CovariantMovementPrimitive.cpp:273 fullPivLu().inverse(), MultivariateGaussian.hpp:81
// llt().matrixL()).  Row-major, no third-party headers (none exist in this image).
#pragma once
#include <cstddef>
#include <vector>

namespace stomp_b200 {
namespace host {

class Dense {
public:
    Dense() : n_rows_(0), n_cols_(0) {}
    Dense(int rows, int cols) : n_rows_(rows), n_cols_(cols), v_((size_t)rows * cols, 0.0) {}
    int rows() const { return n_rows_; }
    int cols() const { return n_cols_; }
    double& at(int i, int j) { return v_[(size_t)i * n_cols_ + j]; }
    double at(int i, int j) const { return v_[(size_t)i * n_cols_ + j]; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }
private:
    int n_rows_, n_cols_;
    std::vector<double> v_;
};

// inverse through an LU factorisation with complete pivoting (same pivoting rule as Eigen's
// FullPivLU, which the reference uses); returns false if a zero pivot is met
bool invert_full_pivot(const Dense& a, Dense& inverse);
// lower Cholesky factor of a symmetric positive definite matrix; false if not positive definite
bool cholesky_lower(const Dense& a, Dense& lower);

}  // namespace host
}  // namespace stomp_b200
