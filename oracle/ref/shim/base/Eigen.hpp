// ORACLE / TEST INFRASTRUCTURE ONLY: the two typedefs of Rock base-types' base/Eigen.hpp that the STOMP core uses.
#ifndef STOMP_B200_ORACLE_BASE_EIGEN_SHIM
#define STOMP_B200_ORACLE_BASE_EIGEN_SHIM
#include <Eigen/Core>
namespace base {
typedef Eigen::VectorXd VectorXd;
typedef Eigen::MatrixXd MatrixXd;
}
#endif
