// Constants of stomp/StompUtils.hpp (reference src/planners/stomp/include/stomp/StompUtils.hpp:48-81).
#pragma once
#include <base/Eigen.hpp>

#define STOMP_VERIFY(cond) cond
#define STOMP_VERIFY_MSG(cond, ...) cond

namespace stomp {

static const int DIFF_RULE_LENGTH = 7;
static const int TRAJECTORY_PADDING = DIFF_RULE_LENGTH - 1;
static const int NUM_DIFF_RULES = 4;

enum CostComponents { STOMP_POSITION = 0, STOMP_VELOCITY = 1, STOMP_ACCELERATION = 2, STOMP_JERK = 3 };

// dense differentiation matrix with index clamping at both ends (StompUtils.cpp:6-23)
void getDifferentiationMatrix(int num_time_steps, CostComponents order, double dt, base::MatrixXd& diff_matrix);

}  // namespace stomp
