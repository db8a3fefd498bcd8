#include <stomp/StompUtils.hpp>

#include <algorithm>

#include "policy_core.hpp"

namespace stomp {

void getDifferentiationMatrix(int num_time_steps, CostComponents order, double dt, base::MatrixXd& diff_matrix)
{
    const stomp_b200::host::DiffBand band = stomp_b200::host::differentiation_band(num_time_steps, (int)order, dt);
    diff_matrix = base::MatrixXd::Zero(num_time_steps, num_time_steps);
    for (int i = 0; i < num_time_steps; ++i)
        for (int j = std::max(0, i - 3); j <= std::min(num_time_steps - 1, i + 3); ++j) diff_matrix(i, j) = band.entry(i, j);
}

}  // namespace stomp
