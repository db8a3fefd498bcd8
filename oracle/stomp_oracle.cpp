// ORACLE — TEST INFRASTRUCTURE ONLY.  See stomp_oracle.hpp for the header comment, the parity
// status (pinned against the reference's own code, oracle/ref) and the rule on who may use this code.
#include "stomp_oracle.hpp"

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

// =================================================================================================
// Linear algebra helpers
// =================================================================================================

// Eigen::FullPivLU<MatrixXd>::inverse() as used at CovariantMovementPrimitive.cpp:273 and
// PolicyImprovement.cpp:797: complete pivoting (largest |entry| of the remaining corner), in-place
// LU, then A^-1 = Q U^-1 L^-1 P applied to the identity.
Mat full_piv_lu_inverse(const Mat& A)
{
    const int n = A.rows;
    assert(A.rows == A.cols);
    Mat lu = A;
    std::vector<int> row_tr(n), col_tr(n);
    for (int k = 0; k < n; ++k) {
        int pr = k, pc = k;
        double biggest = 0.0;
        for (int i = k; i < n; ++i)
            for (int j = k; j < n; ++j) {
                double v = std::fabs(lu(i, j));
                if (v > biggest) { biggest = v; pr = i; pc = j; }
            }
        if (biggest == 0.0) {
            for (int i = k; i < n; ++i) { row_tr[i] = i; col_tr[i] = i; }
            break;
        }
        row_tr[k] = pr; col_tr[k] = pc;
        if (pr != k) for (int j = 0; j < n; ++j) std::swap(lu(k, j), lu(pr, j));
        if (pc != k) for (int i = 0; i < n; ++i) std::swap(lu(i, k), lu(i, pc));
        double piv = lu(k, k);
        for (int i = k + 1; i < n; ++i) lu(i, k) /= piv;
        for (int i = k + 1; i < n; ++i) {
            double lik = lu(i, k);
            if (lik == 0.0) continue;
            for (int j = k + 1; j < n; ++j) lu(i, j) -= lik * lu(k, j);
        }
    }
    // permutation P (rows): apply the transpositions in order to the identity
    std::vector<int> p(n), q(n);
    for (int i = 0; i < n; ++i) { p[i] = i; q[i] = i; }
    // P = T_{n-1} ... T_0 ; row i of (P*B) = row p_idx[i] of B
    for (int k = 0; k < n; ++k) std::swap(p[k], p[row_tr[k]]);
    for (int k = 0; k < n; ++k) std::swap(q[k], q[col_tr[k]]);
    Mat inv(n, n);
    Vec c(n);
    for (int col = 0; col < n; ++col) {
        // c = P * e_col
        for (int i = 0; i < n; ++i) c[i] = (p[i] == col) ? 1.0 : 0.0;
        // forward substitution, unit lower
        for (int i = 0; i < n; ++i) {
            double s = c[i];
            for (int j = 0; j < i; ++j) s -= lu(i, j) * c[j];
            c[i] = s;
        }
        // back substitution
        for (int i = n - 1; i >= 0; --i) {
            double s = c[i];
            for (int j = i + 1; j < n; ++j) s -= lu(i, j) * c[j];
            c[i] = s / lu(i, i);
        }
        // x = Q * c  :  x[q[i]] = c[i]
        for (int i = 0; i < n; ++i) inv(q[i], col) = c[i];
    }
    return inv;
}

// Eigen::LLT<MatrixXd>::matrixL() (MultivariateGaussian.hpp:81): plain Cholesky, no pivoting
Mat llt_lower(const Mat& A)
{
    const int n = A.rows;
    Mat L(n, n);
    for (int j = 0; j < n; ++j) {
        double s = A(j, j);
        for (int k = 0; k < j; ++k) s -= L(j, k) * L(j, k);
        double d = std::sqrt(s);
        L(j, j) = d;
        for (int i = j + 1; i < n; ++i) {
            double t = A(i, j);
            for (int k = 0; k < j; ++k) t -= L(i, k) * L(j, k);
            L(i, j) = t / d;
        }
    }
    return L;
}

// StompUtils.hpp:60-66
const double DIFF_RULES[NUM_DIFF_RULES][DIFF_RULE_LENGTH] = {
    {0, 0, 0, 1, 0, 0, 0},
    {0, 0, -1, 1, 0, 0, 0},
    {0, -1 / 12.0, 16 / 12.0, -30 / 12.0, 16 / 12.0, -1 / 12.0, 0},
    {0, 1 / 12.0, -17 / 12.0, 46 / 12.0, -46 / 12.0, 17 / 12.0, -1 / 12.0}};

// StompUtils.cpp:6-23
void getDifferentiationMatrix(int num_time_steps, int order, double dt, Mat& diff_matrix)
{
    diff_matrix = Mat(num_time_steps, num_time_steps);
    double multiplier = 1.0 / std::pow(dt, (int)order);
    for (int i = 0; i < num_time_steps; ++i) {
        for (int j = -DIFF_RULE_LENGTH / 2; j <= DIFF_RULE_LENGTH / 2; ++j) {
            int index = i + j;
            if (index < 0) index = 0;
            if (index >= num_time_steps) index = num_time_steps - 1;
            diff_matrix(i, index) += multiplier * DIFF_RULES[order][j + DIFF_RULE_LENGTH / 2];
        }
    }
}

// =================================================================================================
// CovariantMovementPrimitive
// =================================================================================================

bool CovariantMovementPrimitive::initialize(int num_time_steps, int num_dimensions, double movement_duration,
                                            const std::vector<Mat>& derivative_costs,
                                            const std::vector<Vec>& initial_trajectory)
{
    num_time_steps_ = num_time_steps;
    num_dimensions_ = num_dimensions;
    movement_duration_ = movement_duration;
    derivative_costs_ = derivative_costs;
    parameters_all_ = initial_trajectory;

    // initializeVariables  :202-225
    movement_dt_ = movement_duration_ / (num_time_steps_ + 1);
    num_vars_free_ = num_time_steps_;
    num_vars_all_ = num_vars_free_ + 2 * (DIFF_RULE_LENGTH - 1);
    free_vars_start_index_ = DIFF_RULE_LENGTH - 1;
    free_vars_end_index_ = free_vars_start_index_ + num_vars_free_ - 1;

    // initializeCosts  :242-280
    differentiation_matrices_.assign(NUM_DIFF_RULES, Mat());
    for (int d = 0; d < NUM_DIFF_RULES; ++d)
        getDifferentiationMatrix(num_vars_all_, d, movement_dt_, differentiation_matrices_[d]);

    control_costs_all_.clear(); control_costs_.clear(); inv_control_costs_.clear(); derivative_costs_sqrt_.clear();
    const int N = num_vars_all_, T = num_vars_free_;
    for (int d = 0; d < num_dimensions_; ++d) {
        Mat sq(N, NUM_DIFF_RULES);
        for (int i = 0; i < N; ++i)
            for (int r = 0; r < NUM_DIFF_RULES; ++r) sq(i, r) = std::sqrt(derivative_costs_[d](i, r));
        derivative_costs_sqrt_.push_back(sq);

        // every dimension of the shipped task carries the same derivative costs
        // (OptimizationTask.cpp:32-33); reuse the factorisation when that is the case
        bool same_as_first = d > 0 && derivative_costs_[d].a == derivative_costs_[0].a;
        if (same_as_first) {
            control_costs_all_.push_back(control_costs_all_[0]);
            control_costs_.push_back(control_costs_[0]);
            inv_control_costs_.push_back(inv_control_costs_[0]);
            continue;
        }
        Mat cost_all(N, N);
        for (int r = 0; r < NUM_DIFF_RULES; ++r) {
            const Mat& Dm = differentiation_matrices_[r];
            // cost_all += dt * (D^T diag(w) D)
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j) {
                    double s = 0.0;
                    int lo = std::max(0, std::max(i, j) - DIFF_RULE_LENGTH);
                    int hi = std::min(N - 1, std::min(i, j) + DIFF_RULE_LENGTH);
                    for (int k = lo; k <= hi; ++k) s += Dm(k, i) * derivative_costs_[d](k, r) * Dm(k, j);
                    cost_all(i, j) += movement_dt_ * s;
                }
        }
        control_costs_all_.push_back(cost_all);
        Mat cost_free(T, T);
        for (int i = 0; i < T; ++i)
            for (int j = 0; j < T; ++j)
                cost_free(i, j) = cost_all(DIFF_RULE_LENGTH - 1 + i, DIFF_RULE_LENGTH - 1 + j);
        control_costs_.push_back(cost_free);
        inv_control_costs_.push_back(full_piv_lu_inverse(cost_free));
    }
    computeLinearControlCosts();
    return true;
}

bool CovariantMovementPrimitive::setToMinControlCost()
{
    computeMinControlCostParameters();
    return true;
}

// :136-172
bool CovariantMovementPrimitive::computeLinearControlCosts()
{
    const int T = num_vars_free_, P = DIFF_RULE_LENGTH - 1;
    linear_control_costs_.assign(num_dimensions_, Vec(T, 0.0));
    constant_control_costs_.assign(num_dimensions_, 0.0);
    for (int d = 0; d < num_dimensions_; ++d) {
        const Mat& C = control_costs_all_[d];
        Vec& lin = linear_control_costs_[d];
        for (int j = 0; j < T; ++j) {
            double s = 0.0;
            for (int i = 0; i < P; ++i) s += parameters_all_[d][i] * C(i, free_vars_start_index_ + j);
            lin[j] = s;
        }
        for (int j = 0; j < T; ++j) {
            double s = 0.0;
            for (int i = 0; i < P; ++i)
                s += parameters_all_[d][free_vars_end_index_ + 1 + i] * C(free_vars_end_index_ + 1 + i, free_vars_start_index_ + j);
            lin[j] += s;
        }
        for (int j = 0; j < T; ++j) lin[j] *= 2.0;
        for (int j = 0; j < T; ++j)
            lin[j] += -movement_dt_ * 2.0 * (parameters_all_[d][free_vars_start_index_ + j] *
                                             derivative_costs_[d](free_vars_start_index_ + j, 0));
        // constant part (:157-166)
        Vec cp(2 * TRAJECTORY_PADDING);
        Mat cm(2 * TRAJECTORY_PADDING, 2 * TRAJECTORY_PADDING);
        const int N = num_vars_all_, Q = TRAJECTORY_PADDING;
        for (int i = 0; i < Q; ++i) { cp[i] = parameters_all_[d][i]; cp[Q + i] = parameters_all_[d][free_vars_end_index_ + 1 + i]; }
        for (int i = 0; i < Q; ++i)
            for (int j = 0; j < Q; ++j) {
                cm(i, j) = C(i, j);
                cm(Q + i, Q + j) = C(N - Q + i, N - Q + j);
                cm(i, Q + j) = C(i, N - Q + j);
                cm(Q + i, j) = C(N - Q + i, j);
            }
        double acc = 0.0;
        for (int i = 0; i < 2 * Q; ++i) {
            double s = 0.0;
            for (int j = 0; j < 2 * Q; ++j) s += cm(i, j) * cp[j];
            acc += cp[i] * s;
        }
        constant_control_costs_[d] = movement_dt_ * acc;
    }
    return true;
}

// :174-189
bool CovariantMovementPrimitive::computeMinControlCostParameters()
{
    const int T = num_vars_free_;
    for (int d = 0; d < num_dimensions_; ++d) {
        for (int i = 0; i < T; ++i) {
            double s = 0.0;
            for (int j = 0; j < T; ++j) s += inv_control_costs_[d](i, j) * linear_control_costs_[d][j];
            parameters_all_[d][free_vars_start_index_ + i] = -0.5 * s;
        }
    }
    return updateMinControlCostParameters(parameters_all_);
}

// :191-200
bool CovariantMovementPrimitive::updateMinControlCostParameters(const std::vector<Vec>& parameters_all)
{
    min_control_cost_parameters_all_ = parameters_all;
    min_control_cost_parameters_free_.resize(num_dimensions_);
    for (int d = 0; d < num_dimensions_; ++d)
        min_control_cost_parameters_free_[d].assign(
            min_control_cost_parameters_all_[d].begin() + free_vars_start_index_,
            min_control_cost_parameters_all_[d].begin() + free_vars_start_index_ + num_vars_free_);
    return true;
}

bool CovariantMovementPrimitive::getParameters(std::vector<Vec>& parameters) const
{
    parameters.resize(num_dimensions_);
    for (int d = 0; d < num_dimensions_; ++d)
        parameters[d].assign(parameters_all_[d].begin() + free_vars_start_index_,
                             parameters_all_[d].begin() + free_vars_start_index_ + num_vars_free_);
    return true;
}

// :327-412.  costs_all += dt*weight*(Ax*Ax), Ax = (D_i * params_all) .* sqrt(w_i); then the padding
// rows are folded into the first / last free timestep.
bool CovariantMovementPrimitive::computeControlCosts(const std::vector<Vec>& parameters, const std::vector<Vec>& noise,
                                                     double weight, std::vector<Vec>& control_costs,
                                                     bool dense_form) const
{
    const int N = num_vars_all_, T = num_vars_free_;
    control_costs.resize(num_dimensions_);
    Vec params_all(N), costs_all(N), Dx(N);
    for (int d = 0; d < num_dimensions_; ++d) {
        params_all = parameters_all_[d];
        for (int t = 0; t < T; ++t) params_all[free_vars_start_index_ + t] = parameters[d][t] + noise[d][t];
        std::fill(costs_all.begin(), costs_all.end(), 0.0);
        const double dtw = movement_dt_ * weight;
        for (int r = 0; r < NUM_DIFF_RULES; ++r) {
            const Mat& Dm = differentiation_matrices_[r];
            if (dense_form) {
                // the reference's dense N x N product, evaluated column by column (axpy form)
                std::fill(Dx.begin(), Dx.end(), 0.0);
                for (int j = 0; j < N; ++j) {
                    const double xj = params_all[j];
                    for (int i = 0; i < N; ++i) Dx[i] += Dm(i, j) * xj;
                }
            } else {
                for (int i = 0; i < N; ++i) {
                    double s = 0.0;
                    int lo = std::max(0, i - DIFF_RULE_LENGTH / 2), hi = std::min(N - 1, i + DIFF_RULE_LENGTH / 2);
                    for (int j = lo; j <= hi; ++j) s += Dm(i, j) * params_all[j];
                    Dx[i] = s;
                }
            }
            for (int i = 0; i < N; ++i) {
                double Ax = Dx[i] * derivative_costs_sqrt_[d](i, r);
                costs_all[i] += dtw * (Ax * Ax);
            }
        }
        Vec& cc = control_costs[d];
        cc.assign(costs_all.begin() + free_vars_start_index_, costs_all.begin() + free_vars_start_index_ + T);
        for (int i = 0; i < free_vars_start_index_; ++i) {
            cc[0] += costs_all[i];
            cc[T - 1] += costs_all[N - (i + 1)];
        }
    }
    return true;
}

// :463-520  (divisor = 1, only row 0 of each update matrix is used)
bool CovariantMovementPrimitive::updateParameters(const std::vector<Mat>& updates)
{
    const double divisor = 1.0;
    for (int d = 0; d < num_dimensions_; ++d)
        for (int t = 0; t < num_vars_free_; ++t)
            parameters_all_[d][free_vars_start_index_ + t] += divisor * updates[d](0, t);
    return true;
}

// =================================================================================================
// MultivariateGaussian
// =================================================================================================

MultivariateGaussian::MultivariateGaussian(const Vec& mean, const Mat& covariance, uint64_t seed)
    : covariance_cholesky_(llt_lower(covariance)), mean_(mean), size_((int)mean.size()), rng_(seed)
{
}

MultivariateGaussian MultivariateGaussian::fromFactor(const Vec& mean, const Mat& factor, uint64_t seed)
{
    MultivariateGaussian g;
    g.covariance_cholesky_ = factor;
    g.mean_ = mean;
    g.size_ = (int)mean.size();
    g.rng_.seed(seed);
    return g;
}

double MultivariateGaussian::normal()
{
    if (have_spare_) { have_spare_ = false; return spare_; }
    // Box-Muller on two 53-bit uniforms in (0,1)
    double u1 = ((double)(rng_() >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    double u2 = ((double)(rng_() >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    double rad = std::sqrt(-2.0 * std::log(u1));
    double ang = 6.283185307179586476925286766559 * u2;
    spare_ = rad * std::sin(ang);
    have_spare_ = true;
    return rad * std::cos(ang);
}

// MultivariateGaussian.hpp:91-97
void MultivariateGaussian::transform(const Vec& eps, Vec& output) const
{
    output.resize(size_);
    for (int i = 0; i < size_; ++i) {
        double s = 0.0;
        for (int j = 0; j <= i; ++j) s += covariance_cholesky_(i, j) * eps[j];
        output[i] = mean_[i] + s;
    }
}

void MultivariateGaussian::sample(Vec& output)
{
    Vec eps(size_);
    for (int i = 0; i < size_; ++i) eps[i] = normal();
    transform(eps, output);
}

// =================================================================================================
// PolicyImprovement
// =================================================================================================

PolicyImprovement::PolicyImprovement()
{
    cost_scaling_h_ = 10.0;
    use_cumulative_costs_ = true;
    use_projection_ = false;
}

bool PolicyImprovement::initialize(int num_time_steps, int min_rollouts, int max_rollouts, int num_rollouts_per_iteration,
                                   std::shared_ptr<CovariantMovementPrimitive> policy, bool use_noise_adaptation,
                                   const Vec& noise_min_stddev, double control_cost_weight, uint64_t seed)
{
    num_time_steps_ = num_time_steps;
    noise_min_stddev_ = noise_min_stddev;
    policy_ = policy;
    use_covariance_matrix_adaptation_ = use_noise_adaptation;
    adapted_covariance_valid_ = false;
    control_cost_weight_ = control_cost_weight;

    control_costs_ = policy_->control_costs_;
    num_dimensions_ = policy_->num_dimensions_;
    policy_->getParameters(parameters_);
    inv_control_costs_ = policy_->inv_control_costs_;

    noise_generators_.clear();
    adapted_stddevs_.assign(num_dimensions_, 1.0);
    for (int d = 0; d < num_dimensions_; ++d) {
        // the reference factorises per dimension (:91-96); every dimension of the shipped task has the
        // same R^-1, so the factor of dimension 0 is reused when the matrices are identical
        if (d > 0 && inv_control_costs_[d].a == inv_control_costs_[0].a)
            noise_generators_.push_back(MultivariateGaussian::fromFactor(
                Vec(num_time_steps_, 0.0), noise_generators_[0].covariance_cholesky_, seed + 7919ull * d));
        else
            noise_generators_.push_back(MultivariateGaussian(Vec(num_time_steps_, 0.0), inv_control_costs_[d], seed + 7919ull * d));
    }
    noiseless_rollout_valid_ = false;
    setNumRollouts(min_rollouts, max_rollouts, num_rollouts_per_iteration);

    // preAllocateTempVariables :731-748
    tmp_noise_.assign(num_dimensions_, Vec(num_time_steps_, 0.0));
    parameter_updates_.assign(num_dimensions_, Mat(num_time_steps_, num_time_steps_));
    time_step_weights_.assign(num_dimensions_, Vec(num_time_steps_, 0.0));
    preComputeProjectionMatrices();
    return true;
}

bool PolicyImprovement::setNumRollouts(int min_rollouts, int max_rollouts, int num_rollouts_per_iteration)
{
    min_rollouts_ = min_rollouts;
    max_rollouts_ = max_rollouts;
    num_rollouts_per_iteration_ = num_rollouts_per_iteration;
    num_rollouts_ = 0;
    num_rollouts_gen_ = 0;
    Rollout rollout;
    const int T = num_time_steps_, D = num_dimensions_;
    rollout.parameters_.assign(D, Vec(T, 0.0));
    rollout.parameters_noise_.assign(D, Vec(T, 0.0));
    rollout.parameters_noise_projected_.assign(D, Vec(T, 0.0));
    rollout.noise_.assign(D, Vec(T, 0.0));
    rollout.noise_projected_.assign(D, Vec(T, 0.0));
    rollout.control_costs_.assign(D, Vec(T, 0.0));
    rollout.total_costs_.assign(D, Vec(T, 0.0));
    rollout.cumulative_costs_.assign(D, Vec(T, 0.0));
    rollout.probabilities_.assign(D, Vec(T, 0.0));
    rollout.full_probabilities_.assign(D, 0.0);
    rollout.full_costs_.assign(D, 0.0);
    rollout.state_costs_.assign(T, 0.0);
    rollouts_.assign(max_rollouts_ + 1, rollout);
    reused_rollouts_.assign(max_rollouts_ + 1, rollout);
    noiseless_rollout_ = rollout;
    rollout_cost_sorter_.reserve(max_rollouts_);
    return true;
}

// :158-311
bool PolicyImprovement::generateRollouts(const Vec& noise_stddev, const NoiseSource& src)
{
    if (!adapted_covariance_valid_) adapted_stddevs_ = noise_stddev;
    policy_->getParameters(parameters_);   // copyParametersFromPolicy :819-827

    const int T = num_time_steps_, D = num_dimensions_;
    int num_rollouts_discard = 0;
    int num_rollouts_reused = num_rollouts_;
    int prev_num_rollouts = num_rollouts_;
    num_rollouts_gen_ = num_rollouts_per_iteration_;
    if (num_rollouts_ + num_rollouts_gen_ < min_rollouts_) {
        num_rollouts_gen_ = min_rollouts_ - num_rollouts_;
        num_rollouts_discard = 0;
        num_rollouts_reused = num_rollouts_;
    }
    if (num_rollouts_ + num_rollouts_gen_ > max_rollouts_) {
        num_rollouts_discard = num_rollouts_ + num_rollouts_gen_ - max_rollouts_;
        num_rollouts_reused = num_rollouts_ - num_rollouts_discard;
    }
    num_rollouts_ = num_rollouts_reused + num_rollouts_gen_;

    if (num_rollouts_reused > 0) {
        double min_cost = rollouts_[0].total_cost_;
        double max_cost = min_cost;
        for (int r = 1; r < prev_num_rollouts; ++r) {
            double c = rollouts_[r].total_cost_;
            if (c < min_cost) min_cost = c;
            if (c > max_cost) max_cost = c;
        }
        double cost_denom = max_cost - min_cost;
        if (cost_denom < 1e-8) cost_denom = 1e-8;

        rollout_cost_sorter_.clear();
        for (int r = 0; r < prev_num_rollouts; ++r) {
            rollouts_[r].parameters_ = parameters_;
            for (int d = 0; d < D; ++d) {
                for (int t = 0; t < T; ++t)
                    rollouts_[r].noise_projected_[d][t] = rollouts_[r].parameters_noise_projected_[d][t] - parameters_[d][t];
                // noise = inv_projection * noise_projected
                if (!use_projection_) {
                    rollouts_[r].noise_[d] = rollouts_[r].noise_projected_[d];
                } else {
                    for (int i = 0; i < T; ++i) {
                        double s = 0.0;
                        for (int j = 0; j < T; ++j) s += inv_projection_matrix_[d](i, j) * rollouts_[r].noise_projected_[d][j];
                        rollouts_[r].noise_[d][i] = s;
                    }
                }
                for (int t = 0; t < T; ++t)
                    rollouts_[r].parameters_noise_[d][t] = parameters_[d][t] + rollouts_[r].noise_[d][t];
            }
            rollouts_[r].importance_weight_ = 1.0;
            rollouts_[r].log_likelihood_ = 0.0;
            double cost_prob = std::exp(-cost_scaling_h_ * (rollouts_[r].total_cost_ - min_cost) / cost_denom);
            double weighted_cost = cost_prob * rollouts_[r].importance_weight_;
            rollout_cost_sorter_.push_back(std::make_pair(-weighted_cost, r));
        }
        std::sort(rollout_cost_sorter_.begin(), rollout_cost_sorter_.end());
        for (int r = 0; r < num_rollouts_reused; ++r) reused_rollouts_[r] = rollouts_[rollout_cost_sorter_[r].second];
        for (int r = 0; r < num_rollouts_reused; ++r) rollouts_[num_rollouts_gen_ + r] = reused_rollouts_[r];
    }

    // generate new rollouts: dimension outer, rollout inner (:258-286)
    for (int d = 0; d < D; ++d) {
        double l1 = control_cost_weight_;
        double l2 = 1.0 / (adapted_stddevs_[d] * adapted_stddevs_[d]);
        double new_stddev = 1.0 / std::sqrt(l1 + l2);
        double p1 = l1 / (l1 + l2);
        double p2 = l2 / (l1 + l2);
        const Vec& mincc = policy_->min_control_cost_parameters_free_[d];
        for (int r = 0; r < num_rollouts_gen_; ++r) {
            if (src.injected) {
                const double* n = src.injected + ((size_t)r * D + d) * T;
                tmp_noise_[d].assign(n, n + T);
            } else if (src.epsilon) {
                const double* e = src.epsilon + ((size_t)r * D + d) * T;
                noise_generators_[d].transform(Vec(e, e + T), tmp_noise_[d]);
            } else {
                noise_generators_[d].sample(tmp_noise_[d]);
            }
            Rollout& ro = rollouts_[r];
            for (int t = 0; t < T; ++t)
                ro.parameters_noise_[d][t] = p1 * mincc[t] + p2 * parameters_[d][t] + new_stddev * tmp_noise_[d][t];
            ro.parameters_[d] = parameters_[d];
            for (int t = 0; t < T; ++t) ro.noise_[d][t] = ro.parameters_noise_[d][t] - ro.parameters_[d][t];
        }
    }
    for (int r = 0; r < num_rollouts_gen_; ++r) rollouts_[r].importance_weight_ = 1.0;

    if (noiseless_rollout_valid_) {
        rollouts_[num_rollouts_] = noiseless_rollout_;
        ++num_rollouts_;
    }
    return true;
}

bool PolicyImprovement::getRollouts(std::vector<std::vector<Vec>>& rollouts, const Vec& noise_stddev, const NoiseSource& src)
{
    if (!generateRollouts(noise_stddev, src)) return false;
    rollouts.clear();
    for (int r = 0; r < num_rollouts_gen_; ++r) rollouts.push_back(rollouts_[r].parameters_noise_);
    return true;
}

bool PolicyImprovement::getProjectedRollouts(std::vector<std::vector<Vec>>& rollouts)
{
    rollouts.clear();
    for (int r = 0; r < num_rollouts_gen_; ++r) rollouts.push_back(rollouts_[r].parameters_noise_projected_);
    return true;
}

bool PolicyImprovement::setRollouts(const std::vector<std::vector<Vec>>& rollouts)
{
    for (int r = 0; r < num_rollouts_gen_; ++r) {
        rollouts_[r].parameters_noise_ = rollouts[r];
        computeNoise(rollouts_[r]);
    }
    return true;
}

bool PolicyImprovement::setRolloutCosts(const Mat& costs, double control_cost_weight, Vec& rollout_costs_total)
{
    control_cost_weight_ = control_cost_weight;
    computeRolloutControlCosts();
    for (int r = 0; r < num_rollouts_gen_; ++r)
        for (int t = 0; t < num_time_steps_; ++t) rollouts_[r].state_costs_[t] = costs(r, t);
    computeRolloutCumulativeCosts(rollout_costs_total);
    return true;
}

bool PolicyImprovement::setNoiselessRolloutCosts(const Vec& costs, double& total_cost)
{
    policy_->getParameters(noiseless_rollout_.parameters_);
    for (int d = 0; d < num_dimensions_; ++d) {
        noiseless_rollout_.noise_[d].assign(num_time_steps_, 0.0);
        noiseless_rollout_.noise_projected_[d].assign(num_time_steps_, 0.0);
        noiseless_rollout_.parameters_noise_[d] = noiseless_rollout_.parameters_[d];
        noiseless_rollout_.parameters_noise_projected_[d] = noiseless_rollout_.parameters_[d];
    }
    noiseless_rollout_.state_costs_ = costs;
    noiseless_rollout_.importance_weight_ = 1.0;
    computeRolloutControlCosts(noiseless_rollout_);
    computeRolloutCumulativeCosts(noiseless_rollout_);
    total_cost = noiseless_rollout_.total_cost_;
    noiseless_rollout_valid_ = true;
    return true;
}

bool PolicyImprovement::computeProjectedNoise()
{
    for (int r = 0; r < num_rollouts_; ++r) computeProjectedNoise(rollouts_[r]);
    return true;
}

bool PolicyImprovement::computeProjectedNoise(Rollout& rollout)
{
    const int T = num_time_steps_;
    for (int d = 0; d < num_dimensions_; ++d) {
        if (!use_projection_) {
            rollout.noise_projected_[d] = rollout.noise_[d];   // identity * noise
        } else {
            for (int i = 0; i < T; ++i) {
                double s = 0.0;
                for (int j = 0; j < T; ++j) s += projection_matrix_[d](i, j) * rollout.noise_[d][j];
                rollout.noise_projected_[d][i] = s;
            }
        }
        for (int t = 0; t < T; ++t)
            rollout.parameters_noise_projected_[d][t] = rollout.parameters_[d][t] + rollout.noise_projected_[d][t];
    }
    return true;
}

bool PolicyImprovement::computeRolloutControlCosts()
{
    for (int r = 0; r < num_rollouts_; ++r) computeRolloutControlCosts(rollouts_[r]);
    return true;
}

bool PolicyImprovement::computeRolloutControlCosts(Rollout& rollout)
{
    policy_->computeControlCosts(rollout.parameters_, rollout.noise_projected_, control_cost_weight_,
                                 rollout.control_costs_, dense_control_costs_);
    return true;
}

// :451-484
bool PolicyImprovement::computeRolloutCumulativeCosts(Rollout& rollout)
{
    const int T = num_time_steps_;
    double state_cost = 0.0;
    for (int t = 0; t < T; ++t) state_cost += rollout.state_costs_[t];
    double cost = state_cost;
    for (int d = 0; d < num_dimensions_; ++d) {
        double cc_sum = 0.0;
        for (int t = 0; t < T; ++t) cc_sum += rollout.control_costs_[d][t];
        rollout.full_costs_[d] = state_cost + cc_sum;
        cost += cc_sum;
    }
    rollout.total_cost_ = cost;
    for (int d = 0; d < num_dimensions_; ++d) {
        for (int t = 0; t < T; ++t) rollout.total_costs_[d][t] = rollout.state_costs_[t] + rollout.control_costs_[d][t];
        rollout.cumulative_costs_[d] = rollout.total_costs_[d];
        if (use_cumulative_costs_ && forward_cumulation_) {
            // "this is forward cumulation": the variant the reference keeps commented out at :473-477 (cost-to-go)
            for (int t = T - 2; t >= 0; --t) rollout.cumulative_costs_[d][t] += rollout.cumulative_costs_[d][t + 1];
        } else if (use_cumulative_costs_) {
            double s = 0.0;
            for (int t = 0; t < T; ++t) s += rollout.total_costs_[d][t];
            for (int t = 0; t < T; ++t) rollout.cumulative_costs_[d][t] = 1.0 * s;
        }
    }
    return true;
}

bool PolicyImprovement::computeRolloutCumulativeCosts(Vec& rollout_costs_total)
{
    rollout_costs_total.resize(num_rollouts_);
    for (int r = 0; r < num_rollouts_; ++r) {
        computeRolloutCumulativeCosts(rollouts_[r]);
        rollout_costs_total[r] = rollouts_[r].total_cost_;
    }
    return true;
}

// :497-582
bool PolicyImprovement::computeRolloutProbabilities()
{
    const int T = num_time_steps_;
    for (int d = 0; d < num_dimensions_; ++d) {
        double min_cost = *std::min_element(rollouts_[0].cumulative_costs_[d].begin(), rollouts_[0].cumulative_costs_[d].end());
        double max_cost = *std::max_element(rollouts_[0].cumulative_costs_[d].begin(), rollouts_[0].cumulative_costs_[d].end());
        for (int r = 1; r < num_rollouts_; ++r) {
            double min_r = *std::min_element(rollouts_[r].cumulative_costs_[d].begin(), rollouts_[r].cumulative_costs_[d].end());
            double max_r = *std::max_element(rollouts_[r].cumulative_costs_[d].begin(), rollouts_[r].cumulative_costs_[d].end());
            if (min_cost > min_r) min_cost = min_r;
            if (max_cost < max_r) max_cost = max_r;
        }
        for (int t = 0; t < T; ++t) {
            if (per_timestep_minmax_) {   // the variant the reference keeps commented out at :518-528
                min_cost = rollouts_[0].cumulative_costs_[d][t];
                max_cost = min_cost;
                for (int r = 1; r < num_rollouts_; ++r) {
                    double c = rollouts_[r].cumulative_costs_[d][t];
                    if (c < min_cost) min_cost = c;
                    if (c > max_cost) max_cost = c;
                }
            }
            double denom = max_cost - min_cost;
            time_step_weights_[d][t] = 1.0;
            if (denom < 1e-8) denom = 1e-8;
            double p_sum = 0.0;
            for (int r = 0; r < num_rollouts_; ++r) {
                rollouts_[r].probabilities_[d][t] = rollouts_[r].importance_weight_ *
                    std::exp(-cost_scaling_h_ * (rollouts_[r].cumulative_costs_[d][t] - min_cost) / denom);
                p_sum += rollouts_[r].probabilities_[d][t];
            }
            for (int r = 0; r < num_rollouts_; ++r) rollouts_[r].probabilities_[d][t] /= p_sum;
        }
        // "total" probabilities
        min_cost = rollouts_[0].full_costs_[d];
        max_cost = min_cost;
        for (int r = 1; r < num_rollouts_; ++r) {
            double c = rollouts_[r].full_costs_[d];
            if (c < min_cost) min_cost = c;
            if (c > max_cost) max_cost = c;
        }
        double cost_denom = max_cost - min_cost;
        if (cost_denom < 1e-8) cost_denom = 1e-8;
        double p_sum = 0.0;
        for (int r = 0; r < num_rollouts_; ++r) {
            rollouts_[r].full_probabilities_[d] = rollouts_[r].importance_weight_ *
                std::exp(-cost_scaling_h_ * (rollouts_[r].full_costs_[d] - min_cost) / cost_denom);
            p_sum += rollouts_[r].full_probabilities_[d];
        }
        for (int r = 0; r < num_rollouts_; ++r) rollouts_[r].full_probabilities_[d] /= p_sum;
    }
    return true;
}

// :584-711
bool PolicyImprovement::computeParameterUpdates()
{
    const int T = num_time_steps_;
    for (int d = 0; d < num_dimensions_; ++d) {
        parameter_updates_[d] = Mat(T, T);
        for (int r = 0; r < num_rollouts_; ++r)
            for (int t = 0; t < T; ++t)
                parameter_updates_[d](0, t) += rollouts_[r].noise_[d][t] * rollouts_[r].probabilities_[d][t];

        if (use_covariance_matrix_adaptation_) {
            double frob_stddev = 0.0, numer = 0.0, denom = 0.0;
            const Mat& Rm = control_costs_[d];
            for (int r = 0; r < num_rollouts_; ++r) {
                denom += rollouts_[r].full_probabilities_[d];
                // noise^T * control_costs * noise  (dense in the reference; R is 9-banded, the zero
                // entries contribute exact zeros)
                const Vec& n = rollouts_[r].noise_[d];
                double q = 0.0;
                if (dense_control_costs_) {
                    for (int i = 0; i < T; ++i) {
                        double s = 0.0;
                        for (int j = 0; j < T; ++j) s += Rm(i, j) * n[j];
                        q += n[i] * s;
                    }
                } else {
                    for (int i = 0; i < T; ++i) {
                        double s = 0.0;
                        int lo = std::max(0, i - DIFF_RULE_LENGTH), hi = std::min(T - 1, i + DIFF_RULE_LENGTH);
                        for (int j = lo; j <= hi; ++j) s += Rm(i, j) * n[j];
                        q += n[i] * s;
                    }
                }
                numer += rollouts_[r].full_probabilities_[d] * q;
            }
            frob_stddev = std::sqrt(numer / (denom * T));
            double update_rate = 0.2;
            adapted_stddevs_[d] = (1.0 - update_rate) * adapted_stddevs_[d] + update_rate * frob_stddev;
            if (adapted_stddevs_[d] < noise_min_stddev_[d]) adapted_stddevs_[d] = noise_min_stddev_[d];
            adapted_covariance_valid_ = true;
        }

        double weight = 0.0, weight_sum = 0.0, max_weight = 0.0;
        for (int t = 0; t < T; ++t) {
            weight = time_step_weights_[d][t];
            weight_sum += weight;
            parameter_updates_[d](0, t) *= weight;
            if (weight > max_weight) max_weight = weight;
        }
        if (weight_sum < 1e-6) weight_sum = 1e-6;
        double divisor = weight_sum / T;
        if (max_weight > divisor) divisor = max_weight;
        for (int t = 0; t < T; ++t) parameter_updates_[d](0, t) /= divisor;

        if (use_projection_) {
            Vec row(T);
            for (int i = 0; i < T; ++i) {
                double s = 0.0;
                for (int j = 0; j < T; ++j) s += projection_matrix_[d](i, j) * parameter_updates_[d](0, j);
                row[i] = s;
            }
            for (int t = 0; t < T; ++t) parameter_updates_[d](0, t) = row[t];
        }
    }
    return true;
}

bool PolicyImprovement::improvePolicy(std::vector<Mat>& parameter_updates)
{
    computeRolloutProbabilities();
    computeParameterUpdates();
    parameter_updates = parameter_updates_;
    return true;
}

// :750-801
bool PolicyImprovement::preComputeProjectionMatrices()
{
    projection_matrix_.resize(num_dimensions_);
    inv_projection_matrix_.resize(num_dimensions_);
    const int T = num_time_steps_;
    if (!use_projection_) {
        for (int d = 0; d < num_dimensions_; ++d) {
            projection_matrix_[d] = Mat::identity(T);
            inv_projection_matrix_[d] = projection_matrix_[d];
        }
        return true;
    }
    for (int d = 0; d < num_dimensions_; ++d) {
        if (d > 0 && inv_control_costs_[d].a == inv_control_costs_[0].a) {
            projection_matrix_[d] = projection_matrix_[0];
            inv_projection_matrix_[d] = inv_projection_matrix_[0];
            continue;
        }
        projection_matrix_[d] = inv_control_costs_[d];
        for (int p = 0; p < T; ++p) {
            double column_max = projection_matrix_[d](p, p);
            double f = 1.0 / (T * column_max);
            for (int i = 0; i < T; ++i) projection_matrix_[d](i, p) *= f;
        }
        inv_projection_matrix_[d] = full_piv_lu_inverse(projection_matrix_[d]);
    }
    return true;
}

bool PolicyImprovement::computeNoise(Rollout& rollout)
{
    for (int d = 0; d < num_dimensions_; ++d)
        for (int t = 0; t < num_time_steps_; ++t)
            rollout.noise_[d][t] = rollout.parameters_noise_[d][t] - rollout.parameters_[d][t];
    return true;
}

// =================================================================================================
// SphereSdfTask
// =================================================================================================

bool SphereSdfTask::stompInitialize()
{
    const int N = stomp_config_.num_time_steps_ + 2 * TRAJECTORY_PADDING;
    derivative_costs_.assign(stomp_config_.num_dimensions_, Mat(N, NUM_DIFF_RULES));
    initial_trajectory_.assign(stomp_config_.num_dimensions_, Vec(N, 0.0));
    for (int d = 0; d < stomp_config_.num_dimensions_; ++d)
        for (int i = 0; i < N; ++i) derivative_costs_[d](i, STOMP_ACCELERATION) = 1.0;
    return true;
}

void SphereSdfTask::updateTrajectory(const Vec& start, const Vec& goal)
{
    const int T = stomp_config_.num_time_steps_, P = TRAJECTORY_PADDING;
    for (int d = 0; d < stomp_config_.num_dimensions_; ++d) {
        for (int i = 0; i < P; ++i) {
            initial_trajectory_[d][i] = 1.0 * start[d] * 1.0;
            initial_trajectory_[d][P + T + i] = 1.0 * goal[d] * 1.0;
        }
        double increment = (goal[d] - start[d]) / (T - 1);
        for (int i = 0; i < T; ++i) initial_trajectory_[d][P + i] = start[d] + (i * increment);
    }
}

void SphereSdfTask::createPolicy()
{
    policy_.reset(new CovariantMovementPrimitive());
    policy_->initialize(stomp_config_.num_time_steps_, stomp_config_.num_dimensions_, stomp_config_.movement_duration_,
                        derivative_costs_, initial_trajectory_);
    policy_->setToMinControlCost();
    initial_trajectory_ = policy_->parameters_all_;
}

void SphereSdfTask::updatePolicy()
{
    policy_.reset(new CovariantMovementPrimitive());
    policy_->initialize(stomp_config_.num_time_steps_, stomp_config_.num_dimensions_, stomp_config_.movement_duration_,
                        derivative_costs_, initial_trajectory_);
    policy_->updateMinControlCostParameters(initial_trajectory_);
}

void SphereSdfTask::sphereCentres(const double* q, double* centres) const
{
    Frame f; frame_identity(f);
    size_t s = 0;
    for (size_t d = 0; d < joints_.size(); ++d) {
        apply_joint(f, joints_[d], q[d]);
        while (s < spheres_.size() && spheres_[s].link == (int)d) {
            sphere_centre(f, spheres_[s], centres + 3 * s);
            ++s;
        }
    }
}

bool SphereSdfTask::stateCollides(const double* q) const
{
    Frame f; frame_identity(f);
    size_t s = 0;
    bool hit = false;
    for (size_t d = 0; d < joints_.size(); ++d) {
        apply_joint(f, joints_[d], q[d]);
        while (s < spheres_.size() && spheres_[s].link == (int)d) {
            double c[3];
            sphere_centre(f, spheres_[s], c);
            if (sphere_collides(sdf_, c, spheres_[s].r)) hit = true;
            ++s;
        }
    }
    if (!self_pairs_.empty()) {
        std::vector<double> centres(3 * spheres_.size());
        sphereCentres(q, centres.data());
        for (const SelfPairSpec& pr : self_pairs_)
            if (self_pair_collides(centres.data(), pr)) hit = true;
    }
    return hit;
}

double SphereSdfTask::statePenetration(const double* q) const
{
    Frame f; frame_identity(f);
    size_t s = 0;
    double pen = 0.0;
    for (size_t d = 0; d < joints_.size(); ++d) {
        apply_joint(f, joints_[d], q[d]);
        while (s < spheres_.size() && spheres_[s].link == (int)d) {
            double c[3];
            sphere_centre(f, spheres_[s], c);
            const double soft = spheres_[s].r + smooth_margin_;
            const double depth = soft - sphere_distance(sdf_, c);
            if (depth > 0.0) pen = pen + depth;
            ++s;
        }
    }
    return smooth_weight_ * pen;
}

// OptimizationTask::getConstrainDifference (OptimizationTask.cpp:218-237), same statement order
double SphereSdfTask::jointConstraintCost(const double* q) const
{
    double constraint_cost = 0.0;
    for (size_t i = 0; i < jc_value_.size(); ++i) {
        const double diff_value = jc_tolerance_[i] - std::fabs(jc_value_[i] - q[i]);
        if (diff_value < 0.0) constraint_cost = constraint_cost + (-1.0 * diff_value);
    }
    return jc_weight_ * constraint_cost;
}

// OptimizationTask.cpp:137-204: the trajectory evaluated is `parameters` (not the projected one);
// validity is overwritten at every timestep and so reports the last timestep only.
bool SphereSdfTask::execute(const std::vector<Vec>& parameters, const std::vector<Vec>& /*projected_parameters*/,
                            Vec& costs, int, int, int, bool& validity)
{
    const int T = stomp_config_.num_time_steps_, D = stomp_config_.num_dimensions_;
    costs.assign(T, 0.0);
    validity = true;
    std::vector<double> q(D);
    for (int t = 0; t < T; ++t) {
        for (int d = 0; d < D; ++d) q[d] = parameters[d][t];
        double collision_cost;
        if (stateCollides(q.data())) { collision_cost = 1.0; validity = false; }
        else { collision_cost = 0.0; validity = true; }
        if (smooth_cost_) collision_cost = statePenetration(q.data());
        costs[t] = collision_cost;
        if (joint_constraint_) costs[t] += jointConstraintCost(q.data());     // computeJointsConstraintCost (:206-216)
    }
    return true;
}

bool SphereSdfTask::filter(std::vector<Vec>& parameters, int, int)
{
    bool filtered = false;
    for (size_t d = 0; d < parameters.size(); ++d)
        for (int t = 0; t < stomp_config_.num_time_steps_; ++t) {
            if (parameters[d][t] < lower_limits_.at(d)) { parameters[d][t] = lower_limits_.at(d); filtered = true; }
            if (parameters[d][t] > upper_limits_.at(d)) { parameters[d][t] = upper_limits_.at(d); filtered = true; }
        }
    return filtered;
}

// =================================================================================================
// Stomp
// =================================================================================================

bool Stomp::initialize(const StompConfig& config, std::shared_ptr<StompTask> task, uint64_t seed)
{
    stomp_config_ = config;
    stomp_task_ = task;
    stomp_task_->getPolicy(policy_);
    stomp_config_.num_time_steps_ = policy_->num_time_steps_;
    control_cost_weight_ = stomp_task_->getControlCostWeight();
    stomp_config_.num_dimensions_ = policy_->num_dimensions_;
    policy_improvement_.initialize(stomp_config_.num_time_steps_, stomp_config_.min_rollouts_, stomp_config_.max_rollouts_,
                                   stomp_config_.num_rollouts_per_iteration_, policy_, stomp_config_.use_noise_adaptation_,
                                   stomp_config_.noise_min_stddev_, control_cost_weight_, seed);
    rollout_costs_ = Mat(stomp_config_.max_rollouts_, stomp_config_.num_time_steps_);
    policy_iteration_counter_ = 0;
#ifdef _OPENMP
    stomp_config_.num_threads_ = omp_get_max_threads();
#else
    stomp_config_.num_threads_ = 1;
#endif
    if (!stomp_config_.use_openmp_) stomp_config_.num_threads_ = 1;   // the reference also calls omp_set_num_threads(1)
    tmp_rollout_cost_.assign(stomp_config_.max_rollouts_, Vec(stomp_config_.num_time_steps_, 0.0));
    best_noiseless_cost_ = std::numeric_limits<double>::max();
    return true;
}

bool Stomp::doGenRollouts(int iteration_number, const NoiseSource& src)
{
    Vec noise(stomp_config_.num_dimensions_);
    for (int i = 0; i < stomp_config_.num_dimensions_; ++i)
        noise[i] = stomp_config_.noise_stddev_[i] * std::pow(stomp_config_.noise_decay_[i], iteration_number - 1);
    policy_improvement_.getRollouts(rollouts_, noise, src);
    bool filtered = false;
    for (size_t r = 0; r < rollouts_.size(); ++r)
        if (stomp_task_->filter(rollouts_[r], (int)r, 0)) filtered = true;
    if (filtered) policy_improvement_.setRollouts(rollouts_);
    policy_improvement_.computeProjectedNoise();
    policy_improvement_.getProjectedRollouts(projected_rollouts_);
    return true;
}

bool Stomp::doExecuteRollouts(int iteration_number)
{
    const int n = (int)rollouts_.size();
    rollout_validity_.assign(n, 0);
#pragma omp parallel for num_threads(stomp_config_.num_threads_) schedule(static)
    for (int r = 0; r < n; ++r) {
        bool validity;
        stomp_task_->execute(rollouts_[r], projected_rollouts_[r], tmp_rollout_cost_[r], iteration_number, r, 0, validity);
        rollout_validity_[r] = validity ? 1 : 0;
    }
    for (int r = 0; r < n; ++r)
        for (int t = 0; t < stomp_config_.num_time_steps_; ++t) rollout_costs_(r, t) = tmp_rollout_cost_[r][t];
    return true;
}

bool Stomp::doUpdate(int)
{
    Vec all_costs;
    policy_improvement_.setRolloutCosts(rollout_costs_, control_cost_weight_, all_costs);
    policy_improvement_.improvePolicy(parameter_updates_);
    policy_->updateParameters(parameter_updates_);
    return true;
}

bool Stomp::doNoiselessRollout(int iteration_number)
{
    policy_->getParameters(parameters_);
    bool validity = false;
    stomp_task_->execute(parameters_, parameters_, tmp_rollout_cost_[0], iteration_number, -1, 0, validity);
    double total_cost;
    policy_improvement_.setNoiselessRolloutCosts(tmp_rollout_cost_[0], total_cost);
    if (total_cost < best_noiseless_cost_) {
        best_noiseless_parameters_ = parameters_;
        best_noiseless_cost_ = total_cost;
    }
    last_noiseless_rollout_valid_ = validity;
    return true;
}

bool Stomp::runSingleIteration(int iteration_number, const NoiseSource& src)
{
    policy_iteration_counter_++;
    doGenRollouts(iteration_number, src);
    doExecuteRollouts(iteration_number);
    doUpdate(iteration_number);
    doNoiselessRollout(iteration_number);
    return true;
}

}  // namespace oracle
