// ORACLE / TEST INFRASTRUCTURE ONLY — never linked into the product libraries.
//
// Drives the REFERENCE'S OWN, UNMODIFIED STOMP core — /root/reference/src/planners/stomp/src/{Stomp,
// PolicyImprovement,CovariantMovementPrimitive,StompUtils}.cpp and their headers, compiled where they lie against
// the Eigen / Boost stand-ins of oracle/ref/shim — so that the CPU restatement in oracle/stomp_oracle.cpp (and through
// it the CUDA path) is pinned to what the reference's code actually computes: rollout bookkeeping and reuse,
// mean-shifted sampling, control costs, cumulative costs, probabilities, the parameter update, the noise adaptation
// and the noise-less rollout.
//
// What is NOT the reference here, and why:
//  * the task (stomp::StompTask implementation).  The reference's OptimizationTask (src/planners/src/wrappers/stomp/
//    OptimizationTask.cpp) needs robot_model / FCL / KDL, none of which exist in this image; RefTask below restates its
//    stompInitialize (:22-44), updateTrajectory (:46-66), filter (:85-106), createPolicy / updatePolicy (:108-135) and
//    the execute protocol (:137-204: cost 1.0 / 0.0 per time step, validity = last time step), and takes the verdict
//    of a state from the oracle's sphere / SDF task, which is what stands in for robot_model::isStateValid.
//  * Eigen's two factorisations, bound to the oracle's restatements (full-pivot LU inverse, plain LLT).
//  * the standard normals, read from a tape (shim/boost/random/variate_generator.hpp) so that both sides see the same
//    numbers; the driver returns L * eps exactly as MultivariateGaussian::sample formed it.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>

#include <Eigen/Core>
#include <boost/shared_ptr.hpp>
#include <boost/random/variate_generator.hpp>
#include <boost/random/normal_distribution.hpp>
#include <boost/random/mersenne_twister.hpp>

#include "../stomp_oracle.hpp"

// the reference's classes keep everything of interest private; the driver reads it for the comparison
#define private public
#define protected public
#include <stomp/Stomp.hpp>
#undef private
#undef protected

namespace stomp_ref_tape {
std::vector<double> tape;
std::size_t cursor = 0;
}

namespace Eigen {
static oracle::Mat to_oracle(const MatrixXd& a)
{
    oracle::Mat m((int)a.rows(), (int)a.cols());
    for (int i = 0; i < m.rows; ++i) for (int j = 0; j < m.cols; ++j) m(i, j) = a(i, j);
    return m;
}
static MatrixXd from_oracle(const oracle::Mat& m)
{
    MatrixXd a(m.rows, m.cols);
    for (int i = 0; i < m.rows; ++i) for (int j = 0; j < m.cols; ++j) a(i, j) = m(i, j);
    return a;
}
MatrixXd shim_full_piv_lu_inverse(const MatrixXd& a) { return from_oracle(oracle::full_piv_lu_inverse(to_oracle(a))); }
MatrixXd shim_llt_lower(const MatrixXd& a) { return from_oracle(oracle::llt_lower(to_oracle(a))); }
}

extern "C" {
struct oracle_config {   // oracle/oracle_capi.cpp
    int32_t num_time_steps, num_dimensions;
    int32_t min_rollouts, max_rollouts, num_rollouts_per_iteration, num_iterations;
    double movement_duration, control_cost_weight, min_cost_improvement;
    double noise_stddev[32], noise_decay[32], noise_min_stddev[32];
    int32_t use_noise_adaptation, use_openmp;
    int32_t use_cumulative_costs, use_projection, per_timestep_minmax, dense_control_costs;
    uint64_t seed;
};
void* oracle_task(void* hp);
const void* oracle_config_of(void* hp);
}

namespace {

class RefTask : public stomp::StompTask {
public:
    RefTask(const stomp::StompConfig& c, oracle::SphereSdfTask* kin) : stomp_config_(c), kin_(kin) {}

    // OptimizationTask.cpp:22-44
    bool stompInitialize(int, int) override
    {
        const int N = stomp_config_.num_time_steps_ + 2 * stomp::TRAJECTORY_PADDING;
        derivative_costs_.clear();
        derivative_costs_.resize(stomp_config_.num_dimensions_, base::MatrixXd::Zero(N, stomp::NUM_DIFF_RULES));
        initial_trajectory_.clear();
        initial_trajectory_.resize(stomp_config_.num_dimensions_, base::VectorXd::Zero(N));
        for (int d = 0; d < stomp_config_.num_dimensions_; ++d)
            derivative_costs_[d].col(stomp::STOMP_ACCELERATION) = base::VectorXd::Ones(N);
        rollout_validity_.assign(stomp_config_.max_rollouts_ + 1, 0);      // sized here: execute runs under OpenMP
        return true;
    }
    // :46-66
    void updateTrajectory(const double* start, const double* goal)
    {
        for (int d = 0; d < stomp_config_.num_dimensions_; ++d) {
            initial_trajectory_[d].head(stomp::TRAJECTORY_PADDING) = 1.0 * start[d] * base::VectorXd::Ones(stomp::TRAJECTORY_PADDING);
            initial_trajectory_[d].tail(stomp::TRAJECTORY_PADDING) = 1.0 * goal[d] * base::VectorXd::Ones(stomp::TRAJECTORY_PADDING);
            const double increment = (goal[d] - start[d]) / (stomp_config_.num_time_steps_ - 1);
            for (int i = 0; i < stomp_config_.num_time_steps_; i++)
                initial_trajectory_[d]((stomp::TRAJECTORY_PADDING) + i) = start[d] + (i * increment);
        }
    }
    // :108-119
    void createPolicy()
    {
        policy_.reset(new stomp::CovariantMovementPrimitive());
        policy_->initialize(stomp_config_.num_time_steps_, stomp_config_.num_dimensions_, stomp_config_.movement_duration_,
                            derivative_costs_, initial_trajectory_);
        policy_->setToMinControlCost();
        policy_->getParametersAll(initial_trajectory_);
    }
    // :121-135
    void updatePolicy()
    {
        policy_.reset(new stomp::CovariantMovementPrimitive());
        policy_->initialize(stomp_config_.num_time_steps_, stomp_config_.num_dimensions_, stomp_config_.movement_duration_,
                            derivative_costs_, initial_trajectory_);
        policy_->updateMinControlCostParameters(initial_trajectory_);
    }
    // :137-204; the state verdict comes from the oracle's sphere / SDF task
    bool execute(std::vector<base::VectorXd>& parameters, std::vector<base::VectorXd>&, base::VectorXd& costs, base::MatrixXd&,
                 const int, const int rollout_number, int, bool, std::vector<base::VectorXd>&, bool& validity) override
    {
        const int D = stomp_config_.num_dimensions_, T = stomp_config_.num_time_steps_;
        costs = base::VectorXd::Zero(T);
        std::vector<double> q(D);
        validity = true;
        for (int t = 0; t < T; ++t) {
            for (int d = 0; d < D; ++d) q[d] = parameters[d](t);
            if (kin_->stateCollides(q.data())) { costs(t) = 1.0; validity = false; }
            else { costs(t) = 0.0; validity = true; }
        }
        if (rollout_number >= 0 && (size_t)rollout_number < rollout_validity_.size()) rollout_validity_[rollout_number] = validity ? 1 : 0;
        return true;
    }
    // :85-106
    bool filter(std::vector<base::VectorXd>& parameters, int, int) override
    {
        bool filtered = false;
        for (unsigned int d = 0; d < parameters.size(); ++d)
            for (int t = 0; t < stomp_config_.num_time_steps_; ++t) {
                if (parameters[d](t) < kin_->lower_limits_.at(d)) { parameters[d](t) = kin_->lower_limits_.at(d); filtered = true; }
                if (parameters[d](t) > kin_->upper_limits_.at(d)) { parameters[d](t) = kin_->upper_limits_.at(d); filtered = true; }
            }
        return filtered;
    }
    bool getPolicy(boost::shared_ptr<stomp::CovariantMovementPrimitive>& policy) override { policy = policy_; return true; }
    bool setPolicy(const boost::shared_ptr<stomp::CovariantMovementPrimitive> policy) override { policy_ = policy; return true; }
    double getControlCostWeight() override { return stomp_config_.control_cost_weight_; }

    stomp::StompConfig stomp_config_;
    oracle::SphereSdfTask* kin_;
    boost::shared_ptr<stomp::CovariantMovementPrimitive> policy_;
    std::vector<base::MatrixXd> derivative_costs_;
    std::vector<base::VectorXd> initial_trajectory_;
    std::vector<uint8_t> rollout_validity_;
};

struct RefHandle {
    stomp::StompConfig sc;
    std::shared_ptr<RefTask> task;
    std::unique_ptr<stomp::Stomp> stomp;
    std::vector<double> unit_noise;   // [G][D][T] of the last iteration: L * eps as sample() formed it
    double old_cost = 0, cost_improvement = 0, current_cost = 0;
    int num_iterations = 0;
};

}  // namespace

extern "C" {

// the reference configured like `oracle_handle` (whose chain / spheres / SDF / limits it borrows for the task)
void* ref_create(void* oracle_handle)
{
    const oracle_config& c = *static_cast<const oracle_config*>(oracle_config_of(oracle_handle));
    RefHandle* h = new RefHandle();
    stomp::StompConfig& s = h->sc;
    s.num_threads_ = 1;
    s.num_time_steps_ = c.num_time_steps; s.num_dimensions_ = c.num_dimensions;
    s.min_rollouts_ = c.min_rollouts; s.max_rollouts_ = c.max_rollouts; s.num_rollouts_per_iteration_ = c.num_rollouts_per_iteration;
    s.num_iterations_ = c.num_iterations;
    s.movement_duration_ = c.movement_duration; s.control_cost_weight_ = c.control_cost_weight;
    s.delay_per_iteration_ = 0; s.resolution_ = 0; s.min_cost_improvement_ = c.min_cost_improvement;
    s.noise_stddev_.assign(c.noise_stddev, c.noise_stddev + c.num_dimensions);
    s.noise_decay_.assign(c.noise_decay, c.noise_decay + c.num_dimensions);
    s.noise_min_stddev_.assign(c.noise_min_stddev, c.noise_min_stddev + c.num_dimensions);
    s.use_noise_adaptation_ = c.use_noise_adaptation != 0;
    s.use_openmp_ = c.use_openmp != 0;          // Stomp.cpp:80-85,210: OpenMP over the rollouts of Task::execute
    h->task.reset(new RefTask(s, static_cast<oracle::SphereSdfTask*>(oracle_task(oracle_handle))));
    h->task->stompInitialize(1, s.max_rollouts_);
    return h;
}
void ref_destroy(void* hp) { delete static_cast<RefHandle*>(hp); }

// StompPlanner::setStartGoalTrajectory (StompPlanner.cpp:177-184)
int ref_set_start_goal(void* hp, const double* start, const double* goal)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    h->task->updateTrajectory(start, goal);
    h->task->createPolicy();
    return 0;
}

// StompPlanner::updateInitialTrajectory (StompPlanner.cpp:186-208); trajectory is [D][T]
int ref_set_initial_trajectory(void* hp, const double* traj)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    const int D = h->sc.num_dimensions_, T = h->sc.num_time_steps_, P = stomp::TRAJECTORY_PADDING;
    for (int d = 0; d < D; ++d) {
        for (int i = 0; i < P; ++i) {
            h->task->initial_trajectory_[d](i) = traj[(size_t)d * T] * 1.0;
            h->task->initial_trajectory_[d](P + T + i) = traj[(size_t)d * T + T - 1] * 1.0;
        }
        for (int i = 0; i < T; ++i) h->task->initial_trajectory_[d](P + i) = traj[(size_t)d * T + i];
    }
    h->task->updatePolicy();
    return 0;
}

int ref_get_policy(void* hp, double* R, double* Rinv, double* L, double* params_all, double* mincc, double* linear)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    if (!h->task->policy_) return -1;
    stomp::CovariantMovementPrimitive& p = *h->task->policy_;
    const int T = p.num_vars_free_, N = p.num_vars_all_, D = p.num_dimensions_;
    if (R) std::memcpy(R, p.control_costs_[0].data(), sizeof(double) * T * T);
    if (Rinv) std::memcpy(Rinv, p.inv_control_costs_[0].data(), sizeof(double) * T * T);
    if (L) { Eigen::MatrixXd Lm = p.inv_control_costs_[0].llt().matrixL(); std::memcpy(L, Lm.data(), sizeof(double) * T * T); }
    for (int d = 0; d < D; ++d) {
        if (params_all) std::memcpy(params_all + (size_t)d * N, p.parameters_all_[d].data(), sizeof(double) * N);
        if (mincc) std::memcpy(mincc + (size_t)d * T, p.min_control_cost_parameters_free_[d].data(), sizeof(double) * T);
        if (linear) std::memcpy(linear + (size_t)d * T, p.linear_control_costs_[d].data(), sizeof(double) * T);
    }
    return 0;
}

// start of StompPlanner::solve (StompPlanner.cpp:65-73,96-99): a new stomp::Stomp per solve
int ref_begin_solve(void* hp)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    if (!h->task->policy_) return -1;
    h->stomp.reset(new stomp::Stomp());
    h->stomp->initialize(h->sc, h->task);
    h->old_cost = 0; h->cost_improvement = 0; h->current_cost = 0; h->num_iterations = 0;
    return 0;
}

// Stomp::setCostCumulation (Stomp.cpp / PolicyImprovement.cpp:451-495): false = per-time-step costs and probabilities
int ref_set_cost_cumulation(void* hp, int use_cumulative_costs)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    if (!h->stomp) return -1;
    h->stomp->setCostCumulation(use_cumulative_costs != 0);
    return 0;
}

// PolicyImprovement::use_projection_ has no setter in the reference (constructor default false, PolicyImprovement.cpp:57);
// the driver flips the private member and recomputes the projection matrices (:750-801) so that the restatement's
// use_projection branch (M = R^-1 with scaled columns, noise_projected = M noise) is pinned as well.
int ref_set_projection(void* hp, int use_projection)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    if (!h->stomp) return -1;
    h->stomp->policy_improvement_.use_projection_ = use_projection != 0;
    h->stomp->policy_improvement_.preComputeProjectionMatrices();
    return 0;
}

// how many rollouts the next runSingleIteration will generate (PolicyImprovement.cpp:170-186)
int ref_next_num_generated(void* hp)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    const stomp::PolicyImprovement& pi = h->stomp->policy_improvement_;
    int num_rollouts = pi.num_rollouts_, gen = pi.num_rollouts_per_iteration_;
    if (num_rollouts + gen < pi.min_rollouts_) gen = pi.min_rollouts_ - num_rollouts;   // first iteration
    (void)num_rollouts;
    return gen;
}

// one pass of the loop body of StompPlanner::solve (:101-118) with the standard normals `epsilon` [G][D][T]; the
// reference draws them joint by joint, rollout by rollout (PolicyImprovement.cpp:258-286).  Returns 1 when the
// wrapper's stop rule fires.
int ref_iterate(void* hp, int iteration, const double* epsilon, int G)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    if (!h->stomp) return -1;
    const int D = h->sc.num_dimensions_, T = h->sc.num_time_steps_;
    stomp_ref_tape::tape.assign((size_t)G * D * T, 0.0);
    stomp_ref_tape::cursor = 0;
    size_t w = 0;
    for (int d = 0; d < D; ++d)
        for (int r = 0; r < G; ++r)
            for (int t = 0; t < T; ++t) stomp_ref_tape::tape[w++] = epsilon[((size_t)r * D + d) * T + t];
    // L * eps exactly as MultivariateGaussian::sample forms it (same product routine, same L)
    h->unit_noise.assign((size_t)G * D * T, 0.0);
    for (int d = 0; d < D; ++d) {
        const Eigen::MatrixXd& Lm = h->stomp->policy_improvement_.noise_generators_[d].covariance_cholesky_;
        const Eigen::VectorXd& mean = h->stomp->policy_improvement_.noise_generators_[d].mean_;
        for (int r = 0; r < G; ++r) {
            Eigen::VectorXd e(T);
            for (int t = 0; t < T; ++t) e(t) = epsilon[((size_t)r * D + d) * T + t];
            Eigen::VectorXd out = mean + Lm * e;
            for (int t = 0; t < T; ++t) h->unit_noise[((size_t)r * D + d) * T + t] = out(t);
        }
    }
    h->num_iterations++;
    h->stomp->runSingleIteration(iteration);
    const bool tape_used_up = stomp_ref_tape::cursor == stomp_ref_tape::tape.size();
    stomp_ref_tape::tape.clear(); stomp_ref_tape::cursor = 0;
    if (!tape_used_up) return -3;          // the reference drew a different number of normals than G*D*T
    h->current_cost = h->stomp->getNoiselessRolloutTotalCost();
    h->cost_improvement = h->current_cost - h->old_cost;
    h->old_cost = h->current_cost;
    if ((h->current_cost < 1) && (std::fabs(h->cost_improvement) < h->sc.min_cost_improvement_)) return 1;
    return 0;
}

int ref_get_unit_noise(void* hp, double* out) { RefHandle* h = static_cast<RefHandle*>(hp); std::memcpy(out, h->unit_noise.data(), sizeof(double) * h->unit_noise.size()); return 0; }

int ref_num_rollouts(void* hp, int32_t* num_rollouts, int32_t* num_rollouts_gen)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    if (!h->stomp) return -1;
    *num_rollouts = h->stomp->policy_improvement_.num_rollouts_;
    *num_rollouts_gen = h->stomp->policy_improvement_.num_rollouts_gen_;
    return 0;
}

// same field ids as oracle_get_rollout_field
int ref_get_rollout_field(void* hp, int field, double* out)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    if (!h->stomp) return -1;
    const stomp::PolicyImprovement& pi = h->stomp->policy_improvement_;
    const int n = pi.num_rollouts_, D = pi.num_dimensions_, T = pi.num_time_steps_;
    for (int r = 0; r < n; ++r) {
        const stomp::Rollout& ro = pi.rollouts_[r];
        const std::vector<base::VectorXd>* f = nullptr;
        switch (field) {
            case 0: f = &ro.parameters_noise_; break;
            case 1: f = &ro.noise_; break;
            case 2: f = &ro.control_costs_; break;
            case 3: f = &ro.probabilities_; break;
            case 4: f = &ro.cumulative_costs_; break;
            case 5: f = &ro.total_costs_; break;
            case 10: f = &ro.parameters_noise_projected_; break;
            case 11: f = &ro.noise_projected_; break;
            case 6: std::memcpy(out + (size_t)r * T, ro.state_costs_.data(), sizeof(double) * T); continue;
            case 7: std::memcpy(out + (size_t)r * D, ro.full_probabilities_.data(), sizeof(double) * D); continue;
            case 8: std::memcpy(out + (size_t)r * D, ro.full_costs_.data(), sizeof(double) * D); continue;
            case 9: out[r] = ro.total_cost_; continue;
            default: return -2;
        }
        for (int d = 0; d < D; ++d) std::memcpy(out + ((size_t)r * D + d) * T, (*f)[d].data(), sizeof(double) * T);
    }
    return 0;
}

int ref_get_rollout_validity(void* hp, uint8_t* out, int G)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    for (int r = 0; r < G; ++r) out[r] = (size_t)r < h->task->rollout_validity_.size() ? h->task->rollout_validity_[r] : 0;
    return 0;
}

int ref_get_updates(void* hp, double* out)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    if (!h->stomp) return -1;
    const stomp::PolicyImprovement& pi = h->stomp->policy_improvement_;
    for (int d = 0; d < pi.num_dimensions_; ++d)
        for (int t = 0; t < pi.num_time_steps_; ++t) out[(size_t)d * pi.num_time_steps_ + t] = pi.parameter_updates_[d](0, t);
    return 0;
}

int ref_get_parameters(void* hp, double* out)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    std::vector<base::VectorXd> p;
    h->task->policy_->getParameters(p);
    for (size_t d = 0; d < p.size(); ++d) std::memcpy(out + d * p[d].size(), p[d].data(), sizeof(double) * p[d].size());
    return 0;
}

int ref_get_stddevs(void* hp, double* out)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    if (!h->stomp) return -1;
    const std::vector<double>& s = h->stomp->policy_improvement_.adapted_stddevs_;
    std::memcpy(out, s.data(), sizeof(double) * s.size());
    return 0;
}

int ref_get_noiseless(void* hp, double* total_cost, int32_t* valid, double* state_costs, double* control_costs, double* best_cost)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    if (!h->stomp) return -1;
    const stomp::PolicyImprovement& pi = h->stomp->policy_improvement_;
    *total_cost = pi.noiseless_rollout_.total_cost_;
    *valid = h->stomp->last_noiseless_rollout_valid_ ? 1 : 0;
    const int T = pi.num_time_steps_;
    if (state_costs) std::memcpy(state_costs, pi.noiseless_rollout_.state_costs_.data(), sizeof(double) * T);
    if (control_costs)
        for (int d = 0; d < pi.num_dimensions_; ++d)
            std::memcpy(control_costs + (size_t)d * T, pi.noiseless_rollout_.control_costs_[d].data(), sizeof(double) * T);
    if (best_cost) *best_cost = h->stomp->best_noiseless_cost_;
    return 0;
}

// end of StompPlanner::solve (:148-173)
int ref_finish_solve(void* hp, double* solution, int32_t* iterations_used)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    const int D = h->sc.num_dimensions_, T = h->sc.num_time_steps_;
    stomp::CovariantMovementPrimitive& p = *h->task->policy_;
    if (solution)
        for (int d = 0; d < D; ++d)
            for (int i = 0; i < T; ++i) solution[(size_t)d * T + i] = p.parameters_all_[d](i + (stomp::DIFF_RULE_LENGTH - 1));
    if (iterations_used) *iterations_used = h->num_iterations;
    h->stomp.reset();
    return ((h->current_cost < 1) && (std::fabs(h->cost_improvement) <= h->sc.min_cost_improvement_)) ? 1 : 0;
}

// the reference's CPU loop with its own sampler (std::mt19937 behind the boost stand-ins), timed where the reference
// times (MotionPlanners.cpp:506-512 brackets solve()); used by bench.py --impl reference
int ref_solve(void* hp, int max_iterations, int honour_stop, double* seconds_out)
{
    RefHandle* h = static_cast<RefHandle*>(hp);
    if (ref_begin_solve(hp) != 0) return -1;
    stomp_ref_tape::tape.clear(); stomp_ref_tape::cursor = 0;
    const double t0 = omp_get_wtime();
    int it = 0;
    for (; it < max_iterations; ++it) {
        h->num_iterations++;
        h->stomp->runSingleIteration(it);
        h->current_cost = h->stomp->getNoiselessRolloutTotalCost();
        h->cost_improvement = h->current_cost - h->old_cost;
        h->old_cost = h->current_cost;
        if (honour_stop && (h->current_cost < 1) && (std::fabs(h->cost_improvement) < h->sc.min_cost_improvement_)) { ++it; break; }
    }
    if (seconds_out) *seconds_out = omp_get_wtime() - t0;
    return it;
}

}  // extern "C"
