// ORACLE — TEST INFRASTRUCTURE ONLY (see kinematics_spec.hpp for the usage rule).
//
// Plain C++17 CPU restatement of the reference's STOMP rollout loop, written from the reference's
// behaviour with simple ordered loops (no Eigen / Boost: neither exists in this image, and the
// reference itself cannot be compiled here — SURVEY.md §8c).  Each function cites the reference
// file:line it follows (paths relative to /root/reference/).
//
// PARITY PINNED AGAINST THE REFERENCE'S OWN CODE.  The reference ships no golden vectors for this path
//  and its full build needs Rock / Eigen / Boost / robot_model / FCL, none of which are in this
// image — but its STOMP core (src/planners/stomp/src/{Stomp,PolicyImprovement,CovariantMovementPrimitive,
// StompUtils}.cpp + headers) only needs a slice of Eigen and Boost.  oracle/ref/ compiles those UNMODIFIED
// sources where they lie against small stand-ins (oracle/ref/shim) into oracle/_ref/libstomp_ref.so, and
// tests/test_reference_pin.py runs whole solves through both with the same standard normals: every rollout
// field, the update, the parameters, the adapted noise, the noise-less rollout and the stop rule agree to
// 1e-12 relative (bit for bit in practice), with and without rollout reuse, clamping, noise decay, warm start.
// tests/golden/*.npz are vectors written by that build (tests/golden/make_golden.py) for machines without the
// reference.  Further pins: (1) the invariants that follow from the reference code
// (tests/test_oracle_invariants.py), (2) an independently written NumPy restatement (oracle/numpy_ref.py).
// What stays unpinned is the state verdict (FK + collision), see kinematics_spec.hpp.
#pragma once
#include <cstdint>
#include <memory>
#include <random>
#include <utility>
#include <vector>

#include "kinematics_spec.hpp"

namespace oracle {

typedef std::vector<double> Vec;

// -------------------------------------------------------------------------------------------------
// Dense row-major matrix + the two factorisations the path needs
// -------------------------------------------------------------------------------------------------
struct Mat {
    int rows = 0, cols = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
    double& operator()(int i, int j) { return a[(size_t)i * cols + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * cols + j]; }
    static Mat identity(int n) { Mat m(n, n); for (int i = 0; i < n; ++i) m(i, i) = 1.0; return m; }
};

Mat full_piv_lu_inverse(const Mat& A);   // Eigen FullPivLU::inverse(), CovariantMovementPrimitive.cpp:273
Mat llt_lower(const Mat& A);             // Eigen LLT::matrixL(), MultivariateGaussian.hpp:81

// StompUtils.hpp:56-66 / StompUtils.cpp:6-23
static const int DIFF_RULE_LENGTH = 7;
static const int TRAJECTORY_PADDING = DIFF_RULE_LENGTH - 1;
static const int NUM_DIFF_RULES = 4;
extern const double DIFF_RULES[NUM_DIFF_RULES][DIFF_RULE_LENGTH];
enum CostComponents { STOMP_POSITION = 0, STOMP_VELOCITY = 1, STOMP_ACCELERATION = 2, STOMP_JERK = 3 };
void getDifferentiationMatrix(int num_time_steps, int order, double dt, Mat& diff_matrix);

// StompConfig.hpp:19-42
struct StompConfig {
    int num_threads_ = 1;
    int min_rollouts_ = 0, max_rollouts_ = 0, num_rollouts_per_iteration_ = 0;
    int num_time_steps_ = 0, num_dimensions_ = 0, num_iterations_ = 0;
    double movement_duration_ = 0, control_cost_weight_ = 0, delay_per_iteration_ = 0;
    double resolution_ = 0, min_cost_improvement_ = 0;
    Vec noise_stddev_, noise_decay_, noise_min_stddev_;
    bool use_noise_adaptation_ = false;
    bool use_openmp_ = false;
};

// -------------------------------------------------------------------------------------------------
// CovariantMovementPrimitive  (CovariantMovementPrimitive.cpp / .hpp)
// -------------------------------------------------------------------------------------------------
class CovariantMovementPrimitive {
public:
    // :57-74
    bool initialize(int num_time_steps, int num_dimensions, double movement_duration,
                    const std::vector<Mat>& derivative_costs, const std::vector<Vec>& initial_trajectory);
    bool setToMinControlCost();                                                   // :128-132
    bool computeLinearControlCosts();                                             // :136-172
    bool computeMinControlCostParameters();                                       // :174-189
    bool updateMinControlCostParameters(const std::vector<Vec>& parameters_all);  // :191-200
    bool getParameters(std::vector<Vec>& parameters) const;                       // :227-240
    // :327-412 (dense_form=true keeps the reference's O(N^2) mat-vec evaluation; false walks only the
    // non-zero band of the same matrices in the same column order and returns identical doubles)
    bool computeControlCosts(const std::vector<Vec>& parameters, const std::vector<Vec>& noise,
                             double weight, std::vector<Vec>& control_costs, bool dense_form) const;
    bool updateParameters(const std::vector<Mat>& updates);                       // :463-520

    int num_time_steps_ = 0, num_vars_free_ = 0, num_vars_all_ = 0;
    int free_vars_start_index_ = 0, free_vars_end_index_ = 0, num_dimensions_ = 0;
    double movement_duration_ = 0, movement_dt_ = 0;
    std::vector<Mat> derivative_costs_;        // [D] N x 4
    std::vector<Mat> derivative_costs_sqrt_;
    std::vector<Vec> parameters_all_;          // [D] N
    std::vector<Vec> min_control_cost_parameters_all_, min_control_cost_parameters_free_;
    std::vector<Mat> differentiation_matrices_;  // [4] N x N
    std::vector<Mat> control_costs_all_;         // [D] N x N
    std::vector<Mat> control_costs_;             // [D] T x T   (R)
    std::vector<Mat> inv_control_costs_;         // [D] T x T   (R^-1)
    std::vector<Vec> linear_control_costs_;      // [D] T
    Vec constant_control_costs_;
};

// -------------------------------------------------------------------------------------------------
// MultivariateGaussian (MultivariateGaussian.hpp:77-97).  Boost's mt19937 + normal_distribution
// stream is not reproducible without Boost and is seeded from an un-seeded rand() in the reference,
// so parity tests inject the noise; this generator (std::mt19937_64 + Box-Muller) is the default
// source when no noise is injected.
// -------------------------------------------------------------------------------------------------
class MultivariateGaussian {
public:
    MultivariateGaussian() {}
    MultivariateGaussian(const Vec& mean, const Mat& covariance, uint64_t seed);
    static MultivariateGaussian fromFactor(const Vec& mean, const Mat& factor, uint64_t seed);
    void sample(Vec& output);                 // output = mean + L * eps
    void transform(const Vec& eps, Vec& output) const;
    Mat covariance_cholesky_;
private:
    Vec mean_;
    int size_ = 0;
    std::mt19937_64 rng_;
    bool have_spare_ = false;
    double spare_ = 0;
    double normal();
};

// PolicyImprovement.hpp:49-68
struct Rollout {
    std::vector<Vec> parameters_, noise_, noise_projected_, parameters_noise_, parameters_noise_projected_;
    Vec state_costs_;
    std::vector<Vec> control_costs_, total_costs_, cumulative_costs_, probabilities_;
    Vec full_probabilities_, full_costs_;
    double importance_weight_ = 1.0, log_likelihood_ = 0.0, total_cost_ = 0.0;
};

// How the unit noise (the output of MultivariateGaussian::sample, i.e. L*eps with zero mean) is obtained
struct NoiseSource {
    const double* injected = nullptr;  // [num_rollouts_gen][D][T] post-Cholesky noise, or null
    const double* epsilon = nullptr;   // [num_rollouts_gen][D][T] standard normals to push through L, or null
};

// -------------------------------------------------------------------------------------------------
// PolicyImprovement (PolicyImprovement.cpp)
// -------------------------------------------------------------------------------------------------
class PolicyImprovement {
public:
    PolicyImprovement();                                                           // :52-58
    bool initialize(int num_time_steps, int min_rollouts, int max_rollouts, int num_rollouts_per_iteration,
                    std::shared_ptr<CovariantMovementPrimitive> policy, bool use_noise_adaptation,
                    const Vec& noise_min_stddev, double control_cost_weight, uint64_t seed);   // :64-105
    bool setNumRollouts(int min_rollouts, int max_rollouts, int num_rollouts_per_iteration);     // :107-156
    bool getRollouts(std::vector<std::vector<Vec>>& rollouts, const Vec& noise_stddev, const NoiseSource& src);  // :313-328
    bool setRollouts(const std::vector<std::vector<Vec>>& rollouts);              // :342-351
    bool getProjectedRollouts(std::vector<std::vector<Vec>>& rollouts);           // :330-339
    bool computeProjectedNoise();                                                  // :421-428
    bool setRolloutCosts(const Mat& costs, double control_cost_weight, Vec& rollout_costs_total);  // :358-378
    bool setNoiselessRolloutCosts(const Vec& costs, double& total_cost);          // :401-419
    bool improvePolicy(std::vector<Mat>& parameter_updates);                      // :713-729
    void clearReusedRollouts() { num_rollouts_ = 0; }                             // :353-356
    void setCostCumulation(bool b) { use_cumulative_costs_ = b; }                 // :859-862
    void resetAdaptiveNoise() { adapted_covariance_valid_ = false; }              // :864-867

    // state (public: the ctypes getters read it)
    int num_dimensions_ = 0, num_time_steps_ = 0;
    int num_rollouts_ = 0, max_rollouts_ = 0, min_rollouts_ = 0, num_rollouts_per_iteration_ = 0;
    int num_rollouts_gen_ = 0;
    double cost_scaling_h_;
    bool use_cumulative_costs_, use_projection_;
    bool dense_control_costs_ = false;   // evaluation form only, same doubles either way
    bool forward_cumulation_ = false;    // cumulative_costs_ = cost-to-go: the commented-out variant at :473-477 (switch, default off)
    bool per_timestep_minmax_ = false;   // the commented-out variant at :518-528 (switch, default off)
    std::shared_ptr<CovariantMovementPrimitive> policy_;
    std::vector<Mat> control_costs_, inv_control_costs_, projection_matrix_, inv_projection_matrix_;
    double control_cost_weight_ = 0;
    std::vector<Vec> parameters_;
    std::vector<Rollout> rollouts_, reused_rollouts_;
    Rollout noiseless_rollout_;
    bool noiseless_rollout_valid_ = false;
    std::vector<MultivariateGaussian> noise_generators_;
    std::vector<Mat> parameter_updates_;
    std::vector<Vec> time_step_weights_;
    Vec adapted_stddevs_;
    bool adapted_covariance_valid_ = false, use_covariance_matrix_adaptation_ = false;
    Vec noise_min_stddev_;
    std::vector<Vec> tmp_noise_;
    std::vector<std::pair<double, int>> rollout_cost_sorter_;

    bool preComputeProjectionMatrices();                                           // :750-801
    bool computeRolloutControlCosts();                                             // :442-449
    bool computeRolloutControlCosts(Rollout& rollout);                             // :812-817
    bool computeRolloutCumulativeCosts(Rollout& rollout);                          // :451-484
    bool computeRolloutCumulativeCosts(Vec& rollout_costs_total);                  // :486-495
    bool computeRolloutProbabilities();                                            // :497-582
    bool computeParameterUpdates();                                                // :584-711
    bool computeNoise(Rollout& rollout);                                           // :803-810
    bool computeProjectedNoise(Rollout& rollout);                                  // :430-440
    bool generateRollouts(const Vec& noise_stddev, const NoiseSource& src);        // :158-311
};

// -------------------------------------------------------------------------------------------------
// StompTask interface (StompTask.hpp:48-116) and the sphere-vs-SDF task that replaces
// OptimizationTask's external FK + FCL query (OptimizationTask.cpp:137-204)
// -------------------------------------------------------------------------------------------------
class StompTask {
public:
    virtual ~StompTask() {}
    virtual bool execute(const std::vector<Vec>& parameters, const std::vector<Vec>& projected_parameters,
                         Vec& costs, int iteration_number, int rollout_number, int thread_id, bool& validity) = 0;
    virtual bool filter(std::vector<Vec>& parameters, int rollout_id, int thread_id) = 0;
    virtual bool getPolicy(std::shared_ptr<CovariantMovementPrimitive>& policy) = 0;
    virtual double getControlCostWeight() = 0;
};

class SphereSdfTask : public StompTask {
public:
    explicit SphereSdfTask(const StompConfig& config) : stomp_config_(config) {}
    bool stompInitialize();                                                        // OptimizationTask.cpp:22-44
    void updateTrajectory(const Vec& start, const Vec& goal);                      // :46-66
    void createPolicy();                                                           // :108-119
    void updatePolicy();                                                           // :121-135
    bool execute(const std::vector<Vec>& parameters, const std::vector<Vec>& projected_parameters,
                 Vec& costs, int iteration_number, int rollout_number, int thread_id, bool& validity) override;  // :137-204
    bool filter(std::vector<Vec>& parameters, int rollout_id, int thread_id) override;  // :85-106
    bool getPolicy(std::shared_ptr<CovariantMovementPrimitive>& policy) override { policy = policy_; return true; }
    double getControlCostWeight() override { return stomp_config_.control_cost_weight_; }
    // verdict of one joint configuration (what robot_model's updateJointGroup + isStateValid returned)
    bool stateCollides(const double* q) const;
    // Alternative state costs, both off by default:
    //  smooth obstacle cost — the cost of a state grows with the penetration of the link spheres into the clearance band
    //    instead of jumping to 1 (the shape of the non-boolean obstacles of stomp/test/stomp_2d_test.cpp:337-363, carried
    //    over to spheres and a distance field): weight * sum_s max(0, (r_s + margin) - d_s), spheres in index order;
    //  joint-constraint cost — OptimizationTask::computeJointsConstraintCost / getConstrainDifference
    //    (src/wrappers/stomp/OptimizationTask.cpp:206-237; its call at :169-172 is commented out in the reference):
    //    weight * sum_d max(0, |value_d - q_d| - tolerance_d), added to the state cost.
    // Validity stays the binary verdict either way.
    double statePenetration(const double* q) const;
    double jointConstraintCost(const double* q) const;
    bool smooth_cost_ = false;
    double smooth_margin_ = 0.0, smooth_weight_ = 1.0;
    bool joint_constraint_ = false;
    Vec jc_value_, jc_tolerance_;
    double jc_weight_ = 1.0;
    void sphereCentres(const double* q, double* centres /*[S][3]*/) const;

    StompConfig stomp_config_;
    std::shared_ptr<CovariantMovementPrimitive> policy_;
    std::vector<Vec> initial_trajectory_, input_initial_trajectory_;
    std::vector<Mat> derivative_costs_;
    std::vector<JointSpec> joints_;
    std::vector<SphereSpec> spheres_;   // sorted by link
    std::vector<SelfPairSpec> self_pairs_;   // empty: world collisions only
    SdfSpec sdf_ {};
    std::vector<float> sdf_storage_;
    Vec lower_limits_, upper_limits_;
};

// -------------------------------------------------------------------------------------------------
// Stomp loop driver (Stomp.cpp) and the StompPlanner::solve loop (StompPlanner.cpp:65-174)
// -------------------------------------------------------------------------------------------------
class Stomp {
public:
    bool initialize(const StompConfig& config, std::shared_ptr<StompTask> task, uint64_t seed);   // Stomp.cpp:56-94
    bool runSingleIteration(int iteration_number, const NoiseSource& src);                        // :274-301
    bool doGenRollouts(int iteration_number, const NoiseSource& src);                             // :172-204
    bool doExecuteRollouts(int iteration_number);                                                 // :206-229
    bool doUpdate(int iteration_number);                                                          // :239-251
    bool doNoiselessRollout(int iteration_number);                                                // :253-272
    double getNoiselessRolloutTotalCost() { return policy_improvement_.noiseless_rollout_.total_cost_; }

    StompConfig stomp_config_;
    std::shared_ptr<StompTask> stomp_task_;
    std::shared_ptr<CovariantMovementPrimitive> policy_;
    PolicyImprovement policy_improvement_;
    std::vector<Vec> best_noiseless_parameters_;
    double best_noiseless_cost_ = 0;
    bool last_noiseless_rollout_valid_ = false;
    std::vector<std::vector<Vec>> rollouts_, projected_rollouts_;
    std::vector<Mat> parameter_updates_;
    std::vector<Vec> parameters_;
    Mat rollout_costs_;
    std::vector<uint8_t> rollout_validity_;   // per generated rollout, for inspection only
    double control_cost_weight_ = 0;
    std::vector<Vec> tmp_rollout_cost_;
    int policy_iteration_counter_ = 0;
};

}  // namespace oracle
