// stomp::StompConfig / stomp::DebugConfig with the reference's fields
// (reference src/planners/stomp/include/stomp/StompConfig.hpp:11-42).
#pragma once
#include <string>
#include <vector>

namespace stomp {

struct DebugConfig {
    std::string output_dir_;
    bool save_noisy_trajectories_ = false;
    bool save_noiseless_trajectories_ = false;
    bool save_cost_function_ = false;
    bool write_to_file_ = false;
};

struct StompConfig {
    int num_threads_ = 1;
    int min_rollouts_ = 0;
    int max_rollouts_ = 0;
    int num_rollouts_per_iteration_ = 0;
    int num_time_steps_ = 0;
    int num_dimensions_ = 0;
    int num_iterations_ = 0;

    double movement_duration_ = 0.0;
    double control_cost_weight_ = 0.0;
    double delay_per_iteration_ = 0.0;
    double resolution_ = 0.0;
    double min_cost_improvement_ = 0.0;

    std::vector<double> noise_stddev_;
    std::vector<double> noise_decay_;
    std::vector<double> noise_min_stddev_;

    bool use_noise_adaptation_ = false;
    bool use_openmp_ = false;     // accepted for compatibility; rollouts run on the GPU

    // additive keys of this build (absent keys keep the defaults)
    int device_ = 0;              // CUDA device ordinal
    unsigned long long seed_ = 2024;
};

}  // namespace stomp
