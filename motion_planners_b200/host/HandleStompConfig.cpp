// YAML -> StompConfig / DebugConfig (reference src/planners/src/wrappers/stomp/HandleStompConfig.cpp:7-63).
#include <wrapper/stomp/HandleStompConfig.hpp>

namespace handle_stomp_config {

static void read_list(const YAML::Node& node, std::vector<double>& out)
{
    out.resize(node.size());
    for (std::size_t i = 0; i < node.size(); i++) out.at(i) = node[i].as<double>();
}

stomp::StompConfig getStompConfig(const YAML::Node& yaml_data)
{
    using motion_planners::getValue;
    stomp::StompConfig config;
    config.num_threads_ = getValue<int, double>(yaml_data, "num_thread_");
    config.min_rollouts_ = getValue<int, double>(yaml_data, "min_rollouts_");
    config.max_rollouts_ = getValue<int, double>(yaml_data, "max_rollouts_");
    config.num_rollouts_per_iteration_ = getValue<int, double>(yaml_data, "num_rollouts_per_iteration_");
    config.num_time_steps_ = getValue<int, double>(yaml_data, "num_time_steps_");
    config.num_dimensions_ = getValue<int, double>(yaml_data, "num_dimensions_");
    config.num_iterations_ = getValue<int, double>(yaml_data, "num_iterations_");
    read_list(yaml_data["noise_stddev_"], config.noise_stddev_);
    read_list(yaml_data["noise_decay_"], config.noise_decay_);
    read_list(yaml_data["noise_min_stddev_"], config.noise_min_stddev_);
    config.movement_duration_ = getValue<double>(yaml_data, "movement_duration_");
    config.control_cost_weight_ = getValue<double>(yaml_data, "control_cost_weight_");
    config.delay_per_iteration_ = getValue<double>(yaml_data, "delay_per_iteration_");
    config.resolution_ = getValue<double>(yaml_data, "resolution_");
    config.min_cost_improvement_ = getValue<double>(yaml_data, "min_cost_improvement_");
    config.use_noise_adaptation_ = getValue<bool>(yaml_data, "use_noise_adaptation_");
    config.use_openmp_ = getValue<bool>(yaml_data, "use_openmp_");
    // additive keys of this build
    config.device_ = getValue<int>(yaml_data, "device_", 0);
    config.seed_ = (unsigned long long)getValue<double>(yaml_data, "seed_", 2024.0);
    return config;
}

stomp::DebugConfig getDebugConfig(const YAML::Node& yaml_data)
{
    using motion_planners::getValue;
    stomp::DebugConfig config;
    config.output_dir_ = getValue<std::string>(yaml_data, "output_dir_");
    config.save_noisy_trajectories_ = getValue<bool>(yaml_data, "save_noisy_trajectories_");
    config.save_noiseless_trajectories_ = getValue<bool>(yaml_data, "save_noiseless_trajectories_");
    config.save_cost_function_ = getValue<bool>(yaml_data, "save_cost_function_");
    config.write_to_file_ = getValue<bool>(yaml_data, "write_to_file_");
    return config;
}

}  // namespace handle_stomp_config
