// YAML -> stomp::StompConfig / stomp::DebugConfig, same keys as the reference
// (reference src/planners/src/wrappers/stomp/HandleStompConfig.cpp:7-63, test/config/stomp.yml).
#pragma once
#include <yaml-cpp/yaml.h>
#include <stomp/StompConfig.hpp>
#include <abstract/AbstractPlanner.hpp>

namespace handle_stomp_config {
stomp::StompConfig getStompConfig(const YAML::Node& yaml_data);
stomp::DebugConfig getDebugConfig(const YAML::Node& yaml_data);
}  // namespace handle_stomp_config
