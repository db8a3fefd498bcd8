// Shim for base-logging: stream-style log macros (LOG_DEBUG_S << ...).  Debug output is off unless
// BASE_LOG_LEVEL=DEBUG is set in the environment.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

namespace base_logging {
struct Line {
    std::ostringstream s;
    const char* tag;
    bool on;
    Line(const char* t, bool enabled) : tag(t), on(enabled) {}
    ~Line() { if (on) std::cerr << "[" << tag << "] " << s.str() << std::endl; }
    template <class T> Line& operator<<(const T& v) { if (on) s << v; return *this; }
};
inline bool debug_enabled()
{
    static const bool on = [] { const char* e = std::getenv("BASE_LOG_LEVEL"); return e && std::string(e) == "DEBUG"; }();
    return on;
}
}  // namespace base_logging

#define LOG_DEBUG_S ::base_logging::Line("DEBUG", ::base_logging::debug_enabled())
#define LOG_INFO_S ::base_logging::Line("INFO", ::base_logging::debug_enabled())
#define LOG_WARN_S ::base_logging::Line("WARN", true)
#define LOG_ERROR_S ::base_logging::Line("ERROR", true)
#define LOG_FATAL_S ::base_logging::Line("FATAL", true)
#define LOG_ERROR(...) do { std::fprintf(stderr, "[ERROR] "); std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
