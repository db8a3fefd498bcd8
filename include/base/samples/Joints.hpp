// Shim for base/samples/Joints.hpp (Rock base-types): names + per-joint state, as used by
// AbstractPlanner::setStartGoalTrajectory and test_motion_planners.cpp:69-79.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

namespace base {

struct JointState {
    double position, speed, effort, raw, acceleration;
    JointState() : position(NAN), speed(NAN), effort(NAN), raw(NAN), acceleration(NAN) {}
    static JointState Position(double p) { JointState s; s.position = p; return s; }
};

namespace samples {

struct Joints {
    std::vector<std::string> names;
    std::vector<JointState> elements;

    void resize(size_t n) { names.resize(n); elements.resize(n); }
    void clear() { names.clear(); elements.clear(); }
    size_t size() const { return elements.size(); }
    bool empty() const { return elements.empty(); }
    size_t mapNameToIndex(const std::string& name) const
    {
        for (size_t i = 0; i < names.size(); ++i)
            if (names[i] == name) return i;
        throw std::runtime_error("base::samples::Joints: no joint named " + name);
    }
    const JointState& getElementByName(const std::string& name) const { return elements.at(mapNameToIndex(name)); }
    JointState& getElementByName(const std::string& name) { return elements.at(mapNameToIndex(name)); }
    const JointState& operator[](size_t i) const { return elements.at(i); }
    JointState& operator[](size_t i) { return elements.at(i); }
};

}  // namespace samples
}  // namespace base
