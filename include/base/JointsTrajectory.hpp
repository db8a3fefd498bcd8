// Shim for base/JointsTrajectory.hpp (Rock base-types): elements[joint][time step].
#pragma once
#include <string>
#include <vector>

#include <base/samples/Joints.hpp>

namespace base {

struct JointsTrajectory {
    std::vector<std::string> names;
    std::vector<std::vector<JointState> > elements;

    bool empty() const { return elements.empty() || elements.front().empty(); }
    size_t getTimeSteps() const { return elements.empty() ? 0 : elements.front().size(); }
    size_t getNumberOfJoints() const { return elements.size(); }
    void resize(size_t num_joints, size_t num_steps)
    {
        names.resize(num_joints);
        elements.assign(num_joints, std::vector<JointState>(num_steps));
    }
    void clear() { names.clear(); elements.clear(); }
    void getJointsAtTimeStep(size_t step, base::samples::Joints& joints) const
    {
        joints.resize(elements.size());
        joints.names = names;
        for (size_t j = 0; j < elements.size(); ++j) joints.elements[j] = elements[j].at(step);
    }
};

}  // namespace base
