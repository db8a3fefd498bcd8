// Shim: the reference spells the policy pointer boost::shared_ptr (StompTask.hpp:96-103); Boost is absent
// from this image, so the name maps onto the standard one and the signatures stay as written there.
#pragma once
#include <memory>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
template <class T> using enable_shared_from_this = std::enable_shared_from_this<T>;
}  // namespace boost
