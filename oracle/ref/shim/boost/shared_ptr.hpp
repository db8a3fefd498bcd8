// ORACLE / TEST INFRASTRUCTURE ONLY: boost::shared_ptr as std::shared_ptr.
#ifndef STOMP_B200_ORACLE_BOOST_SHARED_PTR_SHIM
#define STOMP_B200_ORACLE_BOOST_SHARED_PTR_SHIM
#include <memory>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
}
#endif
