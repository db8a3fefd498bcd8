// ORACLE / TEST INFRASTRUCTURE ONLY.
#ifndef STOMP_B200_ORACLE_BOOST_NORMAL_SHIM
#define STOMP_B200_ORACLE_BOOST_NORMAL_SHIM
#include <random>
namespace boost {
template <class T = double> struct normal_distribution : std::normal_distribution<T> {
    normal_distribution(T m = 0, T s = 1) : std::normal_distribution<T>(m, s) {}
};
}
#endif
