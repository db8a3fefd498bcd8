// ORACLE / TEST INFRASTRUCTURE ONLY: boost::mt19937 as std::mt19937.
#ifndef STOMP_B200_ORACLE_BOOST_MT_SHIM
#define STOMP_B200_ORACLE_BOOST_MT_SHIM
#include <random>
namespace boost { typedef std::mt19937 mt19937; }
#endif
