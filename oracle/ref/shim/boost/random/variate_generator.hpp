// ORACLE / TEST INFRASTRUCTURE ONLY.
// The reference draws its standard normals from boost::variate_generator<mt19937, normal_distribution<>> seeded with
// rand() (MultivariateGaussian.hpp:86-96).  For the parity pin the draws are read from a TAPE that the driver fills
// (so that the oracle can be fed the very same numbers); with an empty tape the generator falls back to std.
#ifndef STOMP_B200_ORACLE_BOOST_VARGEN_SHIM
#define STOMP_B200_ORACLE_BOOST_VARGEN_SHIM
#include <cstddef>
#include <vector>
namespace stomp_ref_tape {
extern std::vector<double> tape;
extern std::size_t cursor;
}
namespace boost {
template <class Engine, class Dist> class variate_generator {
public:
    variate_generator(Engine e, Dist d) : e_(e), d_(d) {}
    double operator()()
    {
        if (stomp_ref_tape::cursor < stomp_ref_tape::tape.size()) return stomp_ref_tape::tape[stomp_ref_tape::cursor++];
        return d_(e_);
    }
private:
    Engine e_;
    Dist d_;
};
}
#endif
