// oracle/shim/boost/array.hpp — NOT Boost.  TEST INFRASTRUCTURE ONLY (see oracle/shim/eigen3/Eigen/Core).
// boost::array as used by the reference (multirotor_model.hpp:122: at(), copy assignment) is std::array.
#ifndef ORACLE_SHIM_BOOST_ARRAY
#define ORACLE_SHIM_BOOST_ARRAY
#include <array>
#include <cstddef>
namespace boost {
template <class T, std::size_t N>
using array = std::array<T, N>;
}
#endif
