// oracle/ref_shim/boost/algorithm/clamp.hpp -- TEST INFRASTRUCTURE ONLY: boost::algorithm::clamp for
// include/eth_mav_msgs/common.h (boost is absent here; the helper that uses it is not on the path).
#pragma once
namespace boost { namespace algorithm {
template <class T> const T& clamp(const T& v, const T& lo, const T& hi) { return v < lo ? lo : (hi < v ? hi : v); }
} }
