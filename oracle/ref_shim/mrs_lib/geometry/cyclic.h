// oracle/ref_shim/mrs_lib/geometry/cyclic.h -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the part of ctu-mrs/mrs_lib's cyclic.h that the reference's eth/vertex.cpp uses
// (mrs_lib::geometry::radians::dist, vertex.cpp:457,536) plus its siblings, so that the reference file compiles unmodified.
// mrs_lib is a dependency of the reference (package.xml:22) that is not vendored under /root/reference; semantics restated:
// radians live in [0, 2 pi), sradians in [-pi, pi); diff(a, b) = signed shortest way from b to a in [-pi, pi);
// dist = |diff|; unwrap(what, from) = from + diff(what, from); interp(a, b, c) = wrap(a + c diff(b, a)).
#ifndef ORACLE_REF_SHIM_MRS_LIB_CYCLIC_H_
#define ORACLE_REF_SHIM_MRS_LIB_CYCLIC_H_
#include <cmath>
namespace mrs_lib {
namespace geometry {
template <typename flt, class spec>
struct cyclic {
  static constexpr flt minimum() { return spec::minimum; }
  static constexpr flt supremum() { return spec::supremum; }
  static constexpr flt range() { return spec::supremum - spec::minimum; }
  static constexpr flt half_range() { return range() / flt(2); }
  static flt wrap(const flt val) {
    flt rem = std::fmod(val - minimum(), range());
    if (rem < flt(0)) rem += range();
    return rem + minimum();
  }
  static flt diff(const flt minuend, const flt subtrahend) {
    const flt d = minuend - subtrahend;
    if (d < -half_range()) return d + range();
    if (d >= half_range()) return d - range();
    return d;
  }
  static flt dist(const flt from, const flt to) { return std::abs(diff(from, to)); }
  static flt unwrap(const flt what, const flt from) { return from + diff(what, from); }
  static flt interpUnwrapped(const flt from, const flt to, const flt coeff) { return from + coeff * diff(to, from); }
  static flt interp(const flt from, const flt to, const flt coeff) { return wrap(interpUnwrapped(from, to, coeff)); }
};
struct radians : public cyclic<double, radians> {
  static constexpr double minimum = 0;
  static constexpr double supremum = 2 * M_PI;
};
struct sradians : public cyclic<double, sradians> {
  static constexpr double minimum = -M_PI;
  static constexpr double supremum = M_PI;
};
}  // namespace geometry
}  // namespace mrs_lib
#endif
