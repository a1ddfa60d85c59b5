// tg_common.cuh -- shared definitions for the sm_100a trajectory kernels.
//
// Every numeric routine in csrc/ is written as a TG_HD (host + device) function over raw pointers so that
// (a) nvcc compiles it into the CUDA kernels of libtg_b200.so (the product; -fmad=false), and
// (b) tests/host_emu compiles the very same functions with g++ to check bit-parity against the CPU oracle
//     on machines without a GPU.  (b) is test infrastructure: the shipped library has no CPU path.
//
// Numeric contract: IEEE-754 binary64, no FMA contraction, the operation order written in each
// routine; transcendental functions come from include/tg_detmath.h.
#ifndef TG_COMMON_CUH_
#define TG_COMMON_CUH_

#include <stdint.h>

#include "../../include/tg_detmath.h"

#define TG_N 10     // coefficients per polynomial (reference: node.cpp:1063, PolynomialOptimization<10>)
#define TG_HALF 5   // derivative slots per vertex (lin_impl.h:206)
#define TG_D 4      // x, y, z, heading (node.cpp:902)

// per-(segment, time) record produced by setup_segment_record(), consumed by the solve warp
#define TG_REC_DINV 0    // 5x5 inverse of the lower-right block of A
#define TG_REC_X 25      // 5x5 lower-left block of A^-1 : (-Dinv*C)*diag(1/k!)
#define TG_REC_Q 50      // non-zero block of Q, (N-r) x (N-r), SYMMETRIC bit for bit (B_i B_j commute): packed upper triangle, tg_qtri()
#define TG_REC_H 86      // 10x10 H = (A^-T Q) A^-1
#define TG_REC_SIZE 186  // doubles (1488 B, 16 B aligned)

#define TG_DBL_EPSILON 2.2204460492503131e-16
#define TG_DBL_MIN 2.2250738585072014e-308
#define TG_FLT_MIN 1.17549435082228750797e-38
#define TG_FLT_MAX 3.40282346638528859812e+38
#define TG_DBL_MAX 1.7976931348623157e+308
#define TG_DBL_LOWEST (-1.7976931348623157e+308)
#define TG_PI 3.14159265358979323846

namespace tg {

using tgdm::dabs;
using tgdm::dsqrt;

// base_coefficients_(k, i) = i!/(i-k)!  (eth/polynomial.cpp:155-170); exact small integers
TG_HD double bcoef(int k, int i) {
  int p = 1;
  for (int m = 0; m < k; ++m) p *= (i - m);
  return (double)p;
}

// index of Q[k][b], k <= b < 8, in the packed upper triangle (row k starts at 8k - k(k-1)/2; the same layout for every r)
TG_HD constexpr int tg_qtri(int k, int b) { return k * 8 - (k * (k - 1)) / 2 + (b - k); }
TG_HD constexpr int tg_qsym(int k, int b) { return k <= b ? tg_qtri(k, b) : tg_qtri(b, k); }

TG_HD double dmax(double a, double b) { return (a < b) ? b : a; }  // std::max(a, b)
TG_HD double dmin(double a, double b) { return (b < a) ? b : a; }  // std::min(a, b)
TG_HD int imin(int a, int b) { return a < b ? a : b; }
TG_HD int imax(int a, int b) { return a > b ? a : b; }
TG_HD bool dfinite(double x) { return !(tgdm::disnan(x) || tgdm::disinf(x)); }

// two doubles moved as one 16-byte access (every buffer this is used on is 16-byte aligned)
struct alignas(16) Dbl2 {
  double x, y;
};
TG_HD void store2(double* p, double a, double b) {
  Dbl2 t;
  t.x = a;
  t.y = b;
  *reinterpret_cast<Dbl2*>(p) = t;
}
// Parameters of one batch call; mirrors tg_params in include/tg_b200.h field by field.
struct Params {
  int derivative_to_optimize;
  int max_evals;
  double f_rel, x_rel;
  double limits[9];  // v_h, v_v, a_h, a_v, j_h, j_v, v_hdg, a_hdg, j_hdg
  double dt;
  int check_deviation;
  double max_deviation;
  int max_deviation_iters;
  int first_segment_checked;
  double max_len_factor, min_len_factor;
  int run_time_alloc;
  int override_heading_atan2;
};

}  // namespace tg

#endif  // TG_COMMON_CUH_
