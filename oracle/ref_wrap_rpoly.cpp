// oracle/ref_wrap_rpoly.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry point around the REFERENCE's own findRootsJenkinsTraub (compiled from
// /root/reference/src/eth_trajectory_generation/rpoly/rpoly_ak1.cpp by oracle/Makefile into oracle/_ref/).
#include <eth_trajectory_generation/rpoly/rpoly_ak1.h>

extern "C" int ref_find_roots_jt(const double* coeffs_increasing, int n, double* re, double* im, int* ok) {
  Eigen::VectorXd c(n);
  for (int i = 0; i < n; ++i) c(i) = coeffs_increasing[i];
  Eigen::VectorXcd roots;
  const bool success = eth_trajectory_generation::findRootsJenkinsTraub(c, &roots);
  *ok = success ? 1 : 0;
  for (long i = 0; i < roots.size(); ++i) {
    re[i] = roots[i].real();
    im[i] = roots[i].imag();
  }
  return (int)roots.size();
}
