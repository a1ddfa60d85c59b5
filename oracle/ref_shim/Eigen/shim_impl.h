// oracle/ref_shim/Eigen/shim_impl.h -- TEST INFRASTRUCTURE ONLY.
//
// A small stand-in for the subset of Eigen that the reference's eth_trajectory_generation library touches
// (src/eth_trajectory_generation/*.cpp, include/eth_trajectory_generation/impl/*.h, include/eth_mav_msgs/*.h).
// Eigen itself is absent from this image.  With this directory on the include path, oracle/Makefile compiles those
// reference files UNMODIFIED from /root/reference into oracle/_ref/libref_eth.so, so that the CPU restatement in oracle/
// can be checked against the reference's own control flow and scalar code.  It is not a general Eigen replacement:
//   * everything is evaluated eagerly (no expression templates); static sizes are tracked in the types only as far as
//     the reference's code needs them (vector/row-vector/1x1 distinctions, fixed-size blocks);
//   * arithmetic ORDER is this project's numeric contract, the same one the restatement follows: every dense
//     product entry is sum_k a(i,k) b(k,j) accumulated over ascending k starting from the k = 0 term; norms and sums
//     accumulate over ascending index; inverse() is LU with first-maximum row pivoting and per-column substitution.
//     Real Eigen may order some of these sums differently (SIMD packets); that cannot be known without its source.
//   * SparseMatrix is dense-backed with a structural mask; SparseQR::solve is, by default, the band LU of the numeric
//     contract (so that results are bit-comparable), or a dense Householder QR when REF_SHIM_QR_HOUSEHOLDER is defined
//     (used to measure how far LU lands from a QR solve of the same system).
#ifndef ORACLE_REF_SHIM_EIGEN_IMPL_
#define ORACLE_REF_SHIM_EIGEN_IMPL_

#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

typedef std::ptrdiff_t Index;
enum { Dynamic = -1 };
enum NoChange_t { NoChange };
template <class T>
using aligned_allocator = std::allocator<T>;

struct IOFormat {
  int precision;
  std::string coeffSeparator, rowSeparator, rowPrefix, rowSuffix;
  IOFormat(int prec = 6, int = 0, const std::string& cs = " ", const std::string& rs = "\n", const std::string& rp = "", const std::string& rsx = "")
      : precision(prec), coeffSeparator(cs), rowSeparator(rs), rowPrefix(rp), rowSuffix(rsx) {}
};

template <class S, int R, int C>
class Matrix;
template <class S, int R, int C>
class Block;

namespace internal {
template <class T>
struct traits;
template <class S, int R, int C>
struct traits<Matrix<S, R, C>> {
  typedef S Scalar;
  enum { Rows = R, Cols = C };
};
template <class S, int R, int C>
struct traits<Block<S, R, C>> {
  typedef S Scalar;
  enum { Rows = R, Cols = C };
};
constexpr int pick(int a, int b) { return a != Dynamic ? a : b; }
template <class T>
inline T abs_(const T& x) { return x < T(0) ? -x : x; }
inline double real_abs(double x) { return std::fabs(x); }
inline double real_abs(const std::complex<double>& x) { return std::abs(x); }
}  // namespace internal

template <class D, class S>
struct WithFormat {
  const D& m;
  IOFormat f;
};

// ---- read-only interface shared by Matrix and Block -------------------------------------------------------------------
template <class Derived>
class MatrixBase {
 public:
  typedef typename internal::traits<Derived>::Scalar Scalar;
  enum { RowsAtCompileTime = internal::traits<Derived>::Rows, ColsAtCompileTime = internal::traits<Derived>::Cols };
  typedef Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> PlainObject;
  typedef Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> TransposeReturn;
  static constexpr bool kIsVector = (RowsAtCompileTime == 1 || ColsAtCompileTime == 1);
  static constexpr bool kIsRowVector = (RowsAtCompileTime == 1 && ColsAtCompileTime != 1);
  typedef Matrix<Scalar, (kIsRowVector ? 1 : Dynamic), (kIsRowVector ? Dynamic : 1)> SegmentReturn;

  const Derived& derived() const { return *static_cast<const Derived*>(this); }
  Derived& derived() { return *static_cast<Derived*>(this); }
  Index rows() const { return derived().rows_(); }
  Index cols() const { return derived().cols_(); }
  Index size() const { return rows() * cols(); }
  const Scalar& coeff(Index i, Index j) const { return derived().at(i, j); }
  const Scalar& operator()(Index i, Index j) const { return derived().at(i, j); }
  const Scalar& lin(Index i) const { return (cols() == 1) ? derived().at(i, 0) : derived().at(0, i); }
  const Scalar& operator()(Index i) const { return lin(i); }
  const Scalar& operator[](Index i) const { return lin(i); }
  const Scalar& x() const { return lin(0); }
  const Scalar& y() const { return lin(1); }
  const Scalar& z() const { return lin(2); }
  const Scalar& w() const { return lin(3); }

  PlainObject eval() const {
    PlainObject r(rows(), cols());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.at(i, j) = coeff(i, j);
    return r;
  }
  TransposeReturn transpose() const {
    TransposeReturn r(cols(), rows());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.at(j, i) = coeff(i, j);
    return r;
  }
  PlainObject operator-() const {
    PlainObject r(rows(), cols());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.at(i, j) = -coeff(i, j);
    return r;
  }
  // reductions: ascending column-major index, starting from the first element
  Scalar sum() const {
    Scalar s = Scalar(0);
    bool first = true;
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) {
        s = first ? coeff(i, j) : s + coeff(i, j);
        first = false;
      }
    return s;
  }
  Scalar squaredNorm() const {
    Scalar s = Scalar(0);
    bool first = true;
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) {
        const Scalar t = coeff(i, j) * coeff(i, j);
        s = first ? t : s + t;
        first = false;
      }
    return s;
  }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  Scalar trace() const {
    Scalar s = Scalar(0);
    for (Index i = 0; i < std::min(rows(), cols()); ++i) s = (i == 0) ? coeff(i, i) : s + coeff(i, i);
    return s;
  }
  template <class O>
  Scalar dot(const MatrixBase<O>& o) const {
    Scalar s = Scalar(0);
    for (Index i = 0; i < size(); ++i) s = (i == 0) ? lin(i) * o.lin(i) : s + lin(i) * o.lin(i);
    return s;
  }
  template <class O>
  Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const {
    Matrix<Scalar, 3, 1> r;
    r.at(0, 0) = lin(1) * o.lin(2) - lin(2) * o.lin(1);
    r.at(1, 0) = lin(2) * o.lin(0) - lin(0) * o.lin(2);
    r.at(2, 0) = lin(0) * o.lin(1) - lin(1) * o.lin(0);
    return r;
  }
  template <class O>
  PlainObject cwiseProduct(const MatrixBase<O>& o) const {
    PlainObject r(rows(), cols());
    if (o.rows() == rows() && o.cols() == cols()) {
      for (Index j = 0; j < cols(); ++j)
        for (Index i = 0; i < rows(); ++i) r.at(i, j) = coeff(i, j) * o.coeff(i, j);
    } else {  // vectors of the same length
      for (Index i = 0; i < size(); ++i) r.linw(i) = lin(i) * o.lin(i);
    }
    return r;
  }
  PlainObject cwiseInverse() const {
    PlainObject r(rows(), cols());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.at(i, j) = Scalar(1) / coeff(i, j);
    return r;
  }
  PlainObject cwiseAbs() const {
    PlainObject r(rows(), cols());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.at(i, j) = internal::abs_(coeff(i, j));
    return r;
  }
  Scalar maxCoeff() const {
    Scalar m = coeff(0, 0);
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) m = std::max(m, coeff(i, j));
    return m;
  }
  Scalar minCoeff() const {
    Scalar m = coeff(0, 0);
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) m = std::min(m, coeff(i, j));
    return m;
  }
  Matrix<Scalar, internal::pick(RowsAtCompileTime, ColsAtCompileTime), 1> diagonal() const {
    const Index n = std::min(rows(), cols());
    Matrix<Scalar, internal::pick(RowsAtCompileTime, ColsAtCompileTime), 1> r(n, 1);
    for (Index i = 0; i < n; ++i) r.at(i, 0) = coeff(i, i);
    return r;
  }
  // vector -> full square matrix with the vector on the diagonal (the reference only ever multiplies / assigns it)
  Matrix<Scalar, internal::pick(RowsAtCompileTime, ColsAtCompileTime) == 1 ? Dynamic : internal::pick(RowsAtCompileTime, ColsAtCompileTime),
         internal::pick(RowsAtCompileTime, ColsAtCompileTime) == 1 ? Dynamic : internal::pick(RowsAtCompileTime, ColsAtCompileTime)>
  asDiagonal() const {
    const Index n = size();
    Matrix<Scalar, internal::pick(RowsAtCompileTime, ColsAtCompileTime) == 1 ? Dynamic : internal::pick(RowsAtCompileTime, ColsAtCompileTime),
           internal::pick(RowsAtCompileTime, ColsAtCompileTime) == 1 ? Dynamic : internal::pick(RowsAtCompileTime, ColsAtCompileTime)>
        r(n, n);
    r.setZero();
    for (Index i = 0; i < n; ++i) r.at(i, i) = lin(i);
    return r;
  }
  PlainObject reverse() const {
    PlainObject r(rows(), cols());
    const Index n = size();
    for (Index i = 0; i < n; ++i) r.linw(i) = lin(n - 1 - i);
    return r;
  }
  PlainObject normalized() const {
    PlainObject r = eval();
    r.normalize();
    return r;
  }
  bool isZero(double tol = 1e-12) const {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i)
        if (internal::real_abs(coeff(i, j)) > tol) return false;
    return true;
  }
  // const sub-blocks are returned by value
  SegmentReturn head(Index n) const { return segment(0, n); }
  SegmentReturn tail(Index n) const { return segment(size() - n, n); }
  SegmentReturn segment(Index i0, Index n) const {
    SegmentReturn r(kIsRowVector ? 1 : n, kIsRowVector ? n : 1);
    for (Index i = 0; i < n; ++i) r.linw(i) = lin(i0 + i);
    return r;
  }
  template <int N>
  Matrix<Scalar, (kIsRowVector ? 1 : N), (kIsRowVector ? N : 1)> head() const {
    Matrix<Scalar, (kIsRowVector ? 1 : N), (kIsRowVector ? N : 1)> r;
    for (Index i = 0; i < N; ++i) r.linw(i) = lin(i);
    return r;
  }
  template <int N>
  Matrix<Scalar, (kIsRowVector ? 1 : N), (kIsRowVector ? N : 1)> tail() const {
    Matrix<Scalar, (kIsRowVector ? 1 : N), (kIsRowVector ? N : 1)> r;
    for (Index i = 0; i < N; ++i) r.linw(i) = lin(size() - N + i);
    return r;
  }
  Matrix<Scalar, Dynamic, Dynamic> block(Index i0, Index j0, Index nr, Index nc) const {
    Matrix<Scalar, Dynamic, Dynamic> r(nr, nc);
    for (Index j = 0; j < nc; ++j)
      for (Index i = 0; i < nr; ++i) r.at(i, j) = coeff(i0 + i, j0 + j);
    return r;
  }
  template <int NR, int NC>
  Matrix<Scalar, NR, NC> block(Index i0, Index j0) const {
    Matrix<Scalar, NR, NC> r;
    for (Index j = 0; j < NC; ++j)
      for (Index i = 0; i < NR; ++i) r.at(i, j) = coeff(i0 + i, j0 + j);
    return r;
  }
  Matrix<Scalar, 1, ColsAtCompileTime> row(Index i) const {
    Matrix<Scalar, 1, ColsAtCompileTime> r(1, cols());
    for (Index j = 0; j < cols(); ++j) r.at(0, j) = coeff(i, j);
    return r;
  }
  Matrix<Scalar, RowsAtCompileTime, 1> col(Index j) const {
    Matrix<Scalar, RowsAtCompileTime, 1> r(rows(), 1);
    for (Index i = 0; i < rows(); ++i) r.at(i, 0) = coeff(i, j);
    return r;
  }
  // LU with first-maximum row pivoting (right-looking, multipliers by division), then for every column of the identity a
  // forward and a back substitution -- the numeric contract for PartialPivLU::inverse()
  PlainObject inverse() const {
    const Index n = rows();
    std::vector<Scalar> lu((size_t)n * n);
    std::vector<Index> perm(n);
    auto L = [&](Index i, Index j) -> Scalar& { return lu[(size_t)i * n + j]; };
    for (Index i = 0; i < n; ++i) {
      perm[i] = i;
      for (Index j = 0; j < n; ++j) L(i, j) = coeff(i, j);
    }
    for (Index k = 0; k < n; ++k) {
      Index piv = k;
      double best = internal::real_abs(L(k, k));
      for (Index i = k + 1; i < n; ++i) {
        const double a = internal::real_abs(L(i, k));
        if (a > best) {
          best = a;
          piv = i;
        }
      }
      if (piv != k) {
        for (Index j = 0; j < n; ++j) std::swap(L(k, j), L(piv, j));
        std::swap(perm[k], perm[piv]);
      }
      for (Index i = k + 1; i < n; ++i) L(i, k) = L(i, k) / L(k, k);
      for (Index i = k + 1; i < n; ++i)
        for (Index j = k + 1; j < n; ++j) L(i, j) = L(i, j) - L(i, k) * L(k, j);
    }
    PlainObject out(n, n);
    std::vector<Scalar> yv(n);
    for (Index c = 0; c < n; ++c) {
      for (Index i = 0; i < n; ++i) {
        Scalar s = (perm[i] == c) ? Scalar(1) : Scalar(0);
        for (Index j = 0; j < i; ++j) s = s - L(i, j) * yv[j];
        yv[i] = s;
      }
      for (Index i = n - 1; i >= 0; --i) {
        Scalar s = yv[i];
        for (Index j = i + 1; j < n; ++j) s = s - L(i, j) * yv[j];
        yv[i] = s / L(i, i);
      }
      for (Index i = 0; i < n; ++i) out.at(i, c) = yv[i];
    }
    return out;
  }
  Scalar determinant() const {  // cofactor expansion is enough for the 3x3 case the reference touches
    const Index n = rows();
    if (n == 1) return coeff(0, 0);
    if (n == 2) return coeff(0, 0) * coeff(1, 1) - coeff(0, 1) * coeff(1, 0);
    Scalar d = Scalar(0);
    for (Index c = 0; c < n; ++c) {
      Matrix<Scalar, Dynamic, Dynamic> minor(n - 1, n - 1);
      for (Index i = 1; i < n; ++i) {
        Index cc = 0;
        for (Index j = 0; j < n; ++j) {
          if (j == c) continue;
          minor.at(i - 1, cc++) = coeff(i, j);
        }
      }
      const Scalar t = coeff(0, c) * minor.determinant();
      d = (c % 2 == 0) ? d + t : d - t;
    }
    return d;
  }
  WithFormat<Derived, Scalar> format(const IOFormat& f) const { return WithFormat<Derived, Scalar>{derived(), f}; }
};

template <class D>
std::ostream& print_matrix(std::ostream& os, const MatrixBase<D>& m, const IOFormat& f) {
  std::ostringstream ss;
  ss << std::setprecision(f.precision);
  for (Index i = 0; i < m.rows(); ++i) {
    if (i) ss << f.rowSeparator;
    ss << f.rowPrefix;
    for (Index j = 0; j < m.cols(); ++j) {
      if (j) ss << f.coeffSeparator;
      ss << m.coeff(i, j);
    }
    ss << f.rowSuffix;
  }
  return os << ss.str();
}
template <class D>
std::ostream& operator<<(std::ostream& os, const MatrixBase<D>& m) { return print_matrix(os, m, IOFormat()); }
template <class D, class S>
std::ostream& operator<<(std::ostream& os, const WithFormat<D, S>& w) { return print_matrix(os, w.m, w.f); }

// ---- mutable interface ---------------------------------------------------------------------------------------------------
template <class Derived>
class CommaInitializer {
 public:
  typedef typename internal::traits<Derived>::Scalar Scalar;
  CommaInitializer(Derived& m) : m_(m), k_(0) {}
  CommaInitializer& add(const Scalar& s) {
    // row-major fill order, as Eigen's comma initializer
    const Index c = m_.cols();
    m_.at(k_ / c, k_ % c) = s;
    ++k_;
    return *this;
  }
  template <class O>
  CommaInitializer& add_vec(const MatrixBase<O>& v) {  // only vectors stacked into a vector are needed
    for (Index i = 0; i < v.size(); ++i) add(v.lin(i));
    return *this;
  }
  CommaInitializer& operator,(const Scalar& s) { return add(s); }
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value && !std::is_same<T, Scalar>::value>::type>
  CommaInitializer& operator,(const T& s) { return add(Scalar(s)); }
  template <class O>
  CommaInitializer& operator,(const MatrixBase<O>& v) { return add_vec(v); }

 private:
  Derived& m_;
  Index k_;
};

template <class Derived>
class DenseBase : public MatrixBase<Derived> {
 public:
  typedef MatrixBase<Derived> Base;
  typedef typename Base::Scalar Scalar;
  using Base::cols;
  using Base::derived;
  using Base::rows;
  using Base::size;
  using Base::operator();
  using Base::operator[];
  using Base::x;
  using Base::y;
  using Base::z;
  using Base::w;
  using Base::block;
  using Base::col;
  using Base::head;
  using Base::row;
  using Base::segment;
  using Base::tail;
  Scalar& coeffRef(Index i, Index j) { return derived().at(i, j); }
  Scalar& operator()(Index i, Index j) { return derived().at(i, j); }
  Scalar& linw(Index i) { return (cols() == 1) ? derived().at(i, 0) : derived().at(0, i); }
  Scalar& operator()(Index i) { return linw(i); }
  Scalar& operator[](Index i) { return linw(i); }
  Scalar& x() { return linw(0); }
  Scalar& y() { return linw(1); }
  Scalar& z() { return linw(2); }
  Scalar& w() { return linw(3); }
  Derived& setConstant(const Scalar& v) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().at(i, j) = v;
    return derived();
  }
  Derived& setZero() { return setConstant(Scalar(0)); }
  Derived& setOnes() { return setConstant(Scalar(1)); }
  Derived& setIdentity() {
    setZero();
    for (Index i = 0; i < std::min(rows(), cols()); ++i) derived().at(i, i) = Scalar(1);
    return derived();
  }
  // element-wise copy; a vector may be assigned to a vector of the other orientation (Eigen transposes vectors implicitly)
  template <class O>
  void copy_from(const MatrixBase<O>& o) {
    if (o.rows() == rows() && o.cols() == cols()) {
      for (Index j = 0; j < cols(); ++j)
        for (Index i = 0; i < rows(); ++i) derived().at(i, j) = o.coeff(i, j);
    } else {
      assert(o.size() == size() && (rows() == 1 || cols() == 1) && (o.rows() == 1 || o.cols() == 1));
      for (Index i = 0; i < size(); ++i) linw(i) = o.lin(i);
    }
  }
  template <class O>
  Derived& operator+=(const MatrixBase<O>& o) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().at(i, j) = derived().at(i, j) + o.coeff(i, j);
    return derived();
  }
  template <class O>
  Derived& operator-=(const MatrixBase<O>& o) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().at(i, j) = derived().at(i, j) - o.coeff(i, j);
    return derived();
  }
  Derived& operator*=(const Scalar& s) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().at(i, j) = derived().at(i, j) * s;
    return derived();
  }
  Derived& operator/=(const Scalar& s) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().at(i, j) = derived().at(i, j) / s;
    return derived();
  }
  void normalize() {  // Eigen: no-op on a zero vector
    const Scalar n2 = this->squaredNorm();
    if (n2 > Scalar(0)) *this /= std::sqrt(n2);
  }
  CommaInitializer<Derived> operator<<(const Scalar& s) {
    CommaInitializer<Derived> ci(derived());
    ci.add(s);
    return ci;
  }
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value && !std::is_same<T, Scalar>::value>::type>
  CommaInitializer<Derived> operator<<(const T& s) {
    CommaInitializer<Derived> ci(derived());
    ci.add(Scalar(s));
    return ci;
  }
  template <class O>
  CommaInitializer<Derived> operator<<(const MatrixBase<O>& v) {
    CommaInitializer<Derived> ci(derived());
    ci.add_vec(v);
    return ci;
  }
  // writable views
  typedef Block<Scalar, (Base::kIsRowVector ? 1 : Dynamic), (Base::kIsRowVector ? Dynamic : 1)> SegmentView;
  SegmentView head(Index n) { return segment(0, n); }
  SegmentView tail(Index n) { return segment(size() - n, n); }
  SegmentView segment(Index i0, Index n) {
    if (cols() == 1) return SegmentView(&derived().at(i0, 0), n, 1, derived().ld_());
    return SegmentView(&derived().at(0, i0), 1, n, derived().ld_());
  }
  Block<Scalar, Dynamic, Dynamic> block(Index i0, Index j0, Index nr, Index nc) {
    return Block<Scalar, Dynamic, Dynamic>(&derived().at(i0, j0), nr, nc, derived().ld_());
  }
  template <int NR, int NC>
  Block<Scalar, NR, NC> block(Index i0, Index j0) {
    return Block<Scalar, NR, NC>(&derived().at(i0, j0), NR, NC, derived().ld_());
  }
  Block<Scalar, 1, Base::ColsAtCompileTime> row(Index i) { return Block<Scalar, 1, Base::ColsAtCompileTime>(&derived().at(i, 0), 1, cols(), derived().ld_()); }
  Block<Scalar, Base::RowsAtCompileTime, 1> col(Index j) { return Block<Scalar, Base::RowsAtCompileTime, 1>(&derived().at(0, j), rows(), 1, derived().ld_()); }
};

// ---- a view into someone else's column-major storage ---------------------------------------------------------------------
template <class S, int R, int C>
class Block : public DenseBase<Block<S, R, C>> {
 public:
  typedef DenseBase<Block<S, R, C>> Base;
  Block(S* p, Index r, Index c, Index ld) : p_(p), r_(r), c_(c), ld_v(ld) {}
  Block(const Block&) = default;
  Index rows_() const { return r_; }
  Index cols_() const { return c_; }
  Index ld_() const { return ld_v; }
  S& at(Index i, Index j) { return p_[i + j * ld_v]; }
  const S& at(Index i, Index j) const { return p_[i + j * ld_v]; }
  Block& operator=(const Block& o) {
    const Matrix<S, R, C> tmp = o.eval();  // views may alias
    this->copy_from(tmp);
    return *this;
  }
  template <class O>
  Block& operator=(const MatrixBase<O>& o) {
    const typename MatrixBase<O>::PlainObject tmp = o.eval();
    this->copy_from(tmp);
    return *this;
  }

 private:
  S* p_;
  Index r_, c_, ld_v;
};

// 1x1 results convert to their scalar (c^T Q c in computeCost)
template <class Derived, class S, bool OneByOne>
struct ScalarConversion {};
template <class Derived, class S>
struct ScalarConversion<Derived, S, true> {
  operator S() const { return static_cast<const Derived*>(this)->at(0, 0); }
};

// ---- owning matrix -------------------------------------------------------------------------------------------------------
template <class S, int R, int C>
class Matrix : public DenseBase<Matrix<S, R, C>>, public ScalarConversion<Matrix<S, R, C>, S, (R == 1 && C == 1)> {
 public:
  typedef DenseBase<Matrix<S, R, C>> Base;
  typedef S Scalar;
  Matrix() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C), d_((size_t)r_ * c_) {}
  // Matrix(n): vector of size n (dynamic vectors); Matrix(r, c): sizes -- or the two coefficients of a fixed 2-vector
  template <class T, class = typename std::enable_if<std::is_integral<T>::value>::type>
  explicit Matrix(T n) : r_(R == Dynamic ? (C == 1 || C == Dynamic ? (Index)n : 1) : R), c_(C == Dynamic ? (R == 1 ? (Index)n : 1) : C), d_((size_t)r_ * c_) {
    if (R == Dynamic && C == Dynamic) {
      r_ = (Index)n;
      c_ = 1;
      d_.assign((size_t)r_, S());
    }
  }
  Matrix(Index r, Index c) : r_(R == Dynamic ? r : R), c_(C == Dynamic ? c : C), d_((size_t)r_ * c_) {
    static_assert(!(R * C == 2 && R != Dynamic && C != Dynamic), "2-vector coefficient constructor not provided by the stand-in");
  }
  Matrix(const S& a, const S& b, const S& c) : r_(R == Dynamic ? 3 : R), c_(C == Dynamic ? 1 : C), d_(3) {
    d_[0] = a;
    d_[1] = b;
    d_[2] = c;
  }
  Matrix(const S& a, const S& b, const S& c, const S& d) : r_(R == Dynamic ? 4 : R), c_(C == Dynamic ? 1 : C), d_(4) {
    d_[0] = a;
    d_[1] = b;
    d_[2] = c;
    d_[3] = d;
  }
  Matrix(const Matrix&) = default;
  Matrix(Matrix&&) = default;
  template <class O>
  Matrix(const MatrixBase<O>& o) : r_(0), c_(0) {
    assign_from(o);
  }
  Matrix& operator=(const Matrix&) = default;
  Matrix& operator=(Matrix&&) = default;
  template <class O>
  Matrix& operator=(const MatrixBase<O>& o) {
    assign_from(o);
    return *this;
  }
  Index rows_() const { return r_; }
  Index cols_() const { return c_; }
  Index ld_() const { return r_; }
  S& at(Index i, Index j) { return d_[(size_t)(i + j * r_)]; }
  const S& at(Index i, Index j) const { return d_[(size_t)(i + j * r_)]; }
  S* data() { return d_.data(); }
  const S* data() const { return d_.data(); }

  void resize(Index n) {
    if (C == 1 || (R == Dynamic && C == Dynamic)) resize(n, 1);
    else resize(1, n);
  }
  void resize(Index r, Index c) {
    r_ = r;
    c_ = c;
    d_.assign((size_t)r * c, S());
  }
  void resize(Index r, NoChange_t) { resize(r, c_); }
  void resize(NoChange_t, Index c) { resize(r_, c); }
  void conservativeResize(Index n) {
    std::vector<S> old = d_;
    const Index on = (Index)old.size();
    resize(n);
    for (Index i = 0; i < std::min(on, n); ++i) d_[(size_t)i] = old[(size_t)i];
  }

  static Matrix Constant(Index r, Index c, const S& v) {
    Matrix m(r, c);
    m.setConstant(v);
    return m;
  }
  static Matrix Constant(Index n, const S& v) {
    Matrix m;
    m.resize(n);
    m.setConstant(v);
    return m;
  }
  static Matrix Constant(const S& v) {
    Matrix m;
    m.setConstant(v);
    return m;
  }
  static Matrix Zero() { return Constant(S(0)); }
  static Matrix Zero(Index n) { return Constant(n, S(0)); }
  static Matrix Zero(Index r, Index c) { return Constant(r, c, S(0)); }
  static Matrix Ones() { return Constant(S(1)); }
  static Matrix Ones(Index n) { return Constant(n, S(1)); }
  static Matrix Ones(Index r, Index c) { return Constant(r, c, S(1)); }
  static Matrix Identity() {
    Matrix m;
    m.setIdentity();
    return m;
  }
  static Matrix Identity(Index r, Index c) {
    Matrix m(r, c);
    m.setIdentity();
    return m;
  }
  static Matrix Unit(Index k) {
    Matrix m;
    m.setZero();
    m.linw(k) = S(1);
    return m;
  }
  static Matrix UnitX() { return Unit(0); }
  static Matrix UnitY() { return Unit(1); }
  static Matrix UnitZ() { return Unit(2); }

 private:
  template <class O>
  void assign_from(const MatrixBase<O>& o) {
    Index r = o.rows(), c = o.cols();
    // vector <-> row vector: keep this type's orientation
    if ((R == 1 && C != 1 && c == 1 && r != 1) || (C == 1 && R != 1 && r == 1 && c != 1)) std::swap(r, c);
    assert((R == Dynamic || R == r) && (C == Dynamic || C == c));
    std::vector<S> tmp((size_t)r * c);
    if (r == o.rows()) {
      for (Index j = 0; j < c; ++j)
        for (Index i = 0; i < r; ++i) tmp[(size_t)(i + j * r)] = o.coeff(i, j);
    } else {
      for (Index i = 0; i < r * c; ++i) tmp[(size_t)i] = o.lin(i);
    }
    r_ = r;
    c_ = c;
    d_.swap(tmp);
  }
  Index r_, c_;
  std::vector<S> d_;
};

// ---- arithmetic ---------------------------------------------------------------------------------------------------------
template <class A, class B>
Matrix<typename MatrixBase<A>::Scalar, internal::pick(MatrixBase<A>::RowsAtCompileTime, MatrixBase<B>::RowsAtCompileTime),
       internal::pick(MatrixBase<A>::ColsAtCompileTime, MatrixBase<B>::ColsAtCompileTime)>
operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  Matrix<typename MatrixBase<A>::Scalar, internal::pick(MatrixBase<A>::RowsAtCompileTime, MatrixBase<B>::RowsAtCompileTime),
         internal::pick(MatrixBase<A>::ColsAtCompileTime, MatrixBase<B>::ColsAtCompileTime)>
      r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.at(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return r;
}
template <class A, class B>
Matrix<typename MatrixBase<A>::Scalar, internal::pick(MatrixBase<A>::RowsAtCompileTime, MatrixBase<B>::RowsAtCompileTime),
       internal::pick(MatrixBase<A>::ColsAtCompileTime, MatrixBase<B>::ColsAtCompileTime)>
operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  Matrix<typename MatrixBase<A>::Scalar, internal::pick(MatrixBase<A>::RowsAtCompileTime, MatrixBase<B>::RowsAtCompileTime),
         internal::pick(MatrixBase<A>::ColsAtCompileTime, MatrixBase<B>::ColsAtCompileTime)>
      r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.at(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return r;
}
// dense product: entry (i, j) = a(i,0) b(0,j), then + a(i,k) b(k,j) for k = 1, 2, ... (numeric contract)
template <class A, class B>
Matrix<typename MatrixBase<A>::Scalar, MatrixBase<A>::RowsAtCompileTime, MatrixBase<B>::ColsAtCompileTime> operator*(const MatrixBase<A>& a,
                                                                                                                   const MatrixBase<B>& b) {
  typedef typename MatrixBase<A>::Scalar S;
  assert(a.cols() == b.rows());
  Matrix<S, MatrixBase<A>::RowsAtCompileTime, MatrixBase<B>::ColsAtCompileTime> r(a.rows(), b.cols());
  const Index K = a.cols();
  for (Index j = 0; j < b.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) {
      S s = S(0);
      for (Index k = 0; k < K; ++k) s = (k == 0) ? a.coeff(i, 0) * b.coeff(0, j) : s + a.coeff(i, k) * b.coeff(k, j);
      r.at(i, j) = s;
    }
  return r;
}
template <class A>
typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A>& a, const typename MatrixBase<A>::Scalar& s) {
  typename MatrixBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.at(i, j) = a.coeff(i, j) * s;
  return r;
}
template <class A>
typename MatrixBase<A>::PlainObject operator*(const typename MatrixBase<A>::Scalar& s, const MatrixBase<A>& a) {
  typename MatrixBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.at(i, j) = s * a.coeff(i, j);
  return r;
}
template <class A>
typename MatrixBase<A>::PlainObject operator/(const MatrixBase<A>& a, const typename MatrixBase<A>::Scalar& s) {
  typename MatrixBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.at(i, j) = a.coeff(i, j) / s;
  return r;
}
template <class A, class B>
bool operator==(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) return false;
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i)
      if (!(a.coeff(i, j) == b.coeff(i, j))) return false;
  return true;
}
template <class A, class B>
bool operator!=(const MatrixBase<A>& a, const MatrixBase<B>& b) { return !(a == b); }

typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, 1, Dynamic> RowVectorXd;
typedef Matrix<std::complex<double>, Dynamic, 1> VectorXcd;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;

// ---- geometry ------------------------------------------------------------------------------------------------------------
template <class S>
class AngleAxis {
 public:
  AngleAxis() : angle_(0), axis_(Matrix<S, 3, 1>::UnitX()) {}
  template <class O>
  AngleAxis(const S& angle, const MatrixBase<O>& axis) : angle_(angle), axis_(axis) {}
  const S& angle() const { return angle_; }
  const Matrix<S, 3, 1>& axis() const { return axis_; }

 private:
  S angle_;
  Matrix<S, 3, 1> axis_;
};
typedef AngleAxis<double> AngleAxisd;

template <class S>
class Quaternion {
 public:
  Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
  Quaternion(const S& w, const S& x, const S& y, const S& z) : w_(w), x_(x), y_(y), z_(z) {}
  // Eigen/src/Geometry/Quaternion.h: ha = 0.5 * angle; w = cos(ha); vec = sin(ha) * axis
  explicit Quaternion(const AngleAxis<S>& aa) {
    const S ha = S(0.5) * aa.angle();
    w_ = std::cos(ha);
    const S s = std::sin(ha);
    x_ = s * aa.axis().x();
    y_ = s * aa.axis().y();
    z_ = s * aa.axis().z();
  }
  template <class O>
  explicit Quaternion(const MatrixBase<O>& m) { *this = m; }
  // rotation matrix -> quaternion (Shoemake), as Eigen's quaternionbase_assign_impl<3,3>
  template <class O>
  Quaternion& operator=(const MatrixBase<O>& m) {
    S t = m.trace();
    if (t > S(0)) {
      t = std::sqrt(t + S(1.0));
      w_ = S(0.5) * t;
      t = S(0.5) / t;
      x_ = (m.coeff(2, 1) - m.coeff(1, 2)) * t;
      y_ = (m.coeff(0, 2) - m.coeff(2, 0)) * t;
      z_ = (m.coeff(1, 0) - m.coeff(0, 1)) * t;
    } else {
      Index i = 0;
      if (m.coeff(1, 1) > m.coeff(0, 0)) i = 1;
      if (m.coeff(2, 2) > m.coeff(i, i)) i = 2;
      const Index j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m.coeff(i, i) - m.coeff(j, j) - m.coeff(k, k) + S(1.0));
      S v[3];
      v[i] = S(0.5) * t;
      t = S(0.5) / t;
      w_ = (m.coeff(k, j) - m.coeff(j, k)) * t;
      v[j] = (m.coeff(j, i) + m.coeff(i, j)) * t;
      v[k] = (m.coeff(k, i) + m.coeff(i, k)) * t;
      x_ = v[0];
      y_ = v[1];
      z_ = v[2];
    }
    return *this;
  }
  static Quaternion Identity() { return Quaternion(); }
  Quaternion& setIdentity() {
    *this = Quaternion();
    return *this;
  }
  const S& w() const { return w_; }
  const S& x() const { return x_; }
  const S& y() const { return y_; }
  const S& z() const { return z_; }
  S& w() { return w_; }
  S& x() { return x_; }
  S& y() { return y_; }
  S& z() { return z_; }
  S squaredNorm() const { return ((x_ * x_ + y_ * y_) + z_ * z_) + w_ * w_; }
  S norm() const { return std::sqrt(squaredNorm()); }
  void normalize() {
    const S n = norm();
    w_ /= n;
    x_ /= n;
    y_ /= n;
    z_ /= n;
  }
  Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
  Quaternion inverse() const {
    const S n2 = squaredNorm();
    if (n2 > S(0)) return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2);
    return Quaternion(S(0), S(0), S(0), S(0));
  }
  Matrix<S, 3, 3> toRotationMatrix() const {
    Matrix<S, 3, 3> r;
    const S tx = S(2) * x_, ty = S(2) * y_, tz = S(2) * z_;
    const S twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const S txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const S tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    r.at(0, 0) = S(1) - (tyy + tzz);
    r.at(0, 1) = txy - twz;
    r.at(0, 2) = txz + twy;
    r.at(1, 0) = txy + twz;
    r.at(1, 1) = S(1) - (txx + tzz);
    r.at(1, 2) = tyz - twx;
    r.at(2, 0) = txz - twy;
    r.at(2, 1) = tyz + twx;
    r.at(2, 2) = S(1) - (txx + tyy);
    return r;
  }
  Quaternion operator*(const Quaternion& b) const {
    return Quaternion(w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_, w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                      w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_, w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_);
  }
  template <class O>
  Matrix<S, 3, 1> operator*(const MatrixBase<O>& v) const {
    return toRotationMatrix() * v;
  }

 private:
  S w_, x_, y_, z_;
};
typedef Quaternion<double> Quaterniond;
// rotation matrix * quaternion -> rotation matrix (RotationBase's friend operator*)
template <class A, class S>
Matrix<S, 3, 3> operator*(const MatrixBase<A>& m, const Quaternion<S>& q) { return m * q.toRotationMatrix(); }

enum TransformTraits { Isometry = 1, Affine = 2, AffineCompact = 0x10 | Affine, Projective = 0x20 };
template <class S, int Dim, int Mode>
class Transform {
 public:
  Transform() : lin_(Matrix<S, Dim, Dim>::Identity()), t_(Matrix<S, Dim, 1>::Zero()) {}
  static Transform Identity() { return Transform(); }
  Matrix<S, Dim, Dim> rotation() const { return lin_; }
  const Matrix<S, Dim, Dim>& linear() const { return lin_; }
  Matrix<S, Dim, Dim>& linear() { return lin_; }
  const Matrix<S, Dim, 1>& translation() const { return t_; }
  Matrix<S, Dim, 1>& translation() { return t_; }
  template <class O>
  Matrix<S, Dim, 1> operator*(const MatrixBase<O>& v) const {
    return lin_ * v + t_;
  }

 private:
  Matrix<S, Dim, Dim> lin_;
  Matrix<S, Dim, 1> t_;
};
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<double, 3, Isometry> Isometry3d;

// ---- sparse (dense-backed) -------------------------------------------------------------------------------------------------
template <class S>
class Triplet {
 public:
  Triplet() : r_(0), c_(0), v_(0) {}
  Triplet(Index r, Index c, const S& v = S(0)) : r_(r), c_(c), v_(v) {}
  Index row() const { return r_; }
  Index col() const { return c_; }
  const S& value() const { return v_; }

 private:
  Index r_, c_;
  S v_;
};

template <class S>
class SparseMatrix {
 public:
  typedef S Scalar;
  SparseMatrix() : r_(0), c_(0) {}
  SparseMatrix(Index r, Index c) : r_(r), c_(c), v_((size_t)r * c, S(0)), nz_((size_t)r * c, 0) {}
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index nonZeros() const { return (Index)std::count(nz_.begin(), nz_.end(), (unsigned char)1); }
  void resize(Index r, Index c) { *this = SparseMatrix(r, c); }
  const S& val(Index i, Index j) const { return v_[(size_t)(i + j * r_)]; }
  bool has(Index i, Index j) const { return nz_[(size_t)(i + j * r_)] != 0; }
  S coeff(Index i, Index j) const { return val(i, j); }
  void set(Index i, Index j, const S& v) {
    v_[(size_t)(i + j * r_)] = v;
    nz_[(size_t)(i + j * r_)] = 1;
  }
  // duplicates are summed in list order (Eigen: set_from_triplets sums duplicates)
  template <class It>
  void setFromTriplets(It b, It e) {
    std::fill(v_.begin(), v_.end(), S(0));
    std::fill(nz_.begin(), nz_.end(), 0);
    for (It it = b; it != e; ++it) {
      const size_t k = (size_t)(it->row() + it->col() * r_);
      v_[k] = nz_[k] ? v_[k] + it->value() : it->value();
      nz_[k] = 1;
    }
  }
  SparseMatrix transpose() const {
    SparseMatrix t(c_, r_);
    for (Index j = 0; j < c_; ++j)
      for (Index i = 0; i < r_; ++i)
        if (has(i, j)) t.set(j, i, val(i, j));
    return t;
  }
  SparseMatrix operator-() const {
    SparseMatrix t(r_, c_);
    for (Index j = 0; j < c_; ++j)
      for (Index i = 0; i < r_; ++i)
        if (has(i, j)) t.set(i, j, -val(i, j));
    return t;
  }
  SparseMatrix block(Index i0, Index j0, Index nr, Index nc) const {
    SparseMatrix t(nr, nc);
    for (Index j = 0; j < nc; ++j)
      for (Index i = 0; i < nr; ++i)
        if (has(i0 + i, j0 + j)) t.set(i, j, val(i0 + i, j0 + j));
    return t;
  }
  // sparse * sparse: result column j = sum over the structural non-zeros k of rhs column j (ascending k) of lhs column k
  SparseMatrix operator*(const SparseMatrix& b) const {
    assert(c_ == b.r_);
    SparseMatrix t(r_, b.c_);
    for (Index j = 0; j < b.c_; ++j)
      for (Index k = 0; k < c_; ++k) {
        if (!b.has(k, j)) continue;
        const S bk = b.val(k, j);
        for (Index i = 0; i < r_; ++i) {
          if (!has(i, k)) continue;
          const S term = val(i, k) * bk;
          if (t.has(i, j)) t.set(i, j, t.val(i, j) + term);
          else t.set(i, j, term);
        }
      }
    return t;
  }
  // sparse * dense vector / matrix: res = 0, then for ascending column k: res += lhs.col(k) * rhs(k)
  template <class O>
  Matrix<S, Dynamic, MatrixBase<O>::ColsAtCompileTime> operator*(const MatrixBase<O>& x) const {
    assert(c_ == x.rows());
    Matrix<S, Dynamic, MatrixBase<O>::ColsAtCompileTime> r(r_, x.cols());
    r.setZero();
    for (Index j = 0; j < x.cols(); ++j)
      for (Index k = 0; k < c_; ++k)
        for (Index i = 0; i < r_; ++i)
          if (has(i, k)) r.at(i, j) = r.at(i, j) + val(i, k) * x.coeff(k, j);
    return r;
  }
  Matrix<S, Dynamic, Dynamic> toDense() const {
    Matrix<S, Dynamic, Dynamic> m(r_, c_);
    for (Index j = 0; j < c_; ++j)
      for (Index i = 0; i < r_; ++i) m.at(i, j) = val(i, j);
    return m;
  }
  operator Matrix<S, Dynamic, Dynamic>() const { return toDense(); }

 private:
  Index r_, c_;
  std::vector<S> v_;
  std::vector<unsigned char> nz_;
};
template <class S>
std::ostream& operator<<(std::ostream& os, const SparseMatrix<S>& m) { return os << m.toDense(); }

enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
template <class I>
struct COLAMDOrdering {};
template <class I>
struct NaturalOrdering {};
template <class I>
struct AMDOrdering {};

// SparseQR stand-in.  Default: LU without pivoting on the full (two-sided) band, multipliers through the reciprocal pivot,
// back substitution from the far columns inwards, x_i = s * (1 / a_ii) -- the numeric contract of this project's solvers.
// With REF_SHIM_QR_HOUSEHOLDER: dense Householder QR, no column pivoting (an honest QR solve of the same matrix).
template <class MatrixType, class Ordering>
class SparseQR {
 public:
  typedef typename MatrixType::Scalar S;
  SparseQR() : n_(0), hbw_(0) {}
  explicit SparseQR(const MatrixType& a) { compute(a); }
  ComputationInfo info() const { return Success; }
  void compute(const MatrixType& a) {
    n_ = a.rows();
    assert(a.rows() == a.cols());
#ifdef REF_SHIM_QR_HOUSEHOLDER
    qr_.assign((size_t)n_ * n_, S(0));
    beta_.assign((size_t)n_, S(0));
    for (Index j = 0; j < n_; ++j)
      for (Index i = 0; i < n_; ++i) qr_[(size_t)(i + j * n_)] = a.val(i, j);
    for (Index k = 0; k < n_; ++k) {
      S norm2 = S(0);
      for (Index i = k; i < n_; ++i) norm2 += Q(i, k) * Q(i, k);
      const S alpha = (Q(k, k) > S(0)) ? -std::sqrt(norm2) : std::sqrt(norm2);
      if (alpha == S(0)) continue;
      const S v0 = Q(k, k) - alpha;
      S vnorm2 = v0 * v0;
      for (Index i = k + 1; i < n_; ++i) vnorm2 += Q(i, k) * Q(i, k);
      if (vnorm2 == S(0)) continue;
      beta_[(size_t)k] = S(2) / vnorm2;
      // apply H = I - beta v v^T to the trailing columns; v = (v0, Q(k+1..,k))
      for (Index j = k + 1; j < n_; ++j) {
        S dotv = v0 * Q(k, j);
        for (Index i = k + 1; i < n_; ++i) dotv += Q(i, k) * Q(i, j);
        const S f = beta_[(size_t)k] * dotv;
        Q(k, j) -= f * v0;
        for (Index i = k + 1; i < n_; ++i) Q(i, j) -= f * Q(i, k);
      }
      Q(k, k) = alpha;
      v0_.resize((size_t)n_);
      v0_[(size_t)k] = v0;
    }
#else
    hbw_ = 0;
    for (Index j = 0; j < n_; ++j)
      for (Index i = 0; i < n_; ++i)
        if (a.has(i, j)) hbw_ = std::max<Index>(hbw_, i > j ? i - j : j - i);
    lu_.assign((size_t)n_ * n_, S(0));
    rinv_.assign((size_t)n_, S(0));
    for (Index j = 0; j < n_; ++j)
      for (Index i = 0; i < n_; ++i) lu_[(size_t)(i + j * n_)] = a.val(i, j);
    // the multipliers are kept in the strict lower band; the right-hand sides replay them in solve()
    for (Index k = 0; k < n_; ++k) {
      rinv_[(size_t)k] = S(1) / LU(k, k);
      const Index iend = std::min(n_ - 1, k + hbw_);
      for (Index i = k + 1; i <= iend; ++i) {
        const S l = LU(i, k) * rinv_[(size_t)k];
        LU(i, k) = l;
        for (Index j = k + 1; j <= iend; ++j) LU(i, j) = LU(i, j) - l * LU(k, j);
      }
    }
#endif
  }
  template <class O>
  Matrix<S, Dynamic, 1> solve(const MatrixBase<O>& b) const {
    Matrix<S, Dynamic, 1> x(n_, 1);
    std::vector<S> y((size_t)n_);
    for (Index i = 0; i < n_; ++i) y[(size_t)i] = b.lin(i);
#ifdef REF_SHIM_QR_HOUSEHOLDER
    for (Index k = 0; k < n_; ++k) {
      if (beta_[(size_t)k] == S(0)) continue;
      S dotv = v0_[(size_t)k] * y[(size_t)k];
      for (Index i = k + 1; i < n_; ++i) dotv += Q(i, k) * y[(size_t)i];
      const S f = beta_[(size_t)k] * dotv;
      y[(size_t)k] -= f * v0_[(size_t)k];
      for (Index i = k + 1; i < n_; ++i) y[(size_t)i] -= f * Q(i, k);
    }
    for (Index i = n_ - 1; i >= 0; --i) {
      S s = y[(size_t)i];
      for (Index j = i + 1; j < n_; ++j) s -= Q(i, j) * x.at(j, 0);
      x.at(i, 0) = s / Q(i, i);
    }
#else
    for (Index k = 0; k < n_; ++k) {
      const Index iend = std::min(n_ - 1, k + hbw_);
      for (Index i = k + 1; i <= iend; ++i) y[(size_t)i] = y[(size_t)i] - LU(i, k) * y[(size_t)k];
    }
    for (Index i = n_ - 1; i >= 0; --i) {
      S s = y[(size_t)i];
      for (Index j = std::min(n_ - 1, i + hbw_); j > i; --j) s = s - LU(i, j) * x.at(j, 0);
      x.at(i, 0) = s * rinv_[(size_t)i];
    }
#endif
    return x;
  }

 private:
  Index n_, hbw_;
  std::vector<S> lu_, rinv_, qr_, beta_, v0_;
  S& LU(Index i, Index j) { return lu_[(size_t)(i + j * n_)]; }
  const S& LU(Index i, Index j) const { return lu_[(size_t)(i + j * n_)]; }
  S& Q(Index i, Index j) { return qr_[(size_t)(i + j * n_)]; }
  const S& Q(Index i, Index j) const { return qr_[(size_t)(i + j * n_)]; }
};

}  // namespace Eigen
#endif
