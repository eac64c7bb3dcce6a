// ORACLE — TEST INFRASTRUCTURE ONLY. Never included, linked or executed by the product path (hso_b200/).
//
// Stand-in for the Eigen3 headers the reference is written against. Eigen is a system dependency of the reference
// (CMakeLists.txt:54 FIND_PACKAGE(Eigen3 REQUIRED), version unpinned) that is absent from this image and cannot be fetched; this header
// provides the subset of its API that the reference's hot-path translation units use, so that those files — src/CoarseTracker.cpp,
// src/feature_alignment.cpp, src/pose_optimizer.cpp, src/matcher.cpp, src/frame.cpp, src/point.cpp, src/camera.cpp, src/vikit/*.cpp and the
// vendored thirdparty/Sophus — compile UNMODIFIED, from where they lie under /root/reference, into oracle/_ref (see oracle/Makefile).
//
// What it is: dense fixed/dynamic-size matrices with eager evaluation. Every operator evaluates the same scalar expression per coefficient
// that Eigen's expression templates evaluate lazily (same association, same order of the inner sums for products), so results agree with a
// real Eigen build up to the compiler's own floating-point contraction choices. The two non-trivial algorithms are restated from Eigen's
// published sources: LDLT (Eigen/src/Cholesky/LDLT.h: unblocked in-place factorisation with largest-|diagonal| pivoting, solve through
// P^T L^-T D^+ L^-1 P) and the fixed-size inverses (cofactors for 2x2 / 3x3, Eigen/src/LU/InverseImpl.h); Quaternion follows
// Eigen/src/Geometry/Quaternion.h (Hamilton product, _transformVector, toRotationMatrix, conversion from a rotation matrix).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <memory>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_DEFINE_STL_VECTOR_SPECIALIZATION(...)
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3
#define EIGEN_MINOR_VERSION 0
#define EIGEN_ALIGN16 alignas(16)

namespace Eigen {

const int Dynamic = -1;
enum NoChange_t { NoChange };
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum { Lower = 1, Upper = 2 };
typedef std::ptrdiff_t Index;

template <class T> using aligned_allocator = std::allocator<T>;

template <class T, int R, int C, int Opt = 0, int MR = R, int MC = C> class Matrix;
template <class X, int BR, int BC> class Block;
template <class D> struct traits;

template <class T, int R, int C, int Opt, int MR, int MC>
struct traits<Matrix<T, R, C, Opt, MR, MC>> {
  typedef T Scalar;
  enum { Rows = R, Cols = C };
};
template <class X, int BR, int BC>
struct traits<Block<X, BR, BC>> {
  typedef typename traits<X>::Scalar Scalar;
  enum { Rows = BR, Cols = BC };
};

namespace internal {
template <int A, int B> struct pick_dim { enum { value = (A == Dynamic) ? B : A }; };
template <class S> struct is_scalar : std::is_arithmetic<S> {};
template <class A, class B> struct promote { typedef decltype(A() * B()) type; };
}  // namespace internal

template <class Derived> class LDLT_;
template <class T> struct CommaInit;

namespace internal {
// fixed sizes live in the object (like Eigen's DenseStorage), dynamic ones on the heap
template <class T, int N> struct FixedStore {
  T v[N > 0 ? N : 1];
  T& operator[](Index i) { return v[i]; }
  const T& operator[](Index i) const { return v[i]; }
  void assign(size_t, const T&) {}
};
template <class T, int R, int C> struct no_conversion { template <class U> no_conversion(const U&) {} };
template <class T> struct DynStore {
  std::vector<T> v;
  T& operator[](Index i) { return v[(size_t)i]; }
  const T& operator[](Index i) const { return v[(size_t)i]; }
  void assign(size_t n, const T& x) { v.assign(n, x); }
};
}  // namespace internal

// ---- CRTP base: everything readable through rows() / cols() / coeff(i, j) ---------------------------------------------------------
template <class Derived>
class MatrixBase {
 public:
  typedef typename traits<Derived>::Scalar Scalar;
  enum { RowsAtCompileTime = traits<Derived>::Rows, ColsAtCompileTime = traits<Derived>::Cols,
         IsVector = (traits<Derived>::Rows == 1 || traits<Derived>::Cols == 1) };
  typedef Matrix<Scalar, traits<Derived>::Rows, traits<Derived>::Cols> PlainObject;
  typedef Matrix<Scalar, traits<Derived>::Cols, traits<Derived>::Rows> TransposedObject;

  const Derived& derived() const { return *static_cast<const Derived*>(this); }
  Derived& derived() { return *static_cast<Derived*>(this); }
  Index rows() const { return derived().rows_(); }
  Index cols() const { return derived().cols_(); }
  Index size() const { return rows() * cols(); }
  Scalar coeff(Index i, Index j) const { return derived().get(i, j); }
  Scalar& coeffRef(Index i, Index j) { return derived().ref(i, j); }
  Scalar operator()(Index i, Index j) const { return derived().get(i, j); }
  Scalar& operator()(Index i, Index j) { return derived().ref(i, j); }
  // vector access
  Scalar vget(Index i) const { return cols() == 1 ? derived().get(i, 0) : derived().get(0, i); }
  Scalar& vref(Index i) { return cols() == 1 ? derived().ref(i, 0) : derived().ref(0, i); }
  Scalar operator()(Index i) const { return vget(i); }
  Scalar& operator()(Index i) { return vref(i); }
  Scalar operator[](Index i) const { return vget(i); }
  Scalar& operator[](Index i) { return vref(i); }
  Scalar x() const { return vget(0); }
  Scalar y() const { return vget(1); }
  Scalar z() const { return vget(2); }
  Scalar w() const { return vget(3); }
  Scalar& x() { return vref(0); }
  Scalar& y() { return vref(1); }
  Scalar& z() { return vref(2); }
  Scalar& w() { return vref(3); }

  PlainObject eval() const {
    PlainObject r;
    r.resize(rows(), cols());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.ref(i, j) = coeff(i, j);
    return r;
  }
  TransposedObject transpose() const {
    TransposedObject r;
    r.resize(cols(), rows());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.ref(j, i) = coeff(i, j);
    return r;
  }
  TransposedObject adjoint() const { return transpose(); }
  template <class T2>
  Matrix<T2, traits<Derived>::Rows, traits<Derived>::Cols> cast() const {
    Matrix<T2, traits<Derived>::Rows, traits<Derived>::Cols> r;
    r.resize(rows(), cols());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.ref(i, j) = static_cast<T2>(coeff(i, j));
    return r;
  }

  // reductions (Eigen's default traversal: column-major, sequential)
  Scalar sum() const {
    Scalar s = Scalar(0);
    bool first = true;
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) { if (first) { s = coeff(i, j); first = false; } else s += coeff(i, j); }
    return s;
  }
  Scalar squaredNorm() const {
    Scalar s = Scalar(0);
    bool first = true;
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) { const Scalar v = coeff(i, j); if (first) { s = v * v; first = false; } else s += v * v; }
    return s;
  }
  Scalar norm() const { using std::sqrt; return sqrt(squaredNorm()); }
  Scalar trace() const { Scalar s = coeff(0, 0); for (Index i = 1; i < rows(); ++i) s += coeff(i, i); return s; }
  Scalar maxCoeff() const {
    Scalar m = coeff(0, 0);
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (coeff(i, j) > m) m = coeff(i, j);
    return m;
  }
  Scalar minCoeff() const {
    Scalar m = coeff(0, 0);
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (coeff(i, j) < m) m = coeff(i, j);
    return m;
  }
  Scalar mean() const { return sum() / Scalar(size()); }
  PlainObject normalized() const { PlainObject r = eval(); const Scalar n = r.norm(); if (n > Scalar(0)) r /= n; return r; }
  PlainObject cwiseAbs() const { PlainObject r = eval(); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.ref(i, j) = std::abs(r.get(i, j)); return r; }
  template <class O> PlainObject cwiseProduct(const MatrixBase<O>& o) const {
    PlainObject r = eval();
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.ref(i, j) *= o.coeff(i, j);
    return r;
  }
  template <class O> Scalar dot(const MatrixBase<O>& o) const {
    Scalar s = vget(0) * o.vget(0);
    for (Index i = 1; i < size(); ++i) s += vget(i) * o.vget(i);
    return s;
  }
  template <class O> Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const {
    Matrix<Scalar, 3, 1> r;
    r[0] = vget(1) * o.vget(2) - vget(2) * o.vget(1);
    r[1] = vget(2) * o.vget(0) - vget(0) * o.vget(2);
    r[2] = vget(0) * o.vget(1) - vget(1) * o.vget(0);
    return r;
  }
  Scalar value() const { return coeff(0, 0); }
  Scalar determinant() const;
  PlainObject inverse() const;
  LDLT_<PlainObject> ldlt() const;
  bool allFinite() const { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (!std::isfinite(coeff(i, j))) return false; return true; }
  template <class O> bool isApprox(const MatrixBase<O>& o, Scalar prec = Scalar(1e-12)) const {
    Scalar d = 0, a = 0, b = 0;
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) {
      const Scalar e = coeff(i, j) - o.coeff(i, j); d += e * e; a += coeff(i, j) * coeff(i, j); b += o.coeff(i, j) * o.coeff(i, j);
    }
    return d <= prec * prec * std::min(a, b);
  }
  PlainObject array() const { return eval(); }
  PlainObject matrix() const { return eval(); }

  // ---- sub-matrix views (readable and writable) --------------------------------------------------------------------------------------
  template <int BR, int BC> Block<Derived, BR, BC> block(Index i, Index j) { return Block<Derived, BR, BC>(derived(), i, j, BR, BC); }
  template <int BR, int BC> Block<const Derived, BR, BC> block(Index i, Index j) const { return Block<const Derived, BR, BC>(derived(), i, j, BR, BC); }
  Block<Derived, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) { return Block<Derived, Dynamic, Dynamic>(derived(), i, j, r, c); }
  Block<const Derived, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) const { return Block<const Derived, Dynamic, Dynamic>(derived(), i, j, r, c); }
  Block<Derived, traits<Derived>::Rows, 1> col(Index j) { return Block<Derived, traits<Derived>::Rows, 1>(derived(), 0, j, rows(), 1); }
  Block<const Derived, traits<Derived>::Rows, 1> col(Index j) const { return Block<const Derived, traits<Derived>::Rows, 1>(derived(), 0, j, rows(), 1); }
  Block<Derived, 1, traits<Derived>::Cols> row(Index i) { return Block<Derived, 1, traits<Derived>::Cols>(derived(), i, 0, 1, cols()); }
  Block<const Derived, 1, traits<Derived>::Cols> row(Index i) const { return Block<const Derived, 1, traits<Derived>::Cols>(derived(), i, 0, 1, cols()); }
  // vector segments keep the orientation of the vector
  template <int N> Block<Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> segment(Index s) {
    return traits<Derived>::Cols == 1 ? Block<Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)>(derived(), s, 0, N, 1)
                                      : Block<Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)>(derived(), 0, s, 1, N);
  }
  template <int N> Block<const Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> segment(Index s) const {
    return traits<Derived>::Cols == 1 ? Block<const Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)>(derived(), s, 0, N, 1)
                                      : Block<const Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)>(derived(), 0, s, 1, N);
  }
  Block<Derived, Dynamic, Dynamic> segment(Index s, Index n) { return cols() == 1 ? block(s, 0, n, 1) : block(0, s, 1, n); }
  Block<const Derived, Dynamic, Dynamic> segment(Index s, Index n) const { return cols() == 1 ? block(s, 0, n, 1) : block(0, s, 1, n); }
  template <int N> auto head() { return this->template segment<N>(0); }
  template <int N> auto head() const { return this->template segment<N>(0); }
  template <int N> auto tail() { return this->template segment<N>(size() - N); }
  template <int N> auto tail() const { return this->template segment<N>(size() - N); }
  auto head(Index n) { return segment(0, n); }
  auto head(Index n) const { return segment(0, n); }
  auto tail(Index n) { return segment(size() - n, n); }
  auto tail(Index n) const { return segment(size() - n, n); }
  template <int BR, int BC> auto topLeftCorner() { return this->template block<BR, BC>(0, 0); }
  template <int BR, int BC> auto topLeftCorner() const { return this->template block<BR, BC>(0, 0); }
  template <int BR, int BC> auto topRightCorner() { return this->template block<BR, BC>(0, cols() - BC); }
  template <int BR, int BC> auto topRightCorner() const { return this->template block<BR, BC>(0, cols() - BC); }
  template <int BR, int BC> auto bottomLeftCorner() { return this->template block<BR, BC>(rows() - BR, 0); }
  template <int BR, int BC> auto bottomLeftCorner() const { return this->template block<BR, BC>(rows() - BR, 0); }
  template <int BR, int BC> auto bottomRightCorner() { return this->template block<BR, BC>(rows() - BR, cols() - BC); }
  template <int BR, int BC> auto bottomRightCorner() const { return this->template block<BR, BC>(rows() - BR, cols() - BC); }
  auto topLeftCorner(Index r, Index c) { return block(0, 0, r, c); }
  auto topLeftCorner(Index r, Index c) const { return block(0, 0, r, c); }
  auto topRightCorner(Index r, Index c) { return block(0, cols() - c, r, c); }
  auto topRightCorner(Index r, Index c) const { return block(0, cols() - c, r, c); }
  auto bottomLeftCorner(Index r, Index c) { return block(rows() - r, 0, r, c); }
  auto bottomRightCorner(Index r, Index c) { return block(rows() - r, cols() - c, r, c); }
  template <int N> auto leftCols() { return this->template block<traits<Derived>::Rows, N>(0, 0); }
  template <int N> auto leftCols() const { return this->template block<traits<Derived>::Rows, N>(0, 0); }
  template <int N> auto rightCols() { return this->template block<traits<Derived>::Rows, N>(0, cols() - N); }
  template <int N> auto topRows() { return this->template block<N, traits<Derived>::Cols>(0, 0); }
  template <int N> auto topRows() const { return this->template block<N, traits<Derived>::Cols>(0, 0); }
  Matrix<Scalar, Dynamic, 1> diagonal() const;
  Matrix<Scalar, Dynamic, Dynamic> asDiagonal() const {
    Matrix<Scalar, Dynamic, Dynamic> d;
    d.resize(size(), size());
    d.setZero();
    for (Index i = 0; i < size(); ++i) d.ref(i, i) = vget(i);
    return d;
  }
  template <class O> bool operator==(const MatrixBase<O>& o) const {
    if (rows() != o.rows() || cols() != o.cols()) return false;
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (!(coeff(i, j) == o.coeff(i, j))) return false;
    return true;
  }
  template <class O> bool operator!=(const MatrixBase<O>& o) const { return !(*this == o); }

  // ---- in-place arithmetic (works for matrices and views) ------------------------------------------------------------------------------
  template <class O> Derived& operator+=(const MatrixBase<O>& o) { assign_op(o, [](Scalar& a, Scalar b) { a += b; }); return derived(); }
  template <class O> Derived& operator-=(const MatrixBase<O>& o) { assign_op(o, [](Scalar& a, Scalar b) { a -= b; }); return derived(); }
  template <class S, class = typename std::enable_if<internal::is_scalar<S>::value>::type>
  Derived& operator*=(S s) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) *= Scalar(s); return derived(); }
  template <class S, class = typename std::enable_if<internal::is_scalar<S>::value>::type>
  Derived& operator/=(S s) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) /= Scalar(s); return derived(); }
  template <class O> Derived& operator*=(const MatrixBase<O>& o) { PlainObject t = (*this) * o; assign_op(t, [](Scalar& a, Scalar b) { a = b; }); return derived(); }
  Derived& noalias() { return derived(); }
  Derived& setZero() { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = Scalar(0); return derived(); }
  Derived& setOnes() { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = Scalar(1); return derived(); }
  Derived& setConstant(Scalar v) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = v; return derived(); }
  Derived& fill(Scalar v) { return setConstant(v); }
  Derived& setIdentity() { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = (i == j) ? Scalar(1) : Scalar(0); return derived(); }
  void normalize() { const Scalar n = norm(); if (n > Scalar(0)) (*this) /= n; }
  template <class O> void swap(MatrixBase<O>& o) {
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) std::swap(coeffRef(i, j), o.coeffRef(i, j));
  }
  CommaInit<Derived> operator<<(Scalar v);
  template <class O> CommaInit<Derived> operator<<(const MatrixBase<O>& o);

 protected:
  // element-wise update from another expression; a row vector may be assigned to a column vector and vice versa (Eigen transposes vectors)
  template <class O, class F> void assign_op(const MatrixBase<O>& o, F f) {
    if (o.rows() == rows() && o.cols() == cols()) {
      for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) f(coeffRef(i, j), Scalar(o.coeff(i, j)));
    } else {
      assert(o.rows() == cols() && o.cols() == rows() && (rows() == 1 || cols() == 1));
      for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) f(coeffRef(i, j), Scalar(o.coeff(j, i)));
    }
  }
};

// ---- dense storage ------------------------------------------------------------------------------------------------------------------------
template <class T, int R, int C, int Opt, int MR, int MC>
class Matrix : public MatrixBase<Matrix<T, R, C, Opt, MR, MC>> {
  typedef MatrixBase<Matrix<T, R, C, Opt, MR, MC>> Base;
  static constexpr bool kFixed = (R != Dynamic && C != Dynamic);
  static constexpr bool kRowMajor = (Opt & RowMajor) != 0 && !(R != 1 && C == 1);
  typename std::conditional<kFixed, internal::FixedStore<T, (kFixed ? R * C : 1)>, internal::DynStore<T>>::type m_;
  Index r_ = (R == Dynamic ? 0 : R), c_ = (C == Dynamic ? 0 : C);

  template <bool F = kFixed> typename std::enable_if<F>::type init_zero() { for (int k = 0; k < R * C; ++k) m_[k] = T(); }
  template <bool F = kFixed> typename std::enable_if<!F>::type init_zero() {}

 public:
  typedef T Scalar;
  using Base::operator();
  using Base::operator+=;
  using Base::operator-=;

  Matrix() { init_zero(); }
  Matrix(const Matrix& o) = default;
  Matrix& operator=(const Matrix& o) = default;
  // sizes (dynamic) or coefficients (fixed-size vectors)
  explicit Matrix(Index n) { init_zero(); if (!kFixed) { if (C == 1 || (R == Dynamic && C == Dynamic)) resize(n, C == 1 ? 1 : n); else resize(1, n); } }
  Matrix(const T& a, const T& b) { init_two(a, b); }
  template <class A, class B, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value &&
                                                               !(std::is_same<A, T>::value && std::is_same<B, T>::value)>::type>
  Matrix(const A& a, const B& b) { init_two(a, b); }
  Matrix(const T& a, const T& b, const T& c) { init_zero(); static_assert(R * C == 3 || !kFixed, "3 coefficients"); resize(R == 1 ? 1 : 3, R == 1 ? 3 : 1); vec(0) = a; vec(1) = b; vec(2) = c; }
  Matrix(const T& a, const T& b, const T& c, const T& d) { init_zero(); resize(R == 1 ? 1 : 4, R == 1 ? 4 : 1); vec(0) = a; vec(1) = b; vec(2) = c; vec(3) = d; }
  explicit Matrix(const T* data) { init_zero(); for (Index k = 0; k < r_ * c_; ++k) m_[k] = data[k]; }
  template <class O> Matrix(const MatrixBase<O>& o) { init_zero(); assign(o); }
  template <class O> Matrix& operator=(const MatrixBase<O>& o) { assign(o); return *this; }

  // a 1x1 matrix (inner product) converts to its coefficient, as in Eigen
  operator typename std::conditional<(R == 1 && C == 1), T, internal::no_conversion<T, R, C>>::type() const { return m_[0]; }

  Index rows_() const { return R == Dynamic ? r_ : Index(R); }  // compile-time constants for fixed sizes: the loops below unroll
  Index cols_() const { return C == Dynamic ? c_ : Index(C); }
  T get(Index i, Index j) const { return m_[idx(i, j)]; }
  T& ref(Index i, Index j) { return m_[idx(i, j)]; }
  T* data() { return &m_[0]; }
  const T* data() const { return &m_[0]; }
  void resize(Index r, Index c) { resize_impl(r, c); }
  void resize(Index n) { if (C == 1) resize_impl(n, 1); else resize_impl(1, n); }
  void resize(NoChange_t, Index c) { resize_impl(r_, c); }
  void resize(Index r, NoChange_t) { resize_impl(r, c_); }
  void conservativeResize(Index r, Index c) {
    Matrix t; t.resize(r, c);
    for (Index j = 0; j < std::min(c, c_); ++j) for (Index i = 0; i < std::min(r, r_); ++i) t.ref(i, j) = get(i, j);
    *this = t;
  }

  static Matrix Zero() { Matrix r; r.setZero(); return r; }
  static Matrix Zero(Index rr, Index cc) { Matrix r; r.resize(rr, cc); r.setZero(); return r; }
  static Matrix Zero(Index n) { Matrix r(n); r.setZero(); return r; }
  static Matrix Ones() { Matrix r; r.setOnes(); return r; }
  static Matrix Ones(Index rr, Index cc) { Matrix r; r.resize(rr, cc); r.setOnes(); return r; }
  static Matrix Constant(const T& v) { Matrix r; r.setConstant(v); return r; }
  static Matrix Identity() { Matrix r; r.setIdentity(); return r; }
  static Matrix Identity(Index rr, Index cc) { Matrix r; r.resize(rr, cc); r.setIdentity(); return r; }
  static Matrix UnitX() { Matrix r; r.setZero(); r.vec(0) = T(1); return r; }
  static Matrix UnitY() { Matrix r; r.setZero(); r.vec(1) = T(1); return r; }
  static Matrix UnitZ() { Matrix r; r.setZero(); r.vec(2) = T(1); return r; }

 private:
  T& vec(Index i) { return m_[i]; }
  Index idx(Index i, Index j) const { return kRowMajor ? i * cols_() + j : j * rows_() + i; }
  template <class A, class B> void init_two(const A& a, const B& b) {
    init_zero();
    if (kFixed && R * C == 2) { m_[0] = T(a); m_[1] = T(b); }
    else { resize_impl(Index(a), Index(b)); }
  }
  template <bool F = kFixed> typename std::enable_if<F>::type resize_impl(Index r, Index c) { assert(r == R && c == C); (void)r; (void)c; }
  template <bool F = kFixed> typename std::enable_if<!F>::type resize_impl(Index r, Index c) {
    assert((R == Dynamic || r == R) && (C == Dynamic || c == C));
    if (r * c != r_ * c_) m_.assign((size_t)(r * c), T());
    r_ = r; c_ = c;
  }
  template <class O> void assign(const MatrixBase<O>& o) {
    Index orr = o.rows(), oc = o.cols();
    bool tr = false;
    // vectors are transposed on assignment when the orientation differs (Eigen does the same)
    if ((R == 1 && C != 1 && oc == 1 && orr != 1) || (C == 1 && R != 1 && orr == 1 && oc != 1)) { tr = true; std::swap(orr, oc); }
    // the source may be a view of this matrix: evaluate it completely first (on the stack for fixed sizes, like Eigen's own temporaries)
    if (kFixed) {
      T tmp[kFixed ? (R * C > 0 ? R * C : 1) : 1];
      for (Index j = 0; j < oc; ++j) for (Index i = 0; i < orr; ++i) tmp[j * orr + i] = T(tr ? o.coeff(j, i) : o.coeff(i, j));
      for (Index j = 0; j < oc; ++j) for (Index i = 0; i < orr; ++i) ref(i, j) = tmp[j * orr + i];
    } else {
      std::vector<T> tmp((size_t)(orr * oc));
      for (Index j = 0; j < oc; ++j) for (Index i = 0; i < orr; ++i) tmp[(size_t)(j * orr + i)] = T(tr ? o.coeff(j, i) : o.coeff(i, j));
      resize_impl(orr, oc);
      for (Index j = 0; j < oc; ++j) for (Index i = 0; i < orr; ++i) ref(i, j) = tmp[(size_t)(j * orr + i)];
    }
  }
};

// ---- writable view of a rectangular part of a matrix ----------------------------------------------------------------------------------
template <class X, int BR, int BC>
class Block : public MatrixBase<Block<X, BR, BC>> {
  typedef MatrixBase<Block<X, BR, BC>> Base;
  X& x_;
  Index i0_, j0_, r_, c_;

 public:
  typedef typename traits<typename std::remove_const<X>::type>::Scalar Scalar;
  using Base::operator+=;
  using Base::operator-=;
  Block(X& x, Index i0, Index j0, Index r, Index c) : x_(x), i0_(i0), j0_(j0), r_(r), c_(c) {}
  Block(const Block&) = default;
  Index rows_() const { return BR == Dynamic ? r_ : Index(BR); }
  Index cols_() const { return BC == Dynamic ? c_ : Index(BC); }
  Scalar get(Index i, Index j) const { return x_.get(i0_ + i, j0_ + j); }
  Scalar& ref(Index i, Index j) { return const_cast<typename std::remove_const<X>::type&>(x_).ref(i0_ + i, j0_ + j); }
  template <class O> Block& operator=(const MatrixBase<O>& o) {
    typename MatrixBase<O>::PlainObject t = o.eval();  // the source may alias the viewed matrix
    this->assign_op(t, [](Scalar& a, Scalar b) { a = b; });
    return *this;
  }
  Block& operator=(const Block& o) {
    typename Base::PlainObject t = o.eval();
    this->assign_op(t, [](Scalar& a, Scalar b) { a = b; });
    return *this;
  }
};
template <class X, int BR, int BC>
struct traits<Block<const X, BR, BC>> {
  typedef typename traits<X>::Scalar Scalar;
  enum { Rows = BR, Cols = BC };
};

// ---- comma initialiser: m << a, b, c, ... scalars and blocks placed left to right, top to bottom (Eigen/src/Core/CommaInitializer.h) ----
template <class D>
struct CommaInit {
  typedef typename traits<D>::Scalar Scalar;
  D& m;
  Index row, col, block_rows;
  CommaInit(D& mm, Scalar v) : m(mm), row(0), col(0), block_rows(1) { m.coeffRef(0, 0) = v; col = 1; }
  template <class O> CommaInit(D& mm, const MatrixBase<O>& o) : m(mm), row(0), col(0), block_rows(o.rows()) { place(o); }
  template <class O> void place(const MatrixBase<O>& o) {
    if (col == m.cols()) { row += block_rows; col = 0; block_rows = o.rows(); }
    for (Index j = 0; j < o.cols(); ++j) for (Index i = 0; i < o.rows(); ++i) m.coeffRef(row + i, col + j) = o.coeff(i, j);
    col += o.cols();
  }
  template <class S, class = typename std::enable_if<internal::is_scalar<S>::value>::type>
  CommaInit& operator,(S v) {
    if (col == m.cols()) { row += block_rows; col = 0; block_rows = 1; }
    m.coeffRef(row, col) = (Scalar)v;
    ++col;
    return *this;
  }
  template <class O> CommaInit& operator,(const MatrixBase<O>& o) { place(o); return *this; }
};
template <class Derived>
CommaInit<Derived> MatrixBase<Derived>::operator<<(Scalar v) { return CommaInit<Derived>(derived(), v); }
template <class Derived>
template <class O>
CommaInit<Derived> MatrixBase<Derived>::operator<<(const MatrixBase<O>& o) { return CommaInit<Derived>(derived(), o); }

// ---- arithmetic (eager) ---------------------------------------------------------------------------------------------------------------------
template <class A, class B>
Matrix<typename traits<A>::Scalar, internal::pick_dim<traits<A>::Rows, traits<B>::Rows>::value, internal::pick_dim<traits<A>::Cols, traits<B>::Cols>::value>
operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  Matrix<typename traits<A>::Scalar, internal::pick_dim<traits<A>::Rows, traits<B>::Rows>::value, internal::pick_dim<traits<A>::Cols, traits<B>::Cols>::value> r;
  r.resize(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.ref(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return r;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, internal::pick_dim<traits<A>::Rows, traits<B>::Rows>::value, internal::pick_dim<traits<A>::Cols, traits<B>::Cols>::value>
operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  Matrix<typename traits<A>::Scalar, internal::pick_dim<traits<A>::Rows, traits<B>::Rows>::value, internal::pick_dim<traits<A>::Cols, traits<B>::Cols>::value> r;
  r.resize(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.ref(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return r;
}
template <class A>
typename MatrixBase<A>::PlainObject operator-(const MatrixBase<A>& a) {
  typename MatrixBase<A>::PlainObject r;
  r.resize(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.ref(i, j) = -a.coeff(i, j);
  return r;
}
template <class A, class S, class = typename std::enable_if<internal::is_scalar<S>::value>::type>
typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A>& a, S s) {
  typename MatrixBase<A>::PlainObject r;
  r.resize(a.rows(), a.cols());
  const typename traits<A>::Scalar ss = (typename traits<A>::Scalar)s;
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.ref(i, j) = a.coeff(i, j) * ss;
  return r;
}
template <class A, class S, class = typename std::enable_if<internal::is_scalar<S>::value>::type>
typename MatrixBase<A>::PlainObject operator*(S s, const MatrixBase<A>& a) {
  typename MatrixBase<A>::PlainObject r;
  r.resize(a.rows(), a.cols());
  const typename traits<A>::Scalar ss = (typename traits<A>::Scalar)s;
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.ref(i, j) = ss * a.coeff(i, j);
  return r;
}
template <class A, class S, class = typename std::enable_if<internal::is_scalar<S>::value>::type>
typename MatrixBase<A>::PlainObject operator/(const MatrixBase<A>& a, S s) {
  typename MatrixBase<A>::PlainObject r;
  r.resize(a.rows(), a.cols());
  const typename traits<A>::Scalar ss = (typename traits<A>::Scalar)s;
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.ref(i, j) = a.coeff(i, j) / ss;
  return r;
}
// matrix product: r(i,j) = a(i,0) b(0,j) + a(i,1) b(1,j) + ... (Eigen's coefficient-based product for small matrices, same order)
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> r;
  assert(a.cols() == b.rows());
  r.resize(a.rows(), b.cols());
  for (Index j = 0; j < b.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) {
      typename traits<A>::Scalar s = a.coeff(i, 0) * b.coeff(0, j);
      for (Index k = 1; k < a.cols(); ++k) s += a.coeff(i, k) * b.coeff(k, j);
      r.ref(i, j) = s;
    }
  return r;
}
template <class A>
std::ostream& operator<<(std::ostream& os, const MatrixBase<A>& a) {
  for (Index i = 0; i < a.rows(); ++i) {
    for (Index j = 0; j < a.cols(); ++j) os << (j ? " " : "") << a.coeff(i, j);
    if (i + 1 < a.rows()) os << "\n";
  }
  return os;
}

template <class Derived>
Matrix<typename MatrixBase<Derived>::Scalar, Dynamic, 1> MatrixBase<Derived>::diagonal() const {
  Matrix<Scalar, Dynamic, 1> d;
  const Index n = std::min(rows(), cols());
  d.resize(n, 1);
  for (Index i = 0; i < n; ++i) d.ref(i, 0) = coeff(i, i);
  return d;
}

// ---- determinant / inverse: cofactors for 2x2 and 3x3 (Eigen/src/LU/InverseImpl.h), Gauss-Jordan with partial pivoting otherwise ----
template <class Derived>
typename MatrixBase<Derived>::Scalar MatrixBase<Derived>::determinant() const {
  const Index n = rows();
  if (n == 1) return coeff(0, 0);
  if (n == 2) return coeff(0, 0) * coeff(1, 1) - coeff(1, 0) * coeff(0, 1);
  if (n == 3) {
    return coeff(0, 0) * (coeff(1, 1) * coeff(2, 2) - coeff(1, 2) * coeff(2, 1)) - coeff(1, 0) * (coeff(0, 1) * coeff(2, 2) - coeff(0, 2) * coeff(2, 1)) +
           coeff(2, 0) * (coeff(0, 1) * coeff(1, 2) - coeff(0, 2) * coeff(1, 1));
  }
  PlainObject m = eval();
  Scalar det = Scalar(1);
  for (Index k = 0; k < n; ++k) {
    Index p = k;
    for (Index i = k + 1; i < n; ++i) if (std::abs(m.get(i, k)) > std::abs(m.get(p, k))) p = i;
    if (m.get(p, k) == Scalar(0)) return Scalar(0);
    if (p != k) { for (Index j = 0; j < n; ++j) std::swap(m.ref(k, j), m.ref(p, j)); det = -det; }
    det *= m.get(k, k);
    for (Index i = k + 1; i < n; ++i) {
      const Scalar f = m.get(i, k) / m.get(k, k);
      for (Index j = k; j < n; ++j) m.ref(i, j) -= f * m.get(k, j);
    }
  }
  return det;
}
template <class Derived>
typename MatrixBase<Derived>::PlainObject MatrixBase<Derived>::inverse() const {
  const Index n = rows();
  PlainObject r;
  r.resize(n, n);
  if (n == 1) { r.ref(0, 0) = Scalar(1) / coeff(0, 0); return r; }
  if (n == 2) {  // compute_inverse<.., 2>: invdet = 1 / det; result = cofactors * invdet
    const Scalar invdet = Scalar(1) / determinant();
    r.ref(0, 0) = coeff(1, 1) * invdet;
    r.ref(1, 0) = -coeff(1, 0) * invdet;
    r.ref(0, 1) = -coeff(0, 1) * invdet;
    r.ref(1, 1) = coeff(0, 0) * invdet;
    return r;
  }
  if (n == 3) {  // compute_inverse<.., 3>: first cofactor column, det = cofactors_col0 . matrix.col(0), invdet, remaining cofactors
    auto cof = [&](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return coeff(i1, j1) * coeff(i2, j2) - coeff(i1, j2) * coeff(i2, j1);
    };
    const Scalar c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const Scalar det = c00 * coeff(0, 0) + c10 * coeff(1, 0) + c20 * coeff(2, 0);
    const Scalar invdet = Scalar(1) / det;
    r.ref(0, 0) = c00 * invdet; r.ref(0, 1) = c10 * invdet; r.ref(0, 2) = c20 * invdet;
    r.ref(1, 0) = cof(0, 1) * invdet; r.ref(1, 1) = cof(1, 1) * invdet; r.ref(1, 2) = cof(2, 1) * invdet;
    r.ref(2, 0) = cof(0, 2) * invdet; r.ref(2, 1) = cof(1, 2) * invdet; r.ref(2, 2) = cof(2, 2) * invdet;
    return r;
  }
  PlainObject m = eval();
  r.setIdentity();
  for (Index k = 0; k < n; ++k) {
    Index p = k;
    for (Index i = k + 1; i < n; ++i) if (std::abs(m.get(i, k)) > std::abs(m.get(p, k))) p = i;
    if (p != k) for (Index j = 0; j < n; ++j) { std::swap(m.ref(k, j), m.ref(p, j)); std::swap(r.ref(k, j), r.ref(p, j)); }
    const Scalar d = m.get(k, k);
    for (Index j = 0; j < n; ++j) { m.ref(k, j) /= d; r.ref(k, j) /= d; }
    for (Index i = 0; i < n; ++i) {
      if (i == k) continue;
      const Scalar f = m.get(i, k);
      if (f == Scalar(0)) continue;
      for (Index j = 0; j < n; ++j) { m.ref(i, j) -= f * m.get(k, j); r.ref(i, j) -= f * r.get(k, j); }
    }
  }
  return r;
}

// ---- LDLT: Eigen/src/Cholesky/LDLT.h, ldlt_inplace<Lower>::unblocked + LDLT::_solve_impl --------------------------------------------------
template <class M>
class LDLT_ {
  typedef typename M::Scalar Scalar;
  M m_;
  std::vector<Index> tr_;
  bool zero_ = false;

 public:
  explicit LDLT_(const M& a) : m_(a) {
    const Index n = m_.rows();
    tr_.resize((size_t)n);
    std::vector<Scalar> temp((size_t)n);
    for (Index k = 0; k < n; ++k) {
      Index big = k;
      Scalar bigv = std::abs(m_.get(k, k));
      for (Index i = k + 1; i < n; ++i) if (std::abs(m_.get(i, i)) > bigv) { bigv = std::abs(m_.get(i, i)); big = i; }
      tr_[(size_t)k] = big;
      if (big != k) {  // symmetric row / column swap on the lower triangle
        const Index s = n - big - 1;
        for (Index j = 0; j < k; ++j) std::swap(m_.ref(k, j), m_.ref(big, j));
        for (Index i = 0; i < s; ++i) std::swap(m_.ref(big + 1 + i, k), m_.ref(big + 1 + i, big));
        std::swap(m_.ref(k, k), m_.ref(big, big));
        for (Index i = k + 1; i < big; ++i) std::swap(m_.ref(i, k), m_.ref(big, i));
      }
      const Index rs = n - k - 1;
      if (k > 0) {
        for (Index j = 0; j < k; ++j) temp[(size_t)j] = m_.get(j, j) * m_.get(k, j);
        Scalar s = m_.get(k, 0) * temp[0];
        for (Index j = 1; j < k; ++j) s += m_.get(k, j) * temp[(size_t)j];
        m_.ref(k, k) -= s;
        for (Index i = 0; i < rs; ++i) {
          Scalar s2 = m_.get(k + 1 + i, 0) * temp[0];
          for (Index j = 1; j < k; ++j) s2 += m_.get(k + 1 + i, j) * temp[(size_t)j];
          m_.ref(k + 1 + i, k) -= s2;
        }
      }
      const Scalar akk = m_.get(k, k);
      const bool valid = std::abs(akk) > Scalar(0);
      if (k == 0 && !valid) {
        for (Index j = 0; j < n; ++j) tr_[(size_t)j] = j;
        zero_ = true;
        break;
      }
      if (rs > 0 && valid) for (Index i = 0; i < rs; ++i) m_.ref(k + 1 + i, k) /= akk;
    }
  }
  template <class B>
  typename MatrixBase<B>::PlainObject solve(const MatrixBase<B>& b) const {
    typename MatrixBase<B>::PlainObject y = b.eval();
    const Index n = m_.rows();
    for (Index c = 0; c < y.cols(); ++c) {
      if (zero_) { for (Index i = 0; i < n; ++i) y.ref(i, c) = Scalar(0); continue; }
      for (Index k = 0; k < n; ++k) std::swap(y.ref(k, c), y.ref(tr_[(size_t)k], c));                      // P b
      for (Index i = 0; i < n; ++i) for (Index j = 0; j < i; ++j) y.ref(i, c) -= m_.get(i, j) * y.get(j, c);  // L^-1
      const Scalar tol = (std::numeric_limits<Scalar>::min)();
      for (Index i = 0; i < n; ++i) { if (std::abs(m_.get(i, i)) > tol) y.ref(i, c) /= m_.get(i, i); else y.ref(i, c) = Scalar(0); }
      for (Index i = n - 1; i >= 0; --i) for (Index j = i + 1; j < n; ++j) y.ref(i, c) -= m_.get(j, i) * y.get(j, c);  // L^-T
      for (Index k = n - 1; k >= 0; --k) std::swap(y.ref(k, c), y.ref(tr_[(size_t)k], c));                 // P^T
    }
    return y;
  }
  Matrix<Scalar, Dynamic, 1> vectorD() const { return m_.diagonal(); }
  bool isPositive() const { for (Index i = 0; i < m_.rows(); ++i) if (m_.get(i, i) < Scalar(0)) return false; return true; }
};
template <class Derived>
LDLT_<typename MatrixBase<Derived>::PlainObject> MatrixBase<Derived>::ldlt() const { return LDLT_<PlainObject>(eval()); }

// ---- Quaternion: Eigen/src/Geometry/Quaternion.h ------------------------------------------------------------------------------------------
template <class T>
class Quaternion {
  T x_, y_, z_, w_;  // Eigen stores (x, y, z, w)

 public:
  typedef Matrix<T, 3, 1> Vector3;
  typedef Matrix<T, 3, 3> Matrix3;
  Quaternion() : x_(0), y_(0), z_(0), w_(1) {}
  Quaternion(const T& w, const T& x, const T& y, const T& z) : x_(x), y_(y), z_(z), w_(w) {}
  template <class D> explicit Quaternion(const MatrixBase<D>& m) { *this = m; }
  template <class D> Quaternion& operator=(const MatrixBase<D>& mat) {  // quaternionbase_assign_impl<.., 3, 3>: "Quaternion Calculus and Fast Animation"
    using std::sqrt;
    T t = mat.trace();
    if (t > T(0)) {
      t = sqrt(t + T(1.0));
      w_ = T(0.5) * t;
      t = T(0.5) / t;
      x_ = (mat.coeff(2, 1) - mat.coeff(1, 2)) * t;
      y_ = (mat.coeff(0, 2) - mat.coeff(2, 0)) * t;
      z_ = (mat.coeff(1, 0) - mat.coeff(0, 1)) * t;
    } else {
      int i = 0;
      if (mat.coeff(1, 1) > mat.coeff(0, 0)) i = 1;
      if (mat.coeff(2, 2) > mat.coeff(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = sqrt(mat.coeff(i, i) - mat.coeff(j, j) - mat.coeff(k, k) + T(1.0));
      T v[3];
      v[i] = T(0.5) * t;
      t = T(0.5) / t;
      w_ = (mat.coeff(k, j) - mat.coeff(j, k)) * t;
      v[j] = (mat.coeff(j, i) + mat.coeff(i, j)) * t;
      v[k] = (mat.coeff(k, i) + mat.coeff(i, k)) * t;
      x_ = v[0]; y_ = v[1]; z_ = v[2];
    }
    return *this;
  }
  T w() const { return w_; }
  T x() const { return x_; }
  T y() const { return y_; }
  T z() const { return z_; }
  T& w() { return w_; }
  T& x() { return x_; }
  T& y() { return y_; }
  T& z() { return z_; }
  Vector3 vec() const { return Vector3(x_, y_, z_); }
  Matrix<T, 4, 1> coeffs() const { return Matrix<T, 4, 1>(x_, y_, z_, w_); }
  Quaternion& setIdentity() { x_ = y_ = z_ = T(0); w_ = T(1); return *this; }
  static Quaternion Identity() { return Quaternion(); }
  T squaredNorm() const { return x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_; }
  T norm() const { using std::sqrt; return sqrt(squaredNorm()); }
  void normalize() { const T n = norm(); x_ /= n; y_ /= n; z_ /= n; w_ /= n; }
  Quaternion normalized() const { Quaternion q = *this; q.normalize(); return q; }
  Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
  Quaternion inverse() const { const T n2 = squaredNorm(); return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2); }
  Quaternion operator*(const Quaternion& b) const {  // quat_product
    const Quaternion& a = *this;
    return Quaternion(a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_, a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
                      a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_, a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_);
  }
  Quaternion& operator*=(const Quaternion& b) { *this = (*this) * b; return *this; }
  template <class D> Vector3 _transformVector(const MatrixBase<D>& v) const {  // v + w * (2 q x v) + q x (2 q x v)
    Vector3 uv = vec().cross(v);
    uv += uv;
    return Vector3(v) + w_ * uv + vec().cross(uv);
  }
  template <class D> Vector3 operator*(const MatrixBase<D>& v) const { return _transformVector(v); }
  Matrix3 toRotationMatrix() const {
    Matrix3 res;
    const T tx = T(2) * x_, ty = T(2) * y_, tz = T(2) * z_;
    const T twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const T txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const T tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    res.ref(0, 0) = T(1) - (tyy + tzz); res.ref(0, 1) = txy - twz; res.ref(0, 2) = txz + twy;
    res.ref(1, 0) = txy + twz; res.ref(1, 1) = T(1) - (txx + tzz); res.ref(1, 2) = tyz - twx;
    res.ref(2, 0) = txz - twy; res.ref(2, 1) = tyz + twx; res.ref(2, 2) = T(1) - (txx + tyy);
    return res;
  }
  Matrix3 matrix() const { return toRotationMatrix(); }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <class T>
class AngleAxis {
  Matrix<T, 3, 1> axis_;
  T angle_;

 public:
  AngleAxis() : angle_(0) {}
  template <class D> AngleAxis(const T& angle, const MatrixBase<D>& axis) : axis_(axis), angle_(angle) {}
  T angle() const { return angle_; }
  const Matrix<T, 3, 1>& axis() const { return axis_; }
  Matrix<T, 3, 3> toRotationMatrix() const {  // AngleAxis::toRotationMatrix
    Matrix<T, 3, 3> res;
    const T s = std::sin(angle_), c = std::cos(angle_);
    Matrix<T, 3, 1> sin_axis = s * axis_;
    Matrix<T, 3, 1> cos1_axis = (T(1) - c) * axis_;
    T tmp = cos1_axis.x() * axis_.y();
    res.ref(0, 1) = tmp - sin_axis.z(); res.ref(1, 0) = tmp + sin_axis.z();
    tmp = cos1_axis.x() * axis_.z();
    res.ref(0, 2) = tmp + sin_axis.y(); res.ref(2, 0) = tmp - sin_axis.y();
    tmp = cos1_axis.y() * axis_.z();
    res.ref(1, 2) = tmp - sin_axis.x(); res.ref(2, 1) = tmp + sin_axis.x();
    res.ref(0, 0) = cos1_axis.x() * axis_.x() + c; res.ref(1, 1) = cos1_axis.y() * axis_.y() + c; res.ref(2, 2) = cos1_axis.z() * axis_.z() + c;
    return res;
  }
  Matrix<T, 3, 3> matrix() const { return toRotationMatrix(); }
};
typedef AngleAxis<double> AngleAxisd;

template <class T, int N, int MaxN = N>
class DiagonalMatrix : public Matrix<T, N, N> {
 public:
  DiagonalMatrix() { this->setZero(); }
  DiagonalMatrix(const T& a, const T& b) { this->setZero(); this->ref(0, 0) = a; this->ref(1, 1) = b; }
  DiagonalMatrix(const T& a, const T& b, const T& c) { this->setZero(); this->ref(0, 0) = a; this->ref(1, 1) = b; this->ref(2, 2) = c; }
};
template <class T, int N, int MaxN>
struct traits<DiagonalMatrix<T, N, MaxN>> : traits<Matrix<T, N, N>> {};

#define MINI_EIGEN_TYPEDEFS(T, S)                                  \
  typedef Matrix<T, 2, 2> Matrix2##S;                              \
  typedef Matrix<T, 3, 3> Matrix3##S;                              \
  typedef Matrix<T, 4, 4> Matrix4##S;                              \
  typedef Matrix<T, Dynamic, Dynamic> MatrixX##S;                  \
  typedef Matrix<T, 2, 1> Vector2##S;                              \
  typedef Matrix<T, 3, 1> Vector3##S;                              \
  typedef Matrix<T, 4, 1> Vector4##S;                              \
  typedef Matrix<T, Dynamic, 1> VectorX##S;                        \
  typedef Matrix<T, 1, 2> RowVector2##S;                           \
  typedef Matrix<T, 1, 3> RowVector3##S;                           \
  typedef Matrix<T, 1, Dynamic> RowVectorX##S;
MINI_EIGEN_TYPEDEFS(double, d)
MINI_EIGEN_TYPEDEFS(float, f)
MINI_EIGEN_TYPEDEFS(int, i)
#undef MINI_EIGEN_TYPEDEFS

}  // namespace Eigen
