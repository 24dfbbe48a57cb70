// include/msckf_mono/pod_linalg.h -- Eigen-free stand-ins for the handful of Eigen types that appear in the boundary of
// msckf_mono::MSCKF<_S> (reference include/msckf_mono/types.h:8-46).  Used only when <Eigen/Dense> is not on the include
// path (this image has no Eigen): fixed- and dynamic-size row-major matrices, vectors and a quaternion with the accessor
// subset the reference's callers use on these types (src/ros_interface.cpp:250-262, datasets/asl_msckf.cpp:57-160):
// (i), (i,j), [i], x()/y()/z()/w(), setZero/setIdentity, Zero/Identity, comma initialiser, block<R,C>(i,j), transpose(),
// asDiagonal(), + - * (matrix, scalar), norm(), dot(), Quaternion (w,x,y,z) / from a rotation matrix, inverse(),
// toRotationMatrix(), q * v, q * q, normalize(), operator<< to a stream.  Host-side convenience only: no filter numerics
// go through these types (all floating-point work of the filter runs on the device behind include/msckf_b200.h).
#ifndef MSCKF_MONO_POD_LINALG_H_
#define MSCKF_MONO_POD_LINALG_H_

#include <cmath>
#include <cstddef>
#include <ostream>
#include <stdexcept>
#include <vector>

namespace msckf_mono {
namespace pod {

constexpr int Dynamic = -1;

template <typename S, int R, int C> struct Mat;

namespace detail {
template <typename S, int R, int C, bool Dyn = (R == Dynamic || C == Dynamic)>
struct Storage {  // fixed size
  S d[R * C];
  Storage() { for (int i = 0; i < R * C; ++i) d[i] = S(0); }
  int rows() const { return R; }
  int cols() const { return C; }
  void resize(int r, int c) { if (r != R || c != C) throw std::length_error("pod::Mat: fixed size"); }
  S* data() { return d; }
  const S* data() const { return d; }
};
template <typename S, int R, int C>
struct Storage<S, R, C, true> {  // at least one dynamic dimension
  std::vector<S> d;
  int r_ = (R == Dynamic ? 0 : R), c_ = (C == Dynamic ? 0 : C);
  int rows() const { return r_; }
  int cols() const { return c_; }
  void resize(int r, int c) { r_ = r; c_ = c; d.assign((size_t)r * c, S(0)); }
  S* data() { return d.data(); }
  const S* data() const { return d.data(); }
};
}  // namespace detail

template <typename M>
struct CommaInit {  // Eigen's "m << a, b, c;"
  M& m; int k;
  CommaInit(M& mm, typename M::Scalar v) : m(mm), k(0) { put(v); }
  CommaInit& operator,(typename M::Scalar v) { put(v); return *this; }
  void put(typename M::Scalar v) { if (k >= m.rows() * m.cols()) throw std::out_of_range("pod::Mat: too many coefficients"); m.data()[k++] = v; }
};
template <typename V> struct Diag { const V& v; };

template <typename S, int R, int C, int BR, int BC>
struct BlockRef {  // writable view of a BR x BC block
  Mat<S, R, C>& m; int i0, j0;
  template <class Other> BlockRef& operator=(const Other& o) {
    for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) m(i0 + i, j0 + j) = o(i, j);
    return *this;
  }
  BlockRef& operator=(const BlockRef& o) {
    for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) m(i0 + i, j0 + j) = o(i, j);
    return *this;
  }
  S operator()(int i, int j) const { return m(i0 + i, j0 + j); }
  operator Mat<S, BR, BC>() const { Mat<S, BR, BC> o; for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) o(i, j) = m(i0 + i, j0 + j); return o; }
};

template <typename S, int R, int C>
struct Mat : detail::Storage<S, R, C> {
  typedef S Scalar;
  typedef detail::Storage<S, R, C> Base;
  using Base::rows; using Base::cols; using Base::data;
  Mat() {}
  Mat(int r, int c) { Base::resize(r, c); }
  explicit Mat(int n) { Base::resize(R == Dynamic ? n : R, C == Dynamic ? n : C); }
  template <int BR, int BC, int OR, int OC> Mat(const BlockRef<S, OR, OC, BR, BC>& b) { *this = (Mat<S, BR, BC>)b; }
  int size() const { return rows() * cols(); }
  S& operator()(int i, int j) { return data()[(size_t)i * cols() + j]; }
  const S& operator()(int i, int j) const { return data()[(size_t)i * cols() + j]; }
  S& operator()(int i) { return data()[i]; }
  const S& operator()(int i) const { return data()[i]; }
  S& operator[](int i) { return data()[i]; }
  const S& operator[](int i) const { return data()[i]; }
  S& x() { return data()[0]; } const S& x() const { return data()[0]; }
  S& y() { return data()[1]; } const S& y() const { return data()[1]; }
  S& z() { return data()[2]; } const S& z() const { return data()[2]; }
  S& w() { return data()[3]; } const S& w() const { return data()[3]; }
  Mat& setZero() { for (int i = 0; i < size(); ++i) data()[i] = S(0); return *this; }
  Mat& setIdentity() { setZero(); for (int i = 0; i < rows() && i < cols(); ++i) (*this)(i, i) = S(1); return *this; }
  static Mat Zero() { return Mat(); }
  static Mat Zero(int r, int c) { return Mat(r, c); }
  static Mat Identity() { Mat m; m.setIdentity(); return m; }
  static Mat Identity(int r, int c) { Mat m(r, c); m.setIdentity(); return m; }
  CommaInit<Mat> operator<<(S v) { return CommaInit<Mat>(*this, v); }
  Mat<S, C, R> transpose() const {
    Mat<S, C, R> t; t.resize(cols(), rows());
    for (int i = 0; i < rows(); ++i) for (int j = 0; j < cols(); ++j) t(j, i) = (*this)(i, j);
    return t;
  }
  template <int BR, int BC> BlockRef<S, R, C, BR, BC> block(int i, int j) { return BlockRef<S, R, C, BR, BC>{*this, i, j}; }
  template <int BR, int BC> Mat<S, BR, BC> block(int i0, int j0) const {
    Mat<S, BR, BC> o; for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) o(i, j) = (*this)(i0 + i, j0 + j); return o;
  }
  Diag<Mat> asDiagonal() const { return Diag<Mat>{*this}; }
  template <class V> Mat& operator=(const Diag<V>& dg) {
    const int n = dg.v.size(); Base::resize(R == Dynamic ? n : R, C == Dynamic ? n : C); setZero();
    for (int i = 0; i < n; ++i) (*this)(i, i) = dg.v(i);
    return *this;
  }
  S squaredNorm() const { S s = 0; for (int i = 0; i < size(); ++i) s += data()[i] * data()[i]; return s; }
  S norm() const { return std::sqrt(squaredNorm()); }
  S dot(const Mat& o) const { S s = 0; for (int i = 0; i < size(); ++i) s += data()[i] * o.data()[i]; return s; }
  Mat operator-() const { Mat o(*this); for (int i = 0; i < size(); ++i) o.data()[i] = -o.data()[i]; return o; }
  Mat& operator+=(const Mat& o) { for (int i = 0; i < size(); ++i) data()[i] += o.data()[i]; return *this; }
  Mat& operator-=(const Mat& o) { for (int i = 0; i < size(); ++i) data()[i] -= o.data()[i]; return *this; }
  Mat& operator*=(S s) { for (int i = 0; i < size(); ++i) data()[i] *= s; return *this; }
  friend Mat operator+(Mat a, const Mat& b) { a += b; return a; }
  friend Mat operator-(Mat a, const Mat& b) { a -= b; return a; }
  friend Mat operator*(Mat a, S s) { a *= s; return a; }
  friend Mat operator*(S s, Mat a) { a *= s; return a; }
  friend Mat operator/(Mat a, S s) { for (int i = 0; i < a.size(); ++i) a.data()[i] /= s; return a; }
  friend std::ostream& operator<<(std::ostream& os, const Mat& m) {
    for (int i = 0; i < m.rows(); ++i) { for (int j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j); if (i + 1 < m.rows()) os << "\n"; }
    return os;
  }
};
template <typename S, int R, int K, int C>
Mat<S, R, C> operator*(const Mat<S, R, K>& a, const Mat<S, K, C>& b) {
  Mat<S, R, C> o; o.resize(a.rows(), b.cols());
  for (int i = 0; i < a.rows(); ++i) for (int j = 0; j < b.cols(); ++j) { S s = 0; for (int k = 0; k < a.cols(); ++k) s += a(i, k) * b(k, j); o(i, j) = s; }
  return o;
}
template <typename S, int N> using Vec = Mat<S, N, 1>;

template <typename S>
struct Quat {  // constructor order (w,x,y,z) like Eigen::Quaternion; coefficients stored (x,y,z,w) like Eigen's coeffs()
  S x_, y_, z_, w_;
  Quat() : x_(0), y_(0), z_(0), w_(1) {}
  Quat(S w, S x, S y, S z) : x_(x), y_(y), z_(z), w_(w) {}
  explicit Quat(const Mat<S, 3, 3>& m) {  // Eigen's rotation-matrix -> quaternion branch structure
    const S t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > S(0)) {
      S s = std::sqrt(t + S(1)); w_ = S(0.5) * s; s = S(0.5) / s;
      x_ = (m(2, 1) - m(1, 2)) * s; y_ = (m(0, 2) - m(2, 0)) * s; z_ = (m(1, 0) - m(0, 1)) * s;
    } else {
      int i = 0; if (m(1, 1) > m(0, 0)) i = 1; if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      S s = std::sqrt(m(i, i) - m(j, j) - m(k, k) + S(1));
      S v[3]; v[i] = S(0.5) * s; s = S(0.5) / s;
      w_ = (m(k, j) - m(j, k)) * s; v[j] = (m(j, i) + m(i, j)) * s; v[k] = (m(k, i) + m(i, k)) * s;
      x_ = v[0]; y_ = v[1]; z_ = v[2];
    }
  }
  S& x() { return x_; } S& y() { return y_; } S& z() { return z_; } S& w() { return w_; }
  const S& x() const { return x_; } const S& y() const { return y_; } const S& z() const { return z_; } const S& w() const { return w_; }
  static Quat Identity() { return Quat(); }
  S squaredNorm() const { return x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_; }
  S norm() const { return std::sqrt(squaredNorm()); }
  void normalize() { const S n = norm(); x_ /= n; y_ /= n; z_ /= n; w_ /= n; }
  Quat normalized() const { Quat q(*this); q.normalize(); return q; }
  Quat conjugate() const { return Quat(w_, -x_, -y_, -z_); }
  Quat inverse() const { const S n2 = squaredNorm(); return Quat(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2); }
  Mat<S, 3, 3> toRotationMatrix() const {
    Mat<S, 3, 3> R;
    const S tx = 2 * x_, ty = 2 * y_, tz = 2 * z_, twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_,
            tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1 - (txx + tyy);
    return R;
  }
  friend Quat operator*(const Quat& a, const Quat& b) {
    return Quat(a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_, a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
                a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_, a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_);
  }
  friend Mat<S, 3, 1> operator*(const Quat& q, const Mat<S, 3, 1>& v) { return q.toRotationMatrix() * v; }
};

// Eigen::Transform<S,3,Isometry> look-alike: linear() / translation() / matrix() / inverse() / composition
template <typename S>
struct Iso3 {
  Mat<S, 3, 3> R; Mat<S, 3, 1> t;
  Iso3() { R.setIdentity(); }
  static Iso3 Identity() { return Iso3(); }
  Mat<S, 3, 3>& linear() { return R; } const Mat<S, 3, 3>& linear() const { return R; }
  Mat<S, 3, 1>& translation() { return t; } const Mat<S, 3, 1>& translation() const { return t; }
  Mat<S, 4, 4> matrix() const { Mat<S, 4, 4> m; m.setIdentity(); for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) m(i, j) = R(i, j); m(i, 3) = t(i); } return m; }
  Iso3 inverse() const { Iso3 o; o.R = R.transpose(); o.t = -(o.R * t); return o; }
  friend Iso3 operator*(const Iso3& a, const Iso3& b) { Iso3 o; o.R = a.R * b.R; o.t = a.R * b.t + a.t; return o; }
  friend Mat<S, 3, 1> operator*(const Iso3& a, const Mat<S, 3, 1>& p) { return a.R * p + a.t; }
};

}  // namespace pod
}  // namespace msckf_mono
#endif
