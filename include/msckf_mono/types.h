// include/msckf_mono/types.h -- boundary data types of the drop-in msckf_mono::MSCKF<_S> (this repo).
// Mirrors the reference's include/msckf_mono/types.h:8-126 name for name.  When Eigen is available
// (the reference's callers always have it) the aliases ARE the Eigen types, so existing callers
// (src/ros_interface.cpp, datasets/asl_msckf.cpp) compile unchanged.  On a box without Eigen (this
// image) small POD look-alikes with the accessor subset the filter uses stand in, so the class and its
// tests build with nothing but a C++17 compiler.
#ifndef MSCKF_MONO_SENSOR_TYPES_H_
#define MSCKF_MONO_SENSOR_TYPES_H_

#include <cstddef>
#include <vector>

#if !defined(MSCKF_B200_NO_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Dense>)
#define MSCKF_B200_HAVE_EIGEN 1
#endif
#endif

#ifdef MSCKF_B200_HAVE_EIGEN
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <Eigen/StdVector>
namespace msckf_mono {
template <typename _Scalar> using Quaternion = Eigen::Quaternion<_Scalar>;
template <typename _Scalar> using Matrix3 = Eigen::Matrix<_Scalar, 3, 3>;
template <typename _Scalar> using Vector2 = Eigen::Matrix<_Scalar, 2, 1>;
template <typename _Scalar> using Vector3 = Eigen::Matrix<_Scalar, 3, 1>;
template <typename _Scalar> using Point = Vector3<_Scalar>;
template <typename _Scalar> using GyroscopeReading = Vector3<_Scalar>;
template <typename _Scalar> using AccelerometerReading = Vector3<_Scalar>;
template <typename _Scalar, int R, int C> using FixedMatrix = Eigen::Matrix<_Scalar, R, C>;
template <typename T> using aligned_vector = std::vector<T, Eigen::aligned_allocator<T>>;
}  // namespace msckf_mono
#define MSCKF_B200_ALIGNED_NEW EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#else
namespace msckf_mono {
namespace pod {
template <typename S, int N>
struct Vec {
  S d[N];
  Vec() { for (int i = 0; i < N; ++i) d[i] = S(0); }
  S& operator()(int i) { return d[i]; }
  const S& operator()(int i) const { return d[i]; }
  S& operator[](int i) { return d[i]; }
  const S& operator[](int i) const { return d[i]; }
  S& x() { return d[0]; }
  S& y() { return d[1]; }
  S& z() { return d[2]; }
  const S& x() const { return d[0]; }
  const S& y() const { return d[1]; }
  const S& z() const { return d[2]; }
};
template <typename S, int R, int C>
struct Mat {
  S d[R * C];
  Mat() { for (int i = 0; i < R * C; ++i) d[i] = S(0); }
  S& operator()(int i, int j) { return d[i * C + j]; }
  const S& operator()(int i, int j) const { return d[i * C + j]; }
  void setZero() { for (int i = 0; i < R * C; ++i) d[i] = S(0); }
};
template <typename S>
struct Quat {  // constructor order (w,x,y,z) like Eigen::Quaternion
  S x_, y_, z_, w_;
  Quat() : x_(0), y_(0), z_(0), w_(1) {}
  Quat(S w, S x, S y, S z) : x_(x), y_(y), z_(z), w_(w) {}
  S& x() { return x_; }
  S& y() { return y_; }
  S& z() { return z_; }
  S& w() { return w_; }
  const S& x() const { return x_; }
  const S& y() const { return y_; }
  const S& z() const { return z_; }
  const S& w() const { return w_; }
};
}  // namespace pod
template <typename _Scalar> using Quaternion = pod::Quat<_Scalar>;
template <typename _Scalar> using Matrix3 = pod::Mat<_Scalar, 3, 3>;
template <typename _Scalar> using Vector2 = pod::Vec<_Scalar, 2>;
template <typename _Scalar> using Vector3 = pod::Vec<_Scalar, 3>;
template <typename _Scalar> using Point = Vector3<_Scalar>;
template <typename _Scalar> using GyroscopeReading = Vector3<_Scalar>;
template <typename _Scalar> using AccelerometerReading = Vector3<_Scalar>;
template <typename _Scalar, int R, int C> using FixedMatrix = pod::Mat<_Scalar, R, C>;
template <typename T> using aligned_vector = std::vector<T>;
}  // namespace msckf_mono
#define MSCKF_B200_ALIGNED_NEW
#endif

namespace msckf_mono {
// types.h:48-56
template <typename _Scalar>
struct Camera {
  MSCKF_B200_ALIGNED_NEW
  _Scalar c_u, c_v, f_u, f_v, b;
  Quaternion<_Scalar> q_CI;
  Point<_Scalar> p_C_I;
};
// types.h:58-68
template <typename _Scalar>
struct camState {
  MSCKF_B200_ALIGNED_NEW
  Point<_Scalar> p_C_G;
  Quaternion<_Scalar> q_CG;
  _Scalar time;
  int state_id;
  int last_correlated_id;
  std::vector<size_t> tracked_feature_ids;
};
// types.h:70-77
template <typename _Scalar>
struct imuState {
  MSCKF_B200_ALIGNED_NEW
  Point<_Scalar> p_I_G, p_I_G_null;
  Vector3<_Scalar> v_I_G, b_g, b_a, g, v_I_G_null;
  Quaternion<_Scalar> q_IG, q_IG_null;
};
// types.h:79-85
template <typename _Scalar>
struct imuReading {
  MSCKF_B200_ALIGNED_NEW
  GyroscopeReading<_Scalar> omega;
  AccelerometerReading<_Scalar> a;
  _Scalar dT;
};
// types.h:87-93
template <typename _Scalar>
struct noiseParams {
  MSCKF_B200_ALIGNED_NEW
  _Scalar u_var_prime, v_var_prime;
  FixedMatrix<_Scalar, 12, 12> Q_imu;
  FixedMatrix<_Scalar, 15, 15> initial_imu_covar;
};
// types.h:95-100
template <typename _Scalar>
struct MSCKFParams {
  _Scalar max_gn_cost_norm, min_rcond, translation_threshold;
  _Scalar redundancy_angle_thresh, redundancy_distance_thresh;
  int min_track_length, max_track_length, max_cam_states;
};
// types.h:102-113.  The reference stores copies of the observing clones; here the clones are device
// resident and a residualised track carries their POSITIONS in the window (cam_state_indices, msckf.h:1481).
template <typename _Scalar>
struct featureTrackToResidualize {
  size_t feature_id;
  aligned_vector<Vector2<_Scalar>> observations;
  std::vector<size_t> cam_state_indices;
  bool initialized;
  Vector3<_Scalar> p_f_G;
  featureTrackToResidualize() : feature_id(0), initialized(false) {}
};
// types.h:115-126
template <typename _Scalar>
struct featureTrack {
  size_t feature_id;
  aligned_vector<Vector2<_Scalar>> observations;
  std::vector<size_t> cam_state_indices;  // state_ids of cam states corresponding to observations
  bool initialized = false;
  Vector3<_Scalar> p_f_G;
};
}  // namespace msckf_mono
#endif
