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
// reference types.h:8-46, name for name
template <typename _Scalar> using Quaternion = Eigen::Quaternion<_Scalar>;
template <typename _Scalar> using Matrix3 = Eigen::Matrix<_Scalar, 3, 3>;
template <typename _Scalar> using Matrix4 = Eigen::Matrix<_Scalar, 4, 4>;
template <typename _Scalar> using MatrixX = Eigen::Matrix<_Scalar, Eigen::Dynamic, Eigen::Dynamic>;
template <typename _Scalar> using RowVector3 = Eigen::Matrix<_Scalar, 1, 3>;
template <typename _Scalar> using Vector2 = Eigen::Matrix<_Scalar, 2, 1>;
template <typename _Scalar> using Vector3 = Eigen::Matrix<_Scalar, 3, 1>;
template <typename _Scalar> using Vector4 = Eigen::Matrix<_Scalar, 4, 1>;
template <typename _Scalar> using VectorX = Eigen::Matrix<_Scalar, Eigen::Dynamic, 1>;
template <typename _Scalar> using Point = Vector3<_Scalar>;
template <typename _Scalar> using GyroscopeReading = Vector3<_Scalar>;
template <typename _Scalar> using AccelerometerReading = Vector3<_Scalar>;
template <typename _Scalar> using Isometry3 = Eigen::Transform<_Scalar, 3, Eigen::Isometry>;
// not in the reference: helpers of this header
template <typename _Scalar, int R, int C> using FixedMatrix = Eigen::Matrix<_Scalar, R, C>;
template <typename T> using aligned_vector = std::vector<T, Eigen::aligned_allocator<T>>;
using IndexVector = Eigen::VectorXi;  // matrix_utils.h:61,79
constexpr int DynamicSize = Eigen::Dynamic;
}  // namespace msckf_mono
#define MSCKF_B200_ALIGNED_NEW EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#else
#include <msckf_mono/pod_linalg.h>
namespace msckf_mono {
// reference types.h:8-46, name for name, over the Eigen-free stand-ins of pod_linalg.h
template <typename _Scalar> using Quaternion = pod::Quat<_Scalar>;
template <typename _Scalar> using Matrix3 = pod::Mat<_Scalar, 3, 3>;
template <typename _Scalar> using Matrix4 = pod::Mat<_Scalar, 4, 4>;
template <typename _Scalar> using MatrixX = pod::Mat<_Scalar, pod::Dynamic, pod::Dynamic>;
template <typename _Scalar> using RowVector3 = pod::Mat<_Scalar, 1, 3>;
template <typename _Scalar> using Vector2 = pod::Mat<_Scalar, 2, 1>;
template <typename _Scalar> using Vector3 = pod::Mat<_Scalar, 3, 1>;
template <typename _Scalar> using Vector4 = pod::Mat<_Scalar, 4, 1>;
template <typename _Scalar> using VectorX = pod::Mat<_Scalar, pod::Dynamic, 1>;
template <typename _Scalar> using Point = Vector3<_Scalar>;
template <typename _Scalar> using GyroscopeReading = Vector3<_Scalar>;
template <typename _Scalar> using AccelerometerReading = Vector3<_Scalar>;
template <typename _Scalar> using Isometry3 = pod::Iso3<_Scalar>;
template <typename _Scalar, int R, int C> using FixedMatrix = pod::Mat<_Scalar, R, C>;
template <typename T> using aligned_vector = std::vector<T>;
using IndexVector = pod::Mat<int, pod::Dynamic, 1>;
constexpr int DynamicSize = pod::Dynamic;
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
// types.h:102-113.  The reference stores copies of the observing clones in `cam_states`; here the clones are
// device resident and a residualised track carries their POSITIONS in the window (cam_state_indices,
// msckf.h:1481).  The member is kept for source compatibility; the drop-in class leaves it empty.
template <typename _Scalar>
struct featureTrackToResidualize {
  size_t feature_id;
  aligned_vector<Vector2<_Scalar>> observations;
  std::vector<camState<_Scalar>> cam_states;
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
