// msckf_mono_b200/asl/asl_io.hpp
// ASL / EuRoC "mav0" folder I/O for the ROS-free player.  Mirrors what the reference's
// datasets/asl_readers.cpp reads (imu0/data.csv, cam0/sensor.yaml, state_groundtruth_estimate0/data.csv,
// /root/reference/datasets/asl_readers.cpp:27-33 for the T_BS convention) WITHOUT its OpenCV / image front end: the
// feature tracks come pre-extracted from cam0/tracks.csv (normalised, undistorted coordinates -- what the reference's
// TrackHandler hands to MSCKF::update / addFeatures, src/corner_detector.cpp).  Plain C++17, no dependencies.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace msckf_b200 {
namespace asl {

struct ImuRow { int64_t t_ns; double w[3], a[3]; };
struct GtRow { int64_t t_ns; double p[3], q_wxyz[4], v[3], bw[3], ba[3]; };
struct TrackRow { int64_t t_ns; uint64_t id; double x, y; };
struct CameraInfo {
  double T_BS[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // camera pose in the body (IMU) frame
  double fu = 1, fv = 1, cu = 0, cv = 0;
  int width = 0, height = 0;
};
// filter configuration (the reference hard-codes / reads these in datasets/asl_msckf.cpp:73-125)
struct FilterConfig {
  double feature_cov = 7.0;  // pixels
  double w_var = 1e-4, dbg_var = 3.6733e-5, a_var = 1e-2, dba_var = 7e-2;
  double q_var_init = 1e-5, bg_var_init = 1e-2, v_var_init = 1e-2, ba_var_init = 1e-2, p_var_init = 1e-12;
  double max_gn_cost_norm = 11.0, min_rcond = 3e-12, translation_threshold = 0.01;  // max_gn_cost_norm in pixels
  double redundancy_angle_thresh = 0.05, redundancy_distance_thresh = 0.05;
  int min_track_length = 3, max_track_length = 50, max_cam_states = 30;
  double gravity[3] = {0, 0, -9.81};
};

inline std::vector<std::string> split_csv(const std::string& line) {
  std::vector<std::string> out;
  std::string cur;
  for (char ch : line) {
    if (ch == ',') { out.push_back(cur); cur.clear(); }
    else if (ch != '\r' && ch != '\n') cur.push_back(ch);
  }
  out.push_back(cur);
  return out;
}

template <class Row, class Fn>
inline std::vector<Row> read_csv(const std::string& path, size_t min_cols, Fn&& make) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::vector<Row> rows;
  std::string line;
  size_t ln = 0;
  while (std::getline(f, line)) {
    ++ln;
    if (line.empty() || line[0] == '#') continue;
    const auto c = split_csv(line);
    if (c.size() < min_cols) throw std::runtime_error(path + ":" + std::to_string(ln) + ": expected " + std::to_string(min_cols) + " columns");
    rows.push_back(make(c));
  }
  return rows;
}

inline std::vector<ImuRow> read_imu(const std::string& mav0) {
  return read_csv<ImuRow>(mav0 + "/imu0/data.csv", 7, [](const std::vector<std::string>& c) {
    ImuRow r;
    r.t_ns = std::stoll(c[0]);
    for (int i = 0; i < 3; ++i) { r.w[i] = std::stod(c[1 + i]); r.a[i] = std::stod(c[4 + i]); }
    return r;
  });
}
inline std::vector<GtRow> read_groundtruth(const std::string& mav0) {
  return read_csv<GtRow>(mav0 + "/state_groundtruth_estimate0/data.csv", 17, [](const std::vector<std::string>& c) {
    GtRow r;
    r.t_ns = std::stoll(c[0]);
    for (int i = 0; i < 3; ++i) { r.p[i] = std::stod(c[1 + i]); r.v[i] = std::stod(c[8 + i]); r.bw[i] = std::stod(c[11 + i]); r.ba[i] = std::stod(c[14 + i]); }
    for (int i = 0; i < 4; ++i) r.q_wxyz[i] = std::stod(c[4 + i]);
    return r;
  });
}
inline std::vector<TrackRow> read_tracks(const std::string& mav0) {
  return read_csv<TrackRow>(mav0 + "/cam0/tracks.csv", 4, [](const std::vector<std::string>& c) {
    TrackRow r;
    r.t_ns = std::stoll(c[0]);
    r.id = std::stoull(c[1]);
    r.x = std::stod(c[2]);
    r.y = std::stod(c[3]);
    return r;
  });
}
inline std::vector<int64_t> read_frame_times(const std::string& mav0) {
  return read_csv<int64_t>(mav0 + "/cam0/data.csv", 1, [](const std::vector<std::string>& c) { return (int64_t)std::stoll(c[0]); });
}

// the few "key: value" / "key: [a, b, ...]" forms EuRoC's sensor.yaml uses (data: may span several lines)
inline std::map<std::string, std::vector<double>> read_yaml_numbers(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string txt = ss.str();
  std::map<std::string, std::vector<double>> out;
  size_t pos = 0;
  while (pos < txt.size()) {
    size_t eol = txt.find('\n', pos);
    if (eol == std::string::npos) eol = txt.size();
    std::string line = txt.substr(pos, eol - pos);
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    const size_t colon = line.find(':');
    size_t next = eol + 1;
    if (colon != std::string::npos) {
      std::string key = line.substr(0, colon);
      key.erase(0, key.find_first_not_of(" \t"));
      key.erase(key.find_last_not_of(" \t") + 1);
      std::string val = line.substr(colon + 1);
      const size_t lb = val.find('[');
      if (lb != std::string::npos) {
        size_t abs_lb = pos + colon + 1 + lb;
        size_t rb = txt.find(']', abs_lb);
        if (rb == std::string::npos) throw std::runtime_error(path + ": unterminated list for " + key);
        val = txt.substr(abs_lb + 1, rb - abs_lb - 1);
        next = txt.find('\n', rb);
        next = (next == std::string::npos) ? txt.size() : next + 1;
      }
      for (char& ch : val) if (ch == ',' || ch == '\n' || ch == '\r') ch = ' ';
      std::stringstream vs(val);
      std::vector<double> nums;
      std::string tok;
      bool ok = true;
      while (vs >> tok) {
        try { size_t used = 0; const double d = std::stod(tok, &used); if (used != tok.size()) { ok = false; break; } nums.push_back(d); }
        catch (...) { ok = false; break; }
      }
      if (ok && !nums.empty() && !key.empty()) out[key] = nums;
    }
    pos = next;
  }
  return out;
}

inline CameraInfo read_camera(const std::string& mav0) {
  const auto y = read_yaml_numbers(mav0 + "/cam0/sensor.yaml");
  CameraInfo c;
  auto it = y.find("data");
  if (it == y.end() || it->second.size() != 16) throw std::runtime_error("cam0/sensor.yaml: T_BS data[16] missing");
  for (int i = 0; i < 16; ++i) c.T_BS[i] = it->second[i];
  it = y.find("intrinsics");
  if (it == y.end() || it->second.size() != 4) throw std::runtime_error("cam0/sensor.yaml: intrinsics[4] missing");
  c.fu = it->second[0]; c.fv = it->second[1]; c.cu = it->second[2]; c.cv = it->second[3];
  it = y.find("resolution");
  if (it != y.end() && it->second.size() == 2) { c.width = (int)it->second[0]; c.height = (int)it->second[1]; }
  return c;
}

inline FilterConfig read_filter_config(const std::string& mav0) {
  FilterConfig fc;
  std::ifstream probe(mav0 + "/msckf.yaml");
  if (!probe) return fc;  // the reference's defaults
  const auto y = read_yaml_numbers(mav0 + "/msckf.yaml");
  auto get = [&](const char* k, double& v) { auto it = y.find(k); if (it != y.end() && !it->second.empty()) v = it->second[0]; };
  auto geti = [&](const char* k, int& v) { auto it = y.find(k); if (it != y.end() && !it->second.empty()) v = (int)it->second[0]; };
  get("feature_cov", fc.feature_cov);
  get("w_var", fc.w_var); get("dbg_var", fc.dbg_var); get("a_var", fc.a_var); get("dba_var", fc.dba_var);
  get("q_var_init", fc.q_var_init); get("bg_var_init", fc.bg_var_init); get("v_var_init", fc.v_var_init);
  get("ba_var_init", fc.ba_var_init); get("p_var_init", fc.p_var_init);
  get("max_gn_cost_norm", fc.max_gn_cost_norm); get("min_rcond", fc.min_rcond); get("translation_threshold", fc.translation_threshold);
  get("redundancy_angle_thresh", fc.redundancy_angle_thresh); get("redundancy_distance_thresh", fc.redundancy_distance_thresh);
  geti("min_track_length", fc.min_track_length); geti("max_track_length", fc.max_track_length); geti("max_cam_states", fc.max_cam_states);
  auto it = y.find("gravity");
  if (it != y.end() && it->second.size() == 3) for (int i = 0; i < 3; ++i) fc.gravity[i] = it->second[i];
  return fc;
}

// rotation matrix (row-major 3x3) -> quaternion (w, x, y, z)
inline void rot_to_quat(const double R[9], double q[4]) {
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    const double s = std::sqrt(tr + 1.0) * 2;
    q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
    q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    const double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
    q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
  } else {
    const double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
    q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
  }
}

struct PoseRow { int64_t t_ns; double p[3], q_wxyz[4]; int n_clones; int n_tracks_residualized; };

inline void write_trajectory(const std::string& path, const std::vector<PoseRow>& rows) {
  FILE* f = std::fopen(path.c_str(), "w");
  if (!f) throw std::runtime_error("cannot write " + path);
  std::fprintf(f, "#timestamp [ns],p_x,p_y,p_z,q_w,q_x,q_y,q_z,n_clones\n");
  for (const auto& r : rows)
    std::fprintf(f, "%lld,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%d\n", (long long)r.t_ns, r.p[0], r.p[1], r.p[2], r.q_wxyz[0],
                 r.q_wxyz[1], r.q_wxyz[2], r.q_wxyz[3], r.n_clones);
  std::fclose(f);
}

// position RMSE against the ground truth rows with the same timestamps (no alignment: the filter starts from the truth)
inline double position_rmse(const std::vector<PoseRow>& est, const std::vector<GtRow>& gt, int* n_matched) {
  std::map<int64_t, const GtRow*> by_t;
  for (const auto& g : gt) by_t[g.t_ns] = &g;
  double ss = 0;
  int n = 0;
  for (const auto& e : est) {
    auto it = by_t.find(e.t_ns);
    if (it == by_t.end()) continue;
    for (int i = 0; i < 3; ++i) { const double d = e.p[i] - it->second->p[i]; ss += d * d; }
    ++n;
  }
  if (n_matched) *n_matched = n;
  return n ? std::sqrt(ss / n) : NAN;
}

}  // namespace asl
}  // namespace msckf_b200
