// include/msckf_mono/msckf.h -- drop-in replacement of the reference's msckf_mono::MSCKF<_S>
// (/root/reference/include/msckf_mono/msckf.h:31-1512; "ref :N" below = line N of that file).
//
// Same namespace, class name, template parameter and public member functions.  All integer bookkeeping
// (feature tracks, clone ids, residualisation and pruning decisions) runs on the host exactly as in the
// reference; all floating-point work goes through the C-ABI of libmsckf_b200.so (include/msckf_b200.h)
// into hand-written sm_100a kernels, with the covariance, clone poses and IMU state device resident.
// There is no CPU fallback: a failed device call throws std::runtime_error.
//
// Call-order contract (same as the reference's drivers, src/ros_interface.cpp:92-116,
// datasets/asl_msckf.cpp:233-294): per image propagate()* -> augmentState -> update -> addFeatures ->
// marginalize -> [pruneRedundantStates] -> pruneEmptyStates.  marginalize() residualises the tracks queued
// by the preceding update()/finish() at the clone poses current at that time (the reference snapshots
// copies of the clones in update(); the two coincide unless a measurement update is inserted between
// update() and marginalize(), which this class rejects loudly).
#ifndef MSCKF_HPP_
#define MSCKF_HPP_

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include <msckf_b200.h>
#include <msckf_mono/types.h>

namespace msckf_mono {

template <typename _S>
class MSCKF {
 public:
  MSCKF_B200_ALIGNED_NEW
  MSCKF() {}
  ~MSCKF() { if (engine_) msckf_b200_destroy(engine_); }
  MSCKF(const MSCKF&) = delete;
  MSCKF& operator=(const MSCKF&) = delete;

  // Engine capacities (call before initialize; 0 = derive from MSCKFParams).
  void setEngineOptions(int device, int max_clones = 0, int max_tracks = 0, int max_obs = 0) {
    device_ = device; cap_clones_ = max_clones; cap_tracks_ = max_tracks; cap_obs_ = max_obs;
  }

  // ref :72-97
  void initialize(const Camera<_S>& camera, const noiseParams<_S>& noise_params, const MSCKFParams<_S>& msckf_params,
                  const imuState<_S>& imu_state) {
    camera_ = camera;
    noise_params_ = noise_params;
    msckf_params_ = msckf_params;
    imu_state_ = imu_state;
    imu_state_.p_I_G_null = imu_state_.p_I_G;
    imu_state_.v_I_G_null = imu_state_.v_I_G;
    imu_state_.q_IG_null = imu_state_.q_IG;
    last_feature_id_ = 0;
    feature_tracks_.clear(); tracked_feature_ids_.clear(); feature_tracks_to_residualize_.clear();
    tracks_to_remove_.clear(); cam_states_.clear(); pruned_states_.clear(); map_.clear();
    if (engine_) { msckf_b200_destroy(engine_); engine_ = nullptr; }
    msckf_b200_config cfg;
    cfg.dtype = sizeof(_S) == 4 ? MSCKF_B200_F32 : MSCKF_B200_F64;
    cfg.device = device_;
    const int lcap = std::min(std::max(msckf_params.max_track_length, 2), 98);  // chi-square table has 99 entries (ref :91)
    cfg.max_clones = cap_clones_ > 0 ? cap_clones_ : std::min(std::max(msckf_params.max_cam_states, lcap), 98) + 8;
    cfg.max_tracks = cap_tracks_ > 0 ? cap_tracks_ : 512;
    cfg.max_obs = cap_obs_ > 0 ? cap_obs_ : cfg.max_tracks * lcap;
    check(msckf_b200_create(&cfg, &engine_), "msckf_b200_create");
    _S cam[7] = {camera.q_CI.x(), camera.q_CI.y(), camera.q_CI.z(), camera.q_CI.w(), camera.p_C_I(0), camera.p_C_I(1), camera.p_C_I(2)};
    std::vector<_S> nz(2 + 144 + 225);
    nz[0] = noise_params.u_var_prime; nz[1] = noise_params.v_var_prime;
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) nz[2 + 12 * i + j] = noise_params.Q_imu(i, j);
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) nz[146 + 15 * i + j] = noise_params.initial_imu_covar(i, j);
    _S pr[2] = {msckf_params.max_gn_cost_norm, msckf_params.translation_threshold};
    _S im[19];
    for (int i = 0; i < 3; ++i) {
      im[i] = imu_state.p_I_G(i); im[3 + i] = imu_state.v_I_G(i); im[6 + i] = imu_state.b_g(i);
      im[9 + i] = imu_state.b_a(i); im[12 + i] = imu_state.g(i);
    }
    im[15] = imu_state.q_IG.x(); im[16] = imu_state.q_IG.y(); im[17] = imu_state.q_IG.z(); im[18] = imu_state.q_IG.w();
    check(msckf_b200_initialize(engine_, cam, nz.data(), pr, im), "msckf_b200_initialize");
    state_dirty_ = false;
    epoch_ = 0; epoch_at_queue_ = 0;
  }

  // ref :101-145
  void propagate(imuReading<_S>& measurement_) {
    const _S r[7] = {measurement_.omega(0), measurement_.omega(1), measurement_.omega(2), measurement_.a(0), measurement_.a(1),
                     measurement_.a(2), measurement_.dT};
    check(msckf_b200_propagate(engine_, r), "msckf_b200_propagate");
    state_dirty_ = true;
  }

  // ref :148-212
  void augmentState(const int& state_id, const _S& time) {
    map_.clear();
    check(msckf_b200_augment(engine_), "msckf_b200_augment");
    camState<_S> cam_state;
    cam_state.last_correlated_id = -1;
    cam_state.time = time;
    cam_state.state_id = state_id;
    cam_states_.push_back(cam_state);  // pose filled lazily from the device (refreshState)
    state_dirty_ = true;
  }

  // ref :215-299
  void update(const aligned_vector<Vector2<_S>>& measurements, const std::vector<size_t>& feature_ids) {
    feature_tracks_to_residualize_.clear();
    tracks_to_remove_.clear();
    int id_iter = 0;
    for (auto feature_id : tracked_feature_ids_) {
      auto input_it = std::find(feature_ids.begin(), feature_ids.end(), feature_id);
      const bool is_valid = (input_it != feature_ids.end());
      auto track = feature_tracks_.begin() + id_iter;
      if (is_valid) {
        const size_t d = std::distance(feature_ids.begin(), input_it);
        track->observations.push_back(measurements[d]);
        auto cam_state_iter = cam_states_.end() - 1;
        cam_state_iter->tracked_feature_ids.push_back(feature_id);
        track->cam_state_indices.push_back(cam_state_iter->state_id);
      }
      if (!is_valid || (track->observations.size() >= (size_t)msckf_params_.max_track_length)) {
        featureTrackToResidualize<_S> ttr;
        removeTrackedFeature(feature_id, ttr.cam_state_indices);
        if (ttr.cam_state_indices.size() >= (size_t)msckf_params_.min_track_length) {
          ttr.feature_id = track->feature_id;
          ttr.observations = track->observations;
          ttr.initialized = track->initialized;
          if (track->initialized) ttr.p_f_G = track->p_f_G;
          feature_tracks_to_residualize_.push_back(ttr);
        }
        tracks_to_remove_.push_back(feature_id);
      }
      id_iter++;
    }
    for (auto feature_id : tracks_to_remove_) {
      auto track_iter = feature_tracks_.begin();
      while (track_iter != feature_tracks_.end()) {
        if (track_iter->feature_id == feature_id) {
          const size_t last_id = track_iter->cam_state_indices.back();
          for (size_t index : track_iter->cam_state_indices)
            for (auto& camstate : cam_states_)
              if (!camstate.tracked_feature_ids.size() && (size_t)camstate.state_id == index) camstate.last_correlated_id = (int)last_id;
          track_iter = feature_tracks_.erase(track_iter);
          break;
        } else
          track_iter++;
      }
      auto cid = std::find(tracked_feature_ids_.begin(), tracked_feature_ids_.end(), feature_id);
      if (cid != tracked_feature_ids_.end()) tracked_feature_ids_.erase(cid);
    }
    epoch_at_queue_ = epoch_;
  }

  // ref :302-332
  void addFeatures(const aligned_vector<Vector2<_S>>& features, const std::vector<size_t>& feature_ids) {
    for (size_t i = 0; i < features.size(); i++) {
      const size_t id = feature_ids[i];
      if (std::find(tracked_feature_ids_.begin(), tracked_feature_ids_.end(), id) == tracked_feature_ids_.end()) {
        featureTrack<_S> track;
        track.feature_id = feature_ids[i];
        track.observations.push_back(features[i]);
        auto cam_state_last = cam_states_.end() - 1;
        cam_state_last->tracked_feature_ids.push_back(feature_ids[i]);
        track.cam_state_indices.push_back(cam_state_last->state_id);
        feature_tracks_.push_back(track);
        tracked_feature_ids_.push_back(feature_ids[i]);
      } else {
        std::cout << "Error, added new feature that was already being tracked" << std::endl;
        return;
      }
    }
  }

  // ref :336-449.  Loop A, loop B, gating, compression and the Kalman update all run on the device.
  void marginalize() {
    last_report_.clear();
    if (feature_tracks_to_residualize_.empty()) return;
    if (epoch_at_queue_ != epoch_)
      throw std::logic_error("msckf_mono::MSCKF::marginalize: a measurement update was applied between update()/finish() and "
                             "marginalize(); the queued tracks' clone snapshots (ref :250) would differ from the device state");
    packTracks(feature_tracks_to_residualize_);
    const int N = (int)feature_tracks_to_residualize_.size();
    msckf_b200_tracks tr;
    tr.n_tracks = N; tr.obs_offset = pk_off_.data(); tr.obs = pk_obs_.data(); tr.clone_index = pk_idx_.data(); tr.p_f_G = nullptr;
    last_report_.resize(N);
    rp_cm_.resize(N); rp_tri_.resize(N); rp_valid_.resize(N); rp_acc_.resize(N); rp_gamma_.resize(N); rp_pfg_.resize(3 * (size_t)N);
    msckf_b200_report rep;
    rep.cm_ok = rp_cm_.data(); rep.tri_ok = rp_tri_.data(); rep.valid = rp_valid_.data(); rep.accepted = rp_acc_.data();
    rep.gamma = rp_gamma_.data(); rep.p_f_G = rp_pfg_.data();
    check(msckf_b200_update(engine_, MSCKF_B200_MARGINALIZE, &tr, &rep), "msckf_b200_update(MARGINALIZE)");
    for (int t = 0; t < N; ++t) {
      auto& track = feature_tracks_to_residualize_[t];
      TrackReport& r = last_report_[t];
      r.cm_ok = rp_cm_[t]; r.tri_ok = rp_tri_[t]; r.valid = rp_valid_[t]; r.accepted = rp_acc_[t]; r.gamma = rp_gamma_[t];
      for (int k = 0; k < 3; ++k) r.p_f_G(k) = rp_pfg_[3 * (size_t)t + k];
      r.rows = r.accepted ? (2 * (int)track.observations.size() - 3) : 0;
      if (r.valid) {  // ref :368-372
        track.initialized = true;
        track.p_f_G = r.p_f_G;
        map_.push_back(r.p_f_G);
      }
    }
    last_m_ = rep.m; last_rank_ = rep.rank;
    if (rep.m > 0) epoch_++;
    state_dirty_ = true;
  }

  // ref :453-682
  void pruneRedundantStates() {
    if (cam_states_.size() < 20) return;
    refreshState();
    std::vector<size_t> rm_cam_state_ids;
    findRedundantCamStates(rm_cam_state_ids);
    // ---- first loop (ref :466-534): drop single observations, triangulate uninitialised features
    std::vector<size_t> tri_feat;  // indices into feature_tracks_ that need checkMotion + initializePosition
    std::vector<std::vector<size_t>> tri_involved;
    std::vector<featureTrackToResidualize<_S>> tri_batch;
    for (size_t fi = 0; fi < feature_tracks_.size(); ++fi) {
      auto& feature = feature_tracks_[fi];
      std::vector<size_t> involved;
      size_t obs_id = 0;
      for (const auto& cam_id : rm_cam_state_ids) {
        auto it = std::find(feature.cam_state_indices.begin(), feature.cam_state_indices.end(), cam_id);
        if (it != feature.cam_state_indices.end()) { involved.push_back(cam_id); obs_id = std::distance(feature.cam_state_indices.begin(), it); }
      }
      if (involved.size() == 0) continue;
      if (involved.size() == 1) {
        feature.observations.erase(feature.observations.begin() + obs_id);
        feature.cam_state_indices.erase(feature.cam_state_indices.begin() + obs_id);
        continue;
      }
      if (!feature.initialized) {
        featureTrackToResidualize<_S> q;
        q.feature_id = feature.feature_id;
        q.observations = feature.observations;
        for (size_t pos = 0; pos < cam_states_.size(); ++pos)
          if (std::find(feature.cam_state_indices.begin(), feature.cam_state_indices.end(), (size_t)cam_states_[pos].state_id) !=
              feature.cam_state_indices.end())
            q.cam_state_indices.push_back(pos);
        if (q.cam_state_indices.size() != q.observations.size())
          throw std::logic_error("pruneRedundantStates: observations and associated clones differ in number");
        tri_feat.push_back(fi);
        tri_involved.push_back(involved);
        tri_batch.push_back(q);
      }
    }
    if (!tri_batch.empty()) {
      packTracks(tri_batch);
      const int N = (int)tri_batch.size();
      msckf_b200_tracks tr;
      tr.n_tracks = N; tr.obs_offset = pk_off_.data(); tr.obs = pk_obs_.data(); tr.clone_index = pk_idx_.data(); tr.p_f_G = nullptr;
      rp_cm_.resize(N); rp_tri_.resize(N); rp_pfg_.resize(3 * (size_t)N);
      msckf_b200_report rep = {};
      rep.cm_ok = rp_cm_.data(); rep.tri_ok = rp_tri_.data(); rep.p_f_G = rp_pfg_.data();
      check(msckf_b200_update(engine_, MSCKF_B200_TRIANGULATE, &tr, &rep), "msckf_b200_update(TRIANGULATE)");
      for (int k = 0; k < N; ++k) {
        auto& feature = feature_tracks_[tri_feat[k]];
        // checkMotion on fewer than two clones is false (ref :982-984)
        const bool cm = tri_batch[k].cam_state_indices.size() >= 2 && rp_cm_[k];
        if (!cm || !rp_tri_[k]) {
          for (const auto& cam_id : tri_involved[k]) {
            auto it = std::find(feature.cam_state_indices.begin(), feature.cam_state_indices.end(), cam_id);
            if (it != feature.cam_state_indices.end()) {
              const size_t oi = std::distance(feature.cam_state_indices.begin(), it);
              feature.cam_state_indices.erase(it);
              feature.observations.erase(feature.observations.begin() + oi);
            }
          }
        } else {
          feature.initialized = true;
          for (int c = 0; c < 3; ++c) feature.p_f_G(c) = rp_pfg_[3 * (size_t)k + c];
          map_.push_back(feature.p_f_G);
        }
      }
    }
    // ---- second loop (ref :545-607): residualise the involved observations
    std::vector<featureTrackToResidualize<_S>> batch;
    for (auto& feature : feature_tracks_) {
      std::vector<size_t> involved;
      aligned_vector<Vector2<_S>> involved_obs;
      for (const auto& cam_id : rm_cam_state_ids) {
        auto it = std::find(feature.cam_state_indices.begin(), feature.cam_state_indices.end(), cam_id);
        if (it != feature.cam_state_indices.end()) {
          involved.push_back(cam_id);
          involved_obs.push_back(feature.observations[std::distance(feature.cam_state_indices.begin(), it)]);
        }
      }
      if (involved.size() == 0) continue;
      featureTrackToResidualize<_S> q;
      q.feature_id = feature.feature_id;
      q.observations = involved_obs;
      for (size_t pos = 0; pos < cam_states_.size(); ++pos)
        if (std::find(involved.begin(), involved.end(), (size_t)cam_states_[pos].state_id) != involved.end()) q.cam_state_indices.push_back(pos);
      q.initialized = true;
      q.p_f_G = feature.p_f_G;
      if (q.cam_state_indices.size() != q.observations.size())
        throw std::logic_error("pruneRedundantStates: involved observations and clones differ in number");
      batch.push_back(q);
      for (const auto& cam_id : involved) {
        auto it = std::find(feature.cam_state_indices.begin(), feature.cam_state_indices.end(), cam_id);
        if (it != feature.cam_state_indices.end()) {
          const size_t oi = std::distance(feature.cam_state_indices.begin(), it);
          feature.cam_state_indices.erase(it);
          feature.observations.erase(feature.observations.begin() + oi);
        }
      }
    }
    if (!batch.empty()) {
      packTracks(batch);
      const int N = (int)batch.size();
      std::vector<_S> pf(3 * (size_t)N);
      for (int k = 0; k < N; ++k) for (int c = 0; c < 3; ++c) pf[3 * (size_t)k + c] = batch[k].p_f_G(c);
      msckf_b200_tracks tr;
      tr.n_tracks = N; tr.obs_offset = pk_off_.data(); tr.obs = pk_obs_.data(); tr.clone_index = pk_idx_.data(); tr.p_f_G = pf.data();
      msckf_b200_report rep = {};
      check(msckf_b200_update(engine_, MSCKF_B200_RESIDUALIZE, &tr, &rep), "msckf_b200_update(RESIDUALIZE)");
      if (rep.m > 0) epoch_++;
      state_dirty_ = true;
    }
    // ---- delete the clones (ref :616-681)
    refreshState();
    std::vector<int> keep;
    std::vector<camState<_S>> kept;
    for (size_t pos = 0; pos < cam_states_.size(); ++pos) {
      if (std::find(rm_cam_state_ids.begin(), rm_cam_state_ids.end(), (size_t)cam_states_[pos].state_id) != rm_cam_state_ids.end())
        pruned_states_.push_back(cam_states_[pos]);
      else { keep.push_back((int)pos); kept.push_back(cam_states_[pos]); }
    }
    if (keep.size() != cam_states_.size()) {
      check(msckf_b200_prune(engine_, keep.data(), (int)keep.size()), "msckf_b200_prune");
      cam_states_.swap(kept);
    }
  }

  // ref :685-761
  void pruneEmptyStates() {
    const int max_states = msckf_params_.max_cam_states;
    if ((int)cam_states_.size() < max_states) return;
    const int num_cam_states = (int)cam_states_.size();
    int last_to_remove = num_cam_states - max_states - 1;
    if (cam_states_.front().tracked_feature_ids.size()) return;
    for (int i = 1; i < num_cam_states - max_states; i++)
      if (cam_states_[i].tracked_feature_ids.size()) { last_to_remove = i - 1; break; }
    const int ndel = last_to_remove + 1;
    if (ndel <= 0) return;
    refreshState();
    for (int i = 0; i < ndel; ++i) pruned_states_.push_back(cam_states_[i]);
    cam_states_.erase(cam_states_.begin(), cam_states_.begin() + ndel);
    std::vector<int> keep;
    for (int i = ndel; i < num_cam_states; ++i) keep.push_back(i);
    check(msckf_b200_prune(engine_, keep.data(), (int)keep.size()), "msckf_b200_prune");
  }

  // ref :765-807
  void finish() {
    for (size_t i = 0; i < tracked_feature_ids_.size(); i++) {
      std::vector<size_t> camStateIndices;
      removeTrackedFeature(tracked_feature_ids_[i], camStateIndices);
      if (camStateIndices.size() >= (size_t)msckf_params_.min_track_length) {
        featureTrackToResidualize<_S> track;
        const featureTrack<_S>* src = &feature_tracks_[i];
        if (src->feature_id != tracked_feature_ids_[i])
          for (auto& ft : feature_tracks_)
            if (ft.feature_id == tracked_feature_ids_[i]) { src = &ft; break; }
        track.feature_id = src->feature_id;
        track.observations = src->observations;
        track.initialized = src->initialized;
        if (src->initialized) track.p_f_G = src->p_f_G;
        track.cam_state_indices = camStateIndices;
        feature_tracks_to_residualize_.push_back(track);
      }
      tracks_to_remove_.push_back(tracked_feature_ids_[i]);
    }
    epoch_at_queue_ = epoch_;
    marginalize();
  }

  // ---- getters (ref :810-848), all by value like the reference
  inline size_t getNumCamStates() { return cam_states_.size(); }
  inline imuState<_S> getImuState() { refreshState(); return imu_state_; }
  inline aligned_vector<Vector3<_S>> getMap() { return map_; }
  inline Camera<_S> getCamera() { return camera_; }
  inline camState<_S> getCamState(size_t i) { refreshState(); return cam_states_[i]; }
  inline std::vector<camState<_S>> getCamStates() { refreshState(); return cam_states_; }
  inline std::vector<camState<_S>> getPrunedStates() {
    std::sort(pruned_states_.begin(), pruned_states_.end(), [](const camState<_S>& a, const camState<_S>& b) { return a.state_id < b.state_id; });
    return pruned_states_;
  }

  // ---- extras (not in the reference): diagnostics used by the parity tests and the bench
  struct TrackReport { int cm_ok = 0, tri_ok = 0, valid = 0, accepted = 0, rows = 0; _S gamma = 0; Vector3<_S> p_f_G; };
  const std::vector<TrackReport>& lastReport() const { return last_report_; }
  int lastStackedRows() const { return last_m_; }
  int lastRank() const { return last_rank_; }
  const std::vector<size_t>& trackedFeatureIds() const { return tracked_feature_ids_; }
  const std::vector<featureTrackToResidualize<_S>>& tracksToResidualize() const { return feature_tracks_to_residualize_; }
  msckf_b200_engine* engine() { return engine_; }
  // full (15+6M)^2 covariance, row-major (the reference keeps it private in three blocks, ref :52-54)
  std::vector<_S> getCovariance() {
    const size_t n = 15 + 6 * cam_states_.size();
    std::vector<_S> P(n * n);
    check(msckf_b200_get_covariance(engine_, P.data()) < 0 ? -1 : 0, "msckf_b200_get_covariance");
    return P;
  }
  // marginalize() split in two for pipelining many filters over one GPU: launch, then collect.
  void marginalizeLaunch() {
    last_report_.clear();
    launch_pending_ = false;
    if (feature_tracks_to_residualize_.empty()) return;
    if (epoch_at_queue_ != epoch_) throw std::logic_error("marginalizeLaunch: state changed since update()");
    packTracks(feature_tracks_to_residualize_);
    msckf_b200_tracks tr;
    tr.n_tracks = (int)feature_tracks_to_residualize_.size(); tr.obs_offset = pk_off_.data(); tr.obs = pk_obs_.data();
    tr.clone_index = pk_idx_.data(); tr.p_f_G = nullptr;
    check(msckf_b200_update_async(engine_, MSCKF_B200_MARGINALIZE, &tr), "msckf_b200_update_async");
    launch_pending_ = true;
  }
  void marginalizeCollect() {
    if (!launch_pending_) return;
    launch_pending_ = false;
    const int N = (int)feature_tracks_to_residualize_.size();
    last_report_.resize(N);
    rp_cm_.resize(N); rp_tri_.resize(N); rp_valid_.resize(N); rp_acc_.resize(N); rp_gamma_.resize(N); rp_pfg_.resize(3 * (size_t)N);
    msckf_b200_report rep;
    rep.cm_ok = rp_cm_.data(); rep.tri_ok = rp_tri_.data(); rep.valid = rp_valid_.data(); rep.accepted = rp_acc_.data();
    rep.gamma = rp_gamma_.data(); rep.p_f_G = rp_pfg_.data();
    check(msckf_b200_fetch(engine_, &rep), "msckf_b200_fetch");
    for (int t = 0; t < N; ++t) {
      auto& track = feature_tracks_to_residualize_[t];
      TrackReport& r = last_report_[t];
      r.cm_ok = rp_cm_[t]; r.tri_ok = rp_tri_[t]; r.valid = rp_valid_[t]; r.accepted = rp_acc_[t]; r.gamma = rp_gamma_[t];
      for (int k = 0; k < 3; ++k) r.p_f_G(k) = rp_pfg_[3 * (size_t)t + k];
      r.rows = r.accepted ? (2 * (int)track.observations.size() - 3) : 0;
      if (r.valid) { track.initialized = true; track.p_f_G = r.p_f_G; map_.push_back(r.p_f_G); }
    }
    last_m_ = rep.m; last_rank_ = rep.rank;
    if (rep.m > 0) epoch_++;
    state_dirty_ = true;
  }

 private:
  Camera<_S> camera_;
  noiseParams<_S> noise_params_;
  MSCKFParams<_S> msckf_params_;
  std::vector<featureTrack<_S>> feature_tracks_;
  std::vector<size_t> tracked_feature_ids_;
  std::vector<featureTrackToResidualize<_S>> feature_tracks_to_residualize_;
  std::vector<size_t> tracks_to_remove_;
  size_t last_feature_id_ = 0;
  imuState<_S> imu_state_;             // host mirror of the device state (refreshState)
  std::vector<camState<_S>> cam_states_;  // bookkeeping on host; poses mirrored from the device
  std::vector<camState<_S>> pruned_states_;
  aligned_vector<Vector3<_S>> map_;
  msckf_b200_engine* engine_ = nullptr;
  int device_ = 0, cap_clones_ = 0, cap_tracks_ = 0, cap_obs_ = 0;
  bool state_dirty_ = false, launch_pending_ = false;
  unsigned long long epoch_ = 0, epoch_at_queue_ = 0;
  std::vector<TrackReport> last_report_;
  int last_m_ = 0, last_rank_ = 0;
  std::vector<int> pk_off_, pk_idx_, rp_cm_, rp_tri_, rp_valid_, rp_acc_;
  std::vector<_S> pk_obs_, rp_gamma_, rp_pfg_;

  static void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + msckf_b200_last_error());
  }

  void packTracks(const std::vector<featureTrackToResidualize<_S>>& tracks) {
    const size_t N = tracks.size();
    pk_off_.resize(N + 1);
    size_t tot = 0;
    for (size_t t = 0; t < N; ++t) { pk_off_[t] = (int)tot; tot += tracks[t].observations.size(); }
    pk_off_[N] = (int)tot;
    pk_obs_.resize(2 * tot);
    pk_idx_.resize(tot);
    for (size_t t = 0; t < N; ++t) {
      const auto& tr = tracks[t];
      if (tr.cam_state_indices.size() != tr.observations.size())
        throw std::logic_error("track observations and clone indices differ in number");
      // the reference clears the queue in update() only (:218): finish() after a marginalize() would re-residualise the
      // last update's tracks with their old positional indices (undefined behaviour there once clones were pruned)
      for (size_t ci : tr.cam_state_indices)
        if (ci >= cam_states_.size()) throw std::logic_error("stale residualisation queue: clone index out of range (reference: undefined behaviour)");
      for (size_t i = 0; i < tr.observations.size(); ++i) {
        pk_obs_[2 * (pk_off_[t] + i)] = tr.observations[i](0);
        pk_obs_[2 * (pk_off_[t] + i) + 1] = tr.observations[i](1);
        pk_idx_[pk_off_[t] + i] = (int)tr.cam_state_indices[i];
      }
    }
  }

  // pull the IMU state and clone poses back from the device when something changed them
  void refreshState() {
    if (!state_dirty_ || !engine_) return;
    _S im[29];
    std::vector<_S> poses(7 * std::max<size_t>(cam_states_.size(), 1));
    check(msckf_b200_get_state(engine_, im, poses.data()), "msckf_b200_get_state");
    for (int i = 0; i < 3; ++i) {
      imu_state_.p_I_G(i) = im[i]; imu_state_.v_I_G(i) = im[3 + i]; imu_state_.b_g(i) = im[6 + i]; imu_state_.b_a(i) = im[9 + i];
      imu_state_.g(i) = im[12 + i]; imu_state_.p_I_G_null(i) = im[19 + i]; imu_state_.v_I_G_null(i) = im[22 + i];
    }
    imu_state_.q_IG = Quaternion<_S>(im[18], im[15], im[16], im[17]);
    imu_state_.q_IG_null = Quaternion<_S>(im[28], im[25], im[26], im[27]);
    for (size_t k = 0; k < cam_states_.size(); ++k) {
      for (int i = 0; i < 3; ++i) cam_states_[k].p_C_G(i) = poses[7 * k + i];
      cam_states_[k].q_CG = Quaternion<_S>(poses[7 * k + 6], poses[7 * k + 3], poses[7 * k + 4], poses[7 * k + 5]);
    }
    state_dirty_ = false;
  }

  // ref :1469-1485 (positions only: the clone copies are the device-resident poses)
  void removeTrackedFeature(const size_t featureID, std::vector<size_t>& camStateIndices) {
    camStateIndices.clear();
    for (size_t c_i = 0; c_i < cam_states_.size(); c_i++) {
      auto it = std::find(cam_states_[c_i].tracked_feature_ids.begin(), cam_states_[c_i].tracked_feature_ids.end(), featureID);
      if (it != cam_states_[c_i].tracked_feature_ids.end()) {
        cam_states_[c_i].tracked_feature_ids.erase(it);
        camStateIndices.push_back(c_i);
      }
    }
  }

  static _S angularDistance(const Quaternion<_S>& a, const Quaternion<_S>& b) {  // Eigen 3.3 semantics
    const _S bx = -b.x(), by = -b.y(), bz = -b.z(), bw = b.w();
    const _S dw = a.w() * bw - a.x() * bx - a.y() * by - a.z() * bz;
    const _S dx = a.w() * bx + a.x() * bw + a.y() * bz - a.z() * by;
    const _S dy = a.w() * by + a.y() * bw + a.z() * bx - a.x() * bz;
    const _S dz = a.w() * bz + a.z() * bw + a.x() * by - a.y() * bx;
    return _S(2) * std::atan2(std::sqrt(dx * dx + dy * dy + dz * dz), std::abs(dw));
  }

  // ref :1049-1098 (host: integer logic + two norms per clone)
  void findRedundantCamStates(std::vector<size_t>& rm_cam_state_ids) {
    if (cam_states_.size() < 5) return;
    const _S dist_thresh = msckf_params_.redundancy_distance_thresh;
    const _S angle_thresh = msckf_params_.redundancy_angle_thresh;
    size_t kf = 0;
    size_t next = 1;
    const size_t prot = cam_states_.size() - 3;
    while (next != prot) {
      _S d2 = 0;
      for (int i = 0; i < 3; ++i) { const _S d = cam_states_[next].p_C_G(i) - cam_states_[kf].p_C_G(i); d2 += d * d; }
      const _S distance = std::sqrt(d2);
      const _S angle = angularDistance(cam_states_[kf].q_CG, cam_states_[next].q_CG);
      if (distance < dist_thresh && angle < angle_thresh) rm_cam_state_ids.push_back(cam_states_[next].state_id);
      else kf = next;
      ++next;
      const int num_remaining = (int)(cam_states_.size() - rm_cam_state_ids.size());
      if (num_remaining <= msckf_params_.max_cam_states) break;
    }
    const int num_over_max = (int)(cam_states_.size() - rm_cam_state_ids.size()) - msckf_params_.max_cam_states;
    for (int i = 0; i < num_over_max; i++)
      if (rm_cam_state_ids.end() == std::find(rm_cam_state_ids.begin(), rm_cam_state_ids.end(), (size_t)cam_states_[i].state_id))
        rm_cam_state_ids.push_back(cam_states_[i].state_id);
    if (rm_cam_state_ids.size() < 2) rm_cam_state_ids.clear();
    std::sort(rm_cam_state_ids.begin(), rm_cam_state_ids.end());
  }
};

}  // namespace msckf_mono
#endif /* MSCKF_HPP_ */
