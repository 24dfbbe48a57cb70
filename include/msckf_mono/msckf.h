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
  // The filter owns device memory: it moves (so that filters can live in containers and be returned by value) but does not
  // copy -- the reference's implicit copy duplicated a few KB of host state, here it would have to duplicate a device-resident
  // covariance behind the caller's back (use msckf_b200_copy_state on engine() for an explicit checkpoint).
  MSCKF(const MSCKF&) = delete;
  MSCKF& operator=(const MSCKF&) = delete;
  MSCKF(MSCKF&& o) noexcept { moveFrom(o); }
  MSCKF& operator=(MSCKF&& o) noexcept {
    if (this != &o) {
      if (engine_) msckf_b200_destroy(engine_);
      engine_ = nullptr;
      moveFrom(o);
    }
    return *this;
  }

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
    // initial sizes only: like the reference's std::vectors, the engine grows on demand (window, tracks, observations)
    cfg.max_clones = cap_clones_ > 0 ? cap_clones_ : std::min(std::max(msckf_params.max_cam_states + 8, 32), 104);
    cfg.max_tracks = cap_tracks_ > 0 ? cap_tracks_ : 512;
    cfg.max_obs = cap_obs_ > 0 ? cap_obs_ : cfg.max_tracks * std::min(lcap, 32);
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
    imu_queue_.clear();
    epoch_ = 0; epoch_at_queue_ = 0;
  }

  // ref :101-145.  The reading is queued on the host; the queue is handed to the device in one call
  // (msckf_b200_propagate_n: one kernel launch for up to 16 readings, same arithmetic reading by reading) by whichever
  // member function next needs the device state -- augmentState() in the reference's ROS loop (src/ros_interface.cpp:92-108),
  // getImuState() in its ASL loop (datasets/asl_msckf.cpp:229-234, one reading per call there).
  void propagate(imuReading<_S>& measurement_) {
    const _S r[7] = {measurement_.omega(0), measurement_.omega(1), measurement_.omega(2), measurement_.a(0), measurement_.a(1),
                     measurement_.a(2), measurement_.dT};
    imu_queue_.insert(imu_queue_.end(), r, r + 7);
    state_dirty_ = true;
  }

  // ref :148-212
  void augmentState(const int& state_id, const _S& time) {
    map_.clear();
    check(msckf_b200_augment(eng()), "msckf_b200_augment");
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
    marginalizeLaunch();
    marginalizeCollect();
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
      msckf_b200_tracks tr;
      packTracks(tri_batch, nullptr, tr);
      const int N = (int)tri_batch.size();
      rp_cm_.resize(N); rp_tri_.resize(N); rp_pfg_.resize(3 * (size_t)N);
      msckf_b200_report rep = {};
      rep.cm_ok = rp_cm_.data(); rep.tri_ok = rp_tri_.data(); rep.p_f_G = rp_pfg_.data();
      check(msckf_b200_update(eng(), MSCKF_B200_TRIANGULATE, &tr, &rep), "msckf_b200_update(TRIANGULATE)");
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
      msckf_b200_tracks tr;
      packTracks(batch, &batch, tr);
      msckf_b200_report rep = {};
      checkNumeric(msckf_b200_update(eng(), MSCKF_B200_RESIDUALIZE, &tr, &rep), "msckf_b200_update(RESIDUALIZE)");
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
      check(msckf_b200_prune(eng(), keep.data(), (int)keep.size()), "msckf_b200_prune");
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
    check(msckf_b200_prune(eng(), keep.data(), (int)keep.size()), "msckf_b200_prune");
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
  inline std::vector<camState<_S>> getCamStates() const { refreshState(); return cam_states_; }  // const like ref :835
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
  msckf_b200_engine* engine() { return eng(); }  // (queued IMU readings are handed over first)
  // full (15+6M)^2 covariance, row-major (the reference keeps it private in three blocks, ref :52-54)
  std::vector<_S> getCovariance() {
    const size_t n = 15 + 6 * cam_states_.size();
    std::vector<_S> P(n * n);
    check(msckf_b200_get_covariance(eng(), P.data()) < 0 ? -1 : 0, "msckf_b200_get_covariance");
    return P;
  }
  // marginalize() = launch + collect (kept separate for pipelining and for batches of filters, MSCKFBatch below).
  void marginalizeLaunch() {
    msckf_b200_tracks tr;
    if (!marginalizePrepare(tr, true)) return;
    check(msckf_b200_update_async(eng(), MSCKF_B200_MARGINALIZE, &tr), "msckf_b200_update_async");
    launch_pending_ = true;
  }
  void marginalizeCollect() {
    if (!launch_pending_) return;
    launch_pending_ = false;
    msckf_b200_report rep;
    marginalizeReport(rep);
    checkNumeric(msckf_b200_fetch(engine_, &rep), "msckf_b200_fetch");
    marginalizeAbsorb(rep);
  }
  // pieces of marginalize() for a caller that runs the device part of several filters as one batch (MSCKFBatch):
  // Prepare packs the queued tracks (false: nothing queued), Report points a report at this filter's buffers, Absorb
  // folds the fetched report back into the host bookkeeping (ref :368-372).
  bool marginalizePrepare(msckf_b200_tracks& tr, bool in_engine_buffer) {
    last_report_.clear();
    launch_pending_ = false;
    if (feature_tracks_to_residualize_.empty()) return false;
    if (epoch_at_queue_ != epoch_)
      throw std::logic_error("msckf_mono::MSCKF::marginalize: a measurement update was applied between update()/finish() and "
                             "marginalize(); the queued tracks' clone snapshots (ref :250) would differ from the device state");
    packTracks(feature_tracks_to_residualize_, nullptr, tr, in_engine_buffer);
    return true;
  }
  void marginalizeReport(msckf_b200_report& rep) {
    const int N = (int)feature_tracks_to_residualize_.size();
    rp_cm_.resize(N); rp_tri_.resize(N); rp_valid_.resize(N); rp_acc_.resize(N); rp_gamma_.resize(N); rp_pfg_.resize(3 * (size_t)N);
    rep.cm_ok = rp_cm_.data(); rep.tri_ok = rp_tri_.data(); rep.valid = rp_valid_.data(); rep.accepted = rp_acc_.data();
    rep.gamma = rp_gamma_.data(); rep.p_f_G = rp_pfg_.data(); rep.m = 0; rep.rank = 0;
  }
  void marginalizeAbsorb(const msckf_b200_report& rep) {
    const int N = (int)feature_tracks_to_residualize_.size();
    last_report_.resize(N);
    for (int t = 0; t < N; ++t) {
      auto& track = feature_tracks_to_residualize_[t];
      TrackReport& r = last_report_[t];
      r.cm_ok = rp_cm_[t]; r.tri_ok = rp_tri_[t]; r.valid = rp_valid_[t]; r.accepted = rp_acc_[t]; r.gamma = rp_gamma_[t];
      for (int k = 0; k < 3; ++k) r.p_f_G(k) = rp_pfg_[3 * (size_t)t + k];
      r.rows = r.accepted ? (2 * (int)track.observations.size() - 3) : 0;
      if (r.valid) { track.initialized = true; track.p_f_G = r.p_f_G; map_.push_back(r.p_f_G); }  // ref :368-372
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
  mutable imuState<_S> imu_state_;             // host mirror of the device state (refreshState)
  mutable std::vector<camState<_S>> cam_states_;  // bookkeeping on host; poses mirrored from the device
  std::vector<camState<_S>> pruned_states_;
  aligned_vector<Vector3<_S>> map_;
  msckf_b200_engine* engine_ = nullptr;
  int device_ = 0, cap_clones_ = 0, cap_tracks_ = 0, cap_obs_ = 0;
  mutable bool state_dirty_ = false;
  bool launch_pending_ = false;
  mutable std::vector<_S> imu_queue_;  // IMU readings not yet handed to the device (7 scalars each)
  unsigned long long epoch_ = 0, epoch_at_queue_ = 0;
  std::vector<TrackReport> last_report_;
  int last_m_ = 0, last_rank_ = 0;
  std::vector<int> pk_off_, pk_idx_, rp_cm_, rp_tri_, rp_valid_, rp_acc_;
  std::vector<_S> pk_obs_, pk_pfg_, rp_gamma_, rp_pfg_;

  static void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + msckf_b200_last_error());
  }
  // a non-finite update is not an error in the reference: its only anomaly handling is a print (ref :1405-1409)
  static void checkNumeric(int rc, const char* what) {
    if (rc == MSCKF_B200_ERR_NUMERIC) { std::cout << "[msckf_b200] " << what << ": " << msckf_b200_last_error() << std::endl; return; }
    check(rc, what);
  }
  void moveFrom(MSCKF& o) {
    camera_ = o.camera_; noise_params_ = o.noise_params_; msckf_params_ = o.msckf_params_;
    feature_tracks_ = std::move(o.feature_tracks_); tracked_feature_ids_ = std::move(o.tracked_feature_ids_);
    feature_tracks_to_residualize_ = std::move(o.feature_tracks_to_residualize_); tracks_to_remove_ = std::move(o.tracks_to_remove_);
    last_feature_id_ = o.last_feature_id_; imu_state_ = o.imu_state_; cam_states_ = std::move(o.cam_states_);
    pruned_states_ = std::move(o.pruned_states_); map_ = std::move(o.map_);
    engine_ = o.engine_; o.engine_ = nullptr;
    device_ = o.device_; cap_clones_ = o.cap_clones_; cap_tracks_ = o.cap_tracks_; cap_obs_ = o.cap_obs_;
    state_dirty_ = o.state_dirty_; launch_pending_ = o.launch_pending_; imu_queue_ = std::move(o.imu_queue_);
    epoch_ = o.epoch_; epoch_at_queue_ = o.epoch_at_queue_; last_report_ = std::move(o.last_report_);
    last_m_ = o.last_m_; last_rank_ = o.last_rank_;
  }
  // the engine handle for a call that needs the device state: queued IMU readings are applied first
  msckf_b200_engine* eng() const {
    if (!imu_queue_.empty()) {
      const int k = (int)(imu_queue_.size() / 7);
      const int rc = msckf_b200_propagate_n(engine_, imu_queue_.data(), k);
      imu_queue_.clear();
      check(rc, "msckf_b200_propagate_n");
    }
    return engine_;
  }

  // Flat SoA form of a track batch.  in_engine_buffer: packed straight into the engine's pinned input block
  // (msckf_b200_input_buffer: no second copy on the way to the device); otherwise into this object's own vectors.
  // with_pfg: the tracks' positions ride along (RESIDUALIZE).
  void packTracks(const std::vector<featureTrackToResidualize<_S>>& tracks, const std::vector<featureTrackToResidualize<_S>>* with_pfg,
                  msckf_b200_tracks& tr, bool in_engine_buffer = true) {
    const size_t N = tracks.size();
    size_t tot = 0;
    for (size_t t = 0; t < N; ++t) tot += tracks[t].observations.size();
    int* off; int* idx; _S* obs; _S* pfg = nullptr;
    if (in_engine_buffer) {
      check(msckf_b200_input_buffer(eng(), (int)N, (int)tot, &tr), "msckf_b200_input_buffer");
      off = const_cast<int*>(tr.obs_offset); idx = const_cast<int*>(tr.clone_index);
      obs = static_cast<_S*>(const_cast<void*>(tr.obs)); pfg = static_cast<_S*>(const_cast<void*>(tr.p_f_G));
    } else {
      pk_off_.resize(N + 1); pk_obs_.resize(2 * tot); pk_idx_.resize(tot); pk_pfg_.resize(3 * N);
      off = pk_off_.data(); idx = pk_idx_.data(); obs = pk_obs_.data(); pfg = pk_pfg_.data();
      tr.n_tracks = (int)N; tr.obs_offset = off; tr.obs = obs; tr.clone_index = idx; tr.p_f_G = pfg;
    }
    if (!with_pfg) tr.p_f_G = nullptr;
    tot = 0;
    for (size_t t = 0; t < N; ++t) { off[t] = (int)tot; tot += tracks[t].observations.size(); }
    off[N] = (int)tot;
    for (size_t t = 0; t < N; ++t) {
      const auto& trk = tracks[t];
      if (trk.cam_state_indices.size() != trk.observations.size())
        throw std::logic_error("track observations and clone indices differ in number");
      // the reference clears the queue in update() only (:218): finish() after a marginalize() would re-residualise the
      // last update's tracks with their old positional indices (undefined behaviour there once clones were pruned)
      for (size_t ci : trk.cam_state_indices)
        if (ci >= cam_states_.size()) throw std::logic_error("stale residualisation queue: clone index out of range (reference: undefined behaviour)");
      for (size_t i = 0; i < trk.observations.size(); ++i) {
        obs[2 * (off[t] + i)] = trk.observations[i](0);
        obs[2 * (off[t] + i) + 1] = trk.observations[i](1);
        idx[off[t] + i] = (int)trk.cam_state_indices[i];
      }
      if (with_pfg) for (int c = 0; c < 3; ++c) pfg[3 * t + c] = (*with_pfg)[t].p_f_G(c);
    }
  }

  // pull the IMU state and clone poses back from the device when something changed them
  void refreshState() const {
    if (!state_dirty_ || !engine_) return;
    _S im[29];
    std::vector<_S> poses(7 * std::max<size_t>(cam_states_.size(), 1));
    check(msckf_b200_get_state(eng(), im, poses.data()), "msckf_b200_get_state");
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

// Not in the reference: marginalize() on several independent filters (sequences, Monte-Carlo trials) as ONE device batch --
// one launch per kernel with the filter index in blockIdx.z, one CUDA graph, one packed copy each way
// (msckf_b200_batch_* in include/msckf_b200.h).  Results are bit-identical to calling marginalize() on each filter.
// While the object lives the filters share its CUDA stream; all their other member functions stay usable.
template <typename _S>
class MSCKFBatch {
 public:
  explicit MSCKFBatch(const std::vector<MSCKF<_S>*>& filters, int host_threads = 1) : filters_(filters), threads_(host_threads) {
    std::vector<msckf_b200_engine*> es;
    for (auto* f : filters_) es.push_back(f->engine());
    if (msckf_b200_batch_create(es.data(), (int)es.size(), &batch_) != 0)
      throw std::runtime_error(std::string("msckf_b200_batch_create failed: ") + msckf_b200_last_error());
  }
  ~MSCKFBatch() { if (batch_) msckf_b200_batch_destroy(batch_); }
  MSCKFBatch(const MSCKFBatch&) = delete;
  MSCKFBatch& operator=(const MSCKFBatch&) = delete;
  void marginalize() {
    const size_t n = filters_.size();
    tracks_.assign(n, msckf_b200_tracks());
    reports_.assign(n, msckf_b200_report());
    active_.assign(n, 0);
    bool any = false;
    for (size_t i = 0; i < n; ++i) {
      filters_[i]->engine();  // hand queued IMU readings over before the batch runs
      active_[i] = filters_[i]->marginalizePrepare(tracks_[i], false) ? 1 : 0;
      if (active_[i]) { filters_[i]->marginalizeReport(reports_[i]); any = true; } else tracks_[i].n_tracks = 0;
    }
    if (!any) return;
    const int rc = msckf_b200_batch_update(batch_, MSCKF_B200_MARGINALIZE, tracks_.data(), reports_.data(), threads_);
    if (rc == MSCKF_B200_ERR_NUMERIC) std::cout << "[msckf_b200] batch update: " << msckf_b200_last_error() << std::endl;
    else if (rc != 0) throw std::runtime_error(std::string("msckf_b200_batch_update failed: ") + msckf_b200_last_error());
    for (size_t i = 0; i < n; ++i)
      if (active_[i]) filters_[i]->marginalizeAbsorb(reports_[i]);
  }
  msckf_b200_batch* handle() { return batch_; }

 private:
  std::vector<MSCKF<_S>*> filters_;
  std::vector<msckf_b200_tracks> tracks_;
  std::vector<msckf_b200_report> reports_;
  std::vector<char> active_;
  msckf_b200_batch* batch_ = nullptr;
  int threads_ = 1;
};

}  // namespace msckf_mono
#endif /* MSCKF_HPP_ */
