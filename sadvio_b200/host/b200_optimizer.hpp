// Host-side C++ adapter above the C ABI (include/sdv.h): mirrors the reference optimizer interface for the entry
// points this repository replaces,
//     bool isae::AOptimizer::localMapBA(std::shared_ptr<LocalMap>&, size_t fixed_frame_number = 0)
//     bool isae::AOptimizer::localMapVIOptimization(std::shared_ptr<LocalMap>&, size_t fixed_frame_number = 0)
//     bool isae::AOptimizer::landmarkOptimization(std::shared_ptr<Frame>&)          (masks of the window solve)
//     bool isae::AOptimizer::singleFrameOptimization(std::shared_ptr<Frame>&)
//     bool isae::AOptimizer::singleFrameVIOptimization(std::shared_ptr<Frame>&)
//     bool isae::AOptimizer::marginalize(std::shared_ptr<Frame>& frame0, std::shared_ptr<Frame>& frame1, bool enable_sparsif)
// (reference cpp/include/isaeslam/optimizers/AOptimizer.h:22-30), same argument meaning and error behaviour
// (bool, no exceptions; when no solve could run — no device, malformed window — the state is left untouched and the call
// returns false; a solve that ends in Ceres' FAILURE termination writes back and returns true like the reference).
//
// The reference data model (isae::Frame / ImageSensor / IMU / ALandmark / AFeature / LocalMap, Eigen based) is not
// available in this image, so this header carries a minimal mirror of the pointer graph with the reference's getter
// names; INTEGRATION.md shows the same adapter written against the real isae:: types.  What matters — and what
// tests/test_host_adapter.py pins — is the FLATTENING ORDER, which must reproduce the walk of
//     addResidualsLocalMap   (AngularAdjustmentCERESAnalytic.cpp:212-339)
//     addIMUResiduals        (AOptimizer.cpp:22-96)
// so that the (landmark, frame, camera) visibility triplets are bit-exact, and the WRITE-BACK of AOptimizer.cpp:391-434.
#pragma once
#include "../../include/sdv.h"

#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <unordered_map>
#include <vector>

namespace sdvhost {

using Mat34 = std::array<double, 12>; // row-major [R | t]

struct Frame;
struct Landmark;

struct ImageSensor { // isae::ImageSensor / Camera
    std::weak_ptr<Frame> frame;
    Mat34 T_s_f{};              // getFrame2SensorTransform()
    std::array<double, 4> K{};  // fx, fy, cx, cy
    std::shared_ptr<Frame> getFrame() const { return frame.lock(); }
    double getFocal() const { return (K[0] + K[1]) / 2; } // Camera.h:46
};

struct IMU { // isae::IMU (only what the window solve reads or writes)
    std::array<double, 3> v{}, ba{}, bg{};
    std::array<double, 9> delta_R{{1, 0, 0, 0, 1, 0, 0, 0, 1}};
    std::array<double, 3> delta_v{}, delta_p{};
    std::array<double, 81> Sigma{};
    std::array<double, 9> J_dR_bg{}, J_dv_ba{}, J_dv_bg{}, J_dp_ba{}, J_dp_bg{};
    double bacc_noise = 0, bgyr_noise = 0;
    std::shared_ptr<Frame> last_kf; // getLastKF()
};

struct Frame { // isae::Frame
    Mat34 T_f_w{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}};
    uint64_t timestamp_ns = 0;
    bool keyframe = false, has_prior = false;
    Mat34 T_prior{};
    std::array<double, 6> inf_prior{};
    std::vector<std::shared_ptr<ImageSensor>> sensors;
    std::shared_ptr<IMU> imu; // getIMU(), may be null
    std::vector<std::shared_ptr<Landmark>> pointxd; // getLandmarks()["pointxd"]: the landmarks this frame observes
    bool isKeyFrame() const { return keyframe; }
};

struct Feature { // isae::AFeature
    std::weak_ptr<ImageSensor> sensor;
    std::array<double, 3> bearing{}; // getBearingVectors().at(0)
    std::array<double, 2> uv{};      // getPoints().at(0)
    double sigma = 1.0;              // getSigma()
};

struct Landmark { // isae::ALandmark ("pointxd")
    std::array<double, 3> t_w{};
    bool initialized = true, outlier = false;
    bool in_map = true, has_prior = false, is_marg = false; // isInMap(), hasPrior() / setPrior(), isMarg() / setMarg() (ALandmark.h:96-117)
    std::vector<std::weak_ptr<Feature>> features; // getFeatures()
    bool isInitialized() const { return initialized; }
    bool isOutlier() const { return outlier; }
    inline bool sanityCheck(); // ALandmark.cpp:130-147 (defined below)
};

struct LocalMap { // isae::LocalMap
    std::deque<std::shared_ptr<Frame>> frames;            // oldest -> newest (localmap.cpp:9-26)
    std::vector<std::shared_ptr<Landmark>> pointxd;       // getLandmarks()["pointxd"]
    // amap.h:28-32: newest first
    void getLastNFramesIn(size_t n, std::vector<std::shared_ptr<Frame>> &out) const {
        for (size_t k = 0; k < n && k < frames.size(); k++) out.push_back(frames[frames.size() - 1 - k]);
    }
    size_t getMapSize() const { return frames.size(); }
};

// isae::Marginalization — only the members addMarginalizationResiduals reads (marginalization.hpp:54-85).  Matrices are
// row-major here; the Eigen members of the reference are column-major (INTEGRATION.md shows the copy).
struct Marginalization {
    int n = 0, n_full = 0;                                              // _n, _n_full
    std::shared_ptr<Frame> frame_to_keep;                               // _frame_to_keep (null in the VO case)
    std::vector<std::shared_ptr<Landmark>> lmk_to_keep;                 // _lmk_to_keep["pointxd"]
    size_t n_landmark_types = 1;                                        // _lmk_to_keep.size(): landmark TYPES with kept landmarks
    std::unordered_map<const Frame *, int> map_frame_idx;               // _map_frame_idx
    std::unordered_map<const Landmark *, int> map_lmk_idx;              // _map_lmk_idx
    std::vector<double> marginalization_jacobian;                       // [n_full][n]
    std::vector<double> marginalization_residual;                       // [n_full]
    std::unordered_map<const Frame *, std::array<double, 225>> map_frame_inf;  // _map_frame_inf (sparsifyVIO)
    std::unordered_map<const Landmark *, std::array<double, 9>> map_lmk_inf;   // _map_lmk_inf
    std::unordered_map<const Landmark *, std::array<double, 3>> map_lmk_prior; // _map_lmk_prior
    std::shared_ptr<Landmark> lmk_with_prior;                           // _lmk_with_prior (sparsifyVO)
    std::array<double, 9> info_lmk{};                                   // _info_lmk
    std::array<double, 3> prior_lmk{};                                  // _prior_lmk
};

// Structure-of-arrays image of one window + the bookkeeping needed for the write-back.
struct FlatWindow {
    std::vector<double> T_f_w, v, ba, bg, T_prior, inf_prior, T_s_f, K, lmk_t, obs_bearing, obs_uv, obs_sigma;
    std::vector<uint8_t> has_imu, has_prior;
    std::vector<int32_t> obs_lmk, obs_frame, obs_cam, imu_i, imu_j;
    std::vector<double> imu_dt, imu_dR, imu_dv, imu_dp, imu_cov, J_dR_bg, J_dv_ba, J_dv_bg, J_dp_ba, J_dp_bg, sigma_ba, sigma_bg;
    std::vector<std::shared_ptr<Frame>> frame_vector;     // newest -> oldest
    std::vector<std::shared_ptr<Landmark>> landmarks;     // in parameter-block order
    std::vector<std::shared_ptr<Frame>> imu_frame_j;      // frame j of each IMU factor
    // biasDeltaCorrection list (AOptimizer.cpp:421-434): EVERY frame with an IMU whose getLastKF() owns dba/dbg blocks,
    // i.e. is in the window with an IMU — no dt <= 1 s test, no framei != framej test (unlike the factor list above)
    std::vector<std::shared_ptr<Frame>> corr_frame;
    std::vector<int32_t> corr_prev;                       // window index of its previous keyframe
    std::vector<int32_t> keep_lmk, keep_col, p2l_lmk, l2l_a, l2l_b;  // marginal prior (addMarginalizationResiduals)
    std::vector<double> p2l_delta, p2l_sqrt_inf, l2l_delta, l2l_sqrt_inf;
    sdv_dense_prior dense{};
    sdv_sparse_prior sparse{};
    sdv_window view{};
};

namespace detail {
// distinct sensor models -> camera table
inline int cam_of(FlatWindow &fw, const ImageSensor &s) {
    for (size_t c = 0; c * 12 < fw.T_s_f.size(); c++)
        if (std::memcmp(&fw.T_s_f[12 * c], s.T_s_f.data(), 96) == 0 && std::memcmp(&fw.K[4 * c], s.K.data(), 32) == 0) return (int)c;
    fw.T_s_f.insert(fw.T_s_f.end(), s.T_s_f.begin(), s.T_s_f.end());
    fw.K.insert(fw.K.end(), s.K.begin(), s.K.end());
    return (int)(fw.K.size() / 4 - 1);
}
// per-frame arrays of one frame (the caller appends it to frame_vector)
inline void push_frame_state(FlatWindow &fw, const Frame &f, bool with_prior) {
    fw.T_f_w.insert(fw.T_f_w.end(), f.T_f_w.begin(), f.T_f_w.end());
    fw.has_prior.push_back(with_prior && f.has_prior ? 1 : 0); // …Analytic.cpp:239
    fw.T_prior.insert(fw.T_prior.end(), f.T_prior.begin(), f.T_prior.end());
    fw.inf_prior.insert(fw.inf_prior.end(), f.inf_prior.begin(), f.inf_prior.end());
    const IMU *imu = f.imu.get();
    fw.has_imu.push_back(imu ? 1 : 0);
    for (int k = 0; k < 3; k++) {
        fw.v.push_back(imu ? imu->v[k] : 0.0);
        fw.ba.push_back(imu ? imu->ba[k] : 0.0);
        fw.bg.push_back(imu ? imu->bg[k] : 0.0);
    }
    for (auto &s : f.sensors) cam_of(fw, *s);
}
// IMUFactor + IMUBiasFactor list of addIMUResiduals (AOptimizer.cpp:55-94) over fw.frame_vector
// (dt_test = false: the pair list of VIInit, AOptimizer.cpp:485-502, which has no dt <= 1 s test)
inline void imu_factors(FlatWindow &fw, const std::unordered_map<const Frame *, int> &frame_idx, bool dt_test = true) {
    const int F = (int)fw.frame_vector.size();
    for (int i = 0; i < F; i++) {
        const std::shared_ptr<Frame> &framej = fw.frame_vector[i];
        if (!framej->imu) continue;                                   // :60
        std::shared_ptr<Frame> framei = framej->imu->last_kf;         // :62
        if (!framei) continue;                                        // :65
        if (dt_test && (double)(framej->timestamp_ns - framei->timestamp_ns) * 1e-9 > 1) continue; // :69
        auto it = frame_idx.find(framei.get());
        if (it == frame_idx.end() || !framei->imu || framei == framej) continue; // :72
        const IMU &m = *framej->imu;
        fw.imu_i.push_back(it->second);
        fw.imu_j.push_back(i);
        fw.imu_dt.push_back((double)(framej->timestamp_ns - framei->timestamp_ns) * 1e-9);
        fw.imu_dR.insert(fw.imu_dR.end(), m.delta_R.begin(), m.delta_R.end());
        fw.imu_dv.insert(fw.imu_dv.end(), m.delta_v.begin(), m.delta_v.end());
        fw.imu_dp.insert(fw.imu_dp.end(), m.delta_p.begin(), m.delta_p.end());
        fw.imu_cov.insert(fw.imu_cov.end(), m.Sigma.begin(), m.Sigma.end());
        fw.J_dR_bg.insert(fw.J_dR_bg.end(), m.J_dR_bg.begin(), m.J_dR_bg.end());
        fw.J_dv_ba.insert(fw.J_dv_ba.end(), m.J_dv_ba.begin(), m.J_dv_ba.end());
        fw.J_dv_bg.insert(fw.J_dv_bg.end(), m.J_dv_bg.begin(), m.J_dv_bg.end());
        fw.J_dp_ba.insert(fw.J_dp_ba.end(), m.J_dp_ba.begin(), m.J_dp_ba.end());
        fw.J_dp_bg.insert(fw.J_dp_bg.end(), m.J_dp_bg.begin(), m.J_dp_bg.end());
        fw.sigma_ba.push_back(framei->imu->bacc_noise); // residuals.hpp:259 reads imu_i's config
        fw.sigma_bg.push_back(framei->imu->bgyr_noise); // residuals.hpp:261
        fw.imu_frame_j.push_back(framej);
    }
}
// the sdv_window view over the arrays of fw
inline void fill_view(FlatWindow &fw, bool vio, int factor_kind, size_t fixed_frame_number, bool have_dense, bool have_sparse) {
    sdv_window &w = fw.view;
    std::memset(&w, 0, sizeof(w));
    w.abi_version = SDV_ABI_VERSION;
    w.vio = vio ? 1 : 0;
    w.factor_kind = factor_kind;
    w.n_frames = (int32_t)fw.frame_vector.size();
    w.n_fixed = (int32_t)fixed_frame_number;
    w.n_cams = (int32_t)(fw.K.size() / 4);
    w.n_lmks = (int32_t)fw.landmarks.size();
    w.n_obs = (int32_t)fw.obs_lmk.size();
    w.n_imu = (int32_t)fw.imu_i.size();
    w.T_f_w = fw.T_f_w.data();
    w.v = fw.v.data(); w.ba = fw.ba.data(); w.bg = fw.bg.data();
    w.has_imu = fw.has_imu.data(); w.has_prior = fw.has_prior.data();
    w.T_prior = fw.T_prior.data(); w.inf_prior = fw.inf_prior.data();
    w.T_s_f = fw.T_s_f.data(); w.K = fw.K.data(); w.lmk_t = fw.lmk_t.data();
    w.obs_lmk = fw.obs_lmk.data(); w.obs_frame = fw.obs_frame.data(); w.obs_cam = fw.obs_cam.data();
    w.obs_bearing = fw.obs_bearing.data(); w.obs_uv = fw.obs_uv.data();
    w.obs_sigma = fw.obs_sigma.empty() ? nullptr : fw.obs_sigma.data();
    w.imu_i = fw.imu_i.data(); w.imu_j = fw.imu_j.data(); w.imu_dt = fw.imu_dt.data();
    w.imu_dR = fw.imu_dR.data(); w.imu_dv = fw.imu_dv.data(); w.imu_dp = fw.imu_dp.data(); w.imu_cov = fw.imu_cov.data();
    w.imu_J_dR_bg = fw.J_dR_bg.data(); w.imu_J_dv_ba = fw.J_dv_ba.data(); w.imu_J_dv_bg = fw.J_dv_bg.data();
    w.imu_J_dp_ba = fw.J_dp_ba.data(); w.imu_J_dp_bg = fw.J_dp_bg.data();
    w.imu_sigma_ba = fw.sigma_ba.data(); w.imu_sigma_bg = fw.sigma_bg.data();
    w.dense_prior = have_dense ? &fw.dense : nullptr;
    w.sparse_prior = have_sparse ? &fw.sparse : nullptr;
}
} // namespace detail

// Returns false where the reference would throw out of an unordered_map::at (a prior that names a frame outside the
// window, or a sparsified prior with missing per-landmark entries): the caller maps that to `false`, state untouched.
inline bool flatten(const LocalMap &map, size_t fixed_frame_number, bool vio, int factor_kind, FlatWindow &fw,
                    const Marginalization *marg = nullptr, bool enable_sparsif = false) {
    fw = FlatWindow();
    map.getLastNFramesIn(map.getMapSize(), fw.frame_vector); // AOptimizer.cpp:366-367
    const int F = (int)fw.frame_vector.size();
    std::unordered_map<const Frame *, int> frame_idx; // _map_frame_posepar (…Analytic.cpp:224-226)
    for (int i = 0; i < F; i++) frame_idx[fw.frame_vector[i].get()] = i;
    auto cam_of = [&](const ImageSensor &s) { return detail::cam_of(fw, s); };
    for (int i = 0; i < F; i++) detail::push_frame_state(fw, *fw.frame_vector[i], true);
    // landmarks + visual residual blocks in reference walk order (…Analytic.cpp:247-289)
    std::unordered_map<const Landmark *, int> lmk_idx; // _map_lmk_ptpar
    for (auto &landmark : map.pointxd) {
        if (!landmark->isInitialized() || landmark->isOutlier()) continue; // :254
        const int l = (int)fw.landmarks.size();
        lmk_idx[landmark.get()] = l;
        fw.landmarks.push_back(landmark);
        fw.lmk_t.insert(fw.lmk_t.end(), landmark->t_w.begin(), landmark->t_w.end());
        for (auto &wfeature : landmark->features) { // :266
            std::shared_ptr<Feature> feature = wfeature.lock();
            if (!feature) continue;
            std::shared_ptr<ImageSensor> cam = feature->sensor.lock();
            if (!cam) continue;
            std::shared_ptr<Frame> frame = cam->getFrame();
            if (!frame || !frame->isKeyFrame() || frame_idx.find(frame.get()) == frame_idx.end()) continue; // :272-275
            fw.obs_lmk.push_back(l);
            fw.obs_frame.push_back(frame_idx[frame.get()]);
            fw.obs_cam.push_back(cam_of(*cam));
            fw.obs_bearing.insert(fw.obs_bearing.end(), feature->bearing.begin(), feature->bearing.end());
            fw.obs_uv.insert(fw.obs_uv.end(), feature->uv.begin(), feature->uv.end());
        }
    }
    // IMU factors (AOptimizer.cpp:55-94)
    if (vio) {
        detail::imu_factors(fw, frame_idx);
        for (int i = 0; i < F; i++) { // AOptimizer.cpp:421-434
            const std::shared_ptr<Frame> &frame = fw.frame_vector[i];
            if (!frame->imu || !frame->imu->last_kf) continue;            // :422-426
            auto it = frame_idx.find(frame->imu->last_kf.get());
            if (it == frame_idx.end() || !frame->imu->last_kf->imu) continue; // _map_frame_dbapar.find(previous_frame), :430
            fw.corr_frame.push_back(frame);
            fw.corr_prev.push_back(it->second);
        }
    }
    // Marginal prior (addMarginalizationResiduals, AngularAdjustmentCERESAnalytic.cpp:341-486).  A kept landmark that has no
    // parameter block yet gets one, at zero, with no observation (the "supposed to be in the map" branches, :369-373,
    // :413-417, :441-445); it is written back like any other landmark.
    bool have_dense = false, have_sparse = false;
    // BundleAdjustmentCERESAnalytic wires the sparsified factors only if `_lmk_to_keep.size() > 1` — the number of landmark
    // TYPES in the typed map, not of landmarks (BundleAdjustmentCERESAnalytic.cpp:364): with point landmarks alone the pixel
    // optimizer adds no prior at all when sparsification is on.  Reproduced as is.
    const bool wired = marg && !marg->lmk_to_keep.empty() &&
                       !(enable_sparsif && factor_kind == SDV_FACTOR_PIXEL && marg->n_landmark_types <= 1);
    if (wired) {
        auto block_of = [&](const std::shared_ptr<Landmark> &lmk) {
            auto it = lmk_idx.find(lmk.get());
            if (it != lmk_idx.end()) return it->second;
            const int l = (int)fw.landmarks.size();
            lmk_idx[lmk.get()] = l;
            fw.landmarks.push_back(lmk);
            fw.lmk_t.insert(fw.lmk_t.end(), lmk->t_w.begin(), lmk->t_w.end());
            return l;
        };
        int keep_frame = -1;
        if (marg->frame_to_keep) {
            auto it = frame_idx.find(marg->frame_to_keep.get());
            if (it == frame_idx.end() || (vio && !marg->frame_to_keep->imu)) return false; // _map_frame_posepar.at / _map_frame_velpar.at
            keep_frame = it->second;
        }
        if (!enable_sparsif) { // dense MarginalizationFactor, :346-384
            sdv_dense_prior &dp = fw.dense;
            dp.n_full = marg->n_full;
            dp.n = marg->n;
            dp.J = marg->marginalization_jacobian.data();
            dp.r0 = marg->marginalization_residual.data();
            dp.frame = keep_frame;
            dp.frame_col = 0;
            if (marg->frame_to_keep) {
                auto it = marg->map_frame_idx.find(marg->frame_to_keep.get());
                if (it == marg->map_frame_idx.end()) return false;
                dp.frame_col = it->second;
            }
            for (auto &lmk : marg->lmk_to_keep) { // parameter-block order of the factor, marginalization.hpp:101-107
                auto it = marg->map_lmk_idx.find(lmk.get());
                if (it == marg->map_lmk_idx.end()) return false;
                fw.keep_lmk.push_back(block_of(lmk));
                fw.keep_col.push_back(it->second);
            }
            dp.n_keep = (int32_t)fw.keep_lmk.size();
            dp.keep_lmk = fw.keep_lmk.data();
            dp.keep_col = fw.keep_col.data();
            have_dense = true;
        } else if (marg->frame_to_keep) { // sparsified, VIO case, :390-424
            const Frame &f = *marg->frame_to_keep;
            auto inf = marg->map_frame_inf.find(&f);
            if (inf == marg->map_frame_inf.end() || !f.imu) return false;
            sdv_sparse_prior &sp = fw.sparse;
            sp.has_imu_prior = 1;
            sp.frame = keep_frame;
            std::memcpy(sp.T_prior, f.T_f_w.data(), sizeof(sp.T_prior)); // IMUPriordx(T_f_w, T_f_w, v, v, ba, ba, bg, bg, …), :396-397
            std::memcpy(sp.v_prior, f.imu->v.data(), 24);
            std::memcpy(sp.ba_prior, f.imu->ba.data(), 24);
            std::memcpy(sp.bg_prior, f.imu->bg.data(), 24);
            std::memcpy(sp.imu_sqrt_inf, inf->second.data(), sizeof(sp.imu_sqrt_inf));
            for (auto &lmk : marg->lmk_to_keep) {
                auto d = marg->map_lmk_prior.find(lmk.get());
                auto w = marg->map_lmk_inf.find(lmk.get());
                if (d == marg->map_lmk_prior.end() || w == marg->map_lmk_inf.end()) return false;
                fw.p2l_lmk.push_back(block_of(lmk));
                fw.p2l_delta.insert(fw.p2l_delta.end(), d->second.begin(), d->second.end());
                fw.p2l_sqrt_inf.insert(fw.p2l_sqrt_inf.end(), w->second.begin(), w->second.end());
            }
            sp.n_p2l = (int32_t)fw.p2l_lmk.size();
            have_sparse = true;
        } else { // sparsified, VO case: unary factor + chain, :427-481
            if (!marg->lmk_with_prior) return false;
            sdv_sparse_prior &sp = fw.sparse;
            sp.has_lmk_prior = 1;
            sp.lmk0 = block_of(marg->lmk_with_prior);
            std::memcpy(sp.lmk_prior, marg->prior_lmk.data(), 24);
            std::memcpy(sp.lmk_sqrt_inf, marg->info_lmk.data(), 72);
            for (size_t k = 0; k + 1 < marg->lmk_to_keep.size(); k++) {
                const std::shared_ptr<Landmark> &lmk_k = marg->lmk_to_keep[k], &lmk_kp1 = marg->lmk_to_keep[k + 1];
                const int a = block_of(lmk_k), b = block_of(lmk_kp1);
                if (lmk_k == lmk_kp1) continue; // :463-464
                auto d = marg->map_lmk_prior.find(lmk_kp1.get());
                auto w = marg->map_lmk_inf.find(lmk_kp1.get());
                if (d == marg->map_lmk_prior.end() || w == marg->map_lmk_inf.end()) return false;
                fw.l2l_a.push_back(a);
                fw.l2l_b.push_back(b);
                fw.l2l_delta.insert(fw.l2l_delta.end(), d->second.begin(), d->second.end());
                fw.l2l_sqrt_inf.insert(fw.l2l_sqrt_inf.end(), w->second.begin(), w->second.end());
            }
            sp.n_l2l = (int32_t)fw.l2l_a.size();
            have_sparse = true;
        }
        if (have_sparse) {
            sdv_sparse_prior &sp = fw.sparse;
            sp.p2l_lmk = fw.p2l_lmk.data(); sp.p2l_delta = fw.p2l_delta.data(); sp.p2l_sqrt_inf = fw.p2l_sqrt_inf.data();
            sp.l2l_a = fw.l2l_a.data(); sp.l2l_b = fw.l2l_b.data(); sp.l2l_delta = fw.l2l_delta.data(); sp.l2l_sqrt_inf = fw.l2l_sqrt_inf.data();
        }
    }
    detail::fill_view(fw, vio, factor_kind, fixed_frame_number, have_dense, have_sparse);
    return true;
}

namespace detail {
inline void exp_so3(const double *v, double *R) { // geometry.h:131-147
    double a = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double K[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
    if (a < 1e-9) {
        for (int i = 0; i < 9; i++) R[i] = K[i] + (i % 4 == 0 ? 1.0 : 0.0);
        return;
    }
    for (int i = 0; i < 9; i++) K[i] /= a;
    double K2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) K2[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + (1 - std::cos(a)) * K2[i] + std::sin(a) * K[i];
}
inline void mul33(const double *A, const double *B, double *C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
} // namespace detail

// State write-back of AOptimizer.cpp:391-434 (pose on the right, additive landmark / velocity / biases, then
// IMU::biasDeltaCorrection of every frame with its PREVIOUS keyframe's dba/dbg, IMU.cpp:104-108).
inline void write_back(FlatWindow &fw, const sdv_delta &d, bool vio) {
    const int F = (int)fw.frame_vector.size();
    for (int f = 0; f < F; f++) {
        Frame &fr = *fw.frame_vector[f];
        double dR[9], R[9], Rn[9];
        detail::exp_so3(d.dpose + 6 * f, dR);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) R[i * 3 + j] = fr.T_f_w[i * 4 + j];
        detail::mul33(R, dR, Rn);
        for (int i = 0; i < 3; i++) {
            double t = fr.T_f_w[i * 4 + 3];
            for (int j = 0; j < 3; j++) t += R[i * 3 + j] * d.dpose[6 * f + 3 + j];
            for (int j = 0; j < 3; j++) fr.T_f_w[i * 4 + j] = Rn[i * 3 + j];
            fr.T_f_w[i * 4 + 3] = t;
        }
        if (vio && fr.imu)
            for (int k = 0; k < 3; k++) {
                fr.imu->v[k] += d.dv[3 * f + k];
                fr.imu->ba[k] += d.dba[3 * f + k];
                fr.imu->bg[k] += d.dbg[3 * f + k];
            }
    }
    for (size_t l = 0; l < fw.landmarks.size(); l++)
        for (int k = 0; k < 3; k++) fw.landmarks[l]->t_w[k] += d.dlmk[3 * l + k];
    if (!vio) return;
    for (size_t p = 0; p < fw.corr_frame.size(); p++) {
        IMU &m = *fw.corr_frame[p]->imu;
        const double *dba = d.dba + 3 * fw.corr_prev[p], *dbg = d.dbg + 3 * fw.corr_prev[p];
        double phi[3], E[9], Rn[9];
        for (int i = 0; i < 3; i++) {
            double sp = 0, sv = 0;
            phi[i] = 0;
            for (int j = 0; j < 3; j++) {
                sp += m.J_dp_ba[i * 3 + j] * dba[j] + m.J_dp_bg[i * 3 + j] * dbg[j];
                sv += m.J_dv_ba[i * 3 + j] * dba[j] + m.J_dv_bg[i * 3 + j] * dbg[j];
                phi[i] += m.J_dR_bg[i * 3 + j] * dbg[j];
            }
            m.delta_p[i] += sp;
            m.delta_v[i] += sv;
        }
        detail::exp_so3(phi, E);
        detail::mul33(m.delta_R.data(), E, Rn);
        for (int i = 0; i < 9; i++) m.delta_R[i] = Rn[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The other three visual solves of AOptimizer as masks of the window solve (include/sdv.h, the
// table under sdv_window).
// ---------------------------------------------------------------------------------------------------------------------

// ALandmark::sanityCheck (ALandmark.cpp:130-147): fewer than two features -> outlier; average over the features of the mean
// squared pixel error / sigma^2 (chi2err, :98-116; 1000 when Camera::project fails, Camera.cpp:27-52) above 2 -> outlier.
inline bool Landmark::sanityCheck() {
    if (features.size() < 2) {
        outlier = true;
        return false;
    }
    double mean = 0.0;
    for (auto &wf : features) {
        double chi2 = 1000.0;
        std::shared_ptr<Feature> f = wf.lock();
        std::shared_ptr<ImageSensor> cam = f ? f->sensor.lock() : nullptr;
        std::shared_ptr<Frame> fr = cam ? cam->getFrame() : nullptr;
        if (fr) {
            double pf[3], pc[3];
            for (int i = 0; i < 3; i++) pf[i] = fr->T_f_w[4 * i] * t_w[0] + fr->T_f_w[4 * i + 1] * t_w[1] + fr->T_f_w[4 * i + 2] * t_w[2] + fr->T_f_w[4 * i + 3];
            for (int i = 0; i < 3; i++) pc[i] = cam->T_s_f[4 * i] * pf[0] + cam->T_s_f[4 * i + 1] * pf[1] + cam->T_s_f[4 * i + 2] * pf[2] + cam->T_s_f[4 * i + 3];
            const double u = (cam->K[0] * pc[0] + cam->K[2] * pc[2]) / pc[2], v = (cam->K[1] * pc[1] + cam->K[3] * pc[2]) / pc[2];
            const bool ok = !(pc[2] < 0.1) && !(u < 0 || v < 0 || u > 2 * cam->K[2] || v > 2 * cam->K[3]) && std::isfinite(u) && std::isfinite(v);
            if (ok) {
                const double e0 = (u - f->uv[0]) / f->sigma, e1 = (v - f->uv[1]) / f->sigma;
                chi2 = e0 * e0 + e1 * e1;
            }
        }
        mean += chi2;
    }
    outlier = mean / (double)features.size() > 2.0;
    return !outlier;
}

// landmarkOptimization(frame): addLandmarkResiduals (AngularAdjustmentCERESAnalytic.cpp:106-209) — a parameter block per
// initialised inlier landmark of frame->getLandmarks(), one visual block per live feature on ANY keyframe, every pose block
// constant (:141-145), sigma 1.5 / focal (:152).  Frames enter frame_vector in first-appearance order.
inline void flatten_landmark_cloud(const Frame &frame, int factor_kind, FlatWindow &fw) {
    fw = FlatWindow();
    std::unordered_map<const Frame *, int> frame_idx;
    for (auto &landmark : frame.pointxd) {
        if (!landmark->isInitialized() || landmark->isOutlier()) continue; // :117
        const int l = (int)fw.landmarks.size();
        fw.landmarks.push_back(landmark);
        fw.lmk_t.insert(fw.lmk_t.end(), landmark->t_w.begin(), landmark->t_w.end());
        for (auto &wfeature : landmark->features) {
            std::shared_ptr<Feature> feature = wfeature.lock();
            if (!feature) continue; // (the reference dereferences before it tests, :129-134: an expired feature is undefined behaviour there)
            std::shared_ptr<ImageSensor> cam = feature->sensor.lock();
            std::shared_ptr<Frame> fr = cam ? cam->getFrame() : nullptr;
            if (!fr || !fr->isKeyFrame()) continue; // :134
            auto it = frame_idx.find(fr.get());
            if (it == frame_idx.end()) { // :138-142
                it = frame_idx.emplace(fr.get(), (int)fw.frame_vector.size()).first;
                fw.frame_vector.push_back(fr);
                detail::push_frame_state(fw, *fr, false);
            }
            fw.obs_lmk.push_back(l);
            fw.obs_frame.push_back(it->second);
            fw.obs_cam.push_back(detail::cam_of(fw, *cam));
            fw.obs_bearing.insert(fw.obs_bearing.end(), feature->bearing.begin(), feature->bearing.end());
            fw.obs_uv.insert(fw.obs_uv.end(), feature->uv.begin(), feature->uv.end());
        }
    }
    detail::fill_view(fw, false, factor_kind, fw.frame_vector.size(), false, false); // n_fixed = n_frames: every pose constant
    fw.view.visual_loss_huber_a = std::sqrt(1.345); // AOptimizer.cpp:102
    fw.view.max_num_iterations = 10;                // :113
}

// singleFrameOptimization / singleFrameVIOptimization: addSingleFrameResiduals (…Analytic.cpp:6-102) for the moving frame
// and — VI, when both frames carry an IMU (AOptimizer.cpp:231-240) — for its previous keyframe over the SAME cloud (the moving
// frame's landmarks): a free pose block per frame, every landmark block constant, sigma 1 / focal (:48).  The reference adds
// the blocks frame by frame; the C ABI wants observations landmark-major, so each landmark lists its features on the moving
// frame, then those on the previous keyframe — the same residual blocks, summed in another order.
inline void flatten_single_frame(const std::shared_ptr<Frame> &moving, bool vi, int factor_kind, FlatWindow &fw) {
    fw = FlatWindow();
    fw.frame_vector.push_back(moving);
    const bool with_imu = vi && moving->imu && moving->imu->last_kf && moving->imu->last_kf->imu; // AOptimizer.cpp:231
    if (with_imu) fw.frame_vector.push_back(moving->imu->last_kf);
    std::unordered_map<const Frame *, int> frame_idx;
    for (size_t i = 0; i < fw.frame_vector.size(); i++) {
        frame_idx[fw.frame_vector[i].get()] = (int)i;
        detail::push_frame_state(fw, *fw.frame_vector[i], false);
    }
    for (auto &landmark : moving->pointxd) {
        if (!landmark->isInitialized() || landmark->isOutlier()) continue; // :20
        int l = -1;
        for (size_t i = 0; i < fw.frame_vector.size(); i++)
            for (auto &wfeature : landmark->features) {
                std::shared_ptr<Feature> feature = wfeature.lock();
                if (!feature) continue; // :29
                std::shared_ptr<ImageSensor> cam = feature->sensor.lock();
                if (!cam || cam->getFrame() != fw.frame_vector[i]) continue; // :36
                if (l < 0) { // :38-42
                    l = (int)fw.landmarks.size();
                    fw.landmarks.push_back(landmark);
                    fw.lmk_t.insert(fw.lmk_t.end(), landmark->t_w.begin(), landmark->t_w.end());
                }
                fw.obs_lmk.push_back(l);
                fw.obs_frame.push_back((int)i);
                fw.obs_cam.push_back(detail::cam_of(fw, *cam));
                fw.obs_bearing.insert(fw.obs_bearing.end(), feature->bearing.begin(), feature->bearing.end());
                fw.obs_uv.insert(fw.obs_uv.end(), feature->uv.begin(), feature->uv.end());
                if (factor_kind == SDV_FACTOR_ANGULAR) fw.obs_sigma.push_back(1.0 / cam->getFocal()); // :48 (the pixel functor has sigma 1)
            }
    }
    if (with_imu) detail::imu_factors(fw, frame_idx); // addIMUResiduals(problem, nullptr, ordering, frame_vec, 0), AOptimizer.cpp:239
    detail::fill_view(fw, with_imu, factor_kind, 0, false, false);
    fw.view.landmarks_constant = 1;
    fw.view.visual_loss_huber_a = vi ? std::sqrt(1.345) : 0.0; // AOptimizer.cpp:223 / :156
    fw.view.max_num_iterations = 5;                            // :166, :247
}

// marginalize(frame0, frame1, enable_sparsif) (AngularAdjustmentCERESAnalytic.cpp:488-739): the window sdv_marginalize wants —
// every landmark of frame 0 that preMarginalize looks at (marginalization.cpp:51-56) with ALL its live features, whatever frame
// they sit on (a feature outside frame 0 makes the landmark "not lonely", :60-68), then the landmarks of the previous prior
// that frame 0 no longer sees ("resurrected", :118-139), frames in first-appearance order with frame 1 and frame 0 LAST.
// Returns false where the reference would throw (previous prior on another frame).  `discard_last`: an outlier among the
// resurrected landmarks makes the reference drop the previous prior altogether (:124-127, :141-142).
inline bool flatten_marginalization(const std::shared_ptr<Frame> &frame0, const std::shared_ptr<Frame> &frame1, int factor_kind, const Marginalization *last,
                                    FlatWindow &fw, std::vector<uint8_t> &lmk_has_prior, bool &discard_last) {
    fw = FlatWindow();
    discard_last = false;
    lmk_has_prior.clear();
    const bool vio = frame0->imu && frame1->imu;
    std::vector<std::shared_ptr<Frame>> others;
    std::unordered_map<const Frame *, int> tmp_idx;
    std::unordered_map<const Landmark *, int> lmk_idx;
    struct Obs { int l; const Frame *f; std::shared_ptr<ImageSensor> cam; std::shared_ptr<Feature> ft; };
    std::vector<Obs> obs;
    for (auto &lmk : frame0->pointxd) {
        if (lmk->isOutlier() || !lmk->in_map || !lmk->isInitialized()) continue; // marginalization.cpp:53
        const int l = (int)fw.landmarks.size();
        lmk_idx[lmk.get()] = l;
        fw.landmarks.push_back(lmk);
        fw.lmk_t.insert(fw.lmk_t.end(), lmk->t_w.begin(), lmk->t_w.end());
        lmk_has_prior.push_back(lmk->has_prior ? 1 : 0);
        for (auto &wf : lmk->features) {
            std::shared_ptr<Feature> ft = wf.lock();
            std::shared_ptr<ImageSensor> cam = ft ? ft->sensor.lock() : nullptr;
            std::shared_ptr<Frame> fr = cam ? cam->getFrame() : nullptr;
            if (!fr) continue; // (the reference dereferences unconditionally, :60)
            if (fr != frame0 && fr != frame1 && tmp_idx.emplace(fr.get(), (int)others.size()).second) others.push_back(fr);
            obs.push_back({l, fr.get(), cam, ft});
        }
    }
    const bool have_last = last && !last->lmk_to_keep.empty();
    if (have_last) {
        for (auto &lmk : last->lmk_to_keep)
            if (lmk_idx.find(lmk.get()) == lmk_idx.end()) {
                if (lmk->isOutlier()) { // :124-127
                    discard_last = true;
                    break;
                }
                lmk_idx[lmk.get()] = (int)fw.landmarks.size();
                fw.landmarks.push_back(lmk);
                fw.lmk_t.insert(fw.lmk_t.end(), lmk->t_w.begin(), lmk->t_w.end());
                lmk_has_prior.push_back(lmk->has_prior ? 1 : 0);
            }
    }
    fw.frame_vector = others;
    fw.frame_vector.push_back(frame1);
    fw.frame_vector.push_back(frame0);
    std::unordered_map<const Frame *, int> frame_idx;
    const int F = (int)fw.frame_vector.size();
    for (int i = 0; i < F; i++) {
        frame_idx[fw.frame_vector[i].get()] = i;
        detail::push_frame_state(fw, *fw.frame_vector[i], i >= F - 2); // only the pose priors of frame 0 / frame 1 matter (:664-687)
    }
    for (auto &o : obs) {
        fw.obs_lmk.push_back(o.l);
        fw.obs_frame.push_back(frame_idx[o.f]);
        fw.obs_cam.push_back(detail::cam_of(fw, *o.cam));
        fw.obs_bearing.insert(fw.obs_bearing.end(), o.ft->bearing.begin(), o.ft->bearing.end());
        fw.obs_uv.insert(fw.obs_uv.end(), o.ft->uv.begin(), o.ft->uv.end());
    }
    if (vio) { // IMUFactor(frame0->getIMU(), frame1->getIMU()) (…Analytic.cpp:539-547): frame 1's pre-integration, whatever the distance
        const IMU &m = *frame1->imu;
        fw.imu_i.push_back(F - 1);
        fw.imu_j.push_back(F - 2);
        fw.imu_dt.push_back((double)(frame1->timestamp_ns - frame0->timestamp_ns) * 1e-9);
        fw.imu_dR.insert(fw.imu_dR.end(), m.delta_R.begin(), m.delta_R.end());
        fw.imu_dv.insert(fw.imu_dv.end(), m.delta_v.begin(), m.delta_v.end());
        fw.imu_dp.insert(fw.imu_dp.end(), m.delta_p.begin(), m.delta_p.end());
        fw.imu_cov.insert(fw.imu_cov.end(), m.Sigma.begin(), m.Sigma.end());
        fw.J_dR_bg.insert(fw.J_dR_bg.end(), m.J_dR_bg.begin(), m.J_dR_bg.end());
        fw.J_dv_ba.insert(fw.J_dv_ba.end(), m.J_dv_ba.begin(), m.J_dv_ba.end());
        fw.J_dv_bg.insert(fw.J_dv_bg.end(), m.J_dv_bg.begin(), m.J_dv_bg.end());
        fw.J_dp_ba.insert(fw.J_dp_ba.end(), m.J_dp_ba.begin(), m.J_dp_ba.end());
        fw.J_dp_bg.insert(fw.J_dp_bg.end(), m.J_dp_bg.begin(), m.J_dp_bg.end());
        fw.sigma_ba.push_back(frame0->imu->bacc_noise);
        fw.sigma_bg.push_back(frame0->imu->bgyr_noise);
    }
    bool have_dense = false;
    if (have_last && !discard_last) { // the previous prior, always in its dense form (…Analytic.cpp:631-660)
        sdv_dense_prior &dp = fw.dense;
        dp.n_full = last->n_full;
        dp.n = last->n;
        dp.J = last->marginalization_jacobian.data();
        dp.r0 = last->marginalization_residual.data();
        dp.frame = -1;
        dp.frame_col = 0;
        if (last->frame_to_keep) {
            if (last->frame_to_keep != frame0) return false; // _map_frame_posepar.at(_frame_to_keep) throws
            auto it = last->map_frame_idx.find(frame0.get());
            if (it == last->map_frame_idx.end()) return false;
            dp.frame = F - 1;
            dp.frame_col = it->second;
        }
        for (auto &lmk : last->lmk_to_keep) {
            auto it = last->map_lmk_idx.find(lmk.get());
            if (it == last->map_lmk_idx.end()) return false;
            fw.keep_lmk.push_back(lmk_idx.at(lmk.get()));
            fw.keep_col.push_back(it->second);
        }
        dp.n_keep = (int32_t)fw.keep_lmk.size();
        dp.keep_lmk = fw.keep_lmk.data();
        dp.keep_col = fw.keep_col.data();
        have_dense = true;
    }
    detail::fill_view(fw, vio, factor_kind, 0, have_dense, false);
    fw.view.lmk_has_prior = lmk_has_prior.data();
    return true;
}

// VIInit (AOptimizer.cpp:448-529): every frame of the local map newest -> oldest (:456-457), one velocity block per frame with an
// IMU (:459-464), one IMUFactorInit per frame whose getLastKF() is another frame of the map with an IMU (:485-502, no dt test).
// No landmark, no observation crosses the ABI: the solve is inertial only.  False when a frame has no IMU (the reference
// dereferences getIMU() unconditionally at :487).
inline bool flatten_viinit(const LocalMap &map, FlatWindow &fw) {
    fw = FlatWindow();
    map.getLastNFramesIn(map.getMapSize(), fw.frame_vector);
    const int F = (int)fw.frame_vector.size();
    std::unordered_map<const Frame *, int> frame_idx;
    for (int i = 0; i < F; i++) {
        if (!fw.frame_vector[i]->imu) return false;
        frame_idx[fw.frame_vector[i].get()] = i;
        detail::push_frame_state(fw, *fw.frame_vector[i], true);
    }
    detail::imu_factors(fw, frame_idx, false);
    detail::fill_view(fw, true, SDV_FACTOR_ANGULAR, 0, false, false);
    return F > 0;
}

// State update of VIInit (AOptimizer.cpp:531-567).  The reference adds dba — a constant block, zero — to the accelerometer bias
// twice and never touches the gyroscope bias (:535-536): nothing to do for the biases.
inline void viinit_write_back(LocalMap &map, FlatWindow &fw, const sdv_viinit_result &r) {
    const int F = (int)fw.frame_vector.size();
    const double s = std::exp(r.lambda), *Rw = r.R_w_i;
    for (int f = 0; f < F; f++) {
        Frame &fr = *fw.frame_vector[f];
        for (int k = 0; k < 3; k++) fr.imu->v[k] += r.dv[3 * f + k];                     // :533-534
        double R[9], Rn[9];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) R[i * 3 + j] = fr.T_f_w[i * 4 + j];
        detail::mul33(R, Rw, Rn);                                                            // :550, T_w_i = [R_w_i | 0]
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) fr.T_f_w[i * 4 + j] = Rn[i * 3 + j];
            fr.T_f_w[i * 4 + 3] *= s;                                                   // :549
        }
        if (fr.has_prior) {                                                             // :553-555
            fr.T_prior = fr.T_f_w;
            fr.inf_prior.fill(100.0);
        }
    }
    for (auto &lm : map.pointxd) {                                                      // :559-567
        if (lm->isOutlier()) continue;
        double t[3];
        for (int i = 0; i < 3; i++) t[i] = Rw[0 * 3 + i] * lm->t_w[0] + Rw[1 * 3 + i] * lm->t_w[1] + Rw[2 * 3 + i] * lm->t_w[2]; // R_w_i^T t
        for (int i = 0; i < 3; i++) lm->t_w[i] = s * t[i];
    }
}

// Drop-in for the reference optimizer object (one instance = one sdv_handle, as the back-end optimizer instance,
// slamParameters.cpp:273-274).  `factor_kind` picks what the reference picks by class: AngularAdjustmentCERESAnalytic
// (SDV_FACTOR_ANGULAR) or BundleAdjustmentCERESAnalytic (SDV_FACTOR_PIXEL).
class B200Optimizer {
  public:
    explicit B200Optimizer(int factor_kind = SDV_FACTOR_ANGULAR, int device = 0) : _kind(factor_kind) {
        sdv_config cfg;
        sdv_default_config(&cfg);
        cfg.device = device;
        if (sdv_create(&_h, &cfg) != SDV_OK) _h = nullptr; // no GPU -> every solve returns false (there is no CPU path)
    }
    ~B200Optimizer() {
        if (_h) sdv_destroy(_h);
    }
    B200Optimizer(const B200Optimizer &) = delete;
    B200Optimizer &operator=(const B200Optimizer &) = delete;

    bool localMapBA(std::shared_ptr<LocalMap> &local_map, const size_t fixed_frame_number = 0) { return solve(*local_map, fixed_frame_number, false); }
    bool localMapVIOptimization(std::shared_ptr<LocalMap> &local_map, const size_t fixed_frame_number = 0) {
        return solve(*local_map, fixed_frame_number, true);
    }
    // Landmark refinement with every pose constant and a Huber loss (AOptimizer.cpp:98-150).  Only landmarks that pass
    // ALandmark::sanityCheck are updated (:132-140); always true, like the reference (false only when no solve could run).
    bool landmarkOptimization(std::shared_ptr<Frame> &frame) {
        if (!_h) return false;
        FlatWindow fw;
        flatten_landmark_cloud(*frame, _kind, fw);
        std::vector<double> buf;
        sdv_delta d;
        if (fw.view.n_obs > 0) { // (an empty ceres::Problem solves to itself)
            const int rc = run(fw, buf, d);
            if (rc != SDV_OK && rc != SDV_ERR_NUMERICAL_FAILURE) return false;
        }
        size_t l = 0;
        for (auto &ldmk : frame->pointxd) {
            if (!ldmk->isInitialized() || ldmk->isOutlier()) continue; // same check as in addLandmarkResiduals, :128
            const bool can_be_updated = ldmk->sanityCheck();           // :132 (marks the landmark inlier / outlier)
            if (can_be_updated && fw.view.n_obs > 0)
                for (int k = 0; k < 3; k++) ldmk->t_w[k] += d.dlmk[3 * l + k]; // :136
            l++;
        }
        return true;
    }
    // Pose-only refinement of one frame against constant landmarks (AOptimizer.cpp:152-217): always writes back, returns true.
    bool singleFrameOptimization(std::shared_ptr<Frame> &moving_frame) { return single_frame(moving_frame, false); }
    // The same with the IMU factor to the previous keyframe and a Huber loss on the visual blocks (AOptimizer.cpp:219-297):
    // false without write-back when the summary is not usable (Ceres FAILURE, :259), else poses, velocities and biases.
    bool singleFrameVIOptimization(std::shared_ptr<Frame> &moving_frame) { return single_frame(moving_frame, true); }
    // Visual-inertial initialisation (AOptimizer.cpp:448-581): gravity direction R_w_i (row-major 3x3 out), keyframe velocities and —
    // with optim_scale — the metric scale; velocities, poses, priors and landmarks of the map are updated as the reference does
    // (:531-567) and exp(lambda) is returned.  NaN with the state untouched when no solve could run (no GPU, a frame without IMU).
    double VIInit(std::shared_ptr<LocalMap> &local_map, double *R_w_i, bool optim_scale) {
        FlatWindow fw;
        if (!_h || !flatten_viinit(*local_map, fw)) return std::nan("");
        std::vector<double> dv(3 * fw.frame_vector.size(), 0.0);
        sdv_viinit_result res;
        std::memset(&res, 0, sizeof(res));
        res.dv = dv.data();
        const int rc = sdv_viinit(_h, &fw.view, optim_scale ? 1 : 0, &res, &_stats);
        if (rc != SDV_OK && rc != SDV_ERR_NUMERICAL_FAILURE) return std::nan(""); // (the reference ignores the summary: FAILURE still writes back)
        viinit_write_back(*local_map, fw, res);
        if (R_w_i) std::memcpy(R_w_i, res.R_w_i, sizeof(res.R_w_i));
        return res.scale;
    }
    // Marginal prior of the keyframe that leaves the window (AngularAdjustmentCERESAnalytic.cpp:488-739 /
    // BundleAdjustmentCERESAnalytic.cpp:431-660) on the GPU; fills _marginalization (what the next window solve wires in) and
    // _marginalization_last (what the next marginalisation folds in) like the reference does (:714-736).
    bool marginalize(std::shared_ptr<Frame> &frame0, std::shared_ptr<Frame> &frame1, bool enable_sparsif) {
        _enable_sparsif = enable_sparsif; // :491
        if (!_h || !frame0 || !frame1) return false;
        if ((frame0->imu != nullptr) != (frame1->imu != nullptr)) return false; // (one IMU only: not a configuration the back ends produce)
        FlatWindow fw;
        std::vector<uint8_t> has_prior;
        bool discard_last = false;
        if (!flatten_marginalization(frame0, frame1, _kind, _marginalization_last.get(), fw, has_prior, discard_last)) return false;
        if (discard_last) _marginalization_last->lmk_to_keep.clear(); // marginalization.cpp:141-142
        sdv_marginal_sizes sz;
        if (sdv_marginalize(_h, &fw.view, enable_sparsif ? 1 : 0, &sz) != SDV_OK) return false;
        _marg_sizes = sz;
        Marginalization &M = *_marginalization;
        M = Marginalization();
        if (!sz.ok) { // computeSchurComplement() == false: the scheme is reset (…Analytic.cpp:690-695)
            _marginalization_last->lmk_to_keep.clear();
            return false;
        }
        std::vector<int32_t> keep(std::max(sz.n_keep, 1)), marg(std::max(sz.n_marg, 1)), chain(std::max(sz.n_chain, 1));
        std::vector<double> p2d(3 * (size_t)std::max(sz.n_keep, 1)), p2s(9 * (size_t)std::max(sz.n_keep, 1)), l2d(3 * (size_t)std::max(sz.n_chain, 1)),
            l2s(9 * (size_t)std::max(sz.n_chain, 1));
        std::array<double, 225> imu_inf{};
        M.n = sz.n;
        M.n_full = sz.n_full;
        M.marginalization_jacobian.assign((size_t)sz.n_full * sz.n, 0.0);
        M.marginalization_residual.assign(sz.n_full, 0.0);
        sdv_marginal out;
        std::memset(&out, 0, sizeof(out));
        out.J = M.marginalization_jacobian.data();
        out.r0 = M.marginalization_residual.data();
        out.keep_lmk = keep.data();
        out.marg_lmk = marg.data();
        out.imu_sqrt_inf = imu_inf.data();
        out.p2l_delta = p2d.data();
        out.p2l_sqrt_inf = p2s.data();
        out.chain = chain.data();
        out.l2l_delta = l2d.data();
        out.l2l_sqrt_inf = l2s.data();
        if (sdv_marginal_fetch(_h, &out) != SDV_OK) return false;
        // flags preMarginalize leaves on the landmarks (marginalization.cpp:72-86)
        for (int k = 0; k < sz.n_marg; k++) fw.landmarks[marg[k]]->is_marg = true;
        for (auto &lmk : frame0->pointxd) {
            if (lmk->isOutlier() || !lmk->in_map || !lmk->isInitialized()) continue;
            int num_cam = 0;
            for (auto &wf : lmk->features) {
                std::shared_ptr<Feature> ft = wf.lock();
                std::shared_ptr<ImageSensor> cam = ft ? ft->sensor.lock() : nullptr;
                if (cam && cam->getFrame() == frame0) num_cam++;
            }
            if (num_cam != 2 && !lmk->has_prior) lmk->is_marg = true; // :72-75
        }
        const int first = sz.frame >= 0 ? 15 : 0;
        if (sz.frame >= 0) {
            M.frame_to_keep = frame1;
            M.map_frame_idx[frame1.get()] = 0; // after the shift by _m (marginalization.cpp:251-256)
        }
        for (int k = 0; k < sz.n_keep; k++) {
            const std::shared_ptr<Landmark> &lmk = fw.landmarks[keep[k]];
            lmk->has_prior = true; // setPrior(), :79 (resurrected landmarks carry it already)
            M.lmk_to_keep.push_back(lmk);
            M.map_lmk_idx[lmk.get()] = first + 3 * k;
        }
        if (enable_sparsif && sz.n_full > 0) { // :703-708
            if (sz.frame >= 0) { // sparsifyVIO
                M.map_frame_inf[frame1.get()] = imu_inf;
                for (int k = 0; k < sz.n_keep; k++) {
                    const Landmark *l = fw.landmarks[keep[k]].get();
                    std::memcpy(M.map_lmk_prior[l].data(), &p2d[3 * (size_t)k], 24);
                    std::memcpy(M.map_lmk_inf[l].data(), &p2s[9 * (size_t)k], 72);
                }
            } else if (sz.n_chain >= 2) { // sparsifyVO: _lmk_to_keep becomes the chain (marginalization.cpp:461-462)
                M.lmk_to_keep.clear();
                for (int k = 0; k < sz.n_chain; k++) M.lmk_to_keep.push_back(fw.landmarks[chain[k]]);
                M.lmk_with_prior = fw.landmarks[out.lmk_with_prior];
                M.prior_lmk = M.lmk_with_prior->t_w;
                std::memcpy(M.info_lmk.data(), out.lmk_sqrt_inf, 72);
                for (int k = 0; k + 1 < sz.n_chain; k++) { // keyed on lmk_kp1 (:505-506)
                    const Landmark *l = fw.landmarks[chain[k + 1]].get();
                    std::memcpy(M.map_lmk_prior[l].data(), &l2d[3 * (size_t)k], 24);
                    std::memcpy(M.map_lmk_inf[l].data(), &l2s[9 * (size_t)k], 72);
                }
            }
        }
        *_marginalization_last = M; // …Analytic.cpp:714-736
        return true;
    }
    const sdv_marginal_sizes &lastMarginalSizes() const { return _marg_sizes; }
    std::shared_ptr<Marginalization> _marginalization_last = std::make_shared<Marginalization>(); // AOptimizer::_marginalization_last
    const sdv_stats &lastStats() const { return _stats; }
    // AOptimizer::_marginalization / _enable_sparsif (AOptimizer.h:88-89): what marginalize() left for the next window solve
    std::shared_ptr<Marginalization> _marginalization = std::make_shared<Marginalization>();
    bool _enable_sparsif = false;

  private:
    int run(FlatWindow &fw, std::vector<double> &buf, sdv_delta &d) {
        const size_t F = fw.frame_vector.size(), L = fw.landmarks.size();
        buf.assign(15 * F + 3 * L + 1, 0.0);
        d.dpose = buf.data();
        d.dv = d.dpose + 6 * F;
        d.dba = d.dv + 3 * F;
        d.dbg = d.dba + 3 * F;
        d.dlmk = d.dbg + 3 * F;
        return sdv_solve_window(_h, &fw.view, &d, &_stats);
    }
    bool single_frame(std::shared_ptr<Frame> &moving_frame, bool vi) {
        if (!_h) return false;
        FlatWindow fw;
        flatten_single_frame(moving_frame, vi, _kind, fw);
        if (fw.view.n_obs == 0 && fw.view.n_imu == 0) return true; // nothing to optimise: Ceres returns at once, the state stays
        std::vector<double> buf;
        sdv_delta d;
        const int rc = run(fw, buf, d);
        if (vi && rc == SDV_ERR_NUMERICAL_FAILURE) return false; // !summary.IsSolutionUsable(), AOptimizer.cpp:259
        if (rc != SDV_OK && rc != SDV_ERR_NUMERICAL_FAILURE) return false;
        // poses of _map_frame_posepar, and — VI with both IMUs — v / ba / bg (:263-285); no biasDeltaCorrection here
        FlatWindow only_state = std::move(fw);
        only_state.landmarks.clear();
        only_state.corr_frame.clear();
        write_back(only_state, d, only_state.view.vio != 0);
        return true;
    }
    bool solve(LocalMap &map, size_t fixed, bool vio) {
        if (!_h) return false;
        FlatWindow fw;
        if (!flatten(map, fixed, vio, _kind, fw, _marginalization.get(), _enable_sparsif)) return false;
        std::vector<double> buf;
        sdv_delta d;
        int rc = run(fw, buf, d);
        // SDV_ERR_NUMERICAL_FAILURE is Ceres' TerminationType FAILURE: the reference ignores the summary, writes back the
        // last accepted x and returns true (AOptimizer.cpp:388-445); the termination is in lastStats().  Every other
        // non-zero status means no solve happened: state untouched, false.
        if (rc != SDV_OK && rc != SDV_ERR_NUMERICAL_FAILURE) return false;
        write_back(fw, d, vio);
        return true;
    }
    sdv_handle *_h = nullptr;
    int _kind;
    sdv_stats _stats{};
    sdv_marginal_sizes _marg_sizes{};
};

} // namespace sdvhost
