// Test driver for sadvio_b200/host/b200_optimizer.hpp: reads a pointer-graph description (text, produced by
// tests/test_host_adapter.py), rebuilds the graph with shared_ptr / weak_ptr objects exactly as SaDVIO holds it,
// and either prints the flattened window ("flatten": indices only, "dump": every array of the sdv_window) or runs
// localMapVIOptimization / localMapBA through the C ABI and prints the updated state ("solve"), or applies a synthetic
// solution with the write-back alone ("writeback", no GPU).  "dump_lmkopt" / "dump_single" / "dump_singlevi" print the window
// of landmarkOptimization / singleFrameOptimization / singleFrameVIOptimization(frame), "lmkopt" / "single" / "singlevi" run
// them (frame = argv[5], a position in the file).  An optional trailing
// section describes the isae::Marginalization object the optimizer holds (dense or sparsified prior).
#include "b200_optimizer.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

using namespace sdvhost;

template <size_t N> static void rd(std::istream &in, std::array<double, N> &a) {
    for (auto &x : a) in >> x;
}

int main(int argc, char **argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: adapter_check <flatten|solve> <vio 0/1> <fixed> [kind]\n");
        return 2;
    }
    std::string mode = argv[1];
    bool vio = std::atoi(argv[2]) != 0;
    size_t fixed = (size_t)std::atoi(argv[3]);
    int kind = argc > 4 ? std::atoi(argv[4]) : 0;
    int f2_frame = argc > 5 ? std::atoi(argv[5]) : 0;
    std::istream &in = std::cin;
    int NF;
    in >> NF;
    std::vector<std::shared_ptr<Frame>> all_frames(NF);
    std::vector<int> in_window(NF), lastkf(NF, -1);
    auto map = std::make_shared<LocalMap>();
    for (int f = 0; f < NF; f++) {
        auto fr = std::make_shared<Frame>();
        int kf, hp, hi, ns;
        in >> fr->timestamp_ns >> kf >> in_window[f] >> hp >> hi >> ns;
        fr->keyframe = kf;
        fr->has_prior = hp;
        rd(in, fr->T_f_w);
        rd(in, fr->T_prior);
        rd(in, fr->inf_prior);
        if (hi) {
            fr->imu = std::make_shared<IMU>();
            IMU &m = *fr->imu;
            rd(in, m.v); rd(in, m.ba); rd(in, m.bg); rd(in, m.delta_R); rd(in, m.delta_v); rd(in, m.delta_p); rd(in, m.Sigma);
            rd(in, m.J_dR_bg); rd(in, m.J_dv_ba); rd(in, m.J_dv_bg); rd(in, m.J_dp_ba); rd(in, m.J_dp_bg);
            in >> m.bacc_noise >> m.bgyr_noise >> lastkf[f];
        }
        for (int s = 0; s < ns; s++) {
            auto cam = std::make_shared<ImageSensor>();
            cam->frame = fr;
            rd(in, cam->T_s_f);
            rd(in, cam->K);
            fr->sensors.push_back(cam);
        }
        all_frames[f] = fr;
        if (in_window[f]) map->frames.push_back(fr); // listed oldest -> newest
    }
    for (int f = 0; f < NF; f++)
        if (all_frames[f]->imu && lastkf[f] >= 0) all_frames[f]->imu->last_kf = all_frames[lastkf[f]];
    int NL;
    in >> NL;
    std::vector<std::shared_ptr<Feature>> keep_alive; // features are owned by their frames in SaDVIO; landmarks hold weak_ptrs
    for (int l = 0; l < NL; l++) {
        auto lm = std::make_shared<Landmark>();
        int init, outl, nfeat;
        rd(in, lm->t_w);
        in >> init >> outl >> nfeat;
        lm->initialized = init;
        lm->outlier = outl;
        for (int q = 0; q < nfeat; q++) {
            int fi, si, alive;
            auto ft = std::make_shared<Feature>();
            in >> fi >> si;
            rd(in, ft->bearing);
            rd(in, ft->uv);
            in >> alive;
            ft->sensor = all_frames[fi]->sensors[si];
            lm->features.push_back(ft);
            if (alive) {
                keep_alive.push_back(ft); // a dead weak_ptr models a feature whose frame was dropped
                auto &cloud = all_frames[fi]->pointxd; // frame->getLandmarks()["pointxd"]: the landmarks the frame observes
                if (cloud.empty() || cloud.back() != lm) cloud.push_back(lm);
            }
        }
        map->pointxd.push_back(lm);
    }
    // optional marginal prior: "<sparsif> <frame_to_keep|-1> <frame_col> <n> <n_full> <n_keep>", per kept landmark
    // "<landmark> <col> <delta 3> <sqrt_inf 9>", then J, r0, the 15x15 frame information, "<lmk_with_prior|-1> <prior 3> <info 9>"
    auto marg = std::make_shared<Marginalization>();
    bool sparsif = false;
    int has_prior_section = 0;
    if (in >> has_prior_section && has_prior_section) {
        int sp, ftk, fcol, nkeep;
        in >> sp >> ftk >> fcol >> marg->n >> marg->n_full >> nkeep;
        sparsif = sp != 0;
        if (ftk >= 0) {
            marg->frame_to_keep = all_frames[ftk];
            marg->map_frame_idx[all_frames[ftk].get()] = fcol;
        }
        for (int k = 0; k < nkeep; k++) {
            int l, col;
            in >> l >> col;
            auto &lm = map->pointxd[l];
            marg->lmk_to_keep.push_back(lm);
            marg->map_lmk_idx[lm.get()] = col;
            rd(in, marg->map_lmk_prior[lm.get()]);
            rd(in, marg->map_lmk_inf[lm.get()]);
        }
        marg->marginalization_jacobian.resize((size_t)marg->n * marg->n_full);
        marg->marginalization_residual.resize(marg->n_full);
        for (auto &x : marg->marginalization_jacobian) in >> x;
        for (auto &x : marg->marginalization_residual) in >> x;
        if (ftk >= 0) rd(in, marg->map_frame_inf[all_frames[ftk].get()]);
        int lwp;
        in >> lwp;
        if (lwp >= 0) marg->lmk_with_prior = map->pointxd[lwp];
        rd(in, marg->prior_lmk);
        rd(in, marg->info_lmk);
        if (!in) {
            std::fprintf(stderr, "malformed prior section\n");
            return 2;
        }
    } else {
        in.clear();
    }
    if (!in) {
        std::fprintf(stderr, "malformed input\n");
        return 2;
    }
    if (std::getenv("SDV_TEST_NAN_OBS")) { // a NaN measurement makes every step invalid: the solve ends in Ceres' FAILURE termination
        for (auto &lm : map->pointxd)
            for (auto &wf : lm->features)
                if (auto ft = wf.lock()) {
                    ft->bearing[0] = std::nan("");
                    goto poisoned;
                }
    poisoned:;
    }
    std::printf("%s\n", mode.c_str());
    auto dump_view = [](const FlatWindow &fw, bool ok) {
        const sdv_window &w = fw.view;
        std::printf("ok %d\n", ok ? 1 : 0);
        if (!ok) return;
        auto pd = [](const char *name, const double *a, size_t n) {
            std::printf("%s %zu", name, a ? n : 0);
            for (size_t i = 0; a && i < n; i++) std::printf(" %.17g", a[i]);
            std::printf("\n");
        };
        auto pi = [](const char *name, const int32_t *a, size_t n) {
            std::printf("%s %zu", name, a ? n : 0);
            for (size_t i = 0; a && i < n; i++) std::printf(" %d", a[i]);
            std::printf("\n");
        };
        auto pu = [](const char *name, const uint8_t *a, size_t n) {
            std::printf("%s %zu", name, a ? n : 0);
            for (size_t i = 0; a && i < n; i++) std::printf(" %d", (int)a[i]);
            std::printf("\n");
        };
        const size_t F = w.n_frames, C = w.n_cams, L = w.n_lmks, O = w.n_obs, P = w.n_imu;
        std::printf("dims 8 %d %d %d %d %d %d %d %d\n", w.vio, w.factor_kind, w.n_frames, w.n_fixed, w.n_cams, w.n_lmks, w.n_obs, w.n_imu);
        std::printf("masks 3 %.17g %d %d\n", w.visual_loss_huber_a, w.landmarks_constant, w.max_num_iterations);
        pd("T_f_w", w.T_f_w, 12 * F); pd("v", w.v, 3 * F); pd("ba", w.ba, 3 * F); pd("bg", w.bg, 3 * F);
        pu("has_imu", w.has_imu, F); pu("has_prior", w.has_prior, F); pd("T_prior", w.T_prior, 12 * F); pd("inf_prior", w.inf_prior, 6 * F);
        pd("T_s_f", w.T_s_f, 12 * C); pd("K", w.K, 4 * C); pd("lmk_t", w.lmk_t, 3 * L);
        pi("obs_lmk", w.obs_lmk, O); pi("obs_frame", w.obs_frame, O); pi("obs_cam", w.obs_cam, O);
        pd("obs_bearing", w.obs_bearing, 3 * O); pd("obs_uv", w.obs_uv, 2 * O); pd("obs_sigma", w.obs_sigma, O);
        pi("imu_i", w.imu_i, P); pi("imu_j", w.imu_j, P); pd("imu_dt", w.imu_dt, P); pd("imu_dR", w.imu_dR, 9 * P);
        pd("imu_dv", w.imu_dv, 3 * P); pd("imu_dp", w.imu_dp, 3 * P); pd("imu_cov", w.imu_cov, 81 * P);
        pd("imu_J_dR_bg", w.imu_J_dR_bg, 9 * P); pd("imu_J_dv_ba", w.imu_J_dv_ba, 9 * P); pd("imu_J_dv_bg", w.imu_J_dv_bg, 9 * P);
        pd("imu_J_dp_ba", w.imu_J_dp_ba, 9 * P); pd("imu_J_dp_bg", w.imu_J_dp_bg, 9 * P);
        pd("imu_sigma_ba", w.imu_sigma_ba, P); pd("imu_sigma_bg", w.imu_sigma_bg, P);
        if (w.dense_prior) {
            const sdv_dense_prior &d = *w.dense_prior;
            std::printf("dense 5 %d %d %d %d %d\n", d.n_full, d.n, d.frame, d.frame_col, d.n_keep);
            pd("dense_J", d.J, (size_t)d.n_full * d.n); pd("dense_r0", d.r0, d.n_full);
            pi("dense_keep_lmk", d.keep_lmk, d.n_keep); pi("dense_keep_col", d.keep_col, d.n_keep);
        }
        if (w.sparse_prior) {
            const sdv_sparse_prior &q = *w.sparse_prior;
            std::printf("sparse 6 %d %d %d %d %d %d\n", q.has_imu_prior, q.frame, q.n_p2l, q.has_lmk_prior, q.lmk0, q.n_l2l);
            pd("sp_T_prior", q.T_prior, 12); pd("sp_v_prior", q.v_prior, 3); pd("sp_ba_prior", q.ba_prior, 3); pd("sp_bg_prior", q.bg_prior, 3);
            pd("sp_imu_sqrt_inf", q.imu_sqrt_inf, 225);
            pi("sp_p2l_lmk", q.p2l_lmk, q.n_p2l); pd("sp_p2l_delta", q.p2l_delta, 3 * (size_t)q.n_p2l); pd("sp_p2l_sqrt_inf", q.p2l_sqrt_inf, 9 * (size_t)q.n_p2l);
            pd("sp_lmk_prior", q.lmk_prior, 3); pd("sp_lmk_sqrt_inf", q.lmk_sqrt_inf, 9);
            pi("sp_l2l_a", q.l2l_a, q.n_l2l); pi("sp_l2l_b", q.l2l_b, q.n_l2l); pd("sp_l2l_delta", q.l2l_delta, 3 * (size_t)q.n_l2l);
            pd("sp_l2l_sqrt_inf", q.l2l_sqrt_inf, 9 * (size_t)q.n_l2l);
        }
    };
    if (mode == "dump") {
        FlatWindow fw;
        bool ok = flatten(*map, fixed, vio, kind, fw, marg.get(), sparsif);
        dump_view(fw, ok);
        return 0;
    }
    if (mode == "dump_lmkopt" || mode == "dump_single" || mode == "dump_singlevi") {
        FlatWindow fw;
        if (mode == "dump_lmkopt") flatten_landmark_cloud(*all_frames[f2_frame], kind, fw);
        else flatten_single_frame(all_frames[f2_frame], mode == "dump_singlevi", kind, fw);
        dump_view(fw, true);
        // the frames of the sub-window as positions in the file (the Python mirror works with window indices)
        std::printf("frames %zu", fw.frame_vector.size());
        for (auto &fr : fw.frame_vector)
            for (int f = 0; f < NF; f++)
                if (all_frames[f] == fr) std::printf(" %d", f);
        std::printf("\n");
        return 0;
    }
    if (mode == "dump_viinit") {
        FlatWindow fw;
        dump_view(fw, flatten_viinit(*map, fw));
        return 0;
    }
    if (mode == "flatten") {
        FlatWindow fw;
        flatten(*map, fixed, vio, kind, fw);
        const sdv_window &w = fw.view;
        std::printf("%d %d %d %d %d\n", w.n_frames, w.n_cams, w.n_lmks, w.n_obs, w.n_imu);
        for (int o = 0; o < w.n_obs; o++) std::printf("%d %d %d\n", w.obs_lmk[o], w.obs_frame[o], w.obs_cam[o]);
        for (int p = 0; p < w.n_imu; p++) std::printf("%d %d %.17g\n", w.imu_i[p], w.imu_j[p], w.imu_dt[p]);
        for (int f = 0; f < w.n_frames; f++) std::printf("%.17g\n", w.T_f_w[12 * f + 3]);
        return 0;
    }
    if (mode == "writeback") {
        // no GPU needed: flatten, then apply a synthetic solution (entry k of the concatenated blocks = 1e-3 sin(k + 1)) with the
        // adapter's write-back.  tests/test_host_adapter.py applies the same solution with the Python mirror.
        FlatWindow fw;
        if (!flatten(*map, fixed, vio, kind, fw, marg.get(), sparsif)) return 3;
        const size_t F = fw.frame_vector.size(), L = fw.landmarks.size();
        std::vector<double> buf(15 * F + 3 * L + 1, 0.0);
        for (size_t k = 0; k < 15 * F + 3 * L; k++) buf[k] = 1e-3 * std::sin((double)(k + 1));
        sdv_delta d;
        d.dpose = buf.data();
        d.dv = d.dpose + 6 * F;
        d.dba = d.dv + 3 * F;
        d.dbg = d.dba + 3 * F;
        d.dlmk = d.dbg + 3 * F;
        write_back(fw, d, vio);
        std::printf("1 0\n");
    } else if (mode == "viinit_writeback") {
        // no GPU needed: VIInit's state update with a synthetic result (dv entry k = 1e-2 sin(k + 1), r_wi = (0.1, -0.05), lambda = log 2)
        FlatWindow fw;
        if (!flatten_viinit(*map, fw)) return 3;
        std::vector<double> dv(3 * fw.frame_vector.size());
        for (size_t k = 0; k < dv.size(); k++) dv[k] = 1e-2 * std::sin((double)(k + 1));
        sdv_viinit_result r;
        std::memset(&r, 0, sizeof(r));
        r.dv = dv.data();
        r.r_wi[0] = 0.1;
        r.r_wi[1] = -0.05;
        r.lambda = std::log(2.0);
        const double w3[3] = {0.1, -0.05, 0.0};
        detail::exp_so3(w3, r.R_w_i);
        viinit_write_back(*map, fw, r);
        std::printf("1 0\n");
    } else if (mode == "viinit") {
        // VIInit(local_map, R_w_i, optim_scale = argv[5])
        B200Optimizer opt(kind, 0);
        double Rwi[9] = {0};
        const double scale = opt.VIInit(map, Rwi, f2_frame != 0);
        std::printf("%d %d %.17g", std::isnan(scale) ? 0 : 1, opt.lastStats().iterations, scale);
        for (double x : Rwi) std::printf(" %.17g", x);
        std::printf("\n");
    } else if (mode == "marg") {
        // marginalize(frame0 = oldest keyframe of the window, frame1 = the next one, enable_sparsif = argv[5]), then the window
        // without frame 0 (and without the landmarks only frame 0 saw, unless the prior keeps them) is solved with the prior
        B200Optimizer opt(kind, 0);
        std::shared_ptr<Frame> frame0 = map->frames[0], frame1 = map->frames[1];
        const bool okm = opt.marginalize(frame0, frame1, f2_frame != 0);
        const sdv_marginal_sizes &sz = opt.lastMarginalSizes();
        std::printf("%d %d %d %d %d %d\n", okm ? 1 : 0, sz.m, sz.n, sz.n_full, sz.n_keep, sz.n_marg);
        map->frames.pop_front();
        std::vector<std::shared_ptr<Landmark>> alive;
        for (auto &lm : map->pointxd) {
            bool seen = false;
            for (auto &wf : lm->features)
                if (auto ft = wf.lock())
                    if (auto cam = ft->sensor.lock())
                        for (auto &fr : map->frames) seen |= cam->getFrame() == fr;
            for (auto &k : opt._marginalization->lmk_to_keep) seen |= k == lm;
            if (seen) alive.push_back(lm);
            else lm->t_w = {1e300, 1e300, 1e300}; // (printed below: marks a landmark that left the map)
        }
        auto all_lmks = map->pointxd;
        map->pointxd = alive;
        bool ok = vio ? opt.localMapVIOptimization(map, fixed) : opt.localMapBA(map, fixed);
        map->pointxd = all_lmks;
        map->frames.push_front(frame0);
        std::printf("%d %d\n", ok ? 1 : 0, opt.lastStats().iterations);
    } else if (mode == "lmkopt" || mode == "single" || mode == "singlevi") {
        B200Optimizer opt(kind, 0);
        bool ok = mode == "lmkopt" ? opt.landmarkOptimization(all_frames[f2_frame])
                                   : (mode == "single" ? opt.singleFrameOptimization(all_frames[f2_frame]) : opt.singleFrameVIOptimization(all_frames[f2_frame]));
        std::printf("%d %d\n", ok ? 1 : 0, opt.lastStats().iterations);
    } else {
        B200Optimizer opt(kind, 0);
        opt._marginalization = marg;
        opt._enable_sparsif = sparsif;
        bool ok = vio ? opt.localMapVIOptimization(map, fixed) : opt.localMapBA(map, fixed);
        std::printf("%d %d\n", ok ? 1 : 0, opt.lastStats().iterations);
    }
    for (auto &fr : map->frames) {
        for (double x : fr->T_f_w) std::printf("%.17g ", x);
        if (fr->imu) {
            for (double x : fr->imu->v) std::printf("%.17g ", x);
            for (double x : fr->imu->ba) std::printf("%.17g ", x);
            for (double x : fr->imu->bg) std::printf("%.17g ", x);
            for (double x : fr->imu->delta_p) std::printf("%.17g ", x);
            for (double x : fr->imu->delta_v) std::printf("%.17g ", x);
            for (double x : fr->imu->delta_R) std::printf("%.17g ", x);
        }
        std::printf("\n");
    }
    for (auto &lm : map->pointxd) std::printf("%.17g %.17g %.17g %d\n", lm->t_w[0], lm->t_w[1], lm->t_w[2], lm->outlier ? 1 : 0);
    return 0;
}
