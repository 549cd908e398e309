// Test driver for sadvio_b200/host/b200_optimizer.hpp: reads a pointer-graph description (text, produced by
// tests/test_host_adapter.py), rebuilds the graph with shared_ptr / weak_ptr objects exactly as SaDVIO holds it,
// and either prints the flattened window ("flatten") or runs localMapVIOptimization / localMapBA through the C ABI
// and prints the updated state ("solve").
#include "b200_optimizer.hpp"

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
            if (alive) keep_alive.push_back(ft); // a dead weak_ptr models a feature whose frame was dropped
        }
        map->pointxd.push_back(lm);
    }
    if (!in) {
        std::fprintf(stderr, "malformed input\n");
        return 2;
    }
    std::printf("%s\n", mode.c_str());
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
    B200Optimizer opt(kind, 0);
    bool ok = vio ? opt.localMapVIOptimization(map, fixed) : opt.localMapBA(map, fixed);
    std::printf("%d %d\n", ok ? 1 : 0, opt.lastStats().iterations);
    for (auto &fr : map->frames) {
        for (double x : fr->T_f_w) std::printf("%.17g ", x);
        if (fr->imu) {
            for (double x : fr->imu->v) std::printf("%.17g ", x);
            for (double x : fr->imu->ba) std::printf("%.17g ", x);
            for (double x : fr->imu->bg) std::printf("%.17g ", x);
            for (double x : fr->imu->delta_p) std::printf("%.17g ", x);
        }
        std::printf("\n");
    }
    for (auto &lm : map->pointxd) std::printf("%.17g %.17g %.17g\n", lm->t_w[0], lm->t_w[1], lm->t_w[2]);
    return 0;
}
