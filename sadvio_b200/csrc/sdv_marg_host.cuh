// Host side of sdv_marginalize / sdv_marginal_fetch / sdv_schur_prior (included at the end of sdv_lib.cu: it uses the handle,
// the upload path and the factor kernels of the window solve).  Replaces, behind the C ABI:
//   Marginalization::preMarginalize                 cpp/src/optimizers/marginalization.cpp:23-143   (node selection: host, integers only)
//   AngularAdjustmentCERESAnalytic::marginalize     cpp/src/optimizers/AngularAdjustmentCERESAnalytic.cpp:488-739
//   BundleAdjustmentCERESAnalytic::marginalize      cpp/src/optimizers/BundleAdjustmentCERESAnalytic.cpp:431-660
//   Marginalization::computeSchurComplement / rankReveallingDecomposition / computeJacobiansAndResiduals / sparsifyVIO / sparsifyVO
// All arithmetic runs on the device (sdv_marg.cuh); the host selects nodes, sorts the eigenvalues it thresholds and walks the
// greedy chain of sparsifyVO over the coupling matrix the device computed.
#pragma once

struct MargState {
    // device scratch (grown on demand)
    unsigned char *d = nullptr, *hp = nullptr; // device arena, pinned host mirror of the result part
    size_t d_cap = 0, hp_cap = 0;
    // result of the last call
    bool valid = false, have_A = false;
    sdv_marginal_sizes sz{};
    int sparsify = 0, vio = 0;
    std::vector<int> keep, marg, chain;
    int with_prior = -1;
    // offsets into the pinned mirror (bytes)
    size_t o_J = 0, o_r0 = 0, o_Ak = 0, o_bk = 0, o_U = 0, o_Lam = 0, o_A = 0, o_b = 0, o_imu = 0, o_p2d = 0, o_p2s = 0, o_lsq = 0, o_l2d = 0, o_l2s = 0, o_lpr = 0;
};

namespace {

constexpr int EIG_MAX_SWEEPS = 40;

int eig_pad(int n) { return std::max(2 * EB, (n + EB - 1) / EB * EB); }

// symmetric eigen-decomposition of the np x np matrix in G (overwritten by G V); eigenvectors in V, eigenvalues in w
int eig_sym(sdv_handle *h, double *G, double *V, int np, double *w, unsigned long long *flags, int *sweeps_out) {
    cudaStream_t s = h->stream;
    const int nb = np / EB, nbp = nb + (nb & 1), steps = nbp - 1, pairs = nbp / 2;
    CK(cudaMemsetAsync(flags, 0, sizeof(unsigned long long) * EIG_MAX_SWEEPS, s));
    double *wmax = reinterpret_cast<double *>(flags + EIG_MAX_SWEEPS);
    k_max_colnorm2<<<1, 256, 0, s>>>(G, np, wmax);
    const double negl = np * 2.220446049250313e-16; // entries below this x (largest column norm) are rounding noise
    static const int inner = getenv("SDV_EIG_INNER") ? std::max(1, atoi(getenv("SDV_EIG_INNER"))) : 2;
    k_set_identity<<<std::min(1024, (np * np + 255) / 256), 256, 0, s>>>(V, np);
    h->launches++;
    const double tol = 2.220446049250313e-16 * np; // |h_pq| / sqrt(h_pp h_qq): the rounding floor of an np-term dot product
    unsigned long long *hflags = reinterpret_cast<unsigned long long *>(h->h_rb); // (pinned, >= 256 bytes + LMState)
    int sweep = 0;
    // cluster size of k_jacobi_pairs_cl: the largest one (<= 8) that keeps every block pair of a step resident at once; rows split in
    // multiples of the 32-row tiles.  SDV_EIG_CLUSTER=1 (or a matrix of one or two tiles) selects the one-CTA kernel.
    static const int cs_max = getenv("SDV_EIG_CLUSTER") ? std::max(1, std::min(8, atoi(getenv("SDV_EIG_CLUSTER")))) : 8;
    int cs = 1, rows_per_cta = np;
    cudaLaunchConfig_t lc = {};
    cudaLaunchAttribute lat[1];
    if (np > 64) {
        for (int c = cs_max; c >= 2; c--) {
            lc.gridDim = dim3(pairs * c);
            lc.blockDim = dim3(ET);
            lc.dynamicSmemBytes = 0;
            lc.stream = s;
            lat[0].id = cudaLaunchAttributeClusterDimension;
            lat[0].val.clusterDim.x = c;
            lat[0].val.clusterDim.y = 1;
            lat[0].val.clusterDim.z = 1;
            lc.attrs = lat;
            lc.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, k_jacobi_pairs_cl, &lc) == cudaSuccess && ncl >= pairs) {
                cs = c;
                break;
            }
            cudaGetLastError();
        }
        if (cs > 1) rows_per_cta = ((np + cs - 1) / cs + 31) / 32 * 32;
    }
    for (; sweep < EIG_MAX_SWEEPS; sweep++) {
        for (int st = 0; st < steps; st++) {
            if (cs > 1) {
                cudaError_t e = cudaLaunchKernelEx(&lc, k_jacobi_pairs_cl, G, V, np, nb, nbp, st, flags, sweep, tol, inner, (const double *)wmax, negl, rows_per_cta);
                if (e != cudaSuccess) return fail(h, SDV_ERR_CUDA, std::string("k_jacobi_pairs_cl launch: ") + cudaGetErrorString(e));
            } else
                k_jacobi_pairs<<<pairs, ET, 0, s>>>(G, V, np, nb, nbp, st, flags, sweep, tol, inner, wmax, negl);
        }
        h->launches += steps;
        if (sweep >= 3) { // look at the convergence flag (one small D2H per sweep from here on)
            CK(cudaMemcpyAsync(hflags, flags + sweep, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            double off;
            std::memcpy(&off, hflags, sizeof(double));
            if (off <= tol) {
                sweep++;
                break;
            }
        }
    }
    if (sweeps_out) *sweeps_out = sweep;
    k_eig_values<<<(np + 127) / 128, 128, 0, s>>>(G, V, np, w);
    h->launches++;
    CK(cudaGetLastError());
    return SDV_OK;
}

void mm(sdv_handle *h, double *C, int ldc, const double *A, int lda, const double *B, int ldb, int transB, const double *d, int M, int N, int K, double alpha,
        const double *Cin, int ldcin, double beta) {
    dim3 grid((N + 31) / 32, (M + 31) / 32);
    k_mm<<<grid, 256, 0, h->stream>>>(C, ldc, A, lda, B, ldb, transB, d, M, N, K, alpha, Cin, ldcin, beta);
    h->launches++;
}

struct MargLayout { // offsets (doubles) into the device arena
    size_t A, b, G1, V1, w1, winv, Bv, cv, Ak, bk, G2, V2, w2, J, r0, U, Lam, flags, order, mi, ent, imu, p2d, p2s, lsq, l2d, l2s, end;
};

// Schur complement + rank-revealing decomposition + J / r0 of the information (A, b) that sits at L.A / L.b of the device arena.
// marginalization.cpp:213-265, :318-342, :516-530.
int schur_core(sdv_handle *h, MargState &ms, const MargLayout &L, int m, int n, double eps) {
    cudaStream_t s = h->stream;
    double *base = reinterpret_cast<double *>(ms.d);
    const int N = m + n, np1 = eig_pad(m), np2 = eig_pad(n);
    double *A = base + L.A, *b = base + L.b;
    unsigned long long *flags = reinterpret_cast<unsigned long long *>(base + L.flags);
    int sw1 = 0, sw2 = 0, rc;
    if (m > 0) {
        // Amm = 1/2 (A_mm + A_mm^T), eigen-decomposition, pseudo-inverse with the eigenvalues <= eps dropped (:228-240)
        k_sym_block<<<std::min(2048, (np1 * np1 + 255) / 256), 256, 0, s>>>(A, N, 0, m, base + L.G1, np1);
        h->launches++;
        if ((rc = eig_sym(h, base + L.G1, base + L.V1, np1, base + L.w1, flags, &sw1)) != SDV_OK) return rc;
        k_pinv_diag<<<(np1 + 127) / 128, 128, 0, s>>>(base + L.w1, np1, eps, base + L.winv);
        h->launches++;
        // Ak = Arr - Arm Amm^+ Arm^T, bk = brr - Arm Amm^+ bmm (:242-248), formed as (Arm V) diag(1/w) (Arm V)^T
        mm(h, base + L.Bv, np1, A + (size_t)m * N, N, base + L.V1, np1, 0, nullptr, n, np1, m, 1.0, nullptr, 0, 0.0);   // Bv = Arm V   [n x np1]
        mm(h, base + L.cv, np1, b, m, base + L.V1, np1, 0, nullptr, 1, np1, m, 1.0, nullptr, 0, 0.0);                   // cv = bmm^T V [1 x np1]
        mm(h, base + L.Ak, n, base + L.Bv, np1, base + L.Bv, np1, 1, base + L.winv, n, n, np1, -1.0, A + (size_t)m * N + m, N, 1.0);
        mm(h, base + L.bk, 1, base + L.Bv, np1, base + L.cv, np1, 1, base + L.winv, n, 1, np1, -1.0, b + m, 1, 1.0);
    } else { // nothing to marginalise (only through sdv_schur_prior): Ak = Arr, bk = brr
        mm(h, base + L.Ak, n, A, N, A, N, 0, nullptr, n, n, 0, 0.0, A, N, 1.0);
        mm(h, base + L.bk, 1, A, N, A, N, 0, nullptr, n, 1, 0, 0.0, b, 1, 1.0);
    }
    // rank-revealing decomposition of Ak (:318-342)
    k_sym_block<<<std::min(2048, (np2 * np2 + 255) / 256), 256, 0, s>>>(base + L.Ak, n, 0, n, base + L.G2, np2);
    h->launches++;
    if ((rc = eig_sym(h, base + L.G2, base + L.V2, np2, base + L.w2, flags, &sw2)) != SDV_OK) return rc;
    std::vector<double> w(np2);
    CK(cudaMemcpyAsync(w.data(), base + L.w2, sizeof(double) * np2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<int> order;
    for (int i = 0; i < np2; i++)
        if (w[i] > eps) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return w[a] < w[c]; }); // ascending, like SelfAdjointEigenSolver
    const int n_full = (int)order.size();
    ms.sz.n_full = n_full;
    ms.sz.eig_sweeps_m = sw1;
    ms.sz.eig_sweeps_n = sw2;
    if (n_full > 0) {
        int *d_order = reinterpret_cast<int *>(base + L.order);
        CK(cudaMemcpyAsync(d_order, order.data(), sizeof(int) * n_full, cudaMemcpyHostToDevice, s));
        k_marg_build<<<n_full, 128, 0, s>>>(base + L.V2, np2, base + L.w2, d_order, n, n_full, base + L.bk, base + L.J, base + L.r0, base + L.U, base + L.Lam);
        h->launches++;
    }
    CK(cudaGetLastError());
    return SDV_OK;
}

MargLayout marg_layout(int m, int n, int K) {
    MargLayout L;
    size_t o = 0;
    auto add = [&](size_t cnt) {
        size_t r = o;
        o += (cnt + 31) & ~size_t(31);
        return r;
    };
    const size_t N = (size_t)m + n, np1 = eig_pad(m), np2 = eig_pad(n);
    L.A = add(N * N); L.b = add(N);
    L.G1 = add(np1 * np1); L.V1 = add(np1 * np1); L.w1 = add(np1); L.winv = add(np1);
    L.Bv = add((size_t)n * np1); L.cv = add(np1);
    L.Ak = add((size_t)n * n); L.bk = add(n);
    L.G2 = add(np2 * np2); L.V2 = add(np2 * np2); L.w2 = add(np2);
    L.J = add((size_t)n * n); L.r0 = add(n); L.U = add((size_t)n * n); L.Lam = add(n);
    L.flags = add(EIG_MAX_SWEEPS + 8); L.order = add(np2);
    L.mi = add((size_t)K * K); L.ent = add(K);
    L.imu = add(225); L.p2d = add(3 * (size_t)K); L.p2s = add(9 * (size_t)K); L.lsq = add(9); L.l2d = add(3 * (size_t)K); L.l2s = add(9 * (size_t)K);
    L.end = o; // (the int lists of a call — selections, column maps, chain — follow the doubles)
    return L;
}

// result part of the device arena -> pinned mirror -> (on fetch) the caller's buffers
int marg_readback(sdv_handle *h, MargState &ms, const MargLayout &L, int m, int n, bool with_A) {
    const size_t D = sizeof(double);
    const int N = m + n, nf = ms.sz.n_full, K = (int)ms.keep.size();
    Arena R;
    ms.o_J = R.add(D * (size_t)std::max(nf, 1) * n); ms.o_r0 = R.add(D * std::max(nf, 1));
    ms.o_Ak = R.add(D * (size_t)n * n); ms.o_bk = R.add(D * n);
    ms.o_U = R.add(D * (size_t)n * std::max(nf, 1)); ms.o_Lam = R.add(D * std::max(nf, 1));
    ms.o_A = R.add(D * (size_t)N * N); ms.o_b = R.add(D * N);
    ms.o_imu = R.add(D * 225); ms.o_p2d = R.add(D * 3 * std::max(K, 1)); ms.o_p2s = R.add(D * 9 * std::max(K, 1));
    ms.o_lsq = R.add(D * 9); ms.o_l2d = R.add(D * 3 * std::max(K, 1)); ms.o_l2s = R.add(D * 9 * std::max(K, 1));
    int rc;
    if ((rc = ensure(h, &ms.hp, &ms.hp_cap, R.size, true)) != SDV_OK) return rc;
    cudaStream_t s = h->stream;
    double *base = reinterpret_cast<double *>(ms.d);
    auto get = [&](size_t off, size_t src, size_t cnt) { return cnt == 0 ? cudaSuccess : cudaMemcpyAsync(ms.hp + off, base + src, D * cnt, cudaMemcpyDeviceToHost, s); };
    CK(get(ms.o_J, L.J, (size_t)nf * n)); CK(get(ms.o_r0, L.r0, nf));
    CK(get(ms.o_Ak, L.Ak, (size_t)n * n)); CK(get(ms.o_bk, L.bk, n));
    CK(get(ms.o_U, L.U, (size_t)n * nf)); CK(get(ms.o_Lam, L.Lam, nf));
    if (with_A) {
        CK(get(ms.o_A, L.A, (size_t)N * N)); CK(get(ms.o_b, L.b, N));
    }
    ms.have_A = with_A;
    if (ms.sparsify && ms.vio && nf > 0) {
        CK(get(ms.o_imu, L.imu, 225)); CK(get(ms.o_p2d, L.p2d, 3 * (size_t)K)); CK(get(ms.o_p2s, L.p2s, 9 * (size_t)K));
    }
    if (ms.sparsify && !ms.vio && ms.chain.size() >= 2) {
        const size_t nl = ms.chain.size() - 1;
        CK(get(ms.o_lsq, L.lsq, 9)); CK(get(ms.o_l2d, L.l2d, 3 * nl)); CK(get(ms.o_l2s, L.l2s, 9 * nl));
    }
    CK(cudaStreamSynchronize(s));
    return SDV_OK;
}

MargState &marg_state(sdv_handle *h) {
    if (!h->marg) h->marg = new MargState();
    return *h->marg;
}

} // namespace

extern "C" {

int sdv_marginalize(sdv_handle *h, const sdv_window *win, int32_t sparsify, sdv_marginal_sizes *sizes) {
    if (!h || !win || !sizes) return SDV_ERR_INVALID_ARGUMENT;
    MargState &ms = marg_state(h);
    ms.valid = false;
    std::memset(&ms.sz, 0, sizeof(ms.sz));
    const int F = win->n_frames, L = win->n_lmks, O = win->n_obs;
    if (F < 2) return fail(h, SDV_ERR_INVALID_ARGUMENT, "marginalisation needs frame 0 (the oldest keyframe) and frame 1");
    if (win->sparse_prior) return fail(h, SDV_ERR_INVALID_ARGUMENT, "the previous prior is always propagated in its dense form (_marginalization_last)");
    const double eps = 1e-12; // Marginalization::_eps (marginalization.hpp:79)
    auto t0 = std::chrono::steady_clock::now();
    // ---- the window goes to the device through the upload path of the solve, every block free (the factors are evaluated at
    //      dx = 0 for ALL their parameter blocks: a marginalisation block knows no constant parameter)
    sdv_window w2 = *win;
    w2.n_fixed = 0;
    w2.visual_loss_huber_a = 0.0;
    w2.landmarks_constant = 0;
    int rc = upload_impl(h, &w2, false);
    if (rc != SDV_OK) return rc;
    const int f0 = F - 1, f1 = F - 2; // frames are ordered newest -> oldest (amap.h:28-32)
    const bool vio = win->vio != 0;
    const sdv_dense_prior *last = win->dense_prior;
    // ---- Marginalization::preMarginalize for point landmarks (marginalization.cpp:23-143)
    std::vector<char> with_prior(std::max(L, 1), 0); // ALandmark::hasPrior()
    if (win->lmk_has_prior) {
        for (int l = 0; l < L; l++) with_prior[l] = win->lmk_has_prior[l] != 0;
    } else if (last) {
        for (int k = 0; k < last->n_keep; k++) with_prior[last->keep_lmk[k]] = 1;
    }
    ms.keep.clear();
    ms.marg.clear();
    std::vector<int> col(std::max(L, 1), -1);
    {
        int o = 0;
        while (o < O) {
            const int l = win->obs_lmk[o];
            int num_cam = 0, others = 0, e = o;
            for (; e < O && win->obs_lmk[e] == l; e++) (win->obs_frame[e] == f0 ? num_cam : others)++;
            o = e;
            if (num_cam == 0) continue;                          // not one of frame 0's landmarks
            if (num_cam != 2 && !with_prior[l]) continue;        // no stereo pair, no prior: ignored (:58-77)
            (others == 0 ? ms.marg : ms.keep).push_back(l);      // :80-89
        }
    }
    int last_idx = 6 + (vio ? 9 : 0); // :40-48
    const int col_f0 = 0;
    for (int l : ms.marg) { // :93-98
        col[l] = last_idx;
        last_idx += 3;
    }
    const int m = last_idx;
    int n = 0, col_f1 = -1;
    if (vio) { // :101-106
        col_f1 = last_idx;
        last_idx += 15;
        n += 15;
    }
    for (int l : ms.keep) { // :109-114
        col[l] = last_idx;
        last_idx += 3;
        n += 3;
    }
    if (last)
        for (int k = 0; k < last->n_keep; k++) { // "resurrected" landmarks of the previous prior, :118-139
            const int l = last->keep_lmk[k];
            if (col[l] < 0) {
                ms.keep.push_back(l);
                col[l] = last_idx;
                last_idx += 3;
                n += 3;
            }
        }
    const int N = m + n, K = (int)ms.keep.size();
    ms.sz.m = m;
    ms.sz.n = n;
    ms.sz.n_marg = (int)ms.marg.size();
    ms.sz.n_keep = K;
    ms.sz.frame = vio ? f1 : -1;
    ms.sparsify = sparsify;
    ms.vio = vio;
    ms.chain.clear();
    ms.with_prior = -1;
    if (n < 4) { // computeSchurComplement returns false (:215): the caller resets its marginalisation scheme
        ms.sz.ok = 0;
        ms.valid = true;
        *sizes = ms.sz;
        return SDV_OK;
    }
    // ---- selection lists
    std::vector<int> sel_obs, sel_col, mp_col;
    for (int o = 0; o < O; o++)
        if (win->obs_frame[o] == f0 && col[win->obs_lmk[o]] >= 0) {
            // every feature of frame 0 on a kept or marginalised landmark (…Analytic.cpp:565-629); resurrected landmarks have none
            sel_obs.push_back(o - h->P.o0);
            sel_col.push_back(col[win->obs_lmk[o]]);
        }
    if (last && last->n_keep > 0) { // …Analytic.cpp:631-660 (the block exists only when the previous prior kept landmarks)
        if (last->frame >= 0) {
            if (last->frame != f0) return fail(h, SDV_ERR_INVALID_ARGUMENT, "the previous prior must sit on the frame that is marginalised now");
            for (int q = 0; q < 15; q++) mp_col.push_back(vio ? col_f0 + q : (q < 6 ? col_f0 + q : -1));
        }
        for (int k = 0; k < last->n_keep; k++)
            if (last->keep_col[k] >= 0)
                for (int q = 0; q < 3; q++) mp_col.push_back(col[last->keep_lmk[k]] + q);
        if ((int)mp_col.size() != h->P.mp_nmap) return fail(h, SDV_ERR_INVALID_ARGUMENT, "previous prior: column map mismatch");
    }
    int imu_pair = -1;
    if (vio)
        for (int p = 0; p < win->n_imu; p++)
            if (win->imu_i[p] == f0 && win->imu_j[p] == f1) imu_pair = p;
    if (vio && imu_pair < 0) // the reference builds the IMUFactor from the two IMU objects whatever their distance (…Analytic.cpp:505-559)
        return fail(h, SDV_ERR_UNSUPPORTED, "VIO marginalisation needs the pre-integration between frame 0 and frame 1 in the window's IMU list");
    // ---- device arena
    const MargLayout Lo = marg_layout(m, n, K);
    const size_t ints_needed = sel_obs.size() * 2 + mp_col.size() + 3 * (size_t)K + 64;
    if ((rc = ensure(h, &ms.d, &ms.d_cap, sizeof(double) * Lo.end + sizeof(int) * ints_needed)) != SDV_OK) return rc;
    cudaStream_t s = h->stream;
    double *base = reinterpret_cast<double *>(ms.d);
    int *d_ints = reinterpret_cast<int *>(base + Lo.end);
    std::vector<int> hints;
    hints.insert(hints.end(), sel_obs.begin(), sel_obs.end());
    hints.insert(hints.end(), sel_col.begin(), sel_col.end());
    hints.insert(hints.end(), mp_col.begin(), mp_col.end());
    hints.insert(hints.end(), ms.keep.begin(), ms.keep.end());
    if (!hints.empty()) CK(cudaMemcpyAsync(d_ints, hints.data(), sizeof(int) * hints.size(), cudaMemcpyHostToDevice, s));
    MargPlan M;
    M.N = N; M.m = m; M.n = n;
    M.col_f0 = col_f0; M.col_f1 = col_f1;
    M.nsel = (int)sel_obs.size();
    M.sel_obs = d_ints;
    M.sel_col = d_ints + sel_obs.size();
    M.imu_pair = imu_pair;
    // pose priors: the angular optimizer adds frame 0's and frame 1's (…Analytic.cpp:664-687), the pixel one only frame 0's
    // (BundleAdjustmentCERESAnalytic.cpp:606-617); frame 1 has a column block only in the VIO case
    M.prior_f0 = (win->has_prior && win->has_prior[f0]) ? f0 : -1;
    M.prior_f1 = (win->has_prior && win->has_prior[f1] && win->factor_kind == SDV_FACTOR_ANGULAR && col_f1 >= 0) ? f1 : -1;
    M.nmap = (int)mp_col.size();
    M.mp_col = d_ints + 2 * sel_obs.size();
    const int *d_keep = d_ints + 2 * sel_obs.size() + mp_col.size();
    CK(cudaEventRecord(h->ev[2], s));
    // ---- factor values at dx = 0 (the prologue of a solve), then the information matrix
    k_reset<<<64, 256, 0, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc);
    k_prep_table<<<(h->P.F * h->P.C + 127) / 128, 128, 0, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, 0);
    h->launches += 2;
    launch_lin_factors(h, 0, s);
    CK(cudaMemsetAsync(base + Lo.A, 0, sizeof(double) * ((size_t)N * N), s));
    CK(cudaMemsetAsync(base + Lo.b, 0, sizeof(double) * N, s));
    if (M.nsel > 0) {
        const int grid = std::min(h->num_sms * 4, (M.nsel + 127) / 128);
        if (win->factor_kind == SDV_FACTOR_ANGULAR) k_marg_visual<0><<<grid, 128, 0, s>>>(h->d_P, h->B[0], M, base + Lo.A, base + Lo.b);
        else k_marg_visual<1><<<grid, 128, 0, s>>>(h->d_P, h->B[0], M, base + Lo.A, base + Lo.b);
        h->launches++;
    }
    k_marg_factors<<<1, 256, 0, s>>>(h->d_P, h->B[0], M, base + Lo.A, base + Lo.b);
    h->launches++;
    CK(cudaGetLastError());
    // ---- Schur complement, rank-revealing decomposition, J and r0
    if ((rc = schur_core(h, ms, Lo, m, n, eps)) != SDV_OK) return rc;
    const int nf = ms.sz.n_full;
    // ---- sparsification (…Analytic.cpp:703-708)
    if (sparsify && nf > 0) {
        if (vio) {
            k_sparsify_vio<<<1 + K, 32, 0, s>>>(h->d_P, f1, d_keep, K, base + Lo.U, base + Lo.Lam, nf, eps, base + Lo.imu, base + Lo.p2d, base + Lo.p2s);
            h->launches++;
        } else if (K >= 2) {
            // sparsifyVO (marginalization.cpp:410-514): greedy chain over the coupling of the kept landmarks, unary factor on the
            // landmark of least entropy
            const int first = 0;
            k_vo_coupling<<<K, 32, 0, s>>>(base + Lo.Ak, n, first, K, base + Lo.U, base + Lo.Lam, nf, base + Lo.mi, base + Lo.ent);
            h->launches++;
            std::vector<double> mi((size_t)K * K), ent(K);
            CK(cudaMemcpyAsync(mi.data(), base + Lo.mi, sizeof(double) * mi.size(), cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(ent.data(), base + Lo.ent, sizeof(double) * K, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            for (int k = 0; k < K; k++) // the reference fills (k, l) and (l, k) from the block it meets first (:420-431)
                for (int l = k + 1; l < K; l++) mi[(size_t)l * K + k] = mi[(size_t)k * K + l];
            // Eigen's maxCoeff visits a column-major matrix column by column and keeps the first maximum (:439)
            int max_row = 0, max_col = 0;
            double best = -1.0;
            for (int c = 0; c < K; c++)
                for (int r = 0; r < K; r++)
                    if (mi[(size_t)r * K + c] > best) {
                        best = mi[(size_t)r * K + c];
                        max_row = r;
                        max_col = c;
                    }
            std::vector<int> order = {max_row, max_col}; // positions in keep (:440-441)
            for (int r = 0; r < K; r++) mi[(size_t)r * K + max_row] = 0.0; // :444-446
            for (int c = 0; c < K; c++) mi[(size_t)max_row * K + c] = 0.0;
            for (int r = 0; r < K; r++) mi[(size_t)r * K + max_col] = 0.0;
            int cur = max_col;
            for (;;) { // :451-459
                int c = 0;
                for (int q = 1; q < K; q++)
                    if (mi[(size_t)cur * K + q] > mi[(size_t)cur * K + c]) c = q;
                if (mi[(size_t)cur * K + c] == 0.0) break;
                order.push_back(c);
                for (int q = 0; q < K; q++) mi[(size_t)cur * K + q] = 0.0;
                for (int r = 0; r < K; r++) mi[(size_t)r * K + c] = 0.0;
                cur = c;
            }
            int wp = order[0];
            for (int q : order) // :469-470 (first minimum along the chain)
                if (ent[q] < ent[wp]) wp = q;
            std::vector<int> chain_ints;
            for (int q : order) chain_ints.push_back(first + 3 * q);
            for (int q : order) {
                chain_ints.push_back(ms.keep[q]);
                ms.chain.push_back(ms.keep[q]);
            }
            ms.with_prior = ms.keep[wp];
            int *d_chain = d_ints + hints.size(); // (2 K ints reserved above)
            CK(cudaMemcpyAsync(d_chain, chain_ints.data(), sizeof(int) * chain_ints.size(), cudaMemcpyHostToDevice, s));
            const int nc = (int)order.size();
            k_sparsify_vo<<<nc, 32, 0, s>>>(h->d_P, first + 3 * wp, d_chain, d_chain + nc, nc, base + Lo.U, base + Lo.Lam, nf, eps, base + Lo.lsq, base + Lo.l2d, base + Lo.l2s);
            h->launches++;
            CK(cudaStreamSynchronize(s));
        }
    }
    CK(cudaEventRecord(h->ev[3], s));
    if ((rc = marg_readback(h, ms, Lo, m, n, true)) != SDV_OK) return rc;
    float msdev = 0;
    cudaEventElapsedTime(&msdev, h->ev[2], h->ev[3]);
    ms.sz.ok = 1;
    ms.sz.n_chain = (int)ms.chain.size();
    ms.sz.ms_device = msdev;
    ms.sz.ms_total_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    ms.valid = true;
    *sizes = ms.sz;
    return SDV_OK;
}

// The dense core alone on a caller-provided information matrix (row-major (m + n)^2, the m marginalised parameters first) — the
// entry point the reference's own marginalisation KAT (cpp/tests/marginalization_test.cpp:219-223, :300-313) is pinned through.
int sdv_schur_prior(sdv_handle *h, const double *A, const double *b, int32_t m, int32_t n, double eps, sdv_marginal_sizes *sizes) {
    if (!h || !A || !b || !sizes || m < 0 || n <= 0) return SDV_ERR_INVALID_ARGUMENT;
    cudaSetDevice(h->device);
    MargState &ms = marg_state(h);
    ms.valid = false;
    std::memset(&ms.sz, 0, sizeof(ms.sz));
    ms.keep.clear();
    ms.marg.clear();
    ms.chain.clear();
    ms.sparsify = 0;
    ms.vio = 0;
    ms.sz.m = m;
    ms.sz.n = n;
    ms.sz.frame = -1;
    if (n < 4) {
        ms.valid = true;
        *sizes = ms.sz;
        return SDV_OK;
    }
    int rc;
    if ((rc = ensure(h, &h->h_rb, &h->rb_cap, sizeof(LMState) + sizeof(Accum) + 256, true)) != SDV_OK) return rc;
    const MargLayout Lo = marg_layout(m, n, 0);
    if ((rc = ensure(h, &ms.d, &ms.d_cap, sizeof(double) * Lo.end + 1024)) != SDV_OK) return rc;
    double *base = reinterpret_cast<double *>(ms.d);
    const int N = m + n;
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(base + Lo.A, A, sizeof(double) * (size_t)N * N, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(base + Lo.b, b, sizeof(double) * N, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    if ((rc = schur_core(h, ms, Lo, m, n, eps)) != SDV_OK) return rc;
    if ((rc = marg_readback(h, ms, Lo, m, n, false)) != SDV_OK) return rc;
    ms.sz.ok = 1;
    ms.valid = true;
    *sizes = ms.sz;
    return SDV_OK;
}

int sdv_marginal_fetch(sdv_handle *h, sdv_marginal *out) {
    if (!h || !out) return SDV_ERR_INVALID_ARGUMENT;
    if (!h->marg || !h->marg->valid || !h->marg->sz.ok) return fail(h, SDV_ERR_INVALID_ARGUMENT, "no marginalisation result to fetch");
    const MargState &ms = *h->marg;
    const size_t D = sizeof(double);
    const int n = ms.sz.n, nf = ms.sz.n_full, N = ms.sz.m + ms.sz.n, K = ms.sz.n_keep;
    auto put = [&](double *dst, size_t off, size_t cnt) {
        if (dst && cnt) std::memcpy(dst, ms.hp + off, D * cnt);
    };
    put(out->J, ms.o_J, (size_t)nf * n); put(out->r0, ms.o_r0, nf);
    put(out->Ak, ms.o_Ak, (size_t)n * n); put(out->bk, ms.o_bk, n);
    put(out->U, ms.o_U, (size_t)n * nf); put(out->Lambda, ms.o_Lam, nf);
    if (ms.have_A) {
        put(out->A, ms.o_A, (size_t)N * N); put(out->b, ms.o_b, N);
    }
    if (out->keep_lmk && K) std::memcpy(out->keep_lmk, ms.keep.data(), sizeof(int) * K);
    if (out->marg_lmk && ms.sz.n_marg) std::memcpy(out->marg_lmk, ms.marg.data(), sizeof(int) * ms.sz.n_marg);
    if (ms.sparsify && ms.vio && nf > 0) {
        put(out->imu_sqrt_inf, ms.o_imu, 225); put(out->p2l_delta, ms.o_p2d, 3 * (size_t)K); put(out->p2l_sqrt_inf, ms.o_p2s, 9 * (size_t)K);
    }
    out->lmk_with_prior = ms.with_prior;
    if (ms.sparsify && !ms.vio && ms.chain.size() >= 2) {
        const size_t nl = ms.chain.size() - 1;
        if (out->chain) std::memcpy(out->chain, ms.chain.data(), sizeof(int) * ms.chain.size());
        std::memcpy(out->lmk_sqrt_inf, ms.hp + ms.o_lsq, D * 9);
        put(out->l2l_delta, ms.o_l2d, 3 * nl); put(out->l2l_sqrt_inf, ms.o_l2s, 9 * nl);
    }
    return SDV_OK;
}

} // extern "C"
