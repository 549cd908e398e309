// Dense FP64 Cholesky + triangular solves of the reduced (pose/velocity/bias) system in ONE launch:
// a single thread-block cluster (up to 16 CTAs, hardware cluster barrier ~0.2 us) owns the matrix; tile rows are
// distributed cyclically over the CTAs.  Per 32-column panel:
//   A  owner CTA factors the 32x32 diagonal tile in registers (one warp, shuffle broadcast of the pivot column)
//   -- barrier.cluster --
//   B  every CTA solves its own tiles of the panel column (register-resident TRSM, one warp per tile)
//   -- barrier.cluster --
//   C  every CTA updates its own tile rows of the trailing matrix with FP64 tensor-core MMAs
//      (mma.sync.m8n8k4.f64 — tcgen05 has no f64 kind, DMMA is the FP64 tensor path on sm_100a)
// The right-hand side rides along as an extra tile row, so L y = g comes out of phase B; the backward solve
// L^T z = y runs in the same kernel (left-looking, partial products reduced through global scratch + cluster barrier),
// followed by the reduced-parameter update and the candidate frame-camera table on cluster rank 0.
#pragma once
#include "sdv_kernels.cuh"

namespace sdv {

constexpr int CC_MAX = 16;   // largest cluster used
constexpr int CCT = 256;     // threads per CTA
constexpr int TSTR = 33;     // shared-memory tile row stride in doubles: one row per lane is conflict-free (DMMA fragment loads 4-way)
constexpr unsigned FULL = 0xffffffffu;

SDV_DEV unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
SDV_DEV unsigned cluster_size() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
SDV_DEV void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
SDV_DEV void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
SDV_DEV void cluster_sync_all() {
    cluster_arrive();
    cluster_wait();
}

// Software cluster barriers on mbarriers (explicit phase parity, so warps of one CTA may be in different phases —
// the hardware barrier.cluster tolerates no such skew inside a CTA, which the look-ahead needs).
SDV_DEV void mbar_remote_arrive(uint64_t *bar, unsigned target_rank) {
    unsigned local = smem_u32(bar), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(target_rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// one arrival on the barrier `bar` of EVERY CTA of the cluster; called by a full (converged) warp: lane r signals CTA r
SDV_DEV void mbar_arrive_all_warp(uint64_t *bar, int cs, int lane) {
    __syncwarp();
    if (lane < cs) mbar_remote_arrive(bar, (unsigned)lane);
}
SDV_DEV void mbar_arrive_all(uint64_t *bar, int cs) {
    for (int r = 0; r < cs; r++) mbar_remote_arrive(bar, (unsigned)r);
}
SDV_DEV void mbar_wait_cluster(uint64_t *bar, unsigned parity) {
    // spin at CTA scope (a cluster-scope acquire inside the loop invalidates L1 on every poll), then ONE cluster-scope fence
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITC_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEC_%=;\n"
        "bra WAITC_%=;\n"
        "DONEC_%=:\n"
        "fence.acq_rel.cluster;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

SDV_DEV void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// The instruction cache matters here: fully unrolled register-array versions of these routines made the kernel ~29k SASS
// instructions (466 KB), every panel re-fetched its code from L2. The versions below are rolled loops over a tile in
// shared memory (a few hundred instructions in total).

// In-place Cholesky of the 32x32 tile S (row stride TSTR) by one warp, lane = row. Left-looking; the term of the
// previous column is taken from registers (shuffle) and the older terms of the NEXT column are accumulated while the
// reciprocal square root of the current pivot is in flight. sinv[c] = 1 / L[c][c]. Returns false if not positive definite.
SDV_DEV bool chol32_smem(double *S, double *sinv, int lane) {
    bool ok = true;
    double lprev = 0.0;
    double part = S[lane * TSTR]; // column 0 has no earlier terms
#pragma unroll 1
    for (int c = 0; c < 32; c++) {
        double v = part;
        if (c > 0) v -= lprev * __shfl_sync(FULL, lprev, c);
        double d = __shfl_sync(FULL, v, c);
        if (!(d > 0.0) || !isfinite(d)) {
            ok = false;
            d = 1.0;
        }
        double inv = rsqrt(d);
        // older terms of column c+1 (columns <= c-1 are already in shared memory) overlap the rsqrt latency
        double p0 = 0.0, p1 = 0.0;
        if (c + 1 < 32) {
            const double *rr = S + lane * TSTR, *rc = S + (c + 1) * TSTR;
            p0 = rr[c + 1];
            int q = 0;
#pragma unroll 4
            for (; q + 1 < c; q += 2) {
                p0 -= rr[q] * rc[q];
                p1 -= rr[q + 1] * rc[q + 1];
            }
            if (q < c) p0 -= rr[q] * rc[q];
        }
        double l = lane == c ? d * inv : (lane > c ? v * inv : 0.0);
        S[lane * TSTR + c] = l;
        if (lane == c) sinv[c] = inv;
        lprev = l;
        part = p0 + p1;
        __syncwarp();
    }
    return ok;
}

// In-place X <- X L^-T for the tile X (shared memory, row stride TSTR), one row per lane; L and 1/diag in shared memory.
SDV_DEV void trsm32_smem(double *X, const double *sL, const double *sinv, int lane) {
    double *xr = X + lane * TSTR;
    double xprev = 0.0;
    double part = xr[0];
#pragma unroll 1
    for (int c = 0; c < 32; c++) {
        double v = part;
        if (c > 0) v -= xprev * sL[c * TSTR + c - 1];
        double x = v * sinv[c];
        // older terms of column c+1
        double p0 = 0.0, p1 = 0.0;
        if (c + 1 < 32) {
            const double *rc = sL + (c + 1) * TSTR;
            p0 = xr[c + 1];
            int q = 0;
#pragma unroll 4
            for (; q + 1 < c; q += 2) {
                p0 -= xr[q] * rc[q];
                p1 -= xr[q + 1] * rc[q + 1];
            }
            if (q < c) p0 -= xr[q] * rc[q];
        }
        xr[c] = x;
        xprev = x;
        part = p0 + p1;
    }
}

// Register-resident variants (one tile row per lane, fully unrolled: fast issue, large code).
SDV_DEV bool chol32_reg(double (&a)[32], int lane, double *invd) {
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 32; c++) {
        double dcc = __shfl_sync(FULL, a[c], c);
        if (!(dcc > 0.0) || !isfinite(dcc)) {
            ok = false;
            dcc = 1.0;
        }
        double inv = rsqrt(dcc);
        double l = a[c] * inv;
        if (lane == c) {
            l = dcc * inv;
            *invd = inv;
        }
        a[c] = lane >= c ? l : 0.0;
#pragma unroll
        for (int c2 = c + 1; c2 < 32; c2++) {
            double lc2 = __shfl_sync(FULL, l, c2); // L[c2][c]
            a[c2] -= l * lc2;
        }
    }
    return ok;
}
SDV_DEV void trsm32_reg(double (&a)[32], const double *sL, const double *sinv) {
#pragma unroll
    for (int c = 0; c < 32; c++) {
        double x = a[c] * sinv[c];
        a[c] = x;
#pragma unroll
        for (int c2 = c + 1; c2 < 32; c2++) a[c2] -= x * sL[c2 * TSTR + c];
    }
}

SDV_DEV void load_row32(const double *src, double (&a)[32]) { // 256 contiguous bytes per lane, L2 path (written by other CTAs)
#pragma unroll
    for (int q = 0; q < 16; q++) {
        double2 v = __ldcg(reinterpret_cast<const double2 *>(src) + q);
        a[2 * q] = v.x;
        a[2 * q + 1] = v.y;
    }
}
SDV_DEV void store_row32(double *dst, const double (&a)[32]) {
#pragma unroll
    for (int q = 0; q < 16; q++) reinterpret_cast<double2 *>(dst)[q] = make_double2(a[2 * q], a[2 * q + 1]);
}

// C(32x32, global, leading dim ld) -= A(32x32, shared, stride TSTR) * B(32x32)^T, B from shared (stride TSTR) or global (ld)
SDV_DEV void tile_update_dmma(double *Cg, int ld, const double *As, const double *Bp, int bstride, bool b_global, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double c[4][4][2];
#pragma unroll
    for (int ib = 0; ib < 4; ib++)
#pragma unroll
        for (int jb = 0; jb < 4; jb++) {
            double2 v = *reinterpret_cast<const double2 *>(Cg + (size_t)(ib * 8 + g) * ld + jb * 8 + 2 * t);
            c[ib][jb][0] = v.x;
            c[ib][jb][1] = v.y;
        }
    double b[4][8];
#pragma unroll
    for (int jb = 0; jb < 4; jb++)
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
            const double *p = Bp + (size_t)(jb * 8 + g) * bstride + kk * 4 + t;
            b[jb][kk] = b_global ? __ldcg(p) : *p;
        }
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
#pragma unroll
        for (int ib = 0; ib < 4; ib++) {
            double a = -As[(ib * 8 + g) * TSTR + kk * 4 + t];
#pragma unroll
            for (int jb = 0; jb < 4; jb++) dmma(c[ib][jb][0], c[ib][jb][1], a, b[jb][kk]);
        }
    }
#pragma unroll
    for (int ib = 0; ib < 4; ib++)
#pragma unroll
        for (int jb = 0; jb < 4; jb++)
            *reinterpret_cast<double2 *>(Cg + (size_t)(ib * 8 + g) * ld + jb * 8 + 2 * t) = make_double2(c[ib][jb][0], c[ib][jb][1]);
}

// C(32x32, global) -= A(32x32) * B(32x32)^T with both operands read from global memory (L2) — used when trailing tiles are
// spread over every update warp of the cluster and a CTA no longer owns whole tile rows.
SDV_DEV void tile_update_dmma_gg(double *Cg, const double *Ag, const double *Bg, int ld, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double b[4][8], a[4][8];
#pragma unroll
    for (int jb = 0; jb < 4; jb++)
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
            b[jb][kk] = __ldcg(Bg + (size_t)(jb * 8 + g) * ld + kk * 4 + t);
            a[jb][kk] = __ldcg(Ag + (size_t)(jb * 8 + g) * ld + kk * 4 + t);
        }
    double c[4][4][2];
#pragma unroll
    for (int ib = 0; ib < 4; ib++)
#pragma unroll
        for (int jb = 0; jb < 4; jb++) {
            double2 v = *reinterpret_cast<const double2 *>(Cg + (size_t)(ib * 8 + g) * ld + jb * 8 + 2 * t);
            c[ib][jb][0] = v.x;
            c[ib][jb][1] = v.y;
        }
#pragma unroll
    for (int kk = 0; kk < 8; kk++)
#pragma unroll
        for (int ib = 0; ib < 4; ib++) {
            const double na = -a[ib][kk];
#pragma unroll
            for (int jb = 0; jb < 4; jb++) dmma(c[ib][jb][0], c[ib][jb][1], na, b[jb][kk]);
        }
#pragma unroll
    for (int ib = 0; ib < 4; ib++)
#pragma unroll
        for (int jb = 0; jb < 4; jb++)
            *reinterpret_cast<double2 *>(Cg + (size_t)(ib * 8 + g) * ld + jb * 8 + 2 * t) = make_double2(c[ib][jb][0], c[ib][jb][1]);
}

// A  : (n_pad + 32) x ld reduced system, lower triangle + right-hand side in row n_pad; overwritten by the trailing updates.
// Lo : receives L (and y^T = (L^-1 g)^T in row n_pad).  dinv : [n_pad] reciprocal diagonal of L.
// partial : [T][CC_MAX][32] scratch for the backward solve.
//
// Synchronisation per panel k (mbarriers in every CTA's shared memory, remote arrives through DSMEM):
//   b1[k&1] "L_kk is published"      : count 1 — the owner's warp 0 arrives on every CTA right after factoring the tile
//   b2      "panel column published" : count CS — each CTA arrives (on every CTA) after its own triangular solves; the owner of
//                                      tile row k+1 then updates + factors tile (k+1,k+1) BEFORE waiting (look-ahead), the others
//                                      wait and run their trailing updates, which depend only on their own rows + the panel column.
#define SDV_TICK(slot) do { tn = clock64(); tp[slot] += tn - tc; tc = tn; } while (0)

// Second half shared by the factorisation kernels: failure vote, backward solve L^T z = y (left-looking over tile columns,
// partial products reduced through global scratch + hardware cluster barrier), reduced-parameter update, model-decrease
// terms and the candidate frame-camera table on cluster rank 0.  Called by every thread of the cluster.
SDV_DEV void chol_backward_and_update(const DevProblem &P, const LinBuf &B0, const LinBuf &B1, LMState *st, Accum *acc, double *Lo,
                                      double *dinv, double *partial, const double *damp_p, const double *graw_p, double *dxp, double *sK,
                                      double *sinv, double *xs, double *wsum, bool fail, double *prof, long long *tp, long long tc) {
    const int ld = P.ld, T = P.n_pad / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = (int)cluster_rank(), CS = (int)cluster_size();
    constexpr int NW = CCT / 32;
    long long tn;
    if (fail) acc->chol_fail = 1;
    cluster_sync_all();
    const bool bad = __ldcg(&acc->chol_fail) != 0 || __ldcg(&acc->schur_fail) != 0; // same answer on every CTA
    if (bad) {
        if (rank == 0 && threadIdx.x == 0) {
            st->step_valid = 0;
            st->model_cost_change = 0.0;
        }
        return;
    }
    // ---------------- backward solve L^T z = y, left-looking over tile columns
    const double *y = Lo + (size_t)(T * 32) * ld;
    tc = clock64();
    for (int k = T - 1; k >= 0; k--) {
        const int owner = k % CS;
        __syncthreads(); // xs of the previous step visible to every warp
        const int first = k + 1 + ((rank - (k + 1)) % CS + CS) % CS;
        double s = 0.0;
        for (int i = first + warp * CS; i < T; i += NW * CS) {
            const double *Lt = Lo + (size_t)(i * 32) * ld + k * 32 + lane;
            const double *xi = xs + (size_t)(i / CS) * 32;
            double v[32];
#pragma unroll
            for (int r = 0; r < 32; r++) v[r] = __ldcg(Lt + (size_t)r * ld); // 32 independent L2 loads in flight
#pragma unroll
            for (int r = 0; r < 32; r++) s += v[r] * xi[r];
        }
        wsum[warp * 32 + lane] = s;
        if (rank == owner) { // the diagonal tile for the triangular solve
            for (int e = threadIdx.x; e < 512; e += CCT) {
                int r = e >> 4, q = e & 15;
                double2 v = __ldcg(reinterpret_cast<const double2 *>(Lo + (size_t)(k * 32 + r) * ld + k * 32) + q);
                sK[r * TSTR + 2 * q] = v.x;
                sK[r * TSTR + 2 * q + 1] = v.y;
            }
            if (threadIdx.x < 32) sinv[threadIdx.x] = __ldcg(dinv + k * 32 + threadIdx.x);
        }
        __syncthreads();
        if (warp == 0) {
            double tot = 0;
#pragma unroll
            for (int w2 = 0; w2 < NW; w2++) tot += wsum[w2 * 32 + lane];
            partial[((size_t)k * CC_MAX + rank) * 32 + lane] = tot;
        }
        cluster_sync_all();
        if (rank == owner && warp == 0) {
            double yy = __ldcg(y + k * 32 + lane);
            for (int r = 0; r < CS; r++) yy -= __ldcg(partial + ((size_t)k * CC_MAX + r) * 32 + lane);
            double x = 0;
#pragma unroll
            for (int c = 31; c >= 0; c--) {
                double xc = __shfl_sync(FULL, yy, c) * sinv[c];
                if (lane == c) x = xc;
                if (lane < c) yy -= sK[c * TSTR + lane] * xc;
            }
            xs[(size_t)(k / CS) * 32 + lane] = x;
            dxp[k * 32 + lane] = -x; // S delta = -g
        }
    }
    cluster_sync_all();
    SDV_TICK(6);
    if (prof && threadIdx.x == 0)
        for (int q = 0; q < 7; q++) prof[rank * 8 + q] = (double)tp[q];
    if (rank != 0) return;
    // ---------------- reduced-parameter update, model-decrease terms, candidate frame-camera table (cluster rank 0)
    const LinBuf &Bx = st->cur ? B1 : B0;
    const LinBuf &Bc = st->cur ? B0 : B1;
    const int n = P.n;
    double gd = 0, dd = 0, sn = 0, cn = 0;
    for (int i = threadIdx.x; i < P.n_pad; i += CCT) {
        double d = i < n ? __ldcg(dxp + i) : 0.0;
        if (i >= n) dxp[i] = 0.0;
        double xc = Bx.xp[i] + d;
        Bc.xp[i] = i < n ? xc : 0.0;
        if (i < n) {
            gd += graw_p[i] * d;
            dd += damp_p[i] * d * d;
            sn += d * d;
            cn += xc * xc;
        }
    }
    gd = warp_sum(gd);
    dd = warp_sum(dd);
    sn = warp_sum(sn);
    cn = warp_sum(cn);
    if (lane == 0 && P.rank == 0) {
        atomicAdd(&acc->model_gd, gd);
        atomicAdd(&acc->model_dd, dd);
        atomicAdd(&acc->step_norm2, sn);
        atomicAdd(&acc->cand_norm2, cn);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P.F * P.C; i += CCT) compute_fct_row(P, Bc.xp, i / P.C, i % P.C, Bc.fct + (size_t)i * FCT_ROW);
    if (threadIdx.x == 0) st->step_valid = 1;
}

template <bool REG>
__global__ void __launch_bounds__(CCT, 1) k_chol_cluster(DevProblem P, LinBuf B0, LinBuf B1, LMState *st, Accum *acc, double *A, double *Lo,
                                                         double *dinv, double *partial, const double *damp_p, const double *graw_p, double *dxp,
                                                         int max_rows, double *prof) {
    if (st->status != 0) return; // uniform over the cluster
    extern __shared__ __align__(16) double csm[];
    __shared__ uint64_t b1[2], b2;
    double *sK = csm;                               // [32][TSTR] diagonal tile L_kk
    double *sinv = sK + 32 * TSTR;                  // [32]
    double *sRow = sinv + 32;                       // [max_rows][32][TSTR] own tiles of the current panel column
    double *xs = sRow + (size_t)max_rows * 32 * TSTR; // [max_rows][32] solved blocks of own tile rows
    double *wsum = xs + (size_t)max_rows * 32;      // [8][32]
    const int ld = P.ld, T = P.n_pad / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = (int)cluster_rank(), CS = (int)cluster_size();
    constexpr int NW = CCT / 32;
    bool fail = false;
    // optional phase timing (cycles, summed over panels): [rank][0 wait L_kk, 1 load L_kk, 2 trsm, 3 look-ahead, 4 wait column, 5 update, 6 backward]
    long long tp[7] = {0, 0, 0, 0, 0, 0, 0}, tc = clock64(), tn;

    if (threadIdx.x == 0) {
        mbar_init(&b1[0], 1);
        mbar_init(&b1[1], 1);
        mbar_init(&b2, CS);
    }
    __syncthreads();
    cluster_sync_all(); // every CTA's barriers are initialised before anybody arrives remotely

    // k = -1 is the prologue: only the "look-ahead" part runs and factors tile (0,0)
    for (int k = -1; k < T; k++) {
        const int owner = (k + CS) % CS, next_owner = (k + 1) % CS;
        int first = 0, nown = 0;
        if (k >= 0) {
            tc = clock64();
            mbar_wait_cluster(&b1[k & 1], (unsigned)((k >> 1) & 1)); // L_kk and its reciprocal diagonal are visible
            SDV_TICK(0);
            if (rank != owner) {
                for (int e = threadIdx.x; e < 512; e += CCT) {
                    int r = e >> 4, q = e & 15;
                    double2 v = __ldcg(reinterpret_cast<const double2 *>(Lo + (size_t)(k * 32 + r) * ld + k * 32) + q);
                    sK[r * TSTR + 2 * q] = v.x;
                    sK[r * TSTR + 2 * q + 1] = v.y;
                }
                if (threadIdx.x < 32) sinv[threadIdx.x] = __ldcg(dinv + k * 32 + threadIdx.x);
            }
            __syncthreads();
            SDV_TICK(1);
            // ---------------- own tiles of the panel column (rows i > k, i % CS == rank; i == T is the right-hand side)
            first = k + 1 + ((rank - (k + 1)) % CS + CS) % CS;
            nown = first <= T ? (T - first) / CS + 1 : 0;
            for (int s = warp; s < nown; s += NW) {
                int i = first + s * CS;
                double *X = sRow + (size_t)s * 32 * TSTR;
                {
                    double a[32];
                    load_row32(A + (size_t)(i * 32 + lane) * ld + k * 32, a);
                    if (REG) trsm32_reg(a, sK, sinv);
#pragma unroll
                    for (int c = 0; c < 32; c++) X[lane * TSTR + c] = a[c];
                }
                if (!REG) trsm32_smem(X, sK, sinv, lane);
                __syncwarp();
                // publish L_ik, 16 lanes per row -> coalesced 256-byte row segments
                for (int e = lane; e < 512; e += 32) {
                    int r = e >> 4, q = e & 15;
                    reinterpret_cast<double2 *>(Lo + (size_t)(i * 32 + r) * ld + k * 32)[q] = make_double2(X[r * TSTR + 2 * q], X[r * TSTR + 2 * q + 1]);
                }
            }
            __syncthreads();
            SDV_TICK(2);
            if (threadIdx.x == 32) mbar_arrive_all(&b2, CS); // this CTA's part of the panel column is published
        }
        const bool lookahead = (k + 1 < T) && rank == next_owner; // then slot 0 is tile row k+1, whose only tile is the diagonal
        if (lookahead && warp == 0) {
            // tile (k+1,k+1) -= L_{k+1,k} L_{k+1,k}^T with this lane's row in registers, then factor it.
            // sK still holds L_kk, which this CTA no longer needs (its triangular solves are done), so it is the work tile.
            {
                double a[32];
                load_row32(A + (size_t)((k + 1) * 32 + lane) * ld + (k + 1) * 32, a);
                if (k >= 0) {
#pragma unroll 1
                    for (int q = 0; q < 32; q++) {
                        double lq = sRow[(size_t)lane * TSTR + q];
#pragma unroll
                        for (int c = 0; c < 32; c++) a[c] -= lq * sRow[(size_t)c * TSTR + q];
                    }
                }
                if (REG) {
                    double inv = 1.0;
                    if (!chol32_reg(a, lane, &inv)) fail = true;
                    sinv[lane] = inv;
                }
#pragma unroll
                for (int c = 0; c < 32; c++) sK[lane * TSTR + c] = a[c];
            }
            __syncwarp();
            if (!REG && !chol32_smem(sK, sinv, lane)) fail = true;
            __syncwarp();
            dinv[(k + 1) * 32 + lane] = sinv[lane];
            for (int e = lane; e < 512; e += 32) {
                int r = e >> 4, q = e & 15;
                reinterpret_cast<double2 *>(Lo + (size_t)((k + 1) * 32 + r) * ld + (k + 1) * 32)[q] =
                    make_double2(sK[r * TSTR + 2 * q], sK[r * TSTR + 2 * q + 1]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_all(&b1[(k + 1) & 1], CS);
        }
        if (k < 0) continue;
        SDV_TICK(3);
        mbar_wait_cluster(&b2, (unsigned)(k & 1)); // every L_jk of this panel column is visible
        SDV_TICK(4);
        // ---------------- trailing update of own tile rows, tiles (i, j) with k < j <= min(i, T-1)
        int base = 0; // flat tile index of the first tile of row slot s
        for (int s = lookahead ? 1 : 0; s < nown; s++) {
            int i = first + s * CS;
            int nj = min(i, T - 1) - k;
            int q0 = ((warp - base) % NW + NW) % NW;
            for (int q = q0; q < nj; q += NW) {
                int j = k + 1 + q;
                bool j_own = (j % CS) == rank; // then L_jk sits in sRow as well
                const double *Bp = j_own ? sRow + (size_t)((j - first) / CS) * 32 * TSTR : Lo + (size_t)(j * 32) * ld + k * 32;
                tile_update_dmma(A + (size_t)(i * 32) * ld + j * 32, ld, sRow + (size_t)s * 32 * TSTR, Bp, j_own ? TSTR : ld, !j_own, lane);
            }
            base += nj;
        }
        __syncthreads(); // sRow / sK are rewritten in the next panel
        SDV_TICK(5);
    }
    chol_backward_and_update(P, B0, B1, st, acc, Lo, dinv, partial, damp_p, graw_p, dxp, sK, sinv, xs, wsum, fail, prof, tp, tc);
}

// Hybrid register Cholesky: the pivot column is broadcast through a 32-double shared-memory buffer (one store + broadcast
// loads per column) instead of 2 SHFL per element — SHFL issue was the throughput limit of chol32_reg (8.3k cycles/tile);
// only the element of the NEXT pivot column, which sits on the dependency chain, still goes through a shuffle.
SDV_DEV bool chol32_hyb(double (&a)[32], int lane, double *invd, double *colbuf /* shared, [2][32] */) {
    // Software-pipelined: the next pivot (shuffle + rsqrt, the dependency chain) is issued BEFORE the bulk of the current
    // column's rank-1 update, which then fills the latency of the reciprocal square root.
    bool ok = true;
    double d = __shfl_sync(FULL, a[0], 0);
    if (!(d > 0.0) || !isfinite(d)) {
        ok = false;
        d = 1.0;
    }
    double inv = rsqrt(d);
#pragma unroll
    for (int c = 0; c < 32; c++) {
        double l = a[c] * inv;
        if (lane == c) {
            l = d * inv;
            *invd = inv;
        }
        if (lane < c) l = 0.0;
        a[c] = l;
        if (c + 1 < 32) {
            a[c + 1] -= l * __shfl_sync(FULL, l, c + 1); // the next pivot column first
            d = __shfl_sync(FULL, a[c + 1], c + 1);
            if (!(d > 0.0) || !isfinite(d)) {
                ok = false;
                d = 1.0;
            }
            inv = rsqrt(d);                               // in flight during the updates below
            if (c + 2 < 32) {
                double *cb = colbuf + (c & 1) * 32;
                cb[lane] = l;
                __syncwarp();
#pragma unroll
                for (int c2 = c + 2; c2 < 32; c2++) a[c2] -= l * cb[c2];
            }
        }
    }
    return ok;
}

SDV_DEV void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// cluster-scope release: the panel warps read the updated tiles back through L2 (__ldcg)
SDV_DEV void mbar_arrive_local(uint64_t *bar) { asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
SDV_DEV void mbar_wait_local(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITL_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEL_%=;\n"
        "bra WAITL_%=;\n"
        "DONEL_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// DSMEM store of one double into the shared memory of CTA `target_rank` of the cluster
SDV_DEV void dsmem_store(double *local_addr, unsigned target_rank, double v) {
    unsigned local = smem_u32(local_addr), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(target_rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
}

// Backward solve, version 2 (replaces the global-scratch + cluster-barrier scheme of chol_backward_and_update):
//  * every CTA first inverts the diagonal tiles it owns (X = L_kk^-T by a register TRSM on the identity, all CTAs in
//    parallel), so that a step is a 32x32 matrix-vector product instead of a 32-step substitution;
//  * the partial products sum_{own i > k} L_ik^T x_i are sent straight into the owner's shared memory (DSMEM stores)
//    followed by one remote mbarrier arrive; only the owner of step k waits, the others run ahead.
// tiles : shared, >= ceil(T/CS) tiles of [32][TSTR] (the sRow region); red : shared [CC_MAX][32]; bk : mbarrier, count CS.
SDV_DEV void chol_backward_v2(const DevProblem &P, const LinBuf &B0, const LinBuf &B1, LMState *st, Accum *acc, double *Lo, double *dinv,
                              const double *damp_p, const double *graw_p, double *dxp, double *tiles, double *tinv, double *xs, double *wsum,
                              double *red, uint64_t *bk, bool fail, double *prof, long long *tp, long long tc) {
    const int ld = P.ld, T = P.n_pad / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = (int)cluster_rank(), CS = (int)cluster_size();
    constexpr int NW = CCT / 32;
    long long tn;
    if (threadIdx.x == 0) mbar_init(bk, CS);
    if (fail) acc->chol_fail = 1;
    __syncthreads();
    cluster_sync_all(); // L complete and visible, failure flag visible, bk initialised everywhere
    const bool bad = __ldcg(&acc->chol_fail) != 0 || __ldcg(&acc->schur_fail) != 0; // same answer on every CTA
    if (bad) {
        if (rank == 0 && threadIdx.x == 0) {
            st->step_valid = 0;
            st->model_cost_change = 0.0;
        }
        return;
    }
    tc = clock64();
    // ---- inverse of the own diagonal tiles: tile q of this CTA is tile row kk = rank + q*CS
    const int ndiag = rank < T ? (T - 1 - rank) / CS + 1 : 0;
    for (int q = warp; q < ndiag; q += NW) {
        const int kk = rank + q * CS;
        double *Lt = tiles + (size_t)q * 32 * TSTR, *iv = tinv + q * 32;
        for (int e = lane; e < 512; e += 32) {
            int r = e >> 4, c2 = e & 15;
            double2 v = __ldcg(reinterpret_cast<const double2 *>(Lo + (size_t)(kk * 32 + r) * ld + kk * 32) + c2);
            Lt[r * TSTR + 2 * c2] = v.x;
            Lt[r * TSTR + 2 * c2 + 1] = v.y;
        }
        iv[lane] = __ldcg(dinv + kk * 32 + lane);
        __syncwarp();
        double a[32];
#pragma unroll
        for (int c = 0; c < 32; c++) a[c] = (c == lane) ? 1.0 : 0.0;
        trsm32_reg(a, Lt, iv); // row `lane` of I * L^-T
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; c++) Lt[lane * TSTR + c] = a[c];
    }
    __syncthreads();
    const double *y = Lo + (size_t)(T * 32) * ld;
    for (int k = T - 1; k >= 0; k--) {
        const int owner = k % CS;
        // partial product of this CTA for step k over its own rows i > k (their x_i were solved by this CTA)
        const int first = k + 1 + ((rank - (k + 1)) % CS + CS) % CS;
        double s = 0.0;
        for (int i = first + warp * CS; i < T; i += NW * CS) {
            const double *Lc = Lo + (size_t)(i * 32) * ld + k * 32 + lane;
            const double *xi = xs + (size_t)(i / CS) * 32;
            double v[32];
#pragma unroll
            for (int r = 0; r < 32; r++) v[r] = __ldcg(Lc + (size_t)r * ld);
#pragma unroll
            for (int r = 0; r < 32; r++) s += v[r] * xi[r];
        }
        wsum[warp * 32 + lane] = s;
        __syncthreads();
        if (warp == 0) {
            double tot = 0;
#pragma unroll
            for (int w2 = 0; w2 < NW; w2++) tot += wsum[w2 * 32 + lane];
            dsmem_store(red + rank * 32 + lane, (unsigned)owner, tot);
            __syncwarp();
            if (lane == 0) mbar_remote_arrive(bk, (unsigned)owner);
            if (rank == owner) {
                mbar_wait_cluster(bk, (unsigned)(((T - 1 - k) / CS) & 1));
                double yy = __ldcg(y + k * 32 + lane);
#pragma unroll
                for (int r = 0; r < CC_MAX; r++)
                    if (r < CS) yy -= red[r * 32 + lane];
                wsum[lane] = yy; // y' for the matrix-vector product
                __syncwarp();
                const double *X = tiles + (size_t)(k / CS) * 32 * TSTR + lane * TSTR; // row `lane` of L_kk^-T (upper triangular)
                double x0 = 0, x1 = 0;
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    x0 += X[c] * wsum[c];
                    x1 += X[c + 1] * wsum[c + 1];
                }
                const double x = x0 + x1;
                xs[(size_t)(k / CS) * 32 + lane] = x;
                dxp[k * 32 + lane] = -x; // S delta = -g
            }
        }
        __syncthreads(); // wsum reuse; xs of this step visible to every warp of the owner
    }
    cluster_sync_all();
    SDV_TICK(6);
    if (prof && threadIdx.x == 0)
        for (int q = 0; q < 7; q++) prof[rank * 8 + q] = (double)tp[q];
    if (rank != 0) return;
    // ---------------- reduced-parameter update, model-decrease terms, candidate frame-camera table (cluster rank 0)
    const LinBuf &Bx = st->cur ? B1 : B0;
    const LinBuf &Bc = st->cur ? B0 : B1;
    const int n = P.n;
    double gd = 0, dd = 0, sn = 0, cn = 0;
    for (int i = threadIdx.x; i < P.n_pad; i += CCT) {
        double d = i < n ? __ldcg(dxp + i) : 0.0;
        if (i >= n) dxp[i] = 0.0;
        double xc = Bx.xp[i] + d;
        Bc.xp[i] = i < n ? xc : 0.0;
        if (i < n) {
            gd += graw_p[i] * d;
            dd += damp_p[i] * d * d;
            sn += d * d;
            cn += xc * xc;
        }
    }
    gd = warp_sum(gd);
    dd = warp_sum(dd);
    sn = warp_sum(sn);
    cn = warp_sum(cn);
    if (lane == 0 && P.rank == 0) {
        atomicAdd(&acc->model_gd, gd);
        atomicAdd(&acc->model_dd, dd);
        atomicAdd(&acc->step_norm2, sn);
        atomicAdd(&acc->cand_norm2, cn);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P.F * P.C; i += CCT) compute_fct_row(P, Bc.xp, i / P.C, i % P.C, Bc.fct + (size_t)i * FCT_ROW);
    if (threadIdx.x == 0) st->step_valid = 1;
}

constexpr int NPW = 2;             // panel warps: triangular solves of the panel column + look-ahead factorisation
constexpr int NUW = CCT / 32 - NPW; // update warps: trailing tile updates (FP64 tensor-core MMA)

// Warp-specialised variant of the cluster factorisation.  The per-panel critical chain
//     L_kk published -> load L_kk -> TRSM of tile row k+1 -> update + factor tile (k+1,k+1) -> L_(k+1)(k+1) published
// runs on the two PANEL warps of the CTAs involved and never waits for the bulk of the trailing update, which the six
// UPDATE warps of every CTA perform one panel behind.  Each trailing tile is always updated by the same warp (no
// read-modify-write races across panels); sRow is double-buffered by panel parity; dependencies:
//   b1[k&1] (cluster) : L_kk published                           owner panel warp 0  -> every panel warp
//   b2      (cluster) : panel column k published                 every CTA           -> every update warp
//   udone   (CTA)     : tiles of columns <= k+2 updated for panel k (and all of panel k-1), so tile column k+1 and the
//                       diagonal tile (k+2,k+2) are final          update warps        -> panel warps of the same CTA
// HYB selects the shared-memory-broadcast Cholesky of the diagonal tile.
template <bool HYB>
__global__ void __launch_bounds__(CCT, 1) k_chol_ws(DevProblem P, LinBuf B0, LinBuf B1, LMState *st, Accum *acc, double *A, double *Lo, double *dinv,
                                                    double *partial, const double *damp_p, const double *graw_p, double *dxp, int max_rows,
                                                    double *prof) {
    if (st->status != 0) return; // uniform over the cluster
    extern __shared__ __align__(16) double csm[];
    __shared__ uint64_t b1[2], b2, udone, bk;
    double *sK = csm;                                   // [32][TSTR] diagonal tile L_kk
    double *sinv = sK + 32 * TSTR;                      // [32]
    double *colbuf = sinv + 32;                         // [2][32]
    double *sRow = colbuf + 64;                         // [2][max_rows][32][TSTR] own tiles of the panel column, by panel parity
    double *xs = sRow + (size_t)2 * max_rows * 32 * TSTR; // [max_rows][32]
    double *wsum = xs + (size_t)max_rows * 32;          // [8][32]
    double *red = wsum + 8 * 32;                        // [CC_MAX][32] partial products received from the other CTAs
    double *tinv = red + CC_MAX * 32;                   // [max_rows][32] reciprocal diagonals of the own diagonal tiles
    const int ld = P.ld, T = P.n_pad / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = (int)cluster_rank(), CS = (int)cluster_size();
    const size_t row_buf = (size_t)max_rows * 32 * TSTR;
    bool fail = false;
    long long tp[7] = {0, 0, 0, 0, 0, 0, 0}, tc = clock64(), tn;

    if (threadIdx.x == 0) {
        mbar_init(&b1[0], 1);
        mbar_init(&b1[1], 1);
        mbar_init(&b2, CS);
        mbar_init(&udone, NUW);
    }
    __syncthreads();
    cluster_sync_all(); // every CTA's barriers are initialised before anybody arrives remotely

    // Warps 0 and 4 share SM sub-partition 0: making them the panel warps keeps the critical look-ahead factorisation off the
    // sub-partitions that run the tensor-core updates.
    const bool is_panel = (warp & 3) == 0;
    const int pw = warp >> 2;                  // panel warp index 0 / 1
    if (is_panel) {
        // =================================================== panel warps
        for (int k = -1; k < T; k++) {
            const int owner = (k + CS) % CS, next_owner = (k + 1) % CS;
            double *rows = sRow + (size_t)(k & 1) * row_buf;
            if (k >= 0) {
                tc = clock64();
                mbar_wait_cluster(&b1[k & 1], (unsigned)((k >> 1) & 1)); // L_kk and its reciprocal diagonal are visible
                SDV_TICK(0);
                if (rank != owner && pw == 0) {
                    for (int e = lane; e < 512; e += 32) {
                        int r = e >> 4, q = e & 15;
                        double2 v = __ldcg(reinterpret_cast<const double2 *>(Lo + (size_t)(k * 32 + r) * ld + k * 32) + q);
                        sK[r * TSTR + 2 * q] = v.x;
                        sK[r * TSTR + 2 * q + 1] = v.y;
                    }
                    sinv[lane] = __ldcg(dinv + k * 32 + lane);
                }
                if (k >= 1) mbar_wait_local(&udone, (unsigned)((k - 1) & 1)); // own tile column k is final, sRow[k&1] is free
                named_bar_sync(1, NPW * 32);
                SDV_TICK(1);
                const int first = k + 1 + ((rank - (k + 1)) % CS + CS) % CS;
                const int nown = first <= T ? (T - first) / CS + 1 : 0;
                for (int s = pw; s < nown; s += NPW) {
                    int i = first + s * CS;
                    double *X = rows + (size_t)s * 32 * TSTR;
                    {
                        double a[32];
                        load_row32(A + (size_t)(i * 32 + lane) * ld + k * 32, a);
                        trsm32_reg(a, sK, sinv);
#pragma unroll
                        for (int c = 0; c < 32; c++) X[lane * TSTR + c] = a[c];
                    }
                    __syncwarp();
                    for (int e = lane; e < 512; e += 32) { // publish L_ik, coalesced 256-byte row segments
                        int r = e >> 4, q = e & 15;
                        reinterpret_cast<double2 *>(Lo + (size_t)(i * 32 + r) * ld + k * 32)[q] = make_double2(X[r * TSTR + 2 * q], X[r * TSTR + 2 * q + 1]);
                    }
                }
                named_bar_sync(1, NPW * 32);
                SDV_TICK(2);
                if (pw == 0) mbar_arrive_all_warp(&b2, CS, lane); // this CTA's part of the panel column is published
            }
            if ((k + 1 < T) && rank == next_owner && pw == 0) {
                // look-ahead: tile (k+1,k+1) -= L_{k+1,k} L_{k+1,k}^T (row slot 0 of this panel), then factor it.
                // sK held L_kk, which this CTA no longer needs: it becomes L_{k+1,k+1} for the next panel.
                double a[32];
                load_row32(A + (size_t)((k + 1) * 32 + lane) * ld + (k + 1) * 32, a);
                if (k >= 0) {
#pragma unroll 1
                    for (int q = 0; q < 32; q++) {
                        double lq = rows[(size_t)lane * TSTR + q];
#pragma unroll
                        for (int c = 0; c < 32; c++) a[c] -= lq * rows[(size_t)c * TSTR + q];
                    }
                }
                double inv = 1.0;
                bool ok = HYB ? chol32_hyb(a, lane, &inv, colbuf) : chol32_reg(a, lane, &inv);
                if (!ok) fail = true;
#pragma unroll
                for (int c = 0; c < 32; c++) sK[lane * TSTR + c] = a[c];
                sinv[lane] = inv;
                dinv[(k + 1) * 32 + lane] = inv;
                __syncwarp();
                for (int e = lane; e < 512; e += 32) {
                    int r = e >> 4, q = e & 15;
                    reinterpret_cast<double2 *>(Lo + (size_t)((k + 1) * 32 + r) * ld + (k + 1) * 32)[q] =
                        make_double2(sK[r * TSTR + 2 * q], sK[r * TSTR + 2 * q + 1]);
                }
                mbar_arrive_all_warp(&b1[(k + 1) & 1], CS, lane);
                SDV_TICK(3);
            }
        }
    } else {
        // =================================================== update warps
        const int uw = warp - 1 - (warp >> 2); // warps 1,2,3,5,6,7 -> 0..5
        for (int k = 0; k < T; k++) {
            const double *rows = sRow + (size_t)(k & 1) * row_buf;
            const int first = k + 1 + ((rank - (k + 1)) % CS + CS) % CS;
            const int nown = first <= T ? (T - first) / CS + 1 : 0;
            const bool skip0 = (k + 1 < T) && rank == (k + 1) % CS; // tile (k+1,k+1) belongs to the look-ahead
            tc = clock64();
            mbar_wait_cluster(&b2, (unsigned)(k & 1)); // every L_jk of this panel column is visible (incl. our own rows in sRow)
            SDV_TICK(4);
            bool arrived = false;
            for (int j = k + 1; j < T; j++) { // tile columns in the order the panel warps will need them
                if (!arrived && j > k + 2) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_local(&udone);
                    arrived = true;
                }
                for (int s = skip0 ? 1 : 0; s < nown; s++) {
                    const int i = first + s * CS;
                    if (i < j) continue;
                    if ((j + 3 * (i / CS)) % NUW != uw) continue; // a tile is always updated by the same warp
                    const bool j_own = (j % CS) == rank;      // then L_jk sits in sRow as well
                    const double *Bp = j_own ? rows + (size_t)((j - first) / CS) * 32 * TSTR : Lo + (size_t)(j * 32) * ld + k * 32;
                    tile_update_dmma(A + (size_t)(i * 32) * ld + j * 32, ld, rows + (size_t)s * 32 * TSTR, Bp, j_own ? TSTR : ld, !j_own, lane);
                }
            }
            if (!arrived) {
                __syncwarp();
                if (lane == 0) mbar_arrive_local(&udone);
            }
            SDV_TICK(5);
        }
    }
    __syncthreads();
    if (partial) // legacy backward solve (global scratch + hardware cluster barrier), kept for A/B timing
        chol_backward_and_update(P, B0, B1, st, acc, Lo, dinv, partial, damp_p, graw_p, dxp, sK, sinv, xs, wsum, fail, prof, tp, tc);
    else
        chol_backward_v2(P, B0, B1, st, acc, Lo, dinv, damp_p, graw_p, dxp, sRow, tinv, xs, wsum, red, &bk, fail, prof, tp, tc);
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-level role split.  In k_chol_ws the panel warp shares its SM with six tensor-core update warps; their shared-memory
// and LSU traffic stretches the critical tile factorisation 2x in the early (update-heavy) panels.  Here the first NP CTAs of
// the cluster are PANEL CTAs (L_kk factorisation, all triangular solves of the panel column) and the other CS-NP are UPDATE
// CTAs (trailing tile updates only), so the dependency chain
//     L_kk published -> load L_kk -> TRSM tile row k+1 -> update + factor tile (k+1,k+1) -> published
// runs on SMs that do nothing else.
//   b1[k&1] : count 1      owner panel CTA (k % NP), warp 0          -> every panel CTA      "L_kk published"
//   b2      : count NP     every panel CTA after its TRSMs           -> every update CTA     "panel column k published"
//   udone   : count NU*8   every update warp after its urgent tiles  -> every panel CTA      "tile columns <= k+2 final"
// Tile row i is solved by panel CTA i % NP (warp (i / NP) % 8) and updated by update CTA NP + i % NU; a trailing tile is
// always updated by the same warp.  Backward solve: chol_backward_v2 on all CS CTAs.
template <int NP>
__global__ void __launch_bounds__(CCT, 1) k_chol_roles(DevProblem P, LinBuf B0, LinBuf B1, LMState *st, Accum *acc, double *A, double *Lo, double *dinv,
                                                       const double *damp_p, const double *graw_p, double *dxp, int max_rows, double *prof) {
    if (st->status != 0) return; // uniform over the cluster
    extern __shared__ __align__(16) double csm[];
    __shared__ uint64_t b1[2], b2, udone, bk;
    constexpr int NW = CCT / 32;
    double *sK = csm;                                   // [32][TSTR] diagonal tile L_kk (panel CTAs)
    double *sinv = sK + 32 * TSTR;                      // [32]
    double *colbuf = sinv + 32;                         // [2][32]
    double *sDiag = colbuf + 64;                        // [32][TSTR] prefetched tile (k+1,k+1) (panel CTAs)
    double *sRow = sDiag + 32 * TSTR;                   // [2][max_rows][32][TSTR]: panel CTA: look-ahead operand; backward solve: tile inverses
    double *xs = sRow + (size_t)2 * max_rows * 32 * TSTR;
    double *wsum = xs + (size_t)max_rows * 32;
    double *red = wsum + 8 * 32;
    double *tinv = red + CC_MAX * 32;
    const int ld = P.ld, T = P.n_pad / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = (int)cluster_rank(), CS = (int)cluster_size();
    const int NU = CS - NP;
    const size_t row_buf = (size_t)max_rows * 32 * TSTR;
    bool fail = false;
    long long tp[7] = {0, 0, 0, 0, 0, 0, 0}, tc = clock64(), tn;

    if (threadIdx.x == 0) {
        mbar_init(&b1[0], 1);
        mbar_init(&b1[1], 1);
        mbar_init(&b2, NP);
        mbar_init(&udone, NU * NW);
    }
    __syncthreads();
    cluster_sync_all();

    if (rank < NP) {
        // =================================================== panel CTA
        for (int k = -1; k < T; k++) {
            const int owner = (k + NP) % NP, next_owner = (k + 1) % NP;
            const bool lookahead = (k + 1 < T) && rank == next_owner && warp == 0;
            double *rows = sRow + (size_t)(k & 1) * row_buf;
            double a[32];
            if (k >= 0) {
                // rows i > k with i % NP == rank; slot s handles row first + s*NP; row k+1 (if ours) is slot 0 -> warp 0
                const int first = k + 1 + ((rank - (k + 1)) % NP + NP) % NP;
                const int nown = first <= T ? (T - first) / NP + 1 : 0;
                tc = clock64();
                if (k >= 1) mbar_wait_cluster(&udone, (unsigned)((k - 1) & 1)); // tile column k and tile (k+1,k+1) are final
                SDV_TICK(1);
                // prefetch everything that does not depend on L_kk: the TRSM operand row (registers) and, for the look-ahead,
                // this lane's row of tile (k+1,k+1) (asynchronous copy into shared memory)
                if (warp < nown) load_row32(A + (size_t)((first + warp * NP) * 32 + lane) * ld + k * 32, a);
                if (lookahead) {
                    const double *src = A + (size_t)((k + 1) * 32 + lane) * ld + (k + 1) * 32;
                    const unsigned dst = smem_u32(sDiag + lane * TSTR);
#pragma unroll
                    for (int c = 0; c < 32; c++) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * c), "l"(src + c) : "memory");
                    asm volatile("cp.async.commit_group;" ::: "memory");
                }
                mbar_wait_cluster(&b1[k & 1], (unsigned)((k >> 1) & 1)); // L_kk and its reciprocal diagonal are visible
                SDV_TICK(0);
                if (rank != owner) {
                    for (int e = threadIdx.x; e < 512; e += CCT) {
                        int r = e >> 4, q = e & 15;
                        double2 v = __ldcg(reinterpret_cast<const double2 *>(Lo + (size_t)(k * 32 + r) * ld + k * 32) + q);
                        sK[r * TSTR + 2 * q] = v.x;
                        sK[r * TSTR + 2 * q + 1] = v.y;
                    }
                    if (threadIdx.x < 32) sinv[threadIdx.x] = __ldcg(dinv + k * 32 + threadIdx.x);
                }
                __syncthreads();
                for (int s = warp; s < nown; s += NW) {
                    const int i = first + s * NP;
                    if (s != warp) load_row32(A + (size_t)(i * 32 + lane) * ld + k * 32, a);
                    trsm32_reg(a, sK, sinv);
                    if (s == 0) {
#pragma unroll
                        for (int c = 0; c < 32; c++) rows[lane * TSTR + c] = a[c]; // the look-ahead operand
                    }
                    store_row32(Lo + (size_t)(i * 32 + lane) * ld + k * 32, a);
                }
                __syncthreads();
                SDV_TICK(2);
                if (warp == 1) { // this CTA's part of the panel column is published -> every update CTA
                    __syncwarp();
                    if (lane < NU) mbar_remote_arrive(&b2, (unsigned)(NP + lane));
                }
            }
            if (lookahead) {
                if (k >= 0) {
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 32; c++) a[c] = sDiag[lane * TSTR + c];
#pragma unroll 1
                    for (int q = 0; q < 32; q++) {
                        double lq = rows[(size_t)lane * TSTR + q];
#pragma unroll
                        for (int c = 0; c < 32; c++) a[c] -= lq * rows[(size_t)c * TSTR + q];
                    }
                } else {
                    load_row32(A + (size_t)lane * ld, a);
                }
                SDV_TICK(4); // (panel CTA) diagonal update
                double inv = 1.0;
                if (!chol32_hyb(a, lane, &inv, colbuf)) fail = true;
                SDV_TICK(5); // (panel CTA) tile factorisation alone
#pragma unroll
                for (int c = 0; c < 32; c++) sK[lane * TSTR + c] = a[c];
                sinv[lane] = inv;
                dinv[(k + 1) * 32 + lane] = inv;
                store_row32(Lo + (size_t)((k + 1) * 32 + lane) * ld + (k + 1) * 32, a);
                __syncwarp();
                if (lane < NP) mbar_remote_arrive(&b1[(k + 1) & 1], (unsigned)lane);
                SDV_TICK(3); // (panel CTA) publish
            }
        }
    } else {
        // =================================================== update CTA: every warp is an independent worker
        // Trailing tile (i, j) belongs to global update warp (5 i + j) mod NUWT for the whole factorisation (no read-modify-write
        // races between panels, load balanced over all update warps of the cluster); per tile column j a warp owns the rows
        // i = i0 + m NUWT with 5 i0 + j = gw (mod NUWT). Operands come from L2.
        const int NUWT = NU * NW;
        const int gw = (rank - NP) * NW + warp;
        int inv5 = 1; // modular inverse of 5 (NUWT = 96 -> 77)
        while ((5 * inv5) % NUWT != 1 % NUWT && inv5 < NUWT) inv5++;
        for (int k = 0; k < T; k++) {
            tc = clock64();
            mbar_wait_cluster(&b2, (unsigned)(k & 1)); // the whole panel column k is in Lo
            SDV_TICK(4);
            bool arrived = false;
            for (int j = k + 1; j < T; j++) { // tile columns in the order the panel CTAs will need them
                if (!arrived && j > k + 2) {
                    __syncwarp();
                    if (lane < NP) mbar_remote_arrive(&udone, (unsigned)lane);
                    arrived = true;
                }
                int i0 = (int)(((long long)((gw - j) % NUWT + NUWT) * inv5) % NUWT);
                for (int i = i0; i <= T; i += NUWT) {
                    if (i < j || (i == k + 1 && j == k + 1)) continue; // tile (k+1,k+1) belongs to the look-ahead
                    tile_update_dmma_gg(A + (size_t)(i * 32) * ld + j * 32, Lo + (size_t)(i * 32) * ld + k * 32, Lo + (size_t)(j * 32) * ld + k * 32, ld, lane);
                }
            }
            if (!arrived) {
                __syncwarp();
                if (lane < NP) mbar_remote_arrive(&udone, (unsigned)lane);
            }
            SDV_TICK(5);
        }
    }
    __syncthreads();
    chol_backward_v2(P, B0, B1, st, acc, Lo, dinv, damp_p, graw_p, dxp, sRow, tinv, xs, wsum, red, &bk, fail, prof, tp, tc);
}

// C(32x32) -= A(32x32) B(32x32)^T with plain FP64 FMAs, all three tiles in global memory (L2), one warp: lane r owns row r of C
// and of A in registers, B is staged transposed in this warp's shared-memory buffer and read back as broadcast 16-byte loads.
// On B200 mma.sync.m8n8k4.f64 issues at ~1 per 25 cycles per warp and does not scale with the warps of an SM (measured: 3.3k
// cycles per tile alone, ~18k with 8 warps per SM), the FP64 FMA pipe of each SM sub-partition does.
constexpr int BTS = 34; // row stride of the transposed B tile (16-byte aligned rows)
constexpr int CHAIN_SMEM_DOUBLES = 8 * 32 * BTS; // >= 2 * 1024 + 64 + 2 * 1024
__device__ __noinline__ void tile_update_dfma_gg(double *Cg, const double *Ag, const double *Bg, int ld, int lane, double *Bt) {
    double a[32], c[32];
    {
        double b[32];
        load_row32(Bg + (size_t)lane * ld, b);
#pragma unroll
        for (int q = 0; q < 32; q++) Bt[q * BTS + lane] = b[q];
    }
    load_row32(Ag + (size_t)lane * ld, a);
    load_row32(Cg + (size_t)lane * ld, c);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 32; q++) {
        const double aq = a[q];
#pragma unroll
        for (int c2 = 0; c2 < 32; c2 += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(Bt + q * BTS + c2);
            c[c2] -= aq * v.x;
            c[c2 + 1] -= aq * v.y;
        }
    }
    store_row32(Cg + (size_t)lane * ld, c);
    __syncwarp();
}

// =====================================================================================================================
// Variant 4: one "chain" CTA + panel CTAs + update CTAs.
//
// The factorisation of an n ~ 750 system is one long dependency chain: tile k's Cholesky -> triangular solve of tile row k+1
// -> update of tile (k+1,k+1) -> tile k+1's Cholesky ...  Here that chain never leaves one CTA and never touches L2: two warps
// of the chain CTA alternate.  While warp A factorises tile (k,k) it publishes every finished column in shared memory; warp B
// holds tile row k+1 (tiles (k+1,k) and (k+1,k+1)) in registers and consumes the columns as they appear (column-oriented
// triangular solve + rank-1 update of its diagonal tile), so it can start factorising tile (k+1,k+1) ~one column after A ends.
// A then becomes the B of tile row k+2.  The other tile rows are solved by the panel CTAs (from L2, after the tile is
// published) and the trailing tiles are updated by the update CTAs with DMMA; the two tiles the chain needs next,
// (k+2,k+1) and (k+2,k+2), are updated first and signalled separately.
// =====================================================================================================================
__host__ __device__ constexpr int mod_inverse(int a, int m) {
    for (int x = 1; x < m; x++)
        if ((a * x) % m == 1) return x;
    return 0;
}
// CTA-local mbarrier helpers (SYNCS.ARRIVE / try_wait, no MEMBAR: a st.release.cta flag costs a MEMBAR.ALL.CTA per column)
SDV_DEV void mbar_arrive_cta(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
SDV_DEV bool mbar_test_cta(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
SDV_DEV void mbar_wait_cta(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITL_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEL_%=;\n"
        "bra WAITL_%=;\n"
        "DONEL_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// chol32_hyb that also publishes: colT[c*32 + i] = L[i][c], sinvT[c] = 1/L[c][c], then one arrival on colbar[c]
SDV_DEV bool chol32_stream(double (&a)[32], int lane, double *colT, double *sinvT, uint64_t *colbar) {
    bool ok = true;
    double d = __shfl_sync(FULL, a[0], 0);
    if (!(d > 0.0) || !isfinite(d)) {
        ok = false;
        d = 1.0;
    }
    double inv = rsqrt(d);
#pragma unroll
    for (int c = 0; c < 32; c++) {
        double l = a[c] * inv;
        if (lane == c) {
            l = d * inv;
            sinvT[c] = inv;
        }
        if (lane < c) l = 0.0;
        a[c] = l;
        colT[c * 32 + lane] = l;
        if (c + 1 < 32) {
            a[c + 1] -= l * __shfl_sync(FULL, l, c + 1); // the next pivot column first
            d = __shfl_sync(FULL, a[c + 1], c + 1);
            if (!(d > 0.0) || !isfinite(d)) {
                ok = false;
                d = 1.0;
            }
            inv = rsqrt(d); // in flight during the updates below
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(colbar + c);
#pragma unroll
        for (int c2 = c + 2; c2 < 32; c2 += 2) {
            if (c2 + 1 < 32 && (c2 & 1) == 0) {
                const double2 v = *reinterpret_cast<const double2 *>(colT + c * 32 + c2);
                a[c2] -= l * v.x;
                a[c2 + 1] -= l * v.y;
            } else {
                a[c2] -= l * colT[c * 32 + c2];
                if (c2 + 1 < 32) a[c2 + 1] -= l * colT[c * 32 + c2 + 1];
            }
        }
    }
    return ok;
}

// Consumer: t = row `lane` of tile (r, k) (becomes L[r][k]), d = row `lane` of tile (r, r) (receives -= L_rk L_rk^T).
// xb[c*32 + i] receives L[r][k](i, c) (the finished tile, column-major) for the publishing warp.
template <bool DIAG>
SDV_DEV void trsm_stream(double (&t)[32], double (&d)[32], int lane, const double *colT, const double *sinvT, uint64_t *colbar, unsigned parity,
                         double *xb /* shared [32][32], private to this warp */) {
    const bool all_done = mbar_test_cta(colbar + 31, parity); // catching up with a finished tile: no per-column waits
#pragma unroll
    for (int c = 0; c < 32; c++) {
        if (!all_done) mbar_wait_cta(colbar + c, parity);
        const double x = t[c] * sinvT[c];
        t[c] = x;
#pragma unroll
        for (int j = c + 1; j < 32; j++) t[j] -= x * colT[c * 32 + j];
        xb[c * 32 + lane] = x;
        if (DIAG) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(xb + c * 32 + j);
                d[j] -= x * v.x;
                d[j + 1] -= x * v.y;
            }
        }
    }
}

template <int NPC, bool DFMA>
__global__ void __launch_bounds__(CCT, 1) k_chol_chain(DevProblem P, LinBuf B0, LinBuf B1, LMState *st, Accum *acc, double *A, double *Lo, double *dinv,
                                                       const double *damp_p, const double *graw_p, double *dxp, int max_rows, double *prof) {
    if (st->status != 0) return; // uniform over the cluster
    extern __shared__ __align__(16) double csm[];
    __shared__ uint64_t b1[2], b2, ucol, urg[2], bk; // urg[k&1]: each chain warp waits on its own barrier, phase after phase
    __shared__ uint64_t colbar[2][32], pubT[2], pubD[2], ackT[2], ackD[2]; // chain CTA only
    __shared__ uint32_t snz[129 * 4];                                      // structural tile pattern of L (host-side symbolic factorisation)
    constexpr int NW = CCT / 32;
    double *sK = csm;                                   // [32][TSTR] diagonal tile L_kk (panel CTAs)
    double *sinv = sK + 32 * TSTR;                      // [32]
    double *chain = sinv + 32;                          // chain CTA: colT[2][1024], sinvT[2][32], xb[2][1024]; update CTAs: Bt[8][32][BTS]
    double *sRow = chain + CHAIN_SMEM_DOUBLES;    // backward solve: tile inverses [max_rows][32][TSTR] (x2 kept for layout parity)
    double *xs = sRow + (size_t)2 * max_rows * 32 * TSTR;
    double *wsum = xs + (size_t)max_rows * 32;
    double *red = wsum + 8 * 32;
    double *tinv = red + CC_MAX * 32;
    const int ld = P.ld, T = P.n_pad / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = (int)cluster_rank();
    constexpr int NU = CC_MAX - 1 - NPC, NUWT = NU * NW; // the host launches this kernel only with a full 16-CTA cluster
    bool fail = false;
    long long tp[7] = {0, 0, 0, 0, 0, 0, 0}, tc = clock64(), tn;

    for (int e = threadIdx.x; e < (T + 1) * 4; e += CCT) snz[e] = __ldg(P.tile_nz + e);
    auto nz = [&](int i, int k) { return ((snz[i * 4 + (k >> 5)] >> (k & 31)) & 1u) != 0; };
    if (threadIdx.x == 0) {
        mbar_init(&b1[0], 1);
        mbar_init(&b1[1], 1);
        mbar_init(&b2, NPC + 1);
        mbar_init(&ucol, NUWT);
        mbar_init(&urg[0], 2);
        mbar_init(&urg[1], 2);
        for (int i = 0; i < 64; i++) mbar_init(&colbar[0][0] + i, 1);
        for (int i = 0; i < 2; i++) {
            mbar_init(&pubT[i], 1);
            mbar_init(&pubD[i], 1);
            mbar_init(&ackT[i], 1);
            mbar_init(&ackD[i], 1);
        }
    }
    __syncthreads();
    cluster_sync_all();

    if (rank == 0) {
        // =================================================== chain CTA: warps 0 and 1 alternate, warps 2 and 3 publish for them
        double *colT = chain, *sinvT = chain + 2048;
        if (warp < 2) {
            const int w = warp;                       // tile k with k % 2 == w: colT[w], sinvT[w], colbar[w] are produced by this warp
            double *xb = chain + 2048 + 64 + w * 1024;
            double t[32], d[32];
            int nB = 0, nD = 0;                        // solve / factor phases finished by this warp (phases of pubT/ackT, pubD/ackD)
            for (int k = w; k <= T; k += 2) {          // k == T: only the last triangular solve of the right-hand-side row
                if (k == 0) {
                    load_row32(A + (size_t)lane * ld, d);
                } else {
                    const int pq = (k - 1) & 1;
                    tc = clock64();
                    if (k >= 2) mbar_wait_cluster(&urg[k & 1], (unsigned)(((k - 2) >> 1) & 1)); // tiles (k,k-1), (k,k) final through panel k-2
                    SDV_TICK(0);
                    load_row32(A + (size_t)(k * 32 + lane) * ld + (k - 1) * 32, t);
                    if (k < T) load_row32(A + (size_t)(k * 32 + lane) * ld + k * 32, d);
                    if (nB > 0) mbar_wait_cta(&ackT[w], (unsigned)((nB - 1) & 1)); // the publisher has copied the previous xb
                    SDV_TICK(1);
                    if (k < T) trsm_stream<true>(t, d, lane, colT + pq * 1024, sinvT + pq * 32, &colbar[pq][0], (unsigned)(((k - 1) >> 1) & 1), xb);
                    else trsm_stream<false>(t, d, lane, colT + pq * 1024, sinvT + pq * 32, &colbar[pq][0], (unsigned)(((k - 1) >> 1) & 1), xb);
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta(&pubT[w]);
                    nB++;
                    SDV_TICK(2);
                }
                if (k < T) {
                    if (nD > 0) mbar_wait_cta(&ackD[w], (unsigned)((nD - 1) & 1)); // the publisher has copied the previous tile
                    if (!chol32_stream(d, lane, colT + w * 1024, sinvT + w * 32, &colbar[w][0])) fail = true;
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta(&pubD[w]);
                    nD++;
                    SDV_TICK(3);
                }
            }
        } else if (warp < 4) {
            // publisher of chain warp w: copies finished tiles from shared memory to L2 and signals the other CTAs, so the
            // chain warps never wait for a global store
            const int w = warp - 2;
            const double *xb = chain + 2048 + 64 + w * 1024;
            double a[32];
            int nB = 0, nD = 0;
            for (int k = w; k <= T; k += 2) {
                if (k >= 1) {
                    mbar_wait_cta(&pubT[w], (unsigned)(nB & 1));
#pragma unroll
                    for (int c = 0; c < 32; c++) a[c] = xb[c * 32 + lane];
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta(&ackT[w]);
                    store_row32(Lo + (size_t)(k * 32 + lane) * ld + (k - 1) * 32, a);
                    __syncwarp();
                    if (k < T && lane < NU) mbar_remote_arrive(&b2, (unsigned)(1 + NPC + lane)); // chain's part of panel column k-1
                    nB++;
                }
                if (k < T) {
                    mbar_wait_cta(&pubD[w], (unsigned)(nD & 1));
#pragma unroll
                    for (int c = 0; c < 32; c++) a[c] = colT[w * 1024 + c * 32 + lane];
                    const double inv = sinvT[w * 32 + lane];
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta(&ackD[w]);
                    dinv[k * 32 + lane] = inv;
                    store_row32(Lo + (size_t)(k * 32 + lane) * ld + k * 32, a);
                    __syncwarp();
                    if (lane < NPC) mbar_remote_arrive(&b1[w], (unsigned)(1 + lane));
                    nD++;
                }
            }
        }
    } else if (rank <= NPC) {
        // =================================================== panel CTA: tile rows i >= k+2 with i % NPC == rank-1
        const int pr = rank - 1;
        for (int k = 0; k + 2 <= T; k++) {
            double a[32];
            const int first = k + 2 + ((pr - (k + 2)) % NPC + NPC) % NPC;
            const int nown = first <= T ? (T - first) / NPC + 1 : 0;
            tc = clock64();
            if (k >= 1) mbar_wait_cluster(&ucol, (unsigned)((k - 1) & 1)); // tiles (i,k), i >= k+2, final through panel k-1
            SDV_TICK(1);
            const bool mine = warp < nown && nz(first + warp * NPC, k); // structurally zero tiles stay zero in Lo (cleared at upload)
            if (mine) load_row32(A + (size_t)((first + warp * NPC) * 32 + lane) * ld + k * 32, a);
            mbar_wait_cluster(&b1[k & 1], (unsigned)((k >> 1) & 1)); // L_kk and its reciprocal diagonal are in L2
            SDV_TICK(0);
            for (int e = threadIdx.x; e < 512; e += CCT) {
                int r = e >> 4, q = e & 15;
                double2 v = __ldcg(reinterpret_cast<const double2 *>(Lo + (size_t)(k * 32 + r) * ld + k * 32) + q);
                sK[r * TSTR + 2 * q] = v.x;
                sK[r * TSTR + 2 * q + 1] = v.y;
            }
            if (threadIdx.x < 32) sinv[threadIdx.x] = __ldcg(dinv + k * 32 + threadIdx.x);
            __syncthreads();
            for (int s2 = warp; s2 < nown; s2 += NW) {
                const int i = first + s2 * NPC;
                if (!nz(i, k)) continue;
                if (s2 != warp) load_row32(A + (size_t)(i * 32 + lane) * ld + k * 32, a);
                trsm32_reg(a, sK, sinv);
                store_row32(Lo + (size_t)(i * 32 + lane) * ld + k * 32, a);
            }
            __syncthreads();
            SDV_TICK(2);
            if (warp == 1) { // this CTA's part of panel column k is published -> every update CTA
                __syncwarp();
                if (lane < NU) mbar_remote_arrive(&b2, (unsigned)(1 + NPC + lane));
            }
        }
    } else {
        // =================================================== update CTA: every warp is an independent worker
        // Tile (i, j) belongs to global update warp (5 i + j) mod NUWT for the whole factorisation; operands come from L2.
        const int gw = (rank - 1 - NPC) * NW + warp;
        constexpr int inv5 = mod_inverse(5, NUWT); // compile-time constants: no integer division in the tile loops
        static_assert((5 * inv5) % NUWT == 1, "5 must be invertible modulo the number of update warps");
        auto owner = [&](int i, int j) { return (5 * i + j) % NUWT; };
        double *Bt = chain + warp * 32 * BTS;
        auto update = [&](int i, int j, int k) {
            if (!nz(i, k) || !nz(j, k)) return; // a structurally zero operand: nothing to subtract
            if (DFMA)
                tile_update_dfma_gg(A + (size_t)(i * 32) * ld + j * 32, Lo + (size_t)(i * 32) * ld + k * 32, Lo + (size_t)(j * 32) * ld + k * 32, ld, lane, Bt);
            else
                tile_update_dmma_gg(A + (size_t)(i * 32) * ld + j * 32, Lo + (size_t)(i * 32) * ld + k * 32, Lo + (size_t)(j * 32) * ld + k * 32, ld, lane);
        };
        for (int k = 0; k + 2 <= T; k++) { // panel T-1 has no trailing tile
            tc = clock64();
            mbar_wait_cluster(&b2, (unsigned)(k & 1)); // the whole panel column k is in Lo
            SDV_TICK(4);
            // 1. the two tiles the chain needs next
            if (owner(k + 2, k + 1) == gw) {
                update(k + 2, k + 1, k);
                __syncwarp();
                if (lane == 0) mbar_remote_arrive(&urg[k & 1], 0u);
            }
            if (owner(k + 2, k + 2) == gw) {
                if (k + 2 <= T - 1) update(k + 2, k + 2, k);
                __syncwarp();
                if (lane == 0) mbar_remote_arrive(&urg[k & 1], 0u);
            }
            // 2. the rest of tile column k+1 (the next panel), then the others in the order they will be needed
            for (int j = k + 1; j < T; j++) {
                if (j == k + 2) {
                    __syncwarp();
                    if (lane < NPC) mbar_remote_arrive(&ucol, (unsigned)(1 + lane));
                }
                const int i0 = (((gw - j) % NUWT + NUWT) * inv5) % NUWT;
                for (int i = i0; i <= T; i += NUWT) {
                    if (i < j || i <= k + 1) continue;                    // tile (k+1,k+1) belongs to the chain
                    if (i == k + 2 && (j == k + 1 || j == k + 2)) continue; // done above
                    update(i, j, k);
                }
            }
            if (k + 2 >= T) { // column loop did not reach j == k+2
                __syncwarp();
                if (lane < NPC) mbar_remote_arrive(&ucol, (unsigned)(1 + lane));
            }
            SDV_TICK(5);
        }
    }
    __syncthreads();
    chol_backward_v2(P, B0, B1, st, acc, Lo, dinv, damp_p, graw_p, dxp, sRow, tinv, xs, wsum, red, &bk, fail, prof, tp, tc);
}

// Developer micro-benchmark: cycles of the tile routines, single warp, 5 repetitions each (the first one has cold code).
// out[routine * 8 + rep]; routines: 0 chol32_reg, 1 chol32_smem, 2 trsm32_reg, 3 trsm32_smem, 4 diag update (rolled q),
// 5 tile_update_dmma (C in global), 6 load_row32 (L2), 7 store tile rows
__global__ void k_chol_micro(double *scratch /* >= 4 * 32 * 64 doubles */, double *out) {
    __shared__ double sA[32 * TSTR], sL[32 * TSTR], sinv[32], sX[32 * TSTR];
    const int lane = threadIdx.x;
    // SPD tile: 40 I + small symmetric part
    for (int c = 0; c < 32; c++) sA[lane * TSTR + c] = (lane == c ? 40.0 : 0.0) + 0.01 * ((lane * 7 + c * 3) % 11 + (c * 7 + lane * 3) % 11);
    __syncwarp();
    for (int e = lane; e < 32 * 64; e += 32) scratch[e] = 0.001 * (e % 97);
    __syncwarp();
    double sink = 0;
    for (int rep = 0; rep < 5; rep++) {
        long long t0, t1;
        {   // 0: chol32_reg
            double a[32], inv = 1;
            for (int c = 0; c < 32; c++) a[c] = sA[lane * TSTR + c];
            __syncwarp();
            t0 = clock64();
            chol32_reg(a, lane, &inv);
            t1 = clock64();
            for (int c = 0; c < 32; c++) sL[lane * TSTR + c] = a[c];
            sinv[lane] = inv;
            sink += a[lane & 31 ? 1 : 0];
            if (lane == 0) out[0 * 8 + rep] = (double)(t1 - t0);
            __syncwarp();
        }
        {   // 1: chol32_hyb
            double a[32], inv = 1;
            for (int c = 0; c < 32; c++) a[c] = sA[lane * TSTR + c];
            __syncwarp();
            t0 = clock64();
            chol32_hyb(a, lane, &inv, sX);
            t1 = clock64();
            sink += a[lane & 31 ? 2 : 0] + inv;
            if (lane == 0) out[1 * 8 + rep] = (double)(t1 - t0);
            __syncwarp();
        }
        {   // 2: trsm32_reg
            double a[32];
            for (int c = 0; c < 32; c++) a[c] = sA[lane * TSTR + c];
            t0 = clock64();
            trsm32_reg(a, sL, sinv);
            t1 = clock64();
            for (int c = 0; c < 32; c++) sink += a[c];
            if (lane == 0) out[2 * 8 + rep] = (double)(t1 - t0);
            __syncwarp();
        }
        {   // 3: trsm32_smem
            for (int c = 0; c < 32; c++) sX[lane * TSTR + c] = sA[lane * TSTR + c];
            __syncwarp();
            t0 = clock64();
            trsm32_smem(sX, sL, sinv, lane);
            t1 = clock64();
            sink += sX[lane * TSTR + 5];
            if (lane == 0) out[3 * 8 + rep] = (double)(t1 - t0);
            __syncwarp();
        }
        {   // 4: diag update, rolled over q
            double a[32];
            for (int c = 0; c < 32; c++) a[c] = sA[lane * TSTR + c];
            t0 = clock64();
#pragma unroll 1
            for (int q = 0; q < 32; q++) {
                double lq = sL[lane * TSTR + q];
#pragma unroll
                for (int c = 0; c < 32; c++) a[c] -= lq * sL[c * TSTR + q];
            }
            t1 = clock64();
            sink += a[7];
            if (lane == 0) out[4 * 8 + rep] = (double)(t1 - t0);
            __syncwarp();
        }
        {   // 5: DMMA tile update with C in global memory, B in shared
            t0 = clock64();
            tile_update_dmma(scratch, 64, sL, sA, TSTR, false, lane);
            t1 = clock64();
            if (lane == 0) out[5 * 8 + rep] = (double)(t1 - t0);
            __syncwarp();
        }
        {   // 6: load one row per lane from L2
            double a[32];
            t0 = clock64();
            load_row32(scratch + (size_t)lane * 64, a);
            sink += a[9] + a[31];
            t1 = clock64();
            if (lane == 0) out[6 * 8 + rep] = (double)(t1 - t0);
            __syncwarp();
        }
        {   // 7: coalesced tile store
            t0 = clock64();
            for (int e = lane; e < 512; e += 32) {
                int r = e >> 4, q = e & 15;
                reinterpret_cast<double2 *>(scratch + (size_t)r * 64 + 32)[q] = make_double2(sL[r * TSTR + 2 * q], sL[r * TSTR + 2 * q + 1]);
            }
            t1 = clock64();
            if (lane == 0) out[7 * 8 + rep] = (double)(t1 - t0);
            __syncwarp();
        }
    }
    if (sink == 123.456) out[63] = sink;
}

// Developer micro-benchmark: throughput of the two trailing-update kernels when `nact` warps of one CTA run them at once.
// out[0] = cycles per tile seen by warp 0 (4 tiles per warp, warm).
__global__ void k_update_micro(double *scratch /* >= 8 * 3 * 32 * 64 doubles */, double *out, int dfma, int nact) {
    extern __shared__ __align__(16) double usm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *base = scratch + (size_t)warp * 3 * 32 * 64;
    for (int e = lane; e < 3 * 32 * 64; e += 32) base[e] = 0.001 * ((e + warp) % 97);
    __syncthreads();
    if (warp >= nact) return;
    long long t0 = 0;
    for (int rep = 0; rep < 5; rep++) {
        if (rep == 1) t0 = clock64();
        if (dfma) tile_update_dfma_gg(base, base + 32 * 64, base + 2 * 32 * 64, 64, lane, usm + warp * 32 * BTS);
        else tile_update_dmma_gg(base, base + 32 * 64, base + 2 * 32 * 64, 64, lane);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (double)(t1 - t0) / 4.0;
}

} // namespace sdv
