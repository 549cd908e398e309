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
// In-place X <- X L^-T for the tile X (shared memory, row stride TSTR), one row per lane; L and 1/diag in shared memory.
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
// Hybrid register Cholesky: the pivot column is broadcast through a 32-double shared-memory buffer (one store + broadcast
// loads per column) instead of 2 SHFL per element — SHFL issue was the throughput limit of chol32_reg (8.3k cycles/tile);
// only the element of the NEXT pivot column, which sits on the dependency chain, still goes through a shuffle.
// cluster-scope release: the panel warps read the updated tiles back through L2 (__ldcg)
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
__global__ void __launch_bounds__(CCT, 1) k_chol_chain(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, LMState *st, Accum *acc, double *A, double *Lo, double *dinv,
                                                       const double *damp_p, const double *graw_p, double *dxp, int max_rows, double *prof) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
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
// Developer micro-benchmark: throughput of the two trailing-update kernels when `nact` warps of one CTA run them at once.
// out[0] = cycles per tile seen by warp 0 (4 tiles per warp, warm).
} // namespace sdv
