// Banded FP64 Cholesky + triangular solves of the reduced system in ONE CTA (variant 6, the default whenever it applies).
//
// Why: the reduced (pose/velocity/bias) system of a sliding window is block-banded — a landmark couples only the keyframes
// that see it, an IMU factor only consecutive keyframes (AOptimizer.cpp:59-85) — and the envelope of a Cholesky factor is
// the envelope of the matrix.  For the headline window (n = 735) the half-bandwidth is 59 columns: the whole factorisation
// is ONE dependency chain of 735 pivots with a working set of 80 x 80 doubles.  Spreading that over a cluster (variants 0-5)
// pays an L2 / DSMEM round trip per tile step; here the active window never leaves the shared memory of one SM and the
// pivot chain never leaves the registers of one warp.
//
// Blocks are 16 x 16 (BN).  Half-bandwidth bw (in blocks, computed on the host from the factor graph, sdv_lib.cu).
// Right-looking, per block step k:
//   C  warp 0 ("chain"): rows of block k AND block k+1 are stacked in its registers (16 lanes each; the two half-warps swap
//      roles every step, so the rows of block k+1 are already in place when they become pivots).  One pass over the 16
//      pivots factors D_k, solves P_(k+1,k) = W_(k+1,k) L_kk^-T and applies D_(k+1) -= P P^T on the fly: the next diagonal
//      block is final the moment the last pivot is done.  Per pivot the chain is  mul -> fma -> shfl -> rsqrt  (the next
//      pivot is formed on its own lane before the broadcast; rsqrt = rsqrt.approx.f64 + one third-order correction).
//      Meanwhile the 15 worker warps write the previous panel to global memory (L is needed again by the backward solve),
//      invert the previous diagonal block, and prefetch the next block row of S with cp.async.
//   T  workers: P_(k+d,k) = W_(k+d,k) L_kk^-T for d = 2..bw, and the right-hand side row (L y = g rides along).
//   U  workers: W_(k+di,k+dj) -= P_di P_dj^T with FP64 tensor-core MMAs (mma.sync.m8n8k4.f64, SASS DMMA), g -= P y.
// Backward solve L^T x = y: one warp, block columns of L streamed back from L2 with TMA bulk copies (4-stage mbarrier
// ring); with the explicit inverse of every diagonal block a step is two small matrix-vector products.
#pragma once
#include "sdv_chol.cuh"

namespace sdv {

constexpr int BN = 16;          // block size
constexpr int WSTR = 18;        // row stride (doubles) of a window block: 16-byte aligned rows, conflict-free LDS.128 per row
constexpr int WBLK = BN * WSTR; // doubles per window block
constexpr int BCT = 512;        // threads of the CTA
constexpr int BNW = BCT / 32;
constexpr int BAND_MAX_BW = 7;  // task tables / shared memory are sized for this
constexpr int BAND_MAX_STAGES = 16; // backward-solve ring (TMA bulk copies in flight)
// EXPERIMENT (off): block rows k+1 .. k+3 apply ALL their updates as rank-1 updates behind the chain ("streaming rows", roles 7
// and 1), the tensor-core update warps keep rows >= k+4.  Correct (the parity tests pass with -DSDV_BAND_STREAM_UPDATES=1)
// but slower at C3 (259 us against 135 us): under the 128-register cap of a 512-thread CTA the d = 3 row spills (t[16] +
// 3 x 8 accumulators + operands) and needs ~430 cycles per pivot where the chain needs 116, so the chain waits for it.
// To pay it needs 256 threads per CTA (255 registers) or the d = 3 row split over two warps.
#ifndef SDV_BAND_STREAM_UPDATES
#define SDV_BAND_STREAM_UPDATES 0
#endif
constexpr bool BAND_STREAM_UPDATES = SDV_BAND_STREAM_UPDATES != 0;
// FIRST-COLUMN STREAMING (-DSDV_BAND_STREAM1=1; parity green on B200, OFF by default: measured 117.7 us against 112.4 us at C3).
// Idea: what gates a block step is not only the pivot chain but the two tensor-core tasks behind it — per-warp clock64 traces
// (SDV_BAND_PROF, round 2, 26 steps on CTA 0): chain 2.4 k cycles per step, and per step the chain waits 0.75 k for (1,1), 0.5 k
// for (2,1), 0.26 k for the update warps of step k-2, 0.5 k once for the hand-over of the two-way dissection.  With this switch
// every first-column task is applied as rank-1 updates behind the chain, one pivot column at a time, by a warp that has its
// operands anyway (measured: the two waits shrink to 0.46 k each — the streaming warps consume the pivot columns in groups of
// four, so they still end ~0.5 k cycles after the last pivot — while the chain itself slows from 2.4 k to 3.0 k cycles per
// step under the extra shared-memory traffic of five more warps reading every published column: a net loss):
//   (1,1)  D_(k+1) -= P_1 P_1^T          warp 15 (role 7): 8 FMAs per lane and pivot
//   (d,1)  W_(k+d,k+1) -= P_d P_1^T      the row-solve warp of distance d, whose lanes produce P_d[r][c] one column at a time; the
//                                        two half-warps duplicate the solve and split the 16 columns (band_trsm16<true>)
// Both accumulate from zero and are added to the window block at the end of the step, so they never wait at its start.  The
// blocks the next step reads first are final ~100 cycles after the last pivot; the tensor-core warps keep the tasks (di, dj >= 2),
// which nobody needs before the step after the next.  Unlike SDV_BAND_STREAM_UPDATES (all blocks of rows k+1 .. k+3 streamed:
// the d = 3 row could not keep the chain's pace under the 128-register cap) every warp adds 8 FMAs per pivot to what it did.
#ifndef SDV_BAND_STREAM1
#define SDV_BAND_STREAM1 0
#endif
#if SDV_BAND_STREAM1 && SDV_BAND_STREAM_UPDATES
#error "SDV_BAND_STREAM1 excludes SDV_BAND_STREAM_UPDATES"
#endif
constexpr bool BAND_STREAM1 = SDV_BAND_STREAM1 != 0;
// Specialised full steps in the backward solve (band_backward_full_step below).  Default since round 2 (measured on B200:
// 142 -> 135 us at C3 alone; all parity tests green); -DSDV_BAND_BACKWARD_V2=0 restores the generic loop.
#ifndef SDV_BAND_BACKWARD_V2
#define SDV_BAND_BACKWARD_V2 1
#endif
// EXPERIMENT (-DSDV_BAND_BACKWARD_V3=1; parity green on B200, OFF by default: no gain).  Software-pipelined full steps of the
// backward solve (band_backward_pipe_step below): the products with x_(k+2) .. x_(k+bw) of step k are formed during step k+1, in
// the shadow of its shuffles and its trip through shared memory, so that the dependence chain of a step is the d = 1 block, one
// reduction and the product with the inverse.  Measured (tools/micro/backward.cu, one warp alone on the SM, cycles per block
// step at bw = 4): shipped loop 683-706, V2 658-668, V3 669-679 with the TMA ring; with the factor resident in shared memory
// V2 517, V3 485, and 386 for the d = 1 block + inverse alone — i.e. the dependence chain itself (two shared-memory round trips,
// two reductions, the publication of x) is ~390 cycles, the blocks d >= 2 cost ~100-130 either way, and the ring (one mbarrier
// try_wait, 90 cycles even when complete, + one expect_tx / bulk-copy issue per step) ~170.  In the kernel: 112.5 us against 112.7 us.
#ifndef SDV_BAND_BACKWARD_V3
#define SDV_BAND_BACKWARD_V3 0
#endif
// First milestone of the two-way dissection, a debugging aid (parity green on B200, round 2): with
// -DSDV_BAND_REV=1 the whole kernel factors P S P instead of S (P = index reversal, row i <-> n_pad-1-i, which keeps the
// 16-column blocks aligned) and scatters the solution back — what the second CTA of the cluster will do on its half.  The
// band of P S P is the band of S, the solution is the same up to rounding: every parity test applies unchanged.
#ifndef SDV_BAND_REV
#define SDV_BAND_REV 0
#endif
// The two-way dissection itself ("burn at both ends"), DEFAULT since round 2 (B200: 142 -> 113 us at C3, 511 -> 384 us at C5,
// all parity tests green; -DSDV_BAND_BABE=0 builds the one-CTA kernel of round 1).  Launched
// as a 2-CTA cluster (sdv_lib.cu does so when the band is long enough) CTA 0 eliminates block rows 0 .. nl-1 of S, CTA 1 block
// rows 0 .. nr-1 of P S P, CTA 0 adds CTA 1's separator update (read through distributed shared memory), finishes the
// separator and the forward substitution, solves the separator unknowns, hands them to CTA 1, and both back-substitute
// their interiors.  Launched as a single CTA the kernel behaves as before.  tools/babe_prototype.py is the numpy model.
#ifndef SDV_BAND_BABE
#define SDV_BAND_BABE 1
#endif
#if SDV_BAND_BABE && (SDV_BAND_REV || SDV_BAND_STREAM_UPDATES)
#error "SDV_BAND_BABE excludes SDV_BAND_REV and SDV_BAND_STREAM_UPDATES"
#endif
#if SDV_BAND_BABE
#define BAND_BNB nbk
#else
#define BAND_BNB nb
#endif
#if SDV_BAND_BABE
#define BAND_KBEG kbeg
#define BAND_KEND kend
#define BAND_KLIM k < kend &&
#else
#define BAND_KBEG 0
#define BAND_KEND nb
#define BAND_KLIM
#endif

// shared-memory plan, identical on host and device.  The panel of a step (L_kk and P_(k+1,k) .. P_(k+bw,k), stacked) is
// stored TRANSPOSED: column c of the stacked panel is contiguous, pan[c * pcs + 16 d + row]; pcs = 16 (bw + 1) + 4 makes the
// FP64 MMA fragment loads conflict-free and lets the chain warp read "the rest of column c" with one base address.
struct BandPlan {
    int nb, bw, R;                 // block rows, half-bandwidth in blocks, window rows kept in shared memory
    int pcs, pan_doubles;          // panel column stride, doubles per panel buffer
    int stages;                    // backward ring depth
    int o_win, o_pan, o_inv, o_ys, o_g, o_dmp, o_end; // offsets in doubles
};
__host__ __device__ inline BandPlan band_plan(int n_pad, int bw) {
    BandPlan p;
    p.nb = n_pad / BN;
    p.bw = bw;
    p.R = p.nb < bw + 3 ? p.nb : bw + 3; // live rows k .. k+bw, plus two slots for the prefetch of block row k+bw+2
    p.pcs = BN * (bw + 1) + 4;
    p.pan_doubles = BN * p.pcs + 32; // + slack: the chain warp reads up to 31 doubles past its column
    const int win = p.R * (bw + 1) * WBLK;
    int fwd = win + 2 * p.pan_doubles;
    // the backward ring reuses window + panels; give it up to BAND_MAX_STAGES stages while the CTA stays below ~200 KB
    const int stage = (bw + 2) * BN * BN;
    int stages = p.nb < BAND_MAX_STAGES ? p.nb : BAND_MAX_STAGES;
    while (stages > 2 && (stages * stage + 2 * n_pad + 64) * 8 > 200 * 1024) stages--;
    p.stages = stages;
    if (fwd < stages * stage) fwd = stages * stage;
    p.o_win = 0;
    p.o_pan = win;
    p.o_inv = fwd;
    p.o_ys = p.o_inv + 2 * BN;
    p.o_g = p.o_ys + 2 * BN;
    p.o_dmp = p.o_g + n_pad; // LM damping of the columns (negative = padding column: unit diagonal)
    p.o_end = p.o_dmp + n_pad;
    return p;
}

// 1/sqrt(d): hardware seed (~2^-22) + one third-order step (error ~ 5/16 e^3), no special cases: a non-positive or
// non-finite pivot is detected separately and poisons the result, which is then discarded.
SDV_DEV double band_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = d * y;
    const double e = fma(-t, y, 1.0);
    const double q = fma(0.375, e, 0.5);
    const double ye = y * e;
    return fma(ye, q, y);
}

// One block step of the chain warp: lanes 0..15 hold row `lane` of the pivot block D_k, lanes 16..31 row `lane - 16` of
// W_(k+1,k), so the triangular solve of the next block row costs nothing (same instructions, otherwise idle lanes).
// Measured on B200 (tools/micro/chain2.cu): mul -> fma -> shfl -> rsqrt is 106 cycles per pivot; every shared-memory
// load issued by this warp adds ~3 cycles, selects on the chain add 19.  Hence: the next pivot is formed on its own lane
// before the broadcast, the zeroing of the rows above the pivot is kept off the chain, the broadcast of the pivot column
// uses 16-byte loads (two pivots per loop iteration so the alignment is static), and the update of D_(k+1) is NOT done
// here but by a worker warp with tensor-core MMAs.  The register file ROTATES by one column per pivot
// (a[j-1] = a[j] - l v[j]), which keeps this a real loop of ~100 instructions — the fully unrolled triangular version
// (2 x 1700 instructions) was instruction-fetch bound.
// Publishes column c of [L_kk ; P_(k+1,k)] at pan[c * pcs + 0..31] and 1/diag in iv[c].
// ptxas schedules inside a basic block but cannot see the chain across pivots; left to itself it issued the publication of
// the column AFTER the broadcast + rsqrt of the next pivot, i.e. it serialised the two dependence chains of a pivot.  The
// memory operations are therefore volatile asm in the order they must be issued: publish, shuffle, sync, reload.
SDV_DEV void st_shared_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
SDV_DEV double ld_shared_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
SDV_DEV double2 ld_shared_v2f64(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
SDV_DEV double shfl_f64_volatile(double x, int src) {
    int lo = __double2loint(x), hi = __double2hiint(x);
    asm volatile("shfl.sync.idx.b32 %0, %0, %2, 0x1f, 0xffffffff;\n\tshfl.sync.idx.b32 %1, %1, %2, 0x1f, 0xffffffff;" : "+r"(hi), "+r"(lo) : "r"(src) : "memory");
    return __hiloint2double(hi, lo);
}

template <int C>
SDV_DEV void band_chain_pivot(double (&a)[16], int lane, uint32_t pan_s, uint32_t iv_s, uint64_t *colbar, int pcs, double &d, double &inv, bool &ok) {
    const double l = a[0] * inv;
    const double lm = lane < C ? 0.0 : l; // rows above the pivot (upper triangle of the block)
    st_shared_f64(pan_s + (uint32_t)(C * pcs + lane) * 8u, lm);
    if (lane == C) st_shared_f64(iv_s + C * 8u, inv);
    const double pc = fma(-l, l, a[1]); // the next pivot, valid on its own lane (row C+1)
    const double dn = shfl_f64_volatile(pc, (C + 1) & 31);
    __syncwarp();
    // rest of column C of the stacked panel: entries C+1 .. C+15 (those past row 15 feed registers that are already dead)
    const uint32_t vb = pan_s + (uint32_t)(C * pcs + C) * 8u;
    double v[16];
    if (C & 1) {
#pragma unroll
        for (int j = 1; j < 16; j += 2) {
            const double2 w = ld_shared_v2f64(vb + j * 8u);
            v[j] = w.x;
            if (j + 1 < 16) v[j + 1] = w.y;
        }
    } else {
        v[1] = ld_shared_f64(vb + 8u);
#pragma unroll
        for (int j = 2; j < 16; j += 2) {
            const double2 w = ld_shared_v2f64(vb + j * 8u);
            v[j] = w.x;
            if (j + 1 < 16) v[j + 1] = w.y;
        }
    }
    // columns C-3 .. C of [L_kk ; P_(k+1,k)] and their 1/diag are visible: the row solves stream behind, four columns at a
    // time (an arrive per pivot cost the chain 29 cycles per pivot), issued after this pivot's own loads
    if ((C & 3) == 3 && lane == 0) mbar_arrive_cta(colbar + (C >> 2));
    if (C < 15) ok = ok && (dn > 0.0) && (dn < 1e300);
    const double invn = band_rsqrt(dn);
#pragma unroll
    for (int j = 1; j < 16 - (C >= 8 ? C - 7 : 0); j++) a[j - 1] = fma(-lm, v[j], a[j]); // columns past the block end are dead
    d = dn;
    inv = invn;
}

SDV_DEV void band_chain_step(double (&a)[16], int lane, double *pan, int pcs, double *iv, uint64_t *colbar, bool &ok) {
    const uint32_t pan_s = smem_u32(pan), iv_s = smem_u32(iv);
    double d = __shfl_sync(FULL, a[0], 0);
    ok = ok && (d > 0.0) && (d < 1e300);
    double inv = band_rsqrt(d);
    band_chain_pivot<0>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<1>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<2>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<3>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<4>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<5>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<6>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<7>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<8>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<9>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<10>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<11>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<12>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<13>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<14>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
    band_chain_pivot<15>(a, lane, pan_s, iv_s, colbar, pcs, d, inv, ok);
}

// X <- X L_kk^-T for one row per lane (t = the row), L_kk read from the transposed panel (column c at pan[c * pcs ..]);
// element c of the result goes to dst[c * dstride].  Fully unrolled (the registers rotate, so every index is static) and
// triangular: measured 98 cycles per pivot as a 2-pivot loop (nothing overlaps across the loop edge), see tools/micro.
// colbar != nullptr: STREAMING — column c is consumed as soon as the chain warp has published it (mbarrier per column), so
// the solve ends a few dozen cycles after the factorisation of the diagonal block instead of ~1 k cycles later.
// UPD (block row k+2 only): the update W_(k+2,k+1) -= P_(k+2,k) P_(k+1,k)^T is applied on the fly as rank-1 updates with the
// columns of P_(k+1,k) the chain publishes (w = 8 entries of this lane's row, columns 8 h .. 8 h + 7; the two half-warps
// duplicate the solve and split the update), so the block the chain needs next is final right behind the factorisation.
template <bool UPD>
SDV_DEV void band_trsm16(double (&t)[16], double (&w)[8], int h, const double *pan, int pcs, const double *iv, double *dst, int dstride,
                         double *grow = nullptr, uint64_t *colbar = nullptr, unsigned parity = 0) {
    double xs[16]; // results are stored after the loop: a store into the panel in between would fence the loads behind it
#pragma unroll
    for (int c = 0; c < 16; c++) {
        if (colbar && (c & 3) == 0) mbar_wait_cta(colbar + (c >> 2), parity); // columns c .. c+3
        const double x = t[0] * iv[c];
        xs[c] = x;
        const double *v = pan + c * pcs + c;
        const int j0 = ((c + 1) & 1) ? 2 : 1, jn = 16 - c; // j0: first j with an even (16-byte aligned) entry index c + j
        if (j0 == 2 && 1 < jn) t[0] = fma(-x, v[1], t[1]);
#pragma unroll
        for (int j = j0; j + 1 < jn; j += 2) {
            const double2 q = *reinterpret_cast<const double2 *>(v + j);
            t[j - 1] = fma(-x, q.x, t[j]);
            t[j] = fma(-x, q.y, t[j + 1]);
        }
        if (jn > j0 && ((jn - j0) & 1)) t[jn - 2] = fma(-x, v[jn - 1], t[jn - 1]);
        if (UPD) {
            const double *p1 = pan + c * pcs + 16 + 8 * h; // P_(k+1,k)[8h .. 8h+7][c]
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                const double2 q = *reinterpret_cast<const double2 *>(p1 + j);
                w[j] = fma(-x, q.x, w[j]);
                w[j + 1] = fma(-x, q.y, w[j + 1]);
            }
        }
    }
    if (dst)
#pragma unroll
        for (int c = 0; c < 16; c++) dst[c * dstride] = xs[c];
    if (grow) // the same row, row-major, into the global band storage of the factor
#pragma unroll
        for (int c = 0; c < 16; c += 2) *reinterpret_cast<double2 *>(grow + c) = make_double2(xs[c], xs[c + 1]);
}

// Streaming block row at distance D (2 or 3) from the pivot block (BAND_STREAM_UPDATES): the solve P_D = W_(k+D,k) L_kk^-T and
// ALL updates of that row in this step, W_(k+D,k+1+b) -= P_D P_(1+b)^T for b = 0..D-1, applied as rank-1 updates behind the
// chain, one pivot column at a time.  Lane (r, hh): row r; the two half-warps duplicate the solve and split the 16 columns of
// every updated block (w[b] = columns 8 hh .. 8 hh + 7 of block b).  Column c of P_D is published right away (the diagonal
// block b = D-1 needs it from all 16 rows, and the next streaming row needs it too).
template <int D>
SDV_DEV void band_stream_row(double (&t)[16], double (&w)[D][8], int r, int hh, int lane, double *pan, int pcs, const double *iv,
                             uint64_t *colbar, uint64_t *colbar_prev, uint64_t *colbar_own, unsigned parity) {
#pragma unroll
    for (int c = 0; c < 16; c++) {
        if ((c & 3) == 0) {
            mbar_wait_cta(colbar + (c >> 2), parity);                       // columns c .. c+3 of L_kk, P_1 (chain)
            if (colbar_prev) mbar_wait_cta(colbar_prev + (c >> 2), parity); // ... and of P_(D-1) (the streaming row before this one)
        }
        const double x = t[0] * iv[c];
        double *col = pan + c * pcs;
        if (hh == 0) col[16 * D + r] = x;
        const double *v = col + c;
        const int j0 = ((c + 1) & 1) ? 2 : 1, jn = 16 - c;
        if (j0 == 2 && 1 < jn) t[0] = fma(-x, v[1], t[1]);
#pragma unroll
        for (int j = j0; j + 1 < jn; j += 2) {
            const double2 q = *reinterpret_cast<const double2 *>(v + j);
            t[j - 1] = fma(-x, q.x, t[j]);
            t[j] = fma(-x, q.y, t[j + 1]);
        }
        if (jn > j0 && ((jn - j0) & 1)) t[jn - 2] = fma(-x, v[jn - 1], t[jn - 1]);
        __syncwarp(); // column c of P_D is complete
#pragma unroll
        for (int b = 0; b < D; b++) {
            const double *pb = col + 16 * (b + 1) + 8 * hh;
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                const double2 q = *reinterpret_cast<const double2 *>(pb + j);
                w[b][j] = fma(-x, q.x, w[b][j]);
                w[b][j + 1] = fma(-x, q.y, w[b][j + 1]);
            }
        }
        if (colbar_own && (c & 3) == 3 && lane == 0) mbar_arrive_cta(colbar_own + (c >> 2));
    }
}

// C (16 x 16 window block, row stride WSTR) -= P_i P_j^T, operands = blocks of the transposed panel; one warp
SDV_DEV void band_update_dmma(double *C, const double *PTi, const double *PTj, int pcs, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double c[2][2][2];
#pragma unroll
    for (int ib = 0; ib < 2; ib++)
#pragma unroll
        for (int jb = 0; jb < 2; jb++) {
            const double2 v = *reinterpret_cast<const double2 *>(C + (ib * 8 + g) * WSTR + jb * 8 + 2 * t);
            c[ib][jb][0] = v.x;
            c[ib][jb][1] = v.y;
        }
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
        double av[2], bv[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            av[q] = -PTi[(kk * 4 + t) * pcs + q * 8 + g];
            bv[q] = PTj[(kk * 4 + t) * pcs + q * 8 + g];
        }
#pragma unroll
        for (int ib = 0; ib < 2; ib++)
#pragma unroll
            for (int jb = 0; jb < 2; jb++) dmma(c[ib][jb][0], c[ib][jb][1], av[ib], bv[jb]);
    }
#pragma unroll
    for (int ib = 0; ib < 2; ib++)
#pragma unroll
        for (int jb = 0; jb < 2; jb++)
            *reinterpret_cast<double2 *>(C + (ib * 8 + g) * WSTR + jb * 8 + 2 * t) = make_double2(c[ib][jb][0], c[ib][jb][1]);
}

SDV_DEV void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
SDV_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#if SDV_BAND_REV || SDV_BAND_BABE
SDV_DEV void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
#endif
#if SDV_BAND_BABE
// one double from the shared memory of CTA `rank` of the cluster (same offset as `local_addr` in this CTA)
SDV_DEV double dsmem_load(const double *local_addr, unsigned rank) {
    unsigned local = smem_u32(local_addr), remote;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
    return v;
}
#endif

#if SDV_BAND_BACKWARD_V2
// (tools/micro/backward.cu times it alone): one FULL block step of the
// backward solve (all BW sub-diagonal blocks present) with every load of the step issued up front and the d = 1 block — the
// only one whose x was produced by the previous step — closing the FMA chains.  Returns x_k[c] in both half-warps.
template <int BW> SDV_DEV double band_backward_full_step(const double *sb, const double *gs, double *rvs, int k, int lane) {
    const int c = lane & 15, hh = lane >> 4;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const double *Mi = sb + (BW + 1) * 256 + hh * 128 + c;
    double m[8];
#pragma unroll
    for (int q = 0; q < 8; q++) m[q] = Mi[16 * q];
#pragma unroll
    for (int d = BW; d >= 2; d--) {
        const double *Lc = sb + d * 256 + hh * 128 + c;
        const double2 *x2 = reinterpret_cast<const double2 *>(gs + (k + d) * BN + hh * 8);
        const double2 xa = x2[0], xb = x2[1], xc = x2[2], xd = x2[3];
        s0 = fma(Lc[0], xa.x, s0);
        s1 = fma(Lc[16], xa.y, s1);
        s2 = fma(Lc[32], xb.x, s2);
        s3 = fma(Lc[48], xb.y, s3);
        s0 = fma(Lc[64], xc.x, s0);
        s1 = fma(Lc[80], xc.y, s1);
        s2 = fma(Lc[96], xd.x, s2);
        s3 = fma(Lc[112], xd.y, s3);
    }
    {
        const double *Lc = sb + 256 + hh * 128 + c;
        double l[8];
#pragma unroll
        for (int q = 0; q < 8; q++) l[q] = Lc[16 * q];
        const double2 *x2 = reinterpret_cast<const double2 *>(gs + (k + 1) * BN + hh * 8);
        const double2 xa = x2[0], xb = x2[1], xc = x2[2], xd = x2[3];
        s0 = fma(l[0], xa.x, s0);
        s1 = fma(l[1], xa.y, s1);
        s2 = fma(l[2], xb.x, s2);
        s3 = fma(l[3], xb.y, s3);
        s0 = fma(l[4], xc.x, s0);
        s1 = fma(l[5], xc.y, s1);
        s2 = fma(l[6], xd.x, s2);
        s3 = fma(l[7], xd.y, s3);
    }
    double sum = (s0 + s1) + (s2 + s3);
    sum += __shfl_xor_sync(FULL, sum, 16);
    const double rv = gs[k * BN + c] - sum;
    if (hh == 0) rvs[c] = rv;
    __syncwarp();
    const double2 *r2 = reinterpret_cast<const double2 *>(rvs + hh * 8);
    const double2 ra = r2[0], rb = r2[1], rc = r2[2], rd = r2[3];
    double x0 = m[0] * ra.x, x1 = m[1] * ra.y, x2v = m[2] * rb.x, x3 = m[3] * rb.y;
    x0 = fma(m[4], rc.x, x0);
    x1 = fma(m[5], rc.y, x1);
    x2v = fma(m[6], rd.x, x2v);
    x3 = fma(m[7], rd.y, x3);
    double x = (x0 + x1) + (x2v + x3);
    x += __shfl_xor_sync(FULL, x, 16);
    return x;
}
#endif

#if SDV_BAND_BACKWARD_V3
// This lane's partial (rows 8 hh .. 8 hh + 7) of sum_{d = 2 .. BW} (L_(k+d,k))^T x_(k+d), column c, from the ring stage `sb` of block column k.
template <int BW> SDV_DEV double band_backward_partial(const double *sb, const double *gs, int k, int lane) {
    const int c = lane & 15, hh = lane >> 4;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int d = BW; d >= 2; d--) {
        const double *Lc = sb + d * 256 + hh * 128 + c;
        const double2 *x2 = reinterpret_cast<const double2 *>(gs + (k + d) * BN + hh * 8);
        const double2 xa = x2[0], xb = x2[1], xc = x2[2], xd = x2[3];
        s0 = fma(Lc[0], xa.x, s0);
        s1 = fma(Lc[16], xa.y, s1);
        s2 = fma(Lc[32], xb.x, s2);
        s3 = fma(Lc[48], xb.y, s3);
        s0 = fma(Lc[64], xc.x, s0);
        s1 = fma(Lc[80], xc.y, s1);
        s2 = fma(Lc[96], xd.x, s2);
        s3 = fma(Lc[112], xd.y, s3);
    }
    return (s0 + s1) + (s2 + s3);
}
// One FULL block step k with qp = band_backward_partial of column k already known; returns x_k[c] in both half-warps and — NEXT —
// leaves the partial of column k-1 (stage `sbn`) in qp: its d = 2 block multiplies x_(k+1), which this step loads anyway, the
// blocks d >= 3 fill the two latency holes of the chain (after the reduction shuffle, after the trip of rv through shared memory).
template <int BW, bool NEXT> SDV_DEV double band_backward_pipe_step(const double *sb, const double *sbn, const double *gs, double *rvs, int k, int lane, double &qp) {
    const int c = lane & 15, hh = lane >> 4;
    const double *Lc = sb + 256 + hh * 128 + c, *Mi = sb + (BW + 1) * 256 + hh * 128 + c;
    double l[8], m[8];
#pragma unroll
    for (int q = 0; q < 8; q++) l[q] = Lc[16 * q];
    const double2 *x2 = reinterpret_cast<const double2 *>(gs + (k + 1) * BN + hh * 8);
    const double2 xa = x2[0], xb = x2[1], xc = x2[2], xd = x2[3];
    const double yk = gs[k * BN + c];
    double s0 = fma(l[0], xa.x, qp), s1 = l[1] * xa.y, s2 = l[2] * xb.x, s3 = l[3] * xb.y;
    s0 = fma(l[4], xc.x, s0);
    s1 = fma(l[5], xc.y, s1);
    s2 = fma(l[6], xd.x, s2);
    s3 = fma(l[7], xd.y, s3);
    double sum = (s0 + s1) + (s2 + s3);
    sum += __shfl_xor_sync(FULL, sum, 16);
    double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
    if (NEXT) { // column k-1, d = 2: x_(k+1) is in registers
        const double *Ln = sbn + 2 * 256 + hh * 128 + c;
        n0 = Ln[0] * xa.x;
        n1 = Ln[16] * xa.y;
        n2 = Ln[32] * xb.x;
        n3 = Ln[48] * xb.y;
        n0 = fma(Ln[64], xc.x, n0);
        n1 = fma(Ln[80], xc.y, n1);
        n2 = fma(Ln[96], xd.x, n2);
        n3 = fma(Ln[112], xd.y, n3);
    }
#pragma unroll
    for (int q = 0; q < 8; q++) m[q] = Mi[16 * q];
    const double rv = yk - sum;
    if (hh == 0) rvs[c] = rv;
    __syncwarp();
    const double2 *r2 = reinterpret_cast<const double2 *>(rvs + hh * 8);
    const double2 ra = r2[0], rb = r2[1], rc = r2[2], rd = r2[3];
    if (NEXT) { // column k-1, d = 3 .. BW: x_(k+2) .. x_(k+BW-1)
#pragma unroll
        for (int d = 3; d <= BW; d++) {
            const double *Ln = sbn + d * 256 + hh * 128 + c;
            const double2 *y2 = reinterpret_cast<const double2 *>(gs + (k - 1 + d) * BN + hh * 8);
            const double2 ya = y2[0], yb = y2[1], yc = y2[2], yd = y2[3];
            n0 = fma(Ln[0], ya.x, n0);
            n1 = fma(Ln[16], ya.y, n1);
            n2 = fma(Ln[32], yb.x, n2);
            n3 = fma(Ln[48], yb.y, n3);
            n0 = fma(Ln[64], yc.x, n0);
            n1 = fma(Ln[80], yc.y, n1);
            n2 = fma(Ln[96], yd.x, n2);
            n3 = fma(Ln[112], yd.y, n3);
        }
    }
    double x0 = m[0] * ra.x, x1 = m[1] * ra.y, x2v = m[2] * rb.x, x3 = m[3] * rb.y;
    x0 = fma(m[4], rc.x, x0);
    x1 = fma(m[5], rc.y, x1);
    x2v = fma(m[6], rd.x, x2v);
    x3 = fma(m[7], rd.y, x3);
    double x = (x0 + x1) + (x2v + x3);
    x += __shfl_xor_sync(FULL, x, 16);
    qp = (n0 + n1) + (n2 + n3);
    return x;
}
#endif

// System preparation for k_chol_band (same arithmetic as k_sysprep): gradient-tolerance test, Jacobi column scales at
// iteration 0, LM damping -> dmp (negative = padding column), right-hand side -> gs.  Returns true (uniformly) when the
// gradient tolerance terminates the solve.  Not inlined: its square roots and divisions would otherwise raise the register
// pressure of the factorisation loops (the kernel is capped at 128 registers).
__device__ __noinline__ bool band_sysprep(int n, int n_pad, int ld, LMState *st, Accum *acc, int jacobi_scaling, double gradient_tolerance,
                                          double min_diag, double max_diag, const double *A, double *scale_p, double *damp_p, double *graw_p,
                                          double *gs, double *dmp
#if SDV_BAND_BABE
                                          , bool babe, bool rev, int nloc, int nsep0 // cluster mode, reversed CTA, local rows, first separator row
#endif
) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *g = A + (size_t)n_pad * ld, *cdiag = g + ld, *graw = cdiag + ld;
    const bool first = st->scaling_done == 0, check = st->need_grad_check != 0;
    const double radius = st->radius;
    if (check) { // gradient tolerance (Ceres: iteration 0 and after every successful step)
        double m = 0.0;
        for (int i = threadIdx.x; i < n; i += BCT) m = fmax(m, fabs(graw[i]));
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
        if (lane == 0) dmp[warp] = m;
    }
    __syncthreads(); // also: every thread has read the state words before thread 0 changes them
    if (check) {
        double mm = __longlong_as_double((long long)acc->grad_max_bits);
        for (int i = 0; i < BNW; i++) mm = fmax(mm, dmp[i]);
        if (mm <= gradient_tolerance) { // uniform: every thread evaluates the same values
            __syncthreads();
#if SDV_BAND_BABE
            if (babe) cluster_sync_all(); // the other CTA has read the state words too
            if (threadIdx.x == 0 && !rev) {
#else
            if (threadIdx.x == 0) {
#endif
                st->status = 1 + 2; // SDV_TERM_GRADIENT_TOLERANCE
                st->iter -= 1;      // the step Ceres never starts was already counted
                st->need_grad_check = 0;
            }
            return true;
        }
    }
    __syncthreads();
#if SDV_BAND_BABE
    // every thread of both CTAs has read the state words and the gradient maximum: only now may CTA 0 change them
    if (babe) cluster_sync_all();
    for (int i = threadIdx.x; i < n_pad; i += BCT) {
        const int il = rev ? n_pad - 1 - i : i; // row of this CTA's (possibly reversed) system
        const bool mine = !babe || il < nloc, sep = babe && rev && il >= nsep0;
        double d = -1.0, gi = 0.0; // padding: identity block, zero right-hand side
        if (i < n) {
            const double c = cdiag[i];
            const double sc = first ? (jacobi_scaling ? 1.0 / (1.0 + sqrt(c)) : 1.0) : scale_p[i];
            if (first && !rev) scale_p[i] = sc;
            d = fmin(fmax(sc * sc * c, min_diag), max_diag) / (radius * sc * sc); // lm_damping()
            gi = g[i];
        }
        if (!rev) {
            damp_p[i] = i < n ? d : 0.0;
            graw_p[i] = i < n ? graw[i] : 0.0;
        }
        if (mine) {
            dmp[il] = sep ? 0.0 : d; // CTA 1 carries only its UPDATE of the separator: no damping, no right-hand side there
            gs[il] = sep ? 0.0 : gi;
        }
    }
    if (threadIdx.x == 0 && !rev) {
        st->need_grad_check = 0;
        st->scaling_done = 1;
        acc->grad_max_bits = 0ull;
    }
    return false;
#elif SDV_BAND_REV
    for (int i = threadIdx.x; i < n_pad; i += BCT) { // shared-memory copies (dmp, gs) in the reversed order, globals as they are
        const int il = n_pad - 1 - i;
        if (i < n) {
            const double c = cdiag[i];
            const double sc = first ? (jacobi_scaling ? 1.0 / (1.0 + sqrt(c)) : 1.0) : scale_p[i];
            if (first) scale_p[i] = sc;
            const double d = fmin(fmax(sc * sc * c, min_diag), max_diag) / (radius * sc * sc); // lm_damping()
            dmp[il] = d;
            damp_p[i] = d;
            graw_p[i] = graw[i];
            gs[il] = g[i];
        } else {
            dmp[il] = -1.0;
            damp_p[i] = 0.0;
            graw_p[i] = 0.0;
            gs[il] = 0.0;
        }
    }
#else
    for (int i = threadIdx.x; i < n_pad; i += BCT) {
        if (i < n) {
            const double c = cdiag[i];
            const double sc = first ? (jacobi_scaling ? 1.0 / (1.0 + sqrt(c)) : 1.0) : scale_p[i];
            if (first) scale_p[i] = sc;
            const double d = fmin(fmax(sc * sc * c, min_diag), max_diag) / (radius * sc * sc); // lm_damping()
            dmp[i] = d;
            damp_p[i] = d;
            graw_p[i] = graw[i];
            gs[i] = g[i];
        } else {
            dmp[i] = -1.0; // padding: identity block, zero right-hand side
            damp_p[i] = 0.0;
            graw_p[i] = 0.0;
            gs[i] = 0.0;
        }
    }
#endif
    if (threadIdx.x == 0) {
        st->need_grad_check = 0;
        st->scaling_done = 1;
        acc->grad_max_bits = 0ull;
    }
    return false;
}

// A : (n_pad + 32) x ld reduced system (lower triangle, right-hand side in row n_pad), read only.
// Lb: band storage of the factor, block column k at Lb + k (bw + 2) 256: [L_kk | L_(k+1,k) .. L_(k+bw,k) | L_kk^-1], each
//     block 16 x 16 row-major.  dxp receives -x (S delta = -g).
// The preparation of the reduced system (k_sysprep for the other variants: gradient-tolerance test, Jacobi column scales at
// iteration 0, LM damping of the pose/velocity/bias columns, identity padding) is done here, on the way into shared memory.
__global__ void __launch_bounds__(BCT, 1) k_chol_band(const DevProblem *__restrict__ Pg, LinBuf B0, LinBuf B1, LMState *st, Accum *acc, SolverOpts opt, const double *A,
                                                      double *Lb, double *scale_p, double *damp_p, double *graw_p, double *dxp, double *prof) {
    const DevProblem &P = *Pg; // device-resident problem description: the launch parameters do not depend on the window (one CUDA graph serves them all)
    if (st->status != 0) return;
    extern __shared__ __align__(16) double bsm[];
    __shared__ uint64_t full[BAND_MAX_STAGES], bar_panel[2], bar_step[2], bar_rhs[2], bar_copy[2], bar_p[2][8], bar_c1[2][8], bar_c2[2][8], colbar[2][16], colbar2[2][4], bar_r3[2];
    __shared__ int s_fail;
    const BandPlan pl = band_plan(P.n_pad, P.band_bw);
#if SDV_BAND_BABE
    __shared__ uint64_t bar_xsep; // CTA 1: the separator unknowns have arrived from CTA 0
    const bool babe = cluster_size() == 2;
    const bool rev = babe && cluster_rank() == 1;
    const int bw = pl.bw, R = pl.R, ld = P.ld, pcs = pl.pcs, nbg = pl.nb;
#ifndef SDV_BAND_SKEW
#define SDV_BAND_SKEW 0
#endif
    const int nl = (nbg - bw + 1) / 2 + SDV_BAND_SKEW, nr = nbg - bw - nl;  // interior block rows of CTA 0 / CTA 1, separator = bw block rows
    const int nint = babe ? (rev ? nr : nl) : nbg;          // block steps this CTA runs before the hand-over
    const int nb = babe ? nint + bw : nbg;                  // block rows of this CTA's system (interior + separator)
    if (rev) Lb += (size_t)nbg * (bw + 2) * 256;            // CTA 1 keeps its factor in the second half of the band storage
#else
    const int nb = pl.nb, bw = pl.bw, R = pl.R, ld = P.ld, pcs = pl.pcs;
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *win = bsm + pl.o_win, *pan0 = bsm + pl.o_pan, *invs = bsm + pl.o_inv, *gs = bsm + pl.o_g, *dmp = bsm + pl.o_dmp;
    const int bwp = bw + 1;
    // window block (i, j) lives in slot (i mod R, j mod (bw+1)); kr = k mod R and kc = k mod (bw+1) are carried along so that
    // no integer division is executed inside the factorisation
    int kr = 0, kc = 0;
    auto Wk = [&](int di, int dj) { // block (k + di, k + dj), 0 <= dj <= di <= bw + 1
        int rs = kr + di, cs = kc + dj;
        rs -= rs >= R ? R : 0;
        cs -= cs >= bwp ? bwp : 0;
        return win + (rs * bwp + cs) * WBLK;
    };
#ifdef SDV_BAND_PROF
    auto rdclk = [] { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; };
    long long tp[6] = {0, 0, 0, 0, 0, 0}, tc = rdclk(), tn;
    long long tw[5] = {0, 0, 0, 0, 0}; // chain warp: its wait split by barrier [step, rhs, copy, (1,1), (2,1)]
#define BAND_TICK(q) do { tn = rdclk(); tp[q] += tn - tc; tc = tn; } while (0)
#define BAND_BUSY(q) do { tb[q] += rdclk() - tc; } while (0)
    long long tbk[4] = {0, 0, 0, 0}, tbc = 0, t_xsep = 0; // backward solve, warp 0: [stage wait, compute, publish, refill]
#define BAND_TICKB(q) do { long long t_ = rdclk(); tbk[q] += t_ - tbc; tbc = t_; } while (0)
#else
#define BAND_TICK(q) do { } while (0)
#define BAND_BUSY(q) do { } while (0)
#define BAND_TICKB(q) do { } while (0)
#endif

    // block row i of S (blocks max(0, i - bw) .. i) -> window, 16-byte cp.async chunks; threads [t0, t0 + nt)
    auto load_row = [&](int i, int rs, int c0, int t0, int nt, int tid = -1) { // rs = i mod R, c0 = max(0, i - bw) mod (bw + 1)
        if (tid < 0) tid = (int)threadIdx.x;
        const int j0 = i - bw > 0 ? i - bw : 0;
        const int nchunk = (i - j0 + 1) * 128;
#pragma unroll 1
        for (int e = tid - t0; e < nchunk; e += nt) {
            const int jj = e >> 7, r = (e >> 3) & 15, q = e & 7;
            int cs = c0 + jj;
            cs -= cs >= bwp ? bwp : 0;
#if SDV_BAND_BABE
            if (rev) {
                double *dst = win + (rs * bwp + cs) * WBLK + r * WSTR + 2 * q;
                if (i >= nint && j0 + jj >= nint) { // separator x separator: CTA 1 accumulates only its update there
                    dst[0] = 0.0;
                    dst[1] = 0.0;
                } else {
                    const int np1 = P.n_pad - 1;
#pragma unroll
                    for (int h = 0; h < 2; h++) { // element (p, c) of P S P = element (n_pad-1-p, n_pad-1-c) of S, from the stored triangle
                        const int gr = np1 - (i * BN + r), gc = np1 - ((j0 + jj) * BN + 2 * q + h);
                        const int hi = gr > gc ? gr : gc, lo = gr > gc ? gc : gr;
                        cp_async8(dst + h, A + (size_t)hi * ld + lo);
                    }
                }
            } else
                cp_async16(win + (rs * bwp + cs) * WBLK + r * WSTR + 2 * q, A + (size_t)(i * BN + r) * ld + (j0 + jj) * BN + 2 * q);
#elif SDV_BAND_REV
            // element (p, c) of P S P is element (n_pad-1-p, n_pad-1-c) of S, read from the stored (lower) triangle
            const int np1 = P.n_pad - 1;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int gr = np1 - (i * BN + r), gc = np1 - ((j0 + jj) * BN + 2 * q + h);
                const int hi = gr > gc ? gr : gc, lo = gr > gc ? gc : gr;
                cp_async8(win + (rs * bwp + cs) * WBLK + r * WSTR + 2 * q + h, A + (size_t)hi * ld + lo);
            }
#else
            cp_async16(win + (rs * bwp + cs) * WBLK + r * WSTR + 2 * q, A + (size_t)(i * BN + r) * ld + (j0 + jj) * BN + 2 * q);
#endif
        }
    };
    // after the copies of block row i have landed: add the damping to the diagonal (unit diagonal for padding columns); every
    // thread patches the 16-byte chunks it copied itself, so no extra synchronisation is needed
    auto fix_row = [&](int i, int rs, int c0, int t0, int nt, int tid = -1) {
        if (tid < 0) tid = (int)threadIdx.x;
        const int j0 = i - bw > 0 ? i - bw : 0;
        int cs = c0 + (i - j0);
        cs -= cs >= bwp ? bwp : 0;
        double *D = win + (rs * bwp + cs) * WBLK; // diagonal block (i, i): its chunks are e = (i - j0) * 128 + r * 8 + q
        for (int e = tid - t0; e < (i - j0 + 1) * 128; e += nt) {
            if ((e >> 7) != i - j0) continue;
            const int r = (e >> 3) & 15, q = e & 7;
            if (q != (r >> 1)) continue;
            const double d = dmp[i * BN + r];
            double *x = D + r * WSTR + r;
            *x = d < 0.0 ? 1.0 : *x + d;
        }
    };
    // ---- roles (warp-specialised dataflow; every warp runs ONE small loop, the hand-offs are mbarriers in shared memory).
    // Warp w runs on SM sub-partition w % 4, each with its own FP64 pipe: the FP64-heavy roles are spread over the four.
    //   warp 0                    chain        factor D_k, solve P_(k+1,k)                      -> bar_panel
    //   warps 1,2,3,7,11,15       row solve    P_(k+d,k) = W_(k+d,k) L_kk^-T, d = 2..7 (if <= bw) -> bar_p[d]
    //   warp 5                    right-hand side  y_k, g_(k+d) -= P_d y_k                        -> bar_rhs
    //   warp 6                    inverse of L_kk (for the backward solve)                       -> bar_copy
    //   up to 6 of 9,10,13,14,7,11,15  updates  W_(k+di,k+dj) -= P_di P_dj^T (tensor-core MMAs) -> bar_c1, bar_step
    //   the rest of those         copy         panel blocks 0,1 -> global, prefetch of block row k+bw+2 -> bar_copy
    //   warps 4, 8, 12            idle (they share the chain's sub-partition)
    // A CTA-wide barrier per phase was tried first: a warp that sleeps at __syncthreads() while the chain warp runs took
    // 1.5-2 k cycles to get going again (per-warp clock64 traces), 3 times per step.
    constexpr int N_COPY = (BAND_STREAM_UPDATES || BAND_STREAM1) ? 3 : 4;
    int role = 6, ridx = 0, n_upd = 0; // 0 chain, 1 row solve (ridx = d), 2 rhs, 3 inverse, 4 update (ridx = u), 5 copy, 6 idle, 7 diagonal update
    {
        // Warp w runs on SM sub-partition w % 4, each with its own (narrow: 16 lanes) FP64 pipe and instruction cache.
        //  * sub-partition 0 belongs to the chain warp alone (warps 4, 8, 12 idle): with three other loops next to it the
        //    ~10 KB of unrolled pivot code ran 45 % slower than alone (tools/micro/roles.cu vs the in-kernel trace);
        //  * the row solves d = 2, 3 — the updates (2,1), (3,1) that gate the next steps wait for them — share their
        //    sub-partitions only with update warps, which are idle while the solves run;
        //  * everything with slack (d = 4 solve, right-hand side, inverse) sits on sub-partition 3; the copy duty (panel ->
        //    global, prefetch of the next block row) is shared by the update warps, which wait most of the time.
        const int solve_warp[6] = {1, 2, 3, 13, 14, 10}; // d = 2 .. 7
        const int upd_pool[6] = {5, 9, 6, 10, 13, 14};   // minus the ones used as row-solve warps (bw > 4)
        if (warp == 0) role = 0;
        else if (warp == 7) role = 2;
        else if (warp == 11) role = 3;
        else if (warp == 15 && (BAND_STREAM_UPDATES || BAND_STREAM1)) role = 7; // streaming update of the next diagonal block
        else if (warp == 15 || (warp & 3) == 0) { // 4, 8, 12 (, 15): light (copies), so the chain's sub-partition can host three of them
            role = 5;
            ridx = warp == 15 ? 3 : (warp >> 2) - 1;
        }
        n_upd = bw <= 4 ? 6 : 10 - bw;
#pragma unroll
        for (int q = 0; q < 6; q++) {
            if (q < n_upd && warp == upd_pool[q]) {
                role = 4;
                ridx = q;
            }
            if (q + 2 <= bw && warp == solve_warp[q]) {
                role = 1;
                ridx = q + 2;
            }
        }
    }
    if (threadIdx.x == 0) {
        s_fail = 0;
        for (int s = 0; s < BAND_MAX_STAGES; s++) mbar_init(&full[s], 1);
        for (int q = 0; q < 2; q++) {
            mbar_init(&bar_panel[q], 1);
            mbar_init(&bar_step[q], n_upd);
            mbar_init(&bar_rhs[q], 1);
            mbar_init(&bar_copy[q], N_COPY + 1); // copy warps + the inverse warp
            mbar_init(&bar_r3[q], 1);
            for (int c = 0; c < 4; c++) mbar_init(&colbar2[q][c], 1);
            for (int c = 0; c < 16; c++) mbar_init(&colbar[q][c], 1);
            for (int d = 0; d < 8; d++) {
                mbar_init(&bar_p[q][d], 1);
                mbar_init(&bar_c1[q][d], 1);
                mbar_init(&bar_c2[q][d], 1);
            }
        }
    }
    for (int i = 0; i < nb && i < R; i++) load_row(i, i, i > bw ? (i - bw) % bwp : 0, 0, BCT); // rows 0 .. R-1 (R = bw + 3, or the whole matrix)
    for (int e = threadIdx.x; e < 2 * pl.pan_doubles; e += BCT) pan0[e] = 0.0;
#if SDV_BAND_BABE
    if (threadIdx.x == 0) mbar_init(&bar_xsep, 1);
    if (band_sysprep(P.n, P.n_pad, ld, st, acc, opt.jacobi_scaling, opt.gradient_tolerance, opt.min_diag, opt.max_diag, A, scale_p, damp_p, graw_p, gs, dmp,
                     babe, rev, nb * BN, nint * BN)) {
#else
    if (band_sysprep(P.n, P.n_pad, ld, st, acc, opt.jacobi_scaling, opt.gradient_tolerance, opt.min_diag, opt.max_diag, A, scale_p, damp_p, graw_p, gs, dmp)) { // gradient tolerance reached: no step
#endif
        cp_async_wait_all();
        return;
    }
    cp_async_wait_all();
    __syncthreads();
    for (int i = 0; i < nb && i < R; i++) fix_row(i, i, i > bw ? (i - bw) % bwp : 0, 0, BCT);
    __syncthreads();
    auto ph = [](int k) { return (unsigned)((k >> 1) & 1); }; // phase parity of the barriers of step k (each is reused every 2 steps)
    auto next_k = [&] {
        if (++kr == R) kr = 0;
        if (++kc == bwp) kc = 0;
    };

    // copy duty of step k, shared by the copy warps: blocks 0 and 1
    // of panel k -> global band storage (transposing), then the prefetch of block row k+bw+2 into the window slot that row
    // k-1 has left (last read by the chain of step k-1), with the damping added on the way in
    auto copy_share = [&](int k, int ct, int nct) {
        const double *pan = pan0 + (k & 1) * pl.pan_doubles;
        double *dst = Lb + (size_t)k * (bw + 2) * 256;
        const int ne = (k + 1 < nb ? 2 : 1) * 256;
#pragma unroll 1
        for (int e = ct; e < ne; e += nct) {
            const int r = e >> 4, c = e & 15; // r = 16 d + row: row of the stacked panel
            dst[e] = pan[c * pcs + r];
        }
        if (k >= 1 && k + bw + 2 < nb) {
            const int rs = kr == 0 ? R - 1 : kr - 1, c0 = kc + 2 >= bwp ? kc + 2 - bwp : kc + 2;
            load_row(k + bw + 2, rs, c0, 0, nct, ct);
            cp_async_wait_all();
            fix_row(k + bw + 2, rs, c0, 0, nct, ct);
        }
    };

    bool ok = true;
    BAND_TICK(0);
#if SDV_BAND_BABE
    // CTA 0 of the cluster runs two phases (interior, then — after adding CTA 1's contribution — the separator), everybody else one
    for (int phase = 0; phase < (babe && !rev ? 2 : 1); phase++) {
    const int kbeg = phase == 0 ? 0 : nint, kend = babe && phase == 0 ? nint : nb;
    if (phase == 1) {
        if (warp == 0 && !ok) s_fail = 1;
        __syncthreads(); // this CTA's interior steps are complete
        if (s_fail) acc->chol_fail = 1;
        __threadfence();
#ifdef SDV_BAND_PROF
        const long long th0 = rdclk();
#endif
        cluster_sync_all(); // #1: ... and so are CTA 1's; its window and right-hand side are visible through DSMEM
#ifdef SDV_BAND_PROF
        if (threadIdx.x == 0 && prof) prof[23] = (double)(rdclk() - th0); // time CTA 0 waited for CTA 1 at the hand-over
        const long long th1 = rdclk();
#endif
        if (__ldcg(&acc->chol_fail) != 0 || __ldcg(&acc->schur_fail) != 0) { // uniform over the cluster (CTA 1 tests the same words after #1)
            if (threadIdx.x == 0) {
                st->step_valid = 0;
                st->model_cost_change = 0.0;
            }
            return;
        }
        // separator blocks (i, j), nint <= j <= i < nb: S_ij += (CTA 1's update of its block (nbg-1-j, nbg-1-i)) anti-transposed
        for (int e = threadIdx.x; e < bw * bw * 256; e += BCT) {
            const int a = (e >> 8) / bw, b = (e >> 8) % bw, r = (e >> 4) & 15, c = e & 15;
            if (b > a) continue;
            const int i = nint + a, j = nint + b, i1 = nbg - 1 - j, j1 = nbg - 1 - i;
            double *dst = win + ((i % R) * bwp + (j % bwp)) * WBLK + r * WSTR + c;
            const double *src = win + ((i1 % R) * bwp + (j1 % bwp)) * WBLK + (15 - c) * WSTR + (15 - r);
            *dst += dsmem_load(src, 1);
        }
        for (int t = threadIdx.x; t < bw * BN; t += BCT) gs[nint * BN + t] += dsmem_load(gs + (P.n_pad - 1 - (nint * BN + t)), 1);
        __syncthreads();
#ifdef SDV_BAND_PROF
        if (threadIdx.x == 0 && prof) prof[31] = (double)(rdclk() - th1); // separator hand-over (flags + DSMEM reads)
#endif
    }
#endif
    if (role == 0) {
        // ------------------------------------------------------------------ chain
        for (int k = BAND_KBEG; k < BAND_KEND; k++) {
            const int par = k & 1;
            const int nd = bw < nb - 1 - k ? bw : nb - 1 - k;
#ifdef SDV_BAND_PROF
#define BAND_WTICK(q) do { long long t_ = rdclk(); tw[q] += t_ - twc; twc = t_; } while (0)
            long long twc = rdclk();
#else
#define BAND_WTICK(q) do { } while (0)
#endif
            if (k >= 2) {
                mbar_wait_cta(&bar_step[par], ph(k - 2)); // every reader of this panel buffer (step k-2) is done
                BAND_WTICK(0);
                mbar_wait_cta(&bar_rhs[par], ph(k - 2));
                BAND_WTICK(1);
                mbar_wait_cta(&bar_copy[par], ph(k - 2)); // ... and it has been written out; block row k+bw is resident
                BAND_WTICK(2);
            }
            if (k >= 1) { // blocks (k,k) and (k+1,k) updated through step k-1 = tasks (1,1) and (2,1) of that step
                mbar_wait_cta(&bar_c1[par ^ 1][1], ph(k - 1));
                BAND_WTICK(3);
                if (bw >= 2 && k + 1 < nb) mbar_wait_cta(&bar_c1[par ^ 1][2], ph(k - 1)); // task (2,1) exists only if row k+1 does and bw >= 2
                BAND_WTICK(4);
            }
            BAND_TICK(1);
            double a[16];
            const double *src = (lane < 16 ? Wk(0, 0) : Wk(nd >= 1 ? 1 : 0, 0)) + (lane & 15) * WSTR; // last block: lanes 16.. redo D_k, unused
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(src + c);
                a[c] = v.x;
                a[c + 1] = v.y;
            }
            band_chain_step(a, lane, pan0 + par * pl.pan_doubles, pcs, invs + par * BN, &colbar[par][0], ok);
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&bar_panel[par]);
            BAND_TICK(2);
            next_k();
        }
    } else if (role == 1) {
        // ------------------------------------------------------------------ row solve
        const int d = ridx;
        for (int k = BAND_KBEG; BAND_KLIM k + d < nb; k++) {
            const int par = k & 1;
            double *pan = pan0 + par * pl.pan_doubles;
            const bool streaming = BAND_STREAM_UPDATES && d <= 3;
            if (!streaming) {
                // block (k+d, k) is final through step k-1 once task (d+1, 1) of that step is done; d = bw: a fresh row, resident
                // once the copy of step k-2 is done.  The pivot columns are then consumed as the chain publishes them.
                if (k >= 1 && d < bw) mbar_wait_cta(&bar_c1[par ^ 1][d + 1], ph(k - 1));
                if (k >= 2 && d == bw) mbar_wait_cta(&bar_copy[par], ph(k - 2));
            } else if (k >= 1) {
                // streaming row: ALL its blocks must be final through step k-1.  d = 2: they were row 3 of step k-1 (the other
                // streaming warp); d = 3: row 4 of step k-1 (update warps); last row of the band: fresh from the copy of step k-2
                if (d < bw) mbar_wait_cta(d == 2 ? &bar_r3[par ^ 1] : &bar_step[par ^ 1], ph(k - 1));
                else if (k >= 2) mbar_wait_cta(&bar_copy[par], ph(k - 2));
            }
            BAND_TICK(1);
            double wdum[8];
            if (streaming) {
                const int r = lane & 15, hh = lane >> 4;
                double t[16];
                const double *src = Wk(d, 0) + r * WSTR;
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(src + c);
                    t[c] = v.x;
                    t[c + 1] = v.y;
                }
                if (d == 2) {
                    double w2[2][8];
#pragma unroll
                    for (int b = 0; b < 2; b++)
#pragma unroll
                        for (int c = 0; c < 8; c++) w2[b][c] = (Wk(2, 1 + b) + r * WSTR + 8 * hh)[c];
                    band_stream_row<2>(t, w2, r, hh, lane, pan, pcs, invs + par * BN, &colbar[par][0], nullptr, bw >= 3 ? &colbar2[par][0] : nullptr, ph(k));
#pragma unroll
                    for (int b = 0; b < 2; b++)
#pragma unroll
                        for (int c = 0; c < 8; c++) (Wk(2, 1 + b) + r * WSTR + 8 * hh)[c] = w2[b][c];
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive_cta(&bar_c1[par][2]); // (2,1): the chain's next block row
                        mbar_arrive_cta(&bar_c2[par][2]); // (2,2): the next diagonal block (role 7)
                    }
                } else {
                    double w3[3][8];
#pragma unroll
                    for (int b = 0; b < 3; b++)
#pragma unroll
                        for (int c = 0; c < 8; c++) w3[b][c] = (Wk(3, 1 + b) + r * WSTR + 8 * hh)[c];
                    band_stream_row<3>(t, w3, r, hh, lane, pan, pcs, invs + par * BN, &colbar[par][0], &colbar2[par][0], nullptr, ph(k));
#pragma unroll
                    for (int b = 0; b < 3; b++)
#pragma unroll
                        for (int c = 0; c < 8; c++) (Wk(3, 1 + b) + r * WSTR + 8 * hh)[c] = w3[b][c];
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta(&bar_r3[par]); // row k+3 final through this step: next step's d = 2 row
                }
                if (hh == 0) { // the row of P_d, row-major, into the global band storage of the factor
                    double *grow = Lb + ((size_t)k * (bw + 2) + d) * 256 + r * 16;
#pragma unroll
                    for (int c = 0; c < 16; c++) grow[c] = pan[c * pcs + 16 * d + r];
                }
            } else if (BAND_STREAM1) {
                // solve + first-column update (d,1), streamed behind the chain: lane (r, h) = row r, columns 8 h .. 8 h + 7 of the update
                const int r = lane & 15, h = lane >> 4;
                double t[16], w[8];
                const double *src = Wk(d, 0) + r * WSTR;
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(src + c);
                    t[c] = v.x;
                    t[c + 1] = v.y;
                }
#pragma unroll
                for (int c = 0; c < 8; c++) w[c] = 0.0;
                band_trsm16<true>(t, w, h, pan, pcs, invs + par * BN, h == 0 ? pan + 16 * d + r : nullptr, pcs,
                                  h == 0 ? Lb + ((size_t)k * (bw + 2) + d) * 256 + r * 16 : nullptr, &colbar[par][0], ph(k));
                // W_(k+d,k+1) += w: the block is final through step k-1 once task (d+1, 2) of that step is done (a tensor-core warp);
                // d = bw: a fresh row (waited for above)
                if (k >= 1 && d < bw) mbar_wait_cta(&bar_c2[par ^ 1][d + 1], ph(k - 1));
                double *tgt = Wk(d, 1) + r * WSTR + 8 * h;
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    double2 v = *reinterpret_cast<double2 *>(tgt + c);
                    v.x += w[c];
                    v.y += w[c + 1];
                    *reinterpret_cast<double2 *>(tgt + c) = v;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_cta(&bar_c1[par][d]); // task (d,1) of this step is done
            } else if (lane < 16) {
                double t[16];
                const double *src = Wk(d, 0) + lane * WSTR;
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(src + c);
                    t[c] = v.x;
                    t[c + 1] = v.y;
                }
                band_trsm16<false>(t, wdum, 0, pan, pcs, invs + par * BN, pan + 16 * d + lane, pcs, Lb + ((size_t)k * (bw + 2) + d) * 256 + lane * 16,
                                   &colbar[par][0], ph(k));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&bar_p[par][d]);
            BAND_TICK(2);
            next_k();
        }
    } else if (role == 2) {
        // ------------------------------------------------------------------ right-hand side: L y = g rides along
        for (int k = BAND_KBEG; k < BAND_KEND; k++) {
            const int par = k & 1;
            const int nd = bw < nb - 1 - k ? bw : nb - 1 - k;
            const double *pan = pan0 + par * pl.pan_doubles;
            mbar_wait_cta(&bar_panel[par], ph(k));
            BAND_TICK(1);
            if (lane == 0) {
                double t[16];
#pragma unroll
                for (int c = 0; c < 16; c++) t[c] = gs[k * BN + c];
                double wdum[8];
                band_trsm16<false>(t, wdum, 0, pan, pcs, invs + par * BN, gs + k * BN, 1);
            }
            __syncwarp();
            const int r = lane & 15, hh = lane >> 4;
            const double *y = gs + k * BN + hh * 8;
            for (int d = 1; d <= nd; d++) {
                if (d >= 2) mbar_wait_cta(&bar_p[par][d], ph(k));
                const double *Pd = pan + hh * 8 * pcs + 16 * d + r;
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    s0 += Pd[c * pcs] * y[c];
                    s1 += Pd[(c + 1) * pcs] * y[c + 1];
                }
                double sum = s0 + s1;
                sum += __shfl_xor_sync(FULL, sum, 16);
                if (hh == 0) gs[(k + d) * BN + r] -= sum;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&bar_rhs[par]);
            BAND_TICK(2);
        }
    } else if (role == 3) {
        // ------------------------------------------------------------------ L_kk^-1 -> global (block bw+1 of column k)
        for (int k = BAND_KBEG; k < BAND_KEND; k++) {
            const int par = k & 1;
            mbar_wait_cta(&bar_panel[par], ph(k));
            BAND_TICK(1);
            if (lane < 16) {
                double t[16];
#pragma unroll
                for (int c = 0; c < 16; c++) t[c] = c == lane ? 1.0 : 0.0;
                double wdum[8];
                band_trsm16<false>(t, wdum, 0, pan0 + par * pl.pan_doubles, pcs, invs + par * BN, Lb + ((size_t)k * (bw + 2) + bw + 1) * 256 + lane, 16); // Minv[c][r] = (L^-T)[r][c]
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&bar_copy[par]);
            BAND_TICK(2);
        }
    } else if (role == 4) {
        // ------------------------------------------------------------------ trailing updates
        // Block (i, j) is updated at steps k = i-bw .. j-1, every time as task (di, dj) = (i-k, j-k).  Ownership goes with the
        // BLOCK: "virtual worker" (o, s) owns the blocks with i - j = o and j = s mod (bw - o) — exactly one task per step —
        // and virtual worker v runs on update warp v mod N_UPD.  Successive updates of a block are then program-ordered
        // inside one warp and need no barrier; only the first block column (dj = 1), which the chain and the row solves of
        // the next step read, is signalled (bar_c1).
        const int u = ridx;
        constexpr int MAXV = (BAND_MAX_BW * (BAND_MAX_BW + 1) / 2 + 2) / 3; // at least 3 update warps (bw = 7)
        int vo[MAXV], vdj[MAXV], nv = 0; // offset o and current dj of my virtual workers
        {
            int v = 0;
            for (int o = 0; o < bw; o++)
                for (int sl = 0; sl < bw - o; sl++, v++)
                    if (v % n_upd == u) {
#pragma unroll
                        for (int q = 0; q < MAXV; q++)
                            if (q == nv) {
                                vo[q] = o;
                                vdj[q] = sl == 0 ? bw - o : sl; // step 0: j in [1, bw - o], j = sl mod (bw - o)
                            }
                        nv++;
                    }
        }
#if SDV_BAND_BABE
        for (int t = 0; t < kbeg; t++) // second phase: the block ownership rotates with k
#pragma unroll
            for (int q = 0; q < MAXV; q++)
                if (q < nv) vdj[q] = vdj[q] == 1 ? bw - vo[q] : vdj[q] - 1;
#endif
        for (int k = BAND_KBEG; k < BAND_KEND; k++) {
            const int par = k & 1;
            const int nd = bw < nb - 1 - k ? bw : nb - 1 - k;
            const double *pan = pan0 + par * pl.pan_doubles;
            mbar_wait_cta(&bar_panel[par], ph(k)); // also keeps a warp without tasks from running ahead of the barrier phases
            // two passes: first-column tasks (dj = 1) first, the next step waits for them
            // with BAND_STREAM_UPDATES the block rows 1..3 are not updated here: the streaming warps (roles 7 and 1) do it
#pragma unroll
            // with BAND_STREAM1 the first block column (dj = 1) is not updated here at all (role 7 and the row-solve warps do it); the
            // second one (dj = 2) goes first and is signalled: those blocks are the targets of the next step's streamed updates
            for (int pass = 0; pass < (BAND_STREAM_UPDATES ? 3 : 2); pass++)
#pragma unroll
                for (int q = 0; q < MAXV; q++) { // my virtual workers are sorted by offset, i.e. the tasks of a column come in di order
                    if (q >= nv) continue;
                    const int dj = vdj[q], di = dj + vo[q];
                    if ((BAND_STREAM_UPDATES ? (dj < 3 ? dj - 1 : 2) : (BAND_STREAM1 ? (dj == 2 ? 0 : 1) : (dj == 1 ? 0 : 1))) != pass || di > nd ||
                        (BAND_STREAM_UPDATES && di <= 3) || (BAND_STREAM1 && dj == 1))
                        continue;
                    if (di >= 2) mbar_wait_cta(&bar_p[par][di], ph(k));
                    if (dj >= 2) mbar_wait_cta(&bar_p[par][dj], ph(k));
                    BAND_TICK(1);
                    band_update_dmma(Wk(di, dj), pan + 16 * di, pan + 16 * dj, pcs, lane);
                    if (dj <= ((BAND_STREAM_UPDATES || BAND_STREAM1) ? 2 : 1)) { // the first (two) block column(s) are waited for by the next step
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cta(dj == 1 ? &bar_c1[par][di] : &bar_c2[par][di]);
                    }
                    BAND_TICK(2);
                }
#pragma unroll
            for (int q = 0; q < MAXV; q++)
                if (q < nv) vdj[q] = vdj[q] == 1 ? bw - vo[q] : vdj[q] - 1;
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&bar_step[par]);
            next_k();
        }
    } else if (role == 7) {
        // ------------------------------------------------------------------ D_(k+1) -= P_(k+1,k) P_(k+1,k)^T, streamed behind the chain:
        // lane (r, hh) holds entries 8 hh .. 8 hh + 7 of row r and applies one rank-1 update per published pivot column
        const int r = lane & 15, hh = lane >> 4;
        for (int k = BAND_KBEG; k < BAND_KEND; k++) {
            if (k + 1 >= nb) { // no next diagonal block
                next_k();
                continue;
            }
            const int par = k & 1;
            const double *pan = pan0 + par * pl.pan_doubles;
            if (!BAND_STREAM1 && k >= 1) { // the block is final through step k-1: task (2,2) of that step, or a fresh row when bw == 1
                if (bw >= 2) mbar_wait_cta(&bar_c2[par ^ 1][2], ph(k - 1));
                else if (k >= 2) mbar_wait_cta(&bar_copy[par], ph(k - 2));
            }
            BAND_TICK(1);
            double *dsrc = Wk(1, 1) + r * WSTR + 8 * hh;
            double dd[8];
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                const double2 v = BAND_STREAM1 ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2 *>(dsrc + c);
                dd[c] = v.x;
                dd[c + 1] = v.y;
            }
#pragma unroll
            for (int c = 0; c < 16; c++) {
                if ((c & 3) == 0) mbar_wait_cta(&colbar[par][c >> 2], ph(k));
                const double *p1 = pan + c * pcs + 16;
                const double x = p1[r];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const double2 q = *reinterpret_cast<const double2 *>(p1 + 8 * hh + j);
                    dd[j] = fma(-x, q.x, dd[j]);
                    dd[j + 1] = fma(-x, q.y, dd[j + 1]);
                }
            }
            if (BAND_STREAM1) {
                // accumulated from zero: add to the block once it is final through step k-1 (task (2,2) of that step: long done)
                if (k >= 1) {
                    if (bw >= 2) mbar_wait_cta(&bar_c2[par ^ 1][2], ph(k - 1));
                    else if (k >= 2) mbar_wait_cta(&bar_copy[par], ph(k - 2));
                }
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    double2 v = *reinterpret_cast<double2 *>(dsrc + c);
                    v.x += dd[c];
                    v.y += dd[c + 1];
                    *reinterpret_cast<double2 *>(dsrc + c) = v;
                }
            } else {
#pragma unroll
                for (int c = 0; c < 8; c += 2) *reinterpret_cast<double2 *>(dsrc + c) = make_double2(dd[c], dd[c + 1]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&bar_c1[par][1]); // task (1,1) of this step is done
            BAND_TICK(2);
            next_k();
        }
    } else if (role == 5) {
        // ------------------------------------------------------------------ copy warps
        for (int k = BAND_KBEG; k < BAND_KEND; k++) {
            const int par = k & 1;
            mbar_wait_cta(&bar_panel[par], ph(k));
            BAND_TICK(1);
            copy_share(k, ridx * 32 + lane, N_COPY * 32);
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&bar_copy[par]);
            BAND_TICK(2);
            next_k();
        }
    }
#if SDV_BAND_BABE
    } // phase
    if (rev) { // CTA 1: interior done; publish, meet CTA 0 at barrier #1 and learn whether either half failed
        if (warp == 0 && !ok) s_fail = 1;
        __syncthreads();
        if (s_fail) acc->chol_fail = 1;
        __threadfence();
        cluster_sync_all();
        if (__ldcg(&acc->chol_fail) != 0 || __ldcg(&acc->schur_fail) != 0) return;
    }
#endif
    BAND_TICK(3);
    if (warp == 0 && !ok) s_fail = 1;
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory"); // generic-proxy writes (Lb, window) before the bulk copies below
    __syncthreads();
    if (s_fail) acc->chol_fail = 1; // benign race: every thread writes the same value
    const bool bad = s_fail != 0 || __ldcg(&acc->schur_fail) != 0;
    if (bad) {
        if (threadIdx.x == 0) {
            st->step_valid = 0;
            st->model_cost_change = 0.0;
        }
#if SDV_BAND_BABE
        if (babe) { // only CTA 0 gets here (a separator pivot failed; interior failures returned after barrier #1): release CTA 1
            __threadfence();
            if (threadIdx.x == 0) mbar_remote_arrive(&bar_xsep, 1);
            cluster_sync_all(); // #3
        }
#endif
        return;
    }
    BAND_TICK(4);

    // ---------------------------------------------------------------------- backward solve L^T x = y (warp 0)
    if (warp == 0) {
        const int stage_doubles = (bw + 2) * 256, NS = pl.stages;
        const uint32_t stage_bytes = (uint32_t)stage_doubles * 8u;
        double *ring = bsm;
#if SDV_BAND_BABE
        int nbk = nb; // block columns this CTA back-substitutes: CTA 1 starts below the separator, once its unknowns are there
        if (rev) {
#ifdef SDV_BAND_PROF
            const long long t_x0 = rdclk();
#endif
            mbar_wait_cluster(&bar_xsep, 0); // also: CTA 0 has finished reading this CTA's window (the ring overlays it)
#ifdef SDV_BAND_PROF
            t_xsep = rdclk() - t_x0;
#endif
            nbk = __ldcg(&acc->chol_fail) != 0 ? 0 : nint;
        }
#endif
        if (lane == 0)
            for (int s = 0; s < NS && s < BAND_BNB; s++) {
                const int k = BAND_BNB - 1 - s;
                mbar_expect_tx(&full[s], stage_bytes);
                bulk_g2s(ring + s * stage_doubles, Lb + (size_t)k * stage_doubles, stage_bytes, &full[s]);
            }
        const int c = lane & 15, hh = lane >> 4;
        double *rvs = invs; // 16 doubles of scratch (the reciprocal diagonals are no longer needed)
        int s = 0, ph = 0;
#if SDV_BAND_BACKWARD_V3
        bool have_qp = false;
        double qp = 0.0;
#endif
        for (int it = 0; it < BAND_BNB; it++) {
            const int k = BAND_BNB - 1 - it;
            const int nd = bw < nb - 1 - k ? bw : nb - 1 - k;
            BAND_TICK(5);
#ifdef SDV_BAND_PROF
            tbc = rdclk();
#endif
            mbar_wait(&full[s], (unsigned)ph);
            BAND_TICK(3);
            BAND_TICKB(0);
            const double *sb = ring + s * stage_doubles;
            double x;
#if SDV_BAND_BACKWARD_V3
            if (nd == bw && (bw == 3 || bw == 4)) { // full step of the common band widths, software-pipelined with the next one
                const bool next = it + 1 < BAND_BNB; // block column k-1 exists (and its step is full, too)
                const int sn = s + 1 == NS ? 0 : s + 1;
                const double *sbn = ring + sn * stage_doubles;
                if (!have_qp) qp = bw == 3 ? band_backward_partial<3>(sb, gs, k, lane) : band_backward_partial<4>(sb, gs, k, lane);
                if (next) {
                    mbar_wait(&full[sn], (unsigned)(sn == 0 ? ph ^ 1 : ph));
                    x = bw == 3 ? band_backward_pipe_step<3, true>(sb, sbn, gs, rvs, k, lane, qp) : band_backward_pipe_step<4, true>(sb, sbn, gs, rvs, k, lane, qp);
                } else
                    x = bw == 3 ? band_backward_pipe_step<3, false>(sb, sbn, gs, rvs, k, lane, qp) : band_backward_pipe_step<4, false>(sb, sbn, gs, rvs, k, lane, qp);
                have_qp = next;
            } else
#elif SDV_BAND_BACKWARD_V2
            if (nd == bw && (bw == 3 || bw == 4)) { // full step of the common band widths: specialised body
                x = bw == 3 ? band_backward_full_step<3>(sb, gs, rvs, k, lane) : band_backward_full_step<4>(sb, gs, rvs, k, lane);
            } else
#endif
            {
            // (L_(k+d,k))^T x_(k+d), d = 1..nd: lane (c, hh) sums rows 8 hh .. 8 hh + 7 of every block; four independent chains
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int d = 1; d <= BAND_MAX_BW; d++) {
                if (d > nd) break;
                const double *Lc = sb + d * 256 + hh * 128 + c;
                const double2 *x2 = reinterpret_cast<const double2 *>(gs + (k + d) * BN + hh * 8);
                const double2 xa = x2[0], xb = x2[1], xc = x2[2], xd = x2[3];
                s0 = fma(Lc[0], xa.x, s0);
                s1 = fma(Lc[16], xa.y, s1);
                s2 = fma(Lc[32], xb.x, s2);
                s3 = fma(Lc[48], xb.y, s3);
                s0 = fma(Lc[64], xc.x, s0);
                s1 = fma(Lc[80], xc.y, s1);
                s2 = fma(Lc[96], xd.x, s2);
                s3 = fma(Lc[112], xd.y, s3);
            }
            double sum = (s0 + s1) + (s2 + s3);
            sum += __shfl_xor_sync(FULL, sum, 16);
            const double rv = gs[k * BN + c] - sum; // both half-warps hold rv[c]
            if (hh == 0) rvs[c] = rv;               // x_k = L_kk^-T rv = Minv^T rv, broadcast through shared memory
            __syncwarp();
            const double *Mi = sb + (bw + 1) * 256 + hh * 128 + c;
            const double2 *r2 = reinterpret_cast<const double2 *>(rvs + hh * 8);
            const double2 ra = r2[0], rb = r2[1], rc = r2[2], rd = r2[3];
            double x0 = Mi[0] * ra.x, x1 = Mi[16] * ra.y, x2v = Mi[32] * rb.x, x3 = Mi[48] * rb.y;
            x0 = fma(Mi[64], rc.x, x0);
            x1 = fma(Mi[80], rc.y, x1);
            x2v = fma(Mi[96], rd.x, x2v);
            x3 = fma(Mi[112], rd.y, x3);
            x = (x0 + x1) + (x2v + x3);
            x += __shfl_xor_sync(FULL, x, 16);
            }
            BAND_TICKB(1);
            if (hh == 0) {
                gs[k * BN + c] = x;
#if SDV_BAND_BABE
                dxp[rev ? P.n_pad - 1 - (k * BN + c) : k * BN + c] = -x;
                // x also goes to the other CTA: the separator unknowns to CTA 1 (it needs them), CTA 1's interior to CTA 0 (epilogue)
                if (babe && (rev || k >= nint)) dsmem_store(gs + (P.n_pad - 1 - (k * BN + c)), rev ? 0u : 1u, x);
#elif SDV_BAND_REV
                dxp[P.n_pad - 1 - (k * BN + c)] = -x;
#else
                dxp[k * BN + c] = -x;
#endif
            }
#if SDV_BAND_BABE
            if (babe && !rev && k == nint) { // last separator block: CTA 1 may start
                asm volatile("fence.acq_rel.cluster;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_remote_arrive(&bar_xsep, 1);
            }
#endif
            __syncwarp();
            BAND_TICKB(2);
            if (lane == 0 && it + NS < BAND_BNB) {
                const int k2 = BAND_BNB - 1 - (it + NS);
                mbar_expect_tx(&full[s], stage_bytes);
                bulk_g2s(ring + s * stage_doubles, Lb + (size_t)k2 * stage_doubles, stage_bytes, &full[s]);
            }
            __syncwarp();
            BAND_TICKB(3);
            if (++s == NS) {
                s = 0;
                ph ^= 1;
            }
        }
    }
    __syncthreads();
#if SDV_BAND_BABE
#ifdef SDV_BAND_PROF
    if (rev && prof && threadIdx.x == 0) { // CTA 1's backward solve: slots 6, 7 of warps 6, 7, and its wait for the separator unknowns
        for (int q = 0; q < 4; q++) prof[(6 + (q >> 1)) * 8 + 6 + (q & 1)] = (double)tbk[q];
        prof[8 * 8 + 6] = (double)t_xsep;
    }
    const long long t_sync3 = rdclk();
#endif
    if (babe) { // #3: both halves of x are in CTA 0's shared memory and in dxp
        __threadfence();
        cluster_sync_all();
        if (rev) return;
    }
#ifdef SDV_BAND_PROF
    if (prof && threadIdx.x == 0) prof[8 * 8 + 7] = (double)(rdclk() - t_sync3); // CTA 0 waiting for CTA 1 at the end
#endif
#endif
    BAND_TICK(5);
#ifdef SDV_BAND_PROF
    if (prof && lane == 0)
        for (int q = 0; q < 6; q++) prof[warp * 8 + q] = (double)tp[q]; // warp 0: [setup, wait, chain, -, tail, backward]
    if (prof && lane == 0 && warp == 0) {
        for (int q = 0; q < 5; q++) prof[(q >> 1) * 8 + 6 + (q & 1)] = (double)tw[q]; // free slots 6, 7 of warps 0, 1, 2
        for (int q = 0; q < 4; q++) prof[(4 + (q >> 1)) * 8 + 6 + (q & 1)] = (double)tbk[q]; // slots 6, 7 of warps 4, 5
    }
#endif

    // ---------------------------------------------------------------------- reduced-parameter update, model-decrease terms,
    // candidate frame-camera table (same epilogue as chol_backward_v2)
    const LinBuf &Bx = st->cur ? B1 : B0;
    const LinBuf &Bc = st->cur ? B0 : B1;
    const int n = P.n;
    double gd = 0, dd = 0, sn = 0, cn = 0;
    for (int i = threadIdx.x; i < P.n_pad; i += BCT) {
#if SDV_BAND_REV
        const double d = i < n ? -gs[P.n_pad - 1 - i] : 0.0;
#else
        const double d = i < n ? -gs[i] : 0.0;
#endif
        if (i >= n) dxp[i] = 0.0;
        const double xc = Bx.xp[i] + d;
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
    for (int i = threadIdx.x; i < P.F * P.C; i += BCT) compute_fct_row(P, Bc.xp, i / P.C, i % P.C, Bc.fct + (size_t)i * FCT_ROW);
    if (threadIdx.x == 0) st->step_valid = 1;
}

} // namespace sdv
