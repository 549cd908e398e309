// Backward solve L^T x = y of k_chol_band on one warp, alone on the SM: where do the ~800 cycles per block step go?
// (dependent latency of one step is ~300 cycles.)  All variants read the same band storage — per block
// column k: blocks L_(k,k)^T .. L_(k+bw,k)^T and the explicit inverse of L_kk, (bw + 2) * 256 doubles, exactly the layout the
// kernel streams — and are checked against a host solve.
//   V0  the shipped loop: TMA bulk ring (16 stages), `if (d > nd) break` inside the unrolled block loop
//   V1  same data path, full steps (nd = bw) through a fully unrolled body with the d = 1 block LAST (its x is the only
//       operand produced by the previous step), partial steps through the generic loop
//   V2  V1 with the whole factor resident in shared memory (no TMA, no mbarrier): the pure dependency chain
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o backward backward.cu ; run: ./backward [nb] [bw]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../sadvio_b200/csrc/sdv_chol_band.cuh"
using namespace sdv;
__device__ __forceinline__ long long rdclk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }

constexpr int NS = 16; // ring depth of the shipped kernel at bw = 3

// one block step, generic (what the kernel ships): returns x_k[c] in both half-warps
__device__ __forceinline__ double step_generic(const double *sb, const double *gs, double *rvs, int k, int nd, int bw, int lane) {
    const int c = lane & 15, hh = lane >> 4;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int d = 1; d <= BAND_MAX_BW; d++) {
        if (d > nd) break;
        const double *Lc = sb + d * 256 + hh * 128 + c;
        const double2 *x2 = reinterpret_cast<const double2 *>(gs + (k + d) * BN + hh * 8);
        const double2 xa = x2[0], xb = x2[1], xc = x2[2], xd = x2[3];
        s0 = fma(Lc[0], xa.x, s0); s1 = fma(Lc[16], xa.y, s1); s2 = fma(Lc[32], xb.x, s2); s3 = fma(Lc[48], xb.y, s3);
        s0 = fma(Lc[64], xc.x, s0); s1 = fma(Lc[80], xc.y, s1); s2 = fma(Lc[96], xd.x, s2); s3 = fma(Lc[112], xd.y, s3);
    }
    double sum = (s0 + s1) + (s2 + s3);
    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
    const double rv = gs[k * BN + c] - sum;
    if (hh == 0) rvs[c] = rv;
    __syncwarp();
    const double *Mi = sb + (bw + 1) * 256 + hh * 128 + c;
    const double2 *r2 = reinterpret_cast<const double2 *>(rvs + hh * 8);
    const double2 ra = r2[0], rb = r2[1], rc = r2[2], rd = r2[3];
    double x0 = Mi[0] * ra.x, x1 = Mi[16] * ra.y, x2v = Mi[32] * rb.x, x3 = Mi[48] * rb.y;
    x0 = fma(Mi[64], rc.x, x0); x1 = fma(Mi[80], rc.y, x1); x2v = fma(Mi[96], rd.x, x2v); x3 = fma(Mi[112], rd.y, x3);
    double x = (x0 + x1) + (x2v + x3);
    x += __shfl_xor_sync(0xffffffffu, x, 16);
    return x;
}

// full step (nd = BW) with everything that does not depend on x_(k+1) issued first
template <int BW> __device__ __forceinline__ double step_full(const double *sb, const double *gs, double *rvs, int k, int lane) {
    const int c = lane & 15, hh = lane >> 4;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const double *Mi = sb + (BW + 1) * 256 + hh * 128 + c;
    double m[8];
#pragma unroll
    for (int q = 0; q < 8; q++) m[q] = Mi[16 * q]; // the inverse block does not depend on anything of this step
#pragma unroll
    for (int d = BW; d >= 2; d--) {
        const double *Lc = sb + d * 256 + hh * 128 + c;
        const double2 *x2 = reinterpret_cast<const double2 *>(gs + (k + d) * BN + hh * 8);
        const double2 xa = x2[0], xb = x2[1], xc = x2[2], xd = x2[3];
        s0 = fma(Lc[0], xa.x, s0); s1 = fma(Lc[16], xa.y, s1); s2 = fma(Lc[32], xb.x, s2); s3 = fma(Lc[48], xb.y, s3);
        s0 = fma(Lc[64], xc.x, s0); s1 = fma(Lc[80], xc.y, s1); s2 = fma(Lc[96], xd.x, s2); s3 = fma(Lc[112], xd.y, s3);
    }
    {
        const double *Lc = sb + 256 + hh * 128 + c;
        double l[8];
#pragma unroll
        for (int q = 0; q < 8; q++) l[q] = Lc[16 * q];
        const double2 *x2 = reinterpret_cast<const double2 *>(gs + (k + 1) * BN + hh * 8); // the critical operand
        const double2 xa = x2[0], xb = x2[1], xc = x2[2], xd = x2[3];
        s0 = fma(l[0], xa.x, s0); s1 = fma(l[1], xa.y, s1); s2 = fma(l[2], xb.x, s2); s3 = fma(l[3], xb.y, s3);
        s0 = fma(l[4], xc.x, s0); s1 = fma(l[5], xc.y, s1); s2 = fma(l[6], xd.x, s2); s3 = fma(l[7], xd.y, s3);
    }
    double sum = (s0 + s1) + (s2 + s3);
    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
    const double rv = gs[k * BN + c] - sum;
    if (hh == 0) rvs[c] = rv;
    __syncwarp();
    const double2 *r2 = reinterpret_cast<const double2 *>(rvs + hh * 8);
    const double2 ra = r2[0], rb = r2[1], rc = r2[2], rd = r2[3];
    double x0 = m[0] * ra.x, x1 = m[1] * ra.y, x2v = m[2] * rb.x, x3 = m[3] * rb.y;
    x0 = fma(m[4], rc.x, x0); x1 = fma(m[5], rc.y, x1); x2v = fma(m[6], rd.x, x2v); x3 = fma(m[7], rd.y, x3);
    double x = (x0 + x1) + (x2v + x3);
    x += __shfl_xor_sync(0xffffffffu, x, 16);
    return x;
}

template <int BW> __device__ __forceinline__ double step_any(const double *sb, const double *gs, double *rvs, int k, int nd, int lane, bool full) {
    return full ? step_full<BW>(sb, gs, rvs, k, lane) : step_generic(sb, gs, rvs, k, nd, BW, lane);
}

// MODE 0: shipped loop, TMA ring; 1: specialised full steps, TMA ring; 2: specialised full steps, factor resident in shared memory
template <int MODE, int BW> __global__ void __launch_bounds__(512, 1) k_backward(const double *Lb, const double *y, double *xout, long long *res, int nb) {
    extern __shared__ __align__(128) double sm[];
    __shared__ __align__(8) uint64_t full[NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stage_doubles = (BW + 2) * 256;
    const uint32_t stage_bytes = (uint32_t)stage_doubles * 8u;
    double *gs = sm, *rvs = gs + nb * BN, *ring = rvs + 32; // ring is 128-byte aligned: nb * 16 + 32 doubles
    for (int e = threadIdx.x; e < nb * BN; e += blockDim.x) gs[e] = y[e];
    constexpr bool RES = MODE == 2 || MODE == 4 || MODE == 5; // factor resident in shared memory
    constexpr bool PIPE = MODE == 3 || MODE == 4;             // the kernel's software-pipelined full steps (band_backward_pipe_step)
    if (RES)
        for (int e = threadIdx.x; e < nb * stage_doubles; e += blockDim.x) ring[e] = Lb[e];
    if (threadIdx.x == 0)
        for (int s = 0; s < NS; s++) mbar_init(&full[s], 1);
    __syncthreads();
    if (warp != 0) return; // the other warps of the shipped kernel wait at a __syncthreads() during the backward solve
    const long long t0 = rdclk();
    if (!RES && lane == 0)
        for (int s = 0; s < NS && s < nb; s++) {
            const int k = nb - 1 - s;
            mbar_expect_tx(&full[s], stage_bytes);
            bulk_g2s(ring + s * stage_doubles, Lb + (size_t)k * stage_doubles, stage_bytes, &full[s]);
        }
    const int c = lane & 15, hh = lane >> 4;
    int s = 0, ph = 0;
    long long t_first = 0;
    bool have_qp = false;
    double qp = 0.0;
    for (int it = 0; it < nb; it++) {
        const int k = nb - 1 - it;
        const int nd = BW < nb - 1 - k ? BW : nb - 1 - k;
        const double *sb, *sbn = nullptr;
        if (RES) {
            sb = ring + (size_t)k * stage_doubles;
            sbn = sb - stage_doubles;
        } else {
            mbar_wait(&full[s], (unsigned)ph);
            sb = ring + s * stage_doubles;
        }
        if (it == 0) t_first = rdclk();
        double x;
        if (MODE == 5) { // only what depends on x_(k+1): the floor of the dependence chain (result not checked)
            double q0 = 0.0;
            x = nd == BW ? band_backward_pipe_step<BW, false>(sb, sbn, gs, rvs, k, lane, q0) : step_generic(sb, gs, rvs, k, nd, BW, lane);
        } else if (PIPE && nd == BW) {
            const bool next = it + 1 < nb;
            const int sn = s + 1 == NS ? 0 : s + 1;
            if (!RES) sbn = ring + sn * stage_doubles;
            if (!have_qp) qp = band_backward_partial<BW>(sb, gs, k, lane);
            if (next) {
                if (!RES) mbar_wait(&full[sn], (unsigned)(sn == 0 ? ph ^ 1 : ph));
                x = band_backward_pipe_step<BW, true>(sb, sbn, gs, rvs, k, lane, qp);
            } else
                x = band_backward_pipe_step<BW, false>(sb, sbn, gs, rvs, k, lane, qp);
            have_qp = next;
        } else
            x = MODE == 0 ? step_generic(sb, gs, rvs, k, nd, BW, lane) : step_any<BW>(sb, gs, rvs, k, nd, lane, nd == BW);
        if (hh == 0) {
            gs[k * BN + c] = x;
            xout[k * BN + c] = -x;
        }
        __syncwarp();
        if (!RES) {
            if (lane == 0 && it + NS < nb) {
                const int k2 = nb - 1 - (it + NS);
                mbar_expect_tx(&full[s], stage_bytes);
                bulk_g2s(ring + s * stage_doubles, Lb + (size_t)k2 * stage_doubles, stage_bytes, &full[s]);
            }
            if (++s == NS) {
                s = 0;
                ph ^= 1;
            }
        }
    }
    const long long t1 = rdclk();
    if (lane == 0) {
        res[0] = t1 - t0;
        res[1] = t1 - t_first;
    }
}

template <int MODE, int BW> static void run(const char *name, const double *dLb, const double *dy, double *dx, long long *dres, int nb, const std::vector<double> &xref) {
    const int stage_doubles = (BW + 2) * 256;
    const size_t smem = (size_t)(nb * BN + 32 + ((MODE == 2 || MODE == 4 || MODE == 5) ? nb : NS) * stage_doubles) * 8;
    if (smem > 227 * 1024) {
        printf("%-44s skipped (needs %zu KB of shared memory)\n", name, smem / 1024);
        return;
    }
    cudaFuncSetAttribute(k_backward<MODE, BW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long h[2] = {0, 0};
    for (int rep = 0; rep < 3; rep++) k_backward<MODE, BW><<<1, 512, smem>>>(dLb, dy, dx, dres, nb);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, dres, sizeof(h), cudaMemcpyDeviceToHost);
    std::vector<double> x(nb * BN);
    cudaMemcpy(x.data(), dx, x.size() * 8, cudaMemcpyDeviceToHost);
    double err = 0, mx = 0;
    for (size_t i = 0; i < x.size(); i++) {
        err = fmax(err, fabs(-x[i] - xref[i]));
        mx = fmax(mx, fabs(xref[i]));
    }
    printf("%-44s %7lld cycles total, %6.0f per block step after the first stage landed, max rel err %.1e  (%s)\n", name, h[0], (double)h[1] / nb, err / mx,
           cudaGetErrorString(e));
}

int main(int argc, char **argv) {
    const int nb = argc > 1 ? atoi(argv[1]) : 20, bw = argc > 2 ? atoi(argv[2]) : 3;
    if (bw != 3 && bw != 4) {
        printf("bw must be 3 or 4\n");
        return 1;
    }
    const int n = nb * BN, stage_doubles = (bw + 2) * 256;
    // a well-conditioned banded lower-triangular L (row-major dense on the host), its band storage and the reference solve
    std::vector<double> L((size_t)n * n, 0.0), y(n), xref(n), Lb((size_t)nb * stage_doubles, 0.0);
    srand(5);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j <= i; j++)
            if (i / BN - j / BN <= bw) L[(size_t)i * n + j] = i == j ? 2.0 + (rand() % 100) * 0.01 : ((rand() % 200) - 100) * 5e-4;
        y[i] = ((rand() % 200) - 100) * 0.01;
    }
    for (int i = n - 1; i >= 0; i--) { // L^T x = y
        double s = y[i];
        for (int j = i + 1; j < n; j++) s -= L[(size_t)j * n + i] * xref[j];
        xref[i] = s / L[(size_t)i * n + i];
    }
    for (int k = 0; k < nb; k++) {
        double *st = Lb.data() + (size_t)k * stage_doubles;
        // blocks d = 0 .. bw: element (r, c) of the stage = L_(k+d,k)[r][c] stored as [d][r][c] (row r of the stacked panel)
        for (int d = 0; d <= bw && k + d < nb; d++)
            for (int r = 0; r < BN; r++)
                for (int c = 0; c < BN; c++) st[d * 256 + r * 16 + c] = L[(size_t)((k + d) * BN + r) * n + k * BN + c];
        // explicit inverse of L_kk (lower triangular), stored like a block: Minv[r][c]; x_k = Minv^T rv
        double Mi[BN][BN] = {};
        for (int c = 0; c < BN; c++) {
            for (int r = c; r < BN; r++) {
                double s = r == c ? 1.0 : 0.0;
                for (int q = c; q < r; q++) s -= L[(size_t)(k * BN + r) * n + k * BN + q] * Mi[q][c];
                Mi[r][c] = s / L[(size_t)(k * BN + r) * n + k * BN + r];
            }
        }
        for (int r = 0; r < BN; r++)
            for (int c = 0; c < BN; c++) st[(bw + 1) * 256 + r * 16 + c] = Mi[r][c];
    }
    double *dLb, *dy, *dx;
    long long *dres;
    cudaMalloc(&dLb, Lb.size() * 8);
    cudaMalloc(&dy, n * 8);
    cudaMalloc(&dx, n * 8);
    cudaMalloc(&dres, 64);
    cudaMemcpy(dLb, Lb.data(), Lb.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dy, y.data(), n * 8, cudaMemcpyHostToDevice);
    printf("backward solve, nb = %d block columns, bw = %d\n", nb, bw);
    if (bw == 3) {
        run<0, 3>("V0 shipped loop, TMA ring", dLb, dy, dx, dres, nb, xref);
        run<1, 3>("V1 full steps specialised (d = 1 last), TMA ring", dLb, dy, dx, dres, nb, xref);
        run<2, 3>("V2 specialised, factor resident in smem", dLb, dy, dx, dres, nb, xref);
        run<3, 3>("V3 software-pipelined full steps, TMA ring", dLb, dy, dx, dres, nb, xref);
        run<4, 3>("V4 software-pipelined, factor resident", dLb, dy, dx, dres, nb, xref);
        run<5, 3>("V5 d = 1 block + inverse only, resident (floor)", dLb, dy, dx, dres, nb, xref);
    } else {
        run<0, 4>("V0 shipped loop, TMA ring", dLb, dy, dx, dres, nb, xref);
        run<1, 4>("V1 full steps specialised (d = 1 last), TMA ring", dLb, dy, dx, dres, nb, xref);
        run<2, 4>("V2 specialised, factor resident in smem", dLb, dy, dx, dres, nb, xref);
        run<3, 4>("V3 software-pipelined full steps, TMA ring", dLb, dy, dx, dres, nb, xref);
        run<4, 4>("V4 software-pipelined, factor resident", dLb, dy, dx, dres, nb, xref);
        run<5, 4>("V5 d = 1 block + inverse only, resident (floor)", dLb, dy, dx, dres, nb, xref);
    }
    return 0;
}
