// The chain step, the row solve and the tensor-core block update of k_chol_band, each timed alone on one warp
// (optionally with N other warps spinning on an mbarrier, to see what the pollers cost the workers).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../sadvio_b200/csrc/sdv_chol_band.cuh"
using namespace sdv;
__device__ __forceinline__ long long rdclk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
__global__ void __launch_bounds__(512, 1) k(double *out, long long *res, int spinners, int bw) {
    extern __shared__ __align__(16) double sm[];
    __shared__ uint64_t bar, colbar[16];
    const int pcs = 16 * (bw + 1) + 4;
    double *pan = sm, *iv = sm + 16 * pcs + 32, *win = iv + 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < 16 * pcs + 64 + 4 * WBLK; e += blockDim.x) sm[e] = 1e-3 * (e % 7);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        for (int c = 0; c < 16; c++) mbar_init(&colbar[c], 1);
    }
    __syncthreads();
    if (warp > 0) {
        if (warp <= spinners) mbar_wait_cta(&bar, 0);
        return;
    }
    bool ok = true;
    double a[16];
    long long t0, t1, acc = 0;
    for (int rep = 0; rep < 16; rep++) {
#pragma unroll
        for (int c = 0; c < 16; c++) a[c] = (c == (lane & 15) ? 50.0 : 0.01) + lane * 1e-3;
        *(volatile double *)&win[lane] = a[3];
        t0 = rdclk();
        band_chain_step(a, lane, pan, pcs, iv, colbar, ok);
        *(volatile double *)&win[lane] = a[0] + a[7];
        t1 = rdclk();
        acc += t1 - t0;
    }
    if (lane == 0) res[0] = acc / 16;
    acc = 0;
    for (int rep = 0; rep < 16; rep++) {
        double t[16];
#pragma unroll
        for (int c = 0; c < 16; c++) t[c] = 0.5 + c * 0.01 + lane;
        *(volatile double *)&win[lane] = t[3];
        t0 = rdclk();
        double wdum[8];
        if (lane < 16) band_trsm16<false>(t, wdum, 0, pan, pcs, iv, pan + 32 + lane, pcs);
        __syncwarp();
        *(volatile double *)&win[lane] = t[0] + t[7];
        t1 = rdclk();
        acc += t1 - t0;
    }
    if (lane == 0) res[1] = acc / 16;
    acc = 0;
    for (int rep = 0; rep < 16; rep++) {
        *(volatile double *)&win[lane] = 1.0;
        t0 = rdclk();
        band_update_dmma(win + 64, pan + 16, pan + 32, pcs, lane);
        __syncwarp();
        *(volatile double *)&win[lane] = win[64 + lane];
        t1 = rdclk();
        acc += t1 - t0;
    }
    if (lane == 0) res[2] = acc / 16;
    out[threadIdx.x] = a[0] + (ok ? 1 : 0);
    if (lane == 0) mbar_arrive_cta(&bar);
}
int main() {
    double *out; long long *res;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&res, 64);
    const int bw = 4, smem = (16 * (16 * (bw + 1) + 4) + 64 + 4 * WBLK + 64) * 8;
    for (int sp : {0, 3, 15}) {
        for (int rep = 0; rep < 2; rep++) k<<<1, 512, smem>>>(out, res, sp, bw);
        long long h[3]; cudaMemcpy(h, res, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%2d spinning warps: chain step %lld cycles (%.0f / pivot), row solve %lld (%.0f / pivot), dmma block update %lld\n", sp, h[0], h[0] / 16.0, h[1], h[1] / 16.0, h[2]);
    }
    printf("%s %s\n", cudaGetErrorString(cudaGetLastError()), cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
