// Decomposition of the pivot chain latency (one warp).
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
__device__ __forceinline__ double band_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = d * y, e = fma(-t, y, 1.0), q = fma(0.375, e, 0.5), ye = y * e;
    return fma(ye, q, y);
}
template <int V>
__global__ void k(double *out, long long *cyc) {
    __shared__ double sm[64];
    const int lane = threadIdx.x & 31;
    double a[32];
#pragma unroll
    for (int j = 0; j < 32; j++) a[j] = 40.0 + j + lane * 1e-3;
    double d = 30.0 + lane, inv = 0.2;
    long long t0, t1;
    *(volatile double *)&sm[lane] = d;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
#pragma unroll 1
    for (int c = 0; c < 128; c++) {
        if (V == 0) { // rsqrt only
            d = band_rsqrt(d) + 30.0;
        } else if (V == 1) { // rsqrt + shfl
            d = __shfl_sync(FULL, band_rsqrt(d) + 30.0, (c + 1) & 31);
        } else if (V == 2) { // mul, fma, shfl, rsqrt  (the intended chain, no selects, no updates)
            double l = a[0] * inv;
            double pc = fma(-l, l, a[1]);
            double dn = __shfl_sync(FULL, pc, (c + 1) & 31);
            inv = band_rsqrt(dn);
            a[0] = a[0] + 1e-9 * inv; a[1] += 1e-3;
        } else if (V == 3) { // + selects
            double l = a[0] * inv;
            l = (lane & 15) == (c & 15) ? d * inv : ((lane & 15) < (c & 15) ? 0.0 : l);
            double pc = fma(-l, l, a[1]);
            double dn = __shfl_sync(FULL, pc, (c + 1) & 31);
            inv = band_rsqrt(dn);
            d = dn;
            a[0] = a[0] + 1e-9 * inv; a[1] += 1e-3;
        } else if (V == 4) { // + 31 independent rotating updates
            double l = a[0] * inv;
            l = (lane & 15) == (c & 15) ? d * inv : ((lane & 15) < (c & 15) ? 0.0 : l);
            double pc = fma(-l, l, a[1]);
            double dn = __shfl_sync(FULL, pc, (c + 1) & 31);
            double invn = band_rsqrt(dn);
#pragma unroll
            for (int j = 1; j < 32; j++) a[j - 1] = fma(-l, 1e-3, a[j]);
            a[31] = 60.0; a[0] += 35.0;
            d = dn; inv = invn;
        } else if (V == 5) { // + STS / syncwarp
            double l = a[0] * inv;
            l = (lane & 15) == (c & 15) ? d * inv : ((lane & 15) < (c & 15) ? 0.0 : l);
            double pc = fma(-l, l, a[1]);
            double dn = __shfl_sync(FULL, pc, (c + 1) & 31);
            double invn = band_rsqrt(dn);
            sm[lane] = l;
            __syncwarp();
#pragma unroll
            for (int j = 1; j < 32; j++) a[j - 1] = fma(-l, 1e-3, a[j]);
            a[31] = 60.0; a[0] += 35.0;
            d = dn; inv = invn;
        } else if (V == 6) { // V4 with only 15 updates
            double l = a[0] * inv;
            l = (lane & 15) == (c & 15) ? d * inv : ((lane & 15) < (c & 15) ? 0.0 : l);
            double pc = fma(-l, l, a[1]);
            double dn = __shfl_sync(FULL, pc, (c + 1) & 31);
            double invn = band_rsqrt(dn);
#pragma unroll
            for (int j = 1; j < 16; j++) a[j - 1] = fma(-l, 1e-3, a[j]);
            a[15] = 60.0; a[0] += 35.0;
            d = dn; inv = invn;
        }
    }
    {
        double q = d + inv;
#pragma unroll
        for (int j = 0; j < 32; j++) q += a[j];
        *(volatile double *)&sm[lane + 32] = q;
    }
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
    if (threadIdx.x == 0) cyc[V] = t1 - t0;
    double s = d + inv;
#pragma unroll
    for (int j = 0; j < 32; j++) s += a[j];
    out[threadIdx.x] = s + sm[(lane + 1) & 31];
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8);
    for (int rep = 0; rep < 2; rep++) {
        k<0><<<1, 32>>>(out, cyc); k<1><<<1, 32>>>(out, cyc); k<2><<<1, 32>>>(out, cyc); k<3><<<1, 32>>>(out, cyc);
        k<4><<<1, 32>>>(out, cyc); k<5><<<1, 32>>>(out, cyc); k<6><<<1, 32>>>(out, cyc);
    }
    long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const char *nm[] = {"rsqrt+add", "rsqrt+add+shfl", "mul,fma,shfl,rsqrt", "+selects", "+31 updates", "+STS,syncwarp", "15 updates"};
    for (int v = 0; v < 7; v++) printf("%-20s: %.1f cycles / iteration\n", nm[v], h[v] / 128.0);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
