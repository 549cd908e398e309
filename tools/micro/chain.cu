// Micro-benchmark of the chain-warp pivot loop variants (one warp, shared-memory broadcast of the pivot column).
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
__device__ __forceinline__ void step(double (&a)[32], int lane, int par, double *pan, int pcs, double *iv) {
    const int r = lane & 15;
    const bool piv = (lane >> 4) == par;
    const int pbase = par * 16;
    double *pub = pan + (piv ? r : 16 + r);
    double d = __shfl_sync(FULL, a[0], pbase);
    double inv = band_rsqrt(d);
#pragma unroll 1
    for (int c = 0; c < 16; c++) {
        double l = a[0] * inv;
        if (piv) l = r == c ? d * inv : (r < c ? 0.0 : l);
        const double pc = fma(-l, l, a[1]);
        const double dn = __shfl_sync(FULL, pc, (pbase + c + 1) & 31);
        const double invn = band_rsqrt(dn);
        pub[c * pcs] = l;
        if (lane == pbase + c) iv[c] = inv;
        __syncwarp();
        if (V == 0) { // LDS.64 x 31
            const double *v = pan + c * pcs + c;
#pragma unroll
            for (int j = 1; j < 32; j++) a[j - 1] = fma(-l, v[j], a[j]);
        } else if (V == 1) { // no broadcast loads at all (wrong maths, timing only)
#pragma unroll
            for (int j = 1; j < 32; j++) a[j - 1] = fma(-l, inv, a[j]);
        } else if (V == 2) { // LDS.128 x 16 from an aligned base (wrong maths for odd c, timing only)
            const double2 *v = reinterpret_cast<const double2 *>(pan + c * pcs + (c & ~1));
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const double2 w = v[j];
                if (j > 0) a[2 * j - 1] = fma(-l, w.x, a[2 * j]);
                if (2 * j + 1 < 32) a[2 * j] = fma(-l, w.y, a[2 * j + 1]);
            }
        } else if (V == 3) { // shuffles instead of shared memory: 15 for the pivot block only
#pragma unroll
            for (int j = 1; j < 16; j++) a[j - 1] = fma(-l, __shfl_sync(FULL, l, (pbase + c + j) & 31), a[j]);
        } else if (V == 4) { // LDS.64 x 15 only (pivot block columns), D update dropped
            const double *v = pan + c * pcs + c;
#pragma unroll
            for (int j = 1; j < 16; j++) a[j - 1] = fma(-l, v[j], a[j]);
        }
        d = dn;
        inv = invn;
    }
}
template <int V>
__global__ void k(double *out, long long *cyc, int pcs) {
    extern __shared__ double sm[];
    double *pan = sm, *iv = sm + 16 * pcs + 64;
    const int lane = threadIdx.x & 31;
    for (int e = threadIdx.x; e < 16 * pcs + 128; e += blockDim.x) sm[e] = 1e-3;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    double a[32];
#pragma unroll
    for (int j = 0; j < 32; j++) a[j] = (j == (lane & 15) ? 50.0 : 0.01) + lane * 1e-3;
    long long t0 = clock64();
    for (int k2 = 0; k2 < 8; k2++) {
#pragma unroll
        for (int j = 16; j < 32; j++) a[j] = 40.0 + j;
        a[0] = 30.0;
        step<V>(a, lane, k2 & 1, pan, pcs, iv);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[V] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) s += a[j];
    out[threadIdx.x] = s;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8);
    const int pcs = 84, smem = (16 * pcs + 256) * 8;
    for (int rep = 0; rep < 2; rep++) {
        k<0><<<1, 32, smem>>>(out, cyc, pcs); k<1><<<1, 32, smem>>>(out, cyc, pcs); k<2><<<1, 32, smem>>>(out, cyc, pcs);
        k<3><<<1, 32, smem>>>(out, cyc, pcs); k<4><<<1, 32, smem>>>(out, cyc, pcs);
    }
    long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const char *nm[] = {"LDS.64 x31", "no loads", "LDS.128 x16", "SHFL x15", "LDS.64 x15"};
    for (int v = 0; v < 5; v++) printf("%-12s: %.1f cycles / pivot\n", nm[v], h[v] / 128.0);
    // same with 15 idle warps parked on a barrier? (not needed)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
