// Latency micro-benchmarks on one warp (B200): dependent DFMA / DMUL chains, MUFU.RSQ64H, SHFL, STS->LDS round trip,
// DMMA m8n8k4, and DFMA issue throughput with 1..4 independent chains.  nvcc -arch=sm_100a -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double seed) {
    __shared__ double sm[64];
    const int lane = threadIdx.x & 31;
    double x = seed + lane * 1e-3, y = 1.0000001, z = 1e-9;
    long long t0, t1;
    // 1. dependent DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) x = fma(x, y, z);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 2. 4 independent DFMA chains (issue throughput)
    double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) { a0 = fma(a0, y, z); a1 = fma(a1, y, z); a2 = fma(a2, y, z); a3 = fma(a3, y, z); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    x = a0 + a1 + a2 + a3;
    // 3. 16 independent chains
    double b[16];
#pragma unroll
    for (int j = 0; j < 16; j++) b[j] = x + j;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++)
#pragma unroll
        for (int j = 0; j < 16; j++) b[j] = fma(b[j], y, z);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
#pragma unroll
    for (int j = 0; j < 16; j++) x += b[j];
    // 4. MUFU.RSQ64H dependent chain (rsqrt.approx.ftz.f64)
    double r = fabs(x) + 2.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 128; i++) { double q; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(r)); r = q + 3.0; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;  // = 128 * (MUFU + DADD)
    x += r;
    // 5. SHFL (64-bit = 2 shuffles) dependent chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 128; i++) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 6. STS -> syncwarp -> LDS round trip
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 128; i++) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31]; __syncwarp(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 7. DMMA dependent chain
    double c0 = x, c1 = x;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 128; i++) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(y), "d"(z));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // 8. DMMA 4 independent accumulators
    double d0 = x, d1 = x, d2 = x, d3 = x, d4 = x, d5 = x, d6 = x, d7 = x;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 128; i++) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(y), "d"(z));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d2), "+d"(d3) : "d"(y), "d"(z));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d4), "+d"(d5) : "d"(y), "d"(z));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d6), "+d"(d7) : "d"(y), "d"(z));
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // 9. dependent DMUL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) x = x * y;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[8] = t1 - t0;
    // 10. full CUDA rsqrt() chain
    r = fabs(x) + 2.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) r = rsqrt(r) + 3.0;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[9] = t1 - t0;
    // 11. FP32 FFMA dependent chain for reference
    float f = (float)x, g = 1.0001f, hh = 1e-6f;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) f = fmaf(f, g, hh);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[10] = t1 - t0;
    out[threadIdx.x] = x + c0 + c1 + d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7 + r + f;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8);
    for (int nw = 1; nw <= 16; nw *= 4) {
        for (int rep = 0; rep < 2; rep++) k<<<1, 32 * nw>>>(out, cyc, 1.5);
        long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("warps=%d: DFMA dep %.1f cyc/op | 4 chains %.1f cyc/4ops | 16 chains %.1f cyc/16ops | MUFU.RSQ64H+DADD %.1f | SHFL64 %.1f | STS+LDS %.1f | DMMA dep %.1f | 4 DMMA %.1f | DMUL dep %.1f | rsqrt()+DADD %.1f | FFMA dep %.1f\n", nw,
               h[0] / 256.0, h[1] / 256.0, h[2] / 64.0, h[3] / 128.0, h[4] / 128.0, h[5] / 128.0, h[6] / 128.0, h[7] / 128.0, h[8] / 256.0, h[9] / 64.0, h[10] / 256.0);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
