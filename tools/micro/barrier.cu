// How long does a warp take to get going again after __syncthreads() when ONE warp of the CTA is busy?  (B200)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long rdclk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
template <int MODE>
__global__ void k(double *out, long long *res, int work) {
    __shared__ double sm[512];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double x = 1.0 + threadIdx.x * 1e-6;
    sm[threadIdx.x] = 0.0;
    __syncthreads();
    long long busy1 = 0, busy2 = 0, tot = 0;
    long long tc = rdclk();
    const long long t00 = tc;
    for (int it = 0; it < 32; it++) {
        // phase 1: warp 0 works, the others have nothing to do
        if (warp == 0) {
            x += ((volatile double *)sm)[lane];
            for (int i = 0; i < work; i++) x = fma(x, 1.0000001, 1e-9);
            ((volatile double *)sm)[lane] = x;
        } else if (MODE == 1) {
            sm[threadIdx.x] = x; // touch shared memory
        }
        busy1 += rdclk() - tc;
        __syncthreads();
        tc = rdclk();
        // phase 2: nobody works
        busy2 += rdclk() - tc;
        __syncthreads();
        tc = rdclk();
    }
    tot = rdclk() - t00;
    if (lane == 0) { res[warp * 3] = busy1; res[warp * 3 + 1] = busy2; res[warp * 3 + 2] = tot; }
    out[threadIdx.x] = x + sm[(threadIdx.x + 1) & 511];
}
int main() {
    double *out; long long *res;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&res, 64 * 8);
    for (int work : {0, 250, 1000}) {
        k<0><<<1, 512>>>(out, res, work); k<0><<<1, 512>>>(out, res, work);
        long long h[48]; cudaMemcpy(h, res, sizeof(h), cudaMemcpyDeviceToHost);
        printf("work=%d (x8 cycles): total %lld cycles / 32 iterations = %.0f per iteration\n  busy1 per warp:", work, h[2], h[2] / 32.0);
        for (int w = 0; w < 16; w++) printf(" %lld", h[w * 3] / 32);
        printf("\n  busy2 per warp:");
        for (int w = 0; w < 16; w++) printf(" %lld", h[w * 3 + 1] / 32);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
