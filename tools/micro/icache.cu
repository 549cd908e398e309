// Instruction-cache capacity probe: a loop whose body is N independent FFMAs (straight-line, 16 bytes each), one warp.
// cycles per instruction jumps where the body stops fitting a cache level.
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__global__ void k(float *out, long long *res, int iters) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float m = 1.0001f, c = 1e-6f;
    long long t0 = 0;
    for (int it = 0; it < iters; it++) {
        if (it == 1) t0 = clock64(); // first pass warms the caches
#pragma unroll
        for (int i = 0; i < N / 8; i++) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) res[0] = t1 - t0;
    out[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <int N>
void run(float *out, long long *res, int nwarps) {
    const int iters = 9;
    k<N><<<1, 32 * nwarps>>>(out, res, iters);
    long long h; cudaMemcpy(&h, res, 8, cudaMemcpyDeviceToHost);
    printf("body %6d instr = %5.0f KB, %2d warps: %.2f cycles / instr / warp-pass\n", N, N * 16 / 1024.0, nwarps, (double)h / (iters - 1) / N);
}
int main() {
    float *out; long long *res;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&res, 8);
    for (int nw : {1, 4}) {
        run<256>(out, res, nw); run<512>(out, res, nw); run<1024>(out, res, nw); run<2048>(out, res, nw); run<4096>(out, res, nw);
        run<8192>(out, res, nw); run<16384>(out, res, nw); run<32768>(out, res, nw);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
