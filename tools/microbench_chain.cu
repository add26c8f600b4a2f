// Dev tool: throughput of the sweep pattern  c = fma(nm, c, V[k]); V[k] = c  with NCH independent chains per
// thread over a 16-element register array, W warps on one SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int NCH> __global__ void chain(double* out, long long* cyc, int n) {
    double V[16];
    for (int k = 0; k < 16; ++k) V[k] = out[k] + threadIdx.x * 1e-3;
    const double nm = out[20] + 0.3, ncc = out[21] + 0.2;
    constexpr int L = 16 / NCH;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        double c[NCH];
#pragma unroll
        for (int h = 0; h < NCH; ++h) c[h] = 1e-3 * h;
#pragma unroll
        for (int s = 0; s < L; ++s)
#pragma unroll
            for (int h = 0; h < NCH; ++h) { c[h] = fma(nm, c[h], V[h * L + s]); V[h * L + s] = c[h]; }
#pragma unroll
        for (int h = 0; h < NCH; ++h) c[h] = 1e-3 * h;
#pragma unroll
        for (int s = L - 1; s >= 0; --s)
#pragma unroll
            for (int h = 0; h < NCH; ++h) { c[h] = fma(ncc, c[h], V[h * L + s]); V[h * L + s] = c[h]; }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 16; ++k) s += V[k];
    out[64 + blockIdx.x * blockDim.x + threadIdx.x] = s; if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NCH> void run(double* out, long long* cyc) {
    long long h; int n = 4096;
    for (int w : {4, 8, 12, 16}) {
        chain<NCH><<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("chains/thread=%d warps=%2d: SMSP cycles per DFMA %.2f (per warp %.2f)\n", NCH, w, (double)h / (n * 32.0) / (w / 4.0), (double)h / (n * 32.0));
    }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 1 << 20);
    run<1>(out, cyc); run<2>(out, cyc); run<4>(out, cyc); run<8>(out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
