// Dev tool: FP64 pipe latency / throughput and shuffle latency on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat_dfma(double* out, long long* cyc, int n) {
    double a = out[0], b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = fma(a, b, c);
    }
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_shfl_dfma(double* out, long long* cyc, int n) {
    double a = out[0] + threadIdx.x, b = 1.0000001;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) { double t = __shfl_up_sync(0xffffffffu, a, 1); a = fma(b, t, a); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_shfl(double* out, long long* cyc, int n) {
    double a = out[0] + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) { a = __shfl_up_sync(0xffffffffu, a, 1); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_dsetp_sel(double* out, long long* cyc, int n) {
    double a = out[0] + threadIdx.x, p = out[1];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) { a = (p > a) ? p + 1e-9 * u : a + 1.0; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: W warps per block (1 block per SM), ILP independent chains each
template <int ILP> __global__ void thr_dfma(double* out, long long* cyc, int n) {
    double a[ILP]; for (int j = 0; j < ILP; ++j) a[j] = out[0] + j + threadIdx.x;
    double b = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < ILP; ++j) a[j] = fma(a[j], b, c);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < ILP; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 1 << 20);
    long long h; int n = 4096;
    lat_dfma<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DFMA dependent latency: %.2f cyc\n", (double)h / (n * 16));
    lat_shfl<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("SHFL(double=2x) dependent latency: %.2f cyc\n", (double)h / (n * 16));
    lat_shfl_dfma<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("SHFL+DFMA scan level latency: %.2f cyc\n", (double)h / (n * 16));
    lat_dsetp_sel<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DSETP+SEL+DADD level: %.2f cyc\n", (double)h / (n * 16));
    int ws[] = {1, 2, 4, 8, 12, 16, 32};
    for (int w : ws) {
        thr_dfma<1><<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); double t1 = (double)h / (n * 8);
        thr_dfma<2><<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); double t2 = (double)h / (n * 8 * 2);
        thr_dfma<4><<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); double t4 = (double)h / (n * 8 * 4);
        printf("warps/SM=%2d: cycles per DFMA (per warp) ILP1 %.2f  ILP2 %.2f  ILP4 %.2f   -> SM DFMA warp-instr/cyc ILP4 %.3f\n", w, t1, t2, t4, w / t4);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
