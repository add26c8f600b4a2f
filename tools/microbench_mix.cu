// Dev tool: does non-FP64 work issue for free next to DFMA on sm_100a?  One block of W warps on one SM;
// each iteration = 8 independent DFMA chains x 4 + per-DFMA extra instructions of one kind.
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int NX> __global__ void mix(double* out, long long* cyc, int n) {
    __shared__ double sh[1024];
    double a[8]; for (int j = 0; j < 8; ++j) a[j] = out[0] + j + threadIdx.x;
    int x[8]; for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 7 + j;
    double sv = out[1] + threadIdx.x;
    sh[threadIdx.x] = sv;
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a[j] = fma(a[j], b, c);
#pragma unroll
                for (int e = 0; e < NX; ++e) {
                    if (KIND == 1) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) & 7]), "r"(i)); }
                    if (KIND == 2) { x[j] = __shfl_xor_sync(0xffffffffu, x[j], 1 + e); }
                    if (KIND == 3) { sv += sh[(threadIdx.x + x[j]) & 1023]; }       // LDS.64 + DADD (dadd is fp64 pipe!)
                    if (KIND == 4) { int y; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(y) : "r"((unsigned)__cvta_generic_to_shared(&sh[(threadIdx.x * 2 + j) & 1023]))); x[j] ^= y; }
                    if (KIND == 5) { asm volatile("mov.b32 %0, %1;" : "=r"(x[j]) : "r"(x[(j + 3) & 7])); }
                }
            }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = sv; for (int j = 0; j < 8; ++j) s += a[j] + x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int KIND, int NX> void run(const char* name, double* out, long long* cyc) {
    long long h; int n = 2048;
    for (int w : {4, 8, 16}) {
        mix<KIND, NX><<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double per = (double)h / (n * 32.0);          // cycles per DFMA per warp
        printf("%-28s NX=%d warps=%2d: %.2f cyc per (DFMA+%d extra) per warp -> SMSP cycles per DFMA group %.2f\n", name, NX, w, per, NX, per / (w / 4.0));
    }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 1 << 20);
    run<0, 0>("DFMA only", out, cyc);
    run<1, 1>("DFMA + LOP3", out, cyc); run<1, 2>("DFMA + LOP3", out, cyc); run<1, 3>("DFMA + LOP3", out, cyc);
    run<2, 1>("DFMA + SHFL", out, cyc); run<2, 2>("DFMA + SHFL", out, cyc);
    run<4, 1>("DFMA + LDS.32+LOP", out, cyc); run<4, 2>("DFMA + LDS.32+LOP", out, cyc);
    run<5, 2>("DFMA + MOV", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
