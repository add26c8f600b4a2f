// Dev tool: DFMA throughput with 1, 2 or 3 distinct 64-bit register operands (register-file port pressure).
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND> __global__ void rf(double* out, long long* cyc, int n) {
    double a[8], b[8], c[8];
    for (int j = 0; j < 8; ++j) { a[j] = out[0] + j + threadIdx.x; b[j] = 1.0 + 1e-9 * (j + threadIdx.x) + out[1]; c[j] = 1e-9 * j + out[2]; }
    const double bb = 1.0000001, cc = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (KIND == 0) a[j] = fma(a[j], bb, cc);                 // 1 varying operand
                if (KIND == 1) a[j] = fma(a[j], b[j], cc);               // 2 varying
                if (KIND == 2) a[j] = fma(b[j], a[j], c[j]);             // 3 varying
                if (KIND == 3) a[j] = fma(b[j], a[(j + 1) & 7], c[j]);   // 3 varying, chain across regs (like e = fma(nm, e, V))
                if (KIND == 4) { c[j] = fma(b[j], a[j], c[j]); a[j] = c[j]; }
            }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += a[j] + b[j] + c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int KIND> void run(const char* name, double* out, long long* cyc) {
    long long h; int n = 2048;
    for (int w : {4, 8, 16}) {
        rf<KIND><<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-36s warps=%2d: SMSP cycles per DFMA %.2f\n", name, w, (double)h / (n * 32.0) / (w / 4.0));
    }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 1 << 20);
    run<0>("fma(a, const, const)", out, cyc);
    run<1>("fma(a, b[j], const)", out, cyc);
    run<2>("fma(b[j], a, c[j])", out, cyc);
    run<3>("fma(b[j], a[j+1], c[j])", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
