// Dev tool: throughput of FP64-pipe instruction kinds and operand shapes on sm_100a (16 warps on one SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND> __global__ void k(double* out, long long* cyc, int n) {
    double a[8], b[8], c[8], d[8];
    for (int j = 0; j < 8; ++j) { a[j] = out[0] + j + threadIdx.x; b[j] = 1.0 + 1e-9 * (j + threadIdx.x) + out[1]; c[j] = 1e-9 * j + out[2]; d[j] = out[3] + j; }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (KIND == 0) a[j] = a[j] + b[j];                                  // DADD
                if (KIND == 1) a[j] = a[j] * b[j];                                  // DMUL
                if (KIND == 2) { a[j] = fma(b[j], a[j], c[j]); c[j] = c[j] + b[j]; } // DFMA + DADD
                if (KIND == 3) { d[j] = fma(a[j], b[j], c[j]); a[j] = fma(d[j], b[j], c[j]); }  // dest != src
                if (KIND == 4) a[j] = fma(b[j], a[j], -c[j]);                        // negated addend
                if (KIND == 5) { double g = fma(b[0], a[j], c[j]); a[j] = g + a[j]; c[j] = g; }   // RHS pattern
                if (KIND == 6) { a[j] = a[j] - b[j]; }                                // DADD with negation
            }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += a[j] + b[j] + c[j] + d[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int KIND> void run(const char* name, int per_iter, double* out, long long* cyc) {
    long long h; int n = 2048;
    for (int w : {8, 16}) {
        k<KIND><<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s warps=%2d: SMSP cycles per FP64 instr %.2f\n", name, w, (double)h / (n * 32.0 * per_iter) / (w / 4.0));
    }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 1 << 20);
    run<0>("DADD a+=b", 1, out, cyc);
    run<1>("DMUL a*=b", 1, out, cyc);
    run<2>("DFMA + DADD", 2, out, cyc);
    run<3>("DFMA dest!=src x2", 2, out, cyc);
    run<4>("DFMA negated addend", 1, out, cyc);
    run<5>("RHS pattern fma+add", 2, out, cyc);
    run<6>("DADD a-=b", 1, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
