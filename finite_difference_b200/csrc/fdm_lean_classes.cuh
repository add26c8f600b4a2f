// fdm_lean_classes.cuh -- size classes of fdm_lean_kernel (host and device constexpr; shared with fdm_warp_kernel,
// which takes every size outside these classes and every solve the lean kernel hands back).
#pragma once

namespace lean {

constexpr unsigned long long RETRY_BITS = 0x7ff8b200dead0001ull;   // quiet NaN: "solve me with fdm_warp_kernel"

// Compiled size classes: a kernel with C slots per lane serves 32 (C-1) <= n_int <= 32 C interior rows
// (500 nodes -> 16, 400 -> 13, 300 -> 10).  Other sizes: fdm_warp_kernel.
// Measured on B200 (tools/time_lean.py, bench.py): 10..13 slots per lane (300..416 rows) gain +15..33 % over
// fdm_warp_kernel, 16 slots (500-node grids, cfg3) +5 %; below ~250 nodes the per-phase setup outweighs the extra
// resident warps, so those sizes stay with fdm_warp_kernel.  Define FDM_LEAN_ALL_CLASSES to compile every class
// (parity tests of the layout at the other sizes).
#ifdef FDM_LEAN_ALL_CLASSES
#define FDM_LEAN_CLASSES {4, 7, 8, 10, 13, 16}
#else
#define FDM_LEAN_CLASSES {10, 13, 16}
#endif
constexpr int N_CLASSES_MAX = 8;
struct ClassList { int c[N_CLASSES_MAX]; int n; };
__host__ __device__ constexpr ClassList class_list()
{
    constexpr int src[] = FDM_LEAN_CLASSES;
    ClassList l = {{0, 0, 0, 0, 0, 0, 0, 0}, (int)(sizeof(src) / sizeof(int))};
    for (int i = 0; i < l.n; ++i) l.c[i] = src[i];
    return l;
}
// slots per lane of the class that owns a grid with n_int interior rows (the smallest that fits); 0 = none
__host__ __device__ constexpr int class_of(int n_int)
{
    constexpr ClassList l = class_list();
    for (int i = 0; i < l.n; ++i)
        if (n_int >= 32 * (l.c[i] - 1) && n_int <= 32 * l.c[i]) return l.c[i];
    return 0;
}

}  // namespace lean
