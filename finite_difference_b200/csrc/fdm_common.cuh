// fdm_common.cuh -- helpers shared by every kernel family of libfdm_b200 (warp, lean, CTA kernels and the book
// fan-out): grid nodes, payoffs, Dirichlet values, knock-out thresholds, the growth-factor matching, lane carries,
// the American projection, row layouts and the natural-cubic-spline dividend jump of the warp-per-solve kernels.
// Every translation unit includes it; everything lives in an anonymous namespace (one copy per unit, same source,
// so every caller gets bit-identical node values).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fdm_b200.h"
#include "fdm_internal.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SCAN_LEVELS = 5;
constexpr int DIRECT_TERMS = 6;
// tag for compile-time dispatch of the time loops (generic lambdas instantiated per carry form / warp role)
template <int N> struct IntC { static constexpr int value = N; };

#define FDM_CUDA(call)                                                                \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) return fdm_fail(FDM_ERR_CUDA, #call ": %s", cudaGetErrorString(e_)); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------

// s_j = exp(x_min + j*dx), fd_american_option_pricer.py:239-240 (two roundings, no FMA).
// noinline: every caller must get bit-identical node values.
__device__ __noinline__ double node_s(double x_min, double dx, int j)
{
    return exp(__dadd_rn(x_min, __dmul_rn((double)j, dx)));
}

// the same value inlined (explicit roundings: identical bits), for loops that want several exponentials in flight
__device__ __forceinline__ double node_s_inl(double x_min, double dx, int j)
{
    return exp(__dadd_rn(x_min, __dmul_rn((double)j, dx)));
}

__device__ __forceinline__ double payoff(double s, double k, bool is_call)
{
    double v = is_call ? __dsub_rn(s, k) : __dsub_rn(k, s);
    return v > 0.0 ? v : 0.0;
}

// Dirichlet values at time-to-maturity tau: fd_american_option_pricer.py:286-298,
// fd_american_black76.py:291-314, discrete_barrier_fdm_pricer_cn.py:153-171.
// put_smin: the date-aware barrier pricer's put value K e^{-r tau} * S_min * e^{(b-r) tau}
// (discrete_barrier_fdm_pricer.py:391, evaluated left to right).
__device__ __forceinline__ void boundary_values(bool is_call, bool b76, double r, double b, double k,
                                                double s_top, double tau, double &vmin, double &vmax,
                                                bool put_smin = false, double s_bot = 0.0)
{
    if (b76) {
        double disc = exp(__dmul_rn(-r, tau));
        if (is_call) { vmin = 0.0; vmax = __dmul_rn(disc, __dsub_rn(s_top, k)); }
        else         { vmin = __dmul_rn(disc, k); vmax = 0.0; }
    } else if (is_call) {
        vmin = 0.0;
        vmax = __dsub_rn(__dmul_rn(s_top, exp(__dmul_rn(__dsub_rn(b, r), tau))),
                         __dmul_rn(k, exp(__dmul_rn(-r, tau))));
    } else {
        vmin = __dmul_rn(k, exp(__dmul_rn(-r, tau)));
        if (put_smin) vmin = __dmul_rn(__dmul_rn(vmin, s_bot), exp(__dmul_rn(__dsub_rn(b, r), tau)));
        vmax = 0.0;
    }
}

// The moving Dirichlet value of a solve (call: top node, put: bottom node), out of line: the time loops evaluate it
// once per 32 steps, and inlined its exponentials made up a third of every specialised loop's code (instruction
// cache: several such loops are live per SM).
__device__ __noinline__ double moving_boundary_value(bool is_call, bool b76, double r, double b, double k, double s_top,
                                                     double tau, bool put_smin, double s_bot)
{
    double vmin, vmax;
    boundary_values(is_call, b76, r, b, k, s_top, tau, vmin, vmax, put_smin, s_bot);
    return is_call ? vmax : vmin;
}

// largest node index j in [-1, n] with s_j <= level (-1: none)
__device__ __forceinline__ int last_node_le(double x_min, double dx, int n, double level)
{
    if (!(level > 0.0)) return -1;
    double xj = (log(level) - x_min) / dx;
    if (xj < -2.0) return -1;
    if (xj > (double)n + 2.0) return n;
    int j = min(max((int)floor(xj), -1), n);
    while (j + 1 <= n && node_s(x_min, dx, j + 1) <= level) ++j;
    while (j >= 0 && node_s(x_min, dx, j) > level) --j;
    return j;
}

// smallest node index j in [0, n+1] with s_j >= level (n+1: none)
__device__ __forceinline__ int first_node_ge(double x_min, double dx, int n, double level)
{
    if (!(level > 0.0)) return 0;
    double xj = (log(level) - x_min) / dx;
    if (xj < -2.0) return 0;
    if (xj > (double)n + 2.0) return n + 1;
    int j = min(max((int)ceil(xj), 0), n + 1);
    while (j - 1 >= 0 && node_s(x_min, dx, j - 1) >= level) --j;
    while (j <= n && node_s(x_min, dx, j) < level) ++j;
    return j;
}

// Knock-out set of a solve as two thresholds: node j is knocked out iff j <= ko_a or j >= ko_b
// (discrete_barrier_fdm_pricer_cn.py:203-213; double-out: discrete_barrier_fdm_pricer.py:433-435).
__device__ __forceinline__ void ko_thresholds(const FdmBatch &B, int64_t sid, int flags, double x_min, double dx, int n,
                                              int &ko_a, int &ko_b)
{
    ko_a = -1;
    ko_b = n + 1;
    const bool dn = flags & FDM_FLAG_KO_DOWN, up = flags & FDM_FLAG_KO_UP;
    if (dn && B.barrier) ko_a = last_node_le(x_min, dx, n, B.barrier[sid]);
    if (up) {
        const double *lv = (dn && B.barrier_up) ? B.barrier_up : B.barrier;
        if (lv) ko_b = first_node_ge(x_min, dx, n, lv[sid]);
    }
}

// ---------------------------------------------------------------------------------------------
// matching the reference's discrete growth factor
// ---------------------------------------------------------------------------------------------
// In the row sums of A = I - theta dt L and B = I + (1-theta) dt L the large terms +-alpha dt cancel and
// leave 1 + O(r dt).  Any LU representation stores its factors rounded to 2^-53 relative, i.e. it solves
// with a matrix whose row sum is off by ~ulp(alpha dt); per step that scales smooth solution components
// by 1 + O(alpha dt 2^-53), and over the march by alpha*T*2^-53 (2e-10 at 8192 nodes).  The reference's
// Thomas sweep (fd_american_option_pricer.py:438-457) has such a bias too -- a different one.  Both are
// fixed numbers, known once the steady-state factors are known, so the difference is folded into the
// coefficient of V in the right-hand side:
//     reference, smooth component:  x = S_B / S_ref * V         S_ref = (A_L + den)(1 + c')   [den, c' = its factors]
//     here:                         x = (pd + q) / S_our * V - q V,   S_our = (1 - nm)(1 - ncc) / inv
//     =>  pd = S_our (q + S_B / S_ref) - q
// evaluated in double-double arithmetic.  pd = 1 + O(alpha dt 2^-53).

struct dd { double hi, lo; };

__device__ __forceinline__ dd dd_two_sum(double a, double b)
{
    const double s = __dadd_rn(a, b);
    const double bb = __dsub_rn(s, a);
    const double e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
    return {s, e};
}
__device__ __forceinline__ dd dd_two_prod(double a, double b)
{
    const double p = __dmul_rn(a, b);
    return {p, __fma_rn(a, b, -p)};
}
__device__ __forceinline__ dd dd_add(dd a, dd b)
{
    dd s = dd_two_sum(a.hi, b.hi);
    const double lo = __dadd_rn(s.lo, __dadd_rn(a.lo, b.lo));
    return dd_two_sum(s.hi, lo);
}
__device__ __forceinline__ dd dd_mul(dd a, dd b)
{
    dd p = dd_two_prod(a.hi, b.hi);
    const double lo = __dadd_rn(p.lo, __dadd_rn(__dmul_rn(a.hi, b.lo), __dmul_rn(a.lo, b.hi)));
    return dd_two_sum(p.hi, lo);
}
__device__ __forceinline__ dd dd_neg(dd a) { return {-a.hi, -a.lo}; }
__device__ __forceinline__ dd dd_div(dd a, dd b)
{
    const double q1 = a.hi / b.hi;
    dd r = dd_add(a, dd_neg(dd_mul(b, dd{q1, 0.0})));
    const double q2 = r.hi / b.hi;
    r = dd_add(r, dd_neg(dd_mul(b, dd{q2, 0.0})));
    const double q3 = r.hi / b.hi;
    return dd_add(dd_two_sum(q1, q2), dd{q3, 0.0});
}

// The reference's factors at their bit-wise fixed point: c'_0 = A_U/A_C, den_i = A_C - A_L c'_{i-1}, c'_i = A_U/den_i
// (fd_american_option_pricer.py:443-451).  The floating-point map den -> fl(A_C - fl(A_L fl(A_U / den))) is monotone
// non-decreasing and the reference's pivots fall from den_0 = A_C, so they stop at the LARGEST fixed point below
// A_C -- and so does the iteration from any start between that fixed point and A_C.  `den_start` lets a caller that
// knows the root of the quadratic start a hair above it (cta::setup_phase: 3x fewer iterations); A_C reproduces the
// reference's own sequence.  Returns false if no fixed point is reached within max_rows iterations.
__device__ __noinline__ bool reference_steady_state(double A_L, double A_C, double A_U, double den_start, int max_rows,
                                                    double &den_out, double &cp_out)
{
    double den = den_start, cp = A_U / den_start;
    bool fixed = false;
    for (int i = 1; i < max_rows && !fixed; ++i) {
        const double den_i = __dsub_rn(A_C, __dmul_rn(A_L, cp));
        const double cp_i = A_U / den_i;
        fixed = (den_i == den) && (cp_i == cp);
        den = den_i;
        cp = cp_i;
    }
    den_out = den;
    cp_out = cp;
    return fixed;
}

// pd from the reference's steady-state factors (den, cp) and this engine's converged pivot reciprocal rinv_inf
__device__ __noinline__ double growth_pd_from_ref(double A_L, double den, double cp, double A_U, double B_L, double B_C,
                                                  double B_U, double q, double rinv_inf)
{
    const double nm = -__dmul_rn(A_L, rinv_inf), ncc = -__dmul_rn(A_U, rinv_inf);
    const dd S_ref = dd_mul(dd_two_sum(A_L, den), dd_two_sum(1.0, cp));
    const dd S_our = dd_div(dd_mul(dd_two_sum(1.0, -nm), dd_two_sum(1.0, -ncc)), dd{rinv_inf, 0.0});
    const dd S_B = dd_add(dd_two_sum(B_L, B_C), dd{B_U, 0.0});
    const dd g = dd_add(dd{q, 0.0}, dd_div(S_B, S_ref));
    const dd pd = dd_add(dd_mul(S_our, g), dd{-q, 0.0});
    return pd.hi + pd.lo;
}

// rinv_inf: this engine's converged pivot reciprocal.  Returns pd.
// Grids too short for the reference's recurrence to reach its fixed point have no steady rows to match (and alpha dt
// is small there): pd = 1.
__device__ __forceinline__ double growth_matched_pd(double A_L, double A_C, double A_U, double B_L, double B_C,
                                                    double B_U, double q, double rinv_inf, int max_rows)
{
    double den, cp;
    if (!reference_steady_state(A_L, A_C, A_U, A_C, max_rows, den, cp)) return 1.0;
    return growth_pd_from_ref(A_L, den, cp, A_U, B_L, B_C, B_U, q, rinv_inf);
}

// The same fold against the EXACT row sums of A and B instead of the reference's rounded steady-state factors: the
// march then reproduces the exact-arithmetic growth factor S_B / S_A of smooth components, and what is left is the
// reference's own rounding bias, at most ~alpha T 2^-53 over the march.  fdm_lean_kernel uses it: its grids have at
// most 513 nodes, where alpha T <= 513^2 / (2 * 4.5^2) = 6500 bounds that bias by 7e-13 -- two orders below the 1e-10
// parity bar -- and the iteration to the reference's bit-wise fixed point (above) would cost as much as 10-20 time
// steps per phase.  The CTA kernels (up to 8193 nodes, bias up to 2e-10) keep growth_matched_pd.
__device__ __noinline__ double growth_exact_pd(double A_L, double A_C, double A_U, double B_L, double B_C, double B_U,
                                               double q, double rinv_inf)
{
    const double nm = -__dmul_rn(A_L, rinv_inf), ncc = -__dmul_rn(A_U, rinv_inf);
    const dd S_A = dd_add(dd_two_sum(A_L, A_C), dd{A_U, 0.0});
    const dd S_our = dd_div(dd_mul(dd_two_sum(1.0, -nm), dd_two_sum(1.0, -ncc)), dd{rinv_inf, 0.0});
    const dd S_B = dd_add(dd_two_sum(B_L, B_C), dd{B_U, 0.0});
    const dd g = dd_add(dd{q, 0.0}, dd_div(S_B, S_A));
    const dd pd = dd_add(dd_mul(S_our, g), dd{-q, 0.0});
    return pd.hi + pd.lo;
}

// carry entering each lane's chunk from the lanes below (UP = true) / above (UP = false)
template <bool UP>
__device__ __forceinline__ double lane_carry(const double e_end, const int terms, const double *__restrict__ W,
                                             const double *__restrict__ A, const int lane)
{
    if (terms) {
        // all shuffles are independent: ~one shuffle latency instead of five dependent levels
        double acc0, acc1;
        {
            const double t1 = UP ? __shfl_up_sync(FULL, e_end, 1) : __shfl_down_sync(FULL, e_end, 1);
            const double t2 = UP ? __shfl_up_sync(FULL, e_end, 2) : __shfl_down_sync(FULL, e_end, 2);
            acc0 = W[0 * 32 + lane] * t1;
            acc1 = W[1 * 32 + lane] * t2;
        }
        if (terms > 2) {
            const double t3 = UP ? __shfl_up_sync(FULL, e_end, 3) : __shfl_down_sync(FULL, e_end, 3);
            const double t4 = UP ? __shfl_up_sync(FULL, e_end, 4) : __shfl_down_sync(FULL, e_end, 4);
            acc0 = fma(W[2 * 32 + lane], t3, acc0);
            acc1 = fma(W[3 * 32 + lane], t4, acc1);
        }
        if (terms > 4) {
            const double t5 = UP ? __shfl_up_sync(FULL, e_end, 5) : __shfl_down_sync(FULL, e_end, 5);
            const double t6 = UP ? __shfl_up_sync(FULL, e_end, 6) : __shfl_down_sync(FULL, e_end, 6);
            acc0 = fma(W[4 * 32 + lane], t5, acc0);
            acc1 = fma(W[5 * 32 + lane], t6, acc1);
        }
        return acc0 + acc1;
    }
    double e = e_end;
#pragma unroll
    for (int lv = 0; lv < SCAN_LEVELS; ++lv) {
        const double t = UP ? __shfl_up_sync(FULL, e, 1 << lv) : __shfl_down_sync(FULL, e, 1 << lv);
        e = fma(A[lv * 32 + lane], t, e);
    }
    e = UP ? __shfl_up_sync(FULL, e, 1) : __shfl_down_sync(FULL, e, 1);
    return (UP ? lane == 0 : lane == 31) ? 0.0 : e;
}

// American right-hand side and projection (fdm_warp_kernel and fdm_cta_kernel; fdm_lean_kernel has its own scaled
// form).  The state carries Lq = 2 dt lambda: the relu of the projection is then ONE FP64 add, d + |d| (exactly 2 d
// or +0), instead of three integer instructions on the sign mask, and the factor 1/2 rides on FMAs that exist anyway
// (0.5 Lq is exact, so every value is bit-identical to the (V, dt lambda) form).  Measured on B200: fdm_cta_kernel
// +7 % at 8192 nodes, +12..17 % at 1000..4096 nodes.
#ifdef FDM_RELU_MASK            // the former form, for timing comparisons
constexpr bool RELU_FP64 = false;
#else
constexpr bool RELU_FP64 = true;
#endif
// fdm_warp_kernel (two warps per scheduler, latency-bound): +4 % with 16 slots per lane, no change with 8, -3 % with 4
template <int C>
__host__ __device__ constexpr bool warp_relu_fp64() { return RELU_FP64 && C > 8; }

template <int C, bool FP64_RELU>
__device__ __forceinline__ void american_rhs(double (&V)[C], double (&L)[C], const double q, const double pd)
{
    if (FP64_RELU) {
        if (q == 1.0) {                                   // theta = 1/2 (uniform across the solve)
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const double g = fma(0.5, L[k], V[k]);
                V[k] = fma(pd, V[k], g);
                L[k] = g;
            }
        } else {
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const double g = fma(q, V[k], 0.5 * L[k]);
                V[k] = fma(pd, V[k], g);
                L[k] = g;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const double g = fma(q, V[k], L[k]);
            V[k] = fma(pd, V[k], g);
            L[k] = g;
        }
    }
}

// exercise iff phi > vc (fd_american_option_pricer.py:496): dt lambda_new = max(phi - vc, 0), V = vc + dt lambda_new
template <bool FP64_RELU>
__device__ __forceinline__ void american_project(const double phi, const double vc, double &Vk, double &Lk)
{
    const double d = phi - vc;
    if (FP64_RELU) {
        Lk = d + fabs(d);
        Vk = fma(0.5, Lk, vc);
    } else {
        const int msk = ~(__double2hiint(d) >> 31);
        Lk = __hiloint2double(__double2hiint(d) & msk, __double2loint(d) & msk);
        Vk = vc + Lk;
    }
}

// How the n-1 interior rows are dealt to the 32 lanes of a warp: lane l holds rows row0 .. row0+cnt-1 in slots
// 0 .. cnt-1 of its C-slot register arrays.
//   uniform(C): every lane C slots, rows beyond n-2 are padding            (fdm_warp_kernel)
//   packed(n_int): lane l holds base + (l < extra) rows, no padding        (fdm_lean_kernel)
struct RowLayout {
    int row0, cnt;       // this lane
    int base, extra;     // base rows per lane; the first `extra` lanes hold one more
    __device__ static RowLayout uniform(int C, int lane) { return {lane * C, C, C, 0}; }
    __device__ static RowLayout packed(int n_int, int lane)
    {
        const int base = n_int >> 5, extra = n_int & 31;
        return {lane * base + min(lane, extra), base + (lane < extra ? 1 : 0), base, extra};
    }
    __device__ void owner(int row, int &src, int &kk) const
    {
        const int head = extra * (base + 1);
        if (row < head) { src = row / (base + 1); kk = row - src * (base + 1); }
        else { const int r2 = row - head; const int d = r2 / base; src = extra + d; kk = r2 - d * base; }
        if (src > 31) { src = 31; kk = -1; }
    }
};

// value at node j (0..n) from the distributed vector
template <int C>
__device__ __forceinline__ double node_value(const double (&V)[C], double V0, double VN, int j, int n,
                                             const RowLayout &ly)
{
    const int row = j - 1;
    int src = 0, kk = -1;
    if (row >= 0 && row < n - 1) ly.owner(row, src, kk);
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < C; ++k) mine = (k == kk) ? V[k] : mine;
    double v = __shfl_sync(FULL, mine, src);
    if (j <= 0) v = V0;
    if (j >= n) v = VN;
    return v;
}

// first-minimum argmin of |s_j - level| over all nodes j = 0..n (fd_american_option_pricer.py:253)
// or over interior nodes only (discrete_barrier_fdm_pricer_cn.py:335).
template <int C>
__device__ __forceinline__ int argmin_node(const double (&s)[C], double s0, double sn, double level,
                                           int n, const RowLayout &ly, bool interior_only)
{
    double best = INFINITY;
    int bidx = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = ly.row0 + k + 1;
        const double v = fabs(__dsub_rn(s[k], level));
        if (k < ly.cnt && j <= n - 1 && v < best) { best = v; bidx = j; }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const double ov = __shfl_xor_sync(FULL, best, d);
        const int oi = __shfl_xor_sync(FULL, bidx, d);
        if (ov < best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
    if (!interior_only) {
        const double v0 = fabs(__dsub_rn(s0, level));
        if (v0 <= best) { best = v0; bidx = 0; }
        const double vn = fabs(__dsub_rn(sn, level));
        if (vn < best) { best = vn; bidx = n; }
    }
    return bidx;
}

// Natural cubic spline through (s_j, V_j) and the ex-dividend jump V(S) <- V(S - D):
// fd_american_option_pricer.py:321-386 and :517-548, same operation order (no FMA).
// xs, ys, mu, cz: shared-memory arrays of n+1 doubles.
template <int C, class PhiAt>
__device__ __forceinline__ void dividend_jump(double (&V)[C], double &V0, double &VN, const PhiAt &phi,
                                           double phi0, double phin, double cash, bool is_call, int n,
                                           int lane, const RowLayout &ly, double x_min, double dx,
                                           double *xs, double *ys, double *mu, double *cz)
{
    const int nn = n + 1;   // number of knots
    __syncwarp();
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = ly.row0 + k + 1;
        if (k < ly.cnt && j <= n - 1) { xs[j] = node_s(x_min, dx, j); ys[j] = V[k]; }
    }
    if (lane == 0) {
        xs[0] = node_s(x_min, dx, 0); ys[0] = V0;
        xs[n] = node_s(x_min, dx, n); ys[n] = VN;
    }
    __syncwarp();
    // The spline's tridiagonal system (:338-359) is diagonally dominant: its forward recurrences forget their
    // start like 0.27^k, and so does the back substitution.  Each lane warm-starts G rows outside its own
    // rows (exactly at the true ends) and keeps only its own: all rows are eliminated in parallel, in the
    // reference's operation order, to within 0.27^G of the serial sweep.
    constexpr int G = 40;
    const int j_first = ly.row0 + 1;                       // this lane's interior knots [j_first, j_last]
    const int j_last = min(ly.row0 + ly.cnt, n - 1);
    const bool has_rows = ly.cnt > 0 && j_first <= n - 1;
    if (has_rows) {
        double mu_prev = 0.0, z_prev = 0.0;
        for (int i = max(1, j_first - G); i <= j_last; ++i) {
            const double h_i = __dsub_rn(xs[i + 1], xs[i]);
            const double h_im1 = __dsub_rn(xs[i], xs[i - 1]);
            const double alpha = __dsub_rn(
                __dmul_rn(3.0 / h_i, __dsub_rn(ys[i + 1], ys[i])),
                __dmul_rn(3.0 / h_im1, __dsub_rn(ys[i], ys[i - 1])));
            const double l = __dsub_rn(__dmul_rn(2.0, __dsub_rn(xs[i + 1], xs[i - 1])),
                                       __dmul_rn(h_im1, mu_prev));
            mu_prev = h_i / l;
            z_prev = __dsub_rn(alpha, __dmul_rn(h_im1, z_prev)) / l;
            if (i >= j_first) { mu[i] = mu_prev; cz[i] = z_prev; }
        }
    }
    if (lane == 0) { mu[0] = 0.0; cz[0] = 0.0; cz[nn - 1] = 0.0; }
    __syncwarp();
    // back substitution c[j] = z[j] - mu[j] c[j+1] (:358-359); results parked in V until every lane has read z
    if (has_rows) {
        double c_next = 0.0;
        for (int j = min(n - 1, j_last + G); j > j_last; --j) c_next = __dsub_rn(cz[j], __dmul_rn(mu[j], c_next));
#pragma unroll
        for (int k = C - 1; k >= 0; --k) {
            const int j = ly.row0 + k + 1;
            if (k < ly.cnt && j <= n - 1) { c_next = __dsub_rn(cz[j], __dmul_rn(mu[j], c_next)); V[k] = c_next; }
        }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = ly.row0 + k + 1;
        if (k < ly.cnt && j <= n - 1) cz[j] = V[k];
    }
    __syncwarp();
    const double s_lo = xs[0], s_hi = xs[n];
    auto eval = [&](double Sj, double phi_j) -> double {
        const double Sq = __dsub_rn(Sj, cash);
        double cont;
        if (Sq <= s_lo) cont = ys[0];
        else if (Sq >= s_hi) cont = ys[n];
        else {
            int lo = 0, hi = nn - 1;                       // :375-382
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (Sq < xs[mid]) hi = mid; else lo = mid;
            }
            const int j = lo;
            const double h = __dsub_rn(xs[j + 1], xs[j]);
            const double cj = cz[j], cj1 = cz[j + 1];
            // :360-364
            const double bj = __dsub_rn(__dsub_rn(ys[j + 1], ys[j]) / h,
                                        __dmul_rn(h, __dadd_rn(cj1, __dmul_rn(2.0, cj))) / 3.0);
            const double dj = __dsub_rn(cj1, cj) / __dmul_rn(3.0, h);
            const double dq = __dsub_rn(Sq, xs[j]);
            // :384  a + b*dx + c*dx*dx + d*dx*dx*dx, left to right
            double acc = __dadd_rn(ys[j], __dmul_rn(bj, dq));
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(cj, dq), dq));
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(dj, dq), dq), dq));
            cont = acc;
        }
        if (is_call) cont = cont > phi_j ? cont : phi_j;   // :542-544
        return cont;
    };
    double newV[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = ly.row0 + k + 1;
        newV[k] = (k < ly.cnt && j <= n - 1) ? eval(xs[j], phi(k)) : 0.0;
    }
    const double n0 = eval(s_lo, phi0);
    const double nN = eval(s_hi, phin);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < C; ++k) V[k] = newV[k];
    V0 = n0;
    VN = nN;
}

}  // namespace
