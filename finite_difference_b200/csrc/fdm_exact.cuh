// fdm_exact.cuh -- fdm_exact_kernel: the reference's step, literally, for the solves fdm_cta_kernel flags as
// FDM_STATUS_STIFF (pivots that do not settle inside its table, a row-0 tail that does not die inside the grid:
// alpha dt in the thousands, e.g. a dozen time steps on an 8193-node grid).
//
// One CTA per flagged solve; the value vector, the Thomas factors c'_i and the multiplier lambda live in shared
// memory (3 x 64 KB at 8193 nodes).  The tridiagonal solve is the reference's Thomas algorithm with per-row pivots
// (fd_american_option_pricer.py:438-457), run SERIALLY by thread 0 -- right-hand side (:474-483) and forward sweep
// fused in place, then back-substitution -- with the reference's operation order and true divisions; projection
// (:490-503), knock-out (discrete_barrier_fdm_pricer_cn.py:299-300) and the spline evaluation of a dividend jump
// (:517-548) run across the CTA.  Slow by construction (about a millisecond per time step at 8193 nodes) and exact:
// it is the safety net that makes every grid size solvable, not a path any benchmark configuration takes.
//
// Included by fdm_cta_inst.cu (the 16-warp translation unit hosts it); uses fdm_cta.cuh's helpers.

#pragma once

#include "fdm_cta.cuh"

namespace cta {

constexpr int EXACT_THREADS = 256;
__host__ __device__ constexpr size_t exact_smem_bytes(int n_space_max) { return (size_t)3 * (n_space_max + 2) * sizeof(double) + 64; }

static __global__ void __launch_bounds__(EXACT_THREADS, 1)
fdm_exact_kernel(const FdmBatch B, const FdmOut O, const int64_t first)
{
    extern __shared__ double smem[];
    __shared__ int sh[8];                                   // [0] todo count  [1..] scratch: ex lo / hi
    const int t = threadIdx.x;
    // one CTA screens EXACT_THREADS consecutive solves and takes the flagged ones in turn
    const int64_t base = first + (int64_t)blockIdx.x * EXACT_THREADS;
    for (int off = 0; off < EXACT_THREADS; ++off) {
        const int64_t sid = base + off;
        if (sid >= B.n_solves) break;
        if (!O.status || O.status[sid] != FDM_STATUS_STIFF) continue;      // uniform across the CTA
        const int flags = B.flags[sid];
        const int n = B.n_space[sid];
        const int n_int = n - 1;
        double *A = smem, *Cp = A + (n + 2), *Lm = Cp + (n + 2);
        const bool amer = flags & FDM_FLAG_AMERICAN, is_call = flags & FDM_FLAG_CALL, b76 = flags & FDM_FLAG_BLACK76;
        const bool snap = flags & FDM_FLAG_SNAP, ko = flags & (FDM_FLAG_KO_DOWN | FDM_FLAG_KO_UP);
        const bool put_smin = flags & FDM_FLAG_PUT_BDRY_SMIN;
        const double x_min = B.x_min[sid], dx = B.dx[sid], sigma = B.sigma[sid], r = B.r[sid], bcar = B.b[sid];
        const double spot = B.spot[sid], strike = B.strike[sid];
        const double rebate = (ko && B.rebate) ? B.rebate[sid] : 0.0;
        const int rannacher = B.rannacher[sid];
        int spot_idx = 0, strike_idx = 0;
        if (snap) {
            spot_idx = nearest_node(x_min, dx, n, spot, false);
            strike_idx = nearest_node(x_min, dx, n, strike, false);
        }
        int ko_a, ko_b;
        ko_thresholds(B, sid, flags, x_min, dx, n, ko_a, ko_b);
        const double k_eff = snap ? node_s(x_min, dx, strike_idx) : strike;
        const double s0 = node_s(x_min, dx, 0), sn = node_s(x_min, dx, n);
        const double sn_bdry = (flags & FDM_FLAG_BDRY_TOP_NEXT) ? node_s(x_min, dx, n + 1) : sn;
        const double b_bdry = B.b_bdry ? B.b_bdry[sid] : bcar;
        __syncthreads();
        for (int j = t; j <= n; j += EXACT_THREADS) A[j] = payoff(node_s(x_min, dx, j), k_eff, is_call);    // :558

        const double sig2 = __dmul_rn(sigma, sigma);
        const double mu_x = b76 ? __dmul_rn(-0.5, sig2) : __dsub_rn(bcar, __dmul_rn(0.5, sig2));
        const double alpha = __dmul_rn(0.5, sig2) / __dmul_rn(dx, dx);
        const double beta = mu_x / __dmul_rn(2.0, dx);
        const double ca = __dsub_rn(alpha, beta), cc = __dadd_rn(alpha, beta);
        const double bcoef = __dsub_rn(__dmul_rn(-2.0, alpha), r);

        const int seg_b = B.seg_begin[sid], seg_e = B.seg_begin[sid + 1];
        int mon_i = 0, mon_e = 0;
        if (ko && B.mon_begin) { mon_i = B.mon_begin[sid]; mon_e = B.mon_end ? B.mon_end[sid] : B.mon_begin[sid + 1]; }
        int next_mon = (mon_i < mon_e) ? B.mon_step[mon_i] : -1;
        int gstep = 0, ex_lo = -1, ex_hi = -1;
        int32_t *trace = O.ex_trace ? O.ex_trace + sid * O.ex_trace_stride : nullptr;
        for (int seg = seg_b; seg < seg_e; ++seg) {
            const int n_steps = B.seg_steps[seg];
            const double dt = B.seg_dt[seg], tau0 = B.seg_tau0[seg], cash = B.seg_div[seg];
            const int base_ran = B.seg_restart[seg] ? rannacher : 0;
            __syncthreads();
            for (int j = t; j < n_int; j += EXACT_THREADS) Lm[j] = 0.0;                             // :463
            double theta_built = -1.0;
            for (int m = 0; m < n_steps; ++m) {
                const double theta = (m < base_ran) ? 1.0 : 0.5;                                    // :469
                const double A_L = __dmul_rn(__dmul_rn(-theta, dt), ca);                            // :429-435
                const double A_C = __dsub_rn(1.0, __dmul_rn(__dmul_rn(theta, dt), bcoef));
                const double A_U = __dmul_rn(__dmul_rn(-theta, dt), cc);
                const double B_L = __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), ca);
                const double B_C = __dadd_rn(1.0, __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), bcoef));
                const double B_U = __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), cc);
                const double tau_next = fma((double)(m + 1), dt, tau0);
                double V_min, V_max;
                boundary_values(is_call, b76, r, b_bdry, k_eff, sn_bdry, tau_next, V_min, V_max, put_smin, s0);
                __syncthreads();
                if (t == 0) {
                    if (theta != theta_built) {                                                     // c'_i, :443-449
                        double den = A_C;
                        Cp[0] = A_U / den;
                        for (int i = 1; i < n_int; ++i) {
                            den = __dsub_rn(A_C, __dmul_rn(A_L, Cp[i - 1]));
                            Cp[i] = A_U / den;
                        }
                    }
                    // right-hand side and forward sweep fused, in place on the interior of A (row i = node i + 1)
                    double v_prev = A[0], d_prev = 0.0;
                    for (int i = 0; i < n_int; ++i) {
                        const double v_here = A[i + 1], v_next = A[i + 2];
                        double rhs = __dadd_rn(__dadd_rn(__dmul_rn(B_L, v_prev), __dmul_rn(B_C, v_here)), __dmul_rn(B_U, v_next));
                        if (amer) rhs = __dadd_rn(rhs, __dmul_rn(dt, Lm[i]));                      // + dt lambda_i, :480
                        if (i == 0) rhs = __dsub_rn(rhs, __dmul_rn(A_L, V_min));                    // :482-483
                        if (i == n_int - 1) rhs = __dsub_rn(rhs, __dmul_rn(A_U, V_max));
                        const double den = (i == 0) ? A_C : __dsub_rn(A_C, __dmul_rn(A_L, Cp[i - 1]));
                        d_prev = (i == 0) ? rhs / den : __dsub_rn(rhs, __dmul_rn(A_L, d_prev)) / den;
                        v_prev = v_here;
                        A[i + 1] = d_prev;
                    }
                    for (int i = n_int - 2; i >= 0; --i) A[i + 1] = __dsub_rn(A[i + 1], __dmul_rn(Cp[i], A[i + 2]));   // :455-456
                    A[0] = V_min;                                                                   // :505-506
                    A[n] = V_max;
                    sh[1] = 0x7fffffff;
                    sh[2] = -1;
                }
                theta_built = theta;
                __syncthreads();
                if (amer) {                                                                         // :490-503
                    int lo = 0x7fffffff, hi = -1;
                    for (int k = t; k < n_int; k += EXACT_THREADS) {
                        const double tilde = A[k + 1], phi = payoff(node_s(x_min, dx, k + 1), k_eff, is_call);
                        const double lam_old = Lm[k];
                        const double vc = __dsub_rn(tilde, __dmul_rn(dt, lam_old));                 // :495
                        const bool ex = phi > vc;                                                   // :496
                        const double ln = __dadd_rn(lam_old, __dsub_rn(phi, tilde) / dt);           // :498-500
                        Lm[k] = ln > 0.0 ? ln : 0.0;
                        A[k + 1] = ex ? phi : vc;
                        if (ex) { lo = min(lo, k); hi = max(hi, k); }
                    }
                    if (hi >= 0) { atomicMin(&sh[1], lo); atomicMax(&sh[2], hi); }
                    __syncthreads();
                    ex_lo = sh[2] >= 0 ? sh[1] : -1;
                    ex_hi = sh[2];
                }
                ++gstep;
                if (trace && t == 0 && 2 * (int64_t)gstep <= O.ex_trace_stride) {
                    trace[2 * (gstep - 1)] = ex_lo;
                    trace[2 * (gstep - 1) + 1] = ex_hi;
                }
                if (gstep == next_mon) {                                                            // knock-out projection
                    const double reb = B.mon_rebate ? B.mon_rebate[mon_i] : rebate;
                    __syncthreads();
                    for (int j = t; j <= n; j += EXACT_THREADS) if (j <= ko_a || j >= ko_b) A[j] = reb;
                    ++mon_i;
                    next_mon = (mon_i < mon_e) ? B.mon_step[mon_i] : -1;
                }
            }
            if (cash != 0.0) {
                // natural cubic spline, serial like the reference (:338-364): mu -> Cp, z then c -> Lm
                __syncthreads();
                if (t == 0) {
                    double x_im1 = s0, x_i = node_s(x_min, dx, 1), mu_prev = 0.0, z_prev = 0.0;
                    Cp[0] = 0.0;
                    Lm[0] = 0.0;
                    for (int i = 1; i < n; ++i) {
                        const double x_ip1 = node_s(x_min, dx, i + 1);
                        const double h_i = __dsub_rn(x_ip1, x_i), h_im1 = __dsub_rn(x_i, x_im1);
                        const double al = __dsub_rn(__dmul_rn(3.0 / h_i, __dsub_rn(A[i + 1], A[i])),
                                                    __dmul_rn(3.0 / h_im1, __dsub_rn(A[i], A[i - 1])));
                        const double l = __dsub_rn(__dmul_rn(2.0, __dsub_rn(x_ip1, x_im1)), __dmul_rn(h_im1, mu_prev));
                        mu_prev = h_i / l;
                        z_prev = __dsub_rn(al, __dmul_rn(h_im1, z_prev)) / l;
                        Cp[i] = mu_prev;
                        Lm[i] = z_prev;
                        x_im1 = x_i;
                        x_i = x_ip1;
                    }
                    Lm[n] = 0.0;
                    for (int j = n - 1; j >= 0; --j) Lm[j] = __dsub_rn(Lm[j], __dmul_rn(Cp[j], Lm[j + 1]));
                }
                __syncthreads();
                for (int jn = t; jn <= n; jn += EXACT_THREADS) {                                     // :517-548
                    const double Sj = node_s(x_min, dx, jn), Sq = __dsub_rn(Sj, cash);
                    double cont;
                    if (Sq <= s0) cont = A[0];
                    else if (Sq >= sn) cont = A[n];
                    else {
                        const int j = min(max(last_node_le(x_min, dx, n, Sq), 0), n - 1);
                        const double xj = node_s(x_min, dx, j), h = __dsub_rn(node_s(x_min, dx, j + 1), xj);
                        const double cj = Lm[j], cj1 = Lm[j + 1];
                        const double bj = __dsub_rn(__dsub_rn(A[j + 1], A[j]) / h, __dmul_rn(h, __dadd_rn(cj1, __dmul_rn(2.0, cj))) / 3.0);
                        const double dj = __dsub_rn(cj1, cj) / __dmul_rn(3.0, h);
                        const double dq = __dsub_rn(Sq, xj);
                        double acc = __dadd_rn(A[j], __dmul_rn(bj, dq));
                        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(cj, dq), dq));
                        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(dj, dq), dq), dq));
                        cont = acc;
                    }
                    if (is_call) { const double pv = payoff(Sj, k_eff, true); cont = cont > pv ? cont : pv; }
                    Cp[jn] = cont;                                               // mu is dead: park the new values
                }
                __syncthreads();
                for (int j = t; j <= n; j += EXACT_THREADS) A[j] = Cp[j];
            }
        }
        __syncthreads();
        if (O.v_dump) {
            double *dst = O.v_dump + sid * O.v_dump_stride;
            for (int j = t; j <= n && j < O.v_dump_stride; j += EXACT_THREADS) dst[j] = A[j];
        }
        if (t == 0) read_out(O, sid, A, n, snap, x_min, dx, spot, spot_idx, strike_idx, amer ? ex_lo : -1, amer ? ex_hi : -1, 0);
        __syncthreads();
    }
}

}  // namespace cta
