// fdm_rowvar.cuh -- fdm_rowvar_kernel: theta-scheme time march with a ROW-VARYING tridiagonal matrix, one warp per solve.
//
// Reference: DiscreteBarrierFDMPricer2._solve_pde_backward (discrete_barrier_fdm_pricer_2.py:336-428): Black-Scholes in
// S (not log S) on a uniform grid with snapped nodes, so every row has its own coefficients; the two rows next to the
// knock-out barrier use a non-symmetric three-point stencil (:383-409); Dirichlet rows 0 and N are part of the
// system; knock-out projection after the steps of the monitoring map (:423-425, :241-257).  The analytic pricer's
// stepper (discrete_barrier_analytic_pricer.py:384) has the same shape.
//
// What one step computes, restated for the device (the reference rebuilds and factors the matrix in every step):
//   * the matrix depends on the row and on theta only, so the Thomas factorisation (:262-288: beta_i = main_i -
//     sub_i c*_{i-1}, c*_i = sup_i / beta_i) is hoisted out of the time loop, once per theta phase;
//   * rhs_i = a_expl_i V_{i-1} + b_expl_i V_i + c_expl_i V_{i+1} with the reference's own explicit coefficients
//     (standard rows carry its sign convention, :373-375), Dirichlet values on rows 0 and N (:359-367);
//   * d_i = rhs_i / beta_i - (sub_i / beta_i) d_{i-1} and x_i = d_i - c*_i x_{i+1} are first-order linear recurrences
//     with per-row multipliers: local pass over the lane's C rows, Kogge-Stone scan of the lane end values (the level
//     multipliers are tabulated per phase), second local pass with the carry.
// State: V in registers (C rows per lane, row = lane * C + k); six per-row coefficient tables in shared memory
// ([k][lane]: conflict-free).  HBM traffic: the grid and the terminal payoff in (16 B per node), one price out.
//
// Setup per phase: every lane runs the pivot recurrence over a warm-up of G rows below its first row (the recurrence
// forgets its start like c*^2 per row; G is sized per solve for 1e-18, see setup_phase) and then over its own rows --
// all lanes in parallel, equal to the serial sweep to rounding.
//
// Included by fdm_rowvar_inst.cu (one translation unit per size class).

#pragma once

#include "fdm_common.cuh"

namespace rowvar {

enum { T_AE = 0, T_BE, T_CE, T_RB, T_MM, T_NCS, T_N };        // explicit sub/diag/sup, 1/beta, -sub/beta, -c*

// IMPL: the kernel instance for marches whose every step is fully implicit (rannacher_steps >= n_steps -- the regime
// in which the reference class prices at all).  With theta = 1 the explicit operator is the
// identity (a_expl = c_expl = 0, b_expl = 1, :373-375 and :404-406), so the right-hand side is V itself: three tables
// and three of the eight shared-memory loads per node-step go, and so does half of the shared memory per warp (the
// march is bound by shared-memory bandwidth and by the number of resident warps).  Marches with a theta = 1/2 phase run
// in the general instance; every launch issues both and each warp decides by its solve.
// per-lane constants of a phase: products of the lane's multipliers over its lower / upper half of rows, forward and
// backward (the local passes run as two half-chains), and over all its rows backward
enum { LC_PF1 = 0, LC_PF2, LC_PB1, LC_PB2, LC_PBALL, LC_N };

template <bool IMPL> __host__ __device__ constexpr int n_tables() { return IMPL ? 3 : (int)T_N; }
template <bool IMPL> __host__ __device__ constexpr int tab_of(const int t) { return IMPL ? t - T_RB : t; }

template <int C, bool IMPL>
__host__ __device__ constexpr int smem_doubles() { return n_tables<IMPL>() * 32 * C + (2 * SCAN_LEVELS + LC_N) * 32; }

#ifndef FDM_ROWVAR_IMPL_BLOCKS
#define FDM_ROWVAR_IMPL_BLOCKS 12
#endif
template <int C, bool IMPL>
constexpr int min_blocks() { return (IMPL && C > 10 && C <= 19) ? FDM_ROWVAR_IMPL_BLOCKS : (C <= 10 ? 8 : (C <= 19 ? 4 : 2)); }

struct RowCoef { double sub, diag, sup, ae, be, ce; };

// coefficients of row i (discrete_barrier_fdm_pricer_2.py:369-416); rows 0 and N are Dirichlet rows (:364-365), rows
// beyond N are padding (identity, zero right-hand side)
__device__ __forceinline__ RowCoef row_coef(const int i, const int N, const double *__restrict__ s, const double sig,
                                            const double r, const double dt, const double theta, const double dS,
                                            const int bar_j, const double hm, const double hp)
{
    RowCoef c = {0.0, 1.0, 0.0, 0.0, 0.0, 0.0};
    if (i <= 0 || i >= N) return c;
    const double S = s[i];
    const double sS = __dmul_rn(sig, S), sig2S2 = __dmul_rn(sS, sS);
    double a_impl, b_impl, c_impl;
    if (bar_j < 0 || (i != bar_j && i != bar_j + 1)) {
        const double dS2 = __dmul_rn(dS, dS);
        const double diff = sig2S2 / dS2, conv = __dmul_rn(r, S) / dS;
        const double hi = __dmul_rn(__dmul_rn(0.5, dt), theta), he = __dmul_rn(__dmul_rn(-0.5, dt), __dsub_rn(1.0, theta));
        a_impl = __dmul_rn(hi, __dsub_rn(diff, conv));
        b_impl = __dadd_rn(1.0, __dmul_rn(__dmul_rn(dt, theta), __dadd_rn(diff, r)));
        c_impl = __dmul_rn(hi, __dadd_rn(diff, conv));
        c.ae = __dmul_rn(he, __dsub_rn(diff, conv));                            // the reference's sign, :373
        c.be = __dsub_rn(1.0, __dmul_rn(__dmul_rn(dt, __dsub_rn(1.0, theta)), __dadd_rn(diff, r)));
        c.ce = __dmul_rn(he, __dadd_rn(diff, conv));
    } else {
        const double a1 = hp / __dmul_rn(hm, __dadd_rn(hm, hp));
        const double b1 = __dsub_rn(hp, hm) / __dmul_rn(hm, hp);
        const double c1 = -hm / __dmul_rn(hp, __dadd_rn(hm, hp));
        const double d2 = 2.0 / __dmul_rn(hm, __dadd_rn(hm, hp));
        const double e2 = -2.0 / __dmul_rn(hm, hp);
        const double f2 = 2.0 / __dmul_rn(hp, __dadd_rn(hm, hp));
        const double rS = __dmul_rn(r, S), hs = __dmul_rn(0.5, sig2S2);
        const double L_left = __dadd_rn(__dmul_rn(hs, f2), __dmul_rn(rS, c1));
        const double L_center = __dsub_rn(__dadd_rn(__dmul_rn(hs, e2), __dmul_rn(rS, b1)), r);
        const double L_right = __dadd_rn(__dmul_rn(hs, d2), __dmul_rn(rS, a1));
        const double ti = __dmul_rn(-theta, dt), te = __dmul_rn(__dsub_rn(1.0, theta), dt);
        a_impl = __dmul_rn(ti, L_left);
        b_impl = __dsub_rn(1.0, __dmul_rn(__dmul_rn(theta, dt), L_center));
        c_impl = __dmul_rn(ti, L_right);
        c.ae = __dmul_rn(te, L_left);
        c.be = __dadd_rn(1.0, __dmul_rn(te, L_center));
        c.ce = __dmul_rn(te, L_right);
    }
    c.sub = -a_impl;
    c.diag = b_impl;
    c.sup = -c_impl;
    return c;
}

// Coefficient tables and scan multipliers of one theta phase.  Returns false on a vanishing pivot (the reference raises
// ZeroDivisionError, :271-279).
template <int C, bool IMPL>
__device__ __noinline__ bool setup_phase(double *tab, double *Af, double *Ab, const int lane, const int N,
                                         const double *__restrict__ s, const double sig, const double r, const double dt,
                                         const double theta, const double dS, const int bar_j, const double hm, const double hp)
{
    // Warm-up length: an error in c*_{i-1} shrinks by sub_i sup_{i-1} / beta^2 ~ c*^2 per row, and c* grows with
    // x = theta dt sigma^2 S^2 / (2 dS^2) towards the root of x c^2 - (1 + 2x) c + x = 0.  The top row has the largest x:
    // its c*^2 bounds the contraction everywhere, G rows leave c*^(2G) < 1e-18 (x = 30: 112 rows, x = 800: 540).
    int G;
    {
        const double St = s[max(N - 1, 1)];
        const double x = 0.5 * theta * dt * (sig * St) * (sig * St) / (dS * dS);
        const double cinf = (1.0 + 2.0 * x - sqrt(1.0 + 4.0 * x)) / fmax(2.0 * x, 1e-300);
        const double g = (cinf > 1e-9 && cinf < 1.0) ? -41.5 / (2.0 * log(cinf)) : 8.0;
        G = (g < 32.0 * C) ? max(16, (int)g + 2) : 32 * C;
    }
    const int row_first = lane * C;
    bool ok = true;
    // pivot recurrence: warm-up below the lane's first row, then the own rows
    double cstar = 0.0;
    for (int i = max(0, row_first - G); i < row_first; ++i) {
        const RowCoef c = row_coef(i, N, s, sig, r, dt, theta, dS, bar_j, hm, hp);
        const double beta = __dsub_rn(c.diag, __dmul_rn(c.sub, cstar));
        cstar = (i < N) ? c.sup / beta : 0.0;
    }
    double pm = 1.0, pc = 1.0;                                   // products of the lane's forward / backward multipliers
    double pf1 = 1.0, pf2 = 1.0, pb1 = 1.0, pb2 = 1.0;           // ... over its lower (k < C/2) and upper rows
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k < C; ++k) {
        const int i = row_first + k;
        const RowCoef c = row_coef(i, N, s, sig, r, dt, theta, dS, bar_j, hm, hp);
        const double beta = __dsub_rn(c.diag, __dmul_rn(c.sub, cstar));
        ok &= fabs(beta) >= 1e-14;
        // Rows N (the upper Dirichlet row, x_N = w_hi, :361-365) and beyond (padding) take no right-hand side (1 / beta
        // stored as 0) and pass the backward carry through (-c* stored as 1): w_hi enters as the carry into the top of the
        // warp and arrives at row N unchanged, wherever row N sits in its lane -- no per-row select in the time loop.
        const double rb = (i < N) ? 1.0 / beta : 0.0;
        cstar = (i < N) ? c.sup / beta : -1.0;                   // c*_N = 0 (:280)
        const double mm = (i < N) ? -__dmul_rn(c.sub, 1.0 / beta) : 0.0;
        if (!IMPL) {
            tab[(T_AE * C + k) * 32 + lane] = c.ae;
            tab[(T_BE * C + k) * 32 + lane] = c.be;
            tab[(T_CE * C + k) * 32 + lane] = c.ce;
        }
        tab[(tab_of<IMPL>(T_RB) * C + k) * 32 + lane] = rb;
        tab[(tab_of<IMPL>(T_MM) * C + k) * 32 + lane] = mm;
        tab[(tab_of<IMPL>(T_NCS) * C + k) * 32 + lane] = -cstar;
        pm *= mm;
        pc *= -cstar;
        if (k < C / 2) { pf1 *= mm; pb2 *= -cstar; } else { pf2 *= mm; pb1 *= -cstar; }
        if (i >= N) cstar = 0.0;                                 // (what the pivot recurrence of a later row would see)
    }
    double *lc = Ab + SCAN_LEVELS * 32;
    lc[LC_PF1 * 32 + lane] = pf1; lc[LC_PF2 * 32 + lane] = pf2;
    lc[LC_PB1 * 32 + lane] = pb1; lc[LC_PB2 * 32 + lane] = pb2; lc[LC_PBALL * 32 + lane] = pc;
    // Kogge-Stone level multipliers (as fdm_warp_kernel's setup_phase)
    double af = pm, ab = pc;
#pragma unroll
    for (int lv = 0; lv < SCAN_LEVELS; ++lv) {
        const int d = 1 << lv;
        const bool has_f = lane >= d, has_b = lane + d < 32;
        Af[lv * 32 + lane] = has_f ? af : 0.0;
        Ab[lv * 32 + lane] = has_b ? ab : 0.0;
        const double tf = __shfl_up_sync(FULL, af, d), tb = __shfl_down_sync(FULL, ab, d);
        if (has_f) af *= tf;
        if (has_b) ab *= tb;
    }
    __syncwarp();
    return !__any_sync(FULL, !ok);
}

template <int C, bool IMPL>
__global__ void __launch_bounds__(32, min_blocks<C, IMPL>())
fdm_rowvar_kernel(const FdmRowVarBatch B, const FdmRowVarOut O, const int64_t first, const int n_lo)
{
    extern __shared__ double smem[];
    double *tab = smem, *Af = tab + n_tables<IMPL>() * 32 * C, *Ab = Af + SCAN_LEVELS * 32;
    constexpr int tRB = tab_of<IMPL>(T_RB), tMM = tab_of<IMPL>(T_MM), tNCS = tab_of<IMPL>(T_NCS);

    const int64_t sid = first + blockIdx.x;
    if (sid >= B.n_solves) return;
    const int lane = threadIdx.x;
    const int64_t nb = B.node_begin[sid];
    const int n_nodes = (int)(B.node_begin[sid + 1] - nb);
    if (n_nodes <= n_lo || n_nodes > 32 * C) return;             // another size class
    const int N = n_nodes - 1;
    const double *__restrict__ s = B.s_nodes + nb;
    const int flags = B.flags[sid];
    const bool is_call = flags & FDM_FLAG_CALL;
    const double sig = B.sigma[sid], r = B.r[sid], strike = B.strike[sid], tenor = B.tenor[sid], dt = B.dt[sid];
    const int M = B.n_steps[sid], rann = B.rannacher[sid];
    if ((rann >= M) != IMPL) return;                             // the other instance's solve
    const int bar_j = B.bar_j[sid];
    const double hm = B.h_minus[sid], hp = B.h_plus[sid];
    const double dS = __dsub_rn(s[1], s[0]);                     // :345
    const double s_top = s[N];

    // knocked-out rows of this lane as a bit mask (:241-257; the levels may be absent: NaN compares false)
    unsigned ko_mask = 0;
    {
        const double lo = (flags & FDM_FLAG_KO_DOWN) ? B.lo_bar[sid] : __longlong_as_double(0x7ff8000000000000ll);
        const double up = (flags & FDM_FLAG_KO_UP) ? B.up_bar[sid] : __longlong_as_double(0x7ff8000000000000ll);
        for (int k = 0; k < C; ++k) {
            const int i = lane * C + k;
            if (i <= N && (s[i] <= lo || s[i] >= up)) ko_mask |= 1u << k;
        }
    }
    // IMPL: the right-hand side g = V / beta is formed in place (no R[]), and up to 19 rows per lane the two sweeps'
    // multipliers live in the registers that frees -- one shared-memory load per node-step (1 / beta) instead of five.
    constexpr bool REGTAB = IMPL && C <= 19;
    double V[C], R[IMPL ? 1 : C], MMr[REGTAB ? C : 1], NCr[REGTAB ? C : 1];
    double (&D)[C] = *reinterpret_cast<double (*)[C]>(IMPL ? &V[0] : &R[0]);
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int i = lane * C + k;
        V[k] = (i <= N) ? B.v_terminal[nb + i] : 0.0;
    }
    auto mm_of = [&](const int k) -> double { return REGTAB ? MMr[k] : tab[(tMM * C + k) * 32 + lane]; };
    auto ncs_of = [&](const int k) -> double { return REGTAB ? NCr[k] : tab[(tNCS * C + k) * 32 + lane]; };
    int mon_i = B.mon_begin ? B.mon_begin[sid] : 0;
    const int mon_e = B.mon_begin ? B.mon_begin[sid + 1] : 0;
    int next_mon = (mon_i < mon_e) ? B.mon_step[mon_i] : -1;

    bool bad = !(dS != 0.0) || !(__dmul_rn(dS, dS) != 0.0);      // the reference divides by dS and dS^2 (:379): ZeroDivisionError
    int step = 0;                                                // completed steps; the reference's m = M - step
    int fill_base = -64;
    double my_w = 0.0;
    while (step < M && !bad) {
        const bool implicit_phase = step < rann;                 // theta = 1 while (M - m) < rannacher_steps, :351
        const int step_end = implicit_phase ? min(rann, M) : M;
        if (!setup_phase<C, IMPL>(tab, Af, Ab, lane, N, s, sig, r, dt, implicit_phase ? 1.0 : 0.5, dS, bar_j, hm, hp)) {
            bad = true;
            break;
        }
        if (REGTAB) {
#pragma unroll
            for (int k = 0; k < C; ++k) { MMr[k] = tab[(tMM * C + k) * 32 + lane]; NCr[k] = tab[(tNCS * C + k) * 32 + lane]; }
        }
#pragma unroll 1
        for (; step < step_end; ++step) {
            // Dirichlet values of this step, tau_left = T - (m - 1) dt (:359-363): 32 steps per refill, one exponential per lane
            if (step - fill_base >= 32) {
                fill_base = step;
                const double tau_left = __dsub_rn(tenor, __dmul_rn((double)(M - (step + lane) - 1), dt));
                my_w = __dmul_rn(strike, exp(__dmul_rn(-r, tau_left)));
            }
            const double kd = __shfl_sync(FULL, my_w, step - fill_base);
            const double w_lo = is_call ? 0.0 : kd, w_hi = is_call ? __dsub_rn(s_top, kd) : 0.0;
            // right-hand side
            double below = 0.0, above = 0.0;
            if (!IMPL) { below = __shfl_up_sync(FULL, V[C - 1], 1); above = __shfl_down_sync(FULL, V[0], 1); }
#pragma unroll
            for (int k = 0; k < C; ++k) {
                double rhs = V[k];                               // theta = 1: 0 V_{i-1} + 1 V_i + 0 V_{i+1}
                if (!IMPL) {
                    const double vm = (k == 0) ? below : V[k - 1], vp = (k == C - 1) ? above : V[k + 1];
                    rhs = __dmul_rn(tab[(T_AE * C + k) * 32 + lane], vm);
                    rhs = fma(tab[(T_BE * C + k) * 32 + lane], V[k], rhs);
                    rhs = fma(tab[(T_CE * C + k) * 32 + lane], vp, rhs);
                }
                if (k == 0 && lane == 0) rhs = w_lo;
                D[k] = __dmul_rn(rhs, tab[(tRB * C + k) * 32 + lane]);           // g_i = rhs_i / beta_i (0 from row N on)
            }
            constexpr int H = C / 2;
            const double *lc = Ab + SCAN_LEVELS * 32;
            // forward: d_i = g_i + mm_i d_{i-1}; two half-chains per lane
            double e1 = 0.0, e2 = 0.0;
#pragma unroll
            for (int k = 0; k < H; ++k) e1 = fma(mm_of(k), e1, D[k]);
#pragma unroll
            for (int k = H; k < C; ++k) e2 = fma(mm_of(k), e2, D[k]);
            double e = fma(lc[LC_PF2 * 32 + lane], e1, e2);
#pragma unroll
            for (int lv = 0; lv < SCAN_LEVELS; ++lv) e = fma(Af[lv * 32 + lane], __shfl_up_sync(FULL, e, 1 << lv), e);
            double c = __shfl_up_sync(FULL, e, 1);
            c = (lane == 0) ? 0.0 : c;
            {
                double c1 = c, c2 = fma(lc[LC_PF1 * 32 + lane], c, e1);
#pragma unroll
                for (int k = 0; k < H; ++k) { c1 = fma(mm_of(k), c1, D[k]); D[k] = c1; }
#pragma unroll
                for (int k = H; k < C; ++k) { c2 = fma(mm_of(k), c2, D[k]); D[k] = c2; }
            }
            // backward: x_i = d_i - c*_i x_{i+1}; w_hi is the carry into the top of the warp
            e1 = 0.0; e2 = 0.0;
#pragma unroll
            for (int k = C - 1; k >= H; --k) e1 = fma(ncs_of(k), e1, D[k]);
#pragma unroll
            for (int k = H - 1; k >= 0; --k) e2 = fma(ncs_of(k), e2, D[k]);
            e = fma(lc[LC_PB2 * 32 + lane], e1, e2);
            if (lane == 31) e = fma(lc[LC_PBALL * 32 + lane], w_hi, e);
#pragma unroll
            for (int lv = 0; lv < SCAN_LEVELS; ++lv) e = fma(Ab[lv * 32 + lane], __shfl_down_sync(FULL, e, 1 << lv), e);
            c = __shfl_down_sync(FULL, e, 1);
            c = (lane == 31) ? w_hi : c;
            {
                double c1 = c, c2 = fma(lc[LC_PB1 * 32 + lane], c, e1);
#pragma unroll
                for (int k = C - 1; k >= H; --k) { c1 = fma(ncs_of(k), c1, D[k]); V[k] = c1; }
#pragma unroll
                for (int k = H - 1; k >= 0; --k) { c2 = fma(ncs_of(k), c2, D[k]); V[k] = c2; }
            }
            if (step + 1 == next_mon) {                          // knock-out projection, :423-425
#pragma unroll
                for (int k = 0; k < C; ++k) if (ko_mask & (1u << k)) V[k] = 0.0;
                ++mon_i;
                next_mon = (mon_i < mon_e) ? B.mon_step[mon_i] : -1;
            }
        }
    }

    // ---- read-out: linear interpolation at the effective spot (:433-449, :478-482) ----------------------------------
    if (O.v_dump) {
        double *dst = O.v_dump + sid * O.v_dump_stride;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int i = lane * C + k;
            if (i <= N && i < O.v_dump_stride) dst[i] = V[k];
        }
    }
    const double x = B.s_eff[sid];
    int lo = 0, hi = N;
    if (x > s[0] && x < s[N]) {
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (x < s[mid]) hi = mid; else lo = mid;
        }
    } else if (x <= s[0]) { lo = 0; hi = 0; }
    else { lo = N; hi = N; }
    auto value_at = [&](const int i) -> double {
        const int src = i / C, kk = i - src * C;
        double mine = 0.0;
#pragma unroll
        for (int k = 0; k < C; ++k) mine = (k == kk) ? V[k] : mine;
        return __shfl_sync(FULL, mine, src);
    };
    const double y0 = value_at(lo), y1 = value_at(hi);
    double price = y0;
    if (hi != lo) {
        const double x0 = s[lo], x1 = s[hi];
        const double w = __dsub_rn(x, x0) / __dsub_rn(x1, x0);
        price = __dadd_rn(__dmul_rn(__dsub_rn(1.0, w), y0), __dmul_rn(w, y1));
    }
    if (lane == 0) {
        O.price[sid] = bad ? __longlong_as_double(0x7ff8000000000000ll) : price;
        if (O.status) O.status[sid] = bad ? FDM_STATUS_SINGULAR : (isfinite(price) ? FDM_STATUS_OK : FDM_STATUS_NONFINITE);
    }
}

}  // namespace rowvar
