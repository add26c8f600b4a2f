// fdm_book.cuh -- a whole book of American options as one call: the fan-out of PDE solves behind
// price_log / price_log2 / greeks_log2 (fd_american_option_pricer.py:659-751, fd_american_black76.py:533-625)
// is expanded ON THE DEVICE from per-option scalars, solved, and combined by an epilogue kernel in the
// reference's operation order.  Host <-> device traffic is ~56 B in (+ 16 B per dividend) and ~60 B out per OPTION
// instead of ~104 B + ~108 B per SOLVE.  Barrier books (base solve + sigma bumps) the same way.
//
// Included by fdm_api.cu.

#pragma once

#include "fdm_common.cuh"

namespace book {

struct Role { int fine; double bump; };   // fine: 0 = n_time, 1 = price_log2's fine steps, 2 = 2*n_time

// role tables, in the order the reference executes the solves
__host__ __device__ inline int roles_of(int what, int rich, Role *r)
{
    int n = 0;
    if (what == FDM_BOOK_PRICE_LOG) { r[n++] = {0, 0.0}; return n; }
    if (what == FDM_BOOK_PRICE_LOG2 || what == FDM_BOOK_SCENARIO) {
        r[n++] = {0, 0.0};                              // pN,  :669
        if (rich) r[n++] = {1, 0.0};                    // p2N, :671 (2*num_space_nodes; Black-76 2*num_time_steps)
    }
    if (what == FDM_BOOK_GREEKS_LOG2 || what == FDM_BOOK_SCENARIO) {
        r[n++] = {0, 0.0};                              // V_N, :695
        if (rich) {
            r[n++] = {2, 0.0};                          // V_2N, :700
            r[n++] = {0, 1.0}; r[n++] = {0, -1.0};      // sigma +- h, :717-718
            r[n++] = {0, 2.0}; r[n++] = {0, -2.0};      // sigma +- 2h, :721-722
        } else {
            r[n++] = {0, 1.0}; r[n++] = {0, -1.0};      // :727-728
        }
    }
    return n;
}

struct RoleTable { Role r[8]; int n; };

// one thread per (role, option): descriptor of that solve (hostlogic.build_american_batch on the device), its
// time segments included (hostlogic.american_segments; fd_american_option_pricer.py:559-600).
// Segment slots are laid out role-major: solve (k, i) owns slots k S + div_begin[i] + i .. + n_div(i), S = n_div_total + n_options.
__global__ void expand_kernel(const FdmAmericanBook bk, const FdmBatch out, const RoleTable roles)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_opt = bk.n_options;
    const int64_t S = n_opt + (bk.div_begin ? bk.n_div_total : 0);
    if (gid == 0) ((int32_t *)out.seg_begin)[n_opt * roles.n] = (int32_t)(S * roles.n);
    if (gid >= n_opt * roles.n) return;
    const int k = (int)(gid / n_opt);
    const int64_t i = gid - (int64_t)k * n_opt;
    const Role role = roles.r[k];
    const int fl = bk.flags[i];
    const bool b76 = fl & FDM_FLAG_BLACK76;
    const bool is_call = fl & FDM_FLAG_CALL;
    const int n_space = max(bk.n_space[i], 3);                          // :142
    const int n_time0 = max(bk.n_time[i], 4);                           // :143
    int n_time = n_time0;
    if (role.fine == 1) n_time = b76 ? 2 * n_time0 : 2 * n_space;
    if (role.fine == 2) n_time = 2 * n_time0;
    const double spot = bk.spot[i], strike = bk.strike[i], T = bk.T[i];
    const double sigma = __dadd_rn(bk.sigma[i], __dmul_rn(role.bump, bk.dv_sigma));
    // what the reference's constructor rejects (:68-69, :123-124) cannot raise from a kernel: the solve is flagged
    bool invalid = !(spot > 0.0) || !(strike > 0.0) || !(sigma > 0.0) || !(T > 0.0);
    // _configure_grid, :209-229
    const double s_low = fmin(spot, strike), s_high = fmax(spot, strike);
    const double s_c = sqrt(fmax(__dmul_rn(s_low, s_high), 1e-12));
    const double band = __dmul_rn(__dmul_rn(bk.s_max_mult, sigma), sqrt(fmax(T, 1e-12)));
    const double x_c = log(s_c);
    double S_min = exp(__dsub_rn(x_c, __dmul_rn(0.5, band)));
    double S_max = exp(__dadd_rn(x_c, __dmul_rn(0.5, band)));
    S_min = fmax(fmin(S_min, __dmul_rn(0.5, s_low)), 1e-8);
    S_max = fmax(S_max, __dmul_rn(2.0, s_high));
    const double x_min = log(S_min);
    const double dx = __dsub_rn(log(S_max), x_min) / (double)n_space;    // :233-237
    ((double *)out.x_min)[gid] = x_min;
    ((double *)out.dx)[gid] = dx;
    ((double *)out.sigma)[gid] = sigma;
    ((double *)out.r)[gid] = bk.r[i];
    ((double *)out.b)[gid] = bk.b[i];
    ((double *)out.spot)[gid] = spot;
    ((double *)out.strike)[gid] = strike;
    ((int32_t *)out.n_space)[gid] = n_space;
    ((int32_t *)out.rannacher)[gid] = max(bk.rannacher[i], 0);          // :144

    // ---- time segments, :559-600 ------------------------------------------------------------------
    const int64_t d0 = bk.div_begin ? bk.div_begin[i] : 0;
    const int n_div = bk.div_begin ? bk.div_begin[i + 1] - bk.div_begin[i] : 0;
    if (b76 && n_div > 0) invalid = true;                               // fd_american_black76.py:89: no dividends
    const int64_t seg0 = (int64_t)k * S + d0 + i;
    int32_t *seg_steps = (int32_t *)out.seg_steps, *seg_restart = (int32_t *)out.seg_restart;
    double *seg_dt = (double *)out.seg_dt, *seg_tau0 = (double *)out.seg_tau0, *seg_div = (double *)out.seg_div;
    ((int32_t *)out.seg_begin)[gid] = (int32_t)seg0;
    const double base_dt = T / (double)n_time;                          // :563
    double tau_prev = 0.0;
    int remaining = n_time;
    for (int j = 0; j < n_div; ++j) {
        const double tau_j = bk.div_tau[d0 + j];
        const double len = __dsub_rn(tau_j, tau_prev);                  // :566-569
        const double ratio = len / base_dt;
        int ns = 1;
        if (ratio < 2.0e9) ns = max(1, (int)rint(ratio));               // int(round(.)): half to even, :572
        remaining -= ns;
        const double dt = len / (double)ns;                             // :412
        if (!(dt > 0.0)) invalid = true;                                // :498 divides by dt
        seg_steps[seg0 + j] = ns;
        seg_dt[seg0 + j] = dt;
        seg_tau0[seg0 + j] = tau_prev;
        seg_restart[seg0 + j] = (j == 0 || is_call) ? 1 : 0;            // :583-586
        seg_div[seg0 + j] = bk.div_cash[d0 + j];
        tau_prev = tau_j;
    }
    {
        const int ns = max(1, remaining);                               // :575
        const double dt = __dsub_rn(T, tau_prev) / (double)ns;
        if (!(dt > 0.0)) invalid = true;
        seg_steps[seg0 + n_div] = ns;
        seg_dt[seg0 + n_div] = dt;
        seg_tau0[seg0 + n_div] = tau_prev;
        seg_restart[seg0 + n_div] = (n_div == 0 || is_call || b76) ? 1 : 0;
        seg_div[seg0 + n_div] = 0.0;
    }
    ((int32_t *)out.flags)[gid] = (fl & (FDM_FLAG_CALL | FDM_FLAG_BLACK76)) | FDM_FLAG_AMERICAN | FDM_FLAG_SNAP |
                                  (invalid ? FDM_FLAG_INVALID : 0);
}

// one thread per option: Richardson / bump combinations in the reference's operation order
__global__ void combine_kernel(const FdmAmericanBook bk, const FdmOut res, const FdmBookOut o, const int n_roles)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = bk.n_options;
    if (i >= n) return;
    const bool rich = bk.use_richardson != 0;
    const int what = bk.what;
    const double h = bk.dv_sigma;
    const int64_t oi = o.out_index ? o.out_index[i] : i;               // where this option's results go
    auto P = [&](int k) { return res.price[(int64_t)k * n + i]; };
    auto D = [&](int k) { return res.delta[(int64_t)k * n + i]; };
    auto G = [&](int k) { return res.gamma[(int64_t)k * n + i]; };
    int st = 0;
    bool unsolved = false;
    for (int k = 0; k < n_roles; ++k) {
        const int sk = res.status[(int64_t)k * n + i];
        unsolved |= sk < 0;
        st = max(st, sk);
    }
    if (unsolved) {                                                     // a solve no kernel took: nothing to combine
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        if (o.status) o.status[oi] = FDM_STATUS_NOT_SOLVED;
        if (o.ex_lo) o.ex_lo[oi] = -1;
        if (o.ex_hi) o.ex_hi[oi] = -1;
        if (o.price_log2) o.price_log2[oi] = nan;
        if (o.price) o.price[oi] = nan;
        if (o.delta) o.delta[oi] = nan;
        if (o.gamma) o.gamma[oi] = nan;
        if (o.vega) o.vega[oi] = nan;
        if (o.theta) o.theta[oi] = nan;
        return;
    }
    if (o.status) o.status[oi] = st;
    if (o.ex_lo) o.ex_lo[oi] = res.idx[4 * i + 2];                       // exercise interval of the first (base) solve
    if (o.ex_hi) o.ex_hi[oi] = res.idx[4 * i + 3];
    int k0 = 0;
    if (what == FDM_BOOK_PRICE_LOG) {
        if (o.price_log2) o.price_log2[oi] = P(0);
        if (o.price) o.price[oi] = P(0);
        if (o.delta) o.delta[oi] = D(0);
        if (o.gamma) o.gamma[oi] = G(0);
        return;
    }
    if (what == FDM_BOOK_PRICE_LOG2 || what == FDM_BOOK_SCENARIO) {
        const double pN = P(0);
        double v = pN;
        if (rich) v = __dsub_rn(__dmul_rn(4.0, P(1)), pN) / 3.0;         // (4 p2N - pN) / 3, :672
        if (o.price_log2) o.price_log2[oi] = v;
        k0 = rich ? 2 : 1;
        if (what == FDM_BOOK_PRICE_LOG2) return;
    }
    double price, delta, gamma, dV;
    if (rich) {
        price = __dsub_rn(__dmul_rn(4.0, P(k0 + 1)), P(k0)) / 3.0;       // :704-706
        delta = __dsub_rn(__dmul_rn(4.0, D(k0 + 1)), D(k0)) / 3.0;
        gamma = __dsub_rn(__dmul_rn(4.0, G(k0 + 1)), G(k0)) / 3.0;
        const double D_h = __dsub_rn(P(k0 + 2), P(k0 + 3)) / __dmul_rn(2.0, h);      // :719
        const double D_2h = __dsub_rn(P(k0 + 4), P(k0 + 5)) / __dmul_rn(4.0, h);     // :723
        dV = __dsub_rn(__dmul_rn(4.0, D_h), D_2h) / 3.0;                 // :725
    } else {
        price = P(k0); delta = D(k0); gamma = G(k0);
        dV = __dsub_rn(P(k0 + 1), P(k0 + 2)) / __dmul_rn(2.0, h);        // :729
    }
    const double s0 = bk.sigma[i], S0 = bk.spot[i], r = bk.r[i], b = bk.b[i];
    // theta: :737-743 (q = 0); Black-76 has no carry term, fd_american_black76.py:614-617
    const double diff = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(0.5, s0), s0), S0), S0), gamma);
    double theta;
    if (bk.flags[i] & FDM_FLAG_BLACK76)
        theta = -__dsub_rn(diff, __dmul_rn(r, price));
    else
        theta = -__dsub_rn(__dadd_rn(diff, __dmul_rn(__dmul_rn(__dsub_rn(b, 0.0), S0), delta)), __dmul_rn(r, price));
    if (o.price) o.price[oi] = price;
    if (o.delta) o.delta[oi] = delta;
    if (o.gamma) o.gamma[oi] = gamma;
    if (o.vega) o.vega[oi] = dV / 100.0;                                  // :733
    if (o.theta) o.theta[oi] = theta;
}

// ---- barrier books (DiscreteBarrierCrankNicolsonLog.price / greeks, discrete_barrier_fdm_pricer_cn.py:429-466) ----

// one thread per (bump, option): the base solve and the sigma +- dv solves share the grid (frozen at the base
// volatility, :59-118) and the monitor table
__global__ void barrier_expand_kernel(const FdmBarrierBook bk, const FdmBatch out, const int n_roles)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_opt = bk.n_options;
    if (gid == 0) ((int32_t *)out.seg_begin)[n_opt * n_roles] = (int32_t)(n_opt * n_roles);
    if (gid >= n_opt * n_roles) return;
    const int k = (int)(gid / n_opt);
    const int64_t i = gid - (int64_t)k * n_opt;
    const double bump = k == 0 ? 0.0 : (k == 1 ? 1.0 : -1.0);                          // :450, :454
    const double sigma = __dadd_rn(bk.sigma[i], __dmul_rn(bump, bk.dv_sigma));
    const double T = bk.T[i], S0 = bk.S0[i];
    const int n_space = bk.n_space[i], n_time = bk.n_time[i];
    const bool invalid = !(T > 0.0) || !(sigma > 0.0) || !(S0 > 0.0) || n_space < 3 || n_time < 1;   // :61-66
    ((double *)out.x_min)[gid] = bk.x_min[i];
    ((double *)out.dx)[gid] = bk.dx[i];
    ((double *)out.sigma)[gid] = sigma;
    ((double *)out.r)[gid] = bk.r_disc[i];
    ((double *)out.b)[gid] = bk.b_carry[i];
    ((double *)out.spot)[gid] = S0;
    ((double *)out.strike)[gid] = bk.K[i];
    ((double *)out.barrier)[gid] = bk.barrier[i];
    ((double *)out.rebate)[gid] = bk.rebate[i];
    if (out.barrier_up) ((double *)out.barrier_up)[gid] = bk.barrier_up[i];
    ((int32_t *)out.n_space)[gid] = max(n_space, 3);
    ((int32_t *)out.flags)[gid] = (bk.flags[i] & (FDM_FLAG_CALL | FDM_FLAG_KO_DOWN | FDM_FLAG_KO_UP)) |
                                  (invalid ? FDM_FLAG_INVALID : 0);
    ((int32_t *)out.rannacher)[gid] = 0;                                               // pure CN, :242-248
    ((int32_t *)out.seg_begin)[gid] = (int32_t)gid;
    ((int32_t *)out.seg_steps)[gid] = max(n_time, 1);
    ((double *)out.seg_dt)[gid] = T / (double)n_time;                                  // :230
    ((double *)out.seg_tau0)[gid] = 0.0;
    ((int32_t *)out.seg_restart)[gid] = 0;
    ((double *)out.seg_div)[gid] = 0.0;
    if (out.mon_begin) {
        ((int32_t *)out.mon_begin)[gid] = bk.mon_begin[i];
        ((int32_t *)out.mon_end)[gid] = bk.mon_begin[i + 1];
    }
}

// one thread per option: vega by central bump and theta from the PDE identity, :440-457
__global__ void barrier_combine_kernel(const FdmBarrierBook bk, const FdmOut res, const FdmBookOut o, const int n_roles)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = bk.n_options;
    if (i >= n) return;
    int st = 0;
    bool unsolved = false;
    for (int k = 0; k < n_roles; ++k) {
        const int sk = res.status[(int64_t)k * n + i];
        unsolved |= sk < 0;
        st = max(st, sk);
    }
    const int64_t oi = o.out_index ? o.out_index[i] : i;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double price = nan, delta = nan, gamma = nan, vega = nan, theta = nan;
    if (!unsolved) {
        price = res.price[i]; delta = res.delta[i]; gamma = res.gamma[i];
        const double sg = bk.sigma[i], S0 = bk.S0[i], r = bk.r_disc[i], b = bk.b_carry[i];
        theta = -__dsub_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(0.5, sg), sg), S0), S0), gamma),
                                     __dmul_rn(__dmul_rn(b, S0), delta)),
                           __dmul_rn(r, price));                                       // :440-444
        if (n_roles == 3) vega = __dsub_rn(res.price[n + i], res.price[2 * n + i]) / __dmul_rn(2.0, bk.dv_sigma);   // :457
    }
    if (o.status) o.status[oi] = unsolved ? FDM_STATUS_NOT_SOLVED : st;
    if (o.price) o.price[oi] = price;
    if (o.delta) o.delta[oi] = delta;
    if (o.gamma) o.gamma[oi] = gamma;
    if (o.vega) o.vega[oi] = vega;
    if (o.theta) o.theta[oi] = theta;
}

}  // namespace book
