// fdm_book.cuh -- a whole book of American options as one call: the fan-out of PDE solves behind
// price_log / price_log2 / greeks_log2 (fd_american_option_pricer.py:659-751, fd_american_black76.py:533-625)
// is expanded ON THE DEVICE from per-option scalars, solved, and combined by an epilogue kernel in the
// reference's operation order.  Host <-> device traffic is ~56 B in and ~60 B out per OPTION instead of
// ~104 B + ~108 B per SOLVE.  Dividend-free books only (dividend segment tables are built on the host).
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

// one thread per (role, option): descriptor of that solve (hostlogic.build_american_batch on the device)
__global__ void expand_kernel(const FdmAmericanBook bk, const FdmBatch out, const RoleTable roles)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_opt = bk.n_options;
    if (gid == 0) ((int32_t *)out.seg_begin)[n_opt * roles.n] = (int32_t)(n_opt * roles.n);
    if (gid >= n_opt * roles.n) return;
    const int k = (int)(gid / n_opt);
    const int64_t i = gid - (int64_t)k * n_opt;
    const Role role = roles.r[k];
    const int fl = bk.flags[i];
    const bool b76 = fl & FDM_FLAG_BLACK76;
    const int n_space = max(bk.n_space[i], 3);                          // :142
    const int n_time0 = max(bk.n_time[i], 4);                           // :143
    int n_time = n_time0;
    if (role.fine == 1) n_time = b76 ? 2 * n_time0 : 2 * n_space;
    if (role.fine == 2) n_time = 2 * n_time0;
    const double spot = bk.spot[i], strike = bk.strike[i], T = bk.T[i];
    const double sigma = __dadd_rn(bk.sigma[i], __dmul_rn(role.bump, bk.dv_sigma));
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
    ((int32_t *)out.flags)[gid] = (fl & (FDM_FLAG_CALL | FDM_FLAG_BLACK76)) | FDM_FLAG_AMERICAN | FDM_FLAG_SNAP;
    ((int32_t *)out.rannacher)[gid] = max(bk.rannacher[i], 0);          // :144
    ((int32_t *)out.seg_begin)[gid] = (int32_t)gid;
    ((int32_t *)out.seg_steps)[gid] = n_time;
    ((double *)out.seg_dt)[gid] = T / (double)n_time;                   // :412 with tau_start = 0
    ((double *)out.seg_tau0)[gid] = 0.0;
    ((int32_t *)out.seg_restart)[gid] = 1;
    ((double *)out.seg_div)[gid] = 0.0;
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
    auto P = [&](int k) { return res.price[(int64_t)k * n + i]; };
    auto D = [&](int k) { return res.delta[(int64_t)k * n + i]; };
    auto G = [&](int k) { return res.gamma[(int64_t)k * n + i]; };
    int st = 0;
    for (int k = 0; k < n_roles; ++k) st = max(st, res.status[(int64_t)k * n + i]);
    if (o.status) o.status[i] = st;
    if (o.ex_lo) o.ex_lo[i] = res.idx[4 * i + 2];                       // exercise interval of the first (base) solve
    if (o.ex_hi) o.ex_hi[i] = res.idx[4 * i + 3];
    int k0 = 0;
    if (what == FDM_BOOK_PRICE_LOG) {
        if (o.price_log2) o.price_log2[i] = P(0);
        if (o.price) o.price[i] = P(0);
        if (o.delta) o.delta[i] = D(0);
        if (o.gamma) o.gamma[i] = G(0);
        return;
    }
    if (what == FDM_BOOK_PRICE_LOG2 || what == FDM_BOOK_SCENARIO) {
        const double pN = P(0);
        double v = pN;
        if (rich) v = __dsub_rn(__dmul_rn(4.0, P(1)), pN) / 3.0;         // (4 p2N - pN) / 3, :672
        if (o.price_log2) o.price_log2[i] = v;
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
    if (o.price) o.price[i] = price;
    if (o.delta) o.delta[i] = delta;
    if (o.gamma) o.gamma[i] = gamma;
    if (o.vega) o.vega[i] = dV / 100.0;                                  // :733
    if (o.theta) o.theta[i] = theta;
}

}  // namespace book
