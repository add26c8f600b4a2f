/*
 * fdm_oracle.c -- CPU restatement of the reference's 1-D log-space theta/Crank-Nicolson
 * PDE solvers.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product package
 * (finite_difference_b200/) never imports, links or calls anything in oracle/.
 *
 * It follows the reference's arithmetic line for line: same operation order, true
 * division, no FMA contraction (build with -ffp-contract=off), libm exp/log (what CPython's
 * math.exp/math.log call).  It is pinned against outputs of the unmodified reference
 * generated in the build container (tests/golden/, made by oracle/gen_golden.py).
 *
 * Citations are to files under the reference checkout:
 *   AM  = fd_american_option_pricer.py  (fd_american_equity.py is an arithmetic clone)
 *   B76 = fd_american_black76.py
 *   BAR = discrete_barrier_fdm_pricer_cn.py lines 1-537
 *         (fd_discrete_barrier_option.py lines 1-423 hold the same solver)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define FDO_OK 0
#define FDO_EINVAL -1
#define FDO_ENOMEM -2

/* ------------------------------------------------------------------------------------ */
/* Thomas algorithm with scalar (Toeplitz) diagonals.  AM:438-457, BAR:250-269.          */
/* c_prime / d_prime are caller-provided scratch of length n.                            */
static void thomas_toeplitz(double A_L, double A_C, double A_U, const double *rhs,
                            double *x, double *c_prime, double *d_prime, int n)
{
    double denom = A_C;
    c_prime[0] = A_U / denom;
    d_prime[0] = rhs[0] / denom;
    for (int i = 1; i < n; ++i) {
        denom = A_C - A_L * c_prime[i - 1];
        if (i < n - 1)
            c_prime[i] = A_U / denom;
        d_prime[i] = (rhs[i] - A_L * d_prime[i - 1]) / denom;
    }
    x[n - 1] = d_prime[n - 1];
    for (int i = n - 2; i >= 0; --i)
        x[i] = d_prime[i] - c_prime[i] * x[i + 1];
}

/* ------------------------------------------------------------------------------------ */
/* American pricers                                                                      */

typedef struct {
    double spot;        /* AM:73  (Black-76: the forward, B76:74-75) */
    double strike;      /* AM:74 */
    double T;           /* time_to_expiry, AM:117 */
    double sigma;       /* AM:77 */
    double r;           /* discount_rate_nacc, AM:127 */
    double b;           /* carry_rate_nacc, AM:131-136 (ignored for Black-76) */
    double s_max_mult;  /* AM:145 */
    int32_t is_call;    /* AM:79 */
    int32_t model;      /* 0 = equity (AM / fd_american_equity), 1 = Black-76 forward (B76) */
    int32_t n_space;    /* num_space_nodes, AM:142 */
    int32_t n_time;     /* base_N for this solve, AM:562 */
    int32_t rannacher;  /* rannacher_steps, AM:144 */
    int32_t n_div;      /* dividends already mapped to tau by the caller (AM:304-315) */
    const double *div_tau;   /* ascending tau, strictly inside (0,T) */
    const double *div_cash;
} FdoAmerican;

typedef struct {
    int32_t spot_idx;       /* AM:253-255 */
    int32_t strike_idx;     /* AM:261-263 */
    int32_t n_steps_total;  /* sum of segment steps actually marched */
    int32_t ex_lo;          /* first / last interior index k (0-based in the interior vector,  */
    int32_t ex_hi;          /* i.e. node k+1) with phi_k > v_candidate at the LAST step; -1 if none */
    double dx;
    double S_min;
    double S_max;
    double price;           /* _interp_price, AM:613-630 */
    int32_t seg_steps[16];  /* AM:569-575 (first 16 segments) */
    int32_t n_seg;
} FdoAmericanInfo;

/* AM:209-229 / B76:209-234 */
static void am_configure_grid(const FdoAmerican *p, double *S_min_out, double *S_max_out)
{
    double s_low = p->spot < p->strike ? p->spot : p->strike;
    double s_high = p->spot > p->strike ? p->spot : p->strike;
    double prod = s_low * s_high;
    double s_c = sqrt(prod > 1e-12 ? prod : 1e-12);
    double band = p->s_max_mult * p->sigma * sqrt(p->T > 1e-12 ? p->T : 1e-12);
    double x_c = log(s_c);
    double x_min = x_c - 0.5 * band;
    double x_max = x_c + 0.5 * band;
    double S_min = exp(x_min);
    double S_max = exp(x_max);
    double half_low = 0.5 * s_low;
    double twice_high = 2.0 * s_high;
    if (half_low < S_min) S_min = half_low;
    if (twice_high > S_max) S_max = twice_high;
    *S_min_out = S_min > 1e-8 ? S_min : 1e-8;
    *S_max_out = S_max;
}

/* first-minimum argmin of |s[i]-level| over all nodes, AM:253, AM:261 */
static int argmin_abs(const double *s, int count, double level)
{
    int best = 0;
    double bestv = fabs(s[0] - level);
    for (int i = 1; i < count; ++i) {
        double v = fabs(s[i] - level);
        if (v < bestv) { bestv = v; best = i; }
    }
    return best;
}

static inline double am_payoff(double S, double k, int is_call)   /* AM:277-281 */
{
    double v = is_call ? S - k : k - S;
    return v > 0.0 ? v : 0.0;
}

/* AM:286-298 and B76:291-314 */
static void am_boundary(const FdoAmerican *p, double S_max_node, double k, double tau,
                        double *V_min, double *V_max)
{
    if (p->model == 1) {
        double disc = exp(-p->r * tau);
        if (p->is_call) { *V_min = 0.0; *V_max = disc * (S_max_node - k); }
        else            { *V_min = disc * k; *V_max = 0.0; }
    } else {
        if (p->is_call) {
            *V_min = 0.0;
            *V_max = S_max_node * exp((p->b - p->r) * tau) - k * exp(-p->r * tau);
        } else {
            *V_min = k * exp(-p->r * tau);
            *V_max = 0.0;
        }
    }
}

typedef struct {
    double *rhs, *tilde, *cp, *dp, *lam, *payoff;   /* length n+1 each */
    double *sp_alpha, *sp_l, *sp_mu, *sp_z, *sp_c, *sp_b, *sp_d, *sp_h, *Vtmp;
} AmScratch;

static int am_scratch_alloc(AmScratch *w, int n_nodes)
{
    size_t bytes = (size_t)n_nodes * sizeof(double);
    double **slots[] = {&w->rhs, &w->tilde, &w->cp, &w->dp, &w->lam, &w->payoff,
                        &w->sp_alpha, &w->sp_l, &w->sp_mu, &w->sp_z, &w->sp_c, &w->sp_b,
                        &w->sp_d, &w->sp_h, &w->Vtmp};
    memset(w, 0, sizeof(*w));
    for (size_t i = 0; i < sizeof(slots) / sizeof(slots[0]); ++i) {
        *slots[i] = (double *)malloc(bytes);
        if (!*slots[i]) return FDO_ENOMEM;
    }
    return FDO_OK;
}

static void am_scratch_free(AmScratch *w)
{
    double *slots[] = {w->rhs, w->tilde, w->cp, w->dp, w->lam, w->payoff,
                       w->sp_alpha, w->sp_l, w->sp_mu, w->sp_z, w->sp_c, w->sp_b,
                       w->sp_d, w->sp_h, w->Vtmp};
    for (size_t i = 0; i < sizeof(slots) / sizeof(slots[0]); ++i) free(slots[i]);
}

/*
 * _solve_segment, AM:392-511 (B76:320-444).  V has N+1 entries.  ex_trace, if not NULL,
 * receives (lo, hi) per step at ex_trace[2*step], ex_trace[2*step+1].
 */
static void am_solve_segment(const FdoAmerican *p, const double *s, double k_pde, double dx,
                             int N, double *V, double tau_start, double tau_end, int n_steps,
                             int restart_rannacher, AmScratch *w, int32_t *ex_trace,
                             int32_t *ex_lo_last, int32_t *ex_hi_last)
{
    if (n_steps < 1) return;                                          /* AM:404-405 */

    double dt = (tau_end - tau_start) / (double)n_steps;              /* AM:412 */
    double sig2 = p->sigma * p->sigma;
    double r = p->r;
    double mu_x;
    if (p->model == 1) {
        mu_x = -0.5 * sig2;                                           /* B76:351 */
    } else {
        double q = 0.0;
        mu_x = (p->b - q) - 0.5 * sig2;                               /* AM:419 */
    }
    double alpha = 0.5 * sig2 / (dx * dx);                            /* AM:421 */
    double beta_adv = mu_x / (2.0 * dx);                              /* AM:422 */
    double a = alpha - beta_adv;                                      /* AM:424 */
    double c = alpha + beta_adv;
    double bcoef = -2.0 * alpha - r;

    int n_int = N - 1;
    double *payoff_int = w->payoff;                                   /* AM:460-461 */
    for (int j = 0; j < n_int; ++j) payoff_int[j] = am_payoff(s[j + 1], k_pde, p->is_call);
    double *lam = w->lam;
    for (int j = 0; j < n_int; ++j) lam[j] = 0.0;                     /* AM:463 */
    int base_ran = restart_rannacher ? p->rannacher : 0;              /* AM:464 */

    double tau = tau_start;
    for (int m = 0; m < n_steps; ++m) {
        double tau_next = tau + dt;                                   /* AM:468 */
        double theta = (m < base_ran) ? 1.0 : 0.5;                    /* AM:469 */

        double A_L = -theta * dt * a;                                 /* AM:429-435 */
        double A_C = 1.0 - theta * dt * bcoef;
        double A_U = -theta * dt * c;
        double B_L = (1.0 - theta) * dt * a;
        double B_C = 1.0 + (1.0 - theta) * dt * bcoef;
        double B_U = (1.0 - theta) * dt * c;

        double V_min, V_max;
        am_boundary(p, s[N], k_pde, tau_next, &V_min, &V_max);        /* AM:472 */

        double *rhs = w->rhs;
        for (int j = 1; j < N; ++j)                                   /* AM:475-480 */
            rhs[j - 1] = B_L * V[j - 1] + B_C * V[j] + B_U * V[j + 1] + dt * lam[j - 1];
        rhs[0] -= A_L * V_min;                                        /* AM:482-483 */
        rhs[n_int - 1] -= A_U * V_max;

        thomas_toeplitz(A_L, A_C, A_U, rhs, w->tilde, w->cp, w->dp, n_int);   /* AM:485 */

        int lo = -1, hi = -1;
        for (int k = 0; k < n_int; ++k) {                             /* AM:490-503 */
            double tilde_vk = w->tilde[k];
            double phi_k = payoff_int[k];
            double lam_old = lam[k];
            double v_candidate = tilde_vk - dt * lam_old;
            double v_new;
            if (phi_k > v_candidate) {
                v_new = phi_k;
                if (lo < 0) lo = k;
                hi = k;
            } else {
                v_new = v_candidate;
            }
            double lam_new = lam_old + (phi_k - tilde_vk) / dt;
            if (lam_new < 0.0) lam_new = 0.0;
            lam[k] = lam_new;
            V[k + 1] = v_new;
        }
        V[0] = V_min;                                                 /* AM:505-506 */
        V[N] = V_max;
        tau = tau_next;                                               /* AM:509 */
        if (ex_trace) { ex_trace[2 * m] = lo; ex_trace[2 * m + 1] = hi; }
        *ex_lo_last = lo;
        *ex_hi_last = hi;
    }
}

/*
 * _build_natural_cubic_spline + _apply_dividend_jump, AM:321-386, AM:517-548.
 * numpy float64 scalar arithmetic in the reference == IEEE double here.
 */
static void am_dividend_jump(const FdoAmerican *p, const double *s, double k_pde, int n_nodes,
                             double *V, double cash_div, AmScratch *w)
{
    int n = n_nodes;
    double *h = w->sp_h, *alpha = w->sp_alpha, *l = w->sp_l, *mu = w->sp_mu, *z = w->sp_z;
    double *cc = w->sp_c, *bb = w->sp_b, *dd = w->sp_d;
    const double *x = s, *y = V;

    for (int i = 0; i < n - 1; ++i) h[i] = x[i + 1] - x[i];           /* AM:334 */
    for (int i = 0; i < n; ++i) { alpha[i] = 0.0; l[i] = 1.0; mu[i] = 0.0; z[i] = 0.0; cc[i] = 0.0; }
    for (int i = 1; i < n - 1; ++i)                                   /* AM:339-343 */
        alpha[i] = 3.0 / h[i] * (y[i + 1] - y[i]) - 3.0 / h[i - 1] * (y[i] - y[i - 1]);
    for (int i = 1; i < n - 1; ++i) {                                 /* AM:349-352 */
        l[i] = 2.0 * (x[i + 1] - x[i - 1]) - h[i - 1] * mu[i - 1];
        mu[i] = h[i] / l[i];
        z[i] = (alpha[i] - h[i - 1] * z[i - 1]) / l[i];
    }
    for (int j = n - 2; j >= 0; --j) {                                /* AM:358-364 */
        cc[j] = z[j] - mu[j] * cc[j + 1];
        bb[j] = (y[j + 1] - y[j]) / h[j] - h[j] * (cc[j + 1] + 2.0 * cc[j]) / 3.0;
        dd[j] = (cc[j + 1] - cc[j]) / (3.0 * h[j]);
    }

    double *Vn = w->Vtmp;
    for (int jn = 0; jn < n; ++jn) {                                  /* AM:533-546 */
        double Sj = s[jn];
        double Sj_minus = Sj - cash_div;
        double cont_val;
        if (Sj_minus <= s[0]) {
            cont_val = V[0];
        } else if (Sj_minus >= s[n - 1]) {
            cont_val = V[n - 1];
        } else {
            /* spline_eval, AM:368-384 (the two outer guards cannot trigger here) */
            int lo = 0, hi = n - 1;
            while (hi - lo > 1) {
                int mid = (lo + hi) / 2;
                if (Sj_minus < x[mid]) hi = mid; else lo = mid;
            }
            int j = lo;
            double dxq = Sj_minus - x[j];
            cont_val = y[j] + bb[j] * dxq + cc[j] * dxq * dxq + dd[j] * dxq * dxq * dxq;
        }
        if (p->is_call) {
            double ex_val = am_payoff(Sj, k_pde, 1);
            Vn[jn] = cont_val > ex_val ? cont_val : ex_val;           /* AM:542-544 */
        } else {
            Vn[jn] = cont_val;
        }
    }
    memcpy(V, Vn, (size_t)n * sizeof(double));
}

/* Python round(): half to even.  nearbyint under the default rounding mode. */
static inline long py_round(double v) { return (long)nearbyint(v); }

/*
 * _solve_grid, AM:554-602 (B76:450-472: single segment, dividends ignored).
 * V_out: n_space+1 doubles.  s_out (optional): n_space+1 doubles.
 * ex_trace (optional): 2*n_steps_total int32 (lo, hi per marched step, segment by segment).
 */
int fdo_american_solve(const FdoAmerican *p, double *V_out, double *s_out,
                       int32_t *ex_trace, FdoAmericanInfo *info)
{
    if (!p || !V_out || !info) return FDO_EINVAL;
    if (!(p->spot > 0.0) || !(p->strike > 0.0) || !(p->sigma > 0.0) || !(p->T > 0.0))
        return FDO_EINVAL;
    int n = p->n_space;
    if (n < 3 || p->n_time < 1) return FDO_EINVAL;
    int n_nodes = n + 1;

    AmScratch w;
    if (am_scratch_alloc(&w, n_nodes + 1) != FDO_OK) { am_scratch_free(&w); return FDO_ENOMEM; }
    double *s = (double *)malloc((size_t)n_nodes * sizeof(double));
    if (!s) { am_scratch_free(&w); return FDO_ENOMEM; }

    /* _build_log_grid, AM:231-244 */
    double S_min, S_max;
    am_configure_grid(p, &S_min, &S_max);
    double x_min = log(S_min);
    double x_max = log(S_max);
    double dx = (x_max - x_min) / (double)n;
    for (int i = 0; i < n_nodes; ++i) {
        double x = x_min + (double)i * dx;
        s[i] = exp(x);
    }
    int spot_idx = argmin_abs(s, n_nodes, p->spot);                   /* AM:253 */
    int strike_idx = argmin_abs(s, n_nodes, p->strike);               /* AM:261 */
    double k_pde = s[strike_idx];                                     /* AM:272-275 */

    double *V = V_out;
    for (int i = 0; i < n_nodes; ++i) V[i] = am_payoff(s[i], k_pde, p->is_call);   /* AM:558 */

    int n_div = (p->model == 1) ? 0 : p->n_div;                       /* B76:89 */
    int base_N = p->n_time;                                           /* AM:562 */
    double T = p->T;
    double base_dt = T / (double)base_N;                              /* AM:563 */
    int n_segments = n_div + 1;

    double *tau_points = (double *)malloc((size_t)(n_segments + 1) * sizeof(double));
    int *seg_steps = (int *)malloc((size_t)n_segments * sizeof(int));
    if (!tau_points || !seg_steps) {
        free(tau_points); free(seg_steps); free(s); am_scratch_free(&w);
        return FDO_ENOMEM;
    }
    tau_points[0] = 0.0;                                              /* AM:565 */
    for (int i = 0; i < n_div; ++i) tau_points[i + 1] = p->div_tau[i];
    tau_points[n_segments] = T;

    int remaining = base_N;                                           /* AM:568-575 */
    for (int i = 0; i < n_segments - 1; ++i) {
        double L = tau_points[i + 1] - tau_points[i];
        long nn = py_round(L / base_dt);
        if (nn < 1) nn = 1;
        seg_steps[i] = (int)nn;
        remaining -= (int)nn;
    }
    seg_steps[n_segments - 1] = remaining > 1 ? remaining : 1;

    int step_off = 0;
    int32_t ex_lo = -1, ex_hi = -1;
    for (int seg = 0; seg < n_segments; ++seg) {                      /* AM:578-600 */
        int restart = (seg == 0) || (seg > 0 && p->is_call);          /* AM:583-586 */
        if (p->model == 1) restart = 1;                               /* B76:470 */
        am_solve_segment(p, s, k_pde, dx, n, V, tau_points[seg], tau_points[seg + 1],
                         seg_steps[seg], restart, &w,
                         ex_trace ? ex_trace + 2 * step_off : NULL, &ex_lo, &ex_hi);
        step_off += seg_steps[seg];
        if (seg < n_div)                                              /* AM:598-600 */
            am_dividend_jump(p, s, k_pde, n_nodes, V, p->div_cash[seg], &w);
    }

    /* _interp_price at the snapped spot, AM:613-630 */
    double S0 = s[spot_idx];
    double price;
    if (S0 <= s[0]) price = V[0];
    else if (S0 >= s[n]) price = V[n];
    else {
        int lo = 0, hi = n;
        while (hi - lo > 1) {
            int mid = (lo + hi) / 2;
            if (S0 < s[mid]) hi = mid; else lo = mid;
        }
        double wgt = (S0 - s[lo]) / (s[hi] - s[lo]);
        price = (1.0 - wgt) * V[lo] + wgt * V[hi];
    }

    memset(info, 0, sizeof(*info));
    info->spot_idx = spot_idx;
    info->strike_idx = strike_idx;
    info->n_steps_total = step_off;
    info->ex_lo = ex_lo;
    info->ex_hi = ex_hi;
    info->dx = dx;
    info->S_min = S_min;
    info->S_max = S_max;
    info->price = price;
    info->n_seg = n_segments;
    for (int i = 0; i < n_segments && i < 16; ++i) info->seg_steps[i] = seg_steps[i];
    if (s_out) memcpy(s_out, s, (size_t)n_nodes * sizeof(double));

    free(tau_points); free(seg_steps); free(s); am_scratch_free(&w);
    return FDO_OK;
}

/* ------------------------------------------------------------------------------------ */
/* Discrete barrier Crank-Nicolson dataclass                                             */

typedef struct {
    double S0, K, T, sigma, r_disc, b_carry;     /* BAR:28-33 */
    double lower_barrier;                        /* NaN == None, BAR:39 */
    double upper_barrier;                        /* NaN == None, BAR:40 */
    double rebate;                               /* BAR:41 */
    int32_t is_call;                             /* BAR:34 */
    int32_t ko_kind;       /* 0 none, 1 down-and-out, 2 up-and-out (BAR:203-213), applied only if apply_KO */
    int32_t apply_KO;
    int32_t N_space;                             /* must be set (see fdo_barrier_configure) */
    int32_t N_time;
    int32_t n_monitor;
    const double *monitor_times;                 /* calendar years, BAR:44 */
} FdoBarrier;

typedef struct {
    double S_min, S_max, dx;
    double price;            /* BAR:308-327 */
    double delta, gamma;     /* BAR:329-353 */
    int32_t n_monitor_steps; /* size of the monitor index set */
    int32_t greek_idx;       /* interior node used for delta/gamma */
} FdoBarrierInfo;

/* configure_grid, BAR:59-118.  N_space_in / N_time_in <= 0 mean None. */
int fdo_barrier_configure(const FdoBarrier *p, int32_t N_space_in, int32_t N_time_in,
                          double *S_min_out, double *S_max_out,
                          int32_t *N_space_out, int32_t *N_time_out)
{
    if (!(p->T > 0.0) || !(p->sigma > 0.0) || !(p->S0 > 0.0)) return FDO_EINVAL;   /* BAR:61-66 */
    double s_low = p->S0 < p->K ? p->S0 : p->K;                        /* BAR:69-76 */
    double s_high = p->S0 > p->K ? p->S0 : p->K;
    if (!isnan(p->lower_barrier) && p->lower_barrier > 0) {
        if (p->lower_barrier < s_low) s_low = p->lower_barrier;
        if (p->lower_barrier > s_high) s_high = p->lower_barrier;
    }
    if (!isnan(p->upper_barrier) && p->upper_barrier > 0) {
        if (p->upper_barrier < s_low) s_low = p->upper_barrier;
        if (p->upper_barrier > s_high) s_high = p->upper_barrier;
    }
    double S_min = s_low / 4.0;                                        /* BAR:79-85 */
    if (S_min < 1e-8) S_min = 1e-8;
    double S_max = s_high * 4.0;
    if (S_min >= S_max) { S_min = p->S0 / 5.0; S_max = p->S0 * 5.0; }
    *S_min_out = S_min;
    *S_max_out = S_max;

    double x_min = log(S_min), x_max = log(S_max);                     /* BAR:90-95 */
    double x_range = x_max - x_min;
    double dx_target = p->sigma * sqrt(p->T) / 12;
    if (dx_target <= 0.0) dx_target = x_range / 300.0;

    int32_t N_space = N_space_in;
    if (N_space <= 0) {                                                /* BAR:97-99 */
        N_space = (int32_t)ceil(x_range / dx_target);
        if (N_space < 300) N_space = 300;
    }
    int32_t N_time = N_time_in;
    if (N_time <= 0) {                                                 /* BAR:102-118 */
        double dx = x_range / N_space;
        double lambda_target = 0.4;
        int32_t Ntime_opt = (int32_t)ceil(0.5 * p->sigma * p->sigma * p->T / (lambda_target * dx * dx));
        int valid = 0;
        for (int i = 0; i < p->n_monitor; ++i)
            if (0.0 < p->monitor_times[i] && p->monitor_times[i] < p->T) ++valid;
        int32_t M_intervals = valid + 1;
        N_time = Ntime_opt;
        if (N_space > N_time) N_time = N_space;
        if (10 * M_intervals > N_time) N_time = 10 * M_intervals;
    }
    *N_space_out = N_space;
    *N_time_out = N_time;
    return FDO_OK;
}

/* _solve_grid, BAR:219-302, then price/delta/gamma, BAR:308-353.  V_out: N_space+1. */
/* Engine options beyond the importable dataclass, after the non-parsing fragments of
 * fd_discrete_barrier_option.py: `rannacher` fully implicit start-up steps (theta = 1 for the first steps, :912-1007),
 * `rebate_discounted` (value at a knocked-out node = rebate * exp(-r_disc * tau), the rebate being paid at expiry,
 * :851-867) and ko_kind 3 = double-out (:887-893).  No runnable reference exists for them: PARITY UNPINNED; this
 * restatement keeps the operation order of the pinned path and is what the CUDA kernel is checked against.
 * With (0, 0) and ko_kind <= 2 the arithmetic is exactly the pinned fdo_barrier_solve. */
static int barrier_solve_impl(const FdoBarrier *p, int rannacher, int rebate_discounted, double *V_out, double *s_out,
                              int32_t *monitor_steps_out, FdoBarrierInfo *info);

int fdo_barrier_solve(const FdoBarrier *p, double *V_out, double *s_out,
                      int32_t *monitor_steps_out /* optional, N_time flags */,
                      FdoBarrierInfo *info)
{
    return barrier_solve_impl(p, 0, 0, V_out, s_out, monitor_steps_out, info);
}

int fdo_barrier_solve_ext(const FdoBarrier *p, int32_t rannacher, int32_t rebate_discounted, double *V_out,
                          double *s_out, int32_t *monitor_steps_out, FdoBarrierInfo *info)
{
    return barrier_solve_impl(p, rannacher, rebate_discounted, V_out, s_out, monitor_steps_out, info);
}

static int barrier_solve_impl(const FdoBarrier *p, int rannacher, int rebate_discounted, double *V_out, double *s_out,
                              int32_t *monitor_steps_out, FdoBarrierInfo *info)
{
    if (!p || !V_out || !info) return FDO_EINVAL;
    int N = p->N_space;
    int M = p->N_time;
    if (N < 3 || M < 1) return FDO_EINVAL;

    double S_min, S_max;
    int32_t ns, nt;
    int rc = fdo_barrier_configure(p, N, M, &S_min, &S_max, &ns, &nt);   /* BAR:224 */
    if (rc != FDO_OK) return rc;

    int n_nodes = N + 1;
    double *s = (double *)malloc((size_t)n_nodes * sizeof(double));
    double *rhs = (double *)malloc((size_t)n_nodes * sizeof(double));
    double *xi = (double *)malloc((size_t)n_nodes * sizeof(double));
    double *cp = (double *)malloc((size_t)n_nodes * sizeof(double));
    double *dp = (double *)malloc((size_t)n_nodes * sizeof(double));
    unsigned char *is_mon = (unsigned char *)calloc((size_t)M + 1, 1);
    if (!s || !rhs || !xi || !cp || !dp || !is_mon) {
        free(s); free(rhs); free(xi); free(cp); free(dp); free(is_mon);
        return FDO_ENOMEM;
    }

    double x_min = log(S_min), x_max = log(S_max);                     /* BAR:129-134 */
    double dx = (x_max - x_min) / N;
    for (int i = 0; i < n_nodes; ++i) {
        double x = x_min + i * dx;
        s[i] = exp(x);
    }

    double dt = p->T / M;                                              /* BAR:228 */
    double sig2 = p->sigma * p->sigma;
    double mu_x = p->b_carry - 0.5 * sig2;
    double alpha = 0.5 * sig2 / (dx * dx);
    double beta_adv = mu_x / (2.0 * dx);
    double a = alpha - beta_adv;
    double c = alpha + beta_adv;
    double bcoef = -2.0 * alpha - p->r_disc;

    double A_L = -0.5 * dt * a;                                        /* BAR:242-248 */
    double A_C = 1.0 - 0.5 * dt * bcoef;
    double A_U = -0.5 * dt * c;
    double B_L = 0.5 * dt * a;
    double B_C = 1.0 + 0.5 * dt * bcoef;
    double B_U = 0.5 * dt * c;

    double *V = V_out;
    for (int i = 0; i < n_nodes; ++i) {                                /* BAR:142-151 */
        double v = p->is_call ? s[i] - p->K : p->K - s[i];
        V[i] = v > 0.0 ? v : 0.0;
    }

    int n_mon_steps = 0;                                               /* BAR:177-193 */
    for (int i = 0; i < p->n_monitor; ++i) {
        double t_mon = p->monitor_times[i];
        if (t_mon <= 0.0 || t_mon >= p->T) continue;
        double tau_mon = p->T - t_mon;
        long k = py_round(tau_mon / dt);
        if (0 < k && k < M) {
            if (!is_mon[k]) ++n_mon_steps;
            is_mon[k] = 1;
        }
    }

    int n_int = N - 1;
    for (int m = 0; m < M; ++m) {                                      /* BAR:278-300 */
        if (rannacher > 0 && (m == 0 || m == rannacher)) {             /* extension: theta = 1 start-up steps */
            const double theta = m < rannacher ? 1.0 : 0.5;
            A_L = -theta * dt * a;
            A_C = 1.0 - theta * dt * bcoef;
            A_U = -theta * dt * c;
            B_L = (1.0 - theta) * dt * a;
            B_C = 1.0 + (1.0 - theta) * dt * bcoef;
            B_U = (1.0 - theta) * dt * c;
        }
        double tau_next = (m + 1) * dt;
        double V_min_next, V_max_next;
        if (p->is_call) {                                              /* BAR:153-171 */
            V_min_next = 0.0;
            V_max_next = s[N] * exp((p->b_carry - p->r_disc) * tau_next)
                         - p->K * exp(-p->r_disc * tau_next);
        } else {
            V_max_next = 0.0;
            V_min_next = p->K * exp(-p->r_disc * tau_next);
        }
        for (int j = 1; j < N; ++j)
            rhs[j - 1] = B_L * V[j - 1] + B_C * V[j] + B_U * V[j + 1];
        rhs[0] -= A_L * V_min_next;
        rhs[n_int - 1] -= A_U * V_max_next;

        thomas_toeplitz(A_L, A_C, A_U, rhs, xi, cp, dp, n_int);

        V[0] = V_min_next;
        V[N] = V_max_next;
        memcpy(V + 1, xi, (size_t)n_int * sizeof(double));

        if (p->apply_KO && is_mon[m + 1]) {                            /* BAR:299-300, 199-213 */
            const double reb = rebate_discounted ? p->rebate * exp(-p->r_disc * tau_next) : p->rebate;
            if (p->ko_kind == 1 && !isnan(p->lower_barrier)) {
                for (int i = 0; i < n_nodes; ++i)
                    if (s[i] <= p->lower_barrier) V[i] = reb;
            } else if (p->ko_kind == 2 && !isnan(p->upper_barrier)) {
                for (int i = 0; i < n_nodes; ++i)
                    if (s[i] >= p->upper_barrier) V[i] = reb;
            } else if (p->ko_kind == 3 && !isnan(p->lower_barrier) && !isnan(p->upper_barrier)) {
                for (int i = 0; i < n_nodes; ++i)
                    if (s[i] <= p->lower_barrier || s[i] >= p->upper_barrier) V[i] = reb;
            }
        }
    }

    /* _interp_price_from_grid, BAR:308-327 */
    double S0 = p->S0, price;
    if (S0 <= s[0]) price = V[0];
    else if (S0 >= s[N]) price = V[N];
    else {
        int lo = 0, hi = N;
        while (hi - lo > 1) {
            int mid = (lo + hi) / 2;
            if (S0 < s[mid]) hi = mid; else lo = mid;
        }
        double wgt = (S0 - s[lo]) / (s[hi] - s[lo]);
        price = (1.0 - wgt) * V[lo] + wgt * V[hi];
    }

    /* _delta_gamma_from_grid, BAR:329-353: interior node closest to S0, first minimum */
    int idx = 1;
    {
        double bestv = fabs(s[1] - S0);
        for (int i = 2; i < N; ++i) {
            double v = fabs(s[i] - S0);
            if (v < bestv) { bestv = v; idx = i; }
        }
    }
    double h1 = s[idx] - s[idx - 1];
    double h2 = s[idx + 1] - s[idx];
    double delta = -h2 / (h1 * (h1 + h2)) * V[idx - 1]
                   + (h2 - h1) / (h1 * h2) * V[idx]
                   + h1 / (h2 * (h1 + h2)) * V[idx + 1];
    double gamma = 2.0 * (V[idx - 1] / (h1 * (h1 + h2))
                          - V[idx] / (h1 * h2)
                          + V[idx + 1] / (h2 * (h1 + h2)));

    info->S_min = S_min;
    info->S_max = S_max;
    info->dx = dx;
    info->price = price;
    info->delta = delta;
    info->gamma = gamma;
    info->n_monitor_steps = n_mon_steps;
    info->greek_idx = idx;
    if (s_out) memcpy(s_out, s, (size_t)n_nodes * sizeof(double));
    if (monitor_steps_out)
        for (int k = 0; k <= M; ++k) monitor_steps_out[k] = is_mon[k];

    free(s); free(rhs); free(xi); free(cp); free(dp); free(is_mon);
    return FDO_OK;
}


/* ------------------------------------------------------------------------------------ */
/* Date/curve-aware barrier pricer DiscreteBarrierFDMPricer                              */
/* DBF = discrete_barrier_fdm_pricer.py.  The file does not parse as shipped; the class  */
/* body used for the goldens is its lines 1-745 + 883-1027 (see oracle/ref_loader.py).    */

typedef struct {
    double S_min, S_max;        /* choose_grid_parameters, DBF:270-320 (host: needs scipy norm.ppf) */
    double S0_eff;              /* spot - pv_divs, DBF:632 */
    double S_greek;             /* self.spot: centre of the delta/gamma stencil, DBF:952 */
    double K, T, sigma;
    double r, b, q;             /* discount_rate_nacc, carry_rate_nacc, div_yield_nacc, DBF:457-459 */
    double lower_barrier, upper_barrier;   /* NaN == None */
    double rebate_amount;
    int32_t rebate_at_hit;      /* DBF:421-424 */
    int32_t is_call;
    int32_t ko_kind;            /* 0 no projection (none / knock-in types, DBF:415-416), 1 down-and-out,
                                   2 up-and-out, 3 double-out */
    int32_t apply_KO;
    int32_t num_space_nodes;    /* n: s_nodes has n+1 entries, DBF:352-357 */
    int32_t num_time_steps;     /* dt = T / num_time_steps, DBF:451 */
    int32_t N_time;             /* steps marched, DBF:449 */
    int32_t rannacher_steps;    /* DBF:510-518 */
    int32_t n_monitor;
    const double *monitor_times;   /* self.monitor_times, DBF:257-268 */
} FdoDbfp;

typedef struct {
    double dx, price, delta, gamma;
    int32_t v_len;              /* length of the value vector after the march (n, see below) */
    int32_t greek_idx;
    int32_t n_monitor_steps;
} FdoDbfpInfo;

/* _solve_grid, DBF:442-547; _interp_price, DBF:629-646; _delta_gamma_from_grid, DBF:949-977.
 * Note DBF:448 N = num_space_nodes - 1 and DBF:540 V[1:-1] = V_int: the slice assignment drops one
 * element in the first step, so from then on V has n entries (nodes 0..n-1) while s_nodes keeps n+1;
 * the top Dirichlet value (evaluated with S_max = s_nodes[n]) sits at index n-1.
 * V_out: n+1 doubles (v_len are valid).  s_out: n+1. */
int fdo_dbfp_solve(const FdoDbfp *p, double *V_out, double *s_out, int32_t *monitor_steps_out, FdoDbfpInfo *info)
{
    if (!p || !V_out || !info) return FDO_EINVAL;
    const int n = p->num_space_nodes;
    const int N = n - 1;                                               /* DBF:448 */
    const int M = p->N_time;
    if (n < 4 || M < 1 || p->num_time_steps < 1) return FDO_EINVAL;
    const int n_nodes = n + 1;
    double *s = (double *)malloc((size_t)n_nodes * sizeof(double));
    double *rhs = (double *)malloc((size_t)n_nodes * sizeof(double));
    double *xi = (double *)malloc((size_t)n_nodes * sizeof(double));
    double *cp = (double *)malloc((size_t)n_nodes * sizeof(double));
    double *dp = (double *)malloc((size_t)n_nodes * sizeof(double));
    unsigned char *is_mon = (unsigned char *)calloc((size_t)p->num_time_steps + 2, 1);
    if (!s || !rhs || !xi || !cp || !dp || !is_mon) {
        free(s); free(rhs); free(xi); free(cp); free(dp); free(is_mon);
        return FDO_ENOMEM;
    }
    const double x_min = log(p->S_min), x_max = log(p->S_max);         /* DBF:349-357 */
    const double dx = (x_max - x_min) / n;
    for (int i = 0; i < n_nodes; ++i) {
        double x = x_min + i * dx;
        s[i] = exp(x);
    }
    const double dt = p->T / p->num_time_steps;                        /* DBF:451 */
    const double sig2 = p->sigma * p->sigma;
    const double mu_x = (p->b - p->q) - 0.5 * sig2;                    /* DBF:461-462 */
    const double alpha = 0.5 * sig2 / (dx * dx);
    const double beta_adv = mu_x / (2.0 * dx);
    const double a = alpha - beta_adv;
    const double c = alpha + beta_adv;
    const double bcoef = -2.0 * alpha - p->r;

    double *V = V_out;
    int v_len = n_nodes;
    for (int i = 0; i < n_nodes; ++i) {                                /* DBF:366-370 */
        double v = p->is_call ? s[i] - p->K : p->K - s[i];
        V[i] = v > 0.0 ? v : 0.0;
    }
    int n_mon_steps = 0;                                               /* DBF:395-411 */
    for (int i = 0; i < p->n_monitor; ++i) {
        double t_mon = p->monitor_times[i];
        if (t_mon <= 0.0 || t_mon > p->T) continue;
        double tau_mon = p->T - t_mon;
        long k = (long)floor(tau_mon / dt + 1e-9);
        if (k < 1) k = 1;
        if (k > p->num_time_steps) k = p->num_time_steps;
        if (!is_mon[k]) ++n_mon_steps;
        is_mon[k] = 1;
    }
    int rannacher_left = p->rannacher_steps;
    const int n_int = N - 1;
    for (int m = 0; m < M; ++m) {                                      /* DBF:513-546 */
        double theta;
        if (rannacher_left > 0) { theta = 1.0; --rannacher_left; } else theta = 0.5;
        const double A_L = -theta * dt * a;                            /* DBF:475-483 */
        const double A_C = 1.0 - theta * dt * bcoef;
        const double A_U = -theta * dt * c;
        const double B_L = (1.0 - theta) * dt * a;
        const double B_C = 1.0 + (1.0 - theta) * dt * bcoef;
        const double B_U = (1.0 - theta) * dt * c;
        const double tau_next = (m + 1) * dt;
        double V_min_next, V_max_next;                                 /* DBF:372-393 */
        if (p->is_call) {
            V_min_next = 0.0;
            V_max_next = s[n] * exp((p->b - p->r) * tau_next) - p->K * exp(-p->r * tau_next);
        } else {
            V_max_next = 0.0;
            V_min_next = p->K * exp(-p->r * tau_next) * s[0] * exp((p->b - p->r) * tau_next);
        }
        for (int j = 1; j < N; ++j)
            rhs[j - 1] = B_L * V[j - 1] + B_C * V[j] + B_U * V[j + 1];
        rhs[0] -= A_L * V_min_next;
        rhs[n_int - 1] -= A_U * V_max_next;
        thomas_toeplitz(A_L, A_C, A_U, rhs, xi, cp, dp, n_int);
        V[0] = V_min_next;                                             /* DBF:538-540 */
        memcpy(V + 1, xi, (size_t)n_int * sizeof(double));
        V[n_int + 1] = V_max_next;
        v_len = n_int + 2;                                             /* == n */
        if (p->apply_KO && is_mon[m + 1] && p->ko_kind != 0) {         /* DBF:542-543, 413-440 */
            const double rebate = p->rebate_at_hit ? p->rebate_amount
                                                   : p->rebate_amount * exp(-p->b * tau_next);
            const int cnt = v_len < n_nodes ? v_len : n_nodes;
            for (int i = 0; i < cnt; ++i) {
                int out = 0;
                if (p->ko_kind == 1 && !isnan(p->lower_barrier) && s[i] <= p->lower_barrier) out = 1;
                else if (p->ko_kind == 2 && !isnan(p->upper_barrier) && s[i] >= p->upper_barrier) out = 1;
                else if (p->ko_kind == 3 && ((!isnan(p->lower_barrier) && s[i] <= p->lower_barrier) ||
                                             (!isnan(p->upper_barrier) && s[i] >= p->upper_barrier))) out = 1;
                if (out) V[i] = rebate;
            }
        }
    }
    /* _interp_price, DBF:629-646 (binary search over the n+1 s_nodes; V indexed alike) */
    const double S0 = p->S0_eff;
    double price;
    if (S0 <= s[0]) price = V[0];
    else if (S0 >= s[n]) price = V[v_len - 1];
    else {
        int lo = 0, hi = n;
        while (hi - lo > 1) {
            int mid = (lo + hi) / 2;
            if (S0 < s[mid]) hi = mid; else lo = mid;
        }
        if (hi >= v_len) { free(s); free(rhs); free(xi); free(cp); free(dp); free(is_mon); return FDO_EINVAL; }
        double w = (S0 - s[lo]) / (s[hi] - s[lo]);
        price = (1.0 - w) * V[lo] + w * V[hi];
    }
    /* _delta_gamma_from_grid, DBF:949-977: interior node of s_nodes (1..n-1) closest to self.spot, first minimum */
    int idx = 1;
    {
        double bestv = fabs(s[1] - p->S_greek);
        for (int i = 2; i < n; ++i) {
            double v = fabs(s[i] - p->S_greek);
            if (v < bestv) { bestv = v; idx = i; }
        }
    }
    if (idx + 1 >= v_len) { free(s); free(rhs); free(xi); free(cp); free(dp); free(is_mon); return FDO_EINVAL; }
    {
        const double h1 = s[idx] - s[idx - 1];
        const double h2 = s[idx + 1] - s[idx];
        info->delta = -h2 / (h1 * (h1 + h2)) * V[idx - 1] + (h2 - h1) / (h1 * h2) * V[idx]
                      + h1 / (h2 * (h1 + h2)) * V[idx + 1];
        info->gamma = 2.0 * (V[idx - 1] / (h1 * (h1 + h2)) - V[idx] / (h1 * h2) + V[idx + 1] / (h2 * (h1 + h2)));
        info->greek_idx = idx;
    }
    info->dx = dx;
    info->price = price;
    info->v_len = v_len;
    info->n_monitor_steps = n_mon_steps;
    if (s_out) memcpy(s_out, s, (size_t)n_nodes * sizeof(double));
    if (monitor_steps_out)
        for (int k = 0; k <= p->num_time_steps; ++k) monitor_steps_out[k] = is_mon[k];
    free(s); free(rhs); free(xi); free(cp); free(dp); free(is_mon);
    return FDO_OK;
}

/* ------------------------------------------------------------------------------------ */
/* Batch drivers (OpenMP over solves) for the CPU baseline.  Results: price per solve.    */

int fdo_american_batch(int64_t n_solves, const FdoAmerican *params, double *price_out,
                       int32_t *ex_lo_out, int32_t *ex_hi_out, int32_t n_threads)
{
    int failures = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : failures)
    for (int64_t i = 0; i < n_solves; ++i) {
        double *V = (double *)malloc((size_t)(params[i].n_space + 1) * sizeof(double));
        FdoAmericanInfo info;
        int rc = V ? fdo_american_solve(&params[i], V, NULL, NULL, &info) : FDO_ENOMEM;
        if (rc == FDO_OK) {
            price_out[i] = info.price;
            if (ex_lo_out) ex_lo_out[i] = info.ex_lo;
            if (ex_hi_out) ex_hi_out[i] = info.ex_hi;
        } else {
            price_out[i] = NAN;
            ++failures;
        }
        free(V);
    }
    return failures ? FDO_EINVAL : FDO_OK;
}

int fdo_barrier_batch(int64_t n_solves, const FdoBarrier *params, double *price_out,
                      double *delta_out, double *gamma_out, int32_t n_threads)
{
    int failures = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : failures)
    for (int64_t i = 0; i < n_solves; ++i) {
        double *V = (double *)malloc((size_t)(params[i].N_space + 1) * sizeof(double));
        FdoBarrierInfo info;
        int rc = V ? fdo_barrier_solve(&params[i], V, NULL, NULL, &info) : FDO_ENOMEM;
        if (rc == FDO_OK) {
            price_out[i] = info.price;
            if (delta_out) delta_out[i] = info.delta;
            if (gamma_out) gamma_out[i] = info.gamma;
        } else {
            price_out[i] = NAN;
            ++failures;
        }
        free(V);
    }
    return failures ? FDO_EINVAL : FDO_OK;
}

/* ------------------------------------------------------------------------------------ */
/* DiscreteBarrierFDMPricer2 (P2 = discrete_barrier_fdm_pricer_2.py): S-space theta scheme with a ROW-VARYING        */
/* tridiagonal matrix, a non-symmetric stencil on the two rows next to the barrier and knock-out projection.         */
/* Restates _solve_pde_backward (P2:336-428), _solve_tridiagonal (P2:262-288) and _apply_knockout_projection         */
/* (P2:241-257) operation for operation.  The explicit half of the standard rows carries the reference's sign        */
/* (P2:373-375: a_expl, c_expl negative), which makes theta = 1/2 steps unstable; the march is well behaved when     */
/* every step is implicit (rannacher_steps >= num_time_steps).  pow(x, 2.0) is what CPython's `x ** 2` calls.        */

typedef struct {
    double sigma, r, strike, tenor, dt;
    double lo_bar, up_bar;          /* effective knock-out levels (NaN = None) */
    double h_minus, h_plus;         /* non-symmetric stencil of rows bar_j, bar_j + 1 */
    int32_t n_nodes;                /* N + 1 */
    int32_t n_steps, rannacher, is_call;
    int32_t ko_kind;                /* 0 none, 1 down-and-out, 2 up-and-out, 3 double-out (P2:246-253) */
    int32_t bar_j;                  /* left index of the interval holding the active barrier, -1: none (P2:297-321) */
    const double *s_nodes;          /* [n_nodes] shifted grid (P2:468) */
    const double *v_terminal;       /* [n_nodes] terminal payoff, smoothed (P2:218-235) */
    const uint8_t *project_after;   /* [n_steps + 1]: project_after[k] != 0 <=> step index k is in the monitoring map (P2:423-425) */
} FdoPricer2;

/* CPython evaluates `x ** 2` with libm's pow(); gcc would fold pow(x, 2.0) into x * x, which is not always the same
 * double (glibc's pow is faithfully, not correctly, rounded).  The volatile pointer keeps the library call. */
static double (*volatile p2_pow)(double, double) = pow;

static int p2_tridiag(const double *sub, const double *diag, const double *sup, const double *rhs, double *x,
                      double *c_star, double *d_star, int n)
{
    double beta = diag[0];
    if (fabs(beta) < 1e-14) return FDO_EINVAL;                         /* P2:271-272 */
    c_star[0] = sup[0] / beta;
    d_star[0] = rhs[0] / beta;
    for (int i = 1; i < n; ++i) {
        beta = diag[i] - sub[i] * c_star[i - 1];
        if (fabs(beta) < 1e-14) return FDO_EINVAL;
        c_star[i] = (i < n - 1) ? (sup[i] / beta) : 0.0;
        d_star[i] = (rhs[i] - sub[i] * d_star[i - 1]) / beta;
    }
    x[n - 1] = d_star[n - 1];
    for (int i = n - 2; i >= 0; --i) x[i] = d_star[i] - c_star[i] * x[i + 1];
    return FDO_OK;
}

int fdo_pricer2_solve(const FdoPricer2 *p, double *V_out)
{
    if (!p || !V_out || p->n_nodes < 3 || p->n_steps < 1) return FDO_EINVAL;
    const int n = p->n_nodes, N = n - 1, M = p->n_steps;
    const double *s = p->s_nodes;
    double *buf = (double *)malloc((size_t)n * 7 * sizeof(double));
    if (!buf) return FDO_ENOMEM;
    double *sub = buf, *mainv = buf + n, *sup = buf + 2 * n, *rhs = buf + 3 * n, *V = buf + 4 * n, *cs = buf + 5 * n, *ds = buf + 6 * n;
    const double dt = p->dt, r = p->r, sig = p->sigma;
    const double dS = s[1] - s[0];                                     /* P2:345 */
    memcpy(V, p->v_terminal, (size_t)n * sizeof(double));
    int rc = FDO_OK;
    if (dS == 0.0 || p2_pow(dS, 2.0) == 0.0) rc = FDO_EINVAL;         /* CPython raises ZeroDivisionError at P2:379 */
    for (int m = M; m >= 1 && rc == FDO_OK; --m) {
        const double theta = ((M - m) < p->rannacher) ? 1.0 : 0.5;     /* P2:351 */
        for (int i = 0; i < n; ++i) { sub[i] = 0.0; mainv[i] = 0.0; sup[i] = 0.0; rhs[i] = 0.0; }
        const double tau_left = p->tenor - (m - 1) * dt;               /* P2:359 */
        if (p->is_call) { rhs[0] = 0.0; rhs[N] = s[N] - p->strike * exp(-r * tau_left); }
        else            { rhs[0] = p->strike * exp(-r * tau_left); rhs[N] = 0.0; }
        mainv[0] = 1.0;
        mainv[N] = 1.0;
        for (int i = 1; i < N; ++i) {
            const double S = s[i];
            const double sig2S2 = p2_pow(sig * S, 2.0);
            double a_impl, b_impl, c_impl, a_expl, b_expl, c_expl;
            if (p->bar_j < 0 || (i != p->bar_j && i != p->bar_j + 1)) {
                const double dS2 = p2_pow(dS, 2.0);                    /* P2:373-380 */
                a_impl = 0.5 * dt * theta * (sig2S2 / dS2 - r * S / dS);
                b_impl = 1.0 + dt * theta * (sig2S2 / dS2 + r);
                c_impl = 0.5 * dt * theta * (sig2S2 / dS2 + r * S / dS);
                a_expl = -0.5 * dt * (1 - theta) * (sig2S2 / dS2 - r * S / dS);
                b_expl = 1.0 - dt * (1 - theta) * (sig2S2 / dS2 + r);
                c_expl = -0.5 * dt * (1 - theta) * (sig2S2 / dS2 + r * S / dS);
            } else {
                const double hm = p->h_minus, hp = p->h_plus;          /* P2:383-409 */
                const double a1 = hp / (hm * (hm + hp));
                const double b1 = (hp - hm) / (hm * hp);
                const double c1 = -hm / (hp * (hm + hp));
                const double d2 = 2.0 / (hm * (hm + hp));
                const double e2 = -2.0 / (hm * hp);
                const double f2 = 2.0 / (hp * (hm + hp));
                const double L_left = 0.5 * sig2S2 * f2 + r * S * c1;
                const double L_center = 0.5 * sig2S2 * e2 + r * S * b1 - r;
                const double L_right = 0.5 * sig2S2 * d2 + r * S * a1;
                a_impl = -theta * dt * L_left;
                b_impl = 1.0 - theta * dt * L_center;
                c_impl = -theta * dt * L_right;
                a_expl = (1 - theta) * dt * L_left;
                b_expl = 1.0 + (1 - theta) * dt * L_center;
                c_expl = (1 - theta) * dt * L_right;
            }
            sub[i] = -a_impl;
            mainv[i] = b_impl;
            sup[i] = -c_impl;
            rhs[i] = a_expl * V[i - 1] + b_expl * V[i] + c_expl * V[i + 1];
        }
        rc = p2_tridiag(sub, mainv, sup, rhs, V, cs, ds, n);
        if (rc == FDO_OK && p->project_after && p->project_after[m - 1]) {   /* P2:423-425, P2:241-257 */
            for (int i = 0; i < n; ++i) {
                int ko = 0;
                if (p->ko_kind == 1 && !isnan(p->lo_bar) && s[i] <= p->lo_bar) ko = 1;
                else if (p->ko_kind == 2 && !isnan(p->up_bar) && s[i] >= p->up_bar) ko = 1;
                else if (p->ko_kind == 3 && ((!isnan(p->lo_bar) && s[i] <= p->lo_bar) || (!isnan(p->up_bar) && s[i] >= p->up_bar))) ko = 1;
                if (ko) V[i] = 0.0;
            }
        }
    }
    memcpy(V_out, V, (size_t)n * sizeof(double));
    free(buf);
    return rc;
}

int fdo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
