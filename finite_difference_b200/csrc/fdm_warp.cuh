// fdm_warp.cuh -- fdm_warp_kernel: one warp owns one PDE solve, per-row LU coefficients in registers.
//
// One warp owns one PDE solve.  The n-1 interior unknowns are split into 32 contiguous chunks
// of C nodes; each lane keeps its chunk of V, dt*lambda, payoff and the LU coefficients in
// registers for the whole time march.  Nothing but the per-solve scalars is read from HBM and
// nothing but prices / stencils / exercise intervals is written back.
//
// What one time step computes (reference: fd_american_option_pricer.py:467-509,
// discrete_barrier_fdm_pricer_cn.py:278-300), restated so that the work per node is 9 FP64
// instructions instead of the reference's 22 flops:
//
//   reference                                    here (p = 1/theta, q = (1-theta)/theta, L = dt*lambda)
//   rhs = B V + dt*lam  (+ boundary terms)       B = p I - q A  =>  A^-1 (B V + L) = A^-1 (p V + L) - q V
//   Vt  = A^-1 rhs      (Thomas, 2 divisions)    exact LU of the Toeplitz A hoisted out of the loop:
//                                                   u_i  = R_i - m_i u_{i-1}            m_i  = A_L / den_{i-1}
//                                                   xs_i = u_i - cc_i xs_{i+1}          cc_i = A_U / den_{i+1}
//                                                   Vt_i = inv_i xs_i                   inv_i = 1 / den_i
//                                                with den_i the reference's pivots (same recurrence, :443-451)
//   vc  = Vt - dt*lam ; V = max(phi, vc)         vc = inv*xs - (q V + L) ; L = max(phi - vc, 0) ; V = vc + L
//   lam = max(0, lam + (phi - Vt)/dt)            (identical in exact arithmetic; V = phi up to one rounding when exercised)
//
// The two first-order recurrences run across the warp as (local pass, 5-level shuffle scan of
// affine maps, local pass with the carry).  Dirichlet boundary terms enter as the carries of
// lane 0 / lane 31, so no per-node masking is needed.
//
// No CPU fallback exists: every entry point fails with FDM_ERR_NO_DEVICE / FDM_ERR_CUDA if the
// GPU path cannot run.

#pragma once

#include "fdm_common.cuh"
#include "fdm_lean_classes.cuh"

namespace {

struct PhaseScalars {
    // Boundary term of this lane: lane 0 owns row 0 (V_min side), lane 31 the top row (V_max side).
    // carry = cB * V_bdry(now) - cA * V_bdry(next step): rhs -= A * V_next (:482-483) plus the
    // boundary column of the explicit operator B (:476-480); zero on every other lane.
    double cA, cB;
    double p, q;                 // 1/theta, (1-theta)/theta
    double pd;                   // 1 + delta: coefficient of V in R = G + pd V (see setup_phase)
    int terms;                   // carries: 2/4/6-term direct weighted sum; 0 = Kogge-Stone scan
};

// Shared-memory tables of one warp (doubles):
//   tab  [32*C+2]            pivot reciprocals during setup; value vector for read-out / dividends
//   Af,Ab[SCAN_LEVELS][32]   Kogge-Stone multipliers, forward / backward sweep
//   Wf,Wb[DIRECT_TERMS][32]  weights of the direct carry sum
//   Hf,Hb[2][32]             half-chunk multipliers
template <int C>
struct WarpTables {
    double *tab, *Af, *Ab, *Wf, *Wb, *Hf, *Hb;
    __device__ explicit WarpTables(double *base)
    {
        tab = base;
        Af = tab + (32 * C + 2);
        Ab = Af + SCAN_LEVELS * 32;
        Wf = Ab + SCAN_LEVELS * 32;
        Wb = Wf + DIRECT_TERMS * 32;
        Hf = Wb + DIRECT_TERMS * 32;
        Hb = Hf + 2 * 32;
    }
    __device__ double *end() const { return Hb + 2 * 32; }
};

template <int C>
constexpr int warp_smem_doubles() { return (32 * C + 2) + 2 * SCAN_LEVELS * 32 + 2 * DIRECT_TERMS * 32 + 4 * 32; }

// Builds the per-lane LU coefficient registers for one (theta, dt) phase and the carry tables
// (shared memory).  The pivots follow the reference's recurrence den_0 = A_C,
// den_i = A_C - A_L * A_U / den_{i-1} (fd_american_option_pricer.py:443-451) with one
// reciprocal per row; they converge bit-wise to a fixed point after a few dozen rows, after
// which the table is constant.
template <int C, bool AMER>
__device__ __forceinline__ void setup_phase(double theta, double dt, double a, double c, double bcoef,
                                            int n_int, int lane, double (&nm)[C], double (&ncc)[C],
                                            double (&inv)[C], const WarpTables<C> &T, PhaseScalars &ph)
{
    constexpr int H = (C + 1) / 2;
    const double A_L = __dmul_rn(__dmul_rn(-theta, dt), a);
    const double A_C = __dsub_rn(1.0, __dmul_rn(__dmul_rn(theta, dt), bcoef));
    const double A_U = __dmul_rn(__dmul_rn(-theta, dt), c);
    const double B_L = __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), a);
    const double B_U = __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), c);
    ph.p = 1.0 / theta;
    ph.q = (1.0 - theta) / theta;
    const double B_C = __dadd_rn(1.0, __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), bcoef));

    double *tab = T.tab;
    __syncwarp();
    double den = A_C;
    double rinv = __drcp_rn(den);
    if (lane == 0) tab[0] = rinv;
    int conv = n_int;                       // rows >= conv-1 share the pivot of row conv-1
    for (int i = 1; i < n_int; ++i) {
        const double den_i = __dsub_rn(A_C, __dmul_rn(A_L, __dmul_rn(A_U, rinv)));
        if (den_i == den) { conv = i; break; }
        den = den_i;
        rinv = __drcp_rn(den);
        if (lane == 0) tab[i] = rinv;
    }
    __syncwarp();
    // growth-factor matching (see growth_matched_pd): pd = 1 + delta is the coefficient of V in R = G + pd V
    ph.pd = (conv < n_int) ? growth_matched_pd(A_L, A_C, A_U, B_L, B_C, B_U, ph.q, rinv, n_int) : 1.0;
    const double delta = ph.pd - 1.0;
    const double bscale = AMER ? 1.0 : 1.0 / (ph.p + delta);     // the linear march solves for R / (p + delta)
    ph.cA = (lane == 0) ? A_L * bscale : ((lane == 31) ? A_U * bscale : 0.0);
    ph.cB = (lane == 0) ? B_L * bscale : ((lane == 31) ? B_U * bscale : 0.0);

    const double scale = AMER ? 1.0 : ph.p + delta;   // linear march: p folded into the final scaling
    double af1 = 1.0, af2 = 1.0, ab1 = 1.0, ab2 = 1.0;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int row = lane * C + k;
        double v_inv = 0.0, v_nm = 0.0, v_ncc = 1.0;
        if (row < n_int) {
            v_inv = tab[min(row, conv - 1)] * scale;
            v_nm = (row == 0) ? 1.0 : -(A_L * tab[min(row - 1, conv - 1)]);
            if (row < n_int - 1) v_ncc = -(A_U * tab[min(row + 1, conv - 1)]);
        }
        inv[k] = v_inv;
        nm[k] = v_nm;
        ncc[k] = v_ncc;
        if (k < H) { af1 *= v_nm; ab2 *= v_ncc; } else { af2 *= v_nm; ab1 *= v_ncc; }
    }
    // forward sweep visits k = 0..H-1 (half 1) then H..C-1 (half 2); backward sweep visits
    // k = C-1..H (half 1) then H-1..0 (half 2)
    T.Hf[lane] = af1;  T.Hf[32 + lane] = af2;
    T.Hb[lane] = ab1;  T.Hb[32 + lane] = ab2;
    double af = af1 * af2, ab = ab1 * ab2;
    const double af_lane = af, ab_lane = ab;
    // Kogge-Stone multipliers: at level lv the incoming partial result is scaled by the product
    // of this lane's current span; lanes without a partner at that distance get 0.
#pragma unroll
    for (int lv = 0; lv < SCAN_LEVELS; ++lv) {
        const int d = 1 << lv;
        const bool has_f = lane >= d;
        const bool has_b = lane + d < 32;
        T.Af[lv * 32 + lane] = has_f ? af : 0.0;
        T.Ab[lv * 32 + lane] = has_b ? ab : 0.0;
        const double tf = __shfl_up_sync(FULL, af, d);
        const double tb = __shfl_down_sync(FULL, ab, d);
        if (has_f) af *= tf;
        if (has_b) ab *= tb;
    }
    // Direct carry: carry_in(l) = sum_j w_j(l) * end(l -/+ j), w_1 = 1, w_{j+1} = w_j * a(l -/+ j).
    // The first neglected weight bounds the truncation error relative to the solution scale;
    // 1e-20 is four orders below the rounding noise of the sweeps themselves.
    double wf = (lane >= 1) ? 1.0 : 0.0, wb = (lane <= 30) ? 1.0 : 0.0;
    int terms = 0;
#pragma unroll
    for (int j = 1; j <= DIRECT_TERMS; ++j) {
        T.Wf[(j - 1) * 32 + lane] = wf;
        T.Wb[(j - 1) * 32 + lane] = wb;
        const double tf = __shfl_up_sync(FULL, af_lane, j);
        const double tb = __shfl_down_sync(FULL, ab_lane, j);
        wf = (lane >= j + 1) ? wf * tf : 0.0;
        wb = (lane + j + 1 <= 31) ? wb * tb : 0.0;
        if ((j & 1) == 0 && terms == 0 && !__any_sync(FULL, fabs(wf) > 1e-20 || fabs(wb) > 1e-20)) terms = j;
    }
    ph.terms = terms;
    __syncwarp();
}

#ifdef FDM_PROFILE_CYCLES
#define FDM_TICK(var, slot)                                             \
    do {                                                                \
        asm volatile("mov.b64 %0, %1;" : "=d"(var) : "d"(var));         \
        long long t_ = clock64();                                       \
        prof_acc[slot] += t_ - prof_t;                                  \
        prof_t = t_;                                                    \
    } while (0)
#define FDM_PROF_PARAMS , long long (&prof_acc)[8], long long &prof_t
#define FDM_PROF_ARGS , prof_acc, prof_t
#else
#define FDM_TICK(var, slot)
#define FDM_PROF_PARAMS
#define FDM_PROF_ARGS
#endif

// One theta-scheme step for the whole warp.  cin_f / cin_b: boundary contribution entering row 0
// (non-zero on lane 0 only) and the top row (non-zero on lane 31 only, through the padded rows).
// Each lane's chunk is swept as two independent half-chains to halve the dependent-FMA latency.
template <int C, bool AMER>
__device__ __forceinline__ void march_step(double (&V)[C], double (&L)[C], const double (&phi)[C],
                                           const double (&nm)[C], const double (&ncc)[C],
                                           const double (&inv)[C], const double pd, const double q,
                                           const double cin_f, const double cin_b, const int terms,
                                           const WarpTables<C> &T, const int lane FDM_PROF_PARAMS)
{
    constexpr int H = (C + 1) / 2;
    // subtrahend G = q V + L and right-hand side R = (p + delta) V + L = G + pd V (p - q = 1 for every theta)
    if (AMER) {
        american_rhs<C, warp_relu_fp64<C>()>(V, L, q, pd);
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k) L[k] = q * V[k];          // p is folded into inv[] for the linear (European) march
    }
    // ---- forward elimination: u_i = R_i - m_i u_{i-1} ------------------------------------------
    double e1 = cin_f, e2 = 0.0;
#pragma unroll
    for (int k = 0; k < H; ++k) e1 = fma(nm[k], e1, V[k]);
#pragma unroll
    for (int k = H; k < C; ++k) e2 = fma(nm[k], e2, V[k]);
    double e_end = fma(T.Hf[32 + lane], e1, e2);
    FDM_TICK(e_end, 0);
    double cs = lane_carry<true>(e_end, terms, T.Wf, T.Af, lane);
    FDM_TICK(cs, 1);
    {
        double c1 = cs + cin_f;                       // carry into half 1
        double c2 = fma(T.Hf[lane], cs, e1);          // carry into half 2 = true end of half 1
#pragma unroll
        for (int k = 0; k < H; ++k) { c1 = fma(nm[k], c1, V[k]); V[k] = c1; }
#pragma unroll
        for (int k = H; k < C; ++k) { c2 = fma(nm[k], c2, V[k]); V[k] = c2; }
        FDM_TICK(c2, 2);
    }
    // ---- back substitution on xs_i = den_i x_i: xs_i = u_i - cc_i xs_{i+1} ----------------------
    e1 = cin_b;
    e2 = 0.0;
#pragma unroll
    for (int k = C - 1; k >= H; --k) e1 = fma(ncc[k], e1, V[k]);
#pragma unroll
    for (int k = H - 1; k >= 0; --k) e2 = fma(ncc[k], e2, V[k]);
    e_end = fma(T.Hb[32 + lane], e1, e2);
    FDM_TICK(e_end, 3);
    cs = lane_carry<false>(e_end, terms, T.Wb, T.Ab, lane);
    FDM_TICK(cs, 4);
    double c1 = cs + cin_b;
    double c2 = fma(T.Hb[lane], cs, e1);
    // the two half-chains advance together: node C-1-i of half 1 with node H-1-i of half 2
#pragma unroll
    for (int i = 0; i < H; ++i) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int k = half == 0 ? C - 1 - i : H - 1 - i;
            if (half == 0 && k < H) continue;
            double xs;
            if (half == 0) { c1 = fma(ncc[k], c1, V[k]); xs = c1; }
            else           { c2 = fma(ncc[k], c2, V[k]); xs = c2; }
            const double vc = fma(inv[k], xs, -L[k]);
            if (AMER) {
                // ex = phi > vc (fd_american_option_pricer.py:496); V = ex ? phi : vc;
                // dt*lambda_new = ex ? phi - vc : 0 (:498-500)
                american_project<warp_relu_fp64<C>()>(phi[k], vc, V[k], L[k]);
            } else {
                V[k] = vc;
            }
        }
    }
    FDM_TICK(V[0], 5);
}

// ---------------------------------------------------------------------------------------------
// the solve kernel: one warp (one CTA of 32 threads) per PDE solve
// ---------------------------------------------------------------------------------------------

// resident warps per SM the register allocation is sized for
#ifndef FDM_MINB
#define FDM_MINB 0
#endif
template <int C, bool AMER>
constexpr int min_blocks()
{
    return FDM_MINB ? FDM_MINB : (C <= 4 ? 16 : C <= 8 ? 12 : 8);
}


// one solve by the calling warp.  only_rest != 0: only what fdm_lean_kernel left -- sizes outside its classes and solves
// it marked for retry
template <int C, bool AMER, bool TRACE>
__device__ __forceinline__ void warp_solve(const FdmBatch &B, const FdmOut &O, const int64_t sid, const int want_american,
                                           const int div_smem, const int only_rest, double *smem)
{
    const WarpTables<C> T(smem);
    double *divbuf = T.end();

    if (sid >= B.n_solves) return;
    const int flags = B.flags[sid];
    // each launch handles one family (American / linear); the other family's solves are skipped
    if (((flags & FDM_FLAG_AMERICAN) != 0) != (want_american != 0)) return;
    if (flags & FDM_FLAG_INVALID) return;                       // rejected by the book fan-out: status stays -1

    const int lane = threadIdx.x;
    const int n = B.n_space[sid];
    if (n > FDM_WARP_MAX_N_SPACE) return;                       // larger grids: fdm_cta_kernel
    if (only_rest && lean::class_of(n - 1) != 0 &&
        (unsigned long long)__double_as_longlong(O.price[sid]) != lean::RETRY_BITS) return;
    const RowLayout ly = RowLayout::uniform(C, lane);
    const int n_int = n - 1;
    const bool is_call = flags & FDM_FLAG_CALL;
    const bool b76 = flags & FDM_FLAG_BLACK76;
    const bool snap = flags & FDM_FLAG_SNAP;
    const bool ko = flags & (FDM_FLAG_KO_DOWN | FDM_FLAG_KO_UP);

    const double x_min = B.x_min[sid];
    const double dx = B.dx[sid];
    const double sigma = B.sigma[sid];
    const double r = B.r[sid];
    const double bcar = B.b[sid];
    const double spot = B.spot[sid];
    const double strike = B.strike[sid];
    const double rebate = (ko && B.rebate) ? B.rebate[sid] : 0.0;
    const int rannacher = B.rannacher[sid];

    // ---- grid, snapping, payoff -------------------------------------------------------------
    double V[C], L[C], phi[C], nm[C], ncc[C], inv[C];
    const double s0 = node_s(x_min, dx, 0);
    const double sn = node_s(x_min, dx, n);
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = lane * C + k + 1;
        phi[k] = (j <= n - 1) ? node_s(x_min, dx, j) : 0.0;     // node prices, for now
    }
    int spot_idx = 0, strike_idx = 0;
    double k_eff = strike;
    if (snap) {
        spot_idx = argmin_node<C>(phi, s0, sn, spot, n, ly, false);
        strike_idx = argmin_node<C>(phi, s0, sn, strike, n, ly, false);
        k_eff = node_s(x_min, dx, strike_idx);                  // fd_american_option_pricer.py:272-275
    }
    int ko_a, ko_b;
    ko_thresholds(B, sid, flags, x_min, dx, n, ko_a, ko_b);
    const bool put_smin = flags & FDM_FLAG_PUT_BDRY_SMIN;
    const double sn_bdry = (flags & FDM_FLAG_BDRY_TOP_NEXT) ? node_s(x_min, dx, n + 1) : sn;
    const double b_bdry = B.b_bdry ? B.b_bdry[sid] : bcar;
    const double phi0 = payoff(s0, k_eff, is_call);
    const double phin = payoff(sn, k_eff, is_call);
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = lane * C + k + 1;
        const double pv = (j <= n - 1) ? payoff(phi[k], k_eff, is_call) : 0.0;
        phi[k] = pv;
        V[k] = pv;                                              // terminal condition, :558
        L[k] = 0.0;
    }
    // boundary node values live on the lane that owns the adjacent row: V[0] on lane 0, V[n] on lane 31
    double Vb = (lane == 0) ? phi0 : ((lane == 31) ? phin : 0.0);
    const bool bdry_lane = is_call ? (lane == 31) : (lane == 0);

    // ---- operator coefficients, fd_american_option_pricer.py:413-426 -------------------------
    const double sig2 = __dmul_rn(sigma, sigma);
    const double mu_x = b76 ? __dmul_rn(-0.5, sig2) : __dsub_rn(bcar, __dmul_rn(0.5, sig2));
    const double alpha = __dmul_rn(0.5, sig2) / __dmul_rn(dx, dx);
    const double beta = mu_x / __dmul_rn(2.0, dx);
    const double ca = __dsub_rn(alpha, beta);
    const double cc = __dadd_rn(alpha, beta);
    const double bcoef = __dsub_rn(__dmul_rn(-2.0, alpha), r);

    // ---- time march ---------------------------------------------------------------------------
    const int seg_b = B.seg_begin[sid], seg_e = B.seg_begin[sid + 1];
    int mon_i = 0, mon_e = 0;
    if (ko && B.mon_begin) { mon_i = B.mon_begin[sid]; mon_e = B.mon_end ? B.mon_end[sid] : B.mon_begin[sid + 1]; }
    int next_mon = (mon_i < mon_e) ? B.mon_step[mon_i] : -1;
    int gstep = 0;                                              // steps marched so far (all segments)
#ifdef FDM_PROFILE_CYCLES
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long prof_start = clock64();
    long long prof_t = prof_start;
#endif
    int32_t *trace = (TRACE && O.ex_trace) ? O.ex_trace + sid * O.ex_trace_stride : nullptr;

    for (int seg = seg_b; seg < seg_e; ++seg) {
        const int n_steps = B.seg_steps[seg];
        const double dt = B.seg_dt[seg];
        const double tau0 = B.seg_tau0[seg];
        const int base_ran = B.seg_restart[seg] ? rannacher : 0;   // :464
        const double cash = B.seg_div[seg];
        if (AMER) {
#pragma unroll
            for (int k = 0; k < C; ++k) L[k] = 0.0;             // lambda restarts every segment, :463
        }
        int m = 0;
        // Boundary values.  A call has V_min = 0 and a put V_max = 0 at every tau (:286-298), so one
        // sequence w(tau) describes the moving boundary.  Every 32 steps each lane evaluates the
        // exponentials of one future step (tau = tau_start + (m+1) dt); a shuffle hands them out.
        double my_w = 0.0;
        int fill_base = -64;                                     // forces a refill at the first fetch
        auto fetch_boundary = [&](int step) -> double {
            if (step - fill_base >= 32 || step < fill_base) {
                fill_base = step;
                const double tau = fma((double)(step + lane + 1), dt, tau0);
                double vmin, vmax;
                boundary_values(is_call, b76, r, b_bdry, k_eff, sn_bdry, tau, vmin, vmax, put_smin, s0);
                my_w = is_call ? vmax : vmin;
            }
            const double w = __shfl_sync(FULL, my_w, step - fill_base);
            return bdry_lane ? w : 0.0;                          // value on this lane's boundary node
        };
        while (m < n_steps) {
            const bool implicit_phase = m < base_ran;
            const int m_end = implicit_phase ? min(base_ran, n_steps) : n_steps;
            PhaseScalars ph;
            setup_phase<C, AMER>(implicit_phase ? 1.0 : 0.5, dt, ca, cc, bcoef, n_int, lane, nm, ncc, inv, T, ph);
            double w_next = fetch_boundary(m);
            double cin = ph.cB * Vb - ph.cA * w_next;            // non-zero on lanes 0 and 31 only
            // (one loop for every carry form: compiling it per form, as fdm_lean_kernel does, measured +2 % on 128..256
            // node grids and -5 % on 500-node dividend solves -- at 254 registers the specialised loops spill)
            for (; m < m_end; ++m) {
                const double w_now = w_next;
                // the following step's boundary value is fetched early: its latency hides behind this step
                w_next = fetch_boundary(m + 1);
                double cin_n = ph.cB * w_now - ph.cA * w_next;
                march_step<C, AMER>(V, L, phi, nm, ncc, inv, ph.pd, ph.q, (lane == 0) ? cin : 0.0,
                                    (lane == 31) ? cin : 0.0, ph.terms, T, lane FDM_PROF_ARGS);
                Vb = w_now;                                     // V[0] = V_min, V[-1] = V_max, :505-506
                ++gstep;
                if (TRACE && trace) {
                    int lo = 0x7fffffff, hi = -1;
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        const int row = lane * C + k;
                        if (AMER && row < n_int && L[k] > 0.0) { lo = min(lo, row); hi = max(hi, row); }
                    }
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) {
                        lo = min(lo, __shfl_xor_sync(FULL, lo, d));
                        hi = max(hi, __shfl_xor_sync(FULL, hi, d));
                    }
                    if (lane == 0 && 2 * (int64_t)gstep <= O.ex_trace_stride) {
                        trace[2 * (gstep - 1)] = (hi < 0) ? -1 : lo;
                        trace[2 * (gstep - 1) + 1] = hi;
                    }
                }
                if (gstep == next_mon) {                        // discrete_barrier_fdm_pricer_cn.py:299-300
                    const double reb = B.mon_rebate ? B.mon_rebate[mon_i] : rebate;
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        const int j = lane * C + k + 1;
                        if (j <= n - 1 && (j <= ko_a || j >= ko_b)) V[k] = reb;
                    }
                    if (lane == 0 && 0 <= ko_a) Vb = reb;
                    if (lane == 31 && n >= ko_b) Vb = reb;
                    ++mon_i;
                    next_mon = (mon_i < mon_e) ? B.mon_step[mon_i] : -1;
                    cin_n = ph.cB * Vb - ph.cA * w_next;
                }
                cin = cin_n;
            }
        }
        if (cash != 0.0) {
            if (div_smem) {
                double *xs = divbuf;
                double *ys = xs + (32 * C + 2);
                double *mu = ys + (32 * C + 2);
                double *cz = mu + (32 * C + 2);
                double V0 = __shfl_sync(FULL, Vb, 0), VN = __shfl_sync(FULL, Vb, 31);
                dividend_jump<C>(V, V0, VN, [&](int k) { double v = 0.0;
#pragma unroll
                    for (int kk = 0; kk < C; ++kk) v = (kk == k) ? phi[kk] : v;
                    return v; }, phi0, phin, cash, is_call, n, lane, ly, x_min, dx, xs, ys, mu, cz);
                Vb = (lane == 0) ? V0 : ((lane == 31) ? VN : 0.0);
            } else {
#pragma unroll
                for (int k = 0; k < C; ++k) V[k] = NAN;         // launch without dividend scratch: flag it
            }
        }
    }

#ifdef FDM_PROFILE_CYCLES
    if (blockIdx.x == 0 && lane == 0) {
        const long long tot = clock64() - prof_start;
        printf("C=%d AMER=%d steps=%d total=%lld cyc/step=%.1f | rg+fwd1 %.1f carry1 %.1f fwd2 %.1f bwd1 %.1f carry2 %.1f bwd2+proj %.1f\n",
               C, (int)AMER, gstep, tot, (double)tot / gstep, (double)prof_acc[0] / gstep, (double)prof_acc[1] / gstep,
               (double)prof_acc[2] / gstep, (double)prof_acc[3] / gstep, (double)prof_acc[4] / gstep, (double)prof_acc[5] / gstep);
    }
#endif
    // ---- read-out -------------------------------------------------------------------------------
    const double V0 = __shfl_sync(FULL, Vb, 0), VN = __shfl_sync(FULL, Vb, 31);
    int ex_lo = -1, ex_hi = -1;
    if (AMER) {
        int lo = 0x7fffffff, hi = -1;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int row = lane * C + k;
            if (row < n_int && L[k] > 0.0) { lo = min(lo, row); hi = max(hi, row); }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            lo = min(lo, __shfl_xor_sync(FULL, lo, d));
            hi = max(hi, __shfl_xor_sync(FULL, hi, d));
        }
        if (hi >= 0) { ex_lo = lo; ex_hi = hi; }
    }

    if (O.v_dump) {
        double *dst = O.v_dump + sid * O.v_dump_stride;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int j = lane * C + k + 1;
            if (j <= n - 1 && j < O.v_dump_stride) dst[j] = V[k];
        }
        if (lane == 0) {
            dst[0] = V0;
            if (n < O.v_dump_stride) dst[n] = VN;
        }
    }

    double price, delta, gamma;
    double sv[4], ss[4];
    int idx_a, idx_b;
    if (snap) {
        // _interp_price at the snapped spot is V at that node; _local_cubic_delta_gamma,
        // fd_american_option_pricer.py:632-653
        int i_near = min(max(spot_idx, 1), n - 2);
        const double Sx = node_s(x_min, dx, spot_idx);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = i_near - 1 + t;
            sv[t] = node_value<C>(V, V0, VN, j, n, ly);
            ss[t] = node_s(x_min, dx, j);
        }
        price = node_value<C>(V, V0, VN, spot_idx, n, ly);
        // cubic through the four points in Newton form, derivatives at z = 0 (z = S - Sx)
        const double z0 = ss[0] - Sx, z1 = ss[1] - Sx, z2 = ss[2] - Sx, z3 = ss[3] - Sx;
        const double f01 = (sv[1] - sv[0]) / (z1 - z0);
        const double f12 = (sv[2] - sv[1]) / (z2 - z1);
        const double f23 = (sv[3] - sv[2]) / (z3 - z2);
        const double f012 = (f12 - f01) / (z2 - z0);
        const double f123 = (f23 - f12) / (z3 - z1);
        const double f0123 = (f123 - f012) / (z3 - z0);
        delta = f01 + f012 * (-(z0 + z1)) + f0123 * (z0 * z1 + z0 * z2 + z1 * z2);
        gamma = 2.0 * f012 - 2.0 * f0123 * (z0 + z1 + z2);
        idx_a = spot_idx;
        idx_b = strike_idx;
    } else {
        // discrete_barrier_fdm_pricer_cn.py:308-353
        double sl[C];
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int j = lane * C + k + 1;
            sl[k] = (j <= n - 1) ? node_s(x_min, dx, j) : 0.0;
            if (j <= n - 1 && sl[k] <= spot) ++cnt;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
        int lo = cnt;                        // largest j with s_j <= spot, given s_0 <= spot
        if (spot <= s0) { price = V0; lo = 0; }
        else if (spot >= sn) { price = VN; lo = n; }
        else {
            const double s_lo = node_s(x_min, dx, lo), s_hi = node_s(x_min, dx, lo + 1);
            const double v_lo = node_value<C>(V, V0, VN, lo, n, ly);
            const double v_hi = node_value<C>(V, V0, VN, lo + 1, n, ly);
            const double w = __dsub_rn(spot, s_lo) / __dsub_rn(s_hi, s_lo);
            price = __dadd_rn(__dmul_rn(__dsub_rn(1.0, w), v_lo), __dmul_rn(w, v_hi));
        }
        const int gi = argmin_node<C>(sl, s0, sn, spot, n, ly, true);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            sv[t] = node_value<C>(V, V0, VN, gi - 1 + t, n, ly);
            ss[t] = node_s(x_min, dx, gi - 1 + t);
        }
        sv[3] = 0.0;
        ss[3] = 0.0;
        const double h1 = __dsub_rn(ss[1], ss[0]);
        const double h2 = __dsub_rn(ss[2], ss[1]);
        const double h12 = __dadd_rn(h1, h2);
        delta = __dadd_rn(__dadd_rn(__dmul_rn(-h2 / __dmul_rn(h1, h12), sv[0]),
                                    __dmul_rn(__dsub_rn(h2, h1) / __dmul_rn(h1, h2), sv[1])),
                          __dmul_rn(h1 / __dmul_rn(h2, h12), sv[2]));
        gamma = __dmul_rn(2.0, __dadd_rn(__dsub_rn(sv[0] / __dmul_rn(h1, h12), sv[1] / __dmul_rn(h1, h2)),
                                         sv[2] / __dmul_rn(h2, h12)));
        idx_a = lo;
        idx_b = gi;
    }

    if (lane == 0) {
        O.price[sid] = price;
        if (O.delta) O.delta[sid] = delta;
        if (O.gamma) O.gamma[sid] = gamma;
        if (O.stencil_v) {
#pragma unroll
            for (int t = 0; t < 4; ++t) O.stencil_v[4 * sid + t] = sv[t];
        }
        if (O.stencil_s) {
#pragma unroll
            for (int t = 0; t < 4; ++t) O.stencil_s[4 * sid + t] = ss[t];
        }
        if (O.idx) {
            O.idx[4 * sid + 0] = idx_a;
            O.idx[4 * sid + 1] = idx_b;
            O.idx[4 * sid + 2] = ex_lo;
            O.idx[4 * sid + 3] = ex_hi;
        }
        if (O.status) O.status[sid] = isfinite(price) ? 0 : 1;
    }
}

// whether a solve is this kernel's in only_rest mode (cheap test on the descriptors and the lean kernel's hand-back mark)
__device__ __forceinline__ bool rest_candidate(const FdmBatch &B, const FdmOut &O, const int64_t sid, const int want_american)
{
    if (sid >= B.n_solves) return false;
    const int flags = B.flags[sid];
    if (((flags & FDM_FLAG_AMERICAN) != 0) != (want_american != 0) || (flags & FDM_FLAG_INVALID)) return false;
    const int n = B.n_space[sid];
    if (n > FDM_WARP_MAX_N_SPACE) return false;
    return lean::class_of(n - 1) == 0 || (unsigned long long)__double_as_longlong(O.price[sid]) == lean::RETRY_BITS;
}

// Launch shapes: normally one CTA (= one warp) per solve.  In only_rest mode -- after the lean kernels, when most or
// all solves are already done -- one warp screens 32 consecutive solves (one coalesced read of their descriptors)
// and solves the few that are left, so a launch over 800 000 finished solves is 25 000 CTAs that exit at once
// instead of 800 000 (1.0 ms -> ~0.03 ms on B200).
template <int C, bool AMER, bool TRACE>
__global__ void __launch_bounds__(32, min_blocks<C, AMER>())
fdm_warp_kernel(const FdmBatch B, const FdmOut O, const int64_t first, const int want_american,
                const int div_smem, const int only_rest)
{
    extern __shared__ double smem[];
    int64_t base;
    unsigned todo;
    if (only_rest) {
        base = first + (int64_t)blockIdx.x * 32;
        todo = __ballot_sync(FULL, rest_candidate(B, O, base + threadIdx.x, want_american));
    } else {
        base = first + blockIdx.x;
        todo = 1u;
    }
    while (todo) {
        const int l = __ffs(todo) - 1;
        todo &= todo - 1;
        warp_solve<C, AMER, TRACE>(B, O, base + l, want_american, div_smem, only_rest, smem);
        __syncwarp();
    }
}

}  // namespace
