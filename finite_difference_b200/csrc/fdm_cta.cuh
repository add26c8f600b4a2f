// fdm_cta.cuh -- one CTA per PDE solve, for grids too large for one warp's registers
// (514 .. 8193 space nodes; BASELINE config 4 is 8192 nodes x 10 000 steps).
//
// Same step algebra as fdm_warp_kernel (see the header of fdm_warp.cuh), different residency:
//   * thread t of the CTA owns C contiguous slots; slot s holds interior row s - pad, so the padding
//     sits BELOW row 0 and the last slot of the last thread is the top row.  V and dt*lambda stay in
//     registers, the payoff lives in shared memory ([k][t] layout, conflict-free);
//   * the LU pivots of the Toeplitz matrix converge bit-wise after a few hundred rows at most, so
//     only the warps that hold padding or the first P rows read per-slot coefficients (staged per
//     thread in shared memory); every other warp runs with three scalar coefficients, and warps that
//     hold nothing but padding only keep the barriers.  The choice is warp-uniform: no divergence;
//   * each first-order recurrence runs as  local pass -> intra-warp shuffle scan -> one __syncthreads ->
//     short Horner sum over the preceding warps' totals -> local pass.  The warp-to-warp multipliers decay like
//     |c'|^(32 C) (1e-17 at C = 16), so the Horner sum is truncated after K terms with the neglected
//     weight below 1e-22 (K is found per phase).
//   * the time loop is compiled once per warp role (padding, table, scalar, scalar in row-0 mode, owner of row 0,
//     correction-reach) and per form of the warp-to-warp sums (one term / general), picked by a switch outside the
//     loop: between two block barriers every warp runs straight-line code (see the kernel; 512 -> ~350 instructions
//     per warp-step, +50..75 % over the one-loop form on every class);
//   Two block-wide barriers per time step.  (Tried and dropped: passing the warp totals through a shared-memory
//   ring with per-warp progress counters instead of the barriers -- the warps need their nearest neighbours only --
//   measured 8..10 % slower on 2048..8192-node grids: the warps stay in lock-step anyway and the polls cost issue slots.)
//   * size classes (warps x slots per thread): 4x12, 4x16, 5x14, 5x16, 8x12, 8x16, 16x12, 16x16 serve up to 1536 / 2048 /
//     2240 / 2560 / 3072 / 4096 / 6144 / 8192 rows; a solve runs in the smallest class that holds it (2133-node barrier
//     grids: 5x14, four CTAs per SM).
//   * American state between steps: (vc, 2 dt lambda), not (V, 2 dt lambda) -- 9 FP64 instructions per node-step; the
//     linear march keeps V only and rebuilds V_old from u in its last pass (lin_vc; not in the 16-warp class).
//   What was measured and dropped (more slots per thread with fewer warps, shared-memory progress counters instead of the
//   barriers, named barriers per warp pair, the lean kernel's step order, ...) is listed in DESIGN.md section 3.
//
// Included by fdm_cta_inst.cu (one translation unit per size class).

#pragma once

#include "fdm_common.cuh"

namespace cta {

#ifdef FDM_CTA_EXP_NOSYNC           // timing experiment only (wrong results): no block barriers in the step
#define FDM_CTA_STEP_SYNC() __syncwarp()
#else
#define FDM_CTA_STEP_SYNC() __syncthreads()
#endif

constexpr double TRUNC_EPS = 1e-22;

// Row-0 correction (see setup_phase) instead of per-slot table warps, in every class.  (Before the time loop was
// compiled per warp role the 16-warp linear variant kept the tables: at its 128-register cap the correction's extra
// live values spilled inside the loop, 4.1 against 5.5e11 on 8192-node barrier grids.  With the per-role loops the
// correction wins there too: 7.8e11 -> 1.09e12.)  FDM_CTA_NO_ROW0_CORR builds the table-only kernel.
template <int C, int WARPS, bool AMER>
__host__ __device__ constexpr bool row0_mode()
{
#ifdef FDM_CTA_NO_ROW0_CORR
    return false;
#else
    return true;
#endif
}

// One-term warp-to-warp sums compiled as straight-line code (K1 loops): +15..30 % in every class.  (A first version
// ran a 16-warp American batch that mixed K1 and general solves 28 % SLOWER, 8192 nodes: 3.70 -> 5.18 ms, while every
// solve alone was faster: the loops were ~1000 instructions each and several were live per SM.  Fixed by shrinking the
// loops -- boundary exponentials out of line, general Horner loops not unrolled; FDM_CTA_K1_NOT_16A restores the
// exception.)
template <int C, int WARPS, bool AMER>
__host__ __device__ constexpr bool k1_mode()
{
#ifdef FDM_CTA_K1_NOT_16A
    return !(WARPS == 16 && AMER);
#else
    return true;
#endif
}

// nearest node to `level` with the reference's first-minimum tie-break
// (fd_american_option_pricer.py:253; interior_only: discrete_barrier_fdm_pricer_cn.py:335).
// |s_j - level| is V-shaped in j, so a window around the analytic position holds the argmin.
__device__ __forceinline__ int nearest_node(double x_min, double dx, int n, double level, bool interior_only)
{
    const int lo_j = interior_only ? 1 : 0, hi_j = interior_only ? n - 1 : n;
    double xj = (log(level) - x_min) / dx;
    if (!(xj > -1e9)) xj = -1e9;
    if (xj > 1e9) xj = 1e9;
    int j0 = (int)floor(xj);
    j0 = min(max(j0, lo_j), hi_j);
    double best = INFINITY;
    int bidx = lo_j;
    for (int j = max(lo_j, j0 - 2); j <= min(hi_j, j0 + 3); ++j) {
        const double v = fabs(__dsub_rn(node_s(x_min, dx, j), level));
        if (v < best) { best = v; bidx = j; }
    }
    return bidx;
}

template <int C, int WARPS>
struct Layout {
    static constexpr int NT = 32 * WARPS;
    static constexpr int NP = NT * C + 16;                           // n + 1 nodes, rounded up to a block of jsw()
    // warps that may carry per-slot coefficients: 6 x 512 slots hold the padding plus a head of up to ~3000
    // rows (alpha*dt up to ~4e4 at 8193 nodes); stiffer solves are flagged with status 2
    // Per-slot coefficient tables serve the rare table mode only (pivots not converged / far-reaching row-0 tail); two
    // staged table warps (1024 head rows) keep shared memory small enough for 4 / 3 / 2 resident CTAs of 4 / 5 / 8 warps
    // (more staged slots would cost a CTA per SM); the 16-warp class has the SM to itself and stages six.
#ifndef FDM_CTA_SLOTS_SMALL
#define FDM_CTA_SLOTS_SMALL 2
#endif
    static constexpr int SLOTS = WARPS <= 8 ? (WARPS < FDM_CTA_SLOTS_SMALL ? WARPS : FDM_CTA_SLOTS_SMALL) : 6;
    static constexpr int SLOT_DOUBLES = 3 * C * 32 + 8 * 32;         // nm, ncc, inv [k][lane] + 8 per-lane constants
    static constexpr int R2 = NP > SLOTS * SLOT_DOUBLES ? NP : SLOTS * SLOT_DOUBLES;
    static constexpr int SMALL = 4 * WARPS + 16 + 64 + 32 + 8 + C + 8 + 320;   // part, cst, PAc, wb, ishare, nmp, smv, scw
    static constexpr int TOTAL = 2 * NP + R2 + SMALL;           // American
    static constexpr int TOTAL_LINEAR = NP + R2 + SMALL;        // no payoff table r0 (and no dividend jumps: see the kernel)
};

struct Sm {
    double *r0;     // payoff by slot [k*NT + t]                      | dividend jump: cz
    double *r1;     // pivot tables of the phase                      | dividend jump: ys; read-out: V by node
    double *r2;     // per-slot coefficients of the table warps       | dividend jump: mu
    double *part;   // pE_f, pA_f, pE_b, pA_b [WARPS]
    double *cst;    // phase constants of the scalar-coefficient warps: NM NCC INV aF aB hF1 hF2 hB1 hB2 top_mult pd
    double *PAc;    // [64] exclusive lane products of those warps, forward / backward
    double *wb;     // [32] moving Dirichlet values of the next 32 steps
    int *ish;       // [13] row-0 Dirichlet term by slot  [14,15] scan levels  [0] P [1] spot node [2] strike node [3,4] KO range [5] status [6,7] ex lo/hi [8] Kf [9] Kb
                    // [10] first table warp [11] last table warp [12] last warp the row-0 correction reaches (-1: table mode)
    double *nmp;    // [C] powers NM^k of the steady-state forward multiplier (row-0 correction)
    double *scw;    // [2][5][32] masked level weights of the scalar warps' scans (forward, backward): a^(2^lv) or 0
    double *smv;    // [0] GF NCC^-k0, factor of the row-0 correction  [1] xs at slot 0 of the owner of row 0, current step
                    // [2] Kc: what the Dirichlet carry leaves in [1] through the dead slots below row 0
                    // [3] [4] the carry enters below slot 0 / below slot C/2 (1 or 0)  [5] its weight at the thread's end
};

enum { CST_NM = 0, CST_NCC, CST_INV, CST_AF, CST_AB, CST_HF1, CST_HF2, CST_HB1, CST_HB2, CST_TOP, CST_PD, CST_BS };
enum { LC_AF = 0, LC_AB, LC_HF1, LC_HF2, LC_HB1, LC_HB2, LC_PAF, LC_PAB };

// warp roles of the specialised time loops (kernel): fixed per warp for a phase
enum { ROLE_ANY = -1, ROLE_PAD = 0, ROLE_TBL, ROLE_SCALAR, ROLE_SCALAR_R0, ROLE_OWNER, ROLE_SMC, ROLE_OWNER_H, ROLE_OWNER_SLOT };
// how the scalar owner warp of row 0 receives the Dirichlet term of row 0 (setup_phase)
enum { INJ_NONE = 0, INJ_CARRY0, INJ_CARRYH, INJ_SLOT, INJ_RUNTIME };   // carry below slot 0 / below slot C/2 / by slot

struct Ph {
    double q, pd, NM, NCC, INV;   // pd: growth-matched p - q + delta (linear); American: kappa (setup_phase)
    double gz;              // row-0 correction weight of this thread: GF * NM^(row of slot 0), 0 where it does not reach
    int smc;                // this warp applies the row-0 correction in step_c
    int bar3;               // the correction reaches beyond the warp of row 0: a third block barrier per step
    double cF;              // owner of row 0, carry form: NM^-(k0+1), turns the Dirichlet term of row 0 into a forward carry; else 0
    int inj;                // owner warp: the Dirichlet term of row 0 is added at its slot instead (setup_phase)
    int e0;                 // row of this thread's slot 0 (negative inside the padding below row 0)
    double cA, cB;          // boundary coefficients of this thread (row-0 owner: A_L, B_L; last thread: A_U, B_U)
    int Kf, Kb;             // terms of the warp-to-warp Horner sums
    int cls;                // 0 padding only, 1 scalar coefficients, 2 per-slot coefficients
    const double *slot;     // this warp's coefficient block (cls == 2)
};

template <int C, int WARPS, bool AMER>
__device__ __forceinline__ void setup_phase(double theta, double dt, double a, double c, double bcoef, int n_int,
                                            int pad, const Sm &sm, Ph &ph)
{
    using LY = Layout<C, WARPS>;
    constexpr int H = C / 2;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const double A_L = __dmul_rn(__dmul_rn(-theta, dt), a);
    const double A_C = __dsub_rn(1.0, __dmul_rn(__dmul_rn(theta, dt), bcoef));
    const double A_U = __dmul_rn(__dmul_rn(-theta, dt), c);
    const double B_L = __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), a);
    const double B_C = __dadd_rn(1.0, __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), bcoef));
    const double B_U = __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), c);
    const double p = 1.0 / theta;
    ph.q = (1.0 - theta) / theta;
    const int cap = LY::NP / 3 - 2;                     // head rows the pivot tables can hold
    double *pv_nm = sm.r1, *pv_ncc = sm.r1 + (cap + 1), *pv_inv = sm.r1 + 2 * (cap + 1);
    __syncthreads();
    if (t == 0) {
        // pivots: den_0 = A_C, den_i = A_C - A_L*A_U/den_{i-1} (fd_american_option_pricer.py:443-451);
        // entry i < P is row i, entry P the steady state
        double r_prev = 0.0, r_cur = __drcp_rn(A_C);
        int P = -1;
        // Fast path (the usual case): with the row-0 correction only the STEADY state of the pivots is used, so the
        // table below is not needed -- and the steady state need not be walked to from row 0.  The root of
        // den^2 - A_C den + A_L A_U = 0 says where it is; the reference's own recurrence, started a hair above the root,
        // settles on the same bit-wise fixed point as the reference's sweep from A_C (reference_steady_state) in a third
        // of the iterations.  (Thread 0 walked both recurrences over 100..500 rows before: 11 % of a 2500-step solve
        // at 8192 nodes with its five phases, 6 % of the 2133 x 500 barrier-runner solves.)
        bool fast = false;
        double den_ref = 0.0, cp_ref = 0.0;
        if (row0_mode<C, WARPS, AMER>()) {
            const double PLU = __dmul_rn(A_L, A_U), PLU_lo = __fma_rn(A_L, A_U, -PLU);
            const double disc = __fma_rn(A_C, A_C, -4.0 * PLU);
            if (disc > 0.0) {
                double ds = 0.5 * (A_C + sqrt(disc));
                for (int it = 0; it < 2; ++it) {
                    const double res = __fma_rn(__dsub_rn(ds, A_C), ds, PLU) + PLU_lo;
                    ds -= res / (2.0 * ds - A_C);
                }
                const double rs = 1.0 / ds, nmc = -(A_L * rs), nccc = -(A_U * rs), rho = nmc * nccc;
                if (rho > 1e-20 && rho < 0.999 && fabs(nmc) < 1.0 && fabs(nmc) > 1e-10 && fabs(nccc) > 1e-10) {
                    const double rows_conv = 37.0 / -log(rho);              // rows until the reference's sweep settles
                    const double kappa = 1.0 / (1.0 - rho), dl = A_L * A_U * rs * rs;
                    const double gf = dl * kappa / (1.0 + dl * kappa);
                    const double rows_z = (log(1e-24) - log(fmax(fabs(gf), 1e-300))) / log(fabs(nmc));
                    if (isfinite(gf) && rows_conv < 0.5 * n_int && rows_conv < 0.5 * cap && rows_z + 2.0 < 0.9 * n_int) {
                        const double start = fmin(ds * (1.0 + 0x1p-42), A_C);
                        fast = reference_steady_state(A_L, A_C, A_U, start, 4096, den_ref, cp_ref);
                    }
                }
            }
        }
        if (fast) {
            P = 0;
            r_cur = __drcp_rn(den_ref);
            pv_inv[0] = r_cur; pv_nm[0] = -(A_L * r_cur); pv_ncc[0] = -(A_U * r_cur);
        }
        for (int i = 0; i < cap && !fast; ++i) {
            const double den_next = __dsub_rn(A_C, __dmul_rn(A_L, __dmul_rn(A_U, r_cur)));
            const double r_next = __drcp_rn(den_next);
            pv_inv[i] = r_cur;
            pv_nm[i] = (i == 0) ? 0.0 : -(A_L * r_prev);            // row 0 has no lower neighbour
            pv_ncc[i] = -(A_U * r_next);
            if (i > 0 && r_cur == r_prev && r_next == r_cur) { P = i; break; }
            r_prev = r_cur;
            r_cur = r_next;
        }
        const bool converged = P >= 0;
        if (P < 0) {                                    // not converged within the table: flagged, best effort
            P = cap;
            sm.ish[5] = 2;
            pv_inv[P] = r_cur; pv_nm[P] = -(A_L * r_cur); pv_ncc[P] = -(A_U * r_cur);
        }
        // Row-0 correction (Sherman-Morrison).  The factors built from the steady-state pivot den_inf reproduce A except
        // for entry (0,0): A = A_inf + delta e0 e0^T, delta = A_C - den_inf = A_L A_U / den_inf.  So A x = R is
        // A_inf x = R - g e0 with g = delta x_0, and in the sweeps' variables (u forward, xs = den_inf x backward)
        //     xs'_i = xs_i - g kappa NM^i (1 - rho^(n-i)),  rho = NM NCC,  kappa = 1 / (1 - rho),
        //     g = delta r_inf xs_0 / (1 + delta r_inf kappa)   =>   xs'_i = xs_i - GF xs_0 NM^i.
        // Every warp then runs with the three scalar coefficients; the warps the NM^i tail reaches (normally only the
        // warp of row 0) subtract it in step_c.  Used when the pivots converged and the tail dies inside the grid
        // (rho^(n-i) term below 1e-24); otherwise the per-slot tables below stay in charge.
        double GF = 0.0;
        int wz = -1;
        if (row0_mode<C, WARPS, AMER>() && converged) {
            const double rinf = pv_inv[P], nmc = pv_nm[P], nccc = pv_ncc[P];
            const double kappa = 1.0 / (1.0 - nmc * nccc);
            const double dl = A_L * A_U * rinf * rinf;                  // delta * r_inf
            const double gf = dl * kappa / (1.0 + dl * kappa);
            const double anm = fabs(nmc);
            if (anm > 1e-10 && anm < 1.0 && fabs(nccc) > 1e-10 && isfinite(gf)) {
                const double rows = (log(1e-24) - log(fmax(fabs(gf), 1e-300))) / log(anm);
                const int Pz = rows < 1.0 ? 1 : (rows > 1e9 ? 0x3fffffff : (int)rows + 1);
                if (Pz < n_int) { GF = gf; wz = min((pad + Pz - 1) / (32 * C), WARPS - 1); }
            }
            double pw = 1.0;
            for (int k = 0; k < C; ++k) { sm.nmp[k] = pw; pw *= nmc; }
        }
        {
            // the owner of row 0 publishes xs at its slot 0 = NCC^k0 xs_0 (k0 padding slots below row 0)
            double unw = 1.0;
            for (int k = 0; k < pad % C; ++k) unw /= pv_ncc[P];
            sm.smv[0] = GF * unw;
            // How the Dirichlet term of row 0 reaches its slot k0 = pad % C of the scalar owner warp (row-0 correction
            // mode).  As a carry scaled by NM^-(k0'+1) that enters the thread's forward chains k0' dead slots below row 0
            // -- below slot 0, or below slot C/2 (the second half-chain) when row 0 sits in the upper half: free, but the
            // k0' dead slots then hold u_j = NM^(j+1) carry, which the backward pass sums into [1]; the predicted sum
            // [2] * carry is taken out again, a cancellation that costs a factor rho^-k0' in precision (2e-9 on a
            // 555-node barrier grid when the carry always entered below slot 0: k0 = 10, rho = 0.2).  So only while
            // rho^-k0' <= 1e3; otherwise the term is added at its slot (ish[13]).
            const int k0s = pad % C, kinj = (k0s >= C / 2) ? C / 2 : 0;
            const double nmc = pv_nm[P], rho = pv_nm[P] * pv_ncc[P];
            double kc = 0.0, rj = 1.0, lead = nmc, afx = 1.0;
            for (int k = 0; k < k0s - kinj; ++k) { kc += rj; rj *= rho; }
            for (int k = 0; k < kinj; ++k) lead *= pv_ncc[P];
            for (int k = kinj; k < C; ++k) afx *= nmc;
            const bool by_slot = !(fabs(rj) >= 1e-3);            // rj = rho^k0'
            sm.smv[2] = by_slot ? 0.0 : lead * kc;
            sm.smv[3] = kinj == 0 ? 1.0 : 0.0;                   // the carry enters below slot 0 ...
            sm.smv[4] = kinj == 0 ? 0.0 : 1.0;                   // ... or below slot C/2
            sm.smv[5] = afx;                                     // what it leaves at the end of the thread: NM^(C - kinj)
            sm.ish[13] = by_slot ? 1 : 0;
        }
        sm.ish[12] = wz;
        // growth-factor matching (fdm_b200.cu, growth_matched_pd); the linear march folds p + delta into inv
        const double pd = fast ? growth_pd_from_ref(A_L, den_ref, cp_ref, A_U, B_L, B_C, B_U, ph.q, pv_inv[P])
                          : ((P < cap && P < n_int) ? growth_matched_pd(A_L, A_C, A_U, B_L, B_C, B_U, ph.q, pv_inv[P], n_int) : 1.0);
        double bs = 1.0;
        if (!AMER) {
            const double scale = p + (pd - 1.0);
            for (int i = 0; i <= P; ++i) pv_inv[i] *= scale;
            bs = 1.0 / scale;
        } else {
            // American state between steps: (vc, L = 2 dt lambda) with V = vc + L / 2, not (V, L): the right-hand side
            // R = (pd + q) V + dt lambda = (pd + q) vc + s L, s = (pd + q + 1) / 2, is ONE fma once the system is
            // solved for R / s = kappa vc + L (pivot scaling times s, Dirichlet terms over s), and the projection
            // no longer forms V: 9 FP64 instructions per node-step instead of 10.
            const double s = 0.5 * (pd + ph.q + 1.0);
            for (int i = 0; i <= P; ++i) pv_inv[i] *= s;
            bs = 1.0 / s;
        }
        const int wf = pad / (32 * C);                              // warp that holds row 0
        int wl = min((pad + P - 1) / (32 * C), WARPS - 1);          // last warp with a head row
        if (wz >= 0) wl = wf - 1;                                   // row-0 correction: no table warps
        else if (wl - wf + 1 > LY::SLOTS) sm.ish[5] = 2;            // more table warps than staged slots
        // theta alpha dt beyond ~1.5e4 (a dozen time steps on thousands of nodes): the hoisted factorisation's growth
        // matching is good for ~1e-10 there and no better (1.2e-10 seen at alpha dt = 6e4 on 7346 nodes x 21 steps in the
        // random sweep); such a solve is left to the literal sweep of fdm_exact_kernel as well
        if (fabs(A_L) > 1.5e4 || fabs(A_U) > 1.5e4) sm.ish[5] = 2;
        sm.ish[0] = P; sm.ish[10] = wf; sm.ish[11] = min(wl, wf + LY::SLOTS - 1);
        sm.cst[CST_NM] = pv_nm[P]; sm.cst[CST_NCC] = pv_ncc[P]; sm.cst[CST_INV] = pv_inv[P];
        sm.cst[CST_TOP] = (sm.ish[11] == WARPS - 1) ? 1.0 : 1.0 / pv_ncc[P];
        sm.cst[CST_PD] = AMER ? (pd + ph.q) * bs : pd;            // American: kappa
        sm.cst[CST_BS] = bs;
    }
    __syncthreads();
    const int P = sm.ish[0], wf = sm.ish[10], wl = sm.ish[11];
    ph.NM = sm.cst[CST_NM]; ph.NCC = sm.cst[CST_NCC]; ph.INV = sm.cst[CST_INV]; ph.pd = sm.cst[CST_PD];
    const double bscale = sm.cst[CST_BS];               // the march solves for R / (p + delta) (linear), R / s (American)
    const bool owner0 = (t == pad / C);
    ph.cA = owner0 ? A_L * bscale : ((t == LY::NT - 1) ? A_U * bscale : 0.0);
    ph.cB = owner0 ? B_L * bscale : ((t == LY::NT - 1) ? B_U * bscale : 0.0);
    ph.cls = (w < wf) ? 0 : (w <= wl ? 2 : 1);
#ifdef FDM_CTA_EXP_NOTBL          // timing experiment only (wrong results): no table warps
    if (ph.cls == 2) ph.cls = 1;
#endif
    ph.gz = 0.0; ph.cF = 0.0; ph.inj = 0; ph.smc = 0; ph.bar3 = 0; ph.e0 = t * C - pad;
    if (row0_mode<C, WARPS, AMER>()) {
        const int wz = sm.ish[12];
        ph.smc = (wz >= 0 && w >= wf && w <= wz) ? 1 : 0;
        ph.bar3 = (wz > wf) ? 1 : 0;
        ph.gz = 0.0;
        if (wz >= 0) {
            // dead lanes below row 0 stay at exactly zero with a zero pivot scaling (V = 0 * xs - 0); the owner of
            // row 0 has dead AND live slots: step_cs tells them apart by ph.e0
            const int t0 = pad / C, k0 = pad % C;
            ph.inj = sm.ish[13];
            if (t == t0 && !ph.inj) ph.cF = 1.0 / pow(ph.NM, (double)((k0 >= C / 2 ? k0 - C / 2 : k0) + 1));
            if (t < t0) ph.INV = 0.0;
        }
        const int e0 = t * C - pad;                                  // row of slot 0 (negative inside the padding)
        if (ph.smc && e0 > -C) {
            const double nmc = ph.NM;
            const double mag = pow(fabs(nmc), (double)e0);
            ph.gz = sm.smv[0] * ((nmc < 0.0 && (e0 & 1)) ? -mag : mag);
        }
    }
    ph.slot = sm.r2 + (size_t)max(w - wf, 0) * LY::SLOT_DOUBLES;

    if (ph.cls != 0) {
        double f1 = 1.0, f2 = 1.0, b1 = 1.0, b2 = 1.0;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            double vnm = ph.NM, vncc = ph.NCC;
            if (ph.cls == 2) {
                const int row = t * C + k - pad;
                double vinv = 0.0;
                vnm = 0.0; vncc = 0.0;                               // padding slots are inert
                if (row >= 0) {
                    const int ci = min(row, P);
                    vnm = pv_nm[ci];
                    vncc = (row == n_int - 1) ? 1.0 : pv_ncc[ci];    // top row inside a table warp: carry passes
                    vinv = pv_inv[ci];
                }
                double *sl = sm.r2 + (size_t)(w - wf) * LY::SLOT_DOUBLES;
                sl[k * 32 + lane] = vnm;
                sl[C * 32 + k * 32 + lane] = vncc;
                sl[2 * C * 32 + k * 32 + lane] = vinv;
            }
            if (k < H) { f1 *= vnm; b2 *= vncc; } else { f2 *= vnm; b1 *= vncc; }
        }
        const double aF = f1 * f2, aB = b1 * b2;
        double pf = aF, pb = aB;                         // inclusive lane products, then exclusive
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double tf = __shfl_up_sync(FULL, pf, d);
            const double tb = __shfl_down_sync(FULL, pb, d);
            if (lane >= d) pf *= tf;
            if (lane + d < 32) pb *= tb;
        }
        if (lane == 31) sm.part[1 * WARPS + w] = pf;     // pA_f
        if (lane == 0) sm.part[3 * WARPS + w] = pb;      // pA_b
        const double ef = __shfl_up_sync(FULL, pf, 1), eb = __shfl_down_sync(FULL, pb, 1);
        const double PAf = (lane == 0) ? 1.0 : ef, PAb = (lane == 31) ? 1.0 : eb;
        if (ph.cls == 2) {
            double *lc = sm.r2 + (size_t)(w - wf) * LY::SLOT_DOUBLES + 3 * C * 32;
            lc[LC_AF * 32 + lane] = aF;   lc[LC_AB * 32 + lane] = aB;
            lc[LC_HF1 * 32 + lane] = f1;  lc[LC_HF2 * 32 + lane] = f2;
            lc[LC_HB1 * 32 + lane] = b1;  lc[LC_HB2 * 32 + lane] = b2;
            lc[LC_PAF * 32 + lane] = PAf; lc[LC_PAB * 32 + lane] = PAb;
        } else if (w == wl + 1) {                        // every scalar warp has the same constants
            if (lane == 0) {
                sm.cst[CST_AF] = aF; sm.cst[CST_AB] = aB; sm.cst[CST_HF1] = f1; sm.cst[CST_HF2] = f2;
                sm.cst[CST_HB1] = b1; sm.cst[CST_HB2] = b2;
            }
            sm.PAc[lane] = PAf;
            sm.PAc[32 + lane] = PAb;
            double af = aF, ab = aB;                     // level weights by repeated squaring, as warp_scan forms them
#pragma unroll
            for (int lv = 0; lv < 5; ++lv) {
                sm.scw[lv * 32 + lane] = (lane >= (1 << lv)) ? af : 0.0;
                sm.scw[160 + lv * 32 + lane] = (lane + (1 << lv) < 32) ? ab : 0.0;
                af *= af;
                ab *= ab;
            }
        }
    } else if (lane == 0) {
        sm.part[0 * WARPS + w] = 0.0; sm.part[1 * WARPS + w] = 0.0;
        sm.part[2 * WARPS + w] = 0.0; sm.part[3 * WARPS + w] = 0.0;
    }
    __syncthreads();
    if (t == 0) {
        // terms of the warp-to-warp sums: smallest K whose neglected weight is below TRUNC_EPS for every warp
        int Kf = WARPS - 1, Kb = WARPS - 1;
        bool got_f = false, got_b = false;
        for (int K = 1; K < WARPS && !(got_f && got_b); ++K) {
            double worst_f = 0.0, worst_b = 0.0;
            for (int s0 = wf; s0 + K <= WARPS; ++s0) {
                double qf = 1.0, qb = 1.0;
                for (int u = s0; u < s0 + K; ++u) { qf *= sm.part[1 * WARPS + u]; qb *= sm.part[3 * WARPS + u]; }
                worst_f = fmax(worst_f, fabs(qf));
                worst_b = fmax(worst_b, fabs(qb));
            }
            if (!got_f && worst_f < TRUNC_EPS) { Kf = K; got_f = true; }
            if (!got_b && worst_b < TRUNC_EPS) { Kb = K; got_b = true; }
        }
        sm.ish[8] = Kf; sm.ish[9] = Kb;
        // levels of the scalar warps' lane scans: level lv adds the terms 2^lv .. 2^(lv+1) - 1 lanes away
        auto scan_levels = [](double a) {
            a = fabs(a);
            const double a8 = (a * a) * (a * a) * ((a * a) * (a * a));
            return !(a8 >= TRUNC_EPS) ? 3 : (!(a8 * a8 >= TRUNC_EPS) ? 4 : 5);
        };
        sm.ish[14] = scan_levels(sm.cst[CST_AF]); sm.ish[15] = scan_levels(sm.cst[CST_AB]);
    }
    __syncthreads();
    ph.Kf = sm.ish[8]; ph.Kb = sm.ish[9];
}

// inclusive scan of affine maps across the warp; returns the inclusive end value E_l (zero carry
// into the warp).  TBL: multipliers differ per lane; else they are powers of one constant.
template <bool TBL, bool UP>
__device__ __forceinline__ double warp_scan(double e, double a, int lane)
{
    double E = e, A = a;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double tE = UP ? __shfl_up_sync(FULL, E, d) : __shfl_down_sync(FULL, E, d);
        const bool has = UP ? (lane >= d) : (lane + d < 32);
        if (TBL) {
            const double tA = UP ? __shfl_up_sync(FULL, A, d) : __shfl_down_sync(FULL, A, d);
            if (has) { E = fma(A, tE, E); A *= tA; }
        } else {
            E = fma(has ? A : 0.0, tE, E);
            A *= A;
        }
    }
    return E;
}

// the scalar warps' scan with the masked level weights read from shared memory (one LDS per level instead of a
// compare, two selects and a multiply); same values as warp_scan<false, UP>
#ifdef FDM_CTA_NO_SCAN_TABLE
constexpr bool SCAN_TABLE = false;
#else
constexpr bool SCAN_TABLE = true;
#endif
// `levels` (3..5, uniform over the CTA): the lane-to-lane multiplier a = NM^C is small, and after L levels the scan
// holds the 2^L nearest lanes' terms; what is left out weighs |a|^(2^L), below TRUNC_EPS for the levels skipped
// (setup_phase).  A level is a dependent SHFL + DFMA, ~36 cycles in which the warp has nothing else to do: 1000-node
// American grids +8 %, 3000 / 6000 nodes +3 / +2 %, 8192-node barrier grids +2 %.  Not (SKIP = false) in the 16 x 16
// American class, whose loops sit at the 128-register cap: the two uniform branches cost it 3..5 % (cfg4: 8.25e11
// without, 7.84e11 with) whether levels are skipped or not.
template <int C, int WARPS, bool AMER>
__host__ __device__ constexpr bool scan_level_skip()
{
#ifdef FDM_CTA_NO_LEVEL_SKIP
    return false;
#else
    return !(AMER && WARPS == 16 && C == 16);
#endif
}
template <bool UP, bool SKIP>
__device__ __forceinline__ double warp_scan_tab(double e, const double *__restrict__ wt, int lane, const int levels)
{
    double E = e;
#pragma unroll
    for (int lv = 0; lv < 5; ++lv) {
        if (SKIP && lv >= 3 && lv >= levels) break;
        const double tE = UP ? __shfl_up_sync(FULL, E, 1 << lv) : __shfl_down_sync(FULL, E, 1 << lv);
        E = fma(wt[lv * 32 + lane], tE, E);
    }
    return E;
}

// ---- the three phases of one time step (barriers sit between them, outside any branch) -----------

// Register-light linear march (see lin_vc): no copy of V_old.
// (Not in the 16-warp class: one CTA per SM either way, and the extra instruction costs 8192-node barrier grids 8 %.)
template <int C, int WARPS, bool AMER>
__host__ __device__ constexpr bool lin_light()
{
#ifdef FDM_CTA_NO_LIN_LIGHT
    return false;
#else
    return !AMER && WARPS != 16;
#endif
}

// RHS, forward local pass, intra-warp scan.  k0 >= 0 on the thread that owns row 0.  TBL: its Dirichlet term is added
// to R at that slot.  INJ (the scalar warp that holds row 0, row-0 correction mode): at its slot too (INJ_SLOT), or as
// the carry carryF into one of the thread's two half-chains (INJ_CARRY0 / INJ_CARRYH); setup_phase says which, and
// each form has its own time loop (the owner is the warp the others wait for: every instruction on it shows).
template <int C, int WARPS, bool AMER, bool TBL, int INJ>
__device__ __forceinline__ void step_a(double (&V)[C], double (&L)[C], const Ph &ph, const Sm &sm, const int k0,
                                       const double cin_f, const double carryF, double &e1_keep, double &Eprev)
{
    constexpr int H = C / 2;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const double *nm = ph.slot;
    const double *lc = ph.slot + 3 * C * 32;
    if (AMER) {
        // state (vc, L = 2 dt lambda): subtrahend G = q V + dt lambda = q vc + (q + 1) / 2 L, scaled rhs kappa vc + L
        if (ph.q == 1.0) {
#pragma unroll
            for (int k = 0; k < C; ++k) { const double g = V[k] + L[k]; V[k] = fma(ph.pd, V[k], L[k]); L[k] = g; }
        } else {
            const double hq = 0.5 * (ph.q + 1.0);
#pragma unroll
            for (int k = 0; k < C; ++k) { const double g = fma(ph.q, V[k], hq * L[k]); V[k] = fma(ph.pd, V[k], L[k]); L[k] = g; }
        }
    } else if (!lin_light<C, WARPS, AMER>()) {
        if (ph.q == 1.0) {                                 // theta = 1/2 (CTA-uniform): the old V itself, a register rename
#pragma unroll
            for (int k = 0; k < C; ++k) L[k] = V[k];
        } else {
#pragma unroll
            for (int k = 0; k < C; ++k) L[k] = ph.q * V[k];
        }
    }
    // (lin_light: the linear march keeps no subtrahend q V_old, the last pass rebuilds V_old from u, see lin_vc)
    if (TBL && k0 >= 0) {
#pragma unroll
        for (int k = 0; k < C; ++k) if (k == k0) V[k] += cin_f;
    }
    if (INJ == INJ_SLOT || (INJ == INJ_RUNTIME && ph.inj)) {
        // predicated adds (k0 = -1 off the owner thread), in PTX: from `if (k == k0) V[k] += cin` the compiler builds a
        // tree of indirect branches through constant-memory jump tables on the one warp every other warp waits for
#pragma unroll
        for (int k = 0; k < C; ++k)
            asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\t@p add.rn.f64 %0, %0, %3;\n\t}"
                : "+d"(V[k]) : "r"(k0), "r"(k), "d"(cin_f));
    }
    double e1 = 0.0, e2 = 0.0;
#pragma unroll
    for (int k = 0; k < H; ++k) e1 = fma(TBL ? nm[k * 32 + lane] : ph.NM, e1, V[k]);
#pragma unroll
    for (int k = H; k < C; ++k) e2 = fma(TBL ? nm[k * 32 + lane] : ph.NM, e2, V[k]);
    double e_end = fma(TBL ? lc[LC_HF2 * 32 + lane] : sm.cst[CST_HF2], e1, e2);
    if (INJ == INJ_CARRY0) e_end = fma(sm.cst[CST_AF], carryF, e_end);
    if (INJ == INJ_CARRYH) e_end = fma(sm.cst[CST_HF2], carryF, e_end);
    if (INJ == INJ_RUNTIME) e_end = fma(sm.smv[5], carryF, e_end);               // Dirichlet term of row 0 as a carry into its owner
    const double E = (!TBL && SCAN_TABLE) ? warp_scan_tab<true, scan_level_skip<C, WARPS, AMER>()>(e_end, sm.scw, lane, sm.ish[14])
                                          : warp_scan<TBL, true>(e_end, TBL ? lc[LC_AF * 32 + lane] : sm.cst[CST_AF], lane);
    if (lane == 31) sm.part[0 * WARPS + w] = E;           // pE_f
    const double up = __shfl_up_sync(FULL, E, 1);
    Eprev = (lane == 0) ? 0.0 : up;
    e1_keep = e1;
}

template <int C, int WARPS, bool AMER, bool TBL, int INJ, bool K1>
__device__ __forceinline__ void step_b(double (&V)[C], const Ph &ph, const Sm &sm, const int wf, const double cin_b,
                                       const double e1_fwd, const double Eprev, const double carryF, double &e1_keep,
                                       double &Enext, double &eend_keep, double &csf_keep, double &cmid_keep)
{
    constexpr int H = C / 2;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const double *nm = ph.slot, *ncc = ph.slot + C * 32;
    const double *lc = ph.slot + 3 * C * 32;
    // carry into this warp from the warps below (truncated Horner sum)
    double S = 0.0;
    if (K1) { if (w > wf) S = sm.part[0 * WARPS + w - 1]; }       // one term: the total of the warp below
    else {
#pragma unroll 1
        for (int v = max(wf, w - ph.Kf); v < w; ++v) S = fma(sm.part[1 * WARPS + v], S, sm.part[0 * WARPS + v]);
    }
    // the owner's Dirichlet carry enters below slot 0 or below slot C/2; it is known before the barrier, so its sums
    // with Eprev and e1_fwd are off the path that waits for S
    double ep = Eprev, e1x = e1_fwd;
    if (INJ == INJ_CARRYH) e1x = e1_fwd + carryF;
    if (INJ == INJ_RUNTIME) { ep = fma(sm.smv[3], carryF, Eprev); e1x = fma(sm.smv[4], carryF, e1_fwd); }
    double cs = fma(TBL ? lc[LC_PAF * 32 + lane] : sm.PAc[lane], S, ep);
    if (INJ == INJ_CARRY0) cs += carryF;
    csf_keep = cs;
    {
        double c1 = cs;
        double c2 = fma(TBL ? lc[LC_HF1 * 32 + lane] : sm.cst[CST_HF1], cs, e1x);
        cmid_keep = c2;                                     // what slot C/2 sees below it (the owner's carry included)
#pragma unroll
        for (int k = 0; k < H; ++k) { c1 = fma(TBL ? nm[k * 32 + lane] : ph.NM, c1, V[k]); V[k] = c1; }
#pragma unroll
        for (int k = H; k < C; ++k) { c2 = fma(TBL ? nm[k * 32 + lane] : ph.NM, c2, V[k]); V[k] = c2; }
    }
    // back substitution, local pass: upper half (k = C-1..H) then lower half (H-1..0)
    double e1 = cin_b, e2 = 0.0;
#pragma unroll
    for (int k = C - 1; k >= H; --k) e1 = fma(TBL ? ncc[k * 32 + lane] : ph.NCC, e1, V[k]);
#pragma unroll
    for (int k = H - 1; k >= 0; --k) e2 = fma(TBL ? ncc[k * 32 + lane] : ph.NCC, e2, V[k]);
    const double e_end = fma(TBL ? lc[LC_HB2 * 32 + lane] : sm.cst[CST_HB2], e1, e2);
    const double E = (!TBL && SCAN_TABLE) ? warp_scan_tab<false, scan_level_skip<C, WARPS, AMER>()>(e_end, sm.scw + 160, lane, sm.ish[15])
                                          : warp_scan<TBL, false>(e_end, TBL ? lc[LC_AB * 32 + lane] : sm.cst[CST_AB], lane);
    if (lane == 0) sm.part[2 * WARPS + w] = E;            // pE_b
    const double dn = __shfl_down_sync(FULL, E, 1);
    Enext = (lane == 31) ? 0.0 : dn;
    e1_keep = e1;
    eend_keep = e_end;
}

// Linear march, last pass.  The new value is vc = INV xs - q V_old, and V_old is not kept: the system is solved for
// R = V_old (setup_phase scales it), so V_old = u_k - nm_k u_{k-1} with u still in the registers; below slot 0 and below
// slot C/2, u_{k-1} is the value the half-chain was started with (cs_f, c_mid: the forward carry into the thread, the
// owner's Dirichlet carry included -- not V[C/2 - 1], which lacks that carry when row 0 sits at slot C/2).  One FP64 instruction more per node than with a copy of V_old, but half the register file:
// 4 / 4 / 3 resident CTAs of 4 / 5 / 8 warps instead of 4 / 3 / 2, and no 2 C register moves per step.
//   u_k, um: u of this slot and of the slot below;  returns vc.
template <bool Q1>
__device__ __forceinline__ double lin_vc(const double inv, const double xs, const double nm, const double um, const double u_k,
                                         const double q)
{
    const double t = fma(nm, um, -u_k);                      // -V_old
    return Q1 ? fma(inv, xs, t) : fma(inv, xs, q * t);
}

template <int C, int WARPS, bool AMER, bool TBL, bool K1>
__device__ __forceinline__ void step_c(double (&V)[C], double (&L)[C], const Ph &ph, const Sm &sm, const double cin_b,
                                       const double e1_bwd, const double Enext, const double cs_f, const double c_mid,
                                       const int k0, const double cin_f)
{
    constexpr int H = C / 2;
    constexpr int NT = 32 * WARPS;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const double *ncc = ph.slot + C * 32, *inv = ph.slot + 2 * C * 32;
    const double *lc = ph.slot + 3 * C * 32;
    double S = 0.0;
    if (K1) { if (w < WARPS - 1) S = sm.part[2 * WARPS + w + 1]; }
    else {
#pragma unroll 1
        for (int v = min(WARPS - 1, w + ph.Kb); v > w; --v) S = fma(sm.part[3 * WARPS + v], S, sm.part[2 * WARPS + v]);
    }
    const double cs = fma(TBL ? lc[LC_PAB * 32 + lane] : sm.PAc[32 + lane], S, Enext);
    double c1 = cs + cin_b;
    double c2 = fma(TBL ? lc[LC_HB1 * 32 + lane] : sm.cst[CST_HB1], cs, e1_bwd);
    const double *nm = ph.slot;
    constexpr bool LIGHT = lin_light<C, WARPS, AMER>();
    auto slots = [&](auto q1_tag) {
    constexpr bool Q1 = decltype(q1_tag)::value != 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int k = half == 0 ? C - 1 - i : H - 1 - i;
            double xs;
            if (half == 0) { c1 = fma(TBL ? ncc[k * 32 + lane] : ph.NCC, c1, V[k]); xs = c1; }
            else           { c2 = fma(TBL ? ncc[k * 32 + lane] : ph.NCC, c2, V[k]); xs = c2; }
            double vc;
            if (!LIGHT) vc = fma(TBL ? inv[k * 32 + lane] : ph.INV, xs, -L[k]);
            else {
                vc = lin_vc<Q1>(TBL ? inv[k * 32 + lane] : ph.INV, xs, TBL ? nm[k * 32 + lane] : ph.NM,
                                k == 0 ? cs_f : (k == H ? c_mid : V[k - 1]), V[k], ph.q);
                // a table warp adds the Dirichlet term of row 0 to R at its slot: R = V_old + cin there
                if (TBL) asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\t@p fma.rn.f64 %0, %3, %4, %0;\n\t}"
                             : "+d"(vc) : "r"(k0), "r"(k), "d"(ph.q), "d"(cin_f));
            }
            if (AMER) {
                const double d = sm.r0[k * NT + t] - vc;         // exercise iff phi > vc (fd_american_option_pricer.py:496)
                L[k] = d + fabs(d);                              // 2 dt lambda_new, ONE FP64 instruction
                V[k] = vc;
            } else {
                V[k] = vc;
            }
        }
    }
    };
    if (!LIGHT || ph.q == 1.0) slots(IntC<1>{}); else slots(IntC<0>{});   // theta = 1/2 but for the Rannacher steps
}

// step_c of a scalar-coefficient warp with the backward carry `cs` already summed (bwd_carry).  CORR: the row-0
// correction reaches this warp: xs'_k = xs_k - gz xs_0 NM^k (one extra FMA per slot, off the recurrence chain).
// INVT: the warp of row 0 (pivot scaling by slot, zero on the dead slots).
template <int WARPS, bool K1>
__device__ __forceinline__ double bwd_carry(const Ph &ph, const Sm &sm, const double Enext)
{
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    double S = 0.0;
    if (K1) { if (w < WARPS - 1) S = sm.part[2 * WARPS + w + 1]; }
    else {
#pragma unroll 1
        for (int v = min(WARPS - 1, w + ph.Kb); v > w; --v) S = fma(sm.part[3 * WARPS + v], S, sm.part[2 * WARPS + v]);
    }
    return fma(sm.PAc[32 + lane], S, Enext);
}

// SLOTINJ: this warp added the Dirichlet term of row 0 to R at its slot (R = V_old + cin there).
template <int C, int WARPS, bool AMER, bool CORR, bool INVT, bool SLOTINJ>
__device__ __forceinline__ void step_cs(double (&V)[C], double (&L)[C], const Ph &ph, const Sm &sm, const double cs,
                                        const double cin_b, const double e1_bwd, const double cs_f, const double c_mid,
                                        const int k0, const double cin_f, const double xs_row0)
{
    constexpr int H = C / 2;
    constexpr int NT = 32 * WARPS;
    const int t = threadIdx.x;
    double c1 = cs + cin_b;
    double c2 = fma(sm.cst[CST_HB1], cs, e1_bwd);
    const double gzx = CORR ? -(ph.gz * xs_row0) : 0.0;
    constexpr bool LIGHT = lin_light<C, WARPS, AMER>();
    // Slots in the order the two half-chains visit them; the payoff and the correction's NM^k are fetched PF slots
    // ahead.  (Left to the compiler, each load sat right in front of its use at the 128-register cap: the owner warp of
    // row 0, the one with the correction, spent 890 cycles in this pass on 2133-node barrier grids against 270 for the
    // other warps -- and every other warp of the CTA waits for it at the next barrier.)
    auto kseq = [](const int j) -> int { return (j & 1) == 0 ? C - 1 - (j >> 1) : H - 1 - (j >> 1); };
    constexpr int PF = 4;
    auto slots = [&](auto q1_tag) {
    constexpr bool Q1 = decltype(q1_tag)::value != 0;
    double pphi[AMER ? C + PF : 1], pnm[CORR ? C + PF : 1];
#pragma unroll
    for (int j = 0; j < PF; ++j) {
        if (AMER) pphi[j] = sm.r0[kseq(j) * NT + t];
        if (CORR) pnm[j] = sm.nmp[kseq(j)];
    }
#pragma unroll
    for (int j = 0; j < C; ++j) {
        const int k = kseq(j);
        if (j + PF < C) {
            if (AMER) pphi[j + PF] = sm.r0[kseq(j + PF) * NT + t];
            if (CORR) pnm[j + PF] = sm.nmp[kseq(j + PF)];
        }
        double xs;
        if ((j & 1) == 0) { c1 = fma(ph.NCC, c1, V[k]); xs = c1; }
        else              { c2 = fma(ph.NCC, c2, V[k]); xs = c2; }
        if (CORR) xs = fma(gzx, pnm[j], xs);
        // the warp of row 0: zero pivot scaling on the dead slots below row 0 (ph.INV is 0 on its dead lanes already)
        const bool live = !INVT || ph.e0 + k >= 0;
        const double iv = live ? ph.INV : 0.0;
        double vc;
        if (!LIGHT) vc = fma(iv, xs, -L[k]);
        else {
            vc = lin_vc<Q1>(iv, xs, ph.NM, k == 0 ? cs_f : (k == H ? c_mid : V[k - 1]), V[k], ph.q);
            if (SLOTINJ) asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\t@p fma.rn.f64 %0, %3, %4, %0;\n\t}"
                             : "+d"(vc) : "r"(k0), "r"(k), "d"(ph.q), "d"(cin_f));
            // dead slots stay at exactly zero: their u is not (carry form), and the rebuilt V_old there is rounding
            // noise of it
            if (INVT) vc = live ? vc : 0.0;
        }
        if (AMER) {
            const double d = pphi[j] - vc;                       // exercise iff phi > vc (fd_american_option_pricer.py:496)
            L[k] = d + fabs(d);                                  // 2 dt lambda_new, ONE FP64 instruction
            V[k] = vc;
        } else {
            V[k] = vc;
        }
    }
    };
    if (!LIGHT || ph.q == 1.0) slots(IntC<1>{}); else slots(IntC<0>{});   // theta = 1/2 but for the Rannacher steps
}

// ---- dividend jump (fd_american_option_pricer.py:321-386, :517-548) ------------------------------
// Where node j lives in the jump's shared arrays.  Thread t works on C consecutive nodes, so in the plain [j] layout
// the lanes of a warp are C doubles apart: at C = 16 all 32 lanes hit one pair of banks and every access took 32
// wavefronts -- the jump was bound by exactly that (8192 nodes: 237k cycles per jump, ~60 % of it bank conflicts).
// XOR-ing the low four bits with the next four spreads a warp's access over all banks within the same 16-node block.
__device__ __forceinline__ int jsw(const int j) { return j ^ ((j >> 4) & 15); }

// ys, mu, cz: shared arrays of n+1 doubles in jsw() order.
// Works on shared memory only (value vector in ys[], result returned in mu[]) and is not inlined, so its
// registers do not weigh on the time loop.  Returns the new boundary node value of this thread.
template <int C, int WARPS>
__device__ __noinline__ double dividend_jump(double Vb, double cash, bool is_call, double k_eff, int n, int pad,
                                             double x_min, double dx, double *ys, double *mu, double *cz)
{
    double V[C];
    constexpr int NT = 32 * WARPS;
    const int t = threadIdx.x;
    const int nn = n + 1;
    const bool owner0 = (t == pad / C);
#ifdef FDM_CTA_PROFILE
    long long jt[6]; int ji = 0;
#define JTICK() jt[ji++] = clock64()
#else
#define JTICK()
#endif
    JTICK();
    if (owner0) ys[jsw(0)] = Vb;
    if (t == NT - 1) ys[jsw(n)] = Vb;
    __syncthreads();
    // The spline's tridiagonal system (:338-359) is diagonally dominant: the forward recurrences forget
    // their start like 0.27^k and so does the back substitution.  Every thread therefore warm-starts
    // G rows outside its own rows (exactly at the true ends) and keeps only its own rows: all rows are
    // eliminated in parallel, in the reference's operation order, to within 0.27^G of the serial sweep.
    // Per row: the knots follow s_{i+1} = s_i e^{dx} from an exact start (an exponential per row before), one
    // reciprocal serves the two quotients by l and 3 / h_{i-1} is the previous row's 3 / h_i (four IEEE divisions per
    // row before); G = 30 warm-up rows leave 0.27^30 = 1e-17 of a wrong start.  237k -> see DESIGN.md cycles per jump
    // at 8192 nodes; the quotients differ from the reference's in the last bit, the jumped values by ~1e-14 relative.
    constexpr int G = 30;
    const int j_first = max(t * C - pad + 1, 1);           // this thread's interior knots [j_first, j_last]
    const int j_last = t * C + C - pad;
    const bool has_rows = j_last >= 1;
    if (has_rows) {
        const int i0 = max(1, j_first - G);
        const double edx = exp(dx);
        double x_im1 = node_s(x_min, dx, i0 - 1), x_i = node_s(x_min, dx, i0);
        double mu_prev = 0.0, z_prev = 0.0;
        double y_im1 = ys[jsw(i0 - 1)], y_i = ys[jsw(i0)];
        double r3_prev = 3.0 * __drcp_rn(__dsub_rn(x_i, x_im1));
#pragma unroll 2
        for (int i = i0; i <= j_last; ++i) {
            const double x_ip1 = x_i * edx;
            const double y_ip1 = ys[jsw(i + 1)];
            const double h_i = __dsub_rn(x_ip1, x_i);
            const double h_im1 = __dsub_rn(x_i, x_im1);
            const double r3 = 3.0 * __drcp_rn(h_i);
            const double alpha = __dsub_rn(__dmul_rn(r3, __dsub_rn(y_ip1, y_i)), __dmul_rn(r3_prev, __dsub_rn(y_i, y_im1)));
            const double l = __dsub_rn(__dmul_rn(2.0, __dsub_rn(x_ip1, x_im1)), __dmul_rn(h_im1, mu_prev));
            const double rl = __drcp_rn(l);
            mu_prev = h_i * rl;
            z_prev = __dsub_rn(alpha, __dmul_rn(h_im1, z_prev)) * rl;
            if (i >= j_first) { mu[jsw(i)] = mu_prev; cz[jsw(i)] = z_prev; }
            x_im1 = x_i; x_i = x_ip1;
            y_im1 = y_i; y_i = y_ip1;
            r3_prev = r3;
        }
    }
    if (owner0) { mu[jsw(0)] = 0.0; cz[jsw(0)] = 0.0; }
    if (t == NT - 1) cz[jsw(nn - 1)] = 0.0;
    JTICK();
    __syncthreads();
    JTICK();
    // back substitution c[j] = z[j] - mu[j] c[j+1] (:358-359); results parked in V until every thread has read z
    if (has_rows) {
        const int i1 = min(n - 1, j_last + G);
        double c_next = 0.0;
        for (int j = i1; j > j_last; --j) c_next = __dsub_rn(cz[jsw(j)], __dmul_rn(mu[jsw(j)], c_next));
#pragma unroll
        for (int k = C - 1; k >= 0; --k) {
            const int j = t * C + k - pad + 1;
            if (j >= 1) { c_next = __dsub_rn(cz[jsw(j)], __dmul_rn(mu[jsw(j)], c_next)); V[k] = c_next; }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = t * C + k - pad + 1;
        if (j >= 1) cz[jsw(j)] = V[k];
    }
    __syncthreads();
    JTICK();
    const double s_lo = node_s(x_min, dx, 0), s_hi = node_s(x_min, dx, n);
    const double edx_eval = exp(dx);
    // A thread's nodes are consecutive, so S_j - cash rises from slot to slot and so does its bracket: the first node
    // of a thread takes the node value and the bracket from the exponential / the logarithm, the following ones walk
    // (S_j and the bracket's knots by multiplication with e^dx: a few ulp of drift over a thread's C nodes, against
    // four exponentials, a logarithm and three divisions per node before; the spline is C2 across a knot).
    int br_j = -1;
    double br_x = 0.0, br_x1 = 0.0;
    auto eval = [&](const double Sj, const bool walk) -> double {
        const double Sq = __dsub_rn(Sj, cash);
        double cont;
        if (Sq <= s_lo) cont = ys[jsw(0)];
        else if (Sq >= s_hi) cont = ys[jsw(n)];
        else {
            if (!walk || br_j < 0) {
                int j = last_node_le(x_min, dx, n, Sq);       // == the reference's binary search (:375-382)
                br_j = min(max(j, 0), nn - 2);
                br_x = node_s(x_min, dx, br_j);
                br_x1 = br_x * edx_eval;
            } else {
                while (br_j < nn - 2 && br_x1 <= Sq) { ++br_j; br_x = br_x1; br_x1 *= edx_eval; }
            }
            const int j = br_j;
            // (one reciprocal: the quotients differ from the reference's in the last bit)
            const double xj = br_x, xj1 = br_x1;
            const double h = __dsub_rn(xj1, xj), rh = __drcp_rn(h);
            const double cj = cz[jsw(j)], cj1 = cz[jsw(j + 1)];
            const double bj = __dsub_rn(__dmul_rn(__dsub_rn(ys[jsw(j + 1)], ys[jsw(j)]), rh),
                                        __dmul_rn(__dmul_rn(h, __dadd_rn(cj1, __dmul_rn(2.0, cj))), 1.0 / 3.0));
            const double dj = __dmul_rn(__dmul_rn(__dsub_rn(cj1, cj), rh), 1.0 / 3.0);
            const double dq = __dsub_rn(Sq, xj);
            double acc = __dadd_rn(ys[jsw(j)], __dmul_rn(bj, dq));
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(cj, dq), dq));
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(dj, dq), dq), dq));
            cont = acc;
        }
        if (is_call) {
            const double ph = payoff(Sj, k_eff, true);
            cont = cont > ph ? cont : ph;                        // :542-544
        }
        return cont;
    };
    // mu[] is free once cz[] holds the spline's c coefficients: the jumped values go there
    {
        double Sj = 0.0;
        bool first = true;
#pragma unroll 1
        for (int k = 0; k < C; ++k) {
            const int j = t * C + k - pad + 1;
            if (j < 1) continue;
            Sj = first ? node_s(x_min, dx, j) : Sj * edx_eval;
            mu[jsw(j)] = eval(Sj, !first);
            first = false;
        }
    }
    double nb = Vb;
    if (owner0) nb = eval(s_lo, false);
    if (t == NT - 1) nb = eval(s_hi, false);
    JTICK();
    __syncthreads();
    JTICK();
#ifdef FDM_CTA_PROFILE
    if (blockIdx.x == 40 && (threadIdx.x == 32 * (WARPS - 1) || threadIdx.x == 32 * (WARPS / 2) + 5))
        printf("jump t=%d: elimination %lld | wait %lld | back-subst+store %lld | eval %lld | wait %lld\n", t, jt[1] - jt[0], jt[2] - jt[1],
               jt[3] - jt[2], jt[4] - jt[3], jt[5] - jt[4]);
#endif
    return nb;
}

// Read-out by ONE thread from the value vector staged by node (fd_american_option_pricer.py:613-653 with snapping;
// discrete_barrier_fdm_pricer_cn.py:308-353 without): price, stencil, delta, gamma, indices, status.
static __device__ __noinline__ void read_out(const FdmOut &O, const int64_t sid, const double *Vs, const int n, const bool snap,
                                      const double x_min, const double dx, const double spot, const int spot_idx,
                                      const int strike_idx, const int ex_lo, const int ex_hi, const int flagged)
{
    const double sn = node_s(x_min, dx, n);
    double price, delta, gamma, sv[4], ss[4];
    int idx_a, idx_b;
    if (snap) {
        const int i_near = min(max(spot_idx, 1), n - 2);
        const double Sx = node_s(x_min, dx, spot_idx);
        for (int q = 0; q < 4; ++q) { sv[q] = Vs[i_near - 1 + q]; ss[q] = node_s(x_min, dx, i_near - 1 + q); }
        price = Vs[spot_idx];
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
        const double s0 = node_s(x_min, dx, 0);
        int lo = 0;
        if (spot <= s0) { price = Vs[0]; lo = 0; }
        else if (spot >= sn) { price = Vs[n]; lo = n; }
        else {
            lo = min(max(last_node_le(x_min, dx, n, spot), 0), n - 1);
            const double s_lo = node_s(x_min, dx, lo), s_hi = node_s(x_min, dx, lo + 1);
            const double wgt = __dsub_rn(spot, s_lo) / __dsub_rn(s_hi, s_lo);
            price = __dadd_rn(__dmul_rn(__dsub_rn(1.0, wgt), Vs[lo]), __dmul_rn(wgt, Vs[lo + 1]));
        }
        const int gi = nearest_node(x_min, dx, n, spot, true);
        for (int q = 0; q < 3; ++q) { sv[q] = Vs[gi - 1 + q]; ss[q] = node_s(x_min, dx, gi - 1 + q); }
        sv[3] = 0.0; ss[3] = 0.0;
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
    O.price[sid] = price;
    if (O.delta) O.delta[sid] = delta;
    if (O.gamma) O.gamma[sid] = gamma;
    if (O.stencil_v) for (int q = 0; q < 4; ++q) O.stencil_v[4 * sid + q] = sv[q];
    if (O.stencil_s) for (int q = 0; q < 4; ++q) O.stencil_s[4 * sid + q] = ss[q];
    if (O.idx) {
        O.idx[4 * sid + 0] = idx_a;
        O.idx[4 * sid + 1] = idx_b;
        O.idx[4 * sid + 2] = ex_lo;
        O.idx[4 * sid + 3] = ex_hi;
    }
    if (O.status) O.status[sid] = flagged ? flagged : (isfinite(price) ? 0 : 1);
}

// ---- the kernel ------------------------------------------------------------------------------------

template <int C, int WARPS>
constexpr int smem_doubles(const bool american) { return american ? Layout<C, WARPS>::TOTAL : Layout<C, WARPS>::TOTAL_LINEAR; }

// resident CTAs per SM the registers are sized for.  More, smaller-budget CTAs beat fewer fat ones: every CTA is
// latency-bound between its two block barriers, and independent CTAs fill each other's gaps (2133-node barrier
// grids, 5 warps: 2 CTAs x 168 registers 6.4e11 -> 3 CTAs x 128 registers 7.2e11 node-steps/s).
#ifndef FDM_CTA4_BLOCKS
#define FDM_CTA4_BLOCKS 4
#endif
#ifndef FDM_CTA5_BLOCKS
#define FDM_CTA5_BLOCKS 3
#endif
#ifndef FDM_CTA8_BLOCKS
#define FDM_CTA8_BLOCKS 2
#endif
// The linear march keeps one register array (V), not two: a resident CTA more in the 5- and 8-warp classes.
#ifndef FDM_CTA5_BLOCKS_LINEAR
#define FDM_CTA5_BLOCKS_LINEAR 4
#endif
#ifndef FDM_CTA8_BLOCKS_LINEAR
#define FDM_CTA8_BLOCKS_LINEAR 3
#endif
template <int C, int WARPS, bool AMER>
constexpr int cta_min_blocks()
{
    constexpr bool light = lin_light<C, WARPS, AMER>();
    return WARPS == 4 ? FDM_CTA4_BLOCKS : (WARPS == 5 ? (light ? FDM_CTA5_BLOCKS_LINEAR : FDM_CTA5_BLOCKS)
                      : (WARPS == 8 ? (light ? FDM_CTA8_BLOCKS_LINEAR : FDM_CTA8_BLOCKS) : 1));
}

template <int C, int WARPS, bool AMER, bool TRACE>
__global__ void __launch_bounds__(32 * WARPS, cta_min_blocks<C, WARPS, AMER>())
fdm_cta_kernel(const FdmBatch B, const FdmOut O, const int64_t first, const int want_american, const int n_lo)
{
    using LY = Layout<C, WARPS>;
    constexpr int NT = LY::NT;
    constexpr int NP = LY::NP;
    extern __shared__ double smem[];
    Sm sm;
    sm.r0 = smem;                                       // American only
    sm.r1 = sm.r0 + (AMER ? NP : 0);
    sm.r2 = sm.r1 + NP;
    sm.part = sm.r2 + LY::R2;
    sm.cst = sm.part + 4 * WARPS;
    sm.PAc = sm.cst + 16;
    sm.wb = sm.PAc + 64;
    sm.ish = (int *)(sm.wb + 32);
    sm.nmp = sm.wb + 32 + 8;
    sm.smv = sm.nmp + C;
    sm.scw = sm.smv + 8;

    const int64_t sid = first + blockIdx.x;
    if (sid >= B.n_solves) return;
    const int flags = B.flags[sid];
    if (((flags & FDM_FLAG_AMERICAN) != 0) != (want_american != 0)) return;
    if (flags & FDM_FLAG_INVALID) return;                       // rejected by the book fan-out: status stays -1
    const int n = B.n_space[sid];
    if (n <= n_lo || n - 1 > NT * C) return;           // other size classes are other launches
    const int n_int = n - 1;
    const int pad = NT * C - n_int;                    // slots below row 0

    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const bool is_call = flags & FDM_FLAG_CALL;
    const bool b76 = flags & FDM_FLAG_BLACK76;
    const bool snap = flags & FDM_FLAG_SNAP;
    const bool ko = flags & (FDM_FLAG_KO_DOWN | FDM_FLAG_KO_UP);
    const double x_min = B.x_min[sid], dx = B.dx[sid], sigma = B.sigma[sid], r = B.r[sid], bcar = B.b[sid];
    const double spot = B.spot[sid], strike = B.strike[sid];
    const double rebate = (ko && B.rebate) ? B.rebate[sid] : 0.0;
    const int rannacher = B.rannacher[sid];
    const int t0 = pad / C;                            // thread that owns row 0 (and boundary node 0)
    const int k0 = (t == t0) ? pad % C : -1;

    if (t == 0) {
        int spot_idx = 0, strike_idx = 0, ka, kb;
        if (snap) {
            spot_idx = nearest_node(x_min, dx, n, spot, false);
            strike_idx = nearest_node(x_min, dx, n, strike, false);
        }
        ko_thresholds(B, sid, flags, x_min, dx, n, ka, kb);
        sm.ish[1] = spot_idx; sm.ish[2] = strike_idx; sm.ish[3] = ka; sm.ish[4] = kb;
        sm.ish[5] = 0; sm.ish[6] = 0x7fffffff; sm.ish[7] = -1;
    }
    __syncthreads();
    const int spot_idx = sm.ish[1], strike_idx = sm.ish[2], ko_a = sm.ish[3], ko_b = sm.ish[4];
    const double k_eff = snap ? node_s(x_min, dx, strike_idx) : strike;
    const double sn = node_s(x_min, dx, n);
    const bool put_smin = flags & FDM_FLAG_PUT_BDRY_SMIN;
    const double sn_bdry = (flags & FDM_FLAG_BDRY_TOP_NEXT) ? node_s(x_min, dx, n + 1) : sn;
    const double b_bdry = B.b_bdry ? B.b_bdry[sid] : bcar;

    double V[C], L[C];
    auto fill_payoff = [&](bool set_v) {
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int j = t * C + k - pad + 1;
            const double pv = (j >= 1) ? payoff(node_s(x_min, dx, j), k_eff, is_call) : 0.0;
            if (AMER) sm.r0[k * NT + t] = pv;
            if (set_v) { V[k] = pv; L[k] = 0.0; }
        }
    };
    fill_payoff(true);
    // boundary node values: V[0] on the owner of row 0, V[n] on the last thread
    double Vb = (t == t0) ? payoff(node_s(x_min, dx, 0), k_eff, is_call)
                          : ((t == NT - 1) ? payoff(sn, k_eff, is_call) : 0.0);
    const bool bdry_thread = is_call ? (t == NT - 1) : (t == t0);
    const bool bdry_warp = is_call ? (w == WARPS - 1) : (w == t0 / 32);

    const double sig2 = __dmul_rn(sigma, sigma);
    const double mu_x = b76 ? __dmul_rn(-0.5, sig2) : __dsub_rn(bcar, __dmul_rn(0.5, sig2));
    const double alpha = __dmul_rn(0.5, sig2) / __dmul_rn(dx, dx);
    const double beta = mu_x / __dmul_rn(2.0, dx);
    const double ca = __dsub_rn(alpha, beta);
    const double cc = __dadd_rn(alpha, beta);
    const double bcoef = __dsub_rn(__dmul_rn(-2.0, alpha), r);

    // -DFDM_CTA_PROFILE: per-region cycle counts of one warp (dev tool; printf from block 40)
#ifdef FDM_CTA_PROFILE
    long long prof_acc[3] = {0, 0, 0};
    long long prof_t = clock64();
    long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ph_t = 0;            // step_a | barrier | step_b | barrier | step_c
#define PH_TICK(slot) do { const long long t_ = clock64(); ph_acc[slot] += t_ - ph_t; ph_t = t_; } while (0)
#define PH_START() ph_t = clock64()
#define CTA_TICK(slot) do { const long long t_ = clock64(); prof_acc[slot] += t_ - prof_t; prof_t = t_; } while (0)
#else
#define CTA_TICK(slot)
#define PH_TICK(slot)
#define PH_START()
#endif
    const int seg_b = B.seg_begin[sid], seg_e = B.seg_begin[sid + 1];
    if (!AMER) {
        // The linear kernel has no shared memory for the spline of a dividend jump (what it saves buys a resident CTA
        // more per SM).  No pricer of the reference marches a European / barrier solve through cash dividends -- they
        // are escrowed -- so such a solve, if a caller of the C ABI builds one, is left to fdm_exact_kernel.
        bool has_div = false;
        for (int seg = seg_b; seg < seg_e; ++seg) has_div |= B.seg_div[seg] != 0.0;
        if (has_div) {                                          // CTA-uniform
            if (t == 0) {
                if (O.status) O.status[sid] = FDM_STATUS_STIFF;
                else if (O.price) O.price[sid] = __longlong_as_double(0x7ff8000000000000ll);   // nobody to hand it to
            }
            return;
        }
    }
    int mon_i = 0, mon_e = 0;
    if (ko && B.mon_begin) { mon_i = B.mon_begin[sid]; mon_e = B.mon_end ? B.mon_end[sid] : B.mon_begin[sid + 1]; }
    int next_mon = (mon_i < mon_e) ? B.mon_step[mon_i] : -1;
    int gstep = 0;
    int32_t *trace = (TRACE && O.ex_trace) ? O.ex_trace + sid * O.ex_trace_stride : nullptr;

    for (int seg = seg_b; seg < seg_e; ++seg) {
        const int n_steps = B.seg_steps[seg];
        const double dt = B.seg_dt[seg];
        const double tau0 = B.seg_tau0[seg];
        const int base_ran = B.seg_restart[seg] ? rannacher : 0;
        const double cash = B.seg_div[seg];
        if (AMER) {
#pragma unroll
            for (int k = 0; k < C; ++k) L[k] = 0.0;             // :463
        }
        int m = 0;
        int fill_base = -64;
        // moving Dirichlet value: 32 steps at a time, one exponential per lane of the warp that owns the
        // boundary, parked in shared memory (only that warp touches wb[])
        auto fetch_boundary = [&](int step) -> double {
            if (!bdry_warp) return 0.0;
            if (step - fill_base >= 32 || step < fill_base) {
                fill_base = step;
                const double tau = fma((double)(step + lane + 1), dt, tau0);
                const double wv = moving_boundary_value(is_call, b76, r, b_bdry, k_eff, sn_bdry, tau, put_smin, node_s(x_min, dx, 0));
                __syncwarp();
                sm.wb[lane] = wv;
                __syncwarp();
            }
            return bdry_thread ? sm.wb[step - fill_base] : 0.0;
        };
        while (m < n_steps) {
            const bool implicit_phase = m < base_ran;
            const int m_end = implicit_phase ? min(base_ran, n_steps) : n_steps;
            Ph ph;
            CTA_TICK(0);
            setup_phase<C, WARPS, AMER>(implicit_phase ? 1.0 : 0.5, dt, ca, cc, bcoef, n_int, pad, sm, ph);
            CTA_TICK(1);
            const int wf = sm.ish[10];
            constexpr bool R0 = row0_mode<C, WARPS, AMER>();
            const bool row0_corr = R0 && sm.ish[12] >= 0;       // CTA-uniform
            const double top_mult = (t == NT - 1) ? sm.cst[CST_TOP] : 0.0;
            double w_next = fetch_boundary(m);
            double cin = ph.cB * Vb - ph.cA * w_next;           // non-zero on the two boundary threads only
            // The time loop is compiled once per warp role (warp-uniform, fixed for the phase), so a step has no
            // branches on the role: between two block barriers each warp runs straight-line code.  Every role executes
            // the same sequence of block barriers (different BAR.SYNC sites on barrier 0, as in any warp-specialised
            // kernel).  Role-less form: FDM_CTA_ONE_LOOP.
            auto time_loop = [&](auto role_tag, auto k1_tag) {
            constexpr int ROLE = decltype(role_tag)::value;
            constexpr bool K1 = decltype(k1_tag)::value != 0;   // warp-to-warp sums of one term (the usual case)
            for (; m < m_end; ++m) {
                const double w_now = w_next;
                w_next = fetch_boundary(m + 1);
                double cin_n = ph.cB * w_now - ph.cA * w_next;
                const double cin_b = cin * top_mult;            // enters the top row through its ncc
                double e1f = 0.0, Eprev = 0.0, e1b = 0.0, Enext = 0.0, eend_b = 0.0, csf = 0.0, cmid = 0.0;
                constexpr bool ANY = ROLE == ROLE_ANY;
                const bool tbl = ANY ? ph.cls == 2 : ROLE == ROLE_TBL;
                const bool scl = ANY ? ph.cls == 1 : (ROLE != ROLE_TBL && ROLE != ROLE_PAD);
                // the scalar warp that holds row 0 (row-0 correction mode) and how the Dirichlet term of row 0 enters it
                constexpr int INJ = !R0 ? INJ_NONE : (ANY ? INJ_RUNTIME : (ROLE == ROLE_OWNER ? INJ_CARRY0
                                    : (ROLE == ROLE_OWNER_H ? INJ_CARRYH : (ROLE == ROLE_OWNER_SLOT ? INJ_SLOT : INJ_NONE))));
                constexpr bool OWN = INJ != INJ_NONE;
                const double carryF = (OWN && INJ != INJ_SLOT) ? cin * ph.cF : 0.0;   // non-zero on the owner thread
                PH_START();
                if (tbl) step_a<C, WARPS, AMER, true, INJ_NONE>(V, L, ph, sm, k0, cin, 0.0, e1f, Eprev);
                else if (scl) step_a<C, WARPS, AMER, false, INJ>(V, L, ph, sm, k0, cin, carryF, e1f, Eprev);
                PH_TICK(0);
                FDM_CTA_STEP_SYNC();
                PH_TICK(1);
                if (tbl) step_b<C, WARPS, AMER, true, INJ_NONE, K1>(V, ph, sm, wf, cin_b, e1f, Eprev, 0.0, e1b, Enext, eend_b, csf, cmid);
                else if (scl) step_b<C, WARPS, AMER, false, INJ, K1>(V, ph, sm, wf, cin_b, e1f, Eprev, carryF, e1b, Enext, eend_b, csf, cmid);
                PH_TICK(2);
                FDM_CTA_STEP_SYNC();
                PH_TICK(3);
                const bool r0c = ANY || ROLE == ROLE_PAD ? row0_corr : (ROLE == ROLE_SCALAR_R0 || OWN || ROLE == ROLE_SMC);
                if (tbl) step_c<C, WARPS, AMER, true, K1>(V, L, ph, sm, cin_b, e1b, Enext, csf, cmid, k0, cin);
                else if (!r0c) {
                    if (scl) step_c<C, WARPS, AMER, false, K1>(V, L, ph, sm, cin_b, e1b, Enext, csf, cmid, -1, 0.0);
                } else {
                    // row-0 correction: the owner of row 0 knows xs at its slot 0 (= NCC^k0 xs_0: its padding slots pass
                    // the value through) from the local pass and its carry, before the last pass starts
                    // (xs_row0: within the owner warp by shuffle; through shared memory and a third block barrier only when
                    // the correction reaches beyond it)
                    // (not in the 16 x 16 American class, see scan_level_skip: there the shuffle form measured 2.5 % slower)
                    constexpr bool BY_SHFL = scan_level_skip<C, WARPS, AMER>();
                    double cs_b = 0.0, xs_row0 = 0.0;
                    if (scl) {
                        cs_b = bwd_carry<WARPS, K1>(ph, sm, Enext);
                        if (OWN && (BY_SHFL || k0 >= 0)) {
                            xs_row0 = fma(-sm.smv[2], carryF, fma(sm.cst[CST_AB], cs_b, eend_b));
                            if ((!BY_SHFL || ph.bar3) && k0 >= 0) sm.smv[1] = xs_row0;
                        }
                    }
                    PH_TICK(5);
                    if (ph.bar3) { __syncthreads(); xs_row0 = sm.smv[1]; }
                    else if (!BY_SHFL) { __syncwarp(); if (OWN) xs_row0 = sm.smv[1]; }
                    else if (OWN) xs_row0 = __shfl_sync(FULL, xs_row0, (pad / C) & 31);
                    PH_TICK(6);
                    if (scl) {
                        const bool own = ANY ? w == wf : OWN;
                        const bool smc = ANY ? ph.smc != 0 : ROLE == ROLE_SMC;
                        constexpr bool SLOTINJ = INJ == INJ_SLOT || INJ == INJ_RUNTIME;
                        const int k0s = (INJ == INJ_SLOT || (INJ == INJ_RUNTIME && ph.inj)) ? k0 : -1;
                        if (own) step_cs<C, WARPS, AMER, true, true, SLOTINJ>(V, L, ph, sm, cs_b, cin_b, e1b, csf, cmid, k0s, cin, xs_row0);
                        else if (smc) step_cs<C, WARPS, AMER, true, false, false>(V, L, ph, sm, cs_b, cin_b, e1b, csf, cmid, -1, 0.0, xs_row0);
                        else step_cs<C, WARPS, AMER, false, false, false>(V, L, ph, sm, cs_b, cin_b, e1b, csf, cmid, -1, 0.0, 0.0);
                    }
                }
                PH_TICK(4);
                Vb = w_now;                                     // V[0] = V_min, V[-1] = V_max (:505-506)
                ++gstep;
                if (TRACE && trace) {
                    int lo = 0x7fffffff, hi = -1;
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        const int row = t * C + k - pad;
                        if (AMER && row >= 0 && L[k] > 0.0) { lo = min(lo, row); hi = max(hi, row); }
                    }
                    __syncthreads();
                    if (t == 0) { sm.ish[6] = 0x7fffffff; sm.ish[7] = -1; }
                    __syncthreads();
                    if (hi >= 0) { atomicMin(&sm.ish[6], lo); atomicMax(&sm.ish[7], hi); }
                    __syncthreads();
                    if (t == 0 && 2 * (int64_t)gstep <= O.ex_trace_stride) {
                        trace[2 * (gstep - 1)] = (sm.ish[7] < 0) ? -1 : sm.ish[6];
                        trace[2 * (gstep - 1) + 1] = sm.ish[7];
                    }
                }
                if (gstep == next_mon) {                        // discrete_barrier_fdm_pricer_cn.py:299-300
                    const double reb = B.mon_rebate ? B.mon_rebate[mon_i] : rebate;
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        const int j = t * C + k - pad + 1;
                        if (j >= 1 && (j <= ko_a || j >= ko_b)) V[k] = reb;
                    }
                    if (t == t0 && 0 <= ko_a) Vb = reb;
                    if (t == NT - 1 && n >= ko_b) Vb = reb;
                    ++mon_i;
                    next_mon = (mon_i < mon_e) ? B.mon_step[mon_i] : -1;
                    cin_n = ph.cB * Vb - ph.cA * w_next;
                }
                cin = cin_n;
            }
            };
#ifdef FDM_CTA_ONE_LOOP
            time_loop(IntC<ROLE_ANY>{}, IntC<0>{});
#else
            const int role = ph.cls == 0 ? ROLE_PAD : (ph.cls == 2 ? ROLE_TBL : (!row0_corr ? ROLE_SCALAR
                             : (w == wf ? (ph.inj ? ROLE_OWNER_SLOT : (sm.smv[4] != 0.0 ? ROLE_OWNER_H : ROLE_OWNER)) : (ph.smc ? ROLE_SMC : ROLE_SCALAR_R0))));
            auto by_role = [&](auto k1_tag) {
                switch (role) {
                case ROLE_PAD: time_loop(IntC<ROLE_PAD>{}, k1_tag); break;
                case ROLE_TBL: time_loop(IntC<ROLE_TBL>{}, k1_tag); break;
                case ROLE_SCALAR: time_loop(IntC<ROLE_SCALAR>{}, k1_tag); break;
                case ROLE_OWNER: if constexpr (R0) time_loop(IntC<ROLE_OWNER>{}, k1_tag); break;
                case ROLE_OWNER_H: if constexpr (R0) time_loop(IntC<ROLE_OWNER_H>{}, k1_tag); break;
                case ROLE_OWNER_SLOT: if constexpr (R0) time_loop(IntC<ROLE_OWNER_SLOT>{}, k1_tag); break;
                case ROLE_SMC: if constexpr (R0) time_loop(IntC<ROLE_SMC>{}, k1_tag); break;
                default: if constexpr (R0) time_loop(IntC<ROLE_SCALAR_R0>{}, k1_tag); break;
                }
            };
#ifdef FDM_CTA_NO_K1
            by_role(IntC<0>{});
#else
            // (not with the third barrier of a far-reaching row-0 correction: that combination measured slower than the
            // general loops -- 2048 nodes x 400 steps, sigma 0.25..0.30: 1.65 ms against 1.2..1.4)
            if (k1_mode<C, WARPS, AMER>() && ph.Kf == 1 && ph.Kb == 1 && !ph.bar3) by_role(IntC<1>{}); else by_role(IntC<0>{});
#endif
#endif
        }
        if (AMER) {                                             // V = vc + dt lambda (L keeps lambda for the read-out)
#pragma unroll
            for (int k = 0; k < C; ++k) V[k] = fma(0.5, L[k], V[k]);
        }
        if (AMER && cash != 0.0) {
            CTA_TICK(0);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const int j = t * C + k - pad + 1;
                if (j >= 1) sm.r1[jsw(j)] = V[k];
            }
            Vb = dividend_jump<C, WARPS>(Vb, cash, is_call, k_eff, n, pad, x_min, dx, sm.r1, sm.r2, sm.r0);
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const int j = t * C + k - pad + 1;
                V[k] = (j >= 1) ? sm.r2[jsw(j)] : 0.0;
            }
            __syncthreads();
            fill_payoff(false);                                 // r0 held cz during the jump
            CTA_TICK(2);
        }
    }

    // ---- read-out: stage the value vector by node in shared memory, thread 0 finishes -------------
    CTA_TICK(0);
#ifdef FDM_CTA_PROFILE
    if (blockIdx.x == 40 && (threadIdx.x & 31) == 0)
        printf("  warp %2d per step: step_a %lld | wait %lld | step_b %lld | wait %lld | step_c %lld (carry+publish %lld, sync %lld)\n", threadIdx.x >> 5, ph_acc[0] / max(gstep, 1),
               ph_acc[1] / max(gstep, 1), ph_acc[2] / max(gstep, 1), ph_acc[3] / max(gstep, 1), ph_acc[4] / max(gstep, 1), ph_acc[5] / max(gstep, 1), ph_acc[6] / max(gstep, 1));
    if (blockIdx.x == 40 && threadIdx.x == 32 * (WARPS - 1))
        printf("cta<%d warps> n=%d steps=%d cycles: prologue+loops %lld | setup_phase %lld | dividend jumps %lld\n", WARPS, n, gstep,
               prof_acc[0], prof_acc[1], prof_acc[2]);
#endif
    __syncthreads();
    double *Vs = sm.r1;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = t * C + k - pad + 1;
        if (j >= 1) Vs[j] = V[k];
    }
    if (t == t0) Vs[0] = Vb;
    if (t == NT - 1) Vs[n] = Vb;
    if (t == 0) { sm.ish[6] = 0x7fffffff; sm.ish[7] = -1; }
    __syncthreads();
    if (AMER) {
        int lo = 0x7fffffff, hi = -1;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int row = t * C + k - pad;
            if (row >= 0 && L[k] > 0.0) { lo = min(lo, row); hi = max(hi, row); }
        }
        if (hi >= 0) { atomicMin(&sm.ish[6], lo); atomicMax(&sm.ish[7], hi); }
    }
    if (O.v_dump) {
        double *dst = O.v_dump + sid * O.v_dump_stride;
        for (int j = t; j <= n && j < O.v_dump_stride; j += NT) dst[j] = Vs[j];
    }
    __syncthreads();
    if (t != 0) return;
    const int ex_lo = (AMER && sm.ish[7] >= 0) ? sm.ish[6] : -1;
    const int ex_hi = (AMER && sm.ish[7] >= 0) ? sm.ish[7] : -1;
    read_out(O, sid, Vs, n, snap, x_min, dx, spot, spot_idx, strike_idx, ex_lo, ex_hi, sm.ish[5]);
}

}  // namespace cta
