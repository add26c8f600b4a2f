// fdm_lean.cuh -- the register-lean warp-per-solve kernel: same step algebra as fdm_warp_kernel (see the header
// of fdm_b200.cu), but NO per-row coefficient registers, so 12-16 solves are resident per SM instead of 8.
// fdm_warp_kernel is latency-bound at two warps per scheduler (a lone warp needs ~1270 cycles per time step for
// ~290 cycles of FP64 pipe work); this kernel buys the missing warps with registers.
//
//   * Constant-coefficient LU.  The pivots of the Toeplitz matrix A converge to a fixed point den_inf; the
//     factors built from it, L_inf U_inf = A_inf, reproduce A except for entry (0,0) = den_inf instead of A_C.
//     Sherman-Morrison puts the missing delta = A_C - den_inf back exactly:
//         A x = R   <=>   A_inf x = R - delta x_0 e0,     x_0 = y_0 / (1 + delta z_0),  y = A_inf^-1 R, z = A_inf^-1 e0
//     i.e. the same sweeps with R_0 lowered by g = delta x_0.  In the forward sweep that is -g nm^i on row i --
//     a change of every lane's carry-in by a per-lane constant times g -- and in the backward sweep -g xsz_i,
//     again a per-lane constant on the carry-in.  So the correction costs a few scalar FMAs per step, no per-node
//     work and no per-node registers.
//   * Each sweep is local pass -> carries -> local pass.  Here the forward local pass stores its result u_loc and
//     the second forward pass becomes independent FMAs  u_k = u_loc_k + nm^(k+1) carry  (powers of nm: C/2
//     warp-uniform registers), which also lets the backward local pass run on u_loc while the forward carries
//     are still in flight, and lets x_0 (one shuffle) be known before the last pass.
//     Per node and step: 8 FP64 instructions American, 6 linear (fdm_warp_kernel: 9 / 6).
//   * No padding.  Lane l holds base + (l < extra) rows (base = n_int / 32): only the last slot of a lane is
//     optional and it is predicated, so three scalars (nm, ncc, inv) serve every slot of every lane.
//     A kernel compiled for C slots serves 32 (C-1) <= n_int <= 32 C.
//   * The payoff lives in shared memory (one conflict-free LDS per node and step), scaled by 3 (below).
//   * American state is (vc, 3 dt lambda) instead of (V, dt lambda): with V = vc + dt lambda,
//         theta = 1/2:  R = 2 V + L = 2 vc + Ls          G = V + L = vc + (2/3) Ls
//         theta = 1  :  R =   V + L =   vc + (2/3) Ls    G =     L =      (1/3) Ls          (Ls = 3 L)
//     and Ls_new = relu(3 phi - 3 vc): one FP64 instruction fewer per node than carrying V.  Where a node is
//     exercised, V = phi exactly (fd_american_option_pricer.py:497), rebuilt as phi3 / 3 at read-out.
//     The growth-factor matching of fdm_b200.cu (pd) is folded into inv, as the linear march always did.
//
// Solves this kernel does not take (dividend jumps, pivots not converged within MAX_PIVOT_ITERS rows, a zero
// off-diagonal) are
// marked by a NaN payload in price[] and solved by fdm_warp_kernel in a following launch.
//
// Included by fdm_lean_inst.cu (one translation unit per size class).

#pragma once

#include "fdm_common.cuh"
#include "fdm_lean_classes.cuh"

namespace lean {



enum { LC_KAP = 0, LC_KAPL, LC_KAPU, LC_D, LC_ZC, LC_AB1, LC_AB2, LC_ABL, LC_CA, LC_CB, LC_N };
constexpr int PW_MAX = 16;      // powers of NM kept in shared memory (warp-uniform, broadcast reads)

// shared-memory tables of one warp (doubles): phi3 [C][32] | Wf, Wb [DIRECT_TERMS][32] | Af, Ab [SCAN_LEVELS][32]
template <int C>
struct Tables {
    double *phi3, *Wf, *Wb, *Af, *Ab, *par;
    __device__ explicit Tables(double *base)
    {
        phi3 = base;
        Wf = phi3 + 32 * C;
        Wb = Wf + DIRECT_TERMS * 32;
        Af = Wb + DIRECT_TERMS * 32;
        Ab = Af + SCAN_LEVELS * 32;
        par = Ab + SCAN_LEVELS * 32;        // [0] k_eff [1] s_n [2] s_0: inputs of the boundary refill, parked out of registers
        lc = par + 4;                       // [LC_N][32] per-lane phase constants read once per step
    }
    __device__ double *end() const { return lc + LC_N * 32 + PW_MAX; }
    double *lc;
    // P[j] = NM^j, j = 0..C-C/2 for the 16-slot class.  Kept out of registers on purpose: at 128 registers the
    // compiler would otherwise rebuild the powers with a dependent DMUL chain in every time step.
    __device__ double pw(int j) const { return ((const volatile double *)(lc + LC_N * 32))[j]; }
};
template <int C>
constexpr int smem_doubles() { return 32 * C + 2 * DIRECT_TERMS * 32 + 2 * SCAN_LEVELS * 32 + 4 + LC_N * 32 + PW_MAX; }

template <int C>
struct Ph {
    static constexpr int H = C / 2, PW = C - C / 2;
#ifndef FDM_LEAN_PSMEM_MIN
#define FDM_LEAN_PSMEM_MIN 14
#endif
    // where the powers of NM live: registers win up to 13 slots (re-measured with the per-form loops: 400x400
    // American 8.49e11 with registers, 7.60e11 with shared memory), shared memory from 14
    static constexpr bool P_SMEM = C >= FDM_LEAN_PSMEM_MIN;
    double NM, NCC, INV;      // -A_L/den_inf, -A_U/den_inf, scale/den_inf
    double P[PW + 1];         // P[j] = NM^j, j = 0..PW (registers, C <= 13; else Tables::pw)
    double cA, cB;            // boundary coefficients of this lane, already divided by NM (lane 0) / NCC (lane 31)
    double q;                 // (1-theta)/theta  (linear march)
    double kap;               // g = kap * xs_0
    double D, Zc;             // Sherman-Morrison: forward carry-in -= g D, backward carry-in -= g Zc   (per lane)
    double kapL, kapU;        // sum_k NCC^k NM^(k+1) over the lower / upper half of this lane
    double ab1, ab2, abl;     // NCC^(cnt-H), NCC^H, NCC^cnt
    int terms;                // direct carry terms (0: Kogge-Stone scan)
    bool half;                // theta == 1/2
};

__device__ __forceinline__ double ipow(double a, int n)
{
    double r = 1.0;
    for (int i = 0; i < n; ++i) r *= a;
    return r;
}

// Sweeps of the constant-coefficient factors on a register vector with exact Kogge-Stone carries: on return V
// holds xs (den_inf * A_inf^-1 V).  Used once per phase for z; not on the per-step path.
// nmL / nccL: multipliers of the optional last slot (NM / NCC where the lane holds it, else 1 with V[C-1] = 0).
template <int C>
__device__ __forceinline__ void const_sweeps(double (&V)[C], const bool last, const int cnt, const double NM,
                                             const double NCC, const int lane)
{
    const double nmL = last ? NM : 1.0, nccL = last ? NCC : 1.0;
    double e = 0.0;
#pragma unroll
    for (int k = 0; k < C; ++k) e = fma(k < C - 1 ? NM : nmL, e, V[k]);
    double E = e, A = ipow(NM, cnt);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double tE = __shfl_up_sync(FULL, E, d), tA = __shfl_up_sync(FULL, A, d);
        const bool has = lane >= d;
        E = fma(has ? A : 0.0, tE, E);
        A *= has ? tA : 1.0;
    }
    double c = __shfl_up_sync(FULL, E, 1);
    c = (lane == 0) ? 0.0 : c;
#pragma unroll
    for (int k = 0; k < C; ++k) { c = fma(k < C - 1 ? NM : nmL, c, V[k]); V[k] = (k < C - 1 || last) ? c : 0.0; }
    e = 0.0;
#pragma unroll
    for (int k = C - 1; k >= 0; --k) e = fma(k < C - 1 ? NCC : nccL, e, V[k]);
    E = e;
    A = ipow(NCC, cnt);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double tE = __shfl_down_sync(FULL, E, d), tA = __shfl_down_sync(FULL, A, d);
        const bool has = lane + d < 32;
        E = fma(has ? A : 0.0, tE, E);
        A *= has ? tA : 1.0;
    }
    c = __shfl_down_sync(FULL, E, 1);
    c = (lane == 31) ? 0.0 : c;
#pragma unroll
    for (int k = C - 1; k >= 0; --k) { c = fma(k < C - 1 ? NCC : nccL, c, V[k]); V[k] = (k < C - 1 || last) ? c : 0.0; }
}

// Phase constants and carry tables.  Returns false if this solve has to go to fdm_warp_kernel.
template <int C, bool AMER>
__device__ __forceinline__ bool setup_phase(double theta, double dt, double a, double c, double bcoef, int n_int,
                                            int lane, const RowLayout &ly, const Tables<C> &T, Ph<C> &ph)
{
    constexpr int H = C / 2, PW = C - H;
    const bool last = ly.cnt == C;
    const int cntU = ly.cnt - H;
    const double A_L = __dmul_rn(__dmul_rn(-theta, dt), a);
    const double A_C = __dsub_rn(1.0, __dmul_rn(__dmul_rn(theta, dt), bcoef));
    const double A_U = __dmul_rn(__dmul_rn(-theta, dt), c);
    const double B_L = __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), a);
    const double B_C = __dadd_rn(1.0, __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), bcoef));
    const double B_U = __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), c);
    const double p = 1.0 / theta;
    ph.q = (1.0 - theta) / theta;
    ph.half = theta == 0.5;

    // pivots of the reference's recurrence (fd_american_option_pricer.py:443-451) to their bit-wise fixed point
    double den = A_C, rinv = __drcp_rn(den);
    bool conv = false;
    for (int i = 1; i < MAX_PIVOT_ITERS; ++i) {
        const double den_i = __dsub_rn(A_C, __dmul_rn(A_L, __dmul_rn(A_U, rinv)));
        if (den_i == den) { conv = true; break; }
        den = den_i;
        rinv = __drcp_rn(den);
    }
    const double NM = -(A_L * rinv), NCC = -(A_U * rinv);
    // the verdict is taken by a warp vote at the end, so that the compiler sees uniform control flow around it
    const bool bad = !conv || !(fabs(NM) > 0.0) || !(fabs(NCC) > 0.0) || !isfinite(rinv);
    const double delta_sm = __dmul_rn(A_L, __dmul_rn(A_U, rinv));       // A_C - den_inf, as the recurrence forms it

    // growth-factor matching (fdm_b200.cu): the V coefficient carries 1 + delta; folded into inv
    const double pd = growth_matched_pd(A_L, A_C, A_U, B_L, B_C, B_U, ph.q, rinv, n_int);
    const double scale = AMER ? 1.0 + (pd - 1.0) / (1.0 + ph.q) : p + (pd - 1.0);
    const double bscale = 1.0 / scale;
    ph.NM = NM;
    ph.NCC = NCC;
    // American, theta = 1/2: the step solves for 2 R (see march_step), so the Dirichlet terms double and the pivot
    // scaling halves -- powers of two, exact
    const double dbl = (AMER && ph.half) ? 2.0 : 1.0;
    ph.INV = rinv * scale / dbl;
    ph.cA = (lane == 0) ? A_L * bscale / NM * dbl : ((lane == 31) ? A_U * bscale / NCC * dbl : 0.0);
    ph.cB = (lane == 0) ? B_L * bscale / NM * dbl : ((lane == 31) ? B_U * bscale / NCC * dbl : 0.0);

    ph.P[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= PW; ++j) ph.P[j] = ph.P[j - 1] * NM;
    __syncwarp();
    if (Ph<C>::P_SMEM && lane == 0) {
#pragma unroll
        for (int j = 0; j <= PW; ++j) T.lc[LC_N * 32 + j] = ph.P[j];
    }
    const double af1 = ipow(NM, H), af2 = ipow(NM, cntU);
    ph.ab1 = ipow(NCC, cntU);
    ph.ab2 = ipow(NCC, H);
    const double af_lane = af1 * af2, ab_lane = ph.ab1 * ph.ab2;
    ph.abl = ab_lane;
    // backward local end of  NM^(k+1) carry : sum_k NCC^k NM^(k+1) over the slots of a half
    {
        double sL = 0.0, sU = 0.0, t = NM;
        for (int k = 0; k < H; ++k) { sL += t; t *= NCC * NM; }
        t = NM;
        for (int k = 0; k < cntU; ++k) { sU += t; t *= NCC * NM; }
        ph.kapL = sL;
        ph.kapU = sU;
    }

    __syncwarp();
    // Kogge-Stone multipliers and direct-carry weights, as in fdm_warp_kernel's setup_phase
    double af = af_lane, ab = ab_lane;
#pragma unroll
    for (int lv = 0; lv < SCAN_LEVELS; ++lv) {
        const int d = 1 << lv;
        const bool has_f = lane >= d, has_b = lane + d < 32;
        T.Af[lv * 32 + lane] = has_f ? af : 0.0;
        T.Ab[lv * 32 + lane] = has_b ? ab : 0.0;
        const double tf = __shfl_up_sync(FULL, af, d), tb = __shfl_down_sync(FULL, ab, d);
        if (has_f) af *= tf;
        if (has_b) ab *= tb;
    }
    double wf = (lane >= 1) ? 1.0 : 0.0, wb = (lane <= 30) ? 1.0 : 0.0;
    int terms = 0;
#pragma unroll
    for (int j = 1; j <= DIRECT_TERMS; ++j) {
        T.Wf[(j - 1) * 32 + lane] = wf;
        T.Wb[(j - 1) * 32 + lane] = wb;
        const double tf = __shfl_up_sync(FULL, af_lane, j), tb = __shfl_down_sync(FULL, ab_lane, j);
        wf = (lane >= j + 1) ? wf * tf : 0.0;
        wb = (lane + j + 1 <= 31) ? wb * tb : 0.0;
        if ((j & 1) == 0 && terms == 0 && !__any_sync(FULL, fabs(wf) > 1e-20 || fabs(wb) > 1e-20)) terms = j;
    }
    ph.terms = terms;
    __syncwarp();

    // Sherman-Morrison constants.  xsz = sweeps of e0 (den_inf * A_inf^-1 e0), exact carries.
    double xsz[C];
#pragma unroll
    for (int k = 0; k < C; ++k) xsz[k] = (lane == 0 && k == 0) ? 1.0 : 0.0;
    const_sweeps<C>(xsz, last, ly.cnt, NM, NCC, lane);
    const double xs00 = __shfl_sync(FULL, xsz[0], 0);
    ph.kap = delta_sm * rinv / (1.0 + delta_sm * (rinv * xs00));       // g = delta x_0 = kap * xs_0(uncorrected)
    // forward: rows of lane l change by -g NM^(row0 + k) = NM^(k+1) * (-g NM^(row0 - 1))
    double Dl = (lane == 0) ? 1.0 / NM : ipow(NM, ly.row0 - 1);
    // backward: the carry into lane l is xs of the first row of lane l+1, which changes by -g xsz there
    double Zl = __shfl_down_sync(FULL, xsz[0], 1);
    if (lane == 31) Zl = 0.0;
    // Contributions below 1e-30 of the leading one are dropped: the direct carries neglect terms below 1e-20 as
    // well, and where they leave an exact zero (far out-of-the-money rows in the first steps) the correction must
    // not turn it negative -- the reference's exercise test phi > vc (:496) is strict.
    if (!(fabs(Dl) > 1e-30)) Dl = 0.0;
    if (!(fabs(Zl) > 1e-30 * fabs(xs00))) Zl = 0.0;
    ph.D = Dl;
    ph.Zc = Zl;
    if (Ph<C>::P_SMEM) {
        T.lc[LC_KAP * 32 + lane] = ph.kap;   T.lc[LC_KAPL * 32 + lane] = ph.kapL; T.lc[LC_KAPU * 32 + lane] = ph.kapU;
        T.lc[LC_D * 32 + lane] = ph.D;       T.lc[LC_ZC * 32 + lane] = ph.Zc;     T.lc[LC_AB1 * 32 + lane] = ph.ab1;
        T.lc[LC_AB2 * 32 + lane] = ph.ab2;   T.lc[LC_ABL * 32 + lane] = ph.abl;   T.lc[LC_CA * 32 + lane] = ph.cA;
        T.lc[LC_CB * 32 + lane] = ph.cB;
    }
    __syncwarp();
    return !__any_sync(FULL, bad);
}

// Per-lane phase constants that are used once per step.  In the 16-slot class they are read from shared memory at
// the point of use (volatile: not hoisted back into registers), which frees ~20 registers in the time loop -- there
// it is the difference between losing to fdm_warp_kernel and beating it; with 13 slots or fewer there is room and
// registers are faster (measured both ways).
#define LCV(ph, T, lane, field, idx) \
    (Ph<C>::P_SMEM ? ((const volatile double *)(T).lc)[(idx) * 32 + (lane)] : (ph).field)

// direct carry, split in two so that independent work can be issued between the shuffles and their use
template <bool UP>
struct CarryShf {
    double t1, t2, t3, t4, t5, t6;       // scalars (an array here ends up in local memory); combine() reads exactly
                                         // the ones issue() wrote for the same `terms`, so none is pre-set
    __device__ __forceinline__ void issue(const double e_end, const int terms)
    {
        if (terms) {
            t1 = UP ? __shfl_up_sync(FULL, e_end, 1) : __shfl_down_sync(FULL, e_end, 1);
            t2 = UP ? __shfl_up_sync(FULL, e_end, 2) : __shfl_down_sync(FULL, e_end, 2);
            if (terms > 2) {
                t3 = UP ? __shfl_up_sync(FULL, e_end, 3) : __shfl_down_sync(FULL, e_end, 3);
                t4 = UP ? __shfl_up_sync(FULL, e_end, 4) : __shfl_down_sync(FULL, e_end, 4);
            }
            if (terms > 4) {
                t5 = UP ? __shfl_up_sync(FULL, e_end, 5) : __shfl_down_sync(FULL, e_end, 5);
                t6 = UP ? __shfl_up_sync(FULL, e_end, 6) : __shfl_down_sync(FULL, e_end, 6);
            }
        }
    }
    __device__ __forceinline__ double combine(const double e_end, const int terms, const double *__restrict__ W,
                                              const double *__restrict__ A, const int lane) const
    {
        if (terms) {
            double acc0 = W[0 * 32 + lane] * t1, acc1 = W[1 * 32 + lane] * t2;
            if (terms > 2) { acc0 = fma(W[2 * 32 + lane], t3, acc0); acc1 = fma(W[3 * 32 + lane], t4, acc1); }
            if (terms > 4) { acc0 = fma(W[4 * 32 + lane], t5, acc0); acc1 = fma(W[5 * 32 + lane], t6, acc1); }
            return acc0 + acc1;
        }
        return lane_carry<UP>(e_end, 0, W, A, lane);        // Kogge-Stone scan
    }
};

// One theta-scheme step.  V: vc (American) / V (linear) on entry and exit; L: 6 dt lambda (American) / scratch.
// The American state keeps Lq = 6 dt lambda = 2 relu(3 phi - 3 vc) because that is ONE FP64 instruction, d + |d|
// (exact: 2 d or +0), where relu itself costs three integer instructions per node on the sign mask; the kernel is
// issue-limited, not FP64-pipe-limited.  The factor two is carried by exact power-of-two constants: with theta = 1/2
// the right-hand side becomes 2 R = 4 vc + Lq (the system is linear: INV is halved, the Dirichlet terms doubled in
// setup_phase) and G = vc + Lq / 3; with theta = 1, R = vc + Lq / 3 and G = Lq / 6.  Every intermediate is the old
// one times a power of two, so the results are bit-identical to the (vc, 3 dt lambda) form.
// The optional last slot is handled without branches: on lanes that do not hold it its multipliers are 1 / 0 and
// its V, L stay 0, which makes it the identity in both sweeps.
template <int C, bool AMER, int TERMS>
__device__ __forceinline__ void march_step(double (&V)[C], double (&L)[C], const Ph<C> &ph, const double cin_f,
                                           const double cin_b, const bool last, const Tables<C> &T, const int lane)
{
    constexpr int H = C / 2, PW = C - H;
    constexpr int terms = TERMS;                        // the time loop is specialised per carry form (see the kernel)
    const double NM = ph.NM, NCC = ph.NCC;
    const double nmL = last ? NM : 1.0, nccL = last ? NCC : 1.0;
    // right-hand side R (into V) and subtrahend G (into L)
    if (AMER) {
        if (ph.half) {
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const double g = fma(1.0 / 3.0, L[k], V[k]);
                V[k] = fma(4.0, V[k], L[k]);
                L[k] = g;
            }
        } else {
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const double g = (1.0 / 6.0) * L[k];
                V[k] = fma(1.0 / 3.0, L[k], V[k]);
                L[k] = g;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k) L[k] = ph.q * V[k];
    }
    // ---- forward local pass (zero carry-in; the Dirichlet term of row 0 is lane 0's own start value),
    //      result u_loc kept in V ------------------------------------------------------------------------------
    double a = cin_f, b = 0.0;
#pragma unroll
    for (int k = 0; k < H; ++k) { a = fma(NM, a, V[k]); V[k] = a; }
#pragma unroll
    for (int k = H; k < C - 1; ++k) { b = fma(NM, b, V[k]); V[k] = b; }
    b = fma(nmL, b, V[C - 1]);
    V[C - 1] = last ? b : 0.0;
    auto pw = [&](const int j) -> double { return Ph<C>::P_SMEM ? T.pw(j) : ph.P[j]; };
    const double af2 = last ? pw(PW) : pw(PW - 1);
    const double e_end = fma(af2, a, b);
    CarryShf<true> shf_f;
    shf_f.issue(e_end, terms);
    // ---- backward local pass on u_loc, while the forward carries are in flight ------------------------
    double e1 = cin_b, e2 = 0.0;
#pragma unroll
    for (int k = C - 1; k >= H; --k) e1 = fma(k < C - 1 ? NCC : nccL, e1, V[k]);
#pragma unroll
    for (int k = H - 1; k >= 0; --k) e2 = fma(NCC, e2, V[k]);
    const double c_in = shf_f.combine(e_end, terms, T.Wf, T.Af, lane);            // carry into slot 0
    const double pH = pw(H);
    const double c_mid = fma(pH, c_in, a);                                           // carry into slot H (= u_{H-1})
    // true u = u_loc + NM^(k+1) carry  =>  backward local ends pick up kap * carry
    const double kapU = LCV(ph, T, lane, kapU, LC_KAPU);
    double e_end_b = fma(LCV(ph, T, lane, ab2, LC_AB2), fma(kapU, c_mid, e1), fma(LCV(ph, T, lane, kapL, LC_KAPL), c_in, e2));
    CarryShf<false> shf_b;
    shf_b.issue(e_end_b, terms);
    double cb = shf_b.combine(e_end_b, terms, T.Wb, T.Ab, lane);
    // Sherman-Morrison: g = delta x_0 from lane 0's inclusive end value xs_0; R_0 -= g
    const double g = LCV(ph, T, lane, kap, LC_KAP) * __shfl_sync(FULL, fma(LCV(ph, T, lane, abl, LC_ABL), cb, e_end_b), 0);
    const double c_in2 = fma(-g, LCV(ph, T, lane, D, LC_D), c_in);
    const double c_mid2 = fma(pH, c_in2, a);
    cb = fma(-g, LCV(ph, T, lane, Zc, LC_ZC), cb);
    double c1 = cb + cin_b;                                       // xs above the top slot
    double c2 = fma(LCV(ph, T, lane, ab1, LC_AB1), cb, fma(kapU, c_mid2, e1));        // xs of slot H
    // ---- second passes: u = u_loc + P carry (independent), xs chain, projection --------------------------
    auto slot_of = [](const int i) -> int {
        if (i == 0) return C - 1;
        const int j = (i - 1) >> 1;
        const bool upper_exists = C - 2 - j >= H;
        if (!upper_exists) return H - 1 - ((i - 1) - (C - 1 - H));      // only the lower chain is left
        return ((i - 1) & 1) == 0 ? C - 2 - j : H - 1 - j;
    };
    constexpr int PF = 4;                                         // payoff prefetch distance (LDS latency off the chain)
    double pf[C + PF];
    if (AMER) {
#pragma unroll
        for (int i = 0; i < PF; ++i) pf[i] = T.phi3[slot_of(i < C ? i : C - 1) * 32 + lane];
    }
    const double pL = last ? pw(PW) : 0.0;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        const int k = slot_of(i);
        double xs;
        bool keep = true;
        if (k >= H) {
            const double u = fma(k < C - 1 ? pw(k - H + 1) : pL, c_mid2, V[k]);
            c1 = fma(k < C - 1 ? NCC : nccL, c1, u);
            xs = c1;
            keep = k < C - 1 ? true : last;
        } else {
            const double u = fma(pw(k + 1), c_in2, V[k]);
            c2 = fma(NCC, c2, u);
            xs = c2;
        }
        const double vc = fma(ph.INV, xs, -L[k]);
        if (AMER) {
            if (i + PF < C) pf[i + PF] = T.phi3[slot_of(i + PF) * 32 + lane];
            // exercise iff phi > vc (fd_american_option_pricer.py:496); Lq = 6 dt lambda_new = 2 relu(3 phi - 3 vc)
            const double d = fma(-3.0, vc, pf[i]);
            const double lq = d + fabs(d);
            L[k] = keep ? lq : 0.0;
        }
        V[k] = keep ? vc : 0.0;
    }
}

#ifndef FDM_LEAN_MINB
#define FDM_LEAN_MINB 0
#endif
template <int C>
constexpr int min_blocks() { return FDM_LEAN_MINB ? FDM_LEAN_MINB : (C <= 4 ? 24 : C <= 8 ? 20 : 16); }

template <int C, bool AMER, bool TRACE>
__global__ void __launch_bounds__(32, min_blocks<C>())
fdm_lean_kernel(const FdmBatch B, const FdmOut O, const int64_t first, const int want_american)
{
    extern __shared__ double smem[];
    const Tables<C> T(smem);

    const int64_t sid = first + blockIdx.x;
    if (sid >= B.n_solves) return;
    const int flags = B.flags[sid];
    if (((flags & FDM_FLAG_AMERICAN) != 0) != (want_american != 0)) return;
    if (flags & FDM_FLAG_INVALID) return;                       // rejected by the book fan-out: status stays -1
    const int lane = threadIdx.x;
    const int n = B.n_space[sid];
    const int n_int = n - 1;
    if (class_of(n_int) != C) return;                           // another size class, or fdm_warp_kernel's
    const RowLayout ly = RowLayout::packed(n_int, lane);
    const bool last = ly.cnt == C;
    const bool is_call = flags & FDM_FLAG_CALL;
    const bool b76 = flags & FDM_FLAG_BLACK76;
    const bool snap = flags & FDM_FLAG_SNAP;
    const bool ko = flags & (FDM_FLAG_KO_DOWN | FDM_FLAG_KO_UP);

    const double x_min = B.x_min[sid], dx = B.dx[sid], sigma = B.sigma[sid], r = B.r[sid], bcar = B.b[sid];
    const double spot = B.spot[sid], strike = B.strike[sid];
    const double rebate = (ko && B.rebate) ? B.rebate[sid] : 0.0;
    const int rannacher = B.rannacher[sid];

    // ---- grid, snapping, payoff -------------------------------------------------------------
    double V[C], L[C];
    const double s0 = node_s(x_min, dx, 0);
    const double sn = node_s(x_min, dx, n);
#pragma unroll
    for (int k = 0; k < C; ++k) V[k] = (k < ly.cnt) ? node_s(x_min, dx, ly.row0 + k + 1) : 0.0;   // node prices, for now
    int spot_idx = 0, strike_idx = 0;
    double k_eff = strike;
    if (snap) {
        spot_idx = argmin_node<C>(V, s0, sn, spot, n, ly, false);
        strike_idx = argmin_node<C>(V, s0, sn, strike, n, ly, false);
        k_eff = node_s(x_min, dx, strike_idx);                  // fd_american_option_pricer.py:272-275
    }
    int ko_a, ko_b;
    ko_thresholds(B, sid, flags, x_min, dx, n, ko_a, ko_b);
    const bool put_smin = flags & FDM_FLAG_PUT_BDRY_SMIN;
    const double phi0 = payoff(s0, k_eff, is_call);
    const double phin = payoff(sn, k_eff, is_call);
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const double pv = (k < ly.cnt) ? payoff(V[k], k_eff, is_call) : 0.0;
        T.phi3[k * 32 + lane] = 3.0 * pv;
        V[k] = pv;                                              // terminal condition, :558
        L[k] = 0.0;
    }
    // exact payoff of slot k, recomputed (rare uses: dividend jump, read-out); node_s is not inlined
    auto phi_at = [&](int k) -> double {
        return (k < ly.cnt) ? payoff(node_s(x_min, dx, ly.row0 + k + 1), k_eff, is_call) : 0.0;
    };
    // value vector from the (vc, 3 dt lambda) state: V = phi where exercised (:497)
    auto materialise = [&]() {
        if (AMER) {
#pragma unroll
            for (int k = 0; k < C; ++k) if (L[k] > 0.0) V[k] = phi_at(k);
        }
    };
    double Vb = (lane == 0) ? phi0 : ((lane == 31) ? phin : 0.0);
    const bool bdry_lane = is_call ? (lane == 31) : (lane == 0);
    if (lane == 0) {
        T.par[0] = k_eff;
        T.par[1] = (flags & FDM_FLAG_BDRY_TOP_NEXT) ? node_s(x_min, dx, n + 1) : sn;
        T.par[2] = s0;
    }
    __syncwarp();

    // ---- operator coefficients, fd_american_option_pricer.py:413-426 -------------------------
    const double sig2 = __dmul_rn(sigma, sigma);
    const double mu_x = b76 ? __dmul_rn(-0.5, sig2) : __dsub_rn(bcar, __dmul_rn(0.5, sig2));
    const double alpha = __dmul_rn(0.5, sig2) / __dmul_rn(dx, dx);
    const double beta = mu_x / __dmul_rn(2.0, dx);
    const double ca = __dsub_rn(alpha, beta);
    const double cc = __dadd_rn(alpha, beta);
    const double bcoef = __dsub_rn(__dmul_rn(-2.0, alpha), r);

    const int seg_b = B.seg_begin[sid], seg_e = B.seg_begin[sid + 1];
    int mon_i = 0, mon_e = 0;
    if (ko && B.mon_begin) { mon_i = B.mon_begin[sid]; mon_e = B.mon_end ? B.mon_end[sid] : B.mon_begin[sid + 1]; }
    int next_mon = (mon_i < mon_e) ? B.mon_step[mon_i] : -1;
    int gstep = 0;
    int32_t *trace = (TRACE && O.ex_trace) ? O.ex_trace + sid * O.ex_trace_stride : nullptr;
    // solves with a dividend jump go to fdm_warp_kernel (its spline scratch would cut this kernel's occupancy)
    bool retry = false;
    for (int seg = seg_b; seg < seg_e; ++seg) retry |= B.seg_div[seg] != 0.0;

    for (int seg = seg_b; seg < seg_e && !retry; ++seg) {
        const int n_steps = B.seg_steps[seg];
        const double dt = B.seg_dt[seg];
        const int base_ran = B.seg_restart[seg] ? rannacher : 0;   // :464
        if (AMER) {
#pragma unroll
            for (int k = 0; k < C; ++k) L[k] = 0.0;             // lambda restarts every segment, :463
        }
        int m = 0;
        double my_w = 0.0;
        int fill_base = -64;
        auto fetch_boundary = [&](int step) -> double {
            if (step - fill_base >= 32 || step < fill_base) {
                // once per 32 steps: the inputs are re-read (global / shared) instead of living in registers
                fill_base = step;
                const double dt_ = *(volatile const double *)&B.seg_dt[seg];
                const double tau0_ = *(volatile const double *)&B.seg_tau0[seg];
                const double r_ = *(volatile const double *)&B.r[sid];
                const double b_ = B.b_bdry ? *(volatile const double *)&B.b_bdry[sid] : *(volatile const double *)&B.b[sid];
                const double tau = fma((double)(step + lane + 1), dt_, tau0_);
                my_w = moving_boundary_value(is_call, b76, r_, b_, ((volatile double *)T.par)[0],
                                             ((volatile double *)T.par)[1], tau, put_smin, ((volatile double *)T.par)[2]);
            }
            const double w = __shfl_sync(FULL, my_w, step - fill_base);
            return bdry_lane ? w : 0.0;
        };
        while (m < n_steps) {
            const bool implicit_phase = m < base_ran;
            const int m_end = implicit_phase ? min(base_ran, n_steps) : n_steps;
            Ph<C> ph;
            if (!setup_phase<C, AMER>(implicit_phase ? 1.0 : 0.5, dt, ca, cc, bcoef, n_int, lane, ly, T, ph)) {
                retry = true;
                break;
            }
            double w_next = fetch_boundary(m);
            double cin = LCV(ph, T, lane, cB, LC_CB) * Vb - LCV(ph, T, lane, cA, LC_CA) * w_next;   // non-zero on lanes 0 and 31 only
            // The time loop is compiled once per carry form (TT = 2 / 4 / 6 direct terms, 0 = scan), so the step has no
            // branches on ph.terms: ~35 fewer instructions per warp-step (BRA / UISETP / PLOP3 and the shuffle
            // temporaries' moves) and a tighter schedule.  Measured on B200: +5..14 % on every lean shape.
            auto time_loop = [&](auto tt) {
            constexpr int TT = decltype(tt)::value;
            for (; m < m_end; ++m) {
                const double w_now = w_next;
                w_next = fetch_boundary(m + 1);
                double cin_n = LCV(ph, T, lane, cB, LC_CB) * w_now - LCV(ph, T, lane, cA, LC_CA) * w_next;
                march_step<C, AMER, TT>(V, L, ph, (lane == 0) ? cin : 0.0, (lane == 31) ? cin : 0.0, last, T, lane);
                Vb = w_now;                                     // V[0] = V_min, V[-1] = V_max, :505-506
                ++gstep;
                if (TRACE && trace) {
                    int lo = 0x7fffffff, hi = -1;
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        const int row = ly.row0 + k;
                        if (AMER && k < ly.cnt && L[k] > 0.0) { lo = min(lo, row); hi = max(hi, row); }
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
                        const int j = ly.row0 + k + 1;
                        if (k < ly.cnt && (j <= ko_a || j >= ko_b)) V[k] = reb;
                    }
                    if (lane == 0 && 0 <= ko_a) Vb = reb;
                    if (lane == 31 && n >= ko_b) Vb = reb;
                    ++mon_i;
                    next_mon = (mon_i < mon_e) ? B.mon_step[mon_i] : -1;
                    cin_n = LCV(ph, T, lane, cB, LC_CB) * Vb - LCV(ph, T, lane, cA, LC_CA) * w_next;
                }
                cin = cin_n;
            }
            };
            switch (ph.terms) {
            case 2: time_loop(IntC<2>{}); break;
            case 4: time_loop(IntC<4>{}); break;
            case 6: time_loop(IntC<6>{}); break;
            default: time_loop(IntC<0>{}); break;
            }
        }
        if (retry) break;
    }
    if (retry) {
        if (lane == 0) O.price[sid] = __longlong_as_double((long long)RETRY_BITS);
        return;
    }

    // ---- read-out (as fdm_warp_kernel) ------------------------------------------------------------
    const double V0 = __shfl_sync(FULL, Vb, 0), VN = __shfl_sync(FULL, Vb, 31);
    int ex_lo = -1, ex_hi = -1;
    if (AMER) {
        int lo = 0x7fffffff, hi = -1;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int row = ly.row0 + k;
            if (k < ly.cnt && L[k] > 0.0) { lo = min(lo, row); hi = max(hi, row); }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            lo = min(lo, __shfl_xor_sync(FULL, lo, d));
            hi = max(hi, __shfl_xor_sync(FULL, hi, d));
        }
        if (hi >= 0) { ex_lo = lo; ex_hi = hi; }
    }
    materialise();

    if (O.v_dump) {
        double *dst = O.v_dump + sid * O.v_dump_stride;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int j = ly.row0 + k + 1;
            if (k < ly.cnt && j < O.v_dump_stride) dst[j] = V[k];
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
        // fd_american_option_pricer.py:613-653
        const int i_near = min(max(spot_idx, 1), n - 2);
        const double Sx = node_s(x_min, dx, spot_idx);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = i_near - 1 + t;
            sv[t] = node_value<C>(V, V0, VN, j, n, ly);
            ss[t] = node_s(x_min, dx, j);
        }
        price = node_value<C>(V, V0, VN, spot_idx, n, ly);
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
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            L[k] = (k < ly.cnt) ? node_s(x_min, dx, ly.row0 + k + 1) : 0.0;    // L is free now: node prices
            if (k < ly.cnt && L[k] <= spot) ++cnt;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
        int lo = cnt;
        if (spot <= s0) { price = V0; lo = 0; }
        else if (spot >= sn) { price = VN; lo = n; }
        else {
            const double s_lo = node_s(x_min, dx, lo), s_hi = node_s(x_min, dx, lo + 1);
            const double v_lo = node_value<C>(V, V0, VN, lo, n, ly);
            const double v_hi = node_value<C>(V, V0, VN, lo + 1, n, ly);
            const double w = __dsub_rn(spot, s_lo) / __dsub_rn(s_hi, s_lo);
            price = __dadd_rn(__dmul_rn(__dsub_rn(1.0, w), v_lo), __dmul_rn(w, v_hi));
        }
        const int gi = argmin_node<C>(L, s0, sn, spot, n, ly, true);
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

}  // namespace lean

