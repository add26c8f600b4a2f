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
// Dividend jumps run here too: the spline's scratch overlays the tables between two segments (dividend_jump below).
// Solves this kernel does not take (a degenerate operator: zero off-diagonal, non-finite pivot) are
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
__host__ __device__ constexpr int smem_doubles() { return 32 * C + 2 * DIRECT_TERMS * 32 + 2 * SCAN_LEVELS * 32 + 4 + LC_N * 32 + PW_MAX; }

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

// a^n by repeated squaring (n up to ~500: per-lane constants of the row-0 correction; ~2 log2 n multiplications
// instead of a dependent chain of n)
__device__ __forceinline__ double ipow_sq(double a, int n)
{
    double r = 1.0;
#pragma unroll 1
    while (n > 0) {
        if (n & 1) r *= a;
        a *= a;
        n >>= 1;
    }
    return r;
}

// Phase constants and carry tables.  Returns false if this solve has to go to fdm_warp_kernel.
//
// Everything here is closed form -- no iteration to the pivots' bit-wise fixed point, no sweeps of e0.  (The first
// version iterated the pivot recurrence and the reference's (den, c') recurrence to their fixed points, 15..200
// dependent divisions each, swept e0 through the factors with Kogge-Stone carries and built NM^row with a chain of up
// to 500 multiplications: 37 000..51 000 cycles per phase, 5 % of a 500-step solve with its two phases and 17 % of a
// two-dividend solve with its five or six.)
//   * den_inf is the larger root of den^2 - A_C den + A_L A_U = 0 (the stable fixed point of den = A_C - A_L A_U / den),
//     from the quadratic formula and one Newton step with the residual formed by FMAs: correct to an ulp.
//   * the factors built from it reproduce A except for entry (0,0): delta = A_C - den_inf (exact subtraction).
//   * sweeping e0 through the constant factors gives xs_i = NM^i kappa (1 - rho^(n-i)), rho = NM NCC, kappa = 1/(1-rho):
//     xs_0 and the per-lane constants of the row-0 correction follow from powers by repeated squaring.
template <int C, bool AMER>
__device__ __forceinline__ bool setup_phase(double theta, double dt, double a, double c, double bcoef, int n_int,
                                            int lane, const RowLayout &ly, const Tables<C> &T, Ph<C> &ph)
{
    constexpr int H = C / 2, PW = C - H;
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

    // steady-state pivot of the reference's recurrence den_i = A_C - A_L A_U / den_{i-1} (fd_american_option_pricer.py:443-451)
    const double PLU = __dmul_rn(A_L, A_U), PLU_lo = __fma_rn(A_L, A_U, -PLU);
    const double disc = __fma_rn(A_C, A_C, -4.0 * PLU);
    double den = 0.5 * (A_C + sqrt(disc));
#pragma unroll
    for (int it = 0; it < 2; ++it) {                               // Newton on the quadratic, residual in extra precision
        const double res = __fma_rn(__dsub_rn(den, A_C), den, PLU) + PLU_lo;
        den -= res / (2.0 * den - A_C);
    }
    const double rinv = 1.0 / den;
    const double NM = -(A_L * rinv), NCC = -(A_U * rinv);
    const double rho = NM * NCC, kappa = 1.0 / (1.0 - rho);
    // the verdict is taken by a warp vote at the end, so that control flow stays uniform
    const bool bad = !(disc > 0.0) || !(fabs(NM) > 0.0) || !(fabs(NCC) > 0.0) || !isfinite(rinv) || !(fabs(rho) < 1.0) ||
                     !isfinite(kappa);
    const double delta_sm = __dsub_rn(A_C, den);                   // A = A_inf + delta e0 e0^T

    // growth factor of smooth components matched to the exact row sums (growth_exact_pd); folded into inv
    const double pd = growth_exact_pd(A_L, A_C, A_U, B_L, B_C, B_U, ph.q, rinv);
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
        for (int k = 0; k < H; ++k) { sL += t; t *= rho; }
        t = NM;
        for (int k = 0; k < cntU; ++k) { sU += t; t *= rho; }
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

    // Sherman-Morrison constants.  xs = sweeps of e0 through the constant factors: xs_i = NM^i kappa (1 - rho^(n-i)).
    const double xs00 = kappa * (1.0 - ipow_sq(rho, n_int));
    ph.kap = delta_sm * rinv / (1.0 + delta_sm * (rinv * xs00));       // g = delta x_0 = kap * xs_0(uncorrected)
    // forward: rows of lane l change by -g NM^(row0 + k) = NM^(k+1) * (-g NM^(row0 - 1))
    double Dl = (lane == 0) ? 1.0 / NM : ipow_sq(NM, ly.row0 - 1);
    // backward: the carry into lane l is xs of the first row of lane l+1, which changes by -g xs there
    const int r_next = ly.row0 + ly.cnt;
    double Zl = (lane == 31) ? 0.0 : ipow_sq(NM, r_next) * kappa * (1.0 - ipow_sq(rho, n_int - r_next));
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
// HALF: 1 / 0 = theta is / is not 1/2, known at compile time (the linear time loops are compiled for both: left to a
// runtime test, ptxas if-converts the C multiplications q V of the other case into predicated DMULs that issue in every
// step -- 13 of 234 instructions of a 400-node Crank-Nicolson step); -1 = ph.half decides (American: a real branch).
template <int C, bool AMER, int TERMS, int HALF>
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
    } else if (HALF == 1 || (HALF < 0 && ph.half)) {
        // theta = 1/2: q = 1, the subtrahend is the old V itself -- a register rename, not an FP64 instruction
        // (13 of the 111 FP64 instructions of a 400-node Crank-Nicolson step)
#pragma unroll
        for (int k = 0; k < C; ++k) L[k] = V[k];
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


// ---- dividend jump (fd_american_option_pricer.py:321-386, :517-548) for the packed row layout --------------------
// Natural cubic spline through (s_j, V_j) and V(S) <- spline(S - D).  Three shared-memory arrays of n + 2 doubles
// OVERLAY the kernel's tables (a new phase rebuilds them anyway): A = knots s_j, Bv = values V_j on entry and the
// jumped values on exit, Z = the spline's c coefficients (scratch before that).  Not inlined and no register arrays
// in the signature, so the time loop's register allocation does not see it.  Returns this lane's new boundary node
// value (lanes 0 and 31).
//
// The reference solves the spline's tridiagonal system serially (:338-359).  Here every lane eliminates only its own
// rows (a first version swept 40 warm-up rows per lane and cost as much as 200 time steps per jump):
//   * mu_i = h_i / (2 (x_{i+1} - x_{i-1}) - h_{i-1} mu_{i-1}) is the pivot recurrence; an error in mu_{i-1} shrinks by
//     mu^2 ~ 0.07 per row, so GM = 14 warm-up rows of the mu recurrence alone (one division each) reproduce the serial
//     values to 0.07^14 ~ 1e-16;
//   * z_i = (alpha_i - h_{i-1} z_{i-1}) / l_i and c_j = z_j - mu_j c_{j+1} are LINEAR first-order recurrences with
//     multipliers ~ -0.27: local pass with zero carry-in, the running product of the multipliers parked in Z, then
//     true = local + product * carry, the carry being the neighbour's end value with its own two predecessors folded
//     in (a lane's product is below 0.27^9 ~ 1e-5: three terms leave 1e-16).
//   * one reciprocal of l_i serves mu_i and z_i, and 3 / h_i is carried to the next row.
//   * every division is a reciprocal and a multiplication (an IEEE division costs ~40 instructions; the first
//     compact version spent 40 % of its instructions there).
// The spline agrees with the reference's serial sweep to a few ulps (not bit for bit).  The bracket of S_j - D is found once per lane from the logarithm and then
// walked upwards (it is monotone in j) instead of one binary search per node.
template <int C>
__device__ __noinline__ double dividend_jump(const double Vb, const double cash, const bool is_call, const double k_eff,
                                             const int n, const int lane, const int row0, const int cnt, const double x_min,
                                             const double dx, double *A, double *Bv, double *Z)
{
    static_assert(3 * (32 * C + 3) <= smem_doubles<C>(), "the spline scratch must fit in the warp's tables");
    // CODE SIZE matters more than instruction count here: 16 warps share a 32 KB instruction cache with the time loop
    // they spend their life in, and a straight-line (unrolled) version of this function was 8000 instructions -- every
    // jump of any warp flushed the loop out of the cache (ncu: no_instruction became the top stall, 8 warps per issue
    // cycle).  So: runtime loops, the per-row values in small dynamically indexed local arrays.
    constexpr int GM = 14;
    const int j_first = row0 + 1;                              // this lane's interior knots j_first .. j_first + cnt - 1
    double mur[C], zr[C];
#ifdef FDM_LEAN_PROFILE
    long long jt[8]; int ji = 0;
#define LJ_TICK() jt[ji++] = clock64()
#else
#define LJ_TICK()
#endif
    LJ_TICK();
    if (lane == 0) { Bv[0] = Vb; A[0] = node_s(x_min, dx, 0); }
    if (lane == 31) { Bv[n] = Vb; A[n] = node_s(x_min, dx, n); }
    // Knots of the own rows: one exponential for the first, then s_{j+1} = s_j e^dx (16 multiplications instead of 16
    // exponentials; the knots drift by at most cnt ulps from the reference's, which moves the spline by ~1e-15
    // relative).  Where the caller marked a node as exercised (Z[j] > 0: its 6 dt lambda) the value is the payoff
    // (fd_american_option_pricer.py:497) -- the state carries vc there.  The payoff TABLE of the time loop is not
    // rebuilt from these knots: the caller keeps it in registers across the call.
    {
        const double qx = exp(dx);
        double sj = node_s(x_min, dx, j_first);
#pragma unroll 1
        for (int k = 0; k < cnt; ++k) {
            const int j = j_first + k;
            A[j] = sj;
            if (Z[j] > 0.0) Bv[j] = payoff(sj, k_eff, is_call);
            sj = __dmul_rn(sj, qx);
        }
    }
    __syncwarp();
    LJ_TICK();
    // ---- forward: warm-up of the pivot recurrence, then the lane's own rows ----------------------------------
    double mu_prev = 0.0, z_prev = 0.0, prod = 1.0;
    {
        const int i0 = max(1, j_first - GM);
        double x_im1 = A[i0 - 1], x_i = A[i0];
#pragma unroll 1
        for (int i = i0; i < j_first; ++i) {
            const double x_ip1 = A[i + 1];
            const double l = __dsub_rn(__dmul_rn(2.0, __dsub_rn(x_ip1, x_im1)), __dmul_rn(__dsub_rn(x_i, x_im1), mu_prev));
            mu_prev = __dmul_rn(__dsub_rn(x_ip1, x_i), __drcp_rn(l));
            x_im1 = x_i;
            x_i = x_ip1;
        }
        double y_im1 = Bv[j_first - 1], y_i = Bv[j_first];
        double t_im1 = 3.0 * __drcp_rn(__dsub_rn(x_i, x_im1));
#pragma unroll 1
        for (int k = 0; k < cnt; ++k) {
            const int i = j_first + k;
            const double x_ip1 = A[i + 1], y_ip1 = Bv[i + 1];
            const double h_i = __dsub_rn(x_ip1, x_i), h_im1 = __dsub_rn(x_i, x_im1);
            const double t_i = 3.0 * __drcp_rn(h_i);
            const double alpha = __dsub_rn(__dmul_rn(t_i, __dsub_rn(y_ip1, y_i)), __dmul_rn(t_im1, __dsub_rn(y_i, y_im1)));
            const double rl = __drcp_rn(__dsub_rn(__dmul_rn(2.0, __dsub_rn(x_ip1, x_im1)), __dmul_rn(h_im1, mu_prev)));
            mu_prev = __dmul_rn(h_i, rl);
            z_prev = __dmul_rn(__dsub_rn(alpha, __dmul_rn(h_im1, z_prev)), rl);
            prod = __dmul_rn(prod, -__dmul_rn(h_im1, rl));
            mur[k] = mu_prev;
            zr[k] = z_prev;
            Z[i] = prod;                                       // z_i(true) = z_i(local) + prod_i * carry
            x_im1 = x_i; x_i = x_ip1; y_im1 = y_i; y_i = y_ip1; t_im1 = t_i;
        }
    }
    LJ_TICK();
    {
        // carry into this lane = true z at the last row of the lane below
        const double e1 = __shfl_up_sync(FULL, z_prev, 1), e2 = __shfl_up_sync(FULL, z_prev, 2), e3 = __shfl_up_sync(FULL, z_prev, 3);
        const double p1 = __shfl_up_sync(FULL, prod, 1), p2 = __shfl_up_sync(FULL, prod, 2);
        double carry = (lane >= 3) ? e3 : 0.0;
        carry = (lane >= 2) ? fma(p2, carry, e2) : 0.0;
        carry = (lane >= 1) ? fma(p1, carry, e1) : 0.0;        // z_0 = 0, :346
        // ---- backward: c_j = z_j - mu_j c_{j+1}, :358-359; c_n = 0 -----------------------------------------------
        double c_next = 0.0;
        prod = 1.0;
#pragma unroll 1
        for (int k = cnt - 1; k >= 0; --k) {
            const double z_true = fma(Z[j_first + k], carry, zr[k]);
            c_next = __dsub_rn(z_true, __dmul_rn(mur[k], c_next));
            prod = __dmul_rn(prod, -mur[k]);
            zr[k] = c_next;
            Z[j_first + k] = prod;
        }
        const double f1 = __shfl_down_sync(FULL, c_next, 1), f2 = __shfl_down_sync(FULL, c_next, 2), f3 = __shfl_down_sync(FULL, c_next, 3);
        const double q1 = __shfl_down_sync(FULL, prod, 1), q2 = __shfl_down_sync(FULL, prod, 2);
        carry = (lane <= 28) ? f3 : 0.0;
        carry = (lane <= 29) ? fma(q2, carry, f2) : 0.0;
        carry = (lane <= 30) ? fma(q1, carry, f1) : 0.0;
#pragma unroll 1
        for (int k = 0; k < cnt; ++k) Z[j_first + k] = fma(Z[j_first + k], carry, zr[k]);
    }
    if (lane == 0) Z[0] = 0.0;
    if (lane == 31) Z[n] = 0.0;
    __syncwarp();
    LJ_TICK();
    // ---- evaluation at S_j - D, :517-548: nodes q = -1 (node 0, lane 0), 0 .. cnt-1 (own), cnt (node n, lane 31) -------
    const double s_lo = A[0], s_hi = A[n];
    int jj = -1;                                               // bracket of the previous node (monotone in j)
    double nb = Vb;
#pragma unroll 1
    for (int q = (lane == 0 ? -1 : 0); q < cnt + (lane == 31 ? 1 : 0); ++q) {
        const int jn = q < 0 ? 0 : (q == cnt ? n : j_first + q);
        const double Sj = A[jn];
        const double Sq = __dsub_rn(Sj, cash);
        double cont;
        if (Sq <= s_lo) cont = Bv[0];                          // :535-538
        else if (Sq >= s_hi) cont = Bv[n];
        else {
            if (jj < 0) jj = min(max((int)floor((log(Sq) - x_min) / dx), 0), n - 1);
            while (jj + 1 <= n - 1 && A[jj + 1] <= Sq) ++jj;   // the reference's binary search (:375-382) ends on the
            while (jj > 0 && A[jj] > Sq) --jj;                 // knot interval with x_j <= Sq < x_{j+1}
            const int j = jj;
            const double xj = A[j], h = __dsub_rn(A[j + 1], xj);
            const double cj = Z[j], cj1 = Z[j + 1];
            const double yj = Bv[j];
            const double rh = __drcp_rn(h), third = 1.0 / 3.0;  // one reciprocal for the three divisions of :360-364
            const double bj = __dsub_rn(__dmul_rn(__dsub_rn(Bv[j + 1], yj), rh),
                                        __dmul_rn(__dmul_rn(h, __dadd_rn(cj1, __dmul_rn(2.0, cj))), third));
            const double dj = __dmul_rn(__dmul_rn(__dsub_rn(cj1, cj), rh), third);
            const double dq = __dsub_rn(Sq, xj);
            double acc = __dadd_rn(yj, __dmul_rn(bj, dq));                                          // :384, left to right
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(cj, dq), dq));
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(dj, dq), dq), dq));
            cont = acc;
        }
        if (is_call) {
            const double pv = payoff(Sj, k_eff, true);
            cont = cont > pv ? cont : pv;                      // :542-544
        }
        if (q < 0 || q == cnt) nb = cont; else zr[q] = cont;
    }
    __syncwarp();                                              // every lane has read the old values
#pragma unroll 1
    for (int k = 0; k < cnt; ++k) Bv[j_first + k] = zr[k];
    __syncwarp();
    LJ_TICK();
#ifdef FDM_LEAN_PROFILE
    if (blockIdx.x == 1000 && lane == 5)
        printf("lean jump: knots %lld | elimination %lld | back-substitution %lld | evaluation %lld\n", jt[1] - jt[0], jt[2] - jt[1],
               jt[3] - jt[2], jt[4] - jt[3]);
#endif
    return nb;
}

#ifndef FDM_LEAN_MINB
#define FDM_LEAN_MINB 0
#endif

// -DFDM_LEAN_PROFILE: per-region cycle counts of one warp (dev tool; printf from block 1000)
#ifdef FDM_LEAN_PROFILE
#define LEAN_TICK(slot) do { const long long t_ = clock64(); prof_acc[slot] += t_ - prof_t; prof_t = t_; } while (0)
#else
#define LEAN_TICK(slot)
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
    bool retry = false;
#ifdef FDM_LEAN_PROFILE
    long long prof_acc[5] = {0, 0, 0, 0, 0};
    long long prof_t = clock64();
#endif

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
            LEAN_TICK(0);
            const bool phase_ok = setup_phase<C, AMER>(implicit_phase ? 1.0 : 0.5, dt, ca, cc, bcoef, n_int, lane, ly, T, ph);
            LEAN_TICK(1);
            if (!phase_ok) {
                retry = true;
                break;
            }
            double w_next = fetch_boundary(m);
            double cin = LCV(ph, T, lane, cB, LC_CB) * Vb - LCV(ph, T, lane, cA, LC_CA) * w_next;   // non-zero on lanes 0 and 31 only
            // The time loop is compiled once per carry form (TT = 2 / 4 / 6 direct terms, 0 = scan), so the step has no
            // branches on ph.terms: ~35 fewer instructions per warp-step (BRA / UISETP / PLOP3 and the shuffle
            // temporaries' moves) and a tighter schedule.  Measured on B200: +5..14 % on every lean shape.
            auto time_loop = [&](auto tt, auto hh) {
            constexpr int TT = decltype(tt)::value, HH = decltype(hh)::value;
            for (; m < m_end; ++m) {
                const double w_now = w_next;
                w_next = fetch_boundary(m + 1);
                double cin_n = LCV(ph, T, lane, cB, LC_CB) * w_now - LCV(ph, T, lane, cA, LC_CA) * w_next;
                march_step<C, AMER, TT, HH>(V, L, ph, (lane == 0) ? cin : 0.0, (lane == 31) ? cin : 0.0, last, T, lane);
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
            auto by_terms = [&](auto hh) {
                switch (ph.terms) {
                case 2: time_loop(IntC<2>{}, hh); break;
                case 4: time_loop(IntC<4>{}, hh); break;
                case 6: time_loop(IntC<6>{}, hh); break;
                default: time_loop(IntC<0>{}, hh); break;
                }
            };
            if (AMER) by_terms(IntC<-1>{});
            else if (ph.half) by_terms(IntC<1>{});
            else by_terms(IntC<0>{});
        }
        if (retry) break;
        const double cash = B.seg_div[seg];
        if (cash != 0.0) {
            // ex-dividend jump.  The spline's scratch overlays this warp's tables: the next phase rebuilds the carry
            // tables, and the payoff table and the boundary inputs are refilled here.
            LEAN_TICK(2);
            double *A = smem, *Bv = A + (n + 2), *Zs = Bv + (n + 2);
            __syncwarp();
            // Zs lies beyond the payoff table: park the exercise marks (6 dt lambda > 0) there first, then take the exact
            // payoff table into the registers of L, and only then let the values overwrite its place
#pragma unroll
            for (int k = 0; k < C; ++k) if (k < ly.cnt) Zs[ly.row0 + k + 1] = AMER ? L[k] : 0.0;
#pragma unroll
            for (int k = 0; k < C; ++k) L[k] = T.phi3[k * 32 + lane];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < C; ++k) if (k < ly.cnt) Bv[ly.row0 + k + 1] = V[k];
            LEAN_TICK(3);
            Vb = dividend_jump<C>(Vb, cash, is_call, k_eff, n, lane, ly.row0, ly.cnt, x_min, dx, A, Bv, Zs);
            LEAN_TICK(4);
#pragma unroll
            for (int k = 0; k < C; ++k) V[k] = (k < ly.cnt) ? Bv[ly.row0 + k + 1] : 0.0;
            __syncwarp();
#pragma unroll
            for (int k = 0; k < C; ++k) {
                T.phi3[k * 32 + lane] = L[k];
                L[k] = 0.0;
            }
            if (lane == 0) {
                T.par[0] = k_eff;
                T.par[1] = (flags & FDM_FLAG_BDRY_TOP_NEXT) ? node_s(x_min, dx, n + 1) : sn;
                T.par[2] = s0;
            }
            __syncwarp();
        }
    }
    if (retry) {
        if (lane == 0) O.price[sid] = __longlong_as_double((long long)RETRY_BITS);
        return;
    }

    // ---- read-out (as fdm_warp_kernel) ------------------------------------------------------------
    LEAN_TICK(2);
#ifdef FDM_LEAN_PROFILE
    if (blockIdx.x == 1000 && lane == 0)
        printf("lean<%d> steps=%d cycles: prologue+loops %lld | setup_phase %lld | time loops %lld | materialise %lld | jump %lld\n",
               C, gstep, prof_acc[0], prof_acc[1], prof_acc[2], prof_acc[3], prof_acc[4]);
#endif
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

