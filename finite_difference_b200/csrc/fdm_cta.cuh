// fdm_cta.cuh -- one CTA per PDE solve, for grids too large for one warp's registers
// (514 .. 8193 space nodes; BASELINE config 4 is 8192 nodes x 10 000 steps).
//
// Same step algebra as fdm_warp_kernel (see the header of fdm_b200.cu), different residency:
//   * thread t of the CTA owns the C contiguous interior rows [t*C, (t+1)*C); V and dt*lambda stay
//     in registers, the payoff lives in shared memory ([k][t] layout, conflict-free);
//   * the LU pivots of the Toeplitz matrix converge bit-wise after a few hundred rows at most, so
//     only the warps that contain the first P rows (or the top boundary row / padding) read
//     per-row coefficients from a small shared-memory table; every other warp runs with three
//     scalar coefficients.  The choice is warp-uniform: no divergence;
//   * each first-order recurrence runs as  local pass -> intra-warp shuffle scan -> one
//     __syncthreads -> short Horner sum over the preceding warps' totals -> local pass.
//     The warp-to-warp multipliers decay like |c'|^(32 C) (1e-17 at C = 16), so the Horner sum is
//     truncated after K terms with the neglected weight below 1e-22 (K is found per phase).
//   Two barriers per time step in total.
//
// Included by fdm_b200.cu (single translation unit); uses its helpers.

#pragma once

namespace cta {

constexpr double TRUNC_EPS = 1e-22;

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

struct Tables {
    // coefficient tables of one (theta, dt) phase, index = row class:
    //   [0, P)  head rows (pivots still changing), [P] steady state, [P+1] the top real row
    //   (ncc = 1 so the Dirichlet carry passes), [P+2] padding rows above the grid (inert).
    double *nm, *ncc, *inv;
    int P, n_int;
    __device__ __forceinline__ int cls(int row) const
    {
        return row >= n_int ? P + 2 : (row == n_int - 1 ? P + 1 : min(row, P));
    }
};

struct Phase {
    double p, q, pd;
    double cAf, cBf, cAb, cBb;       // boundary carry coefficients (thread 0 / last thread)
    double NM, NCC, INV;             // steady-state coefficients
    double aF, aB;                   // product of the chunk's multipliers (forward / backward)
    double hF1, hF2, hB1, hB2;       // half-chunk products
    double PAf, PAb;                 // product over the lanes before (fwd) / after (bwd) this lane in the warp
    int Kf, Kb;                      // terms of the warp-to-warp Horner sums
    bool tbl;                        // this warp reads per-row coefficients from the tables
};

template <int C, int WARPS, bool AMER>
__device__ __forceinline__ void setup_phase(double theta, double dt, double a, double c, double bcoef,
                                            int n_int, double *tab_base, int tab_cap, double *part,
                                            int *status_flag, Phase &ph, Tables &T)
{
    constexpr int NT = 32 * WARPS;
    constexpr int H = C / 2;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const double A_L = __dmul_rn(__dmul_rn(-theta, dt), a);
    const double A_C = __dsub_rn(1.0, __dmul_rn(__dmul_rn(theta, dt), bcoef));
    const double A_U = __dmul_rn(__dmul_rn(-theta, dt), c);
    const double B_L = __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), a);
    const double B_U = __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), c);
    ph.p = 1.0 / theta;
    ph.q = (1.0 - theta) / theta;
    const double B_C = __dadd_rn(1.0, __dmul_rn(__dmul_rn(__dsub_rn(1.0, theta), dt), bcoef));

    const int cap = tab_cap / 3 - 3;               // head rows the tables can hold
    T.nm = tab_base;
    T.ncc = tab_base + (cap + 3);
    T.inv = tab_base + 2 * (cap + 3);
    T.n_int = n_int;
    int *ishare = (int *)(part + 4 * WARPS);
    __syncthreads();
    if (t == 0) {
        // pivots: den_0 = A_C, den_i = A_C - A_L*A_U/den_{i-1} (fd_american_option_pricer.py:443-451)
        double r_prev = 0.0, r_cur = __drcp_rn(A_C);
        int P = -1;
        for (int i = 0; i < cap; ++i) {
            const double den_next = __dsub_rn(A_C, __dmul_rn(A_L, __dmul_rn(A_U, r_cur)));
            const double r_next = __drcp_rn(den_next);
            T.inv[i] = r_cur;
            T.nm[i] = (i == 0) ? 1.0 : -(A_L * r_prev);
            T.ncc[i] = -(A_U * r_next);
            if (i > 0 && r_cur == r_prev && r_next == r_cur) { P = i; break; }   // row i is already steady
            r_prev = r_cur;
            r_cur = r_next;
        }
        if (P < 0) { P = cap; *status_flag = 2; }   // pivots not converged within the table: flagged
        // steady state = entry P (written by the loop when it broke at i == P)
        if (P == cap) {
            T.inv[P] = r_cur; T.nm[P] = -(A_L * r_cur); T.ncc[P] = -(A_U * r_cur);
        }
        const int top = n_int - 1;
        const int tc = min(top, P);
        T.inv[P + 1] = T.inv[tc]; T.nm[P + 1] = T.nm[tc]; T.ncc[P + 1] = 1.0;
        T.inv[P + 2] = 0.0; T.nm[P + 2] = 0.0; T.ncc[P + 2] = 1.0;
        // growth-factor matching (fdm_b200.cu, growth_matched_pd); the linear march folds p + delta into inv
        const double pd = (P < cap && P < n_int) ? growth_matched_pd(A_L, A_C, A_U, B_L, B_C, B_U, ph.q, T.inv[P], n_int) : 1.0;
        part[4 * WARPS + 4] = pd;
        if (!AMER) {
            const double scale = ph.p + (pd - 1.0);
            for (int i = 0; i <= P + 1; ++i) T.inv[i] *= scale;
        }
        ishare[0] = P;
    }
    __syncthreads();
    T.P = ishare[0];
    const int P = T.P;
    ph.pd = part[4 * WARPS + 4];
    const double bscale = AMER ? 1.0 : 1.0 / (ph.p + (ph.pd - 1.0));   // the linear march solves for R / (p + delta)
    ph.cAf = A_L * bscale; ph.cBf = B_L * bscale;
    ph.cAb = A_U * bscale; ph.cBb = B_U * bscale;
    ph.NM = T.nm[P]; ph.NCC = T.ncc[P]; ph.INV = T.inv[P];
    const int wfirst = w * 32 * C, wlast = wfirst + 32 * C - 1;
    ph.tbl = (wfirst < P) || (wlast >= n_int - 1);
#ifdef FDM_CTA_FORCE_CONST
    ph.tbl = false;     // timing experiment only: wrong results
#endif

    double f1 = 1.0, f2 = 1.0, b1 = 1.0, b2 = 1.0;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int ci = T.cls(t * C + k);
        const double vnm = ph.tbl ? T.nm[ci] : ph.NM;
        const double vncc = ph.tbl ? T.ncc[ci] : ph.NCC;
        if (k < H) { f1 *= vnm; b2 *= vncc; } else { f2 *= vnm; b1 *= vncc; }
    }
    ph.hF1 = f1; ph.hF2 = f2; ph.hB1 = b1; ph.hB2 = b2;
    ph.aF = f1 * f2; ph.aB = b1 * b2;
    // exclusive products across the warp (forward: lanes below, backward: lanes above)
    double pf = ph.aF, pb = ph.aB;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double tf = __shfl_up_sync(FULL, pf, d);
        const double tb = __shfl_down_sync(FULL, pb, d);
        if (lane >= d) pf *= tf;
        if (lane + d < 32) pb *= tb;
    }
    // pf / pb are inclusive; the warp totals go to shared memory, the exclusive ones stay in registers
    if (lane == 31) part[1 * WARPS + w] = pf;      // pA_f
    if (lane == 0) part[3 * WARPS + w] = pb;       // pA_b
    double ef = __shfl_up_sync(FULL, pf, 1), eb = __shfl_down_sync(FULL, pb, 1);
    ph.PAf = (lane == 0) ? 1.0 : ef;
    ph.PAb = (lane == 31) ? 1.0 : eb;
    __syncthreads();
    // terms of the warp-to-warp sums: smallest K whose neglected weight is below TRUNC_EPS for every warp
    int Kf = WARPS - 1, Kb = WARPS - 1;
    for (int K = 1; K < WARPS; ++K) {
        double worst_f = 0.0, worst_b = 0.0;
        for (int s = 0; s + K <= WARPS; ++s) {
            double qf = 1.0, qb = 1.0;
            for (int u = s; u < s + K; ++u) { qf *= part[1 * WARPS + u]; qb *= part[3 * WARPS + u]; }
            worst_f = fmax(worst_f, fabs(qf));
            worst_b = fmax(worst_b, fabs(qb));
        }
        if (worst_f < TRUNC_EPS && Kf == WARPS - 1) Kf = K;
        if (worst_b < TRUNC_EPS && Kb == WARPS - 1) Kb = K;
    }
    ph.Kf = Kf; ph.Kb = Kb;
    (void)NT;
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

// ---- the three phases of one time step (barriers sit between them, outside any branch) -----------

template <int C, int WARPS, bool AMER, bool TBL>
__device__ __forceinline__ void step_a(double (&V)[C], double (&L)[C], const Phase &ph, const Tables &T,
                                       const double cin_f, double *part, double &e1_keep, double &Eprev)
{
    constexpr int H = C / 2;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        if (AMER) {
            const double g = fma(ph.q, V[k], L[k]);
            V[k] = fma(ph.pd, V[k], g);
            L[k] = g;
        } else {
            L[k] = ph.q * V[k];
        }
    }
    double e1 = cin_f, e2 = 0.0;
#pragma unroll
    for (int k = 0; k < H; ++k) e1 = fma(TBL ? T.nm[T.cls(t * C + k)] : ph.NM, e1, V[k]);
#pragma unroll
    for (int k = H; k < C; ++k) e2 = fma(TBL ? T.nm[T.cls(t * C + k)] : ph.NM, e2, V[k]);
    const double e_end = fma(ph.hF2, e1, e2);
    const double E = warp_scan<TBL, true>(e_end, ph.aF, lane);
    if (lane == 31) part[0 * WARPS + w] = E;          // pE_f
    const double up = __shfl_up_sync(FULL, E, 1);
    Eprev = (lane == 0) ? 0.0 : up;
    e1_keep = e1;
}

template <int C, int WARPS, bool AMER, bool TBL>
__device__ __forceinline__ void step_b(double (&V)[C], const Phase &ph, const Tables &T, const double cin_f,
                                       const double cin_b, double *part, const double e1_fwd, const double Eprev,
                                       double &e1_keep, double &Enext)
{
    constexpr int H = C / 2;
    constexpr int NT = 32 * WARPS;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    // carry into this warp from the warps below (truncated Horner sum)
    double S = 0.0;
    for (int v = max(0, w - ph.Kf); v < w; ++v) S = fma(part[1 * WARPS + v], S, part[0 * WARPS + v]);
    const double cs = fma(ph.PAf, S, Eprev);           // carry into this thread's chunk, excluding cin_f
    {
        double c1 = cs + cin_f;
        double c2 = fma(ph.hF1, cs, e1_fwd);
#pragma unroll
        for (int k = 0; k < H; ++k) { c1 = fma(TBL ? T.nm[T.cls(t * C + k)] : ph.NM, c1, V[k]); V[k] = c1; }
#pragma unroll
        for (int k = H; k < C; ++k) { c2 = fma(TBL ? T.nm[T.cls(t * C + k)] : ph.NM, c2, V[k]); V[k] = c2; }
    }
    // back substitution, local pass: upper half (k = C-1..H) then lower half (H-1..0)
    double e1 = cin_b, e2 = 0.0;
#pragma unroll
    for (int k = C - 1; k >= H; --k) e1 = fma(TBL ? T.ncc[T.cls(t * C + k)] : ph.NCC, e1, V[k]);
#pragma unroll
    for (int k = H - 1; k >= 0; --k) e2 = fma(TBL ? T.ncc[T.cls(t * C + k)] : ph.NCC, e2, V[k]);
    const double e_end = fma(ph.hB2, e1, e2);
    const double E = warp_scan<TBL, false>(e_end, ph.aB, lane);
    if (lane == 0) part[2 * WARPS + w] = E;           // pE_b
    const double dn = __shfl_down_sync(FULL, E, 1);
    Enext = (lane == 31) ? 0.0 : dn;
    e1_keep = e1;
    (void)NT;
}

template <int C, int WARPS, bool AMER, bool TBL>
__device__ __forceinline__ void step_c(double (&V)[C], double (&L)[C], const double *__restrict__ phi_s,
                                       const Phase &ph, const Tables &T, const double cin_b, double *part,
                                       const double e1_bwd, const double Enext)
{
    constexpr int H = C / 2;
    constexpr int NT = 32 * WARPS;
    const int t = threadIdx.x, w = t >> 5;
    double S = 0.0;
    for (int v = min(WARPS - 1, w + ph.Kb); v > w; --v) S = fma(part[3 * WARPS + v], S, part[2 * WARPS + v]);
    const double cs = fma(ph.PAb, S, Enext);
    double c1 = cs + cin_b;
    double c2 = fma(ph.hB1, cs, e1_bwd);
#pragma unroll
    for (int i = 0; i < H; ++i) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int k = half == 0 ? C - 1 - i : H - 1 - i;
            const int ci = TBL ? T.cls(t * C + k) : 0;
            double xs;
            if (half == 0) { c1 = fma(TBL ? T.ncc[ci] : ph.NCC, c1, V[k]); xs = c1; }
            else           { c2 = fma(TBL ? T.ncc[ci] : ph.NCC, c2, V[k]); xs = c2; }
            const double vc = fma(TBL ? T.inv[ci] : ph.INV, xs, -L[k]);
            if (AMER) {
                // exercise iff phi > vc (:496): dt*lambda_new = max(phi - vc, 0) on the sign mask, V = vc + dt*lambda_new
                const double d = phi_s[k * NT + t] - vc;
                const int msk = ~(__double2hiint(d) >> 31);
                L[k] = __hiloint2double(__double2hiint(d) & msk, __double2loint(d) & msk);
                V[k] = vc + L[k];
            } else {
                V[k] = vc;
            }
        }
    }
}

// ---- dividend jump (fd_american_option_pricer.py:321-386, :517-548) ------------------------------
// ys, mu, cz: shared arrays of n+1 doubles.  The knot abscissae are staged in mu[] and overwritten
// in place by the forward sweep; thread 0 runs the two serial sweeps in the reference's order.
template <int C, int WARPS>
__device__ __forceinline__ void dividend_jump(double (&V)[C], double &Vb, double cash, bool is_call, double k_eff,
                                              int n, double x_min, double dx, double *ys, double *mu, double *cz)
{
    constexpr int NT = 32 * WARPS;
    const int t = threadIdx.x;
    const int nn = n + 1;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = t * C + k + 1;
        if (j <= n - 1) { ys[j] = V[k]; mu[j] = node_s(x_min, dx, j); }
    }
    if (t == 0) { ys[0] = Vb; mu[0] = node_s(x_min, dx, 0); }
    if (t == NT - 1) { ys[n] = Vb; mu[n] = node_s(x_min, dx, n); }
    __syncthreads();
    if (t == 0) {
        double x_im1 = mu[0], x_i = mu[1];
        double mu_prev = 0.0, z_prev = 0.0;
        mu[0] = 0.0;
        cz[0] = 0.0;
        for (int i = 1; i < nn - 1; ++i) {
            const double x_ip1 = mu[i + 1];
            const double h_i = __dsub_rn(x_ip1, x_i);
            const double h_im1 = __dsub_rn(x_i, x_im1);
            const double alpha = __dsub_rn(__dmul_rn(3.0 / h_i, __dsub_rn(ys[i + 1], ys[i])),
                                           __dmul_rn(3.0 / h_im1, __dsub_rn(ys[i], ys[i - 1])));
            const double l = __dsub_rn(__dmul_rn(2.0, __dsub_rn(x_ip1, x_im1)), __dmul_rn(h_im1, mu_prev));
            mu_prev = h_i / l;
            z_prev = __dsub_rn(alpha, __dmul_rn(h_im1, z_prev)) / l;
            mu[i] = mu_prev;
            cz[i] = z_prev;
            x_im1 = x_i;
            x_i = x_ip1;
        }
        double c_next = 0.0;
        cz[nn - 1] = 0.0;
        for (int j = nn - 2; j >= 0; --j) {
            c_next = __dsub_rn(cz[j], __dmul_rn(mu[j], c_next));
            cz[j] = c_next;
        }
    }
    __syncthreads();
    const double s_lo = node_s(x_min, dx, 0), s_hi = node_s(x_min, dx, n);
    auto eval = [&](int jnode) -> double {
        const double Sj = node_s(x_min, dx, jnode);
        const double Sq = __dsub_rn(Sj, cash);
        double cont;
        if (Sq <= s_lo) cont = ys[0];
        else if (Sq >= s_hi) cont = ys[n];
        else {
            int j = last_node_le(x_min, dx, n, Sq);           // == the reference's binary search (:375-382)
            j = min(max(j, 0), nn - 2);
            const double xj = node_s(x_min, dx, j), xj1 = node_s(x_min, dx, j + 1);
            const double h = __dsub_rn(xj1, xj);
            const double cj = cz[j], cj1 = cz[j + 1];
            const double bj = __dsub_rn(__dsub_rn(ys[j + 1], ys[j]) / h,
                                        __dmul_rn(h, __dadd_rn(cj1, __dmul_rn(2.0, cj))) / 3.0);
            const double dj = __dsub_rn(cj1, cj) / __dmul_rn(3.0, h);
            const double dq = __dsub_rn(Sq, xj);
            double acc = __dadd_rn(ys[j], __dmul_rn(bj, dq));
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
    double newV[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = t * C + k + 1;
        newV[k] = (j <= n - 1) ? eval(j) : 0.0;
    }
    double nb = Vb;
    if (t == 0) nb = eval(0);
    if (t == NT - 1) nb = eval(n);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < C; ++k) V[k] = newV[k];
    Vb = nb;
}

// ---- the kernel ------------------------------------------------------------------------------------

template <int C, int WARPS>
constexpr int smem_doubles(bool with_div)
{
    return (with_div ? 3 : 2) * (32 * WARPS * C + 2) + 4 * WARPS + 16;
}

template <int C, int WARPS, bool AMER, bool TRACE>
__global__ void __launch_bounds__(32 * WARPS, 1)
fdm_cta_kernel(const FdmBatch B, const FdmOut O, const int64_t first, const int want_american, const int div_smem,
               const int n_lo)
{
    constexpr int NT = 32 * WARPS;
    constexpr int NP = NT * C + 2;
    extern __shared__ double smem[];
    double *r0 = smem;                 // payoff [k*NT + t]          | dividend jump: cz
    double *r1 = r0 + NP;              // coefficient tables         | dividend jump: ys; read-out: V by node
    double *r2 = r1 + NP;              // dividend jump: mu (only allocated when div_smem)
    double *part = div_smem ? r2 + NP : r1 + NP;     // pE_f, pA_f, pE_b, pA_b [WARPS] + scratch
    int *ishare = (int *)(part + 4 * WARPS);         // [0] P, [1..4] snapping / KO indices, [5] status, [6,7] ex lo/hi

    const int64_t sid = first + blockIdx.x;
    if (sid >= B.n_solves) return;
    const int flags = B.flags[sid];
    if (((flags & FDM_FLAG_AMERICAN) != 0) != (want_american != 0)) return;
    const int n = B.n_space[sid];
    if (n <= n_lo || n - 1 > NT * C) return;           // other size classes are other launches
    const int n_int = n - 1;

    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const bool is_call = flags & FDM_FLAG_CALL;
    const bool b76 = flags & FDM_FLAG_BLACK76;
    const bool snap = flags & FDM_FLAG_SNAP;
    const int ko = (flags & FDM_FLAG_KO_DOWN) ? 1 : ((flags & FDM_FLAG_KO_UP) ? 2 : 0);
    const double x_min = B.x_min[sid], dx = B.dx[sid], sigma = B.sigma[sid], r = B.r[sid], bcar = B.b[sid];
    const double spot = B.spot[sid], strike = B.strike[sid];
    const double barrier = (ko && B.barrier) ? B.barrier[sid] : 0.0;
    const double rebate = (ko && B.rebate) ? B.rebate[sid] : 0.0;
    const int rannacher = B.rannacher[sid];

    if (t == 0) {
        int spot_idx = 0, strike_idx = 0, klo = 1, khi = 0;
        if (snap) {
            spot_idx = nearest_node(x_min, dx, n, spot, false);
            strike_idx = nearest_node(x_min, dx, n, strike, false);
        }
        if (ko == 1) { klo = 0; khi = last_node_le(x_min, dx, n, barrier); }
        if (ko == 2) { klo = first_node_ge(x_min, dx, n, barrier); khi = n; }
        ishare[1] = spot_idx; ishare[2] = strike_idx; ishare[3] = klo; ishare[4] = khi;
        ishare[5] = 0; ishare[6] = 0x7fffffff; ishare[7] = -1;
    }
    __syncthreads();
    const int spot_idx = ishare[1], strike_idx = ishare[2], ko_lo = ishare[3], ko_hi = ishare[4];
    const double k_eff = snap ? node_s(x_min, dx, strike_idx) : strike;
    const double sn = node_s(x_min, dx, n);

    double V[C], L[C];
    auto fill_payoff = [&](bool set_v) {
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int j = t * C + k + 1;
            const double pv = (j <= n - 1) ? payoff(node_s(x_min, dx, j), k_eff, is_call) : 0.0;
            r0[k * NT + t] = pv;
            if (set_v) { V[k] = pv; L[k] = 0.0; }
        }
    };
    fill_payoff(true);
    // boundary node values: V[0] on thread 0, V[n] on the last thread
    double Vb = (t == 0) ? payoff(node_s(x_min, dx, 0), k_eff, is_call)
                         : ((t == NT - 1) ? payoff(sn, k_eff, is_call) : 0.0);
    const bool bdry_thread = is_call ? (t == NT - 1) : (t == 0);
    const bool bdry_warp = is_call ? (w == WARPS - 1) : (w == 0);

    const double sig2 = __dmul_rn(sigma, sigma);
    const double mu_x = b76 ? __dmul_rn(-0.5, sig2) : __dsub_rn(bcar, __dmul_rn(0.5, sig2));
    const double alpha = __dmul_rn(0.5, sig2) / __dmul_rn(dx, dx);
    const double beta = mu_x / __dmul_rn(2.0, dx);
    const double ca = __dsub_rn(alpha, beta);
    const double cc = __dadd_rn(alpha, beta);
    const double bcoef = __dsub_rn(__dmul_rn(-2.0, alpha), r);

    const int seg_b = B.seg_begin[sid], seg_e = B.seg_begin[sid + 1];
    int mon_i = 0, mon_e = 0;
    if (ko && B.mon_begin) { mon_i = B.mon_begin[sid]; mon_e = B.mon_begin[sid + 1]; }
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
        double my_w = 0.0;
        int fill_base = -64;
        // moving Dirichlet value: computed 32 steps at a time by the warp that owns that boundary
        auto fetch_boundary = [&](int step) -> double {
            if (!bdry_warp) return 0.0;
            if (step - fill_base >= 32 || step < fill_base) {
                fill_base = step;
                const double tau = fma((double)(step + lane + 1), dt, tau0);
                double vmin, vmax;
                boundary_values(is_call, b76, r, bcar, k_eff, sn, tau, vmin, vmax);
                my_w = is_call ? vmax : vmin;
            }
            const double wv = __shfl_sync(FULL, my_w, step - fill_base);
            return bdry_thread ? wv : 0.0;
        };
        while (m < n_steps) {
            const bool implicit_phase = m < base_ran;
            const int m_end = implicit_phase ? min(base_ran, n_steps) : n_steps;
            Phase ph;
            Tables T;
            setup_phase<C, WARPS, AMER>(implicit_phase ? 1.0 : 0.5, dt, ca, cc, bcoef, n_int, r1, NP, part,
                                        &ishare[5], ph, T);
            const double cA = (t == 0) ? ph.cAf : ((t == NT - 1) ? ph.cAb : 0.0);
            const double cB = (t == 0) ? ph.cBf : ((t == NT - 1) ? ph.cBb : 0.0);
            double w_next = fetch_boundary(m);
            double cin = cB * Vb - cA * w_next;                 // non-zero on the two boundary threads only
            for (; m < m_end; ++m) {
                const double w_now = w_next;
                w_next = fetch_boundary(m + 1);
                double cin_n = cB * w_now - cA * w_next;
                const double cin_f = (t == 0) ? cin : 0.0;
                const double cin_b = (t == NT - 1) ? cin : 0.0;
                double e1f, Eprev, e1b, Enext;
                if (ph.tbl) step_a<C, WARPS, AMER, true>(V, L, ph, T, cin_f, part, e1f, Eprev);
                else        step_a<C, WARPS, AMER, false>(V, L, ph, T, cin_f, part, e1f, Eprev);
                __syncthreads();
                if (ph.tbl) step_b<C, WARPS, AMER, true>(V, ph, T, cin_f, cin_b, part, e1f, Eprev, e1b, Enext);
                else        step_b<C, WARPS, AMER, false>(V, ph, T, cin_f, cin_b, part, e1f, Eprev, e1b, Enext);
                __syncthreads();
                if (ph.tbl) step_c<C, WARPS, AMER, true>(V, L, r0, ph, T, cin_b, part, e1b, Enext);
                else        step_c<C, WARPS, AMER, false>(V, L, r0, ph, T, cin_b, part, e1b, Enext);
                Vb = w_now;                                     // V[0] = V_min, V[-1] = V_max (:505-506)
                ++gstep;
                if (TRACE && trace) {
                    int lo = 0x7fffffff, hi = -1;
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        const int row = t * C + k;
                        if (AMER && row < n_int && L[k] > 0.0) { lo = min(lo, row); hi = max(hi, row); }
                    }
                    __syncthreads();
                    if (t == 0) { ishare[6] = 0x7fffffff; ishare[7] = -1; }
                    __syncthreads();
                    if (hi >= 0) { atomicMin(&ishare[6], lo); atomicMax(&ishare[7], hi); }
                    __syncthreads();
                    if (t == 0 && 2 * (int64_t)gstep <= O.ex_trace_stride) {
                        trace[2 * (gstep - 1)] = (ishare[7] < 0) ? -1 : ishare[6];
                        trace[2 * (gstep - 1) + 1] = ishare[7];
                    }
                }
                if (gstep == next_mon) {                        // discrete_barrier_fdm_pricer_cn.py:299-300
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        const int j = t * C + k + 1;
                        if (j <= n - 1 && j >= ko_lo && j <= ko_hi) V[k] = rebate;
                    }
                    if (t == 0 && ko_lo <= 0 && 0 <= ko_hi) Vb = rebate;
                    if (t == NT - 1 && ko_lo <= n && n <= ko_hi) Vb = rebate;
                    ++mon_i;
                    next_mon = (mon_i < mon_e) ? B.mon_step[mon_i] : -1;
                    cin_n = cB * Vb - cA * w_next;
                }
                cin = cin_n;
            }
        }
        if (cash != 0.0) {
            if (div_smem) {
                dividend_jump<C, WARPS>(V, Vb, cash, is_call, k_eff, n, x_min, dx, r1, r2, r0);
                __syncthreads();
                fill_payoff(false);                             // r0 held cz during the jump
            } else {
#pragma unroll
                for (int k = 0; k < C; ++k) V[k] = NAN;
            }
        }
    }

    // ---- read-out: stage the value vector by node in shared memory, thread 0 finishes -------------
    __syncthreads();
    double *Vs = r1;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = t * C + k + 1;
        if (j <= n - 1) Vs[j] = V[k];
    }
    if (t == 0) { Vs[0] = Vb; ishare[6] = 0x7fffffff; ishare[7] = -1; }
    if (t == NT - 1) Vs[n] = Vb;
    __syncthreads();
    if (AMER) {
        int lo = 0x7fffffff, hi = -1;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int row = t * C + k;
            if (row < n_int && L[k] > 0.0) { lo = min(lo, row); hi = max(hi, row); }
        }
        if (hi >= 0) { atomicMin(&ishare[6], lo); atomicMax(&ishare[7], hi); }
    }
    if (O.v_dump) {
        double *dst = O.v_dump + sid * O.v_dump_stride;
        for (int j = t; j <= n && j < O.v_dump_stride; j += NT) dst[j] = Vs[j];
    }
    __syncthreads();
    if (t != 0) return;

    const int ex_lo = (AMER && ishare[7] >= 0) ? ishare[6] : -1;
    const int ex_hi = (AMER && ishare[7] >= 0) ? ishare[7] : -1;
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
    if (O.status) O.status[sid] = ishare[5] ? ishare[5] : (isfinite(price) ? 0 : 1);
}

}  // namespace cta
