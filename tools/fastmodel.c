/* Dev tool (not product, not oracle): serial CPU model of the arithmetic the CUDA kernel uses,
 * to measure its rounding distance from the oracle before spending GPU time.
 *   y = A^-1 (p V + L + bdry) ; vc = inv*xs - (q V + L) ; d = phi - vc ; V' = max(phi,vc) ; L' = max(d,0)
 * with p = 1/theta, q = (1-theta)/theta, L = dt*lambda, exact-LU coefficients m_i, cc_i, inv_i. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

int fastmodel_solve(double x_min, double dx, int n, double sigma, double r, double b, int model,
                    int is_call, int american, int rannacher, int M, double T,
                    double spot, double strike, int snap,
                    int ko_kind, double barrier, double rebate, const int *mon, int n_mon,
                    double *V_out /* n+1 */, int *ex_lo, int *ex_hi)
{
    int ni = n - 1;
    double *s = malloc((n + 1) * sizeof(double)), *phi = malloc((n + 1) * sizeof(double));
    double *V = malloc(ni * sizeof(double)), *L = calloc(ni, sizeof(double));
    double *m = malloc(ni * sizeof(double)), *cc = malloc(ni * sizeof(double)), *inv = malloc(ni * sizeof(double));
    double *G = malloc(ni * sizeof(double));
    for (int i = 0; i <= n; ++i) s[i] = exp(x_min + (double)i * dx);
    double k = strike;
    if (snap) {
        int best = 0; double bv = fabs(s[0] - strike);
        for (int i = 1; i <= n; ++i) { double v = fabs(s[i] - strike); if (v < bv) { bv = v; best = i; } }
        k = s[best];
    }
    for (int i = 0; i <= n; ++i) { double v = is_call ? s[i] - k : k - s[i]; phi[i] = v > 0 ? v : 0; }
    for (int i = 0; i < ni; ++i) V[i] = phi[i + 1];
    double V0 = phi[0], VN = phi[n];
    double dt = T / M;
    double sig2 = sigma * sigma;
    double mu = (model == 1) ? -0.5 * sig2 : (b - 0.0) - 0.5 * sig2;
    double alpha = 0.5 * sig2 / (dx * dx), beta = mu / (2.0 * dx);
    double a = alpha - beta, c = alpha + beta, bc = -2.0 * alpha - r;
    int mi = 0;
    double cur_theta = -1;
    double A_L = 0, A_C = 0, A_U = 0, B_L = 0, B_U = 0, p = 0, q = 0;
    for (int step = 0; step < M; ++step) {
        double theta = (step < rannacher) ? 1.0 : 0.5;
        if (theta != cur_theta) {
            cur_theta = theta;
            A_L = -theta * dt * a; A_C = 1.0 - theta * dt * bc; A_U = -theta * dt * c;
            B_L = (1.0 - theta) * dt * a; B_U = (1.0 - theta) * dt * c;
            p = 1.0 / theta; q = (1.0 - theta) / theta;
            double den = A_C, cp = A_U / den;
            inv[0] = 1.0 / den; m[0] = 0.0;
            for (int i = 1; i < ni; ++i) {
                double den_i = A_C - A_L * cp;
                m[i] = A_L * inv[i - 1];
                inv[i] = 1.0 / den_i;
                cp = A_U / den_i;
                cc[i - 1] = A_U * inv[i];
            }
            cc[ni - 1] = 0.0;
        }
        double tau = (step + 1) * dt;
        double Vmin, Vmax;
        if (model == 1) { double disc = exp(-r * tau); if (is_call) { Vmin = 0; Vmax = disc * (s[n] - k); } else { Vmin = disc * k; Vmax = 0; } }
        else if (is_call) { Vmin = 0; Vmax = s[n] * exp((b - r) * tau) - k * exp(-r * tau); }
        else { Vmin = k * exp(-r * tau); Vmax = 0; }
        double bl = B_L * V0 - A_L * Vmin, bu = B_U * VN - A_U * Vmax;
        for (int i = 0; i < ni; ++i) { G[i] = fma(q, V[i], L[i]); V[i] = fma(p, V[i], L[i]); }
        V[0] += bl; V[ni - 1] += bu;
        double e = 0;
        for (int i = 0; i < ni; ++i) { e = fma(-m[i], e, V[i]); V[i] = e; }
        e = 0;
        int lo = -1, hi = -1;
        for (int i = ni - 1; i >= 0; --i) {
            e = fma(-cc[i], e, V[i]);
            double vc = fma(inv[i], e, -G[i]);
            if (american) {
                double d = phi[i + 1] - vc;
                V[i] = phi[i + 1] > vc ? phi[i + 1] : vc;
                L[i] = d > 0 ? d : 0;
                if (d > 0) { if (hi < 0) hi = i; lo = i; }
            } else V[i] = vc;
        }
        V0 = Vmin; VN = Vmax;
        if (ex_lo) { *ex_lo = lo; *ex_hi = hi; }
        if (ko_kind && mi < n_mon && mon[mi] == step + 1) {
            ++mi;
            for (int i = 0; i < ni; ++i)
                if ((ko_kind == 1 && s[i + 1] <= barrier) || (ko_kind == 2 && s[i + 1] >= barrier)) V[i] = rebate;
            if ((ko_kind == 1 && s[0] <= barrier) || (ko_kind == 2 && s[0] >= barrier)) V0 = rebate;
            if ((ko_kind == 1 && s[n] <= barrier) || (ko_kind == 2 && s[n] >= barrier)) VN = rebate;
        }
    }
    V_out[0] = V0; V_out[n] = VN;
    memcpy(V_out + 1, V, ni * sizeof(double));
    free(s); free(phi); free(V); free(L); free(m); free(cc); free(inv); free(G);
    return 0;
}
