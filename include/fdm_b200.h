/*
 * fdm_b200.h -- C ABI of the B200 batched finite-difference PDE engine (libfdm_b200.so).
 *
 * The reference (rwx-gigaba-sonwabo/Finite_Difference) is pure Python and has no FFI.  The
 * natural cut is its `_solve_grid` method: one call = one 1-D log-space theta/Crank-Nicolson
 * PDE solve returning the value vector at valuation, from which the pricer classes read a
 * price and a small stencil:
 *
 *   fd_american_option_pricer.py:554-602   AmericanFDMPricer._solve_grid(N_time)
 *   fd_american_equity.py:778-843          AmericanFDMPricer._solve_grid(n_time)
 *   fd_american_black76.py:450-472         AmericanFwdFDMPricer._solve_grid(N_time)
 *   discrete_barrier_fdm_pricer_cn.py:219-302, fd_discrete_barrier_option.py:182-262
 *                                          DiscreteBarrierCrankNicolsonLog._solve_grid(apply_KO)
 *
 * `fdm_solve_batch*` replaces a batch of those calls plus the read-out that follows them
 * (`_interp_price` / `_local_cubic_delta_gamma`, fd_american_option_pricer.py:613-653;
 * `_interp_price_from_grid` / `_delta_gamma_from_grid`,
 * discrete_barrier_fdm_pricer_cn.py:308-353).  Everything above `_solve_grid` (dates, curves,
 * Richardson and bump fan-out, knock-in parity) stays in the host language and only builds
 * the descriptors below.
 *
 * Plain C, plain pointers and sizes: bindable from ctypes / cffi / cgo / JNI alike.
 * All functions return FDM_OK (0) or a negative error code, never throw, and set a
 * thread-local message readable through fdm_last_error().
 */
#ifndef FDM_B200_H
#define FDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDM_ABI_VERSION 3

/* error codes */
#define FDM_OK                0
#define FDM_ERR_INVALID      -1   /* bad argument (NULL pointer, n_space < 3, ...)          */
#define FDM_ERR_CUDA         -2   /* a CUDA runtime call failed; see fdm_last_error()       */
#define FDM_ERR_UNSUPPORTED  -3   /* grid larger than any compiled kernel variant           */
#define FDM_ERR_NO_DEVICE    -4   /* no CUDA device: there is NO CPU fallback               */

/* per-solve flags */
#define FDM_FLAG_CALL      0x01   /* payoff max(S-K,0); else put                            */
#define FDM_FLAG_AMERICAN  0x02   /* Ikonen-Toivanen projection, fd_american_option_pricer.py:490-503 */
#define FDM_FLAG_BLACK76   0x04   /* drift -sigma^2/2 and Black-76 boundaries, fd_american_black76.py:291-314,:351 */
#define FDM_FLAG_SNAP      0x08   /* snap spot and strike to the nearest node, fd_american_option_pricer.py:246-266;
                                     price = V[spot node], 4-point stencil.  Without it (barrier dataclass):
                                     raw strike, linear interpolation at spot, 3-point stencil              */
#define FDM_FLAG_KO_DOWN   0x10   /* down-and-out projection at monitor steps, discrete_barrier_fdm_pricer_cn.py:203-207 */
#define FDM_FLAG_KO_UP     0x20   /* up-and-out projection, discrete_barrier_fdm_pricer_cn.py:209-213.
                                     KO_DOWN | KO_UP together: double-out (discrete_barrier_fdm_pricer.py:433-435),
                                     lower level in barrier[], upper level in barrier_up[] */
/* The date/curve-aware barrier pricer DiscreteBarrierFDMPricer (discrete_barrier_fdm_pricer.py:442-547) marches a
   vector that is one node shorter than its s_nodes (V[1:-1] = V_int with N = num_space_nodes - 1, :448, :540):
   the kernel grid is then n_space = num_space_nodes - 1, but the Dirichlet value at the top is still evaluated at
   s_nodes[-1], one node ABOVE the kernel grid. */
#define FDM_FLAG_BDRY_TOP_NEXT 0x40  /* top Dirichlet value uses S = exp(x_min + (n_space+1) dx), :378 with :540 */
#define FDM_FLAG_PUT_BDRY_SMIN 0x80  /* put: V_min = K e^{-r tau} * S_min * e^{(b-r) tau}, discrete_barrier_fdm_pricer.py:391 */
#define FDM_FLAG_INVALID 0x100       /* set by the book fan-out on a solve whose inputs the reference rejects (a zero-length
                                        dividend segment divides by zero at fd_american_option_pricer.py:498; non-positive
                                        spot / strike / sigma / T raise at :68-69, :123-124): no kernel touches it, its status
                                        stays FDM_STATUS_NOT_SOLVED */

/* per-solve status (FdmOut.status, FdmBookOut.status) */
#define FDM_STATUS_OK           0
#define FDM_STATUS_NONFINITE    1   /* price is not finite */
#define FDM_STATUS_STIFF        2   /* CTA kernel: pivots not converged inside its table.  With FdmOut.status given, such solves
                                       are solved again by the exact (serial Thomas) kernel and end with status 0; without a
                                       status array the best-effort result stands */
#define FDM_STATUS_SINGULAR     3   /* row-varying solves: a pivot of the Thomas factorisation vanished (the reference raises
                                       ZeroDivisionError, discrete_barrier_fdm_pricer_2.py:271-279); price is NaN */
#define FDM_STATUS_NOT_SOLVED  -1   /* no kernel took the solve: n_space outside [3, n_space_max], launch hints that miss
                                       its size class, or FDM_FLAG_INVALID.  status[] is filled with this value before
                                       the launches, so a skipped solve is always visible */

/* largest num_space_nodes any compiled kernel variant handles (one CTA per solve above
   FDM_WARP_MAX_N_SPACE, one warp per solve up to it) */
#define FDM_MAX_N_SPACE 8193
#define FDM_WARP_MAX_N_SPACE 513

/*
 * A batch of PDE solves, structure-of-arrays.  Every pointer addresses `n_solves` elements
 * unless noted.  For fdm_solve_batch() all pointers are DEVICE pointers owned by the caller;
 * for fdm_solve_batch_host() they are HOST pointers.
 */
typedef struct FdmBatch {
    int64_t n_solves;
    int32_t n_space_max;   /* upper bound of n_space[] in this call: selects the kernel variant */
    int32_t n_seg_total;   /* length of the seg_* tables  */
    int32_t n_mon_total;   /* length of mon_step (0 if none) */
    int32_t reserved;      /* optional launch hints from fdm_launch_hints(); 0 = unknown mix (every kernel family and
                              size class is launched; solves that are not theirs exit at once) */

    const double *x_min;   /* log(S_min) of the grid, fd_american_option_pricer.py:233 */
    const double *dx;      /* (log S_max - log S_min)/n_space, :237 */
    const double *sigma;   /* :413 */
    const double *r;       /* discount rate (NACC), :415 */
    const double *b;       /* carry rate (NACC), :416; ignored with FDM_FLAG_BLACK76 */
    const double *spot;    /* :73 / S0 */
    const double *strike;  /* :74 / K  */
    const double *barrier; /* KO level; may be NULL if no solve has a KO flag */
    const double *rebate;  /* undiscounted rebate, discrete_barrier_fdm_pricer_cn.py:207; may be NULL */
    const int32_t *n_space;    /* num_space_nodes >= 3 */
    const int32_t *flags;      /* FDM_FLAG_* */
    const int32_t *rannacher;  /* rannacher_steps >= 0, :464 */
    const int32_t *seg_begin;  /* CSR, n_solves+1 entries into seg_* */
    const int32_t *mon_begin;  /* CSR, n_solves+1 entries into mon_step; NULL if n_mon_total == 0 */

    /* time segments, fd_american_option_pricer.py:565-600 (one per solve when there are no dividends) */
    const int32_t *seg_steps;    /* n_steps >= 1 */
    const double *seg_dt;        /* (tau_end - tau_start)/n_steps, :412 */
    const double *seg_tau0;      /* tau_start */
    const int32_t *seg_restart;  /* restart_rannacher, :583-586 */
    const double *seg_div;       /* cash dividend applied AFTER the segment (spline jump, :517-548); 0 = none */

    /* KO monitor steps, discrete_barrier_fdm_pricer_cn.py:177-193: 1-based global step index k,
       strictly ascending per solve; projection runs after step k */
    const int32_t *mon_step;

    /* ABI 2 (all optional, NULL = not used) */
    const double *b_bdry;      /* carry rate inside the Dirichlet values when it differs from the drift's b
                                  (discrete_barrier_fdm_pricer.py:382 uses carry_rate_nacc, the drift :461 b - q) */
    const double *barrier_up;  /* upper KO level of a double-out solve (KO_DOWN | KO_UP) */
    const double *mon_rebate;  /* [n_mon_total] value written at knocked-out nodes at each monitor step, e.g.
                                  rebate * exp(-b tau) (discrete_barrier_fdm_pricer.py:421-424); NULL: rebate[] */
    /* ABI 3 (optional) */
    const int32_t *mon_end;    /* [n_solves] end of each solve's monitor range; NULL: mon_begin[i+1].  Lets several solves
                                  (the sigma bumps of one option) share one monitor table */
} FdmBatch;

/*
 * Results, structure-of-arrays, `n_solves` elements each (stencils: 4 per solve).
 * Optional pointers may be NULL.
 */
typedef struct FdmOut {
    double *price;      /* SNAP: V[spot node]; else linear interpolation at spot */
    double *delta;      /* SNAP: cubic through the 4-point stencil at the snapped spot; else 3-point formula */
    double *gamma;
    double *stencil_v;  /* [4*n_solves] V at the stencil nodes (3 used without SNAP) */
    double *stencil_s;  /* [4*n_solves] S at the stencil nodes */
    int32_t *idx;       /* [4*n_solves]: {spot node (SNAP) or lower bracket node, strike node (SNAP) or
                           stencil centre node, ex_lo, ex_hi}.  ex_lo/ex_hi: first/last interior index k
                           (node k+1) in the exercise set after the last step, -1 if empty / not American */
    int32_t *status;    /* FDM_STATUS_* */
    double *v_dump;     /* optional: full value vector, v_dump_stride doubles per solve */
    int64_t v_dump_stride;
    int32_t *ex_trace;  /* optional: (lo,hi) per marched step, ex_trace_stride int32 per solve */
    int64_t ex_trace_stride;
} FdmOut;

/* ABI version of the loaded library (== FDM_ABI_VERSION of the header it was built from). */
int fdm_version(void);

/* sizeof() of the ABI structs as this library was compiled: 0 FdmBatch, 1 FdmOut, 2 FdmAmericanBook, 3 FdmBookOut,
   4 FdmBarrierBook, 5 FdmRowVarBatch, 6 FdmRowVarOut (-1: unknown).  Lets a binding verify its own struct declarations at load time. */
int fdm_sizeof(int which);

/* Message of the last failure on the calling thread ("" if none). */
const char *fdm_last_error(void);

/* Number of visible CUDA devices, or a negative error code. */
int fdm_device_count(void);

/*
 * Launch all solves of `batch` on the current CUDA device, asynchronously on `stream`
 * (a cudaStream_t passed as void*; NULL = default stream).  Device pointers in, device
 * pointers out; no allocation, no synchronisation.  `n_launches`, if not NULL, receives the
 * number of kernels launched.  `status[]` (if given) is set to FDM_STATUS_NOT_SOLVED on the stream
 * first: the arrays cannot be validated on the host here, and every kernel skips a solve outside
 * its size class.
 * Threading: re-entrant; any number of host threads and devices (each call runs on the caller's
 * current device and stream and owns no state).
 */
int fdm_solve_batch(const FdmBatch *batch, const FdmOut *out, void *stream, int32_t *n_launches);

/*
 * Launch hints for FdmBatch.reserved, from HOST copies of n_space[], flags[] (n_solves entries) and seg_div[]
 * (n_seg entries, may be NULL): which kernel families (American / linear, dividend scratch) and grid-size classes
 * occur, so that fdm_solve_batch() launches only those.  Pure host function, no device needed.
 */
int fdm_launch_hints(const int32_t *n_space, const int32_t *flags, int64_t n_solves, const double *seg_div,
                     int64_t n_seg);

/*
 * Host-buffer entry point: copies the descriptors to the device, solves, copies the results
 * back and synchronises.  `batch` and `out` hold HOST pointers.  `kernel_ms`, if not NULL,
 * receives the device time of the solve kernels alone (CUDA events).
 * n_space[i] outside [3, n_space_max], seg_steps < 1 or seg_dt <= 0 are rejected with FDM_ERR_INVALID.
 * Threading: one grow-only workspace and one lock PER DEVICE; host threads driving different devices
 * run concurrently, threads sharing a device serialise (the call is synchronous anyway).
 */
int fdm_solve_batch_host(const FdmBatch *batch, const FdmOut *out, int32_t *n_launches, float *kernel_ms);

/* ------------------------------------------------------------------------------------------------
 * Books: the fan-out behind the pricer classes' public methods, expanded and combined on the device.
 * Replaces N x { price_log2(); greeks_log2() } (run_american_scenarios.py:144-172) -- or any one of
 * price_log / price_log2 / greeks_log2 (fd_american_option_pricer.py:659-751; Black-76:
 * fd_american_black76.py:533-625) -- for N options, with or without discrete cash dividends.  Per-option inputs are the scalars
 * the reference constructor derives (:117-149): T = time_to_expiry, r = discount_rate_nacc,
 * b = carry_rate_nacc.
 */
#define FDM_BOOK_PRICE_LOG   0   /* 1 solve  */
#define FDM_BOOK_PRICE_LOG2  1   /* 2 solves (1 without Richardson) */
#define FDM_BOOK_GREEKS_LOG2 2   /* 6 solves (3 without Richardson) */
#define FDM_BOOK_SCENARIO    3   /* price_log2 then greeks_log2: 8 solves (4 without Richardson) */

typedef struct FdmAmericanBook {
    int64_t n_options;
    const double *spot, *strike, *T, *sigma, *r, *b;
    const int32_t *flags;        /* FDM_FLAG_CALL | FDM_FLAG_BLACK76 per option */
    const int32_t *n_space, *n_time, *rannacher;
    double s_max_mult;           /* 4.5 in the reference */
    double dv_sigma;             /* volatility bump h of greeks_log2 (0.01) */
    int32_t what;                /* FDM_BOOK_* */
    int32_t use_richardson;
    /* ABI 3: discrete cash dividends (fd_american_option_pricer.py:304-315, :554-602).  CSR per option, entries
       ascending in tau = time to maturity at the ex-date, as _div_times_tau returns them; ignored for
       FDM_FLAG_BLACK76 options (fd_american_black76.py:89).  NULL div_begin = no dividends. */
    const int32_t *div_begin;    /* [n_options + 1] */
    const double *div_tau, *div_cash;
    int64_t n_div_total;         /* div_begin[n_options] (the device entry cannot read it) */
    /* device entry only (the host entry derives them) */
    int32_t n_space_max;         /* upper bound of max(n_space[i], 3) */
    int32_t hints;               /* fdm_launch_hints() of (max(n_space,3), FDM_FLAG_AMERICAN); 0 = unknown */
} FdmAmericanBook;

typedef struct FdmBookOut {      /* n_options entries each; any pointer may be NULL */
    double *price_log2;          /* price_log() / price_log2() (what = 0, 1, 3) */
    double *price, *delta, *gamma, *vega, *theta;   /* greeks_log2() (what = 2, 3; what = 0: price/delta/gamma) */
    int32_t *ex_lo, *ex_hi;      /* exercise interval of the base solve after its last step */
    int32_t *status;             /* max over the option's solves */
    /* ABI 3 (optional): option i of the book writes its results to entry out_index[i] of every array above instead
       of entry i.  A caller that lays a book out in a launch order of its own choosing (e.g. ascending stiffness, so
       that neighbouring CTAs run the same specialised time loop) gets the results back in its original order with
       no pass of its own.  NULL = identity.  Host pointer for the *_host entries, device pointer otherwise. */
    const int32_t *out_index;
} FdmBookOut;

/* solves per option for (what, use_richardson) */
int fdm_book_solves_per_option(int32_t what, int32_t use_richardson);

/*
 * HOST pointers in `book` and `out`.  Copies ~56 B per option (+ 16 B per dividend) to the device, expands, solves,
 * combines, copies the results back and synchronises.  Raises FDM_ERR_INVALID where the reference raises
 * (non-positive spot / strike / T / bumped sigma; a zero-length dividend segment).  `node_steps` (optional) receives the metric unit
 * sum(num_space_nodes x time steps) over all executed solves.
 */
int fdm_american_book_host(const FdmAmericanBook *book, const FdmBookOut *out, int64_t *node_steps,
                           int32_t *n_launches, float *kernel_ms);

/*
 * Device-pointer book entry: `book` and `out` hold DEVICE pointers (per-option arrays already in HBM), `workspace`
 * is a device buffer of at least fdm_american_book_workspace() bytes, 256-byte aligned.  Fan-out, solves and
 * epilogue are queued on `stream`; no allocation, no synchronisation, no host copies.  book->n_space_max must be
 * set; book->hints may be 0.  Inputs the reference would reject cannot raise here: the option's solves are flagged
 * FDM_FLAG_INVALID and its status is FDM_STATUS_NOT_SOLVED, its outputs NaN.  Re-entrant.
 */
size_t fdm_american_book_workspace(int64_t n_options, int64_t n_div_total, int32_t what, int32_t use_richardson);
int fdm_american_book(const FdmAmericanBook *book, const FdmBookOut *out, void *workspace, size_t workspace_bytes,
                      void *stream, int32_t *n_launches);

/* ------------------------------------------------------------------------------------------------
 * Barrier books: the fan-out behind DiscreteBarrierCrankNicolsonLog.price() / greeks()
 * (discrete_barrier_fdm_pricer_cn.py:429-466: base solve, sigma + dv, sigma - dv on the grid frozen at the base
 * volatility, :59-118) for N options, expanded and combined on the device.  Knock-in parity and the closed-form
 * `none` leg (:489-500) stay with the caller.
 */
#define FDM_BARRIER_BOOK_SOLVE  0   /* 1 knock-out solve: price, delta, gamma */
#define FDM_BARRIER_BOOK_GREEKS 1   /* 3 solves: + vega = (p(sigma+dv) - p(sigma-dv)) / (2 dv), theta (:457-464) */

typedef struct FdmBarrierBook {
    int64_t n_options;
    const double *S0, *K, *T, *sigma, *r_disc, *b_carry;
    const double *x_min, *dx;        /* log grid of configure_grid() (:59-131) at the base volatility */
    const double *barrier;           /* knock-out level (lower level of a double-out) */
    const double *barrier_up;        /* upper level of a double-out; may be NULL */
    const double *rebate;
    const int32_t *flags;            /* FDM_FLAG_CALL | FDM_FLAG_KO_DOWN | FDM_FLAG_KO_UP */
    const int32_t *n_space, *n_time; /* N_space, N_time (resolved) */
    const int32_t *mon_begin;        /* CSR [n_options + 1] into mon_step; shared by the option's solves */
    const int32_t *mon_step;         /* 1-based monitor steps, ascending (:177-193) */
    int64_t n_mon_total;
    double dv_sigma;                 /* 1e-3 in the reference */
    int32_t what;                    /* FDM_BARRIER_BOOK_* */
    int32_t n_space_max;             /* device entry: upper bound of n_space[] */
    int32_t hints;                   /* device entry: fdm_launch_hints() bits or 0 */
    int32_t reserved;
} FdmBarrierBook;

int fdm_barrier_book_solves_per_option(int32_t what);
size_t fdm_barrier_book_workspace(int64_t n_options, int32_t what);
/* DEVICE pointers; FdmBookOut: price, delta, gamma, vega, theta, status are written (others ignored). */
int fdm_barrier_book(const FdmBarrierBook *book, const FdmBookOut *out, void *workspace, size_t workspace_bytes,
                     void *stream, int32_t *n_launches);
/* HOST pointers; copies in, solves, copies out, synchronises. */
int fdm_barrier_book_host(const FdmBarrierBook *book, const FdmBookOut *out, int64_t *node_steps,
                          int32_t *n_launches, float *kernel_ms);

/* ------------------------------------------------------------------------------------------------
 * Row-varying solves: DiscreteBarrierFDMPricer2._solve_pde_backward (discrete_barrier_fdm_pricer_2.py:336-428) --
 * Black-Scholes in S on a grid with snapped nodes, so every row of the tridiagonal system has its own coefficients;
 * non-symmetric stencil on the two rows next to the barrier (:383-409); Dirichlet rows 0 and N inside the system;
 * knock-out projection after the steps of the monitoring map (:423-425).  One call = a batch of such marches plus
 * the linear read-out at the effective spot (:433-449).  The caller supplies what the class derives outside the
 * march: the (escrow-shifted) grid, the terminal payoff, the effective barrier levels and the barrier interval.
 */
#define FDM_ROWVAR_MAX_NODES 1024

typedef struct FdmRowVarBatch {
    int64_t n_solves;
    int32_t n_nodes_max;          /* upper bound of the per-solve node counts N + 1 (<= FDM_ROWVAR_MAX_NODES) */
    int32_t n_mon_total;
    const int64_t *node_begin;    /* CSR [n_solves + 1] into s_nodes / v_terminal */
    const double *s_nodes;        /* grid S_0 .. S_N per solve (:468), ascending */
    const double *v_terminal;     /* terminal payoff per node (:231-249) */
    const double *sigma, *r, *strike, *tenor, *dt;
    const double *lo_bar, *up_bar;   /* effective knock-out levels (:291-294); read only where the flag is set */
    const double *h_minus, *h_plus;  /* non-symmetric stencil of rows bar_j and bar_j + 1 (:297-321) */
    const double *s_eff;          /* read-out point: spot minus escrowed dividends (:470) */
    const int32_t *n_steps;       /* num_time_steps M */
    const int32_t *rannacher;     /* fully implicit steps at the start of the march (:351) */
    const int32_t *flags;         /* FDM_FLAG_CALL | FDM_FLAG_KO_DOWN | FDM_FLAG_KO_UP */
    const int32_t *bar_j;         /* left index of the interval holding the active barrier, -1 = none */
    const int32_t *mon_begin;     /* CSR [n_solves + 1] into mon_step; NULL if n_mon_total == 0 */
    const int32_t *mon_step;      /* 1-based count of completed steps after which the projection runs, ascending:
                                     M - k for every key k < M of the reference's monitoring map (:455-466) */
} FdmRowVarBatch;

typedef struct FdmRowVarOut {
    double *price;                /* linear interpolation at s_eff */
    int32_t *status;              /* FDM_STATUS_*; may be NULL */
    double *v_dump;               /* optional: value vector, v_dump_stride doubles per solve */
    int64_t v_dump_stride;
} FdmRowVarOut;

/* DEVICE pointers, asynchronous on `stream`, no allocation. */
int fdm_rowvar_solve(const FdmRowVarBatch *batch, const FdmRowVarOut *out, void *stream, int32_t *n_launches);
/* HOST pointers: H2D, kernels, D2H, synchronises. */
int fdm_rowvar_solve_host(const FdmRowVarBatch *batch, const FdmRowVarOut *out, int32_t *n_launches, float *kernel_ms);

/*
 * The host-buffer entry points keep one grow-only device workspace per device (no cudaMalloc per call).
 * This releases the workspaces of every device.
 */
int fdm_release_workspace(void);

/*
 * Register-resident dependent/independent DFMA micro-benchmark: the FP64 roofline
 * denominator MEASURED_PEAKS.json lacks.  Writes achieved TFLOP/s (FMA = 2 flops).
 */
int fdm_fp64_peak(double *tflops_out, float *ms_out);

#ifdef __cplusplus
}
#endif

#endif /* FDM_B200_H */
