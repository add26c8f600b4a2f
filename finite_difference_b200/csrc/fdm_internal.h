// fdm_internal.h -- what the translation units of libfdm_b200 share on the HOST side: the error reporter and one launch
// function per compiled kernel instance (fdm_warp_inst.cu / fdm_lean_inst.cu / fdm_cta_inst.cu are compiled once per
// size class, in parallel; fdm_api.cu holds the C ABI and dispatches to them).  Not part of the public ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "fdm_b200.h"

// sets the thread-local message of fdm_last_error() and returns `code`
int fdm_fail(int code, const char *fmt, const char *detail = "");

// fdm_warp_kernel<C>: C slots per lane, up to 32 C interior rows.  only_rest: take only what the lean kernels left.
#define FDM_DECL_WARP(C)                                                                                         \
    int fdm_launch_warp_##C(const FdmBatch &b, const FdmOut &o, cudaStream_t st, bool any_american, bool any_linear, \
                            bool any_div, int only_rest, int32_t *n_launches);
FDM_DECL_WARP(4) FDM_DECL_WARP(8) FDM_DECL_WARP(13) FDM_DECL_WARP(16)

// fdm_lean_kernel<C>: 32 (C-1) <= interior rows <= 32 C
#define FDM_DECL_LEAN(C)                                                                                         \
    int fdm_launch_lean_##C(const FdmBatch &b, const FdmOut &o, cudaStream_t st, bool any_american, bool any_linear, \
                            bool any_div, int32_t *n_launches);
FDM_DECL_LEAN(4) FDM_DECL_LEAN(7) FDM_DECL_LEAN(8) FDM_DECL_LEAN(10) FDM_DECL_LEAN(13) FDM_DECL_LEAN(16)

// fdm_cta_kernel<C, W>: W warps of C slots per thread serve up to 32 W C interior rows; n_lo = largest n_space of the
// class below.  The classes, in ascending capacity (FDM_CTA_CLASSES(X) expands X(W, C) for each):
#define FDM_CTA_CLASSES(X) X(4, 12) X(4, 16) X(5, 14) X(5, 16) X(8, 12) X(8, 16) X(16, 12) X(16, 16)
#define FDM_DECL_CTA(W, C)                                                                                       \
    int fdm_launch_cta_##W##x##C(const FdmBatch &b, const FdmOut &o, cudaStream_t st, bool any_american,        \
                                 bool any_linear, bool any_div, int n_lo, int32_t *n_launches);
FDM_CTA_CLASSES(FDM_DECL_CTA)
// fdm_exact_kernel: the reference's step literally, for solves the CTA kernels flagged FDM_STATUS_STIFF
int fdm_launch_cta_exact(const FdmBatch &b, const FdmOut &o, cudaStream_t st, int32_t *n_launches);

// fdm_rowvar_kernel<C>: C rows per lane, up to 32 C nodes; n_lo = largest node count of the class below
#define FDM_DECL_ROWVAR(C)                                                                                       \
    int fdm_launch_rowvar_##C(const FdmRowVarBatch &b, const FdmRowVarOut &o, cudaStream_t st, int n_lo, int32_t *n_launches);
FDM_DECL_ROWVAR(7) FDM_DECL_ROWVAR(10) FDM_DECL_ROWVAR(13) FDM_DECL_ROWVAR(19) FDM_DECL_ROWVAR(26) FDM_DECL_ROWVAR(32)
