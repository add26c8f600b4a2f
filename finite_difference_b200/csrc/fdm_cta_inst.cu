// fdm_cta_inst.cu -- one size class of fdm_cta_kernel (compiled with -DFDM_INST_W=<warps per CTA> and
// -DFDM_INST_C=<slots per thread>, default 16) and its launcher fdm_launch_cta_<W>x<C>.
#include "fdm_cta.cuh"

#ifndef FDM_INST_W
#error "compile with -DFDM_INST_W=<warps per CTA>"
#endif
#ifndef FDM_INST_C
#define FDM_INST_C 16
#endif
#define FDM_HOSTS_EXACT (FDM_INST_W == 16 && FDM_INST_C == 16)
#if FDM_HOSTS_EXACT
#include "fdm_exact.cuh"
#endif
#define FDM_CAT2(a, b) a##b
#define FDM_CAT(a, b) FDM_CAT2(a, b)

int FDM_CAT(FDM_CAT(FDM_CAT(fdm_launch_cta_, FDM_INST_W), x), FDM_INST_C)(
    const FdmBatch &b, const FdmOut &o, cudaStream_t st, bool any_american, bool any_linear, bool any_div, int n_lo,
    int32_t *n_launches)
{
    constexpr int C = FDM_INST_C, WARPS = FDM_INST_W;
    (void)any_div;
    const bool trace = o.ex_trace != nullptr;
    const size_t smem = (size_t)cta::smem_doubles<C, WARPS>(true) * sizeof(double);
    const size_t smem_lin = (size_t)cta::smem_doubles<C, WARPS>(false) * sizeof(double);
    auto k_am = cta::fdm_cta_kernel<C, WARPS, true, false>;
    auto k_amt = cta::fdm_cta_kernel<C, WARPS, true, true>;
    auto k_lin = cta::fdm_cta_kernel<C, WARPS, false, false>;
    FDM_CUDA(cudaFuncSetAttribute(k_am, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FDM_CUDA(cudaFuncSetAttribute(k_amt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FDM_CUDA(cudaFuncSetAttribute(k_lin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lin));
    const int64_t max_grid = 0x7fffffff;
    int launches = 0;
    for (int64_t first = 0; first < b.n_solves; first += max_grid) {
        const int64_t count = (b.n_solves - first < max_grid) ? (b.n_solves - first) : max_grid;
        const dim3 grid((unsigned)count), block(32 * WARPS);
        if (any_american) {
            if (trace) k_amt<<<grid, block, smem, st>>>(b, o, first, 1, n_lo);
            else       k_am<<<grid, block, smem, st>>>(b, o, first, 1, n_lo);
            ++launches;
        }
        if (any_linear) {
            k_lin<<<grid, block, smem_lin, st>>>(b, o, first, 0, n_lo);
            ++launches;
        }
        FDM_CUDA(cudaGetLastError());
    }
    if (n_launches) *n_launches += launches;
    return FDM_OK;
}

#if FDM_HOSTS_EXACT
// fdm_exact_kernel over the solves the CTA kernels flagged FDM_STATUS_STIFF (fdm_exact.cuh); needs out.status
int fdm_launch_cta_exact(const FdmBatch &b, const FdmOut &o, cudaStream_t st, int32_t *n_launches)
{
    if (!o.status) return FDM_OK;
    const size_t smem = cta::exact_smem_bytes(b.n_space_max);
    FDM_CUDA(cudaFuncSetAttribute(cta::fdm_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t per = cta::EXACT_THREADS, max_grid = 0x7fffffff;
    int launches = 0;
    for (int64_t first = 0; first < b.n_solves; first += max_grid * per) {
        const int64_t left = b.n_solves - first;
        const int64_t count = (left < max_grid * per) ? (left + per - 1) / per : max_grid;
        cta::fdm_exact_kernel<<<dim3((unsigned)count), dim3(cta::EXACT_THREADS), smem, st>>>(b, o, first);
        ++launches;
        FDM_CUDA(cudaGetLastError());
    }
    if (n_launches) *n_launches += launches;
    return FDM_OK;
}
#endif
