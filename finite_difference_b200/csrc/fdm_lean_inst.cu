// fdm_lean_inst.cu -- one size class of fdm_lean_kernel (compiled with -DFDM_INST_C=<slots per lane>) and its launcher.
#include "fdm_lean.cuh"

#ifndef FDM_INST_C
#error "compile with -DFDM_INST_C=<slots per lane>"
#endif
#define FDM_CAT2(a, b) a##b
#define FDM_CAT(a, b) FDM_CAT2(a, b)

int FDM_CAT(fdm_launch_lean_, FDM_INST_C)(const FdmBatch &b, const FdmOut &o, cudaStream_t st, bool any_american,
                                          bool any_linear, bool any_div, int32_t *n_launches)
{
    constexpr int C = FDM_INST_C;
    const bool trace = o.ex_trace != nullptr;
    (void)any_div;
    const size_t smem = (size_t)lean::smem_doubles<C>() * sizeof(double);
    auto k_am = lean::fdm_lean_kernel<C, true, false>;
    auto k_amt = lean::fdm_lean_kernel<C, true, true>;
    auto k_lin = lean::fdm_lean_kernel<C, false, false>;
    if (smem > 48 * 1024) {
        FDM_CUDA(cudaFuncSetAttribute(k_am, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FDM_CUDA(cudaFuncSetAttribute(k_amt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FDM_CUDA(cudaFuncSetAttribute(k_lin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int64_t max_grid = 0x7fffffff;
    int launches = 0;
    for (int64_t first = 0; first < b.n_solves; first += max_grid) {
        const int64_t count = (b.n_solves - first < max_grid) ? (b.n_solves - first) : max_grid;
        const dim3 grid((unsigned)count), block(32);
        if (any_american) {
            if (trace) k_amt<<<grid, block, smem, st>>>(b, o, first, 1);
            else       k_am<<<grid, block, smem, st>>>(b, o, first, 1);
            ++launches;
        }
        if (any_linear) {
            k_lin<<<grid, block, smem, st>>>(b, o, first, 0);
            ++launches;
        }
        FDM_CUDA(cudaGetLastError());
    }
    if (n_launches) *n_launches += launches;
    return FDM_OK;
}
