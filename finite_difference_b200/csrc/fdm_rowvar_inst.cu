// fdm_rowvar_inst.cu -- one size class of fdm_rowvar_kernel (compiled with -DFDM_INST_C=<rows per lane>) and its launcher.
#include "fdm_rowvar.cuh"

#ifndef FDM_INST_C
#error "compile with -DFDM_INST_C=<rows per lane>"
#endif
#define FDM_CAT2(a, b) a##b
#define FDM_CAT(a, b) FDM_CAT2(a, b)

int FDM_CAT(fdm_launch_rowvar_, FDM_INST_C)(const FdmRowVarBatch &b, const FdmRowVarOut &o, cudaStream_t st, int n_lo,
                                            int32_t *n_launches)
{
    constexpr int C = FDM_INST_C;
    // two instances per class: marches that are fully implicit throughout (half the shared memory, rowvar::IMPL) and the
    // rest; each warp exits at once when its solve belongs to the other one
    const size_t smem_i = (size_t)rowvar::smem_doubles<C, true>() * sizeof(double);
    const size_t smem_g = (size_t)rowvar::smem_doubles<C, false>() * sizeof(double);
    auto kern_i = rowvar::fdm_rowvar_kernel<C, true>;
    auto kern_g = rowvar::fdm_rowvar_kernel<C, false>;
    if (smem_i > 48 * 1024) FDM_CUDA(cudaFuncSetAttribute(kern_i, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_i));
    if (smem_g > 48 * 1024) FDM_CUDA(cudaFuncSetAttribute(kern_g, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
    const int64_t max_grid = 0x7fffffff;
    int launches = 0;
    for (int64_t first = 0; first < b.n_solves; first += max_grid) {
        const int64_t count = (b.n_solves - first < max_grid) ? (b.n_solves - first) : max_grid;
        kern_i<<<dim3((unsigned)count), dim3(32), smem_i, st>>>(b, o, first, n_lo);
        kern_g<<<dim3((unsigned)count), dim3(32), smem_g, st>>>(b, o, first, n_lo);
        launches += 2;
        FDM_CUDA(cudaGetLastError());
    }
    if (n_launches) *n_launches += launches;
    return FDM_OK;
}
