// fdm_warp_inst.cu -- one instance of fdm_warp_kernel (compiled with -DFDM_INST_C=4|8|13|16) and its launcher.
#include "fdm_warp.cuh"

#ifndef FDM_INST_C
#error "compile with -DFDM_INST_C=<slots per lane>"
#endif
#define FDM_CAT2(a, b) a##b
#define FDM_CAT(a, b) FDM_CAT2(a, b)

int FDM_CAT(fdm_launch_warp_, FDM_INST_C)(const FdmBatch &b, const FdmOut &o, cudaStream_t st, bool any_american,
                                          bool any_linear, bool any_div, int only_rest, int32_t *n_launches)
{
    constexpr int C = FDM_INST_C;
    const bool trace = o.ex_trace != nullptr;
    const int base = warp_smem_doubles<C>();
    const int div_doubles = any_div ? 4 * (32 * C + 2) : 0;
    const size_t smem = (size_t)(base + div_doubles) * sizeof(double);
    const int64_t max_grid = 0x7fffffff;
    int launches = 0;
    const int per_cta = only_rest ? 32 : 1;                  // only_rest: one warp screens 32 solves (fdm_warp.cuh)
    for (int64_t first = 0; first < b.n_solves; first += max_grid * per_cta) {
        const int64_t left = b.n_solves - first;
        const int64_t count = (left < max_grid * per_cta) ? (left + per_cta - 1) / per_cta : max_grid;
        const dim3 grid((unsigned)count), block(32);
        if (any_american) {
            if (trace) fdm_warp_kernel<C, true, true><<<grid, block, smem, st>>>(b, o, first, 1, any_div, only_rest);
            else       fdm_warp_kernel<C, true, false><<<grid, block, smem, st>>>(b, o, first, 1, any_div, only_rest);
            ++launches;
        }
        if (any_linear) {
            fdm_warp_kernel<C, false, false><<<grid, block, smem, st>>>(b, o, first, 0, any_div, only_rest);
            ++launches;
        }
        FDM_CUDA(cudaGetLastError());
    }
    if (n_launches) *n_launches += launches;
    return FDM_OK;
}
