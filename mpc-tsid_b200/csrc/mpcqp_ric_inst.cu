// mpcqp_ric_inst.cu -- one horizon of the stage-wise solver per translation unit (RIC_N = 16, 32 or 64), so that the three
// instantiations of riccati_kernel compile in parallel.  Defines the two entry points mpcqp_api.cu dispatches to.
#ifndef RIC_N
#error "compile with -DRIC_N=16|32|64"
#endif
#ifndef MPCQP_SINGLE_TU
#include "mpcqp_riccati.cuh"
#endif

#define RIC_CAT2(a, b) a##b
#define RIC_CAT(a, b) RIC_CAT2(a, b)

namespace mpcqp {

// shared memory per CTA and resident CTAs per SM of this horizon's kernels
cudaError_t RIC_CAT(ric_configure_, RIC_N)(int* ctas_per_sm) {
    constexpr int N = RIC_N;
    cudaError_t e;
    const int smem = (int)(RIC_PER_CTA * sizeof(RicInst<N>));
    int a = 1 << 30, b = 0;
    if ((e = cudaFuncSetAttribute(riccati_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    if ((e = cudaFuncSetAttribute(riccati_kernel<N, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, riccati_kernel<N, true>, 32 * RIC_WARPS, smem))) return e;
#if RIC_N != 64
    if ((e = cudaFuncSetAttribute(riccati_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    if ((e = cudaFuncSetAttribute(riccati_kernel<N, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, riccati_kernel<N, false>, 32 * RIC_WARPS, smem))) return e;
#endif
    *ctas_per_sm = a < b ? a : b;
    return cudaSuccess;
}

// N = 64 builds the instance with the stage-wise ADMM stage only (build time); the stage itself is a run-time flag
void RIC_CAT(ric_launch_, RIC_N)(int grid, cudaStream_t s, const DevParams& dp, const DevState& st, const DevScenario& sc, const double* dx,
                                 const double* df, double* ws, int* ctr, int first, int off, int n_inst) {
    constexpr int N = RIC_N;
    const size_t smem = RIC_PER_CTA * sizeof(RicInst<N>);
#if RIC_N != 64
    if (!(dp.mode & 8)) {
        riccati_kernel<N, false><<<grid, 32 * RIC_WARPS, smem, s>>>(dp, st, sc, dx, df, ws, ctr, first, off, n_inst);
        return;
    }
#endif
    riccati_kernel<N, true><<<grid, 32 * RIC_WARPS, smem, s>>>(dp, st, sc, dx, df, ws, ctr, first, off, n_inst);
}

}  // namespace mpcqp
#undef RIC_N
