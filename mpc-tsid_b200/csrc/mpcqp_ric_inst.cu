// mpcqp_ric_inst.cu -- one capacity of the stage-wise solver per translation unit (RIC_N = 16, 32 or 64: the largest horizon the
// instantiation can hold), so that the three instantiations compile in parallel.  Defines the entry points mpcqp_api.cu dispatches to.
#ifndef RIC_N
#error "compile with -DRIC_N=16|32|64"
#endif
#ifndef MPCQP_SINGLE_TU
#include "mpcqp_riccati.cuh"
#endif

#define RIC_CAT2(a, b) a##b
#define RIC_CAT(a, b) RIC_CAT2(a, b)

namespace mpcqp {

// shared memory per CTA and resident CTAs per SM of this capacity's kernels
cudaError_t RIC_CAT(ric_configure_, RIC_N)(int* ctas_per_sm, int* ipm_ctas_per_sm) {
    constexpr int N = RIC_N;
    cudaError_t e;
    const int smem = (int)(RIC_PER_CTA * sizeof(RicInst<N>));
    int a = 0, b = 0;
    int a1 = 0;
    if ((e = cudaFuncSetAttribute(riccati_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    if ((e = cudaFuncSetAttribute(riccati_kernel<N, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, riccati_kernel<N, true>, 32 * RIC_WARPS, smem))) return e;
    if ((e = cudaFuncSetAttribute(riccati_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    if ((e = cudaFuncSetAttribute(riccati_kernel<N, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a1, riccati_kernel<N, false>, 32 * RIC_WARPS, smem))) return e;
    a = a1 < a ? a1 : a;
    if ((e = cudaFuncSetAttribute(ipm_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    if ((e = cudaFuncSetAttribute(ipm_kernel<N>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, ipm_kernel<N>, 32 * RIC_WARPS, smem))) return e;
    *ctas_per_sm = a;                      // the two kernels size their persistent grids separately; the workspace holds
    *ipm_ctas_per_sm = b;                  // one slot per resident half-warp of the larger one
    return cudaSuccess;
}

void RIC_CAT(ric_launch_, RIC_N)(int grid, cudaStream_t s, const DevParams& dp, const DevState& st, const DevScenario& sc, const double* dx,
                                 const double* df, double* ws, int* ctr, int first, int off, int n_inst) {
    constexpr int N = RIC_N;
    const size_t smem = RIC_PER_CTA * sizeof(RicInst<N>);
    if (dp.N == N) riccati_kernel<N, true><<<grid, 32 * RIC_WARPS, smem, s>>>(dp, st, sc, dx, df, ws, ctr, first, off, n_inst);
    else riccati_kernel<N, false><<<grid, 32 * RIC_WARPS, smem, s>>>(dp, st, sc, dx, df, ws, ctr, first, off, n_inst);
}

// fallback stage.  pdl: programmatic dependent launch behind the active-set kernel of the same stream (the kernel blocks in
// griddepcontrol.wait before it reads the queue); behind anything else the attribute is an ordinary launch.  Without it the CTAs
// are not scheduled before the active-set kernel has drained (overlapped index ranges: a blocked resident CTA would hold a slot
// another range's robots could use).
void RIC_CAT(ipm_launch_, RIC_N)(int grid, cudaStream_t s, const DevParams& dp, const DevState& st, const DevScenario& sc, const double* dx,
                                 const double* df, double* ws, int first, int pdl) {
    constexpr int N = RIC_N;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32 * RIC_WARPS); cfg.dynamicSmemBytes = RIC_PER_CTA * sizeof(RicInst<N>); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, ipm_kernel<N>, dp, st, sc, dx, df, ws, first);
}

}  // namespace mpcqp
#undef RIC_N
