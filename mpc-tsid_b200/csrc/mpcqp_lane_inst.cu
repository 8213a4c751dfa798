// mpcqp_lane_inst.cu -- the one-robot-per-lane active-set kernel (mpcqp_lane.cuh) and the planner kernel that feeds it in the
// device-resident closed loop, as their own translation unit.  Defines the entry points mpcqp_api.cu dispatches to.
#include "mpcqp_lane.cuh"

namespace mpcqp {

// Device-resident closed loop in front of lane_kernel: the planner (mpcqp_scenario.cuh: gait roll, footsteps, reference trajectory)
// of every robot, half a warp per robot, writing xref / fsteps in the reference's layout to HBM (SC.xref_out / fsteps_out) -- the
// inputs lane_kernel and ipm_kernel then read like a caller's.  Shared memory per robot: ScenarioSmem + 12 (n + 1) + 260 doubles.
__global__ void __launch_bounds__(128)
plan_kernel(DevParams P, DevScenario SC, int inst_offset, int inst_count) {
    extern __shared__ __align__(16) unsigned char plan_raw[];
    const int n = P.N;
    const size_t per = ((sizeof(ScenarioSmem) + 15) / 16 * 16 + (size_t)(12 * (n + 1) + 260) * 8 + 15) / 16 * 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane >> 4;
    unsigned char* mine = plan_raw + (size_t)(warp * 2 + sub) * per;
    ScenarioSmem& sc = *reinterpret_cast<ScenarioSmem*>(mine);
    double* xr = reinterpret_cast<double*>(mine + (sizeof(ScenarioSmem) + 15) / 16 * 16);
    double* fs = xr + 12 * (n + 1);
    const int warps = blockDim.x >> 5;
    for (int w0 = (blockIdx.x * warps + warp) * 2; w0 < inst_count; w0 += gridDim.x * warps * 2) {
        const bool valid = w0 + sub < inst_count;
        const int inst = inst_offset + (valid ? w0 + sub : w0);          // an idle half shadows its neighbour, nothing committed
        scenario_inputs<16>(P, SC, sc, inst, xr, fs, n, valid);
        __syncwarp();
    }
}

size_t plan_smem_bytes(int n) {
    const size_t per = ((sizeof(ScenarioSmem) + 15) / 16 * 16 + (size_t)(12 * (n + 1) + 260) * 8 + 15) / 16 * 16;
    return 8 * per;
}

cudaError_t lane_configure(int* ctas_per_sm) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LANE_SMEM_BYTES))) return e;
    if ((e = cudaFuncSetAttribute(lane_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
    if ((e = cudaFuncSetAttribute(plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan_smem_bytes(64)))) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, lane_kernel, 32, LANE_SMEM_BYTES);
}

// bytes of workspace a launch with `grid` CTAs needs at horizon n
size_t lane_ws_bytes(int grid, int n) {
    return (size_t)grid * 32 * ((size_t)lane_ws_doubles(n) * 8 + (size_t)lane_ws_words(n) * 4);
}

void lane_launch(int grid, cudaStream_t s, const DevParams& dp, const DevState& st, const DevScenario& sc, const double* dx,
                 const double* df, double* ws, int* ctr, int first, int off, int n_inst) {
    lane_kernel<<<grid, 32, LANE_SMEM_BYTES, s>>>(dp, st, sc, dx, df, ws, ctr, first, off, n_inst);
}

void plan_launch(int sms, cudaStream_t s, const DevParams& dp, const DevScenario& sc, int off, int n_inst) {
    int grid = (n_inst + 7) / 8;
    if (grid > 8 * sms) grid = 8 * sms;
    plan_kernel<<<grid, 128, plan_smem_bytes(dp.N), s>>>(dp, sc, off, n_inst);
}

}  // namespace mpcqp
