// what limits the residency of riccati_kernel<16, true>?  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I mpc-tsid_b200/csrc -o /tmp/occ_probe tools/scratch/occ_probe.cu
#include <cstdio>
#include "mpcqp_riccati.cuh"
using namespace mpcqp;
int main() {
    cudaFuncAttributes a;
    auto k = riccati_kernel<16, true>;
    cudaFuncGetAttributes(&a, k);
    const int smem = 2 * (int)sizeof(RicInst<16>);
    printf("regs %d static smem %zu local %zu maxDyn %d  sizeof(RicInst<16>) %zu -> dyn %d\n", a.numRegs, a.sharedSizeBytes, a.localSizeBytes, a.maxDynamicSharedSizeBytes, sizeof(RicInst<16>), smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int carve : {-1, 100}) {
        if (carve >= 0) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        for (int s : {smem, 22000, 20000, 16000, 8000, 0}) {
            int n = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, 32, s);
            printf("carveout %d dyn smem %d -> %d CTAs/SM\n", carve, s, n);
        }
    }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("regsPerSM %d smemPerSM %zu smemPerBlockOptin %zu reservedPerBlock %zu maxBlocksPerSM %d maxThreadsPerSM %d\n", p.regsPerMultiprocessor, p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin, p.reservedSharedMemPerBlock, p.maxBlocksPerMultiProcessor, p.maxThreadsPerMultiProcessor);
    return 0;
}
