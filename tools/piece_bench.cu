// Cycle counts of the solve kernel's building blocks in isolation (one CTA, or several per SM).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../mpc-tsid_b200/csrc -o piece_bench piece_bench.cu
#include <cstdio>
#include <vector>
#include <cmath>
#define MPCQP_TILE_PROF
#include "mpcqp_device.cuh"
using namespace mpcqp;
constexpr int NT = 12, NTILES = 78;

struct Sm { double W[NTILES * 64]; double u[96]; double tmp[96]; unsigned long long mbar; int flag; };

__global__ void __launch_bounds__(128, 4) bench(const double* Wg, long long* out, int reps) {
    extern __shared__ __align__(16) unsigned char raw[];
    Sm& sm = *reinterpret_cast<Sm*>(raw);
    unsigned int phase = 0;
    if (threadIdx.x == 0) mbar_init(&sm.mbar, 1);
    __syncthreads();
    long long t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < reps; ++r) {
        __syncthreads();
        long long a = clock64();
        if (threadIdx.x == 0) { fence_async_smem(); mbar_expect_tx(&sm.mbar, NTILES * 512); bulk_g2s(sm.W, Wg, NTILES * 512, &sm.mbar); }
        mbar_wait(&sm.mbar, phase); phase ^= 1;
        __syncthreads();
        long long b = clock64(); t[0] += b - a;
        bool ok = factor_invert_tiles<NT, 4>(sm.W, &sm.flag);
        long long c = clock64(); t[1] += c - b;
        long long d = clock64(); t[2] += d - c;
        for (int i = threadIdx.x; i < 96; i += blockDim.x) sm.u[i] = 1.0 + i;
        __syncthreads();
        long long e = clock64();
        tri_solve<NT, 4>(sm.W, sm.u, sm.tmp);
        long long f = clock64(); t[3] += f - e;
        // a bare dependent chain for reference: 64 dependent DFMA, 16 rsqrt
        double x = sm.u[threadIdx.x & 63];
#pragma unroll 1
        for (int i = 0; i < 64; ++i) x = fma(x, 1.0000001, 1e-9);
        long long g2 = clock64(); t[4] += g2 - f;
#pragma unroll 1
        for (int i = 0; i < 16; ++i) x = rsqrt(x + 1.5);
        long long h = clock64(); t[5] += h - g2;
        sm.tmp[threadIdx.x & 63] = x + (ok ? 0 : 1);
        __syncthreads();
        long long k2 = clock64();
        for (int i = 0; i < 16; ++i) __syncthreads();
        t[6] += clock64() - k2;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) for (int i = 0; i < 8; ++i) out[i] = t[i] / reps;
    if (threadIdx.x == 0) out[8 + blockIdx.x % 8] = (long long)sm.u[5];
}

int main() {
    // SPD test matrix in tile layout: diagonally dominant
    std::vector<double> W(NTILES * 64, 0.0);
    for (int i = 0; i < 96; ++i) for (int j = 0; j < 96; ++j) {
        if ((i >> 3) < (j >> 3)) continue;
        double v = (i == j) ? 100.0 + i : 1.0 / (1.0 + std::abs(i - j));
        W[tile_index(i >> 3, j >> 3) * 64 + elem_off(i & 7, j & 7)] = v;
    }
    double* dW; long long* dout;
    cudaMalloc(&dW, W.size() * 8); cudaMemcpy(dW, W.data(), W.size() * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 16 * 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm));
    const char* names[] = {"bulk 40KB", "factor+invert", "-", "tri_solve", "64 dep DFMA", "16 dep rsqrt", "16 syncthreads"};
    for (int blocks : {1, 148 * 4}) {
        bench<<<blocks, 128, sizeof(Sm)>>>(dW, dout, 21);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[16]; cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
        printf("blocks=%d (%s):", blocks, cudaGetErrorString(e));
        for (int i = 0; i < 7; ++i) printf("  %s %lld", names[i], h[i]);
        printf("\n");
        unsigned long long tp[8], z[8] = {0};
        cudaMemcpyFromSymbol(tp, g_tile_prof, sizeof(tp)); cudaMemcpyToSymbol(g_tile_prof, z, sizeof(z));
        const double per = 1.0 / (4.0 * blocks * 21);   // summed over 4 warps, blocks CTAs, 21 reps (1 warm-up launch counted)
        printf("   per-warp avg cycles per factorisation: diag(owner) %.0f  shadow(non-owner) %.0f  wait1 %.0f  panel %.0f  wait2 %.0f  finish-col %.0f\n",
               tp[0] * per, tp[1] * per, tp[2] * per, tp[3] * per, tp[4] * per, tp[5] * per);
    }
    return 0;
}
