// FP64 microbenchmarks for B200 (sm_100a): DFMA and DMMA (mma.sync f64) throughput / latency, and
// shared-memory load bandwidth.  These give the roofline denominators SURVEY.md 8(d) says are
// missing from MEASURED_PEAKS.json (FP64 pipe, smem), and decide how the factorisation is mapped.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int j = 0; j < NACC; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) dmma884(c[j][0], c[j][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#if 1
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
template <int NACC>
__global__ void dmma1688_kernel(double* out, int iters, double a, double b) {
    double c[NACC][4];
    double av[4] = {a, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
#pragma unroll
    for (int j = 0; j < NACC; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; c[j][2] = 1; c[j][3] = 2; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) dmma1688(c[j], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

// DMMA and DFMA interleaved: do the two overlap (separate pipes) or share one?
__global__ void mixed_kernel(double* out, int iters, double a, double b) {
    double c[4][2];
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = 4, x5 = 5, x6 = 6, x7 = 7;
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dmma884(c[j][0], c[j][1], a, b);
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// latency of a dependent chain, one warp
__global__ void latency_kernel(long long* out, double a, double b, double* sink) {
    double c0 = threadIdx.x, c1 = 1.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) dmma884(c0, c1, a, b);
    long long t1 = clock64();
    double x = threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) x = fma(x, a, b);
    long long t2 = clock64();
    double r = x;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) r = 1.0 / (r + 1.5);
    long long t3 = clock64();
    double q = x + 2.0;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) q = sqrt(q + 1.5);
    long long t4 = clock64();
    double w = x;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) w = __shfl_xor_sync(0xffffffffu, w, 1) + 1.0;
    long long t5 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; }
    sink[threadIdx.x] = c0 + c1 + x + r + q + w;
}

// shared-memory read bandwidth: every thread streams LDS.64 / LDS.128 over a 32 KB buffer
template <int VEC>
__global__ void lds_kernel(double* out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int base = threadIdx.x * VEC;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int idx = (base + u * 512 + it) & 4095 & ~(VEC - 1);
            if (VEC == 1) { s0 += sm[idx]; }
            else { double2 v = *reinterpret_cast<double2*>(&sm[idx]); s0 += v.x; s1 += v.y; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}

template <typename F>
float time_it(F launch, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 64 * 1024));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int threads = 32 * warps; if (threads > 1024) continue;
        int blocks = sms;
        float ms = time_it([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * 8 * iters * (double)threads * blocks;
        printf("{\"test\": \"dfma\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, fl / ms * 1e-9);
    }
    for (int warps : {4, 8, 16}) {
        int threads = 32 * warps, blocks = sms;
        float ms = time_it([&] { dmma_kernel<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * 256 * 4 * iters * (double)warps * blocks;
        printf("{\"test\": \"dmma_m8n8k4_x4\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, fl / ms * 1e-9);
        ms = time_it([&] { dmma_kernel<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 256 * 1 * iters * (double)warps * blocks;
        printf("{\"test\": \"dmma_m8n8k4_x1\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, fl / ms * 1e-9);
        ms = time_it([&] { dmma1688_kernel<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        fl = 2.0 * 1024 * 2 * iters * (double)warps * blocks;
        printf("{\"test\": \"dmma_m16n8k8_x2\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, fl / ms * 1e-9);
        ms = time_it([&] { mixed_kernel<<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); });
        fl = 2.0 * (256 * 4 + 32 * 8 * 4) * (iters / 4) * (double)warps * blocks;
        printf("{\"test\": \"mixed_dmma_dfma\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, fl / ms * 1e-9);
    }
    {
        long long* lat; CK(cudaMalloc(&lat, 5 * sizeof(long long)));
        latency_kernel<<<1, 32>>>(lat, 1.0000001, 1e-9, out); CK(cudaDeviceSynchronize());
        long long h[5]; CK(cudaMemcpy(h, lat, sizeof(h), cudaMemcpyDeviceToHost));
        printf("{\"test\": \"latency_cycles\", \"dmma_m8n8k4\": %.1f, \"dfma\": %.1f, \"ddiv\": %.1f, \"dsqrt\": %.1f, \"shfl64_plus_dadd\": %.1f}\n",
               h[0] / 256.0, h[1] / 256.0, h[2] / 64.0, h[3] / 64.0, h[4] / 256.0);
    }
    for (int warps : {8, 16, 32}) {
        int threads = 32 * warps, blocks = sms, it2 = 4000;
        float ms = time_it([&] { lds_kernel<1><<<blocks, threads, 32768>>>(out, it2); });
        double bytes = 8.0 * 8 * it2 * (double)threads * blocks;
        printf("{\"test\": \"lds64\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tb_per_s\": %.3f}\n", warps, ms, bytes / ms * 1e-9);
        ms = time_it([&] { lds_kernel<2><<<blocks, threads, 32768>>>(out, it2); });
        bytes = 16.0 * 8 * it2 * (double)threads * blocks;
        printf("{\"test\": \"lds128\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tb_per_s\": %.3f}\n", warps, ms, bytes / ms * 1e-9);
    }
    return 0;
}
