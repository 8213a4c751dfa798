// What the host-input copies of one tick cost: device time (events) and host enqueue time of each shape.
// nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/copy_bench tools/copy_bench.cu
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const size_t B = 4096, XS = 204, FS = 260, BR = 104;
    double *hx, *hf, *dx, *df; int *hflag, *dflag;
    cudaHostAlloc(&hx, B * XS * 8, 0); cudaHostAlloc(&hf, B * FS * 8, 0); cudaHostAlloc(&hflag, 1024, 0);
    cudaMalloc(&dx, B * XS * 8); cudaMalloc(&df, B * FS * 8); cudaMalloc(&dflag, 1024);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto&& body) {
        float best = 1e9f; double hbest = 1e9;
        for (int rep = 0; rep < 20; ++rep) {
            cudaStreamSynchronize(s);
            cudaEventRecord(e0, s);
            double t0 = now(); body(); double t1 = now();
            cudaEventRecord(e1, s); cudaStreamSynchronize(s);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 2) { best = ms < best ? ms : best; hbest = (t1 - t0) < hbest ? (t1 - t0) : hbest; }
        }
        printf("%-58s device %.3f ms   host enqueue %.1f us\n", name, best, hbest * 1e6);
    };
    run("xref contiguous 6.68 MB", [&] { cudaMemcpyAsync(dx, hx, B * XS * 8, cudaMemcpyHostToDevice, s); });
    run("fsteps contiguous 8.52 MB", [&] { cudaMemcpyAsync(df, hf, B * FS * 8, cudaMemcpyHostToDevice, s); });
    run("fsteps 2D 4096 rows x 832 B (pitch 2080) 3.41 MB", [&] { cudaMemcpy2DAsync(df, FS * 8, hf, FS * 8, BR * 8, B, cudaMemcpyHostToDevice, s); });
    run("fsteps 2D 4096 rows x 1024 B (pitch 2080)", [&] { cudaMemcpy2DAsync(df, FS * 8, hf, FS * 8, 1024, B, cudaMemcpyHostToDevice, s); });
    run("fsteps 2D 4096 rows x 512 B (pitch 2080)", [&] { cudaMemcpy2DAsync(df, FS * 8, hf, FS * 8, 512, B, cudaMemcpyHostToDevice, s); });
    run("xref + fsteps 2D", [&] { cudaMemcpyAsync(dx, hx, B * XS * 8, cudaMemcpyHostToDevice, s); cudaMemcpy2DAsync(df, FS * 8, hf, FS * 8, BR * 8, B, cudaMemcpyHostToDevice, s); });
    for (int nc : {2, 8, 16}) {
        char name[96]; snprintf(name, 96, "%d chunks x (xref + fsteps 2D + 4 B flag)", nc);
        run(name, [&] {
            const size_t n = B / nc;
            for (int c = 0; c < nc; ++c) {
                cudaMemcpyAsync(dx + c * n * XS, hx + c * n * XS, n * XS * 8, cudaMemcpyHostToDevice, s);
                cudaMemcpy2DAsync(df + c * n * FS, FS * 8, hf + c * n * FS, FS * 8, BR * 8, n, cudaMemcpyHostToDevice, s);
                cudaMemcpyAsync(dflag + c, hflag + c, 4, cudaMemcpyHostToDevice, s);
            }
        });
    }
    run("16 x 4 B flag copies", [&] { for (int c = 0; c < 16; ++c) cudaMemcpyAsync(dflag + c, hflag + c, 4, cudaMemcpyHostToDevice, s); });
    run("forces back 0.39 MB", [&] { cudaMemcpyAsync(hx, dx, B * 12 * 8, cudaMemcpyDeviceToHost, s); });
    return 0;
}
