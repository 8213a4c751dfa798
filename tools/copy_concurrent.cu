// Host <-> device copy bandwidth with 1, 2, 4, ... GPUs of one box copying AT THE SAME TIME (one host thread and one
// page-locked buffer per GPU): what the end-to-end path of bench.py (10.1 MB in, 0.39 MB out per 4096-robot tick and GPU)
// can get from the host's memory / PCIe path when every rank copies at once.  Prints one JSON line per width.
// nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/copy_concurrent tools/copy_concurrent.cu -lpthread
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Result { double h2d_gbs, d2h_gbs, tick_ms_p50, tick_ms_p99; };

int main(int argc, char** argv) {
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    const size_t IN = 10092544, OUT = 393216;               // bytes per tick and GPU of bench.py's e2e leg
    const int TICKS = 400;
    for (int width = 1; width <= ndev; width *= 2) {
        std::vector<Result> res(width);
        std::atomic<int> arrived{0};
        std::atomic<bool> go{false};
        std::vector<std::thread> th;
        for (int g = 0; g < width; ++g) {
            th.emplace_back([&, g] {
                cudaSetDevice(g);
                char *h_in, *h_out, *d_in, *d_out;
                cudaHostAlloc(&h_in, IN, cudaHostAllocPortable); cudaHostAlloc(&h_out, OUT, cudaHostAllocPortable);
                memset(h_in, 1, IN);
                cudaMalloc(&d_in, IN); cudaMalloc(&d_out, OUT);
                cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
                cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
                for (int i = 0; i < 5; ++i) { cudaMemcpyAsync(d_in, h_in, IN, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s); }
                arrived.fetch_add(1);
                while (!go.load()) std::this_thread::yield();
                std::vector<double> ticks(TICKS);
                double tin = 0, tout = 0;
                for (int i = 0; i < TICKS; ++i) {
                    const double t0 = now();
                    cudaEventRecord(e0, s);
                    cudaMemcpyAsync(d_in, h_in, IN, cudaMemcpyHostToDevice, s);
                    cudaEventRecord(e1, s);
                    cudaMemcpyAsync(h_out, d_out, OUT, cudaMemcpyDeviceToHost, s);
                    cudaEventRecord(e2, s);
                    cudaStreamSynchronize(s);
                    ticks[i] = (now() - t0) * 1e3;
                    float a, b; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
                    tin += a; tout += b;
                }
                std::sort(ticks.begin(), ticks.end());
                res[g] = {IN * TICKS / (tin * 1e-3) * 1e-9, OUT * TICKS / (tout * 1e-3) * 1e-9, ticks[TICKS / 2], ticks[TICKS * 99 / 100]};
                cudaFreeHost(h_in); cudaFreeHost(h_out); cudaFree(d_in); cudaFree(d_out);
            });
        }
        while (arrived.load() < width) std::this_thread::yield();
        go.store(true);
        for (auto& t : th) t.join();
        double sum = 0, mn = 1e30, p50 = 0, p99 = 0;
        for (auto& r : res) { sum += r.h2d_gbs; mn = r.h2d_gbs < mn ? r.h2d_gbs : mn; p50 = r.tick_ms_p50 > p50 ? r.tick_ms_p50 : p50; p99 = r.tick_ms_p99 > p99 ? r.tick_ms_p99 : p99; }
        printf("{\"gpus_copying\": %d, \"h2d_gbs_per_gpu_mean\": %.2f, \"h2d_gbs_per_gpu_min\": %.2f, \"h2d_gbs_aggregate\": %.1f, "
               "\"d2h_gbs_per_gpu_mean\": %.2f, \"copy_only_tick_ms_p50_max_over_gpus\": %.4f, \"copy_only_tick_ms_p99_max_over_gpus\": %.4f, "
               "\"bytes_in\": %zu, \"bytes_out\": %zu, \"ticks\": %d}\n",
               width, sum / width, mn, sum, res[0].d2h_gbs, p50, p99, IN, OUT, TICKS);
        fflush(stdout);
    }
    return 0;
}
