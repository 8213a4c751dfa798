// mpcqp_api.cu -- host side of libmpcqp.so: the C ABI declared in include/mpcqp.h.
// Owns device memory, the stream and the carried warm-start state of a batch of MPC instances and
// launches the kernels of mpcqp_kernels.cu.  There is deliberately no CPU code path: without a
// CUDA device mpcqp_create fails with MPCQP_ERR_NO_DEVICE.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/mpcqp.h"
#include "mpcqp_kernels.cu"

using namespace mpcqp;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(MPCQP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
    } while (0)

// invert an n x n SPD matrix in extended precision (Gauss-Jordan, no pivoting needed)
void invert_spd(std::vector<long double>& a, int n) {
    std::vector<long double> inv((size_t)n * n, 0.0L);
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0L;
    for (int p = 0; p < n; ++p) {
        const long double d = 1.0L / a[(size_t)p * n + p];
        for (int c = 0; c < n; ++c) { a[(size_t)p * n + c] *= d; inv[(size_t)p * n + c] *= d; }
        for (int r = 0; r < n; ++r) {
            if (r == p) continue;
            const long double m = a[(size_t)r * n + p];
            if (m == 0.0L) continue;
            for (int c = 0; c < n; ++c) {
                a[(size_t)r * n + c] -= m * a[(size_t)p * n + c];
                inv[(size_t)r * n + c] -= m * inv[(size_t)p * n + c];
            }
        }
    }
    a.swap(inv);
}

}  // namespace

template <int N>
cudaError_t configure_kernels() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(solve_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<N, false>)))) return e;
    if ((e = cudaFuncSetAttribute(solve_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<N, true>)))) return e;
    if ((e = cudaFuncSetAttribute(solve_kernel<N, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
    return cudaFuncSetAttribute(solve_kernel<N, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

// Stage-wise solver: one translation unit per horizon (mpcqp_ric_inst.cu); in a single-TU build (profiling) they are included here.
#ifdef MPCQP_SINGLE_TU
#include "mpcqp_riccati.cuh"
#define RIC_N 16
#include "mpcqp_ric_inst.cu"
#define RIC_N 32
#include "mpcqp_ric_inst.cu"
#define RIC_N 64
#include "mpcqp_ric_inst.cu"
#else
#include "mpcqp_ric_consts.h"
#endif
namespace mpcqp {
#define RIC_DECL(N)                                                                                                   \
    cudaError_t ric_configure_##N(int* ctas_per_sm, int* ipm_ctas_per_sm);                                                                \
    void ric_launch_##N(int grid, cudaStream_t s, const DevParams& dp, const DevState& st, const DevScenario& sc,      \
                        const double* dx, const double* df, double* ws, int* ctr, int first, int off, int n_inst);     \
    void ipm_launch_##N(int grid, cudaStream_t s, const DevParams& dp, const DevState& st, const DevScenario& sc,      \
                        const double* dx, const double* df, double* ws, int first, int pdl);
RIC_DECL(16) RIC_DECL(32) RIC_DECL(64)
#undef RIC_DECL
// one robot per lane (mpcqp_lane_inst.cu): the active-set stage for large batches, and the planner kernel in front of it
cudaError_t lane_configure(int* ctas_per_sm);
size_t lane_ws_bytes(int grid, int n);
void lane_launch(int grid, cudaStream_t s, const DevParams& dp, const DevState& st, const DevScenario& sc, const double* dx,
                 const double* df, double* ws, int* ctr, int first, int off, int n_inst);
void plan_launch(int sms, cudaStream_t s, const DevParams& dp, const DevScenario& sc, int off, int n_inst);
}  // namespace mpcqp

// the capacity class (compiled instantiation) that holds a horizon of n steps
static int ric_capacity(int n) { return n <= 16 ? 16 : (n <= 32 ? 32 : 64); }

cudaError_t launch_stagewise(int cap, int n_inst, int max_ctas, cudaStream_t s, const DevParams& dp, const DevState& st, const DevScenario& sc,
                             const double* dx, const double* df, double* ws, int* ctr, bool reset_ctr, int first, int off) {
    int grid = (n_inst + RIC_PER_CTA - 1) / RIC_PER_CTA;
    if (grid > max_ctas) grid = max_ctas;                      // persistent: one workspace slot per resident half-warp
    // the kernel's work counter (pairs of robots beyond the first per warp) must be zero at launch: on the handle's main stream
    // the caller's reset of the fallback queue covers it (one 16-byte memset), on a side stream it is reset here
    if (reset_ctr) {
        const cudaError_t e = cudaMemsetAsync(ctr, 0, sizeof(int), s);
        if (e != cudaSuccess) return e;
    }
    if (cap == 16) ric_launch_16(grid, s, dp, st, sc, dx, df, ws, ctr, first, off, n_inst);
    else if (cap == 32) ric_launch_32(grid, s, dp, st, sc, dx, df, ws, ctr, first, off, n_inst);
    else ric_launch_64(grid, s, dp, st, sc, dx, df, ws, ctr, first, off, n_inst);
    return cudaSuccess;
}

void launch_ipm(int cap, int max_ctas, cudaStream_t s, const DevParams& dp, const DevState& st, const DevScenario& sc,
                const double* dx, const double* df, double* ws, int first, int pdl) {
    if (cap == 16) ipm_launch_16(max_ctas, s, dp, st, sc, dx, df, ws, first, pdl);
    else if (cap == 32) ipm_launch_32(max_ctas, s, dp, st, sc, dx, df, ws, first, pdl);
    else ipm_launch_64(max_ctas, s, dp, st, sc, dx, df, ws, first, pdl);
}

template <int N>
void launch_solve(bool admm, int grid, cudaStream_t s, const DevParams& dp, const DevState& st, const DevScenario& sc,
                  const double* dx, const double* df, int first, int off, int n) {
    if (admm) {
        // fallback stage: programmatic dependent launch behind the active-set kernel of the same stream (the kernel blocks in
        // griddepcontrol.wait before it reads the queue); behind anything else the attribute is an ordinary launch
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32 * Cfg<N>::NW); cfg.dynamicSmemBytes = sizeof(Smem<N, true>); cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, solve_kernel<N, true>, dp, st, sc, dx, df, first, off, n);
    }
    else solve_kernel<N, false><<<grid, 32 * Cfg<N>::NW, sizeof(Smem<N, false>), s>>>(dp, st, sc, dx, df, first, off, n);
}

constexpr int MAX_RANGES = 8;        // index ranges of a batch that may advance through the ticks independently (mpcqp_set_overlap)

struct mpcqp_handle {
    mpcqp_params p;
    DevParams dp;
    DevState st;
    cudaStream_t stream = nullptr;
    // side streams.  Host-input runs: copy + solve of alternate chunks overlap on side[0], side[1] (joined every tick).
    // Overlapped index ranges: range r lives on side[r] from tick to tick with its own counters, queue region and workspace.
    cudaStream_t side[MAX_RANGES] = {};
    cudaEvent_t ev_main = nullptr, ev_side[MAX_RANGES] = {};
    int ranges = 0;                 // mpcqp_set_overlap: 0 = automatic inside mpcqp_scenario_run only, 1 = off, 2 .. MAX_RANGES
    bool forked = false;            // ranges are in flight on the side streams: the main stream has not been joined yet
    int forked_ranges = 0;
    int rparity[MAX_RANGES] = {};
    double* d_xref = nullptr;
    double* d_fsteps = nullptr;
    double* d_Minv = nullptr;
    double* d_C2 = nullptr;
    void* d_block = nullptr;        // one allocation behind all DevState arrays
    size_t block_bytes = 0;
    int aw = 0, cw = 0;             // words per instance of the active / contact masks
    bool ran = false;
    int32_t* ctr_base = nullptr;    // two copies of {fallback queue length, three work counters}, 8 ints apart, alternating by tick
    int ctr_parity = 0;
    bool ctr_clean = false;         // this tick's copy was zeroed by the previous tick's stage-wise launch
    bool zero_next = false;         // set by solve(): a main-stream stage-wise launch of this tick zeroes the next tick's copy
    int64_t launches = 0;
    int sms = 0;
    bool direct_ok = true;          // host inputs in page-locked memory are read by the kernels directly (MPCQP_NO_DIRECT=1 switches it off)
    int last_ranges = 1;            // how the last tick was issued (mpcqp_get_fallback_count sums the ranges' queues)

    double* pin[2] = {nullptr, nullptr};            // asynchronous result slots (pinned host memory)
    cudaEvent_t ev_pin[2] = {nullptr, nullptr};
    cudaEvent_t ev_pin_r[2][MAX_RANGES] = {};       // ... when the slot was filled range by range (overlapped ticks)
    int pin_ranges[2] = {0, 0};                     // 0: the slot's copy ran on the main stream, R: on the streams of R index ranges
    bool pin_valid[2] = {false, false};

    DevScenario sc;                 // device-resident closed loop (disabled unless mpcqp_scenario_init was called)
    void* d_scen = nullptr;
    int scen_tick = 0;
    bool scen_ready = false;

    cudaError_t launch_err = cudaSuccess;           // first error of a launch helper since the last check

    // fallback stage of this tick for the robots in the queue: the interior-point kernel (stage-wise path) or the dense ADMM kernel
    bool fallback_is_ipm() const { return (p.mode & MPCQP_MODE_STAGEWISE) && (p.mode & MPCQP_MODE_IPM); }
    bool has_fallback() const { return fallback_is_ipm() || (p.mode & MPCQP_MODE_ADMM); }

    // Large batches: the active-set stage runs one robot per lane (mpcqp_lane.cuh) instead of half a warp per robot
    bool use_lane() const {
        if (!(p.mode & MPCQP_MODE_STAGEWISE) || !fallback_is_ipm()) return false;
        return (p.mode & MPCQP_MODE_LANE) || p.batch >= lane_min;
    }
    // the lane kernel's workspace of side stream `which` (0 = main), grown on demand (first tick only)
    cudaError_t lane_ws(int which, int grid, double** out) {
        const size_t need = lane_ws_bytes(grid, p.n_steps);
        if (need > lane_ws_cap[which]) {
            if (d_lane_ws[which]) { cudaDeviceSynchronize(); cudaFree(d_lane_ws[which]); }
            d_lane_ws[which] = nullptr; lane_ws_cap[which] = 0;
            const cudaError_t e = cudaMalloc(&d_lane_ws[which], need);
            if (e != cudaSuccess) return e;
            lane_ws_cap[which] = need;
        }
        *out = d_lane_ws[which];
        return cudaSuccess;
    }

    void solve(bool admm, int grid, cudaStream_t s, const double* dx, const double* df, int first, int off, int n,
               bool closed_loop = false) {
        DevScenario use = sc;
        use.enabled = closed_loop ? 1 : 0;
        const bool lane = use_lane();
        if (lane && closed_loop) {
            // the planner runs as its own kernel and leaves xref / fsteps in the handle's input buffers; the solve kernels read them
            // like a caller's (enabled == 2: inputs from memory, integration at the end of the solve)
            use.enabled = 2; use.xref_out = d_xref; use.fsteps_out = d_fsteps;
            dx = d_xref; df = d_fsteps;
        }
        if (!admm && lane) {
            const int lane_of_stream = s == side[0] ? 1 : (s == side[1] ? 2 : 0);
            int g = (n + 31) / 32;
            if (g > lane_max_ctas) g = lane_max_ctas;
            double* ws = nullptr;
            cudaError_t e = lane_ws(lane_of_stream, g, &ws);
            DevState stl = st;
            int* ctr = st.fb_count + 1 + lane_of_stream;
            if (lane_of_stream == 0) { stl.fb_next = ctr_base + 8 * (ctr_parity ^ 1); zero_next = true; }
            else if (e == cudaSuccess) e = cudaMemsetAsync(ctr, 0, sizeof(int), s);
            if (e == cudaSuccess) {
                if (closed_loop) { plan_launch(sms, s, dp, use, off, n); ++launches; }
                lane_launch(g, s, dp, stl, use, dx, df, ws, ctr, first, off, n);
            }
            if (e != cudaSuccess && launch_err == cudaSuccess) launch_err = e;
            ++launches;
            return;
        }
        if (admm && fallback_is_ipm()) {
            // the main stream's workspace: every active-set launch of this tick has been joined into `s` by now
            launch_ipm(ric_capacity(p.n_steps), ipm_max_ctas, s, dp, st, use, dx, df, d_ric_ws, first, 1);
        } else if (!admm && (p.mode & MPCQP_MODE_STAGEWISE)) {
            // active-set stage on the stage-wise factorisation: half a warp per robot, persistent grid; every stream
            // that may run it concurrently has its own gain workspace
            const int lane_of_stream = s == side[0] ? 1 : (s == side[1] ? 2 : 0);
            double* ws = d_ric_ws + (size_t)lane_of_stream * ric_ws_doubles;
            DevState stl = st;
            if (lane_of_stream == 0) { stl.fb_next = ctr_base + 8 * (ctr_parity ^ 1); zero_next = true; }
            const cudaError_t e = launch_stagewise(ric_capacity(p.n_steps), n, ric_max_ctas, s, dp, stl, use, dx, df, ws,
                                                   st.fb_count + 1 + lane_of_stream, lane_of_stream != 0, first, off);
            if (e != cudaSuccess && launch_err == cudaSuccess) launch_err = e;
        } else if (p.n_steps == 16) launch_solve<16>(admm, grid, s, dp, st, use, dx, df, first, off, n);
        else launch_solve<32>(admm, grid, s, dp, st, use, dx, df, first, off, n);
        ++launches;
    }
    // One tick of index range r = robots off .. off + n - 1 on its own stream: the active-set kernel, then (ordinary launch: its CTAs
    // must not sit resident while the range's last sweeps finish) the interior-point kernel over the range's own fallback queue.
    // Counters: 16 ints per range behind the main stream's (two copies of {queue length, work counter, 0, 0}, alternating by tick;
    // the active-set kernel clears the copy of the next tick); queue: the range's slice of fb_list; workspace copy 1 + r.
    void solve_range(int r, int off, int n, const double* dx, const double* df, int first, bool closed_loop) {
        DevScenario use = sc;
        use.enabled = closed_loop ? 1 : 0;
        int32_t* ctr = ctr_base + 16 + 16 * r;
        rparity[r] ^= 1;
        DevState stl = st;
        stl.fb_count = ctr + 8 * rparity[r];
        stl.fb_next = ctr + 8 * (rparity[r] ^ 1);
        stl.fb_list = st.fb_list + off;
        double* ws = d_ric_ws + (size_t)(1 + r) * ric_ws_doubles;
        const int cap = ric_capacity(p.n_steps);
        const cudaError_t e = launch_stagewise(cap, n, ric_max_ctas, side[r], dp, stl, use, dx, df, ws, stl.fb_count + 1, false, first, off);
        if (e != cudaSuccess && launch_err == cudaSuccess) launch_err = e;
        ++launches;
        if (fallback_is_ipm()) {
            const int want = (n + RIC_PER_CTA - 1) / RIC_PER_CTA;
            launch_ipm(cap, want < ipm_max_ctas ? want : ipm_max_ctas, side[r], dp, stl, use, dx, df, ws, first, 0);
            ++launches;
        }
    }
    // ranges to use for a tick issued now (1 = the ordinary single-stream tick)
    int ranges_for(bool automatic_ok) const {
        if (!(p.mode & MPCQP_MODE_STAGEWISE) || (has_fallback() && !fallback_is_ipm()) || use_lane()) return 1;
        int r = ranges;
        if (r == 0) {
            // automatic: eight ranges when the batch is one to three waves of resident robots, two up to eight waves -- there the last sweeps of a tick
            // leave most of the GPU idle (a robot that needs a second sweep ends its tick a whole sweep after the others)
            if (!automatic_ok) return 1;
            r = p.batch <= wave() ? 1 : (p.batch <= 3 * wave() ? 8 : (p.batch <= 8 * wave() ? 2 : 1));      // measured (M solves/s with 1 / 2 / 4 / 8 ranges): 4096 robots 25.6 / 31.3 / 32.4 / 33.3, 8192: 32.9 / 40.2 / 37.7 / 37.5, N = 32 4096: 9.5 / 11.8 / 12.9 / 13.3
            if (const char* e = std::getenv("MPCQP_RANGES")) { const int c = std::atoi(e); if (c >= 1 && c <= MAX_RANGES) r = c; }      // tuning hook
        }
        while (r > 1 && p.batch < 2 * RIC_PER_CTA * r) --r;
        return r;
    }
    int ctas_per_sm(bool admm) const { return p.n_steps == 16 ? (admm ? 2 : 4) : 1; }
    // robots of the active-set stage that are resident at once (one wave)
    int wave(void) const { return (p.mode & MPCQP_MODE_STAGEWISE) ? ric_max_ctas * RIC_PER_CTA : ctas_per_sm(false) * sms; }

    int32_t* d_all = nullptr;       // ADMM-only mode: 0 .. B-1, then B (the queue that holds every robot)
    double* d_scratch = nullptr;    // device staging of the diagnostics that are computed on request (export_build, cost components)
    size_t scratch_bytes = 0;
    cudaError_t scratch(size_t bytes) {
        if (bytes <= scratch_bytes) return cudaSuccess;
        if (d_scratch) cudaFree(d_scratch);
        d_scratch = nullptr; scratch_bytes = 0;
        const cudaError_t e = cudaMalloc(&d_scratch, bytes);
        if (e == cudaSuccess) scratch_bytes = bytes;
        return e;
    }

    double* d_lane_ws[3] = {nullptr, nullptr, nullptr};     // one robot per lane: workspace of the main stream and the two host-input side streams
    size_t lane_ws_cap[3] = {0, 0, 0};
    int lane_max_ctas = 0;
    int lane_min = 1 << 30;         // batch from which the active-set stage runs one robot per lane
    double* d_ric_ws = nullptr;     // stage-wise path: per-stage gains of the resident robots, x3 (main + two side streams)
    size_t ric_ws_doubles = 0;
    int ric_max_ctas = 0, ipm_max_ctas = 0;
};

extern "C" {

const char* mpcqp_last_error(void) { return g_err.c_str(); }
const char* mpcqp_version(void) { return "mpcqp 0.1 (sm_100a)"; }

void mpcqp_default_params(mpcqp_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->struct_size = (int32_t)sizeof(*p);
    p->n_steps = 16;
    p->batch = 1;
    p->device = 0;
    p->dt = 0.02;
    p->T_gait = 0.32;
    p->mass = 2.50000279;                                   // MPC.py:28
    p->mu = 0.9;                                            // MPC.py:39
    p->fz_max = 25.0;                                       // MPC.py:228
    p->gravity = 9.81;                                      // MPC.py:201
    const double gI[9] = {3.09249e-2, -8.00101e-7, 1.865287e-5,          // MPC.py:35-37
                          -8.00101e-7, 5.106100e-2, 1.245813e-4,
                          1.865287e-5, 1.245813e-4, 6.939757e-2};
    std::memcpy(p->gI, gI, sizeof(gI));
    const double fh[12] = {0.19, 0.19, -0.19, -0.19,                     // MPC.py:67-70
                           0.15005, -0.15005, 0.15005, -0.15005,
                           0.0, 0.0, 0.0, 0.0};
    std::memcpy(p->footholds, fh, sizeof(fh));
    // MPC.py:255-275
    p->w_state[0] = 0.1; p->w_state[1] = 0.1; p->w_state[2] = 1.0;
    p->w_state[3] = 0.11; p->w_state[4] = 0.11; p->w_state[5] = 0.11;
    for (int i = 0; i < 3; ++i) p->w_state[6 + i] = 2.0 * std::sqrt(p->w_state[i]);
    for (int i = 0; i < 3; ++i) p->w_state[9 + i] = 0.05 * std::sqrt(p->w_state[3 + i]);
    p->w_force = 1e-5;                                      // MPC.py:282-284
    p->mode = MPCQP_MODE_ACTIVE_SET | MPCQP_MODE_STAGEWISE | MPCQP_MODE_IPM;
    p->max_sweeps = 16;
    p->max_iter = 1000;
    p->min_iter = 10;
    p->check_every = 5;
    p->warm_start = 1;
    p->rho = 5e-5;
    p->sigma = 1e-6;
    p->alpha = 1.6;
    p->feas_tol = 1e-9;
    p->dual_tol = 1e-12;
    p->refine = 0;
    p->ipm_max_iter = 60;
}

int mpcqp_destroy(mpcqp_handle* h) {
    if (!h) return MPCQP_OK;
    cudaSetDevice(h->p.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_xref);
    cudaFree(h->d_fsteps);
    cudaFree(h->d_Minv);
    cudaFree(h->d_C2);
    cudaFree(h->d_block);
    cudaFree(h->d_scen);
    cudaFree(h->d_ric_ws);
    for (int i = 0; i < 3; ++i) cudaFree(h->d_lane_ws[i]);
    cudaFree(h->d_all);
    cudaFree(h->d_scratch);
    for (int i = 0; i < 2; ++i) {
        if (h->pin[i]) cudaFreeHost(h->pin[i]);
        if (h->ev_pin[i]) cudaEventDestroy(h->ev_pin[i]);
        for (int r = 0; r < MAX_RANGES; ++r)
            if (h->ev_pin_r[i][r]) cudaEventDestroy(h->ev_pin_r[i][r]);
    }
    for (int i = 0; i < MAX_RANGES; ++i) {
        if (h->side[i]) { cudaStreamSynchronize(h->side[i]); cudaStreamDestroy(h->side[i]); }
        if (h->ev_side[i]) cudaEventDestroy(h->ev_side[i]);
    }
    if (h->ev_main) cudaEventDestroy(h->ev_main);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MPCQP_OK;
}

int mpcqp_create(const mpcqp_params* p, mpcqp_handle** out) {
    if (!p || !out) return fail(MPCQP_ERR_INVALID, "null argument");
    *out = nullptr;
    if (p->struct_size != (int32_t)sizeof(mpcqp_params)) return fail(MPCQP_ERR_INVALID, "mpcqp_params.struct_size mismatch");
    if (p->n_steps < 1 || p->n_steps > 64)
        return fail(MPCQP_ERR_INVALID, "n_steps: horizons of 1 .. 64 steps are supported");
    // the dense kernels (MPCQP_MODE_ACTIVE_SET without MPCQP_MODE_STAGEWISE, and the MPCQP_MODE_ADMM stage) exist for N = 16 and 32 only
    const bool dense_ok = p->n_steps == 16 || p->n_steps == 32;
    const bool ipm_fallback = (p->mode & MPCQP_MODE_STAGEWISE) && (p->mode & MPCQP_MODE_IPM);
    const bool needs_dense = !(p->mode & MPCQP_MODE_STAGEWISE) || ((p->mode & MPCQP_MODE_ADMM) && !ipm_fallback);
    if (needs_dense && !dense_ok)
        return fail(MPCQP_ERR_INVALID, "the dense stages exist for n_steps = 16 and 32 only: use MPCQP_MODE_ACTIVE_SET | MPCQP_MODE_STAGEWISE | MPCQP_MODE_IPM");
    if ((p->mode & MPCQP_MODE_IPM) && !(p->mode & MPCQP_MODE_STAGEWISE))
        return fail(MPCQP_ERR_INVALID, "MPCQP_MODE_IPM is the fallback stage of the stage-wise path: set MPCQP_MODE_STAGEWISE too");
    if ((p->mode & MPCQP_MODE_STAGEWISE) && !(p->mode & MPCQP_MODE_ACTIVE_SET))
        return fail(MPCQP_ERR_INVALID, "MPCQP_MODE_STAGEWISE selects the factorisation of the active-set stage: set MPCQP_MODE_ACTIVE_SET too");
    if (p->batch < 1) return fail(MPCQP_ERR_INVALID, "batch must be >= 1");
    if (!(p->dt > 0) || !(p->mass > 0) || !(p->mu > 0) || !(p->fz_max > 0) || !(p->w_force > 0))
        return fail(MPCQP_ERR_INVALID, "dt, mass, mu, fz_max, w_force must be positive");
    for (int i = 0; i < 12; ++i)
        if (!(p->w_state[i] > 0)) return fail(MPCQP_ERR_INVALID, "state weights must be positive");
    if (!(p->mode & (MPCQP_MODE_ACTIVE_SET | MPCQP_MODE_ADMM))) return fail(MPCQP_ERR_INVALID, "mode selects no solver stage");
    if (!(p->rho > 0) || !(p->sigma >= 0) || !(p->alpha > 0 && p->alpha < 2)) return fail(MPCQP_ERR_INVALID, "bad ADMM settings");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MPCQP_ERR_NO_DEVICE, "no CUDA device visible: libmpcqp has no CPU fallback");
    }
    if (p->device < 0 || p->device >= ndev) return fail(MPCQP_ERR_INVALID, "device ordinal out of range");
    CU(cudaSetDevice(p->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, p->device));
    if (prop.major < 10) return fail(MPCQP_ERR_NO_DEVICE, "libmpcqp is built for sm_100a (B200) only");

    mpcqp_handle* h = new (std::nothrow) mpcqp_handle();
    if (!h) return fail(MPCQP_ERR_INVALID, "out of host memory");
    h->p = *p;
    std::memset(&h->sc, 0, sizeof(h->sc));
    h->sms = prop.multiProcessorCount;
    const int N = p->n_steps, B = p->batch;
    DevParams& d = h->dp;
    std::memset(&d, 0, sizeof(d));
    d.N = N; d.batch = B;
    d.dt = p->dt; d.mass = p->mass; d.mu = p->mu; d.fz_max = p->fz_max; d.gravity = p->gravity; d.w_force = p->w_force;
    {
        std::vector<long double> g(9);
        for (int i = 0; i < 9; ++i) g[i] = p->gI[i];
        invert_spd(g, 3);
        for (int i = 0; i < 9; ++i) d.gIinv[i] = (double)g[i];
    }
    std::memcpy(d.footholds, p->footholds, sizeof(d.footholds));
    for (int c = 0; c < 6; ++c) { d.wp[c] = p->w_state[c]; d.wv[c] = p->w_state[6 + c]; }
    d.rho = p->rho; d.sigma = p->sigma; d.alpha = p->alpha; d.feas_tol = p->feas_tol; d.dual_tol = p->dual_tol;
    d.max_sweeps = p->max_sweeps; d.max_iter = p->max_iter; d.min_iter = p->min_iter > 0 ? p->min_iter : 1;
    d.check_every = p->check_every > 0 ? p->check_every : 1;
    d.warm_start = p->warm_start; d.mode = p->mode; d.refine = p->refine > 0 ? 1 : 0;
#ifdef MPCQP_CANARY
    d.refine = p->refine;           // debug build: refine == 77 makes robot 0 write one element past its state array (detector self-test)
#endif
    d.ipm_max_iter = p->ipm_max_iter > 0 ? p->ipm_max_iter : 60;
    d.fs_rows = 20;
    // the reciprocals make_face needs, with the kernels' former arithmetic: mu * mu rounded, times m exact, plus one rounded, times w rounded, IEEE division
    d.inv_wf = 1.0 / d.w_force;
    for (int m = 0; m < 3; ++m) {
        volatile double t = d.mu * d.mu;
        volatile double s1 = 1.0 + t * (double)m;
        volatile double x = d.w_force * s1;
        d.dz_tab[m] = 1.0 / x;
    }

    // Gram matrices of the double-integrator response and their inverses (constant per handle)
    //   M_c[k,l] = sum_{i >= max(k,l)} ( dt^2 Qp_c (i-k)(i-l) + Qv_c ),  i = 0..N-1
    // (dense kernels only: 6 N must fill whole 8 x 8 tiles)
    const int n = 6 * N, NT = (n + 7) / 8, NTILES = NT * (NT + 1) / 2;
    std::vector<double> C2((size_t)N * N), Mt((size_t)NTILES * 64, 0.0);
    for (int c = 0; c < 6 && dense_ok; ++c) {
        std::vector<long double> m((size_t)N * N);
        for (int k = 0; k < N; ++k)
            for (int l = 0; l < N; ++l) {
                long double c0 = 0, c2 = 0;
                for (int i = (k > l ? k : l); i < N; ++i) { c0 += 1; c2 += (long double)(i - k) * (i - l); }
                m[(size_t)k * N + l] = (long double)p->dt * p->dt * p->w_state[c] * c2 + (long double)p->w_state[6 + c] * c0;
                C2[(size_t)k * N + l] = (double)c2;
            }
        invert_spd(m, N);
        for (int k = 0; k < N; ++k)
            for (int l = 0; l < N; ++l) {
                const int gi = 6 * k + c, gj = 6 * l + c;
                if ((gi >> 3) < (gj >> 3)) continue;                 // lower block triangle (+ full diagonal tiles)
                Mt[(size_t)tile_index(gi >> 3, gj >> 3) * 64 + elem_off(gi & 7, gj & 7)] = (double)m[(size_t)k * N + l];
            }
    }
    auto bail = [&](int code) { mpcqp_destroy(h); return code; };
#define CUH(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return bail(fail(MPCQP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); \
    } while (0)
    CUH(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (int i = 0; i < MAX_RANGES; ++i) {
        CUH(cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking));
        CUH(cudaEventCreateWithFlags(&h->ev_side[i], cudaEventDisableTiming));
    }
    CUH(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
    CUH(cudaMalloc(&h->d_C2, C2.size() * sizeof(double)));
    CUH(cudaMalloc(&h->d_Minv, Mt.size() * sizeof(double)));
    CUH(cudaMemcpy(h->d_C2, C2.data(), C2.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUH(cudaMemcpy(h->d_Minv, Mt.data(), Mt.size() * sizeof(double), cudaMemcpyHostToDevice));
    d.C2 = h->d_C2; d.Minv_tiled = h->d_Minv;
    CUH(cudaMalloc(&h->d_xref, (size_t)B * 12 * (N + 1) * sizeof(double)));
    CUH(cudaMalloc(&h->d_fsteps, (size_t)B * 260 * sizeof(double)));

    h->aw = (20 * N + 31) / 32; h->cw = (4 * N + 31) / 32;
    // carve all per-instance arrays out of one block (doubles first: natural alignment)
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_f = take((size_t)B * 12 * N * 8), o_y = take((size_t)B * 20 * N * 8), o_xs = take((size_t)B * 12 * N * 8);
    const size_t o_f0 = take((size_t)B * 12 * 8), o_x1 = take((size_t)B * 12 * 8), o_qw = take((size_t)B * 6 * 8), o_obj = take((size_t)B * 8);
    const size_t o_status = take((size_t)B * 4), o_sweeps = take((size_t)B * 4), o_iters = take((size_t)B * 4);
    const size_t o_contact = take((size_t)B * h->cw * 4), o_active = take((size_t)B * h->aw * 4);
    const size_t o_list = take((size_t)B * 4), o_count = take(4 * (16 + 16 * MAX_RANGES)), o_sig = take((size_t)B * 4 * N), o_canary = take(256);
    h->block_bytes = off;
    CUH(cudaMalloc(&h->d_block, off));
    CUH(cudaMemset(h->d_block, 0, off));
    char* base = (char*)h->d_block;
    h->st.f = (double*)(base + o_f); h->st.y = (double*)(base + o_y); h->st.xs = (double*)(base + o_xs);
    h->st.f0 = (double*)(base + o_f0); h->st.x1 = (double*)(base + o_x1); h->st.qw = (double*)(base + o_qw); h->st.obj = (double*)(base + o_obj);
    {
        // MPC.py:55-58: every robot's world pose starts at [0, 0, h_ref, 0, 0, 0]
        std::vector<double> qw((size_t)B * 6, 0.0);
        for (int b = 0; b < B; ++b) qw[(size_t)b * 6 + 2] = 0.2027682;
        CUH(cudaMemcpy(h->st.qw, qw.data(), qw.size() * 8, cudaMemcpyHostToDevice));
    }
    h->st.status = (int32_t*)(base + o_status); h->st.sweeps = (int32_t*)(base + o_sweeps); h->st.iters = (int32_t*)(base + o_iters);
    h->st.contact = (uint32_t*)(base + o_contact); h->st.active = (uint32_t*)(base + o_active);
    h->st.fb_list = (int32_t*)(base + o_list); h->st.fb_count = (int32_t*)(base + o_count);
    h->ctr_base = h->st.fb_count; h->st.fb_next = nullptr;
    h->st.sig = (uint8_t*)(base + o_sig);
#ifdef MPCQP_CANARY
    h->st.canary = (unsigned int*)(base + o_canary);
#else
    h->st.canary = nullptr; (void)o_canary;
#endif
    CUH(cudaMemset(h->st.sig, SIG_FREE, (size_t)B * 4 * N));
    int ric_per_sm = 0, ipm_per_sm = 0;
    const int cap = ric_capacity(N);
    if (N == 16) CUH(configure_kernels<16>());
    else if (N == 32) CUH(configure_kernels<32>());
    if (cap == 16) CUH(ric_configure_16(&ric_per_sm, &ipm_per_sm));
    else if (cap == 32) CUH(ric_configure_32(&ric_per_sm, &ipm_per_sm));
    else CUH(ric_configure_64(&ric_per_sm, &ipm_per_sm));
    if (std::getenv("MPCQP_VERBOSE")) std::fprintf(stderr, "mpcqp: stage-wise kernels fit %d (active-set) / %d (interior-point) CTAs per SM\n", ric_per_sm, ipm_per_sm);
    if (const char* e = std::getenv("MPCQP_NO_DIRECT")) h->direct_ok = std::atoi(e) == 0;      // A/B hook
    if (const char* e = std::getenv("MPCQP_RIC_CTAS")) { const int c = std::atoi(e); if (c > 0 && c < ric_per_sm) ric_per_sm = c; }      // tuning hook
    if (p->mode & MPCQP_MODE_STAGEWISE) {
        if (ric_per_sm < 1 || ipm_per_sm < 1) return bail(fail(MPCQP_ERR_CUDA, "the stage-wise kernel does not fit on this device"));
        h->ric_max_ctas = ric_per_sm * h->sms;
        h->ipm_max_ctas = ipm_per_sm * h->sms;
        const int slots = h->ric_max_ctas > h->ipm_max_ctas ? h->ric_max_ctas : h->ipm_max_ctas;
        h->ric_ws_doubles = (size_t)slots * RIC_PER_CTA * ric_ws_slot_doubles(cap);
        CUH(cudaMalloc(&h->d_ric_ws, (1 + MAX_RANGES) * h->ric_ws_doubles * sizeof(double)));
        int lane_per_sm = 0;
        CUH(lane_configure(&lane_per_sm));
        if (lane_per_sm < 1) return bail(fail(MPCQP_ERR_CUDA, "the one-robot-per-lane kernel does not fit on this device"));
        h->lane_max_ctas = lane_per_sm * h->sms;
        // never chosen by batch size: as measured a lane per robot does not beat half a warp per robot at any batch
        h->lane_min = 1 << 30;
        if (const char* e = std::getenv("MPCQP_LANE_MIN")) { const int c = std::atoi(e); if (c > 0) h->lane_min = c; }      // tuning hook
    }
    if (!(p->mode & MPCQP_MODE_ACTIVE_SET)) {
        std::vector<int32_t> all((size_t)B + 1);
        for (int i = 0; i <= B; ++i) all[i] = i;
        CUH(cudaMalloc(&h->d_all, all.size() * 4));
        CUH(cudaMemcpy(h->d_all, all.data(), all.size() * 4, cudaMemcpyHostToDevice));
    }
#undef CUH
    *out = h;
    return MPCQP_OK;
}

// Overlapped index ranges (mpcqp_set_overlap).  fork: the side streams pick up everything issued on the main stream so far.
// join: the main stream waits for every range; every entry point that reads or changes the handle's state calls it first,
// so results, resets and host-input runs see whole ticks.  Between fork and join a range's tick t + 1 is ordered behind its
// own tick t only.
static int fork_ranges(mpcqp_handle* h, int R) {
    if (h->forked && h->forked_ranges == R) return MPCQP_OK;
    if (h->forked) {
        for (int r = 0; r < h->forked_ranges; ++r) {
            CU(cudaEventRecord(h->ev_side[r], h->side[r]));
            CU(cudaStreamWaitEvent(h->stream, h->ev_side[r], 0));
        }
    }
    CU(cudaEventRecord(h->ev_main, h->stream));
    for (int r = 0; r < R; ++r) CU(cudaStreamWaitEvent(h->side[r], h->ev_main, 0));
    h->forked = true;
    h->forked_ranges = R;
    return MPCQP_OK;
}
static int join_ranges(mpcqp_handle* h) {
    if (!h->forked) return MPCQP_OK;
    for (int r = 0; r < h->forked_ranges; ++r) {
        CU(cudaEventRecord(h->ev_side[r], h->side[r]));
        CU(cudaStreamWaitEvent(h->stream, h->ev_side[r], 0));
    }
    h->forked = false;
    return MPCQP_OK;
}
// Host -> device copy of `n` gait tables (20 x 13 doubles each).  A table ends at its first row whose step count is 0
// (MPC.py:646); the reference's tables use 2 .. 7 of the 20 rows, the rest is NaN padding.  If row 7 of every table in the
// batch has a zero count, only rows 0..7 travel (832 of 2080 bytes per robot, one strided copy): the device never reads
// past a table's terminator, so what an earlier tick left in rows 8..19 does not matter.  One read per robot decides.
static cudaError_t copy_gait_tables(double* dst, const double* src, size_t n, cudaStream_t s) {
    constexpr size_t ROW = 13, TABLE = 260, PROBE = 7;
    bool brief = true;
    for (size_t b = 0; b < n && brief; ++b) brief = src[b * TABLE + PROBE * ROW] == 0.0;
    if (!brief) return cudaMemcpyAsync(dst, src, n * TABLE * sizeof(double), cudaMemcpyHostToDevice, s);
    return cudaMemcpy2DAsync(dst, TABLE * sizeof(double), src, TABLE * sizeof(double), (PROBE + 1) * ROW * sizeof(double), n,
                             cudaMemcpyHostToDevice, s);
}

// robots off .. off + n - 1 of range r of R (even sizes: the two robots of a warp stay in one range)
static void range_of(const mpcqp_handle* h, int R, int r, int* off, int* n) {
    const int B = h->p.batch;
    const int per = ((B + R - 1) / R + 1) & ~1;
    *off = r * per;
    *n = B - *off < per ? B - *off : per;
    if (*n < 0) *n = 0;
}
// one tick of the whole batch as R independent index ranges.  hx / hf != null: host inputs; every range first copies its own rows
// into the handle's input buffers on its own stream (the copy engine moves them at the full bus rate while other ranges solve;
// the range's previous tick, which read those rows, is ahead of the copy in the same stream).
static int run_ranges(mpcqp_handle* h, int R, const double* dx, const double* df, int first, bool closed_loop,
                      const double* hx = nullptr, const double* hf = nullptr) {
    int rc = fork_ranges(h, R);
    if (rc) return rc;
    const size_t xs = (size_t)12 * (h->p.n_steps + 1), fs = 260;
    for (int r = 0; r < R; ++r) {
        int off, n;
        range_of(h, R, r, &off, &n);
        if (n <= 0) continue;
        if (hx) {
            CU(cudaMemcpyAsync(h->d_xref + off * xs, hx + off * xs, n * xs * sizeof(double), cudaMemcpyHostToDevice, h->side[r]));
            CU(copy_gait_tables(h->d_fsteps + off * fs, hf + off * fs, n, h->side[r]));
        }
        h->solve_range(r, off, n, dx, df, first, closed_loop);
    }
    h->last_ranges = R;
    return MPCQP_OK;
}

// Start of a tick: switch to the other copy of the counters (fallback queue length + the stage-wise kernel's work counters).  It is
// already zero when the previous tick ran a stage-wise launch on the main stream (that kernel clears the copy it does not use);
// otherwise one 16-byte memset.
static cudaError_t begin_tick(mpcqp_handle* h) {
    h->ctr_parity ^= 1;
    h->st.fb_count = h->ctr_base + 8 * h->ctr_parity;
    h->st.fb_next = nullptr;
    const bool clean = h->ctr_clean && h->zero_next;
    h->ctr_clean = true; h->zero_next = false;
    return clean ? cudaSuccess : cudaMemsetAsync(h->st.fb_count, 0, 4 * sizeof(int32_t), h->stream);
}

static int stage_inputs(mpcqp_handle* h, const double* xref, const double* fsteps, int location,
                        const double** dx, const double** df) {
    const int N = h->p.n_steps, B = h->p.batch;
    if (location == MPCQP_DEVICE) {
        *dx = xref; *df = fsteps;
    } else if (location == MPCQP_HOST) {
        CU(cudaMemcpyAsync(h->d_xref, xref, (size_t)B * 12 * (N + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->d_fsteps, fsteps, (size_t)B * 260 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        *dx = h->d_xref; *df = h->d_fsteps;
    } else {
        return fail(MPCQP_ERR_INVALID, "location must be MPCQP_HOST or MPCQP_DEVICE");
    }
    return MPCQP_OK;
}

int mpcqp_run(mpcqp_handle* h, double k, const double* xref, const double* fsteps, int location) {
    if (!h || !xref || !fsteps) return fail(MPCQP_ERR_INVALID, "null argument");
    if (location != MPCQP_HOST && location != MPCQP_DEVICE) return fail(MPCQP_ERR_INVALID, "location must be MPCQP_HOST or MPCQP_DEVICE");
    CU(cudaSetDevice(h->p.device));
    const int B = h->p.batch, N = h->p.n_steps;
    const int first = (k == 0.0) ? 1 : 0;                       // MPC.py:491, 413: only k == 0 vs k > 0 matters
    const size_t xs = (size_t)12 * (N + 1), fs = 260;
    const bool stageA = (h->p.mode & MPCQP_MODE_ACTIVE_SET) != 0;
    const double *dx = xref, *df = fsteps;
    // Host inputs in page-locked memory (mpcqp_host_alloc, cudaHostAlloc, torch pin_memory): the stage-wise kernels fetch every
    // robot's xref and gait table themselves, with the same bulk asynchronous copies they use on HBM, straight from the caller's
    // buffers over PCIe -- no staging copy in front of the solve: a robot's inputs travel while other robots factorise, and of a gait
    // table only rows 0..7 cross the bus unless it is longer.  The buffers must stay unchanged until the tick's results have been
    // fetched (the same contract as for the asynchronous staging copies).  Pageable memory takes the staged path below.
    bool direct = false;
    if (location == MPCQP_HOST && stageA && (h->p.mode & MPCQP_MODE_STAGEWISE) && (!h->has_fallback() || h->fallback_is_ipm()) &&
        !h->use_lane() && h->direct_ok) {
        cudaPointerAttributes ax, af;
        if (cudaPointerGetAttributes(&ax, xref) == cudaSuccess && cudaPointerGetAttributes(&af, fsteps) == cudaSuccess &&
            ax.type == cudaMemoryTypeHost && af.type == cudaMemoryTypeHost && ax.devicePointer && af.devicePointer) {
            direct = true;
            dx = (const double*)ax.devicePointer; df = (const double*)af.devicePointer;
        } else cudaGetLastError();
    }
    if (stageA) {
        // overlap switched on: the tick is issued as independent index ranges, no join.  Host inputs are staged range by range (the
        // copy engine reaches the full bus rate, 53 GB/s, where robot-by-robot reads by the kernels reach ~36: measured)
        const int R = h->ranges_for(false);
        if (R > 1) {
            const bool host = location == MPCQP_HOST;
            const int rc = host ? run_ranges(h, R, h->d_xref, h->d_fsteps, first, false, xref, fsteps) : run_ranges(h, R, dx, df, first, false);
            if (rc) return rc;
            CU(h->launch_err);
            CU(cudaGetLastError());
            h->ran = true;
            return MPCQP_OK;
        }
    }
    { const int rc = join_ranges(h); if (rc) return rc; }
    h->last_ranges = 1;
    CU(begin_tick(h));
    // Host inputs: the batch is cut into chunks of two full waves (2 x 4 CTAs x #SM instances); chunk
    // c is copied and solved on side stream c & 1, so the H2D copy of one chunk overlaps the solve of
    // the previous one and the two solve kernels fill each other's tails.
    int chunk = (h->p.mode & MPCQP_MODE_STAGEWISE) ? h->wave() : 2 * h->wave();
    if (h->use_lane()) chunk = 2 * h->lane_max_ctas * 32;
    if (const char* e = std::getenv("MPCQP_CHUNK")) { const int c = std::atoi(e); if (c > 0) chunk = c; }      // tuning hook
    if (direct) {
        h->dp.fs_rows = 8;
        h->solve(false, B, h->stream, dx, df, first, 0, B);
        if (h->has_fallback()) h->solve(true, B, h->stream, dx, df, first, 0, B);
        h->dp.fs_rows = 20;
        CU(h->launch_err);
        CU(cudaGetLastError());
        h->ran = true;
        return MPCQP_OK;
    }
    if (location == MPCQP_HOST) {
        dx = h->d_xref; df = h->d_fsteps;
        if (stageA && B > chunk) {
            CU(cudaEventRecord(h->ev_main, h->stream));          // everything issued so far (incl. the last tick)
            for (int i = 0; i < 2; ++i) CU(cudaStreamWaitEvent(h->side[i], h->ev_main, 0));
            int c = 0;
            for (int off = 0; off < B; off += chunk, ++c) {
                const int n = B - off < chunk ? B - off : chunk;
                cudaStream_t s = h->side[c & 1];
                CU(cudaMemcpyAsync(h->d_xref + off * xs, xref + off * xs, n * xs * sizeof(double), cudaMemcpyHostToDevice, s));
                CU(copy_gait_tables(h->d_fsteps + off * fs, fsteps + off * fs, n, s));
                h->solve(false, n, s, dx, df, first, off, n);
            }
            for (int i = 0; i < 2; ++i) {
                CU(cudaEventRecord(h->ev_side[i], h->side[i]));
                CU(cudaStreamWaitEvent(h->stream, h->ev_side[i], 0));
            }
        } else {
            CU(cudaMemcpyAsync(h->d_xref, xref, B * xs * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            CU(copy_gait_tables(h->d_fsteps, fsteps, B, h->stream));
            if (stageA) h->solve(false, B, h->stream, dx, df, first, 0, B);
        }
    } else if (stageA) {
        h->solve(false, B, h->stream, dx, df, first, 0, B);
    }
    if (!stageA) {
        // ADMM only: queue every instance (the identity list and its length were uploaded once, at creation)
        CU(cudaMemcpyAsync(h->st.fb_list, h->d_all, (size_t)B * 4, cudaMemcpyDeviceToDevice, h->stream));
        CU(cudaMemcpyAsync(h->st.fb_count, h->d_all + B, 4, cudaMemcpyDeviceToDevice, h->stream));
    }
    if (h->has_fallback()) {
        const int slots = h->ctas_per_sm(true) * h->sms;
        h->solve(true, B < slots ? B : slots, h->stream, dx, df, first, 0, B);
    }
    CU(h->launch_err);
    CU(cudaGetLastError());
    h->ran = true;
    return MPCQP_OK;
}

static int fetch(mpcqp_handle* h, void* dst, const void* src, size_t bytes, int location) {
    if (!dst) return MPCQP_OK;
    if (location == MPCQP_HOST) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    } else if (location == MPCQP_DEVICE) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, h->stream));
    } else {
        return fail(MPCQP_ERR_INVALID, "location must be MPCQP_HOST or MPCQP_DEVICE");
    }
    return MPCQP_OK;
}

int mpcqp_get_latest_result(mpcqp_handle* h, double* forces, int location) {
    if (!h || !forces) return fail(MPCQP_ERR_INVALID, "null argument");
    if (!h->ran) return fail(MPCQP_ERR_STATE, "no run has been issued on this handle");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    int rc = fetch(h, forces, h->st.f0, (size_t)h->p.batch * 12 * sizeof(double), location);
    if (rc) return rc;
    if (location == MPCQP_HOST) CU(cudaStreamSynchronize(h->stream));
    return MPCQP_OK;
}

// MPC.f_applied and the first predicted state in one call (one synchronisation): what a control loop needs per tick
// (MPC_Wrapper.py:114 reads f_applied; MPC.py:448-450, 503-510 use q_next / v_next).  `next_state` is x_robot[:, 0] (12 per robot).
int mpcqp_get_step_result(mpcqp_handle* h, double* forces, double* next_state, int location) {
    double* dev1 = next_state;
    if (!h || (!forces && !dev1)) return fail(MPCQP_ERR_INVALID, "null argument");
    if (!h->ran) return fail(MPCQP_ERR_STATE, "no run has been issued on this handle");
    if (location != MPCQP_HOST && location != MPCQP_DEVICE) return fail(MPCQP_ERR_INVALID, "bad location");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    const size_t B = h->p.batch, row = 12 * sizeof(double);
    const cudaMemcpyKind kind = location == MPCQP_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (forces) CU(cudaMemcpyAsync(forces, h->st.f0, B * row, kind, h->stream));
    if (dev1) CU(cudaMemcpyAsync(dev1, h->st.x1, B * row, kind, h->stream));
    if (location == MPCQP_HOST) CU(cudaStreamSynchronize(h->stream));
    return MPCQP_OK;
}

// MPC.q_w (MPC.py:58, 503-510): the dead-reckoned world pose of every robot, B x 6; `set` != 0 writes it instead
int mpcqp_world_pose(mpcqp_handle* h, double* qw, int set, int location) {
    if (!h || !qw) return fail(MPCQP_ERR_INVALID, "null argument");
    if (location != MPCQP_HOST && location != MPCQP_DEVICE) return fail(MPCQP_ERR_INVALID, "bad location");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    const size_t bytes = (size_t)h->p.batch * 6 * sizeof(double);
    if (set) CU(cudaMemcpyAsync(h->st.qw, qw, bytes, location == MPCQP_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
    else CU(cudaMemcpyAsync(qw, h->st.qw, bytes, location == MPCQP_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, h->stream));
    if (location == MPCQP_HOST) CU(cudaStreamSynchronize(h->stream));
    return MPCQP_OK;
}

// per-instance status only (4 B per robot)
int mpcqp_get_status(mpcqp_handle* h, int32_t* status, int location) {
    if (!h || !status) return fail(MPCQP_ERR_INVALID, "null argument");
    if (!h->ran) return fail(MPCQP_ERR_STATE, "no run has been issued on this handle");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    int rc = fetch(h, status, h->st.status, (size_t)h->p.batch * 4, location);
    if (rc) return rc;
    if (location == MPCQP_HOST) CU(cudaStreamSynchronize(h->stream));
    return MPCQP_OK;
}

int mpcqp_get_solution(mpcqp_handle* h, double* x, int location) {
    if (!h || !x) return fail(MPCQP_ERR_INVALID, "null argument");
    if (!h->ran) return fail(MPCQP_ERR_STATE, "no run has been issued on this handle");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    const int N = h->p.n_steps, B = h->p.batch;
    const cudaMemcpyKind kind = location == MPCQP_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (location != MPCQP_HOST && location != MPCQP_DEVICE) return fail(MPCQP_ERR_INVALID, "bad location");
    const size_t half = (size_t)12 * N * sizeof(double);
    CU(cudaMemcpy2DAsync(x, 2 * half, h->st.xs, half, half, B, kind, h->stream));
    CU(cudaMemcpy2DAsync((char*)x + half, 2 * half, h->st.f, half, half, B, kind, h->stream));
    if (location == MPCQP_HOST) CU(cudaStreamSynchronize(h->stream));
    return MPCQP_OK;
}

int mpcqp_get_info(mpcqp_handle* h, int32_t* status, int32_t* sweeps, int32_t* iters, double* obj,
                   uint32_t* contact, uint32_t* active, double* y, int location) {
    if (!h) return fail(MPCQP_ERR_INVALID, "null argument");
    if (!h->ran) return fail(MPCQP_ERR_STATE, "no run has been issued on this handle");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    const size_t B = h->p.batch;
    int rc;
    if ((rc = fetch(h, status, h->st.status, B * 4, location))) return rc;
    if ((rc = fetch(h, sweeps, h->st.sweeps, B * 4, location))) return rc;
    if ((rc = fetch(h, iters, h->st.iters, B * 4, location))) return rc;
    if ((rc = fetch(h, obj, h->st.obj, B * 8, location))) return rc;
    if ((rc = fetch(h, contact, h->st.contact, B * h->cw * 4, location))) return rc;
    if ((rc = fetch(h, active, h->st.active, B * h->aw * 4, location))) return rc;
    if ((rc = fetch(h, y, h->st.y, B * 20 * h->p.n_steps * 8, location))) return rc;
    if (location == MPCQP_HOST) CU(cudaStreamSynchronize(h->stream));
    return MPCQP_OK;
}

int mpcqp_get_fallback_count(mpcqp_handle* h, int32_t* count) {
    if (!h || !count) return fail(MPCQP_ERR_INVALID, "null argument");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    if (h->last_ranges > 1) {
        // the last tick ran as index ranges: each has its own queue
        int32_t part[MAX_RANGES] = {};
        for (int r = 0; r < h->last_ranges; ++r)
            CU(cudaMemcpyAsync(&part[r], h->ctr_base + 16 + 16 * r + 8 * h->rparity[r], 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        *count = 0;
        for (int r = 0; r < MAX_RANGES; ++r) *count += part[r];
        return MPCQP_OK;
    }
    CU(cudaMemcpyAsync(count, h->st.fb_count, 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MPCQP_OK;
}

int mpcqp_reset_warm_start(mpcqp_handle* h) {
    if (!h) return fail(MPCQP_ERR_INVALID, "null argument");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    const size_t B = h->p.batch, N = h->p.n_steps;
    CU(cudaMemsetAsync(h->st.f, 0, B * 12 * N * 8, h->stream));
    CU(cudaMemsetAsync(h->st.y, 0, B * 20 * N * 8, h->stream));
    CU(cudaMemsetAsync(h->st.sig, SIG_FREE, B * 4 * N, h->stream));
    return MPCQP_OK;
}

int mpcqp_synchronize(mpcqp_handle* h) {
    if (!h) return fail(MPCQP_ERR_INVALID, "null argument");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    CU(cudaStreamSynchronize(h->stream));
    return MPCQP_OK;
}

void* mpcqp_stream(mpcqp_handle* h) { return h ? (void*)h->stream : nullptr; }

// Overlap of consecutive ticks (see include/mpcqp.h)
int mpcqp_set_overlap(mpcqp_handle* h, int ranges) {
    if (!h || ranges < 0 || ranges > MAX_RANGES) return fail(MPCQP_ERR_INVALID, "ranges must be 0 (automatic) .. 8");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    h->ranges = ranges;
    return MPCQP_OK;
}
int mpcqp_join(mpcqp_handle* h) {
    if (!h) return fail(MPCQP_ERR_INVALID, "null argument");
    CU(cudaSetDevice(h->p.device));
    return join_ranges(h);
}

// page-locked host memory for the caller's input / output arrays (copies from pageable memory are staged by the driver
// and cost ~20 % of a 4096-robot tick)
void* mpcqp_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        fail(MPCQP_ERR_CUDA, "cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}
int mpcqp_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return MPCQP_OK;
}
int64_t mpcqp_launch_count(mpcqp_handle* h) { return h ? h->launches : 0; }

int mpcqp_export_build(mpcqp_handle* h, double k, const double* xref, const double* fsteps, int location,
                       double* B_vals, double* S_vals, double* NK) {
    if (!h || !xref || !fsteps || !B_vals || !S_vals || !NK) return fail(MPCQP_ERR_INVALID, "null argument");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    const double *dx, *df;
    int rc = stage_inputs(h, xref, fsteps, location, &dx, &df);
    if (rc) return rc;
    const size_t B = h->p.batch, N = h->p.n_steps;
    double *dB, *dS, *dN;
    const bool host = location == MPCQP_HOST;
    if (host) {
        CU(h->scratch(B * 72 * N * 8));
        dB = h->d_scratch; dS = dB + B * 48 * N; dN = dS + B * 12 * N;
    } else {
        dB = B_vals; dS = S_vals; dN = NK;
    }
    export_build_kernel<<<(int)B, N > 32 ? 256 : 128, 0, h->stream>>>(h->dp, dx, df, k == 0.0 ? 1 : 0, dB, dS, dN);
    ++h->launches;
    CU(cudaGetLastError());
    if (host) {
        CU(cudaMemcpyAsync(B_vals, dB, B * 48 * N * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(S_vals, dS, B * 12 * N * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(NK, dN, B * 12 * N * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return MPCQP_OK;
}

int mpcqp_measure_fp64_peak(int device, double* dfma_tflops, double* dmma_tflops) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MPCQP_ERR_NO_DEVICE, "no CUDA device visible");
    }
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount, threads = 512, iters = 20000;
    double* out;
    CU(cudaMalloc(&out, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    double best[2] = {0, 0};
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 4; ++rep) {
            CU(cudaEventRecord(e0));
            if (which == 0) peak_dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
            else peak_dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
            CU(cudaEventRecord(e1));
            CU(cudaEventSynchronize(e1));
            float ms;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            const double flop = which == 0 ? 2.0 * 8 * iters * (double)threads * blocks
                                           : 2.0 * 256 * 4 * iters * (double)(threads / 32) * blocks;
            const double tf = flop / (ms * 1e-3) * 1e-12;
            if (rep > 0 && tf > best[which]) best[which] = tf;
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    if (dfma_tflops) *dfma_tflops = best[0];
    if (dmma_tflops) *dmma_tflops = best[1];
    return MPCQP_OK;
}

// ------------------------------------------------------------------------------------------------
// device-resident closed loop (SURVEY 8f rows f1 + f2)
// ------------------------------------------------------------------------------------------------
int mpcqp_scenario_init(mpcqp_handle* h, const uint64_t* seq, const int32_t* phase, const double* vref, const double* state,
                        const double* sigma4, uint64_t seed) {
    if (!h || !seq || !phase || !vref || !state || !sigma4) return fail(MPCQP_ERR_INVALID, "null argument");
    const size_t B = h->p.batch;
    const int N = h->p.n_steps;
    // the gait period is T_gait / dt steps (FootstepPlanner.py:52-63); `seq` holds 4 bits per step, 16 steps per 64-bit word
    const int period = (int)std::lround(h->p.T_gait / h->p.dt);
    if (period < 1 || period > 64 || std::fabs(h->p.T_gait / h->p.dt - period) > 1e-9)
        return fail(MPCQP_ERR_INVALID, "closed loop needs a gait period T_gait / dt of 1 .. 64 whole steps");
    const size_t seq_words = (size_t)(period + 15) / 16;
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    CU(cudaStreamSynchronize(h->stream));
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_state = take(B * 12 * 8), o_frame = take(B * 3 * 8), o_feet = take(B * 8 * 8), o_target = take(B * 8 * 8);
    const size_t o_vref = take(B * 6 * 8), o_seq = take(B * seq_words * 8), o_phase = take(B * 4), o_prev = take(B);
    const size_t o_cmd = take(B * 3 * 8), o_cflag = take(B), o_ctick = take(B * 4);
    if (!h->d_scen) CU(cudaMalloc(&h->d_scen, off));
    CU(cudaMemset(h->d_scen, 0, off));
    char* base = (char*)h->d_scen;
    DevScenario& s = h->sc;
    std::memset(&s, 0, sizeof(s));
    s.state = (double*)(base + o_state); s.frame = (double*)(base + o_frame); s.feet = (double*)(base + o_feet);
    s.target = (double*)(base + o_target); s.vref = (const double*)(base + o_vref);
    s.seq = (const unsigned long long*)(base + o_seq); s.phase = (const int32_t*)(base + o_phase); s.prevc = (uint8_t*)(base + o_prev);
    s.cmd = (double*)(base + o_cmd); s.cmd_flag = (uint8_t*)(base + o_cflag); s.cmd_tick = (int32_t*)(base + o_ctick);
    {
        std::vector<double> cmd(B * 3);
        for (size_t b = 0; b < B; ++b) { cmd[b * 3] = SC_H_ROTATION0; cmd[b * 3 + 1] = SC_H_REF; cmd[b * 3 + 2] = 0.0; }
        CU(cudaMemcpy(base + o_cmd, cmd.data(), B * 3 * 8, cudaMemcpyHostToDevice));
        CU(cudaMemset(base + o_ctick, 0xFF, B * 4));           // -1: no tick applied yet
    }
    std::vector<double> feet(B * 8);
    for (size_t b = 0; b < B; ++b)
        for (int j = 0; j < 4; ++j) { feet[b * 8 + j] = sc_shoulder_x(j); feet[b * 8 + 4 + j] = sc_shoulder_y(j); }
    CU(cudaMemcpy(base + o_state, state, B * 12 * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(base + o_feet, feet.data(), B * 8 * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(base + o_target, feet.data(), B * 8 * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(base + o_vref, vref, B * 6 * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(base + o_seq, seq, B * seq_words * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(base + o_phase, phase, B * 4, cudaMemcpyHostToDevice));
    for (int i = 0; i < 4; ++i) s.sigma[i] = sigma4[i];
    // numpy.linspace(start, stop, N): start + i * step with the last element set to stop exactly
    const double a0 = 0.0, a1 = h->p.T_gait - h->p.dt, b0 = h->p.dt, b1 = h->p.T_gait;
    for (int i = 0; i < N; ++i) {
        s.lin_a[i] = (N == 1) ? a0 : ((i == N - 1) ? a1 : a0 + i * ((a1 - a0) / (N - 1)));
        s.lin_b[i] = (N == 1) ? b0 : ((i == N - 1) ? b1 : b0 + i * ((b1 - b0) / (N - 1)));
    }
    s.seed = seed;
    s.period = period;
    s.seq_words = (int)seq_words;
    h->scen_tick = 0;
    h->scen_ready = true;
    CU(mpcqp_reset_warm_start(h) == 0 ? cudaSuccess : cudaErrorUnknown);
    return MPCQP_OK;
}

// New joystick commands for the following ticks (Joystick.update_v_ref, Joystick.py:29-43): vref B x 6 (HOST; may be NULL to
// keep the current ones) and the `reduced` support-polygon switch (Joystick.py:66-67, FootstepPlanner.py:330-332).
int mpcqp_scenario_set_commands(mpcqp_handle* h, const double* vref, int reduced) {
    if (!h) return fail(MPCQP_ERR_INVALID, "null argument");
    if (!h->scen_ready) return fail(MPCQP_ERR_STATE, "mpcqp_scenario_init has not been called");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    if (vref) {
        CU(cudaMemcpyAsync((void*)h->sc.vref, vref, (size_t)h->p.batch * 6 * 8, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));                  // the caller's buffer is free on return
    }
    h->sc.reduced = reduced ? 1 : 0;
    return MPCQP_OK;
}

int mpcqp_scenario_run(mpcqp_handle* h, int ticks, int emit_inputs) {
    if (!h || ticks < 0) return fail(MPCQP_ERR_INVALID, "bad argument");
    if (!h->scen_ready) return fail(MPCQP_ERR_STATE, "mpcqp_scenario_init has not been called");
    CU(cudaSetDevice(h->p.device));
    const int B = h->p.batch;
    const bool stageA = (h->p.mode & MPCQP_MODE_ACTIVE_SET) != 0;
    if (!stageA) return fail(MPCQP_ERR_INVALID, "the closed loop needs the active-set stage enabled");
    h->sc.xref_out = emit_inputs ? h->d_xref : nullptr;
    h->sc.fsteps_out = emit_inputs ? h->d_fsteps : nullptr;
    // every robot's tick depends on its own previous tick only (planner, solve and integration are per robot): with two or more
    // ticks in one call the batch advances as independent index ranges, joined before the call returns
    const int R = h->ranges_for(ticks >= 2);
    if (R > 1 && ticks >= 1) {
        for (int t = 0; t < ticks; ++t) {
            h->sc.tick = h->scen_tick;
            const int rc = run_ranges(h, R, nullptr, nullptr, h->scen_tick == 0 ? 1 : 0, true);
            if (rc) return rc;
            ++h->scen_tick;
        }
        if (h->ranges == 0) { const int rc = join_ranges(h); if (rc) return rc; }     // automatic: whole ticks at the call boundary
        CU(h->launch_err);
        CU(cudaGetLastError());
        h->ran = true;
        return MPCQP_OK;
    }
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    if (ticks > 0) h->last_ranges = 1;
    for (int t = 0; t < ticks; ++t) {
        h->sc.tick = h->scen_tick;
        const int first = h->scen_tick == 0 ? 1 : 0;
        CU(begin_tick(h));
        h->solve(false, B, h->stream, nullptr, nullptr, first, 0, B, true);
        if (h->has_fallback()) {
            const int slots = h->ctas_per_sm(true) * h->sms;
            h->solve(true, B < slots ? B : slots, h->stream, nullptr, nullptr, first, 0, B, true);
        }
        ++h->scen_tick;
    }
    CU(cudaGetLastError());
    h->ran = true;
    return MPCQP_OK;
}

int mpcqp_scenario_get(mpcqp_handle* h, double* state, double* frame, double* feet) {
    if (!h) return fail(MPCQP_ERR_INVALID, "null argument");
    if (!h->scen_ready) return fail(MPCQP_ERR_STATE, "mpcqp_scenario_init has not been called");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    const size_t B = h->p.batch;
    if (state) CU(cudaMemcpyAsync(state, h->sc.state, B * 12 * 8, cudaMemcpyDeviceToHost, h->stream));
    if (frame) CU(cudaMemcpyAsync(frame, h->sc.frame, B * 3 * 8, cudaMemcpyDeviceToHost, h->stream));
    if (feet) CU(cudaMemcpyAsync(feet, h->sc.feet, B * 8 * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MPCQP_OK;
}

int mpcqp_get_inputs(mpcqp_handle* h, double* xref, double* fsteps) {
    if (!h || !xref || !fsteps) return fail(MPCQP_ERR_INVALID, "null argument");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    const size_t B = h->p.batch, N = h->p.n_steps;
    CU(cudaMemcpyAsync(xref, h->d_xref, B * 12 * (N + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(fsteps, h->d_fsteps, B * 260 * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MPCQP_OK;
}

// Logger.log_cost_function (Logger.py:406-418) for every robot of the last run: cost is B x 13 (12 state components, forces)
int mpcqp_get_cost_components(mpcqp_handle* h, double* cost, int location) {
    if (!h || !cost) return fail(MPCQP_ERR_INVALID, "null argument");
    if (!h->ran) return fail(MPCQP_ERR_STATE, "no run has been issued on this handle");
    if (location != MPCQP_HOST && location != MPCQP_DEVICE) return fail(MPCQP_ERR_INVALID, "location must be MPCQP_HOST or MPCQP_DEVICE");
    CU(cudaSetDevice(h->p.device));
    { const int rc_ = join_ranges(h); if (rc_) return rc_; }
    const size_t B = h->p.batch, bytes = B * 13 * sizeof(double);
    double* d = cost;
    if (location == MPCQP_HOST) { CU(h->scratch(bytes)); d = h->d_scratch; }
    cost_components_kernel<<<(int)((B + 3) / 4), 128, 0, h->stream>>>(h->dp, h->st, h->p.n_steps, d);
    ++h->launches;
    CU(cudaGetLastError());
    if (location == MPCQP_HOST) {
        CU(cudaMemcpyAsync(cost, d, bytes, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return MPCQP_OK;
}

// ---- asynchronous result protocol (the reference's intent in MPC_Wrapper.py:116-260: the control loop
// picks up the forces of the PREVIOUS solve while the current one runs).  mpcqp_result_async enqueues the copy of the forces
// of the run just issued into pinned slot 0 / 1 and returns at once; mpcqp_result_wait blocks on that slot only.
int mpcqp_result_async(mpcqp_handle* h, int slot) {
    if (!h || slot < 0 || slot > 1) return fail(MPCQP_ERR_INVALID, "bad argument");
    if (!h->ran) return fail(MPCQP_ERR_STATE, "no run has been issued on this handle");
    CU(cudaSetDevice(h->p.device));
    const size_t bytes = (size_t)h->p.batch * 12 * sizeof(double);
    if (!h->pin[slot]) {
        CU(cudaMallocHost(&h->pin[slot], bytes));
        CU(cudaEventCreateWithFlags(&h->ev_pin[slot], cudaEventDisableTiming));
    }
    if (h->forked) {
        // overlapped ticks: every index range copies its own forces on its own stream, behind its tick and in front of its next one
        // -- no range waits for another, and a range's next tick cannot overwrite what has not been copied yet
        const int R = h->forked_ranges;
        for (int r = 0; r < R; ++r) {
            int off, n;
            range_of(h, R, r, &off, &n);
            if (!h->ev_pin_r[slot][r]) CU(cudaEventCreateWithFlags(&h->ev_pin_r[slot][r], cudaEventDisableTiming));
            if (n > 0) CU(cudaMemcpyAsync(h->pin[slot] + (size_t)off * 12, h->st.f0 + (size_t)off * 12, (size_t)n * 12 * sizeof(double),
                                          cudaMemcpyDeviceToHost, h->side[r]));
            CU(cudaEventRecord(h->ev_pin_r[slot][r], h->side[r]));
        }
        h->pin_ranges[slot] = R;
        h->pin_valid[slot] = true;
        return MPCQP_OK;
    }
    CU(cudaMemcpyAsync(h->pin[slot], h->st.f0, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->ev_pin[slot], h->stream));
    h->pin_ranges[slot] = 0;
    h->pin_valid[slot] = true;
    return MPCQP_OK;
}

// 1 if the copy requested in `slot` has landed, 0 if it is still in flight (never blocks)
int mpcqp_result_ready(mpcqp_handle* h, int slot) {
    if (!h || slot < 0 || slot > 1) return fail(MPCQP_ERR_INVALID, "bad argument");
    if (!h->pin_valid[slot]) return 0;
    cudaSetDevice(h->p.device);
    const int R = h->pin_ranges[slot];
    for (int r = 0; r < (R > 0 ? R : 1); ++r) {
        const cudaError_t e = cudaEventQuery(R > 0 ? h->ev_pin_r[slot][r] : h->ev_pin[slot]);
        if (e == cudaSuccess) continue;
        if (e == cudaErrorNotReady) { cudaGetLastError(); return 0; }
        return fail(MPCQP_ERR_CUDA, std::string("cudaEventQuery: ") + cudaGetErrorString(e));
    }
    return 1;
}

int mpcqp_result_wait(mpcqp_handle* h, int slot, double* forces) {
    if (!h || !forces || slot < 0 || slot > 1) return fail(MPCQP_ERR_INVALID, "bad argument");
    if (!h->pin_valid[slot]) return fail(MPCQP_ERR_STATE, "no result was requested in this slot");
    CU(cudaSetDevice(h->p.device));
    if (h->pin_ranges[slot] > 0) {
        for (int r = 0; r < h->pin_ranges[slot]; ++r) CU(cudaEventSynchronize(h->ev_pin_r[slot][r]));
    } else CU(cudaEventSynchronize(h->ev_pin[slot]));
    std::memcpy(forces, h->pin[slot], (size_t)h->p.batch * 12 * sizeof(double));
    return MPCQP_OK;
}

#ifdef MPCQP_CANARY
// debug build only (`make canary`): read (and clear) the guard-word / invariant violation counters of the stage-wise kernels
int mpcqp_debug_canary(mpcqp_handle* h, unsigned int* out8) {
    if (!h || !out8) return fail(MPCQP_ERR_INVALID, "null argument");
    CU(cudaSetDevice(h->p.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out8, h->st.canary, sizeof(unsigned int) * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemset(h->st.canary, 0, sizeof(unsigned int) * 8));
    return MPCQP_OK;
}
#endif

#ifdef MPCQP_PROFILE
// debug builds only: read (and clear) the per-phase cycle counters
int mpcqp_debug_profile(unsigned long long* out64) {
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out64, g_prof, sizeof(unsigned long long) * 64));
    unsigned long long z[64] = {0};
    CU(cudaMemcpyToSymbol(g_prof, z, sizeof(z)));
    return MPCQP_OK;
}
#endif

}  // extern "C"
