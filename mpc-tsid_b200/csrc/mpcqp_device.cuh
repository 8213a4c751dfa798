// mpcqp_device.cuh -- device-side building blocks of the batched centroidal-MPC QP engine (sm_100a).
//
// One CTA solves one QP instance (one robot, one MPC tick).  Notation:
//   N            horizon (MPC.py:42), feet j = 0..3 = FL, FR, HL, HR, steps k = 0..N-1
//   f_kj         3-vector contact force of foot j at step k            (x[12N + 12k + 3j ..] in MPC.py)
//   Bv_kj        6x3 map force -> velocity impulse: rows 0..2 = (dt/m) I, rows 3..5 = dt inv(R gI) [r]x
//                (== rows 6..11 of the reference's B block, MPC.py:119, 339-346)
//   u_k          6-vector impulse at step k = sum_j Bv_kj f_kj
//   M            6N x 6N Gram matrix of the double-integrator response, M_c[k,l] = dt^2 Qp_c C2 + Qv_c C0
//   H            = w_f I + Bv' M Bv      condensed Hessian (never formed)
//   W            = M^-1 + sum_j (Bv Z) D^-1 (Bv Z)'   6N x 6N SPD, the only matrix ever factorised
// The tile layout of W in shared memory is the DMMA operand layout: 8x8 tiles, each stored as two
// 8x4 panels (element (r, c) at (c >> 2) * 32 + r * 4 + (c & 3)), lower block triangle only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// tile helpers out of line (small code, call overhead) or inlined (large code, no calls)
#ifdef MPCQP_OUTLINE_TILES
#define MPCQP_TILE_FN __noinline__
#else
#define MPCQP_TILE_FN __forceinline__
#endif
// optional per-phase cycle counters of the factorisation (tools/piece_bench.cu)
#ifdef MPCQP_TILE_PROF
__device__ unsigned long long g_tile_prof[8];
#define TPROF(slot) do { if ((threadIdx.x & 31) == 0) { long long n_ = clock64(); atomicAdd(&g_tile_prof[(slot) + 8 * 0], (unsigned long long)(n_ - tp_)); tp_ = n_; } } while (0)
#define TPROF_T0() long long tp_ = clock64()
#else
#define TPROF(slot) do {} while (0)
#define TPROF_T0() do {} while (0)
#endif

namespace mpcqp {

struct DevParams {
    int N, batch;
    double dt, mass, mu, fz_max, gravity, w_force;
    double gIinv[9];        // inverse body inertia, row major (host, extended precision)
    double footholds[12];   // 3 x 4 row major
    double wp[6], wv[6];    // state weights: position-like (x y z roll pitch yaw) and velocity-like
    double rho, sigma, alpha, feas_tol, dual_tol;
    int max_sweeps, max_iter, min_iter, check_every, warm_start, mode, refine, ipm_max_iter;
    int fs_rows;            // stage-wise kernels: rows of a gait table fetched up front (20 = the whole table; 8 = rows 0..7, the rest only for a
                            // table that does not end within them: inputs read straight from the caller's page-locked host memory)
    double inv_wf;          // 1 / w_force
    double dz_tab[3];       // 1 / (w_force (1 + mu^2 m)), m = sx^2 + sy^2 = 0, 1, 2: the face reciprocals (make_face), IEEE divisions done once on the host
    const double* Minv_tiled;   // lower block triangle of M^-1 in the smem tile layout
    const double* C2;           // N x N: C2[k,l] = sum_{i >= max(k,l)} (i-k)(i-l)   (C0[k,l] = N - max(k,l))
};

// Carried per-instance state and outputs (device pointers, leading dimension = instance)
struct DevState {
    double* f;          // B x N*12   forces of the last solve (also the force half of MPC.x)
    double* y;          // B x N*20   multipliers of the pyramid rows
    uint8_t* sig;       // B x N*4    active-set signature per foot-step
    double* xs;         // B x N*12   state half of MPC.x (X - xref)
    double* f0;         // B x 12     f_applied
    double* x1;         // B x 12     X_1 = x_robot[:, 0], the first predicted state (MPC.q_next, MPC.v_next), packed
    double* qw;         // B x 6      dead-reckoned world pose MPC.q_w (MPC.py:58, 503-510), carried across ticks
    double* obj;        // B
    int32_t* status;    // B
    int32_t* sweeps;    // B
    int32_t* iters;     // B
    uint32_t* contact;  // B x 2 words (4N bits, N <= 16) or more
    uint32_t* active;   // B x ceil(20N/32)
    int32_t* fb_list;   // fallback queue (instance ids)
    int32_t* fb_count;  // its length
    unsigned int* canary;   // debug build (MPCQP_CANARY): 8 violation counters, null otherwise
    int32_t* fb_next;   // the next tick's copy of {fb_count, three work counters}: zeroed by this tick's stage-wise launch on the main stream (null: not this launch)
};

// MPC.py:503-510: the world pose advances by the first predicted pose, rotated by the current world yaw.  One thread per robot.
__device__ __forceinline__ void world_pose_step(double* __restrict__ qw, const double* xn) {
    double s, c;
    sincos(qw[5], &s, &c);
    qw[0] += c * xn[0] - s * xn[1];
    qw[1] += s * xn[0] + c * xn[1];
    qw[2] = xn[2]; qw[3] = xn[3]; qw[4] = xn[4];
    qw[5] += xn[5];
}

__host__ __device__ constexpr int tile_index(int I, int J) { return I * (I + 1) / 2 + J; }
__host__ __device__ constexpr int elem_off(int r, int c) { return (c >> 2) * 32 + r * 4 + (c & 3); }

// signature code per foot-step: sx, sy in {-1,0,+1} (which friction row is active), tz in {0 free,
// 1 apex (f = 0), 2 top (fz = fz_max)}
__device__ __forceinline__ uint8_t sig_pack(int sx, int sy, int tz) { return (uint8_t)((sx + 1) + 3 * (sy + 1) + 9 * tz); }
__device__ __forceinline__ void sig_unpack(uint8_t s, int& sx, int& sy, int& tz) {
    tz = s / 9; int r = s - 9 * tz; sy = r / 3 - 1; sx = r - 3 * (r / 3) - 1;
}
constexpr uint8_t SIG_FREE = 4;   // sx = sy = 0, tz = 0

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// ---------------------------------------------------------------------------------------------
// TMA-style bulk staging (cp.async.bulk global -> shared, completion on an mbarrier)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Blocked Cholesky of the (8 NT x 8 NT) SPD matrix held as lower 8x8 tiles in shared memory.
// Left-looking over tile columns; trailing updates and the panel solve run on the FP64 tensor
// pipe (DMMA m8n8k4).  The 8x8 diagonal tile is factorised redundantly in the registers of every
// lane of its owner warp (no shuffles on the pivot chain) and inverted column-per-lane.
// Tile I of a column belongs to warp I % NWARPS, so the serial diagonal work rotates over the
// warps (and therefore over the SM's four schedulers).
// On exit: off-diagonal tiles hold L_IJ, diagonal tiles hold inv(L_JJ) (lower triangular).
// Returns false (CTA-uniform) if a pivot is not positive.
// ---------------------------------------------------------------------------------------------
// sum_{K in [K0, K1)} L_IK L_JK'  as a C fragment, two independent DMMA chains.  Out of line on purpose:
// the tile loop calls it from six places and the kernel's hot loop has to stay inside the
// instruction cache (four CTAs per SM sit in different phases).
__device__ MPCQP_TILE_FN double2 syrk_sum(const double* __restrict__ rowI, const double* __restrict__ rowJ,
                                         int K0, int K1, int fo) {
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
    int K = K0;
    for (; K + 1 < K1; K += 2) {
        const double a0 = rowI[K * 64 + fo], a1 = rowI[K * 64 + 32 + fo];
        const double b0 = rowJ[K * 64 + fo], b1 = rowJ[K * 64 + 32 + fo];
        const double a2 = rowI[K * 64 + 64 + fo], a3 = rowI[K * 64 + 96 + fo];
        const double b2 = rowJ[K * 64 + 64 + fo], b3 = rowJ[K * 64 + 96 + fo];
        dmma884(c0, c1, a0, b0);
        dmma884(e0, e1, a2, b2);
        dmma884(c0, c1, a1, b1);
        dmma884(e0, e1, a3, b3);
    }
    if (K < K1) {
        const double a0 = rowI[K * 64 + fo], a1 = rowI[K * 64 + 32 + fo];
        const double b0 = rowJ[K * 64 + fo], b1 = rowJ[K * 64 + 32 + fo];
        dmma884(c0, c1, a0, b0);
        dmma884(e0, e1, a1, b1);
    }
    return make_double2(c0 + e0, c1 + e1);
}

// 8x8 diagonal tile in shared memory (both triangles valid) -> inv(chol(tile)), lower triangular,
// written back in place.  Every lane of the calling warp runs the same pivot chain in registers
// (rsqrt, scale, rank-1 update: no shuffle or memory hop between dependent steps); lane c < 8 then
// builds column c of the inverse by forward substitution.  Kept out of line so that its 36-double
// working set does not inflate the register allocation of the tile loop around it.
static __device__ __noinline__ void diag_factor_invert(double* __restrict__ D, int lane, int* __restrict__ flag) {
    double L[36];                                       // packed lower triangle
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) L[r * (r + 1) / 2 + c] = D[elem_off(r, c)];
    __syncwarp();
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double d = L[j * (j + 1) / 2 + j];
        ok = ok && (d > 0.0);
        const double inv = rsqrt(d);
        L[j * (j + 1) / 2 + j] = inv;                   // keep 1 / l_jj on the diagonal
#pragma unroll
        for (int i = j + 1; i < 8; ++i) L[i * (i + 1) / 2 + j] *= inv;
#pragma unroll
        for (int i = j + 1; i < 8; ++i)
#pragma unroll
            for (int c = j + 1; c <= i; ++c)
                L[i * (i + 1) / 2 + c] = fma(-L[i * (i + 1) / 2 + j], L[c * (c + 1) / 2 + j], L[i * (i + 1) / 2 + c]);
    }
    const int cc = lane & 7;
    double x[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < r; ++k) sum = (k >= cc) ? fma(L[r * (r + 1) / 2 + k], x[k], sum) : sum;
        const double dr = L[r * (r + 1) / 2 + r];
        x[r] = (r == cc) ? dr : ((r > cc) ? -sum * dr : 0.0);
    }
    if (lane < 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r) D[elem_off(r, lane)] = x[r];
        if (!ok) *flag = 0;
    }
}

// One tile of X = inv(L):  X_RJ = -inv(L_RR) * sum_{K=J..R-1} L_RK X_KJ   (C fragment in o0, o1).
// Needs rows < R of X and row R of L in shared memory.
__device__ MPCQP_TILE_FN double2 xinv_tile(const double* __restrict__ Wt, int R, int Jc, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int fo = g * 4 + t;                                        // A operand: element (g, t + 4kk)
    const int bo = (g >> 2) * 32 + t * 4 + (g & 3);                  // B operand: element (t + 4kk, g) -> + 16 kk
    const double* rowR = Wt + tile_index(R, 0) * 64;
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
    int K = Jc;
    for (; K + 1 < R; K += 2) {
        const double* XA = Wt + tile_index(K, Jc) * 64;
        const double* XB = Wt + tile_index(K + 1, Jc) * 64;
        const double a0 = rowR[K * 64 + fo], a1 = rowR[K * 64 + 32 + fo];
        const double a2 = rowR[K * 64 + 64 + fo], a3 = rowR[K * 64 + 96 + fo];
        const double b0 = XA[bo], b1 = XA[bo + 16], b2 = XB[bo], b3 = XB[bo + 16];
        dmma884(c0, c1, a0, b0);
        dmma884(e0, e1, a2, b2);
        dmma884(c0, c1, a1, b1);
        dmma884(e0, e1, a3, b3);
    }
    if (K < R) {
        const double* XA = Wt + tile_index(K, Jc) * 64;
        dmma884(c0, c1, rowR[K * 64 + fo], XA[bo]);
        dmma884(e0, e1, rowR[K * 64 + 32 + fo], XA[bo + 16]);
    }
    c0 += e0; c1 += e1;
    // B operand of the last product is the accumulator itself, re-laid by shuffles (rows t, t + 4)
    const double* Drr = Wt + tile_index(R, R) * 64;
    const int s0 = t * 4 + (g >> 1), s1 = s0 + 16;
    const double v00 = shfl_d(c0, s0), v01 = shfl_d(c1, s0);
    const double v10 = shfl_d(c0, s1), v11 = shfl_d(c1, s1);
    const double bb0 = (g & 1) ? v01 : v00;
    const double bb1 = (g & 1) ? v11 : v10;
    double d0 = 0.0, d1 = 0.0;
    dmma884(d0, d1, Drr[fo], bb0);
    dmma884(d0, d1, Drr[32 + fo], bb1);
    return make_double2(-d0, -d1);
}

// ---------------------------------------------------------------------------------------------
// Cholesky W = L L' of the (8 NT x 8 NT) SPD matrix held as lower 8x8 tiles in shared memory,
// fused with the in-place inversion of the factor: on exit every tile holds X = inv(L), so that
// every later solve W v = s is two fully parallel mat-vecs, v = X'(X s)  (tri_solve).
//
//   * left-looking over tile columns; trailing updates, panel solves and the inversion run on the
//     FP64 tensor pipe (DMMA m8n8k4);
//   * the 8x8 diagonal tile is factorised redundantly in the registers of every lane of its owner
//     warp (the pivot chain -- rsqrt, scale, update -- has no shuffle or memory hop in it) and
//     inverted column-per-lane; tile I of a column belongs to warp I % NWARPS so this serial part
//     rotates over the warps / the SM's four schedulers;
//   * the pivot chain is the critical path, so everything else is scheduled into its shadow:
//     while the owner factorises diagonal tile J, the other warps (a) pre-accumulate the update of
//     column J + 1 over the columns that are already final and (b) compute row J - 1 of X, which
//     overwrites row J - 1 of L (dead for the factorisation by then) one barrier later.
// Returns false (CTA-uniform) if a pivot is not positive.
// ---------------------------------------------------------------------------------------------
template <int NT, int NWARPS>
__device__ bool factor_invert_tiles(double* __restrict__ Wt, int* __restrict__ flag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int fo = g * 4 + t;                                   // operand fragment offset inside a panel
    const int co = (t >> 1) * 32 + g * 4 + (t & 1) * 2;         // C fragment offset (cols 2t, 2t+1 of row g)
    constexpr int MAXT = (NT + NWARPS - 1) / NWARPS;
    constexpr int XT = (NT - 1 + NWARPS - 2) / (NWARPS - 1);    // X-row tiles per non-owner warp
    if (threadIdx.x == 0) *flag = 1;
    double cur0[MAXT], cur1[MAXT], nxt0[MAXT], nxt1[MAXT], xr0[XT], xr1[XT];
#pragma unroll
    for (int m = 0; m < MAXT; ++m) {
        const int I = warp + m * NWARPS;
        cur0[m] = 0.0; cur1[m] = 0.0;
        if (I < NT) {
            const double2 w = *reinterpret_cast<const double2*>(Wt + tile_index(I, 0) * 64 + co);
            cur0[m] = w.x; cur1[m] = w.y;
        }
    }
    TPROF_T0();
    for (int J = 0; J < NT; ++J) {
        const int I0 = J + ((warp - J) & (NWARPS - 1));         // first tile row >= J owned by this warp
        const int I1 = J + 1 + ((warp - J - 1) & (NWARPS - 1)); // same for column J + 1
        const bool owner = warp == (J & (NWARPS - 1));
        const int rank = (warp - J - 1) & (NWARPS - 1);         // 0 .. NWARPS-2 among the non-owners
        const double* rowN = Wt + tile_index(J + 1 < NT ? J + 1 : J, 0) * 64;
        double* D = Wt + tile_index(J, J) * 64;
#pragma unroll
        for (int m = 0; m < MAXT; ++m) { nxt0[m] = 0.0; nxt1[m] = 0.0; }
        if (owner) {
            // ---- diagonal tile, factorised redundantly by every lane of the owner warp
            *reinterpret_cast<double2*>(D + co) = make_double2(cur0[0], cur1[0]);
            __syncwarp();
            diag_factor_invert(D, lane, flag);
            TPROF(0);
        } else {
            // ---- in the shadow of the pivot chain: (a) look-ahead of column J + 1 over K < J
            if (J + 1 < NT) {
#pragma unroll
                for (int m = 0; m < MAXT; ++m) {
                    const int I = I1 + m * NWARPS;
                    if (I < NT) { const double2 r = syrk_sum(Wt + tile_index(I, 0) * 64, rowN, 0, J, fo); nxt0[m] = r.x; nxt1[m] = r.y; }
                }
            }
            // (b) row J - 1 of X (kept in registers until the barrier below)
            if (J >= 2) {
#pragma unroll
                for (int m = 0; m < XT; ++m) {
                    const int Jc = rank + m * (NWARPS - 1);
                    if (Jc < J - 1) { const double2 r = xinv_tile(Wt, J - 1, Jc, lane); xr0[m] = r.x; xr1[m] = r.y; }
                }
            }
            TPROF(1);
        }
        __syncthreads();
        TPROF(2);
        // ---- panel: L_IJ = C_IJ * inv(L_JJ)'   (A = C fragment re-laid by shuffles, B[k][n] = Linv[n][k])
        const double b0 = D[fo], b1 = D[32 + fo];
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            const int I = I0 + m * NWARPS;
            if (I < NT && I > J) {
                const int s0 = g * 4 + (t >> 1), s1 = s0 + 2;
                const double v00 = shfl_d(cur0[m], s0), v01 = shfl_d(cur1[m], s0);
                const double v10 = shfl_d(cur0[m], s1), v11 = shfl_d(cur1[m], s1);
                const double a0 = (t & 1) ? v01 : v00;
                const double a1 = (t & 1) ? v11 : v10;
                double d0 = 0.0, d1 = 0.0;
                dmma884(d0, d1, a0, b0);
                dmma884(d0, d1, a1, b1);
                *reinterpret_cast<double2*>(Wt + tile_index(I, J) * 64 + co) = make_double2(d0, d1);
            }
        }
        if (!owner && J >= 2) {
#pragma unroll
            for (int m = 0; m < XT; ++m) {
                const int Jc = rank + m * (NWARPS - 1);
                if (Jc < J - 1) *reinterpret_cast<double2*>(Wt + tile_index(J - 1, Jc) * 64 + co) = make_double2(xr0[m], xr1[m]);
            }
        }
        TPROF(3);
        __syncthreads();
        TPROF(4);
        // ---- finish column J + 1: add the K = J term (the owner of J also its K < J part), C = W - sum
        if (J + 1 < NT) {
#pragma unroll
            for (int m = 0; m < MAXT; ++m) {
                const int I = I1 + m * NWARPS;
                cur0[m] = 0.0; cur1[m] = 0.0;
                if (I < NT) {
                    const double* rowI = Wt + tile_index(I, 0) * 64;
                    const double2 r = syrk_sum(rowI, rowN, owner ? 0 : J, J + 1, fo);
                    const double2 w = *reinterpret_cast<const double2*>(Wt + tile_index(I, J + 1) * 64 + co);
                    cur0[m] = w.x - (nxt0[m] + r.x);
                    cur1[m] = w.y - (nxt1[m] + r.y);
                }
            }
        }
        TPROF(5);
    }
    // ---- last row of X, all warps
    {
        double l0[MAXT], l1[MAXT];
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            const int Jc = warp + m * NWARPS;
            if (Jc < NT - 1) { const double2 r = xinv_tile(Wt, NT - 1, Jc, lane); l0[m] = r.x; l1[m] = r.y; }
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            const int Jc = warp + m * NWARPS;
            if (Jc < NT - 1) *reinterpret_cast<double2*>(Wt + tile_index(NT - 1, Jc) * 64 + co) = make_double2(l0[m], l1[m]);
        }
        __syncthreads();
    }
    return *flag != 0;
}

// v = X' (X s) = inv(L L') s, in place in shared memory (s has 8 NT entries, tmp is scratch of the
// same size).  Every warp reads whole 8x8 tiles with the C-fragment access pattern (one conflict-
// free 16-byte load per lane): lane (g, t) holds X[g][2t], X[g][2t+1].  Warp w owns tile rows
// (pass 1) / tile columns (pass 2) w, w + NWARPS, ...   Ends with a __syncthreads().
template <int NT, int NWARPS>
__device__ void tri_solve(const double* __restrict__ Wt, double* __restrict__ s, double* __restrict__ tmp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int co = (t >> 1) * 32 + g * 4 + (t & 1) * 2;
    constexpr int MAXT = (NT + NWARPS - 1) / NWARPS;
    // ---- y = X s : row 8I + g, partial over columns (2t, 2t+1) of every tile, reduced over t
    {
        double a0[MAXT], a1[MAXT];
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            const int I = warp + m * NWARPS;
            a0[m] = 0.0; a1[m] = 0.0;
            if (I < NT) {
                const double* T = Wt + tile_index(I, 0) * 64 + co;
#pragma unroll 2
                for (int J = 0; J <= I; ++J) {
                    const double2 x = *reinterpret_cast<const double2*>(T + J * 64);
                    const double2 sv = *reinterpret_cast<const double2*>(s + 8 * J + 2 * t);
                    a0[m] = fma(x.x, sv.x, a0[m]);
                    a1[m] = fma(x.y, sv.y, a1[m]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            double v = a0[m] + a1[m];
            v += shfl_xor_d(v, 1);
            v += shfl_xor_d(v, 2);
            const int I = warp + m * NWARPS;
            if (I < NT && t == 0) tmp[8 * I + g] = v;
        }
    }
    __syncthreads();
    // ---- v = X' y : columns 8J + 2t, 8J + 2t + 1, partial over row g of every tile, reduced over g
    {
        double a0[MAXT], a1[MAXT];
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            const int J = warp + m * NWARPS;
            a0[m] = 0.0; a1[m] = 0.0;
            if (J < NT) {
#pragma unroll 2
                for (int I = J; I < NT; ++I) {
                    const double2 x = *reinterpret_cast<const double2*>(Wt + tile_index(I, J) * 64 + co);
                    const double yv = tmp[8 * I + g];
                    a0[m] = fma(x.x, yv, a0[m]);
                    a1[m] = fma(x.y, yv, a1[m]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            double v0 = a0[m], v1 = a1[m];
            v0 += shfl_xor_d(v0, 4);  v1 += shfl_xor_d(v1, 4);
            v0 += shfl_xor_d(v0, 8);  v1 += shfl_xor_d(v1, 8);
            v0 += shfl_xor_d(v0, 16); v1 += shfl_xor_d(v1, 16);
            const int J = warp + m * NWARPS;
            if (J < NT && g == 0) *reinterpret_cast<double2*>(s + 8 * J + 2 * t) = make_double2(v0, v1);
        }
    }
    __syncthreads();
}

}  // namespace mpcqp
