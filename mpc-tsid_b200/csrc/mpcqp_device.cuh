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

namespace mpcqp {

struct DevParams {
    int N, batch;
    double dt, mass, mu, fz_max, gravity, w_force;
    double gIinv[9];        // inverse body inertia, row major (host, extended precision)
    double footholds[12];   // 3 x 4 row major
    double wp[6], wv[6];    // state weights: position-like (x y z roll pitch yaw) and velocity-like
    double rho, sigma, alpha, feas_tol, dual_tol;
    int max_sweeps, max_iter, min_iter, check_every, warm_start, mode, refine;
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
    double* obj;        // B
    int32_t* status;    // B
    int32_t* sweeps;    // B
    int32_t* iters;     // B
    uint32_t* contact;  // B x 2 words (4N bits, N <= 16) or more
    uint32_t* active;   // B x ceil(20N/32)
    int32_t* fb_list;   // fallback queue (instance ids)
    int32_t* fb_count;  // its length
};

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
template <int NT, int NWARPS>
__device__ bool cholesky_tiles(double* __restrict__ Wt, int* __restrict__ flag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int fo = g * 4 + t;                                   // operand fragment offset inside a panel
    const int co = (t >> 1) * 32 + g * 4 + (t & 1) * 2;         // C fragment offset (cols 2t, 2t+1 of row g)
    constexpr int MAXT = (NT + NWARPS - 1) / NWARPS;
    if (threadIdx.x == 0) *flag = 1;
    for (int J = 0; J < NT; ++J) {
        double c0[MAXT], c1[MAXT];
        const int I0 = J + ((warp - J) & (NWARPS - 1));         // first tile row >= J owned by this warp
        const double* rowJ = Wt + tile_index(J, 0) * 64;
        // ---- trailing update of tile column J (two accumulator pairs per tile: even / odd K)
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            const int I = I0 + m * NWARPS;
            c0[m] = 0.0; c1[m] = 0.0;
            if (I < NT) {
                const double* rowI = Wt + tile_index(I, 0) * 64;
                double e0 = 0.0, e1 = 0.0;
                int K = 0;
                for (; K + 1 < J; K += 2) {
                    const double a0 = rowI[K * 64 + fo], a1 = rowI[K * 64 + 32 + fo];
                    const double b0 = rowJ[K * 64 + fo], b1 = rowJ[K * 64 + 32 + fo];
                    const double a2 = rowI[K * 64 + 64 + fo], a3 = rowI[K * 64 + 96 + fo];
                    const double b2 = rowJ[K * 64 + 64 + fo], b3 = rowJ[K * 64 + 96 + fo];
                    dmma884(c0[m], c1[m], a0, b0);
                    dmma884(e0, e1, a2, b2);
                    dmma884(c0[m], c1[m], a1, b1);
                    dmma884(e0, e1, a3, b3);
                }
                if (K < J) {
                    const double a0 = rowI[K * 64 + fo], a1 = rowI[K * 64 + 32 + fo];
                    const double b0 = rowJ[K * 64 + fo], b1 = rowJ[K * 64 + 32 + fo];
                    dmma884(c0[m], c1[m], a0, b0);
                    dmma884(e0, e1, a1, b1);
                }
                const double2 w = *reinterpret_cast<const double2*>(Wt + tile_index(I, J) * 64 + co);
                c0[m] = w.x - (c0[m] + e0);
                c1[m] = w.y - (c1[m] + e1);
            }
        }
        // ---- diagonal tile: owner warp J % NWARPS holds it in slot 0
        double* D = Wt + tile_index(J, J) * 64;
        if (warp == (J & (NWARPS - 1))) {
            *reinterpret_cast<double2*>(D + co) = make_double2(c0[0], c1[0]);
            __syncwarp();
            double L[36];                                       // packed lower triangle, every lane the same
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
            // X = inv(L): lane c (< 8) builds column c by forward substitution, all in registers
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
        __syncthreads();
        // ---- panel: L_IJ = C_IJ * inv(L_JJ)'   (A = C fragment re-laid by shuffles, B[k][n] = Linv[n][k])
        const double b0 = D[fo], b1 = D[32 + fo];
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            const int I = I0 + m * NWARPS;
            if (I < NT && I > J) {
                const int s0 = g * 4 + (t >> 1), s1 = s0 + 2;
                const double v00 = shfl_d(c0[m], s0), v01 = shfl_d(c1[m], s0);
                const double v10 = shfl_d(c0[m], s1), v11 = shfl_d(c1[m], s1);
                const double a0 = (t & 1) ? v01 : v00;
                const double a1 = (t & 1) ? v11 : v10;
                double d0 = 0.0, d1 = 0.0;
                dmma884(d0, d1, a0, b0);
                dmma884(d0, d1, a1, b1);
                *reinterpret_cast<double2*>(Wt + tile_index(I, J) * 64 + co) = make_double2(d0, d1);
            }
        }
        __syncthreads();
    }
    return *flag != 0;
}

// ---------------------------------------------------------------------------------------------
// In-place inverse of the block lower-triangular factor: on entry off-diagonal tiles hold L_IJ and
// diagonal tiles inv(L_JJ); on exit every tile holds X = inv(L).  Row by row,
//     X_IJ = -inv(L_II) * sum_{K=J..I-1} L_IK X_KJ,
// all tiles of a row in parallel on the FP64 tensor pipe (tile J of the row -> warp J % NWARPS).
// With X explicit the two triangular solves of every later linear solve become two fully parallel
// mat-vecs (tri_solve) instead of 2 * NT dependent block steps.
// ---------------------------------------------------------------------------------------------
template <int NT, int NWARPS>
__device__ void invert_tiles(double* __restrict__ Wt) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int fo = g * 4 + t;                                        // A operand: element (g, t + 4kk)
    const int bo = (g >> 2) * 32 + t * 4 + (g & 3);                  // B operand: element (t + 4kk, g) -> + 16 kk
    const int co = (t >> 1) * 32 + g * 4 + (t & 1) * 2;
    constexpr int MAXT = (NT + NWARPS - 1) / NWARPS;
    for (int I = 1; I < NT; ++I) {
        double c0[MAXT], c1[MAXT];
        const double* rowI = Wt + tile_index(I, 0) * 64;
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            const int J = warp + m * NWARPS;
            c0[m] = 0.0; c1[m] = 0.0;
            if (J < I) {
                double e0 = 0.0, e1 = 0.0;
                int K = J;
                for (; K + 1 < I; K += 2) {
                    const double* XA = Wt + tile_index(K, J) * 64;
                    const double* XB = Wt + tile_index(K + 1, J) * 64;
                    const double a0 = rowI[K * 64 + fo], a1 = rowI[K * 64 + 32 + fo];
                    const double a2 = rowI[K * 64 + 64 + fo], a3 = rowI[K * 64 + 96 + fo];
                    const double b0 = XA[bo], b1 = XA[bo + 16], b2 = XB[bo], b3 = XB[bo + 16];
                    dmma884(c0[m], c1[m], a0, b0);
                    dmma884(e0, e1, a2, b2);
                    dmma884(c0[m], c1[m], a1, b1);
                    dmma884(e0, e1, a3, b3);
                }
                if (K < I) {
                    const double* XA = Wt + tile_index(K, J) * 64;
                    const double a0 = rowI[K * 64 + fo], a1 = rowI[K * 64 + 32 + fo];
                    dmma884(c0[m], c1[m], a0, XA[bo]);
                    dmma884(e0, e1, a1, XA[bo + 16]);
                }
                c0[m] += e0; c1[m] += e1;
                // X_IJ = -inv(L_II) * acc : A = inv(L_II) from smem, B = acc re-laid by shuffles
                const double* Dii = Wt + tile_index(I, I) * 64;
                const int s0 = t * 4 + (g >> 1), s1 = s0 + 16;       // lanes holding rows t, t + 4 of acc
                const double v00 = shfl_d(c0[m], s0), v01 = shfl_d(c1[m], s0);
                const double v10 = shfl_d(c0[m], s1), v11 = shfl_d(c1[m], s1);
                const double bb0 = (g & 1) ? v01 : v00;
                const double bb1 = (g & 1) ? v11 : v10;
                double d0 = 0.0, d1 = 0.0;
                dmma884(d0, d1, Dii[fo], bb0);
                dmma884(d0, d1, Dii[32 + fo], bb1);
                c0[m] = -d0; c1[m] = -d1;
            }
        }
        __syncthreads();                                             // every read of row I's L tiles is done
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            const int J = warp + m * NWARPS;
            if (J < I) *reinterpret_cast<double2*>(Wt + tile_index(I, J) * 64 + co) = make_double2(c0[m], c1[m]);
        }
        __syncthreads();
    }
}

// v = X' (X s) = inv(L L') s, in place in shared memory (s has 8 NT entries).  All threads take
// part; `tmp` is 8 NT doubles of scratch.  Ends with a __syncthreads().
template <int NT>
__device__ void tri_solve(const double* __restrict__ Wt, double* __restrict__ s, double* __restrict__ tmp) {
    const int n = 8 * NT;
    // y = X s : one row per thread, long rows first (thread 0 takes the last row)
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int i = n - 1 - idx, I = i >> 3, r = i & 7;
        const double* T = Wt + tile_index(I, 0) * 64 + r * 4;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int J = 0; J <= I; ++J) {
            const double2 p0 = *reinterpret_cast<const double2*>(T + J * 64);
            const double2 p1 = *reinterpret_cast<const double2*>(T + J * 64 + 2);
            const double2 p2 = *reinterpret_cast<const double2*>(T + J * 64 + 32);
            const double2 p3 = *reinterpret_cast<const double2*>(T + J * 64 + 34);
            const double* sv = s + 8 * J;
            a0 = fma(p0.x, sv[0], a0); a1 = fma(p0.y, sv[1], a1); a2 = fma(p1.x, sv[2], a2); a3 = fma(p1.y, sv[3], a3);
            a0 = fma(p2.x, sv[4], a0); a1 = fma(p2.y, sv[5], a1); a2 = fma(p3.x, sv[6], a2); a3 = fma(p3.y, sv[7], a3);
        }
        tmp[i] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    // v = X' y : one column per thread, long columns first
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int J = j >> 3, c = j & 7;
        const int off = (c >> 2) * 32 + (c & 3);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int I = J; I < NT; ++I) {
            const double* T = Wt + tile_index(I, J) * 64 + off;
            const double* yv = tmp + 8 * I;
            a0 = fma(T[0], yv[0], a0); a1 = fma(T[4], yv[1], a1); a2 = fma(T[8], yv[2], a2); a3 = fma(T[12], yv[3], a3);
            a0 = fma(T[16], yv[4], a0); a1 = fma(T[20], yv[5], a1); a2 = fma(T[24], yv[6], a2); a3 = fma(T[28], yv[7], a3);
        }
        s[j] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
}

}  // namespace mpcqp
