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
    const double* M;            // 6 x N x N
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
// Blocked Cholesky of the (6N x 6N) SPD matrix held as lower 8x8 tiles in shared memory.
// Left-looking over tile columns; trailing updates and the panel solve run on the FP64 tensor
// pipe (DMMA m8n8k4); the 8x8 diagonal tile is factorised and inverted in registers.
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
        // ---- trailing update of tile column J
#pragma unroll
        for (int m = 0; m < MAXT; ++m) {
            const int I = J + warp + m * NWARPS;
            c0[m] = 0.0; c1[m] = 0.0;
            if (I < NT) {
                const double* rowI = Wt + tile_index(I, 0) * 64;
                const double* rowJ = Wt + tile_index(J, 0) * 64;
                for (int K = 0; K < J; ++K) {
                    const double a0 = rowI[K * 64 + fo], a1 = rowI[K * 64 + 32 + fo];
                    const double b0 = rowJ[K * 64 + fo], b1 = rowJ[K * 64 + 32 + fo];
                    dmma884(c0[m], c1[m], a0, b0);
                    dmma884(c0[m], c1[m], a1, b1);
                }
                const double2 w = *reinterpret_cast<const double2*>(Wt + tile_index(I, J) * 64 + co);
                c0[m] = w.x - c0[m];
                c1[m] = w.y - c1[m];
            }
        }
        // ---- diagonal tile (owner: warp 0, slot 0): Cholesky + triangular inverse, rows spread over
        //      lanes 0..7 and exchanged by shuffles (8 doubles of state per lane, no local arrays)
        double* D = Wt + tile_index(J, J) * 64;
        if (warp == 0) {
            *reinterpret_cast<double2*>(D + co) = make_double2(c0[0], c1[0]);
            __syncwarp();
            double a[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) a[c] = (lane < 8 && c <= lane) ? D[elem_off(lane & 7, c)] : 0.0;
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double dj = shfl_d(a[j], j);
                ok = ok && (dj > 0.0);
                const double inv = rsqrt(dj);
                const double lij = a[j] * inv;                 // l_ij for lanes i > j
#pragma unroll
                for (int c = j + 1; c < 8; ++c) {
                    const double lcj = shfl_d(lij, c);
                    a[c] = fma(-lij, lcj, a[c]);
                }
                a[j] = (lane == j) ? inv : lij;                // lane j keeps 1 / l_jj
            }
            // X = inv(L): lane c builds column c by forward substitution on the broadcast rows of L
            double x[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                double sum = 0.0;
#pragma unroll
                for (int k = 0; k < r; ++k) {
                    const double lrk = shfl_d(a[k], r);
                    sum = (k >= lane) ? fma(lrk, x[k], sum) : sum;
                }
                const double dr = shfl_d(a[r], r);
                x[r] = (r == lane) ? dr : ((r > lane) ? -sum * dr : 0.0);
            }
            __syncwarp();
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
            const int I = J + warp + m * NWARPS;
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

// Solve L L' v = s in place (s in shared memory, length 8*NT), executed by warp 0 only.
// Diagonal tiles hold inv(L_JJ).  Callers must __syncthreads() before and after.
template <int NT>
__device__ void solve_tiles_warp0(const double* __restrict__ Wt, double* __restrict__ s) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    // forward: L w = s
    for (int J = 0; J < NT; ++J) {
        const double* D = Wt + tile_index(J, J) * 64;
        double vj = 0.0;
        if (lane < 8) {
#pragma unroll
            for (int c = 0; c < 8; ++c) vj = fma(D[elem_off(lane, c)], s[J * 8 + c], vj);
        }
        __syncwarp();
        if (lane < 8) s[J * 8 + lane] = vj;
        __syncwarp();
        double v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = s[J * 8 + c];
        for (int row = (J + 1) * 8 + lane; row < NT * 8; row += 32) {
            const double* T = Wt + tile_index(row >> 3, J) * 64;
            const int r = row & 7;
            double acc = s[row];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc = fma(-T[elem_off(r, c)], v[c], acc);
            s[row] = acc;
        }
        __syncwarp();
    }
    // backward: L' v = w
    for (int J = NT - 1; J >= 0; --J) {
        const double* D = Wt + tile_index(J, J) * 64;
        double vj = 0.0;
        if (lane < 8) {
#pragma unroll
            for (int r = 0; r < 8; ++r) vj = fma(D[elem_off(r, lane)], s[J * 8 + r], vj);   // inv(L)' row = column of inv(L)
        }
        __syncwarp();
        if (lane < 8) s[J * 8 + lane] = vj;
        __syncwarp();
        double v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = s[J * 8 + r];
        for (int col = lane; col < J * 8; col += 32) {
            const double* T = Wt + tile_index(J, col >> 3) * 64;
            const int c = col & 7;
            double acc = s[col];
#pragma unroll
            for (int r = 0; r < 8; ++r) acc = fma(-T[elem_off(r, c)], v[r], acc);
            s[col] = acc;
        }
        __syncwarp();
    }
}

}  // namespace mpcqp
