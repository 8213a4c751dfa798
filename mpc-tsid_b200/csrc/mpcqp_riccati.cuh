// mpcqp_riccati.cuh -- stage-wise (Riccati) active-set sweeps: HALF A WARP solves one robot's QP.
//
// Same mathematics as the dense path (condensed QP on the faces a signature selects,
// KKT guard, warm-started primal-dual active-set sweeps), different factorisation.  Instead of the
// 6N x 6N Woodbury matrix W (O(N^3) per sweep, one CTA per robot) the equality-constrained QP
//     min  sum_s 1/2 (x_s - xref_s)' Q (x_s - xref_s) + 1/2 w_f sum |f|^2
//     s.t. x_{k+1} = A x_k + [0; u_k + g],  u_k = sum_j Bv_kj f_kj,  f_kj = pf_kj + Z_kj q_kj
// is solved by dynamic programming over the horizon (O(N), 6x6 blocks only), x = [p (6); v (6)],
// A = [[I, dt I], [0, I]] (MPC.py:110-111), cost-to-go 1/2 x'P_k x + p_k'x.
//   backward, stage k = N-1..0, with P+ = P_{k+1}, E_k = sum_j (Bv Z) R^-1 (Bv Z)' (6x6), beta_k = ubar_k + g:
//       L L' = Pvv+,   G = I + L' E_k L = M M',   Gamma = (I + E_k Pvv+)^-1 E_k = L^-T (I - G^-1) L^-1
//       every row rho of [Ppv+; Pvv+; pv+'] (13 rows, one per lane) goes through the same four products
//           y = rho L^-T,  v = y G^-1 = (y M^-T) M^-1,  rho (I - Gamma Pvv+) = v L',  rho Gamma = (y - v) L^-1
//       (computed square-root free: Pvv+ = U D U', H = D^-1 + U' E_k U = W Delta W' with U, W unit lower triangular, so
//        that with yt = rho U^-T and e = ((yt D^-1) W^-T Delta^-1) W^-1:  rho (I - Gamma Pvv+) = e U',  rho Gamma = ((yt - e) D^-1) U^-1)
//       which yields Pt = (P+^-1 + [0 0; 0 E])^-1 (its [:, v] block), the closed-loop gain [Ppv; Pvv] Gamma and
//       Gamma pv; the [:, p] block follows from  Pt[:, p] = P+[:, p] - (rho Gamma) Pvp+.
//       P_k = Q + A' Pt A,   p_k = -Q xref_k + A'(Pt[:,v] beta_k + pt)
//   forward:  z = A x_k + [0; beta_k],  w = -([Ppv;Pvv] Gamma)' z - Gamma pv,  x_{k+1} = z + [0; w]
//   costates  lam_s = Q (x_s - xref_s) + A' lam_{s+1};
//   per foot: h = Bv' lam^v_{k+1},  q = -R^-1 Z' h,  f = pf + Z q,  grad = w_f f + h   (= H f + g of the
//   condensed problem), then the same KKT guard as the dense path (mpcqp_foot.cuh).
// Mapping to the machine.  The path is bound by the latency of the two 6x6 pivot chains per stage, so the
// design maximises the number of robots in flight per SM: 16 lanes per robot, two robots per warp that run
// CONVERGED (one instruction stream, full-mask collectives of width 16, warp-uniform control flow; partial masks
// leave the halves diverged and issue everything twice), 12 KB of shared memory per robot, the per-stage gains (the
// only O(N) state of the recursion that must survive until the forward pass) in an L2-resident workspace.
// Code size is part of the design: every per-foot loop is rolled (the unrolled version spent 16 % of its stall
// samples on instruction fetch).
// The two factorisations of a stage are computed redundantly in the REGISTERS of every lane (static indices, no shuffle or
// memory hop on the pivot chain; pivot to pivot is a 4-instruction reciprocal and one fused multiply-add), so every product
// above is "one row in registers times / substituted through a register-resident unit triangular matrix" and a stage needs
// three warp barriers.  Capacities 16 and 32 substitute with the factors themselves; capacity 64 also forms the inverse
// rows of the unit factors in the shadow of the pivots and multiplies by them (see ldl6_regs).
// Replaces MPC.update_ML / update_NK / call_solver / retrieve_result (MPC.py:316-458) like the dense path.
//
// Horizons: the kernels are compiled for three CAPACITIES NC = 16, 32, 64 (array sizes, lanes per robot, feet per lane); the
// horizon itself, n = P.N <= NC, is a run-time value (any n_steps = n_periods T_gait / dt of the reference, main.py:20-23,
// FootstepPlanner.py:52-63): xref is 12 x (n + 1), foot-steps t >= 4 n do not exist, the recursion runs over n stages.
//
// Robots the warm-started sweeps do not certify (cycling signatures: cold starts of long horizons, nearly degenerate faces)
// go to a second kernel, ipm_kernel, that runs a primal-dual INTERIOR-POINT iteration on the same stage-wise factorisation:
// per foot the barrier adds  C' diag(y / s) C  (a full symmetric 3 x 3) to the force weight, whose inverse S is exactly the
// 3 x 3 the stage assembly consumes, so one iteration = one ric_assemble + ric_core + per-foot step rules.  About twenty
// iterations bring the complementarity gap to 1e-8 whatever the active set looks like; the rows with y > s then seed the sweeps,
// which certify the optimum exactly as on the fast path (same KKT guard).  This is the globally convergent stage of the
// stage-wise path, the role OSQP's ADMM iteration has in the reference (MPC.py:414-428).
#pragma once
#include "mpcqp_device.cuh"
#include "mpcqp_foot.cuh"
#include "mpcqp_scenario.cuh"

// Optional phase timing (-DMPCQP_PROFILE): clock64() deltas of lane 0 summed into g_prof[16..] (mpcqp_kernels.cu)
#ifdef MPCQP_PROFILE
#define RPROF_T0() long long rprof_t_ = clock64()
#define RPROF(slot) do { if ((threadIdx.x & 31) == 0) { long long n_ = clock64(); atomicAdd(&g_prof[16 + (slot)], (unsigned long long)(n_ - rprof_t_)); rprof_t_ = n_; } } while (0)
#define RPROF_COUNT(slot) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_prof[16 + (slot)], 1ull); } while (0)
#else
#define RPROF_T0() do {} while (0)
#define RPROF(slot) do {} while (0)
#define RPROF_COUNT(slot) do {} while (0)
#endif

namespace mpcqp {

}  // namespace mpcqp
#include "mpcqp_ric_consts.h"
namespace mpcqp {
#ifdef MPCQP_CANARY
// DevState.canary: [0] a shared-memory guard word was overwritten, [1] the guard behind a slot's gains, [2] behind its
// interior-point state, [3] an index invariant failed (queue slot / robot id out of range)
constexpr unsigned long long RIC_CANARY_WORD = 0xC0FFEE0DDEADBEEFull;
#define RIC_GUARD(name) alignas(16) unsigned long long name[2];
#else
#define RIC_GUARD(name)
#endif
constexpr int RIC_DEPTH = 4;        // stages of gains the forward pass keeps in flight from the workspace (5 and 6 measured slower: profiles/r02_kernel_variants.md)

// cost-to-go of one stage, row major 6x6 blocks (double-buffered: stage k reads one, writes the other)
struct alignas(16) RicCost {
    double Ppp[36], Ppv[36], Pvv[36], pp[6], pv[6];
};

template <int N>                        // N = capacity (16, 32, 64); the run-time horizon n <= N
struct alignas(16) RicInst {
    static constexpr int NF = 4 * N;
    static constexpr int ROUNDS = NF / 16;                     // feet per lane
    static constexpr int AW = (20 * N + 31) / 32, CW = (4 * N + 31) / 32;
    double xr[12 * (N + 1)];            // xref of this robot
    RIC_GUARD(g0)
    union {
        double fs[20 * 13];             // fsteps (dead after decode)
        double E[21 * N];               // per sweep: packed lower triangles of the 6x6 blocks E_k; once the
                                        // backward pass is done the same bytes hold the sweep's forces (3 x NF)
    };
    RIC_GUARD(g1)
    double lev[3 * NF];                 // lever arms foothold - xref[0:3, k], struct of arrays
    RIC_GUARD(g2)
    // inv(R_z(yaw_k) gI) per step, row major -- or, at capacity 64, only cos / sin of yaw_k (the block is re-formed where it is used:
    // 3.5 KB less per robot, which is what lets a third CTA fit on an SM there)
    static constexpr bool COMPACT_II = N >= 64;
    double Ii[COMPACT_II ? 2 * N : 9 * N];
    RIC_GUARD(g3)
    union {
        double beta[6 * N];             // ubar_k + g (dead once the forward pass is done)
        double lam[6 * N];              // velocity costates lam^v_1..lam^v_N (written by the costate pass that follows it)
    };
    RIC_GUARD(g4)
    union {
        double xst[12 * N];             // states x_1..x_N of the forward pass
        ScenarioSmem sc;                // planner scratch of the device-resident closed loop (dead after the inputs exist)
    };
    RIC_GUARD(g5)
    RicCost cost[2];
    RIC_GUARD(g6)
    double T[36];                       // E_k L, row major
    double hp[6];
    double xnext[12];
    RIC_GUARD(g7)
    unsigned long long hist[16];        // hashes of the signatures already tried
    unsigned long long mbar;
    unsigned int amask[AW + CW];
    uint8_t sigb[NF];                   // per foot-step: signature of the sweep being assembled | contact << 7
    RIC_GUARD(g8)
    static_assert(21 * N >= 260 && 21 * N >= 12 * N, "union sizing");
    static_assert(12 * N * 8 >= sizeof(ScenarioSmem), "union sizing");
};

#define RIC_TI(r, c) ((r) * ((r) + 1) / 2 + (c))

#ifdef MPCQP_CANARY
// one lane per half-warp arms / checks the guards of its robot's shared memory and workspace slot
template <int N>
__device__ void canary_arm(RicInst<N>& sm, double* ws) {
    unsigned long long* g[9] = {sm.g0, sm.g1, sm.g2, sm.g3, sm.g4, sm.g5, sm.g6, sm.g7, sm.g8};
    for (int i = 0; i < 9; ++i) { g[i][0] = RIC_CANARY_WORD; g[i][1] = ~RIC_CANARY_WORD; }
    unsigned long long* a = reinterpret_cast<unsigned long long*>(ws + (size_t)RIC_GAIN * N);
    unsigned long long* b = reinterpret_cast<unsigned long long*>(ws + (size_t)ric_ws_slot_doubles(N) - RIC_WS_PAD);
    a[0] = RIC_CANARY_WORD; a[1] = ~RIC_CANARY_WORD; b[0] = RIC_CANARY_WORD; b[1] = ~RIC_CANARY_WORD;
}
template <int N>
__device__ void canary_check(const RicInst<N>& sm, const double* ws, unsigned int* g_canary_errors) {
    const unsigned long long* g[9] = {sm.g0, sm.g1, sm.g2, sm.g3, sm.g4, sm.g5, sm.g6, sm.g7, sm.g8};
    for (int i = 0; i < 9; ++i)
        if (g[i][0] != RIC_CANARY_WORD || g[i][1] != ~RIC_CANARY_WORD) atomicAdd(&g_canary_errors[0], 1u);
    const unsigned long long* a = reinterpret_cast<const unsigned long long*>(ws + (size_t)RIC_GAIN * N);
    const unsigned long long* b = reinterpret_cast<const unsigned long long*>(ws + (size_t)ric_ws_slot_doubles(N) - RIC_WS_PAD);
    if (a[0] != RIC_CANARY_WORD || a[1] != ~RIC_CANARY_WORD) atomicAdd(&g_canary_errors[1], 1u);
    if (b[0] != RIC_CANARY_WORD || b[1] != ~RIC_CANARY_WORD) atomicAdd(&g_canary_errors[2], 1u);
}
#define RIC_CANARY_ARM() do { if (hl == 0) canary_arm<N>(sm, ws); __syncwarp(); } while (0)
#define RIC_CANARY_CHECK() do { __syncwarp(); if (hl == 0) canary_check<N>(sm, ws, st.canary); } while (0)
#define RIC_INVARIANT(cond) do { if (!(cond)) atomicAdd(&st.canary[3], 1u); } while (0)
#else
#define RIC_CANARY_ARM() do {} while (0)
#define RIC_CANARY_CHECK() do {} while (0)
#define RIC_INVARIANT(cond) do {} while (0)
#endif

// The two robots of a warp run CONVERGED: one instruction stream, full-mask collectives, shuffles of width 16.
constexpr unsigned RIC_FULL = 0xffffffffu;
__device__ __forceinline__ double hshfl_d(double v, int src) { return __shfl_sync(RIC_FULL, v, src, 16); }
// true on every lane of a half-warp iff the predicate holds on all / any of its 16 lanes
__device__ __forceinline__ bool half_all(bool p, int sub) { return ((__ballot_sync(RIC_FULL, p) >> (16 * sub)) & 0xFFFFu) == 0xFFFFu; }
__device__ __forceinline__ bool half_any(bool p, int sub) { return ((__ballot_sync(RIC_FULL, p) >> (16 * sub)) & 0xFFFFu) != 0u; }

// 1 / d: hardware seed (relative error ~2^-22) and one third-order correction, three dependent instructions after the seed.
// d <= 0 or non-finite gives garbage; the caller tests d separately.
__device__ __forceinline__ double rcp_fast(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
}

// Square-root-free factorisation a = U D U' of a 6x6 SPD matrix held as packed lower triangle in registers, in place
// (U unit lower triangular: its strict lower part replaces a's, the diagonal of `a` is left as garbage), with
// dinv = 1 / D and ui = inv(U) (unit lower, strict part only is meaningful).  The pivot-to-pivot dependency is
// 1 / d_j (seed + 3) and ONE fused multiply-add: the next pivot is updated from the unscaled column, whose square is ready
// before the reciprocal is.  Row j of the inverse needs only rows < j of it and row j of U, so it is formed in the shadow
// of the pivots.  All indices are static.
// Two row vectors ride along in the shadow of the pivots: `fs` is overwritten by fs U^-T (forward substitution: entry
// j + 1 needs row j + 1 of U up to column j, final right after pivot j) and, if COLP, `cp_out` receives cp U (entry j needs
// column j of U, final at the same moment) -- so both are complete one multiply-add after the last pivot.
// INV: also form ui; without it the callers substitute with U itself (row_solve1), which is what the capacity-16 / 32 kernels do:
// 3.5 % faster (profiles/r02_kernel_variants.md).  The capacity-64 kernel keeps the explicit inverse rows: released from rest under
// 1.5 m/s commands at N = 64 the substitution form left the dynamics rows at 1.7e-8 against the certificate's 1e-8 bar.
template <bool COLP, bool INV>
__device__ __forceinline__ bool ldl6_regs(double (&a)[21], double (&dinv)[6], double (&ui)[21], double (&fs)[6], const double (&cp)[6],
                                          double (&cp_out)[6]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const double d = a[RIC_TI(j, j)];
        // positive, finite and neither tiny nor huge, read off the exponent field with one integer compare (2^-996 <= d < 2^996;
        // negative values and NaNs wrap past the span) instead of two FP64 compares on the pivot chain's pipe
        ok = ok && ((unsigned)(__double2hiint(d) - 0x01B00000) < (unsigned)(0x7E300000 - 0x01B00000));
        double sq = 0.0;
        if (j + 1 < 6) sq = a[RIC_TI(j + 1, j)] * a[RIC_TI(j + 1, j)];
        const double r = rcp_fast(d);
        dinv[j] = r;
        if (j + 1 < 6) a[RIC_TI(j + 1, j + 1)] = fma(-sq, r, a[RIC_TI(j + 1, j + 1)]);       // the pivot chain
        double t[6];                                                                      // unscaled column j
#pragma unroll
        for (int i = j + 1; i < 6; ++i) { t[i] = a[RIC_TI(i, j)]; a[RIC_TI(i, j)] = t[i] * r; }
#pragma unroll
        for (int i = j + 1; i < 6; ++i)
#pragma unroll
            for (int c = j + 1; c <= i; ++c)
                if (!(i == j + 1 && c == j + 1)) a[RIC_TI(i, c)] = fma(-a[RIC_TI(i, j)], t[c], a[RIC_TI(i, c)]);
        // row j of inv(U):  ui[j][c] = -(u[j][c] + sum_{c < k < j} u[j][k] ui[k][c])
        if constexpr (INV) {
#pragma unroll
            for (int c = 0; c < j; ++c) {
                double acc = a[RIC_TI(j, c)];
#pragma unroll
                for (int k = c + 1; k < j; ++k) acc = fma(a[RIC_TI(j, k)], ui[RIC_TI(k, c)], acc);
                ui[RIC_TI(j, c)] = -acc;
            }
        }
        // the carried rows
        if (j + 1 < 6) {
            double acc = fs[j + 1];
#pragma unroll
            for (int r = 0; r <= j; ++r) acc = fma(-fs[r], a[RIC_TI(j + 1, r)], acc);
            fs[j + 1] = acc;
        }
        if (COLP) {
            double acc = cp[j];
#pragma unroll
            for (int r = j + 1; r < 6; ++r) acc = fma(cp[r], a[RIC_TI(r, j)], acc);
            cp_out[j] = acc;
        }
    }
    return ok;
}

// out = in * m   (m UNIT lower triangular, packed; its diagonal entries are not read):  out[c] = in[c] + sum_{r > c} in[r] m[r][c]
__device__ __forceinline__ void row_mul1(double (&out)[6], const double (&in)[6], const double (&m)[21]) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        double t = in[c];
#pragma unroll
        for (int r = c + 1; r < 6; ++r) t = fma(in[r], m[RIC_TI(r, c)], t);
        out[c] = t;
    }
}
// out = in * inv(m)  (m UNIT lower triangular, packed), by substitution from the last column back:  out[c] = in[c] - sum_{r > c} out[r] m[r][c]
__device__ __forceinline__ void row_solve1(double (&out)[6], const double (&in)[6], const double (&m)[21]) {
    double t[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) t[c] = in[c];
#pragma unroll
    for (int r = 5; r >= 1; --r) {
        out[r] = t[r];
#pragma unroll
        for (int c = 0; c < r; ++c) t[c] = fma(-out[r], m[RIC_TI(r, c)], t[c]);
    }
    out[0] = t[0];
}
// out = in * m'  (m UNIT lower triangular, packed):  out[c] = in[c] + sum_{r < c} in[r] m[c][r]
__device__ __forceinline__ void row_mul1T(double (&out)[6], const double (&in)[6], const double (&m)[21]) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        double t = in[c];
#pragma unroll
        for (int r = 0; r < c; ++r) t = fma(in[r], m[RIC_TI(c, r)], t);
        out[c] = t;
    }
}
__device__ __forceinline__ void load_row6(double (&out)[6], const double* p) {       // p 16-byte aligned
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2),
                  c = *reinterpret_cast<const double2*>(p + 4);
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y; out[4] = c.x; out[5] = c.y;
}
__device__ __forceinline__ void store_row6(double* p, const double (&v)[6]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
    *reinterpret_cast<double2*>(p + 4) = make_double2(v[4], v[5]);
}
// the 3x3 angular block of foot-step t: dt inv(R gI) [r]x from the stored lever arm and inertia block
template <int N>
__device__ __forceinline__ void foot_A(const DevParams& P, const RicInst<N>& sm, int t, double A[9]) {
    constexpr int NF = 4 * N;
    const double r[3] = {sm.lev[t], sm.lev[NF + t], sm.lev[2 * NF + t]};
    double Ii[9];
    if constexpr (RicInst<N>::COMPACT_II) {
        const double cs = sm.Ii[2 * (t >> 2)], sn = sm.Ii[2 * (t >> 2) + 1];
#pragma unroll
        for (int a = 0; a < 3; ++a) {                          // step_inertia's arithmetic (mpcqp_foot.cuh)
            Ii[3 * a + 0] = P.gIinv[3 * a + 0] * cs - P.gIinv[3 * a + 1] * sn;
            Ii[3 * a + 1] = P.gIinv[3 * a + 0] * sn + P.gIinv[3 * a + 1] * cs;
            Ii[3 * a + 2] = P.gIinv[3 * a + 2];
        }
    } else {
        const double* src = sm.Ii + 9 * (t >> 2);
#pragma unroll
        for (int i = 0; i < 9; ++i) Ii[i] = src[i];
    }
    lever_block(P, Ii, r, A);
}
// The LQ solve itself: backward recursion over sm.E (packed E_k) and sm.beta, forward pass, velocity costates.
// Returns (uniform over the half-warp) false if a pivot was not positive.  On return sm.xst holds the states
// x_1..x_N and sm.lam the velocity costates lam^v_1..lam^v_N; sm.E is dead.
template <int N>
__device__ __forceinline__ bool ric_core(const DevParams& P, RicInst<N>& sm, double* __restrict__ ws, int sub, int hl, const int n) {
    const double dt = P.dt;
    const int ld = n + 1;                                        // leading dimension of xref
    constexpr bool EXPLICIT_INV = N >= 64;                       // see ldl6_regs
    RPROF_T0();
    // ---- terminal cost-to-go: P_N = Q, p_N = -Q xref_N
    {
        RicCost& c0 = sm.cost[(n - 1) & 1];
        for (int i = hl; i < 36; i += 16) {
            const int rr = i / 6, dg = (i - 6 * rr) == rr;
            c0.Ppp[i] = dg ? P.wp[rr] : 0.0;
            c0.Pvv[i] = dg ? P.wv[rr] : 0.0;
            c0.Ppv[i] = 0.0;
        }
        if (hl < 6) {
            c0.pp[hl] = -P.wp[hl] * sm.xr[hl * ld + n];
            c0.pv[hl] = -P.wv[hl] * sm.xr[(6 + hl) * ld + n];
        }
    }
    __syncwarp();
    RPROF(1);

    // ---- lane roles in a stage: rid 0..5 rows of Ppv, 6..11 rows of Pvv, 12 pv' (13..15 repeat 12);
    //      lanes 0..5 additionally own row hl of T = E_k L
    const int rid = (hl < 13) ? hl : 12;
    const bool prow = rid < 6, vrow = rid >= 6 && rid < 12;
    const int ri = prow ? rid : (vrow ? rid - 6 : 0);          // row index inside its 6x6 block
    int ie[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) ie[q] = (ri >= q) ? ri * (ri + 1) / 2 + q : q * (q + 1) / 2 + ri;
    const double* pe[6];                                       // this lane's row of the packed E_k, walked from stage to stage
#pragma unroll
    for (int q = 0; q < 6; ++q) pe[q] = sm.E + 21 * (n - 1) + ie[q];
    // this lane's row of Q = diag(wp, wv): six registers pairs held across the recursion, so that adding the diagonal weight costs one
    // add per element (the ternary form costs a compare and two selects more, per element and stage, and holds all twelve weights)
    const double wq = prow ? P.wp[ri] : P.wv[ri];
    double dq[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) dq[c] = (c == ri) ? wq : 0.0;

    const double* xrp = sm.xr + (prow ? ri : 6 + ri) * ld;
    bool spd = true;
    for (int k = n - 1; k >= 0; --k) {
        const RicCost& cin = sm.cost[k & 1];
        RicCost& cout = sm.cost[(k & 1) ^ 1];
        const double xrk = xrp[k];                               // this row's reference at stage k, fetched here: it is needed at the very end of the stage
        // (1) U D U' = Pvv, Ui = inv(U): every lane, in registers (square-root free: L = U D^1/2 never appears)
        double L[21], Li[21], dinv[6];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) L[RIC_TI(r, c)] = cin.Pvv[r * 6 + c];
        double rho[6], er[6];
        load_row6(rho, prow ? cin.Ppv + 6 * ri : (vrow ? cin.Pvv + 6 * ri : cin.pv));
#pragma unroll
        for (int q = 0; q < 6; ++q) { er[q] = *pe[q]; pe[q] -= 21; }
        // ... carrying along  y = rho U^-T  (so that rho L^-T = y D^-1/2)  and row ri of T = E U
        double tr[6], y[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) y[q] = rho[q];
        spd = ldl6_regs<true, EXPLICIT_INV>(L, dinv, Li, y, er, tr) && spd;
        RPROF(2);
        // (2) the rows of T -> shared memory
        if (prow) store_row6(sm.T + 6 * ri, tr);
        __syncwarp();
        RPROF(3);
        // (3) H = D^-1 + U' T  (every lane; G = I + L'EL = D^1/2 H D^1/2),  H = W Delta W',  Wi = inv(W)
        double G[21], Mi[21], einv[6];
        {
            double Tl[21];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) Tl[RIC_TI(r, c)] = sm.T[r * 6 + c];
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = 0; c <= a; ++c) {
                    double t = Tl[RIC_TI(a, c)];                                     // r = a term: U[a][a] = 1
                    if (a == c) t += dinv[a];                                        // (no "+ 0.0" off the diagonal: it would be issued)
#pragma unroll
                    for (int r = a + 1; r < 6; ++r) t = fma(L[RIC_TI(r, a)], Tl[RIC_TI(r, c)], t);
                    G[RIC_TI(a, c)] = t;
                }
        }
        RPROF(4);
        // ... carrying along  b1 = (y D^-1) W^-T
        double b1[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) b1[q] = y[q] * dinv[q];
        spd = ldl6_regs<false, EXPLICIT_INV>(G, einv, Mi, b1, b1, b1) && spd;
        RPROF(5);
        // (4) with e = ((y D^-1) W^-T Delta^-1) W^-1:   t = e U' (row of Pt[:, v]),   kr = ((y - e) D^-1) U^-1 (row of
        //     [Ppv; Pvv; pv'] Gamma)
        double v[6], kr[6];
        {
            double dr[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) b1[q] *= einv[q];
            if constexpr (EXPLICIT_INV) row_mul1(v, b1, Mi); else row_solve1(v, b1, G);
#pragma unroll
            for (int q = 0; q < 6; ++q) dr[q] = (y[q] - v[q]) * dinv[q];
            row_mul1T(tr, v, L);
            if constexpr (EXPLICIT_INV) row_mul1(kr, dr, Li); else row_solve1(kr, dr, L);
        }
        if (hl < 13) {
            double* g = ws + (size_t)RIC_GAIN * k + rid;                // gain of impulse component o: coefficient rid
#pragma unroll
            for (int o = 0; o < 6; ++o) g[14 * o] = kr[o];
        }
        RPROF(6);
        if (k > 0) {
            // (5) wr = row of Pt[:, p] = P+[:, p] row - kr Pvp;  wvv = component of pt
            double wr[6], wvv;
            {
                const double* base = prow ? cin.Ppp + 6 * ri : cin.Ppv + ri;
                const int stride = prow ? 1 : 6;
                double pvv[6];
                load_row6(pvv, cin.pv);
                wvv = prow ? cin.pp[ri] : cin.pv[ri];
#pragma unroll
                for (int q = 0; q < 6; ++q) wvv = fma(-kr[q], pvv[q], wvv);
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    double pr[6];
                    load_row6(pr, cin.Ppv + 6 * c);
                    double t = base[c * stride];
#pragma unroll
                    for (int q = 0; q < 6; ++q) t = fma(-kr[q], pr[q], t);
                    wr[c] = t;
                }
            }
            double bk[6];
            load_row6(bk, sm.beta + 6 * k);
            double hb = wvv;                                   // Pt[:, v] beta + pt, this row's component
#pragma unroll
            for (int q = 0; q < 6; ++q) hb = fma(tr[q], bk[q], hb);
            RPROF(7);
            // (6) rows of P_k: lanes 0..5 write Ppp, Ppv, pp; then lanes 6..11 write Pvv, pv
            if (prow) {
                double npp[6], npv[6];
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    npp[c] = wr[c] + dq[c];
                    npv[c] = fma(dt, wr[c], tr[c]);
                }
                store_row6(cout.Ppp + 6 * ri, npp);
                store_row6(cout.Ppv + 6 * ri, npv);
                sm.hp[ri] = hb;
                cout.pp[ri] = hb - wq * xrk;
            }
            __syncwarp();
            if (vrow) {
                double npv[6], nvv[6];
                load_row6(npv, cout.Ppv + 6 * ri);
#pragma unroll
                for (int c = 0; c < 6; ++c) nvv[c] = fma(dt, npv[c] + wr[c], tr[c]) + dq[c];
                store_row6(cout.Pvv + 6 * ri, nvv);
                cout.pv[ri] = fma(dt, sm.hp[ri], hb) - wq * xrk;
            }
            __syncwarp();
            RPROF(8);
        }
    }
    __syncwarp();                                            // the gains in the workspace are visible to the half-warp
    const bool spd_all = half_all(spd, sub);                   // uniform over the half-warp; a robot whose pivots failed
                                                               // runs on with harmless garbage so that the warp stays converged

    // ---- forward pass: x replicated in the registers of every lane; lane o (mod 6) forms impulse component o from
    //      its 13 gain coefficients, fetched RIC_DEPTH stages ahead from the (L2-resident) workspace
    {
        double x[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) x[c] = sm.xr[c * ld];
        const int o = hl % 6;
        const double* gcol = ws + 14 * o;
        double2 g[RIC_DEPTH][7];
#pragma unroll
        for (int d = 0; d < RIC_DEPTH; ++d)
#pragma unroll
            for (int i = 0; i < 7; ++i) g[d][i] = __ldcg(reinterpret_cast<const double2*>(gcol + (size_t)RIC_GAIN * d) + i);
        for (int k0 = 0; k0 < n; k0 += RIC_DEPTH) {
#pragma unroll
            for (int d = 0; d < RIC_DEPTH; ++d) {
                const int k = k0 + d;
                if (k >= n) break;                               // uniform over the warp (n is)
                double bk[6], z[12];
                load_row6(bk, sm.beta + 6 * k);
#pragma unroll
                for (int c = 0; c < 6; ++c) { z[c] = fma(dt, x[6 + c], x[c]); z[6 + c] = x[6 + c] + bk[c]; }
                double a0 = g[d][6].x, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int i = 0; i < 6; i += 2) {
                    a0 = fma(g[d][i].x, z[2 * i], a0); a1 = fma(g[d][i].y, z[2 * i + 1], a1);
                    a2 = fma(g[d][i + 1].x, z[2 * i + 2], a2); a3 = fma(g[d][i + 1].y, z[2 * i + 3], a3);
                }
                const double w = (a0 + a1) + (a2 + a3);          // minus the impulse correction of component o
                if (k + RIC_DEPTH < n) {
#pragma unroll
                    for (int i = 0; i < 7; ++i) g[d][i] = __ldcg(reinterpret_cast<const double2*>(gcol + (size_t)RIC_GAIN * (k + RIC_DEPTH)) + i);
                }
#pragma unroll
                for (int c = 0; c < 6; ++c) { x[c] = z[c]; x[6 + c] = z[6 + c] - hshfl_d(w, c); }
                if (hl == 0) {
                    double* xs = sm.xst + 12 * k;
#pragma unroll
                    for (int c = 0; c < 12; c += 2) *reinterpret_cast<double2*>(xs + c) = make_double2(x[c], x[c + 1]);
                }
            }
        }
    }
    __syncwarp();
    RPROF(9);
    // ---- costates of the velocities: lam_s = Q e_s + A' lam_{s+1}  ->  lam[6 (s-1) ..]
    if (hl < 6) {
        const int c = hl;
        const double wpc = P.wp[c], wvc = P.wv[c];
        double lp = 0.0, lv = 0.0;
#pragma unroll 8
        for (int s = n; s >= 1; --s) {
            const double* xs = sm.xst + 12 * (s - 1);
            const double ep = xs[c] - sm.xr[c * ld + s], ev = xs[6 + c] - sm.xr[(6 + c) * ld + s];
            lv = fma(wvc, ev, fma(dt, lp, lv));
            lp = fma(wpc, ep, lp);
            sm.lam[6 * (s - 1) + c] = lv;
        }
    }
    __syncwarp();
    RPROF(10);
    return spd_all;
}

// E_k = sum_j Bv S Bv' (packed lower triangle) and beta_k = g + sum_j Bv pf for every step, where foot-step t contributes the
// symmetric 3 x 3 S_t (inverse of its force weight on the directions it may move in: Z R^-1 Z' of a face, the full inverse of
// w_f I + C' D C of an interior-point iterate) and the offset pf_t.  One lane per step (lane hl owns steps hl, hl + 16, ...),
// the four feet of a step in sequence, no shuffles: with Bv = [lin I; A],  E = [[lin^2 S, lin (A S)'], [lin A S, A S A']].
// foot_of(t, S, pf) describes foot-step t (S as S00, S10, S11, S20, S21, S22); it must not depend on which lane asks.
template <int N, class FootFn>
__device__ __forceinline__ void ric_assemble(const DevParams& P, RicInst<N>& sm, int hl, const int n, FootFn foot_of) {
    const double lin = P.dt / P.mass, lin2 = lin * lin;
#pragma unroll 1
    for (int k = hl; k < n; k += 16) {
        double e[21], ub[6];
#pragma unroll
        for (int i = 0; i < 21; ++i) e[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) ub[i] = 0.0;
        ub[2] = -P.gravity * P.dt;                             // g: only the z velocity, MPC.py:200-201
        // stance feet only: a swing foot contributes S = 0 and pf = 0, i.e. exactly nothing (sm.sigb holds contact << 7 per foot-step)
        unsigned cm = *reinterpret_cast<const unsigned*>(sm.sigb + 4 * k) & 0x80808080u;
#pragma unroll 1
        for (; cm != 0u; cm &= cm - 1u) {
            const int t = 4 * k + ((__ffs(cm) - 1) >> 3);
            double S[6], pf[3];
            foot_of(t, S, pf);
            double A[9];
            foot_A<N>(P, sm, t, A);
            double M[9];                                       // M = A S
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                M[3 * q + 0] = A[3 * q] * S[0] + A[3 * q + 1] * S[1] + A[3 * q + 2] * S[3];
                M[3 * q + 1] = A[3 * q] * S[1] + A[3 * q + 1] * S[2] + A[3 * q + 2] * S[4];
                M[3 * q + 2] = A[3 * q] * S[3] + A[3 * q + 1] * S[4] + A[3 * q + 2] * S[5];
            }
            e[RIC_TI(0, 0)] = fma(lin2, S[0], e[RIC_TI(0, 0)]);
            e[RIC_TI(1, 0)] = fma(lin2, S[1], e[RIC_TI(1, 0)]);
            e[RIC_TI(1, 1)] = fma(lin2, S[2], e[RIC_TI(1, 1)]);
            e[RIC_TI(2, 0)] = fma(lin2, S[3], e[RIC_TI(2, 0)]);
            e[RIC_TI(2, 1)] = fma(lin2, S[4], e[RIC_TI(2, 1)]);
            e[RIC_TI(2, 2)] = fma(lin2, S[5], e[RIC_TI(2, 2)]);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
#pragma unroll
                for (int c = 0; c < 3; ++c) e[RIC_TI(3 + q, c)] = fma(lin, M[3 * q + c], e[RIC_TI(3 + q, c)]);
#pragma unroll
                for (int c = 0; c <= q; ++c)
                    e[RIC_TI(3 + q, 3 + c)] += M[3 * q] * A[3 * c] + M[3 * q + 1] * A[3 * c + 1] + M[3 * q + 2] * A[3 * c + 2];
            }
            double u6[6];
            bv_apply(A, lin, pf, u6);
#pragma unroll
            for (int i = 0; i < 6; ++i) ub[i] += u6[i];
        }
#pragma unroll
        for (int i = 0; i < 21; ++i) sm.E[21 * k + i] = e[i];
        store_row6(sm.beta + 6 * k, ub);
    }
}
// S = Z R^-1 Z' of a face (closed form)
__device__ __forceinline__ void face_S(const Face& fc, double (&S)[6]) {
    S[0] = fc.dx + fc.dz * fc.czx * fc.czx; S[1] = fc.dz * fc.czx * fc.czy; S[2] = fc.dy + fc.dz * fc.czy * fc.czy;
    S[3] = fc.dz * fc.czx; S[4] = fc.dz * fc.czy; S[5] = fc.dz;
}

// The signatures of a lane's foot-steps (one per round), four to a 32-bit word: rolled loops index them with a run-time round, and an
// array of bytes indexed that way lives in local memory (ncu: 3 % of a launch's samples waited on those loads); at capacity 16 this
// is one register.
template <int R>
struct SigVec {
    unsigned w[(R + 3) / 4] = {};
    __device__ __forceinline__ uint8_t operator[](int r) const { return (uint8_t)((w[r >> 2] >> (8 * (r & 3))) & 0xffu); }
    __device__ __forceinline__ void set(int r, uint8_t v) {
        const int sh = 8 * (r & 3);
        w[r >> 2] = (w[r >> 2] & ~(0xffu << sh)) | ((unsigned)v << sh);
    }
};

constexpr uint8_t SIG_PIN = 27;    // not a face: the foot-step's force is held at a given value (roll-out of a feasible iterate)

// One equality-constrained solve on the faces `sg` selects + KKT guard.  The 16 lanes of the robot call it.
// Returns (uniform over the half-warp) 1 if every foot passes the guard, 0 if not, -1 if a pivot was not positive.
// Both halves of the warp must call it together (full-mask collectives inside).
// On return sm.E holds the forces (3 per foot, foot-major), sm.xst the states, sm.lam the velocity costates.
// A foot-step whose signature is SIG_PIN is held at the force pin[c * NF + t] (c = 0..2) and skips the guard: with every
// stance foot pinned the sweep is the forward roll-out of those forces (E_k = 0), which is how an iterate that was not
// certified still leaves with states that belong to its forces.
template <int N>
__device__ int ric_sweep(const DevParams& P, RicInst<N>& sm, double* __restrict__ ws, int sub, int hl, const int n, unsigned conbits,
                         const SigVec<RicInst<N>::ROUNDS>& sg, SigVec<RicInst<N>::ROUNDS>& nsg, const double* __restrict__ pin = nullptr) {
    using S = RicInst<N>;
    constexpr int ROUNDS = S::ROUNDS, NF = S::NF;
    const double lin = P.dt / P.mass;
    RPROF_T0();
    RPROF_COUNT(0);

#pragma unroll 1
    for (int r = 0; r < ROUNDS; ++r) sm.sigb[hl + 16 * r] = (uint8_t)(sg[r] | (((conbits >> r) & 1u) << 7));
    __syncwarp();
    ric_assemble<N>(P, sm, hl, n, [&](int t, double (&S6)[6], double (&pf)[3]) {
        const uint8_t sb = sm.sigb[t];
        Face fc;
        const bool pinned = (sb & 127) == SIG_PIN;
        make_face(P, (sb >> 7) != 0 && !pinned, sb & 127, fc);
        face_S(fc, S6);
        pf[0] = fc.pf[0]; pf[1] = fc.pf[1]; pf[2] = fc.pf[2];
        if (pinned && (sb >> 7) && pin != nullptr) { pf[0] = pin[t]; pf[1] = pin[NF + t]; pf[2] = pin[2 * NF + t]; }
    });
    __syncwarp();
    RPROF(1);
    const bool spd_all = ric_core<N>(P, sm, ws, sub, hl, n);
    // ---- per foot: forces on the face, gradient, KKT guard
    bool ok = true;
#pragma unroll 1
    for (int r = 0; r < ROUNDS; ++r) {
        const int t = hl + 16 * r, k = t >> 2;
        const bool contact = (conbits >> r) & 1u;
        double f[3] = {0.0, 0.0, 0.0};
        uint8_t ns = sg[r];
        if (contact && sg[r] == SIG_PIN) {
            if (pin != nullptr) { f[0] = pin[t]; f[1] = pin[NF + t]; f[2] = pin[2 * NF + t]; }
        } else if (contact) {
            Face fc;
            make_face(P, true, sg[r], fc);
            double A[9], h[3];
            foot_A<N>(P, sm, t, A);
            bvT_apply(A, lin, sm.lam + 6 * k, h);
            const double qx = fc.zx ? -fc.dx * h[0] : 0.0;
            const double qy = fc.zy ? -fc.dy * h[1] : 0.0;
            const double qz = fc.zz ? -fc.dz * (fc.czx * h[0] + fc.czy * h[1] + h[2]) : 0.0;
            f[0] = fc.pf[0] + qx + fc.czx * qz;
            f[1] = fc.pf[1] + qy + fc.czy * qz;
            f[2] = fc.pf[2] + qz;
            const double grad[3] = {fma(P.w_force, f[0], h[0]), fma(P.w_force, f[1], h[1]), fma(P.w_force, f[2], h[2])};
            FootSol sol;
            ok = kkt_guard(P, sg[r], f, grad, sol, ns) && ok;
        }
        nsg.set(r, ns);
        sm.E[3 * t] = f[0]; sm.E[3 * t + 1] = f[1]; sm.E[3 * t + 2] = f[2];
    }
    const bool all_ok = half_all(ok, sub);
    RPROF(11);
    return !spd_all ? -1 : (all_ok ? 1 : 0);
}

// The active-set iteration: up to max_s sweeps from the signatures `sg`, every failed sweep followed by the primal-dual
// update (every foot adopts the guard's proposal).  A proposal that was tried before would cycle; then, if `careful`, one
// foot changes per sweep in index order (the stage's own safeguard), otherwise the iteration stops (the interior-point
// stage takes over).  `want` says whether this half-warp takes part; both halves call it together.
// done / status / sweeps are updated for the halves that take part; on return with done the results of the accepted
// sweep sit in shared memory ONLY IF the other half did not go on sweeping -- callers that need them re-run one sweep.
template <int N>
__device__ __forceinline__ void ric_active_set(const DevParams& P, RicInst<N>& sm, double* __restrict__ ws, int sub, int hl, const int n,
                                               unsigned conbits, SigVec<RicInst<N>::ROUNDS>& sg, SigVec<RicInst<N>::ROUNDS>& nsg,
                                               bool want, int max_s, bool allow_careful, int& sweeps, bool& done, int& status) {
    constexpr int ROUNDS = RicInst<N>::ROUNDS;
    // order-sensitive hash of a signature, uniform over the half-warp (cycle detection)
    auto sig_hash = [&](const SigVec<ROUNDS>& g) {
        unsigned long long h = 0ull;
#pragma unroll 1
        for (int r = 0; r < ROUNDS; ++r) {
            if ((conbits >> r) & 1u) {
                unsigned long long q = (unsigned long long)(g[r] + 1) * 0x9E3779B97F4A7C15ull;
                q ^= q >> 29; q *= (2ull * (hl + 16 * r) + 0xBF58476D1CE4E5B9ull); q ^= q >> 32;
                h += q;
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) h += __shfl_xor_sync(RIC_FULL, h, o, 16);
        return h;
    };
    int nhist = 0;
    bool careful = false;       // set once the full primal-dual update proposed a signature that was tried before
    bool stop = !want;
    for (int s = 0; s < max_s; ++s) {
        const bool need = !done && !stop;
        if (!__any_sync(RIC_FULL, need)) break;
        const int rc = ric_sweep<N>(P, sm, ws, sub, hl, n, conbits, sg, nsg);
        bool search = false;
        if (need) {
            ++sweeps;
            if (rc < 0) stop = true;
            else if (rc > 0) { done = true; status = 1; }
            else search = true;
        }
        if (!__any_sync(RIC_FULL, search)) continue;      // the common case: nothing to hash, nothing to choose
        // ---- next signature.  The one just tried goes into the history; first choice: every foot adopts its proposal
        // (primal-dual active-set step).
        const unsigned long long h = sig_hash(sg);
        __syncwarp();
        if (search && hl == 0 && nhist < 16) sm.hist[nhist] = h;
        if (search) nhist = (nhist < 16) ? nhist + 1 : nhist;
        __syncwarp();
        auto seen = [&](unsigned long long q) {
            bool f = false;
            for (int i = 0; i < nhist; ++i) f = f || (sm.hist[i] == q);
            return f;
        };
        const unsigned long long hfull = sig_hash(nsg);
        if (search && !careful) {
            if (seen(hfull)) {
                careful = true;
                if (!allow_careful) { stop = true; search = false; }
            } else {
                sg = nsg;
                search = false;
            }
        }
        int tlast = -1;
        while (__any_sync(RIC_FULL, search)) {
            int tm = 0x7fffffff;
#pragma unroll 1
            for (int r = 0; r < ROUNDS; ++r) {
                const int t = hl + 16 * r;
                if (nsg[r] != sg[r] && t > tlast && t < tm) tm = t;
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) { const int q = __shfl_xor_sync(RIC_FULL, tm, o, 16); tm = q < tm ? q : tm; }
            SigVec<ROUNDS> cand;
#pragma unroll 1
            for (int r = 0; r < ROUNDS; ++r) cand.set(r, (hl + 16 * r == tm) ? nsg[r] : sg[r]);
            const unsigned long long hc = sig_hash(cand);
            if (search) {
                if (tm == 0x7fffffff) { stop = true; search = false; }            // every single change was tried before
                else if (!seen(hc)) {
                    sg = cand;
                    search = false;
                } else tlast = tm;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Interior-point stage.  Per stance foot-step the six rows  C6 f <= h6  (the five pyramid rows of MPC.py:136-148 with the
// two-sided fz row split in two):  fx - mu fz, -fx - mu fz, fy - mu fz, -fy - mu fz, -fz <= 0,  fz <= fz_max.
// State per foot-step in the workspace (struct of arrays, stride NF): f (3), y (6); scratch: target fp (3), dy (6).
// One iteration, written for the NEW iterate instead of the step (so that the linear terms of the tracking problem stay
// where ric_core expects them):  with s = h6 - C6 f, D = y / s,
//     (H + C6' D C6) f+ = -g - C6' (y - D h6 + sigma mu / s)
// is the stage-wise LQ problem with per-foot force weight R = w_f I + C6' D C6 and linear term c = C6' (...):  S = R^-1,
// pf = -S c, f+ = pf - S Bv' lam.  Then ds = -C6 (f+ - f), dy = sigma mu / s - y - D ds, separate primal and dual step
// lengths to the boundary (0.995), sigma chosen from the length of the previous step.
// ---------------------------------------------------------------------------------------------------------------------
struct IpmFoot {
    double S[6], c[3], s[6], D[6];
};
__device__ __forceinline__ void ipm_foot(const DevParams& P, const double (&f)[3], const double (&y)[6], double sigmu, IpmFoot& o) {
    const double mu = P.mu, w = P.w_force;
    o.s[0] = mu * f[2] - f[0]; o.s[1] = mu * f[2] + f[0]; o.s[2] = mu * f[2] - f[1]; o.s[3] = mu * f[2] + f[1];
    o.s[4] = f[2]; o.s[5] = P.fz_max - f[2];
    double v[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double is = 1.0 / o.s[i];
        o.D[i] = y[i] * is;
        v[i] = fma(sigmu, is, y[i]);
    }
    v[5] -= o.D[5] * P.fz_max;
    const double a = w + o.D[0] + o.D[1], b = w + o.D[2] + o.D[3];
    const double p = mu * (o.D[1] - o.D[0]), q = mu * (o.D[3] - o.D[2]);
    const double ia = 1.0 / a, ib = 1.0 / b;
    // Schur complement of the fz row, cancellation free:  mu^2 [(D0 + D1) - (D1 - D0)^2 / a] = mu^2 [w (D0 + D1) + 4 D0 D1] / a
    const double cz = w + o.D[4] + o.D[5] + mu * mu * ((w * (o.D[0] + o.D[1]) + 4.0 * o.D[0] * o.D[1]) * ia +
                                                        (w * (o.D[2] + o.D[3]) + 4.0 * o.D[2] * o.D[3]) * ib);
    const double iz = 1.0 / cz, pa = p * ia, qb = q * ib;
    o.S[5] = iz; o.S[3] = -pa * iz; o.S[4] = -qb * iz;
    o.S[0] = fma(pa * pa, iz, ia); o.S[2] = fma(qb * qb, iz, ib); o.S[1] = pa * qb * iz;
    o.c[0] = v[0] - v[1]; o.c[1] = v[2] - v[3]; o.c[2] = -mu * ((v[0] + v[1]) + (v[2] + v[3])) - v[4] + v[5];
}
__device__ __forceinline__ void sym3_apply(const double (&S)[6], const double (&x)[3], double (&o)[3]) {
    o[0] = S[0] * x[0] + S[1] * x[1] + S[3] * x[2];
    o[1] = S[1] * x[0] + S[2] * x[1] + S[4] * x[2];
    o[2] = S[3] * x[0] + S[4] * x[1] + S[5] * x[2];
}

// Interior-point iterations of one robot until the complementarity gap mu_gap = y's / rows falls below `target` or `budget`
// iterations are spent.  `gap` / `sigma` carry over between calls.  Returns false on a numerical breakdown.
template <int N>
__device__ bool ric_ipm(const DevParams& P, RicInst<N>& sm, double* __restrict__ ws, double* __restrict__ adm, int sub, int hl, const int n,
                        unsigned conbits, bool want, double target, int budget, double& gap, double& sigma, int& iters) {
    constexpr int NF = RicInst<N>::NF, ROUNDS = RicInst<N>::ROUNDS;
    const double lin = P.dt / P.mass, tau = 0.995;
    int rows = 0;
#pragma unroll 1
    for (int r = 0; r < ROUNDS; ++r) rows += ((conbits >> r) & 1u) ? 6 : 0;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) rows += __shfl_xor_sync(RIC_FULL, rows, o, 16);
    const double inv_rows = rows > 0 ? 1.0 / (double)rows : 0.0;
    bool ok = true;
    bool live = want && rows > 0;
    for (int it = 0; it < budget; ++it) {
        const bool need = live && ok && gap > target;
        if (!__any_sync(RIC_FULL, need)) break;
        const double sigmu = sigma * gap;
#pragma unroll 1
        for (int r = 0; r < ROUNDS; ++r) sm.sigb[hl + 16 * r] = (uint8_t)(((conbits >> r) & 1u) << 7);
        __syncwarp();
        ric_assemble<N>(P, sm, hl, n, [&](int t, double (&S6)[6], double (&pf)[3]) {
#pragma unroll
            for (int i = 0; i < 6; ++i) S6[i] = 0.0;
            pf[0] = pf[1] = pf[2] = 0.0;
            if (sm.sigb[t] >> 7) {
                double f[3], y[6];
#pragma unroll
                for (int c = 0; c < 3; ++c) f[c] = adm[c * NF + t];
#pragma unroll
                for (int q = 0; q < 6; ++q) y[q] = adm[(3 + q) * NF + t];
                IpmFoot ft;
                ipm_foot(P, f, y, sigmu, ft);
#pragma unroll
                for (int i = 0; i < 6; ++i) S6[i] = ft.S[i];
                double sc[3];
                sym3_apply(ft.S, ft.c, sc);
                pf[0] = -sc[0]; pf[1] = -sc[1]; pf[2] = -sc[2];
            }
        });
        __syncwarp();
        const bool spd = ric_core<N>(P, sm, ws, sub, hl, n);
        // ---- per foot: target, steps, lengths to the boundary
        double ap = 1.0, ad = 1.0;
#pragma unroll 1
        for (int r = 0; r < ROUNDS; ++r) {
            const int t = hl + 16 * r, k = t >> 2;
            if ((conbits >> r) & 1u) {
                double f[3], y[6], A[9], h[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) f[c] = adm[c * NF + t];
#pragma unroll
                for (int q = 0; q < 6; ++q) y[q] = adm[(3 + q) * NF + t];
                IpmFoot ft;
                ipm_foot(P, f, y, sigmu, ft);
                foot_A<N>(P, sm, t, A);
                bvT_apply(A, lin, sm.lam + 6 * k, h);
                const double ch[3] = {ft.c[0] + h[0], ft.c[1] + h[1], ft.c[2] + h[2]};
                double fp[3];
                sym3_apply(ft.S, ch, fp);
                const double df[3] = {-fp[0] - f[0], -fp[1] - f[1], -fp[2] - f[2]};
                const double mz = P.mu * df[2];
                const double ds[6] = {mz - df[0], mz + df[0], mz - df[1], mz + df[1], df[2], -df[2]};
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const double dy = sigmu / ft.s[q] - y[q] - ft.D[q] * ds[q];
                    if (ds[q] < 0.0) ap = fmin(ap, -tau * ft.s[q] / ds[q]);
                    if (dy < 0.0) ad = fmin(ad, -tau * y[q] / dy);
                    adm[(12 + q) * NF + t] = dy;
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) adm[(9 + c) * NF + t] = df[c];
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            ap = fmin(ap, __shfl_xor_sync(RIC_FULL, ap, o, 16));
            ad = fmin(ad, __shfl_xor_sync(RIC_FULL, ad, o, 16));
        }
        // ---- apply, new gap
        double part = 0.0;
        const bool sane = spd && (ap > 0.0) && (ad > 0.0) && isfinite(ap) && isfinite(ad);
#pragma unroll 1
        for (int r = 0; r < ROUNDS; ++r) {
            const int t = hl + 16 * r;
            if ((conbits >> r) & 1u) {
                double f[3], y[6];
#pragma unroll
                for (int c = 0; c < 3; ++c) f[c] = adm[c * NF + t];
#pragma unroll
                for (int q = 0; q < 6; ++q) y[q] = adm[(3 + q) * NF + t];
                if (need && sane) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) { f[c] = fma(ap, adm[(9 + c) * NF + t], f[c]); adm[c * NF + t] = f[c]; }
#pragma unroll
                    for (int q = 0; q < 6; ++q) { y[q] = fma(ad, adm[(12 + q) * NF + t], y[q]); adm[(3 + q) * NF + t] = y[q]; }
                }
                const double mz = P.mu * f[2];
                const double s6[6] = {mz - f[0], mz + f[0], mz - f[1], mz + f[1], f[2], P.fz_max - f[2]};
#pragma unroll
                for (int q = 0; q < 6; ++q) part = fma(y[q], s6[q], part);
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(RIC_FULL, part, o, 16);
        if (need) {
            ++iters;
            if (!sane) ok = false;
            else {
                gap = part * inv_rows;
                const double a = fmin(ap, ad);
                sigma = a > 0.9 ? 0.05 : (a > 0.6 ? 0.15 : (a > 0.3 ? 0.3 : 0.5));
            }
        }
        __syncwarp();
    }
    return ok;
}

// Outputs of one robot (the half-warp version of finish() in mpcqp_kernels.cu)                 [MPC.py:432-458]
// The states are those of the last sweep's forward pass (sm.xst), i.e. the dynamics driven by exactly the impulses of the
// forces returned (sm.E) -- also for a robot that leaves uncertified (status MPCQP_STATUS_MAX_ITER): its last sweep was the
// roll-out of its feasible interior-point iterate (SIG_PIN), `ymax` then points at that iterate's multipliers.
template <int N>
__device__ void ric_finish(const DevParams& P, const DevScenario& SC, RicInst<N>& sm, const DevState& st, int inst, int sub, int hl, const int n,
                           unsigned conbits, const SigVec<RicInst<N>::ROUNDS>& sg, bool solved, int status, int sweeps, int iters,
                           const double* __restrict__ ymax, bool commit) {
    using S = RicInst<N>;
    constexpr int NF = S::NF, ROUNDS = S::ROUNDS, AWC = S::AW, CWC = S::CW;
    const int AW = (20 * n + 31) / 32, CW = (4 * n + 31) / 32, ld = n + 1;
    const double lin = P.dt / P.mass;
    for (int i = hl; i < AWC + CWC; i += 16) sm.amask[i] = 0u;
    // the world pose this tick advances (MPC.py:503-510): fetched now, used by lane 0 at the very end
    double qw[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (commit && hl == 0) {
#pragma unroll
        for (int c = 0; c < 6; c += 2) { const double2 v = *reinterpret_cast<const double2*>(st.qw + (size_t)inst * 6 + c); qw[c] = v.x; qw[c + 1] = v.y; }
    }
    if (!solved) {
        // no forces: the states are the free response  p_{s+1} = p_s + dt v_s, v_{s+1} = v_s + g   (MPC.py:110-111, 200-205)
        if (hl < 6) {
            const int c = hl;
            double p = sm.xr[c * ld], v = sm.xr[(6 + c) * ld];
            const double gc = (c == 2) ? -P.gravity * P.dt : 0.0;
            for (int s = 0; s < n; ++s) {
                const double pn = p + P.dt * v;
                v += gc; p = pn;
                sm.xst[12 * s + c] = p; sm.xst[12 * s + 6 + c] = v;
            }
        }
    }
    __syncwarp();
    double part = 0.0;
    double* xs = st.xs + (size_t)inst * 12 * n;
    {
        // element i = hl + 16 u + 48 m is component (hl + 16 u) % 12 of step 4 m + (hl + 16 u) / 12: the component, its weight and the
        // step offset of a lane's three residues are fixed, so the division and the weight lookup leave the loop (same order of
        // accumulation as i = hl, hl + 16, ...)
        int xo[3];
        double wu[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int jj = hl + 16 * u, ds = jj / 12, c = jj - 12 * ds;
            xo[u] = c * ld + ds + 1;
            wu[u] = 0.5 * (c < 6 ? P.wp[c] : P.wv[c - 6]);
        }
        if (commit && hl < 12) st.x1[(size_t)inst * 12 + hl] = isfinite(sm.xst[hl] - sm.xr[xo[0]]) ? sm.xst[hl] : 0.0;   // MPC.q_next / v_next (MPC.py:448-450)
        for (int i0 = 0, s0 = 0; i0 < 12 * n; i0 += 48, s0 += 4) {
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int i = i0 + hl + 16 * u;
                if (i < 12 * n) {
                    const double e = sm.xst[i] - sm.xr[xo[u] + s0];
                    const double ee = isfinite(e) ? e : 0.0;                                 // malformed input: never NaN out
                    if (commit) xs[i] = ee;                                                  // MPC.x[:12N] (MPC.py:428)
                    part = fma(wu[u] * ee, ee, part);
                }
            }
        }
    }
    if (hl < 12) sm.xnext[hl] = sm.xst[hl];                                                  // MPC.q_next / v_next (MPC.py:448-450)
#pragma unroll 1
    for (int r = 0; r < ROUNDS; ++r) {
        const int t = hl + 16 * r, k = t >> 2, j = t & 3;
        const bool exists = t < 4 * n;
        const bool contact = (conbits >> r) & 1u;
        double f[3] = {0.0, 0.0, 0.0};
        FootSol sol;
#pragma unroll
        for (int q = 0; q < 5; ++q) sol.y[q] = 0.0;
        if (solved && contact) {
            f[0] = sm.E[3 * t]; f[1] = sm.E[3 * t + 1]; f[2] = sm.E[3 * t + 2];
            if (ymax != nullptr) {
                // uncertified interior-point iterate: its own multipliers (the two fz rows share the reference's fifth row)
#pragma unroll
                for (int q = 0; q < 4; ++q) sol.y[q] = ymax[(3 + q) * NF + t];
                sol.y[4] = ymax[7 * NF + t] - ymax[8 * NF + t];
            } else {
                // multipliers: the guard's closed form on the gradient of the accepted sweep
                double A[9], h[3];
                foot_A<N>(P, sm, t, A);
                bvT_apply(A, lin, sm.lam + 6 * k, h);
                const double grad[3] = {fma(P.w_force, f[0], h[0]), fma(P.w_force, f[1], h[1]), fma(P.w_force, f[2], h[2])};
                uint8_t dummy;
                kkt_guard(P, sg[r], f, grad, sol, dummy);
            }
        }
        part += 0.5 * P.w_force * (f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        if (commit && exists) {
            double* fo = st.f + (size_t)inst * 12 * n + 3 * t;
            fo[0] = f[0]; fo[1] = f[1]; fo[2] = f[2];
            double* yo = st.y + (size_t)inst * 20 * n + 5 * t;
#pragma unroll
            for (int q = 0; q < 5; ++q) yo[q] = sol.y[q];
            st.sig[(size_t)inst * 4 * n + t] = sg[r] > 26 ? SIG_FREE : sg[r];
            if (k == 0) {
                double* f0 = st.f0 + (size_t)inst * 12 + 3 * j;
                f0[0] = f[0]; f0[1] = f[1]; f0[2] = f[2];
            }
        }
        // rows that hold with equality; a swing foot is pinned to f = 0 (MPC.py:355-358): all five of its rows do
        const double mu = P.mu, tol = 1e-9;
        const double row[5] = {f[0] - mu * f[2], -f[0] - mu * f[2], f[1] - mu * f[2], -f[1] - mu * f[2], -f[2]};
        const int b0 = 5 * t;
        if (exists) {
            unsigned m5 = 0u;                                   // the five rows of this foot-step, then at most two words to touch
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const bool act = (fabs(row[q]) <= tol) || (q == 4 && fabs(row[4] + P.fz_max) <= tol);
                m5 |= act ? (1u << q) : 0u;
            }
            const unsigned sh = b0 & 31;
            const unsigned lo = m5 << sh, hi = (sh > 27) ? (m5 >> (32 - sh)) : 0u;
            if (lo) atomicOr(&sm.amask[b0 >> 5], lo);
            if (hi) atomicOr(&sm.amask[(b0 >> 5) + 1], hi);
        }
        const unsigned cb = (__ballot_sync(RIC_FULL, contact) >> (16 * sub)) & 0xFFFFu;      // feet 16 r .. 16 r + 15
        if (hl == 0 && cb) atomicOr(&sm.amask[AWC + (r >> 1)], cb << (16 * (r & 1)));
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(RIC_FULL, part, o, 16);
    __syncwarp();
    if (!commit) return;
    for (int i = hl; i < AW; i += 16) st.active[(size_t)inst * AW + i] = sm.amask[i];
    for (int i = hl; i < CW; i += 16) st.contact[(size_t)inst * CW + i] = sm.amask[AWC + i];
    if (hl == 0) {
        st.obj[inst] = part;
        st.status[inst] = status;
        st.sweeps[inst] = sweeps;
        st.iters[inst] = iters;
        if (status != 3) {
            world_pose_step(qw, sm.xnext);
#pragma unroll
            for (int c = 0; c < 6; c += 2) *reinterpret_cast<double2*>(st.qw + (size_t)inst * 6 + c) = make_double2(qw[c], qw[c + 1]);
        }
    }
    // close the loop: the twelve components of the next measured state (and their Gaussian samples) one lane each
    if (SC.enabled && status != 3) scenario_advance_lanes(SC, inst, sm.xnext, hl);
}

// Inputs of one robot -> shared memory, then the decode: contact flags, lever arms, inertia blocks  [MPC.py:316-360, 635-652].
// Both halves of the warp call it together.  Returns (uniform over the half-warp) true if the inputs are malformed.
template <int N>
__device__ __forceinline__ bool ric_load_decode(const DevParams& P, const DevScenario& SC, RicInst<N>& sm, const double* __restrict__ xref_g,
                                                const double* __restrict__ fsteps_g, int inst, bool valid, int first_tick, int sub, int hl,
                                                const int n, unsigned int& phase, unsigned& conbits) {
    constexpr int NF = RicInst<N>::NF, ROUNDS = RicInst<N>::ROUNDS;
    if (SC.enabled == 1) {
        scenario_inputs<16>(P, SC, sm.sc, inst, sm.xr, sm.fs, n, valid);
    } else {
        // rows 0 .. fs_rows - 1 of the gait table up front; the rest only if no row among them ends the table (MPC.py:646) -- over
        // PCIe (inputs in the caller's page-locked memory) the reference's tables (2 .. 7 rows of 20) cost 832 instead of 2080 bytes
        const int head = P.fs_rows * 13;
        if (hl == 0) {
            fence_async_smem();
            mbar_expect_tx(&sm.mbar, (12 * (n + 1) + head) * 8);
            bulk_g2s(sm.xr, xref_g + (size_t)inst * 12 * (n + 1), 12 * (n + 1) * 8, &sm.mbar);
            bulk_g2s(sm.fs, fsteps_g + (size_t)inst * 260, head * 8, &sm.mbar);
        }
        mbar_wait(&sm.mbar, phase);
        phase ^= 1u;
        if (head < 260) {
            bool ends = false;                                  // uniform over the half-warp: every lane reads the same counts
            for (int r = 0; r < P.fs_rows; ++r) ends = ends || !(sm.fs[r * 13] > 0.0);     // a terminator, or a malformed count (decode stops there too)
            if (!ends) {
                if (hl == 0) {
                    mbar_expect_tx(&sm.mbar, (260 - head) * 8);
                    bulk_g2s(sm.fs + head, fsteps_g + (size_t)inst * 260 + head, (260 - head) * 8, &sm.mbar);
                }
                mbar_wait(&sm.mbar, phase);
                phase ^= 1u;
            }
        }
        __syncwarp();
    }
    RPROF_T0();
    RPROF(13);
    bool bad = false;
    conbits = 0u;
    int scan_q = 0, scan_row = -1;                             // a lane's steps grow with r: the row search goes on where it stopped
    double scan_cum = 0.0;
#pragma unroll 1
    for (int r = 0; r < ROUNDS; ++r) {
        const int t = hl + 16 * r, k = t >> 2, j = t & 3;
        double lv[3] = {0.0, 0.0, 0.0};
        bool contact = false;
        if (t < 4 * n) decode_lever_next(P, sm.xr, sm.fs, n, k, j, first_tick != 0, lv, contact, bad, scan_q, scan_cum, scan_row);
        sm.lev[t] = lv[0]; sm.lev[NF + t] = lv[1]; sm.lev[2 * NF + t] = lv[2];
        conbits |= contact ? (1u << r) : 0u;
    }
    for (int k = hl; k < n; k += 16) {
        if constexpr (RicInst<N>::COMPACT_II) {
            double sn, cs;
            sincos(sm.xr[5 * (n + 1) + k], &sn, &cs);                          // MPC.py:330
            sm.Ii[2 * k] = cs; sm.Ii[2 * k + 1] = sn;
        } else {
            double Ii[9];
            step_inertia(P, sm.xr[5 * (n + 1) + k], Ii);
#pragma unroll
            for (int i = 0; i < 9; ++i) sm.Ii[9 * k + i] = Ii[i];
        }
    }
    for (int i = hl; i < 12 * (n + 1); i += 16)               // not finite <=> exponent field all ones: an integer test on the high word
        bad = bad || ((__double2hiint(sm.xr[i]) & 0x7ff00000) == 0x7ff00000);
    const bool any_bad = half_any(bad, sub);
    __syncwarp();                                            // fs is dead from here on (E overwrites it)
    RPROF(14);
    return any_bad;
}

// `work_ctr` must be zero at launch.
// The active-set stage, stage-wise factorisation.  Persistent grid: the two robots 2 m, 2 m + 1 of the launch are
// solved by the two halves of warp (m mod warps), warps = RIC_WARPS * gridDim.x; `ws` holds RIC_GAIN * N doubles per
// half-warp.  The halves share one instruction stream: control flow is warp-uniform, a half that has nothing (left)
// to do shadows the computation with its stores masked.  Robots the sweeps do not certify are queued (st.fb_list) for the
// fallback stage the mode selects: ipm_kernel below (MPCQP_MODE_IPM) or the dense ADMM kernel (MPCQP_MODE_ADMM, N <= 32).
// FULL: the horizon fills the capacity (n = N is a compile-time constant: the headline horizons 16, 32, 64 keep static index
// arithmetic); otherwise n = P.N < N at run time.
#ifndef RIC_MIN_CTAS
#define RIC_MIN_CTAS 1          // tuning hook: a larger value caps the registers (12 -> 168 registers, three warps per scheduler)
#endif
template <int N, bool FULL>
__global__ void __launch_bounds__(32 * RIC_WARPS, RIC_MIN_CTAS)
riccati_kernel(DevParams P, DevState st, DevScenario SC, const double* __restrict__ xref_g, const double* __restrict__ fsteps_g,
               double* __restrict__ ws_g, int* __restrict__ work_ctr, int first_tick, int inst_offset, int inst_count) {
    using S = RicInst<N>;
    constexpr int ROUNDS = S::ROUNDS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = FULL ? N : P.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane >> 4, hl = lane & 15;
    const int gwarp = blockIdx.x * RIC_WARPS + warp;
    S& sm = reinterpret_cast<S*>(smem_raw)[warp * 2 + sub];
    double* ws = ws_g + (size_t)(gwarp * 2 + sub) * ric_ws_slot_doubles(N);      // per half-warp: stage gains, then interior-point state
    if (hl == 0) mbar_init(&sm.mbar, 1);
    RIC_CANARY_ARM();
    // the fallback kernel behind this one is launched programmatically dependent: let it become resident (its prologue runs, then it
    // blocks in griddepcontrol.wait) as the CTAs of this persistent grid retire, instead of after the grid has drained
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (st.fb_next != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        // counters are double-buffered by tick: nobody uses the other copy during this tick
        *reinterpret_cast<int4*>(st.fb_next) = make_int4(0, 0, 0, 0);
    }
    __syncwarp();
    unsigned int phase = 0;
    const bool has_fallback = (P.mode & (2 | 8)) != 0;
    // Work distribution: every warp takes its first pair of robots by position and every further pair from a counter, so a
    // warp whose robots needed a second sweep does not also hold up the robots a static schedule would queue behind them.
    for (int w0 = gwarp * 2; w0 < inst_count;
         w0 = gridDim.x * RIC_PER_CTA + __shfl_sync(RIC_FULL, lane == 0 ? atomicAdd(work_ctr, 2) : 0, 0)) {
        const bool valid = w0 + sub < inst_count;
        const int inst = inst_offset + (valid ? w0 + sub : w0);     // an idle half shadows its neighbour, stores masked
        __syncwarp();
        RPROF_T0();
        RPROF_COUNT(12);
        const bool warm = P.warm_start && !first_tick;
        SigVec<ROUNDS> sg, nsg;
        // warm start: the previous tick's signatures advanced by one step (MPC.py:403-406); all loads in flight at once
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int t = hl + 16 * r, k = t >> 2, j = t & 3, ks = (k + 1 < n) ? k + 1 : 0;
            const uint8_t s8 = (warm && t < 4 * n) ? st.sig[(size_t)inst * 4 * n + 4 * ks + j] : SIG_FREE;
            sg.set(r, s8 > 26 ? SIG_FREE : s8);
        }
        unsigned conbits = 0u;
        const bool any_bad = ric_load_decode<N>(P, SC, sm, xref_g, fsteps_g, inst, valid, first_tick, sub, hl, n, phase, conbits);
#pragma unroll 1
        for (int r = 0; r < ROUNDS; ++r)
            if (!((conbits >> r) & 1u)) sg.set(r, SIG_FREE);
        int sweeps = 0, status = 0;
        bool done = false;
        if (any_bad) {
            status = 3;
            conbits = 0u;
#pragma unroll 1
            for (int r = 0; r < ROUNDS; ++r) sg.set(r, SIG_FREE);
        }
        ric_active_set<N>(P, sm, ws, sub, hl, n, conbits, sg, nsg, valid && !any_bad, P.max_sweeps, !(P.mode & 8), sweeps, done, status);
        // a robot the sweeps gave up on goes to the fallback stage with its carried state untouched
        const bool pushed = valid && !done && !any_bad && has_fallback;
        if (pushed && hl == 0) {
            const int q = atomicAdd(st.fb_count, 1);
            RIC_INVARIANT(q >= 0 && q < P.batch && inst >= 0 && inst < P.batch);
            st.fb_list[q] = inst;
            st.sweeps[inst] = sweeps;
        }
        RPROF(15);
        ric_finish<N>(P, SC, sm, st, inst, sub, hl, n, conbits, sg, done, status, sweeps, 0, nullptr, valid && !pushed);
        RPROF(16);
#ifdef MPCQP_CANARY
        if (P.refine == 77 && hl == 0 && inst == 0) sm.xst[12 * N] = 1.0;      // self-test of the detector: one element past the states
#endif
        RIC_CANARY_CHECK();
    }
}

// The interior-point stage for the robots queued in st.fb_list (launched programmatically dependent on riccati_kernel: resident
// early, reads the queue after griddepcontrol.wait).  Half a warp per robot like the active-set stage, persistent grid, static
// striding over the queue.  Per robot: interior-point iterations to a complementarity gap of 1e-8 (then 1e-11), the rows with
// y > s as signature, active-set sweeps from it (same guard => status SOLVED means the same thing as on the fast path).  A robot
// that is still uncertified leaves with its strictly feasible interior-point forces, the states rolled out from exactly those
// forces, and status MPCQP_STATUS_MAX_ITER.
template <int N>
__global__ void __launch_bounds__(32 * RIC_WARPS)
ipm_kernel(DevParams P, DevState st, DevScenario SC, const double* __restrict__ xref_g, const double* __restrict__ fsteps_g,
           double* __restrict__ ws_g, int first_tick) {
    using S = RicInst<N>;
    constexpr int NF = S::NF, ROUNDS = S::ROUNDS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = P.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane >> 4, hl = lane & 15;
    const int gwarp = blockIdx.x * RIC_WARPS + warp;
    S& sm = reinterpret_cast<S*>(smem_raw)[warp * 2 + sub];
    double* ws = ws_g + (size_t)(gwarp * 2 + sub) * ric_ws_slot_doubles(N);
    double* adm = ws + (size_t)RIC_GAIN * N + RIC_WS_PAD;
    if (hl == 0) mbar_init(&sm.mbar, 1);
    __syncwarp();
    RIC_CANARY_ARM();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int n_work = *st.fb_count;
    unsigned int phase = 0;
    for (int w0 = gwarp * 2; w0 < n_work; w0 += gridDim.x * RIC_PER_CTA) {
        const bool valid = w0 + sub < n_work;
        const int inst = st.fb_list[valid ? w0 + sub : w0];
        __syncwarp();
        unsigned conbits = 0u;
        const bool any_bad = ric_load_decode<N>(P, SC, sm, xref_g, fsteps_g, inst, valid, first_tick, sub, hl, n, phase, conbits);
        SigVec<ROUNDS> sg, nsg;
        int sweeps = valid ? st.sweeps[inst] : 0, status = 0, iters = 0;
        bool done = false;
        // ---- start: every stance force well inside its pyramid, gap 1
#pragma unroll 1
        for (int r = 0; r < ROUNDS; ++r) {
            const int t = hl + 16 * r;
            sg.set(r, SIG_FREE);
            const double f0[3] = {0.0, 0.0, 2.0};
            const double s0[6] = {P.mu * 2.0, P.mu * 2.0, P.mu * 2.0, P.mu * 2.0, 2.0, P.fz_max - 2.0};
#pragma unroll
            for (int c = 0; c < 3; ++c) adm[c * NF + t] = f0[c];
#pragma unroll
            for (int q = 0; q < 6; ++q) adm[(3 + q) * NF + t] = 1.0 / s0[q];
        }
        __syncwarp();
        double gap = 1.0, sigma = 0.3;
        bool alive = valid && !any_bad;
        for (int round = 0; round < 2; ++round) {
            const bool want = alive && !done;
            if (!__any_sync(RIC_FULL, want)) break;
            const bool ok = ric_ipm<N>(P, sm, ws, adm, sub, hl, n, conbits, want, round == 0 ? 1e-8 : 1e-11, P.ipm_max_iter, gap, sigma, iters);
            if (want && !ok) alive = false;
            // ---- signature of the iterate: a row is active iff its multiplier outweighs its slack
#pragma unroll 1
            for (int r = 0; r < ROUNDS; ++r) {
                const int t = hl + 16 * r;
                if (want && ((conbits >> r) & 1u)) {
                    double f[3], y[6];
#pragma unroll
                    for (int c = 0; c < 3; ++c) f[c] = adm[c * NF + t];
#pragma unroll
                    for (int q = 0; q < 6; ++q) y[q] = adm[(3 + q) * NF + t];
                    const double mz = P.mu * f[2];
                    const bool a0 = y[0] > mz - f[0], a1 = y[1] > mz + f[0], a2 = y[2] > mz - f[1], a3 = y[3] > mz + f[1];
                    const bool a4 = y[4] > f[2], a5 = y[5] > P.fz_max - f[2];
                    const bool apex = a4 || (a0 && a1) || (a2 && a3);
                    sg.set(r, sig_pack((a0 ? 1 : 0) - (a1 ? 1 : 0), (a2 ? 1 : 0) - (a3 ? 1 : 0), apex ? 1 : (a5 ? 2 : 0)));
                }
            }
            ric_active_set<N>(P, sm, ws, sub, hl, n, conbits, sg, nsg, want && alive, P.max_sweeps > 8 ? P.max_sweeps : 8, false, sweeps, done, status);
        }
        // The sweeps above ran on both halves of the warp: a robot that got solved first had its results overwritten by the
        // shadow work.  One more sweep restores them; an uncertified robot rolls out its interior-point forces instead.
        const bool failed = valid && !any_bad && !done;
        if (failed) {
            status = 2;
#pragma unroll 1
            for (int r = 0; r < ROUNDS; ++r) sg.set(r, SIG_PIN);
        }
        ric_sweep<N>(P, sm, ws, sub, hl, n, conbits, sg, nsg, adm);
        if (any_bad) status = 3;
        RIC_INVARIANT(inst >= 0 && inst < P.batch);
        ric_finish<N>(P, SC, sm, st, inst, sub, hl, n, any_bad ? 0u : conbits, sg, (done || failed) && !any_bad, status, sweeps, iters,
                      failed ? adm : nullptr, valid);
        RIC_CANARY_CHECK();
    }
}

}  // namespace mpcqp
