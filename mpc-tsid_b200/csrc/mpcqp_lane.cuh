// mpcqp_lane.cuh -- the active-set stage for LARGE batches: ONE LANE solves one robot's QP.
//
// Same mathematics as mpcqp_riccati.cuh (stage-wise Riccati recursion on the faces a signature selects, KKT guard,
// warm-started primal-dual active-set sweeps), mapped the other way round.  riccati_kernel gives a robot
// 16 lanes so that one sweep is short (it is the latency of a tick at 4096 robots), and pays for it: every lane factors the
// same two 6x6 blocks per stage, so ~4.6x of the executed FP64 work is redundant and the registers cap the SM at 16 robots.
// With tens of thousands of robots per GPU there is no latency to buy: here a lane owns a robot outright --
//   * zero redundant arithmetic, no shuffles, no warp barriers inside a sweep: the instruction stream of a lane IS the useful
//     work (~3 k instructions per stage for 32 robots against ~1 k for 2);
//   * the cost-to-go of the stage being processed (78 + 12 doubles) lives in shared memory, double buffered, thread-major
//     (element e of lane l at [e][l]: conflict-free); the two unit-triangular factors of a stage live in registers;
//   * everything of O(N) -- reference trajectory, lever arms, gains, states, costates, forces -- lives in an HBM / L2 workspace laid
//     out [item][slot] so that the 32 lanes of a warp always touch 256 contiguous bytes; E_k is assembled on the fly in the stage
//     that consumes it;
//   * the horizon is a loop bound (any n <= 64, one instantiation);
//   * lanes are persistent: a lane whose robot is certified (or handed to the interior-point stage) writes the outputs and takes
//     the next robot from a counter while its neighbours sweep on, so a warp never waits for its slowest robot.  The three phases
//     (fetch + decode / one sweep / outputs) are warp-synchronous: lanes reconverge at a barrier between them.
// Rows 6 + i and i of [Ppv; Pvv; pv'] are pushed through the triangular solves together (two independent chains per lane).
// Replaces MPC.update_ML / update_NK / call_solver / retrieve_result (MPC.py:316-458) for the robots it certifies; the others go
// to ipm_kernel through the same queue riccati_kernel uses.
#pragma once
#include "mpcqp_device.cuh"
#include "mpcqp_foot.cuh"
#include "mpcqp_scenario.cuh"

namespace mpcqp {

constexpr int LANE_PE = 90;                 // doubles of one cost-to-go: Ppp (21, packed lower), Ppv (36), Pvv (21), pp (6), pv (6)
constexpr int LANE_PPP = 0, LANE_PPV = 21, LANE_PVV = 57, LANE_PP = 78, LANE_PV = 84;
constexpr int LANE_SMEM_BYTES = 2 * LANE_PE * 32 * 8;
// workspace doubles per slot and the unsigned words (signature + contact of the four feet of a step: tried, proposed) behind them
__host__ __device__ constexpr int lane_ws_doubles(int n) { return 140 * n + 12; }
__host__ __device__ constexpr int lane_ws_words(int n) { return 2 * n; }

#define LTI(r, c) ((r) * ((r) + 1) / 2 + (c))

__device__ __forceinline__ double lane_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
}

// a = U D U' in place (U unit lower: strict part of a; diagonal left as garbage), dinv = 1 / D
__device__ __forceinline__ bool lane_ldl6(double (&a)[21], double (&dinv)[6]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const double d = a[LTI(j, j)];
        ok = ok && (d > 1e-300) && (d < 1e300);
        const double r = lane_rcp(d);
        dinv[j] = r;
        double t[6];
#pragma unroll
        for (int i = j + 1; i < 6; ++i) { t[i] = a[LTI(i, j)]; a[LTI(i, j)] = t[i] * r; }
#pragma unroll
        for (int i = j + 1; i < 6; ++i)
#pragma unroll
            for (int c = j + 1; c <= i; ++c) a[LTI(i, c)] = fma(-a[LTI(i, j)], t[c], a[LTI(i, c)]);
    }
    return ok;
}

// One row rho of [Ppv; Pvv; pv'] through the stage (mpcqp_riccati.cuh, header): with Pvv = U D U', H = D^-1 + U'EU = W Delta W',
//   y = rho U^-T,  e = ((y D^-1) W^-T Delta^-1) W^-1,   t = e U'  (row of Pt[:, v]),   kr = ((y - e) D^-1) U^-1  (row of [..] Gamma)
// by substitution (the factors are unit lower triangular, strict parts of u and w).
__device__ __forceinline__ void lane_row(const double (&rho)[6], const double (&u)[21], const double (&dinv)[6], const double (&w)[21],
                                         const double (&einv)[6], double (&t)[6], double (&kr)[6]) {
    double y[6], b[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        double s = rho[c];
#pragma unroll
        for (int r = 0; r < c; ++r) s = fma(-y[r], u[LTI(c, r)], s);
        y[c] = s;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        double s = y[c] * dinv[c];
#pragma unroll
        for (int r = 0; r < c; ++r) s = fma(-b[r], w[LTI(c, r)], s);
        b[c] = s;
    }
    double e[6];
#pragma unroll
    for (int c = 5; c >= 0; --c) {
        double s = b[c] * einv[c];
#pragma unroll
        for (int r = c + 1; r < 6; ++r) s = fma(-e[r], w[LTI(r, c)], s);
        e[c] = s;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        double s = e[c];
#pragma unroll
        for (int r = 0; r < c; ++r) s = fma(e[r], u[LTI(c, r)], s);
        t[c] = s;
    }
#pragma unroll
    for (int c = 5; c >= 0; --c) {
        double s = (y[c] - e[c]) * dinv[c];
#pragma unroll
        for (int r = c + 1; r < 6; ++r) s = fma(-kr[r], u[LTI(r, c)], s);
        kr[c] = s;
    }
}

__device__ __forceinline__ unsigned long long lane_mix(unsigned long long h, unsigned word, int k) {
    unsigned long long q = ((unsigned long long)word + 1ull) * 0x9E3779B97F4A7C15ull;
    q ^= q >> 29; q *= (2ull * (unsigned long long)k + 0xBF58476D1CE4E5B9ull); q ^= q >> 32;
    return h + q;
}

// per-lane view of the workspace
struct LaneWs {
    double* w;              // + slot
    unsigned* g;            // + slot
    size_t stride;          // slots (a multiple of 32)
    int n, lev, cs, beta, gain, xst, lam, frc;
    __device__ __forceinline__ double& d(int off, int i) const { return w[(size_t)(off + i) * stride]; }
    __device__ __forceinline__ unsigned& sig(int k) const { return g[(size_t)k * stride]; }
    __device__ __forceinline__ unsigned& nsig(int k) const { return g[(size_t)(n + k) * stride]; }
};
// signature word of a step: 5 bits of signature per foot (bits 5 j .. 5 j + 4), contact of foot j at bit 20 + j
__device__ __forceinline__ unsigned lane_sig_of(unsigned word, int j) { return (word >> (5 * j)) & 31u; }
__device__ __forceinline__ bool lane_contact_of(unsigned word, int j) { return (word >> (20 + j)) & 1u; }

__device__ __forceinline__ void lane_inertia(const DevParams& P, double cs, double sn, double (&Ii)[9]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        Ii[3 * a + 0] = P.gIinv[3 * a + 0] * cs - P.gIinv[3 * a + 1] * sn;
        Ii[3 * a + 1] = P.gIinv[3 * a + 0] * sn + P.gIinv[3 * a + 1] * cs;
        Ii[3 * a + 2] = P.gIinv[3 * a + 2];
    }
}

// Inputs of robot `inst` -> workspace (reference trajectory, yaw cosines / sines, lever arms, contact bits) and the warm-start
// signature (the previous tick's, advanced by one step, MPC.py:403-406).  Returns true if the inputs are malformed.
// Same decoding rules as decode_lever_next / step_inertia (mpcqp_foot.cuh)                          [MPC.py:316-360, 635-652]
__device__ bool lane_fetch(const DevParams& P, const DevState& st, const LaneWs& L, const double* __restrict__ xref_g,
                           const double* __restrict__ fsteps_g, int inst, int first_tick) {
    const int n = L.n, ld = n + 1;
    const double* xg = xref_g + (size_t)inst * 12 * ld;
    const double* fg = fsteps_g + (size_t)inst * 260;
    bool bad = false;
#pragma unroll 4
    for (int i = 0; i < 12 * ld; ++i) {
        const double v = __ldg(xg + i);
        bad = bad || !isfinite(v);
        L.d(0, i) = v;
    }
    for (int k = 0; k < n; ++k) {
        double sn, cs;
        sincos(__ldg(xg + 5 * ld + k), &sn, &cs);
        L.d(L.cs, k) = cs; L.d(L.cs, n + k) = sn;
    }
    const bool warm = P.warm_start && !first_tick;
    const uint8_t* sg = st.sig + (size_t)inst * 4 * n;
    int k0 = 0;
    for (int r = 0; r <= 20; ++r) {
        // row r of the gait table covers steps k0 .. k1 - 1; r == 20 (or a terminator) stands for "no row": no contact, foothold 0
        double cnt = 0.0;
        if (r < 20) cnt = __ldg(fg + r * 13);
        bool none = (r == 20) || cnt == 0.0;
        if (!none && (!(cnt > 0.0) || cnt != floor(cnt))) { bad = true; none = true; }
        double foot[12];
        unsigned cmask = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double v = 0.0;
                if (!none) {
                    v = __ldg(fg + r * 13 + 1 + 3 * j + c);
                    if (c == 0 && !(isnan(v) || v == 0.0)) cmask |= 1u << j;          // MPC.py:650
                    if (isnan(v)) v = 0.0;                                            // MPC.py:327
                }
                if (first_tick) v = P.footholds[c * 4 + j];                           // MPC.py:176
                foot[3 * j + c] = v;
            }
        }
        const int k1 = none ? n : ((cnt >= (double)(n - k0)) ? n : k0 + (int)cnt);
        for (int k = k0; k < k1; ++k) {
            unsigned word = cmask << 20;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int c = 0; c < 3; ++c) L.d(L.lev, c * 4 * n + 4 * k + j) = foot[3 * j + c] - __ldg(xg + c * ld + k);   // MPC.py:343
                const int ks = (k + 1 < n) ? k + 1 : 0;
                uint8_t s8 = SIG_FREE;
                if (warm && ((cmask >> j) & 1u)) { s8 = sg[4 * ks + j]; if (s8 > 26) s8 = SIG_FREE; }
                word |= (unsigned)s8 << (5 * j);
            }
            L.sig(k) = word;
        }
        k0 = k1;
        if (k0 >= n) break;
    }
    return bad;
}

// One equality-constrained solve on the faces the signature words select + KKT guard (what ric_sweep does with 16 lanes).
// Returns 1 if every foot passes the guard, 0 if not, -1 if a pivot was not positive.  hs / hn: hashes of the signature tried
// and of the one the guard proposes (written to L.nsig).  On return the workspace holds states, velocity costates and forces.
__device__ __noinline__ int lane_sweep(const DevParams& P, const LaneWs& L, double* __restrict__ psm, unsigned long long& hs,
                                       unsigned long long& hn) {
    const int n = L.n, ld = n + 1;
    const int lane = threadIdx.x & 31;
    const double dt = P.dt, lin = P.dt / P.mass, lin2 = lin * lin;
#define PS(buf, e) psm[((buf) * LANE_PE + (e)) * 32 + lane]
    // ---- terminal cost-to-go: P_N = Q, p_N = -Q xref_N
    {
        const int b = (n - 1) & 1;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int c = 0; c <= r; ++c) {
                PS(b, LANE_PPP + LTI(r, c)) = (r == c) ? P.wp[r] : 0.0;
                PS(b, LANE_PVV + LTI(r, c)) = (r == c) ? P.wv[r] : 0.0;
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) PS(b, LANE_PPV + 6 * r + c) = 0.0;
            PS(b, LANE_PP + r) = -P.wp[r] * L.d(0, r * ld + n);
            PS(b, LANE_PV + r) = -P.wv[r] * L.d(0, (6 + r) * ld + n);
        }
    }
    bool spd = true;
#pragma unroll 1
    for (int k = n - 1; k >= 0; --k) {
        const int in = k & 1, out = in ^ 1;
        // (0) E_k = sum_j Bv S Bv' (packed lower), beta_k = g + sum_j Bv pf             [ric_assemble]
        double e[21], ub[6];
#pragma unroll
        for (int i = 0; i < 21; ++i) e[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) ub[i] = 0.0;
        ub[2] = -P.gravity * P.dt;                             // MPC.py:200-201
        {
            const unsigned word = L.sig(k);
            double Ii[9];
            lane_inertia(P, L.d(L.cs, k), L.d(L.cs, n + k), Ii);
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                if (!lane_contact_of(word, j)) continue;
                Face fc;
                make_face(P, true, (uint8_t)lane_sig_of(word, j), fc);
                const double S0 = fc.dx + fc.dz * fc.czx * fc.czx, S1 = fc.dz * fc.czx * fc.czy, S2 = fc.dy + fc.dz * fc.czy * fc.czy;
                const double S3 = fc.dz * fc.czx, S4 = fc.dz * fc.czy, S5 = fc.dz;
                const int t = 4 * k + j;
                const double r[3] = {L.d(L.lev, t), L.d(L.lev, 4 * n + t), L.d(L.lev, 8 * n + t)};
                double A[9], M[9];
                lever_block(P, Ii, r, A);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    M[3 * q + 0] = A[3 * q] * S0 + A[3 * q + 1] * S1 + A[3 * q + 2] * S3;
                    M[3 * q + 1] = A[3 * q] * S1 + A[3 * q + 1] * S2 + A[3 * q + 2] * S4;
                    M[3 * q + 2] = A[3 * q] * S3 + A[3 * q + 1] * S4 + A[3 * q + 2] * S5;
                }
                e[LTI(0, 0)] = fma(lin2, S0, e[LTI(0, 0)]); e[LTI(1, 0)] = fma(lin2, S1, e[LTI(1, 0)]);
                e[LTI(1, 1)] = fma(lin2, S2, e[LTI(1, 1)]); e[LTI(2, 0)] = fma(lin2, S3, e[LTI(2, 0)]);
                e[LTI(2, 1)] = fma(lin2, S4, e[LTI(2, 1)]); e[LTI(2, 2)] = fma(lin2, S5, e[LTI(2, 2)]);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) e[LTI(3 + q, c)] = fma(lin, M[3 * q + c], e[LTI(3 + q, c)]);
#pragma unroll
                    for (int c = 0; c <= q; ++c)
                        e[LTI(3 + q, 3 + c)] += M[3 * q] * A[3 * c] + M[3 * q + 1] * A[3 * c + 1] + M[3 * q + 2] * A[3 * c + 2];
                }
                double u6[6];
                bv_apply(A, lin, fc.pf, u6);
#pragma unroll
                for (int i = 0; i < 6; ++i) ub[i] += u6[i];
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) L.d(L.beta, 6 * k + i) = ub[i];
        }
        // (1) Pvv = U D U'
        double u[21], dinv[6];
#pragma unroll
        for (int i = 0; i < 21; ++i) u[i] = PS(in, LANE_PVV + i);
        spd = lane_ldl6(u, dinv) && spd;
        // (2) H = D^-1 + U' (E U) = W Delta W'
        double w[21], einv[6];
        {
            double T[21];                                      // (E U)(r, c) for r >= c
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) {
                    double t = e[LTI(r, c)];
#pragma unroll
                    for (int q = c + 1; q < 6; ++q) t = fma((r >= q) ? e[LTI(r, q)] : e[LTI(q, r)], u[LTI(q, c)], t);
                    T[LTI(r, c)] = t;
                }
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = 0; c <= a; ++c) {
                    double t = T[LTI(a, c)] + ((a == c) ? dinv[a] : 0.0);
#pragma unroll
                    for (int r = a + 1; r < 6; ++r) t = fma(u[LTI(r, a)], T[LTI(r, c)], t);
                    w[LTI(a, c)] = t;
                }
        }
        spd = lane_ldl6(w, einv) && spd;
        // (3) rows: pv' alone, then rows i of Ppv and Pvv together; the new cost-to-go goes to the other buffer
        double pvin[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) pvin[c] = PS(in, LANE_PV + c);
        {
            double t[6], kr[6];
            lane_row(pvin, u, dinv, w, einv, t, kr);
#pragma unroll
            for (int o = 0; o < 6; ++o) L.d(L.gain, 78 * k + 72 + o) = kr[o];
        }
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {
            double rp[6], rv[6], tp[6], tv[6], kp[6], kv[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                rp[c] = PS(in, LANE_PPV + 6 * i + c);
                rv[c] = PS(in, LANE_PVV + ((i >= c) ? LTI(i, c) : LTI(c, i)));
            }
            lane_row(rp, u, dinv, w, einv, tp, kp);
            lane_row(rv, u, dinv, w, einv, tv, kv);
#pragma unroll
            for (int o = 0; o < 6; ++o) {
                L.d(L.gain, 78 * k + 6 * i + o) = kp[o];
                L.d(L.gain, 78 * k + 36 + 6 * i + o) = kv[o];
            }
            if (k > 0) {
                // rows of Pt[:, p] = P+[:, p] - (rho Gamma) Pvp and the components of pt, Pt[:, v] beta + pt
                double wp_[6], wv_[6];
                double hp = PS(in, LANE_PP + i), hv = PS(in, LANE_PV + i);
#pragma unroll
                for (int q = 0; q < 6; ++q) { hp = fma(-kp[q], pvin[q], hp); hv = fma(-kv[q], pvin[q], hv); }
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    double a = PS(in, LANE_PPP + ((i >= c) ? LTI(i, c) : LTI(c, i)));
                    double b = PS(in, LANE_PPV + 6 * c + i);
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const double pcq = PS(in, LANE_PPV + 6 * c + q);
                        a = fma(-kp[q], pcq, a);
                        b = fma(-kv[q], pcq, b);
                    }
                    wp_[c] = a; wv_[c] = b;
                }
#pragma unroll
                for (int q = 0; q < 6; ++q) { hp = fma(tp[q], ub[q], hp); hv = fma(tv[q], ub[q], hv); }
                // P_k = Q + A' Pt A,  p_k = -Q xref_k + A'(Pt[:, v] beta + pt)
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const double npv = fma(dt, wp_[c], tp[c]);
                    PS(out, LANE_PPV + 6 * i + c) = npv;
                    if (c <= i) {
                        PS(out, LANE_PPP + LTI(i, c)) = wp_[c] + ((c == i) ? P.wp[i] : 0.0);
                        PS(out, LANE_PVV + LTI(i, c)) = fma(dt, npv + wv_[c], tv[c]) + ((c == i) ? P.wv[i] : 0.0);
                    }
                }
                PS(out, LANE_PP + i) = hp - P.wp[i] * L.d(0, i * ld + k);
                PS(out, LANE_PV + i) = fma(dt, hp, hv) - P.wv[i] * L.d(0, (6 + i) * ld + k);
            }
        }
    }
    // ---- forward pass
    {
        double x[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) x[c] = L.d(0, c * ld);
#pragma unroll 1
        for (int k = 0; k < n; ++k) {
            double z[12], wimp[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) { z[c] = fma(dt, x[6 + c], x[c]); z[6 + c] = x[6 + c] + L.d(L.beta, 6 * k + c); }
#pragma unroll
            for (int o = 0; o < 6; ++o) wimp[o] = L.d(L.gain, 78 * k + 72 + o);
#pragma unroll
            for (int r = 0; r < 12; ++r)
#pragma unroll
                for (int o = 0; o < 6; ++o) wimp[o] = fma(L.d(L.gain, 78 * k + 6 * r + o), z[r], wimp[o]);
#pragma unroll
            for (int c = 0; c < 6; ++c) { x[c] = z[c]; x[6 + c] = z[6 + c] - wimp[c]; }
#pragma unroll
            for (int c = 0; c < 12; ++c) L.d(L.xst, 12 * k + c) = x[c];
        }
    }
    // ---- velocity costates lam_s = Q e_s + A' lam_{s+1}, and per foot of step s - 1: force on the face, gradient, KKT guard
    bool ok = true;
    hs = 0ull; hn = 0ull;
    {
        double lp[6], lv[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) { lp[c] = 0.0; lv[c] = 0.0; }
#pragma unroll 1
        for (int s = n; s >= 1; --s) {
            const int k = s - 1;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const double ep = L.d(L.xst, 12 * k + c) - L.d(0, c * ld + s), ev = L.d(L.xst, 12 * k + 6 + c) - L.d(0, (6 + c) * ld + s);
                lv[c] = fma(P.wv[c], ev, fma(dt, lp[c], lv[c]));
                lp[c] = fma(P.wp[c], ep, lp[c]);
                L.d(L.lam, 6 * k + c) = lv[c];
            }
            const unsigned word = L.sig(k);
            unsigned nword = word;
            double Ii[9];
            lane_inertia(P, L.d(L.cs, k), L.d(L.cs, n + k), Ii);
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                if (!lane_contact_of(word, j)) continue;
                const int t = 4 * k + j;
                const uint8_t sgj = (uint8_t)lane_sig_of(word, j);
                Face fc;
                make_face(P, true, sgj, fc);
                const double r[3] = {L.d(L.lev, t), L.d(L.lev, 4 * n + t), L.d(L.lev, 8 * n + t)};
                double A[9], h[3];
                lever_block(P, Ii, r, A);
                bvT_apply(A, lin, lv, h);
                const double qx = fc.zx ? -fc.dx * h[0] : 0.0;
                const double qy = fc.zy ? -fc.dy * h[1] : 0.0;
                const double qz = fc.zz ? -fc.dz * (fc.czx * h[0] + fc.czy * h[1] + h[2]) : 0.0;
                double f[3] = {fc.pf[0] + qx + fc.czx * qz, fc.pf[1] + qy + fc.czy * qz, fc.pf[2] + qz};
                const double grad[3] = {fma(P.w_force, f[0], h[0]), fma(P.w_force, f[1], h[1]), fma(P.w_force, f[2], h[2])};
                FootSol sol;
                uint8_t ns;
                ok = kkt_guard(P, sgj, f, grad, sol, ns) && ok;
                nword = (nword & ~(31u << (5 * j))) | ((unsigned)ns << (5 * j));
                L.d(L.frc, 3 * t) = f[0]; L.d(L.frc, 3 * t + 1) = f[1]; L.d(L.frc, 3 * t + 2) = f[2];
            }
            L.nsig(k) = nword;
            hs = lane_mix(hs, word, k);
            hn = lane_mix(hn, nword, k);
        }
    }
#undef PS
    return !spd ? -1 : (ok ? 1 : 0);
}

// Outputs of one robot (what ric_finish does with 16 lanes)                                      [MPC.py:432-458, 503-510]
__device__ __noinline__ void lane_finish(const DevParams& P, const DevScenario& SC, const DevState& st, const LaneWs& L, int inst, bool solved,
                                         int status, int sweeps) {
    const int n = L.n, ld = n + 1;
    const int AW = (20 * n + 31) / 32, CW = (4 * n + 31) / 32;
    const double lin = P.dt / P.mass;
    double part = 0.0;
    double xn[12];
    double* xs = st.xs + (size_t)inst * 12 * n;
    {
        // not solved (malformed input): no forces, the states are the free response p+ = p + dt v, v+ = v + g
        double p[6], v[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) { p[c] = L.d(0, c * ld); v[c] = L.d(0, (6 + c) * ld); }
        for (int s = 0; s < n; ++s) {
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                double xp, xv;
                if (solved) { xp = L.d(L.xst, 12 * s + c); xv = L.d(L.xst, 12 * s + 6 + c); }
                else { xp = p[c] + P.dt * v[c]; xv = v[c] + ((c == 2) ? -P.gravity * P.dt : 0.0); p[c] = xp; v[c] = xv; }
                const double ep = xp - L.d(0, c * ld + s + 1), ev = xv - L.d(0, (6 + c) * ld + s + 1);
                const double eep = isfinite(ep) ? ep : 0.0, eev = isfinite(ev) ? ev : 0.0;     // malformed input: never NaN out
                xs[12 * s + c] = eep; xs[12 * s + 6 + c] = eev;                                 // MPC.x[:12N] (MPC.py:428)
                part = fma(0.5 * P.wp[c] * eep, eep, part);
                part = fma(0.5 * P.wv[c] * eev, eev, part);
                if (s == 0) { xn[c] = isfinite(ep) ? xp : 0.0; xn[6 + c] = isfinite(ev) ? xv : 0.0; }
            }
        }
#pragma unroll
        for (int c = 0; c < 12; ++c) st.x1[(size_t)inst * 12 + c] = xn[c];                      // MPC.q_next / v_next (MPC.py:448-450)
    }
    unsigned aword = 0u, cword = 0u;
    int aidx = 0;
    for (int k = 0; k < n; ++k) {
        const unsigned word = L.sig(k);
        double Ii[9], lv[6];
        lane_inertia(P, L.d(L.cs, k), L.d(L.cs, n + k), Ii);
#pragma unroll
        for (int c = 0; c < 6; ++c) lv[c] = L.d(L.lam, 6 * k + c);
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int t = 4 * k + j;
            const bool contact = solved && lane_contact_of(word, j);
            const uint8_t sgj = (uint8_t)lane_sig_of(word, j);
            double f[3] = {0.0, 0.0, 0.0};
            FootSol sol;
#pragma unroll
            for (int q = 0; q < 5; ++q) sol.y[q] = 0.0;
            if (contact) {
                f[0] = L.d(L.frc, 3 * t); f[1] = L.d(L.frc, 3 * t + 1); f[2] = L.d(L.frc, 3 * t + 2);
                const double r[3] = {L.d(L.lev, t), L.d(L.lev, 4 * n + t), L.d(L.lev, 8 * n + t)};
                double A[9], h[3];
                lever_block(P, Ii, r, A);
                bvT_apply(A, lin, lv, h);
                const double grad[3] = {fma(P.w_force, f[0], h[0]), fma(P.w_force, f[1], h[1]), fma(P.w_force, f[2], h[2])};
                uint8_t dummy;
                kkt_guard(P, sgj, f, grad, sol, dummy);
            }
            part += 0.5 * P.w_force * (f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
            double* fo = st.f + (size_t)inst * 12 * n + 3 * t;
            fo[0] = f[0]; fo[1] = f[1]; fo[2] = f[2];
            double* yo = st.y + (size_t)inst * 20 * n + 5 * t;
#pragma unroll
            for (int q = 0; q < 5; ++q) yo[q] = sol.y[q];
            st.sig[(size_t)inst * 4 * n + t] = (!contact || sgj > 26) ? SIG_FREE : sgj;
            if (k == 0) {
                double* f0 = st.f0 + (size_t)inst * 12 + 3 * j;
                f0[0] = f[0]; f0[1] = f[1]; f0[2] = f[2];
            }
            // rows that hold with equality; a swing foot is pinned to f = 0 (MPC.py:355-358): all five of its rows do
            const double mu = P.mu, tol = 1e-9;
            const double row[5] = {f[0] - mu * f[2], -f[0] - mu * f[2], f[1] - mu * f[2], -f[1] - mu * f[2], -f[2]};
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const bool act = (fabs(row[q]) <= tol) || (q == 4 && fabs(row[4] + P.fz_max) <= tol);
                const int b = 5 * t + q;
                if (act) aword |= 1u << (b & 31);
                if ((b & 31) == 31) { st.active[(size_t)inst * AW + aidx] = aword; aword = 0u; ++aidx; }
            }
            if (contact) cword |= 1u << (t & 31);
        }
        if ((k & 7) == 7 || k == n - 1) { st.contact[(size_t)inst * CW + (k >> 3)] = cword; cword = 0u; }
    }
    if (aidx < AW) st.active[(size_t)inst * AW + aidx] = aword;
    st.obj[inst] = part;
    st.status[inst] = status;
    st.sweeps[inst] = sweeps;
    st.iters[inst] = 0;
    if (status != 3) {
        double qw[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) qw[c] = st.qw[(size_t)inst * 6 + c];
        world_pose_step(qw, xn);
#pragma unroll
        for (int c = 0; c < 6; ++c) st.qw[(size_t)inst * 6 + c] = qw[c];
        if (SC.enabled) scenario_advance(SC, inst, xn);
    }
}

// The active-set stage, one robot per lane, persistent lanes.  `ws` holds lane_ws_doubles(n) doubles and lane_ws_words(n) words per
// slot (slot = global thread), [item][slot]; `work_ctr` must be zero at launch.  Robots the sweeps do not certify are queued
// (st.fb_list) for ipm_kernel exactly as riccati_kernel queues them.
__global__ void __launch_bounds__(32, 4)
lane_kernel(const __grid_constant__ DevParams P, const __grid_constant__ DevState st, const __grid_constant__ DevScenario SC, const double* __restrict__ xref_g, const double* __restrict__ fsteps_g,
            double* __restrict__ ws, int* __restrict__ work_ctr, int first_tick, int inst_offset, int inst_count) {
    extern __shared__ __align__(16) double lane_psm[];
    const int lane = threadIdx.x & 31;
    const int slots = gridDim.x * 32;
    const int slot = blockIdx.x * 32 + lane;
    const int n = P.N;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (st.fb_next != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<int4*>(st.fb_next) = make_int4(0, 0, 0, 0);
    LaneWs L;
    L.w = ws + slot;
    L.g = reinterpret_cast<unsigned*>(ws + (size_t)lane_ws_doubles(n) * slots) + slot;
    L.stride = (size_t)slots;
    L.n = n;
    L.lev = 12 * (n + 1); L.cs = L.lev + 12 * n; L.beta = L.cs + 2 * n; L.gain = L.beta + 6 * n;
    L.xst = L.gain + 78 * n; L.lam = L.xst + 12 * n; L.frc = L.lam + 6 * n;
    // per-robot state of the lane
    bool have = false, exhausted = false, bad = false, done = false, stop = false;
    int inst = 0, sweeps = 0, nhist = 0, status = 0;
    int next = slot;                         // first robot by position, every further one from the counter
    unsigned long long hist[16];
    while (true) {
        // ---- phase 1: lanes without a robot take one
        if (!have && !exhausted) {
            if (next < 0) next = slots + atomicAdd(work_ctr, 1);
            if (next < inst_count) {
                inst = inst_offset + next;
                bad = lane_fetch(P, st, L, xref_g, fsteps_g, inst, first_tick);
                have = true; done = false; stop = bad; sweeps = 0; nhist = 0; status = bad ? 3 : 0;
            } else exhausted = true;
            next = -1;
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, have)) break;
        // ---- phase 2: one sweep for every lane whose robot is neither certified nor given up on
        if (have && !done && !stop) {
            if (sweeps >= P.max_sweeps) stop = true;
            else {
                unsigned long long hs, hn;
                const int rc = lane_sweep(P, L, lane_psm, hs, hn);
                ++sweeps;
                if (rc < 0) stop = true;
                else if (rc > 0) { done = true; status = 1; }
                else {
                    // the signature just tried goes into the history; every foot adopts its proposal (primal-dual active-set step)
                    // unless that signature was tried before (it would cycle: the interior-point stage takes over)
                    if (nhist < 16) hist[nhist++] = hs;
                    bool seen = false;
                    for (int i = 0; i < nhist; ++i) seen = seen || (hist[i] == hn);
                    if (seen) stop = true;
                    else for (int k = 0; k < n; ++k) L.sig(k) = L.nsig(k);
                }
            }
        }
        __syncwarp();
        // ---- phase 3: outputs of the robots that are through; the lane is free again
        if (have && (done || stop)) {
            if (done || bad) lane_finish(P, SC, st, L, inst, done, status, sweeps);
            else {
                const int q = atomicAdd(st.fb_count, 1);
                st.fb_list[q] = inst;
                st.sweeps[inst] = sweeps;
            }
            have = false;
        }
        __syncwarp();
    }
}

#undef LTI
}  // namespace mpcqp
