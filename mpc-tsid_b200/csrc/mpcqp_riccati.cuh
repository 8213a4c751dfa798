// mpcqp_riccati.cuh -- stage-wise (Riccati) active-set sweeps: one WARP solves one robot's QP.
//
// Same mathematics as the dense path (condensed QP on the faces a signature selects,
// KKT guard, warm-started primal-dual active-set sweeps), different factorisation.  Instead of the
// 6N x 6N Woodbury matrix W (O(N^3) per sweep, one CTA per robot) the equality-constrained QP
//     min  sum_s 1/2 (x_s - xref_s)' Q (x_s - xref_s) + 1/2 w_f sum |f|^2
//     s.t. x_{k+1} = A x_k + [0; u_k + g],  u_k = sum_j Bv_kj f_kj,  f_kj = pf_kj + Z_kj q_kj
// is solved by dynamic programming over the horizon (O(N), 6x6 blocks only), x = [p (6); v (6)],
// A = [[I, dt I], [0, I]] (MPC.py:110-111), cost-to-go 1/2 x'P_k x + p_k'x:
//   backward, stage k = N-1..0, with P+ = P_{k+1}, E_k = sum_j (Bv Z) R^-1 (Bv Z)' (6x6), beta_k = ubar_k + g:
//       L L' = Pvv+,  G = I + L' E L = M M',  X = L M^-T,  Y = Ppv+ L^-T,  Y2 = Y M^-T,  U = M^-1 L^-1
//       Pt = (P+^-1 + [0 0; 0 E])^-1:   Ptvv = X X',  Ptpv = Y2 X',  Ptpp = Ppp+ - Y Y' + Y2 Y2'
//       Gamma = (I + E Pvv+)^-1 E,  Gamma Pvp = (Y L^-1 - Y2 U)',  Gamma Pvv = I - U' X'
//       P_k = Q + A' Pt A,   p_k = -Q xref_k + A'(Pt[:,v] beta_k + pt)
//       closed loop  w_k = -(Kx x_k + k0)   stored per stage (78 doubles)
//   forward: x_{k+1} = A x_k + [0; beta_k + w_k];  costates lam_s = Q (x_s - xref_s) + A' lam_{s+1};
//   per foot: h = Bv' lam^v_{k+1},  q = -R^-1 Z' h,  f = pf + Z q,  grad = w_f f + h   (= H f + g of the
//   condensed problem), then the same KKT guard as the dense path (mpcqp_foot.cuh).
// Everything a stage needs lives in the warp's private shared-memory block; the two 6x6 Cholesky
// factorisations per stage run redundantly in the registers of every lane (no shuffle or memory hop
// on the pivot chain), their inverses come out of the same loop (forward substitution of the unit
// vectors in the shadow of the pivots), all other 6x6 products are one output per lane.
// Replaces MPC.update_ML / update_NK / call_solver / retrieve_result (MPC.py:316-458) like the dense path.
#pragma once
#include "mpcqp_device.cuh"
#include "mpcqp_foot.cuh"
#include "mpcqp_scenario.cuh"

namespace mpcqp {

constexpr int RIC_SLOT = 78;        // doubles kept per stage for the forward pass: Kx (6 x 12), k0 (6)
constexpr int RIC_WARPS = 2;        // robots (= warps) per CTA

template <int N>
struct alignas(16) RicWarp {
    static constexpr int NF = 4 * N;
    static constexpr int ROUNDS = NF / 32;                     // feet per lane
    static constexpr int AW = (20 * N + 31) / 32, CW = (4 * N + 31) / 32;
    double xr[12 * (N + 1)];            // xref of this robot
    union {
        double fs[20 * 13];             // fsteps (dead after decode)
        double E[21 * N];               // per sweep: packed lower triangles of the 6x6 blocks E_k; once the
                                        // backward pass is done the same bytes hold the sweep's forces (3 x NF)
    };
    double A[9 * NF];                   // per foot-step: dt inv(R gI) [r]x, struct of arrays
    double beta[6 * N];                 // ubar_k + g;  in finish(): the impulses of the final forces
    double slot[RIC_SLOT * N];          // per stage Kx, k0; after the forward pass [x_{k+1} (12), lam^v_{k+1} (6), ..]
    double Ppp[36], Ppv[36], Pvv[36];   // cost-to-go of the stage being eliminated
    double L[36], Li[36], Mi[36], T[36], Y[36], G[36], X[36], Y2[36], U[36], Tpp[36], Tpv[36], Tvv[36];
    double pp[6], pv[6], av[6], cv[6], tp[6], tv[6], gam[6];
    unsigned long long hist[16];        // hashes of the signatures already tried
    unsigned long long mbar;
    unsigned int amask[AW + CW];
    ScenarioSmem sc;
    static_assert(21 * N >= 260 && 21 * N >= 12 * N, "union sizing");
};

__device__ __forceinline__ double dot_rr(const double* __restrict__ a, const double* __restrict__ b) {
    double s0 = a[0] * b[0], s1 = a[1] * b[1];
    s0 = fma(a[2], b[2], s0); s1 = fma(a[3], b[3], s1);
    s0 = fma(a[4], b[4], s0); s1 = fma(a[5], b[5], s1);
    return s0 + s1;
}
__device__ __forceinline__ double dot_rc(const double* __restrict__ a, const double* __restrict__ b) {   // b strided by 6
    double s0 = a[0] * b[0], s1 = a[1] * b[6];
    s0 = fma(a[2], b[12], s0); s1 = fma(a[3], b[18], s1);
    s0 = fma(a[4], b[24], s0); s1 = fma(a[5], b[30], s1);
    return s0 + s1;
}
__device__ __forceinline__ double dot_cc(const double* __restrict__ a, const double* __restrict__ b) {   // both strided by 6
    double s0 = a[0] * b[0], s1 = a[6] * b[6];
    s0 = fma(a[12], b[12], s0); s1 = fma(a[18], b[18], s1);
    s0 = fma(a[24], b[24], s0); s1 = fma(a[30], b[30], s1);
    return s0 + s1;
}

// Cholesky S = L L' of a 6x6 SPD matrix in shared memory (row major, both triangles valid), redundantly in
// the registers of every lane, fused with the inversion of L: lane c < 6 carries the unit vector e_c
// through the forward substitution as the pivots appear, so column c of inv(L) is complete one
// multiply after the last pivot.  Lout (if given) receives L (lower triangle only; the strictly upper part of
// the buffer is never written and stays zero), Liout receives inv(L) (full 6x6, zeros included).
__device__ __forceinline__ bool chol6_inv(const double* __restrict__ S, double* __restrict__ Lout, double* __restrict__ Liout, int lane) {
    double a[21];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r * (r + 1) / 2 + c] = S[r * 6 + c];
    const int cc = lane < 6 ? lane : 0;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = (i == cc) ? 1.0 : 0.0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const double d = a[j * (j + 1) / 2 + j];
        ok = ok && (d > 0.0);
        const double inv = rsqrt(d);
#pragma unroll
        for (int i = j + 1; i < 6; ++i) a[i * (i + 1) / 2 + j] *= inv;
#pragma unroll
        for (int i = j + 1; i < 6; ++i)
#pragma unroll
            for (int c = j + 1; c <= i; ++c)
                a[i * (i + 1) / 2 + c] = fma(-a[i * (i + 1) / 2 + j], a[c * (c + 1) / 2 + j], a[i * (i + 1) / 2 + c]);
        const double xj = y[j] * inv;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) y[i] = fma(-a[i * (i + 1) / 2 + j], xj, y[i]);
        if (Lout != nullptr && lane == j) {
            Lout[j * 6 + j] = d * inv;
#pragma unroll
            for (int i = j + 1; i < 6; ++i) Lout[i * 6 + j] = a[i * (i + 1) / 2 + j];
        }
        if (lane < 6) Liout[j * 6 + cc] = xj;
    }
    return ok;
}

// One equality-constrained solve on the faces `sg` selects + KKT guard.  All 32 lanes of the warp call it.
// Returns (warp-uniform) 1 if every foot passes the guard, 0 if not, -1 if a pivot was not positive.
// On return sm.E holds the forces (3 per foot, foot-major), sm.slot[k][12..17] the velocity costates.
template <int N>
__device__ int ric_sweep(const DevParams& P, RicWarp<N>& sm, unsigned conbits, const uint8_t (&sg)[RicWarp<N>::ROUNDS],
                         uint8_t (&nsg)[RicWarp<N>::ROUNDS], int lane) {
    using S = RicWarp<N>;
    constexpr int NF = S::NF, ROUNDS = S::ROUNDS;
    const double lin = P.dt / P.mass, dt = P.dt;

    // ---- E_k and beta_k, feet in parallel (4 lanes = the feet of one step)
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int t = lane + 32 * r, k = t >> 2, j = t & 3;
        Face fc;
        make_face(P, (conbits >> r) & 1u, sg[r], fc);
        double A[9];
        load_A<NF>(sm.A, t, A);
        double b[3][6], d[3];
        const double zx = fc.zx ? 1.0 : 0.0, zy = fc.zy ? 1.0 : 0.0, zz = fc.zz ? 1.0 : 0.0;
        b[0][0] = lin * zx; b[0][1] = 0.0; b[0][2] = 0.0;
        b[1][0] = 0.0; b[1][1] = lin * zy; b[1][2] = 0.0;
        b[2][0] = lin * fc.czx * zz; b[2][1] = lin * fc.czy * zz; b[2][2] = lin * zz;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            b[0][3 + q] = A[3 * q] * zx;
            b[1][3 + q] = A[3 * q + 1] * zy;
            b[2][3 + q] = (A[3 * q] * fc.czx + A[3 * q + 1] * fc.czy + A[3 * q + 2]) * zz;
        }
        d[0] = fc.dx; d[1] = fc.dy; d[2] = fc.dz;
        int e = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int c = 0; c <= a; ++c, ++e) {
                double v = d[0] * b[0][a] * b[0][c] + d[1] * b[1][a] * b[1][c] + d[2] * b[2][a] * b[2][c];
                v += shfl_xor_d(v, 1);
                v += shfl_xor_d(v, 2);
                if ((e & 3) == j) sm.E[21 * k + e] = v;
            }
        double ub[6];
        bv_apply(A, lin, fc.pf, ub);
        ub[2] -= (j == 0) ? P.gravity * dt : 0.0;             // g: only the z velocity, MPC.py:200-201
        step_sum_store(ub, sm.beta + 6 * k, j);
    }
    // ---- terminal cost-to-go: P_N = Q, p_N = -Q xref_N
    for (int i = lane; i < 36; i += 32) {
        const int rr = i / 6, dg = (i - 6 * rr) == rr;
        sm.Ppp[i] = dg ? P.wp[rr] : 0.0;
        sm.Pvv[i] = dg ? P.wv[rr] : 0.0;
        sm.Ppv[i] = 0.0;
    }
    if (lane < 6) {
        sm.pp[lane] = -P.wp[lane] * sm.xr[lane * (N + 1) + N];
        sm.pv[lane] = -P.wv[lane] * sm.xr[(6 + lane) * (N + 1) + N];
    }
    __syncwarp();

    // lane -> output maps of the 6x6 products: output `lane` (i0, c0), outputs 32..35 on lanes 0..3 (row 5),
    // lower-triangle output (si, sj) on lanes 0..20, vector component vv on lanes 26..31
    const int i0 = lane / 6, c0 = lane - 6 * i0;
    const bool has1 = lane < 4;
    const int c1 = has1 ? 2 + lane : 5;
    const int si = (lane >= 15) ? 5 : (lane >= 10) ? 4 : (lane >= 6) ? 3 : (lane >= 3) ? 2 : (lane >= 1) ? 1 : 0;
    const bool hsym = lane < 21;
    const int sj = hsym ? lane - si * (si + 1) / 2 : 0;
    const int sii = hsym ? si : 0;
    const bool hvec = lane >= 26;
    const int vv = hvec ? lane - 26 : 0;
    int ie0[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) ie0[q] = (i0 >= q) ? i0 * (i0 + 1) / 2 + q : q * (q + 1) / 2 + i0;

    bool spd = true;
    for (int k = N - 1; k >= 0; --k) {
        const double* Ek = sm.E + 21 * k;
        const double* bk = sm.beta + 6 * k;
        // (1) L L' = Pvv, Li = inv(L)
        spd = chol6_inv(sm.Pvv, sm.L, sm.Li, lane) && spd;
        __syncwarp();
        // (2) T = E L,  Y = Ppv Li',  av = Li pv
        {
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                t0 = fma(Ek[ie0[q]], sm.L[q * 6 + c0], t0);
                t1 = fma(Ek[15 + q], sm.L[q * 6 + c1], t1);
            }
            const double y0 = dot_rr(sm.Ppv + 6 * i0, sm.Li + 6 * c0);
            const double y1 = dot_rr(sm.Ppv + 30, sm.Li + 6 * c1);
            const double a_ = dot_rr(sm.Li + 6 * vv, sm.pv);
            sm.T[lane] = t0; sm.Y[lane] = y0;
            if (has1) { sm.T[32 + lane] = t1; sm.Y[32 + lane] = y1; }
            if (hvec) sm.av[vv] = a_;
        }
        __syncwarp();
        // (3) G = I + L' T
        {
            const double g_ = dot_cc(sm.L + sii, sm.T + sj) + ((sii == sj) ? 1.0 : 0.0);
            if (hsym) { sm.G[sii * 6 + sj] = g_; sm.G[sj * 6 + sii] = g_; }
        }
        __syncwarp();
        // (4) M M' = G, Mi = inv(M)
        spd = chol6_inv(sm.G, nullptr, sm.Mi, lane) && spd;
        __syncwarp();
        // (5) X = L Mi',  Y2 = Y Mi',  U = Mi Li,  cv = Mi av
        {
            const double x0 = dot_rr(sm.L + 6 * i0, sm.Mi + 6 * c0), x1 = dot_rr(sm.L + 30, sm.Mi + 6 * c1);
            const double y0 = dot_rr(sm.Y + 6 * i0, sm.Mi + 6 * c0), y1 = dot_rr(sm.Y + 30, sm.Mi + 6 * c1);
            const double u0 = dot_rc(sm.Mi + 6 * i0, sm.Li + c0), u1 = dot_rc(sm.Mi + 30, sm.Li + c1);
            const double c_ = dot_rr(sm.Mi + 6 * vv, sm.av);
            sm.X[lane] = x0; sm.Y2[lane] = y0; sm.U[lane] = u0;
            if (has1) { sm.X[32 + lane] = x1; sm.Y2[32 + lane] = y1; sm.U[32 + lane] = u1; }
            if (hvec) sm.cv[vv] = c_;
        }
        __syncwarp();
        // (6) Pt blocks, pt, and the pieces of the closed-loop gain:  KpT = Ppv Gamma (-> sm.G),  G2 = Gamma Pvv (-> sm.T)
        {
            const double tvv = dot_rr(sm.X + 6 * sii, sm.X + 6 * sj);
            const double tpp = sm.Ppp[sii * 6 + sj] - dot_rr(sm.Y + 6 * sii, sm.Y + 6 * sj) + dot_rr(sm.Y2 + 6 * sii, sm.Y2 + 6 * sj);
            const double tpv0 = dot_rr(sm.Y2 + 6 * i0, sm.X + 6 * c0), tpv1 = dot_rr(sm.Y2 + 30, sm.X + 6 * c1);
            const double kp0 = dot_rc(sm.Y + 6 * i0, sm.Li + c0) - dot_rc(sm.Y2 + 6 * i0, sm.U + c0);
            const double kp1 = dot_rc(sm.Y + 30, sm.Li + c1) - dot_rc(sm.Y2 + 30, sm.U + c1);
            const double g20 = ((i0 == c0) ? 1.0 : 0.0) - dot_rc(sm.X + 6 * c0, sm.U + i0);
            const double g21 = ((5 == c1) ? 1.0 : 0.0) - dot_rc(sm.X + 6 * c1, sm.U + 5);
            const double tv_ = dot_rr(sm.X + 6 * vv, sm.cv);
            const double tp_ = sm.pp[vv] - dot_rr(sm.Y + 6 * vv, sm.av) + dot_rr(sm.Y2 + 6 * vv, sm.cv);
            const double gm_ = dot_rc(sm.av, sm.Li + vv) - dot_rc(sm.cv, sm.U + vv);
            if (hsym) {
                sm.Tvv[sii * 6 + sj] = tvv; sm.Tvv[sj * 6 + sii] = tvv;
                sm.Tpp[sii * 6 + sj] = tpp; sm.Tpp[sj * 6 + sii] = tpp;
            }
            sm.Tpv[lane] = tpv0; sm.G[lane] = kp0; sm.T[lane] = g20;
            if (has1) { sm.Tpv[32 + lane] = tpv1; sm.G[32 + lane] = kp1; sm.T[32 + lane] = g21; }
            if (hvec) { sm.tv[vv] = tv_; sm.tp[vv] = tp_; sm.gam[vv] = gm_; }
        }
        __syncwarp();
        // (7) gain of this stage into its slot; cost-to-go of stage k (not needed for k = 0)
        {
            double* slot = sm.slot + RIC_SLOT * k;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int e = lane + 32 * q;
                if (e < 72) {
                    const int o = e / 12, i = e - 12 * o;
                    slot[e] = (i < 6) ? sm.G[i * 6 + o] : fma(dt, sm.G[(i - 6) * 6 + o], sm.T[o * 6 + i - 6]);
                }
            }
            if (hvec) slot[72 + vv] = dot_rr(sm.T + 6 * vv, bk) + sm.gam[vv];
            if (k > 0) {
                const double tpp = sm.Tpp[sii * 6 + sj];
                const double npp_ = tpp + ((sii == sj) ? P.wp[sii] : 0.0);
                const double nvv_ = dt * dt * tpp + dt * (sm.Tpv[sii * 6 + sj] + sm.Tpv[sj * 6 + sii]) + sm.Tvv[sii * 6 + sj]
                                    + ((sii == sj) ? P.wv[sii] : 0.0);
                const double npv0 = fma(dt, sm.Tpp[lane], sm.Tpv[lane]);
                const double npv1 = fma(dt, sm.Tpp[30 + c1], sm.Tpv[30 + c1]);
                const double hp = dot_rr(sm.Tpv + 6 * vv, bk) + sm.tp[vv];
                const double hv = dot_rr(sm.Tvv + 6 * vv, bk) + sm.tv[vv];
                if (hsym) {
                    sm.Ppp[sii * 6 + sj] = npp_; sm.Ppp[sj * 6 + sii] = npp_;
                    sm.Pvv[sii * 6 + sj] = nvv_; sm.Pvv[sj * 6 + sii] = nvv_;
                }
                sm.Ppv[lane] = npv0;
                if (has1) sm.Ppv[32 + lane] = npv1;
                if (hvec) {
                    sm.pp[vv] = hp - P.wp[vv] * sm.xr[vv * (N + 1) + k];
                    sm.pv[vv] = fma(dt, hp, hv) - P.wv[vv] * sm.xr[(6 + vv) * (N + 1) + k];
                }
            }
        }
        __syncwarp();
    }
    if (!__all_sync(0xffffffffu, spd)) return -1;

    // ---- forward pass: x replicated in the registers of every lane
    {
        double x[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) x[c] = sm.xr[c * (N + 1)];
        const int o = lane % 6;
        for (int k = 0; k < N; ++k) {
            double* slot = sm.slot + RIC_SLOT * k;
            const double* row = slot + 12 * o;
            double a0 = slot[72 + o], a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int i = 0; i < 12; i += 4) {
                a0 = fma(row[i], x[i], a0); a1 = fma(row[i + 1], x[i + 1], a1);
                a2 = fma(row[i + 2], x[i + 2], a2); a3 = fma(row[i + 3], x[i + 3], a3);
            }
            const double w = -((a0 + a1) + (a2 + a3));
            const double* bk = sm.beta + 6 * k;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const double wc = shfl_d(w, c);
                const double xp = fma(dt, x[6 + c], x[c]);
                x[6 + c] = x[6 + c] + bk[c] + wc;
                x[c] = xp;
            }
            __syncwarp();                       // everyone has read this stage's gain
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < 12; c += 2) *reinterpret_cast<double2*>(slot + c) = make_double2(x[c], x[c + 1]);
            }
        }
    }
    __syncwarp();
    // ---- costates of the velocities: lam_s = Q e_s + A' lam_{s+1}  ->  slot[s-1][12..17]
    if (lane < 6) {
        const int c = lane;
        double lp = 0.0, lv = 0.0;
        for (int s = N; s >= 1; --s) {
            const double* slot = sm.slot + RIC_SLOT * (s - 1);
            const double ep = slot[c] - sm.xr[c * (N + 1) + s], ev = slot[6 + c] - sm.xr[(6 + c) * (N + 1) + s];
            lv = fma(P.wv[c], ev, fma(dt, lp, lv));
            lp = fma(P.wp[c], ep, lp);
            sm.slot[RIC_SLOT * (s - 1) + 12 + c] = lv;
        }
    }
    __syncwarp();
    // ---- per foot: forces on the face, gradient, KKT guard
    bool ok = true;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int t = lane + 32 * r, k = t >> 2;
        const bool contact = (conbits >> r) & 1u;
        double f[3] = {0.0, 0.0, 0.0};
        nsg[r] = sg[r];
        if (contact) {
            Face fc;
            make_face(P, true, sg[r], fc);
            double A[9], h[3];
            load_A<NF>(sm.A, t, A);
            bvT_apply(A, lin, sm.slot + RIC_SLOT * k + 12, h);
            const double qx = fc.zx ? -fc.dx * h[0] : 0.0;
            const double qy = fc.zy ? -fc.dy * h[1] : 0.0;
            const double qz = fc.zz ? -fc.dz * (fc.czx * h[0] + fc.czy * h[1] + h[2]) : 0.0;
            f[0] = fc.pf[0] + qx + fc.czx * qz;
            f[1] = fc.pf[1] + qy + fc.czy * qz;
            f[2] = fc.pf[2] + qz;
            const double grad[3] = {fma(P.w_force, f[0], h[0]), fma(P.w_force, f[1], h[1]), fma(P.w_force, f[2], h[2])};
            FootSol sol;
            ok = kkt_guard(P, sg[r], f, grad, sol, nsg[r]) && ok;
        }
        sm.E[3 * t] = f[0]; sm.E[3 * t + 1] = f[1]; sm.E[3 * t + 2] = f[2];
    }
    const int all_ok = __all_sync(0xffffffffu, ok);
    return all_ok ? 1 : 0;
}

// Outputs of one robot (the warp version of finish() in mpcqp_kernels.cu)                 [MPC.py:432-458]
template <int N>
__device__ void ric_finish(const DevParams& P, const DevScenario& SC, RicWarp<N>& sm, const DevState& st, int inst, unsigned conbits,
                           const uint8_t (&sg)[RicWarp<N>::ROUNDS], bool solved, int status, int sweeps, int lane) {
    using S = RicWarp<N>;
    constexpr int NF = S::NF, ROUNDS = S::ROUNDS, AW = S::AW, CW = S::CW;
    const double lin = P.dt / P.mass;
    for (int i = lane; i < AW + CW; i += 32) sm.amask[i] = 0u;
    // impulses of the final forces -> sm.beta
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int t = lane + 32 * r, k = t >> 2, j = t & 3;
        double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (solved && ((conbits >> r) & 1u)) {
            double A[9];
            load_A<NF>(sm.A, t, A);
            const double f[3] = {sm.E[3 * t], sm.E[3 * t + 1], sm.E[3 * t + 2]};
            bv_apply(A, lin, f, v);
        }
        step_sum_store(v, sm.beta + 6 * k, j);
    }
    __syncwarp();
    double part = 0.0;
    if (lane < 6) {
        // component c: p_{s+1} = p_s + dt v_s, v_{s+1} = v_s + u_s + g_c                   (MPC.py:110-111, 200-205)
        const int c = lane;
        double p = sm.xr[c * (N + 1)], v = sm.xr[(6 + c) * (N + 1)];
        const double gc = (c == 2) ? -P.gravity * P.dt : 0.0;
        double* xs = st.xs + (size_t)inst * 12 * N;
        for (int s = 0; s < N; ++s) {
            const double pn = p + P.dt * v;
            const double vn = v + sm.beta[6 * s + c] + gc;
            p = pn; v = vn;
            if (s == 0) { sm.sc.xnext[c] = p; sm.sc.xnext[6 + c] = v; }                     // MPC.q_next / v_next (MPC.py:448-450)
            const double ep = p - sm.xr[c * (N + 1) + s + 1], ev = v - sm.xr[(6 + c) * (N + 1) + s + 1];
            xs[12 * s + c] = ep;
            xs[12 * s + 6 + c] = ev;
            part += 0.5 * (P.wp[c] * ep * ep + P.wv[c] * ev * ev);
        }
    }
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int t = lane + 32 * r, k = t >> 2, j = t & 3;
        const bool contact = (conbits >> r) & 1u;
        double f[3] = {0.0, 0.0, 0.0};
        FootSol sol;
#pragma unroll
        for (int q = 0; q < 5; ++q) sol.y[q] = 0.0;
        if (solved && contact) {
            // multipliers: the guard's closed form on the gradient of the accepted sweep
            f[0] = sm.E[3 * t]; f[1] = sm.E[3 * t + 1]; f[2] = sm.E[3 * t + 2];
            double A[9], h[3];
            load_A<NF>(sm.A, t, A);
            bvT_apply(A, lin, sm.slot + RIC_SLOT * k + 12, h);
            const double grad[3] = {fma(P.w_force, f[0], h[0]), fma(P.w_force, f[1], h[1]), fma(P.w_force, f[2], h[2])};
            uint8_t dummy;
            kkt_guard(P, sg[r], f, grad, sol, dummy);
        }
        part += 0.5 * P.w_force * (f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        double* fo = st.f + (size_t)inst * 12 * N + 3 * t;
        fo[0] = f[0]; fo[1] = f[1]; fo[2] = f[2];
        double* yo = st.y + (size_t)inst * 20 * N + 5 * t;
#pragma unroll
        for (int q = 0; q < 5; ++q) yo[q] = sol.y[q];
        st.sig[(size_t)inst * NF + t] = sg[r];
        if (k == 0) {
            double* f0 = st.f0 + (size_t)inst * 12 + 3 * j;
            f0[0] = f[0]; f0[1] = f[1]; f0[2] = f[2];
        }
        // rows that hold with equality; a swing foot is pinned to f = 0 (MPC.py:355-358): all five of its rows do
        const double mu = P.mu, tol = 1e-9;
        const double row[5] = {f[0] - mu * f[2], -f[0] - mu * f[2], f[1] - mu * f[2], -f[1] - mu * f[2], -f[2]};
        const int b0 = 5 * t;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const bool act = (fabs(row[q]) <= tol) || (q == 4 && fabs(row[4] + P.fz_max) <= tol);
            if (act) atomicOr(&sm.amask[(b0 + q) >> 5], 1u << ((b0 + q) & 31));
        }
        const unsigned cb = __ballot_sync(0xffffffffu, contact);
        if (lane == 0) sm.amask[AW + r] = cb;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __syncwarp();
    for (int i = lane; i < AW; i += 32) st.active[(size_t)inst * AW + i] = sm.amask[i];
    for (int i = lane; i < CW; i += 32) st.contact[(size_t)inst * CW + i] = sm.amask[AW + i];
    if (lane == 0) {
        st.obj[inst] = part;
        st.status[inst] = status;
        st.sweeps[inst] = sweeps;
        st.iters[inst] = 0;
        if (SC.enabled && status != 3) scenario_advance(SC, inst, sm.sc.xnext);
    }
}

// The active-set stage, stage-wise factorisation: grid = ceil(instances / RIC_WARPS), one warp per robot.
template <int N>
__global__ void __launch_bounds__(32 * RIC_WARPS)
riccati_kernel(DevParams P, DevState st, DevScenario SC, const double* __restrict__ xref_g, const double* __restrict__ fsteps_g,
               int first_tick, int inst_offset, int inst_count) {
    using S = RicWarp<N>;
    constexpr int NF = S::NF, ROUNDS = S::ROUNDS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    S& sm = reinterpret_cast<S*>(smem_raw)[warp];
    if (lane == 0) mbar_init(&sm.mbar, 1);
    for (int i = lane; i < 36; i += 32) sm.L[i] = 0.0;          // the strictly upper triangle of L stays zero
    __syncwarp();
    unsigned int phase = 0;
    for (int w = blockIdx.x * RIC_WARPS + warp; w < inst_count; w += gridDim.x * RIC_WARPS) {
        const int inst = w + inst_offset;
        __syncwarp();
        if (SC.enabled) {
            scenario_inputs<N, true>(P, SC, sm.sc, inst, sm.xr, sm.fs);
        } else {
            if (lane == 0) {
                fence_async_smem();
                mbar_expect_tx(&sm.mbar, (12 * (N + 1) + 260) * 8);
                bulk_g2s(sm.xr, xref_g + (size_t)inst * 12 * (N + 1), 12 * (N + 1) * 8, &sm.mbar);
                bulk_g2s(sm.fs, fsteps_g + (size_t)inst * 260, 260 * 8, &sm.mbar);
            }
            mbar_wait(&sm.mbar, phase);
            phase ^= 1u;
        }
        // ---- decode: contact flags, lever-arm blocks, warm-start signature      [MPC.py:316-360, 403-406, 635-652]
        const bool warm = P.warm_start && !first_tick;
        bool bad = false;
        unsigned conbits = 0u;
        uint8_t sg[ROUNDS], nsg[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int t = lane + 32 * r, k = t >> 2, j = t & 3;
            double A0[9];
            bool contact = false;
            decode_foot<N>(P, sm.xr, sm.fs, k, j, first_tick != 0, A0, contact, bad);
#pragma unroll
            for (int i = 0; i < 9; ++i) sm.A[i * NF + t] = A0[i];
            conbits |= contact ? (1u << r) : 0u;
            sg[r] = SIG_FREE;
            if (warm && contact) {
                const int ks = (k + 1 < N) ? k + 1 : 0;
                const uint8_t s8 = st.sig[(size_t)inst * NF + 4 * ks + j];
                sg[r] = s8 > 26 ? SIG_FREE : s8;
            }
        }
        for (int i = lane; i < 12 * (N + 1); i += 32) bad = bad || !isfinite(sm.xr[i]);
        const bool any_bad = __any_sync(0xffffffffu, bad);       // also orders the reads of fs before E is written
        int sweeps = 0, status = 0;
        bool done = false;
        if (any_bad) {
            status = 3;
            conbits = 0u;
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) sg[r] = SIG_FREE;
        } else {
            int nhist = 0;
            for (int s = 0; s < P.max_sweeps && !done; ++s) {
                unsigned long long h = 0ull;
#pragma unroll
                for (int r = 0; r < ROUNDS; ++r) {
                    if ((conbits >> r) & 1u) {
                        unsigned long long q = (unsigned long long)(sg[r] + 1) * 0x9E3779B97F4A7C15ull;
                        q ^= q >> 29; q *= (2ull * (lane + 32 * r) + 0xBF58476D1CE4E5B9ull); q ^= q >> 32;
                        h += q;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
                bool seen = false;
                for (int i = 0; i < nhist; ++i) seen = seen || (sm.hist[i] == h);
                if (seen) break;
                __syncwarp();
                if (lane == 0 && nhist < 16) sm.hist[nhist] = h;
                nhist = (nhist < 16) ? nhist + 1 : nhist;
                __syncwarp();
                const int rc = ric_sweep<N>(P, sm, conbits, sg, nsg, lane);
                ++sweeps;
                if (rc < 0) break;
                if (rc > 0) { done = true; status = 1; }
                else {
#pragma unroll
                    for (int r = 0; r < ROUNDS; ++r) sg[r] = nsg[r];
                }
            }
            if (!done) {
                if (P.mode & 2) {
                    if (lane == 0) {
                        const int slot = atomicAdd(st.fb_count, 1);
                        st.fb_list[slot] = inst;
                        st.sweeps[inst] = sweeps;
                    }
                    continue;       // state of this robot is left untouched for the ADMM stage
                }
                status = 0;
            }
        }
        ric_finish<N>(P, SC, sm, st, inst, conbits, sg, done, status, sweeps, lane);
    }
}

}  // namespace mpcqp
