// mpcqp_foot.cuh -- per-foot building blocks shared by the dense (CTA per robot) and the stage-wise
// (warp per robot) solve paths: contact decoding and lever-arm blocks (MPC.py:316-360, 635-652), the affine
// face a signature selects, the force <-> impulse maps, and the KKT guard of the active-set sweeps.
#pragma once
#include "mpcqp_device.cuh"

namespace mpcqp {

// per-foot description of the affine face f = pf + Z q selected by one signature:
// Z has the columns ex (if zx), ey (if zy), (czx, czy, 1) (if zz); D = w_f Z'Z is diagonal.
struct Face {
    double dx, dy, dz;              // 1 / (w_f |z_col|^2), or 0 when the column is absent
    double pf[3];
    double czx, czy;                // sx mu, sy mu
    bool zx, zy, zz;
};

__device__ __forceinline__ void make_face(const DevParams& P, bool contact, uint8_t sig, Face& fc) {
    int sx, sy, tz;
    sig_unpack(sig, sx, sy, tz);
    const bool live = contact && tz != 1;
    fc.zx = live && sx == 0;
    fc.zy = live && sy == 0;
    fc.zz = live && tz == 0;
    fc.czx = sx * P.mu;
    fc.czy = sy * P.mu;
    fc.dx = fc.zx ? P.inv_wf : 0.0;                              // 1 / w
    fc.dy = fc.zy ? P.inv_wf : 0.0;
    fc.dz = fc.zz ? P.dz_tab[sx * sx + sy * sy] : 0.0;           // 1 / (w (1 + mu^2 (sx^2 + sy^2)))
    const bool top = live && tz == 2;
    fc.pf[0] = top ? fc.czx * P.fz_max : 0.0;
    fc.pf[1] = top ? fc.czy * P.fz_max : 0.0;
    fc.pf[2] = top ? P.fz_max : 0.0;
}

template <int NF>
__device__ __forceinline__ void load_A(const double* fa, int t, double A[9]) {
#pragma unroll
    for (int i = 0; i < 9; ++i) A[i] = fa[i * NF + t];
}

// Bv f: rows 0..2 = (dt/m) f, rows 3..5 = A f
__device__ __forceinline__ void bv_apply(const double A[9], double lin, const double f[3], double out[6]) {
    out[0] = lin * f[0]; out[1] = lin * f[1]; out[2] = lin * f[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) out[3 + r] = A[3 * r] * f[0] + A[3 * r + 1] * f[1] + A[3 * r + 2] * f[2];
}
__device__ __forceinline__ void bvT_apply(const double A[9], double lin, const double* v, double out[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = lin * v[c] + A[c] * v[3] + A[3 + c] * v[4] + A[6 + c] * v[5];
}

// sum a 6-vector over the four feet of a step (lanes 4k..4k+3) and let lane j == 0 store it.
// Out of line (code size): called from every per-foot phase.
static __device__ __noinline__ void step_sum_store6(double v0, double v1, double v2, double v3, double v4, double v5, double* dst, int j) {
    v0 += shfl_xor_d(v0, 1); v1 += shfl_xor_d(v1, 1); v2 += shfl_xor_d(v2, 1);
    v3 += shfl_xor_d(v3, 1); v4 += shfl_xor_d(v4, 1); v5 += shfl_xor_d(v5, 1);
    v0 += shfl_xor_d(v0, 2); v1 += shfl_xor_d(v1, 2); v2 += shfl_xor_d(v2, 2);
    v3 += shfl_xor_d(v3, 2); v4 += shfl_xor_d(v4, 2); v5 += shfl_xor_d(v5, 2);
    if (j == 0) { dst[0] = v0; dst[1] = v1; dst[2] = v2; dst[3] = v3; dst[4] = v4; dst[5] = v5; }
}
__device__ __forceinline__ void step_sum_store(double v[6], double* dst, int j) {
    step_sum_store6(v[0], v[1], v[2], v[3], v[4], v[5], dst, j);
}

// -------------------------------------------------------------------------------------------------
// decode: contact flag, foothold, lever arm block for (step k, foot j)      [MPC.py:316-360, 635-652]
// -------------------------------------------------------------------------------------------------
// N is the horizon (xref is 12 x (N + 1)); the dense kernels pass their compile-time horizon
__device__ __forceinline__ void decode_foot(const DevParams& P, const double* xr, const double* fs, const int N, int k, int j,
                                            bool first_tick, double A[9], bool& contact, bool& bad) {
    int row = -1;
    double cum = 0.0;
    for (int r = 0; r < 20; ++r) {
        const double cnt = fs[r * 13];
        if (cnt == 0.0) break;                       // MPC.py:646: first empty row ends the table
        if (!(cnt > 0.0) || cnt != floor(cnt)) { bad = true; break; }
        if ((double)k < cum + cnt) { row = r; break; }
        cum += cnt;
    }
    double foot[3] = {0.0, 0.0, 0.0};
    contact = false;
    if (row >= 0) {
        const double x = fs[row * 13 + 1 + 3 * j];
        contact = !(isnan(x) || x == 0.0);           // MPC.py:650
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double v = fs[row * 13 + 1 + 3 * j + c];
            foot[c] = isnan(v) ? 0.0 : v;            // MPC.py:327
        }
    }
    if (first_tick) {                                 // MPC.py:176: tick 0 uses the default footholds
#pragma unroll
        for (int c = 0; c < 3; ++c) foot[c] = P.footholds[c * 4 + j];
    }
    double r[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) r[c] = foot[c] - xr[c * (N + 1) + k];      // MPC.py:343
    double sn, cs;
    sincos(xr[5 * (N + 1) + k], &sn, &cs);                                   // MPC.py:330
    // inv(R gI) = gI^-1 R'   (MPC.py:339-340: the reference inverts R gI, not R gI R')
    double Ii[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        Ii[3 * a + 0] = P.gIinv[3 * a + 0] * cs - P.gIinv[3 * a + 1] * sn;
        Ii[3 * a + 1] = P.gIinv[3 * a + 0] * sn + P.gIinv[3 * a + 1] * cs;
        Ii[3 * a + 2] = P.gIinv[3 * a + 2];
    }
    // dt * Ii * [r]x   (MPC.py:345-346, utils.py:179-185)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        A[3 * a + 0] = P.dt * (Ii[3 * a + 1] * r[2] - Ii[3 * a + 2] * r[1]);
        A[3 * a + 1] = P.dt * (Ii[3 * a + 2] * r[0] - Ii[3 * a + 0] * r[2]);
        A[3 * a + 2] = P.dt * (Ii[3 * a + 0] * r[1] - Ii[3 * a + 1] * r[0]);
    }
}

// The same decode split in two, for the stage-wise path (it keeps the lever arm and the per-step inertia
// block instead of the 3x3 product, and forms  A = dt inv(R gI) [r]x  on demand with decode_foot's arithmetic):
//   contact flag + lever arm r = foothold - xref[0:3, k] of (step k, foot j)            [MPC.py:327, 343, 635-652]
// `n` is the run-time horizon: xref is 12 x (n + 1).
// A lane asks for increasing steps k: (q, cum, row) carry the row search of decode_foot from one call to the next (rows < q have
// been passed and validated, cum = their counts; row = the row of the last k, -1 if there was none, -2 once the table has ended) --
// the rows are visited and validated in the same order as by a search that starts over, so contact, lever arm and `bad` are the same.
__device__ __forceinline__ void decode_lever_next(const DevParams& P, const double* xr, const double* fs, int n, int k, int j, bool first_tick,
                                                  double r[3], bool& contact, bool& bad, int& q, double& cum, int& row) {
    if (row != -2 && !bad) {
        row = -1;
        for (; q < 20; ++q) {
            const double cnt = fs[q * 13];
            if (cnt == 0.0) { row = -2; break; }
            if (!(cnt > 0.0) || cnt != floor(cnt)) { bad = true; break; }
            if ((double)k < cum + cnt) { row = q; break; }
            cum += cnt;
        }
    }
    double foot[3] = {0.0, 0.0, 0.0};
    contact = false;
    if (row >= 0) {
        const double x = fs[row * 13 + 1 + 3 * j];
        contact = !(isnan(x) || x == 0.0);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double v = fs[row * 13 + 1 + 3 * j + c];
            foot[c] = isnan(v) ? 0.0 : v;
        }
    }
    if (first_tick) {
#pragma unroll
        for (int c = 0; c < 3; ++c) foot[c] = P.footholds[c * 4 + j];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) r[c] = foot[c] - xr[c * (n + 1) + k];
}
//   inv(R_z(yaw) gI) = gI^-1 R'                                                          [MPC.py:330, 339-340]
__device__ __forceinline__ void step_inertia(const DevParams& P, double yaw, double Ii[9]) {
    double sn, cs;
    sincos(yaw, &sn, &cs);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        Ii[3 * a + 0] = P.gIinv[3 * a + 0] * cs - P.gIinv[3 * a + 1] * sn;
        Ii[3 * a + 1] = P.gIinv[3 * a + 0] * sn + P.gIinv[3 * a + 1] * cs;
        Ii[3 * a + 2] = P.gIinv[3 * a + 2];
    }
}
//   A = dt Ii [r]x                                                                       [MPC.py:345-346, utils.py:179-185]
__device__ __forceinline__ void lever_block(const DevParams& P, const double Ii[9], const double r[3], double A[9]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        A[3 * a + 0] = P.dt * (Ii[3 * a + 1] * r[2] - Ii[3 * a + 2] * r[1]);
        A[3 * a + 1] = P.dt * (Ii[3 * a + 2] * r[0] - Ii[3 * a + 0] * r[2]);
        A[3 * a + 2] = P.dt * (Ii[3 * a + 0] * r[1] - Ii[3 * a + 1] * r[0]);
    }
}

// per-foot results of one sweep
struct FootSol {
    double f[3];
    double y[5];
};

// Multipliers, KKT guard and the next active-set guess for one stance foot whose equality-constrained
// solution is f with condensed gradient grad = H f + g.  Returns true if the foot passes; f is snapped to the
// apex when the signature says so.  (Rows of C: +-fx - mu fz, +-fy - mu fz, -fz; MPC.py:136-148, 226-229.)
__device__ __forceinline__ bool kkt_guard(const DevParams& P, uint8_t sig, double f[3], const double grad[3], FootSol& sol,
                                          uint8_t& nsig) {
    bool ok = true;
#pragma unroll
    for (int r = 0; r < 5; ++r) sol.y[r] = 0.0;
    int sx, sy, tz;
    sig_unpack(sig, sx, sy, tz);
    const double mu = P.mu, ytol = P.dual_tol, ftol = P.feas_tol;
    int nsx = sx, nsy = sy, ntz = tz;
    if (tz == 1) {
        // apex: need y >= 0 with C' y = -grad; the sign of the slack on the fz >= 0 row decides
        const double qx = -grad[0], qy = -grad[1], qz = -grad[2];
        sol.y[0] = fmax(qx, 0.0); sol.y[1] = fmax(-qx, 0.0);
        sol.y[2] = fmax(qy, 0.0); sol.y[3] = fmax(-qy, 0.0);
        sol.y[4] = -qz - mu * (fabs(qx) + fabs(qy));
        f[0] = f[1] = f[2] = 0.0;
        if (sol.y[4] < -ytol) {
            // leave the apex by releasing the fz >= 0 row only: the friction rows the gradient pushes against stay active
            ok = false;
            nsx = (qx > ytol) ? 1 : ((qx < -ytol) ? -1 : 0);
            nsy = (qy > ytol) ? 1 : ((qy < -ytol) ? -1 : 0);
            ntz = 0;
        }
    } else {
        const double yx = (sx != 0) ? -sx * grad[0] : 0.0;
        const double yy = (sy != 0) ? -sy * grad[1] : 0.0;
        if (sx > 0) sol.y[0] = yx; else if (sx < 0) sol.y[1] = yx;
        if (sy > 0) sol.y[2] = yy; else if (sy < 0) sol.y[3] = yy;
        const double y4 = grad[2] - mu * (yx + yy);
        if (tz == 2) {
            sol.y[4] = y4;
            if (y4 > ytol) { ok = false; ntz = 0; }
        }
        if (sx != 0 && yx < -ytol) { ok = false; nsx = 0; }
        if (sy != 0 && yy < -ytol) { ok = false; nsy = 0; }
        if (sx == 0) {
            if (f[0] - mu * f[2] > ftol) { ok = false; nsx = 1; }
            else if (-f[0] - mu * f[2] > ftol) { ok = false; nsx = -1; }
        }
        if (sy == 0) {
            if (f[1] - mu * f[2] > ftol) { ok = false; nsy = 1; }
            else if (-f[1] - mu * f[2] > ftol) { ok = false; nsy = -1; }
        }
        if (tz == 0) {
            if (f[2] > P.fz_max + ftol) { ok = false; ntz = 2; }
            else if (f[2] < -ftol) { ok = false; ntz = 1; }
        }
    }
    nsig = sig_pack(nsx, nsy, ntz);
#pragma unroll
    for (int c = 0; c < 3; ++c) sol.f[c] = f[c];
    return ok;
}

}  // namespace mpcqp
