// mpcqp_scenario.cuh -- device-resident closed loop (SURVEY.md 8f rows f1 + f2).
//
// Produces the MPC inputs of one robot inside the solve kernel (no HBM round trip, no host producer)
// and advances the robot after the solve.  It is the device restatement of mpc-tsid_b200/scenario.py,
// which in turn restates the reference's planner:
//     gait roll / run-length table       FootstepPlanner.py:401-425
//     compute_footsteps                  FootstepPlanner.py:284-361
//     compute_next_footstep              FootstepPlanner.py:363-399
//     getRefStates                       FootstepPlanner.py:76-161   (with the vz / roll / pitch command state machine, :128-152)
// and closes the loop on the centroidal model: the next measured state is the MPC's one-step
// prediction (MPC.py:448-450) re-expressed in the next yaw-aligned local frame (Interface.py:100-132)
// plus counter-based Gaussian noise (same generator as scenario.py's noise_kind="hash").
#pragma once
#include "mpcqp_device.cuh"

namespace mpcqp {

struct DevScenario {
    double* state;          // B x 12  measured state in the local frame
    double* frame;          // B x 3   local frame in the world (x, y, yaw)
    double* feet;           // B x 8   feet in the world, [x0..x3, y0..y3]
    double* target;         // B x 8   where each swinging foot is planned to land (world)
    const double* vref;     // B x 6   joystick commands (mpcqp_scenario_set_commands replaces them between ticks)
    double* cmd;            // B x 3   getRefStates' command state: h_rotation_command, and the xref[2, 1:] / xref[8, 1:] it leaves behind
    uint8_t* cmd_flag;      // B       flag_rotation_command (0 idle, 1 commanding, 2 released), FootstepPlanner.py:66, 128-152
    int32_t* cmd_tick;      // B       tick whose command step has been applied (a robot's inputs may be rebuilt by the fallback stage)
    const unsigned long long* seq;   // B x seq_words: bit 4 (s % 16) + j of word s / 16 = foot j in contact at step s of the gait period
    const int32_t* phase;   // B
    uint8_t* prevc;         // B: bits 0..3 contact of the previous tick's first step, bit 7 = valid
    double* xref_out;       // optional B x 12 x (N+1): the inputs generated this tick (parity hook), or null
    double* fsteps_out;     // optional B x 20 x 13
    double sigma[4];        // noise: position, angle, linear velocity, angular velocity
    double lin_a[64];       // numpy.linspace(0, T_gait - dt, N)      (FootstepPlanner.py:95)
    double lin_b[64];       // numpy.linspace(dt, T_gait, N)          (FootstepPlanner.py:114)
    unsigned long long seed;
    int tick;               // closed-loop tick of this launch
    int period;             // steps per gait period (T_gait / dt), <= 64
    int seq_words;          // 64-bit words of `seq` per robot: ceil(period / 16)
    int reduced;            // Joystick.reduced: the smaller support polygon of FootstepPlanner.py:330-332
    int enabled;
};

// shoulder positions, FootstepPlanner.py:23-24 (feet FL, FR, HL, HR)
__host__ __device__ constexpr double sc_shoulder_x(int j) { return j < 2 ? 0.19 : -0.19; }
__host__ __device__ constexpr double sc_shoulder_y(int j) { return (j & 1) ? -0.15005 : 0.15005; }
constexpr double SC_H_REF = 0.2027682, SC_K_FEEDBACK = 0.03, SC_LEG_L = 0.12, SC_T_STANCE = 0.16, SC_G = 9.81;
constexpr double SC_CMD_STEP = 0.05, SC_H_ROTATION0 = 0.20;       // FootstepPlanner.py:131, 69

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// standard normal, a pure function of (seed, instance, tick, component)
__device__ inline double hash_normal(unsigned long long seed, unsigned long long inst, unsigned long long tick, unsigned long long comp) {
    const unsigned long long key = splitmix64(seed + inst * 0x9E3779B97F4A7C15ull + tick * 0xD1B54A32D192ED03ull + comp * 0x8CB92BA72F3D8DD7ull);
    const unsigned long long h1 = splitmix64(key), h2 = splitmix64(key ^ 0xA5A5A5A5A5A5A5A5ull);
    const double u1 = ((double)(h1 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)(h2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

// scratch the planner needs in shared memory
struct ScenarioSmem {
    double st[12], fr[3], vr[6];
    double vx[64], vy[64];
    double xnext[12];
    int cnt[20], mask[20];
    int nrows;
    double zref, vzref;     // xref[2, 1:], xref[8, 1:] of this tick
    int flag;
};

// Build xref (12 x (N+1)) and fsteps (20 x 13) of instance `inst` in shared memory.  Called by every thread of
// the group that owns the instance: the whole CTA (GROUP = 0), one warp (GROUP = 32) or half a warp (GROUP = 16).
// N is the run-time horizon (xref is 12 x (N + 1)).
template <int GROUP = 0>
__device__ void scenario_inputs(const DevParams& P, const DevScenario& S, ScenarioSmem& sc, int inst, double* xr, double* fs,
                                const int N, bool commit = true) {
    constexpr bool WARP = GROUP != 0;
    const int tid = WARP ? (int)(threadIdx.x & (GROUP - 1)) : (int)threadIdx.x;
    const int nthr = WARP ? GROUP : (int)blockDim.x;
    // GROUP = 16: both halves of the warp call this together (one instruction stream), so the barrier is the full warp's
    auto group_sync = [] { if (WARP) __syncwarp(); else __syncthreads(); };
    const double nanv = __longlong_as_double(0x7ff8000000000000ll);
    for (int i = tid; i < 260; i += nthr) fs[i] = (i % 13 == 0) ? 0.0 : nanv;
    for (int i = tid; i < 12 * (N + 1); i += nthr) xr[i] = 0.0;
    if (tid < 12) sc.st[tid] = S.state[(size_t)inst * 12 + tid];
    if (tid < 3) sc.fr[tid] = S.frame[(size_t)inst * 3 + tid];
    if (tid < 6) sc.vr[tid] = S.vref[(size_t)inst * 6 + tid];
    group_sync();
    const double w = sc.vr[5];
    if (tid == 0) {
        // run-length table of the next N steps of the periodic gait (what FootstepPlanner.roll maintains)
        unsigned long long seq[4] = {0ull, 0ull, 0ull, 0ull};
        for (int q = 0; q < S.seq_words; ++q) seq[q] = S.seq[(size_t)inst * S.seq_words + q];
        const int ph = S.phase[inst];
        int rows = 0, prev = -1;
        for (int i = 0; i < N; ++i) {
            const int s = (S.tick + ph + i) % S.period;
            const int m = (int)((seq[s >> 4] >> (4 * (s & 15))) & 15ull);
            if (m != prev) { sc.mask[rows] = m; sc.cnt[rows] = 1; ++rows; prev = m; }
            else sc.cnt[rows - 1] += 1;
        }
        sc.nrows = rows;
        for (int r = 0; r < rows; ++r) fs[r * 13] = (double)sc.cnt[r];
        // the command state machine of getRefStates (FootstepPlanner.py:128-152), advanced once per tick and robot
        int flag = S.cmd_flag[inst];
        double hrot = S.cmd[(size_t)inst * 3], zref = S.cmd[(size_t)inst * 3 + 1], vzref = S.cmd[(size_t)inst * 3 + 2];
        if (S.cmd_tick[inst] != S.tick) {
            const double vz = sc.vr[2];
            const bool big = fabs(vz) > SC_CMD_STEP, small = fabs(vz) < SC_CMD_STEP;
            if (big && flag != 1) flag = 1;
            if (big && flag == 1) { hrot += vz * P.dt; zref = hrot; vzref = vz; }
            else if (small && flag == 1) { vzref = 0.0; flag = 2; }
            else if (flag == 0) { zref = SC_H_REF; vzref = 0.0; }
            if (commit) {
                S.cmd[(size_t)inst * 3] = hrot; S.cmd[(size_t)inst * 3 + 1] = zref; S.cmd[(size_t)inst * 3 + 2] = vzref;
                S.cmd_flag[inst] = (uint8_t)flag;
                S.cmd_tick[inst] = S.tick;
            }
        }
        sc.zref = zref; sc.vzref = vzref; sc.flag = flag;
    }
    for (int i = WARP ? tid : tid - 32; i >= 0 && i < N; i += WARP ? GROUP : N) {
        double sn, cs;
        sincos(S.lin_a[i] * w, &sn, &cs);                                  // FootstepPlanner.py:95-97
        sc.vx[i] = sc.vr[0] * cs - sc.vr[1] * sn;
        sc.vy[i] = sc.vr[0] * sn + sc.vr[1] * cs;
    }
    group_sync();
    if (tid < 4) {
        const int j = tid;
        double fwx = S.feet[(size_t)inst * 8 + j], fwy = S.feet[(size_t)inst * 8 + 4 + j];
        double twx = S.target[(size_t)inst * 8 + j], twy = S.target[(size_t)inst * 8 + 4 + j];
        const uint8_t pc = S.prevc[inst];
        const bool c0 = (sc.mask[0] >> j) & 1;
        if ((pc & 0x80) && c0 && !((pc >> j) & 1)) { fwx = twx; fwy = twy; }     // touchdown on the planned target
        double sf, cf;
        sincos(sc.fr[2], &sf, &cf);
        const double ddx = fwx - sc.fr[0], ddy = fwy - sc.fr[1];
        const double lx = cf * ddx + sf * ddy, ly = -sf * ddx + cf * ddy;
        double cur[3] = {lx, ly, 0.0};
        if (c0) { fs[1 + 3 * j] = lx; fs[2 + 3 * j] = ly; fs[3 + 3 * j] = 0.0; }
        // compute_next_footstep(v_ref, v_ref, h): symmetry + (vanishing) feedback + centrifugal, clamped, + shoulder
        const double h = sc.st[2];
        const double crx = sc.vr[1] * sc.vr[5] - sc.vr[2] * sc.vr[4], cry = sc.vr[2] * sc.vr[3] - sc.vr[0] * sc.vr[5];
        const double kc = 0.5 * sqrt(h / SC_G);
        double nfx = SC_T_STANCE * 0.5 * sc.vr[0] + SC_K_FEEDBACK * (sc.vr[0] - sc.vr[0]) + kc * crx;
        double nfy = SC_T_STANCE * 0.5 * sc.vr[1] + SC_K_FEEDBACK * (sc.vr[1] - sc.vr[1]) + kc * cry;
        nfx = fmin(fmax(nfx, -SC_LEG_L), SC_LEG_L) + sc_shoulder_x(j);
        nfy = fmin(fmax(nfy, -SC_LEG_L), SC_LEG_L) + sc_shoulder_y(j);
        if (S.reduced) { nfx -= (j < 2) ? 0.14 : -0.14; nfy -= (j & 1) ? -0.12 : 0.12; }        // FootstepPlanner.py:330-332
        const double vcx = sc.st[6], vcy = sc.st[7];
        double dt_cum = 0.0;
        bool got = false;
        bool prev_st = c0;
        for (int r = 1; r < sc.nrows; ++r) {
            dt_cum = dt_cum + (double)sc.cnt[r - 1] * P.dt;
            const bool stn = (sc.mask[r] >> j) & 1;
            if (stn) {
                if (!prev_st) {
                    double sa, ca;
                    sincos(w * dt_cum, &sa, &ca);
                    double dx, dy;
                    if (w != 0.0) {
                        dx = (vcx * sa + vcy * (ca - 1.0)) / w;
                        dy = (vcy * sa - vcx * (ca - 1.0)) / w;
                    } else {
                        dx = vcx * dt_cum;
                        dy = vcy * dt_cum;
                    }
                    cur[0] = ca * nfx - sa * nfy + dx;
                    cur[1] = sa * nfx + ca * nfy + dy;
                    cur[2] = 0.0;
                }
                fs[r * 13 + 1 + 3 * j] = cur[0]; fs[r * 13 + 2 + 3 * j] = cur[1]; fs[r * 13 + 3 + 3 * j] = cur[2];
                if (!got) {
                    got = true;
                    twx = cf * cur[0] - sf * cur[1] + sc.fr[0];
                    twy = sf * cur[0] + cf * cur[1] + sc.fr[1];
                }
            }
            prev_st = stn;
        }
        if (commit) {
            S.feet[(size_t)inst * 8 + j] = fwx; S.feet[(size_t)inst * 8 + 4 + j] = fwy;
            S.target[(size_t)inst * 8 + j] = twx; S.target[(size_t)inst * 8 + 4 + j] = twy;
        }
    }
    if (tid == (WARP ? 4 : 32)) {
        // positions: dt * cumsum of the rotated reference velocity, from the measured position (sequential like numpy)
        double cx = 0.0, cy = 0.0;
        for (int i = 0; i < N; ++i) {
            cx += sc.vx[i]; cy += sc.vy[i];
            xr[0 * (N + 1) + 1 + i] = P.dt * cx + sc.st[0];
            xr[1 * (N + 1) + 1 + i] = P.dt * cy + sc.st[1];
            xr[6 * (N + 1) + 1 + i] = sc.vx[i];
            xr[7 * (N + 1) + 1 + i] = sc.vy[i];
        }
    }
    for (int i = WARP ? tid : tid - 64; i >= 0 && i < N; i += WARP ? GROUP : N) {
        xr[2 * (N + 1) + 1 + i] = sc.zref;
        xr[8 * (N + 1) + 1 + i] = sc.vzref;
        xr[5 * (N + 1) + 1 + i] = w * S.lin_b[i];
        xr[11 * (N + 1) + 1 + i] = w;
        if (sc.flag != 0) {                                                                    // FootstepPlanner.py:153-158
            xr[3 * (N + 1) + 1 + i] = sc.st[3] + sc.vr[3] * S.lin_a[i];
            xr[4 * (N + 1) + 1 + i] = sc.st[4] + sc.vr[4] * S.lin_a[i];
            xr[9 * (N + 1) + 1 + i] = sc.vr[3];
            xr[10 * (N + 1) + 1 + i] = sc.vr[4];
        }
    }
    if (WARP ? tid < 12 : (tid >= 96 && tid < 108)) { const int c = WARP ? tid : tid - 96; xr[c * (N + 1)] = sc.st[c]; }
    group_sync();
    if (tid == 0) {
        // remember the first-step contacts for the next tick's touchdown test
        if (commit) S.prevc[inst] = (uint8_t)(0x80 | (sc.mask[0] & 15));
    }
    if (S.xref_out && commit) {
        for (int i = tid; i < 12 * (N + 1); i += nthr) S.xref_out[(size_t)inst * 12 * (N + 1) + i] = xr[i];
        for (int i = tid; i < 260; i += nthr) S.fsteps_out[(size_t)inst * 260 + i] = fs[i];
    }
}

// Close the loop for instance `inst`: x_next (12) is the MPC's predicted next state in the current frame.
// One thread.
__device__ inline void scenario_advance(const DevScenario& S, int inst, const double* xn) {
    double fx = S.frame[(size_t)inst * 3], fy = S.frame[(size_t)inst * 3 + 1], fyaw = S.frame[(size_t)inst * 3 + 2];
    double sf, cf;
    sincos(fyaw, &sf, &cf);
    fx += cf * xn[0] - sf * xn[1];
    fy += sf * xn[0] + cf * xn[1];
    const double dyaw = xn[5];
    fyaw += dyaw;
    double sy, cy;
    sincos(dyaw, &sy, &cy);
    double st[12];
    st[0] = 0.0; st[1] = 0.0; st[2] = xn[2]; st[3] = xn[3]; st[4] = xn[4]; st[5] = 0.0;
    st[6] = cy * xn[6] + sy * xn[7];
    st[7] = -sy * xn[6] + cy * xn[7];
    st[8] = xn[8];
    st[9] = cy * xn[9] + sy * xn[10];
    st[10] = -sy * xn[9] + cy * xn[10];
    st[11] = xn[11];
    for (int c = 0; c < 12; ++c) {
        const bool zeroed = (c == 0 || c == 1 || c == 5);       // frame is centred on, and yaw-aligned with, the robot
        const double sg = S.sigma[c / 3];
        const double n = (zeroed || sg == 0.0) ? 0.0 : hash_normal(S.seed, (unsigned long long)inst, (unsigned long long)S.tick, (unsigned long long)c) * sg;
        S.state[(size_t)inst * 12 + c] = st[c] + n;
    }
    S.frame[(size_t)inst * 3] = fx; S.frame[(size_t)inst * 3 + 1] = fy; S.frame[(size_t)inst * 3 + 2] = fyaw;
}

// The same, spread over the lanes of the group that owns the robot (lanes 0..11 one state component each: the twelve Gaussian
// samples are most of the cost); `lane` = index inside the group, every lane of the group calls it.  Bit-identical results.
__device__ inline void scenario_advance_lanes(const DevScenario& S, int inst, const double* xn, int lane) {
    const double dyaw = xn[5];
    double sy, cy;
    sincos(dyaw, &sy, &cy);
    if (lane < 12) {
        const int c = lane;
        double v;
        if (c == 0 || c == 1 || c == 5) v = 0.0;
        else if (c == 6) v = cy * xn[6] + sy * xn[7];
        else if (c == 7) v = -sy * xn[6] + cy * xn[7];
        else if (c == 9) v = cy * xn[9] + sy * xn[10];
        else if (c == 10) v = -sy * xn[9] + cy * xn[10];
        else v = xn[c];
        const bool zeroed = (c == 0 || c == 1 || c == 5);
        const double sg = S.sigma[c / 3];
        const double n = (zeroed || sg == 0.0) ? 0.0 : hash_normal(S.seed, (unsigned long long)inst, (unsigned long long)S.tick, (unsigned long long)c) * sg;
        S.state[(size_t)inst * 12 + c] = v + n;
    }
    if (lane == 12) {
        double fx = S.frame[(size_t)inst * 3], fy = S.frame[(size_t)inst * 3 + 1], fyaw = S.frame[(size_t)inst * 3 + 2];
        double sf, cf;
        sincos(fyaw, &sf, &cf);
        fx += cf * xn[0] - sf * xn[1];
        fy += sf * xn[0] + cf * xn[1];
        fyaw += dyaw;
        S.frame[(size_t)inst * 3] = fx; S.frame[(size_t)inst * 3 + 1] = fy; S.frame[(size_t)inst * 3 + 2] = fyaw;
    }
}

}  // namespace mpcqp
