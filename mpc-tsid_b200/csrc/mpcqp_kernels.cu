// mpcqp_kernels.cu -- the hot path: per-tick QP build + solve for a batch of robots, one CTA each.
//
// Replaces, for a whole batch at once, what the reference does per tick on the CPU:
//   MPC.construct_gait / construct_S      MPC.py:635-652, 611-633   -> decode_foot()
//   MPC.update_ML  (lever-arm blocks)     MPC.py:316-360            -> decode_foot()
//   MPC.update_NK  (right-hand side)      MPC.py:362-378            -> free_response()  (condensed form)
//   MPC.call_solver -> osqp solve         MPC.py:380-430            -> sweep() / admm stage
//   MPC.retrieve_result                   MPC.py:432-458            -> finish()
// The QP is solved in condensed form (states and swing-foot forces eliminated);
// the optimum is the same point the reference's sparse QP has (strictly convex, unique).
#include "mpcqp_device.cuh"
#include "mpcqp_scenario.cuh"
#include "mpcqp_foot.cuh"

// Optional phase timing (-DMPCQP_PROFILE): per-phase clock64() deltas of thread 0, summed over CTAs
// into g_prof; read back through mpcqp_debug_profile().  Off in the shipped build.
#ifdef MPCQP_PROFILE
__device__ unsigned long long g_prof[64];     // 0..15 dense path, 16..63 stage-wise path
#define PROF_T0() long long prof_t_ = clock64()
#define PROF(slot) do { if (threadIdx.x == 0) { long long n_ = clock64(); atomicAdd(&g_prof[slot], (unsigned long long)(n_ - prof_t_)); prof_t_ = n_; } } while (0)
#define PROF_COUNT(slot) do { if (threadIdx.x == 0) atomicAdd(&g_prof[slot], 1ull); } while (0)
#else
#define PROF_T0() do {} while (0)
#define PROF(slot) do {} while (0)
#define PROF_COUNT(slot) do {} while (0)
#endif

namespace mpcqp {

// Launch shape per horizon.  N = 16: W is 39 KB, four CTAs of four warps share an SM.  N = 32: W is
// 154 KB, one CTA of eight warps per SM, and the ADMM stage cannot hold a second factor, so its
// polish attempts borrow the ADMM buffer and the ADMM system is refactored if the polish fails.
template <int N>
struct Cfg {
    static constexpr int NW = (N <= 16) ? 4 : 8;                // warps per CTA (one CTA = one QP instance)
    static constexpr int MIN_CTAS = (N <= 16) ? 4 : 1;          // CTAs per SM the active-set kernel is built for
    static constexpr int MIN_CTAS_ADMM = (N <= 16) ? 2 : 1;
    static constexpr bool TWO_BUF = (N <= 16);                  // ADMM kernel keeps W (polish) and W2 (ADMM) at once
};

// -------------------------------------------------------------------------------------------------
// shared-memory plan of one CTA
// -------------------------------------------------------------------------------------------------
template <int N, bool ADMM>
struct Smem {
    static constexpr int NDIM = 6 * N;
    static constexpr int NT = NDIM / 8;
    static constexpr int NTILES = NT * (NT + 1) / 2;
    static constexpr int NF = 4 * N;        // foot-steps
    double W[NTILES * 64];                  // sweep / polish system: W, then inv(L) in place
    double W2[(ADMM && Cfg<N>::TWO_BUF) ? NTILES * 64 : 2];   // ADMM system (ADMM stage, when two factors fit)
    double xr[12 * (N + 1)];                // xref of this instance
    union {
        double fs[20 * 13];                 // fsteps of this instance (dead after decode)
        double fa[12 * NF];                 // per foot-step, struct of arrays: A[9] = dt inv(R gI)[r]x, g[3]
    };
    double gam[NDIM];                       // gradient of the tracking cost w.r.t. the impulses at f = 0
    double u[NDIM];                         // impulse-space work vector (rhs / solution of W v = s)
    double ms[NDIM];                        // M u
    double tmp[NDIM];                       // scratch of tri_solve
    double C2[N * N];                       // C2[k,l] = sum_{i >= max(k,l)} (i-k)(i-l), constant
    double red[40];
    unsigned long long hist[16];            // hashes of signatures already tried (cycle detection)
    unsigned long long mbar;                // mbarrier of the bulk (TMA) staging copies
    int flag;
    int pad;
    ScenarioSmem sc;                        // device-resident closed loop (planner scratch, predicted next state)
};

// -------------------------------------------------------------------------------------------------
// free response of the double integrators and its gradient                  [MPC.py:362-378, condensed]
//   gam[6k + c] = sum_{s = k+1..N} ( (s-1-k) dt Qp_c ep_s,c + Qv_c ev_s,c )
//   ep_s = p0 + s dt v0 + dt g_c s(s-1)/2 - xref_p[s],   ev_s = v0 + s g_c - xref_v[s]
// -------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void free_response(const DevParams& P, const double* xr, double* gam) {
    for (int idx = threadIdx.x; idx < 6 * N; idx += blockDim.x) {
        const int k = idx / 6, c = idx - 6 * k;
        const double p0 = xr[c * (N + 1)], v0 = xr[(6 + c) * (N + 1)];
        const double gc = (c == 2) ? -P.gravity * P.dt : 0.0;                 // MPC.py:200-201
        double acc0 = 0.0, acc1 = 0.0;
        for (int s = k + 1; s <= N; ++s) {
            const double ep = p0 + s * P.dt * v0 + P.dt * gc * (0.5 * s * (s - 1)) - xr[c * (N + 1) + s];
            const double ev = v0 + s * gc - xr[(6 + c) * (N + 1) + s];
            acc0 = fma((double)(s - 1 - k) * P.dt * P.wp[c], ep, acc0);
            acc1 = fma(P.wv[c], ev, acc1);
        }
        gam[idx] = acc0 + acc1;
    }
}

// ms = M u   (six independent N x N Gram matrices, one per impulse component):
//   M_c[k,l] = dt^2 Qp_c C2[k,l] + Qv_c C0[k,l],  C0[k,l] = N - max(k,l),  C2 from shared memory
template <int N>
__device__ __forceinline__ void gram_apply(const DevParams& P, const double* C2, const double* u, double* ms) {
    for (int idx = threadIdx.x; idx < 6 * N; idx += blockDim.x) {
        const int k = idx / 6, c = idx - 6 * k;
        const double* row = C2 + k * N;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll 2
        for (int l = 0; l < N; l += 4) {
            const double u0 = u[6 * l + c], u1 = u[6 * (l + 1) + c], u2 = u[6 * (l + 2) + c], u3 = u[6 * (l + 3) + c];
            a0 = fma(row[l], u0, a0);
            a1 = fma(row[l + 1], u1, a1);
            a2 = fma(row[l + 2], u2, a2);
            a3 = fma(row[l + 3], u3, a3);
            b0 = fma((double)(N - (k > l ? k : l)), u0, b0);
            b1 = fma((double)(N - (k > l + 1 ? k : l + 1)), u1, b1);
            b2 = fma((double)(N - (k > l + 2 ? k : l + 2)), u2, b2);
            b3 = fma((double)(N - (k > l + 3 ? k : l + 3)), u3, b3);
        }
        ms[idx] = P.dt * P.dt * P.wp[c] * ((a0 + a1) + (a2 + a3)) + P.wv[c] * ((b0 + b1) + (b2 + b3));
    }
}

// W = M^-1 + blockdiag_k( sum_j Bv Z D^-1 Z' Bv' )   written in tile layout.
// The constant part (40 KB, L2 resident) is staged by one bulk async copy (TMA path); the per-foot
// arithmetic overlaps it and the 6x6 blocks are added once the copy has landed.
template <int N, int NTILES>
__device__ __forceinline__ void assemble_W(const DevParams& P, double* W, const double* fa, unsigned long long* mbar,
                                           unsigned int& phase, const Face& fc, double ddx, double ddz, bool admm,
                                           int k, int j, bool foot_thread) {
    __syncthreads();                                    // nobody still reads the previous contents of W
    if (threadIdx.x == 0) {
        fence_async_smem();
        mbar_expect_tx(mbar, NTILES * 64 * 8);
        bulk_g2s(W, P.Minv_tiled, NTILES * 64 * 8, mbar);
    }
    // b[col][0..5]: the three columns of Bv Z
    double b[3][6], d[3];
    if (foot_thread) {
        double A[9];
        load_A<4 * N>(fa, threadIdx.x, A);
        const double lin = P.dt / P.mass;
        const double zx = fc.zx ? 1.0 : 0.0, zy = fc.zy ? 1.0 : 0.0, zz = fc.zz ? 1.0 : 0.0;
        b[0][0] = lin * zx; b[0][1] = 0.0; b[0][2] = 0.0;
        b[1][0] = 0.0; b[1][1] = lin * zy; b[1][2] = 0.0;
        b[2][0] = lin * fc.czx * zz; b[2][1] = lin * fc.czy * zz; b[2][2] = lin * zz;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            b[0][3 + r] = A[3 * r] * zx;
            b[1][3 + r] = A[3 * r + 1] * zy;
            b[2][3 + r] = (A[3 * r] * fc.czx + A[3 * r + 1] * fc.czy + A[3 * r + 2]) * zz;
        }
        d[0] = admm ? ddx * zx : fc.dx; d[1] = admm ? ddx * zy : fc.dy; d[2] = admm ? ddz * zz : fc.dz;
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
    if (foot_thread) {
        int e = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int c = 0; c <= a; ++c, ++e) {
                double v = d[0] * b[0][a] * b[0][c] + d[1] * b[1][a] * b[1][c] + d[2] * b[2][a] * b[2][c];
                v += shfl_xor_d(v, 1);
                v += shfl_xor_d(v, 2);
                if ((e & 3) == j) {                              // the four feet share the 21 updates
                    const int gi = 6 * k + a, gj = 6 * k + c;
                    const int I = gi >> 3, J = gj >> 3;
                    double* T = W + tile_index(I, J) * 64;
                    T[elem_off(gi & 7, gj & 7)] += v;
                    if (I == J && a != c) T[elem_off(gj & 7, gi & 7)] += v;
                }
            }
    }
    __syncthreads();
}


// One equality-constrained solve on the faces given by `sig`, then the KKT guard and the next
// active-set guess.  Returns (CTA-uniform) 1 if the guard passed for every foot, 0 if not, -1 if W
// was not positive definite.
template <int N, bool ADMM>
__device__ int sweep(const DevParams& P, Smem<N, ADMM>& sm, unsigned int& phase, bool contact, uint8_t sig,
                     uint8_t& nsig, FootSol& sol, int k, int j, bool foot_thread) {
    using S = Smem<N, ADMM>;
    constexpr int NF = 4 * N;
    const int tid = threadIdx.x;
    const double lin = P.dt / P.mass;
    Face fc;
    PROF_T0();
    PROF_COUNT(15);
    make_face(P, foot_thread && contact, sig, fc);
    assemble_W<N, S::NTILES>(P, sm.W, sm.fa, &sm.mbar, phase, fc, 0.0, 0.0, false, k, j, foot_thread);
    PROF(0);
    const bool spd = factor_invert_tiles<S::NT, Cfg<N>::NW>(sm.W, &sm.flag);
    PROF(1);
    if (!spd) return -1;

    // f = pf + Z q.  Pass 0 solves from pf; passes 1..refine are iterative-refinement steps on the
    // reduced system; the last pass only evaluates the gradient at the final point for the guard.
    const bool have_pf = foot_thread && (fc.pf[2] != 0.0);
    const int any_pf = __syncthreads_or(have_pf ? 1 : 0);
    double f[3] = {fc.pf[0], fc.pf[1], fc.pf[2]};
    double grad[3] = {0.0, 0.0, 0.0};
    for (int pass = 0;; ++pass) {
        // grad = H f + g  (H f skipped on pass 0 when f = pf = 0 everywhere)
        if (pass > 0 || any_pf) {
            if (foot_thread) {
                double A[9], v[6];
                load_A<NF>(sm.fa, tid, A);
                bv_apply(A, lin, f, v);
                step_sum_store(v, sm.u + 6 * k, j);
            }
            __syncthreads();
            gram_apply<N>(P, sm.C2, sm.u, sm.ms);
            __syncthreads();
            if (foot_thread) {
                double A[9];
                load_A<NF>(sm.fa, tid, A);
                bvT_apply(A, lin, sm.ms + 6 * k, grad);
#pragma unroll
                for (int c = 0; c < 3; ++c) grad[c] += P.w_force * f[c] + sm.fa[(9 + c) * NF + tid];
            }
            PROF(2);
        } else if (foot_thread) {
#pragma unroll
            for (int c = 0; c < 3; ++c) grad[c] = sm.fa[(9 + c) * NF + tid];
        }
        if (pass == 1 + P.refine) break;
        // reduced rhs r = -Z' grad, t = D^-1 r, s_k = sum_j Bv (Z t)
        double tx = 0.0, ty = 0.0, tz_ = 0.0;
        if (foot_thread) {
            tx = fc.zx ? -grad[0] * fc.dx : 0.0;
            ty = fc.zy ? -grad[1] * fc.dy : 0.0;
            tz_ = fc.zz ? -(fc.czx * grad[0] + fc.czy * grad[1] + grad[2]) * fc.dz : 0.0;
            const double zt[3] = {tx + fc.czx * tz_, ty + fc.czy * tz_, tz_};
            double A[9], v[6];
            load_A<NF>(sm.fa, tid, A);
            bv_apply(A, lin, zt, v);
            step_sum_store(v, sm.u + 6 * k, j);
        }
        __syncthreads();
        PROF(3);
        tri_solve<S::NT, Cfg<N>::NW>(sm.W, sm.u, sm.tmp);
        PROF(4);
        if (foot_thread) {
            double A[9], h[3];
            load_A<NF>(sm.fa, tid, A);
            bvT_apply(A, lin, sm.u + 6 * k, h);                 // Bv' v
            const double qx = tx - fc.dx * (fc.zx ? h[0] : 0.0);
            const double qy = ty - fc.dy * (fc.zy ? h[1] : 0.0);
            const double qz = tz_ - fc.dz * (fc.zz ? (fc.czx * h[0] + fc.czy * h[1] + h[2]) : 0.0);
            f[0] += qx + fc.czx * qz;
            f[1] += qy + fc.czy * qz;
            f[2] += qz;
        }
        __syncthreads();
    }

    // ---- multipliers, KKT guard, next signature (per foot)
    bool ok = true;
    nsig = sig;
#pragma unroll
    for (int c = 0; c < 3; ++c) sol.f[c] = 0.0;
#pragma unroll
    for (int r = 0; r < 5; ++r) sol.y[r] = 0.0;
    if (foot_thread && contact) ok = kkt_guard(P, sig, f, grad, sol, nsig);
    const int all_ok = __syncthreads_and(ok ? 1 : 0);
    PROF(5);
    return all_ok ? 1 : 0;
}

// order-sensitive hash of the CTA's signature (cycle detection for the active-set sweeps)
__device__ __forceinline__ unsigned long long sig_hash(uint8_t sig, bool hashed, unsigned long long* slot) {
    if (threadIdx.x == 0) *slot = 0ull;
    __syncthreads();
    if (hashed) {
        unsigned long long h = (unsigned long long)(sig + 1) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29; h *= (2ull * threadIdx.x + 0xBF58476D1CE4E5B9ull); h ^= h >> 32;
        atomicAdd(slot, h);
    }
    __syncthreads();
    return *slot;
}

// -------------------------------------------------------------------------------------------------
// finish: states by forward simulation, objective, masks, outputs            [MPC.py:432-458]
// -------------------------------------------------------------------------------------------------
template <int N, bool ADMM>
__device__ __forceinline__ void finish(const DevParams& P, const DevScenario& S, Smem<N, ADMM>& sm, const DevState& st, int inst,
                       bool contact, const FootSol& sol, uint8_t sig, int k, int j, bool foot_thread, bool valid_A,
                       int status, int sweeps, int iters) {
    constexpr int NF = 4 * N;
    const double lin = P.dt / P.mass;
    // impulses of the final forces
    if (foot_thread) {
        double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (contact && valid_A) {
            double A[9];
            load_A<NF>(sm.fa, threadIdx.x, A);
            bv_apply(A, lin, sol.f, v);
        }
        step_sum_store(v, sm.u + 6 * k, j);
    }
    __syncthreads();
    double part = 0.0;
    if (threadIdx.x < 6) {
        // component c: p_{s+1} = p_s + dt v_s, v_{s+1} = v_s + u_s + g_c          (MPC.py:110-111, 200-205)
        const int c = threadIdx.x;
        double p = sm.xr[c * (N + 1)], v = sm.xr[(6 + c) * (N + 1)];
        const double gc = (c == 2) ? -P.gravity * P.dt : 0.0;
        double* xs = st.xs + (size_t)inst * 12 * N;
        for (int s = 0; s < N; ++s) {
            const double pn = p + P.dt * v;
            const double vn = v + sm.u[6 * s + c] + gc;
            p = pn; v = vn;
            if (s == 0) { sm.sc.xnext[c] = p; sm.sc.xnext[6 + c] = v; }      // MPC.q_next / v_next (MPC.py:448-450)
            const double ep = p - sm.xr[c * (N + 1) + s + 1], ev = v - sm.xr[(6 + c) * (N + 1) + s + 1];
            xs[12 * s + c] = ep;
            xs[12 * s + 6 + c] = ev;
            if (s == 0) { st.x1[(size_t)inst * 12 + c] = p; st.x1[(size_t)inst * 12 + 6 + c] = v; }
            part += 0.5 * (P.wp[c] * ep * ep + P.wv[c] * ev * ev);
        }
    }
    if (foot_thread) {
        part += 0.5 * P.w_force * (sol.f[0] * sol.f[0] + sol.f[1] * sol.f[1] + sol.f[2] * sol.f[2]);
        double* fo = st.f + (size_t)inst * 12 * N + 12 * k + 3 * j;
        fo[0] = sol.f[0]; fo[1] = sol.f[1]; fo[2] = sol.f[2];
        double* yo = st.y + (size_t)inst * 20 * N + 20 * k + 5 * j;
#pragma unroll
        for (int r = 0; r < 5; ++r) yo[r] = sol.y[r];
        st.sig[(size_t)inst * 4 * N + 4 * k + j] = sig;
        if (k == 0) {
            double* f0 = st.f0 + (size_t)inst * 12 + 3 * j;
            f0[0] = sol.f[0]; f0[1] = sol.f[1]; f0[2] = sol.f[2];
        }
    }
    // objective: block reduction
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) sm.red[threadIdx.x >> 5] = part;
    // masks: contact (4N bits) and rows that hold with equality (20N bits)
    constexpr int AW = (20 * N + 31) / 32, CW = (4 * N + 31) / 32;
    unsigned int* amask = reinterpret_cast<unsigned int*>(sm.ms);
    for (int i = threadIdx.x; i < AW + CW; i += blockDim.x) amask[i] = 0u;
    __syncthreads();
    if (foot_thread) {
        // a swing foot is pinned to f = 0 (MPC.py:355-358), so all five of its rows sit on their bound
        const double mu = P.mu, tol = 1e-9;
        const double fx = sol.f[0], fy = sol.f[1], fz = sol.f[2];
        const double row[5] = {fx - mu * fz, -fx - mu * fz, fy - mu * fz, -fy - mu * fz, -fz};
        const int b0 = 20 * k + 5 * j;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const bool act = (fabs(row[r]) <= tol) || (r == 4 && fabs(row[4] + P.fz_max) <= tol);
            if (act) atomicOr(&amask[(b0 + r) >> 5], 1u << ((b0 + r) & 31));
        }
        if (contact) atomicOr(&amask[AW + ((4 * k + j) >> 5)], 1u << ((4 * k + j) & 31));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < AW; i += blockDim.x) st.active[(size_t)inst * AW + i] = amask[i];
    for (int i = threadIdx.x; i < CW; i += blockDim.x) st.contact[(size_t)inst * CW + i] = amask[AW + i];
    if (threadIdx.x == 0) {
        double o = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) o += sm.red[w];
        st.obj[inst] = o;
        st.status[inst] = status;
        st.sweeps[inst] = sweeps;
        st.iters[inst] = iters;
        if (status != 3) world_pose_step(st.qw + (size_t)inst * 6, sm.sc.xnext);
        // device-resident closed loop: the robot moves to the state the MPC predicted for the next tick
        if (S.enabled && status != 3) scenario_advance(S, inst, sm.sc.xnext);
    }
}

// -------------------------------------------------------------------------------------------------
// The solve kernel.  ADMM = false: active-set stage for every instance (grid = batch).
//                    ADMM = true : fallback stage for the instances queued in st.fb_list.
// -------------------------------------------------------------------------------------------------
template <int N, bool ADMM>
__global__ void __launch_bounds__(32 * Cfg<N>::NW, ADMM ? Cfg<N>::MIN_CTAS_ADMM : Cfg<N>::MIN_CTAS)
solve_kernel(DevParams P, DevState st, DevScenario SC, const double* __restrict__ xref_g, const double* __restrict__ fsteps_g,
             int first_tick, int inst_offset, int inst_count) {
    using S = Smem<N, ADMM>;
    constexpr int NF = 4 * N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S& sm = *reinterpret_cast<S*>(smem_raw);
    const int tid = threadIdx.x;
    const bool foot_thread = tid < NF;
    const int k = tid >> 2, j = tid & 3;
    const double lin = P.dt / P.mass;

    unsigned int phase = 0;
    if (tid == 0) mbar_init(&sm.mbar, 1);
    for (int i = tid; i < N * N; i += blockDim.x) sm.C2[i] = __ldg(P.C2 + i);
    __syncthreads();

    if (ADMM) asm volatile("griddepcontrol.wait;" ::: "memory");     // programmatically dependent on the active-set launch (no-op otherwise)
    int n_work = ADMM ? *st.fb_count : inst_count;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int inst = ADMM ? st.fb_list[w] : w + inst_offset;
        __syncthreads();
#ifdef MPCQP_PROFILE
        const long long inst_t0 = clock64();
#endif
        PROF_T0();
        if (SC.enabled) {
            // ---- device-resident closed loop: the planner runs here, inputs never touch HBM
            scenario_inputs<0>(P, SC, sm.sc, inst, sm.xr, sm.fs, N);
        } else {
            // ---- stage inputs in shared memory: two bulk async copies (16-byte aligned blocks per instance)
            if (tid == 0) {
                fence_async_smem();
                mbar_expect_tx(&sm.mbar, (12 * (N + 1) + 260) * 8);
                bulk_g2s(sm.xr, xref_g + (size_t)inst * 12 * (N + 1), 12 * (N + 1) * 8, &sm.mbar);
                bulk_g2s(sm.fs, fsteps_g + (size_t)inst * 260, 260 * 8, &sm.mbar);
            }
            mbar_wait(&sm.mbar, phase);
            phase ^= 1u;
        }
        bool bad = false, contact = false;
        uint8_t sig = SIG_FREE;
        double A0[9];
        if (foot_thread) decode_foot(P, sm.xr, sm.fs, N, k, j, first_tick != 0, A0, contact, bad);
        for (int i = tid; i < 12 * (N + 1); i += blockDim.x) bad = bad || !isfinite(sm.xr[i]);
        free_response<N>(P, sm.xr, sm.gam);
        const int any_bad = __syncthreads_or(bad ? 1 : 0);       // also: every thread is done reading fs
        const bool warm = P.warm_start && !first_tick;
        if (foot_thread) {
            // per-foot constants move to shared memory (struct of arrays) so that nothing per-foot is
            // live in registers across the factorisation
#pragma unroll
            for (int i = 0; i < 9; ++i) sm.fa[i * NF + tid] = A0[i];
            double g[3];
            bvT_apply(A0, lin, sm.gam + 6 * k, g);
#pragma unroll
            for (int c = 0; c < 3; ++c) sm.fa[(9 + c) * NF + tid] = g[c];
            // warm start: the previous tick's active set, advanced by one step (MPC.py:403-406)
            if (warm && contact) {
                const int ks = (k + 1 < N) ? k + 1 : 0;
                sig = st.sig[(size_t)inst * 4 * N + 4 * ks + j];
                if (sig > 26) sig = SIG_FREE;
            }
        }
        FootSol sol;
#pragma unroll
        for (int c = 0; c < 3; ++c) sol.f[c] = 0.0;
#pragma unroll
        for (int r = 0; r < 5; ++r) sol.y[r] = 0.0;
        PROF(7);

        int sweeps = 0, iters = 0, status = 0;
        bool done = false;
        if (any_bad) {
            // malformed input: report it, emit zero forces (never NaN), skip both solver stages
            status = 3;
            contact = false;
            sig = SIG_FREE;
        } else if (!ADMM) {
            // ---------------- active-set stage
            int nhist = 0;
            for (int s = 0; s < P.max_sweeps && !done; ++s) {
                const unsigned long long h = sig_hash(sig, foot_thread && contact, &sm.hist[15]);
                bool seen = false;
                for (int i = 0; i < nhist; ++i) seen = seen || (sm.hist[i] == h);
                if (seen) break;
                __syncthreads();
                if (tid == 0 && nhist < 15) sm.hist[nhist] = h;
                nhist = (nhist < 15) ? nhist + 1 : nhist;
                uint8_t nsig;
                const int r = sweep<N, ADMM>(P, sm, phase, contact, sig, nsig, sol, k, j, foot_thread);
                ++sweeps;
                if (r < 0) break;
                if (r > 0) { done = true; status = 1; }
                else sig = nsig;
            }
            if (!done) {
                if (P.mode & 2) {
                    if (tid == 0) {
                        const int slot = atomicAdd(st.fb_count, 1);
                        st.fb_list[slot] = inst;
                        st.sweeps[inst] = sweeps;
                    }
                    continue;       // state of this instance is left untouched for the ADMM stage
                }
                status = 0;
#pragma unroll
                for (int c = 0; c < 3; ++c) sol.f[c] = 0.0;
#pragma unroll
                for (int r = 0; r < 5; ++r) sol.y[r] = 0.0;
            }
        } else {
            // ---------------- ADMM stage (fixed rho, Woodbury-form linear solve, guarded polish)
            sweeps = (P.mode & 1) ? st.sweeps[inst] : 0;
            const double rho = P.rho, sigma = P.sigma, alpha = P.alpha, mu = P.mu;
            double z[5] = {0, 0, 0, 0, 0}, f[3] = {0, 0, 0}, y[5] = {0, 0, 0, 0, 0};
            if (foot_thread && contact && warm) {
                const int ks = (k + 1 < N) ? k + 1 : 0;
                const double* fp = st.f + (size_t)inst * 12 * N + 12 * ks + 3 * j;
                const double* yp = st.y + (size_t)inst * 20 * N + 20 * ks + 5 * j;
                f[0] = fp[0]; f[1] = fp[1]; f[2] = fp[2];
#pragma unroll
                for (int r = 0; r < 5; ++r) y[r] = yp[r];
                const double cf[5] = {f[0] - mu * f[2], -f[0] - mu * f[2], f[1] - mu * f[2], -f[1] - mu * f[2], -f[2]};
#pragma unroll
                for (int r = 0; r < 5; ++r) z[r] = fmin(cf[r], 0.0);
                z[4] = fmax(z[4], -P.fz_max);
            }
            // W(rho) = M^-1 + Bv D^-1 Bv' with every force free and D = (w + sigma) I + rho C'C (diagonal)
            const bool live = foot_thread && contact;
            const double ddx = live ? 1.0 / (P.w_force + sigma + 2.0 * rho) : 0.0;
            const double ddz = live ? 1.0 / (P.w_force + sigma + rho * (4.0 * mu * mu + 1.0)) : 0.0;
            Face fa;
            make_face(P, live, SIG_FREE, fa);
            double* const Wadmm = Cfg<N>::TWO_BUF ? sm.W2 : sm.W;
            assemble_W<N, S::NTILES>(P, Wadmm, sm.fa, &sm.mbar, phase, fa, ddx, ddz, true, k, j, foot_thread);
            bool spd = factor_invert_tiles<S::NT, Cfg<N>::NW>(Wadmm, &sm.flag);
            uint8_t prev_sig = 255;
            int stable = 0;
            while (spd && !done && iters < P.max_iter) {
                ++iters;
                double tx = 0, ty = 0, tzz = 0;
                if (foot_thread) {
                    const double v0 = rho * z[0] - y[0], v1 = rho * z[1] - y[1], v2 = rho * z[2] - y[2],
                                 v3 = rho * z[3] - y[3], v4 = rho * z[4] - y[4];
                    tx = ddx * (sigma * f[0] - sm.fa[9 * NF + tid] + (v0 - v1));
                    ty = ddx * (sigma * f[1] - sm.fa[10 * NF + tid] + (v2 - v3));
                    tzz = ddz * (sigma * f[2] - sm.fa[11 * NF + tid] - mu * (v0 + v1 + v2 + v3) - v4);
                    const double tt[3] = {tx, ty, tzz};
                    double A[9], v[6];
                    load_A<NF>(sm.fa, tid, A);
                    bv_apply(A, lin, tt, v);
                    step_sum_store(v, sm.u + 6 * k, j);
                }
                __syncthreads();
                tri_solve<S::NT, Cfg<N>::NW>(Wadmm, sm.u, sm.tmp);
                uint8_t cur = SIG_FREE;
                if (live) {
                    double A[9], h[3];
                    load_A<NF>(sm.fa, tid, A);
                    bvT_apply(A, lin, sm.u + 6 * k, h);
                    const double ftx = tx - ddx * h[0], fty = ty - ddx * h[1], ftz = tzz - ddz * h[2];
                    const double zt[5] = {ftx - mu * ftz, -ftx - mu * ftz, fty - mu * ftz, -fty - mu * ftz, -ftz};
                    f[0] = alpha * ftx + (1.0 - alpha) * f[0];
                    f[1] = alpha * fty + (1.0 - alpha) * f[1];
                    f[2] = alpha * ftz + (1.0 - alpha) * f[2];
                    bool upp[5], low4 = false;
#pragma unroll
                    for (int r = 0; r < 5; ++r) {
                        const double zr = alpha * zt[r] + (1.0 - alpha) * z[r];
                        double zn = fmin(zr + y[r] / rho, 0.0);
                        if (r == 4) zn = fmax(zn, -P.fz_max);
                        y[r] += rho * (zr - zn);
                        z[r] = zn;
                        upp[r] = (0.0 - zn) < y[r];                    // OSQP's polish rule
                        if (r == 4) low4 = (zn + P.fz_max) < -y[r];
                    }
                    const int sx = (upp[0] ? 1 : 0) - (upp[1] ? 1 : 0), sy = (upp[2] ? 1 : 0) - (upp[3] ? 1 : 0);
                    const bool apex = upp[4] || (upp[0] && upp[1]) || (upp[2] && upp[3]);
                    cur = sig_pack(sx, sy, apex ? 1 : (low4 ? 2 : 0));
                }
                __syncthreads();
                if (iters >= P.min_iter && (iters % P.check_every) == 0) {
                    const int same = __syncthreads_and((cur == prev_sig) ? 1 : 0);
                    prev_sig = cur;
                    stable = same ? stable + 1 : 0;
                    if (stable >= 1) {
                        stable = 0;
                        uint8_t nsig;
                        FootSol trial;
                        const int r = sweep<N, ADMM>(P, sm, phase, contact, cur, nsig, trial, k, j, foot_thread);
                        ++sweeps;
                        if (r > 0) { sol = trial; sig = cur; done = true; status = 1; }
                        else if (!Cfg<N>::TWO_BUF) {
                            // the polish borrowed the ADMM buffer: rebuild the ADMM factor and carry on
                            assemble_W<N, S::NTILES>(P, Wadmm, sm.fa, &sm.mbar, phase, fa, ddx, ddz, true, k, j, foot_thread);
                            spd = factor_invert_tiles<S::NT, Cfg<N>::NW>(Wadmm, &sm.flag);
                        }
                    }
                }
            }
            if (!done) {
                status = 2;
                sig = SIG_FREE;
#pragma unroll
                for (int c = 0; c < 3; ++c) sol.f[c] = live ? f[c] : 0.0;
#pragma unroll
                for (int r = 0; r < 5; ++r) sol.y[r] = live ? y[r] : 0.0;
            }
        }
        PROF(8);
        finish<N, ADMM>(P, SC, sm, st, inst, contact, sol, sig, k, j, foot_thread, !any_bad, status, sweeps, iters);
        PROF(9);
#ifdef MPCQP_PROFILE
        if (threadIdx.x == 0) { atomicAdd(&g_prof[ADMM ? 13 : 12], (unsigned long long)(clock64() - inst_t0)); atomicAdd(&g_prof[ADMM ? 11 : 10], 1ull); }
#endif
    }
}

// -------------------------------------------------------------------------------------------------
// Build-half parity hook: the coefficients MPC.update_ML / update_NK write each tick, in the
// reference's own order (MPC.py:154-166, 349, 355-358, 362-378).
// -------------------------------------------------------------------------------------------------
__global__ void export_build_kernel(DevParams P, const double* __restrict__ xref_g, const double* __restrict__ fsteps_g,
                                    int first_tick, double* __restrict__ Bv, double* __restrict__ Sv, double* __restrict__ NK) {
    __shared__ double xr[12 * 65];
    __shared__ double fs[260];
    const int N = P.N;                                   // any horizon up to 64
    const int inst = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < 12 * (N + 1); i += blockDim.x) xr[i] = xref_g[(size_t)inst * 12 * (N + 1) + i];
    for (int i = tid; i < 260; i += blockDim.x) fs[i] = fsteps_g[(size_t)inst * 260 + i];
    __syncthreads();
    for (int t = tid; t < 4 * N; t += blockDim.x) {
        const int k = t >> 2, j = t & 3;
        double A[9];
        bool bad = false, contact = false;
        decode_foot(P, xr, fs, N, k, j, first_tick != 0, A, contact, bad);
        double* b = Bv + (size_t)inst * 48 * N + 48 * k + 12 * j;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            b[4 * c + 0] = P.dt / P.mass;                                    // MPC.py:119
            b[4 * c + 1] = A[0 + c];
            b[4 * c + 2] = A[3 + c];
            b[4 * c + 3] = A[6 + c];
        }
        double* s = Sv + (size_t)inst * 12 * N + 12 * k + 3 * j;
        s[0] = s[1] = s[2] = contact ? 0.0 : 1.0;                            // MPC.py:628-630
    }
    for (int idx = tid; idx < 12 * N; idx += blockDim.x) {
        const int k = idx / 12, i = idx - 12 * k;
        // row k of N: -g - [k == 0] A x0 + X*_{k+1} - [k > 0] A X*_k          (MPC.py:366-376)
        double v = xr[i * (N + 1) + k + 1];
        if (i == 8) v += P.gravity * P.dt;
        double ax = xr[i * (N + 1) + k];
        if (i < 6) ax += P.dt * xr[(i + 6) * (N + 1) + k];
        NK[(size_t)inst * 12 * N + idx] = v - ax;
    }
}

// -------------------------------------------------------------------------------------------------
// Logger-facing output: contribution of each state component and of the forces to the
// cost over the horizon, cost_i = sum_k x_i P_i x_i (Logger.log_cost_function, Logger.py:406-418; no 1/2).
// One warp per robot; cost is B x 13.
// -------------------------------------------------------------------------------------------------
__global__ void cost_components_kernel(DevParams P, DevState st, int N, double* __restrict__ cost) {
    const int robot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (robot >= P.batch) return;
    const double* xs = st.xs + (size_t)robot * 12 * N;
    const double* f = st.f + (size_t)robot * 12 * N;
    double acc = 0.0;                                   // lanes 0..11: state component `lane`; lanes 12..31: forces
    if (lane < 12) {
        const double w = lane < 6 ? P.wp[lane] : P.wv[lane - 6];
        for (int k = 0; k < N; ++k) { const double e = xs[12 * k + lane]; acc = fma(w * e, e, acc); }
    } else {
        for (int i = lane - 12; i < 12 * N; i += 20) acc = fma(P.w_force * f[i], f[i], acc);
    }
    double fsum = lane >= 12 ? acc : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, o);
    if (lane < 12) cost[(size_t)robot * 13 + lane] = acc;
    if (lane == 12) cost[(size_t)robot * 13 + 12] = fsum;
}

// -------------------------------------------------------------------------------------------------
// FP64 peak probes (roofline denominators; MEASURED_PEAKS.json has no FP64 entry)
// -------------------------------------------------------------------------------------------------
__global__ void peak_dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void peak_dmma_kernel(double* out, int iters, double a, double b) {
    double c[4][2];
#pragma unroll
    for (int q = 0; q < 4; ++q) { c[q][0] = threadIdx.x + q; c[q][1] = q; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma884(c[q][0], c[q][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) s += c[q][0] + c[q][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace mpcqp

