/*
 * mpcqp.h -- C ABI of libmpcqp.so, the B200-native batched centroidal-MPC QP engine.
 *
 * This is the drop-in boundary for ONE hot path of thomascbrs/mpc-tsid: the per-tick QP build in
 * MPC.py and the OSQP solve it drives.  Each entry point names the reference interface it
 * replaces (file:line under /root/reference).  Plain pointers and sizes only; no torch types.
 *
 * Conventions (reference: MPC.py:460-514)
 *   - all arrays are float64, C order, one instance after another (leading batch dimension B);
 *   - xref   : B x 12 x (N+1)   column 0 = measured state [x y z roll pitch yaw vx vy vz wx wy wz]
 *                               in the local frame, columns 1..N = reference trajectory;
 *   - fsteps : B x 20 x 13      row r = [steps left in phase r | (x,y,z) of FL, FR, HL, HR],
 *                               NaN (or x == 0.0) marks a swing foot, a 0 in column 0 ends the table;
 *   - forces : B x 12           (FL, FR, HL, HR) x (fx, fy, fz) of the first horizon step
 *                               == MPC.f_applied (MPC.py:441);
 *   - x      : B x 24N          [X_1 - xref_1 .. X_N - xref_N ; f_0 .. f_{N-1}] == MPC.x (MPC.py:428);
 *   - every function returns 0 on success (the reference's methods all `return 0`) or a negative
 *     MPCQP_ERR_* code; mpcqp_last_error() returns a thread-local message.  Inputs are never
 *     written (the reference overwrites NaNs in the caller's fsteps, MPC.py:327; we do not).
 *   - `location` says where a caller buffer lives: MPCQP_HOST or MPCQP_DEVICE (same device as the handle).
 *   - one handle owns one CUDA stream; calls on one handle must not race.  No global state.
 */
#ifndef MPCQP_H_
#define MPCQP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCQP_HOST 0
#define MPCQP_DEVICE 1

#define MPCQP_OK 0
#define MPCQP_ERR_INVALID (-1)      /* bad argument / unsupported size */
#define MPCQP_ERR_CUDA (-2)         /* CUDA runtime error, see mpcqp_last_error() */
#define MPCQP_ERR_NO_DEVICE (-3)    /* no CUDA device: there is NO CPU fallback */
#define MPCQP_ERR_STATE (-4)        /* call order (e.g. result requested before any run) */

/* per-instance status written by mpcqp_run */
#define MPCQP_STATUS_UNSOLVED 0
#define MPCQP_STATUS_SOLVED 1          /* active-set polished, KKT-verified (exact to rounding) */
#define MPCQP_STATUS_MAX_ITER 2        /* the fallback stage ran out of iterations: a feasible (inside the friction pyramid), uncertified
                                          iterate is returned, x[:12N] is the roll-out of exactly these forces */
#define MPCQP_STATUS_BAD_INPUT 3       /* non-finite state / malformed gait table; forces are 0 */

/* solver stages that may run (bit mask in mpcqp_params.mode) */
#define MPCQP_MODE_ACTIVE_SET 1     /* warm-started primal-dual active-set sweeps (fast path)   */
#define MPCQP_MODE_ADMM 2           /* fixed-rho ADMM + guarded polish (globally convergent)    */
#define MPCQP_MODE_STAGEWISE 4      /* active-set stage factorises stage by stage (Riccati recursion over the horizon, one warp per
                                       robot, O(N)) instead of the dense 6N x 6N condensed system (one CTA per robot, O(N^3))      */
#define MPCQP_MODE_IPM 8            /* fallback stage of the stage-wise path: primal-dual interior-point iterations on the stage-wise
                                       factorisation (any horizon, ~20 factorisations whatever the active set), then active-set
                                       sweeps from the rows it identifies; takes precedence over MPCQP_MODE_ADMM when both are set   */

#define MPCQP_MODE_LANE 16          /* run the active-set stage of the stage-wise path with ONE LANE per robot: no redundant arithmetic, the
                                       O(N) data of a robot in an HBM / L2 workspace, any horizon with one instantiation.  Opt-in: it gives
                                       the same answers (tests/test_gpu_lane.py) but, as measured, half the throughput of the default half
                                       warp per robot.  Needs MPCQP_MODE_STAGEWISE | MPCQP_MODE_IPM.          */

typedef struct mpcqp_handle mpcqp_handle;

/* Everything MPC.__init__ hard-codes (MPC.py:22-82) plus solver settings.  Fill with
 * mpcqp_default_params() first, then override. */
typedef struct mpcqp_params {
    int32_t struct_size;        /* = sizeof(mpcqp_params), ABI check */
    int32_t n_steps;            /* horizon N, 1 .. 64   (MPC.py:42; main.py:23: n_periods * T_gait / dt); the dense stages: 16, 32 */
    int32_t batch;              /* B independent robots               */
    int32_t device;             /* CUDA device ordinal                */
    double dt;                  /* MPC.py:25 */
    double T_gait;              /* MPC.py:45 (kept for API parity; unused by the QP) */
    double mass;                /* MPC.py:28 */
    double mu;                  /* MPC.py:39 */
    double fz_max;              /* MPC.py:228 */
    double gravity;             /* MPC.py:201 (9.81) */
    double gI[9];               /* body inertia, row major   (MPC.py:35-37) */
    double footholds[12];       /* 3 x 4 row major, default footholds used at k == 0 (MPC.py:67-70, 176) */
    double w_state[12];         /* diagonal state weights    (MPC.py:255-275) */
    double w_force;             /* force weight              (MPC.py:282-284) */
    /* solver */
    int32_t mode;               /* MPCQP_MODE_* mask, default ACTIVE_SET | STAGEWISE | IPM */
    int32_t max_sweeps;         /* active-set sweeps before the fallback stage takes over */
    int32_t max_iter;           /* ADMM iteration cap */
    int32_t min_iter;           /* ADMM iterations before the first polish attempt */
    int32_t check_every;        /* ADMM iterations between active-set stability checks */
    int32_t warm_start;         /* 1 = shift the previous tick's solution (MPC.py:403-406) */
    double rho;                 /* ADMM penalty on the pyramid rows */
    double sigma;               /* ADMM proximal weight */
    double alpha;               /* ADMM relaxation */
    double feas_tol;            /* primal feasibility tolerance of the KKT guard [N] */
    double dual_tol;            /* multiplier sign tolerance of the KKT guard */
    int32_t refine;             /* iterative-refinement passes per equality-constrained solve (0 or 1) */
    int32_t ipm_max_iter;       /* interior-point iterations per round (two rounds: gap 1e-8, then 1e-11) */
} mpcqp_params;

/* Reference constants of MPC.py:22-82 for the Solo trot configuration (dt 0.02, N 16, T_gait 0.32). */
void mpcqp_default_params(mpcqp_params* p);

/* replaces MPC.__init__ (MPC.py:22-82) / MPC_Wrapper.__init__ (MPC_Wrapper.py:20-37) */
int mpcqp_create(const mpcqp_params* p, mpcqp_handle** out);
int mpcqp_destroy(mpcqp_handle* h);

/* replaces MPC.run(k, xref, fsteps) (MPC.py:460-514) for B instances: build + solve + extract.
 * k == 0 mirrors the reference's first tick (default footholds as lever arms, no warm start,
 * MPC.py:176,413); any k > 0 is a regular tick.  Asynchronous on the handle's stream. */
int mpcqp_run(mpcqp_handle* h, double k, const double* xref, const double* fsteps, int location);

/* replaces MPC_Wrapper.get_latest_result() (MPC_Wrapper.py:57-78) / MPC.f_applied (MPC.py:441):
 * B x 12 forces of the last run.  Synchronises the handle's stream when location == MPCQP_HOST. */
int mpcqp_get_latest_result(mpcqp_handle* h, double* forces, int location);

/* What a control loop reads every tick, in one call and one synchronisation: forces (B x 12, as above; may be NULL) and
 * next_state (B x 12; may be NULL) = x_robot[:, 0], the first predicted state: MPC.q_next = next_state[0:6], MPC.v_next =
 * next_state[6:12] (MPC.py:448-450), from which the dead-reckoned MPC.q_w (MPC.py:503-510) follows. */
int mpcqp_get_step_result(mpcqp_handle* h, double* forces, double* next_state, int location);

/* MPC.q_w (MPC.py:58, 503-510), B x 6: the world pose every run dead-reckons from MPC.q_next on the device (starts at
 * [0, 0, 0.2027682, 0, 0, 0], MPC.py:55-58).  set == 0 reads it, set != 0 overwrites it (e.g. to re-anchor on an estimator). */
int mpcqp_world_pose(mpcqp_handle* h, double* qw, int set, int location);

/* status[B] alone (MPCQP_STATUS_*): the cheap check the reference never makes (MPC.py:427 ignores sol.info.status). */
int mpcqp_get_status(mpcqp_handle* h, int32_t* status, int location);

/* MPC.x (MPC.py:428): B x 24N.  MPC.x_robot (MPC.py:437-447) is x[:12N] + xref[:,1:]. */
int mpcqp_get_solution(mpcqp_handle* h, double* x, int location);

/* Per-instance diagnostics the reference never exposes (it ignores sol.info.status, MPC.py:427):
 * any pointer may be NULL.  status[B], sweeps[B] (active-set factorizations), iters[B] (iterations of the
 * fallback stage: interior-point or ADMM), obj[B] (1/2 x'Px), contact[B*ceil(4N/32)] (4N bits: bit 4k+j = foot j in stance at step k),
 * active[B*ceil(20N/32)] (bit 20k+5j+r = pyramid row r of foot j at step k holds with equality),
 * y[B*20N] multipliers of the pyramid rows (rows of L in MPC.py:136-148). */
int mpcqp_get_info(mpcqp_handle* h, int32_t* status, int32_t* sweeps, int32_t* iters, double* obj,
                   uint32_t* contact, uint32_t* active, double* y, int location);

/* Number of instances the last run sent to the fallback stage (host int). */
int mpcqp_get_fallback_count(mpcqp_handle* h, int32_t* count);

/* Forget the carried solution (what a fresh osqp.OSQP() + k == 0 does in the reference). */
int mpcqp_reset_warm_start(mpcqp_handle* h);

int mpcqp_synchronize(mpcqp_handle* h);
/* the CUDA stream (cudaStream_t) work is enqueued on; owned by the handle */
void* mpcqp_stream(mpcqp_handle* h);
/* Overlap of consecutive ticks.  The robots of a batch are independent (MPC.py holds one robot; tick t + 1 of a robot needs only
 * its own tick t: the shifted warm start, MPC.py:403-406), but a tick issued as one launch ends when its slowest robot does, and a
 * robot whose warm-start guess misses needs a second factorisation sweep -- at 4096 robots a third of a tick's time is that tail.
 * With `ranges` >= 2 (at most 8) a tick issued by mpcqp_run (MPCQP_DEVICE; MPCQP_HOST: every range stages its own rows on its own
 * stream) or by mpcqp_scenario_run is cut into that many
 * contiguous index ranges, each on its own internal stream with its own fallback queue: a range's tick t + 1 is ordered behind
 * its own tick t only, so the tail of one range is filled by the other ranges' robots.  The inputs of such a tick must be complete
 * on mpcqp_stream(h) at the time of the call (or the caller synchronised).  Every entry point that reads or changes the handle's
 * state (results, status, reset, host-input runs, mpcqp_synchronize, mpcqp_join) first makes mpcqp_stream(h) wait for all ranges,
 * so callers of this API always see whole ticks; only work the caller enqueues on mpcqp_stream(h) HIMSELF needs mpcqp_join first.
 * ranges = 1 switches the overlap off; 0 (default) = automatic: off for mpcqp_run, and inside one mpcqp_scenario_run call of two
 * or more ticks eight ranges when the batch is one to three waves of resident robots, two up to eight waves, joined before the call returns.
 * mpcqp_result_async does not join: every range copies its own forces behind its own tick, so a loop that issues tick t + 1 before it
 * waits for the forces of tick t (the asynchronous protocol below) keeps all ranges busy.
 * Results are bit-identical with and without overlap (tests/test_gpu_canary.py). */
int mpcqp_set_overlap(mpcqp_handle* h, int ranges);
/* make mpcqp_stream(h) wait for every index range in flight (no host synchronisation) */
int mpcqp_join(mpcqp_handle* h);
/* Page-locked host memory for arrays that cross the bus every tick (xref, fsteps, forces): copies to / from pageable
 * memory are staged by the driver and cost about a fifth of a 4096-robot tick.  NULL on failure. */
void* mpcqp_host_alloc(size_t bytes);
int mpcqp_host_free(void* p);
/* kernels launched by this handle since creation (for bench.py's gpu_launches claim) */
int64_t mpcqp_launch_count(mpcqp_handle* h);

/* Parity hook for the build half: the coefficients MPC.update_ML / update_NK write each tick.
 *   B_vals : B x N x 48   == ML.data[i_update_B + 96k]      (MPC.py:349)
 *   S_vals : B x 12N      == ML.data[i_update_S]            (MPC.py:355-358)
 *   NK     : B x 12N      == NK[:12N]                        (MPC.py:362-378)
 * Synchronous. */
int mpcqp_export_build(mpcqp_handle* h, double k, const double* xref, const double* fsteps, int location,
                       double* B_vals, double* S_vals, double* NK);

/* ---- device-resident closed loop (SURVEY.md 8f rows f1 + f2; the reference's producer side) --------------
 * The footstep planner (FootstepPlanner.update_fsteps / getRefStates, FootstepPlanner.py:76-161, 284-445) and the
 * closed-loop state update (MPC.q_next / v_next, MPC.py:448-450, moved into the next local frame as
 * Interface.py:100-132 does) run inside the solve kernel, so a tick needs no host producer and no input copy.
 * Per instance: seq = ceil(period / 16) 64-bit words (period = T_gait / dt <= 64 steps: dt = 0.01 gives 32), bit 4*(s % 16) + j of word
 * s / 16 = foot j in contact at step s of the gait period (FootstepPlanner.py:207-282; one word per instance while period <= 16), phase = offset into that period, vref = 6 commanded velocities, state = 12 initial
 * measured states; sigma4 = Gaussian noise (position, angle, linear, angular velocity) drawn from a counter-based
 * generator keyed by (seed, instance, tick, component).  All pointers are HOST pointers. */
int mpcqp_scenario_init(mpcqp_handle* h, const uint64_t* seq, const int32_t* phase, const double* vref,
                        const double* state, const double* sigma4, uint64_t seed);
/* New joystick commands from the next tick on (Joystick.update_v_ref, Joystick.py:29-43): vref B x 6 (HOST pointer, NULL keeps the
 * current ones) -- vz / roll-rate / pitch-rate commands drive getRefStates' state machine (FootstepPlanner.py:128-152) -- and
 * the `reduced` support-polygon switch (Joystick.py:66-67, FootstepPlanner.py:330-332). */
int mpcqp_scenario_set_commands(mpcqp_handle* h, const double* vref, int reduced);
/* Run `ticks` closed-loop ticks (plan -> build -> solve -> integrate) back to back on the device.  With
 * emit_inputs != 0 the xref / fsteps generated by the last tick can be read with mpcqp_get_inputs. */
int mpcqp_scenario_run(mpcqp_handle* h, int ticks, int emit_inputs);
int mpcqp_scenario_get(mpcqp_handle* h, double* state /*B x 12*/, double* frame /*B x 3*/, double* feet /*B x 8*/);
int mpcqp_get_inputs(mpcqp_handle* h, double* xref, double* fsteps);

/* ---- logger-facing outputs ------------------------------------------------------------
 * Logger.log_cost_function (Logger.py:406-418): cost[B x 13], cost_i = sum over the horizon of x_i P_i x_i for the 12 state
 * components, cost_12 = the same over all forces.  (MPC.x_robot, Logger.py:455-461, is mpcqp_get_solution + xref.) */
int mpcqp_get_cost_components(mpcqp_handle* h, double* cost, int location);

/* ---- asynchronous result protocol (replaces the reference's unfinished process-based wrapper,
 * MPC_Wrapper.py:116-260): after mpcqp_run, mpcqp_result_async(h, slot) enqueues the copy of the B x 12 forces into the
 * handle's pinned slot 0 or 1 and returns immediately; mpcqp_result_wait(h, slot, forces) blocks until THAT copy has
 * landed.  Alternating the slots lets a control loop apply the forces of tick k-1 while tick k is being solved. */
int mpcqp_result_async(mpcqp_handle* h, int slot);
int mpcqp_result_ready(mpcqp_handle* h, int slot);      /* 1 landed, 0 in flight, < 0 error; never blocks */
int mpcqp_result_wait(mpcqp_handle* h, int slot, double* forces);

/* Measured FP64 peak of the device in TFLOP/s (DMMA m8n8k4 issue loop), for roofline reporting. */
int mpcqp_measure_fp64_peak(int device, double* dfma_tflops, double* dmma_tflops);

const char* mpcqp_last_error(void);
const char* mpcqp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MPCQP_H_ */
