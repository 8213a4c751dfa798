"""ctypes binding of libmpcqp.so (include/mpcqp.h) -- the only door from Python into the engine.

The library is built in-tree by `__graft_entry__.build()` / `make -C mpc-tsid_b200/csrc`.  If it is
missing, or no B200 is visible, everything here raises: there is no CPU fallback by design.
"""
import ctypes as C
import os

import numpy as np

HOST, DEVICE = 0, 1
MODE_ACTIVE_SET, MODE_ADMM = 1, 2
MODE_STAGEWISE = 4           # active-set stage on the stage-wise (Riccati) factorisation, half a warp per robot
MODE_IPM = 8                 # fallback stage of the stage-wise path: interior-point iterations on the same factorisation (any horizon)
MODE_LANE = 16               # active-set stage with one lane per robot (opt-in variant; same answers, measured slower)
STATUS = {0: "unsolved", 1: "solved", 2: "max_iter", 3: "bad_input"}

# MPCQP_LIB selects another build of the same library (the debug build libmpcqp_canary.so of `make canary`)
_LIB_PATH = os.environ.get("MPCQP_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmpcqp.so")

# every symbol include/mpcqp.h declares (tests check that the built library exports all of them)
EXPORTS = (
    "mpcqp_default_params", "mpcqp_create", "mpcqp_destroy", "mpcqp_run", "mpcqp_get_latest_result",
    "mpcqp_get_solution", "mpcqp_get_info", "mpcqp_get_fallback_count", "mpcqp_reset_warm_start",
    "mpcqp_synchronize", "mpcqp_stream", "mpcqp_launch_count", "mpcqp_export_build",
    "mpcqp_measure_fp64_peak", "mpcqp_last_error", "mpcqp_version",
    "mpcqp_scenario_init", "mpcqp_scenario_run", "mpcqp_scenario_get", "mpcqp_get_inputs",
    "mpcqp_get_cost_components", "mpcqp_result_async", "mpcqp_result_ready", "mpcqp_result_wait",
    "mpcqp_get_step_result", "mpcqp_get_status", "mpcqp_host_alloc", "mpcqp_host_free",
    "mpcqp_world_pose", "mpcqp_scenario_set_commands", "mpcqp_set_overlap", "mpcqp_join",
)


class Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("n_steps", C.c_int32), ("batch", C.c_int32), ("device", C.c_int32),
        ("dt", C.c_double), ("T_gait", C.c_double), ("mass", C.c_double), ("mu", C.c_double),
        ("fz_max", C.c_double), ("gravity", C.c_double),
        ("gI", C.c_double * 9), ("footholds", C.c_double * 12), ("w_state", C.c_double * 12),
        ("w_force", C.c_double),
        ("mode", C.c_int32), ("max_sweeps", C.c_int32), ("max_iter", C.c_int32), ("min_iter", C.c_int32),
        ("check_every", C.c_int32), ("warm_start", C.c_int32),
        ("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
        ("feas_tol", C.c_double), ("dual_tol", C.c_double),
        ("refine", C.c_int32), ("ipm_max_iter", C.c_int32),
    ]


class MpcqpError(RuntimeError):
    pass


_lib = None


def load():
    """dlopen libmpcqp.so and declare prototypes.  Raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise MpcqpError("libmpcqp.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)" % _LIB_PATH)
    lib = C.CDLL(_LIB_PATH)
    vp, dp, i32p, u32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    lib.mpcqp_default_params.argtypes = [C.POINTER(Params)]
    lib.mpcqp_default_params.restype = None
    lib.mpcqp_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    lib.mpcqp_destroy.argtypes = [vp]
    lib.mpcqp_run.argtypes = [vp, C.c_double, dp, dp, C.c_int]
    lib.mpcqp_get_latest_result.argtypes = [vp, dp, C.c_int]
    lib.mpcqp_get_solution.argtypes = [vp, dp, C.c_int]
    lib.mpcqp_get_info.argtypes = [vp, i32p, i32p, i32p, dp, u32p, u32p, dp, C.c_int]
    lib.mpcqp_get_fallback_count.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.mpcqp_reset_warm_start.argtypes = [vp]
    lib.mpcqp_synchronize.argtypes = [vp]
    lib.mpcqp_stream.argtypes = [vp]
    lib.mpcqp_stream.restype = C.c_void_p
    lib.mpcqp_launch_count.argtypes = [vp]
    lib.mpcqp_launch_count.restype = C.c_int64
    lib.mpcqp_export_build.argtypes = [vp, C.c_double, dp, dp, C.c_int, dp, dp, dp]
    lib.mpcqp_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.mpcqp_scenario_init.argtypes = [vp, dp, dp, dp, dp, dp, C.c_uint64]
    lib.mpcqp_scenario_run.argtypes = [vp, C.c_int, C.c_int]
    lib.mpcqp_scenario_set_commands.argtypes = [vp, dp, C.c_int]
    lib.mpcqp_scenario_get.argtypes = [vp, dp, dp, dp]
    lib.mpcqp_get_inputs.argtypes = [vp, dp, dp]
    lib.mpcqp_get_cost_components.argtypes = [vp, dp, C.c_int]
    lib.mpcqp_result_async.argtypes = [vp, C.c_int]
    lib.mpcqp_result_ready.argtypes = [vp, C.c_int]
    lib.mpcqp_result_wait.argtypes = [vp, C.c_int, dp]
    lib.mpcqp_get_step_result.argtypes = [vp, dp, dp, C.c_int]
    lib.mpcqp_get_status.argtypes = [vp, i32p, C.c_int]
    lib.mpcqp_world_pose.argtypes = [vp, dp, C.c_int, C.c_int]
    lib.mpcqp_set_overlap.argtypes = [vp, C.c_int]
    lib.mpcqp_join.argtypes = [vp]
    lib.mpcqp_host_alloc.argtypes = [C.c_size_t]
    lib.mpcqp_host_alloc.restype = C.c_void_p
    lib.mpcqp_host_free.argtypes = [vp]
    lib.mpcqp_last_error.restype = C.c_char_p
    lib.mpcqp_version.restype = C.c_char_p
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise MpcqpError("libmpcqp error %d: %s" % (rc, load().mpcqp_last_error().decode()))


def default_params(**overrides):
    p = Params()
    load().mpcqp_default_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError("mpcqp_params has no field %r" % k)
        cur = getattr(p, k)
        if hasattr(cur, "__len__"):
            arr = np.asarray(v, dtype=np.float64).ravel()
            if len(arr) != len(cur):
                raise ValueError("%s needs %d values" % (k, len(cur)))
            for i, x in enumerate(arr):
                cur[i] = float(x)
        else:
            setattr(p, k, v)
    return p


def _ptr(a):
    """Host numpy array or a raw device pointer (int)."""
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return C.c_void_p(a.ctypes.data)


def _host_f64(a, shape):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != shape:
        raise ValueError("expected shape %r, got %r" % (shape, a.shape))
    return a


def bind_near_gpu(device=0):
    """Pin the calling thread (and the threads it starts later) to the CPU cores NVML reports as local to CUDA device
    `device`, so that the page-locked input / output buffers it allocates afterwards land on that GPU's NUMA node and the
    per-tick copies do not cross the socket interconnect.  One process per GPU: call it first thing in each process.
    Returns {"cpus": n, "first": i, "last": j} or {"error": "..."} (never raises: containers may forbid it)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(int(device)).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return {"cpus": len(cpus), "first": cpus[0], "last": cpus[-1]}
    except Exception as e:          # noqa: BLE001 -- best effort by design
        return {"error": "%s: %s" % (type(e).__name__, e)}


class Engine:
    """One libmpcqp handle: a batch of `batch` independent MPC instances on one GPU."""

    def __init__(self, batch=1, n_steps=16, device=0, **overrides):
        self.lib = load()
        self.params = default_params(batch=int(batch), n_steps=int(n_steps), device=int(device), **overrides)
        self.B, self.N = int(batch), int(n_steps)
        h = C.c_void_p()
        _check(self.lib.mpcqp_create(C.byref(self.params), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mpcqp_synchronize(self._h)
            for ptr in getattr(self, "_pinned", []):
                self.lib.mpcqp_host_free(ptr)
            self._pinned = []
            self.lib.mpcqp_destroy(self._h)
            self._h = None

    def pinned(self, shape, dtype=np.float64):
        """A page-locked numpy array owned by this engine (freed by close()): what xref / fsteps / forces should live in."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = self.lib.mpcqp_host_alloc(n)
        if not ptr:
            raise MpcqpError("mpcqp_host_alloc(%d) failed: %s" % (n, self.lib.mpcqp_last_error().decode()))
        if not hasattr(self, "_pinned"):
            self._pinned = []
        self._pinned.append(ptr)
        return np.frombuffer((C.c_char * n).from_address(ptr), dtype=dtype).reshape(shape)

    __del__ = close

    # ---- run
    def run(self, k, xref, fsteps):
        """Host arrays: xref (B,12,N+1), fsteps (B,20,13).  Asynchronous."""
        xref = _host_f64(xref, (self.B, 12, self.N + 1))
        fsteps = _host_f64(fsteps, (self.B, 20, 13))
        self._keep = (xref, fsteps)          # keep alive until the copy engine is done with them
        _check(self.lib.mpcqp_run(self._h, float(k), _ptr(xref), _ptr(fsteps), HOST))

    def run_device(self, k, xref_ptr, fsteps_ptr):
        """Raw device pointers (e.g. torch.Tensor.data_ptr()) to arrays of the same layout."""
        _check(self.lib.mpcqp_run(self._h, float(k), C.c_void_p(int(xref_ptr)), C.c_void_p(int(fsteps_ptr)), DEVICE))

    # ---- results
    def forces(self, out=None):
        out = np.empty((self.B, 12)) if out is None else out
        _check(self.lib.mpcqp_get_latest_result(self._h, _ptr(out), HOST))
        return out

    def step_result(self, forces=None, next_state=None):
        """Forces (B,12) and the first predicted state x_robot[:, 0] (B,12) of the last run in one call / one synchronisation."""
        forces = np.empty((self.B, 12)) if forces is None else forces
        next_state = np.empty((self.B, 12)) if next_state is None else next_state
        _check(self.lib.mpcqp_get_step_result(self._h, _ptr(forces), _ptr(next_state), HOST))
        return forces, next_state

    def world_pose(self, set_to=None):
        """MPC.q_w of every robot, (B, 6): read it, or overwrite it with `set_to`."""
        if set_to is not None:
            q = _host_f64(set_to, (self.B, 6))
            _check(self.lib.mpcqp_world_pose(self._h, _ptr(q), 1, HOST))
            return q
        q = np.empty((self.B, 6))
        _check(self.lib.mpcqp_world_pose(self._h, _ptr(q), 0, HOST))
        return q

    def status(self):
        st = np.empty(self.B, np.int32)
        _check(self.lib.mpcqp_get_status(self._h, _ptr(st), HOST))
        return st

    def forces_device(self, ptr):
        _check(self.lib.mpcqp_get_latest_result(self._h, C.c_void_p(int(ptr)), DEVICE))

    def solution(self, out=None):
        out = np.empty((self.B, 24 * self.N)) if out is None else out
        _check(self.lib.mpcqp_get_solution(self._h, _ptr(out), HOST))
        return out

    def solution_device(self, ptr):
        _check(self.lib.mpcqp_get_solution(self._h, C.c_void_p(int(ptr)), DEVICE))

    def info(self, with_y=True):
        B, N = self.B, self.N
        aw, cw = (20 * N + 31) // 32, (4 * N + 31) // 32
        status, sweeps, iters = (np.empty(B, np.int32) for _ in range(3))
        obj = np.empty(B)
        contact, active = np.empty((B, cw), np.uint32), np.empty((B, aw), np.uint32)
        y = np.empty((B, 20 * N)) if with_y else None
        _check(self.lib.mpcqp_get_info(self._h, _ptr(status), _ptr(sweeps), _ptr(iters), _ptr(obj), _ptr(contact),
                                       _ptr(active), _ptr(y) if with_y else None, HOST))
        bits = lambda words, n: ((words[:, np.arange(n) // 32] >> (np.arange(n) % 32).astype(np.uint32)) & 1).astype(bool)
        return dict(status=status, sweeps=sweeps, iters=iters, obj=obj, y=y,
                    contact=bits(contact, 4 * N).reshape(B, N, 4), active=bits(active, 20 * N).reshape(B, N, 4, 5))

    def fallback_count(self):
        c = C.c_int32()
        _check(self.lib.mpcqp_get_fallback_count(self._h, C.byref(c)))
        return c.value

    def export_build(self, k, xref, fsteps):
        xref = _host_f64(xref, (self.B, 12, self.N + 1))
        fsteps = _host_f64(fsteps, (self.B, 20, 13))
        Bv, Sv, NK = np.empty((self.B, self.N, 48)), np.empty((self.B, 12 * self.N)), np.empty((self.B, 12 * self.N))
        _check(self.lib.mpcqp_export_build(self._h, float(k), _ptr(xref), _ptr(fsteps), HOST, _ptr(Bv), _ptr(Sv), _ptr(NK)))
        return Bv, Sv, NK

    # ---- device-resident closed loop (planner + integration inside the solve kernel)
    def scenario_init(self, scen):
        """Take the robots of a host `scenario.Scenario` (gaits, phases, commands, initial states, noise
        levels, seed; it must use noise_kind="hash" or zero noise to stay comparable) onto the device."""
        if scen.B != self.B or scen.N != self.N:
            raise ValueError("scenario and engine disagree on batch / horizon")
        seq = np.ascontiguousarray(scen.seq_bits(), dtype=np.uint64)
        phase = np.ascontiguousarray(scen.phase, dtype=np.int32)
        vref = np.ascontiguousarray(scen.current_v_ref(), dtype=np.float64)
        state = np.ascontiguousarray(scen.state, dtype=np.float64)
        sigma = np.ascontiguousarray(scen.noise, dtype=np.float64)
        _check(self.lib.mpcqp_scenario_init(self._h, _ptr(seq), _ptr(phase), _ptr(vref), _ptr(state), _ptr(sigma),
                                            C.c_uint64(int(scen.seed))))
        if getattr(scen, "reduced", False):
            self.scenario_set_commands(None, True)

    def scenario_set_commands(self, v_ref=None, reduced=False):
        """Joystick commands for the following ticks: v_ref (B, 6) or (6,) or None (unchanged), and the `reduced` switch."""
        v = None if v_ref is None else np.ascontiguousarray(np.broadcast_to(np.asarray(v_ref, dtype=np.float64), (self.B, 6)))
        _check(self.lib.mpcqp_scenario_set_commands(self._h, None if v is None else _ptr(v), int(bool(reduced))))

    def scenario_run(self, ticks, emit_inputs=False):
        _check(self.lib.mpcqp_scenario_run(self._h, int(ticks), 1 if emit_inputs else 0))

    def scenario_state(self):
        B = self.B
        state, frame, feet = np.empty((B, 12)), np.empty((B, 3)), np.empty((B, 2, 4))
        _check(self.lib.mpcqp_scenario_get(self._h, _ptr(state), _ptr(frame), _ptr(feet)))
        return dict(state=state, frame=frame, feet=feet)

    def last_inputs(self):
        xref, fsteps = np.empty((self.B, 12, self.N + 1)), np.empty((self.B, 20, 13))
        _check(self.lib.mpcqp_get_inputs(self._h, _ptr(xref), _ptr(fsteps)))
        return xref, fsteps

    def cost_components(self):
        """Logger.log_cost_function (Logger.py:406-418) per robot: (B, 13)."""
        out = np.empty((self.B, 13))
        _check(self.lib.mpcqp_get_cost_components(self._h, _ptr(out), HOST))
        return out

    def result_async(self, slot):
        """Enqueue the copy of the forces of the run just issued into pinned slot 0 / 1; returns at once."""
        _check(self.lib.mpcqp_result_async(self._h, int(slot)))

    def result_ready(self, slot):
        rc = self.lib.mpcqp_result_ready(self._h, int(slot))
        if rc < 0:
            _check(rc)
        return bool(rc)

    def result_wait(self, slot, out=None):
        out = np.empty((self.B, 12)) if out is None else out
        _check(self.lib.mpcqp_result_wait(self._h, int(slot), _ptr(out)))
        return out

    def reset_warm_start(self):
        _check(self.lib.mpcqp_reset_warm_start(self._h))

    def synchronize(self):
        _check(self.lib.mpcqp_synchronize(self._h))

    def set_overlap(self, ranges):
        """Issue device-resident ticks as `ranges` independent index ranges (2 .. 4) so that consecutive ticks overlap; 1 = off,
        0 = automatic (inside scenario_run only).  Results are unchanged; see mpcqp_set_overlap in include/mpcqp.h."""
        _check(self.lib.mpcqp_set_overlap(self._h, int(ranges)))

    def join(self):
        """Make the engine's stream wait for every index range in flight (needed only before enqueuing own work on `stream`)."""
        _check(self.lib.mpcqp_join(self._h))

    @property
    def stream(self):
        return self.lib.mpcqp_stream(self._h)

    @property
    def launches(self):
        return int(self.lib.mpcqp_launch_count(self._h))


def measure_fp64_peak(device=0):
    a, b = C.c_double(), C.c_double()
    _check(load().mpcqp_measure_fp64_peak(int(device), C.byref(a), C.byref(b)))
    return dict(dfma_tflops=a.value, dmma_tflops=b.value)
