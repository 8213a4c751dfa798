"""Headless, batched producer of the MPC inputs (`xref`, `fsteps`) -- host side, numpy only.

The reference feeds its MPC from a footstep planner driven by a joystick and a simulator
(main.py:77-104, processing.py:63-145).  The simulator half (PyBullet, pinocchio, TSID) is outside
the hot path, so benchmarks and tests close the loop on the centroidal model itself instead
(SURVEY.md section 8d).  This module restates, vectorised over a batch of independent robots, the
parts of the planner that define what the MPC sees:

    gait tables + roll            FootstepPlanner.py:207-282, 401-425
    compute_footsteps             FootstepPlanner.py:284-361
    compute_next_footstep         FootstepPlanner.py:363-399
    getRefStates                  FootstepPlanner.py:76-161   (including the vz / roll / pitch command state machine, :128-152)
    velocity ramp                 Joystick.py:82-83

and a closed-loop state update: the next measured state is the MPC's own one-step prediction
(MPC.py:448-450) re-expressed in the next yaw-aligned local frame (the frame rules of
Interface.py:100-132), plus seeded Gaussian noise.  Stance feet stay fixed in the world; a foot
that touches down lands on the planner's target.

Everything here is float64 and produces arrays in exactly the layout MPC.run consumes:
    xref   (B, 12, N+1)   column 0 = measured state, columns 1..N = reference
    fsteps (B, 20, 13)    [steps left in phase | (x, y, z) x 4 feet], NaN for swing feet
"""
import numpy as np

SHOULDERS = np.array([[0.19, 0.19, -0.19, -0.19],
                      [0.15005, -0.15005, 0.15005, -0.15005]])   # FootstepPlanner.py:23-24
H_REF = 0.2027682                                                 # processing.py:127
K_FEEDBACK = 0.03                                                 # FootstepPlanner.py:20
LEG_L = 0.12                                                      # FootstepPlanner.py:33
T_STANCE = 0.16                                                   # FootstepPlanner.py:377
G = 9.81
MAX_ROWS = 20                                                     # FootstepPlanner.py:62, MPC.py:82
CMD_STEP = 0.05                                                   # FootstepPlanner.py:131: joystick dead band of the vz command
H_ROTATION0 = 0.20                                                # FootstepPlanner.py:69: h_rotation_command at start
REDUCED_OFFSET = np.array([[0.14, 0.14, -0.14, -0.14],
                           [0.12, -0.12, 0.12, -0.12]])           # FootstepPlanner.py:331-332 (`reduced` support polygon)

# per-step contact patterns over one gait period of 16 steps, feet = FL, FR, HL, HR
_ALL = (1, 1, 1, 1)


def _two_beat(a, b, half):
    return [_ALL] + [a] * (half - 1) + [_ALL] + [b] * (half - 1)


def gait_sequence(kind, half=8):
    """One period of per-step contact flags, shape (2*half, 4).

    trot / bound / pace follow FootstepPlanner.py:220-229, 252-254, 278-280 ([1, N-1, 1, N-1]).
    `walk` is not in the reference: four phases of half/2 steps, one foot swinging at a time
    (FL, HR, FR, HL).  `static` is FootstepPlanner.py:186-205 (all feet down)."""
    if kind == "trot":
        seq = _two_beat((1, 0, 0, 1), (0, 1, 1, 0), half)
    elif kind == "bound":
        seq = _two_beat((1, 1, 0, 0), (0, 0, 1, 1), half)
    elif kind == "pace":
        seq = _two_beat((1, 0, 1, 0), (0, 1, 0, 1), half)
    elif kind == "walk":
        q = half // 2
        seq = [(0, 1, 1, 1)] * q + [(1, 1, 1, 0)] * q + [(1, 0, 1, 1)] * q + [(1, 1, 0, 1)] * q
    elif kind == "static":
        seq = [_ALL] * (2 * half)
    else:
        raise ValueError("unknown gait %r" % (kind,))
    return np.array(seq, dtype=np.float64)


GAIT_KINDS = ("trot", "pace", "bound", "walk", "static")


def reference_ramp(tick):
    """Joystick.py:82-83 (predefined forward-velocity ramp), `tick` counted in MPC ticks (20 TSID
    ticks each: k_loop = 20 * tick)."""
    return min(max((20.0 * tick - 960.0) / 3500.0, 0.0), 1.0)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def hash_normal(seed, inst, tick, comp):
    """Standard normal as a pure function of (seed, robot, tick, component): splitmix64 + Box-Muller.
    Bit-for-bit the integer path of mpcqp_scenario.cuh::hash_normal (the libm calls may differ by an ulp)."""
    with np.errstate(over="ignore"):
        key = _splitmix64(seed + inst * np.uint64(0x9E3779B97F4A7C15) + tick * np.uint64(0xD1B54A32D192ED03)
                          + comp * np.uint64(0x8CB92BA72F3D8DD7))
        h1, h2 = _splitmix64(key), _splitmix64(key ^ np.uint64(0xA5A5A5A5A5A5A5A5))
    u1 = ((h1 >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
    u2 = ((h2 >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586 * u2)


class Scenario:
    """B independent robots, each with its own gait, gait phase, velocity command and noise stream."""

    def __init__(self, batch, n_steps=16, dt=0.02, T_gait=0.32, gaits="trot", seed=20260,
                 v_ref=None, phase=None, noise=(1e-3, 5e-3, 1e-2, 2e-2), random_commands=True,
                 ramp=False, noise_kind="numpy", reduced=False):
        self.B, self.N, self.dt, self.T_gait = int(batch), int(n_steps), float(dt), float(T_gait)
        B, N = self.B, self.N
        self.period = int(round(T_gait / dt))
        half = self.period // 2
        if isinstance(gaits, str):
            gaits = [gaits]
        self.kinds = [gaits[i % len(gaits)] for i in range(B)]
        table = {k: gait_sequence(k, half) for k in set(self.kinds)}
        self.seq = np.stack([table[k] for k in self.kinds])                   # (B, period, 4)
        self.rng = [np.random.default_rng([seed, i]) for i in range(B)] if B <= 4096 else None
        self._bulk_rng = np.random.default_rng([seed, B])
        if phase is None:
            phase = self._bulk_rng.integers(0, self.period, size=B) if random_commands else np.zeros(B, int)
        self.phase = np.asarray(phase, dtype=np.int64) % self.period
        self.v_ref = np.zeros((B, 6))
        if v_ref is not None:
            self.v_ref[:] = np.asarray(v_ref, dtype=np.float64)
        elif random_commands:
            self.v_ref[:, 0] = self._bulk_rng.uniform(-0.5, 1.0, B)
            self.v_ref[:, 1] = self._bulk_rng.uniform(-0.3, 0.3, B)
            self.v_ref[:, 5] = self._bulk_rng.uniform(-0.4, 0.4, B)
        self.ramp = bool(ramp)
        self.reduced = bool(reduced)                                         # Joystick.reduced (processing.py:78-88)
        # getRefStates' command state machine, one per robot (FootstepPlanner.py:66-69, 128-152): flag_rotation_command,
        # h_rotation_command, and the two reference rows it leaves untouched between ticks (xref[2, 1:], xref[8, 1:])
        self.cmd_flag = np.zeros(B, dtype=np.int64)
        self.h_rot = np.full(B, H_ROTATION0)
        self.z_ref = np.full(B, H_REF)
        self.vz_ref = np.zeros(B)
        self.noise = np.asarray(noise, dtype=np.float64)
        # "numpy": one numpy Generator per robot (configs[0..2]); "hash": a counter-based generator keyed by
        # (seed, robot, tick, component) that the device-resident closed loop reproduces (mpcqp_scenario.cuh)
        self.noise_kind = noise_kind
        self.seed = int(seed)
        self.tick = 0
        # measured state in the local frame: [x, y, z, roll, pitch, yaw, vx, vy, vz, wx, wy, wz]
        self.state = np.zeros((B, 12))
        self.state[:, 2] = H_REF
        # local frame in the world (x, y, yaw) and feet in the world (B, 2, 4)
        self.frame = np.zeros((B, 3))
        self.feet_w = np.broadcast_to(SHOULDERS, (B, 2, 4)).copy()
        self.xref = np.zeros((B, 12, N + 1))
        self.fsteps = np.full((B, MAX_ROWS, 13), np.nan)
        self.fsteps[:, :, 0] = 0.0
        self._prev_contact0 = None
        self._touchdown_target_w = self.feet_w.copy()

    # ------------------------------------------------------------------ gait tables
    def step_contacts(self):
        """Per-step contact flags over the horizon for the current tick: (B, N, 4)."""
        idx = (self.tick + self.phase[:, None] + np.arange(self.N)[None, :]) % self.period
        return np.take_along_axis(self.seq, idx[:, :, None], axis=1)

    @staticmethod
    def run_length_rows(contacts):
        """(B, N, 4) per-step flags -> gait table (B, 20, 5) = what `roll` maintains
        (FootstepPlanner.py:401-425): consecutive identical steps merged into one phase row."""
        B, N, _ = contacts.shape
        gait = np.zeros((B, MAX_ROWS, 5))
        row = np.zeros(B, dtype=np.int64)
        ar = np.arange(B)
        gait[:, 0, 1:] = contacts[:, 0]
        gait[:, 0, 0] = 1.0
        for i in range(1, N):
            same = np.all(contacts[:, i] == contacts[:, i - 1], axis=1)
            row = row + (~same)
            if np.any(row >= MAX_ROWS):
                raise ValueError("gait needs more than %d phase rows" % MAX_ROWS)
            gait[ar, row, 1:] = contacts[:, i]
            gait[ar, row, 0] += 1.0
        return gait

    # ------------------------------------------------------------------ planner
    def _next_footstep(self, v_ref, h):
        """FootstepPlanner.py:363-399 with v_cur = v_ref (as compute_footsteps calls it, :322)."""
        B = v_ref.shape[0]
        nf = np.zeros((B, 3, 4))
        nf[:, 0:2, :] += (T_STANCE * 0.5 * v_ref[:, 0:2])[:, :, None]
        nf[:, 0:2, :] += (K_FEEDBACK * (v_ref[:, 0:2] - v_ref[:, 0:2]))[:, :, None]
        cross = np.cross(v_ref[:, 0:3], v_ref[:, 3:6])
        nf[:, 0:2, :] += (0.5 * np.sqrt(h / G))[:, None, None] * cross[:, 0:2, None]
        nf[:, 0:2, :] = np.clip(nf[:, 0:2, :], -LEG_L, LEG_L)
        nf[:, 0:2, :] += SHOULDERS[None]
        if self.reduced:                                                     # FootstepPlanner.py:330-332
            nf[:, 0:2, :] -= REDUCED_OFFSET[None]
        return nf

    def _compute_footsteps(self, gait, l_feet, v_cur, v_ref, h):
        """FootstepPlanner.py:284-361 (reduced = False)."""
        B = self.B
        fs = np.full((B, MAX_ROWS, 13), np.nan)
        fs[:, :, 0] = gait[:, :, 0]
        st = np.repeat(gait[:, :, 1:] == 1.0, 3, axis=2)                       # (B, rows, 12)
        feet_flat = np.transpose(l_feet, (0, 2, 1)).reshape(B, 12)             # ravel(order='F')
        fs[:, 0, 1:] = np.where(st[:, 0], feet_flat, np.nan)
        nf = self._next_footstep(v_ref, h)
        dt_cum = np.zeros(B)
        alive = np.ones(B, dtype=bool)
        w = v_ref[:, 5]
        safe_w = np.where(w != 0.0, w, 1.0)
        for i in range(1, MAX_ROWS):
            alive = alive & (gait[:, i, 0] != 0.0)
            if not np.any(alive):
                break
            dt_cum = dt_cum + gait[:, i - 1, 0] * self.dt
            keep = st[:, i - 1] & st[:, i]
            new = (~st[:, i - 1]) & st[:, i]
            ang = w * dt_cum
            c, s = np.cos(ang), np.sin(ang)
            dx = np.where(w != 0.0, (v_cur[:, 0] * np.sin(w * dt_cum) + v_cur[:, 1] * (np.cos(w * dt_cum) - 1.0)) / safe_w,
                          v_cur[:, 0] * dt_cum)
            dy = np.where(w != 0.0, (v_cur[:, 1] * np.sin(w * dt_cum) - v_cur[:, 0] * (np.cos(w * dt_cum) - 1.0)) / safe_w,
                          v_cur[:, 1] * dt_cum)
            tx = c[:, None] * nf[:, 0, :] - s[:, None] * nf[:, 1, :] + dx[:, None]
            ty = s[:, None] * nf[:, 0, :] + c[:, None] * nf[:, 1, :] + dy[:, None]
            tz = nf[:, 2, :]
            target = np.stack([tx, ty, tz], axis=2).reshape(B, 12)             # ravel(order='F')
            rowv = np.where(keep, fs[:, i - 1, 1:], np.nan)
            rowv = np.where(new, target, rowv)
            fs[:, i, 1:] = np.where(alive[:, None], rowv, np.nan)
        return fs

    def _ref_states(self, v_ref):
        """FootstepPlanner.py:76-161, including the state machine of the vz / roll / pitch commands (:128-152): it keeps
        per-robot state (cmd_flag, h_rot) and leaves xref[2, 1:] / xref[8, 1:] as they were when no branch rewrites them."""
        B, N, dt = self.B, self.N, self.dt
        xr = np.zeros((B, 12, N + 1))
        to = np.linspace(0.0, self.T_gait - dt, N)[None, :]
        yaw = to * v_ref[:, 5:6]
        xr[:, 6, 1:] = v_ref[:, 0:1] * np.cos(yaw) - v_ref[:, 1:2] * np.sin(yaw)
        xr[:, 7, 1:] = v_ref[:, 0:1] * np.sin(yaw) + v_ref[:, 1:2] * np.cos(yaw)
        xr[:, 0, 1:] = dt * np.cumsum(xr[:, 6, 1:], axis=1) + self.state[:, 0:1]
        xr[:, 1, 1:] = dt * np.cumsum(xr[:, 7, 1:], axis=1) + self.state[:, 1:2]
        xr[:, 5, 1:] = v_ref[:, 5:6] * np.linspace(dt, self.T_gait, N)[None, :]
        xr[:, 11, 1:] = v_ref[:, 5:6]
        xr[:, :, 0] = self.state
        # ---- command state machine
        vz = v_ref[:, 2]
        big, small = np.abs(vz) > CMD_STEP, np.abs(vz) < CMD_STEP
        self.cmd_flag = np.where(big & (self.cmd_flag != 1), 1, self.cmd_flag)                  # :134-135
        commanding = big & (self.cmd_flag == 1)                                                 # :138-143
        releasing = small & (self.cmd_flag == 1)                                                # :144-148
        idle = ~commanding & ~releasing & (self.cmd_flag == 0)                                  # :149-151
        self.h_rot = np.where(commanding, self.h_rot + vz * dt, self.h_rot)
        self.z_ref = np.where(commanding, self.h_rot, np.where(idle, H_REF, self.z_ref))
        self.vz_ref = np.where(commanding, vz, np.where(releasing | idle, 0.0, self.vz_ref))
        self.cmd_flag = np.where(releasing, 2, self.cmd_flag)
        xr[:, 2, 1:] = self.z_ref[:, None]
        xr[:, 8, 1:] = self.vz_ref[:, None]
        on = (self.cmd_flag != 0)[:, None]                                                      # :153-158
        xr[:, 3, 1:] = np.where(on, self.state[:, 3:4] + v_ref[:, 3:4] * to, 0.0)
        xr[:, 4, 1:] = np.where(on, self.state[:, 4:5] + v_ref[:, 4:5] * to, 0.0)
        xr[:, 9, 1:] = np.where(on, v_ref[:, 3:4], 0.0)
        xr[:, 10, 1:] = np.where(on, v_ref[:, 4:5], 0.0)
        return xr

    def set_v_ref(self, v_ref):
        """New joystick commands (Joystick.update_v_ref, Joystick.py:29-43) from the next inputs() on: (6,) or (B, 6)."""
        self.v_ref[:] = np.asarray(v_ref, dtype=np.float64)

    # ------------------------------------------------------------------ loop
    def current_v_ref(self):
        if self.ramp:
            v = np.zeros((self.B, 6))
            v[:, 0] = reference_ramp(self.tick)
            return v
        return self.v_ref

    def local_feet(self):
        """World feet expressed in the current local frame: (B, 3, 4), z = 0."""
        d = self.feet_w - self.frame[:, 0:2, None]
        c, s = np.cos(self.frame[:, 2]), np.sin(self.frame[:, 2])
        lx = c[:, None] * d[:, 0] + s[:, None] * d[:, 1]
        ly = -s[:, None] * d[:, 0] + c[:, None] * d[:, 1]
        return np.stack([lx, ly, np.zeros_like(lx)], axis=1)

    def inputs(self):
        """Planner pass for the current tick.  Returns (xref, fsteps); both are fresh arrays."""
        contacts = self.step_contacts()
        gait = self.run_length_rows(contacts)
        v_ref = self.current_v_ref()
        # touchdown: a foot in stance now that was swinging at the previous tick lands on its target
        c0 = contacts[:, 0] == 1.0
        if self._prev_contact0 is not None:
            landed = c0 & ~self._prev_contact0
            self.feet_w = np.where(landed[:, None, :], self._touchdown_target_w, self.feet_w)
        self._prev_contact0 = c0
        l_feet = self.local_feet()
        v_cur = self.state[:, 6:12]
        self.fsteps = self._compute_footsteps(gait, l_feet, v_cur, v_ref, self.state[:, 2])
        self.xref = self._ref_states(v_ref)
        # remember, in the world, where each swinging foot is planned to land (first row where it is down)
        tgt = np.full((self.B, 2, 4), np.nan)
        got = np.zeros((self.B, 4), dtype=bool)
        for r in range(1, MAX_ROWS):
            xy = self.fsteps[:, r, 1:].reshape(self.B, 4, 3)
            ok = ~np.isnan(xy[:, :, 0]) & ~got & (self.fsteps[:, r, 0:1] != 0.0)
            tgt[:, 0, :] = np.where(ok, xy[:, :, 0], tgt[:, 0, :])
            tgt[:, 1, :] = np.where(ok, xy[:, :, 1], tgt[:, 1, :])
            got |= ok
        c, s = np.cos(self.frame[:, 2]), np.sin(self.frame[:, 2])
        wx = c[:, None] * tgt[:, 0] - s[:, None] * tgt[:, 1] + self.frame[:, 0:1]
        wy = s[:, None] * tgt[:, 0] + c[:, None] * tgt[:, 1] + self.frame[:, 1:2]
        tw = np.stack([wx, wy], axis=1)
        self._touchdown_target_w = np.where(np.isnan(tw), self._touchdown_target_w, tw)
        return self.xref.copy(), self.fsteps.copy()

    def seq_bits(self):
        """The gait period as ceil(period / 16) 64-bit words per robot: bit 4*(s % 16) + j of word s // 16 = foot j in contact
        at step s (libmpcqp's format; shape (B,) while the period fits one word, else (B, words))."""
        words = (self.period + 15) // 16
        bits = np.zeros((self.B, words), dtype=np.uint64)
        for s_ in range(self.period):
            for j in range(4):
                bits[:, s_ // 16] |= (self.seq[:, s_, j] == 1.0).astype(np.uint64) << np.uint64(4 * (s_ % 16) + j)
        return bits[:, 0] if words == 1 else bits

    def _noise(self):
        B = self.B
        if self.noise_kind == "hash":
            inst = np.arange(B, dtype=np.uint64)[:, None]
            comp = np.arange(12, dtype=np.uint64)[None, :]
            n = hash_normal(np.uint64(self.seed), inst, np.uint64(self.tick), comp)
            return n * np.repeat(self.noise, 3)[None, :]
        if self.rng is not None:
            n = np.stack([r.standard_normal(12) for r in self.rng])
        else:
            n = self._bulk_rng.standard_normal((B, 12))
        sig = np.repeat(self.noise, 3)
        return n * sig[None, :]

    def advance(self, x_next):
        """Close the loop: `x_next` (B, 12) is the MPC's predicted next state in the CURRENT local
        frame (x_robot[:, 0], MPC.py:448-450).  Moves the local frame to the predicted (x, y, yaw),
        rotates velocities into it, adds noise, and steps the gait."""
        x_next = np.asarray(x_next, dtype=np.float64).reshape(self.B, 12)
        c, s = np.cos(self.frame[:, 2]), np.sin(self.frame[:, 2])
        self.frame[:, 0] += c * x_next[:, 0] - s * x_next[:, 1]
        self.frame[:, 1] += s * x_next[:, 0] + c * x_next[:, 1]
        dyaw = x_next[:, 5]
        self.frame[:, 2] += dyaw
        cy, sy = np.cos(dyaw), np.sin(dyaw)
        st = np.zeros((self.B, 12))
        st[:, 2:5] = x_next[:, 2:5]
        st[:, 6] = cy * x_next[:, 6] + sy * x_next[:, 7]
        st[:, 7] = -sy * x_next[:, 6] + cy * x_next[:, 7]
        st[:, 8] = x_next[:, 8]
        st[:, 9] = cy * x_next[:, 9] + sy * x_next[:, 10]
        st[:, 10] = -sy * x_next[:, 9] + cy * x_next[:, 10]
        st[:, 11] = x_next[:, 11]
        noise = self._noise()
        noise[:, 0:2] = 0.0       # the local frame is centred on the robot by construction
        noise[:, 5] = 0.0         # ... and yaw-aligned with it
        self.state = st + noise
        self.tick += 1
