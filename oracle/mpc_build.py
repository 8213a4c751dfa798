"""CPU restatement of the reference's per-tick QP build -- TEST INFRASTRUCTURE ONLY.

Restates, in this file's own vectorised form, what /root/reference/MPC.py computes:
    MPC.__init__                 MPC.py:22-82     constants (mass, gI, mu, default footholds)
    construct_gait               MPC.py:635-652   contact table from the NaN / 0.0 pattern of fsteps
    construct_S                  MPC.py:611-633   swing-pin flags
    create_ML / update_ML        MPC.py:98-190, 316-360   sparse constraint matrix (CSC, fixed pattern)
    create_NK / update_NK        MPC.py:192-234, 362-378  bounds
    create_weight_matrices       MPC.py:236-288   diagonal P, q = 0
    call_solver (warm start)     MPC.py:403-406
    retrieve_result              MPC.py:432-458
It is pinned against the reference itself: tests/golden/build_*.npz hold `ML.data`, `NK`, `NK_inf`
and `P.data` produced by importing the unmodified MPC.py (tests/golden/make_golden.py), and
tests/test_oracle_build.py compares this restatement with them.
"""
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp


@dataclass
class Params:
    dt: float = 0.02
    n_steps: int = 16
    T_gait: float = 0.32
    mass: float = 2.50000279                                    # MPC.py:28
    gI: np.ndarray = field(default_factory=lambda: np.array(    # MPC.py:35-37
        [[3.09249e-2, -8.00101e-7, 1.865287e-5],
         [-8.00101e-7, 5.106100e-2, 1.245813e-4],
         [1.865287e-5, 1.245813e-4, 6.939757e-2]]))
    mu: float = 0.9                                             # MPC.py:39
    fz_max: float = 25.0                                        # MPC.py:228
    gravity: float = 9.81                                       # MPC.py:201
    footholds: np.ndarray = field(default_factory=lambda: np.array(   # MPC.py:67-70
        [[0.19, 0.19, -0.19, -0.19],
         [0.15005, -0.15005, 0.15005, -0.15005],
         [0.0, 0.0, 0.0, 0.0]]))
    w_force: float = 1e-5                                       # MPC.py:282-284

    @property
    def w_state(self):
        # MPC.py:255-275
        w = np.zeros(12)
        w[0:3] = [0.1, 0.1, 1.0]
        w[3:6] = 0.11
        w[6:9] = 2.0 * np.sqrt(w[0:3])
        w[9:12] = 0.05 * np.sqrt(w[3:6])
        return w


def contact_table(fsteps, n_steps):
    """(N,4) contact flags and (N,) phase-row index per horizon step.  MPC.py:635-652, 611-633.

    A foot is in contact during phase row r iff its x coordinate is neither NaN nor exactly 0.0.
    Rows are consumed until the first row whose step count is 0."""
    fsteps = np.asarray(fsteps)
    contact = np.zeros((n_steps, 4))
    row_of_step = np.full(n_steps, -1, dtype=np.int64)
    k = 0
    for r in range(fsteps.shape[0]):
        cnt = fsteps[r, 0]
        if cnt == 0:
            break
        cnt = int(cnt)
        x = fsteps[r, 1::3]
        c = 1.0 - (np.isnan(x) | (x == 0.0))
        hi = min(k + cnt, n_steps)
        contact[k:hi] = c
        row_of_step[k:hi] = r
        k += cnt
    return contact, row_of_step


def _pattern(N):
    """Row indices / column pointers of the fixed CSC pattern (MPC.py:151; layout in SURVEY App. A)."""
    indices, indptr = [], [0]
    for k in range(N):
        for i in range(12):
            rows = [12 * k + i]
            if k < N - 1:
                if i >= 6:
                    rows.append(12 * (k + 1) + i - 6)
                rows.append(12 * (k + 1) + i)
            indices += rows
            indptr.append(len(indices))
    for k in range(N):
        for j in range(4):
            for c in range(3):
                rows = [12 * k + 6 + c, 12 * k + 9, 12 * k + 10, 12 * k + 11, 12 * N + 12 * k + 3 * j + c]
                base = 24 * N + 20 * k + 5 * j
                if c == 0:
                    rows += [base + 0, base + 1]
                elif c == 1:
                    rows += [base + 2, base + 3]
                else:
                    rows += [base + 0, base + 1, base + 2, base + 3, base + 4]
                indices += rows
                indptr.append(len(indices))
    return np.array(indices, dtype=np.int32), np.array(indptr, dtype=np.int32)


def lever_blocks(xref, fsteps, p: Params, first_tick=False):
    """dt * inv(R_z(yaw_k) gI) [r_kj]x for every (step, foot): (N,4,3,3).  MPC.py:170-182, 330-349.

    Quirks mirrored: inv(R gI) (not R gI R'), yaw and CoM from xref column k (column 0 = measured
    state), NaN footholds read as 0, tick 0 uses the default footholds."""
    N = p.n_steps
    fs = np.where(np.isnan(fsteps), 0.0, fsteps)
    _, row_of_step = contact_table(fsteps, N)
    out = np.zeros((N, 4, 3, 3))
    for k in range(N):
        c, s = np.cos(xref[5, k]), np.sin(xref[5, k])
        R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        I_inv = np.linalg.inv(R @ p.gI)
        if first_tick:
            feet = p.footholds
        elif row_of_step[k] >= 0:
            feet = fs[row_of_step[k], 1:].reshape(4, 3).T
        else:
            continue
        lever = feet - xref[0:3, k:k + 1]
        for j in range(4):
            r = lever[:, j]
            skew = np.array([[0.0, -r[2], r[1]], [r[2], 0.0, -r[0]], [-r[1], r[0], 0.0]])
            out[k, j] = p.dt * (I_inv @ skew)
    return out


def build_qp(xref, fsteps, p: Params = None, first_tick=False):
    """Returns (Pdiag, A (csc 44N x 24N), l, u, contact (N,4))."""
    p = p or Params()
    N, dt = p.n_steps, p.dt
    xref = np.asarray(xref, dtype=np.float64)
    fsteps = np.asarray(fsteps, dtype=np.float64)
    contact, _ = contact_table(fsteps, N)
    Bang = lever_blocks(xref, fsteps, p, first_tick)
    indices, indptr = _pattern(N)
    data = np.zeros(indices.shape[0])
    pos = 0
    for k in range(N):
        for i in range(12):
            data[pos] = -1.0
            pos += 1
            if k < N - 1:
                if i >= 6:
                    data[pos] = dt
                    pos += 1
                data[pos] = 1.0
                pos += 1
    assert pos == 30 * N - 18
    mu = p.mu
    for k in range(N):
        for j in range(4):
            for c in range(3):
                e = 1.0 - contact[k, j]
                col = [dt / p.mass, Bang[k, j, 0, c], Bang[k, j, 1, c], Bang[k, j, 2, c], e]
                col += ([1.0, -1.0] if c < 2 else [-mu, -mu, -mu, -mu, -1.0])
                data[pos:pos + len(col)] = col
                pos += len(col)
    A = sp.csc_matrix((data, indices, indptr), shape=(44 * N, 24 * N))

    # bounds: MPC.py:200-232, 362-378, 410
    Amat = np.eye(12)
    Amat[np.arange(6), np.arange(6) + 6] = dt
    g = np.zeros(12)
    g[8] = -p.gravity * dt
    u = np.zeros(44 * N)
    X = xref[:, 1:]
    for k in range(N):
        rhs = -g + X[:, k]
        rhs = rhs - Amat @ (xref[:, 0] if k == 0 else X[:, k - 1])
        u[12 * k:12 * k + 12] = rhs
    l = u.copy()
    l[24 * N:] = -np.inf
    l[24 * N + 4::5] = -p.fz_max
    Pdiag = np.concatenate([np.tile(p.w_state, N), np.full(12 * N, p.w_force)])
    return Pdiag, A, l, u, contact


def shift_warm_start(x_prev, N):
    """MPC.py:403-406: previous solution advanced by one stage; last state block zeroed, the old
    f_0 wraps into the last force block."""
    xs = np.roll(x_prev[:12 * N], -12).copy()
    xs[-12:] = 0.0
    fs = np.roll(x_prev[12 * N:], -12).copy()
    return np.concatenate([xs, fs])


def extract(x, xref, N):
    """MPC.py:432-450: predicted trajectory and the forces to apply now."""
    x_robot = x[:12 * N].reshape((12, N), order="F") + xref[:, 1:]
    f_applied = x[12 * N:12 * N + 12].copy()
    return f_applied, x_robot
