"""Vectorised KKT certificate for whole batches -- TEST INFRASTRUCTURE (numpy only, no sparse matrices).

tests/common.certify() builds the reference's sparse QP (oracle.mpc_build, pinned to the reference's own matrices) for ONE robot
and checks the full KKT system; at 20 ms per robot that covers a sample of a 4096- or 65 536-robot batch.  This module checks the
SAME conditions for every robot of a batch at once, in condensed form: with the dynamics multipliers eliminated by the costate
recursion  lam_s = Q e_s + A' lam_{s+1}  (MPC.py:110-111: A = [[I, dt I], [0, I]]), stationarity of the reference's QP
(Px + A'y = 0 with P, A of MPC.py:98-288) reduces, per stance foot-step, to
        w_f f + Bv' lam^v_{k+1} + C' y = 0,      C = the five pyramid rows of MPC.py:136-148,
next to the dynamics rows (primal equality), the pyramid rows (primal inequality), multiplier signs and complementarity.
test_batch_certificate_agrees_with_the_oracle (CPU) pins it to common.certify on the golden fixtures, including that it rejects
what the oracle rejects."""
import numpy as np

from oracle import mpc_build


def decode(xref, fsteps, p, first_tick=False):
    """contact (B, N, 4) bool and Bv angular blocks A3 (B, N, 4, 3, 3) = dt inv(R gI) [r]x      [MPC.py:316-360, 635-652]"""
    B, N = xref.shape[0], p.n_steps
    cnt = np.nan_to_num(fsteps[:, :, 0]).copy()
    dead = np.cumsum(cnt == 0.0, axis=1) > 0                        # rows from the first empty one on (MPC.py:646)
    cnt[dead] = 0.0
    cum = np.cumsum(cnt, axis=1)                                    # (B, 20)
    k = np.arange(N)
    row = (k[None, :, None] >= cum[:, None, :]).sum(axis=2)         # (B, N): phase row of step k
    valid = row < (~dead).sum(axis=1)[:, None]
    row = np.minimum(row, 19)
    feet = np.take_along_axis(fsteps[:, :, 1:], row[:, :, None], axis=1).reshape(B, N, 4, 3)
    contact = valid[:, :, None] & ~(np.isnan(feet[..., 0]) | (feet[..., 0] == 0.0))          # MPC.py:650
    feet = np.nan_to_num(feet)                                      # MPC.py:327
    if first_tick:
        feet = np.broadcast_to(p.footholds.T[None, None], (B, N, 4, 3))                     # MPC.py:176
    r = feet - xref[:, 0:3, :N].transpose(0, 2, 1)[:, :, None, :]   # lever arms (MPC.py:343)
    yaw = xref[:, 5, :N]
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.zeros((B, N, 3, 3))
    R[..., 0, 0], R[..., 0, 1], R[..., 1, 0], R[..., 1, 1], R[..., 2, 2] = c, -s, s, c, 1.0
    Iinv = np.linalg.inv(R @ p.gI)                                  # MPC.py:339-340: inv(R gI)
    skew = np.zeros((B, N, 4, 3, 3))
    skew[..., 0, 1], skew[..., 0, 2] = -r[..., 2], r[..., 1]
    skew[..., 1, 0], skew[..., 1, 2] = r[..., 2], -r[..., 0]
    skew[..., 2, 0], skew[..., 2, 1] = -r[..., 1], r[..., 0]
    A3 = p.dt * np.einsum("bnac,bnjcd->bnjad", Iinv, skew)
    return contact, A3


def certificate(xref, fsteps, x, y, p=None, first_tick=False):
    """-> dict of per-robot residuals (arrays of length B): dyn, prim, stat, comp, bad_sign, obj."""
    p = p or mpc_build.Params()
    B, N, dt, mu, fmax = xref.shape[0], p.n_steps, p.dt, p.mu, p.fz_max
    contact, A3 = decode(xref, fsteps, p, first_tick)
    f = x[:, 12 * N:].reshape(B, N, 4, 3)
    ym = y.reshape(B, N, 4, 5)
    X = x[:, :12 * N].reshape(B, N, 12) + xref[:, :, 1:].transpose(0, 2, 1)                  # states 1..N
    X0 = np.concatenate([xref[:, :, 0][:, None, :], X[:, :-1]], axis=1)                      # states 0..N-1
    lin = dt / p.mass
    u = lin * f.sum(axis=2)                                                                  # (B, N, 3) linear impulse (swing forces are 0)
    ua = np.einsum("bnjac,bnjc->bna", A3, f)
    pred = X0.copy()
    pred[..., 0:6] += dt * X0[..., 6:12]
    pred[..., 6:9] += u
    pred[..., 9:12] += ua
    pred[..., 8] -= p.gravity * dt                                                           # MPC.py:200-205
    dyn = np.abs(X - pred).reshape(B, -1).max(axis=1)
    # costates
    w = p.w_state
    e = x[:, :12 * N].reshape(B, N, 12)
    lam = np.zeros((B, N + 1, 12))
    for s_ in range(N - 1, -1, -1):                                                          # lam[s_] belongs to state s_ + 1
        nxt = lam[:, s_ + 1]
        lam[:, s_, 0:6] = w[0:6] * e[:, s_, 0:6] + nxt[:, 0:6]
        lam[:, s_, 6:12] = w[6:12] * e[:, s_, 6:12] + dt * nxt[:, 0:6] + nxt[:, 6:12]
    lv = lam[:, :N, 6:12]                                                                    # lam^v_{k+1}, k = 0..N-1
    grad = p.w_force * f + lin * lv[:, :, None, 0:3] + np.einsum("bnjca,bnc->bnja", A3, lv[..., 3:6])
    cty = np.stack([ym[..., 0] - ym[..., 1], ym[..., 2] - ym[..., 3],
                    -mu * (ym[..., 0] + ym[..., 1] + ym[..., 2] + ym[..., 3]) - ym[..., 4]], axis=-1)
    stat = (np.abs(grad + cty) * contact[..., None]).reshape(B, -1).max(axis=1)
    cf = np.stack([f[..., 0] - mu * f[..., 2], -f[..., 0] - mu * f[..., 2], f[..., 1] - mu * f[..., 2],
                   -f[..., 1] - mu * f[..., 2], -f[..., 2]], axis=-1)                        # rows <= 0, row 4 also >= -fmax
    viol = np.maximum(cf, 0.0).max(axis=-1)
    viol = np.maximum(viol, np.maximum(-fmax - cf[..., 4], 0.0))
    swing = np.abs(f).max(axis=-1) * (~contact)
    prim = np.maximum(viol * contact, swing).reshape(B, -1).max(axis=1)
    yc = ym * contact[..., None]
    comp4 = np.abs(yc[..., :4] * cf[..., :4]).max(axis=-1)
    comp5 = np.maximum(np.abs(np.maximum(yc[..., 4], 0.0) * cf[..., 4]), np.abs(np.minimum(yc[..., 4], 0.0) * (cf[..., 4] + fmax)))
    comp = np.maximum(comp4, comp5).reshape(B, -1).max(axis=1)
    bad_sign = np.maximum(-yc[..., :4], 0.0).reshape(B, -1).max(axis=1)
    wfull = np.concatenate([np.tile(w, N), np.full(12 * N, p.w_force)])
    obj = 0.5 * (x * x * wfull).sum(axis=1)
    return dict(dyn=dyn, prim=prim, stat=stat, comp=comp, bad_sign=bad_sign, obj=obj, contact=contact)


def assert_batch_certified(cert, where=""):
    """the bars of tests/common.assert_certified, for every robot"""
    for key, tol in (("dyn", 1e-8), ("prim", 1e-8), ("stat", 1e-10), ("comp", 1e-8), ("bad_sign", 1e-9)):
        worst = int(np.argmax(cert[key]))
        assert cert[key][worst] <= tol, (where, key, worst, float(cert[key][worst]))
