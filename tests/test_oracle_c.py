"""oracle/mpc_osqp.c (the plain-C restatement: build + OSQP algorithm + extraction) against
 (a) what the UNMODIFIED reference MPC.py built and extracted (tests/golden/solve_*.npz),
 (b) the numpy/scipy restatement oracle/osqp_port.py, iteration for iteration,
 (c) problems with known answers, and the KKT certificate after the active-set polish."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import c_port, kkt, mpc_build

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "solve_*.npz")))
IDS = [os.path.basename(p)[6:-4] for p in GOLD]


def test_library_exports():
    lib = c_port.load()
    for sym in ("mpc_oracle_create", "mpc_oracle_destroy", "mpc_oracle_build", "mpc_oracle_run", "mpc_oracle_replay_mt",
                "mpc_oracle_solve_qp", "mpc_oracle_nnz"):
        assert getattr(lib, sym)


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_c_build_matches_reference(path):
    g = np.load(path)
    N = g["x"].shape[1] // 24
    m = c_port.MPC(n_steps=N)
    assert m.nnz == 126 * N - 18
    for t in range(len(g["ML_data"])):
        Ap, Ai, Ax, l, u = m.build(g["xref"][t], g["fsteps"][t], first_tick=(g["k"][t] == 0))
        assert np.array_equal(Ai, g["ML_indices"]) and np.array_equal(Ap, g["ML_indptr"])
        np.testing.assert_allclose(Ax, g["ML_data"][t], rtol=0, atol=1e-15)
        np.testing.assert_allclose(u, g["NK"][t], rtol=0, atol=1e-15)
        fin = np.isfinite(g["NK_inf"][t])
        assert np.array_equal(np.isfinite(l), fin)
        np.testing.assert_allclose(l[fin], g["NK_inf"][t][fin], rtol=0, atol=1e-15)
    m.close()


@pytest.mark.parametrize("name", ["trot", "pace", "walk"])
def test_c_run_follows_the_numpy_port(name):
    """The fixtures store the raw eps-1e-8 iterate (`x_admm`) and iteration count the numpy/scipy port produced when
    the reference's MPC.run drove it tick after tick (warm starts, Ax updates, adaptive rho): the C restatement is
    the same algorithm, so it must take the same number of iterations and land on the same iterate."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "solve_%s.npz" % name))
    m = c_port.MPC(n_steps=16)
    T = min(8, len(g["k"]))
    for t in range(T):
        r = m.run(g["k"][t] == 0, g["xref"][t], g["fsteps"][t])
        assert r["status"] == 1
        # same iteration count (one termination-check interval of slack for a different libm / CPU); the iterates
        # differ only because the fixture's MPC.run warm-started from the polished x, this run from its own raw x
        assert abs(r["iter"] - int(g["osqp_iter"][t])) <= 25, (t, r["iter"], g["osqp_iter"][t])
        np.testing.assert_allclose(r["x"], g["x_admm"][t], rtol=0, atol=5e-6)
        np.testing.assert_array_equal(r["f"], r["x"][192:204])                    # MPC.py:440
        assert np.abs(r["x"][192:] - g["x"][t][192:]).max() < 1e-2                   # un-polished: not 1e-4 N sharp
    m.close()


def test_c_run_polished_is_certified_and_matches_golden():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "solve_trot_turn.npz"))
    N = 16
    m = c_port.MPC(n_steps=N)
    for t in range(4):
        r = m.run(g["k"][t] == 0, g["xref"][t], g["fsteps"][t])
        Pd, A, l, u, _ = mpc_build.build_qp(g["xref"][t], g["fsteps"][t], first_tick=(g["k"][t] == 0))
        P, q = sp.diags(Pd).tocsc(), np.zeros(24 * N)
        xp, yp, low, upp = kkt.polish(P, q, A, l, u, r["x"], r["y"])
        cert = kkt.certificate(P, q, A, l, u, xp, yp)
        assert kkt.is_certified(cert), cert
        assert np.abs(xp - g["x"][t]).max() < 1e-9
    m.close()


def test_c_known_answers():
    # min 1/2 |x|^2 - x0 - 3 x1  s.t. 0 <= x <= 2  ->  x = (1, 2), multiplier of the upper bound on x1 = 1
    x, y, it = c_port.solve_qp([1.0, 1.0], [-1.0, -3.0], sp.identity(2, format="csc"), [0.0, 0.0], [2.0, 2.0], eps=1e-9)
    assert it > 0
    np.testing.assert_allclose(x, [1.0, 2.0], atol=1e-6)
    np.testing.assert_allclose(y, [0.0, 1.0], atol=1e-5)
    # min 1/2 |x|^2  s.t. x0 + x1 = 1 -> x = (.5, .5), y = -.5
    x, y, it = c_port.solve_qp([1.0, 1.0], [0.0, 0.0], sp.csc_matrix(np.array([[1.0, 1.0]])), [1.0], [1.0], eps=1e-9)
    np.testing.assert_allclose(x, [0.5, 0.5], atol=1e-7)
    np.testing.assert_allclose(y, [-0.5], atol=1e-6)


def test_c_sparse_ldl_on_a_random_quasidefinite_qp():
    """Exercises the ordering + LDL' on a pattern unlike the MPC's: random sparse A, random bounds; the result must agree
    with the numpy/scipy port (which uses scipy's sparse LU for the same KKT systems)."""
    from oracle.osqp_port import OSQP
    rng = np.random.default_rng(5)
    n, m = 30, 45
    A = sp.random(m, n, density=0.15, random_state=7, format="csc") + sp.vstack([sp.identity(n), sp.csc_matrix((m - n, n))])
    A = sp.csc_matrix(A)
    Pd = rng.uniform(0.5, 2.0, n)
    q = rng.normal(size=n)
    l = -rng.uniform(0.1, 1.0, m)
    u = rng.uniform(0.1, 1.0, m)
    l[:5] = u[:5]                                      # some equality rows
    x, y, it = c_port.solve_qp(Pd, q, A, l, u, eps=1e-9)
    s = OSQP()
    s.setup(P=sp.diags(Pd).tocsc(), q=q, A=A, l=l, u=u, eps_abs=1e-9, eps_rel=1e-9)
    r = s.solve()
    assert it == r.info.iter
    np.testing.assert_allclose(x, r.x, atol=1e-9)
    np.testing.assert_allclose(y, r.y, atol=1e-8)


def test_c_replay_threads_agree_with_single_runs():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "solve_trot.npz"))
    T = 4
    xr = np.stack([g["xref"][:T], g["xref"][:T]])
    fs = np.stack([g["fsteps"][:T], g["fsteps"][:T]])
    sec, f, iters = c_port.replay_mt(xr, fs, warm=1)
    assert sec > 0 and iters > 0
    np.testing.assert_array_equal(f[0], f[1])
    np.testing.assert_allclose(f[0], g["x_admm"][:T, 192:204], rtol=0, atol=1e-8)
