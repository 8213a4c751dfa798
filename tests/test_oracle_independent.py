"""The oracle's certified optima against an independent third-party QP solver.

The reference solves its QP with `osqp`, which is not installable here, so the solve half of the oracle is
pinned by a KKT certificate (oracle/kkt.py) rather than by a comparison.  This file adds the comparison that *is* available:
HiGHS (its active-set QP solver, bundled inside scipy) solves the QP the reference built -- matrices from oracle.mpc_build, which
test_oracle_build.py pins to the reference's own ML.data / NK -- and must land on the golden optimum: same objective to the
accuracy HiGHS reaches (a few 1e-6 relative, always from above: no feasible point below the golden objective), forces as close as
the flat force cost
(w_f = 1e-5: |df| ~ sqrt(2 * dobj / w_f)) allows.  CPU only; nothing here touches the product."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import mpc_build

hc = pytest.importorskip("scipy.optimize._highspy._core")
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(p for p in glob.glob(os.path.join(HERE, "golden", "solve_*.npz")) if "_N" not in os.path.basename(p))


def solve_highs(Pd, A, l, u, tol=1e-10):
    """min 1/2 x' diag(Pd) x  s.t.  l <= A x <= u  with HiGHS; returns (x, model status)."""
    n, m = len(Pd), A.shape[0]
    A = sp.csc_matrix(A)
    model = hc.HighsModel()
    lp = model.lp_
    lp.num_col_, lp.num_row_ = n, m
    lp.col_cost_ = np.zeros(n)
    lp.col_lower_, lp.col_upper_ = np.full(n, -hc.kHighsInf), np.full(n, hc.kHighsInf)
    lp.row_lower_ = np.where(np.isfinite(l), l, -hc.kHighsInf)
    lp.row_upper_ = np.where(np.isfinite(u), u, hc.kHighsInf)
    lp.a_matrix_.format_ = hc.MatrixFormat.kColwise
    lp.a_matrix_.start_ = A.indptr.astype(np.int32)
    lp.a_matrix_.index_ = A.indices.astype(np.int32)
    lp.a_matrix_.value_ = A.data.astype(np.float64)
    hs = model.hessian_
    hs.dim_, hs.format_ = n, hc.HessianFormat.kTriangular
    hs.start_, hs.index_, hs.value_ = np.arange(n + 1, dtype=np.int32), np.arange(n, dtype=np.int32), np.asarray(Pd, dtype=np.float64)
    h = hc._Highs()
    h.setOptionValue("output_flag", False)
    h.setOptionValue("primal_feasibility_tolerance", tol)
    h.setOptionValue("dual_feasibility_tolerance", tol)
    h.passModel(model)
    h.run()
    return np.array(h.getSolution().col_value), h.getModelStatus()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[6:-4] for p in GOLD])
def test_highs_lands_on_the_golden_optimum(path):
    g = np.load(path)
    n = g["x"].shape[1] // 24
    par = mpc_build.Params(n_steps=n)
    for t in range(min(3, len(g["k"]))):
        Pd, A, l, u, _ = mpc_build.build_qp(g["xref"][t], g["fsteps"][t], par, first_tick=(g["k"][t] == 0))
        x, status = solve_highs(Pd, A, l, u)
        assert status == hc.HighsModelStatus.kOptimal
        Ax = A @ x
        assert (Ax >= l - 1e-7).all() and (Ax <= u + 1e-7).all()
        obj, gold = 0.5 * float((Pd * x * x).sum()), float(g["obj"][t])
        scale = max(abs(gold), 1e-12)
        assert abs(obj - gold) <= 5e-6 * scale + 1e-12, (t, obj, gold)        # HiGHS' own accuracy: 6e-7 typical, 2.4e-6 worst (walk)
        assert obj >= gold - 1e-8 * scale - 1e-13, (t, obj, gold)           # the certified optimum is not beaten
        df = np.abs(x[12 * n:] - g["x"][t][12 * n:]).max()
        assert df <= 3.0 * np.sqrt(2.0 * max(obj - gold, 0.0) / par.w_force) + 1e-3, (t, df, obj - gold)
        assert np.abs(x[:12 * n] - g["x"][t][:12 * n]).max() <= 5e-3            # weakest state weight 0.0166: same flatness argument
