"""CPU tests of the oracle itself (the checker must be right before it checks anything).

The reference's tests (tests/TestBasic.cpp) hold no evaluation-level vectors; what they do hold --
the parameter sets, the test's own restatement of the friction rows (:122-123) and of the
superquadric value (:192-194), the equilibrium identities (:127-132) -- is asserted here on the
oracle, together with closed-form points, finite differences and a 50-digit evaluation."""
import numpy as np
import pytest

from centroidalplanner_b200 import synthetic
from oracle import cpl_oracle_py as orc

NAMES4 = synthetic.NAMES4


def _testbasic_ground():
    o = orc.Oracle(NAMES4, orc.ENV_GROUND, 100.0)
    o.set_ground_z(0.1)
    o.set_mu(0.5)
    o.set_com_weight(2.0)
    o.set_force_weight(0.0)
    o.set_wrench(synthetic.TESTBASIC["wrench"])
    for nm in NAMES4:
        o.set_var_bounds(orc.BLOCK_P, nm, synthetic.TESTBASIC["p_lb"], synthetic.TESTBASIC["p_ub"])
    return o


@pytest.mark.parametrize("nc", [1, 2, 4, 8, 13])
@pytest.mark.parametrize("env", [orc.ENV_NONE, orc.ENV_GROUND, orc.ENV_SUPERQUADRIC])
def test_dimensions_match_survey_formulas(nc, env):
    o = orc.Oracle(["c%02d" % i for i in range(nc)], env)
    assert o.n == 3 + 9 * nc
    if env == orc.ENV_NONE:
        assert (o.m, o.nnz) == (6 + 2 * nc, 6 + 27 * nc)
    else:
        assert (o.m, o.nnz) == (6 + 6 * nc, 6 + 42 * nc)
    iRow, jCol = o.structure()
    # IpoptAdapter order: row-major, column ascending within a row, no duplicates
    key = iRow.astype(np.int64) * o.n + jCol
    assert (np.diff(key) > 0).all()


def test_structure_rows_checked_in_the_survey():
    o = orc.Oracle(NAMES4, orc.ENV_GROUND)
    iRow, jCol = o.structure()
    assert list(jCol[iRow == 0]) == [3, 12, 21, 30]
    assert list(jCol[iRow == 6]) == [6, 7, 8]
    assert list(jCol[iRow == 7]) == [6, 7, 8, 9]
    assert list(jCol[iRow == 10]) == [3, 4, 5, 9, 10, 11]
    o2 = orc.Oracle(["r_foot", "l_foot", "r_hand", "l_hand"], orc.ENV_GROUND)
    assert list(o2.sorted_order()) == [1, 3, 0, 2]  # l_foot, l_hand, r_foot, r_hand
    r2, c2 = o2.structure()
    assert list(c2[r2 == 6]) == [15, 16, 17]  # row 6 belongs to l_foot = vector index 1


def test_sorted_order_is_bytewise_like_std_string():
    o = orc.Oracle(["b", "B", "a10", "a2", "a"], orc.ENV_GROUND)
    assert [o.names[k] for k in o.sorted_order()] == ["B", "a", "a10", "a2", "b"]


def test_bounds():
    o = _testbasic_ground()
    lb, ub = o.var_bounds()
    assert lb[0] == -1000.0 and ub[2] == 1000.0
    for k in range(4):
        assert list(lb[3 + 9 * k + 3:3 + 9 * k + 6]) == [-0.3, -0.3, 0.0]
        assert list(ub[3 + 9 * k + 3:3 + 9 * k + 6]) == [0.3, 0.3, 1.0]
    clb, cub = o.con_bounds()
    assert (clb[:6] == 0).all() and (cub == 0).all()
    for j in range(4):
        assert list(clb[6 + 6 * j:12 + 6 * j]) == [0, 0, 0, 0, -1e20, -1e20]


def equilibrium_x(z0=0.1, m=100.0):
    """4 feet at (+-a, +-b, z0), CoM above the centroid, F_i = (0,0,m*9.81/4), n = (0,0,1)."""
    x = np.zeros(39)
    x[0:3] = [0.0, 0.0, 1.0]
    feet = [(0.3, 0.2), (0.3, -0.2), (-0.3, 0.2), (-0.3, -0.2)]
    for k, (a, b) in enumerate(feet):
        x[3 + 9 * k:3 + 9 * k + 3] = [0.0, 0.0, m * 9.81 / 4]
        x[3 + 9 * k + 3:3 + 9 * k + 6] = [a, b, z0]
        x[3 + 9 * k + 6:3 + 9 * k + 9] = [0.0, 0.0, 1.0]
    return x


def test_closed_form_equilibrium_point():
    o = orc.Oracle(NAMES4, orc.ENV_GROUND, 100.0)
    o.set_ground_z(0.1)
    o.set_mu(0.5)
    g = o.eval(equilibrium_x())["g"]
    assert np.abs(g[:6]).max() <= 1e-12           # SigmaF + m g = 0 (TestBasic.cpp:127-129), Sigma tau = 0 (:130-132)
    for j in range(4):
        r = 6 + 6 * j
        assert g[r] == 0.0 and (g[r + 1:r + 4] == 0.0).all()    # on the plane, normal matches
        assert g[r + 4] == -245.25 and g[r + 5] == -0.5 * 245.25  # -F.n + 0 ; 0 - mu F.n


def test_default_start_is_nan_in_friction_rows_only():
    """x = 0 is the reference's start (Variable3D.cpp:8-10): 0/0 in FrictionCone.cpp:85-87,97-99."""
    o = orc.Oracle(NAMES4, orc.ENV_GROUND, 100.0)
    e = o.eval(np.zeros(39))
    iRow, _ = o.structure()
    nan_rows = set(iRow[np.isnan(e["jac"])])
    assert nan_rows == {6 + 6 * j + 5 for j in range(4)}
    assert np.isfinite(e["g"]).all()


def test_friction_rows_equal_the_reference_tests_restatement():
    """TestBasic.cpp:122-123 recomputes both friction rows from the solution; same expressions here."""
    o = _testbasic_ground()
    X = synthetic.ground_batch(64)
    for x in X:
        g = o.eval(x)["g"]
        for j, k in enumerate(o.sorted_order()):
            F = x[3 + 9 * k:3 + 9 * k + 3]
            n = x[3 + 9 * k + 6:3 + 9 * k + 9]
            r0 = -F.dot(n)
            r1 = np.linalg.norm(F - n.dot(F) * n) - 0.5 * F.dot(n)
            assert abs(g[6 + 6 * j + 4] - r0) <= 1e-12 * max(1, abs(r0))
            assert abs(g[6 + 6 * j + 5] - r1) <= 1e-12 * max(1, abs(r1))


def test_superquadric_value_matches_the_reference_tests_restatement():
    """TestBasic.cpp:192-194: pow((p-C)/R, P) summed == 1 on the surface; same expression, any p."""
    sq = synthetic.SUPERQUADRIC
    o = orc.Oracle(NAMES4, orc.ENV_SUPERQUADRIC)
    o.set_superquadric(sq["C"], sq["R"], sq["P"])
    rng = np.random.default_rng(3)
    for _ in range(50):
        p = np.array(sq["C"]) + rng.uniform(-1, 1, 3) * np.array([0.35, 0.35, 0.5])
        want = sum(((p[i] - sq["C"][i]) / sq["R"][i]) ** sq["P"][i] for i in range(3)) - 1.0
        assert abs(o.env_value(p) - want) <= 1e-13 * max(1.0, abs(want))


@pytest.mark.parametrize("env", [orc.ENV_NONE, orc.ENV_GROUND, orc.ENV_SUPERQUADRIC])
@pytest.mark.parametrize("names", [NAMES4, ["r_foot", "l_foot", "r_hand", "l_hand"], synthetic.NAMES8])
def test_jacobian_and_gradient_against_central_differences(env, names):
    """The reference's de-facto derivative check is IPOPT's derivative_test (CentroidalPlanner.cpp:26)."""
    o = orc.Oracle(names, env, 80.0)
    sq = synthetic.SUPERQUADRIC
    o.set_superquadric(sq["C"], sq["R"], sq["P"])
    o.set_ground_z(0.1)
    o.set_mu(0.7)
    o.set_wrench([10, -5, 3, 1, 2, -4])
    for k, nm in enumerate(names):
        o.set_force_threshold(nm, 5.0 + k)
        o.set_pos_ref(nm, [0.1, 0.2, 0.3 * k])
        o.set_contact_force_weight(nm, 0.01)
    nc = len(names)
    gen = synthetic.superquadric_batch if env == orc.ENV_SUPERQUADRIC else synthetic.ground_batch
    x = gen(3, nc, 21)[2]
    e = o.eval(x)
    iRow, jCol = o.structure()
    J = np.zeros((o.m, o.n))
    J[iRow, jCol] = e["jac"]
    Jfd = np.zeros_like(J)
    gfd = np.zeros(o.n)
    for i in range(o.n):
        h = 1e-6 * max(1.0, abs(x[i]))
        xp, xm = x.copy(), x.copy()
        xp[i] += h
        xm[i] -= h
        ep, em = o.eval(xp, want=("g", "cost")), o.eval(xm, want=("g", "cost"))
        Jfd[:, i] = (ep["g"] - em["g"]) / (2 * h)
        gfd[i] = (ep["cost"] - em["cost"]) / (2 * h)
    assert (np.abs(J - Jfd) <= 2e-6 * np.maximum(1.0, np.abs(J))).all()
    mask = np.zeros_like(J, dtype=bool)
    mask[iRow, jCol] = True
    assert np.abs(Jfd[~mask]).max() <= 1e-6     # nothing outside the declared sparsity
    assert (np.abs(gfd - e["grad"]) <= 1e-5 * np.maximum(1.0, np.abs(e["grad"]))).all()


def test_superquadric_normal_jacobian_against_50_digit_evaluation():
    """d(grad f/|grad f|)/dp evaluated with mpmath at 50 digits; away from the expanded-square region
     the fp64 transcription must agree to ~1e-13."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    sq = synthetic.SUPERQUADRIC
    C, R, P = [[mp.mpf(str(v)) for v in sq[k]] for k in ("C", "R", "P")]
    o = orc.Oracle(NAMES4, orc.ENV_SUPERQUADRIC)
    o.set_superquadric(sq["C"], sq["R"], sq["P"])

    def unit_grad(px, py, pz, i):
        p = [px, py, pz]
        g = [P[q] / R[q] ** P[q] * (p[q] - C[q]) ** (P[q] - 1) for q in range(3)]
        return g[i] / mp.sqrt(g[0] ** 2 + g[1] ** 2 + g[2] ** 2)

    for p in ([0.25, -0.2, 1.3], [-0.1, 0.3, 0.6], [0.2, 0.15, 1.45]):
        J = o.env_normal_jacobian(p)
        pm = [mp.mpf(repr(v)) for v in p]
        for i in range(3):
            for j in range(3):
                want = mp.diff(lambda *a: unit_grad(*a, i), tuple(pm), tuple(int(q == j) for q in range(3)))
                assert abs(J[i, j] - float(want)) <= 2e-13 * max(abs(float(want)), 1e-300), (p, i, j, J[i, j], float(want))
        # GetNormalValue = -grad/|grad| (Superquadric.cpp:66-68)
        nrm = o.env_normal(p)
        for i in range(3):
            assert abs(nrm[i] + float(unit_grad(*pm, i))) <= 1e-15


def test_reduction_order_flag_changes_at_most_one_ulp():
    o = _testbasic_ground()
    X = synthetic.ground_batch(256)
    a = o.eval_batch(X)
    o.set_reduction_order(1)
    b = o.eval_batch(X)
    assert not np.array_equal(a["g"], b["g"])  # the flag does something ...
    # ... at the level of one rounding of the reduced quantity: |F.n| < 512 here, ulp(512) = 1.1e-13
    assert np.abs(a["g"] - b["g"]).max() <= 2 * 1.14e-13


def test_skipping_untouched_pairs_changes_nothing():
    for env in (orc.ENV_GROUND, orc.ENV_SUPERQUADRIC):
        o = orc.Oracle(NAMES4, env)
        X = (synthetic.superquadric_batch if env == orc.ENV_SUPERQUADRIC else synthetic.ground_batch)(64)
        a = o.eval_batch(X)
        o.set_call_all_pairs(0)
        b = o.eval_batch(X, nthreads=3)
        for k in ("g", "jac", "cost", "grad"):
            assert np.array_equal(a[k], b[k], equal_nan=True)


def test_gravity_term_is_exact_for_testbasic_mass():
    assert 100.0 * -9.81 == -981.0
