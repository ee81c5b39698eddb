"""CPU checks of the oracle's ColoredICP pieces (no GPU): the reference ships no golden vectors for
EstimateColorGradients / ColoredICP (its test, cpp/tests/t/pipelines/registration/Registration.cpp:246-289,
needs downloaded data and only compares against the legacy pipeline at 0.05 / 0.02), so the oracle is
pinned here through analytic properties; the per-correspondence Jacobian is pinned bit-exactly against the
reference's own header in test_oracle_vs_ref.py."""
import numpy as np

import oracle
from tests.synth import make_colors, make_icp_pair


def _intensity_and_gradient(p):
    """mean of make_colors(.., seed=1) channels on the plane z = const, and its analytic x/y gradient"""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    s = 1
    I = (0.5 + 0.5 * np.sin(2.1 * x + 0.3 * s) * np.cos(1.7 * y) + 0.5 + 0.5 * np.sin(1.3 * x + 2.2 * y + 1.0) +
         0.5 + 0.5 * np.cos(0.9 * x - 1.9 * y + 4.0 * z)) / 3
    gx = (0.5 * 2.1 * np.cos(2.1 * x + 0.3 * s) * np.cos(1.7 * y) + 0.5 * 1.3 * np.cos(1.3 * x + 2.2 * y + 1.0) -
          0.5 * 0.9 * np.sin(0.9 * x - 1.9 * y + 4.0 * z)) / 3
    gy = (-0.5 * 1.7 * np.sin(2.1 * x + 0.3 * s) * np.sin(1.7 * y) + 0.5 * 2.2 * np.cos(1.3 * x + 2.2 * y + 1.0) +
          0.5 * 1.9 * np.sin(0.9 * x - 1.9 * y + 4.0 * z)) / 3
    return I, gx, gy


def test_color_gradient_matches_analytic_gradient_on_a_plane():
    rng = np.random.default_rng(3)
    p = np.zeros((20000, 3))
    p[:, :2] = rng.uniform(0, 2, (20000, 2))
    nrm = np.tile([0.0, 0.0, 1.0], (len(p), 1))
    col = make_colors(p, 1)
    I, gx, gy = _intensity_and_gradient(p)
    np.testing.assert_allclose(col.mean(1), I, atol=1e-6)
    g = oracle.estimate_color_gradients(p, nrm, col, 0.05, 30)
    inner = (p[:, 0] > 0.1) & (p[:, 0] < 1.9) & (p[:, 1] > 0.1) & (p[:, 1] < 1.9)
    assert np.abs(g[:, 2]).max() < 1e-6                        # tangential (orthogonality row)
    err = np.hypot(g[inner, 0] - gx[inner], g[inner, 1] - gy[inner])
    assert np.median(err) < 0.01 and err.max() < 0.08, (np.median(err), err.max())   # |grad| ~ 0.5


def test_color_gradient_zero_with_fewer_than_four_neighbours():
    p = np.array([[0, 0, 0], [0.01, 0, 0], [0, 0.01, 0], [5, 5, 5], [5.01, 5, 5], [5, 5.01, 5], [5.01, 5.01, 5]], float)
    nrm = np.tile([0.0, 0.0, 1.0], (len(p), 1))
    col = np.tile(p[:, :1] * 10, (1, 3))                       # intensity = 10 x
    g = oracle.estimate_color_gradients(p, nrm, col, 0.05, 30)
    assert not g[:3].any()                                     # 3 neighbours (self included) -> zero
    np.testing.assert_allclose(g[3:], np.tile([10.0, 0, 0], (4, 1)), atol=1e-3)


def test_sym3x3_pinv_drops_null_directions():
    n = np.array([0.0, 0.0, 1.0])
    A = np.diag([2.0, 3.0, 0.0])
    x = oracle.solve_sym3x3_pinv(A, [4.0, 9.0, 5.0])           # the z component is in the null space
    np.testing.assert_allclose(x, [2.0, 3.0, 0.0], atol=1e-12)
    R = np.linalg.qr(np.random.default_rng(0).normal(size=(3, 3)))[0]
    A2 = R @ np.diag([5.0, 1e-3, 0.0]) @ R.T
    b = R @ np.array([1.0, 2.0, 3.0])
    np.testing.assert_allclose(oracle.solve_sym3x3_pinv(A2, b), np.linalg.pinv(A2, rcond=1e-9) @ b, rtol=1e-8)


def test_colored_icp_recovers_motion_and_in_plane_shift():
    src, tgt, nrm, T_gt = make_icp_pair(20000, seed=22)
    sc, tc = make_colors(oracle.transform_points(T_gt, src), 1), make_colors(tgt, 1)
    grad = oracle.estimate_color_gradients(tgt, nrm, tc, 0.08, 30)
    res = oracle.icp_colored(src, sc, tgt, nrm, tc, grad, 0.05, max_iteration=15, relative_fitness=0, relative_rmse=0)
    assert res.status == 0 and res.num_iterations == 15
    np.testing.assert_allclose(res.transformation, T_gt, atol=2e-3)
    # a plane: only the photometric term observes the in-plane shift
    rng = np.random.default_rng(5)
    tp = np.zeros((20000, 3), np.float32)
    tp[:, :2] = rng.uniform(0, 3, (20000, 2))
    tn = np.tile(np.float32([0, 0, 1]), (len(tp), 1))
    shift = np.array([0.012, -0.008, 0.0])
    base = np.zeros((15000, 3))
    base[:, :2] = rng.uniform(0.2, 2.8, (15000, 2))
    sp = (base - shift).astype(np.float32)
    g = oracle.estimate_color_gradients(tp, tn, make_colors(tp, 1), 0.1, 30)
    res = oracle.icp_colored(sp, make_colors(base, 1), tp, tn, make_colors(tp, 1), g, 0.05, max_iteration=25,
                             relative_fitness=0, relative_rmse=0, lambda_geometric=0.5)
    np.testing.assert_allclose(res.transformation[:3, 3], shift, atol=1e-3)
    p2l = oracle.icp_p2plane(sp, tp, tn, 0.05, max_iteration=3, relative_fitness=0, relative_rmse=0)
    assert p2l.status != 0 or np.abs(p2l.transformation[:2, 3] - shift[:2]).max() > 5e-3   # geometry alone cannot
