"""Pin the CPU oracle against the reference's own inline known-answer tests.

The fixtures in tests/golden/reference_kats.json are parsed verbatim from
Open3D's C++ tests by tests/golden/make_reference_kats.py (file:line recorded
in every entry).  These tests run on CPU (no GPU marker).
"""
import numpy as np
import pytest

import oracle


def test_robust_kernel_weights(kats):
    k = kats["robust_kernel"]  # cpp/tests/t/pipelines/registration/Registration.cpp:411-491
    for f32 in (False, True):
        for method, expected in k["expected_by_method"].items():
            w = oracle.robust_weight(int(method), k["scaling_parameter"], k["shape_parameter"],
                                     k["residual"], f32=f32)
            assert abs(w - expected) < k["tol"], (method, w, expected)
        for g in k["generalized_by_shape"]:
            w = oracle.robust_weight(6, k["scaling_parameter"], g["shape"], g["residual"], f32=f32)
            assert abs(w - g["expected"]) < g["tol"], (g, w)


@pytest.mark.parametrize("bruteforce", [False, True])
def test_hybrid_search_kat(kats, bruteforce):
    k = kats["hybrid_search"]  # cpp/tests/core/NearestNeighborSearch.cpp:321-353
    idx, d2, cnt = oracle.hybrid_search(k["dataset_points"], k["query_points"], k["radius"],
                                        k["max_knn"], bruteforce=bruteforce)
    assert idx.tolist() == k["gt_indices"]
    # core::Tensor::AllClose defaults: rtol 1e-5, atol 1e-8
    np.testing.assert_allclose(d2, np.array(k["gt_distances"], np.float32), rtol=1e-5, atol=1e-8)
    assert cnt.tolist() == k["gt_counts"]


def _fixture(kats, dtype=np.float32):
    k = kats["transformation_estimation"]
    return (np.array(k["source_points"], dtype), np.array(k["target_points"], dtype),
            np.array(k["target_normals"], dtype), np.array(k["correspondences"], np.int64),
            k["expected"])


def test_p2plane_rmse_kat(kats):
    src, tgt, nrm, corr, exp = _fixture(kats)
    e = exp["p2plane_rmse"]  # TransformationEstimation.cpp:148
    assert abs(oracle.rmse_p2plane(src, tgt, nrm, corr) - e["value"]) < e["tol"]


@pytest.mark.parametrize("acc", ["sums64", "sums32"])
def test_p2plane_compute_transformation_kat(kats, acc):
    """GetJacobianPointToPlane + 29-slot reduction + DecodeAndSolve6x6 +
    PoseToTransformation + TransformPoints, end to end (TransformationEstimation.cpp:176)."""
    src, tgt, nrm, corr, exp = _fixture(kats)
    s = oracle.pose_p2plane_sums(src, tgt, nrm, corr)
    pose, residual, count, singular = oracle.decode_and_solve_6x6(s[acc])
    assert not singular and count == 14
    T = oracle.pose_to_transformation(pose)
    moved = oracle.transform_points(T, src)
    e = exp["p2plane_rmse_after"]
    assert abs(oracle.rmse_p2plane(moved, tgt, nrm, corr) - e["value"]) < e["tol"]


def test_p2plane_f64_kat(kats):
    src, tgt, nrm, corr, exp = _fixture(kats, np.float64)
    s = oracle.pose_p2plane_sums_f64(src, tgt, nrm, corr)
    pose, _, count, singular = oracle.decode_and_solve_6x6(s)
    assert not singular and count == 14
    T = oracle.pose_to_transformation(pose)
    moved = (src @ T[:3, :3].T + T[:3, 3])
    e = exp["p2plane_rmse_after"]
    err = ((moved - tgt[corr]) * nrm[corr]) ** 2
    assert abs(np.sqrt(err.sum() / len(corr)) - e["value"]) < e["tol"]


def test_sums_layout_against_numpy(kats):
    """29-slot packing (RegistrationCPU.cpp:66-76) re-derived independently in numpy."""
    src, tgt, nrm, corr, _ = _fixture(kats, np.float64)
    t, n = tgt[corr], nrm[corr]
    r = ((src - t) * n).sum(1)
    J = np.concatenate([np.cross(src, n), n], axis=1)
    JtJ = J.T @ J
    Jtr = J.T @ r
    want = [JtJ[j, k] for j in range(6) for k in range(j + 1)] + list(Jtr) + [r.sum(), len(r)]
    got = oracle.pose_p2plane_sums_f64(src, tgt, nrm, corr)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    got32 = oracle.pose_p2plane_sums(src, tgt, nrm, corr)
    np.testing.assert_allclose(got32["sums64"], want, rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(got32["sums32"], want, rtol=2e-5, atol=2e-5)


def test_pose_to_transformation_is_rzryrx():
    pose = np.array([0.1, -0.2, 0.3, 1.0, 2.0, 3.0])
    a, b, g = pose[:3]
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(g), -np.sin(g), 0], [np.sin(g), np.cos(g), 0], [0, 0, 1]])
    T = oracle.pose_to_transformation(pose)
    np.testing.assert_allclose(T[:3, :3], Rz @ Ry @ Rx, atol=1e-15)
    np.testing.assert_allclose(T[:3, 3], pose[3:])
    np.testing.assert_allclose(T[3], [0, 0, 0, 1])


def test_solve_matches_numpy_and_flags_singular():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((40, 6))
    b = rng.standard_normal(40)
    AtA, Atb = A.T @ A, A.T @ b
    s = np.zeros(29)
    s[:21] = [AtA[j, k] for j in range(6) for k in range(j + 1)]
    s[21:27] = Atb
    s[27], s[28] = 1.5, 40
    pose, res, cnt, singular = oracle.decode_and_solve_6x6(s)
    np.testing.assert_allclose(pose, np.linalg.solve(AtA, -Atb), rtol=1e-10)
    assert (res, cnt, singular) == (1.5, 40, False)
    _, _, _, singular = oracle.decode_and_solve_6x6(np.zeros(29))
    assert singular


def test_vbg_indexing_set_semantics(kats):
    k = kats["vbg_indexing"]  # cpp/tests/t/geometry/VoxelBlockGrid.cpp:199-219
    table = np.zeros((10, 3), np.int32)
    bi, masks, size, rc = oracle.hashmap_activate(table, 0, k["keys"])
    assert rc == 0 and size == k["expected_unique"] == int(masks.sum())
    keys = np.array(k["keys"], np.int32)
    assert (table[bi] == keys).all()
    # second activation of the same keys inserts nothing (HashMap.cpp Activate semantics)
    bi2, masks2, size2, _ = oracle.hashmap_activate(table, size, k["keys"])
    assert size2 == size and not masks2.any() and (bi2 == bi).all()


def test_hash_functions_known_values():
    # FNV-1a-style MiniVecHash (core/hashmap/Dispatch.h:67-81), re-derived in python ints
    def fnv(k):
        h = 14695981039346656037
        for e in k:
            h ^= e & 0xFFFFFFFFFFFFFFFF
            h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h
    keys = [[0, 0, 0], [-1, 3, 2], [1, 2, 3], [2147483647, -2147483648, -7]]
    assert oracle.minivec_hash(keys).tolist() == [fnv(k) for k in keys]

    # SpatialHash (core/nns/NeighborSearchCommon.h:31-37): int arithmetic, sign-extended
    def sh(x, y, z):
        def i32(v):
            v &= 0xFFFFFFFF
            return v - (1 << 32) if v & 0x80000000 else v
        h = i32(i32(x * 73856096) ^ i32(y * 193649663) ^ i32(z * 83492791))
        return h & 0xFFFFFFFFFFFFFFFF
    cells = [[0, 0, 0], [1, 2, 3], [-5, 17, -300], [100000, -100000, 54321]]
    assert oracle.spatial_hash(cells).tolist() == [sh(*c) for c in cells]
    assert oracle.compute_voxel_index([-0.05, 0.149, 0.1], 10.0).tolist() == [-1, 1, 1]


def test_grid_search_equals_bruteforce_random():
    rng = np.random.default_rng(5)
    pts = rng.uniform(-1, 1, (3000, 3)).astype(np.float32)
    q = rng.uniform(-1.1, 1.1, (500, 3)).astype(np.float32)
    q[:20] = pts[:20]  # coincident points
    pts[100:110] = pts[90:100]  # duplicates -> exact ties, lower index must win
    for r, k in [(0.12, 1), (0.2, 3), (0.05, 1)]:
        a = oracle.hybrid_search(pts, q, r, k)
        b = oracle.hybrid_search(pts, q, r, k, bruteforce=True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_icp_loop_recovers_known_motion():
    from tests.synth import make_icp_pair
    src, tgt, nrm, T_gt = make_icp_pair(20000, seed=1)
    res = oracle.icp_p2plane(src, tgt, nrm, 0.05, max_iteration=20, relative_fitness=0, relative_rmse=0)
    assert res.status == 0 and res.num_iterations == 20 and not res.converged
    assert res.per_iteration.shape == (20, 2)
    assert res.fitness > 0.95
    np.testing.assert_allclose(res.transformation, T_gt, atol=2e-3)
    # f32 accumulation (reference behaviour) lands within the reference's own test tolerance
    res32 = oracle.icp_p2plane(src, tgt, nrm, 0.05, max_iteration=20, relative_fitness=0,
                               relative_rmse=0, accumulate_f64=False)
    assert abs(res32.fitness - res.fitness) < 0.005 and abs(res32.inlier_rmse - res.inlier_rmse) < 0.005


def test_icp_no_correspondences():
    from tests.synth import make_icp_pair
    src, tgt, nrm, _ = make_icp_pair(2000, seed=2)
    res = oracle.icp_p2plane(src + 100.0, tgt, nrm, 0.05, max_iteration=5)
    # Registration.cpp:51-60, 300-306: identity, fitness 0, not converged
    assert res.fitness == 0 and res.inlier_rmse == 0 and not res.converged
    assert res.num_iterations == 0 and (res.correspondences == -1).all()
    np.testing.assert_array_equal(res.transformation, np.eye(4))
