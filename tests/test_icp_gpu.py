"""GPU parity tests of the ICP path: CUDA (through the C ABI) vs the CPU oracle and
vs the reference's golden vectors.  Index / integer results must be bit-exact;
floating-point sums are compared at 1e-5 relative to the natural scale of each
slot (sum of |terms|), which is the tolerance BASELINE.json's north_star states
("JtJ ... within 1e-5 relative fp32").
"""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from tests.synth import make_colors, make_icp_pair

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star tolerance for JtJ / Jtr


@pytest.fixture(scope="module")
def o3d():
    import open3d_b200
    assert torch.cuda.is_available()
    return open3d_b200


def _lib():
    from open3d_b200 import _lib
    return _lib


def _stream():
    return int(torch.cuda.current_stream().cuda_stream)


def _search(points, queries, radius, k=1):
    L = _lib()
    p = torch.from_numpy(np.ascontiguousarray(points, np.float32)).cuda()
    q = torch.from_numpy(np.ascontiguousarray(queries, np.float32)).cuda()
    h = C.c_void_p()
    L.check(L.lib.o3db_nns_create(p.data_ptr(), p.shape[0], float(radius), _stream(), C.byref(h)))
    n = q.shape[0]
    idx = torch.full((n, k), -7, dtype=torch.int32, device="cuda")
    d2 = torch.full((n, k), -7.0, dtype=torch.float32, device="cuda")
    cnt = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    L.check(L.lib.o3db_nns_hybrid_search(h, q.data_ptr(), n, float(radius), k, idx.data_ptr(), d2.data_ptr(),
                                         cnt.data_ptr(), _stream()))
    torch.cuda.synchronize()
    L.lib.o3db_nns_destroy(h)
    return idx.cpu().numpy(), d2.cpu().numpy(), cnt.cpu().numpy()


# ------------------------------------------------------------------ search

def test_hybrid_search_reference_kat(kats):
    k = kats["hybrid_search"]  # cpp/tests/core/NearestNeighborSearch.cpp:321-353
    idx, d2, cnt = _search(k["dataset_points"], k["query_points"], k["radius"], k["max_knn"])
    assert idx.tolist() == k["gt_indices"]
    np.testing.assert_allclose(d2, np.array(k["gt_distances"], np.float32), rtol=1e-5, atol=1e-8)
    assert cnt.tolist() == k["gt_counts"]


@pytest.mark.parametrize("n,r,k", [(20000, 0.05, 1), (20000, 0.03, 1), (5000, 0.08, 4)])
def test_hybrid_search_bit_exact_vs_oracle(n, r, k):
    src, tgt, _, _ = make_icp_pair(n, seed=3)
    src = src.copy()
    src[:50] = tgt[:50]          # coincident points (dist 0)
    tgt = tgt.copy()
    tgt[200:210] = tgt[190:200]  # duplicated targets: exact ties -> lower index
    src[-10:] += 50.0            # far outside the target's bounding box
    gi, gd, gc = _search(tgt, src, r, k)
    oi, od, oc = oracle.hybrid_search(tgt, src, r, k)
    assert np.array_equal(gi, oi)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))   # bit-exact f32 distances
    assert np.array_equal(gc, oc)
    assert (gc[-10:] == 0).all() and (gi[:50, 0] >= 0).all()


def test_hybrid_search_random_volume_and_edge_cases():
    rng = np.random.default_rng(11)
    pts = rng.uniform(-1, 1, (4000, 3)).astype(np.float32)
    q = rng.uniform(-1.2, 1.2, (3000, 3)).astype(np.float32)
    for r in (0.11, 0.26):
        gi, gd, gc = _search(pts, q, r, 1)
        oi, od, oc = oracle.hybrid_search(pts, q, r, 1, bruteforce=True)
        assert np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32)) and np.array_equal(gc, oc)
    # single target point, query exactly at distance r along x (<= is inclusive upstream)
    gi, gd, gc = _search(np.zeros((1, 3), np.float32), np.array([[0.5, 0, 0], [0.5000001, 0, 0]], np.float32), 0.5, 1)
    oi, od, oc = oracle.hybrid_search(np.zeros((1, 3), np.float32), np.array([[0.5, 0, 0], [0.5000001, 0, 0]], np.float32), 0.5, 1)
    assert np.array_equal(gi, oi) and np.array_equal(gc, oc)
    # empty query set is a no-op; bad arguments are errors
    L = _lib()
    p = torch.zeros((4, 3), device="cuda")
    h = C.c_void_p()
    L.check(L.lib.o3db_nns_create(p.data_ptr(), 4, 0.1, _stream(), C.byref(h)))
    assert L.lib.o3db_nns_hybrid_search(h, None, 0, 0.1, 1, None, None, None, _stream()) == 0
    assert L.lib.o3db_nns_hybrid_search(h, p.data_ptr(), 4, 0.2, 1, None, None, None, _stream()) == L.ERR_INVALID
    assert L.lib.o3db_nns_hybrid_search(h, p.data_ptr(), 4, 0.1, 0, None, None, None, _stream()) == L.ERR_INVALID
    L.lib.o3db_nns_destroy(h)


def test_reference_layout_spatial_hash_table():
    """BuildSpatialHashTableCUDA layout: bucket = SpatialHash(floor(p/(2r))) % H, bit-exact."""
    L = _lib()
    src, tgt, _, _ = make_icp_pair(30000, seed=4)
    tgt = tgt - 3.0   # negative coordinates exercise the sign-extended hash
    r = 0.05
    H = max(len(tgt) // 32, 1)   # FixedRadiusIndex.h:516-517
    p = torch.from_numpy(tgt).cuda()
    index = torch.empty(len(tgt), dtype=torch.int32, device="cuda")
    splits = torch.empty(H + 1, dtype=torch.int32, device="cuda")
    L.check(L.lib.o3db_build_spatial_hash_table(p.data_ptr(), len(tgt), r, H, index.data_ptr(), splits.data_ptr(), _stream()))
    torch.cuda.synchronize()
    inv = np.float32(1) / (np.float32(2) * np.float32(r))
    cells = np.floor(tgt * inv).astype(np.int32)
    buckets = (oracle.spatial_hash(cells) % np.uint64(H)).astype(np.int64)
    want = np.concatenate([[0], np.cumsum(np.bincount(buckets, minlength=H))])
    assert np.array_equal(splits.cpu().numpy().astype(np.int64), want)
    idx = index.cpu().numpy().astype(np.int64)
    assert np.array_equal(np.sort(idx), np.arange(len(tgt)))
    got_bucket = np.searchsorted(want, np.arange(len(tgt)), side="right") - 1
    assert np.array_equal(buckets[idx], got_bucket)   # every id sits in its own bucket's segment


# ------------------------------------------------------------ pose kernels

def _pose(src, tgt, nrm, corr, robust=(0, 1.0, 1.0), want_host=True):
    L = _lib()
    s = torch.from_numpy(np.ascontiguousarray(src, np.float32)).cuda()
    t = torch.from_numpy(np.ascontiguousarray(tgt, np.float32)).cuda()
    n = torch.from_numpy(np.ascontiguousarray(nrm, np.float32)).cuda()
    c = torch.from_numpy(np.ascontiguousarray(corr, np.int64)).cuda()
    sums = torch.zeros(29, dtype=torch.float64, device="cuda")
    pose = torch.zeros(6, dtype=torch.float64, device="cuda")
    k = L.RobustKernel(int(robust[0]), float(robust[1]), float(robust[2]))
    res, cnt = C.c_float(0), C.c_int(0)
    rc = L.lib.o3db_compute_pose_point_to_plane(s.data_ptr(), t.data_ptr(), n.data_ptr(), c.data_ptr(), s.shape[0],
                                                C.byref(k), sums.data_ptr(), pose.data_ptr(),
                                                C.byref(res) if want_host else None,
                                                C.byref(cnt) if want_host else None, _stream())
    torch.cuda.synchronize()
    return rc, sums.cpu().numpy(), pose.cpu().numpy(), res.value, cnt.value


def test_compute_transformation_reference_kat(kats, o3d):
    """TransformationEstimation.cpp:148, 176 through the public python surface."""
    k = kats["transformation_estimation"]
    reg = o3d.t.pipelines.registration
    src = o3d.t.geometry.PointCloud(np.array(k["source_points"], np.float32))
    tgt = o3d.t.geometry.PointCloud(np.array(k["target_points"], np.float32))
    tgt.set_point_normals(np.array(k["target_normals"], np.float32))
    corr = np.array(k["correspondences"], np.int64)
    est = reg.TransformationEstimationPointToPlane()
    e = k["expected"]["p2plane_rmse"]
    assert abs(est.compute_rmse(src, tgt, corr) - e["value"]) < e["tol"]
    T = est.compute_transformation(src, tgt, corr)
    moved = src.clone().transform(T)
    e = k["expected"]["p2plane_rmse_after"]
    assert abs(est.compute_rmse(moved, tgt, corr) - e["value"]) < e["tol"]
    # and the 29 sums / pose agree with the oracle on the same fixture
    _, sums, pose, res, cnt = _pose(k["source_points"], k["target_points"], k["target_normals"], corr)
    o = oracle.pose_p2plane_sums(k["source_points"], k["target_points"], k["target_normals"], corr)
    np.testing.assert_allclose(sums, o["sums64"], rtol=0, atol=RTOL * o["abs64"].max())
    op, ores, ocnt, _ = oracle.decode_and_solve_6x6(o["sums64"])
    np.testing.assert_allclose(pose, op, rtol=1e-4, atol=1e-6)
    assert cnt == ocnt == 14 and abs(res - ores) < 1e-5


@pytest.mark.parametrize("robust", [(0, 1.0, 1.0), (1, 1.0, 1.0), (2, 0.01, 1.0), (3, 0.02, 1.0), (4, 0.5, 1.0),
                                    (5, 0.03, 1.0), (6, 0.05, 1.0), (6, 0.05, 0.0), (6, 0.05, 2.0), (6, 0.05, -2.0)])
def test_pose_sums_vs_oracle(robust):
    src, tgt, nrm, _ = make_icp_pair(40000, seed=5)
    idx, _, _ = oracle.hybrid_search(tgt, src, 0.05, 1)
    corr = idx[:, 0].astype(np.int64)
    assert (corr == -1).any() and (corr >= 0).sum() > 1000
    rc, sums, pose, res, cnt = _pose(src, tgt, nrm, corr, robust)
    assert rc == 0
    names = ["L2Loss", "L1Loss", "HuberLoss", "CauchyLoss", "GMLoss", "TukeyLoss", "GeneralizedLoss"]
    o = oracle.pose_p2plane_sums(src, tgt, nrm, corr, (names[robust[0]], robust[1], robust[2]))
    err = np.abs(sums - o["sums64"])
    assert (err <= RTOL * o["abs64"] + 1e-300).all(), (err / o["abs64"]).max()
    assert cnt == int(o["sums64"][28])
    op, _, _, sing = oracle.decode_and_solve_6x6(o["sums64"])
    assert not sing
    np.testing.assert_allclose(pose, op, rtol=1e-3, atol=1e-7)


def test_pose_edge_cases():
    L = _lib()
    src, tgt, nrm, _ = make_icp_pair(1000, seed=6)
    none = np.full(len(src), -1, np.int64)
    rc, sums, pose, res, cnt = _pose(src, tgt, nrm, none)
    # all-zero system is singular: upstream logs the error (raises) and zero-fills
    assert rc == L.ERR_SINGULAR and "Singular" in L.last_error()
    assert (sums == 0).all() and (pose == 0).all() and cnt == 0
    rc, sums, pose, _, _ = _pose(src, tgt, nrm, none, want_host=False)
    assert rc == 0 and (pose == 0).all()
    # n = 0
    assert L.lib.o3db_compute_pose_point_to_plane(None, None, None, None, 0, None, None, None, None, None, _stream()) == 0


def test_colored_pose_vs_oracle():
    L = _lib()
    src, tgt, nrm, _ = make_icp_pair(20000, seed=8)
    sc, tc = make_colors(src, 1), make_colors(tgt, 1)
    rng = np.random.default_rng(0)
    grad = rng.normal(0, 0.5, tgt.shape).astype(np.float32)
    idx, _, _ = oracle.hybrid_search(tgt, src, 0.05, 1)
    corr = idx[:, 0].astype(np.int64)
    dev = [torch.from_numpy(a).cuda() for a in (src, sc, tgt, nrm, tc, grad)]
    c = torch.from_numpy(corr).cuda()
    sums = torch.zeros(29, dtype=torch.float64, device="cuda")
    pose = torch.zeros(6, dtype=torch.float64, device="cuda")
    res, cnt = C.c_float(0), C.c_int(0)
    k = L.RobustKernel(0, 1.0, 1.0)
    L.check(L.lib.o3db_compute_pose_colored_icp(*[d.data_ptr() for d in dev], c.data_ptr(), len(src), C.byref(k),
                                                0.968, sums.data_ptr(), pose.data_ptr(), C.byref(res), C.byref(cnt),
                                                _stream()))
    o = oracle.pose_colored_sums(src, sc, tgt, nrm, tc, grad, corr, 0.968)
    err = np.abs(sums.cpu().numpy() - o["sums64"])
    assert (err <= RTOL * o["abs64"] + 1e-300).all(), (err / o["abs64"]).max()
    assert cnt.value == int(o["sums64"][28])


def test_transform_points_and_normals(o3d):
    rng = np.random.default_rng(2)
    pts = rng.uniform(-5, 5, (10001, 3)).astype(np.float32)
    T = np.eye(4)
    from tests.synth import axis_angle
    T[:3, :3] = axis_angle([0.3, -1, 0.5], 25.0)
    T[:3, 3] = [0.4, -2.0, 1.5]
    pc = o3d.t.geometry.PointCloud(pts.copy()).set_point_normals(pts.copy())
    pc.transform(T)
    torch.cuda.synchronize()
    np.testing.assert_allclose(pc.point["positions"].cpu().numpy(), oracle.transform_points(T, pts), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(pc.point["normals"].cpu().numpy(), oracle.transform_normals(T, pts), rtol=2e-6, atol=2e-6)


# ---------------------------------------------------------------- ICP loop

def _icp(o3d, src, tgt, nrm, r, init=None, crit=None, kernel=None, cb=None):
    reg = o3d.t.pipelines.registration
    s = o3d.t.geometry.PointCloud(src)
    t = o3d.t.geometry.PointCloud(tgt).set_point_normals(nrm)
    est = reg.TransformationEstimationPointToPlane(kernel)
    return reg.icp(s, t, r, np.eye(4) if init is None else init, est, crit or reg.ICPConvergenceCriteria(), -1.0, cb)


@pytest.mark.parametrize("n,iters", [(100000, 1), (100000, 12), (30000, 30)])
def test_icp_loop_vs_oracle(o3d, n, iters):
    """BASELINE config 1 (100k, 1 iteration) and longer runs: trajectory-level parity."""
    reg = o3d.t.pipelines.registration
    src, tgt, nrm, T_gt = make_icp_pair(n, seed=1)
    log = []
    res = _icp(o3d, src, tgt, nrm, 0.05, crit=reg.ICPConvergenceCriteria(0, 0, iters), cb=log.append)
    ref = oracle.icp_p2plane(src, tgt, nrm, 0.05, max_iteration=iters, relative_fitness=0, relative_rmse=0)
    assert res.num_iterations == ref.num_iterations == iters and res.converged == ref.converged
    assert len(log) == iters and [c["iteration_index"] for c in log] == list(range(iters))
    per = np.array([[c["fitness"], c["inlier_rmse"]] for c in log])
    # iteration 0 sees bit-identical inputs: fitness must match exactly, rmse to summation order
    assert per[0, 0] == ref.per_iteration[0, 0]
    # (each warp sums its 32 squared distances as an f32 tree before the f64 accumulation: <= 5 * 2^-24 relative on
    # the sum — the reference itself sums an f32 tensor, Registration.cpp:47-50)
    assert abs(per[0, 1] - ref.per_iteration[0, 1]) < 3e-7 * ref.per_iteration[0, 1]
    # later iterations differ only through the f32 accumulation order of the update
    np.testing.assert_allclose(per[:, 0], ref.per_iteration[:, 0], atol=2e-4)
    np.testing.assert_allclose(per[:, 1], ref.per_iteration[:, 1], atol=2e-6)
    np.testing.assert_allclose(res.transformation, ref.transformation, atol=2e-5)
    assert abs(res.fitness - ref.fitness) < 2e-4 and abs(res.inlier_rmse - ref.inlier_rmse) < 2e-6
    corr = res.correspondence_set.cpu().numpy()
    assert corr.shape == (n_src := len(src),) and corr.dtype == np.int64
    agree = (corr == ref.correspondences).mean()
    assert agree > 0.999, agree
    if iters >= 12:
        np.testing.assert_allclose(res.transformation, T_gt, atol=2e-3)


def test_icp_one_iteration_update_is_tight(o3d):
    """With identical inputs the single update must agree to f32-sum accuracy."""
    reg = o3d.t.pipelines.registration
    src, tgt, nrm, _ = make_icp_pair(50000, seed=9)
    res = _icp(o3d, src, tgt, nrm, 0.05, crit=reg.ICPConvergenceCriteria(0, 0, 1))
    ref = oracle.icp_p2plane(src, tgt, nrm, 0.05, max_iteration=1, relative_fitness=0, relative_rmse=0)
    np.testing.assert_allclose(res.transformation, ref.transformation, atol=1e-7)


def test_icp_with_init_and_convergence(o3d):
    reg = o3d.t.pipelines.registration
    src, tgt, nrm, T_gt = make_icp_pair(40000, seed=10)
    init = T_gt.copy()
    init[:3, 3] += [0.004, -0.003, 0.002]
    crit = reg.ICPConvergenceCriteria(1e-6, 1e-6, 50)
    res = _icp(o3d, src, tgt, nrm, 0.05, init=init, crit=crit)
    ref = oracle.icp_p2plane(src, tgt, nrm, 0.05, init=init, max_iteration=50)
    assert res.converged and ref.converged
    assert abs(res.num_iterations - ref.num_iterations) <= 1 and res.num_iterations < 50
    np.testing.assert_allclose(res.transformation, ref.transformation, atol=5e-5)


def test_icp_robust_kernel(o3d):
    reg = o3d.t.pipelines.registration
    src, tgt, nrm, _ = make_icp_pair(30000, seed=12)
    k = reg.robust_kernel.RobustKernel(reg.robust_kernel.RobustKernelMethod.TukeyLoss, 0.03, 1.0)
    res = _icp(o3d, src, tgt, nrm, 0.05, crit=reg.ICPConvergenceCriteria(0, 0, 8), kernel=k)
    ref = oracle.icp_p2plane(src, tgt, nrm, 0.05, max_iteration=8, relative_fitness=0, relative_rmse=0,
                             robust=("TukeyLoss", 0.03, 1.0))
    np.testing.assert_allclose(res.transformation, ref.transformation, atol=5e-5)


def test_icp_no_correspondences(o3d):
    src, tgt, nrm, _ = make_icp_pair(5000, seed=13)
    res = _icp(o3d, src + 100.0, tgt, nrm, 0.05)
    # Registration.cpp:51-60, 300-306, 434-438
    assert res.fitness == 0 and res.inlier_rmse == 0 and not res.converged and res.num_iterations == 0
    np.testing.assert_array_equal(res.transformation, np.eye(4))
    assert (res.correspondence_set.cpu().numpy() == -1).all()


def test_icp_singular_system_raises(o3d):
    # all target points and normals identical: JtJ has rank 1 -> singular, upstream raises
    tgt = np.zeros((64, 3), np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (64, 1))
    src = np.zeros((64, 3), np.float32)
    with pytest.raises(RuntimeError, match="Singular"):
        _icp(o3d, src, tgt, nrm, 0.05)


def test_icp_host_buffer_entry_point():
    L = _lib()
    src, tgt, nrm, _ = make_icp_pair(20000, seed=14)
    opt = L.IcpOptions()
    opt.max_correspondence_distance, opt.max_iteration = 0.05, 5
    opt.relative_fitness = opt.relative_rmse = 0.0
    opt.kernel = L.RobustKernel(0, 1.0, 1.0)
    res = L.IcpResult()
    corr = np.empty(len(src), np.int64)
    per = np.zeros((5, 2))
    T0 = np.eye(4)
    L.check(L.lib.o3db_icp_point_to_plane_host(src.ctypes.data, len(src), tgt.ctypes.data, nrm.ctypes.data, len(tgt),
                                               L.dptr(T0), C.byref(opt), C.byref(res), corr.ctypes.data, L.dptr(per)))
    ref = oracle.icp_p2plane(src, tgt, nrm, 0.05, max_iteration=5, relative_fitness=0, relative_rmse=0)
    np.testing.assert_allclose(np.array(res.transformation).reshape(4, 4), ref.transformation, atol=2e-5)
    assert res.num_iterations == 5 and (corr == ref.correspondences).mean() > 0.999


def test_icp_reproducibility_and_properties_at_scale(o3d):
    """Full-size property checks (2M points, BASELINE config 2 shape): two runs agree bit for
    bit, and fitness * N is an integer equal to the number of valid correspondences."""
    reg = o3d.t.pipelines.registration
    src, tgt, nrm, _ = make_icp_pair(2_000_000, seed=2)
    crit = reg.ICPConvergenceCriteria(0, 0, 3)
    a = _icp(o3d, src, tgt, nrm, 0.05, crit=crit)
    b = _icp(o3d, src, tgt, nrm, 0.05, crit=crit)
    # (the cell-sort scatter order is not deterministic, so f64 partial sums may differ in the
    # last bits between runs; anything visible beyond that would be a bug)
    assert abs(a.fitness - b.fitness) < 1e-6
    assert (a.correspondence_set == b.correspondence_set).float().mean().item() > 0.99999
    np.testing.assert_allclose(a.transformation, b.transformation, atol=1e-10)
    corr = a.correspondence_set
    valid = int((corr >= 0).sum())
    assert abs(a.fitness * len(src) - valid) < 1e-6 and valid > 0.9 * len(src)
    # every reported correspondence really is within the radius of the moved source point
    s = torch.from_numpy(src).cuda().double()
    T = torch.from_numpy(a.transformation).cuda()
    moved = s @ T[:3, :3].T + T[:3, 3]
    t = torch.from_numpy(tgt).cuda().double()
    m = corr >= 0
    d = (moved[m] - t[corr[m]]).norm(dim=1)
    assert float(d.max()) <= 0.05 * (1 + 1e-4)
    assert abs(float((d ** 2).mean().sqrt()) - a.inlier_rmse) < 1e-5


# ------------------------------------------------- VoxelDownSample

def _by_voxel(pos, vs, *attrs):
    k = np.floor(pos / np.float32(vs)).astype(np.int64)
    order = np.lexsort((k[:, 2], k[:, 1], k[:, 0]))
    return (k[order], pos[order]) + tuple(a[order] for a in attrs)


@pytest.mark.parametrize("n,vs", [(50000, 0.05), (200000, 0.031), (1000, 10.0)])
def test_voxel_down_sample_vs_oracle(o3d, n, vs):
    src, tgt, nrm, _ = make_icp_pair(n, seed=15)
    tgt = tgt - 2.5                                   # negative coordinates too
    col = make_colors(tgt, 2)
    pc = o3d.t.geometry.PointCloud(tgt).set_point_normals(nrm).set_point_colors(col)
    down = pc.voxel_down_sample(vs)
    ref = oracle.voxel_down_sample(tgt, vs, normals=nrm, colors=col)
    gp, gn, gc = (down.point[k].cpu().numpy() for k in ("positions", "normals", "colors"))
    assert len(gp) == len(ref["positions"])           # same set of occupied voxels ...
    kg, gp, gn, gc = _by_voxel(gp, vs, gn, gc)
    assert np.array_equal(kg, ref["keys"].astype(np.int64))   # ... bit-exact voxel keys
    # means: f32 atomics (as the reference's IndexAdd_) vs the oracle's f64 accumulation
    np.testing.assert_allclose(gp, ref["positions"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(gn, ref["normals"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(gc, ref["colors"], rtol=1e-5, atol=2e-6)
    with pytest.raises(RuntimeError, match="voxel_size must be positive"):
        pc.voxel_down_sample(0.0)


def test_multi_scale_icp_with_voxel_pyramid_vs_oracle(o3d):
    """MultiScaleICP with down-sampling (Registration.cpp:221-273, 362-444), BASELINE config-2 style
    (voxel_size 0.02 on the last scale) against the same pipeline assembled from oracle pieces."""
    reg = o3d.t.pipelines.registration
    src, tgt, nrm, T_gt = make_icp_pair(90000, seed=16)
    voxels, radii, iters = [0.06, 0.02], [0.12, 0.05], [6, 8]
    s = o3d.t.geometry.PointCloud(src)
    t = o3d.t.geometry.PointCloud(tgt).set_point_normals(nrm)
    res = reg.multi_scale_icp(s, t, voxels, [reg.ICPConvergenceCriteria(0, 0, k) for k in iters], radii, np.eye(4),
                              reg.TransformationEstimationPointToPlane())
    # oracle pyramid: finest first, coarser levels from the finer ones
    s1, t1 = oracle.voxel_down_sample(src, voxels[1]), oracle.voxel_down_sample(tgt, voxels[1], normals=nrm)
    s0, t0 = oracle.voxel_down_sample(s1["positions"], voxels[0]), oracle.voxel_down_sample(t1["positions"], voxels[0], normals=t1["normals"])
    T = np.eye(4)
    for (ss, tt), r, k in zip(((s0, t0), (s1, t1)), radii, iters):
        ref = oracle.icp_p2plane(ss["positions"], tt["positions"], tt["normals"], r, init=T, max_iteration=k,
                                 relative_fitness=0, relative_rmse=0)
        T = ref.transformation
    assert res.num_iterations == sum(iters)
    np.testing.assert_allclose(res.transformation, T, atol=5e-5)
    assert abs(res.fitness - ref.fitness) < 1e-3 and abs(res.inlier_rmse - ref.inlier_rmse) < 1e-5
    np.testing.assert_allclose(res.transformation, T_gt, atol=3e-3)


# ------------------------------------------------- ColoredICP

def _colored_pair(n, seed):
    src, tgt, nrm, T_gt = make_icp_pair(n, seed=seed)
    # the texture lives in the target frame: a source point carries the colour of where it belongs
    sc = make_colors(oracle.transform_points(T_gt, src), 1)
    return src, sc, tgt, nrm, make_colors(tgt, 1), T_gt


@pytest.mark.parametrize("n,radius,max_nn", [(20000, 0.08, 30), (20000, 0.03, 30), (5000, 0.025, 8)])
def test_color_gradients_vs_oracle(o3d, n, radius, max_nn):
    """EstimateColorGradients, hybrid search (PointCloudImpl.h:1066-1165).  Default solver = the reference's
    solve_svd3x3<float>: BIT-EXACT vs the oracle, which is itself bit-exact vs the reference's own
    EstimatePointWiseColorGradientKernel compiled in oracle/_ref (tests/test_oracle_vs_ref.py) — identical neighbour
    lists, identical f32 normal equations, identical 4-sweep SVD.  solver="exact" (extension): 1e-5 of the oracle's
    exact pseudo-inverse."""
    _, _, tgt, nrm, tc, _ = _colored_pair(n, 21)
    pc = o3d.t.geometry.PointCloud(tgt).set_point_normals(nrm).set_point_colors(tc)
    pc.estimate_color_gradients(max_nn, radius)
    g = pc.point["color_gradients"].cpu().numpy()
    ref = oracle.estimate_color_gradients(tgt, nrm, tc, radius, max_nn)
    assert g.shape == ref.shape and np.isfinite(g).all()
    zero = ~ref.any(axis=1)
    assert np.array_equal(zero, ~g.any(axis=1))           # < 4 neighbours -> exactly zero, same points
    if radius < 0.03:
        assert zero.any() and not zero.all()
    assert g.tobytes() == ref.tobytes(), float(np.abs(g - ref).max())
    pc.estimate_color_gradients(max_nn, radius, solver="exact")
    g = pc.point["color_gradients"].cpu().numpy()
    ref = oracle.estimate_color_gradients(tgt, nrm, tc, radius, max_nn, solver="exact")
    np.testing.assert_allclose(g, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())
    # the exact gradient is tangential: the orthogonality row drives g . n to ~0
    assert np.abs((g * nrm).sum(1)).max() < 1e-3 * max(np.abs(g).max(), 1.0)
    with pytest.raises(RuntimeError, match="colors"):
        o3d.t.geometry.PointCloud(tgt).set_point_normals(nrm).estimate_color_gradients(30, radius)
    with pytest.raises(RuntimeError, match="normals"):
        o3d.t.geometry.PointCloud(tgt).set_point_colors(tc).estimate_color_gradients(30, radius)


def _colored_clouds(o3d, src, sc, tgt, nrm, tc, grad=None):
    s = o3d.t.geometry.PointCloud(src).set_point_colors(sc)
    t = o3d.t.geometry.PointCloud(tgt).set_point_normals(nrm).set_point_colors(tc)
    if grad is not None:
        t.point["color_gradients"] = torch.from_numpy(np.ascontiguousarray(grad, np.float32)).cuda()
    return s, t


@pytest.mark.parametrize("n,iters,robust", [(40000, 1, None), (40000, 12, None), (30000, 8, ("TukeyLoss", 0.05, 1.0))])
def test_colored_icp_loop_vs_oracle(o3d, n, iters, robust):
    """Fused ColoredICP loop vs the oracle's loop on identical inputs (gradients given to both)."""
    reg = o3d.t.pipelines.registration
    src, sc, tgt, nrm, tc, T_gt = _colored_pair(n, 22)
    grad = oracle.estimate_color_gradients(tgt, nrm, tc, 0.08, 30)
    s, t = _colored_clouds(o3d, src, sc, tgt, nrm, tc, grad)
    kernel = None
    if robust:
        kernel = reg.RobustKernel(reg.RobustKernelMethod[robust[0]], robust[1], robust[2])
    est = reg.TransformationEstimationForColoredICP(0.968, kernel)
    log = []
    res = reg.icp(s, t, 0.05, np.eye(4), est, reg.ICPConvergenceCriteria(0, 0, iters), -1.0, log.append)
    ref = oracle.icp_colored(src, sc, tgt, nrm, tc, grad, 0.05, max_iteration=iters, relative_fitness=0,
                             relative_rmse=0, lambda_geometric=0.968, robust=robust or ("L2Loss", 1.0, 1.0))
    assert ref.status == 0 and res.num_iterations == ref.num_iterations == iters
    per = np.array([[c["fitness"], c["inlier_rmse"]] for c in log])
    # (rmse: each warp sums its 32 terms as an f32 tree before the f64 accumulation, see test_icp_loop_vs_oracle)
    assert per[0, 0] == ref.per_iteration[0, 0] and abs(per[0, 1] - ref.per_iteration[0, 1]) < 3e-7 * ref.per_iteration[0, 1]
    np.testing.assert_allclose(per[:, 0], ref.per_iteration[:, 0], atol=2e-4)
    np.testing.assert_allclose(per[:, 1], ref.per_iteration[:, 1], atol=2e-6)
    np.testing.assert_allclose(res.transformation, ref.transformation, atol=1e-7 if iters == 1 else 2e-5)
    agree = (res.correspondence_set.cpu().numpy() == ref.correspondences).mean()
    assert agree > 0.999, agree
    if iters >= 12:
        np.testing.assert_allclose(res.transformation, T_gt, atol=2e-3)
    assert "color_gradients" in t.point and "color_gradients" not in s.point


def test_colored_icp_differs_from_point_to_plane_and_uses_colour(o3d):
    """On a PLANE the geometric term cannot see an in-plane shift; only the photometric term can."""
    reg = o3d.t.pipelines.registration
    rng = np.random.default_rng(5)
    tgt = np.zeros((40000, 3), np.float32)
    tgt[:, :2] = rng.uniform(0, 4, (40000, 2))
    nrm = np.tile(np.float32([0, 0, 1]), (len(tgt), 1))
    shift = np.array([0.012, -0.008, 0.0])
    base = np.zeros((30000, 3))
    base[:, :2] = rng.uniform(0.2, 3.8, (30000, 2))
    src = (base - shift).astype(np.float32)                 # T_gt = translation by +shift
    sc, tc = make_colors(base, 1), make_colors(tgt, 1)
    s, t = _colored_clouds(o3d, src, sc, tgt, nrm, tc)
    crit = reg.ICPConvergenceCriteria(0, 0, 25)
    res_c = reg.icp(s, t, 0.05, np.eye(4), reg.TransformationEstimationForColoredICP(0.5), crit)
    assert "color_gradients" not in t.point                 # computed on a clone (Registration.cpp:235, 257)
    np.testing.assert_allclose(res_c.transformation[:3, 3], shift, atol=1.5e-3)
    grad = oracle.estimate_color_gradients(tgt, nrm, tc, 0.1, 30)   # radius = 2 * max_correspondence_distance
    ref = oracle.icp_colored(src, sc, tgt, nrm, tc, grad, 0.05, max_iteration=25, relative_fitness=0,
                             relative_rmse=0, lambda_geometric=0.5)
    np.testing.assert_allclose(res_c.transformation, ref.transformation, atol=5e-5)


def test_colored_multi_scale_icp_with_pyramid_vs_oracle(o3d):
    """MultiScaleICP + ColoredICP: gradients are estimated on the finest level (radius 4 * voxel) and
    averaged down the pyramid like any other attribute (Registration.cpp:243-270)."""
    reg = o3d.t.pipelines.registration
    src, sc, tgt, nrm, tc, T_gt = _colored_pair(90000, 23)
    voxels, radii, iters = [0.06, 0.03], [0.12, 0.06], [5, 8]
    s, t = _colored_clouds(o3d, src, sc, tgt, nrm, tc)
    res = reg.multi_scale_icp(s, t, voxels, [reg.ICPConvergenceCriteria(0, 0, k) for k in iters], radii, np.eye(4),
                              reg.TransformationEstimationForColoredICP())
    s1 = oracle.voxel_down_sample(src, voxels[1], colors=sc)
    t1 = oracle.voxel_down_sample(tgt, voxels[1], normals=nrm, colors=tc)
    g1 = oracle.estimate_color_gradients(t1["positions"], t1["normals"], t1["colors"], 4.0 * voxels[1], 30)
    s0 = oracle.voxel_down_sample(s1["positions"], voxels[0], colors=s1["colors"])
    t0 = oracle.voxel_down_sample(t1["positions"], voxels[0], normals=t1["normals"], colors=t1["colors"])
    g0 = oracle.voxel_down_sample(t1["positions"], voxels[0], colors=g1)["colors"]
    T = np.eye(4)
    for (ss, tt, gg), r, k in zip(((s0, t0, g0), (s1, t1, g1)), radii, iters):
        ref = oracle.icp_colored(ss["positions"], ss["colors"], tt["positions"], tt["normals"], tt["colors"], gg, r,
                                 init=T, max_iteration=k, relative_fitness=0, relative_rmse=0)
        T = ref.transformation
    assert res.num_iterations == sum(iters)
    # the down-sampled clouds agree to f32 rounding only (atomics vs f64 means), which the ill-conditioned
    # gradient solve amplifies: trajectory-level tolerance
    np.testing.assert_allclose(res.transformation, T, atol=3e-4)
    assert abs(res.fitness - ref.fitness) < 2e-3 and abs(res.inlier_rmse - ref.inlier_rmse) < 2e-5
    np.testing.assert_allclose(res.transformation, T_gt, atol=4e-3)


def test_colored_icp_argument_errors(o3d):
    reg = o3d.t.pipelines.registration
    src, sc, tgt, nrm, tc, _ = _colored_pair(2000, 24)
    est = reg.TransformationEstimationForColoredICP()
    G = o3d.t.geometry.PointCloud
    with pytest.raises(RuntimeError, match="source pointcloud to have colors"):
        reg.icp(G(src), G(tgt).set_point_normals(nrm).set_point_colors(tc), 0.05, np.eye(4), est)
    with pytest.raises(RuntimeError, match="target pointcloud to have colors"):
        reg.icp(G(src).set_point_colors(sc), G(tgt).set_point_normals(nrm), 0.05, np.eye(4), est)
    with pytest.raises(RuntimeError, match="target pointcloud to have normals"):
        reg.icp(G(src).set_point_colors(sc), G(tgt).set_point_colors(tc), 0.05, np.eye(4), est)
    assert reg.TransformationEstimationForColoredICP(1.5).lambda_geometric == 0.968
    # estimator-level seam: compute_transformation == one loop iteration's update
    s, t = _colored_clouds(o3d, src, sc, tgt, nrm, tc, oracle.estimate_color_gradients(tgt, nrm, tc, 0.1, 30))
    one = reg.icp(s, t, 0.05, np.eye(4), est, reg.ICPConvergenceCriteria(0, 0, 1))
    idx, _, _ = oracle.hybrid_search(tgt, src, 0.05, 1)
    T = est.compute_transformation(s, t, torch.from_numpy(idx[:, 0].astype(np.int64)))
    # (the stand-alone pose kernel and the fused loop add the same f32 terms in different association orders)
    np.testing.assert_allclose(T, one.transformation, atol=1e-7)
    assert est.compute_rmse(s, t, torch.from_numpy(idx[:, 0].astype(np.int64))) > 0


def test_get_information_matrix_vs_oracle(o3d):
    """registration.get_information_matrix (Registration.cpp:446-485): transform, hybrid search k = 1, GTG of the matched
    target points.  Same correspondences as the oracle (bit-exact search) => the 21 sums agree to f32-term rounding."""
    reg = o3d.t.pipelines.registration
    src, tgt, nrm, T_gt = make_icp_pair(60000, seed=12)
    s, t = o3d.t.geometry.PointCloud(src), o3d.t.geometry.PointCloud(tgt)
    for T, r in ((T_gt, 0.03), (np.eye(4), 0.05)):
        info = reg.get_information_matrix(s, t, r, T)
        want = oracle.get_information_matrix(src, tgt, r, T)
        assert info.shape == (6, 6) and info.dtype == np.float64 and np.array_equal(info, info.T)
        assert info[3, 3] == want[3, 3] > 1000              # number of correspondences
        np.testing.assert_allclose(info, want, rtol=1e-6, atol=1e-6 * np.abs(want).max())
    far = np.eye(4)
    far[:3, 3] = 100.0
    with pytest.raises(RuntimeError, match="0 correspondence"):
        reg.get_information_matrix(s, t, 0.05, far)
