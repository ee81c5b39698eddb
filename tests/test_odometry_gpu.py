"""GPU parity tests of RGB-D odometry (PointToPlane) and the dense-SLAM loop it completes: CUDA (through the C ABI /
the reference-facing python surface) vs the CPU oracle, which is itself pinned bit-exactly to the reference's
ImageImpl.h / RGBDOdometryJacobianImpl.h (tests/test_oracle_vs_ref.py)."""
import numpy as np
import pytest
import torch

import oracle
from tests.synth import PRIMESENSE_K, camera_pose, render_depth

pytestmark = pytest.mark.gpu

VOXEL, RES, TRUNC = 0.008, 16, 8.0
SCALE, DMIN, DMAX = 1000.0, 0.1, 3.0
NAN = float("nan")


@pytest.fixture(scope="module")
def o3d():
    import open3d_b200
    assert torch.cuda.is_available()
    return open3d_b200


def _pair(i, j, holes=True):
    Ta, Tb = camera_pose(i), camera_pose(j)
    da, db = render_depth(Ta).numpy(), render_depth(Tb).numpy()
    if holes:
        da[40:60, 100:140] = 0
        db[300:330, 200:260] = 0
    return da, db, np.linalg.inv(Ta) @ Tb      # target depth, source depth, T_source_to_target


def _same(a, b):
    a, b = np.ascontiguousarray(a).reshape(-1), np.ascontiguousarray(b).reshape(-1)
    nan = np.isnan(a)
    return a.shape == b.shape and np.array_equal(nan, np.isnan(b)) and \
        np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32))


def test_image_pyramid_kernels_bit_exact_vs_oracle(o3d):
    Image = o3d.t.geometry.Image
    da, _, _ = _pair(100, 103)
    for src in (da, (da.astype(np.float32) / 1000.0)):
        scale = 1000.0 if src.dtype == np.uint16 else 1.0
        got = Image(torch.from_numpy(src).cuda()).clip_transform(scale, 0.0, 3.0, NAN).as_tensor().cpu().numpy()
        want = oracle.clip_transform(src, scale, 0.0, 3.0, NAN)
        assert got.shape == (480, 640, 1) and np.isnan(want).any() and _same(got, want)
    depth = Image(torch.from_numpy(da).cuda()).clip_transform(1000.0, 0.0, 3.0, NAN)
    odepth = oracle.clip_transform(da)
    K = np.array(PRIMESENSE_K, np.float64)
    for level in range(3):
        v = depth.create_vertex_map(K, NAN)
        ov = oracle.create_vertex_map(odepth, K)
        assert _same(v.as_tensor().cpu().numpy(), ov)
        assert _same(v.create_normal_map(NAN).as_tensor().cpu().numpy(), oracle.create_normal_map(ov))
        sm = depth.filter_bilateral(5, 5.0, 10.0).as_tensor().cpu().numpy()[..., 0]
        osm = oracle.filter_bilateral(odepth, 5, 5.0, 10.0)
        assert np.array_equal(np.isnan(sm), np.isnan(osm))
        np.testing.assert_allclose(sm[~np.isnan(sm)], osm[~np.isnan(osm)], rtol=2e-6)     # expf: libm vs CUDA
        depth = depth.pyr_down_depth(0.14, NAN)
        odepth = oracle.pyr_down_depth(odepth, 0.14)
        assert depth.rows == odepth.shape[0] and _same(depth.as_tensor().cpu().numpy(), odepth)
        K = K / 2
        K[2, 2] = 1
    # finite fill value: the `== invalid_fill` branches
    d0 = np.where(np.isnan(oracle.clip_transform(da)), np.float32(0), oracle.clip_transform(da))
    img = Image(torch.from_numpy(d0).cuda())
    assert _same(img.pyr_down_depth(0.14, 0.0).as_tensor().cpu().numpy(), oracle.pyr_down_depth(d0, 0.14, 0.0))
    v0 = img.create_vertex_map(PRIMESENSE_K, 0.0)
    assert _same(v0.as_tensor().cpu().numpy(), oracle.create_vertex_map(d0, PRIMESENSE_K, 0.0))
    assert _same(v0.create_normal_map(0.0).as_tensor().cpu().numpy(),
                 oracle.create_normal_map(oracle.create_vertex_map(d0, PRIMESENSE_K, 0.0), 0.0))
    with pytest.raises(RuntimeError, match="Kernel size must be >= 3"):
        img.filter_bilateral(1)
    with pytest.raises(RuntimeError, match="1 channel"):
        v0.pyr_down_depth(0.1)


def _maps(da, db, level=0):
    ds, dt = oracle.clip_transform(db), oracle.clip_transform(da)
    K = np.array(PRIMESENSE_K, np.float64)
    for _ in range(level):
        ds, dt = oracle.pyr_down_depth(ds, 0.14), oracle.pyr_down_depth(dt, 0.14)
        K = K / 2
        K[2, 2] = 1
    sv, tv = oracle.create_vertex_map(ds, K), oracle.create_vertex_map(dt, K)
    tn = oracle.create_normal_map(oracle.create_vertex_map(oracle.filter_bilateral(dt), K))
    return sv, tv, tn, K


@pytest.mark.parametrize("level,offset", [(0, (0.01, -0.02, 0.015)), (2, (0.0, 0.0, 0.0)), (1, (0.03, 0.03, -0.04))])
def test_compute_odometry_result_point_to_plane_vs_oracle(o3d, level, offset):
    """One Gauss-Newton step on identical maps: the 29 sums within 1e-5 of the oracle's f64 sums (north_star
    bound), hence delta, inlier_rmse (= sum HuberLoss / inliers) and fitness."""
    odo = o3d.t.pipelines.odometry
    da, db, T_gt = _pair(100, 103)
    sv, tv, tn, K = _maps(da, db, level)
    T = T_gt.copy()
    T[:3, 3] += offset
    res = odo.compute_odometry_result_point_to_plane(torch.from_numpy(sv).cuda(), torch.from_numpy(tv).cuda(),
                                                     torch.from_numpy(tn).cuda(), K, T, 0.07, 0.05)
    o = oracle.odometry_p2plane_sums(sv, tv, tn, K, T, 0.07, 0.05)
    assert res.sums29[28] == o["sums64"][28] > 0.3 * sv.shape[0] * sv.shape[1]        # same inlier set
    err = np.abs(res.sums29 - o["sums64"])
    assert (err <= 1e-5 * o["abs64"] + 1e-300).all(), (err / (o["abs64"] + 1e-300)).max()
    rc, dT, rmse, fit = oracle.compute_odometry_result_p2plane(sv, tv, tn, K, T, 0.07, 0.05)
    assert rc == 0 and res.fitness == fit
    np.testing.assert_allclose(res.transformation, dT, atol=2e-6)
    np.testing.assert_allclose(res.inlier_rmse, rmse, rtol=1e-5)
    if level == 0:    # a small perturbation at full resolution: the step moves towards the truth
        before = np.abs(T - T_gt).max()
        assert np.abs(res.transformation @ T - T_gt).max() < 0.5 * before


@pytest.mark.parametrize("i,j,criteria,rel", [(100, 103, (10, 5, 3), 0.0), (400, 404, (6, 3, 1), 0.0),
                                              (100, 101, (4, 0, 2), 0.0), (100, 103, (10, 5, 3), 1e-6)])
def test_rgbd_odometry_multi_scale_vs_oracle(o3d, i, j, criteria, rel):
    odo, geo = o3d.t.pipelines.odometry, o3d.t.geometry
    da, db, T_gt = _pair(i, j)
    src = geo.RGBDImage(None, torch.from_numpy(db).cuda())
    tgt = geo.RGBDImage(None, torch.from_numpy(da).cuda())
    crit = [odo.OdometryConvergenceCriteria(c, rel, rel) for c in criteria]
    res, log = odo.rgbd_odometry_multi_scale(src, tgt, PRIMESENSE_K, np.eye(4), SCALE, DMAX, crit,
                                             odo.Method.PointToPlane, odo.OdometryLossParams(), return_log=True)
    ref = oracle.rgbd_odometry_multi_scale_p2plane(db, da, PRIMESENSE_K, criteria=[(c, rel, rel) for c in criteria])
    assert ref["status"] == 0
    if rel == 0.0:
        assert len(log) == len(ref["per_iteration"]) == sum(criteria)
    else:
        # the default 1e-6 relative test fires at the last-bits noise level of a converged level: the exit
        # iteration may differ by one or two steps; the trajectories agree on their common part
        assert abs(len(log) - len(ref["per_iteration"])) <= 2 and len(log) < sum(criteria)
    m = min(len(log), len(ref["per_iteration"])) if rel == 0.0 else 6
    np.testing.assert_allclose(log[:m, 1], ref["per_iteration"][:m, 1], atol=2e-4)          # fitness per step
    np.testing.assert_allclose(log[:m, 0], ref["per_iteration"][:m, 0], rtol=2e-3, atol=1e-9)
    np.testing.assert_allclose(res.transformation, ref["transformation"], atol=2e-5)
    assert abs(res.fitness - ref["fitness"]) < 2e-4 and abs(res.inlier_rmse - ref["inlier_rmse"]) < 1e-8
    if sum(criteria) >= 10:
        np.testing.assert_allclose(res.transformation, T_gt, atol=5e-4)


def test_rgbd_odometry_early_exit_f32_inputs_and_errors(o3d):
    odo, geo = o3d.t.pipelines.odometry, o3d.t.geometry
    da, db, T_gt = _pair(100, 102, holes=False)
    C = odo.OdometryConvergenceCriteria
    # loose relative criteria: every level stops at its second step (RGBDOdometry.cpp:181-189)
    crit = [C(20, 0.5, 0.5), C(20, 0.5, 0.5), C(20, 0.5, 0.5)]
    src = geo.RGBDImage(None, torch.from_numpy(db.astype(np.float32)).cuda())     # Float32 source, UInt16 target
    tgt = geo.RGBDImage(None, torch.from_numpy(da).cuda())
    res, log = odo.rgbd_odometry_multi_scale(src, tgt, PRIMESENSE_K, None, SCALE, DMAX, crit, odo.Method.PointToPlane,
                                             return_log=True)
    ref = oracle.rgbd_odometry_multi_scale_p2plane(db, da, PRIMESENSE_K, criteria=[(20, 0.5, 0.5)] * 3)
    assert len(log) == len(ref["per_iteration"]) < 20      # 0.5 is far from the noise level: same exit steps
    np.testing.assert_allclose(res.transformation, ref["transformation"], atol=2e-5)
    assert res.fitness == pytest.approx(ref["fitness"], abs=2e-4)
    # nothing to track against: all-invalid target -> the 6x6 system is singular, as upstream raises
    empty = geo.RGBDImage(None, torch.zeros((480, 640), dtype=torch.uint16).cuda())
    with pytest.raises(RuntimeError, match="Singular 6x6"):
        odo.rgbd_odometry_multi_scale(src, empty, PRIMESENSE_K, None, SCALE, DMAX, (3, 2, 1), odo.Method.PointToPlane)
    with pytest.raises(RuntimeError, match="PointToPlane"):
        odo.rgbd_odometry_multi_scale(src, tgt, PRIMESENSE_K)                      # upstream default = Hybrid
    with pytest.raises(RuntimeError, match="same size"):
        odo.rgbd_odometry_multi_scale(src, geo.RGBDImage(None, torch.zeros((240, 320), dtype=torch.uint16)),
                                      PRIMESENSE_K, method=odo.Method.PointToPlane)
    assert [c.max_iteration for c in odo._criteria_list((10, 5, 3))] == [10, 5, 3]


def _oracle_slam(frames):
    cap = 12000
    keys = np.zeros((cap, 3), np.int32)
    tsdf = np.zeros((cap, RES ** 3), np.float32)
    wt = np.zeros((cap, RES ** 3), np.uint16)
    col = np.zeros((cap, RES ** 3, 3), np.uint16)
    size, pose, poses, model_depth = 0, camera_pose(frames[0]).copy(), [], None
    for n, fid in enumerate(frames):
        depth, color = render_depth(camera_pose(fid), with_color=True)
        depth, color = depth.numpy(), color.numpy()
        if n > 0:
            res = oracle.rgbd_odometry_multi_scale_p2plane(depth.astype(np.float32), model_depth, PRIMESENSE_K,
                                                           criteria=[(6, 1e-6, 1e-6), (3, 1e-6, 1e-6), (1, 1e-6, 1e-6)])
            assert res["status"] == 0
            pose = pose @ res["transformation"]
        poses.append(pose.copy())
        E = oracle.inverse_transformation(pose)
        want = oracle.depth_touch(depth, PRIMESENSE_K, E, RES, VOXEL, VOXEL * TRUNC, SCALE, DMAX, 4)
        bi, _, size, r = oracle.hashmap_activate(keys, size, want)
        assert r == 0
        oracle.tsdf_integrate(depth, color, bi, keys, tsdf, wt, col, PRIMESENSE_K, PRIMESENSE_K, E, RES, VOXEL,
                              VOXEL * TRUNC, SCALE, DMAX)
        rng = oracle.estimate_range(want, PRIMESENSE_K, E, 480, 640, 8, RES, VOXEL, DMIN, DMAX)
        rc = oracle.ray_cast(keys, size, tsdf, wt, col, rng, PRIMESENSE_K, E, 480, 640, ("depth", "color"), RES,
                             VOXEL, SCALE, DMIN, DMAX, min(n * 1.0, 3.0), TRUNC, 8)     # Model.cpp:45-47
        model_depth = np.ascontiguousarray(rc["depth"][..., 0])
    return poses


def test_dense_slam_loop_vs_oracle(o3d):
    """examples/python/t_reconstruction_system/dense_slam.py:47-62, per frame:  track_frame_to_model (against the
    model frame synthesized at the end of the previous iteration) -> update_frame_pose -> integrate ->
    synthesize_model_frame, with the poses ESTIMATED (no ground truth enters after frame 0)."""
    slam = o3d.t.pipelines.slam
    frames = list(range(100, 107))
    model = slam.Model(VOXEL, RES, 12000, camera_pose(frames[0]))
    pose, poses = camera_pose(frames[0]).copy(), []
    raycast_frame = slam.Frame(480, 640, PRIMESENSE_K)
    for n, fid in enumerate(frames):
        depth, color = render_depth(camera_pose(fid), with_color=True)
        frame = slam.Frame(480, 640, PRIMESENSE_K)
        frame.set_data("depth", depth.cuda())
        frame.set_data("color", color.cuda())
        if n > 0:
            res = model.track_frame_to_model(frame, raycast_frame, SCALE, DMAX, 0.07)
            assert 0.8 < res.fitness <= 1.0
            pose = pose @ res.transformation
        poses.append(pose.copy())
        model.update_frame_pose(n, pose)
        model.integrate(frame, SCALE, DMAX, TRUNC)
        model.synthesize_model_frame(raycast_frame, SCALE, DMIN, DMAX, TRUNC, False)
    ref = _oracle_slam(frames)
    for n, (p, q) in enumerate(zip(poses, ref)):
        np.testing.assert_allclose(p, q, atol=1e-4, err_msg=f"frame {n}")        # CUDA loop == oracle loop
    # against the ground truth: the nearest-voxel ray march biases the model depth by O(voxel/2) (see
    # test_oracle_raycast.py), so frame-to-model tracking drifts by millimetres per frame, as the algorithm does
    gt = camera_pose(frames[-1])
    assert np.linalg.norm(poses[-1][:3, 3] - gt[:3, 3]) < 0.03 and np.abs(poses[-1][:3, :3] - gt[:3, :3]).max() < 6e-3
    moved = np.linalg.norm(gt[:3, 3] - camera_pose(frames[0])[:3, 3])
    assert np.linalg.norm(poses[-1][:3, 3] - camera_pose(frames[0])[:3, 3]) > 0.4 * moved   # and it did track
