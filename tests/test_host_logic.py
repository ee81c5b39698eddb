"""Host-side logic of the reference-facing Python surface, without a GPU:
argument validation mirrors AssertInputMultiScaleICP (Registration.cpp:119-219)
and the product fails loudly when no CUDA device exists (no CPU fallback)."""
import numpy as np
import pytest
import torch

import open3d_b200 as o3d
from open3d_b200.t.pipelines import registration as reg

needs_no_gpu = pytest.mark.skipif(torch.cuda.is_available(), reason="exercises the no-GPU failure path")


def _clouds(n=64):
    rng = np.random.default_rng(0)
    p = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    src = o3d.t.geometry.PointCloud(p)
    tgt = o3d.t.geometry.PointCloud(p.copy())
    tgt.set_point_normals(np.tile(np.array([[0, 0, 1]], np.float32), (n, 1)))
    return src, tgt


def test_defaults_match_reference():
    c = reg.ICPConvergenceCriteria()
    assert (c.relative_fitness, c.relative_rmse, c.max_iteration) == (1e-6, 1e-6, 30)   # Registration.h:43-48
    k = reg.robust_kernel.RobustKernel()
    assert (int(k.type), k.scaling_parameter, k.shape_parameter) == (0, 1.0, 1.0)
    assert [m.name for m in reg.RobustKernelMethod] == ["L2Loss", "L1Loss", "HuberLoss", "CauchyLoss", "GMLoss",
                                                        "TukeyLoss", "GeneralizedLoss"]
    assert reg.TransformationEstimationForColoredICP(lambda_geometric=7).lambda_geometric == 0.968


def test_input_validation_errors():
    src, tgt = _clouds()
    est = reg.TransformationEstimationPointToPlane()
    with pytest.raises(RuntimeError, match="Max correspondence distance"):
        reg.icp(src, tgt, 0.0, np.eye(4), est)
    with pytest.raises(RuntimeError, match="normal"):
        reg.icp(src, o3d.t.geometry.PointCloud(src.point["positions"]), 0.1, np.eye(4), est)
    with pytest.raises(RuntimeError, match="empty"):
        reg.icp(o3d.t.geometry.PointCloud(), tgt, 0.1, np.eye(4), est)
    with pytest.raises(RuntimeError, match="must be same"):
        reg.multi_scale_icp(src, tgt, [-1, -1], [reg.ICPConvergenceCriteria()], [0.1, 0.1], np.eye(4), est)
    with pytest.raises(RuntimeError, match="decreasing"):
        reg.multi_scale_icp(src, tgt, [0.01, 0.02], [reg.ICPConvergenceCriteria()] * 2, [0.1, 0.1], np.eye(4), est)
    with pytest.raises(RuntimeError, match="strictly decreasing"):      # Registration.cpp:190-200: equal sizes are rejected too
        reg.multi_scale_icp(src, tgt, [0.02, 0.02], [reg.ICPConvergenceCriteria()] * 2, [0.1, 0.1], np.eye(4), est)
    with pytest.raises(RuntimeError, match="in scale: 1"):              # the message names the offending scale
        reg.multi_scale_icp(src, tgt, [0.04, 0.02], [reg.ICPConvergenceCriteria()] * 2, [0.1, 0.0], np.eye(4), est)
    with pytest.raises(RuntimeError, match="empty"):
        reg.get_information_matrix(o3d.t.geometry.PointCloud(), tgt, 0.1, np.eye(4))
    tgt_c = o3d.t.geometry.PointCloud(tgt.point["positions"]).set_point_normals(tgt.point["normals"])
    tgt_c.set_point_colors(np.zeros((64, 3), np.float32))
    with pytest.raises(ValueError, match="solver"):
        tgt_c.estimate_color_gradients(30, 0.1, solver="fast")
    with pytest.raises(RuntimeError, match=r"\[4, 4\]"):
        reg.icp(src, tgt, 0.1, np.eye(3), est)
    with pytest.raises(RuntimeError, match="Float32"):
        o3d.t.geometry.PointCloud(np.zeros((4, 3), np.float64))
    with pytest.raises(RuntimeError, match=r"\[N, 3\]"):
        o3d.t.geometry.PointCloud(np.zeros((4, 2), np.float32))


@needs_no_gpu
def test_no_cpu_fallback_icp():
    src, tgt = _clouds()
    with pytest.raises(RuntimeError, match="(?i)cuda"):
        reg.icp(src, tgt, 0.1, np.eye(4), reg.TransformationEstimationPointToPlane())


@needs_no_gpu
def test_no_cpu_fallback_tsdf():
    with pytest.raises(RuntimeError, match="(?i)cuda"):
        o3d.t.pipelines.slam.Model(0.008, 16, 100)


def test_shard_range_partitions():
    from open3d_b200.distributed import shard_range
    for n, w in [(10, 3), (2_000_000, 8), (5, 8), (0, 2)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1


def test_colored_icp_and_odometry_host_side():
    """argument checks that run before any device work, and the defaults of the mirrored classes"""
    from open3d_b200.t.pipelines import odometry as odo
    src, tgt = _clouds()
    est = reg.TransformationEstimationForColoredICP()
    assert est.lambda_geometric == 0.968 and int(est.kernel.type) == 0
    with pytest.raises(RuntimeError, match="target pointcloud to have colors"):        # Registration.cpp:160-176
        reg.icp(src, tgt, 0.1, np.eye(4), est)
    tgt.set_point_colors(np.zeros((64, 3), np.float32))
    with pytest.raises(RuntimeError, match="source pointcloud to have colors"):
        reg.icp(src, tgt, 0.1, np.eye(4), est)
    with pytest.raises(RuntimeError, match="PointToPlane and"):
        reg.icp(src, tgt, 0.1, np.eye(4), object())
    p = odo.OdometryLossParams()
    assert (p.depth_outlier_trunc, p.depth_huber_delta, p.intensity_huber_delta) == (0.07, 0.05, 0.1)   # RGBDOdometry.h:82-84
    c = odo.OdometryConvergenceCriteria(7)
    assert (c.max_iteration, c.relative_rmse, c.relative_fitness) == (7, 1e-6, 1e-6)                     # :35-37
    assert [m.name for m in odo.Method] == ["PointToPlane", "Intensity", "Hybrid"]
    assert [k.max_iteration for k in odo._criteria_list((6, 3, 1))] == [6, 3, 1]                          # Model.h:83-84
    r = odo.OdometryResult()
    assert np.array_equal(r.transformation, np.eye(4)) and r.inlier_rmse == 0.0 and r.fitness == 0.0
    d = torch.zeros((8, 8), dtype=torch.uint16)
    with pytest.raises(RuntimeError, match="PointToPlane"):                              # default method = Hybrid upstream
        odo.rgbd_odometry_multi_scale(o3d.t.geometry.RGBDImage(None, d), o3d.t.geometry.RGBDImage(None, d), np.eye(3))
    with pytest.raises(RuntimeError, match="Unsupported attribute"):
        o3d.t.geometry.VoxelBlockGrid.ray_cast(object.__new__(o3d.t.geometry.VoxelBlockGrid), None, np.eye(3), np.eye(4),
                                               8, 8, ("albedo",))


@needs_no_gpu
def test_no_cpu_fallback_new_entry_points():
    """every C entry point added for the widened rows fails loudly without a device"""
    import ctypes as C
    from open3d_b200 import _lib as L
    a = np.zeros((16, 16), np.float32)
    out = np.zeros((16, 16, 3), np.float32)
    K = np.ascontiguousarray(np.eye(3))
    T = np.ascontiguousarray(np.eye(4))
    calls = [
        lambda: L.lib.o3db_image_clip_transform(a.ctypes.data, L.DEPTH_F32, 16, 16, 1.0, 0.0, 3.0, 0.0, out.ctypes.data, None),
        lambda: L.lib.o3db_image_create_vertex_map(a.ctypes.data, 16, 16, L.dptr(K), 0.0, out.ctypes.data, None),
        lambda: L.lib.o3db_image_filter_bilateral(a.ctypes.data, 16, 16, 5, 5.0, 10.0, out.ctypes.data, None),
        lambda: L.lib.o3db_estimate_color_gradients(out.ctypes.data, out.ctypes.data, out.ctypes.data, 256, 0.1, 30,
                                                    out.ctypes.data, None),
    ]
    for call in calls:
        rc = call()
        assert rc == L.ERR_CUDA and "cuda" in L.last_error().lower(), (rc, L.last_error())
    res = L.OdometryResult()
    crit = (L.OdometryCriteria * 1)(L.OdometryCriteria(1, 1e-6, 1e-6))
    rc = L.lib.o3db_rgbd_odometry_multi_scale_point_to_plane(a.ctypes.data, L.DEPTH_F32, a.ctypes.data, L.DEPTH_F32, 16, 16,
                                                             L.dptr(K), L.dptr(T), 1.0, 3.0, crit, 1, 0.07, 0.05,
                                                             C.byref(res), None, None)
    assert rc == L.ERR_CUDA


class _FakeFusedVolume:
    """Stand-in for the o3db_vbg_integrate_frame contract (include/open3d_b200.h): the host runs two frames ahead; a
    frame whose blocks do not fit is dropped with every later one; the error surfaces when frame F-2's status is read
    at the submission of frame F, or at a size() call; reserve() re-arms; the first frame after a reserve is sized
    synchronously."""

    class CapacityError(RuntimeError):
        pass

    def __init__(self, capacity):
        self.capacity, self.size, self.frames = capacity, 0, 0
        self.integrated = []          # frame tags that reached the volume, in order
        self.submitted = []           # (library index, tag, dropped?)
        self.dropped_from = None
        self.sync_next = True
        self.reserves = []

    def _raise_if_visible(self, upto):
        if self.dropped_from is not None and self.dropped_from[0] <= upto:
            k, need = self.dropped_from
            raise self.CapacityError(f"voxel block hash map capacity ({self.capacity} blocks) exceeded: fused frame "
                                     f"#{k} needed {need} blocks; that frame and all later ones were dropped")

    def submit(self, tag, blocks):
        self._raise_if_visible(self.frames - 2)
        if self.sync_next:
            self.sync_next = False
            self.capacity = max(self.capacity, self.size + blocks)
        dropped = self.dropped_from is not None or self.size + blocks > self.capacity
        if dropped and self.dropped_from is None:
            self.dropped_from = (self.frames, self.size + blocks)
        if not dropped:
            self.size += blocks
            self.integrated.append(tag)
        self.submitted.append((self.frames, tag, dropped))
        self.frames += 1

    def get_size(self):
        self._raise_if_visible(self.frames)
        return self.size

    def reserve(self, n):
        self.reserves.append(n)
        self.capacity = max(self.capacity, n)
        self.dropped_from = None
        self.sync_next = True


def _replay_for(vol):
    from open3d_b200.t.geometry._replay import FrameReplay
    return FrameReplay(vol.submit, vol.reserve, lambda: vol.capacity, lambda e: isinstance(e, vol.CapacityError))


def test_frame_replay_restores_hashmap_activate_growth():
    """HashMap::Activate grows on demand (HashMap.cpp:166-181): through FrameReplay the caller never sees the fused
    path's dropped-frame error, every frame is integrated exactly once and in order."""
    from open3d_b200.t.geometry._replay import parse_dropped_frame
    assert parse_dropped_frame("... fused frame #12 needed 3456 blocks; that frame") == (12, 3456)
    assert parse_dropped_frame("singular 6x6") is None
    vol = _FakeFusedVolume(1000)
    rp = _replay_for(vol)
    blocks = [800, 50, 50, 5000, 40, 30, 20000, 10, 10]     # two jumps that do not fit
    for tag, b in enumerate(blocks):
        rp.submit(tag, b)
    assert rp.guard(vol.get_size) == sum(blocks)
    assert vol.integrated == list(range(len(blocks)))
    assert rp.recoveries == 2 and len(vol.reserves) == 2
    assert vol.reserves[0] >= 5900 + 2048 and vol.reserves[1] >= 2 * vol.reserves[0]
    # a drop of the very last frames is only visible to the guarded call
    vol = _FakeFusedVolume(1000)
    rp = _replay_for(vol)
    for tag, b in enumerate([500, 100, 9000]):
        rp.submit(tag, b)
    assert vol.integrated == [0, 1]
    assert rp.guard(vol.get_size) == 9600 and vol.integrated == [0, 1, 2]


def test_frame_replay_hands_over_what_it_cannot_recover():
    vol = _FakeFusedVolume(1000)
    rp = _replay_for(vol)
    rp.enabled = False
    with pytest.raises(_FakeFusedVolume.CapacityError, match="fused frame #1 needed 5500 blocks"):
        for tag, b in enumerate([500, 5000, 10, 10, 10]):
            rp.submit(tag, b)
    # other errors pass through untouched, and the failed frame is not remembered
    from open3d_b200.t.geometry._replay import FrameReplay

    def boom(*_):
        raise ValueError("bad image")
    rp = FrameReplay(boom, lambda n: None, lambda: 0, lambda e: False)
    with pytest.raises(ValueError, match="bad image"):
        rp.submit(0, 1)
    assert len(rp._ring) == 0 and rp._next == 0
    # a dropped frame older than the ring cannot be resubmitted from here
    vol = _FakeFusedVolume(1000)
    rp = _replay_for(vol)
    for tag in range(3):
        rp.submit(tag, 10)
    rp._ring.clear()
    vol.dropped_from = (1, 4000)
    with pytest.raises(_FakeFusedVolume.CapacityError):
        rp.guard(vol.get_size)


def test_frame_loop_helpers_of_the_python_mirror():
    """Small host-side pieces on the dense-SLAM loop's critical path: they must stay exact while being cheap."""
    import oracle
    from open3d_b200.t.pipelines import odometry, slam
    # InverseTransformation (t/geometry/Utility.h:77-115): bit-identical to the oracle's restatement
    rng = np.random.default_rng(7)
    for _ in range(50):
        T = np.eye(4)
        T[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        T[:3, 3] = rng.normal(size=3) * 3
        assert np.array_equal(slam._inverse_transformation(T).view(np.uint64), oracle.inverse_transformation(T).view(np.uint64))
    # Frame.get_data: an absent name is an empty tensor (Frame.h), a present one the very object
    f = slam.Frame(4, 4, np.eye(3))
    d = torch.zeros((4, 4), dtype=torch.uint16)
    f.set_data("depth", d)
    assert f.get_data("depth") is d and f.get_data("color").numel() == 0
    # criteria lists: ints convert like vector<OdometryConvergenceCriteria>{6, 3, 1}; the tuple-of-ints case is cached
    a1, n1, it1 = odometry._c_criteria((6, 3, 1))
    a2, n2, it2 = odometry._c_criteria((6, 3, 1))
    assert a1 is a2 and (n1, it1) == (3, 10) and [a1[i].max_iteration for i in range(3)] == [6, 3, 1]
    a3, n3, it3 = odometry._c_criteria([odometry.OdometryConvergenceCriteria(4, 1e-3, 1e-3), 2])
    assert (n3, it3) == (2, 6) and a3[0].relative_rmse == 1e-3 and a3[1].max_iteration == 2
