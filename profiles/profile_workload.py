"""Short workload for ncu captures: 6 ICP iterations on the 2M-point bench clouds and
40 fused TSDF frames (depth + colour); "colored" = colour gradients + 6 ColoredICP iterations at 500 k points;
"raycast" = 5 SynthesizeModelFrame-style ray casts after the 40 frames; "slam" = 6 frames of the dense-SLAM loop.  Usage (under gpurun):
  ncu --set full --clock-control none --import-source on -k regex:icp_iteration_kernel -s 3 -c 2 \
      -o gpurun_out/icp python profiles/profile_workload.py icp
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open3d_b200 import _lib as L  # noqa: E402
from tests.synth import PRIMESENSE_K, camera_pose, make_colors, make_icp_pair, render_depth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
stream = int(torch.cuda.current_stream().cuda_stream)
if what in ("icp", "all"):
    n = int(os.environ.get("ICP_POINTS", 2_000_000))
    src, tgt, nrm, _ = make_icp_pair(n, seed=2)
    d = [torch.from_numpy(a).cuda() for a in (src, tgt, nrm)]
    opt = L.IcpOptions()
    icp_iters = int(os.environ.get("ICP_ITERS", 6))
    opt.max_correspondence_distance, opt.max_iteration = 0.05, icp_iters
    opt.kernel = L.RobustKernel(0, 1.0, 1.0)
    opt.cell_scale = float(os.environ.get("CELL_SCALE", 0))
    opt.search_variant = int(os.environ.get("ICP_VARIANT", 0))
    h = C.c_void_p()
    T0 = np.eye(4)
    L.check(L.lib.o3db_icp_create(d[0].data_ptr(), n, d[1].data_ptr(), d[2].data_ptr(), n, L.dptr(T0), C.byref(opt), None,
                                  stream, C.byref(h)))
    L.check(L.lib.o3db_icp_iterate(h, icp_iters, stream))
    res = L.IcpResult()
    L.check(L.lib.o3db_icp_finish(h, C.byref(res), None, None, stream))
    print("icp", res.fitness, res.inlier_rmse, res.num_iterations)
    L.lib.o3db_icp_destroy(h)
if what in ("colored", "all"):
    n = int(os.environ.get("COLORED_POINTS", 500_000))
    src, tgt, nrm, T_gt = make_icp_pair(n, seed=2)
    sc = make_colors((np.c_[src.astype(np.float64), np.ones(len(src))] @ T_gt.T)[:, :3], 1)
    tc = make_colors(tgt, 1)
    d = [torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda() for a in (src, sc, tgt, nrm, tc)]
    grad = torch.empty_like(d[2])
    L.check(L.lib.o3db_estimate_color_gradients(d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(), len(tgt), 0.08, 30,
                                                grad.data_ptr(), stream))
    opt = L.IcpOptions()
    opt.max_correspondence_distance, opt.max_iteration = 0.05, 6
    opt.kernel = L.RobustKernel(0, 1.0, 1.0)
    h = C.c_void_p()
    L.check(L.lib.o3db_icp_create_colored(d[0].data_ptr(), d[1].data_ptr(), len(src), d[2].data_ptr(), d[3].data_ptr(),
                                          d[4].data_ptr(), grad.data_ptr(), len(tgt), L.dptr(np.eye(4)), C.byref(opt),
                                          0.968, None, stream, C.byref(h)))
    L.check(L.lib.o3db_icp_iterate(h, 6, stream))
    res = L.IcpResult()
    L.check(L.lib.o3db_icp_finish(h, C.byref(res), None, None, stream))
    print("colored icp", res.fitness, res.inlier_rmse, res.num_iterations)
    L.lib.o3db_icp_destroy(h)
if what in ("tsdf", "raycast", "all"):
    v = C.c_void_p()
    with_color = int(os.environ.get("TSDF_COLOR", 1))
    L.check(L.lib.o3db_vbg_create(0.008, 16, 40000, with_color, stream, C.byref(v)))
    K = np.ascontiguousarray(PRIMESENSE_K)
    for i in range(40):
        T = camera_pose(i * 5)
        E = np.eye(4)
        E[:3, :3] = T[:3, :3].T
        E[:3, 3] = -(T[:3, :3].T @ T[:3, 3])
        dep, col = render_depth(T, device="cuda", with_color=True)
        L.check(L.lib.o3db_vbg_integrate_frame(v, dep.data_ptr(), L.DEPTH_U16, col.data_ptr() if with_color else None,
                                               L.COLOR_U8 if with_color else 0, 480, 640, L.dptr(K),
                                               L.dptr(np.ascontiguousarray(E)), 1000.0, 3.0, 8.0, stream))
    print("tsdf blocks", L.lib.o3db_vbg_size(v, stream))
    if what in ("raycast", "all"):
        rd = torch.empty((480, 640, 1), dtype=torch.float32, device="cuda")
        rc = torch.empty((480, 640, 3), dtype=torch.float32, device="cuda")
        ro = L.RaycastOutputs()
        ro.depth, ro.color = rd.data_ptr(), rc.data_ptr()
        for _ in range(5):
            L.check(L.lib.o3db_vbg_ray_cast(v, None, 0, L.dptr(K), L.dptr(np.ascontiguousarray(E)), 640, 480, C.byref(ro),
                                            1000.0, 0.1, 3.0, 3.0, 8.0, 8, None, stream))
        torch.cuda.synchronize()
        print("raycast hit fraction", float((rd > 0).float().mean()))
    L.lib.o3db_vbg_destroy(v)
if what in ("slam", "all"):
    # 6 frames of the dense-SLAM loop (track_frame_to_model + integrate + synthesize_model_frame), public API
    import open3d_b200
    slam = open3d_b200.t.pipelines.slam
    T0 = camera_pose(100)
    model = slam.Model(0.008, 16, 12000, T0)
    pose = T0.copy()
    rc_frame = slam.Frame(480, 640, PRIMESENSE_K)
    for n in range(6):
        dep, col = render_depth(camera_pose(100 + n), device="cuda", with_color=True)
        fr = slam.Frame(480, 640, PRIMESENSE_K)
        fr.set_data("depth", dep.contiguous())
        fr.set_data("color", col.contiguous())
        if n > 0:
            pose = pose @ model.track_frame_to_model(fr, rc_frame, 1000.0, 3.0, 0.07).transformation
        model.update_frame_pose(n, pose)
        model.integrate(fr, 1000.0, 3.0, 8.0)
        model.synthesize_model_frame(rc_frame, 1000.0, 0.1, 3.0, 8.0, False)
    torch.cuda.synchronize()
    print("slam pose drift mm", 1e3 * float(np.linalg.norm(pose[:3, 3] - camera_pose(105)[:3, 3])))
