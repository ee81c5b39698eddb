"""Short workload for ncu captures: 6 ICP iterations on the 2M-point bench clouds and
40 fused TSDF frames (depth + colour).  Usage (under gpurun):
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
from tests.synth import PRIMESENSE_K, camera_pose, make_icp_pair, render_depth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
stream = int(torch.cuda.current_stream().cuda_stream)
if what in ("icp", "all"):
    n = int(os.environ.get("ICP_POINTS", 2_000_000))
    src, tgt, nrm, _ = make_icp_pair(n, seed=2)
    d = [torch.from_numpy(a).cuda() for a in (src, tgt, nrm)]
    opt = L.IcpOptions()
    opt.max_correspondence_distance, opt.max_iteration = 0.05, 6
    opt.kernel = L.RobustKernel(0, 1.0, 1.0)
    opt.cell_scale = float(os.environ.get("CELL_SCALE", 0))
    h = C.c_void_p()
    T0 = np.eye(4)
    L.check(L.lib.o3db_icp_create(d[0].data_ptr(), n, d[1].data_ptr(), d[2].data_ptr(), n, L.dptr(T0), C.byref(opt), None,
                                  stream, C.byref(h)))
    L.check(L.lib.o3db_icp_iterate(h, 6, stream))
    res = L.IcpResult()
    L.check(L.lib.o3db_icp_finish(h, C.byref(res), None, None, stream))
    print("icp", res.fitness, res.inlier_rmse, res.num_iterations)
    L.lib.o3db_icp_destroy(h)
if what in ("tsdf", "all"):
    v = C.c_void_p()
    L.check(L.lib.o3db_vbg_create(0.008, 16, 40000, 1, stream, C.byref(v)))
    K = np.ascontiguousarray(PRIMESENSE_K)
    for i in range(40):
        T = camera_pose(i * 5)
        E = np.eye(4)
        E[:3, :3] = T[:3, :3].T
        E[:3, 3] = -(T[:3, :3].T @ T[:3, 3])
        dep, col = render_depth(T, device="cuda", with_color=True)
        L.check(L.lib.o3db_vbg_integrate_frame(v, dep.data_ptr(), L.DEPTH_U16, col.data_ptr(), L.COLOR_U8, 480, 640, L.dptr(K),
                                               L.dptr(np.ascontiguousarray(E)), 1000.0, 3.0, 8.0, stream))
    print("tsdf blocks", L.lib.o3db_vbg_size(v, stream))
    L.lib.o3db_vbg_destroy(v)
