"""Per-iteration device time of the fused ICP loop on the bench clouds (2M points, r = 0.05):
one o3db_icp_iterate(1) per CUDA-event pair, L2 flushed only before the first iteration (as in
bench.py's step).  Usage (under gpurun):  python profiles/icp_iter_times.py [iterations=30] [reps=3]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open3d_b200 import _lib as L  # noqa: E402
from tests.synth import make_icp_pair  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = int(os.environ.get("ICP_POINTS", 2_000_000))
stream = int(torch.cuda.current_stream().cuda_stream)
src, tgt, nrm, T_gt = make_icp_pair(n, seed=2)
d = [torch.from_numpy(a).cuda() for a in (src, tgt, nrm)]
opt = L.IcpOptions()
opt.max_correspondence_distance, opt.max_iteration = 0.05, iters
opt.kernel = L.RobustKernel(0, 1.0, 1.0)
opt.cell_scale = float(os.environ.get("CELL_SCALE", 0))
opt.search_variant = int(os.environ.get("ICP_VARIANT", 0))     # 0 = default, 1 = direct loads, 2 = staged (TMA + cp.async)
h = C.c_void_p()
L.check(L.lib.o3db_icp_create(d[0].data_ptr(), len(src), d[1].data_ptr(), d[2].data_ptr(), len(tgt), L.dptr(np.eye(4)),
                              C.byref(opt), None, stream, C.byref(h)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = L.IcpResult()
per_iter = np.zeros((iters, 2))
best = None
for rep in range(reps + 1):
    L.check(L.lib.o3db_icp_reset(h, stream))
    flush.fill_(1)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 2)]
    evs[0].record()
    for k in range(iters):
        L.check(L.lib.o3db_icp_iterate(h, 1, stream))
        evs[k + 1].record()
    L.check(L.lib.o3db_icp_finish(h, C.byref(res), None, L.dptr(per_iter), stream))
    evs[iters + 1].record()
    torch.cuda.synchronize()
    t = np.array([evs[k].elapsed_time(evs[k + 1]) * 1e3 for k in range(iters + 1)])
    best = t if best is None or rep == 1 else np.minimum(best, t)
print("iteration_us", " ".join(f"{x:.1f}" for x in best[:iters]))
print("evaluate_us", f"{best[iters]:.1f}", "mean_iter_us", f"{best[:iters].mean():.1f}", "mean_after_5", f"{best[5:iters].mean():.1f}")
print("fitness", " ".join(f"{x:.4f}" for x in per_iter[:, 0]))
print("rmse", " ".join(f"{x:.5f}" for x in per_iter[:, 1]))
print("final", res.fitness, res.inlier_rmse, "err_vs_gt", float(np.abs(np.array(res.transformation).reshape(4, 4) - T_gt).max()))
L.lib.o3db_icp_destroy(h)
