"""Diagnostics (library built with -DICP_TIMING=1): where one steady-state ICP iteration spends its time.
Per block the kernel records %globaltimer at: 0 resident, 1 predecessor complete (PDL wait passed), 2 thread 0's
warp through its loop, 3 all warps of the block through (first barrier of the epilogue), 4 last block: ticket
drawn, 5 grand total ready, 7 solve + pose update done; 6 = SM id."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open3d_b200 import _lib as L  # noqa: E402
from tests.synth import make_icp_pair  # noqa: E402

n = int(os.environ.get("ICP_POINTS", 2_000_000))
stream = int(torch.cuda.current_stream().cuda_stream)
src, tgt, nrm, T_gt = make_icp_pair(n, seed=2)
d = [torch.from_numpy(a).cuda() for a in (src, tgt, nrm)]
opt = L.IcpOptions()
opt.max_correspondence_distance, opt.max_iteration = 0.05, 30
opt.kernel = L.RobustKernel(0, 1.0, 1.0)
opt.search_variant = int(os.environ.get("ICP_VARIANT", 0))
h = C.c_void_p()
L.check(L.lib.o3db_icp_create(d[0].data_ptr(), len(src), d[1].data_ptr(), d[2].data_ptr(), len(tgt), L.dptr(np.eye(4)),
                              C.byref(opt), None, stream, C.byref(h)))
L.lib.o3db_icp_debug_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
for it in (12, 13):
    L.check(L.lib.o3db_icp_iterate(h, it if it == 12 else 1, stream))
    torch.cuda.synchronize()
    buf = np.zeros((2048, 8), np.int64)
    nb = C.c_int(0)
    L.lib.o3db_icp_debug_timing(h, buf.ctypes.data, 2048, C.byref(nb))
    t = buf[:nb.value].astype(np.float64)
    t0 = t[:, 0].min()
    last = np.argmax(t[:, 7])
    rel = lambda x: (x - t0) / 1e3
    print(f"after {it if it == 12 else 13} iterations: blocks {nb.value}")
    print("  resident      min/med/max us", *(f"{v:.1f}" for v in np.percentile(rel(t[:, 0]), [0, 50, 100])))
    print("  pdl released  min/med/max us", *(f"{v:.1f}" for v in np.percentile(rel(t[:, 1]), [0, 50, 100])))
    print("  loop done     min/p10/med/p90/max us", *(f"{v:.1f}" for v in np.percentile(rel(t[:, 2]), [0, 10, 50, 90, 100])))
    print("  loop duration min/med/max us", *(f"{v:.1f}" for v in np.percentile((t[:, 2] - t[:, 1]) / 1e3, [0, 50, 100])))
    dur = (t[:, 2] - t[:, 1]) / 1e3
    sm = buf[:nb.value, 6]
    by_sm = {}
    for b in range(nb.value):
        by_sm.setdefault(int(sm[b]), []).append(dur[b])
    sm_mean = np.array([np.mean(v) for v in by_sm.values()])
    sm_spread = np.array([max(v) - min(v) for v in by_sm.values()])
    cnt = np.bincount([len(v) for v in by_sm.values()])
    print(f"  blocks per SM histogram {cnt.tolist()}; per-SM mean loop us min/med/max",
          *(f"{v:.1f}" for v in np.percentile(sm_mean, [0, 50, 100])), "; within-SM spread med/max",
          *(f"{v:.1f}" for v in np.percentile(sm_spread, [50, 100])))
    order = np.argsort(dur)
    print("  slowest blocks (block, sm, us):", [(int(b), int(sm[b]), round(float(dur[b]), 1)) for b in order[-6:]])
    print("  fastest blocks (block, sm, us):", [(int(b), int(sm[b]), round(float(dur[b]), 1)) for b in order[:6]])
    print("  block done    min/p10/med/p90/max us", *(f"{v:.1f}" for v in np.percentile(rel(t[:, 3]), [0, 10, 50, 90, 100])))
    print(f"  last block {last}: all warps done {rel(t[last, 3]):.1f}, ticket {rel(t[last, 4]):.1f}, total ready "
          f"{rel(t[last, 5]):.1f}, finalize done {rel(t[last, 7]):.1f} us  (publish+ticket {(t[last, 4] - t[last, 3]) / 1e3:.1f}, "
          f"sum {(t[last, 5] - t[last, 4]) / 1e3:.1f}, solve {(t[last, 7] - t[last, 5]) / 1e3:.1f} us)")
L.lib.o3db_icp_destroy(h)
