"""Wall-clock ms/frame of the dense-SLAM loop (track + integrate + ray cast, public API) over N frames, after a warm-up;
also the odometry call alone.  Usage (under gpurun): python profiles/slam_time.py [frames=100]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import open3d_b200  # noqa: E402
from tests.synth import PRIMESENSE_K, camera_pose, render_depth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
start = int(os.environ.get("SLAM_START", 0))      # first frame of the segment (bench.py gives rank r the frames 100 r ...)
if "LOCAL_RANK" in os.environ:                     # under torchrun: one segment per rank, as bench.py does
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    start = 100 * int(os.environ["LOCAL_RANK"])
slam = open3d_b200.t.pipelines.slam
frames = []
for i in range(n):
    d, c = render_depth(camera_pose(start + i), device="cuda", with_color=True)
    frames.append((d.contiguous(), c.contiguous()))


def loop(k, parts=None):
    T0 = camera_pose(start)
    model = slam.Model(0.008, 16, 40000, T0)
    pose = T0.copy()
    rc = slam.Frame(480, 640, PRIMESENSE_K)
    for f in range(k):
        fr = slam.Frame(480, 640, PRIMESENSE_K)
        fr.set_data("depth", frames[f][0])
        fr.set_data("color", frames[f][1])
        t0 = time.perf_counter()
        if f > 0:
            pose = pose @ model.track_frame_to_model(fr, rc, 1000.0, 3.0, 0.07).transformation
        t1 = time.perf_counter()
        model.update_frame_pose(f, pose)
        model.integrate(fr, 1000.0, 3.0, 8.0)
        model.synthesize_model_frame(rc, 1000.0, 0.1, 3.0, 8.0, False)
        if parts is not None:
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            parts.append((t1 - t0, t2 - t1))
    torch.cuda.synchronize()
    return pose


loop(8)
t0 = time.perf_counter()
pose = loop(n)
dt = time.perf_counter() - t0
parts = []
loop(n, parts)
p = np.array(parts[1:]) * 1e3
worst = int(np.argmax(p.sum(1)))
print(f"start {start} slam ms/frame {1e3 * dt / n:.4f}  (track {p[:, 0].mean():.4f} ms, integrate+raycast synced {p[:, 1].mean():.4f} ms)  "
      f"worst frame {worst + 1}: {p[worst, 0]:.3f} + {p[worst, 1]:.3f} ms  drift mm {1e3 * np.linalg.norm(pose[:3, 3] - camera_pose(start + n - 1)[:3, 3]):.2f}")
