"""torchrun --nproc-per-node N tests/multigpu_check.py
N-rank check of the source-sharded ICP (30-double NCCL all-reduce per iteration inside
libo3db200.so): the sharded result must equal the single-GPU result on the whole source,
and every rank must hold the identical transformation."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open3d_b200 import _lib as L  # noqa: E402
from open3d_b200.distributed import Communicator, shard_range  # noqa: E402
from tests.synth import make_colors, make_icp_pair  # noqa: E402


def run(src, tgt, nrm, iters, comm):
    stream = int(torch.cuda.current_stream().cuda_stream)
    d = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (src, tgt, nrm)]
    opt = L.IcpOptions()
    opt.max_correspondence_distance, opt.max_iteration = 0.05, iters
    opt.relative_fitness = opt.relative_rmse = 0.0
    opt.kernel = L.RobustKernel(0, 1.0, 1.0)
    h = C.c_void_p()
    T0 = np.eye(4)
    L.check(L.lib.o3db_icp_create(d[0].data_ptr(), len(src), d[1].data_ptr(), d[2].data_ptr(), len(tgt), L.dptr(T0),
                                  C.byref(opt), comm.handle if comm else None, stream, C.byref(h)))
    L.check(L.lib.o3db_icp_iterate(h, iters, stream))
    res = L.IcpResult()
    per = np.zeros((iters, 2))
    L.check(L.lib.o3db_icp_finish(h, C.byref(res), None, L.dptr(per), stream))
    L.lib.o3db_icp_destroy(h)
    return np.array(res.transformation).reshape(4, 4), res.fitness, res.inlier_rmse, per


def run_colored(src, sc, tgt, nrm, tc, grad, iters, comm):
    """ColoredICP through o3db_icp_create_colored with the communicator (BASELINE configs[3]'s sharding)."""
    stream = int(torch.cuda.current_stream().cuda_stream)
    d = [torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda() for a in (src, sc, tgt, nrm, tc)]
    opt = L.IcpOptions()
    opt.max_correspondence_distance, opt.max_iteration = 0.05, iters
    opt.relative_fitness = opt.relative_rmse = 0.0
    opt.kernel = L.RobustKernel(0, 1.0, 1.0)
    h = C.c_void_p()
    L.check(L.lib.o3db_icp_create_colored(d[0].data_ptr(), d[1].data_ptr(), len(src), d[2].data_ptr(), d[3].data_ptr(),
                                          d[4].data_ptr(), grad.data_ptr(), len(tgt), L.dptr(np.eye(4)), C.byref(opt),
                                          0.968, comm.handle if comm else None, stream, C.byref(h)))
    L.check(L.lib.o3db_icp_iterate(h, iters, stream))
    res = L.IcpResult()
    per = np.zeros((iters, 2))
    L.check(L.lib.o3db_icp_finish(h, C.byref(res), None, L.dptr(per), stream))
    L.lib.o3db_icp_destroy(h)
    return np.array(res.transformation).reshape(4, 4), res.fitness, res.inlier_rmse, per


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    comm = Communicator(rank, world)
    src, tgt, nrm, T_gt = make_icp_pair(400_000, seed=5)
    b, e = shard_range(len(src), rank, world)
    T, fit, rmse, per = run(src[b:e], tgt, nrm, 10, comm)
    # every rank holds the same answer
    t = torch.from_numpy(T).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(t, ref), "ranks disagree on the transformation"
    if rank == 0:
        T1, fit1, rmse1, per1 = run(src, tgt, nrm, 10, None)
        np.testing.assert_allclose(T, T1, atol=1e-9)
        np.testing.assert_allclose(per, per1, atol=1e-9)
        assert abs(fit - fit1) < 1e-12 and abs(rmse - rmse1) < 1e-9
        np.testing.assert_allclose(T, T_gt, atol=2e-3)
        peer = int(L.lib.o3db_comm_uses_peer_memory(comm.handle))
        print(f"multigpu_check ok: world={world} peer_memory_exchange={peer} fitness={fit:.6f} rmse={rmse:.6f} "
              f"|T - T_single|max={np.abs(T - T1).max():.2e}")
    # ColoredICP, same sharding: the target (with colours and gradients) is replicated, the source split
    src, tgt, nrm, T_gt = make_icp_pair(300_000, seed=6)
    sc = make_colors((np.c_[src.astype(np.float64), np.ones(len(src))] @ T_gt.T)[:, :3], 1)
    tc = make_colors(tgt, 1)
    stream = int(torch.cuda.current_stream().cuda_stream)
    dt = [torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda() for a in (tgt, nrm, tc)]
    grad = torch.empty_like(dt[0])
    L.check(L.lib.o3db_estimate_color_gradients(dt[0].data_ptr(), dt[1].data_ptr(), dt[2].data_ptr(), len(tgt), 0.08, 30,
                                                grad.data_ptr(), stream))
    b, e = shard_range(len(src), rank, world)
    T, fit, rmse, per = run_colored(src[b:e], sc[b:e], tgt, nrm, tc, grad, 8, comm)
    t = torch.from_numpy(T).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(t, ref), "ranks disagree on the coloured transformation"
    if rank == 0:
        T1, fit1, rmse1, per1 = run_colored(src, sc, tgt, nrm, tc, grad, 8, None)
        np.testing.assert_allclose(T, T1, atol=1e-9)
        np.testing.assert_allclose(per, per1, atol=1e-9)
        assert abs(fit - fit1) < 1e-12 and abs(rmse - rmse1) < 1e-9
        print(f"multigpu_check colored ok: world={world} fitness={fit:.6f} rmse={rmse:.6f} |T - T_single|max={np.abs(T - T1).max():.2e}")
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
