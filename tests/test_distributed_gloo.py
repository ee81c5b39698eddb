"""world_size-2 gloo tests (CPU) of the multi-process host logic used by the N>1
path: rendezvous, unique-id broadcast, shard partitioning, max/sum reductions of
timings and of a 30-double system (the NCCL all-reduce's CPU stand-in)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from open3d_b200 import distributed as d
    import oracle
    from tests.synth import make_icp_pair
    assert d.env_rank_world() == (rank, world, rank)
    uid = bytes(range(128)) if rank == 0 else None
    got = d.broadcast_bytes(uid, 128, src=0)
    # sharded 29-slot reduction == unsharded (what the per-iteration all-reduce computes)
    src, tgt, nrm, _ = make_icp_pair(4000, seed=7)
    idx, _, _ = oracle.hybrid_search(tgt, src, 0.05, 1)
    corr = idx[:, 0].astype(np.int64)
    b, e = d.shard_range(len(src), rank, world)
    # the full correspondence array indexes the full target; the shard owns source rows [b, e)
    part = oracle.pose_p2plane_sums(src[b:e], tgt, nrm, corr[b:e])["sums64"]
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)
    full = oracle.pose_p2plane_sums(src, tgt, nrm, corr)["sums64"]
    # the same for ColoredICP (o3db_icp_create_colored takes the same communicator): the target side
    # (colours, colour gradients) is replicated, the source colours are sharded with the source points
    from tests.synth import make_colors
    sc, tc = make_colors(src, 1), make_colors(tgt, 1)
    grad = oracle.estimate_color_gradients(tgt, nrm, tc, 0.08, 30)
    cpart = oracle.pose_colored_sums(src[b:e], sc[b:e], tgt, nrm, tc, grad, corr[b:e], 0.968)["sums64"]
    ct = torch.from_numpy(cpart.copy())
    dist.all_reduce(ct)
    cfull = oracle.pose_colored_sums(src, sc, tgt, nrm, tc, grad, corr, 0.968)["sums64"]
    colored_ok = np.allclose(ct.numpy(), cfull, rtol=1e-12, atol=1e-12)
    q.put((rank, got == bytes(range(128)), colored_ok and np.allclose(t.numpy(), full, rtol=1e-12, atol=1e-12),
           d.reduce_max(float(rank + 1)), d.reduce_sum(float(e - b)), len(src)))
    dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, uid_ok, sums_ok, mx, total, n in out:
        assert uid_ok and sums_ok
        assert mx == 2.0 and total == float(n)
