"""GPU parity tests of the TSDF path: CUDA (through the C ABI / the reference-facing
python surface) vs the CPU oracle.  Block keys, hash values and voxel-block sets
must be bit-exact; TSDF / weight / colour values are compared exactly as well
(the kernels evaluate the reference's f32 expressions without FMA contraction),
with the north_star tolerance (1e-5 relative) as the documented bound.
"""
import numpy as np
import pytest
import torch

import oracle
from tests.synth import PRIMESENSE_K, camera_pose, render_depth

pytestmark = pytest.mark.gpu

VOXEL, RES, TRUNC_MULT = 0.008, 16, 8.0   # BASELINE config 3; slam::Model defaults (Model.h)
SCALE, DMAX = 1000.0, 3.0


@pytest.fixture(scope="module")
def o3d():
    import open3d_b200
    assert torch.cuda.is_available()
    return open3d_b200


def _frame(i, color=False, f32=False):
    T = camera_pose(i)
    out = render_depth(T, with_color=color)
    depth, col = (out if color else (out, None))
    depth = depth.numpy()
    if f32:
        depth = depth.astype(np.float32)
        col = None if col is None else (col.numpy().astype(np.float32) / 255.0)
    elif col is not None:
        col = col.numpy()
    return T, oracle.inverse_transformation(T), depth, col


def _sorted_keys(k):
    k = np.asarray(k, np.int32).reshape(-1, 3)
    return k[np.lexsort((k[:, 2], k[:, 1], k[:, 0]))]


def test_key_hash_is_minivec_hash_bit_exact():
    from open3d_b200 import _lib as L
    rng = np.random.default_rng(0)
    keys = rng.integers(-2**31, 2**31 - 1, (4096, 3), dtype=np.int64).astype(np.int32)
    keys[:4] = [[0, 0, 0], [-1, 3, 2], [1, 2, 3], [2**31 - 1, -2**31, -7]]
    k = torch.from_numpy(keys).cuda()
    out = torch.empty(len(keys), dtype=torch.int64, device="cuda")
    L.check(L.lib.o3db_hash_keys(k.data_ptr(), len(keys), out.data_ptr(), 0))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), oracle.minivec_hash(keys))


def test_hashmap_reference_kat_and_semantics(kats, o3d):
    k = kats["vbg_indexing"]   # cpp/tests/t/geometry/VoxelBlockGrid.cpp:199-219
    vbg = o3d.t.geometry.VoxelBlockGrid(voxel_size=3.0 / 512, block_resolution=2, block_count=10)
    hm = vbg.hashmap()
    keys = np.array(k["keys"], np.int32)
    buf, masks = hm.activate(keys)
    assert int(masks.sum()) == k["expected_unique"] == hm.size()
    buf = buf.cpu().numpy()
    table = hm.key_tensor().cpu().numpy()
    assert (table[buf] == keys).all()                      # every input learned the slot of its key
    assert len(set(buf.tolist())) == 3 and set(buf.tolist()) == {0, 1, 2}
    m = masks.cpu().numpy()
    for key in {tuple(r) for r in keys.tolist()}:          # exactly one inserter per unique key
        assert m[[tuple(r) == key for r in keys.tolist()]].sum() == 1
    buf2, masks2 = hm.activate(keys)                       # re-activation inserts nothing
    assert not masks2.any() and np.array_equal(buf2.cpu().numpy(), buf) and hm.size() == 3
    fb, fm = hm.find(np.array([[1, 2, 3], [9, 9, 9]], np.int32))
    assert fm.tolist() == [True, False] and fb[1].item() == -1 and (table[fb[0].item()] == [1, 2, 3]).all()
    assert sorted(hm.active_buf_indices().tolist()) == [0, 1, 2]
    assert vbg.attribute("tsdf").shape == (10, 2, 2, 2, 1) and vbg.attribute("color").shape == (10, 2, 2, 2, 3)


def test_hashmap_growth_from_tiny_capacity(o3d):
    """cpp/tests/core/HashMap.cpp:262-305: capacity 2 -> 20000 distinct keys, with values kept."""
    vbg = o3d.t.geometry.VoxelBlockGrid(voxel_size=0.01, block_resolution=2, block_count=2)
    hm = vbg.hashmap()
    rng = np.random.default_rng(1)
    keys = np.unique(rng.integers(-500, 500, (30000, 3)).astype(np.int32), axis=0)[:20000]
    first = keys[:2]
    buf0, _ = hm.activate(first)
    vbg.attribute("tsdf")[buf0.long(), 0, 0, 0, 0] = torch.tensor([1.5, -2.5], device="cuda")
    dup = np.concatenate([keys, keys[::7]])                 # duplicates inside one call
    buf, masks = hm.activate(dup)
    assert hm.size() == 20000 and hm.capacity() >= 20000
    assert int(masks.sum()) == 20000 - 2
    table = hm.key_tensor().cpu().numpy()
    assert (table[buf.cpu().numpy()] == dup).all()
    assert np.array_equal(_sorted_keys(table[:20000]), _sorted_keys(keys))
    fb, fm = hm.find(first)
    assert fm.all() and np.array_equal(fb.cpu().numpy(), buf0.cpu().numpy())   # slots survive growth
    assert vbg.attribute("tsdf")[fb.long(), 0, 0, 0, 0].tolist() == [1.5, -2.5]


@pytest.mark.parametrize("frame_id,f32", [(0, False), (137, False), (500, True)])
def test_depth_touch_block_set_bit_exact(o3d, frame_id, f32):
    T, E, depth, _ = _frame(frame_id, f32=f32)
    vbg = o3d.t.geometry.VoxelBlockGrid(voxel_size=VOXEL, block_resolution=RES, block_count=1000)
    got = vbg.compute_unique_block_coordinates(torch.from_numpy(depth), PRIMESENSE_K, E, SCALE, DMAX, TRUNC_MULT)
    want = oracle.depth_touch(depth, PRIMESENSE_K, E, RES, VOXEL, VOXEL * TRUNC_MULT, SCALE, DMAX, 4)
    got = got.cpu().numpy()
    assert len(got) == len(want) > 100
    assert len(np.unique(got, axis=0)) == len(got)          # unique
    assert np.array_equal(_sorted_keys(got), want)          # identical set, bit for bit


def test_depth_touch_no_block_is_an_error(o3d):
    vbg = o3d.t.geometry.VoxelBlockGrid(voxel_size=VOXEL, block_resolution=RES, block_count=100)
    with pytest.raises(RuntimeError, match="No block is touched"):
        vbg.compute_unique_block_coordinates(torch.zeros((480, 640), dtype=torch.uint16), PRIMESENSE_K, np.eye(4))


def _oracle_volume(cap, color):
    return (np.zeros((cap, 3), np.int32), np.zeros((cap, RES ** 3), np.float32),
            np.zeros((cap, RES ** 3), np.uint16), np.zeros((cap, RES ** 3, 3), np.uint16) if color else None)


def _compare_volumes(vbg, okeys, otsdf, owt, ocol, osize):
    hm = vbg.hashmap()
    assert hm.size() == osize
    gkeys = hm.key_tensor().cpu().numpy()[:osize]
    assert np.array_equal(_sorted_keys(gkeys), _sorted_keys(okeys[:osize]))   # same block set
    # slot order is implementation-defined on both sides: align by key
    lut = {tuple(k): i for i, k in enumerate(okeys[:osize].tolist())}
    perm = np.array([lut[tuple(k)] for k in gkeys.tolist()])
    gt = vbg.attribute("tsdf").cpu().numpy().reshape(-1, RES ** 3)[:osize]
    gw = vbg.attribute("weight").cpu().numpy().reshape(-1, RES ** 3)[:osize]
    assert np.array_equal(gw, owt[perm])
    np.testing.assert_allclose(gt, otsdf[perm], rtol=1e-5, atol=1e-7)           # north_star bound
    assert np.array_equal(gt.view(np.uint32), otsdf[perm].view(np.uint32))      # and in fact bit-exact
    if ocol is not None:
        gc = vbg.attribute("color").cpu().numpy().reshape(-1, RES ** 3, 3)[:osize]
        assert np.array_equal(gc, ocol[perm])
    return int((gw > 0).sum())


@pytest.mark.parametrize("color,f32", [(False, False), (True, False), (True, True)])
def test_integrate_unfused_api_vs_oracle(o3d, color, f32):
    """GetUniqueBlockCoordinates + Integrate (VoxelBlockGrid.cpp:212-326) over 3 frames."""
    cap = 6000
    vbg = o3d.t.geometry.VoxelBlockGrid(voxel_size=VOXEL, block_resolution=RES, block_count=cap)
    okeys, otsdf, owt, ocol = _oracle_volume(cap, color)
    osize = 0
    for fid in (0, 40, 80):
        T, E, depth, col = _frame(fid, color=color, f32=f32)
        bc = vbg.compute_unique_block_coordinates(torch.from_numpy(depth), PRIMESENSE_K, E, SCALE, DMAX, TRUNC_MULT)
        vbg.integrate(bc, torch.from_numpy(depth), None if col is None else torch.from_numpy(col), PRIMESENSE_K,
                      PRIMESENSE_K, E, SCALE, DMAX, TRUNC_MULT)
        want = oracle.depth_touch(depth, PRIMESENSE_K, E, RES, VOXEL, VOXEL * TRUNC_MULT, SCALE, DMAX, 4)
        bi, _, osize, rc = oracle.hashmap_activate(okeys, osize, want)
        assert rc == 0
        oracle.tsdf_integrate(depth, col, bi, okeys, otsdf, owt, ocol, PRIMESENSE_K, PRIMESENSE_K, E, RES, VOXEL,
                              VOXEL * TRUNC_MULT, SCALE, DMAX)
    updated = _compare_volumes(vbg, okeys, otsdf, owt, ocol, osize)
    assert updated > 100000


@pytest.mark.parametrize("color,host", [(False, False), (True, False), (True, True)])
def test_model_integrate_fused_vs_oracle(o3d, color, host):
    """slam::Model::Integrate (Model.cpp:91-106), fused pipeline, 6 frames incl. revisits."""
    slam = o3d.t.pipelines.slam
    cap = 8000
    model = slam.Model(VOXEL, RES, cap)
    okeys, otsdf, owt, ocol = _oracle_volume(cap, True)
    osize = 0
    for n, fid in enumerate((0, 3, 6, 200, 203, 0)):
        T, E, depth, col = _frame(fid, color=color)
        frame = slam.Frame(480, 640, PRIMESENSE_K)
        d = torch.from_numpy(depth)
        c = None if col is None else torch.from_numpy(col)
        frame.set_data("depth", d if host else d.cuda())
        if c is not None:
            frame.set_data("color", c if host else c.cuda())
        model.update_frame_pose(n, T)
        model.integrate(frame, SCALE, DMAX, TRUNC_MULT)
        want = oracle.depth_touch(depth, PRIMESENSE_K, E, RES, VOXEL, VOXEL * TRUNC_MULT, SCALE, DMAX, 4)
        got = model.frustum_block_coords.cpu().numpy()
        assert np.array_equal(_sorted_keys(got), want)      # Model::frustum_block_coords_
        bi, _, osize, rc = oracle.hashmap_activate(okeys, osize, want)
        assert rc == 0
        oracle.tsdf_integrate(depth, col, bi, okeys, otsdf, owt, ocol if color else None, PRIMESENSE_K, PRIMESENSE_K,
                              E, RES, VOXEL, VOXEL * TRUNC_MULT, SCALE, DMAX)
    _compare_volumes(model.voxel_grid, okeys, otsdf, owt, ocol if color else None, osize)
    if not color:   # colour buffer untouched by depth-only integration
        assert int(model.voxel_grid.attribute("color").view(torch.int16).count_nonzero()) == 0


def test_fused_path_grows_transparently(o3d):
    slam = o3d.t.pipelines.slam
    model = slam.Model(VOXEL, RES, 64)       # far too small: must grow, never drop blocks
    total = set()
    for n, fid in enumerate(range(0, 400, 50)):
        T, E, depth, _ = _frame(fid)
        frame = slam.Frame(480, 640, PRIMESENSE_K)
        frame.set_data("depth", torch.from_numpy(depth).cuda())
        model.update_frame_pose(n, T)
        model.integrate(frame, SCALE, DMAX, TRUNC_MULT)
        total |= {tuple(k) for k in oracle.depth_touch(depth, PRIMESENSE_K, E, RES, VOXEL, VOXEL * TRUNC_MULT,
                                                      SCALE, DMAX, 4).tolist()}
    hm = model.get_hashmap()
    assert hm.size() == len(total) and hm.capacity() >= len(total)
    keys = hm.key_tensor().cpu().numpy()[: hm.size()]
    assert {tuple(k) for k in keys.tolist()} == total


def test_frame_that_outgrows_the_map_is_dropped_whole_and_recoverable():
    """ADVICE r1 (tsdf.cu overflow): the fused path sizes the map ahead of need from what earlier frames added; a camera
    jump that adds far more blocks than that (here ~24 k new blocks after frames that added < 1 k) cannot fit.  The
    reference would grow inside HashMap::Activate (HashMap.cpp:166-181); the async path instead drops that frame AND
    every later one as a whole (every CTA takes the same decision, provisional table entries are released), reports
    which frame it was, and after o3db_vbg_reserve the caller resubmits from there.  The recovered volume must equal
    the oracle's, bit for bit — i.e. the dropped frames left no trace."""
    import ctypes as C
    import re
    from open3d_b200 import _lib as L
    voxel, cap0 = 0.00125, 4000
    stream = int(torch.cuda.current_stream().cuda_stream)
    v = C.c_void_p()
    L.check(L.lib.o3db_vbg_create(voxel, RES, cap0, 0, stream, C.byref(v)))
    K = np.ascontiguousarray(PRIMESENSE_K)
    seq = [(250, 0.95), (250, 0.95), (251, 0.95), (125, DMAX), (126, DMAX), (250, 0.95), (127, DMAX)]
    frames = []
    for fid, dmax in seq:
        T = camera_pose(fid)
        frames.append((render_depth(T, device="cuda").contiguous(), np.ascontiguousarray(oracle.inverse_transformation(T)), dmax))

    def submit(i):
        d, E, dmax = frames[i]
        return L.lib.o3db_vbg_integrate_frame(v, d.data_ptr(), L.DEPTH_U16, None, 0, 480, 640, L.dptr(K), L.dptr(E), SCALE, dmax,
                                              TRUNC_MULT, stream)
    i, first_dropped = 0, None
    while i < len(frames):
        rc = submit(i)
        if rc == 0:
            i += 1
            continue
        assert rc == L.ERR_CAPACITY and first_dropped is None, (rc, L.last_error())
        m = re.search(r"fused frame #(\d+) needed (\d+) blocks", L.last_error())
        assert m, L.last_error()
        first_dropped, needed = int(m.group(1)), int(m.group(2))
        assert first_dropped == 3 and needed > 20000           # the jump
        L.check(L.lib.o3db_vbg_reserve(v, 2 * needed, stream))
        i = first_dropped                                       # resubmit from the dropped frame
    if first_dropped is None:                                   # the last frames were dropped silently so far: ask
        size = L.lib.o3db_vbg_size(v, stream)
        assert size == L.ERR_CAPACITY, size
    assert first_dropped == 3
    size = L.lib.o3db_vbg_size(v, stream)
    cap = 40000
    okeys, otsdf, owt = np.zeros((cap, 3), np.int32), np.zeros((cap, RES ** 3), np.float32), np.zeros((cap, RES ** 3), np.uint16)
    osize = 0
    for d, E, dmax in frames:
        dh = d.cpu().numpy()
        want = oracle.depth_touch(dh, PRIMESENSE_K, E, RES, voxel, voxel * TRUNC_MULT, SCALE, dmax, 4)
        bi, _, osize, rc = oracle.hashmap_activate(okeys, osize, want)
        assert rc == 0
        oracle.tsdf_integrate(dh, None, bi, okeys, otsdf, owt, None, PRIMESENSE_K, PRIMESENSE_K, E, RES, voxel,
                              voxel * TRUNC_MULT, SCALE, dmax)
    assert size == osize
    from open3d_b200.t import _libshim as _s
    capn = int(L.lib.o3db_vbg_capacity(v))
    torch.cuda.synchronize()
    gkeys = _s.device_view(L.lib.o3db_vbg_key_buffer(v), (capn, 3), torch.int32)[:size]
    gt = _s.device_view(L.lib.o3db_vbg_tsdf_buffer(v), (capn, RES ** 3), torch.float32)[:size]
    gw = _s.device_view(L.lib.o3db_vbg_weight_buffer(v), (capn, RES ** 3), torch.uint16)[:size]
    gkeys, gt, gw = gkeys.cpu().numpy(), gt.cpu().numpy(), gw.cpu().numpy()
    assert np.array_equal(_sorted_keys(gkeys), _sorted_keys(okeys[:osize]))
    lut = {tuple(k): j for j, k in enumerate(okeys[:osize].tolist())}
    perm = np.array([lut[tuple(k)] for k in gkeys.tolist()])
    assert np.array_equal(gw, owt[perm]) and np.array_equal(gt.view(np.uint32), otsdf[perm].view(np.uint32))
    L.lib.o3db_vbg_destroy(v)


def test_voxel_block_grid_grows_like_hashmap_activate_after_a_dropped_frame(o3d):
    """The same camera jump through the Python mirror: VoxelBlockGrid.integrate_frame must behave like the reference,
    whose HashMap::Activate grows the map on demand (HashMap.cpp:166-181) — no error reaches the caller, and the
    volume equals the oracle's bit for bit (the dropped frames were resubmitted in order, once)."""
    voxel, cap0 = 0.00125, 4000
    vbg = o3d.t.geometry.VoxelBlockGrid(("tsdf", "weight"), (torch.float32, torch.uint16), ((1,), (1,)), voxel, RES, cap0)
    seq = [(250, 0.95), (250, 0.95), (251, 0.95), (125, DMAX), (126, DMAX), (250, 0.95), (127, DMAX)]
    cap = 40000
    okeys, otsdf, owt = np.zeros((cap, 3), np.int32), np.zeros((cap, RES ** 3), np.float32), np.zeros((cap, RES ** 3), np.uint16)
    osize = 0
    for fid, dmax in seq:
        T = camera_pose(fid)
        d = render_depth(T, device="cuda").contiguous()
        E = np.ascontiguousarray(oracle.inverse_transformation(T))
        vbg.integrate_frame(d, None, PRIMESENSE_K, E, SCALE, dmax, TRUNC_MULT)
        dh = d.cpu().numpy()
        want = oracle.depth_touch(dh, PRIMESENSE_K, E, RES, voxel, voxel * TRUNC_MULT, SCALE, dmax, 4)
        bi, _, osize, rc = oracle.hashmap_activate(okeys, osize, want)
        assert rc == 0
        oracle.tsdf_integrate(dh, None, bi, okeys, otsdf, owt, None, PRIMESENSE_K, PRIMESENSE_K, E, RES, voxel,
                              voxel * TRUNC_MULT, SCALE, dmax)
    hm = vbg.hashmap()
    assert hm.size() == osize                      # also the call that surfaces a drop of the very last frames
    assert vbg._replay.recoveries >= 1 and hm.capacity() > cap0
    gkeys = hm.key_tensor().cpu().numpy()[:osize]
    lut = {tuple(k): j for j, k in enumerate(okeys[:osize].tolist())}
    perm = np.array([lut[tuple(k)] for k in gkeys.tolist()])
    gt = vbg.attribute("tsdf").cpu().numpy().reshape(-1, RES ** 3)[:osize]
    gw = vbg.attribute("weight").cpu().numpy().reshape(-1, RES ** 3)[:osize]
    assert np.array_equal(gw, owt[perm]) and np.array_equal(gt.view(np.uint32), otsdf[perm].view(np.uint32))
    # auto_grow = False hands the documented error to the caller instead
    vbg2 = o3d.t.geometry.VoxelBlockGrid(("tsdf", "weight"), (torch.float32, torch.uint16), ((1,), (1,)), voxel, RES, cap0)
    vbg2.auto_grow = False
    with pytest.raises(RuntimeError, match=r"fused frame #3 needed \d+ blocks"):
        for fid, dmax in seq:
            T = camera_pose(fid)
            vbg2.integrate_frame(render_depth(T, device="cuda").contiguous(), None, PRIMESENSE_K,
                                 np.ascontiguousarray(oracle.inverse_transformation(T)), SCALE, dmax, TRUNC_MULT)
        vbg2.hashmap().size()


def test_generic_block_resolution_vs_oracle(o3d):
    """res = 8 goes through the generic (non-vectorised) integrate path."""
    res, cap = 8, 20000
    vbg = o3d.t.geometry.VoxelBlockGrid(voxel_size=VOXEL, block_resolution=res, block_count=cap)
    T, E, depth, _ = _frame(10)
    bc = vbg.compute_unique_block_coordinates(torch.from_numpy(depth), PRIMESENSE_K, E, SCALE, DMAX, TRUNC_MULT)
    vbg.integrate(bc, torch.from_numpy(depth), None, PRIMESENSE_K, PRIMESENSE_K, E, SCALE, DMAX, TRUNC_MULT)
    want = oracle.depth_touch(depth, PRIMESENSE_K, E, res, VOXEL, VOXEL * TRUNC_MULT, SCALE, DMAX, 4)
    okeys = np.zeros((cap, 3), np.int32)
    otsdf = np.zeros((cap, res ** 3), np.float32)
    owt = np.zeros((cap, res ** 3), np.uint16)
    bi, _, osize, _ = oracle.hashmap_activate(okeys, 0, want)
    oracle.tsdf_integrate(depth, None, bi, okeys, otsdf, owt, None, PRIMESENSE_K, PRIMESENSE_K, E, res, VOXEL,
                          VOXEL * TRUNC_MULT, SCALE, DMAX)
    hm = vbg.hashmap()
    assert hm.size() == osize
    gkeys = hm.key_tensor().cpu().numpy()[:osize]
    lut = {tuple(k): i for i, k in enumerate(okeys[:osize].tolist())}
    perm = np.array([lut[tuple(k)] for k in gkeys.tolist()])
    gt = vbg.attribute("tsdf").cpu().numpy().reshape(-1, res ** 3)[:osize]
    gw = vbg.attribute("weight").cpu().numpy().reshape(-1, res ** 3)[:osize]
    assert np.array_equal(gw, owt[perm]) and np.array_equal(gt.view(np.uint32), otsdf[perm].view(np.uint32))


def test_idempotent_weight_property_full_sequence(o3d):
    """Size-independent property at BASELINE scale (50 frames of the config-3 trajectory):
    integrating the same frame k times gives weight == k on every observed voxel and leaves
    tsdf unchanged (running mean of identical samples), up to f32 rounding."""
    slam = o3d.t.pipelines.slam
    model = slam.Model(VOXEL, RES, 40000)
    T, E, depth, _ = _frame(25)
    frame = slam.Frame(480, 640, PRIMESENSE_K)
    frame.set_data("depth", torch.from_numpy(depth).cuda())
    model.update_frame_pose(0, T)
    model.integrate(frame)
    n = model.get_hashmap().size()
    t1 = model.voxel_grid.attribute("tsdf")[:n].clone()
    w1 = model.voxel_grid.attribute("weight")[:n].clone()
    for _ in range(4):
        model.integrate(frame)
    assert model.get_hashmap().size() == n
    w5 = model.voxel_grid.attribute("weight")[:n]
    assert torch.equal(w5.int(), w1.int() * 5) and set(w1.int().unique().tolist()) == {0, 1}
    torch.testing.assert_close(model.voxel_grid.attribute("tsdf")[:n], t1, rtol=0, atol=2e-6)


def test_integrate_sequence_equals_per_frame_calls(o3d):
    """o3db_vbg_integrate_sequence is the same kernels in the same order: bit-identical volume."""
    import ctypes as C
    from open3d_b200 import _lib as L
    frames = [_frame(f, color=True) for f in (0, 5, 10, 15)]
    K = np.ascontiguousarray(PRIMESENSE_K)

    def volume(use_sequence, host):
        v = C.c_void_p()
        L.check(L.lib.o3db_vbg_create(VOXEL, RES, 6000, 1, 0, C.byref(v)))
        deps = [torch.from_numpy(f[2]) for f in frames]
        cols = [torch.from_numpy(f[3]) for f in frames]
        if host:
            deps, cols = [d.pin_memory() for d in deps], [c.pin_memory() for c in cols]
        else:
            deps, cols = [d.cuda() for d in deps], [c.cuda() for c in cols]
        exts = np.ascontiguousarray(np.stack([f[1] for f in frames]).reshape(-1, 16))
        if use_sequence:
            P = C.c_void_p * len(frames)
            L.check(L.lib.o3db_vbg_integrate_sequence(v, len(frames), P(*[d.data_ptr() for d in deps]), L.DEPTH_U16,
                                                      P(*[c.data_ptr() for c in cols]), L.COLOR_U8, 480, 640, L.dptr(K),
                                                      L.dptr(exts), SCALE, DMAX, TRUNC_MULT, int(host), 0))
        else:
            fn = L.lib.o3db_vbg_integrate_frame_host if host else L.lib.o3db_vbg_integrate_frame
            for i in range(len(frames)):
                L.check(fn(v, deps[i].data_ptr(), L.DEPTH_U16, cols[i].data_ptr(), L.COLOR_U8, 480, 640, L.dptr(K),
                           L.dptr(np.ascontiguousarray(exts[i])), SCALE, DMAX, TRUNC_MULT, 0))
        n = int(L.check(L.lib.o3db_vbg_size(v, 0)))
        from open3d_b200.t._libshim import device_view
        keys = device_view(L.lib.o3db_vbg_key_buffer(v), (n, 3), torch.int32).cpu().numpy().copy()
        tsdf = device_view(L.lib.o3db_vbg_tsdf_buffer(v), (n, 4096), torch.float32).cpu().numpy().copy()
        wt = device_view(L.lib.o3db_vbg_weight_buffer(v), (n, 4096), torch.uint16).cpu().numpy().copy()
        col = device_view(L.lib.o3db_vbg_color_buffer(v), (n, 4096 * 3), torch.uint16).cpu().numpy().copy()
        L.lib.o3db_vbg_destroy(v)
        order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
        return keys[order], tsdf[order], wt[order], col[order]

    ref = volume(False, False)
    for use_seq, host in ((True, False), (True, True), (False, True)):
        got = volume(use_seq, host)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b)


# ------------------------------------------- EstimateRange + RayCast

DMIN = 0.1
ALL_ATTRS = ("depth", "vertex", "color", "normal", "index", "mask", "interp_ratio", "interp_ratio_dx",
             "interp_ratio_dy", "interp_ratio_dz")


@pytest.fixture(scope="module")
def fused_pair(o3d):
    """The same 5 colour frames fused by slam.Model (CUDA) and by the oracle."""
    slam = o3d.t.pipelines.slam
    cap = 9000
    model = slam.Model(VOXEL, RES, cap)
    okeys, otsdf, owt, ocol = _oracle_volume(cap, True)
    osize = 0
    frames = (0, 2, 4, 6, 8)
    for n, fid in enumerate(frames):
        T, E, depth, col = _frame(fid, color=True)
        frame = slam.Frame(480, 640, PRIMESENSE_K)
        frame.set_data("depth", torch.from_numpy(depth).cuda())
        frame.set_data("color", torch.from_numpy(col).cuda())
        model.update_frame_pose(n, T)
        model.integrate(frame, SCALE, DMAX, TRUNC_MULT)
        want = oracle.depth_touch(depth, PRIMESENSE_K, E, RES, VOXEL, VOXEL * TRUNC_MULT, SCALE, DMAX, 4)
        bi, _, osize, rc = oracle.hashmap_activate(okeys, osize, want)
        assert rc == 0
        oracle.tsdf_integrate(depth, col, bi, okeys, otsdf, owt, ocol, PRIMESENSE_K, PRIMESENSE_K, E, RES, VOXEL,
                              VOXEL * TRUNC_MULT, SCALE, DMAX)
    _compare_volumes(model.voxel_grid, okeys, otsdf, owt, ocol, osize)
    return dict(model=model, okeys=okeys, otsdf=otsdf, owt=owt, ocol=ocol, osize=osize, frustum=want,
                last_pose=camera_pose(frames[-1]))


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


@pytest.mark.parametrize("down", [8, 4])
def test_estimate_range_bit_exact_vs_oracle(fused_pair, down):
    vbg = fused_pair["model"].voxel_grid
    T = fused_pair["last_pose"]
    E = oracle.inverse_transformation(T)
    want = oracle.estimate_range(fused_pair["frustum"], PRIMESENSE_K, E, 480, 640, down, RES, VOXEL, DMIN, DMAX)
    # explicit block coordinates (VoxelBlockGrid::RayCast's argument) ...
    bc = torch.from_numpy(fused_pair["frustum"]).cuda()
    got = vbg.ray_cast(bc, PRIMESENSE_K, E, 640, 480, ("depth",), SCALE, DMIN, DMAX, 1.0, TRUNC_MULT, down)["range"]
    assert got.shape == want.shape
    assert np.array_equal(_bits(got.cpu().numpy()), _bits(want))
    # ... and the frustum of the last fused frame taken on the device (slam::Model::frustum_block_coords_)
    got2 = vbg.ray_cast(None, PRIMESENSE_K, E, 640, 480, ("depth",), SCALE, DMIN, DMAX, 1.0, TRUNC_MULT, down)["range"]
    assert np.array_equal(_bits(got2.cpu().numpy()), _bits(want))
    # a partially covered view, and a cloud of blocks all around the camera: corners behind the camera
    # plane are skipped (:398), corners just in front of it project to |u|, |v| ~ 1e9 (saturating casts)
    allk = vbg.hashmap().key_tensor()[: fused_pair["osize"]].contiguous()
    rng_keys = np.random.default_rng(7).integers(-14, 15, (4000, 3)).astype(np.int32) + np.int32([8, 0, 11])
    for T2, keys in ((camera_pose(100), allk.cpu().numpy()), (camera_pose(30), rng_keys)):
        E2 = oracle.inverse_transformation(T2)
        want2 = oracle.estimate_range(keys, PRIMESENSE_K, E2, 480, 640, down, RES, VOXEL, DMIN, DMAX)
        got3 = vbg.ray_cast(torch.from_numpy(keys).cuda(), PRIMESENSE_K, E2, 640, 480, ("depth",), SCALE, DMIN, DMAX,
                            1.0, TRUNC_MULT, down)["range"]
        assert np.array_equal(_bits(got3.cpu().numpy()), _bits(want2))
        covered = want2[..., 0] < want2[..., 1]
        assert covered.any() and (keys is rng_keys or not covered.all())


@pytest.mark.parametrize("fid,threshold", [(8, 3.0), (5, 1.0), (40, 1.0)])
def test_ray_cast_bit_exact_vs_oracle(fused_pair, fid, threshold):
    """All ten renderings of VoxelBlockGrid::RayCast from an integrated pose, an in-between pose and a pose
    that sees the edge of the fused region.  The march is evaluated without FMA contraction: identical."""
    vbg = fused_pair["model"].voxel_grid
    osize = fused_pair["osize"]
    T = camera_pose(fid)
    E = oracle.inverse_transformation(T)
    gkeys = vbg.hashmap().key_tensor()[:osize].contiguous()
    res = vbg.ray_cast(gkeys, PRIMESENSE_K, E, 640, 480, ALL_ATTRS, SCALE, DMIN, DMAX, threshold, TRUNC_MULT, 8)
    rng = oracle.estimate_range(gkeys.cpu().numpy(), PRIMESENSE_K, E, 480, 640, 8, RES, VOXEL, DMIN, DMAX)
    assert np.array_equal(_bits(res["range"].cpu().numpy()), _bits(rng))
    ref = oracle.ray_cast(fused_pair["okeys"], osize, fused_pair["otsdf"], fused_pair["owt"], fused_pair["ocol"],
                          rng, PRIMESENSE_K, E, 480, 640, ALL_ATTRS, RES, VOXEL, SCALE, DMIN, DMAX, threshold,
                          TRUNC_MULT, 8)
    hit = ref["depth"][..., 0] > 0
    assert hit.mean() > (0.5 if fid != 40 else 0.02)
    for name in ("depth", "vertex", "normal", "color", "interp_ratio", "interp_ratio_dx", "interp_ratio_dy",
                 "interp_ratio_dz"):
        got = res[name].cpu().numpy()
        np.testing.assert_allclose(got, ref[name], rtol=1e-5, atol=1e-6, err_msg=name)     # north_star bound
        assert np.array_equal(_bits(got), _bits(ref[name])), name                           # in fact identical
    gmask = res["mask"].cpu().numpy()
    assert np.array_equal(gmask, ref["mask"])
    # voxel indices: slot order differs between the two hash maps -> compare (block key, voxel-in-block)
    gi, oi = res["index"].cpu().numpy(), ref["index"]
    r3 = RES ** 3
    gk = gkeys.cpu().numpy()[gi[gmask] // r3]
    ok = fused_pair["okeys"][oi[gmask] // r3]
    assert np.array_equal(gk, ok) and np.array_equal(gi[gmask] % r3, oi[gmask] % r3)
    assert not gi[~gmask].any()
    if fid == 8:   # against the analytic scene: the O(voxel) staircase of the nearest-voxel march
        truth = render_depth(T).numpy().astype(np.float64)
        d = res["depth"].cpu().numpy()[..., 0].astype(np.float64)
        both = (d > 0) & (truth > 0)
        assert np.median(np.abs(d - truth)[both]) < 0.75 * VOXEL * SCALE


def test_ray_cast_attribute_subsets_and_errors(o3d, fused_pair):
    vbg = fused_pair["model"].voxel_grid
    T = fused_pair["last_pose"]
    E = oracle.inverse_transformation(T)
    full = vbg.ray_cast(None, PRIMESENSE_K, E, 640, 480, ("depth", "vertex", "normal", "color"), SCALE, DMIN, DMAX,
                        1.0, TRUNC_MULT)
    only = vbg.ray_cast(None, PRIMESENSE_K, E, 640, 480, ("depth", "vertex"), SCALE, DMIN, DMAX, 1.0, TRUNC_MULT)
    assert set(only) == {"depth", "vertex", "range"}                       # the no-neighbour fast path (:1000)
    assert torch.equal(only["depth"], full["depth"]) and torch.equal(only["vertex"], full["vertex"])
    assert full["depth"].shape == (480, 640, 1) and full["color"].shape == (480, 640, 3)
    assert not (vbg.ray_cast(None, PRIMESENSE_K, E, 640, 480, ("depth",), SCALE, DMIN, DMAX, 50.0)["depth"] > 0).any()
    with pytest.raises(RuntimeError, match="Unsupported attribute"):
        vbg.ray_cast(None, PRIMESENSE_K, E, 640, 480, ("albedo",))
    # odd image sizes (partial tiles, range cells cut by integer division)
    small = vbg.ray_cast(None, PRIMESENSE_K, E, 333, 250, ("depth", "normal"), SCALE, DMIN, DMAX, 1.0, TRUNC_MULT)
    rng = oracle.estimate_range(fused_pair["frustum"], PRIMESENSE_K, E, 250, 333, 8, RES, VOXEL, DMIN, DMAX)
    ref = oracle.ray_cast(fused_pair["okeys"], fused_pair["osize"], fused_pair["otsdf"], fused_pair["owt"], None, rng,
                          PRIMESENSE_K, E, 250, 333, ("depth", "normal"), RES, VOXEL, SCALE, DMIN, DMAX, 1.0,
                          TRUNC_MULT, 8)
    assert np.array_equal(_bits(small["range"].cpu().numpy()), _bits(rng))
    assert np.array_equal(_bits(small["depth"].cpu().numpy()), _bits(ref["depth"]))
    assert np.array_equal(_bits(small["normal"].cpu().numpy()), _bits(ref["normal"]))
    # a grid without colour renders zeros for "color"
    plain = o3d.t.geometry.VoxelBlockGrid(("tsdf", "weight"), (torch.float32, torch.uint16), ((1,), (1,)), VOXEL, RES, 4000)
    T0, E0, depth, _ = _frame(0)
    bc = plain.compute_unique_block_coordinates(torch.from_numpy(depth), PRIMESENSE_K, E0, SCALE, DMAX, TRUNC_MULT)
    plain.integrate(bc, torch.from_numpy(depth), None, PRIMESENSE_K, PRIMESENSE_K, E0, SCALE, DMAX, TRUNC_MULT)
    out = plain.ray_cast(bc, PRIMESENSE_K, E0, 640, 480, ("depth", "color"), SCALE, DMIN, DMAX, 1.0, TRUNC_MULT)
    assert (out["depth"] > 0).float().mean() > 0.8 and not out["color"].any()


def test_model_synthesize_model_frame(o3d, fused_pair):
    """slam::Model::SynthesizeModelFrame (Model.cpp:38-66)."""
    slam = o3d.t.pipelines.slam
    model = fused_pair["model"]
    rc = slam.Frame(480, 640, PRIMESENSE_K)
    model.synthesize_model_frame(rc, SCALE, DMIN, DMAX, TRUNC_MULT, True)     # weight_threshold = min(frame_id, 3)
    d, c = rc.get_data("depth"), rc.get_data("color")
    assert d.shape == (480, 640, 1) and c.shape == (480, 640, 3) and d.is_cuda
    E = oracle.inverse_transformation(model.get_current_frame_pose())
    rng = oracle.estimate_range(fused_pair["frustum"], PRIMESENSE_K, E, 480, 640, 8, RES, VOXEL, DMIN, DMAX)
    ref = oracle.ray_cast(fused_pair["okeys"], fused_pair["osize"], fused_pair["otsdf"], fused_pair["owt"],
                          fused_pair["ocol"], rng, PRIMESENSE_K, E, 480, 640, ("depth", "color"), RES, VOXEL, SCALE,
                          DMIN, DMAX, 3.0, TRUNC_MULT, 8)
    assert np.array_equal(_bits(d.cpu().numpy()), _bits(ref["depth"]))
    assert np.array_equal(_bits(c.cpu().numpy()), _bits(ref["color"]))
    truth = render_depth(model.get_current_frame_pose()).numpy().astype(np.float64)
    got = d.cpu().numpy()[..., 0].astype(np.float64)
    both = (got > 0) & (truth > 0)
    assert both.mean() > 0.8 and np.median(np.abs(got - truth)[both]) < 0.75 * VOXEL * SCALE
    rc2 = slam.Frame(480, 640, PRIMESENSE_K)
    model.synthesize_model_frame(rc2, SCALE, DMIN, DMAX, TRUNC_MULT, False)
    assert torch.equal(rc2.get_data("depth"), d) and not rc2.get_data("color").any()
