"""ctypes front-end of the CPU oracle (oracle/*.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` leg — never from
``open3d_b200`` (the product).  See oracle/oracle.h for what each function
restates (reference file:line) and its parity status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_SOURCES = ["icp_oracle.c", "tsdf_oracle.c", "oracle.h"]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile if missing or stale."""
    stale = force or not os.path.exists(_LIB_PATH)
    if not stale:
        t = os.path.getmtime(_LIB_PATH)
        stale = any(os.path.getmtime(os.path.join(_HERE, s)) > t for s in _SOURCES)
    if stale:
        subprocess.run(["make", "-B", "-C", _HERE, "liboracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)


class _IcpResult(C.Structure):
    _fields_ = [("num_iterations", C.c_int), ("converged", C.c_int),
                ("fitness", C.c_double), ("inlier_rmse", C.c_double),
                ("transformation", C.c_double * 16), ("loop_seconds", C.c_double),
                ("build_seconds", C.c_double)]


def _declare(L):
    L.orc_minivec_hash_i32x3.restype = C.c_uint64
    L.orc_minivec_hash_i32x3.argtypes = [C.c_int32] * 3
    L.orc_spatial_hash.restype = C.c_uint64
    L.orc_spatial_hash.argtypes = [C.c_int32] * 3
    L.orc_compute_voxel_index_f32.restype = None
    L.orc_compute_voxel_index_f32.argtypes = [_f32p, C.c_float, _i32p]
    L.orc_robust_weight_f64.restype = C.c_double
    L.orc_robust_weight_f64.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    L.orc_robust_weight_f32.restype = C.c_float
    L.orc_robust_weight_f32.argtypes = [C.c_int, C.c_double, C.c_double, C.c_float]
    for name in ("orc_hybrid_search_f32", "orc_hybrid_search_bruteforce_f32"):
        f = getattr(L, name)
        f.restype = None
        f.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_double, C.c_int,
                      _i32p, _f32p, _i32p]
    L.orc_pose_p2plane_sums_f32.restype = None
    L.orc_pose_p2plane_sums_f32.argtypes = [_f32p, _f32p, _f32p, _i64p, C.c_int64,
                                            C.c_int, C.c_double, C.c_double,
                                            _f64p, _f32p, _f64p]
    L.orc_pose_p2plane_sums_f64.restype = None
    L.orc_pose_p2plane_sums_f64.argtypes = [_f64p, _f64p, _f64p, _i64p, C.c_int64,
                                            C.c_int, C.c_double, C.c_double, _f64p]
    L.orc_pose_colored_sums_f32.restype = None
    L.orc_pose_colored_sums_f32.argtypes = [_f32p] * 6 + [_i64p, C.c_int64, C.c_double,
                                                         C.c_int, C.c_double, C.c_double,
                                                         _f64p, _f64p]
    L.orc_decode_and_solve_6x6.restype = C.c_int
    L.orc_decode_and_solve_6x6.argtypes = [_f64p, _f64p, _f64p, C.POINTER(C.c_int)]
    L.orc_pose_to_transformation.restype = None
    L.orc_pose_to_transformation.argtypes = [_f64p, _f64p]
    L.orc_transform_points_f32.restype = None
    L.orc_transform_points_f32.argtypes = [_f64p, _f32p, C.c_int64]
    L.orc_transform_normals_f32.restype = None
    L.orc_transform_normals_f32.argtypes = [_f64p, _f32p, C.c_int64]
    L.orc_rmse_p2plane_f32.restype = C.c_double
    L.orc_rmse_p2plane_f32.argtypes = [_f32p, _f32p, _f32p, _i64p, C.c_int64]
    L.orc_icp_p2plane_f32.restype = C.c_int
    L.orc_icp_p2plane_f32.argtypes = [_f32p, C.c_int64, _f32p, _f32p, C.c_int64,
                                      C.c_double, _f64p, C.c_int, C.c_double, C.c_double,
                                      C.c_int, C.c_double, C.c_double, C.c_int,
                                      C.POINTER(_IcpResult), _f64p, _i64p]
    L.orc_inverse_transformation.restype = None
    L.orc_inverse_transformation.argtypes = [_f64p, _f64p]
    L.orc_depth_touch.restype = C.c_int64
    L.orc_depth_touch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _f64p, _f64p,
                                  C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                  C.c_int, _i32p, C.c_int64]
    L.orc_tsdf_integrate.restype = None
    L.orc_tsdf_integrate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     _i32p, C.c_int64, _i32p, _f32p, _u16p, _u16p,
                                     _f64p, _f64p, _f64p, C.c_int, C.c_float, C.c_float,
                                     C.c_float, C.c_float]
    L.orc_tsdf_integrate_f32_values.restype = None
    L.orc_tsdf_integrate_f32_values.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                _i32p, C.c_int64, _i32p, _f32p, _f32p, _f32p,
                                                _f64p, _f64p, _f64p, C.c_int, C.c_float, C.c_float,
                                                C.c_float, C.c_float]
    L.orc_hashmap_activate.restype = C.c_int
    L.orc_hashmap_activate.argtypes = [_i32p, C.c_int64, _i64p, _i32p, C.c_int64,
                                       _i32p, _u8p]
    L.orc_estimate_range.restype = None
    L.orc_estimate_range.argtypes = [_i32p, C.c_int64, _f64p, _f64p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_float, C.c_float, _f32p]
    L.orc_ray_cast.restype = None
    L.orc_ray_cast.argtypes = [_i32p, C.c_int64, _f32p, _u16p, C.c_void_p, _f32p, _f64p, _f64p, C.c_int, C.c_int,
                               C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                               C.c_int] + [C.c_void_p] * 10
    L.orc_clip_transform.restype = None
    L.orc_clip_transform.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                     C.c_float, _f32p]
    L.orc_pyr_down_depth.restype = None
    L.orc_pyr_down_depth.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, C.c_float, _f32p]
    L.orc_create_vertex_map.restype = None
    L.orc_create_vertex_map.argtypes = [_f32p, C.c_int, C.c_int, _f64p, C.c_float, _f32p]
    L.orc_create_normal_map.restype = None
    L.orc_create_normal_map.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _f32p]
    L.orc_filter_bilateral_f32.restype = None
    L.orc_filter_bilateral_f32.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _f32p]
    L.orc_huber_deriv.restype = C.c_float
    L.orc_huber_deriv.argtypes = [C.c_float, C.c_float]
    L.orc_huber_loss.restype = C.c_float
    L.orc_huber_loss.argtypes = [C.c_float, C.c_float]
    L.orc_odometry_jacobian_p2plane.restype = C.c_int
    L.orc_odometry_jacobian_p2plane.argtypes = [C.c_int, C.c_int, C.c_float, _f32p, _f32p, _f32p, C.c_int, C.c_int,
                                                _f64p, _f64p, _f32p, _f32p]
    L.orc_odometry_p2plane_sums.restype = None
    L.orc_odometry_p2plane_sums.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, _f64p, _f64p, C.c_float, C.c_float,
                                            _f64p, _f64p]
    L.orc_compute_odometry_result_p2plane.restype = C.c_int
    L.orc_compute_odometry_result_p2plane.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, _f64p, _f64p, C.c_float,
                                                      C.c_float, _f64p, _f64p, _f64p]
    L.orc_rgbd_odometry_multi_scale_p2plane.restype = C.c_int
    L.orc_rgbd_odometry_multi_scale_p2plane.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, _f64p, _f64p,
                                                        C.c_float, C.c_float, C.c_int, C.POINTER(C.c_int), _f64p,
                                                        _f64p, C.c_float, C.c_float, _f64p, _f64p, _f64p, _f64p,
                                                        C.POINTER(C.c_int)]
    L.orc_estimate_color_gradients_f32.restype = None
    L.orc_estimate_color_gradients_f32.argtypes = [_f32p, _f32p, _f32p, C.c_int64, C.c_double, C.c_int, _f32p]
    L.orc_estimate_color_gradients_solver_f32.restype = None
    L.orc_estimate_color_gradients_solver_f32.argtypes = [_f32p, _f32p, _f32p, C.c_int64, C.c_double, C.c_int, C.c_int, _f32p]
    L.orc_svd3x3_f32.restype = None
    L.orc_svd3x3_f32.argtypes = [_f32p, _f32p, _f32p, _f32p]
    L.orc_solve_svd3x3_f32.restype = None
    L.orc_solve_svd3x3_f32.argtypes = [_f32p, _f32p, _f32p]
    L.orc_information_matrix_f32.restype = None
    L.orc_information_matrix_f32.argtypes = [_f32p, _i64p, C.c_int64, _f64p]
    L.orc_solve_sym3x3_pinv.restype = None
    L.orc_solve_sym3x3_pinv.argtypes = [_f64p, _f64p, _f64p]
    L.orc_icp_colored_f32.restype = C.c_int
    L.orc_icp_colored_f32.argtypes = [_f32p, _f32p, C.c_int64, _f32p, _f32p, _f32p, _f32p, C.c_int64, C.c_double, _f64p,
                                      C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                      C.POINTER(_IcpResult), _f64p, _i64p]
    L.orc_voxel_down_sample_f32.restype = C.c_int64
    L.orc_voxel_down_sample_f32.argtypes = [_f32p, _f32p, _f32p, C.c_int64, C.c_double, _f32p, _f32p, _f32p, _i32p]
    L.orc_num_threads.restype = C.c_int
    L.orc_num_threads.argtypes = []
    L.orc_set_num_threads.restype = None
    L.orc_set_num_threads.argtypes = [C.c_int]


def _arr(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _p(a, ptype):
    return a.ctypes.data_as(ptype)


ROBUST = {"L2Loss": 0, "L1Loss": 1, "HuberLoss": 2, "CauchyLoss": 3, "GMLoss": 4,
          "TukeyLoss": 5, "GeneralizedLoss": 6}


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> int:
    """OpenMP threads of the CPU legs (process-wide ICV: also covers oracle/_ref, which links the same libgomp)."""
    lib().orc_set_num_threads(int(n))
    return num_threads()


def _cgroup_cpu_quota():
    """CPUs granted by the cgroup (v2 cpu.max / v1 cfs quota), or None."""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            return max(1, int(float(q) / float(per)))
    except (OSError, ValueError):
        pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0 and per > 0:
            return max(1, q // per)
    except (OSError, ValueError):
        pass
    return None


def host_cores() -> int:
    """Hardware threads this process may run on: the affinity mask, capped by the cgroup CPU quota."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    q = _cgroup_cpu_quota()
    return max(1, min(n, q) if q else n)


def physical_cores() -> int:
    """Distinct physical cores inside the affinity mask (SMT siblings counted once), capped like host_cores()."""
    import os
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return host_cores()
    seen = set()
    for c in cpus:
        try:
            sib = open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list").read().strip()
        except OSError:
            sib = str(c)
        seen.add(sib)
    return max(1, min(len(seen), host_cores()))


def thread_candidates():
    """Thread counts worth timing for an "all host cores" CPU leg: every hardware thread, and one per physical core."""
    return sorted({host_cores(), physical_cores()}, reverse=True)


def minivec_hash(keys) -> np.ndarray:
    keys = _arr(keys, np.int32).reshape(-1, 3)
    L = lib()
    return np.array([L.orc_minivec_hash_i32x3(int(k[0]), int(k[1]), int(k[2])) for k in keys],
                    dtype=np.uint64)


def spatial_hash(cells) -> np.ndarray:
    cells = _arr(cells, np.int32).reshape(-1, 3)
    L = lib()
    return np.array([L.orc_spatial_hash(int(k[0]), int(k[1]), int(k[2])) for k in cells],
                    dtype=np.uint64)


def compute_voxel_index(pos, inv_voxel_size) -> np.ndarray:
    pos = _arr(pos, np.float32).reshape(3)
    out = np.zeros(3, np.int32)
    lib().orc_compute_voxel_index_f32(_p(pos, _f32p), float(inv_voxel_size), _p(out, _i32p))
    return out


def robust_weight(method, scale, shape, residual, f32=False) -> float:
    m = ROBUST[method] if isinstance(method, str) else int(method)
    if f32:
        return float(lib().orc_robust_weight_f32(m, scale, shape, residual))
    return float(lib().orc_robust_weight_f64(m, scale, shape, residual))


def hybrid_search(points, queries, radius, max_knn=1, bruteforce=False):
    """-> (idx [N,k] int32, dist2 [N,k] f32, counts [N] int32)."""
    points = _arr(points, np.float32).reshape(-1, 3)
    queries = _arr(queries, np.float32).reshape(-1, 3)
    n = queries.shape[0]
    idx = np.empty((n, max_knn), np.int32)
    d2 = np.empty((n, max_knn), np.float32)
    cnt = np.empty(n, np.int32)
    f = lib().orc_hybrid_search_bruteforce_f32 if bruteforce else lib().orc_hybrid_search_f32
    f(_p(points, _f32p), points.shape[0], _p(queries, _f32p), n, float(radius), int(max_knn),
      _p(idx, _i32p), _p(d2, _f32p), _p(cnt, _i32p))
    return idx, d2, cnt


def pose_p2plane_sums(src, tgt, nrm, corr, robust=("L2Loss", 1.0, 1.0)):
    """-> dict(sums64, sums32, abs64) for f32 clouds."""
    src = _arr(src, np.float32).reshape(-1, 3)
    tgt = _arr(tgt, np.float32).reshape(-1, 3)
    nrm = _arr(nrm, np.float32).reshape(-1, 3)
    corr = _arr(corr, np.int64).reshape(-1)
    s64 = np.zeros(29, np.float64)
    s32 = np.zeros(29, np.float32)
    a64 = np.zeros(29, np.float64)
    lib().orc_pose_p2plane_sums_f32(_p(src, _f32p), _p(tgt, _f32p), _p(nrm, _f32p),
                                    _p(corr, _i64p), src.shape[0], ROBUST[robust[0]],
                                    float(robust[1]), float(robust[2]),
                                    _p(s64, _f64p), _p(s32, _f32p), _p(a64, _f64p))
    return {"sums64": s64, "sums32": s32, "abs64": a64}


def pose_p2plane_sums_f64(src, tgt, nrm, corr, robust=("L2Loss", 1.0, 1.0)):
    src = _arr(src, np.float64).reshape(-1, 3)
    tgt = _arr(tgt, np.float64).reshape(-1, 3)
    nrm = _arr(nrm, np.float64).reshape(-1, 3)
    corr = _arr(corr, np.int64).reshape(-1)
    s64 = np.zeros(29, np.float64)
    lib().orc_pose_p2plane_sums_f64(_p(src, _f64p), _p(tgt, _f64p), _p(nrm, _f64p),
                                    _p(corr, _i64p), src.shape[0], ROBUST[robust[0]],
                                    float(robust[1]), float(robust[2]), _p(s64, _f64p))
    return s64


def pose_colored_sums(src, src_colors, tgt, nrm, tgt_colors, tgt_grad, corr,
                      lambda_geometric=0.968, robust=("L2Loss", 1.0, 1.0)):
    a = [_arr(x, np.float32).reshape(-1, 3) for x in (src, src_colors, tgt, nrm, tgt_colors, tgt_grad)]
    corr = _arr(corr, np.int64).reshape(-1)
    s64 = np.zeros(29, np.float64)
    a64 = np.zeros(29, np.float64)
    lib().orc_pose_colored_sums_f32(*[_p(x, _f32p) for x in a], _p(corr, _i64p), a[0].shape[0],
                                    float(lambda_geometric), ROBUST[robust[0]],
                                    float(robust[1]), float(robust[2]),
                                    _p(s64, _f64p), _p(a64, _f64p))
    return {"sums64": s64, "abs64": a64}


def decode_and_solve_6x6(sums29):
    """-> (pose[6], residual, inlier_count, singular)."""
    s = _arr(sums29, np.float64).reshape(29)
    pose = np.zeros(6, np.float64)
    res = C.c_double(0)
    cnt = C.c_int(0)
    rc = lib().orc_decode_and_solve_6x6(_p(s, _f64p), _p(pose, _f64p), C.byref(res), C.byref(cnt))
    return pose, res.value, cnt.value, bool(rc)


def pose_to_transformation(pose) -> np.ndarray:
    p = _arr(pose, np.float64).reshape(6)
    T = np.zeros(16, np.float64)
    lib().orc_pose_to_transformation(_p(p, _f64p), _p(T, _f64p))
    return T.reshape(4, 4)


def transform_points(T, points) -> np.ndarray:
    T = _arr(T, np.float64).reshape(16)
    pts = np.array(points, dtype=np.float32, order="C", copy=True).reshape(-1, 3)
    lib().orc_transform_points_f32(_p(T, _f64p), _p(pts, _f32p), pts.shape[0])
    return pts


def transform_normals(T, normals) -> np.ndarray:
    T = _arr(T, np.float64).reshape(16)
    nr = np.array(normals, dtype=np.float32, order="C", copy=True).reshape(-1, 3)
    lib().orc_transform_normals_f32(_p(T, _f64p), _p(nr, _f32p), nr.shape[0])
    return nr


def rmse_p2plane(src, tgt, nrm, corr) -> float:
    src = _arr(src, np.float32).reshape(-1, 3)
    tgt = _arr(tgt, np.float32).reshape(-1, 3)
    nrm = _arr(nrm, np.float32).reshape(-1, 3)
    corr = _arr(corr, np.int64).reshape(-1)
    return float(lib().orc_rmse_p2plane_f32(_p(src, _f32p), _p(tgt, _f32p), _p(nrm, _f32p),
                                            _p(corr, _i64p), src.shape[0]))


@dataclass
class IcpResult:
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    converged: bool
    num_iterations: int
    per_iteration: np.ndarray  # [num_iterations_executed, 2] (fitness, rmse)
    correspondences: np.ndarray  # [N] int64
    status: int  # 0 ok, 1 singular system (the reference raises)
    loop_seconds: float = 0.0
    build_seconds: float = 0.0


def icp_p2plane(source, target, target_normals, max_corr_dist, init=None, max_iteration=30,
                relative_fitness=1e-6, relative_rmse=1e-6, robust=("L2Loss", 1.0, 1.0),
                accumulate_f64=True) -> IcpResult:
    src = _arr(source, np.float32).reshape(-1, 3)
    tgt = _arr(target, np.float32).reshape(-1, 3)
    nrm = _arr(target_normals, np.float32).reshape(-1, 3)
    T0 = _arr(np.eye(4) if init is None else init, np.float64).reshape(16)
    res = _IcpResult()
    per = np.full((max(max_iteration, 1), 2), np.nan, np.float64)
    corr = np.empty(src.shape[0], np.int64)
    rc = lib().orc_icp_p2plane_f32(_p(src, _f32p), src.shape[0], _p(tgt, _f32p), _p(nrm, _f32p),
                                   tgt.shape[0], float(max_corr_dist), _p(T0, _f64p),
                                   int(max_iteration), float(relative_fitness),
                                   float(relative_rmse), ROBUST[robust[0]], float(robust[1]),
                                   float(robust[2]), int(bool(accumulate_f64)),
                                   C.byref(res), _p(per, _f64p), _p(corr, _i64p))
    executed = int(np.sum(~np.isnan(per[:, 0])))
    return IcpResult(np.array(res.transformation, np.float64).reshape(4, 4), res.fitness,
                     res.inlier_rmse, bool(res.converged), res.num_iterations,
                     per[:executed].copy(), corr, rc, res.loop_seconds, res.build_seconds)


def inverse_transformation(T) -> np.ndarray:
    T = _arr(T, np.float64).reshape(16)
    Ti = np.zeros(16, np.float64)
    lib().orc_inverse_transformation(_p(T, _f64p), _p(Ti, _f64p))
    return Ti.reshape(4, 4)


def depth_touch(depth, K, extrinsic, resolution=16, voxel_size=0.008, sdf_trunc=0.064,
                depth_scale=1000.0, depth_max=3.0, stride=4) -> np.ndarray:
    """-> unique block keys [B,3] int32, lexicographically sorted."""
    depth = np.ascontiguousarray(depth)
    assert depth.dtype in (np.uint16, np.float32)
    depth = depth.reshape(depth.shape[0], depth.shape[1])
    rows, cols = depth.shape
    K = _arr(K, np.float64).reshape(9)
    E = _arr(extrinsic, np.float64).reshape(16)
    cap = (rows // stride) * (cols // stride) * 4 + 1
    out = np.empty((cap, 3), np.int32)
    n = lib().orc_depth_touch(depth.ctypes.data, int(depth.dtype == np.float32), rows, cols,
                              _p(K, _f64p), _p(E, _f64p), int(resolution), float(voxel_size),
                              float(sdf_trunc), float(depth_scale), float(depth_max),
                              int(stride), _p(out, _i32p), cap)
    assert n >= 0, n
    return out[:n].copy()


def tsdf_integrate(depth, color, buf_indices, block_keys, tsdf, weight, color_buf, depth_K,
                   color_K, extrinsic, resolution=16, voxel_size=0.008, sdf_trunc=0.064,
                   depth_scale=1000.0, depth_max=3.0) -> None:
    """In-place update of tsdf/weight/color_buf numpy buffers (slam::Model layout)."""
    depth = np.ascontiguousarray(depth)
    f32 = depth.dtype == np.float32
    assert f32 or depth.dtype == np.uint16
    rows, cols = depth.shape[0], depth.shape[1]
    if color is not None:
        color = np.ascontiguousarray(color)
        assert color.dtype == (np.float32 if f32 else np.uint8)
        assert color.shape[0] == rows and color.shape[1] == cols
    bi = _arr(buf_indices, np.int32).reshape(-1)
    assert block_keys.dtype == np.int32 and block_keys.flags.c_contiguous
    assert tsdf.dtype == np.float32 and tsdf.flags.c_contiguous
    values_f32 = weight.dtype == np.float32         # the reference's (Float32, Float32) value layout
    assert weight.dtype in (np.uint16, np.float32) and weight.flags.c_contiguous
    if color_buf is not None:
        assert color_buf.dtype == weight.dtype and color_buf.flags.c_contiguous
    dK = _arr(depth_K, np.float64).reshape(9)
    cK = _arr(color_K if color_K is not None else depth_K, np.float64).reshape(9)
    E = _arr(extrinsic, np.float64).reshape(16)
    fn = lib().orc_tsdf_integrate_f32_values if values_f32 else lib().orc_tsdf_integrate
    vp = _f32p if values_f32 else _u16p
    fn(depth.ctypes.data, None if color is None else color.ctypes.data,
                             int(f32), rows, cols, _p(bi, _i32p), bi.shape[0],
                             _p(block_keys, _i32p), _p(tsdf, _f32p), _p(weight, vp),
                             None if color_buf is None else _p(color_buf, vp),
                             _p(dK, _f64p), _p(cK, _f64p), _p(E, _f64p), int(resolution),
                             float(voxel_size), float(sdf_trunc), float(depth_scale),
                             float(depth_max))


def estimate_range(block_keys, intrinsic, extrinsic, height, width, down_factor=8, resolution=16,
                   voxel_size=0.008, depth_min=0.1, depth_max=3.0) -> np.ndarray:
    """-> [height // down, width // down, 2] f32 (min, max) range map."""
    keys = _arr(block_keys, np.int32).reshape(-1, 3)
    K = _arr(intrinsic, np.float64).reshape(9)
    E = _arr(extrinsic, np.float64).reshape(16)
    out = np.empty((height // down_factor, width // down_factor, 2), np.float32)
    lib().orc_estimate_range(_p(keys, _i32p), keys.shape[0], _p(K, _f64p), _p(E, _f64p), int(height), int(width),
                             int(down_factor), int(resolution), float(voxel_size), float(depth_min),
                             float(depth_max), _p(out, _f32p))
    return out


RAYCAST_ATTRS = {"depth": (1, np.float32), "vertex": (3, np.float32), "color": (3, np.float32),
                 "normal": (3, np.float32), "index": (8, np.int64), "mask": (8, np.uint8),
                 "interp_ratio": (8, np.float32), "interp_ratio_dx": (8, np.float32),
                 "interp_ratio_dy": (8, np.float32), "interp_ratio_dz": (8, np.float32)}


def ray_cast(table_keys, size, tsdf, weight, color_buf, range_map, intrinsic, extrinsic, height, width,
             attrs=("depth", "color"), resolution=16, voxel_size=0.008, depth_scale=1000.0, depth_min=0.1,
             depth_max=3.0, weight_threshold=3.0, trunc_voxel_multiplier=8.0, range_map_down_factor=8) -> dict:
    """VoxelBlockGrid::RayCast on the oracle's buffers -> {attr: [h, w, c] array}."""
    assert table_keys.dtype == np.int32 and table_keys.flags.c_contiguous
    assert tsdf.dtype == np.float32 and weight.dtype == np.uint16
    rng = _arr(range_map, np.float32)
    K = _arr(intrinsic, np.float64).reshape(9)
    E = _arr(extrinsic, np.float64).reshape(16)
    out = {}
    ptrs = []
    for name in ("depth", "vertex", "color", "normal", "index", "mask", "interp_ratio", "interp_ratio_dx",
                 "interp_ratio_dy", "interp_ratio_dz"):
        if name in attrs:
            c, dt = RAYCAST_ATTRS[name]
            out[name] = np.full((height, width, c), 77, dt)
            ptrs.append(out[name].ctypes.data)
        else:
            ptrs.append(None)
    lib().orc_ray_cast(_p(table_keys, _i32p), int(size), _p(tsdf, _f32p), _p(weight, _u16p),
                       None if color_buf is None else color_buf.ctypes.data, _p(rng, _f32p), _p(K, _f64p),
                       _p(E, _f64p), int(height), int(width), int(resolution), float(voxel_size),
                       float(depth_scale), float(depth_min), float(depth_max), float(weight_threshold),
                       float(trunc_voxel_multiplier), int(range_map_down_factor), *ptrs)
    if "mask" in out:
        out["mask"] = out["mask"].astype(bool)
    return out


def hashmap_activate(table_keys, size, keys):
    """table_keys [cap,3] int32 (in place), size int -> (buf_indices, masks, new_size, rc)."""
    assert table_keys.dtype == np.int32 and table_keys.flags.c_contiguous
    keys = _arr(keys, np.int32).reshape(-1, 3)
    n = keys.shape[0]
    bi = np.empty(n, np.int32)
    mk = np.empty(n, np.uint8)
    sz = C.c_int64(int(size))
    rc = lib().orc_hashmap_activate(_p(table_keys, _i32p), table_keys.shape[0], C.byref(sz),
                                    _p(keys, _i32p), n, _p(bi, _i32p), _p(mk, _u8p))
    return bi, mk.astype(bool), sz.value, rc


def voxel_down_sample(positions, voxel_size, normals=None, colors=None):
    """-> dict(positions, normals, colors, keys): one mean point per occupied voxel, voxels sorted."""
    pos = _arr(positions, np.float32).reshape(-1, 3)
    n = pos.shape[0]
    nrm = None if normals is None else _arr(normals, np.float32).reshape(-1, 3)
    col = None if colors is None else _arr(colors, np.float32).reshape(-1, 3)
    po = np.empty((n, 3), np.float32)
    no = np.empty((n, 3), np.float32) if nrm is not None else None
    co = np.empty((n, 3), np.float32) if col is not None else None
    keys = np.empty((n, 3), np.int32)
    null = C.POINTER(C.c_float)()
    m = lib().orc_voxel_down_sample_f32(_p(pos, _f32p), null if nrm is None else _p(nrm, _f32p),
                                        null if col is None else _p(col, _f32p), n, float(voxel_size),
                                        _p(po, _f32p), null if no is None else _p(no, _f32p),
                                        null if co is None else _p(co, _f32p), _p(keys, _i32p))
    assert m >= 0
    return {"positions": po[:m].copy(), "normals": None if no is None else no[:m].copy(),
            "colors": None if co is None else co[:m].copy(), "keys": keys[:m].copy()}


GRADIENT_SOLVERS = {"reference": 0, "exact": 1}


def information_matrix(target, corr) -> np.ndarray:
    """kernel::ComputeInformationMatrix: 6x6 f64 GTG over the matched target points."""
    t = _arr(target, np.float32).reshape(-1, 3)
    c = _arr(corr, np.int64).reshape(-1)
    out = np.zeros(36, np.float64)
    lib().orc_information_matrix_f32(_p(t, _f32p), _p(c, _i64p), c.shape[0], _p(out, _f64p))
    return out.reshape(6, 6)


def get_information_matrix(source, target, max_correspondence_distance, transformation) -> np.ndarray:
    """registration::GetInformationMatrix (Registration.cpp:446-485): transform a clone of the source, hybrid search
    (k = 1) on the target, GTG over the matched target points; raises when there is no correspondence."""
    moved = transform_points(transformation, source)
    idx, _, cnt = hybrid_search(target, moved, max_correspondence_distance, 1)
    if int(cnt.sum()) == 0:
        raise RuntimeError("0 correspondence present between the pointclouds. Try increasing the "
                           "max_correspondence_distance parameter.")
    return information_matrix(target, idx[:, 0].astype(np.int64))


def estimate_color_gradients(points, normals, colors, radius, max_nn=30, solver="reference") -> np.ndarray:
    """solver="reference": upstream's solve_svd3x3<float> (bit-identical to the reference kernel);
    "exact": exact pseudo-inverse of the same f32 normal equations."""
    p = _arr(points, np.float32).reshape(-1, 3)
    nr = _arr(normals, np.float32).reshape(-1, 3)
    c = _arr(colors, np.float32).reshape(-1, 3)
    out = np.zeros_like(p)
    lib().orc_estimate_color_gradients_solver_f32(_p(p, _f32p), _p(nr, _f32p), _p(c, _f32p), p.shape[0], float(radius),
                                                  int(max_nn), GRADIENT_SOLVERS[solver], _p(out, _f32p))
    return out


def svd3x3(A):
    """(U, S, V) of upstream's svd3x3<float> (SVD3x3.h:1131-2168), f32, row-major."""
    A = _arr(A, np.float32).reshape(9)
    U, S, V = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(9, np.float32)
    lib().orc_svd3x3_f32(_p(A, _f32p), _p(U, _f32p), _p(S, _f32p), _p(V, _f32p))
    return U.reshape(3, 3), S, V.reshape(3, 3)


def solve_svd3x3(A, b) -> np.ndarray:
    """upstream's solve_svd3x3<float> (SVD3x3.h:2170-2215)."""
    A = _arr(A, np.float32).reshape(9)
    b = _arr(b, np.float32).reshape(3)
    x = np.zeros(3, np.float32)
    lib().orc_solve_svd3x3_f32(_p(A, _f32p), _p(b, _f32p), _p(x, _f32p))
    return x


def solve_sym3x3_pinv(A, b) -> np.ndarray:
    A = _arr(A, np.float64).reshape(9)
    b = _arr(b, np.float64).reshape(3)
    x = np.zeros(3)
    lib().orc_solve_sym3x3_pinv(_p(A, _f64p), _p(b, _f64p), _p(x, _f64p))
    return x


def icp_colored(source, source_colors, target, target_normals, target_colors, target_color_gradients,
                max_corr_dist, init=None, max_iteration=30, relative_fitness=1e-6, relative_rmse=1e-6,
                lambda_geometric=0.968, robust=("L2Loss", 1.0, 1.0)) -> IcpResult:
    src, sc = (_arr(a, np.float32).reshape(-1, 3) for a in (source, source_colors))
    tgt, nrm, tc, tg = (_arr(a, np.float32).reshape(-1, 3) for a in (target, target_normals, target_colors,
                                                                       target_color_gradients))
    T0 = _arr(np.eye(4) if init is None else init, np.float64).reshape(16)
    res = _IcpResult()
    per = np.full((max(max_iteration, 1), 2), np.nan, np.float64)
    corr = np.empty(src.shape[0], np.int64)
    rc = lib().orc_icp_colored_f32(_p(src, _f32p), _p(sc, _f32p), src.shape[0], _p(tgt, _f32p), _p(nrm, _f32p),
                                   _p(tc, _f32p), _p(tg, _f32p), tgt.shape[0], float(max_corr_dist), _p(T0, _f64p),
                                   int(max_iteration), float(relative_fitness), float(relative_rmse),
                                   float(lambda_geometric), ROBUST[robust[0]], float(robust[1]), float(robust[2]),
                                   C.byref(res), _p(per, _f64p), _p(corr, _i64p))
    executed = int(np.sum(~np.isnan(per[:, 0])))
    return IcpResult(np.array(res.transformation, np.float64).reshape(4, 4), res.fitness, res.inlier_rmse,
                     bool(res.converged), res.num_iterations, per[:executed].copy(), corr, rc, res.loop_seconds,
                     res.build_seconds)


# ---------------------------------------------------------------- RGB-D odometry (PointToPlane)

def clip_transform(depth, scale=1000.0, min_value=0.0, max_value=3.0, clip_fill=float("nan")) -> np.ndarray:
    d = np.ascontiguousarray(depth)
    d = d.reshape(d.shape[0], d.shape[1])
    f32 = d.dtype == np.float32
    assert f32 or d.dtype == np.uint16
    out = np.empty(d.shape, np.float32)
    lib().orc_clip_transform(d.ctypes.data, int(f32), d.shape[0], d.shape[1], float(scale), float(min_value),
                             float(max_value), float(clip_fill), _p(out, _f32p))
    return out


def pyr_down_depth(depth, depth_diff, invalid_fill=float("nan")) -> np.ndarray:
    d = _arr(depth, np.float32)
    d = d.reshape(d.shape[0], d.shape[1])
    out = np.empty((d.shape[0] // 2, d.shape[1] // 2), np.float32)
    lib().orc_pyr_down_depth(_p(d, _f32p), d.shape[0], d.shape[1], float(depth_diff), float(invalid_fill),
                             _p(out, _f32p))
    return out


def create_vertex_map(depth, intrinsic, invalid_fill=float("nan")) -> np.ndarray:
    d = _arr(depth, np.float32)
    d = d.reshape(d.shape[0], d.shape[1])
    K = _arr(intrinsic, np.float64).reshape(9)
    out = np.empty(d.shape + (3,), np.float32)
    lib().orc_create_vertex_map(_p(d, _f32p), d.shape[0], d.shape[1], _p(K, _f64p), float(invalid_fill),
                                _p(out, _f32p))
    return out


def create_normal_map(vertex, invalid_fill=float("nan")) -> np.ndarray:
    v = _arr(vertex, np.float32)
    out = np.empty(v.shape, np.float32)
    lib().orc_create_normal_map(_p(v, _f32p), v.shape[0], v.shape[1], float(invalid_fill), _p(out, _f32p))
    return out


def filter_bilateral(img, kernel_size=5, value_sigma=5.0, dist_sigma=10.0) -> np.ndarray:
    a = _arr(img, np.float32)
    a = a.reshape(a.shape[0], a.shape[1])
    out = np.empty(a.shape, np.float32)
    lib().orc_filter_bilateral_f32(_p(a, _f32p), a.shape[0], a.shape[1], int(kernel_size), float(value_sigma),
                                   float(dist_sigma), _p(out, _f32p))
    return out


def huber_deriv(r, delta) -> float:
    return float(lib().orc_huber_deriv(float(r), float(delta)))


def huber_loss(r, delta) -> float:
    return float(lib().orc_huber_loss(float(r), float(delta)))


def odometry_jacobian_p2plane(x, y, source_vertex, target_vertex, target_normal, intrinsic, T, depth_outlier_trunc=0.07):
    sv, tv, tn = (_arr(a, np.float32) for a in (source_vertex, target_vertex, target_normal))
    K = _arr(intrinsic, np.float64).reshape(9)
    Tm = _arr(T, np.float64).reshape(16)
    J = np.zeros(6, np.float32)
    r = np.zeros(1, np.float32)
    ok = lib().orc_odometry_jacobian_p2plane(int(x), int(y), float(depth_outlier_trunc), _p(sv, _f32p), _p(tv, _f32p),
                                             _p(tn, _f32p), sv.shape[0], sv.shape[1], _p(K, _f64p), _p(Tm, _f64p),
                                             _p(J, _f32p), _p(r, _f32p))
    return bool(ok), J, float(r[0])


def odometry_p2plane_sums(source_vertex, target_vertex, target_normal, intrinsic, T, depth_outlier_trunc=0.07,
                          depth_huber_delta=0.05) -> dict:
    sv, tv, tn = (_arr(a, np.float32) for a in (source_vertex, target_vertex, target_normal))
    K = _arr(intrinsic, np.float64).reshape(9)
    Tm = _arr(T, np.float64).reshape(16)
    s, a = np.zeros(29), np.zeros(29)
    lib().orc_odometry_p2plane_sums(_p(sv, _f32p), _p(tv, _f32p), _p(tn, _f32p), sv.shape[0], sv.shape[1],
                                    _p(K, _f64p), _p(Tm, _f64p), float(depth_outlier_trunc), float(depth_huber_delta),
                                    _p(s, _f64p), _p(a, _f64p))
    return {"sums64": s, "abs64": a}


def compute_odometry_result_p2plane(source_vertex, target_vertex, target_normal, intrinsic, T,
                                    depth_outlier_trunc=0.07, depth_huber_delta=0.05):
    """-> (rc, delta 4x4, inlier_rmse, fitness)"""
    sv, tv, tn = (_arr(a, np.float32) for a in (source_vertex, target_vertex, target_normal))
    K = _arr(intrinsic, np.float64).reshape(9)
    Tm = _arr(T, np.float64).reshape(16)
    dT = np.zeros(16)
    rmse, fit = C.c_double(0), C.c_double(0)
    rc = lib().orc_compute_odometry_result_p2plane(_p(sv, _f32p), _p(tv, _f32p), _p(tn, _f32p), sv.shape[0],
                                                   sv.shape[1], _p(K, _f64p), _p(Tm, _f64p),
                                                   float(depth_outlier_trunc), float(depth_huber_delta), _p(dT, _f64p),
                                                   C.byref(rmse), C.byref(fit))
    return rc, dT.reshape(4, 4), rmse.value, fit.value


def rgbd_odometry_multi_scale_p2plane(source_depth, target_depth, intrinsic, init=None, depth_scale=1000.0,
                                      depth_max=3.0, criteria=((10, 1e-6, 1e-6), (5, 1e-6, 1e-6), (3, 1e-6, 1e-6)),
                                      depth_outlier_trunc=0.07, depth_huber_delta=0.05) -> dict:
    """criteria: (max_iteration, relative_rmse, relative_fitness) per level, coarse to fine."""
    s, t = np.ascontiguousarray(source_depth), np.ascontiguousarray(target_depth)
    s, t = s.reshape(s.shape[0], s.shape[1]), t.reshape(t.shape[0], t.shape[1])
    assert s.dtype == t.dtype and s.shape == t.shape and s.dtype in (np.uint16, np.float32)
    K = _arr(intrinsic, np.float64).reshape(9)
    T0 = _arr(np.eye(4) if init is None else init, np.float64).reshape(16)
    n = len(criteria)
    it = (C.c_int * n)(*[int(c[0]) for c in criteria])
    rr = np.array([c[1] for c in criteria], np.float64)
    rf = np.array([c[2] for c in criteria], np.float64)
    T = np.zeros(16)
    rmse, fit, done = C.c_double(0), C.c_double(0), C.c_int(0)
    per = np.zeros((max(sum(int(c[0]) for c in criteria), 1), 2))
    rc = lib().orc_rgbd_odometry_multi_scale_p2plane(s.ctypes.data, t.ctypes.data, int(s.dtype == np.float32),
                                                     s.shape[0], s.shape[1], _p(K, _f64p), _p(T0, _f64p),
                                                     float(depth_scale), float(depth_max), n, it, _p(rr, _f64p),
                                                     _p(rf, _f64p), float(depth_outlier_trunc),
                                                     float(depth_huber_delta), _p(T, _f64p), C.byref(rmse),
                                                     C.byref(fit), _p(per, _f64p), C.byref(done))
    return {"status": rc, "transformation": T.reshape(4, 4), "inlier_rmse": rmse.value, "fitness": fit.value,
            "per_iteration": per[: done.value].copy()}
