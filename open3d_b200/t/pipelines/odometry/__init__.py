"""Mirror of ``open3d.t.pipelines.odometry`` for the PointToPlane method (cpp/pybind/t/pipelines/odometry/
odometry.cpp; cpp/open3d/t/pipelines/odometry/RGBDOdometry.{h,cpp}) — the tracking step of the dense-SLAM loop
(``slam.Model.track_frame_to_model``).

The whole coarse-to-fine Gauss-Newton loop runs device-resident inside libo3db200.so
(``o3db_rgbd_odometry_multi_scale_point_to_plane``); this module validates arguments like upstream and marshals the
result.  Method.Intensity / Method.Hybrid need the reference's NPP Sobel/Gaussian image filters and are outside this
build (SURVEY.md §8f #2 names the PointToPlane kernel).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from enum import IntEnum

import numpy as np
import torch

from ...geometry import Image, RGBDImage, _depth_dtype, _image_tensor, _k9
from ...._lib import (ERR_NO_INLIERS, ERR_SINGULAR, O3DBError, OdometryCriteria as _CCriteria,
                      OdometryResult as _CResult, check, dptr, lib)
from ....core import as_device_f32_points, as_host_f64_4x4, current_stream_ptr  # noqa: F401


class Method(IntEnum):
    """RGBDOdometry.h:24-30"""
    PointToPlane = 0
    Intensity = 1
    Hybrid = 2


@dataclass
class OdometryConvergenceCriteria:
    """RGBDOdometry.h:33-52 (note the argument order: max_iteration, relative_rmse, relative_fitness)."""
    max_iteration: int
    relative_rmse: float = 1e-6
    relative_fitness: float = 1e-6


@dataclass
class OdometryResult:
    """RGBDOdometry.h:54-78: transformation 4x4 Float64 on the host (source -> target)."""
    transformation: np.ndarray = field(default_factory=lambda: np.eye(4))
    inlier_rmse: float = 0.0
    fitness: float = 0.0


@dataclass
class OdometryLossParams:
    """RGBDOdometry.h:80-114"""
    depth_outlier_trunc: float = 0.07
    depth_huber_delta: float = 0.05
    intensity_huber_delta: float = 0.1

    def __post_init__(self):
        if self.depth_outlier_trunc < 0:
            print("[Open3D WARNING] Depth outlier truncation < 0, outliers will be counted!")
        if self.depth_huber_delta >= self.depth_outlier_trunc:
            print("[Open3D WARNING] Huber delta is greater than truncation, huber norm will degenerate to L2 norm!")


def _criteria_list(criteria_list):
    out = []
    for c in criteria_list:   # a bare int converts implicitly, as the C++ vector<OdometryConvergenceCriteria>{10, 5, 3}
        out.append(c if isinstance(c, OdometryConvergenceCriteria) else OdometryConvergenceCriteria(int(c)))
    return out


_criteria_cache = {}


def _c_criteria(criteria_list):
    """(C array, levels, total iterations) of a criteria list; the frame loop passes the same tuple of ints every
    frame, so that case is cached."""
    key = None
    if isinstance(criteria_list, tuple) and all(type(c) is int for c in criteria_list):
        key = criteria_list
        hit = _criteria_cache.get(key)
        if hit is not None:
            return hit
    crit = _criteria_list(criteria_list)
    arr = (_CCriteria * len(crit))(*[_CCriteria(int(c.max_iteration), float(c.relative_rmse),
                                                float(c.relative_fitness)) for c in crit])
    out = (arr, len(crit), sum(int(c.max_iteration) for c in crit))
    if key is not None and len(_criteria_cache) < 64:
        _criteria_cache[key] = out
    return out


def _raise(rc):
    if rc == ERR_SINGULAR:
        raise O3DBError(rc, "Singular 6x6 linear system detected, tracking failed.")
    if rc == ERR_NO_INLIERS:
        raise O3DBError(rc, "Invalid inlier_count value 0, must be > 0.")
    check(rc)


def rgbd_odometry_multi_scale(source, target, intrinsics, init_source_to_target=None, depth_scale=1000.0,
                              depth_max=3.0, criteria_list=(10, 5, 3), method=Method.Hybrid, params=None,
                              return_log=False):
    """RGBDOdometryMultiScale (RGBDOdometry.cpp:56-113).  source / target: RGBDImage (only depth is read by
    Method.PointToPlane).  Upstream's default method is Hybrid, which this build does not provide: pass
    ``method=Method.PointToPlane``."""
    if method != Method.PointToPlane:
        raise RuntimeError("open3d_b200 implements Method.PointToPlane; Intensity / Hybrid odometry need the "
                           "reference's NPP image filters and are outside this build's scope (SURVEY.md §8f #2).")
    params = params or OdometryLossParams()
    sd = _image_tensor(source.depth if isinstance(source, RGBDImage) else source)
    td = _image_tensor(target.depth if isinstance(target, RGBDImage) else target)
    if sd is None or td is None:
        raise RuntimeError("Invalid shape, expected a 1 channel image, but got an empty depth image")
    rows, cols = int(sd.shape[0]), int(sd.shape[1])
    if (int(td.shape[0]), int(td.shape[1])) != (rows, cols):
        raise RuntimeError("source and target depth images must have the same size")
    K = _k9(intrinsics)
    T0 = as_host_f64_4x4(np.eye(4) if init_source_to_target is None else init_source_to_target,
                         "init_source_to_target")
    arr, n_levels, total_iterations = _c_criteria(criteria_list)
    res = _CResult()
    # the per-iteration log costs a device-to-host copy + stream synchronisation at the end of the track: only on request
    per = np.zeros((max(total_iterations, 1), 2)) if return_log else None
    rc = lib.o3db_rgbd_odometry_multi_scale_point_to_plane(
        sd.data_ptr(), _depth_dtype(sd), td.data_ptr(), _depth_dtype(td), rows, cols, dptr(K), dptr(T0),
        float(depth_scale), float(depth_max), arr, n_levels, float(params.depth_outlier_trunc),
        float(params.depth_huber_delta), C.byref(res), None if per is None else dptr(per), current_stream_ptr())
    _raise(rc)
    out = OdometryResult(np.array(res.transformation[:], np.float64).reshape(4, 4), float(res.inlier_rmse),
                         float(res.fitness))
    return (out, per[: int(res.iterations)].copy()) if return_log else out


def compute_odometry_result_point_to_plane(source_vertex_map, target_vertex_map, target_normal_map, intrinsics,
                                           init_source_to_target, depth_outlier_trunc, depth_huber_delta):
    """ComputeOdometryResultPointToPlane (RGBDOdometry.cpp:432-459): one Gauss-Newton step; returns the DELTA
    transformation with inlier_rmse = sum(HuberLoss)/inliers and fitness = inliers / pixels."""
    maps = []
    for m in (source_vertex_map, target_vertex_map, target_normal_map):
        t = _image_tensor(m)
        if t is None or t.dtype != torch.float32 or t.dim() != 3 or t.shape[2] != 3:
            raise RuntimeError("vertex / normal maps must be [H, W, 3] Float32")   # kernel/RGBDOdometry.cpp:30-36
        maps.append(t)
    rows, cols = int(maps[0].shape[0]), int(maps[0].shape[1])
    if any((int(m.shape[0]), int(m.shape[1])) != (rows, cols) for m in maps):
        raise RuntimeError("vertex / normal maps must have the same size")
    K = _k9(intrinsics)
    T0 = as_host_f64_4x4(init_source_to_target, "init_source_to_target")
    dT = np.zeros((4, 4))
    rmse, fit = C.c_double(0), C.c_double(0)
    sums = np.zeros(29)
    rc = lib.o3db_compute_odometry_result_point_to_plane(maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr(),
                                                         rows, cols, dptr(K), dptr(T0), float(depth_outlier_trunc),
                                                         float(depth_huber_delta), dptr(dT), C.byref(rmse),
                                                         C.byref(fit), dptr(sums), current_stream_ptr())
    _raise(rc)
    res = OdometryResult(dT, rmse.value, fit.value)
    res.sums29 = sums
    return res
