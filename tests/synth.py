"""Synthetic inputs for the ICP and TSDF hot paths.

Test / bench infrastructure.  Deterministic in (size, seed).  The ICP clouds are
numpy; the RGB-D renderer is written in torch so that bench.py can render a
1000-frame sequence on the GPU in well under a second (tests run it on CPU).
"""
from __future__ import annotations

import math

import numpy as np
import torch

# ----------------------------------------------------------------------------- ICP


def _surface(x, y):
    """z = 0.3 sin(1.1x) cos(0.9y) + 0.05 sin(5x+1), with analytic unit normals."""
    z = 0.3 * np.sin(1.1 * x) * np.cos(0.9 * y) + 0.05 * np.sin(5 * x + 1)
    dzdx = 0.33 * np.cos(1.1 * x) * np.cos(0.9 * y) + 0.25 * np.cos(5 * x + 1)
    dzdy = -0.27 * np.sin(1.1 * x) * np.sin(0.9 * y)
    n = np.stack([-dzdx, -dzdy, np.ones_like(z)], axis=1)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return z, n


def _jittered_lattice(side, spacing, rng):
    i, j = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    x = (i.ravel() + 0.5 + rng.uniform(-0.4, 0.4, side * side)) * spacing
    y = (j.ravel() + 0.5 + rng.uniform(-0.4, 0.4, side * side)) * spacing
    return x, y


def axis_angle(axis, deg):
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    t = math.radians(deg)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(t) * K + (1 - math.cos(t)) * (K @ K)


def make_icp_pair(n_points, seed=1, spacing=0.02, angle_deg=None, noise=0.0005,
                  independent_sampling=True, shuffle=False):
    """Target = ~n_points samples (one per `spacing` lattice cell) of a smooth height
    field with analytic normals; source = an independent sampling of the same
    surface moved by a known rigid motion (+ Gaussian noise).

    Returns (source [N,3] f32, target [M,3] f32, target_normals [M,3] f32,
    T_gt [4,4] f64) with T_gt the source->target transformation ICP should find.
    The rotation (about the cloud centre, axis (1,2,-1)/sqrt6) defaults to 3 deg
    for clouds up to ~7 m across and shrinks with the extent so that the largest
    initial displacement stays near the 0.05 m correspondence radius.
    """
    rng = np.random.default_rng(seed)
    side = int(round(math.sqrt(n_points)))
    extent = side * spacing
    x, y = _jittered_lattice(side, spacing, rng)
    z, nrm = _surface(x, y)
    tgt = np.stack([x, y, z], axis=1)

    if independent_sampling:
        rng2 = np.random.default_rng(seed + 1000)
        xs, ys = _jittered_lattice(side, spacing, rng2)
        zs, _ = _surface(xs, ys)
        base = np.stack([xs, ys, zs], axis=1)
    else:
        rng2 = rng
        base = tgt.copy()

    if angle_deg is None:
        angle_deg = min(3.0, math.degrees(0.035 / (0.71 * extent)) if extent > 7 else 3.0)
    R = axis_angle([1, 2, -1], angle_deg)
    t = np.array([0.02, -0.01, 0.015])
    c = np.array([extent / 2, extent / 2, 0.0])
    # world motion M: p -> R (p - c) + c + t ; source = M(base) + noise, so T_gt = M^-1
    src = (base - c) @ R.T + c + t + rng2.normal(0, noise, base.shape)
    M = np.eye(4)
    M[:3, :3] = R
    M[:3, 3] = c + t - R @ c
    T_gt = np.linalg.inv(M)
    if shuffle:
        src = src[rng2.permutation(len(src))]
        p = rng2.permutation(len(tgt))
        tgt, nrm = tgt[p], nrm[p]
    return (np.ascontiguousarray(src, np.float32), np.ascontiguousarray(tgt, np.float32),
            np.ascontiguousarray(nrm, np.float32), T_gt)


def make_colors(points, seed=0):
    """Smooth procedural texture in [0,1] (config 4)."""
    p = np.asarray(points, np.float64)
    r = 0.5 + 0.5 * np.sin(2.1 * p[:, 0] + 0.3 * seed) * np.cos(1.7 * p[:, 1])
    g = 0.5 + 0.5 * np.sin(1.3 * p[:, 0] + 2.2 * p[:, 1] + 1.0)
    b = 0.5 + 0.5 * np.cos(0.9 * p[:, 0] - 1.9 * p[:, 1] + 4.0 * p[:, 2])
    return np.ascontiguousarray(np.stack([r, g, b], 1), np.float32)


# ---------------------------------------------------------------------------- TSDF

PRIMESENSE_K = np.array([[525.0, 0, 319.5], [0, 525.0, 239.5], [0, 0, 1.0]])  # camera/PinholeCameraIntrinsic.cpp:33-34

ROOM_MIN = (-3.0, -2.0, 0.0)
ROOM_MAX = (3.0, 2.0, 3.0)
SPHERE_C = (2.1, 0.2, 1.1)
SPHERE_R = 0.5


def camera_pose(i, n_frames=1000, radius=1.0, height=1.5, dyaw_deg=0.36, pitch_deg=12.0):
    """T_frame_to_world (4x4 f64): camera on a circle of `radius` at `height`, looking
    outward (towards the walls) and slightly down; yaw advances dyaw_deg per frame.
    Camera axes: x right, y down, z forward."""
    yaw = math.radians(dyaw_deg * i)
    pitch = math.radians(pitch_deg)
    pos = np.array([radius * math.cos(yaw), radius * math.sin(yaw), height])
    fwd = np.array([math.cos(yaw) * math.cos(pitch), math.sin(yaw) * math.cos(pitch), -math.sin(pitch)])
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = right, down, fwd, pos
    return T


def render_depth(T_frame_to_world, K=PRIMESENSE_K, width=640, height=480, depth_max=3.0,
                 device="cpu", with_color=False, tilt=None):
    """Analytic z-depth of the room interior + sphere.  Returns depth u16 [H,W] in
    millimetres (0 where z > depth_max or no hit) and optionally color u8 [H,W,3]."""
    dev = torch.device(device)
    T = torch.as_tensor(np.asarray(T_frame_to_world), dtype=torch.float64, device=dev)
    fx, fy, cx, cy = float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])
    v, u = torch.meshgrid(torch.arange(height, dtype=torch.float64, device=dev),
                          torch.arange(width, dtype=torch.float64, device=dev), indexing="ij")
    d_cam = torch.stack([(u - cx) / fx, (v - cy) / fy, torch.ones_like(u)], dim=-1)  # z = 1
    R, o = T[:3, :3], T[:3, 3]
    d = d_cam @ R.T  # world direction per unit z
    lo = torch.tensor(ROOM_MIN, dtype=torch.float64, device=dev)
    hi = torch.tensor(ROOM_MAX, dtype=torch.float64, device=dev)
    # exit distance from inside a box: per axis the far plane in the direction of travel
    inv = 1.0 / torch.where(d.abs() < 1e-12, torch.full_like(d, 1e-12), d)
    t_axis = torch.where(d > 0, (hi - o) * inv, (lo - o) * inv)
    t_wall, axis = t_axis.min(dim=-1)
    # sphere
    c = torch.tensor(SPHERE_C, dtype=torch.float64, device=dev)
    oc = o - c
    a = (d * d).sum(-1)
    b = 2 * (d * oc).sum(-1)
    cc = (oc * oc).sum() - SPHERE_R ** 2
    disc = b * b - 4 * a * cc
    t_s = torch.where(disc > 0, (-b - torch.sqrt(disc.clamp(min=0))) / (2 * a), torch.full_like(a, float("inf")))
    t_s = torch.where(t_s > 1e-6, t_s, torch.full_like(a, float("inf")))
    hit_sphere = t_s < t_wall
    z = torch.where(hit_sphere, t_s, t_wall)  # parametrised per unit camera z => z-depth
    depth = torch.where(z <= depth_max, torch.round(z * 1000.0), torch.zeros_like(z)).to(torch.int32)
    depth = depth.clamp(0, 65535).to(torch.uint16)
    if not with_color:
        return depth
    p = o + d * z.unsqueeze(-1)
    check = ((torch.floor(p[..., 0] * 4) + torch.floor(p[..., 1] * 4) + torch.floor(p[..., 2] * 4)) % 2)
    base = torch.where(check > 0.5, 200.0, 60.0)
    col = torch.stack([base, base * 0.8 + 20 * axis.to(torch.float64), 255.0 - base], dim=-1)
    col = torch.where(hit_sphere.unsqueeze(-1), torch.tensor([220.0, 40.0, 40.0], dtype=torch.float64, device=dev), col)
    return depth, col.clamp(0, 255).to(torch.uint8)
