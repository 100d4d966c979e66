"""Drop-in mirror of the reference's ``common/map_utils.py`` geometry functions on the hot path,
same names / argument meaning / return types, computed by the sm_100a kernels through the C ABI.
Scalar calls cost one kernel launch each; batch them (``is_colliding_car`` accepts (B,3) too).

    create_local_map        common/map_utils.py:391-459
    is_colliding_car        common/map_utils.py:103-115
    is_colliding_parallel   common/map_utils.py:221-329
    is_colliding_maze       common/map_utils.py:139-218
    is_colliding_ant        common/map_utils.py:126-136
"""
from __future__ import annotations

import numpy as np
import torch

from ..runtime import get_context

cc_calls = 0  # collision-check counter the reference's drivers reset and read (map_utils.py:95-105)

_staged = {}


def _ctx_for(maze_map, scale=1.0, device=None):
    """Context with `maze_map` staged (re-uploaded only when the grid or its scale changed)."""
    ctx = get_context(device)
    device = ctx.device.index
    g = np.ascontiguousarray(np.asarray(maze_map, dtype=np.float32))
    key = (g.shape, float(scale), g.tobytes())
    if _staged.get(device) != key:
        ctx.set_map(g, scale)
        _staged[device] = key
    return ctx


def invalidate_staged_map(device=None):
    if device is None:
        _staged.clear()
    else:
        _staged.pop(device, None)


def create_local_map(global_map, x, y, theta, map_size, scale, s_global, map_center):
    """-> ndarray (K, N, N) float32 in {0, 1}.  `map_center` must be the grid centre
    (cols/2*s_global, rows/2*s_global), which is what every reference caller passes."""
    g = np.asarray(global_map)
    R, C = g.shape
    if abs(map_center[0] - C / 2 * s_global) > 1e-12 or abs(map_center[1] - R / 2 * s_global) > 1e-12:
        raise ValueError("create_local_map: map_center must be the centre of the grid")
    if isinstance(x, (int, float, np.generic)):
        x, y, theta = np.array([x]), np.array([y]), np.array([theta])
    n = int(map_size) if isinstance(map_size, (int, float)) else int(map_size[0])
    ctx = _ctx_for(g, s_global)
    poses = np.stack([np.asarray(x, np.float32), np.asarray(y, np.float32), np.asarray(theta, np.float32)], 1)
    out = ctx.local_map(torch.as_tensor(poses), n, scale)
    return out.cpu().numpy().astype(g.dtype if g.dtype.kind == "f" else np.float32)


def is_colliding_car(state, maze_map, ball_radius=0.1, car_length=0.15):
    """state (>=3,) -> bool, or (B,>=3) -> ndarray of bool."""
    global cc_calls
    if ball_radius != 0.1 or car_length != 0.15:
        raise NotImplementedError("the kernel is specialised for the reference's car (r=0.1, length=0.15)")
    st = np.asarray(state, dtype=np.float32)
    single = st.ndim == 1
    st = np.atleast_2d(st)[:, :3]
    cc_calls += len(st)
    ctx = _ctx_for(maze_map, 1.0)
    flags = ctx.collide_car(torch.as_tensor(np.ascontiguousarray(st))).cpu().numpy().astype(bool)
    ctx.sync_status()
    return bool(flags[0]) if single else flags


def is_colliding_parallel(states, maze_grid, maze_size_scaling=1, ball_radius=0.1):
    """(N,2) points -> ndarray (N,) bool, including the reference's whole-batch early return."""
    pts = np.asarray(states, dtype=np.float32)
    if pts.ndim == 1:
        pts = pts[None]
    ctx = _ctx_for(maze_grid, maze_size_scaling)
    flags = ctx.collide_points(torch.as_tensor(np.ascontiguousarray(pts[:, :2])), maze_size_scaling, ball_radius)
    out = flags.cpu().numpy().astype(bool)
    ctx.sync_status()
    return out


def is_colliding_ant(state, maze_map, ant_radius=1, map_scale=1):
    st = np.asarray(state, dtype=np.float32)
    single = st.ndim == 1
    st = np.ascontiguousarray(np.atleast_2d(st)[:, :7])
    ctx = _ctx_for(maze_map, map_scale)
    flags = ctx.collide_ant(torch.as_tensor(st), ant_radius).cpu().numpy().astype(bool)
    return bool(flags[0]) if single else flags


def is_colliding_maze(state, maze_grid, maze_size_scaling=1, ball_radius=0.1):
    st = np.asarray(state, dtype=np.float32)
    single = st.ndim == 1
    st = np.atleast_2d(st)
    full = np.zeros((len(st), 7), np.float32)
    full[:, :2] = st[:, :2]
    full[:, 3] = 1.0  # upright: only the maze test decides
    ctx = _ctx_for(maze_grid, maze_size_scaling)
    flags = ctx.collide_ant(torch.as_tensor(full), ball_radius).cpu().numpy().astype(bool)
    return bool(flags[0]) if single else flags
