"""Drop-in mirror of the reference's ``lidar_sim/lidar_2d_sim.py::Lidar2DSim`` (scan :18-45,
_cast_ray :47-98) on the ray-marching kernel (one warp per bundle of 32 rays)."""
from __future__ import annotations

import numpy as np
import torch

from ..common.map_utils import _ctx_for


class Lidar2DSim:
    def __init__(self, azimuth_fov_deg=360, azimuth_res_deg=2.0, max_range=300, noise_std=0.0, scan_time=0.2):
        if azimuth_fov_deg != 360 or azimuth_res_deg != 2.0 or max_range != 300:
            raise NotImplementedError("the kernel is specialised for the reference's 181-ray, 2 degree, 300-cell lidar")
        self.azimuth_fov = azimuth_fov_deg
        self.azimuth_res = azimuth_res_deg
        self.max_range = max_range
        self.noise_std = noise_std
        self.scan_time = scan_time
        self.angles_deg = np.arange(-self.azimuth_fov / 2, self.azimuth_fov / 2 + self.azimuth_res, self.azimuth_res)

    def scan(self, robot_state, maze_data, debug=False):
        """robot_state (x=col, y=row, yaw) in grid coordinates -> (distances (181,), endpoints (181,2),
        visited cells (n,2) as (x, y)).  The visited cells are returned once each in row-major order
        (the reference lists them per ray with repeats; its callers only use them as an index set)."""
        d, e, v = self.scan_batch(np.asarray(robot_state, dtype=np.float64)[None], maze_data)
        ys, xs = np.nonzero(v[0])
        return d[0], e[0], np.stack([xs, ys], 1)

    def scan_batch(self, poses, maze_data):
        """The kernel marches the rays in float64 from a float32 pose (the device state format; the reference keeps
        the pose in float64): hit cells are exact and distances within 1e-5 for float32-representable poses, which
        is what the planner's float32 states are."""
        ctx = _ctx_for(maze_data, 1.0)
        poses = np.ascontiguousarray(np.asarray(poses, dtype=np.float32)[:, :3])
        dist, end, vis = ctx.lidar_scan(torch.as_tensor(poses))
        dist, end, vis = dist.cpu().numpy(), end.cpu().numpy(), vis.cpu().numpy()
        ctx.sync_status()
        if self.noise_std != 0.0:
            # the reference draws one normal per ray (lidar_2d_sim.py:31) and moves the endpoint along the ray
            noise = np.random.normal(0, self.noise_std, dist.shape)
            nd = np.clip(dist + noise, 0, self.max_range)
            ang = np.deg2rad(poses[:, 2:3].astype(np.float64) + self.angles_deg[None])
            end = np.stack([poses[:, 0:1] + nd * np.cos(ang), poses[:, 1:2] + nd * np.sin(ang)], -1)
            dist = nd
        else:
            np.random.normal(0, 1.0, dist.size)  # keeps NumPy's global RNG stream aligned with the reference (:31)
        return dist, end, vis
