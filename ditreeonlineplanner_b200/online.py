"""Online-replanning helpers of the reference's lidar driver on the device kernels:

* ``scan_and_update_maze``         run_scenarios_with_lidar_DiTree.py:112-127
* ``check_no_obstacles_in_path``   run_scenarios_with_lidar_DiTree.py:158-181

Same arguments and side effects (the known map and the scanned map are updated in place and the
planner receives the new map through ``update_maze``).
"""
from __future__ import annotations

import numpy as np
import torch

from .common.map_utils import _ctx_for


def scan_and_update_maze(planner, maze_data, maze_data_with_obstacle, scanned_maze, debug=False):
    curr_state = planner.env.state
    pose = curr_state.copy()
    pose[:2] = planner.env.cell_xy_to_rowcol(curr_state[:2], floor_enable=False)
    pose[:2] = pose[:2][::-1]  # (col, row): the lidar works in grid coordinates
    distances, endpoints, visited_points = planner.env.lidar2dsim.scan(pose[:3], maze_data_with_obstacle, debug)
    endpoints = np.floor(endpoints).astype("int")
    maze_data[endpoints[:, 1], endpoints[:, 0]] = 1
    if len(visited_points):
        scanned_maze[visited_points[:, 1], visited_points[:, 0]] = 2
    scanned_maze[endpoints[:, 1], endpoints[:, 0]] = 1
    planner.update_maze(maze_data)
    return distances


def check_no_obstacles_in_path(planner, scanned_maze, main_path_array, debug=False):
    """Index of the first path point lying in a scanned obstacle cell, -1 if the path is clear."""
    ctx = _ctx_for(np.asarray(scanned_maze, dtype=np.float32), 1.0)
    path = np.ascontiguousarray(np.asarray(main_path_array, dtype=np.float32)[:, :2])
    idx = int(ctx.path_first_obstacle(torch.as_tensor(path))[0])
    ctx.sync_status()
    return idx
