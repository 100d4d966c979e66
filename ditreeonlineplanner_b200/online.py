"""Online-replanning helpers of the reference's lidar driver on the device kernels:

* ``scan_and_update_maze``         run_scenarios_with_lidar_DiTree.py:112-127
* ``check_no_obstacles_in_path``   run_scenarios_with_lidar_DiTree.py:158-181

Same arguments and side effects (the known map and the scanned map are updated in place and the
planner receives the new map through ``update_maze``).
"""
from __future__ import annotations

import numpy as np
import torch

from .runtime import get_context


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
    """Index of the first path point lying in a scanned obstacle cell, -1 if the path is clear.  The scanned map
    (0 unknown / 1 obstacle / 2 seen) travels with the path as a kernel argument: the planner's staged map stays
    resident (one upload of rows*cols bytes per scan instead of two map re-stagings)."""
    ctx = get_context()
    path = np.ascontiguousarray(np.asarray(main_path_array, dtype=np.float32)[:, :2])
    idx = int(ctx.path_first_obstacle_grid(np.asarray(scanned_maze).astype(np.uint8), torch.as_tensor(path))[0])
    ctx.sync_status()
    return idx


def plan_path(planner, curr_state, goal_state, stats2keep, time_budget=None):
    """run_scenarios_with_lidar_DiTree.py:65-76: one (re)plan from the current state; the planner's
    node / iteration counters are accumulated into ``stats2keep`` and its time budget is restored."""
    planner.plan_count += 1
    keep = planner.time_budget
    planner.reset(start_state=curr_state, goal_state=goal_state)
    if time_budget is not None:
        planner.time_budget = time_budget
    path, actions = planner.plan()
    for k in stats2keep:
        stats2keep[k] += planner.results[k]
    planner.reset(start_state=curr_state, goal_state=goal_state)
    planner.time_budget = keep
    return path, actions


def run_online_episode(planner, start_state, goal_state, maze_data_original, maze_data_with_obstacle, run_type=0,
                       offline_time_budget=None, allowed_trials=5, max_actions=None, clock=None):
    """One iteration of the reference's online driver (run_scenarios_with_lidar_DiTree.py:397-520): initial lidar
    scan, reference ("main") plan on the known map, then execute the plan one action at a time on the true
    map; every ``lidar2dsim.scan_time`` seconds of driving the lidar is fired, the known / scanned maps are
    updated and the remaining main path is checked against the newly seen obstacles -- a hit (or an exhausted
    plan, or, for run_type >= 4, two seconds since the last plan) triggers a replan from the current state; five
    failed replans, a collision or the goal end the episode.

    -> dict(executed_path (n,6), executed_actions (n,2), success (True / False / None on collision),
            replans, scans, stats{number_of_nodes, iterations}, known_maze, scanned_maze)
    ``max_actions`` bounds the episode (the reference loops until one of the three endings); ``clock`` replaces
    ``time.time`` (tests)."""
    import time as _time
    now = clock if clock is not None else _time.time
    env = planner.env
    stats = dict(number_of_nodes=0, iterations=0)
    executed_path, executed_actions = [], []
    planner.init_main_path = None
    planner.plan_count = -1                      # the main path is plan 0
    curr_state = np.array(start_state, dtype=np.float64)
    maze_data = np.array(maze_data_original, dtype=np.float64)
    scanned = np.array(maze_data_original, dtype=np.float64)
    scans = replans = 0
    done = False
    planner.reset(start_state=curr_state, goal_state=goal_state, reset_main_path=True)
    scan_and_update_maze(planner, maze_data, maze_data_with_obstacle, scanned)
    scans += 1
    main_path, main_actions = plan_path(planner, curr_state, goal_state, stats, time_budget=offline_time_budget)
    if main_path is not None and main_actions is not None:
        planner.init_main_path = main_path.copy()
        env.reset_done()
        action_idx = trials = 0
        blocked_at = -1
        last_plan = now()
        while not env.is_done(curr_state):
            forced = run_type >= 4 and now() - last_plan >= 2
            if blocked_at >= 0 or action_idx == main_actions.shape[0] or forced:
                main_actions = None
                while trials < allowed_trials and main_actions is None:
                    main_path, main_actions = plan_path(planner, curr_state, goal_state, stats)
                    replans += 1
                    if main_actions is None:
                        trials += 1
                if trials >= allowed_trials:
                    break
                trials = 0
                env.reset_done()
                planner.reset(start_state=curr_state, goal_state=goal_state)
                action_idx = 0
                last_plan = now()
            env.set_state(curr_state)
            blocked_at = -1
            since_scan = driven = 0.0
            done = False
            while action_idx < main_actions.shape[0] and blocked_at < 0 and not done and \
                    not (run_type >= 4 and now() - last_plan >= 2 and driven >= 2):
                nxt, done, _, visited = planner.propagate_action_sequence_env(curr_state, main_actions[action_idx, np.newaxis])
                if done is None:                 # collision with the true map: the episode is over
                    break
                executed_path.append(visited[0, 1, :])
                executed_actions.append(main_actions[action_idx])
                curr_state = nxt
                env.set_state(curr_state)
                action_idx += 1
                since_scan += env.dt
                driven += env.dt
                if since_scan > env.lidar2dsim.scan_time:
                    scan_and_update_maze(planner, maze_data, maze_data_with_obstacle, scanned)
                    scans += 1
                    blocked_at = check_no_obstacles_in_path(planner, scanned, main_path)
                    since_scan = 0.0
                if max_actions is not None and len(executed_actions) >= max_actions:
                    break
            if done is None or (max_actions is not None and len(executed_actions) >= max_actions):
                break
    success = None if done is None else bool(done)
    return dict(executed_path=np.array(executed_path).reshape(-1, 6), executed_actions=np.array(executed_actions).reshape(-1, 2),
                success=success, replans=replans, scans=scans, stats=stats, known_maze=maze_data, scanned_maze=scanned)
