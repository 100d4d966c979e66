"""CPU oracle for the DiTree tree-expansion hot path (geometry, dynamics, reductions).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``ditreeonlineplanner_b200``) never does; it fails loudly when ``libditree.so`` is missing.

Every function is a NumPy float64 restatement of one reference function; the docstring cites the
reference ``file:line`` (paths relative to the upstream repository root).  The oracle is *pinned*:
``tests/test_oracle_golden.py`` checks every function below against vectors produced by running the
unmodified reference (``tools/gen_golden.py``, run in the build container where the reference is
mounted) and committed under ``tests/golden/``.  The one un-pinned function is ``mppi_reduce``
(the reference imports its MPPI controller from a module that is not in its repository); it says
"parity unpinned" in its docstring and in DESIGN.md.
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------------------------
# coordinates (car_env.py:189-201; same formulas inlined at common/map_utils.py:155-162,247-252)
# --------------------------------------------------------------------------------------------


def map_center(grid_shape, s=1.0):
    """(x_center, y_center) of a grid of shape (rows, cols); car_env.py:83-86."""
    rows, cols = grid_shape
    return cols / 2 * s, rows / 2 * s


def xy_to_rowcol(x, y, grid_shape, s=1.0, floor=True):
    """car_env.py:196-201.  Row 0 is the top (largest y)."""
    cx, cy = map_center(grid_shape, s)
    i = (cy - np.asarray(y, dtype=np.float64)) / s
    j = (np.asarray(x, dtype=np.float64) + cx) / s
    if floor:
        return np.floor(i), np.floor(j)
    return i, j


def rowcol_to_xy(row, col, grid_shape, s=1.0):
    """car_env.py:189-194."""
    cx, cy = map_center(grid_shape, s)
    return (np.asarray(col) + 0.5) * s - cx, cy - (np.asarray(row) + 0.5) * s


# --------------------------------------------------------------------------------------------
# grid collision
# --------------------------------------------------------------------------------------------


def collide_points(points, grid, s=1.0, r=0.1):
    """Vectorised ball-vs-grid test, common/map_utils.py:221-329, including its three quirks:
    the whole-batch early return with only the out-of-bounds mask (:255-259), "diagonal neighbour
    outside the grid collides" (:324,327) and the diagonal *column* index clipped with the ROW
    count (:326; raises IndexError on maps with more rows than columns, like the reference)."""
    pts = np.asarray(points)
    if pts.ndim == 1:
        pts = pts[None]
    ax, ay = pts[:, 0], pts[:, 1]
    R, C = grid.shape
    cx, cy = C / 2 * s, R / 2 * s
    rows = np.floor((cy - ay) / s).astype(int)
    cols = np.floor((ax + cx) / s).astype(int)
    hit = (rows < 0) | (rows >= R) | (cols < 0) | (cols >= C)
    if hit.any():
        return hit
    hit = hit | (grid[rows, cols] == 1)
    if hit.all():
        return hit
    mid_x = (cols + 0.5) * s - cx
    mid_y = cy - (rows + 0.5) * s
    h = s / 2
    x_lo, x_hi, y_lo, y_hi = mid_x - h, mid_x + h, mid_y - h, mid_y + h
    hit = hit | ((ax + r > x_hi) & (grid[rows, np.clip(cols + 1, 0, C - 1)] == 1))
    hit = hit | ((ax - r < x_lo) & (grid[rows, np.clip(cols - 1, 0, C - 1)] == 1))
    hit = hit | ((ay + r > y_hi) & (grid[np.clip(rows - 1, 0, R - 1), cols] == 1))
    hit = hit | ((ay - r < y_lo) & (grid[np.clip(rows + 1, 0, R - 1), cols] == 1))
    if hit.all():
        return hit
    for kx, ky, di, dj in ((x_hi, y_hi, -1, 1), (x_lo, y_hi, -1, -1), (x_hi, y_lo, 1, 1), (x_lo, y_lo, 1, -1)):
        ci, cj = rows + di, cols + dj
        d = np.hypot(kx - ax, ky - ay)
        outside = (ci < 0) | (ci >= R) | (cj < 0) | (cj >= C)
        ci = np.clip(ci, 0, R - 1)
        cj = np.clip(cj, 0, R - 1)  # sic: reference clips the column with the row count
        hit = hit | outside | ((d < r) & (grid[ci, cj] == 1))
    return hit


def collide_car(states, grid, r=0.1, car_length=0.15):
    """Two-ball car test, common/map_utils.py:103-115, for a batch of (x, y, theta) rows.
    Each state is tested on its own two-ball pair exactly as the reference calls it (so the
    batch early-return of ``collide_points`` acts within a pair), then ``.any()``."""
    st = np.asarray(states, dtype=np.float64)
    if st.ndim == 1:
        st = st[None]
    out = np.zeros(len(st), dtype=bool)
    half = car_length * 0.5
    for n in range(len(st)):
        off = half * np.array([np.cos(st[n, 2]), np.sin(st[n, 2])])
        balls = np.array([st[n, :2] + off, st[n, :2] - off])
        out[n] = collide_points(balls, grid, 1.0, r).any()
    return out


def collide_car_batch(states, grid, r=0.1, car_length=0.15):
    """Same result as ``collide_car`` but vectorised over states (used for the full-size checks):
    per state, a ball out of bounds decides the pair; otherwise every test of map_utils.py:262-327
    is OR-ed over both balls.  Equivalence with the per-pair reference call: the early returns at
    :266 and :311 only fire when every ball already collides, which does not change ``.any()``."""
    st = np.asarray(states, dtype=np.float64)
    if st.ndim == 1:
        st = st[None]
    R, C = grid.shape
    if R > C:
        raise IndexError("reference indexes out of range on maps with more rows than columns")
    half = car_length * 0.5
    offx, offy = half * np.cos(st[:, 2]), half * np.sin(st[:, 2])
    res = np.zeros(len(st), dtype=bool)
    oob_any = np.zeros(len(st), dtype=bool)
    full_any = np.zeros(len(st), dtype=bool)
    cx, cy = C / 2, R / 2
    for sign in (1.0, -1.0):
        ax, ay = st[:, 0] + sign * offx, st[:, 1] + sign * offy
        rows = np.floor((cy - ay) / 1.0).astype(int)
        cols = np.floor((ax + cx) / 1.0).astype(int)
        oob = (rows < 0) | (rows >= R) | (cols < 0) | (cols >= C)
        oob_any |= oob
        rows_c, cols_c = np.clip(rows, 0, R - 1), np.clip(cols, 0, C - 1)
        hit = grid[rows_c, cols_c] == 1
        mid_x = (cols + 0.5) - cx
        mid_y = cy - (rows + 0.5)
        x_lo, x_hi, y_lo, y_hi = mid_x - 0.5, mid_x + 0.5, mid_y - 0.5, mid_y + 0.5
        hit |= (ax + r > x_hi) & (grid[rows_c, np.clip(cols + 1, 0, C - 1)] == 1)
        hit |= (ax - r < x_lo) & (grid[rows_c, np.clip(cols - 1, 0, C - 1)] == 1)
        hit |= (ay + r > y_hi) & (grid[np.clip(rows - 1, 0, R - 1), cols_c] == 1)
        hit |= (ay - r < y_lo) & (grid[np.clip(rows + 1, 0, R - 1), cols_c] == 1)
        for kx, ky, di, dj in ((x_hi, y_hi, -1, 1), (x_lo, y_hi, -1, -1), (x_hi, y_lo, 1, 1), (x_lo, y_lo, 1, -1)):
            ci, cj = rows + di, cols + dj
            d = np.hypot(kx - ax, ky - ay)
            outside = (ci < 0) | (ci >= R) | (cj < 0) | (cj >= C)
            ci = np.clip(ci, 0, R - 1)
            cj = np.clip(np.clip(cj, 0, R - 1), 0, C - 1)  # second clip is a no-op when R <= C
            hit |= outside | ((d < r) & (grid[ci, cj] == 1))
        full_any |= hit
    res = np.where(oob_any, True, full_any)
    return res


def collide_maze_scalar(state, grid, s=1.0, r=0.1):
    """Scalar maze test used for point/ant robots, common/map_utils.py:139-218.  Note: no
    inside-wall test; an out-of-range side neighbour collides; diagonals need an in-range cell."""
    x, y = float(state[0]), float(state[1])
    R, C = grid.shape
    cx, cy = C / 2 * s, R / 2 * s
    row = int(np.floor((cy - y) / s))
    col = int(np.floor((x + cx) / s))
    mid_x = (col + 0.5) * s - cx
    mid_y = cy - (row + 0.5) * s
    x_lo, x_hi = mid_x - s / 2, mid_x + s / 2
    y_lo, y_hi = mid_y - s / 2, mid_y + s / 2
    if not (0 <= row < R) or not (0 <= col < C):
        return True
    if x + r > x_hi and (col + 1 >= C or grid[row][col + 1] == 1):
        return True
    if x - r < x_lo and (col - 1 < 0 or grid[row][col - 1] == 1):
        return True
    if y + r > y_hi and (row - 1 < 0 or grid[row - 1][col] == 1):
        return True
    if y - r < y_lo and (row + 1 >= R or grid[row + 1][col] == 1):
        return True
    for kx, ky, ci, cj in ((x_hi, y_hi, row - 1, col + 1), (x_lo, y_hi, row - 1, col - 1),
                           (x_hi, y_lo, row + 1, col + 1), (x_lo, y_lo, row + 1, col - 1)):
        if math.sqrt((kx - x) ** 2 + (ky - y) ** 2) < r:
            if 0 <= ci < R and 0 <= cj < C and grid[ci][cj] == 1:
                return True
    return False


def collide_ant(state, grid, r=1.2, s=4.0):
    """common/map_utils.py:126-136 with common/se3_utils.py:155-164 (the quaternion slots 3..6 are
    unpacked as (w, x, y, z); upside-down iff 1 - 2(x^2 + y^2) < 0)."""
    qw, qx, qy, qz = (float(v) for v in state[3:7])
    if 1.0 - 2.0 * (qx * qx + qy * qy) < 0:
        return True
    return collide_maze_scalar(state[:3], grid, s, r)


def collide_ant_batch(states, grid, r=1.2, s=4.0):
    return np.array([collide_ant(st, grid, r, s) for st in np.asarray(states, dtype=np.float64)], dtype=bool)


# --------------------------------------------------------------------------------------------
# robot-centric local occupancy map
# --------------------------------------------------------------------------------------------


def local_map(grid, x, y, theta, n, scale, s_global, center):
    """common/map_utils.py:391-459: out[k, i, j] samples local point (xs[j], ys[i]) rotated by
    theta[k] and translated to (x[k], y[k]); nearest-cell gather with index clipping."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    y = np.atleast_1d(np.asarray(y, dtype=np.float64))
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    n = int(n)
    L = n * scale
    ax = np.linspace(-L / 2 + scale / 2, L / 2 - scale / 2, n)
    xl, yl = np.meshgrid(ax, ax)
    xl, yl = xl.ravel()[None], yl.ravel()[None]
    c, s_ = np.cos(theta)[:, None], np.sin(theta)[:, None]
    xg = c * xl - s_ * yl + x[:, None]
    yg = s_ * xl + c * yl + y[:, None]
    yi = np.floor((center[1] - yg) / s_global).astype(int)
    xi = np.floor((xg + center[0]) / s_global).astype(int)
    xi = np.clip(xi, 0, grid.shape[1] - 1)
    yi = np.clip(yi, 0, grid.shape[0] - 1)
    return grid[yi, xi].reshape(len(x), n, n)


# --------------------------------------------------------------------------------------------
# bicycle dynamics and the propagate wrapper
# --------------------------------------------------------------------------------------------

CAR = dict(m=0.043, C1=0.5, C2=15.5, Cm1=0.28, Cm2=0.05, Cr0=0.011, Cr2=0.006, dt=1.0 / 50.0)
ACTION_LOW = np.array([-10.0, -2.0], dtype=np.float32)
ACTION_HIGH = np.array([10.0, 2.0], dtype=np.float32)
GOAL_RADIUS = 0.5


def bicycle_step(state, action):
    """One explicit-Euler step of the 6-state bicycle model, car_env.py:356-396 (float64).
    Batched over leading dimensions."""
    s = np.asarray(state, dtype=np.float64)
    u = np.clip(np.asarray(action, dtype=np.float64), ACTION_LOW, ACTION_HIGH)
    psi, v, D, dl = s[..., 2], s[..., 3], s[..., 4], s[..., 5]
    p = CAR
    fxd = (p["Cm1"] - p["Cm2"] * v) * D - p["Cr2"] * (v ** 2) - p["Cr0"] * np.tanh(5.0 * v)
    dot = np.stack([
        v * np.cos(psi + p["C1"] * dl),
        v * np.sin(psi + p["C1"] * dl),
        v * p["C2"] * dl,
        (fxd / p["m"]) * np.cos(p["C1"] * dl),
        u[..., 0],
        u[..., 1],
    ], axis=-1)
    return s + p["dt"] * dot


def rollout_car(state0, actions, goal_xy, grid, stop_on_collision=True, states_for_flags=None):
    """Fused per-candidate restatement of planners/base_planner.py:257-320 over car_env.py:240-282:
    for each step: Euler step, goal test (car_env.py:341-350), record, collision test on the new
    state; a collision ends the edge, reaching the goal ends it after recording.

    state0 (B,6), actions (B,S,2) -> dict(traj (B,S,6) float64 with zero rows after termination,
    final (B,6), first_coll (B,) int32 or -1, done_step (B,) int32 or -1).

    ``states_for_flags`` (B,S,6), when given, teacher-forces the *flags*: collision and goal tests
    are evaluated on those states (e.g. the fp32 trajectory a kernel produced) instead of the
    oracle's own float64 trajectory; used for the bit-exact flag comparison."""
    s0 = np.asarray(state0, dtype=np.float64)
    act = np.asarray(actions, dtype=np.float64)
    B, S = act.shape[0], act.shape[1]
    traj = np.zeros((B, S, 6))
    final = s0.copy()
    first_coll = np.full(B, -1, dtype=np.int32)
    done_step = np.full(B, -1, dtype=np.int32)
    alive = np.ones(B, dtype=bool)
    cur = s0.copy()
    goal_xy = np.asarray(goal_xy, dtype=np.float64)
    for i in range(S):
        if not alive.any():
            break
        idx = np.nonzero(alive)[0]
        nxt = bicycle_step(cur[idx], act[idx, i])
        cur[idx] = nxt
        traj[idx, i] = nxt
        final[idx] = nxt
        probe = nxt if states_for_flags is None else np.asarray(states_for_flags, dtype=np.float64)[idx, i]
        done = np.sqrt(((probe[:, :2] - goal_xy) ** 2).sum(axis=1)) < GOAL_RADIUS
        coll = collide_car_batch(probe[:, :3], grid)
        newly = coll & (first_coll[idx] < 0)
        first_coll[idx[newly]] = i
        if stop_on_collision:
            alive[idx[coll]] = False
        # goal reached (and not colliding at this step when stopping on collisions)
        dn = done & ~(coll & stop_on_collision)
        done_step[idx[dn]] = i
        alive[idx[dn]] = False
    return dict(traj=traj, final=final, first_coll=first_coll, done_step=done_step)


def propagate_action_sequence(state, actions, horizon, goal_xy, grid):
    """Return conventions of planners/base_planner.py:257-320 for one car edge:
    (obs, done in {True, False, None}, actions (<=h,2), states (1,<=h+1,6))."""
    actions = np.array(actions, dtype=np.float64)
    states = np.zeros((horizon + 1, len(state)))
    states[0] = state
    obs = np.asarray(state, dtype=np.float64)
    done = False
    for i in range(len(actions[:horizon])):
        obs = bicycle_step(obs, actions[i])
        done = bool(np.linalg.norm(obs[:2] - np.asarray(goal_xy)) < GOAL_RADIUS)
        states[i + 1] = obs
        if collide_car(obs[:3], grid)[0]:
            return obs, None, actions[:i], states[:i][None]
        if done:
            actions[i + 1:] = 0
            break
    actions = actions[:horizon]
    states = states[:len(actions) + 1]
    return obs, done, actions, states[None]


# --------------------------------------------------------------------------------------------
# sampler conditioning (policies/fm_policy.py:53-162) and flow-matching schedule
# --------------------------------------------------------------------------------------------


def quat_to_rot6d(q):
    """common/se3_utils.py:177-189: slots are read as q = (x, y, z, w); the output is the first
    two columns of the rotation matrix, [r00, r10, r20, r01, r11, r21].  (The sampler applies it to
    the already mean/std-normalised slots, policies/fm_policy.py:77-79 -- kept as is.)"""
    q = np.asarray(q, dtype=np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    r00 = 1 - 2 * (y * y + z * z)
    r10 = 2 * (x * y + z * w)
    r20 = 2 * (x * z - y * w)
    r01 = 2 * (x * y - z * w)
    r11 = 1 - 2 * (x * x + z * z)
    r21 = 2 * (y * z + x * w)
    return np.stack([r00, r10, r20, r01, r11, r21], axis=-1)


def build_cond_car(obs_last, prev_action_last, goal, meta, local_map_size=20.0):
    """Condition vector for the car (policies/fm_policy.py:71-143): [v_n, D_n, delta_n, a0_n, a1_n,
    tanh(R(-yaw)(goal - p) / local_map_size)] as float32.  ``prev_action_last`` None -> zeros
    (left un-normalised, fm_policy.py:114-121).  obs_last (B,6), goal (2,) or (B,2)."""
    obs = np.asarray(obs_last, dtype=np.float64)
    B = len(obs)
    on = (obs - meta["Observations_mean"]) / meta["Observations_std"]
    cond = np.zeros((B, 7), dtype=np.float32)
    cond[:, :3] = on[:, 3:6].astype(np.float32)
    if prev_action_last is not None:
        an = (np.asarray(prev_action_last, dtype=np.float64) - meta["Actions_mean"]) / meta["Actions_std"]
        cond[:, 3:5] = an.astype(np.float32)
    g = (np.asarray(goal, dtype=np.float64) - obs[:, :2]).astype(np.float32)
    yaw = obs[:, 2].astype(np.float32)
    c, s = np.cos(yaw), np.sin(yaw)
    gx = c * g[:, 0] + s * g[:, 1]
    gy = -s * g[:, 0] + c * g[:, 1]
    cond[:, 5] = np.tanh(gx / np.float32(local_map_size))
    cond[:, 6] = np.tanh(gy / np.float32(local_map_size))
    return cond


def build_cond_ant(obs_seq, prev_action_last, goal, meta, obs_history=3, local_map_size=16.0):
    """Condition vector for the ant (policies/fm_policy.py:75-81,95-143): per history slot
    [z_n, rot6d(6), 22 normalised dims] (29), missing history = zeros in the leading slots,
    then 8 normalised previous actions, then tanh((goal - p) / local_map_size) without rotation.
    obs_seq (B,h,29) with h <= obs_history."""
    o = np.array(obs_seq, dtype=np.float64)
    if o.ndim == 2:
        o = o[:, None, :]
    B, h, _ = o.shape
    pos = o[:, -1, :2].copy()
    o[..., 2:] = (o[..., 2:] - meta["Observations_mean"]) / meta["Observations_std"]
    feat = np.concatenate([o[..., :3], quat_to_rot6d(o[..., 3:7]), o[..., 7:]], axis=-1)  # (B,h,31)
    slots = np.zeros((B, obs_history, feat.shape[-1]))
    pad = obs_history - h
    if pad > 0:
        slots[:, pad:] = feat
    else:
        slots[:] = feat[:, -obs_history:]
    parts = [slots[..., 2:].reshape(B, -1)]
    a = np.zeros((B, 8))
    if prev_action_last is not None:
        a = (np.asarray(prev_action_last, dtype=np.float64) - meta["Actions_mean"]) / meta["Actions_std"]
    parts.append(a)
    g = (np.asarray(goal, dtype=np.float64) - pos).astype(np.float32)
    parts.append(np.tanh(g / np.float32(local_map_size)))
    return np.concatenate([p.astype(np.float32) for p in parts], axis=1)


def fm_schedule(k_steps, exp_scale=4.0):
    """common/fm_utils.py:4-17 with schedule 'exp' in float32: returns (t0, dt)."""
    t = np.linspace(0.0, 1.0, k_steps + 1, dtype=np.float32)[:-1]
    dt = np.exp(-t * np.float32(exp_scale)).astype(np.float32)
    dt = (dt / dt.sum(dtype=np.float32)).astype(np.float32)
    t0 = np.concatenate([np.zeros(1, np.float32), np.cumsum(dt, dtype=np.float32)[:-1]])
    return t0, dt


# --------------------------------------------------------------------------------------------
# nearest neighbour / argmin reductions (planners/RRT.py:49-55, 87, 233-253)
# --------------------------------------------------------------------------------------------


def nearest(node_xy, query_xy):
    """1-NN in (x, y) over all nodes, first index on ties (what scipy's KDTree.query returns on
    the reference's data; brute force in float64).  node_xy (n,2), query_xy (Q,2) -> (Q,) int."""
    n = np.asarray(node_xy, dtype=np.float64)
    q = np.asarray(query_xy, dtype=np.float64)
    out = np.empty(len(q), dtype=np.int64)
    for s in range(0, len(q), 1024):
        d = (q[s:s + 1024, None, 0] - n[None, :, 0]) ** 2 + (q[s:s + 1024, None, 1] - n[None, :, 1]) ** 2
        out[s:s + 1024] = np.argmin(d, axis=1)
    return out


def nearest_k(node_xy, query_xy, k):
    """KDTree.query(query, k) indices (planners/RRT.py:50 with k > 1): ascending float64 squared distance,
    lowest index first on ties, n marks a missing neighbour (SciPy's convention) when k > n."""
    node_xy = np.asarray(node_xy, dtype=np.float64)
    query_xy = np.asarray(query_xy, dtype=np.float64)
    n = len(node_xy)
    out = np.full((len(query_xy), k), n, dtype=np.int64)
    for i, q in enumerate(query_xy):
        d2 = (q[0] - node_xy[:, 0]) ** 2 + (q[1] - node_xy[:, 1]) ** 2
        order = np.argsort(d2, kind="stable")[:k]
        out[i, :len(order)] = order
    return out


def final_node_cost_argmin(node_xy, goal_xy, obstacle_ahead):
    """planners/RRT.py:233-237: argmin over nodes of dist-to-goal + 10e3 * obstacle_ahead."""
    d = np.linalg.norm(np.asarray(node_xy, dtype=np.float64) - np.asarray(goal_xy, dtype=np.float64), axis=1)
    return int(np.argmin(d + 10e3 * np.asarray(obstacle_ahead, dtype=int)))


# --------------------------------------------------------------------------------------------
# lidar and obstacle probes
# --------------------------------------------------------------------------------------------

LIDAR_ANGLES_DEG = np.arange(-180.0, 180.0 + 2.0, 2.0)


def lidar_cast_ray(pose, maze, angle_deg):
    """lidar_sim/lidar_2d_sim.py:47-98 for one ray.  pose = (x=col, y=row, yaw) in GRID
    coordinates; the yaw (radians) is added to the angle in degrees before deg2rad (sic, :53-54).
    Returns (distance, hit_point (2,), visited cells (n,2) as (x, y) ints)."""
    x0, y0, yaw = (float(v) for v in pose)
    w, h = maze.shape  # sic: the reference names shape[0] "width"
    ang = np.deg2rad(yaw + angle_deg)
    ray = np.array([np.cos(ang), np.sin(ang)])
    p = np.array([x0, y0])
    borders = [(np.array([0, 0]), np.array([0, h])), (np.array([w, 0]), np.array([w, h])),
               (np.array([0, 0]), np.array([w, 0])), (np.array([0, h]), np.array([w, h]))]
    last = None
    for a, b in borders:
        di = (b - a).astype(float)
        A = np.column_stack((ray, -di))
        try:
            t, s = np.linalg.solve(A, a - p)
        except np.linalg.LinAlgError:
            continue
        if t >= 0 and 1 >= s >= 0:
            last = t * ray + p
            break
    ts = np.arange(0, 1, step=0.1 / np.linalg.norm(last - p))
    dots = p[None] + ts[:, None] * (last - p)[None]
    q = np.floor(dots).astype(int)
    q = np.clip(q, [0, 0], [w - 1, h - 1])
    occ = maze[q[:, 1], q[:, 0]]
    if np.any(occ == 1):
        first = int(np.where(occ == 1)[0][0])
        hit = dots[first]
    else:
        first = len(q)
        hit = last
    return float(np.linalg.norm(hit - p)), hit, q[:first]


def lidar_scan(pose, maze):
    """lidar_sim/lidar_2d_sim.py:18-45 with noise_std = 0: (dist (181,), endpoints (181,2),
    visited cells (*,2)).  Endpoint = pose + dist * (cos, sin)(deg2rad(yaw + angle))."""
    dists, ends, visited = [], [], []
    for a in LIDAR_ANGLES_DEG:
        d, _, cells = lidar_cast_ray(pose, maze, a)
        ang = np.deg2rad(pose[2] + a)
        d = float(np.clip(d, 0, 300))
        dists.append(d)
        ends.append((pose[0] + d * np.cos(ang), pose[1] + d * np.sin(ang)))
        visited.extend(cells)
    return np.array(dists), np.array(ends), np.array(visited, dtype=int).reshape(-1, 2)


def ray_probe(state, maze):
    """planners/RRT.py:61-81: 30 samples on linspace(0,1.5) along (cos(-theta), sin(-theta)) from
    the robot's un-floored (col,row); truncation toward zero (`astype(int)`), clip, any wall."""
    x, y, theta = (float(v) for v in state[:3])
    row, col = xy_to_rowcol(x, y, maze.shape, 1.0, floor=False)
    t = np.linspace(0, 1.5, 30)
    pts = t[:, None] @ np.array([[np.cos(-theta), np.sin(-theta)]]) + np.array([col, row])
    q = np.clip(pts.astype("int"), [0, 0], np.array(maze.shape[::-1]) - 1)
    return bool(np.any(maze[q[:, 1], q[:, 0]]))


def path_first_obstacle(path_xy, scanned_maze):
    """run_scenarios_with_lidar_DiTree.py:158-181: index of the first path point whose floored
    (col,row) cell equals 1 in the scanned map, else -1.  Pinned by tests/golden/online.npz (the reference
    function's own source, exec'ed by tools/gen_golden.py::gen_online, on 48 seeded paths)."""
    for idx, (x, y) in enumerate(np.asarray(path_xy, dtype=np.float64)[:, :2]):
        r, c = xy_to_rowcol(x, y, scanned_maze.shape, 1.0, floor=False)
        if scanned_maze[int(np.floor(r)), int(np.floor(c))] == 1:
            return idx
    return -1


def scan_and_update_maze(state, known_maze, maze_with_obstacle, scanned_maze):
    """run_scenarios_with_lidar_DiTree.py:112-127: lidar scan from the car's state (grid coordinates (col, row),
    un-floored), hit cells -> 1 in the known and the scanned map, crossed cells -> 2 in the scanned map (hits win).
    Updates the two maps in place; pinned by tests/golden/online.npz."""
    r, c = xy_to_rowcol(state[0], state[1], known_maze.shape, 1.0, floor=False)
    pose = np.array([c, r, state[2]], dtype=np.float64)
    _, ends, visited = lidar_scan(pose, maze_with_obstacle)
    ee = np.floor(ends).astype(int)
    known_maze[ee[:, 1], ee[:, 0]] = 1
    if len(visited):
        scanned_maze[visited[:, 1], visited[:, 0]] = 2
    scanned_maze[ee[:, 1], ee[:, 0]] = 1


# --------------------------------------------------------------------------------------------
# MPPI reduction  -- PARITY UNPINNED
# --------------------------------------------------------------------------------------------


def mppi_reduce(cost, noise, lam, u):
    """PARITY UNPINNED: the reference imports ``MPPI.mppi.MPPI`` from a module that is not in its
    repository (call sites run_scenarios_with_lidar_MPPI.py:339-341,422), so this restates the
    textbook MPPI update: w = softmax(-(c - min c)/lambda); u += sum_k w_k * noise_k; plus the
    arg-min rollout.  cost (K,), noise (K,T,A), u (T,A) -> (u_new, argmin, weights)."""
    c = np.asarray(cost, dtype=np.float64)
    beta = c.min()
    w = np.exp(-(c - beta) / lam)
    w = w / w.sum()
    u_new = np.asarray(u, dtype=np.float64) + np.tensordot(w, np.asarray(noise, dtype=np.float64), axes=(0, 0))
    return u_new, int(np.argmin(c)), w


# ---------------------------------------------------------------------------------------------
# probability-map state sampler (run_type >= 2)
# ---------------------------------------------------------------------------------------------
def edt_prior(grid):
    """CarEnv.prior (car_env.py:100-101): distance_transform_edt(1 - maze) / sum, by exhaustive search.
    The exact transform is the square root of an integer squared distance, so this matches SciPy's
    float64 values bit for bit on any map that has a wall cell (SciPy itself is not used: the oracle must
    not depend on which SciPy the GPU box carries)."""
    g = np.asarray(grid)
    rows, cols = g.shape
    fg = g != 1                      # non-zero entries of 1 - maze
    wr, wc = np.nonzero(~fg)
    rr, cc = np.mgrid[0:rows, 0:cols]
    if len(wr) == 0:                 # SciPy's result without any background cell (never the case in the data)
        d2 = (rr + 1) ** 2 + cc ** 2
    else:
        d2 = ((rr[..., None] - wr) ** 2 + (cc[..., None] - wc) ** 2).min(axis=-1)
    edt = np.sqrt(d2.astype(np.float64)) * fg
    return edt / np.sum(edt)


def gaussian_map(robot, goal, size=(20, 20)):
    """prob_sampling_utils.py:48-93: a discrete 2-D Gaussian elongated along robot -> goal; the mean slides
    from the goal (near) to the midpoint (far); zero at the robot's own cell; normalised."""
    H, W = size
    rx, ry = float(robot[0]), float(robot[1])
    gx, gy = float(goal[0]), float(goal[1])
    dx, dy = gx - rx, gy - ry
    d = np.sqrt(dx ** 2 + dy ** 2) + 1e-6
    u = np.array([dx, dy]) / d if d > 1e-6 else np.array([1.0, 0.0])
    v = np.array([-u[1], u[0]])
    mid = np.array([(rx + gx) / 2, (ry + gy) / 2])
    w = -np.exp(-d / 15) + 1
    mean = (1 - w) * np.array([gx, gy]) + w * mid
    s_long = 1.0 + 0.7 * np.log1p(d)
    s_side = 0.7 * s_long
    rot = np.stack([u, v], axis=1)
    sigma = rot @ np.diag([s_long ** 2, s_side ** 2]) @ rot.T
    inv = np.linalg.inv(sigma)
    ys, xs = np.mgrid[0:H, 0:W]
    diff = np.stack([xs, ys], axis=-1) - mean
    expo = np.sum((diff @ inv) * diff, axis=2)
    pdf = np.exp(-0.5 * expo)
    pdf[int(ry), int(rx)] = 0
    return pdf / pdf.sum()


def combine_log_blend(prior, gauss, beta=0.8, eps=1e-12):
    """prob_sampling_utils.py:150-172 without an obstacle mask (the reference never passes one)."""
    post = np.exp(beta * np.log(prior + eps) + (1.0 - beta) * np.log(gauss + eps)) * (prior > 0)
    s = post.sum()
    if s <= eps:
        post = prior.copy()
        s = post.sum()
        if s <= eps:
            post = np.ones_like(post)
            s = post.sum()
    return post / s


def sample_cells(prob, u):
    """np.random.choice(prob.size, p=prob.ravel()) for given uniform draws (legacy RandomState.choice):
    cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, side='right')."""
    cdf = np.cumsum(np.asarray(prob, dtype=np.float64).ravel())
    cdf /= cdf[-1]
    return cdf.searchsorted(np.asarray(u, dtype=np.float64), side="right")
