"""Scenario-suite sharding: the unit of work is one (scenario row, run index) pair of
the reference's benchmark loop (run_scenarios.py:202,336-395).  A single tree stays on one GPU;
ranks take disjoint units, there is no data-path collective, and the per-run result rows (the
reference's CSV fields, run_scenarios.py:392-395) are gathered once with ``all_gather``.

Deliberate, documented deviation from the reference: it seeds the three RNG streams once and
lets them run across scenarios (run_scenarios.py:86-90); here every unit is seeded from
(base seed, scenario index, run index) so results do not depend on the number of GPUs.
"""
from __future__ import annotations

import random
import time

import numpy as np
import torch

from .data import load_maze, load_scenarios

ROW_FIELDS = ["iteration", "success", "runtime", "trajectory_length", "trajectory_time", "avg_velocity",
              "num_states_in_tree", "num_RRT_iterations", "ctrl_effort_max", "ctrl_effort_mean", "ctrl_effort_std"]


def all_units(n_scenarios, total_runs, weights=None, by_weight=False):
    """Every (scenario, run) unit, heaviest scenario first (weights default to equal).  Run-major by default
    (every rank of a round-robin deal gets the same mix); `by_weight` lists ALL runs of the heaviest scenario
    first -- the longest-processing-time order a shared work queue wants."""
    order = list(range(n_scenarios))
    if weights is not None:
        order.sort(key=lambda i: -weights[i])
    if by_weight:
        return [(s, r) for s in order for r in range(total_runs)]
    return [(s, r) for r in range(total_runs) for s in order]


def shard_units(n_scenarios, total_runs, rank, world, weights=None):
    """Units of `rank`: round-robin over the weight-sorted list, so every rank gets the same mix."""
    return all_units(n_scenarios, total_runs, weights)[rank::world]


def unit_seed(scenario_idx, run_idx, base=42):
    return (base * 1_000_003 + scenario_idx * 1009 + run_idx) % (2 ** 31 - 1)


def seed_everything(seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


def scenario_states(row, env):
    """Start / goal states of a test_scenarios_car row (run_scenarios.py:239-246)."""
    start_xy = env.cell_rowcol_to_xy(np.array([int(row["start_row"]), int(row["start_col"])]))
    goal_xy = env.cell_rowcol_to_xy(np.array([int(row["goal_row"]), int(row["goal_col"])]))
    start = np.array([start_xy[0], start_xy[1], np.deg2rad(float(row["start_deg"])), 0.0, 0.0, 0.0])
    goal = np.array([goal_xy[0], goal_xy[1], 0.0, 0.0, 0.0, 0.0])
    return start, goal


def result_row(run_idx, path, actions, results, runtime):
    """The reference's CSV row for one run (run_scenarios.py:345-395)."""
    if path is not None:
        length = float(np.sum(np.linalg.norm(np.diff(path[:, :2], axis=0), axis=1)))
        vel = float(np.mean(np.sqrt(np.square(path[:, 2]) + np.square(path[:, 3]))))
        effort = np.linalg.norm(actions, axis=1)
        return [run_idx + 1, 1, runtime, length, results.get("path_time", 0.0), vel, results["number_of_nodes"],
                results["iterations"], float(effort.max()), float(effort.mean()), float(effort.std())]
    return [run_idx + 1, 0, runtime, -1, 0, -1, -1, results["iterations"], -1, -1, -1]


def run_car_unit(row, scenario_idx, run_idx, sampler, time_budget, planner_kwargs=None):
    """One (scenario, run): build env + planner exactly as the reference driver does and plan."""
    from .car_env import CarEnv
    from .planners.RRT import RRT_Planner
    maze = load_maze(row["maze_name"])
    env = CarEnv(maze_map=maze, collision_checking=False, run_type=int((planner_kwargs or {}).get("run_type", 0)))
    start, goal = scenario_states(row, env)
    kw = dict(env_id="carmaze", environment=env, sampler=sampler, prediction_type="actions", action_horizon=8,
              local_map_size=20, local_map_scale=0.2, global_map_scale=1.0, goal_conditioning_bias=0.85,
              prop_duration=[64], time_budget=time_budget, max_iter=300, verbose=False)
    kw.update(planner_kwargs or {})
    planner = RRT_Planner(start, goal, **kw)
    seed_everything(unit_seed(scenario_idx, run_idx))
    t0 = time.time()
    planner.reset()
    path, actions = planner.plan()
    return result_row(run_idx, path, actions, planner.results, time.time() - t0)


_QUEUE_CALLS = 0


def queued_units(units, world):
    """Dynamic deal: ranks pull the next unit index from one shared counter in the process group's key-value
    store (control plane only -- an integer per unit; no tensor leaves a GPU).  A rank that drew short units
    takes more of them, so the suite ends when the LAST unit ends rather than when the unluckiest static
    share does.  Results do not depend on who ran a unit: every unit is seeded from (scenario, run)."""
    global _QUEUE_CALLS
    import torch.distributed as dist
    from torch.distributed.distributed_c10d import _get_default_store
    _QUEUE_CALLS += 1  # run_suite is collective: every rank is at the same call number
    store, key = _get_default_store(), f"ditree/suite_queue/{_QUEUE_CALLS}"
    while True:
        i = store.add(key, 1) - 1
        if i >= len(units):
            return
        yield units[i]


def gather_rows(local_units, local_rows, n_units_total, device, world, per_rank=None):
    """all_gather the fixed-size result rows; returns {(scenario, run): row} on every rank."""
    import torch.distributed as dist
    if per_rank is None:
        per_rank = (n_units_total + world - 1) // world
    buf = torch.full((per_rank, 2 + len(ROW_FIELDS)), float("nan"), dtype=torch.float32, device=device)
    for i, ((s, r), row) in enumerate(zip(local_units, local_rows)):
        buf[i, 0], buf[i, 1] = s, r
        buf[i, 2:] = torch.as_tensor(row, dtype=torch.float32)
    if world > 1:
        out = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
        allbuf = torch.cat(out).cpu().numpy()
    else:
        allbuf = buf.cpu().numpy()
    table = {}
    for rec in allbuf:
        if not np.isnan(rec[0]):
            table[(int(rec[0]), int(rec[1]))] = rec[2:].tolist()
    return table


_DESC_CACHE = {}


def car_unit_descriptor(row, scenario_idx, run_idx, run_type=0):
    """What the device-resident planner needs to know about one (scenario, run): start / goal exactly as
    run_car_unit derives them (run_scenarios.py:239-246), the maze, the unit's seed and -- for run_type >= 2 -- the
    probability map the planner samples cells from (CarEnv.prob_map as of the start of plan(): the EDT prior for
    run_type 2, its blend with the start -> goal Gaussian for run_type >= 3, car_env.py:98-137, RRT.py:126-127).
    The scenario part (an env built once per row) is cached: the runs of a row differ in their seed only."""
    key = (row["maze_name"], row["start_row"], row["start_col"], row["start_deg"], row["goal_row"], row["goal_col"],
           int(run_type))
    hit = _DESC_CACHE.get(key)
    if hit is None:
        from .car_env import CarEnv
        maze = load_maze(row["maze_name"])
        env = CarEnv(maze_map=maze, collision_checking=False, run_type=int(run_type))
        start, goal = scenario_states(row, env)
        # reset() snaps the goal to its cell centre (base_planner.py:86-90 -> car_env reset)
        goal_xy = env.cell_rowcol_to_xy(env.cell_xy_to_rowcol(goal[:2]))
        hit = dict(start=np.asarray(start, dtype=np.float32), goal=np.asarray(goal_xy, dtype=np.float32),
                   maze=np.float32(maze), maze_name=row["maze_name"])
        if run_type >= 2:
            env.reset(options={"reset_cell": env.cell_xy_to_rowcol(start[:2]), "reset_deg": np.rad2deg(start[2]),
                               "goal_cell": env.cell_xy_to_rowcol(goal[:2])})
            if run_type >= 3:
                env.update_prob_map_by_loc()
            hit["prob_map"] = np.array(env.prob_map, dtype=np.float64)
            hit["prob_key"] = key
        _DESC_CACHE[key] = hit
    return dict(hit, seed=unit_seed(scenario_idx, run_idx), key=(scenario_idx, run_idx))


def run_suite_device(sampler, units_iter, rows, planner_kwargs=None, max_units=4096):
    """Run the units `units_iter` yields ((scenario, run) pairs, possibly from the shared queue) on the device-resident
    multi-scenario planner.  -> (list of (scenario, run), list of result rows, planner stats)."""
    from .planners.device_planner import DevicePlanner
    kw = dict(unit_slots=16, streams=2, iteration_cap=4096, max_units=int(max_units))
    kw.update(planner_kwargs or {})
    kw.pop("batch_size", None)
    run_type = int(kw.get("run_type", 0))
    planner = DevicePlanner(sampler, **kw)
    mine = []

    def source():
        for s, r in units_iter:
            mine.append((s, r))
            yield car_unit_descriptor(rows[s], s, r, run_type)
    try:
        recs = planner.run(source())
    finally:
        stats = dict(planner.stats)
        planner.close()
    local = [result_row(r, rec["path"], rec["actions"], rec["results"], rec["runtime"]) for (s, r), rec in zip(mine, recs)]
    stats["goal_reached"] = int(sum(rec["goal_reached"] for rec in recs))
    stats["collision_rate"] = float(sum(rec["collisions"] for rec in recs)) / max(1, sum(rec["chunks"] for rec in recs))
    return mine, local, stats


LAST_SUITE_STATS = {}


def run_suite(sampler, total_runs=1, time_budget=5.0, kind="test_scenarios_car", rank=0, world=1, device="cuda",
              planner_kwargs=None, unit_fn=run_car_unit, schedule="queue", engine="host"):
    """Run this rank's share of the suite and gather everybody's rows.  -> (table, seconds).
    schedule: "queue" (default for world > 1) -- ranks pull units from a shared counter, heaviest maps first;
    "static" -- the round-robin deal of shard_units.
    engine: "host" -- one unit at a time through `unit_fn` (RRT_Planner.plan); "device" -- the device-resident
    multi-scenario planner (planners/device_planner.py): several units share every device pass, the host only feeds
    the device queue (from the shared counter when schedule == "queue")."""
    rows = load_scenarios(kind)
    weights = [int(np.prod(load_maze(r["maze_name"]).shape)) for r in rows]
    n_total = len(rows) * total_runs
    t0 = time.time()
    queue = world > 1 and schedule == "queue"
    if engine == "device":
        units = queued_units(all_units(len(rows), total_runs, weights, by_weight=True), world) if queue else \
            shard_units(len(rows), total_runs, rank, world, weights)
        mine, local, stats = run_suite_device(sampler, units, rows, planner_kwargs, max_units=n_total + 1)
        LAST_SUITE_STATS.clear()
        LAST_SUITE_STATS.update(stats)
        table = gather_rows(mine, local, n_total, device, world, per_rank=n_total if queue else None)
    elif queue:
        mine, local = [], []
        for s, r in queued_units(all_units(len(rows), total_runs, weights, by_weight=True), world):
            mine.append((s, r))
            local.append(unit_fn(rows[s], s, r, sampler, time_budget, planner_kwargs))
        table = gather_rows(mine, local, n_total, device, world, per_rank=n_total)
    else:
        mine = shard_units(len(rows), total_runs, rank, world, weights)
        local = [unit_fn(rows[s], s, r, sampler, time_budget, planner_kwargs) for s, r in mine]
        table = gather_rows(mine, local, n_total, device, world)
    return table, time.time() - t0


# ---------------------------------------------------------------------------------------------
# wire formats
# ---------------------------------------------------------------------------------------------
# The reference's per-scenario CSV (run_scenarios.py:331-333,392-395): the header names TEN columns but every
# row carries ELEVEN values (trajectory_time sits between trajectory_length and avg_velocity without a
# header cell).  results_process.py reads these files with pandas, so the quirk is kept byte for byte.
CSV_HEADER = ["iteration", "success", "runtime", "trajectory_length", "avg_velocity", "num_states_in_tree",
              "num_RRT_iterations", "ctrl_effort_max", "ctrl_effort_mean", "ctrl_effort_std"]


def scenario_csv_path(root, current_run, scenario_name, planner_name="diffusion_RRT_PD64", env_id="carmaze"):
    """benchmark_results/{current_run}/{scenario}_{planner}_{env_id}.csv (run_scenarios.py:306)."""
    import os
    return os.path.join(root, "benchmark_results", str(current_run), f"{scenario_name}_{planner_name}_{env_id}.csv")


def next_run_index(root):
    """The reference numbers result directories 1, 2, ... (run_scenarios.py:191-195)."""
    import os
    d = os.path.join(root, "benchmark_results")
    os.makedirs(d, exist_ok=True)
    taken = [int(x) for x in os.listdir(d) if x.isdigit()]
    return max(taken) + 1 if taken else 1


def existing_rows(path):
    """Rows already present (header excluded): the reference resumes a scenario file where it stopped
    (run_scenarios.py:308-323)."""
    import csv
    import os
    if not os.path.exists(path):
        return 0
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    return len(rows) - 1 if len(rows) > 1 else 0


def write_suite_csv(table, root, current_run=None, kind="test_scenarios_car", planner_name="diffusion_RRT_PD64",
                    env_id="carmaze"):
    """Write the gathered result table ({(scenario, run): row}) as the reference's per-scenario CSV files.
    Integer-valued fields the reference writes as Python ints (iteration, success, the -1 / 0 sentinels,
    node and iteration counts) are written as ints.  -> list of paths."""
    import csv
    import os
    rows = load_scenarios(kind)
    current_run = next_run_index(root) if current_run is None else current_run
    paths = []
    for s_idx, sc_row in enumerate(rows):
        mine = sorted((r, v) for (s, r), v in table.items() if s == s_idx)
        if not mine:
            continue
        path = scenario_csv_path(root, current_run, sc_row["scenario_name"], planner_name, env_id)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        have = existing_rows(path)
        if have == 0:
            with open(path, "w", newline="") as f:
                csv.writer(f).writerow(CSV_HEADER)
        with open(path, "a", newline="") as f:
            w = csv.writer(f)
            for r, v in mine:
                if r < have:
                    continue
                it, ok, runtime, length, ttime, vel, nodes, iters, emax, emean, estd = v
                ok = int(ok)
                sent = (lambda x: int(x) if float(x) in (-1.0, 0.0) else float(x))
                w.writerow([int(it), ok, float(runtime), float(length) if ok == 1 else sent(length),
                            float(ttime) if ok == 1 else sent(ttime), float(vel) if ok == 1 else sent(vel),
                            int(nodes), int(iters), float(emax) if ok == 1 else sent(emax),
                            float(emean) if ok == 1 else sent(emean), float(estd) if ok == 1 else sent(estd)])
        paths.append(path)
    return paths


def save_path_csv(path_array, filename):
    """np.savetxt(f'path_DP_{i}.csv', path_array, fmt='%.6f', delimiter=',') (run_scenarios.py:352)."""
    np.savetxt(filename, np.asarray(path_array), fmt="%.6f", delimiter=",")


def load_checkpoint_state_dict(path, map_location="cpu"):
    """The reference's training checkpoint (run_scenarios.py:175-176): a torch file whose
    'noise_pred_net_state_dict' entry is the ConditionalUnet1DWithLocalMap state_dict; its keys are exactly the
    names ``dt_load_denoiser`` expects, so the result feeds ``DiffusionSampler`` / ``Context.load_denoiser``
    unchanged.  A bare state_dict file is accepted too."""
    ck = torch.load(path, map_location=map_location, weights_only=True)
    sd = ck.get("noise_pred_net_state_dict", ck) if isinstance(ck, dict) else ck
    if not isinstance(sd, dict) or not any(k.startswith("unet.") for k in sd):
        raise KeyError("checkpoint holds no 'noise_pred_net_state_dict' of the reference architecture")
    return {k: v.float() for k, v in sd.items()}
