#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ and the data fixtures of the package by
running the UNMODIFIED reference (mounted read-only at /root/reference) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tools/gen_golden.py

The reference needs seven third-party packages that are absent here; tools/ref_stubs/ holds
import-only stubs for them (two carry arithmetic the reference delegates to them: casadi's MX.tanh
-> numpy tanh, gymnasium.spaces.Box -> float32 clip bounds).  Nothing from the reference is copied
into the repository: only inputs, outputs and data tables (maze grids, scenario rows, normaliser
statistics) are stored.
"""
import json
import os
import random
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DITREE_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REPO, "tools", "ref_stubs"))
sys.path.insert(0, REF)
sys.path.insert(0, REPO)
os.chdir(REF)  # the reference resolves metadata/{env}.pt relative to the CWD

import torch  # noqa: E402

import car_env  # noqa: E402
import common.map_utils as mu  # noqa: E402
from common.fm_utils import get_timesteps  # noqa: E402
from lidar_sim.lidar_2d_sim import Lidar2DSim  # noqa: E402
from local_map_encoder import ConditionalUnet1DWithLocalMap  # noqa: E402
from planners.RRT import RRT_Planner  # noqa: E402
from policies.fm_policy import DiffusionSampler  # noqa: E402

from oracle import denoiser_ref  # noqa: E402  (only its seeded weight initialiser is used here)

GOLD = os.path.join(REPO, "tests", "golden")
DATA = os.path.join(REPO, "ditreeonlineplanner_b200", "data")
os.makedirs(GOLD, exist_ok=True)
os.makedirs(DATA, exist_ok=True)

MAZES = ["Race_Track", "boxes", "narrow_short", "random_huge", "random_large", "random_xlarge", "shapes",
         "val_maze_10", "val_maze_15", "val_maze_7"]


def load_maze(name):
    return np.loadtxt(f"maps/mazes/{name}.csv", delimiter=",")


def save(name, **arrays):
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **arrays)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


# ---------------------------------------------------------------------------------------------
def gen_data_fixtures():
    mazes = {m: load_maze(m).astype(np.uint8) for m in MAZES}
    np.savez_compressed(os.path.join(DATA, "mazes.npz"), **mazes)
    scen = {}
    for kind in ("test_scenarios_car", "validation_scenarios_car", "test_scenarios_ant", "validation_scenarios_ant"):
        rows = []
        with open(f"experiments/{kind}.csv") as f:
            header = f.readline().strip().split(",")
            for line in f:
                if line.strip():
                    rows.append(dict(zip(header, line.strip().split(","))))
        scen[kind] = rows
    with open(os.path.join(DATA, "scenarios.json"), "w") as f:
        json.dump(scen, f, indent=1)
    for env in ("carmaze", "antmaze"):
        md = torch.load(f"metadata/{env}.pt", weights_only=False)
        np.savez(os.path.join(DATA, f"metadata_{env}.npz"), **{k: np.asarray(v, dtype=np.float64) for k, v in md.items()})
    print("  wrote data fixtures")


def gen_schedule():
    out = {}
    for k in (1, 2, 5, 10):
        t0, dt = get_timesteps("exp", k, exp_scale=4.0)
        out[f"t0_{k}"] = t0.numpy()
        out[f"dt_{k}"] = dt.numpy()
    save("schedule.npz", **out)


def gen_local_map():
    rng = np.random.default_rng(101)
    out = {}
    for m in MAZES:
        g = load_maze(m)
        R, C = g.shape
        for tag, n, scale, sg in (("car", 20, 0.2, 1.0), ("ant", 16, 0.8, 4.0)):
            k = 48
            x = rng.uniform(-C / 2 * sg - 1, C / 2 * sg + 1, k)
            y = rng.uniform(-R / 2 * sg - 1, R / 2 * sg + 1, k)
            th = rng.uniform(-2 * np.pi, 2 * np.pi, k)
            x, y, th = (v.astype(np.float32).astype(np.float64) for v in (x, y, th))
            lm = mu.create_local_map(np.float32(g), x, y, th, n, scale, sg, (C / 2 * sg, R / 2 * sg))
            out[f"{m}.{tag}.pose"] = np.stack([x, y, th], 1).astype(np.float32)
            out[f"{m}.{tag}.map"] = lm.astype(np.uint8)
    # scalar call path (K = 1 via python floats)
    g = load_maze("boxes")
    out["scalar.map"] = mu.create_local_map(np.float32(g), 1.25, -3.5, 0.7, 20, 0.2, 1.0, (10.0, 10.0)).astype(np.uint8)
    save("local_map.npz", **out)


def gen_collide_car():
    rng = np.random.default_rng(202)
    out = {}
    for m in MAZES:
        g = np.float32(load_maze(m))
        R, C = g.shape
        n = 4000
        x = rng.uniform(-C / 2 - 0.3, C / 2 + 0.3, n)
        y = rng.uniform(-R / 2 - 0.3, R / 2 + 0.3, n)
        # half the samples are snapped close to cell edges / corners where the tests bite
        snap = rng.random(n) < 0.5
        x = np.where(snap, np.round(x) + rng.normal(0, 0.12, n), x)
        snap2 = rng.random(n) < 0.5
        y = np.where(snap2, np.round(y) + rng.normal(0, 0.12, n), y)
        th = rng.uniform(-np.pi, np.pi, n)
        st = np.stack([x, y, th], 1).astype(np.float32)
        flags = np.array([mu.is_colliding_car(s.astype(np.float64), g) for s in st])
        out[f"{m}.states"] = st
        out[f"{m}.flags"] = np.packbits(flags)
        out[f"{m}.n"] = np.array(n)
    # the documented corner cases of is_colliding_parallel (single points, r = 0.1, scale 1)
    g = np.float32(load_maze("random_large"))
    out["quirk.random_large"] = np.array(mu.is_colliding_parallel(
        np.array([3.9719289005037552, 0.5062825501784967]), g))
    g = np.float32(load_maze("narrow_short"))
    out["quirk.narrow_short"] = np.array(mu.is_colliding_parallel(
        np.array([2.4975924901541786, -1.0489708935371076]), g))
    free = np.zeros((5, 5), np.float32)
    out["quirk.border"] = np.array([mu.is_colliding_parallel(np.array([0.0, 2.0]), free)[0],
                                    mu.is_colliding_parallel(np.array([0.0, 0.0]), free)[0]])
    wall = np.ones((5, 5), np.float32)
    out["quirk.batch_oob"] = mu.is_colliding_parallel(np.array([[0.0, 0.0], [9.0, 0.0]]), wall)
    # point batches through is_colliding_parallel itself (batch early-return semantics included)
    g = np.float32(load_maze("boxes"))
    pts = np.stack([rng.uniform(-9.4, 9.4, 3000), rng.uniform(-9.4, 9.4, 3000)], 1).astype(np.float32)
    out["points.boxes.pts"] = pts
    out["points.boxes.flags"] = mu.is_colliding_parallel(pts.astype(np.float64), g)
    save("collide_car.npz", **out)


def gen_collide_ant():
    rng = np.random.default_rng(303)
    out = {}
    g = np.zeros((5, 5), np.float32)
    g[2, 2] = 1
    cases = np.zeros((6, 29))
    cases[:, 3] = 1.0  # slot 3 is read as w
    cases[0, :2] = (0, 0)
    cases[1, :2] = (1.5, 0)
    cases[2, :2] = (2.5, 0)
    cases[3, :2] = (6.0, 6.0)
    cases[4, 3:7] = (0, 1, 0, 0)  # upside-down
    cases[5, :2] = (-9.9, 0.0)
    out["small.states"] = cases
    out["small.flags"] = np.array([mu.is_colliding_ant(c, g, 1.2, 4.0) for c in cases])
    big = np.float32(load_maze("random_huge"))
    n = 1500
    st = np.zeros((n, 29))
    st[:, 0] = rng.uniform(-64, 64, n)
    st[:, 1] = rng.uniform(-64, 64, n)
    st[:, 2] = 0.75
    q = rng.normal(size=(n, 4))
    q[: n // 2] = np.array([1.0, 0, 0, 0]) + 0.3 * rng.normal(size=(n // 2, 4))
    st[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    st = st.astype(np.float32).astype(np.float64)
    out["huge.states"] = st[:, :7].astype(np.float32)
    out["huge.flags"] = np.array([mu.is_colliding_ant(s, big, 1.2, 4.0) for s in st])
    save("collide_ant.npz", **out)


def make_env(maze, goal_cell=None, start_cell=(1, 1)):
    env = car_env.CarEnv(maze_map=maze, collision_checking=False)
    env.reset(options={"reset_cell": np.array(start_cell), "reset_deg": 0.0,
                       "goal_cell": np.array(goal_cell) if goal_cell is not None else None})
    return env


def gen_bicycle():
    rng = np.random.default_rng(404)
    maze = load_maze("boxes")
    env = make_env(maze, goal_cell=(1, 1))
    env.goal = np.array([100.0, 100.0])  # far away: pure dynamics
    n, S = 96, 50
    s0 = np.stack([rng.uniform(-8, 8, n), rng.uniform(-8, 8, n), rng.uniform(-np.pi, np.pi, n),
                   rng.uniform(0, 4, n), rng.uniform(0, 1.3, n), rng.uniform(-0.44, 0.44, n)], 1)
    act = np.stack([rng.normal(0.45, 1.0, (n, S)), rng.normal(0, 0.92, (n, S))], -1)
    act[::7] *= 6.0  # exercise the clip
    s0 = s0.astype(np.float32).astype(np.float64)
    act = act.astype(np.float32).astype(np.float64)
    traj = np.zeros((n, S, 6))
    for b in range(n):
        env.done = False
        env.terminated = False
        env.set_state(s0[b].copy())
        for i in range(S):
            traj[b, i] = env.step(act[b, i])[0]
    # goal latch: once within 0.5 m the state freezes
    env.goal = np.array([0.3, 0.0])
    env.done = False
    env.set_state(np.array([-1.0, 0.0, 0.0, 3.0, 0.5, 0.0]))
    latch = np.array([env.step(np.array([0.0, 0.0]))[0] for _ in range(30)])
    succ = []
    env.done = False
    env.set_state(np.array([-1.0, 0.0, 0.0, 3.0, 0.5, 0.0]))
    for _ in range(30):
        succ.append(env.step(np.array([0.0, 0.0]))[4]["success"])
    save("bicycle.npz", s0=s0.astype(np.float32), act=act.astype(np.float32), traj=traj,
         latch=latch, latch_success=np.array(succ))


def make_planner(maze, start, goal, **kw):
    env = car_env.CarEnv(maze_map=maze, collision_checking=False)
    return RRT_Planner(start, goal, env_id="carmaze", environment=env, sampler=kw.pop("sampler", None),
                       action_horizon=kw.pop("action_horizon", 8), local_map_size=20, local_map_scale=0.2,
                       global_map_scale=1.0, time_budget=kw.pop("time_budget", 5), **kw)


def gen_propagate():
    rng = np.random.default_rng(505)
    maze = load_maze("boxes")
    env0 = car_env.CarEnv(maze_map=maze, collision_checking=False)
    start_xy = env0.cell_rowcol_to_xy(np.array([17, 2]))
    goal_xy = env0.cell_rowcol_to_xy(np.array([2, 17]))
    start = np.array([start_xy[0], start_xy[1], np.deg2rad(45.0), 0, 0, 0])
    goal = np.array([goal_xy[0], goal_xy[1], 0, 0, 0, 0])
    out = {"goal_xy": goal_xy, "n_cases": np.array(0)}
    cases = []
    # (state, actions): free run, wall hit, goal reach
    cases.append((start.copy(), np.tile(np.array([[2.0, 0.1]]), (8, 1))))
    wall_xy = env0.cell_rowcol_to_xy(np.array([17, 1]))  # free cell next to the left border wall
    cases.append((np.array([wall_xy[0], wall_xy[1], np.pi, 3.5, 1.0, 0.0]), np.tile(np.array([[5.0, 0.0]]), (8, 1))))
    cases.append((np.array([goal_xy[0] - 0.63, goal_xy[1], 0.0, 3.0, 0.5, 0.0]), np.tile(np.array([[0.0, 0.0]]), (8, 1))))
    for _ in range(40):
        cell = None
        while cell is None:
            r, c = rng.integers(1, 19, 2)
            if maze[r, c] == 0:
                cell = (r, c)
        xy = env0.cell_rowcol_to_xy(np.array(cell)) + rng.uniform(-0.3, 0.3, 2)
        st = np.array([xy[0], xy[1], rng.uniform(-np.pi, np.pi), rng.uniform(1, 4.5), rng.uniform(0, 1.3),
                       rng.uniform(-0.4, 0.4)])
        cases.append((st, np.stack([rng.normal(0.45, 1.0, 8), rng.normal(0, 0.9, 8)], 1)))
    for i, (st, act) in enumerate(cases):
        st = st.astype(np.float32).astype(np.float64)
        act = act.astype(np.float32).astype(np.float64)
        pl = make_planner(maze, start, goal)
        obs, done, a, s = pl.propagate_action_sequence_env(st.copy(), act.copy())
        out[f"{i}.state"] = st
        out[f"{i}.act"] = act
        out[f"{i}.obs"] = obs
        out[f"{i}.done"] = np.array(-1 if done is None else int(done))
        out[f"{i}.a"] = np.asarray(a)
        out[f"{i}.s"] = np.asarray(s)
    out["n_cases"] = np.array(len(cases))
    save("propagate.npz", **out)


class _Capture(torch.nn.Module):
    """Stands in for the denoiser: records what the reference sampler feeds it."""
    def forward(self, sample, local_map, timestep, global_cond):
        self.seen = dict(sample=sample.clone(), local_map=local_map.clone(), timestep=timestep.clone(),
                         global_cond=global_cond.clone())
        return torch.zeros_like(sample)


def gen_cond():
    rng = np.random.default_rng(606)
    out = {}
    cap = _Capture()
    smp = DiffusionSampler(cap, None, "carmaze", policy="flow_matching", pred_horizon=64, action_dim=2,
                           obs_history=1, action_history=1, goal_conditioned=True, num_diffusion_iters=1,
                           local_map_size=20)
    smp.device = "cpu"
    B = 32
    obs = np.stack([rng.uniform(-8, 8, B), rng.uniform(-8, 8, B), rng.uniform(-4, 4, B), rng.uniform(0, 4, B),
                    rng.uniform(0, 1.3, B), rng.uniform(-0.44, 0.44, B)], 1).astype(np.float32).astype(np.float64)
    prev = np.stack([rng.normal(0.45, 1, (B, 8)), rng.normal(0, 0.9, (B, 8))], -1).astype(np.float32).astype(np.float64)
    goal = np.array([4.0, 3.0])
    lm = (rng.random((B, 20, 20)) < 0.3).astype(np.float32)
    torch.manual_seed(7)
    res = smp(obs[:, None, :], prev_actions=prev, goal=goal, local_map=lm)
    out["car.obs"], out["car.prev"], out["car.goal"] = obs, prev, goal
    out["car.cond"] = cap.seen["global_cond"].numpy()
    out["car.map_in"] = cap.seen["local_map"].numpy()
    out["car.result"] = res  # zero velocity: un-normalised noise
    torch.manual_seed(7)
    out["car.noise"] = torch.randn(B, 64, 2).numpy()
    res = smp(obs[:1, None, :], prev_actions=None, goal=goal, local_map=lm[:1])
    out["car.cond_noprev"] = cap.seen["global_cond"].numpy()
    goals = np.stack([rng.uniform(-8, 8, B), rng.uniform(-8, 8, B)], 1)
    smp(obs[:, None, :], prev_actions=prev, goal=goals, local_map=lm)
    out["car.goals"] = goals
    out["car.cond_goals"] = cap.seen["global_cond"].numpy()
    # documented example of SURVEY A.6
    smp(np.array([[[1, -2, 0.5, 2, 0.6, 0.1]]], dtype=np.float64), prev_actions=None, goal=np.array([4.0, 3.0]),
        local_map=lm[:1])
    out["car.cond_example"] = cap.seen["global_cond"].numpy()

    smp_a = DiffusionSampler(cap, None, "antmaze", policy="flow_matching", pred_horizon=16, action_dim=8,
                             obs_history=3, action_history=1, goal_conditioned=True, num_diffusion_iters=1,
                             local_map_size=16)
    smp_a.device = "cpu"
    md = smp_a.metadata
    Ba = 16
    for h in (1, 3):
        o = np.zeros((Ba, h, 29))
        o[..., :2] = rng.uniform(-30, 30, (Ba, h, 2))
        o[..., 2:] = md["Observations_mean"] + md["Observations_std"] * rng.normal(size=(Ba, h, 27))
        q = rng.normal(size=(Ba, h, 4))
        o[..., 3:7] = q / np.linalg.norm(q, axis=-1, keepdims=True)
        o = o.astype(np.float32).astype(np.float64)
        pa = (md["Actions_mean"] + md["Actions_std"] * rng.normal(size=(Ba, 2, 8))).astype(np.float32).astype(np.float64)
        ga = rng.uniform(-30, 30, 2)
        lma = (rng.random((Ba, 16, 16)) < 0.3).astype(np.float32)
        smp_a(o, prev_actions=pa, goal=ga, local_map=lma)
        out[f"ant.h{h}.obs"], out[f"ant.h{h}.prev"], out[f"ant.h{h}.goal"] = o, pa, ga
        out[f"ant.h{h}.cond"] = cap.seen["global_cond"].numpy()
    save("cond.npz", **out)


def gen_denoiser(cases=None):
    """`large_b64_k{1,10}`: SURVEY 8(c) item 7 -- the reference DiffusionSampler.forward of the `large` net at
    B = 64 (whole samples fill the GEMM tiles: the fused-GroupNorm CTA-pair kernels, not the split-K path)."""
    cases = cases or (("small", [64, 128, 256], 3, 3, 11), ("large", [512, 1024, 2048], 1, 1, 12),
                      ("large_b64_k1", [512, 1024, 2048], 64, 1, 13), ("large_b64_k10", [512, 1024, 2048], 64, 10, 14))
    for tag, dims, B, K, seed in cases:
        sd = denoiser_ref.init_params(seed=seed, input_dim=2, cond_dim=7, emb_dim=400, down_dims=dims)
        net = ConditionalUnet1DWithLocalMap(input_dim=2, encoder_name="resnet", embedding_dim=400,
                                            additional_global_cond_dim=7, local_map_size=20, down_dims=dims)
        missing = net.load_state_dict(sd, strict=True)
        net.eval()
        g = torch.Generator().manual_seed(seed + 100)
        lm01 = (torch.rand(B, 20, 20, generator=g) < 0.35).float()
        cond = torch.randn(B, 7, generator=g) * 0.5
        sample = torch.randn(B, 64, 2, generator=g)
        ts = torch.full((B,), 6.7166)
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            enc = net.encoder(lm01 * 2 - 1)
            vel = net(sample=sample, local_map=lm01 * 2 - 1, timestep=ts, global_cond=cond)
        # full sampler: K Euler steps through the reference DiffusionSampler
        smp = DiffusionSampler(net, None, "carmaze", policy="flow_matching", pred_horizon=64, action_dim=2,
                               obs_history=1, action_history=1, goal_conditioned=True, num_diffusion_iters=K,
                               local_map_size=20)
        smp.device = "cpu"
        rng = np.random.default_rng(seed)
        obs = np.stack([rng.uniform(-8, 8, B), rng.uniform(-8, 8, B), rng.uniform(-3, 3, B), rng.uniform(0, 4, B),
                        rng.uniform(0, 1.3, B), rng.uniform(-0.44, 0.44, B)], 1).astype(np.float32).astype(np.float64)
        prev = np.stack([rng.normal(0.45, 1, (B, 8)), rng.normal(0, 0.9, (B, 8))], -1).astype(np.float32).astype(np.float64)
        goal = np.array([4.0, 3.0])
        torch.manual_seed(seed + 200)
        actions = smp(obs[:, None, :], prev_actions=prev, goal=goal, local_map=lm01.numpy())
        torch.manual_seed(seed + 200)
        noise = torch.randn(B, 64, 2)
        checksum = float(sum(float(v.double().sum()) for v in sd.values()))
        save(f"denoiser_{tag}.npz", seed=np.array(seed), dims=np.array(dims), K=np.array(K), lm01=lm01.numpy(),
             cond=cond.numpy(), sample=sample.numpy(), ts=ts.numpy(), enc=enc.numpy(), vel=vel.numpy(),
             obs=obs, prev=prev, goal=goal, noise=noise.numpy(), actions=actions, weight_checksum=np.array(checksum))


def insert_box(maze, row, col, h, w):
    m = maze.copy()
    m[row:row + h, col:col + w] = 1
    return m


def gen_lidar():
    rng = np.random.default_rng(707)
    base = load_maze("boxes")
    mazes = [base, insert_box(base, 10, 15, 1, 4), insert_box(base, 9, 16, 2, 2), insert_box(base, 7, 16, 2, 3),
             insert_box(base, 1, 7, 4, 4)]
    lidar = Lidar2DSim()
    out = {}
    idx = 0
    for mi, m in enumerate(mazes):
        for _ in range(4):
            while True:
                x, y = rng.uniform(1, 19, 2)
                if m[int(y), int(x)] == 0:
                    break
            pose = np.array([x, y, rng.uniform(-np.pi, np.pi)]).astype(np.float32).astype(np.float64)
            np.random.seed(0)
            d, e, v = lidar.scan(pose, m)
            out[f"{idx}.maze"] = np.array(mi)
            out[f"{idx}.pose"] = pose
            out[f"{idx}.dist"] = d
            out[f"{idx}.end"] = e
            out[f"{idx}.visited"] = v.astype(np.int16)
            idx += 1
    out["n"] = np.array(idx)
    for mi, m in enumerate(mazes):
        out[f"maze{mi}"] = m.astype(np.uint8)
    save("lidar.npz", **out)


def gen_probe():
    rng = np.random.default_rng(808)
    maze = load_maze("boxes")
    env0 = car_env.CarEnv(maze_map=maze, collision_checking=False)
    start = np.array([*env0.cell_rowcol_to_xy(np.array([17, 2])), 0.0, 0, 0, 0])
    goal = np.array([*env0.cell_rowcol_to_xy(np.array([2, 17])), 0.0, 0, 0, 0])
    pl = make_planner(maze, start, goal)
    n = 400
    st = np.stack([rng.uniform(-9.5, 9.5, n), rng.uniform(-9.5, 9.5, n), rng.uniform(-2 * np.pi, 2 * np.pi, n)], 1)
    st = st.astype(np.float32).astype(np.float64)
    flags = np.array([pl.check_obstacle_ahead(np.concatenate([s, np.zeros(3)])) for s in st])
    save("probe.npz", states=st.astype(np.float32), flags=flags)


class _FakeClock:
    """Deterministic stand-in for the `time` module inside planners.RRT."""
    def __init__(self, step):
        self.t, self.step = 0.0, step

    def time(self):
        self.t += self.step
        return self.t


def _reference_functions(path, names):
    """The named top-level functions of a reference script whose module cannot be imported here (it imports
    absent packages at module level): their source segments are exec'ed unmodified, with NumPy and a stub
    `plt` as their globals."""
    import ast
    import types
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "plt": types.SimpleNamespace()}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


def gen_online():
    """check_no_obstacles_in_path / scan_and_update_maze (run_scenarios_with_lidar_DiTree.py:112-127,158-181)
    run from their own source on seeded paths, scanned maps and poses."""
    check_path, scan_update = _reference_functions(os.path.join(REF, "run_scenarios_with_lidar_DiTree.py"),
                                                   ["check_no_obstacles_in_path", "scan_and_update_maze"])
    import types
    rng = np.random.default_rng(1212)
    out = {}
    n_cases = 0
    for m in ("boxes", "random_large", "shapes", "val_maze_10"):
        maze = load_maze(m)
        R, C = maze.shape
        env = car_env.CarEnv(maze_map=maze.copy(), collision_checking=False)
        planner = types.SimpleNamespace(env=env)
        for _ in range(12):
            scanned = rng.choice([0.0, 1.0, 2.0], size=maze.shape, p=[0.55, 0.2, 0.25])
            n = int(rng.integers(1, 60))
            # a smooth-ish path inside the map (points outside would index from the end, like NumPy does)
            p0 = np.array([rng.uniform(-C / 2 + 0.6, C / 2 - 0.6), rng.uniform(-R / 2 + 0.6, R / 2 - 0.6)])
            steps = rng.normal(0, 0.25, (n, 2)).cumsum(0)
            xy = np.clip(p0 + steps, [-C / 2 + 0.01, -R / 2 + 0.01], [C / 2 - 0.01, R / 2 - 0.01])
            path = np.concatenate([xy, rng.normal(0, 1, (n, 4))], 1).astype(np.float32).astype(np.float64)
            if n_cases % 5 == 4:   # some clear paths
                scanned[scanned == 1] = 2
            out[f"path{n_cases}.maze"] = np.array(MAZES.index(m))
            out[f"path{n_cases}.scanned"] = scanned.astype(np.uint8)
            out[f"path{n_cases}.path"] = path
            out[f"path{n_cases}.idx"] = np.array(check_path(planner, scanned, path))
            n_cases += 1
    out["n_path"] = np.array(n_cases)
    # scan_and_update_maze: lidar scan from the car's state, write-back into the known map and the scanned map
    maze = load_maze("boxes")
    n_scan = 0
    for (r0, c0, h, w) in ((10, 15, 1, 4), (9, 16, 2, 2), (7, 16, 2, 3), (1, 7, 4, 4)):
        with_obs = insert_box(maze, r0, c0, h, w)
        free = np.argwhere(with_obs == 0)
        for _ in range(3):
            env = car_env.CarEnv(maze_map=maze.copy(), collision_checking=False)
            env.lidar2dsim = Lidar2DSim(noise_std=0.0)
            cell = free[rng.integers(len(free))]
            xy = env.cell_rowcol_to_xy(cell) + rng.uniform(-0.3, 0.3, 2)
            st = np.array([xy[0], xy[1], rng.uniform(-np.pi, np.pi), 1.0, 0.5, 0.0]).astype(np.float32).astype(np.float64)
            env.set_state(st.copy())
            known = maze.copy()
            scanned = np.zeros_like(maze)
            updates = []
            planner = types.SimpleNamespace(env=env, update_maze=lambda mz: updates.append(mz.copy()))
            scan_update(planner, known, with_obs, scanned)
            out[f"scan{n_scan}.state"] = st
            out[f"scan{n_scan}.with_obs"] = with_obs.astype(np.uint8)
            out[f"scan{n_scan}.known"] = known.astype(np.uint8)
            out[f"scan{n_scan}.scanned"] = scanned.astype(np.uint8)
            assert len(updates) == 1 and np.array_equal(updates[0], known)
            n_scan += 1
    out["n_scan"] = np.array(n_scan)
    save("online.npz", **out)


def gen_tree():
    """Whole-tree record at B=1 under a fake clock with the small denoiser, plus every sampler
    call's inputs and outputs so the tree can be replayed teacher-forced."""
    import planners.RRT as rrt_mod
    import planners.base_planner as bp_mod
    dims = [64, 128, 256]
    sd = denoiser_ref.init_params(seed=21, input_dim=2, cond_dim=7, emb_dim=400, down_dims=dims)
    net = ConditionalUnet1DWithLocalMap(input_dim=2, encoder_name="resnet", embedding_dim=400,
                                        additional_global_cond_dim=7, local_map_size=20, down_dims=dims)
    net.load_state_dict(sd, strict=True)
    net.eval()
    smp = DiffusionSampler(net, None, "carmaze", policy="flow_matching", pred_horizon=64, action_dim=2,
                           obs_history=1, action_history=1, goal_conditioned=True, num_diffusion_iters=1,
                           local_map_size=20).eval()
    smp.device = "cpu"
    calls = []
    orig_forward = smp.forward

    def recording_forward(obs_seq, prev_actions, goal=None, local_map=None):
        state = torch.get_rng_state()
        res = orig_forward(obs_seq, prev_actions, goal=goal, local_map=local_map)
        torch.set_rng_state(state)
        noise = torch.randn(1, 64, 2)
        calls.append(dict(obs=np.array(obs_seq, dtype=np.float64)[0, -1], noise=noise.numpy()[0],
                          prev=None if prev_actions is None else np.array(prev_actions)[-1],
                          goal=np.array(goal, dtype=np.float64), actions=res[0]))
        return res
    smp.forward = recording_forward
    maze = load_maze("random_large")
    env = car_env.CarEnv(maze_map=maze, collision_checking=False)
    start_xy = env.cell_rowcol_to_xy(np.array([1, 3]))
    goal_xy = env.cell_rowcol_to_xy(np.array([7, 10]))
    start = np.array([start_xy[0], start_xy[1], 0.0, 0.0, 0.0, 0.0])
    goal = np.array([goal_xy[0], goal_xy[1], 0.0, 0.0, 0.0, 0.0])
    torch.manual_seed(42)
    np.random.seed(42)
    random.seed(42)
    clock = _FakeClock(0.1)
    rrt_mod.time = clock
    bp_mod.time = clock
    pl = RRT_Planner(start, goal, env_id="carmaze", environment=env, sampler=smp, prediction_type="actions",
                     action_horizon=8, local_map_size=20, local_map_scale=0.2, global_map_scale=1.0,
                     goal_conditioning_bias=0.85, prop_duration=[64], time_budget=40, max_iter=300, verbose=False)
    pl.reset()
    path, actions = pl.plan()
    nodes = pl.node_list
    parent = np.array([-1 if n.parent is None else nodes.index(n.parent) for n in nodes])
    states = np.array([n.state for n in nodes])
    visits = np.array([n.num_visit for n in nodes])
    edge_len = np.array([0 if n.parent_action_seq is None else len(n.parent_action_seq) for n in nodes])
    print(f"  tree: {pl.results['iterations']} iterations, {len(nodes)} nodes, {len(calls)} sampler calls, "
          f"path {'none' if path is None else path.shape}")
    save("tree.npz", start=start, goal=goal, parent=parent, states=states, visits=visits, edge_len=edge_len,
         iterations=np.array(pl.results["iterations"]), n_calls=np.array(len(calls)),
         call_obs=np.array([c["obs"] for c in calls]), call_noise=np.array([c["noise"] for c in calls]),
         call_has_prev=np.array([c["prev"] is not None for c in calls]),
         call_prev=np.array([np.zeros(2) if c["prev"] is None else c["prev"] for c in calls]),
         call_goal=np.array([c["goal"] for c in calls]), call_actions=np.array([c["actions"][:8] for c in calls]),
         path=np.zeros((0, 6), np.float32) if path is None else path,
         actions=np.zeros((0, 2), np.float32) if actions is None else actions)


def gen_tree_scenarios(time_budget=30):
    """SURVEY 8c item 10 on EVERY row of experiments/test_scenarios_car.csv: the reference planner at B = 1 under
    a fake clock (0.1 s per time.time() call, `time_budget` fake seconds) with the small denoiser, seeds as run_scenarios.py:86-90; the tree and
    the sampler's inputs / outputs per call (no noise: the replay test teacher-forces the actions)."""
    import csv
    import planners.RRT as rrt_mod
    import planners.base_planner as bp_mod
    dims = [64, 128, 256]
    sd = denoiser_ref.init_params(seed=21, input_dim=2, cond_dim=7, emb_dim=400, down_dims=dims)
    net = ConditionalUnet1DWithLocalMap(input_dim=2, encoder_name="resnet", embedding_dim=400,
                                        additional_global_cond_dim=7, local_map_size=20, down_dims=dims)
    net.load_state_dict(sd, strict=True)
    net.eval()
    with open("experiments/test_scenarios_car.csv") as f:
        rows = list(csv.DictReader(f))
    out = {"n": np.array(len(rows)), "time_budget": np.array(time_budget)}
    for k, row in enumerate(rows):
        smp = DiffusionSampler(net, None, "carmaze", policy="flow_matching", pred_horizon=64, action_dim=2,
                               obs_history=1, action_history=1, goal_conditioned=True, num_diffusion_iters=1,
                               local_map_size=20).eval()
        smp.device = "cpu"
        calls = []
        orig_forward = smp.forward

        def recording_forward(obs_seq, prev_actions, goal=None, local_map=None, _orig=orig_forward, _calls=calls):
            res = _orig(obs_seq, prev_actions, goal=goal, local_map=local_map)
            _calls.append(dict(obs=np.array(obs_seq, dtype=np.float64)[0, -1], has_prev=prev_actions is not None,
                               goal=np.array(goal, dtype=np.float64), actions=np.array(res[0][:8], dtype=np.float64)))
            return res
        smp.forward = recording_forward
        maze = load_maze(row["maze_name"])
        env = car_env.CarEnv(maze_map=maze, collision_checking=False)
        start_xy = env.cell_rowcol_to_xy(np.array([int(row["start_row"]), int(row["start_col"])]))
        goal_xy = env.cell_rowcol_to_xy(np.array([int(row["goal_row"]), int(row["goal_col"])]))
        start = np.array([start_xy[0], start_xy[1], np.deg2rad(float(row["start_deg"])), 0.0, 0.0, 0.0])
        goal = np.array([goal_xy[0], goal_xy[1], 0.0, 0.0, 0.0, 0.0])
        torch.manual_seed(42)
        np.random.seed(42)
        random.seed(42)
        clock = _FakeClock(0.1)
        rrt_mod.time = clock
        bp_mod.time = clock
        pl = RRT_Planner(start, goal, env_id="carmaze", environment=env, sampler=smp, prediction_type="actions",
                         action_horizon=8, local_map_size=20, local_map_scale=0.2, global_map_scale=1.0,
                         goal_conditioning_bias=0.85, prop_duration=[64], time_budget=time_budget, max_iter=300,
                         verbose=False)
        pl.reset()
        path, actions = pl.plan()
        nodes = pl.node_list
        print(f"  {k:2d} {row['scenario_name']:>16s}: {pl.results['iterations']} iterations, {len(nodes)} nodes, "
              f"{len(calls)} sampler calls, path {'none' if path is None else path.shape}")
        pre = f"{k}."
        out[pre + "maze"] = np.array(row["maze_name"])
        out[pre + "start"], out[pre + "goal"] = start, goal
        out[pre + "parent"] = np.array([-1 if n.parent is None else nodes.index(n.parent) for n in nodes], np.int32)
        out[pre + "states"] = np.array([n.state for n in nodes])
        out[pre + "visits"] = np.array([n.num_visit for n in nodes], np.int32)
        out[pre + "edge_len"] = np.array([0 if n.parent_action_seq is None else len(n.parent_action_seq) for n in nodes],
                                         np.int32)
        out[pre + "iterations"] = np.array(pl.results["iterations"])
        out[pre + "call_obs"] = np.array([c["obs"] for c in calls])
        out[pre + "call_has_prev"] = np.array([c["has_prev"] for c in calls])
        out[pre + "call_actions"] = np.array([c["actions"] for c in calls])
        out[pre + "call_goal"] = np.array([c["goal"] for c in calls])
        out[pre + "n_calls"] = np.array(len(calls))
        out[pre + "path"] = np.zeros((0, 6), np.float32) if path is None else path
        out[pre + "actions"] = np.zeros((0, 2), np.float32) if actions is None else actions
    save("tree_scenarios.npz", **out)


def gen_probmap():
    """run_type >= 2 sampler: SciPy EDT prior on every maze, gaussian_map + combine_log_blend for seeded
    (robot, goal) pairs on the 20 x 20 mazes, and np.random.choice draws with the uniform variates that
    produced them (RandomState.choice consumes exactly one random_sample per draw)."""
    from scipy.ndimage import distance_transform_edt
    import prob_sampling_utils as psu
    out = {}
    for m in MAZES:
        maze = load_maze(m)
        pr = distance_transform_edt(1 - maze)
        out[f"{m}.prior"] = pr / np.sum(pr)
    rng = np.random.default_rng(5)
    cases = []
    for m in ("boxes",):
        maze = load_maze(m)
        assert maze.shape == (20, 20)
        prior = out[f"{m}.prior"]
        for k in range(24):
            robot = rng.uniform(0, 20, 2)
            goal = rng.uniform(0, 20, 2) if k else robot.copy()   # k = 0: robot == goal (degenerate direction)
            pdf, _, _ = psu.gaussian_map(robot, goal)
            blend = psu.combine_log_blend(prior, pdf)
            np.random.seed(100 + k)
            st = np.random.get_state()
            idx = np.array([int(np.random.choice(blend.size, size=1, p=blend.ravel())[0]) for _ in range(64)])
            np.random.set_state(st)
            u = np.random.random_sample(64)
            cases.append((m, robot, goal, pdf, blend, idx, u))
    out["n_cases"] = np.array(len(cases))
    for i, (m, robot, goal, pdf, blend, idx, u) in enumerate(cases):
        out[f"{i}.maze"] = np.array(m)
        out[f"{i}.robot"], out[f"{i}.goal"], out[f"{i}.pdf"], out[f"{i}.blend"] = robot, goal, pdf, blend
        out[f"{i}.idx"], out[f"{i}.u"] = idx, u
    # CarEnv itself with run_type 2 / 3: constructor, maze_map setter, update_prob_map_by_loc
    maze = load_maze("boxes")
    env = car_env.CarEnv(maze_map=maze.copy(), collision_checking=False, run_type=3)
    out["env.init_prob"] = env.prob_map.copy()
    env.reset(options={"reset_cell": np.array([12, 15]), "reset_deg": 90.0, "goal_cell": np.array([2, 17])})
    env.update_prob_map_by_loc()
    out["env.loc_prob"], out["env.state"], out["env.goal"] = env.prob_map.copy(), env.state.copy(), np.asarray(env.goal, float)
    maze2 = insert_box(maze.copy(), 10, 15, 1, 4)
    env.maze_map = maze2
    out["env.setter_prob"], out["env.setter_prior"] = env.prob_map.copy(), env.prior.copy()
    env2 = car_env.CarEnv(maze_map=maze.copy(), collision_checking=False, run_type=2)
    out["env2.prob"] = env2.prob_map.copy()
    save("probmap.npz", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["data", "schedule", "local_map", "collide_car", "collide_ant", "bicycle", "propagate",
                             "cond", "denoiser", "lidar", "probe", "online", "tree", "tree_scenarios", "probmap"]
    fns = dict(data=gen_data_fixtures, schedule=gen_schedule, local_map=gen_local_map, collide_car=gen_collide_car,
               collide_ant=gen_collide_ant, bicycle=gen_bicycle, propagate=gen_propagate, cond=gen_cond,
               denoiser=gen_denoiser, lidar=gen_lidar, probe=gen_probe, online=gen_online, tree=gen_tree, tree_scenarios=gen_tree_scenarios, probmap=gen_probmap)
    # only the batched large-net cases (minutes of torch-CPU time), leaving the committed small ones untouched
    fns["denoiser_batch"] = lambda: gen_denoiser(cases=(("large_b64_k1", [512, 1024, 2048], 64, 1, 13),
                                                        ("large_b64_k10", [512, 1024, 2048], 64, 10, 14)))
    for w in which:
        print(f"[{w}]")
        fns[w]()
