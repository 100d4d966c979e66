"""Drop-in API mirrors (CarEnv, map_utils, DiffusionSampler, Lidar2DSim, RRT_Planner) on the GPU
against golden vectors of the unmodified reference, including a whole-tree replay."""
import random

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import denoiser_ref as dref
from oracle import ditree_oracle as orc

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_planner(maze, start, goal, sampler=None, **kw):
    from ditreeonlineplanner_b200.car_env import CarEnv
    from ditreeonlineplanner_b200.planners.RRT import RRT_Planner
    env = CarEnv(maze_map=maze, collision_checking=False)
    return RRT_Planner(start, goal, env_id="carmaze", environment=env, sampler=sampler, prediction_type="actions",
                       action_horizon=8, local_map_size=20, local_map_scale=0.2, global_map_scale=1.0,
                       goal_conditioning_bias=0.85, prop_duration=[64], verbose=False, **kw)


def test_map_utils_mirror(mazes):
    from ditreeonlineplanner_b200.common import map_utils as mu
    g = golden("collide_car.npz")
    grid = mazes["shapes"]
    st = g["shapes.states"].astype(np.float64)
    want = np.unpackbits(g["shapes.flags"])[: int(g["shapes.n"])].astype(bool)
    mu.cc_calls = 0
    assert mu.is_colliding_car(st[0], grid) == want[0] and isinstance(mu.is_colliding_car(st[0], grid), bool)
    assert np.array_equal(mu.is_colliding_car(st, grid), want)
    assert mu.cc_calls == 2 + len(st)
    assert np.array_equal(mu.is_colliding_parallel(g["points.boxes.pts"], mazes["boxes"]), g["points.boxes.flags"])
    gl = golden("local_map.npz")
    lm = mu.create_local_map(mazes["boxes"], 1.25, -3.5, 0.7, 20, 0.2, 1.0, (10.0, 10.0))
    assert lm.shape == (1, 20, 20) and np.array_equal(lm.astype(np.uint8), gl["scalar.map"])
    ga = golden("collide_ant.npz")
    grid5 = np.zeros((5, 5), np.float32)
    grid5[2, 2] = 1
    assert np.array_equal(mu.is_colliding_ant(ga["small.states"], grid5, 1.2, 4.0), ga["small.flags"])
    with pytest.raises(IndexError):
        mu.is_colliding_parallel(np.array([[1.5, 0.2]]), np.zeros((7, 4), np.float32))


def test_car_env_mirror(mazes):
    from ditreeonlineplanner_b200.car_env import CarEnv
    g = golden("bicycle.npz")
    env = CarEnv(maze_map=mazes["boxes"], collision_checking=True)
    obs, _ = env.reset(options={"reset_cell": np.array([17, 2]), "reset_deg": 45.0, "goal_cell": np.array([2, 17])})
    np.testing.assert_allclose(obs[:3], [-7.5, -7.5, np.pi / 4], rtol=1e-6)
    np.testing.assert_allclose(env.goal, [7.5, 7.5])
    assert np.array_equal(env.cell_xy_to_rowcol(np.array([-7.5, -7.5])), [17, 2])
    env.goal = np.array([100.0, 100.0])
    env.collision_checking = False
    env.set_state(g["s0"][0].astype(np.float64))
    for i in range(10):
        obs, reward, terminated, truncated, info = env.step(g["act"][0, i])
        assert rel(obs, g["traj"][0, i]) < 1e-4 and not terminated and truncated is False and not info["success"]
    # goal latch
    env.goal = np.array([0.3, 0.0])
    env.set_state(np.array([-1.0, 0.0, 0.0, 3.0, 0.5, 0.0]))
    succ = [env.step(np.zeros(2))[4]["success"] for _ in range(30)]
    assert succ == list(g["latch_success"])
    assert rel(env.state, g["latch"][-1]) < 1e-4
    # in-env collision terminates and freezes
    env2 = CarEnv(maze_map=mazes["boxes"], collision_checking=True)
    env2.reset(options={"reset_cell": np.array([17, 1]), "reset_deg": 180.0, "goal_cell": np.array([2, 17])})
    env2.set_state(np.array([*env2.cell_rowcol_to_xy(np.array([17, 1])), np.pi, 3.5, 1.0, 0.0]))
    hit = [env2.step(np.array([5.0, 0.0]))[2] for _ in range(12)]
    assert hit[-1] and not hit[0]


def test_propagate_conventions_mirror(mazes):
    g = golden("propagate.npz")
    start = np.array([-7.5, -7.5, np.pi / 4, 0, 0, 0])
    goal = np.array([g["goal_xy"][0], g["goal_xy"][1], 0, 0, 0, 0])
    kinds = set()
    for i in range(int(g["n_cases"])):
        pl = make_planner(mazes["boxes"], start, goal)
        obs, done, a, s = pl.propagate_action_sequence_env(g[f"{i}.state"].copy(), g[f"{i}.act"].copy())
        want = int(g[f"{i}.done"])
        kinds.add(want)
        assert (-1 if done is None else int(done)) == want
        assert a.shape == g[f"{i}.a"].shape and s.shape == g[f"{i}.s"].shape
        np.testing.assert_allclose(a, g[f"{i}.a"], rtol=1e-6, atol=1e-6)
        assert rel(obs, g[f"{i}.obs"]) < 1e-4
        if s.size:
            assert rel(s, g[f"{i}.s"]) < 1e-4
            assert np.array_equal((s == 0).all(axis=-1), (g[f"{i}.s"] == 0).all(axis=-1))
    assert kinds == {-1, 0, 1}
    with pytest.raises(ValueError):
        pl.propagate_action_sequence_env(start, None)


def test_sampler_mirror(car_meta):
    from ditreeonlineplanner_b200.policies.fm_policy import DiffusionSampler
    g = golden("denoiser_small.npz")
    sd = dref.init_params(seed=int(g["seed"]), input_dim=2, cond_dim=7, emb_dim=400, down_dims=[int(v) for v in g["dims"]])
    smp = DiffusionSampler(sd, None, "carmaze", policy="flow_matching", pred_horizon=64, action_dim=2, obs_history=1,
                           action_history=1, goal_conditioned=True, num_diffusion_iters=int(g["K"]), local_map_size=20,
                           max_batch=16).eval()
    obs = g["obs"][:, None, :]
    before = obs.copy()
    out = smp(obs, prev_actions=g["prev"], goal=g["goal"], local_map=g["lm01"], noise=torch.as_tensor(g["noise"]).cuda())
    assert out.dtype == np.float64 and out.shape == g["actions"].shape and np.array_equal(obs, before)
    assert rel(out, g["actions"]) < 2e-2
    # same seeded draw as the reference when both run on the same device type is torch.randn(B,T,A):
    torch.manual_seed(1)
    a = smp(obs, prev_actions=g["prev"], goal=g["goal"], local_map=torch.as_tensor(g["lm01"]))
    torch.manual_seed(1)
    b = smp(obs, prev_actions=g["prev"], goal=g["goal"], local_map=torch.as_tensor(g["lm01"]))
    assert np.array_equal(a, b)
    with pytest.raises(FileNotFoundError):
        DiffusionSampler(sd, None, "nosuchenv", policy="flow_matching", pred_horizon=64, action_dim=2)
    with pytest.raises(NotImplementedError):
        DiffusionSampler(sd, None, "carmaze", policy="diffusion", pred_horizon=64, action_dim=2, local_map_size=20)(
            obs, g["prev"], goal=g["goal"], local_map=g["lm01"])


def test_lidar_mirror():
    from ditreeonlineplanner_b200.lidar_sim.lidar_2d_sim import Lidar2DSim
    g = golden("lidar.npz")
    lidar = Lidar2DSim()
    for i in (0, 5, 11, 19):
        maze = g[f"maze{int(g[f'{i}.maze'])}"].astype(np.float64)
        d, e, v = lidar.scan(g[f"{i}.pose"], maze)
        np.testing.assert_allclose(d, g[f"{i}.dist"], atol=1e-5)
        np.testing.assert_allclose(e, g[f"{i}.end"], atol=1e-5)
        want = {(int(x), int(y)) for x, y in g[f"{i}.visited"]}
        assert {(int(x), int(y)) for x, y in v} == want


class _FakeClock:
    def __init__(self, step):
        self.t, self.step = 0.0, step

    def time(self):
        self.t += self.step
        return self.t


class _ReplaySampler:
    """Teacher-forcing: returns the reference's recorded actions call by call and checks that the
    planner asked for them from the same state / goal the reference did."""
    def __init__(self, g):
        self.g, self.i, self.max_obs_err = g, 0, 0.0
        self.metadata, self.num_diffusion_iters, self.pred_horizon, self.action_dim = None, 1, 64, 2

    def __call__(self, prev_states, prev_actions=None, goal=None, local_map=None):
        g, i = self.g, self.i
        assert i < int(g["n_calls"]), "planner made more sampler calls than the reference"
        obs = np.asarray(prev_states)[0, -1]
        self.max_obs_err = max(self.max_obs_err, rel(obs, g["call_obs"][i]))
        np.testing.assert_allclose(np.asarray(goal, np.float64), g["call_goal"][i], rtol=1e-6, atol=1e-6)
        assert (prev_actions is not None) == bool(g["call_has_prev"][i])
        assert tuple(local_map.shape) == (1, 20, 20)
        out = np.zeros((1, 64, 2))
        out[0, :8] = g["call_actions"][i]
        self.i += 1
        return out


def test_tree_replay_matches_reference(mazes, monkeypatch):
    """Whole-tree parity at batch_size = 1 under a fake clock: with the
    reference's sampled actions teacher-forced, the device NN / dynamics / collision kernels must
    rebuild the reference's tree: same parents, same edge lengths, node states to fp32 tolerance."""
    import ditreeonlineplanner_b200.planners.RRT as rrt_mod
    import ditreeonlineplanner_b200.planners.base_planner as bp_mod
    g = golden("tree.npz")
    clock = _FakeClock(0.1)
    monkeypatch.setattr(rrt_mod, "time", clock)
    monkeypatch.setattr(bp_mod, "time", clock)
    smp = _ReplaySampler(g)
    torch.manual_seed(42)
    np.random.seed(42)
    random.seed(42)
    pl = make_planner(mazes["random_large"], g["start"], g["goal"], sampler=smp, time_budget=40, max_iter=300)
    pl.reset()
    path, actions = pl.plan()
    assert smp.i == int(g["n_calls"]) and pl.results["iterations"] == int(g["iterations"])
    nodes = pl.node_list
    assert len(nodes) == len(g["parent"])
    parent = np.array([-1 if n.parent is None else n.parent.index for n in nodes])
    assert np.array_equal(parent, g["parent"])
    assert np.array_equal([0 if n.parent_action_seq is None else len(n.parent_action_seq) for n in nodes], g["edge_len"])
    assert np.array_equal([n.num_visit for n in nodes], g["visits"])
    assert rel(np.array([n.state for n in nodes]), g["states"]) < 1e-4 and smp.max_obs_err < 1e-4
    assert path.shape == g["path"].shape and actions.shape == g["actions"].shape
    assert rel(path, g["path"]) < 1e-4 and rel(actions, g["actions"]) < 1e-6


class _Prefixed:
    """View of one scenario's entries ('{k}.name') of tree_scenarios.npz."""
    def __init__(self, g, k):
        self.g, self.pre = g, f"{k}."

    def __getitem__(self, name):
        return self.g[self.pre + name]


@pytest.mark.parametrize("k", range(15))
def test_tree_replay_every_scenario(k, monkeypatch):
    """SURVEY 8c item 10 on every row of test_scenarios_car.csv: the reference's tree after 30 fake seconds
    (113-145 iterations at B = 1, seeds 42), rebuilt with its sampled actions teacher-forced -- same sampler
    inputs call by call (so the same RNG stream, NN choices and collision verdicts), same parents, edge
    lengths, visit counts, node states and returned path."""
    import ditreeonlineplanner_b200.planners.RRT as rrt_mod
    import ditreeonlineplanner_b200.planners.base_planner as bp_mod
    from ditreeonlineplanner_b200 import load_maze
    g = _Prefixed(golden("tree_scenarios.npz"), k)
    clock = _FakeClock(0.1)
    monkeypatch.setattr(rrt_mod, "time", clock)
    monkeypatch.setattr(bp_mod, "time", clock)
    smp = _ReplaySampler(g)
    torch.manual_seed(42)
    np.random.seed(42)
    random.seed(42)
    pl = make_planner(load_maze(str(g["maze"])), g["start"], g["goal"], sampler=smp, time_budget=30, max_iter=300)
    pl.reset()
    path, actions = pl.plan()
    assert smp.i == int(g["n_calls"]) and pl.results["iterations"] == int(g["iterations"])
    nodes = pl.node_list
    assert np.array_equal([-1 if n.parent is None else n.parent.index for n in nodes], g["parent"])
    assert np.array_equal([0 if n.parent_action_seq is None else len(n.parent_action_seq) for n in nodes], g["edge_len"])
    assert np.array_equal([n.num_visit for n in nodes], g["visits"])
    assert rel(np.array([n.state for n in nodes]), g["states"]) < 1e-4 and smp.max_obs_err < 1e-4
    assert path.shape == g["path"].shape and actions.shape == g["actions"].shape
    assert rel(path, g["path"]) < 1e-4 and rel(actions, g["actions"]) < 1e-6


def test_tree_sampler_calls_match_reference(car_meta):
    """Every sampler call the reference made while growing that tree, re-issued to the bf16
    denoiser with the same weights and noise: actions within the 2e-2 tolerance."""
    from ditreeonlineplanner_b200.policies.fm_policy import DiffusionSampler
    from ditreeonlineplanner_b200.common.map_utils import create_local_map
    from ditreeonlineplanner_b200 import load_maze
    g = golden("tree.npz")
    sd = dref.init_params(seed=21, input_dim=2, cond_dim=7, emb_dim=400, down_dims=[64, 128, 256])
    smp = DiffusionSampler(sd, None, "carmaze", policy="flow_matching", pred_horizon=64, action_dim=2, obs_history=1,
                           action_history=1, goal_conditioned=True, num_diffusion_iters=1, local_map_size=20, max_batch=256)
    grid = load_maze("random_large").astype(np.float32)
    n = int(g["n_calls"])
    obs = g["call_obs"]
    lm = create_local_map(grid, obs[:, 0], obs[:, 1], obs[:, 2], 20, 0.2, 1.0, (grid.shape[1] / 2, grid.shape[0] / 2))
    has = g["call_has_prev"]
    prev = np.where(has[:, None], g["call_prev"], car_meta["Actions_mean"][None])  # mean -> normalised zero
    out = smp(obs[:, None, :], prev_actions=prev[:, None, :], goal=g["call_goal"], local_map=lm,
              noise=torch.as_tensor(g["call_noise"]).cuda())
    assert rel(out[:, :8], g["call_actions"]) < 2e-2


@pytest.mark.parametrize("mode", ["continuous", "rounds"])
def test_batched_planner_runs(mazes, car_meta, mode):
    """batch_size > 1: a tree grown by batched device passes; structural invariants of every edge."""
    from ditreeonlineplanner_b200.policies.fm_policy import DiffusionSampler
    sd = dref.init_params(seed=21, input_dim=2, cond_dim=7, emb_dim=400, down_dims=[64, 128, 256])
    smp = DiffusionSampler(sd, None, "carmaze", policy="flow_matching", pred_horizon=64, action_dim=2, obs_history=1,
                           action_history=1, goal_conditioned=True, num_diffusion_iters=1, local_map_size=20, max_batch=256)
    grid = mazes["random_large"]
    g = golden("tree.npz")
    torch.manual_seed(0)
    np.random.seed(0)
    random.seed(0)
    pl = make_planner(grid, g["start"], g["goal"], sampler=smp, time_budget=30, batch_size=128, iteration_cap=128 * 8 * 6,
                      batch_mode=mode)
    pl.reset()
    path, actions = pl.plan()
    assert len(pl.node_list) > 10
    assert pl.results["iterations"] >= 128 and all(nd.parent in pl.node_list for nd in pl.node_list[1:])
    for nd in pl.node_list[1:]:
        a, s = nd.parent_action_seq, nd.parent_states_seq[0]
        assert 1 <= len(a) <= 64 and len(s) == len(a) + int(np.ceil(len(a) / 8))
        # replaying the edge's actions from the parent reproduces the node and never collides
        res = orc.rollout_car(nd.parent.state[None], a[None], pl.env.goal, grid)
        assert res["first_coll"][0] == -1
        assert rel(res["final"][0], nd.state) < 1e-3
    if path is not None:
        assert path.shape[1] == 6 and actions.shape[1] == 2
