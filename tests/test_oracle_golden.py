"""Pins the CPU oracle (oracle/) against golden vectors produced by the unmodified reference
(tools/gen_golden.py).  CPU-only; runs in seconds."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import denoiser_ref as dref
from oracle import ditree_oracle as orc

MAZES = ["Race_Track", "boxes", "narrow_short", "random_huge", "random_large", "random_xlarge", "shapes",
         "val_maze_10", "val_maze_15", "val_maze_7"]


def test_schedule():
    g = golden("schedule.npz")
    for k in (1, 2, 5, 10):
        t0, dt = orc.fm_schedule(k)
        np.testing.assert_allclose(t0, g[f"t0_{k}"], rtol=0, atol=2e-7)
        np.testing.assert_allclose(dt, g[f"dt_{k}"], rtol=0, atol=2e-7)
        t0t, dtt = dref.fm_schedule_torch(k)
        assert np.array_equal(t0t.numpy(), g[f"t0_{k}"]) and np.array_equal(dtt.numpy(), g[f"dt_{k}"])
    np.testing.assert_allclose(20 * g["t0_10"][:3], [0, 6.7166, 11.2189], atol=1e-3)


@pytest.mark.parametrize("maze", MAZES)
def test_local_map(mazes, maze):
    g = golden("local_map.npz")
    grid = mazes[maze]
    R, C = grid.shape
    for tag, n, scale, sg in (("car", 20, 0.2, 1.0), ("ant", 16, 0.8, 4.0)):
        pose = g[f"{maze}.{tag}.pose"].astype(np.float64)
        got = orc.local_map(grid, pose[:, 0], pose[:, 1], pose[:, 2], n, scale, sg, (C / 2 * sg, R / 2 * sg))
        assert np.array_equal(got.astype(np.uint8), g[f"{maze}.{tag}.map"])


def test_local_map_scalar(mazes):
    g = golden("local_map.npz")
    got = orc.local_map(mazes["boxes"], 1.25, -3.5, 0.7, 20, 0.2, 1.0, (10.0, 10.0))
    assert np.array_equal(got.astype(np.uint8), g["scalar.map"])


@pytest.mark.parametrize("maze", MAZES)
def test_collide_car(mazes, maze):
    g = golden("collide_car.npz")
    st = g[f"{maze}.states"].astype(np.float64)
    want = np.unpackbits(g[f"{maze}.flags"])[: int(g[f"{maze}.n"])].astype(bool)
    assert 0.05 < want.mean() < 0.999
    assert np.array_equal(orc.collide_car_batch(st, mazes[maze]), want)
    assert np.array_equal(orc.collide_car(st[:300], mazes[maze]), want[:300])


def test_collide_quirks(mazes):
    g = golden("collide_car.npz")
    assert bool(g["quirk.random_large"][0]) is True and bool(g["quirk.narrow_short"][0]) is False
    assert orc.collide_points(np.array([3.9719289005037552, 0.5062825501784967]), mazes["random_large"])[0]
    assert not orc.collide_points(np.array([2.4975924901541786, -1.0489708935371076]), mazes["narrow_short"])[0]
    free = np.zeros((5, 5), np.float32)
    assert list(g["quirk.border"]) == [True, False]
    assert orc.collide_points(np.array([0.0, 2.0]), free)[0] and not orc.collide_points(np.array([0.0, 0.0]), free)[0]
    wall = np.ones((5, 5), np.float32)
    assert list(g["quirk.batch_oob"]) == [False, True]
    assert list(orc.collide_points(np.array([[0.0, 0.0], [9.0, 0.0]]), wall)) == [False, True]
    pts = g["points.boxes.pts"].astype(np.float64)
    assert np.array_equal(orc.collide_points(pts, mazes["boxes"]), g["points.boxes.flags"])
    tall = np.zeros((7, 4), np.float32)
    with pytest.raises(IndexError):
        orc.collide_points(np.array([1.5, 0.2]), tall)


def test_collide_ant(mazes):
    g = golden("collide_ant.npz")
    grid = np.zeros((5, 5), np.float32)
    grid[2, 2] = 1
    assert list(g["small.flags"]) == [False, False, True, False, True, True]
    assert np.array_equal(orc.collide_ant_batch(g["small.states"], grid), g["small.flags"])
    st = np.zeros((len(g["huge.states"]), 29))
    st[:, :7] = g["huge.states"]
    want = g["huge.flags"]
    assert 0.1 < want.mean() < 0.95
    assert np.array_equal(orc.collide_ant_batch(st, mazes["random_huge"]), want)


def test_bicycle(mazes):
    g = golden("bicycle.npz")
    s0, act, want = g["s0"].astype(np.float64), g["act"].astype(np.float64), g["traj"]
    cur = s0.copy()
    for i in range(act.shape[1]):
        cur = orc.bicycle_step(cur, act[:, i])
        np.testing.assert_allclose(cur, want[:, i], rtol=1e-12, atol=1e-12)
    res = orc.rollout_car(s0, act, np.array([100.0, 100.0]), mazes["boxes"], stop_on_collision=False)
    np.testing.assert_allclose(res["traj"], want, rtol=1e-12, atol=1e-12)
    # goal latch: the state freezes at the first step inside the 0.5 m goal disc
    latch = orc.rollout_car(np.array([[-1.0, 0, 0, 3.0, 0.5, 0]]), np.zeros((1, 30, 2)), np.array([0.3, 0.0]),
                            np.zeros((40, 40), np.float32), stop_on_collision=False)
    k = int(latch["done_step"][0])
    assert k == int(np.argmax(g["latch_success"]))
    np.testing.assert_allclose(latch["final"][0], g["latch"][-1], rtol=1e-12)
    np.testing.assert_allclose(latch["traj"][0, : k + 1], g["latch"][: k + 1], rtol=1e-12)


def test_propagate_conventions(mazes):
    g = golden("propagate.npz")
    kinds = set()
    for i in range(int(g["n_cases"])):
        obs, done, a, s = orc.propagate_action_sequence(g[f"{i}.state"], g[f"{i}.act"], 8, g["goal_xy"], mazes["boxes"])
        want_done = int(g[f"{i}.done"])
        kinds.add(want_done)
        assert (-1 if done is None else int(done)) == want_done
        np.testing.assert_allclose(obs, g[f"{i}.obs"], rtol=1e-12, atol=1e-12)
        assert a.shape == g[f"{i}.a"].shape and s.shape == g[f"{i}.s"].shape
        np.testing.assert_allclose(a, g[f"{i}.a"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(s, g[f"{i}.s"], rtol=1e-12, atol=1e-12)
    assert kinds == {-1, 0, 1}


def test_cond_car(car_meta):
    g = golden("cond.npz")
    c = orc.build_cond_car(g["car.obs"], g["car.prev"][:, -1], g["car.goal"], car_meta, 20.0)
    np.testing.assert_allclose(c, g["car.cond"], rtol=0, atol=2e-6)
    c = orc.build_cond_car(g["car.obs"][:1], None, g["car.goal"], car_meta, 20.0)
    np.testing.assert_allclose(c, g["car.cond_noprev"], rtol=0, atol=2e-6)
    c = orc.build_cond_car(g["car.obs"], g["car.prev"][:, -1], g["car.goals"], car_meta, 20.0)
    np.testing.assert_allclose(c, g["car.cond_goals"], rtol=0, atol=2e-6)
    c = orc.build_cond_car(np.array([[1, -2, 0.5, 2, 0.6, 0.1]]), None, np.array([4.0, 3.0]), car_meta, 20.0)
    np.testing.assert_allclose(c, g["car.cond_example"], atol=2e-6)
    np.testing.assert_allclose(c[0], [-0.6, 0.2, 0.25, 0, 0, 0.24632, 0.14642], atol=1e-4)
    # zero-velocity net: the sampler returns the un-normalised noise
    np.testing.assert_allclose(g["car.result"], g["car.noise"].astype(np.float64) * car_meta["Actions_std"]
                               + car_meta["Actions_mean"], rtol=1e-12)
    assert np.array_equal(np.unique(g["car.map_in"]), [-1.0, 1.0])


def test_cond_ant(ant_meta):
    g = golden("cond.npz")
    for h in (1, 3):
        c = orc.build_cond_ant(g[f"ant.h{h}.obs"], g[f"ant.h{h}.prev"][:, -1], g[f"ant.h{h}.goal"], ant_meta, 3, 16.0)
        assert c.shape == (16, 97)
        np.testing.assert_allclose(c, g[f"ant.h{h}.cond"], rtol=0, atol=5e-6)


def test_param_inventory():
    small = dref.param_shapes(2, 7, 400, (64, 128, 256))
    assert len(small) == 210
    large = dref.param_shapes(2, 7, 400, (512, 1024, 2048))
    n = sum(int(np.prod(s)) for s in large.values())
    assert abs(n - 184.1e6) < 0.2e6


@pytest.mark.parametrize("tag", ["small", "large"])
def test_denoiser(tag, car_meta):
    g = golden(f"denoiser_{tag}.npz")
    torch.set_num_threads(max(1, torch.get_num_threads()))
    sd = dref.init_params(seed=int(g["seed"]), input_dim=2, cond_dim=7, emb_dim=400, down_dims=list(g["dims"]))
    checksum = float(sum(float(v.double().sum()) for v in sd.values()))
    assert abs(checksum - float(g["weight_checksum"])) < 1e-6 * max(1.0, abs(checksum)), "seeded weights drifted"
    lm = torch.from_numpy(g["lm01"])
    with torch.no_grad():
        enc = dref.encoder_forward(sd, lm * 2 - 1)
        np.testing.assert_allclose(enc.numpy(), g["enc"], rtol=1e-4, atol=1e-5)
        vel = dref.unet_forward(sd, torch.from_numpy(g["sample"]), torch.from_numpy(g["ts"]),
                                torch.cat([enc, torch.from_numpy(g["cond"])], 1))
        np.testing.assert_allclose(vel.numpy(), g["vel"], rtol=1e-4, atol=1e-5)
    cond = orc.build_cond_car(g["obs"], g["prev"][:, -1], g["goal"], car_meta, 20.0)
    act = dref.fm_sample(sd, torch.from_numpy(g["noise"]), torch.from_numpy(cond), lm, int(g["K"]),
                         car_meta["Actions_mean"], car_meta["Actions_std"])
    assert act.dtype == np.float64 and act.shape == g["actions"].shape
    np.testing.assert_allclose(act, g["actions"], rtol=1e-4, atol=1e-5)


def test_nearest_matches_kdtree():
    from scipy.spatial import KDTree
    rng = np.random.default_rng(5)
    for n in (1, 100, 10000):
        nodes = rng.uniform(-10, 10, (n, 2))
        q = rng.uniform(-10, 10, (512, 2))
        q[::9] = nodes[rng.integers(0, n, len(q[::9]))]  # exact hits
        _, idx = KDTree(nodes).query(q, k=1)
        assert np.array_equal(orc.nearest(nodes, q), idx)


def test_lidar():
    g = golden("lidar.npz")
    for i in range(int(g["n"])):
        maze = g[f"maze{int(g[f'{i}.maze'])}"].astype(np.float64)
        d, e, v = orc.lidar_scan(g[f"{i}.pose"], maze)
        np.testing.assert_allclose(d, g[f"{i}.dist"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(e, g[f"{i}.end"], rtol=1e-9, atol=1e-9)
        assert np.array_equal(v, g[f"{i}.visited"])


def test_ray_probe(mazes):
    g = golden("probe.npz")
    got = np.array([orc.ray_probe(s, mazes["boxes"]) for s in g["states"].astype(np.float64)])
    assert 0.1 < g["flags"].mean() < 0.9
    assert np.array_equal(got, g["flags"])


def test_mppi_reduce_properties():
    rng = np.random.default_rng(3)
    cost = rng.uniform(0, 10, 64)
    noise = rng.normal(size=(64, 16, 2))
    u, amin, w = orc.mppi_reduce(cost, noise, 0.5, np.zeros((16, 2)))
    assert amin == int(np.argmin(cost)) and abs(w.sum() - 1) < 1e-12
    u2, _, _ = orc.mppi_reduce(cost + 100.0, noise, 0.5, np.zeros((16, 2)))
    np.testing.assert_allclose(u, u2, rtol=1e-9)  # shift invariance


# ---- probability-map state sampler (run_type >= 2) ---------------------------------------------
def test_probmap_oracle(mazes):
    g = golden("probmap.npz")
    for m in MAZES:
        assert np.array_equal(orc.edt_prior(mazes[m]), g[f"{m}.prior"])          # exact: sqrt of integer d^2
    free = np.zeros((4, 5))
    want = np.sqrt((np.arange(4)[:, None] + 1.0) ** 2 + np.arange(5)[None, :] ** 2)   # SciPy on a wall-free map
    np.testing.assert_allclose(orc.edt_prior(free), want / want.sum(), rtol=1e-15)
    for i in range(int(g["n_cases"])):
        pdf = orc.gaussian_map(g[f"{i}.robot"], g[f"{i}.goal"])
        np.testing.assert_allclose(pdf, g[f"{i}.pdf"], rtol=1e-13, atol=1e-300)
        blend = orc.combine_log_blend(g[f"{str(g[f'{i}.maze'])}.prior"], g[f"{i}.pdf"])
        np.testing.assert_allclose(blend, g[f"{i}.blend"], rtol=1e-13, atol=1e-300)
        assert np.array_equal(orc.sample_cells(g[f"{i}.blend"], g[f"{i}.u"]), g[f"{i}.idx"])


def test_online_path_check_and_scan_writeback(mazes):
    """check_no_obstacles_in_path / scan_and_update_maze (run_scenarios_with_lidar_DiTree.py:112-127,158-181):
    golden = the reference functions' own source exec'ed on seeded inputs (tools/gen_golden.py::gen_online)."""
    g = golden("online.npz")
    hits = 0
    for i in range(int(g["n_path"])):
        want = int(g[f"path{i}.idx"])
        got = orc.path_first_obstacle(g[f"path{i}.path"], g[f"path{i}.scanned"].astype(np.float64))
        assert got == want, i
        hits += want >= 0
    assert 10 < hits < int(g["n_path"])        # both outcomes are exercised
    for i in range(int(g["n_scan"])):
        base = mazes["boxes"].astype(np.float64)
        known, scanned = base.copy(), np.zeros_like(base)
        orc.scan_and_update_maze(g[f"scan{i}.state"], known, g[f"scan{i}.with_obs"].astype(np.float64), scanned)
        assert np.array_equal(known, g[f"scan{i}.known"]), i
        assert np.array_equal(scanned, g[f"scan{i}.scanned"]), i
