"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI,
against the pinned oracle on the same seeded inputs and against the committed golden vectors."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import ditree_oracle as orc

pytestmark = pytest.mark.gpu

MAZES = ["Race_Track", "boxes", "narrow_short", "random_huge", "random_large", "random_xlarge", "shapes",
         "val_maze_10", "val_maze_15", "val_maze_7"]


@pytest.fixture(scope="module")
def ctx():
    from ditreeonlineplanner_b200 import Context
    c = Context(0)
    yield c
    c.close()


def dev(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float32)).cuda()


# ---- collision -----------------------------------------------------------------------------
@pytest.mark.parametrize("maze", MAZES)
def test_collide_car_golden(ctx, mazes, maze):
    g = golden("collide_car.npz")
    ctx.set_map(mazes[maze])
    st = g[f"{maze}.states"]
    want = np.unpackbits(g[f"{maze}.flags"])[: int(g[f"{maze}.n"])].astype(bool)
    got = ctx.collide_car(dev(st)).cpu().numpy().astype(bool)
    ctx.sync_status()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("maze", ["boxes", "random_large", "narrow_short", "random_huge"])
def test_collide_car_2m_vs_oracle(ctx, mazes, maze):
    """Full-size bit-exact check: 2 M random states per maze."""
    grid = mazes[maze]
    R, C = grid.shape
    rng = np.random.default_rng(hash(maze) % 2**31)
    n = 2_000_000
    st = np.stack([rng.uniform(-C / 2 - 0.2, C / 2 + 0.2, n), rng.uniform(-R / 2 - 0.2, R / 2 + 0.2, n),
                   rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)
    ctx.set_map(grid)
    got = ctx.collide_car(dev(st)).cpu().numpy().astype(bool)
    want = orc.collide_car_batch(st.astype(np.float64), grid)
    assert np.array_equal(got, want), f"{np.sum(got != want)} flags differ"


def test_collide_car_ragged_sizes_and_special_values(ctx, mazes):
    """The four-states-per-thread kernel and its single-state tail: batch sizes around the multiples of four,
    shifted views (the flag array stays aligned, the states do not), non-finite positions, NaN / huge headings
    (must agree with the single-state kernel, which takes the exact code for them)."""
    rng = np.random.default_rng(3)
    for maze in ("boxes", "random_huge", "narrow_short"):
        grid = mazes[maze]
        R, C = grid.shape
        ctx.set_map(grid)
        for n in (1, 3, 4, 5, 7, 1027, 100_003):
            st = np.stack([rng.uniform(-C / 2 - 1, C / 2 + 1, n), rng.uniform(-R / 2 - 1, R / 2 + 1, n),
                           rng.uniform(-7, 7, n)], 1).astype(np.float32)
            special = []
            if n > 100:
                st[7], st[8], st[9], st[10] = [np.nan, 0, 0], [0, 0, np.nan], [0.3, 0.2, 1e6], [np.inf, 0, 0]
                special = [7, 8, 9, 10]
            d = dev(st)
            got = ctx.collide_car(d).cpu().numpy().astype(bool)
            for i in special:
                assert got[i] == bool(ctx.collide_car(d[i:i + 1]).item()), (maze, n, i)
            keep = np.setdiff1d(np.arange(n), special)
            assert np.array_equal(got[keep], orc.collide_car_batch(st[keep].astype(np.float64), grid)), (maze, n)
            if n > 1:
                assert np.array_equal(ctx.collide_car(d[1:]).cpu().numpy().astype(bool), got[1:])


def test_collide_points_semantics(ctx, mazes):
    g = golden("collide_car.npz")
    ctx.set_map(mazes["boxes"])
    got = ctx.collide_points(dev(g["points.boxes.pts"])).cpu().numpy().astype(bool)
    assert np.array_equal(got, g["points.boxes.flags"])
    # documented quirks of is_colliding_parallel
    ctx.set_map(mazes["random_large"])
    assert ctx.collide_points(dev([[3.9719289005037552, 0.5062825501784967]])).item() == 1
    ctx.set_map(mazes["narrow_short"])
    assert ctx.collide_points(dev([[2.4975924901541786, -1.0489708935371076]])).item() == 0
    ctx.set_map(np.zeros((5, 5), np.float32))
    assert ctx.collide_points(dev([[0.0, 2.0]])).item() == 1 and ctx.collide_points(dev([[0.0, 0.0]])).item() == 0
    ctx.set_map(np.ones((5, 5), np.float32))
    assert ctx.collide_points(dev([[0.0, 0.0], [9.0, 0.0]])).cpu().tolist() == [0, 1]  # batch early return
    ctx.sync_status()
    # tall map: the reference raises IndexError
    ctx.set_map(np.zeros((7, 4), np.float32))
    ctx.collide_points(dev([[1.5, 0.2]]))
    with pytest.raises(IndexError):
        ctx.sync_status()
    # empty input
    assert ctx.collide_points(torch.zeros((0, 2), device="cuda")).numel() == 0


def test_collide_ant(ctx, mazes):
    g = golden("collide_ant.npz")
    grid = np.zeros((5, 5), np.float32)
    grid[2, 2] = 1
    ctx.set_map(grid, 4.0)
    got = ctx.collide_ant(dev(g["small.states"])).cpu().numpy().astype(bool)
    assert np.array_equal(got, g["small.flags"])
    ctx.set_map(mazes["random_huge"], 4.0)
    got = ctx.collide_ant(dev(g["huge.states"])).cpu().numpy().astype(bool)
    assert np.array_equal(got, g["huge.flags"])
    # 16384-state batch (config C5) vs the oracle
    rng = np.random.default_rng(9)
    n = 16384
    st = np.zeros((n, 29), np.float32)
    st[:, :2] = rng.uniform(-63, 63, (n, 2))
    q = rng.normal(size=(n, 4))
    st[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    got = ctx.collide_ant(dev(st)).cpu().numpy().astype(bool)
    assert np.array_equal(got, orc.collide_ant_batch(st.astype(np.float64), mazes["random_huge"]))


# ---- local map -----------------------------------------------------------------------------
@pytest.mark.parametrize("maze", MAZES)
def test_local_map_golden(ctx, mazes, maze):
    g = golden("local_map.npz")
    for tag, n, scale, sg in (("car", 20, 0.2, 1.0), ("ant", 16, 0.8, 4.0)):
        ctx.set_map(mazes[maze], sg)
        pose = dev(g[f"{maze}.{tag}.pose"])
        got = ctx.local_map(pose, n, scale)
        assert got.dtype == torch.float32 and got.shape == (pose.shape[0], n, n)
        assert np.array_equal(got.cpu().numpy().astype(np.uint8), g[f"{maze}.{tag}.map"])
        signed = ctx.local_map(pose, n, scale, bf16_signed=True).float().cpu().numpy()
        assert np.array_equal(signed, g[f"{maze}.{tag}.map"].astype(np.float32) * 2 - 1)


def test_local_map_100k_vs_oracle(ctx, mazes):
    grid = mazes["boxes"]
    ctx.set_map(grid)
    rng = np.random.default_rng(3)
    n = 100_000
    pose = np.stack([rng.uniform(-10.5, 10.5, n), rng.uniform(-10.5, 10.5, n), rng.uniform(-7, 7, n)], 1).astype(np.float32)
    got = ctx.local_map(dev(pose), 20, 0.2).cpu().numpy()
    p64 = pose.astype(np.float64)
    want = orc.local_map(grid, p64[:, 0], p64[:, 1], p64[:, 2], 20, 0.2, 1.0, (10.0, 10.0))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n_side", [2, 5, 15, 32])
def test_local_map_other_sizes(ctx, mazes, n_side):
    """Odd point counts take the one-point-per-lane path, even ones the paired path; both output types."""
    grid = mazes["random_huge"]
    ctx.set_map(grid)
    rng = np.random.default_rng(n_side)
    n = 3000
    R, C = grid.shape
    pose = np.stack([rng.uniform(-C / 2 - 1, C / 2 + 1, n), rng.uniform(-R / 2 - 1, R / 2 + 1, n),
                     rng.uniform(-7, 7, n)], 1).astype(np.float32)
    p64 = pose.astype(np.float64)
    want = orc.local_map(grid, p64[:, 0], p64[:, 1], p64[:, 2], n_side, 0.3, 1.0, (C / 2, R / 2))
    assert np.array_equal(ctx.local_map(dev(pose), n_side, 0.3).cpu().numpy(), want)
    signed = ctx.local_map(dev(pose), n_side, 0.3, bf16_signed=True).float().cpu().numpy()
    assert np.array_equal(signed, want.astype(np.float32) * 2 - 1)


@pytest.mark.parametrize("n_side,scale,sg", [(20, 0.2, 1.0), (16, 0.8, 4.0)])
def test_local_map_quad_border_poses(ctx, mazes, n_side, scale, sg):
    """The quad kernel (N = 16 / 20) on poses that put whole rows of lattice points ON cell borders (axis-aligned headings,
    positions on the lattice pitch): every such point is inside the fp32 guard band, so the per-block list of deferred
    float64 decisions overflows and the in-line fallback runs too; plus far-outside and non-finite poses (every point
    exact).  Bit-exact against the oracle, both output types, unaligned tail of the batch included."""
    grid = mazes["random_large"] if sg == 1.0 else mazes["random_huge"]
    ctx.set_map(grid, sg)
    R, C = grid.shape
    rng = np.random.default_rng(n_side)
    n = 20_003
    k = rng.integers(-int(C * sg / scale / 2), int(C * sg / scale / 2), (n, 2))
    pose = np.stack([k[:, 0] * (scale / 2), k[:, 1] * (scale / 2), rng.integers(-4, 5, n) * (np.pi / 2)], 1)
    pose[::7, 2] += rng.uniform(-1e-6, 1e-6, len(pose[::7]))          # a hair off the axis
    pose[5::101, :2] = rng.uniform(-3 * C * sg, 3 * C * sg, (len(pose[5::101]), 2))   # centres far outside the map
    pose[11::503, 2] = 1.0e5                                           # heading beyond the MUFU range
    pose = pose.astype(np.float32)
    p64 = pose.astype(np.float64)
    want = orc.local_map(grid, p64[:, 0], p64[:, 1], p64[:, 2], n_side, scale, sg, (C * sg / 2, R * sg / 2))
    got = ctx.local_map(dev(pose), n_side, scale).cpu().numpy()
    assert np.array_equal(got, want)
    signed = ctx.local_map(dev(pose), n_side, scale, bf16_signed=True).float().cpu().numpy()
    assert np.array_equal(signed, want.astype(np.float32) * 2 - 1)


# ---- dynamics ------------------------------------------------------------------------------
def rel_err(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_bicycle_golden(ctx, mazes):
    g = golden("bicycle.npz")
    ctx.set_map(mazes["boxes"])
    res = ctx.propagate_collide(dev(g["s0"]), dev(g["act"]), (100.0, 100.0), stop_on_collision=False)
    traj = res["traj"].cpu().numpy().astype(np.float64)
    # north-star tolerance: propagated states within 1e-4 relative of the reference's float64 Euler
    for b in range(traj.shape[0]):
        assert rel_err(traj[b], g["traj"][b]) < 1e-4
    assert rel_err(res["final"].cpu().numpy(), g["traj"][:, -1]) < 1e-4
    # SoA layout gives the same bits
    soa = ctx.propagate_collide(dev(g["s0"]).t().contiguous(), dev(g["act"]).permute(1, 2, 0).contiguous(),
                                (100.0, 100.0), soa=True, stop_on_collision=False)
    assert torch.equal(soa["traj"].permute(2, 0, 1), res["traj"])
    assert torch.equal(soa["final"].t(), res["final"])


@pytest.mark.parametrize("maze", ["boxes", "random_large"])
def test_propagate_flags_bit_exact(ctx, mazes, maze):
    """Teacher-forced flags: the oracle's collision and goal tests evaluated on the kernel's own
    fp32 trajectory must reproduce first_coll / done_step exactly; states within 1e-4."""
    grid = mazes[maze]
    R, C = grid.shape
    ctx.set_map(grid)
    rng = np.random.default_rng(17)
    B, S = 20000, 50
    free = np.argwhere(grid == 0)
    cells = free[rng.integers(0, len(free), B)]
    x, y = orc.rowcol_to_xy(cells[:, 0], cells[:, 1], grid.shape)
    s0 = np.stack([x + rng.uniform(-0.3, 0.3, B), y + rng.uniform(-0.3, 0.3, B), rng.uniform(-np.pi, np.pi, B),
                   rng.uniform(0, 4, B), rng.uniform(0, 1.3, B), rng.uniform(-0.44, 0.44, B)], 1).astype(np.float32)
    act = np.stack([rng.normal(0.45, 1.0, (B, S)), rng.normal(0, 0.92, (B, S))], -1).astype(np.float32)
    goal = orc.rowcol_to_xy(free[len(free) // 2][0], free[len(free) // 2][1], grid.shape)
    goal = (float(goal[0]), float(goal[1]))
    for stop in (True, False):
        res = ctx.propagate_collide(dev(s0), dev(act), goal, stop_on_collision=stop)
        traj = res["traj"].cpu().numpy()
        first, done = res["first_coll"].cpu().numpy(), res["done_step"].cpu().numpy()
        forced = orc.rollout_car(s0, act, goal, grid, stop_on_collision=stop, states_for_flags=traj)
        assert np.array_equal(first, forced["first_coll"])
        assert np.array_equal(done, forced["done_step"])
        assert (first >= 0).mean() > 0.2 and (done >= 0).sum() > 0
        # rows after the edge ended are zero; rows before match the float64 reference dynamics
        free_run = orc.rollout_car(s0, act, goal, grid, stop_on_collision=stop)
        same = (free_run["first_coll"] == first) & (free_run["done_step"] == done)
        assert same.mean() > 0.995  # flags computed from fp32 vs fp64 states may differ near a boundary
        err = np.linalg.norm(traj[same] - free_run["traj"][same], axis=(1, 2)) / np.maximum(
            np.linalg.norm(free_run["traj"][same], axis=(1, 2)), 1e-9)
        assert err.max() < 1e-4
        end = np.where(first >= 0, first, np.where(done >= 0, done, S - 1)) if stop else np.where(done >= 0, done, S - 1)
        for b in range(0, B, 997):
            assert np.all(traj[b, end[b] + 1:] == 0)
            np.testing.assert_array_equal(res["final"][b].cpu().numpy(), traj[b, end[b]])


def test_propagate_golden_conventions(ctx, mazes):
    g = golden("propagate.npz")
    ctx.set_map(mazes["boxes"])
    for i in range(int(g["n_cases"])):
        res = ctx.propagate_collide(dev(g[f"{i}.state"][None]), dev(g[f"{i}.act"][None]), g["goal_xy"])
        first, done = int(res["first_coll"][0]), int(res["done_step"][0])
        want = int(g[f"{i}.done"])
        if want == -1:  # collision: the reference returns done=None and actions[:i]
            assert first >= 0 and first == g[f"{i}.a"].shape[0]
        elif want == 1:
            assert first < 0 and done >= 0
            assert np.all(g[f"{i}.a"][done + 1:] == 0)
        else:
            assert first < 0 and done < 0
        assert rel_err(res["final"][0].cpu().numpy(), g[f"{i}.obs"]) < 1e-4


def test_propagate_edge_cases(ctx, mazes):
    ctx.set_map(mazes["boxes"])
    res = ctx.propagate_collide(torch.zeros((0, 6), device="cuda"), torch.zeros((0, 50, 2), device="cuda"), (0, 0))
    assert res["traj"].shape == (0, 50, 6)
    # sampler-shaped actions (B,64,2) with S = 50 < T
    s0 = torch.tensor([[-7.5, -7.5, 0.8, 1.0, 0.5, 0.0]], device="cuda").repeat(33, 1)
    act = torch.randn(33, 64, 2, device="cuda")
    a = ctx.propagate_collide(s0, act, (50.0, 50.0), S=50)
    b = ctx.propagate_collide(s0, act[:, :50].contiguous(), (50.0, 50.0))
    assert torch.equal(a["traj"], b["traj"])
    # no trajectory requested
    c = ctx.propagate_collide(s0, act, (50.0, 50.0), S=50, want_traj=False)
    assert c["traj"] is None and torch.equal(c["final"], a["final"])


# ---- conditioning --------------------------------------------------------------------------
def test_cond_car(ctx, car_meta):
    g = golden("cond.npz")
    c = ctx.build_cond_car(dev(g["car.obs"]), dev(g["car.prev"][:, -1]), dev(g["car.goal"]), car_meta, 20.0)
    np.testing.assert_allclose(c.cpu().numpy(), g["car.cond"], rtol=0, atol=3e-6)
    c = ctx.build_cond_car(dev(g["car.obs"][:1]), None, dev(g["car.goal"]), car_meta, 20.0)
    np.testing.assert_allclose(c.cpu().numpy(), g["car.cond_noprev"], rtol=0, atol=3e-6)
    c = ctx.build_cond_car(dev(g["car.obs"]), dev(g["car.prev"][:, -1]), dev(g["car.goals"]), car_meta, 20.0)
    np.testing.assert_allclose(c.cpu().numpy(), g["car.cond_goals"], rtol=0, atol=3e-6)


def test_cond_ant(ctx, ant_meta):
    g = golden("cond.npz")
    for h in (1, 3):
        c = ctx.build_cond_ant(dev(g[f"ant.h{h}.obs"]), dev(g[f"ant.h{h}.prev"][:, -1]), dev(g[f"ant.h{h}.goal"]),
                               ant_meta, 3, 16.0)
        np.testing.assert_allclose(c.cpu().numpy(), g[f"ant.h{h}.cond"], rtol=0, atol=2e-5)


# ---- reductions ----------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 100, 10_000, 100_000])
def test_nearest(ctx, n):
    from scipy.spatial import KDTree
    rng = np.random.default_rng(n)
    nodes = rng.uniform(-10, 10, (n, 2)).astype(np.float32)
    q = rng.uniform(-10, 10, (4096, 2)).astype(np.float32)
    q[::5] = nodes[rng.integers(0, n, len(q[::5]))]  # exact hits
    q[1::50] = np.array([7.5, -7.5], np.float32)  # the repeated goal query
    got = ctx.nearest(dev(nodes[:, 0]), dev(nodes[:, 1]), dev(q)).cpu().numpy()
    assert np.array_equal(got, orc.nearest(nodes, q))  # bit-exact vs the brute-force oracle
    _, idx = KDTree(nodes.astype(np.float64)).query(q.astype(np.float64), k=1)
    d_got = np.linalg.norm(nodes[got].astype(np.float64) - q, axis=1)
    d_kd = np.linalg.norm(nodes[idx].astype(np.float64) - q, axis=1)
    assert np.array_equal(d_got, d_kd)  # same distances as scipy; indices differ only on exact ties
    assert (got == idx).mean() > 0.999


@pytest.mark.parametrize("n,k", [(1, 1), (3, 5), (100, 4), (5000, 8), (100_000, 16)])
def test_nearest_k_vs_oracle_and_kdtree(ctx, n, k):
    from scipy.spatial import KDTree
    rng = np.random.default_rng(n + k)
    nodes = rng.uniform(-10, 10, (n, 2)).astype(np.float32)
    if n > 50:
        nodes[7] = nodes[3]                       # exact duplicates: lowest index first
        nodes[11] = nodes[3]
    q = np.concatenate([rng.uniform(-10, 10, (300, 2)).astype(np.float32), nodes[:min(n, 20)]])
    got = ctx.nearest_k(dev(nodes[:, 0]), dev(nodes[:, 1]), dev(q), k).cpu().numpy()
    want = orc.nearest_k(nodes, q, k)
    bad = np.nonzero(np.any(got != want, axis=1))[0]
    assert len(bad) == 0, (bad[:5], got[bad[:5]], want[bad[:5]])
    # SciPy's KD-tree agrees wherever the neighbour distances are distinct (its tie order is unspecified)
    kk = min(k, n)
    k1 = min(k + 1, n)                             # one more neighbour: a tie at the cut-off is a tie too
    d, idx = KDTree(nodes.astype(np.float64)).query(q.astype(np.float64), k=k1)
    d, idx = d.reshape(len(q), k1), idx.reshape(len(q), k1)
    distinct = np.ones(len(q), bool) if k1 == 1 else np.all(np.diff(d, axis=1) > 0, axis=1)
    assert distinct.mean() > 0.5
    assert np.array_equal(got[distinct][:, :kk], idx[distinct][:, :kk])
    # k = 1 is dt_nearest
    one = ctx.nearest(dev(nodes[:, 0]), dev(nodes[:, 1]), dev(q)).cpu().numpy()
    assert np.array_equal(one, got[:, 0])


def test_nearest_ties_and_argmin(ctx):
    nodes = np.array([[0, 0], [1, 0], [0, 0], [1, 0]], np.float32)
    got = ctx.nearest(dev(nodes[:, 0]), dev(nodes[:, 1]), dev([[0.1, 0], [0.9, 0], [0.5, 0]])).cpu().tolist()
    assert got == [0, 1, 0]
    rng = np.random.default_rng(4)
    xy = rng.uniform(-10, 10, (5000, 2)).astype(np.float32)
    ahead = rng.random(5000) < 0.5
    goal = (3.25, -1.5)
    got = int(ctx.goal_cost_argmin(dev(xy[:, 0]), dev(xy[:, 1]), goal, torch.as_tensor(ahead).cuda())[0])
    assert got == orc.final_node_cost_argmin(xy, goal, ahead)
    got = int(ctx.goal_cost_argmin(dev(xy[:, 0]), dev(xy[:, 1]), goal)[0])
    assert got == orc.final_node_cost_argmin(xy, goal, np.zeros(5000, bool))


def test_mppi_reduce(ctx):
    rng = np.random.default_rng(8)
    for K, T in ((8192, 16), (100, 8), (1, 4)):
        cost = rng.uniform(0, 50, K).astype(np.float32)
        noise = rng.normal(size=(K, T, 2)).astype(np.float32)
        u0 = rng.normal(size=(T, 2)).astype(np.float32)
        u, amin, w = ctx.mppi_reduce(dev(cost), dev(noise), 2.0, dev(u0), want_weights=True)
        wu, wamin, ww = orc.mppi_reduce(cost, noise, 2.0, u0)
        assert int(amin[0]) == wamin
        np.testing.assert_allclose(w.cpu().numpy(), ww, rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(u.cpu().numpy(), wu, rtol=1e-4, atol=1e-5)


# ---- probes and lidar ----------------------------------------------------------------------
def test_ray_probe(ctx, mazes):
    g = golden("probe.npz")
    ctx.set_map(mazes["boxes"])
    got = ctx.ray_probe(dev(g["states"])).cpu().numpy().astype(bool)
    assert np.array_equal(got, g["flags"])
    rng = np.random.default_rng(2)
    st = np.stack([rng.uniform(-9.9, 9.9, 50000), rng.uniform(-9.9, 9.9, 50000), rng.uniform(-7, 7, 50000)], 1).astype(np.float32)
    got = ctx.ray_probe(dev(st)).cpu().numpy().astype(bool)
    want = np.array([orc.ray_probe(s, mazes["boxes"]) for s in st[:5000].astype(np.float64)])
    assert np.array_equal(got[:5000], want)


def test_path_first_obstacle(ctx, mazes):
    grid = mazes["boxes"].copy()
    ctx.set_map(grid)
    rng = np.random.default_rng(6)
    for trial in range(5):
        path = np.cumsum(rng.normal(0, 0.25, (300, 2)), axis=0).astype(np.float32)
        path = np.clip(path, -9.4, 9.4)
        got = int(ctx.path_first_obstacle(dev(path))[0])
        assert got == orc.path_first_obstacle(path.astype(np.float64), grid)
    free_path = np.tile(np.array([[-7.5, -7.5]], np.float32), (10, 1))
    assert int(ctx.path_first_obstacle(dev(free_path))[0]) == -1


def test_lidar_golden(ctx):
    g = golden("lidar.npz")
    for i in range(int(g["n"])):
        maze = g[f"maze{int(g[f'{i}.maze'])}"].astype(np.float32)
        ctx.set_map(maze)
        dist, end, vis = ctx.lidar_scan(dev(g[f"{i}.pose"][None]))
        np.testing.assert_allclose(dist[0].cpu().numpy(), g[f"{i}.dist"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(end[0].cpu().numpy(), g[f"{i}.end"], rtol=0, atol=1e-5)
        # hit cells exact
        assert np.array_equal(np.floor(end[0].cpu().numpy()), np.floor(g[f"{i}.end"]))
        want = np.zeros(maze.shape, np.uint8)
        v = g[f"{i}.visited"]
        want[v[:, 1], v[:, 0]] = 1
        assert np.array_equal(vis[0].cpu().numpy(), want)
    ctx.sync_status()


def test_lidar_batch_vs_oracle(ctx):
    g = golden("lidar.npz")
    maze = g["maze2"].astype(np.float32)
    ctx.set_map(maze)
    rng = np.random.default_rng(12)
    poses = []
    while len(poses) < 64:
        x, y = rng.uniform(1, 19, 2)
        if maze[int(y), int(x)] == 0:
            poses.append([x, y, rng.uniform(-3, 3)])
    poses = np.array(poses, np.float32)
    dist, end, vis = ctx.lidar_scan(dev(poses))
    for b in range(0, 64, 8):
        d, e, v = orc.lidar_scan(poses[b].astype(np.float64), maze.astype(np.float64))
        np.testing.assert_allclose(dist[b].cpu().numpy(), d, rtol=0, atol=1e-5)
        assert np.array_equal(np.floor(end[b].cpu().numpy()), np.floor(e))


def test_propagate_full_size_properties(ctx, mazes):
    """BASELINE-size batch (2^20 candidates x 50 steps, 1.7 GB of traffic): size-independent properties.
    (1) a candidate's result does not depend on its position in the batch (permutation invariance),
    (2) the row and struct-of-arrays kernels give the same bits, (3) a strided sample of edges is
    teacher-forced against the oracle (flags bit-exact), (4) rows past the end of an edge are zero."""
    grid = mazes["boxes"]
    ctx.set_map(grid)
    rng = np.random.default_rng(99)
    B, S = 1 << 20, 50
    free = np.argwhere(grid == 0)
    cells = free[rng.integers(0, len(free), B)]
    x, y = orc.rowcol_to_xy(cells[:, 0], cells[:, 1], grid.shape)
    s0 = np.stack([x + rng.uniform(-0.3, 0.3, B), y + rng.uniform(-0.3, 0.3, B), rng.uniform(-np.pi, np.pi, B),
                   rng.uniform(0, 4, B), rng.uniform(0, 1.3, B), rng.uniform(-0.44, 0.44, B)], 1).astype(np.float32)
    act = torch.randn((B, S, 2), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * \
        torch.tensor([1.0, 0.92], device="cuda") + torch.tensor([0.45, 0.0], device="cuda")
    goal = (7.5, 7.5)
    st = dev(s0)
    res = ctx.propagate_collide(st, act, goal)
    perm = torch.randperm(B, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    res_p = ctx.propagate_collide(st[perm].contiguous(), act[perm].contiguous(), goal)
    for k in ("final", "first_coll", "done_step"):
        assert torch.equal(res[k][perm], res_p[k]), k
    assert torch.equal(res["traj"][perm[:4096]], res_p["traj"][:4096])
    soa = ctx.propagate_collide(st.t().contiguous(), act.permute(1, 2, 0).contiguous(), goal, soa=True, want_traj=False)
    assert torch.equal(soa["final"].t(), res["final"])
    assert torch.equal(soa["first_coll"], res["first_coll"]) and torch.equal(soa["done_step"], res["done_step"])
    sel = np.arange(0, B, 53)
    traj = res["traj"][torch.as_tensor(sel, device="cuda")].cpu().numpy()
    forced = orc.rollout_car(s0[sel], act[torch.as_tensor(sel, device="cuda")].cpu().numpy(), goal, grid, states_for_flags=traj)
    first = res["first_coll"].cpu().numpy()[sel]
    done = res["done_step"].cpu().numpy()[sel]
    assert np.array_equal(first, forced["first_coll"]) and np.array_equal(done, forced["done_step"])
    end = np.where(first >= 0, first, np.where(done >= 0, done, S - 1))
    for i in range(0, len(sel), 101):
        assert np.all(traj[i, end[i] + 1:] == 0)
    frac_coll = float((res["first_coll"] >= 0).float().mean())
    assert 0.2 < frac_coll < 0.9
