"""Multi-process (gloo, world_size 2, CPU) test of the scenario-suite sharding and result gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ditreeonlineplanner_b200 import scenarios as sc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_unit(row, s, r, sampler, time_budget, planner_kwargs):
    sc.seed_everything(sc.unit_seed(s, r))
    return [r + 1, s % 2, 0.1, float(np.random.rand()), 0.0, 1.0, 10 + s, 100 + r, 1.0, 0.5, 0.1]


_RAN = []


def _recording_unit(row, s, r, sampler, time_budget, planner_kwargs):
    import time
    _RAN.append((s, r))
    time.sleep(0.002 * (1 + s % 3))  # uneven units: the queue must still hand out each exactly once
    return _fake_unit(row, s, r, sampler, time_budget, planner_kwargs)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table, _ = sc.run_suite(None, total_runs=3, rank=rank, world=world, device="cpu", unit_fn=_recording_unit)
    queued = list(_RAN)
    static, _ = sc.run_suite(None, total_runs=3, rank=rank, world=world, device="cpu", unit_fn=_fake_unit,
                             schedule="static")
    again, _ = sc.run_suite(None, total_runs=2, rank=rank, world=world, device="cpu", unit_fn=_fake_unit)  # fresh counter
    q.put((rank, sorted(table.items()), sorted(static.items()), queued, len(again)))
    dist.destroy_process_group()


def test_partition_is_exact_cover():
    for world in (1, 2, 4, 8):
        units = [u for r in range(world) for u in sc.shard_units(15, 10, r, world)]
        assert sorted(units) == sorted(sc.all_units(15, 10)) and len(set(units)) == 150
        sizes = [len(sc.shard_units(15, 10, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1  # unit-level sharding balances 150 units over 8 ranks


def test_queue_order_is_heaviest_first_and_complete():
    w = [9, 400, 25, 400, 961]
    order = sc.all_units(5, 3, w, by_weight=True)
    assert sorted(order) == sorted(sc.all_units(5, 3)) and len(set(order)) == 15
    assert [s for s, _ in order[:3]] == [4, 4, 4] and [r for _, r in order[:3]] == [0, 1, 2]  # all runs of the largest map first
    weights_seen = [w[s] for s, _ in order]
    assert weights_seen == sorted(weights_seen, reverse=True)


def test_unit_seeds_are_world_size_invariant():
    seeds = {(s, r): sc.unit_seed(s, r) for s in range(15) for r in range(10)}
    assert len(set(seeds.values())) == 150


def test_gather_world2_matches_world1():
    single, _ = sc.run_suite(None, total_runs=3, rank=0, world=1, device="cpu", unit_fn=_fake_unit)
    assert len(single) == 45 and all(len(v) == len(sc.ROW_FIELDS) for v in single.values())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, items, static_items, queued, n_again in got:
        assert n_again == 30
        for table in (items, static_items):  # shared-queue deal and round-robin deal: the same rows on every rank
            assert dict(table).keys() == single.keys()
            for k, v in table:
                np.testing.assert_allclose(v, single[k], rtol=1e-6)
    ran = [u for g in got for u in g[3]]
    assert sorted(ran) == sorted(single.keys()) and len(set(ran)) == 45  # every unit ran on exactly one rank
    assert all(len(g[3]) > 0 for g in got)


def test_result_row_schema():
    path = np.array([[0, 0, 0, 1.0, 0, 0], [1, 0, 0, 1.0, 0, 0], [1, 1, 0, 1.0, 0, 0]], np.float32)
    acts = np.array([[1.0, 0.0], [0.0, 2.0]], np.float32)
    row = sc.result_row(0, path, acts, {"path_time": 0.06, "number_of_nodes": 3, "iterations": 16}, 1.5)
    assert len(row) == 11 and row[1] == 1 and abs(row[3] - 2.0) < 1e-6 and row[8] == 2.0
    assert sc.result_row(4, None, None, {"iterations": 7, "number_of_nodes": 2}, 0.5)[:2] == [5, 0]


def test_reference_csv_wire_format(tmp_path):
    """The per-scenario CSV the reference writes (run_scenarios.py:331-333,392-395): 10 header cells, 11 values
    per row, ints for counters / sentinels, append-resume; readable the way results_process.py reads it."""
    import csv
    from ditreeonlineplanner_b200 import scenarios as sc
    from ditreeonlineplanner_b200.data import load_scenarios
    rows = load_scenarios("test_scenarios_car")
    ok = sc.result_row(0, np.array([[0, 0, 1.0, 0.5, 0, 0], [3, 4, 1.0, 0.5, 0, 0]], float), np.array([[3.0, 4.0], [0.0, 1.0]]),
                       {"number_of_nodes": 17, "iterations": 285, "path_time": 1.28}, 2.5)
    fail = sc.result_row(1, None, None, {"number_of_nodes": 40, "iterations": 300}, 4.0)
    table = {(0, 0): ok, (0, 1): fail, (3, 0): ok}
    paths = sc.write_suite_csv(table, str(tmp_path), current_run=1)
    assert [p.split("/")[-1] for p in paths] == [f"{rows[0]['scenario_name']}_diffusion_RRT_PD64_carmaze.csv",
                                                 f"{rows[3]['scenario_name']}_diffusion_RRT_PD64_carmaze.csv"]
    with open(paths[0], newline="") as f:
        lines = list(csv.reader(f))
    assert lines[0] == ["iteration", "success", "runtime", "trajectory_length", "avg_velocity", "num_states_in_tree",
                        "num_RRT_iterations", "ctrl_effort_max", "ctrl_effort_mean", "ctrl_effort_std"]
    assert len(lines) == 3 and all(len(r) == 11 for r in lines[1:])
    assert lines[1][:2] == ["1", "1"] and float(lines[1][3]) == 5.0 and lines[1][6:8] == ["17", "285"]
    assert lines[2] == ["2", "0", "4.0", "-1", "0", "-1", "-1", "300", "-1", "-1", "-1"]   # the reference's failure row
    # resume: rows already present are not rewritten, new runs are appended
    table[(0, 2)] = sc.result_row(2, None, None, {"number_of_nodes": 1, "iterations": 5}, 1.0)
    sc.write_suite_csv(table, str(tmp_path), current_run=1)
    assert sc.existing_rows(paths[0]) == 3
    assert sc.next_run_index(str(tmp_path)) == 2
    sc.save_path_csv(np.array([[1.0, 2.5], [3.25, -4.0]]), str(tmp_path / "path_DP_0.csv"))
    assert open(tmp_path / "path_DP_0.csv").read() == "1.000000,2.500000\n3.250000,-4.000000\n"


def test_checkpoint_loader(tmp_path):
    import torch
    from oracle import denoiser_ref as dref
    from ditreeonlineplanner_b200 import scenarios as sc
    sd = dref.init_params(seed=1, input_dim=2, cond_dim=7, emb_dim=400, down_dims=[64, 128, 256])
    torch.save({"noise_pred_net_state_dict": sd, "epoch": 3, "optimizer_state_dict": {}}, tmp_path / "ck.pt")
    got = sc.load_checkpoint_state_dict(str(tmp_path / "ck.pt"))
    assert set(got) == set(sd) and all(torch.equal(got[k], torch.as_tensor(sd[k]).float()) for k in sd)
    torch.save({"something": torch.zeros(1)}, tmp_path / "bad.pt")
    import pytest
    with pytest.raises(KeyError):
        sc.load_checkpoint_state_dict(str(tmp_path / "bad.pt"))
