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


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table, _ = sc.run_suite(None, total_runs=3, rank=rank, world=world, device="cpu", unit_fn=_fake_unit)
    q.put((rank, sorted(table.items())))
    dist.destroy_process_group()


def test_partition_is_exact_cover():
    for world in (1, 2, 4, 8):
        units = [u for r in range(world) for u in sc.shard_units(15, 10, r, world)]
        assert sorted(units) == sorted(sc.all_units(15, 10)) and len(set(units)) == 150
        sizes = [len(sc.shard_units(15, 10, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1  # unit-level sharding balances 150 units over 8 ranks


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
    for rank, items in got:
        assert dict(items).keys() == single.keys()
        for k, v in items:
            np.testing.assert_allclose(v, single[k], rtol=1e-6)


def test_result_row_schema():
    path = np.array([[0, 0, 0, 1.0, 0, 0], [1, 0, 0, 1.0, 0, 0], [1, 1, 0, 1.0, 0, 0]], np.float32)
    acts = np.array([[1.0, 0.0], [0.0, 2.0]], np.float32)
    row = sc.result_row(0, path, acts, {"path_time": 0.06, "number_of_nodes": 3, "iterations": 16}, 1.5)
    assert len(row) == 11 and row[1] == 1 and abs(row[3] - 2.0) < 1e-6 and row[8] == 2.0
    assert sc.result_row(4, None, None, {"iterations": 7, "number_of_nodes": 2}, 0.5)[:2] == [5, 0]
