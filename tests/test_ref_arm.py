"""The reference arm (oracle/_ref: the UNMODIFIED reference, byte-compiled by oracle/build_ref.py) against the oracle
port on the benchmark's workload -- the two CPU baselines bench.py can report must be the same computation."""
import os

import numpy as np
import pytest
import torch

from oracle import denoiser_ref as dref
from oracle import ditree_oracle as orc
from oracle import ref_arm

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not ref_arm.available(), reason="oracle/_ref not staged (python oracle/build_ref.py)")
def test_staged_reference_matches_oracle(mazes, car_meta):
    import sys
    sys.path.insert(0, REPO)
    import bench
    grid = mazes["boxes"]
    dims = [64, 128, 256]
    sd = dref.init_params(seed=0, input_dim=2, cond_dim=7, emb_dim=400, down_dims=dims)
    goal = bench.goal_of(grid)
    K, S, B = 3, 50, 32
    run = ref_arm.Reference().expansion(grid, sd, dims, K, S, goal)
    st, prev = bench.synth_candidates(grid, B, 7)
    r = run(st, prev, seed=5)
    noise = torch.randn(B, 64, 2, generator=torch.manual_seed(5))
    s64 = st.astype(np.float64)
    lm = orc.local_map(grid, s64[:, 0], s64[:, 1], s64[:, 2], 20, 0.2, 1.0, (10.0, 10.0))
    assert np.array_equal(lm, r["local_map"])
    cond = orc.build_cond_car(s64, prev.astype(np.float64), goal, car_meta, 20.0)
    act = dref.fm_sample(sd, noise, torch.from_numpy(cond), torch.from_numpy(lm), K, car_meta["Actions_mean"],
                         car_meta["Actions_std"])
    assert np.abs(act - r["actions"]).max() < 1e-4
    o = orc.rollout_car(s64, r["actions"][:, :S], goal, grid)
    assert np.array_equal(o["first_coll"], r["first_coll"]) and np.array_equal(o["done_step"], r["done_step"])
    assert (r["first_coll"] >= 0).any() and (r["first_coll"] < 0).any()
    np.testing.assert_allclose(o["final"], r["final"], rtol=0, atol=1e-12)


def test_manifest_holds_no_sources():
    """oracle/_ref is compiled output only (and git-ignored): no reference source file may sit in it."""
    d = ref_arm.REF_DIR
    if not os.path.isdir(d):
        pytest.skip("oracle/_ref not staged")
    for root, _, files in os.walk(d):
        for f in files:
            assert not f.endswith(".py"), os.path.join(root, f)
    assert "oracle/_ref/" in open(os.path.join(REPO, ".gitignore")).read()
