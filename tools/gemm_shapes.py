#!/usr/bin/env python
"""Per-launch GEMM efficiency of one planner pass (K = 1, large denoiser) at a given batch: the library's own
CUDA-event profiler (M, N, K, epilogue, ms per launch) -> TFLOP/s per launch, sorted by time."""
import csv, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_candidates
from ditreeonlineplanner_b200 import get_context, load_maze, load_metadata
from ditreeonlineplanner_b200.weights import UNET_DIMS, random_init
ctx = get_context(0)
grid = load_maze("boxes").astype(np.float32); ctx.set_map(grid); meta = load_metadata("carmaze")
dims = UNET_DIMS["large"]
ctx.load_denoiser(random_init(seed=0, input_dim=2, cond_dim=7, emb_dim=400, down_dims=dims), action_dim=2, horizon=64,
                  cond_dim=7, emb_dim=400, map_size=20, down_dims=dims, max_batch=4096)
B = int(os.environ.get("B", "256"))
st, prev = synth_candidates(grid, B, 1)
st = torch.as_tensor(st).cuda(); prev = torch.as_tensor(prev).cuda()
goal = torch.as_tensor(np.array([7.5, 7.5], np.float32)).cuda()
noise = torch.randn((B, 64, 2), device="cuda")
def one():
    lm = ctx.local_map(st, 20, 0.2, bf16_signed=True)
    cond = ctx.build_cond_car(st, prev, goal, meta, 20.0)
    return ctx.fm_sample(noise, cond, lm, 1, meta["Actions_mean"], meta["Actions_std"])
for _ in range(3): one()
torch.cuda.synchronize()
ctx.profile_begin(); one(); torch.cuda.synchronize()
ms, n = ctx.profile_end()
path = os.environ.get("OUT", "gpurun_out/gemm_shapes.csv")
ctx.profile_csv(path)
recs = list(csv.DictReader(open(path)))
print(f"B={B}: {n} GEMM launches, {ms:.3f} ms")
tot = 0.0
for r in sorted(recs, key=lambda r: -float(r["ms"])):
    M, N, K, t = int(r["M"]), int(r["N"]), int(r["K"]), float(r["ms"])
    print(f"M={M:6d} N={N:5d} K={K:5d} epi={r['epi']} {t*1e3:7.1f} us {2.0*M*N*K/t/1e9:8.1f} TFLOP/s")
