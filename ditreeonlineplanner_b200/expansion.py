"""One batched tree expansion on the device: the body of ``RRT_Planner.plan()``'s inner chunk
(planners/RRT.py:157-178) for B candidate nodes at once:

    create_local_map (common/map_utils.py:391-459)  ->  DiffusionSampler.forward (policies/fm_policy.py:53-212)
    ->  propagate_action_sequence_env + is_colliding_car (planners/base_planner.py:257-320)

Everything between the host inputs and the host outputs stays in HBM; the only host work is
enqueueing ~60 kernels per ODE step on torch's current stream.
"""
from __future__ import annotations

import numpy as np
import torch

from .runtime import Context


class TreeExpander:
    def __init__(self, ctx: Context, meta, local_map_size=20, local_map_scale=0.2, num_diffusion_iters=1,
                 pred_horizon=64, action_horizon=8, action_dim=2):
        self.ctx = ctx
        self.meta = meta
        self.N = int(local_map_size)
        self.scale = float(local_map_scale)
        self.K = int(num_diffusion_iters)
        self.T = int(pred_horizon)
        self.S = int(action_horizon)
        self.A = int(action_dim)
        self._pin = {}

    # ---- device-resident pass ---------------------------------------------------------------
    def expand_device(self, states, prev_actions, goal_xy, noise=None, want_traj=True, generator=None):
        """states (B,6) f32 cuda rows; prev_actions (B,2) f32 cuda or None; goal_xy (2,) floats.
        Returns dict of device tensors: actions (B,T,A), traj (B,S,6)|None, final (B,6), first_coll (B,),
        done_step (B,)."""
        ctx = self.ctx
        B = states.shape[0]
        lm = ctx.local_map(states, self.N, self.scale, bf16_signed=True)
        goal = torch.as_tensor(np.asarray(goal_xy, dtype=np.float32), device=ctx.device)
        cond = ctx.build_cond_car(states, prev_actions, goal, self.meta, float(self.N))
        if noise is None:
            # the reference draws the initial sample with torch.randn on the sampler's device (fm_policy.py:158)
            noise = torch.randn((B, self.T, self.A), device=ctx.device, generator=generator)
        actions = ctx.fm_sample(noise, cond, lm, self.K, self.meta["Actions_mean"], self.meta["Actions_std"])
        res = ctx.propagate_collide(states, actions, goal_xy, S=self.S, want_traj=want_traj)
        res["actions"] = actions
        return res

    # ---- host-facing pass (what a planner calls with NumPy arrays) ----------------------------
    def _pinned(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        if key not in self._pin:
            self._pin[key] = torch.empty(shape, dtype=dtype).pin_memory()
        return self._pin[key]

    def expand(self, states, prev_actions, goal_xy, want_traj=False):
        """NumPy in, NumPy out; host<->device copies through pinned staging buffers.
        Returns (dict of ndarrays, h2d_bytes, d2h_bytes)."""
        ctx = self.ctx
        st = np.asarray(states, dtype=np.float32)
        B = st.shape[0]
        h_st = self._pinned("st", (B, 6), torch.float32)
        h_st.numpy()[...] = st
        d_st = h_st.to(ctx.device, non_blocking=True)
        h2d = h_st.numel() * 4
        d_prev = None
        if prev_actions is not None:
            h_pa = self._pinned("pa", (B, 2), torch.float32)
            h_pa.numpy()[...] = np.asarray(prev_actions, dtype=np.float32)
            d_prev = h_pa.to(ctx.device, non_blocking=True)
            h2d += h_pa.numel() * 4
        res = self.expand_device(d_st, d_prev, goal_xy, want_traj=want_traj)
        out = {}
        d2h = 0
        for k in ("final", "first_coll", "done_step", "actions") + (("traj",) if want_traj else ()):
            t = res[k] if k != "actions" else res[k][:, : self.S].contiguous()
            h = self._pinned("o_" + k, tuple(t.shape), t.dtype)
            h.copy_(t, non_blocking=True)
            out[k] = h
            d2h += t.numel() * t.element_size()
        torch.cuda.current_stream(ctx.device).synchronize()
        return {k: v.numpy() for k, v in out.items()}, h2d, d2h
