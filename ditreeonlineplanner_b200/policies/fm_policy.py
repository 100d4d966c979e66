"""Drop-in mirror of the reference's ``policies/fm_policy.py::DiffusionSampler`` (constructor
:11-24, forward :53-212) whose forward pass runs on the sm_100a denoiser.

Same constructor arguments, same call signature and return value
(``ndarray (B, pred_horizon, action_dim)`` float64), same exceptions (FileNotFoundError for
unknown env metadata, NotImplementedError for policies other than 'flow_matching').
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from ..data import load_metadata
from ..runtime import get_context


def _infer_cfg(sd):
    dims = []
    i = 0
    while f"unet.down_modules.{i}.0.blocks.0.block.0.weight" in sd:
        dims.append(int(sd[f"unet.down_modules.{i}.0.blocks.0.block.0.weight"].shape[0]))
        i += 1
    emb = int(sd["encoder.resnet18.fc.weight"].shape[0])
    gdim = int(sd["unet.mid_modules.0.cond_encoder.1.weight"].shape[1])
    dsed = int(sd["unet.diffusion_step_encoder.3.weight"].shape[0])
    A = int(sd["unet.final_conv.1.weight"].shape[0])
    return dims, emb, gdim - dsed - emb, A


_tokens = []   # [(state_dict object, token)]: a monotonic token per distinct weights object, kept alive here so that
_next_token = [0]  # an id() can never be recycled for another dict (what keying on id() alone would risk)


def _weights_token(sd):
    for obj, tok in _tokens:
        if obj is sd:
            return tok
    _next_token[0] += 1
    _tokens.append((sd, _next_token[0]))
    if len(_tokens) > 8:      # bounded: old weights are released, their tokens never reused
        _tokens.pop(0)
    return _next_token[0]


class DiffusionSampler(nn.Module):
    def __init__(self, noise_pred_net, noise_scheduler, env_id, policy, pred_horizon, action_dim,
                 prediction_type="actions", obs_history=1, action_history=1, num_diffusion_iters=100,
                 position_conditioned=False, goal_conditioned=True, local_map_conditioned=True, local_map_size=16,
                 max_batch=4096):
        super().__init__()
        self.device = "cuda" if torch.cuda.is_available() else "cpu"
        self.metadata = load_metadata(env_id)  # FileNotFoundError like fm_policy.py:32
        self.action_dim = action_dim
        self.prediction_type = prediction_type
        self.env_id = env_id
        self.policy = policy
        self.num_diffusion_iters = num_diffusion_iters
        self.pred_horizon = pred_horizon
        self.obs_history = obs_history
        self.action_history = action_history
        self.position_conditioned = position_conditioned
        self.goal_conditioned = goal_conditioned
        self.local_map_conditioned = local_map_conditioned
        self.local_map_size = local_map_size
        self.noise_scheduler = noise_scheduler
        # `noise_pred_net` is the reference module (anything with state_dict()) or a state_dict
        self._state_dict = noise_pred_net.state_dict() if hasattr(noise_pred_net, "state_dict") else dict(noise_pred_net)
        self._max_batch = max_batch
        self._ctx = None
        self._key = None

    _accepts_signed_bf16_map = True   # forward() takes the planners' pre-scaled bf16 local map (RRT_Planner._local_map)

    # nn.Module.to()/eval() keep working; the packed weights live in the device context
    def _context(self):
        """The process-wide device context with THIS sampler's weights packed in it.  Contexts are shared per
        device, so another sampler (other weights, horizon or map size) may have re-packed it since the last call:
        the key is compared on every call (one tuple comparison) and the weights are re-packed when it differs.
        Samplers built from the same state_dict object share one key, hence one packing."""
        ctx = get_context()
        if self._key is None:
            dims, emb, cond_dim, A = _infer_cfg(self._state_dict)
            if A != self.action_dim:
                raise ValueError(f"action_dim {self.action_dim} does not match the network ({A})")
            self._cfg = dict(action_dim=A, horizon=self.pred_horizon, cond_dim=cond_dim, emb_dim=emb,
                             map_size=int(self.local_map_size), down_dims=dims, max_batch=self._max_batch)
            self._key = (_weights_token(self._state_dict), self.pred_horizon, int(self.local_map_size), self._max_batch)
        if getattr(ctx, "_loaded_key", None) != self._key:
            ctx.load_denoiser(self._state_dict, **self._cfg)
            ctx._loaded_key = self._key
        self._ctx = ctx
        return ctx

    def _twin_context(self, max_batch):
        """A second device context holding the same packed weights (own activation arena), so two independent
        sampler passes can run concurrently on two streams (the continuous-refill planner's two slot groups)."""
        twin = getattr(self, "_twin", None)
        if twin is None or self._twin_batch < max_batch:
            from ..runtime import Context
            main = self._context()
            if twin is not None:
                twin.close()
            twin = Context(main.device.index)
            dims, emb, cond_dim, A = _infer_cfg(self._state_dict)
            twin.load_denoiser(self._state_dict, action_dim=A, horizon=self.pred_horizon, cond_dim=cond_dim, emb_dim=emb,
                               map_size=int(self.local_map_size), down_dims=dims, max_batch=int(max_batch))
            self._twin, self._twin_batch = twin, int(max_batch)
        return self._twin

    def _check_supported(self):
        if self.policy != "flow_matching":
            raise NotImplementedError("only the flow_matching policy runs on the B200 path")
        if self.prediction_type != "actions" or self.position_conditioned or not self.goal_conditioned \
                or not self.local_map_conditioned or self.action_history != 1:
            raise NotImplementedError("the B200 path implements the reference's carmaze/antmaze 'actions' configuration")

    def build_cond(self, obs_seq, prev_actions, goal, ctx=None):
        """Condition vectors (B,G) on the device (fm_policy.py:71-143)."""
        ctx = ctx if ctx is not None else self._context()
        env = self.env_id.lower()
        if "car" in env:
            if self.obs_history != 1:
                raise NotImplementedError("carmaze uses obs_history = 1")
            # states | previous actions | goal(s) cross the bus as ONE array (three small copies cost 3 x ~10 us per
            # call of the reference's B = 1 loop)
            B = len(obs_seq)
            g = np.asarray(goal, dtype=np.float32)
            n_prev = 0 if prev_actions is None else 2 * B
            flat = np.empty(6 * B + n_prev + g.size, dtype=np.float32)
            flat[:6 * B] = obs_seq[:, -1, :].reshape(-1)
            if n_prev:
                flat[6 * B: 8 * B] = prev_actions[:, -1, :].reshape(-1)
            flat[6 * B + n_prev:] = g.reshape(-1)
            dev = torch.from_numpy(flat).to(ctx.device)
            return ctx.build_cond_car(dev[:6 * B].view(B, 6), dev[6 * B: 8 * B].view(B, 2) if n_prev else None,
                                      dev[6 * B + n_prev:].view(g.shape), self.metadata, float(self.local_map_size))
        if "ant" in env:
            seq = np.ascontiguousarray(obs_seq[:, -self.obs_history:, :], dtype=np.float32)
            prev = None if prev_actions is None else np.ascontiguousarray(prev_actions[:, -1, :], dtype=np.float32)
            return ctx.build_cond_ant(torch.as_tensor(seq), None if prev is None else torch.as_tensor(prev),
                                      torch.as_tensor(np.asarray(goal, dtype=np.float32)), self.metadata,
                                      self.obs_history, float(self.local_map_size))
        raise NotImplementedError(f"env {self.env_id!r} is not on the B200 path")

    def forward(self, obs_seq, prev_actions, goal=None, local_map=None, noise=None):
        """obs_seq (B, obs_history, obs_dim) | (B, obs_dim) | (obs_dim,); prev_actions
        (B, action_history, A) | (action_history, A) | None; goal (2,) | (B,2); local_map (B,N,N)
        ndarray / tensor in {0,1}.  Returns ndarray (B, pred_horizon, A) float64."""
        self._check_supported()
        obs_seq = np.array(obs_seq, dtype=np.float64)  # copy: inputs are never mutated (fm_policy.py:60)
        if obs_seq.ndim == 1:
            obs_seq = obs_seq[None]
        if obs_seq.ndim == 2:
            obs_seq = obs_seq[:, None, :]
        if prev_actions is not None:
            prev_actions = np.asarray(prev_actions, dtype=np.float64)
            if prev_actions.ndim == 2:
                prev_actions = prev_actions[None]
        B = len(obs_seq)
        ctx = self._context()
        cond = self.build_cond(obs_seq, prev_actions, goal, ctx)
        lm = None
        if getattr(local_map, "_ditree_signed_bf16", False) and local_map.dtype == torch.bfloat16 \
                and local_map.device == ctx.device:
            lm = local_map                                     # already 2 m - 1 in bf16 (the planners' _local_map)
        if lm is None:
            if isinstance(local_map, np.ndarray):
                local_map = torch.from_numpy(local_map)
            local_map = local_map.to(ctx.device, dtype=torch.float32)
            if local_map.dim() == 2:
                local_map = local_map.unsqueeze(0)
            lm = (local_map * 2 - 1).to(torch.bfloat16)  # scale to [-1, 1] (fm_policy.py:152)
        elif lm.dim() == 2:
            lm = lm.unsqueeze(0)
        if noise is None:
            noise = torch.randn((B, self.pred_horizon, self.action_dim), device=ctx.device)
        naction = ctx.fm_sample(noise, cond, lm, self.num_diffusion_iters)
        naction = naction.detach().to("cpu").numpy()
        return naction * self.metadata["Actions_std"] + self.metadata["Actions_mean"]
