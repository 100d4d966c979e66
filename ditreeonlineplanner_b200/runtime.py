"""Device runtime: a thin object over the C ABI that takes torch CUDA tensors (PyTorch is only
the owner of device memory and streams here) and enqueues the sm_100a kernels on torch's current
stream.  No computation happens in Python and nothing falls back to the CPU."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Context:
    """One per device: owns the staged occupancy grid, the packed denoiser and scratch."""

    def __init__(self, device=0):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise RuntimeError("ditreeonlineplanner_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)  # make sure the primary context exists
        h = C.c_void_p()
        rc = self.lib.dt_ctx_create(self.device.index, C.byref(h))
        if rc != 0:
            raise L.DitreeError(rc, "dt_ctx_create failed (is this an sm_100 device?)")
        self.h = h
        self.map_shape = None
        self.s_global = 1.0
        self.model_cfg = None

    # -- plumbing ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.dt_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        # raw handle of torch's current stream on this device (the C call behind torch.cuda.current_stream,
        # without building a Stream object: this runs five times per planner pass)
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(self.device.index))

    def _norm_vec(self, meta):
        """[obs mean | obs std | action mean | action std] float64, built once per metadata dict."""
        cache = self.__dict__.setdefault("_norm_cache", {})
        hit = cache.get(id(meta))
        if hit is None or hit[0] is not meta:
            vec = np.concatenate([meta["Observations_mean"], meta["Observations_std"], meta["Actions_mean"],
                                  meta["Actions_std"]]).astype(np.float64)
            hit = cache[id(meta)] = (meta, vec)
        return hit[1]

    def _check(self, rc):
        if rc != 0:
            raise L.DitreeError(rc, self.lib.dt_last_error(self.h).decode())

    def _f32(self, t):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t, dtype=np.float32))
        if t.device != self.device or t.dtype != torch.float32:
            t = t.to(self.device, torch.float32)
        return t

    def sync_status(self):
        """Raise what the reference would have raised (IndexError) for earlier launches."""
        rc = self.lib.dt_sync_status(self.h, self._stream())
        if rc == L.DT_E_INDEX:
            raise IndexError(self.lib.dt_last_error(self.h).decode())
        self._check(rc)

    def profile_begin(self):
        self._check(self.lib.dt_profile_begin(self.h))

    def profile_end(self):
        """-> (summed GEMM kernel milliseconds, GEMM launches) since profile_begin."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        self._check(self.lib.dt_profile_end(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile_csv(self, path):
        self._check(self.lib.dt_profile_csv(self.h, str(path).encode()))

    def profile_records(self):
        """Per-launch records of the last profile_begin/profile_end window: list of dicts with the CSV's columns
        (BN, epi: 0 bias / 1 GroupNorm+Mish / 2 split-K reduction, gw = 10 * group width + CTAs per tile, M, N, K,
        ms, tflops, ksplit)."""
        import csv
        import os
        import tempfile
        fd, path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        try:
            self.profile_csv(path)
            with open(path) as f:
                return [{k: (float(v) if k in ("ms", "tflops") else int(v)) for k, v in row.items()}
                        for row in csv.DictReader(f)]
        finally:
            os.unlink(path)

    @property
    def launches(self):
        return int(self.lib.dt_launch_count(self.h))

    def set_option(self, name, value):
        self._check(self.lib.dt_set_option(self.h, name.encode(), int(value)))

    # -- map --------------------------------------------------------------------------------
    def set_map(self, grid, s_global=1.0):
        g = np.ascontiguousarray(np.asarray(grid, dtype=np.float32))
        if g.ndim != 2:
            raise ValueError("grid must be 2-D")
        self._check(self.lib.dt_set_map(self.h, g.ctypes.data_as(C.c_void_p), g.shape[0], g.shape[1],
                                        float(s_global), self._stream()))
        self.map_shape = g.shape
        self.s_global = float(s_global)

    def set_map_slot(self, slot, grid, s_global=1.0):
        """Stage `grid` in map slot `slot` (0..31) next to the main map: the per-group maps of multi-scenario passes."""
        g = np.ascontiguousarray(np.asarray(grid, dtype=np.float32))
        if g.ndim != 2:
            raise ValueError("grid must be 2-D")
        self._check(self.lib.dt_set_map_slot(self.h, int(slot), g.ctypes.data_as(C.c_void_p), g.shape[0], g.shape[1],
                                             float(s_global), self._stream()))

    def local_map_slots(self, poses, n, scale, slot_of_group, group_size):
        """poses (B,>=3) rows in groups of `group_size`, group g cropping from map slot slot_of_group[g]
        -> (B,n,n) bf16 holding 2m-1."""
        st = self._f32(poses)
        (x, y, th), stride = self._xyz(st)
        B = st.shape[0]
        sog = torch.as_tensor(np.asarray(slot_of_group, dtype=np.int32)).to(self.device) \
            if not isinstance(slot_of_group, torch.Tensor) else slot_of_group.to(self.device, torch.int32)
        out = torch.empty((B, n, n), dtype=torch.bfloat16, device=self.device)
        self._check(self.lib.dt_local_map_slots(self.h, _ptr(x), _ptr(y), _ptr(th), stride, B, int(n), float(scale),
                                                _ptr(sog.contiguous()), int(group_size), _ptr(out), self._stream()))
        return out

    # -- geometry ---------------------------------------------------------------------------
    @staticmethod
    def _xyz(states, cols=(0, 1, 2)):
        """(B,>=3) row tensor -> three column views + the common element stride."""
        assert states.dim() == 2 and states.stride(1) == 1
        return tuple(states[:, c] for c in cols), states.stride(0)

    def collide_car(self, states):
        st = self._f32(states)
        (x, y, th), stride = self._xyz(st)
        out = torch.empty(st.shape[0], dtype=torch.uint8, device=self.device)
        self._check(self.lib.dt_collide_car(self.h, _ptr(x), _ptr(y), _ptr(th), stride, st.shape[0], _ptr(out),
                                            self._stream()))
        return out

    def collide_points(self, points, scale=1.0, r=0.1):
        p = self._f32(points)
        assert p.dim() == 2 and p.stride(1) == 1
        out = torch.empty(p.shape[0], dtype=torch.uint8, device=self.device)
        self._check(self.lib.dt_collide_points(self.h, _ptr(p[:, 0]), _ptr(p[:, 1]), p.stride(0), p.shape[0],
                                               float(scale), float(r), _ptr(out), self._stream()))
        return out

    def collide_ant(self, states, radius=1.2):
        st = self._f32(states)
        assert st.dim() == 2 and st.shape[1] >= 7 and st.stride(1) == 1
        out = torch.empty(st.shape[0], dtype=torch.uint8, device=self.device)
        self._check(self.lib.dt_collide_ant(self.h, _ptr(st), st.stride(0), st.shape[0], float(radius), _ptr(out),
                                            self._stream()))
        return out

    def local_map(self, poses, n, scale, bf16_signed=False):
        """poses (B,>=3) rows (x, y, theta) -> (B,n,n) float32 {0,1}, or bf16 2m-1 for the encoder."""
        st = self._f32(poses)
        (x, y, th), stride = self._xyz(st)
        B = st.shape[0]
        out = torch.empty((B, n, n), dtype=torch.bfloat16 if bf16_signed else torch.float32, device=self.device)
        self._check(self.lib.dt_local_map(self.h, _ptr(x), _ptr(y), _ptr(th), stride, B, int(n), float(scale),
                                          L.DT_BF16 if bf16_signed else L.DT_F32, _ptr(out), self._stream()))
        return out

    def ray_probe(self, states):
        st = self._f32(states)
        (x, y, th), stride = self._xyz(st)
        out = torch.empty(st.shape[0], dtype=torch.uint8, device=self.device)
        self._check(self.lib.dt_ray_probe(self.h, _ptr(x), _ptr(y), _ptr(th), stride, st.shape[0], _ptr(out),
                                          self._stream()))
        return out

    def path_first_obstacle(self, path_xy):
        p = self._f32(path_xy)
        out = torch.empty(1, dtype=torch.int32, device=self.device)
        self._check(self.lib.dt_path_first_obstacle(self.h, _ptr(p[:, 0]), _ptr(p[:, 1]), p.stride(0), p.shape[0],
                                                    _ptr(out), self._stream()))
        return out

    def path_first_obstacle_grid(self, grid_u8, path_xy):
        """Same test against a caller-supplied (rows, cols) uint8 grid (the online driver's scanned map, values
        0 / 1 / 2) instead of the staged planner map, which stays resident."""
        g = grid_u8 if isinstance(grid_u8, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(grid_u8, dtype=np.uint8))
        g = g.to(self.device, torch.uint8).contiguous()
        p = self._f32(path_xy)
        out = torch.empty(1, dtype=torch.int32, device=self.device)
        self._check(self.lib.dt_path_first_obstacle_grid(self.h, _ptr(g), g.shape[0], g.shape[1], _ptr(p[:, 0]),
                                                         _ptr(p[:, 1]), p.stride(0), p.shape[0], _ptr(out), self._stream()))
        return out

    def lidar_scan(self, poses, want_visited=True):
        p = self._f32(poses).contiguous()
        assert p.dim() == 2 and p.shape[1] == 3
        B = p.shape[0]
        dist = torch.empty((B, 181), dtype=torch.float64, device=self.device)
        end = torch.empty((B, 181, 2), dtype=torch.float64, device=self.device)
        vis = None
        if want_visited:
            vis = torch.empty((B, self.map_shape[0], self.map_shape[1]), dtype=torch.uint8, device=self.device)
        self._check(self.lib.dt_lidar_scan(self.h, _ptr(p), B, _ptr(dist), _ptr(end), _ptr(vis), self._stream()))
        return dist, end, vis

    # -- dynamics ---------------------------------------------------------------------------
    def propagate_collide(self, state0, actions, goal_xy, S=None, want_traj=True, soa=False,
                          stop_on_collision=True, packed_out=False):
        """state0: (B,6) rows, or (6,B) when soa.  actions: (B,T,2) rows (T >= S), or (S,2,B) when soa.
        Returns dict(traj, final, first_coll, done_step); traj is (B,S,6) rows or (S,6,B) when soa."""
        s0 = self._f32(state0)
        act = self._f32(actions)
        if soa:
            B = s0.shape[1]
            assert s0.shape[0] == 6 and s0.is_contiguous() and act.is_contiguous() and act.shape[1:] == (2, B)
            S = act.shape[0] if S is None else S
            s_str = (1, B)
            a_str = (1, 2 * B, B)
            final = torch.empty_like(s0)
            traj = torch.empty((S, 6, B), dtype=torch.float32, device=self.device) if want_traj else None
            t_str = (1, 6 * B, B)
        else:
            B = s0.shape[0]
            assert s0.shape[1] == 6 and s0.is_contiguous() and act.dim() == 3 and act.shape[2] == 2
            assert act.stride(2) == 1
            S = act.shape[1] if S is None else S
            assert S <= act.shape[1]
            s_str = (6, 1)
            a_str = (act.stride(0), act.stride(1), 1)
            final = None if (packed_out and want_traj) else torch.empty_like(s0)
            # (B, S, 6) rows; the row pitch is padded to a whole number of 32-byte sectors so that every
            # 4-step piece the kernel writes covers complete sectors (1200-byte rows would leave every other
            # row straddling them: partial-sector writes cost L2 read-modify-writes)
            pitch = -(-(S * 6) // 8) * 8
            traj = None
            if want_traj and not packed_out:
                traj = torch.empty((B, pitch), dtype=torch.float32, device=self.device)[:, : S * 6].unflatten(1, (S, 6))
            t_str = (pitch, 6, 1)
        if packed_out and not soa and want_traj:
            # one allocation [traj rows | final | first | done] so that a caller who needs everything on the host (the
            # reference's B = 1 loop) fetches it with ONE device->host copy instead of four
            blob = torch.empty(B * pitch + B * 8, dtype=torch.float32, device=self.device)
            traj = blob[: B * pitch].view(B, pitch)[:, : S * 6].unflatten(1, (S, 6))
            final = blob[B * pitch: B * pitch + B * 6].view(B, 6)
            first = blob[B * pitch + B * 6: B * pitch + B * 7].view(torch.int32)
            done = blob[B * pitch + B * 7:].view(torch.int32)
        else:
            first = torch.empty(B, dtype=torch.int32, device=self.device)
            done = torch.empty(B, dtype=torch.int32, device=self.device)
        self._check(self.lib.dt_propagate_collide(
            self.h, _ptr(s0), s_str[0], s_str[1], _ptr(act), a_str[0], a_str[1], a_str[2], B, int(S),
            float(goal_xy[0]), float(goal_xy[1]), _ptr(traj), t_str[0], t_str[1], t_str[2], _ptr(final), _ptr(first),
            _ptr(done), L.DT_PROP_STOP_ON_COLLISION if stop_on_collision else 0, self._stream()))
        res = dict(traj=traj, final=final, first_coll=first, done_step=done)
        if packed_out and not soa and want_traj:
            res["blob"], res["pitch"] = blob, pitch
        return res

    # -- conditioning -----------------------------------------------------------------------
    def build_cond_car(self, states, prev_action, goal, meta, map_size=20.0):
        st = self._f32(states)
        assert st.dim() == 2 and st.shape[1] == 6 and st.is_contiguous()
        B = st.shape[0]
        prev = None if prev_action is None else self._f32(prev_action).contiguous()
        g = self._f32(goal).contiguous()
        gstride = 0 if g.dim() == 1 else 2
        norm = self._norm_vec(meta)
        out = torch.empty((B, 7), dtype=torch.float32, device=self.device)
        self._check(self.lib.dt_build_cond_car(self.h, _ptr(st), 6, 1, _ptr(prev), _ptr(g), gstride, B,
                                               norm.ctypes.data_as(C.c_void_p), float(map_size), _ptr(out),
                                               self._stream()))
        return out

    def build_cond_ant(self, obs_seq, prev_action, goal, meta, obs_history=3, map_size=16.0):
        o = self._f32(obs_seq).contiguous()
        if o.dim() == 2:
            o = o[:, None, :]
        B, h, d = o.shape
        assert d == 29
        prev = None if prev_action is None else self._f32(prev_action).contiguous()
        g = self._f32(goal).contiguous()
        gstride = 0 if g.dim() == 1 else 2
        norm = self._norm_vec(meta)
        out = torch.empty((B, obs_history * 29 + 10), dtype=torch.float32, device=self.device)
        self._check(self.lib.dt_build_cond_ant(self.h, _ptr(o), h, obs_history, _ptr(prev), _ptr(g), gstride, B,
                                               norm.ctypes.data_as(C.c_void_p), float(map_size), _ptr(out),
                                               self._stream()))
        return out

    # -- reductions -------------------------------------------------------------------------
    def nearest(self, node_x, node_y, queries):
        nx, ny = self._f32(node_x).contiguous(), self._f32(node_y).contiguous()
        q = self._f32(queries)
        assert q.dim() == 2 and q.stride(1) == 1
        out = torch.empty(q.shape[0], dtype=torch.int32, device=self.device)
        self._check(self.lib.dt_nearest(self.h, _ptr(nx), _ptr(ny), nx.shape[0], _ptr(q[:, 0]), _ptr(q[:, 1]),
                                        q.stride(0), q.shape[0], _ptr(out), self._stream()))
        return out

    def nearest_k(self, node_x, node_y, queries, k):
        """KDTree.query(queries, k) indices: (Q, k) int32, ascending distance, n = missing neighbour."""
        nx, ny = self._f32(node_x).contiguous(), self._f32(node_y).contiguous()
        q = self._f32(queries)
        assert q.dim() == 2 and q.stride(1) == 1
        out = torch.empty((q.shape[0], int(k)), dtype=torch.int32, device=self.device)
        self._check(self.lib.dt_nearest_k(self.h, _ptr(nx), _ptr(ny), nx.shape[0], _ptr(q[:, 0]), _ptr(q[:, 1]),
                                          q.stride(0), q.shape[0], int(k), _ptr(out), self._stream()))
        return out

    def goal_cost_argmin(self, node_x, node_y, goal_xy, ahead=None):
        nx, ny = self._f32(node_x).contiguous(), self._f32(node_y).contiguous()
        if ahead is not None:
            ahead = ahead.to(self.device, torch.uint8).contiguous()
        out = torch.empty(1, dtype=torch.int32, device=self.device)
        self._check(self.lib.dt_goal_cost_argmin(self.h, _ptr(nx), _ptr(ny), nx.shape[0], float(goal_xy[0]),
                                                 float(goal_xy[1]), _ptr(ahead), _ptr(out), self._stream()))
        return out

    def mppi_reduce(self, cost, noise, lam, u, want_weights=False):
        c = self._f32(cost).contiguous()
        n = self._f32(noise).contiguous()
        K = c.shape[0]
        TA = n.numel() // K
        u = self._f32(u).contiguous().clone()
        amin = torch.empty(1, dtype=torch.int32, device=self.device)
        w = torch.empty(K, dtype=torch.float32, device=self.device) if want_weights else None
        self._check(self.lib.dt_mppi_reduce(self.h, _ptr(c), _ptr(n), K, TA, float(lam), _ptr(u), _ptr(amin), _ptr(w),
                                            self._stream()))
        return u, amin, w

    def mppi_rollout_cost(self, state, u, noise, ref_xy, lookahead, goal_xy, collision_cost, effort_cost):
        """-> (cost (K,), target (2,)) for K rollouts of T steps from one state (fused rollout + cost kernel)."""
        st = self._f32(state).contiguous()
        uu = self._f32(u).contiguous()
        nz = self._f32(noise).contiguous()
        ref = self._f32(ref_xy).contiguous()
        K, T = nz.shape[0], nz.shape[1]
        cost = torch.empty(K, dtype=torch.float32, device=self.device)
        target = torch.empty(2, dtype=torch.float32, device=self.device)
        self._check(self.lib.dt_mppi_rollout_cost(self.h, _ptr(st), _ptr(uu), _ptr(nz), K, T, _ptr(ref), ref.shape[0],
                                                  int(lookahead), float(goal_xy[0]), float(goal_xy[1]),
                                                  float(collision_cost), float(effort_cost), _ptr(cost), _ptr(target),
                                                  self._stream()))
        return cost, target

    def mppi_shift(self, u):
        """In place: u <- u shifted left by one step (last step repeated); returns the action that was u[0]."""
        assert u.is_contiguous() and u.dtype == torch.float32 and u.device == self.device
        act = torch.empty(u.shape[1], dtype=torch.float32, device=self.device)
        self._check(self.lib.dt_mppi_shift(self.h, _ptr(u), u.shape[0], u.shape[1], _ptr(act), self._stream()))
        return act

    # -- probability-map state sampler (run_type >= 2) -----------------------------------------
    def _f64(self, t):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t, dtype=np.float64))
        if t.device != self.device or t.dtype != torch.float64:
            t = t.to(self.device, torch.float64)
        return t.contiguous()

    def edt_prior(self):
        """CarEnv.prior of the staged map: (rows, cols) float64 device tensor."""
        out = torch.empty(tuple(self.map_shape), dtype=torch.float64, device=self.device)
        self._check(self.lib.dt_edt_prior(self.h, _ptr(out), self._stream()))
        return out

    def prob_map(self, prior, robot_xy, goal_xy, beta=0.8):
        """-> (prob_map, gaussian_pdf), both the prior's shape, float64 device tensors."""
        pr = self._f64(prior)
        assert pr.dim() == 2
        out = torch.empty_like(pr)
        gauss = torch.empty_like(pr)
        rc = self.lib.dt_prob_map(self.h, pr.shape[0], pr.shape[1], _ptr(pr), float(robot_xy[0]), float(robot_xy[1]),
                                  float(goal_xy[0]), float(goal_xy[1]), float(beta), _ptr(out), _ptr(gauss), self._stream())
        if rc == L.DT_E_INDEX:
            raise IndexError(self.lib.dt_last_error(self.h).decode())
        self._check(rc)
        return out, gauss

    def log_blend(self, prior, gauss, beta=0.8, obstacle_mask=None, eps=1e-12):
        pr, ga = self._f64(prior), self._f64(gauss)
        assert pr.shape == ga.shape
        mask = None if obstacle_mask is None else torch.as_tensor(np.asarray(obstacle_mask, dtype=bool)).to(self.device, torch.uint8).contiguous()
        out = torch.empty_like(pr)
        self._check(self.lib.dt_log_blend(self.h, _ptr(pr), _ptr(ga), _ptr(mask), pr.numel(), float(beta), float(eps),
                                          _ptr(out), self._stream()))
        return out

    def sample_cells(self, prob, u):
        """np.random.choice(prob.size, p=prob.ravel()) for the uniform draws u (B,) -> (B,) int32 flat indices."""
        p = self._f64(prob).reshape(-1)
        uu = self._f64(u).reshape(-1)
        out = torch.empty(uu.shape[0], dtype=torch.int32, device=self.device)
        self._check(self.lib.dt_sample_cells(self.h, _ptr(p), p.shape[0], _ptr(uu), uu.shape[0], _ptr(out), self._stream()))
        return out

    # -- denoiser ---------------------------------------------------------------------------
    def load_denoiser(self, state_dict, action_dim, horizon, cond_dim, emb_dim, map_size, down_dims, max_batch):
        self._loaded_key = None   # DiffusionSampler._context's cache key: whoever packs weights directly invalidates it
        keep = []
        descs = (L.TensorDesc * len(state_dict))()
        for i, (k, v) in enumerate(state_dict.items()):
            a = v.detach().to("cpu", torch.float32).contiguous() if isinstance(v, torch.Tensor) else \
                torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
            keep.append(a)
            name = k.encode()
            keep.append(name)
            descs[i].name = name
            descs[i].data = a.data_ptr()
            descs[i].ndim = a.dim()
            for j in range(4):
                descs[i].shape[j] = a.shape[j] if j < a.dim() else 1
        cfg = L.ModelCfg(action_dim, horizon, cond_dim, emb_dim, map_size, (C.c_int * 3)(*down_dims), max_batch)
        self._check(self.lib.dt_load_denoiser(self.h, descs, len(state_dict), C.byref(cfg), self._stream()))
        self.model_cfg = dict(action_dim=action_dim, horizon=horizon, cond_dim=cond_dim, emb_dim=emb_dim,
                              map_size=map_size, down_dims=tuple(down_dims), max_batch=max_batch)

    def fm_sample(self, noise, cond, local_map_bf16, K, act_mean=None, act_std=None, exp_scale=4.0):
        cfg = self.model_cfg
        if cfg is None:
            raise L.DitreeError(L.DT_E_NOMODEL, "load_denoiser has not been called")
        nz = self._f32(noise).contiguous()
        cd = self._f32(cond).contiguous()
        lm = local_map_bf16.contiguous()
        assert lm.dtype == torch.bfloat16 and lm.device == self.device
        B = nz.shape[0]
        out = torch.empty_like(nz)
        norm = None
        if act_mean is not None:
            norm = np.concatenate([np.asarray(act_mean, np.float64), np.asarray(act_std, np.float64)])
        self._check(self.lib.dt_fm_sample(self.h, _ptr(nz), _ptr(cd), _ptr(lm), B, int(K), float(exp_scale),
                                          norm.ctypes.data_as(C.c_void_p) if norm is not None else C.c_void_p(0),
                                          _ptr(out), self._stream()))
        return out

    def encode_map(self, local_map_bf16):
        lm = local_map_bf16.contiguous()
        B = lm.shape[0]
        out = torch.empty((B, self.model_cfg["emb_dim"]), dtype=torch.float32, device=self.device)
        self._check(self.lib.dt_encode_map(self.h, _ptr(lm), B, _ptr(out), self._stream()))
        return out

    def unet_forward(self, sample, emb, cond, timestep):
        s = self._f32(sample).contiguous()
        e = self._f32(emb).contiguous()
        c = self._f32(cond).contiguous()
        out = torch.empty_like(s)
        self._check(self.lib.dt_unet_forward(self.h, _ptr(s), _ptr(e), _ptr(c), s.shape[0], float(timestep), _ptr(out),
                                             self._stream()))
        return out

    def conv2d_gn(self, x, w, gamma, beta, k, stride, pad, resid=None, relu=True):
        """Test hook: x (B,H,W,Cin) bf16, w (N, k*k*Cin) bf16 tap-major -> (B,OH,OW,N) bf16 (conv -> GroupNorm(N/16) -> +resid -> ReLU)."""
        assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.is_contiguous() and w.is_contiguous()
        B, H, W, Cin = x.shape
        N = w.shape[0]
        OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        out = torch.empty((B, OH, OW, N), dtype=torch.bfloat16, device=self.device)
        ga, be = self._f32(gamma).contiguous(), self._f32(beta).contiguous()
        self._check(self.lib.dt_conv2d_gn_bf16(self.h, _ptr(x), B, H, W, Cin, _ptr(w), N, int(k), int(stride), int(pad), _ptr(ga),
                                               _ptr(be), _ptr(resid.contiguous() if resid is not None else None), int(bool(relu)),
                                               _ptr(out), self._stream()))
        return out

    def gemm_bf16(self, a, w):
        assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.is_contiguous() and w.is_contiguous()
        M, K = a.shape
        N = w.shape[0]
        out = torch.empty((M, N), dtype=torch.float32, device=self.device)
        self._check(self.lib.dt_gemm_bf16(self.h, _ptr(a), _ptr(w), M, N, K, _ptr(out), self._stream()))
        return out


_default = {}
_cuda_ok = False


def get_context(device=None):
    """Process-wide context per device (the reference's modules are process-wide singletons too).
    ``device=None`` means torch's current CUDA device, so one-process-per-GPU launches (torchrun sets
    the device from LOCAL_RANK) get the context of their own GPU."""
    if device is None:
        global _cuda_ok
        if not _cuda_ok:   # checked until it succeeds once (torch.cuda.is_available() costs 5 us a call: NVML + getenv)
            if not torch.cuda.is_available():
                raise RuntimeError("ditreeonlineplanner_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
            _cuda_ok = True
        device = torch.cuda.current_device()
    idx = device if isinstance(device, int) else (torch.device(device).index or 0)
    if idx not in _default:
        _default[idx] = Context(idx)
    return _default[idx]
