"""TEST / BENCH INFRASTRUCTURE, not product code: runs the UNMODIFIED reference staged in ``oracle/_ref`` (see
oracle/build_ref.py) on the benchmark workload through the reference's own public API and stock code path:

    create_local_map                          common/map_utils.py:391-459
    DiffusionSampler.forward                  policies/fm_policy.py:53-212   (torch CPU, all host threads)
    BasePlanner.propagate_action_sequence_env planners/base_planner.py:257-320
        -> CarEnv.set_state / step            car_env.py:240-282,306-312,341-396
        -> is_colliding_car                   common/map_utils.py:103-115

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product never does.
"""
from __future__ import annotations

import contextlib
import importlib.abc
import importlib.util
import marshal
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")
STUBS = os.path.join(REPO, "tools", "ref_stubs")
_PROTECTED = ("car_env", "common", "local_map_encoder", "planners", "policies", "lidar_sim", "prob_sampling_utils",
              "model", "plot_logger", "casadi", "gymnasium", "matplotlib", "minari", "diffusers", "spatialmath",
              "termcolor")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "MANIFEST.json"))


class _CodeFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Imports `a.b` from oracle/_ref/a/b.code (a marshalled code object of the unmodified reference module, written
    by oracle/build_ref.py with this same interpreter version); directories are namespace-style packages."""

    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        base = os.path.join(self.root, *fullname.split("."))
        if os.path.isfile(base + ".code"):
            return importlib.util.spec_from_loader(fullname, self, origin=base + ".code")
        if os.path.isfile(os.path.join(base, "__init__.code")):
            return importlib.util.spec_from_loader(fullname, self, origin=os.path.join(base, "__init__.code"), is_package=True)
        if os.path.isdir(base) and fullname.split(".")[0] in _REF_TOP:
            spec = importlib.util.spec_from_loader(fullname, self, origin=base, is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        origin = module.__spec__.origin
        if os.path.isdir(origin):
            return
        module.__file__ = origin
        with open(origin, "rb") as f:
            code = marshal.loads(f.read())
        exec(code, module.__dict__)


_REF_TOP = ("common", "planners", "policies", "lidar_sim", "model")


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


class Reference:
    """The staged reference's modules, imported once under their own top-level names.  They share names with nothing in
    this repository's import space (the package lives under ditreeonlineplanner_b200.*)."""

    def __init__(self):
        if not available():
            raise FileNotFoundError("oracle/_ref is not staged: run `python oracle/build_ref.py` in the build container")
        clash = [m for m in _PROTECTED if m in sys.modules and m != "matplotlib"
                 and not str(getattr(sys.modules[m], "__file__", None) or REF_DIR).startswith((REF_DIR, STUBS))]
        if clash:
            raise RuntimeError(f"oracle/ref_arm: modules {clash} are already imported from elsewhere")
        sys.path.insert(0, STUBS)
        sys.meta_path.insert(0, _CodeFinder(REF_DIR))
        with _cwd(REF_DIR):
            import car_env
            import common.map_utils as map_utils
            from local_map_encoder import ConditionalUnet1DWithLocalMap
            from planners.RRT import RRT_Planner
            from policies.fm_policy import DiffusionSampler
        self.car_env, self.map_utils = car_env, map_utils
        self.Net, self.RRT_Planner, self.DiffusionSampler = ConditionalUnet1DWithLocalMap, RRT_Planner, DiffusionSampler

    def expansion(self, grid, state_dict, dims, K, S, goal_xy):
        """-> run(states (B,6) f64, prev_actions (B,2) f64, seed) -> dict like oracle.rollout_car's, computed by the
        reference: one batched sampler call, then the reference's own per-candidate propagate loop."""
        import torch
        net = self.Net(input_dim=2, encoder_name="resnet", embedding_dim=400, additional_global_cond_dim=7,
                       local_map_size=20, down_dims=list(dims))
        net.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()}, strict=True)
        net.eval()
        with _cwd(REF_DIR):  # metadata/carmaze.pt is resolved relative to the CWD (fm_policy.py:28)
            smp = self.DiffusionSampler(net, None, "carmaze", policy="flow_matching", pred_horizon=64, action_dim=2,
                                        obs_history=1, action_history=1, goal_conditioned=True, num_diffusion_iters=K,
                                        local_map_size=20)
        smp.device = "cpu"
        maze = np.asarray(grid, dtype=np.float64)
        env = self.car_env.CarEnv(maze_map=maze.copy(), collision_checking=False)
        R, C = maze.shape
        start = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
        goal = np.array([goal_xy[0], goal_xy[1], 0.0, 0.0, 0.0, 0.0])
        planner = self.RRT_Planner(start, goal, env_id="carmaze", environment=env, sampler=smp, action_horizon=S,
                                   local_map_size=20, local_map_scale=0.2, global_map_scale=1.0, time_budget=1)
        # the reference picks 'cuda' whenever a GPU is visible (base_planner.py, fm_policy.py:26) and moves the sampler
        # there; this arm is the reference's CPU path, so pin everything back to the host
        planner.device = "cpu"
        smp.to("cpu")
        smp.device = "cpu"
        mu = self.map_utils

        def run(states, prev_actions, seed=None):
            st = np.asarray(states, dtype=np.float64)
            B = st.shape[0]
            lm = mu.create_local_map(planner.maze, st[:, 0], st[:, 1], st[:, 2], 20, 0.2, 1.0, (C / 2, R / 2))
            if seed is not None:
                torch.manual_seed(seed)
            with torch.no_grad():
                act = smp(st[:, None, :], prev_actions=np.asarray(prev_actions, dtype=np.float64)[:, None, :],
                          goal=np.asarray(goal_xy, dtype=np.float64), local_map=lm)
            act = np.asarray(act)
            final = np.zeros((B, 6))
            first_coll = np.full(B, -1, np.int32)
            done_step = np.full(B, -1, np.int32)
            n_states = np.zeros(B, np.int32)
            for b in range(B):
                env.done = False
                env.terminated = False
                obs, done, a_seq, s_seq = planner.propagate_action_sequence_env(st[b].copy(), act[b, :S].copy())
                final[b] = obs
                if done is None:
                    first_coll[b] = len(a_seq)       # the colliding step's index (base_planner.py:309-311)
                elif done:
                    nz = np.flatnonzero(~(np.asarray(a_seq) == 0).all(axis=1))
                    done_step[b] = (nz[-1] if len(nz) else -1)
                n_states[b] = s_seq.shape[1]
            return dict(actions=act, final=final, first_coll=first_coll, done_step=done_step, n_states=n_states,
                        local_map=lm)
        return run
