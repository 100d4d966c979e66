"""Data tables of the reference (maze grids maps/mazes/*.csv, scenario rows experiments/*.csv,
normaliser statistics metadata/*.pt) repacked by tools/gen_golden.py."""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def load_maze(name):
    """(rows, cols) float64 grid, 1 = wall (what np.loadtxt gives the reference, run_scenarios.py:213)."""
    if "mazes" not in _cache:
        z = np.load(os.path.join(_HERE, "mazes.npz"))
        _cache["mazes"] = {k: z[k] for k in z.files}
    if name not in _cache["mazes"]:
        raise FileNotFoundError(f"maze {name!r} not found")
    return _cache["mazes"][name].astype(np.float64)


def maze_names():
    load_maze("boxes")
    return sorted(_cache["mazes"])


def load_metadata(env_id):
    """Normaliser statistics dict (policies/fm_policy.py:28-30); raises FileNotFoundError like the reference."""
    path = os.path.join(_HERE, f"metadata_{env_id}.npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"Metadata not found at {path}")
    z = np.load(path)
    return {k: z[k] for k in z.files}


def load_scenarios(kind="test_scenarios_car"):
    with open(os.path.join(_HERE, "scenarios.json")) as f:
        return json.load(f)[kind]
