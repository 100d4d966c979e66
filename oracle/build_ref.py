#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE, not product code: stages the UNMODIFIED reference for the reference arm.

The reference (JJKK1313/DiTreeOnlinePlanner, mounted read-only at /root/reference in the build container) is a
directory of Python scripts; it cannot be pip-installed and does not exist on the GPU box.  This recipe byte-compiles
exactly the modules of the tree-expansion path (planners/base_planner.py:257-320 `propagate_action_sequence_env`,
policies/fm_policy.py:53-212 `DiffusionSampler.forward`, car_env.py `CarEnv.step`, common/map_utils.py
`create_local_map` / `is_colliding_car`, local_map_encoder.py + model/diffusion/* the network) and whatever they import
from the reference tree into ``oracle/_ref/`` as marshalled code objects (``<module path>.code``, loaded by the
finder in oracle/ref_arm.py; ``.pyc`` files do not survive the snapshot to the GPU box) -- compiled outputs only, like
a C reference's ``.so``; no reference SOURCE is copied into the repository, and ``oracle/_ref/`` is git-ignored (it travels to the GPU
box with the gpurun snapshot, like the built libditree.so).  ``oracle/_ref/metadata/carmaze.pt`` is re-created from the
normaliser statistics the package already ships as a data fixture (ditreeonlineplanner_b200/data/metadata_carmaze.npz).

    python oracle/build_ref.py            # no-op (exit 0, message) when /root/reference is absent

`bench.py --impl reference` and `bench.py`'s cpu_baseline leg load it through oracle/ref_arm.py; when ``oracle/_ref``
is missing they fall back to the oracle port and say so (`cpu_baseline.kind = "port"`).
"""
from __future__ import annotations

import json
import os
import marshal
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("DITREE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
STUBS = os.path.join(REPO, "tools", "ref_stubs")
ROOTS = ["car_env", "common.map_utils", "common.fm_utils", "local_map_encoder", "planners.base_planner",
         "planners.RRT", "policies.fm_policy", "lidar_sim.lidar_2d_sim", "prob_sampling_utils"]

_PROBE = r"""
import json, os, sys
sys.dont_write_bytecode = True
sys.path[:0] = [{stubs!r}, {ref!r}]
os.chdir({ref!r})
import importlib
for m in {roots!r}:
    importlib.import_module(m)
ref = os.path.realpath({ref!r}) + os.sep
files = sorted({{os.path.realpath(m.__file__) for m in list(sys.modules.values())
                if getattr(m, "__file__", None) and os.path.isabs(m.__file__) and m.__file__.endswith(".py")
                and os.path.isfile(m.__file__) and os.path.realpath(m.__file__).startswith(ref)}})
print("FILES=" + json.dumps(files))
"""


def build(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f"oracle/build_ref: {REF} not present (GPU box): keeping the prebuilt oracle/_ref as is")
        return os.path.isdir(OUT)
    code = _PROBE.format(stubs=STUBS, ref=REF, roots=ROOTS)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    line = [l for l in out.stdout.splitlines() if l.startswith("FILES=")]
    if out.returncode != 0 or not line:
        raise RuntimeError("oracle/build_ref: importing the reference failed:\n" + out.stdout[-2000:] + out.stderr[-4000:])
    files = json.loads(line[0][6:])
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    ref_root = os.path.realpath(REF)
    manifest = []
    for src in files:
        rel = os.path.relpath(src, ref_root)
        dst = os.path.join(OUT, rel[:-3] + ".code")  # x.py -> x.code: marshal.dumps(compile(source))
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(src, "rb") as f:
            code = compile(f.read(), "reference:" + rel, "exec", dont_inherit=True, optimize=0)
        with open(dst, "wb") as f:
            f.write(marshal.dumps(code))
        manifest.append(rel)
    # the normaliser statistics the sampler loads relative to the CWD (policies/fm_policy.py:28-30)
    import numpy as np
    import torch
    os.makedirs(os.path.join(OUT, "metadata"), exist_ok=True)
    for envname in ("carmaze", "antmaze"):
        z = np.load(os.path.join(REPO, "ditreeonlineplanner_b200", "data", f"metadata_{envname}.npz"))
        torch.save({k: np.asarray(z[k], dtype=np.float64) for k in z.files}, os.path.join(OUT, "metadata", f"{envname}.pt"))
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"reference": REF, "python": sys.version.split()[0], "modules": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/build_ref: staged {len(manifest)} compiled reference modules into {OUT}")
    return True


if __name__ == "__main__":
    build()
