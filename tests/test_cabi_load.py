"""CPU-side checks of the boundary: the C-ABI library is built, loads, and exports every symbol
include/ditree.h declares (no compute call is made without a GPU)."""
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ditreeonlineplanner_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from ditreeonlineplanner_b200 import _lib
    hdr = open(os.path.join(REPO, "include", "ditree.h")).read()
    declared = set(re.findall(r"\b(dt_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"dt_ctx", "dt_tensor_desc", "dt_model_cfg"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_null_ctx(lib):
    assert b"sm_100a" in lib.dt_version()
    assert lib.dt_last_error(None) == b"null context"
    assert lib.dt_launch_count(None) == 0


def test_no_cpu_fallback():
    """The product package never imports the oracle and refuses to run without CUDA."""
    import torch
    pkg = os.path.join(REPO, "ditreeonlineplanner_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
    if not torch.cuda.is_available():
        from ditreeonlineplanner_b200 import Context
        with pytest.raises(RuntimeError):
            Context(0)


def test_sass_is_sm100a():
    import subprocess
    from ditreeonlineplanner_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_sass_opcodes_are_blackwell_native():
    """The SASS of the built library, not the PTX source: every k_conv_gemm instantiation issues tcgen05.mma
    (UTCHMMA; .2CTA in the CTA-pair ones), reads its accumulator with tcgen05.ld (LDTM) and loads tiles with TMA
    (UTMALDG); the geometry kernels stage the occupancy grid with a bulk TMA copy (UBLKCP); nothing uses the
    legacy mma.sync path (HMMA).  profiles/r02_sass_opcodes.md is this histogram, committed."""
    import sys
    sys.path.insert(0, os.path.join(REPO, "tools"))
    import sass_histogram as sh
    per = sh.histogram()
    gemm = {k: c for k, c in per.items() if "k_conv_gemm" in k}
    assert len(gemm) >= 12, len(gemm)
    for k, c in gemm.items():
        assert c["UTCHMMA"] >= 4 and c["LDTM"] >= 2 and c["UTMALDG"] >= 2 and c["UTCBAR"] >= 2, (k, dict(c))
        assert c["STL"] == 0 and c["LDL"] == 0, (k, "register spills in the GEMM kernel")
        # programmatic dependent launch: griddepcontrol.launch_dependents / .wait (SASS PREEXIT / ACQBULK)
        assert c["PREEXIT"] >= 1 and c["ACQBULK"] >= 1, (k, dict(c))
    assert sum(1 for c in gemm.values() if c["UTCHMMA.2CTA"] >= 4) >= len(gemm) // 2
    assert all(c["HMMA"] == 0 for c in per.values())
    # every kernel of the denoiser chain that is launched with the programmatic attribute waits before its first access
    for name in ("k_splitk_epi", "k_prep_sample", "k_final_euler", "k_film_input", "k_time_mlp1", "k_time_mlp2", "k_film_time",
                 "k_im2col", "k_maxpool3s2", "k_avgpool", "k_copy_cols"):
        hit = [c for k, c in per.items() if name in k]
        assert hit and all(c["ACQBULK"] >= 1 for c in hit), name
    for name in ("k_propagate_rows", "k_collide_car4", "k_local_map", "k_lidar_scan"):
        hit = [c for k, c in per.items() if name in k]
        assert hit and all(c["UBLKCP"] >= 1 for c in hit), name
