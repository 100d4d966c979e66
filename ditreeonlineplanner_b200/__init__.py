"""ditreeonlineplanner_b200 -- B200 (sm_100a) implementation of DiTree's tree-expansion hot path
behind the reference's planner / policy API.  See DESIGN.md and INTEGRATION.md."""
from . import _lib  # noqa: F401
from .runtime import Context, get_context  # noqa: F401
from .data import load_maze, load_metadata, load_scenarios  # noqa: F401

__all__ = ["Context", "get_context", "load_maze", "load_metadata", "load_scenarios"]
