"""Import-only stub: names the reference's se3 helpers import but the pinned paths never call."""
def r2q(*a, **k):
    raise NotImplementedError("spatialmath stub")
from . import transforms3d  # noqa
