"""Host-side ODE schedule helper with the reference's name and return convention
(``common/fm_utils.py:4-17``): ``get_timesteps(schedule, K, exp_scale) -> (t0, dt)``, two float32 tensors of
length K with ``dt`` summing to one and ``t0`` the left end of every step.  The device sampler computes the
same 'exp' schedule in fp32 itself (csrc/denoiser.cu, dt_fm_sample); this helper serves callers and tests."""
import torch

_WEIGHTS = {
    "linear": lambda u, scale: torch.ones_like(u),
    "cosine": lambda u, scale: 1 + torch.cos(torch.pi * u),
    "exp": lambda u, scale: torch.exp(-scale * u),
}


def get_timesteps(schedule: str, k_steps: int, exp_scale: float = 1.0):
    if schedule not in _WEIGHTS:
        raise ValueError(f"Invalid schedule: {schedule}")
    u = torch.linspace(0, 1, k_steps + 1)[:-1]          # left ends of K equal sub-intervals of [0, 1)
    w = _WEIGHTS[schedule](u, exp_scale)
    dt = w / k_steps if schedule == "linear" else w / w.sum()
    starts = torch.zeros(k_steps)
    starts[1:] = torch.cumsum(dt, dim=0)[:-1]
    return starts, dt
