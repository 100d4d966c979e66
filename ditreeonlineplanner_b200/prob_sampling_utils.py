"""Mirror of the reference's ``prob_sampling_utils.py`` functions that ``CarEnv`` uses for the
probability-map state sampler (run_type >= 2), computed by the device kernels of csrc/probmap.cu:

    gaussian_map        prob_sampling_utils.py:48-93    -> (pdf, mean, Sigma)
    combine_log_blend   prob_sampling_utils.py:150-172
    sample_from_pdf     prob_sampling_utils.py:176-180

``gaussian_map`` + ``combine_log_blend`` run as ONE kernel (``dt_prob_map``) when called through
``CarEnv``; the stand-alone functions below exist for API parity and return NumPy arrays.
"""
from __future__ import annotations

import numpy as np
import torch

from .runtime import get_context


def _mean_sigma(robot, goal):
    rx, ry = float(robot[0]), float(robot[1])
    gx, gy = float(goal[0]), float(goal[1])
    d = np.hypot(gx - rx, gy - ry) + 1e-6
    u = np.array([gx - rx, gy - ry]) / d if d > 1e-6 else np.array([1.0, 0.0])
    v = np.array([-u[1], u[0]])
    w = 1.0 - np.exp(-d / 15)
    mean = (1 - w) * np.array([gx, gy]) + w * np.array([(rx + gx) / 2, (ry + gy) / 2])
    s_long = 1.0 + 0.7 * np.log1p(d)
    rot = np.stack([u, v], axis=1)
    return mean, rot @ np.diag([s_long ** 2, (0.7 * s_long) ** 2]) @ rot.T


def blended_prob_map(prior, robot, goal, beta=0.8):
    """(combine_log_blend(prior, gaussian_map(robot, goal, prior.shape)), gaussian pdf) as float64 ndarrays."""
    ctx = get_context()
    prob, gauss = ctx.prob_map(np.asarray(prior, dtype=np.float64), robot, goal, beta)
    return prob.cpu().numpy(), gauss.cpu().numpy()


def gaussian_map(robot, goal, size=(20, 20)):
    uniform = np.full(size, 1.0 / (size[0] * size[1]))
    _, pdf = blended_prob_map(uniform, robot, goal)
    mean, sigma = _mean_sigma(robot, goal)
    return pdf, mean, sigma


def combine_log_blend(prior, gauss, beta=0.8, obstacle_mask=None, eps=1e-12):
    """Stand-alone blend of caller-supplied arrays (``dt_log_blend``), the reference's fallbacks included."""
    prior = np.asarray(prior, dtype=np.float64)
    gauss = np.asarray(gauss, dtype=np.float64)
    assert prior.shape == gauss.shape
    return get_context().log_blend(prior, gauss, beta, obstacle_mask, eps).cpu().numpy()


def sample_from_pdf(pdf, n_samples=30):
    """-> (xs, ys): n_samples cells drawn with np.random.choice's algorithm, the search on the device."""
    pdf = np.asarray(pdf, dtype=np.float64)
    u = np.random.random_sample(n_samples)
    idx = get_context().sample_cells(pdf, u).cpu().numpy()
    ys, xs = np.unravel_index(idx, pdf.shape)
    return xs, ys
