"""Parameter inventory of the reference denoiser (``ConditionalUnet1DWithLocalMap`` with the
'resnet' encoder, local_map_encoder.py:78-122 + model/diffusion/conditional_unet1d.py:145-266):
names and shapes of its ``state_dict`` -- used to validate checkpoints before packing them for the
GPU and to draw random-init weights of the reference architecture for benchmarks (no checkpoint
is available offline)."""
from __future__ import annotations

import math

import numpy as np
import torch

UNET_DIMS = {"small": [64, 128, 256], "medium": [256, 512, 1024], "large": [512, 1024, 2048],
             "xlarge": [1024, 2048, 4096]}  # run_scenarios.py:92-97


def state_dict_shapes(input_dim=2, cond_dim=7, emb_dim=400, down_dims=(512, 1024, 2048), dsed=256):
    s = {}
    enc = "encoder.resnet18."
    s[enc + "conv1.weight"] = (64, 3, 7, 7)
    for nm in ("weight", "bias"):
        s[enc + "bn1." + nm] = (64,)
    prev = 64
    for li, c in enumerate((64, 128, 256, 512), start=1):
        for b in range(2):
            q = f"{enc}layer{li}.{b}."
            s[q + "conv1.weight"] = (c, prev if b == 0 else c, 3, 3)
            s[q + "bn1.weight"] = s[q + "bn1.bias"] = (c,)
            s[q + "conv2.weight"] = (c, c, 3, 3)
            s[q + "bn2.weight"] = s[q + "bn2.bias"] = (c,)
            if b == 0 and li > 1:
                s[q + "downsample.0.weight"] = (c, prev, 1, 1)
                s[q + "downsample.1.weight"] = s[q + "downsample.1.bias"] = (c,)
        prev = c
    s[enc + "fc.weight"] = (emb_dim, 512)
    s[enc + "fc.bias"] = (emb_dim,)
    gdim = dsed + emb_dim + cond_dim
    dims = [input_dim] + list(down_dims)

    def block(prefix, ci, co):
        for j, cc in enumerate((ci, co)):
            s[f"{prefix}blocks.{j}.block.0.weight"] = (co, cc, 3)
            s[f"{prefix}blocks.{j}.block.0.bias"] = (co,)
            s[f"{prefix}blocks.{j}.block.1.weight"] = s[f"{prefix}blocks.{j}.block.1.bias"] = (co,)
        s[f"{prefix}cond_encoder.1.weight"] = (2 * co, gdim)
        s[f"{prefix}cond_encoder.1.bias"] = (2 * co,)
        if ci != co:
            s[f"{prefix}residual_conv.weight"] = (co, ci, 1)
            s[f"{prefix}residual_conv.bias"] = (co,)

    for i in range(2):
        block(f"unet.mid_modules.{i}.", dims[-1], dims[-1])
    s["unet.diffusion_step_encoder.1.weight"] = (4 * dsed, dsed)
    s["unet.diffusion_step_encoder.1.bias"] = (4 * dsed,)
    s["unet.diffusion_step_encoder.3.weight"] = (dsed, 4 * dsed)
    s["unet.diffusion_step_encoder.3.bias"] = (dsed,)
    pairs = list(zip(dims[:-1], dims[1:]))
    for i, (ci, co) in enumerate(reversed(pairs[1:])):
        block(f"unet.up_modules.{i}.0.", 2 * co, ci)
        block(f"unet.up_modules.{i}.1.", ci, ci)
        s[f"unet.up_modules.{i}.2.conv.weight"] = (ci, ci, 4)
        s[f"unet.up_modules.{i}.2.conv.bias"] = (ci,)
    for i, (ci, co) in enumerate(pairs):
        block(f"unet.down_modules.{i}.0.", ci, co)
        block(f"unet.down_modules.{i}.1.", co, co)
        if i < len(pairs) - 1:
            s[f"unet.down_modules.{i}.2.conv.weight"] = (co, co, 3)
            s[f"unet.down_modules.{i}.2.conv.bias"] = (co,)
    c0 = dims[1]
    s["unet.final_conv.0.block.0.weight"] = (c0, c0, 3)
    s["unet.final_conv.0.block.0.bias"] = s["unet.final_conv.0.block.1.weight"] = s["unet.final_conv.0.block.1.bias"] = (c0,)
    s["unet.final_conv.1.weight"] = (input_dim, c0, 1)
    s["unet.final_conv.1.bias"] = (input_dim,)
    return s


def random_init(seed=0, **cfg):
    """Random weights with torch's default bounds (U(-1/sqrt(fan_in), 1/sqrt(fan_in))); norm layers
    get non-trivial affine parameters so every code path is exercised."""
    g = torch.Generator().manual_seed(seed)
    shapes = state_dict_shapes(**cfg)
    out = {}
    for name, shp in shapes.items():
        norm = ".bn" in name or "downsample.1." in name or ".block.1." in name
        if norm:
            t = torch.rand(shp, generator=g) + 0.5 if name.endswith("weight") else (torch.rand(shp, generator=g) - 0.5) * 0.4
        else:
            wshape = shapes[name.rsplit(".", 1)[0] + ".weight"]
            fan_in = wshape[1] * wshape[2] if ("up_modules" in name and ".2.conv." in name) else int(np.prod(wshape[1:]))
            t = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
        out[name] = t.float()
    return out


def validate_state_dict(sd, **cfg):
    """Raise KeyError / ValueError (like ``load_state_dict(strict=True)``) on a mismatching checkpoint."""
    want = state_dict_shapes(**cfg)
    missing = [k for k in want if k not in sd]
    if missing:
        raise KeyError(f"checkpoint is missing {len(missing)} tensors, e.g. {missing[:3]}")
    for k, shp in want.items():
        if tuple(sd[k].shape) != tuple(shp):
            raise ValueError(f"size mismatch for {k}: checkpoint {tuple(sd[k].shape)} vs model {tuple(shp)}")


def denoiser_flops(K, input_dim=2, cond_dim=7, emb_dim=400, down_dims=(512, 1024, 2048), horizon=64, map_size=20,
                   dsed=256):
    """Algorithmic FLOPs (2 x MAC of every conv / linear, as torch.utils.flop_counter counts the
    reference modules) of one K-step sample of ONE candidate with the encoder evaluated once:
    returns (encoder_flops, unet_flops_per_step)."""
    # encoder: resnet18 on map_size x map_size, 3 input channels
    def osz(h, k, s, p):
        return (h + 2 * p - k) // s + 1
    mac = 0
    h = osz(map_size, 7, 2, 3)
    mac += h * h * 64 * 3 * 49
    h = osz(h, 3, 2, 1)
    prev = 64
    for li, c in enumerate((64, 128, 256, 512)):
        for b in range(2):
            stride = 2 if (b == 0 and li > 0) else 1
            ho = osz(h, 3, stride, 1)
            mac += ho * ho * c * (prev if b == 0 else c) * 9
            mac += ho * ho * c * c * 9
            if b == 0 and li > 0:
                mac += ho * ho * c * prev
            h = ho
        prev = c
    mac += 512 * emb_dim
    enc = 2 * mac
    # U-Net
    gdim = dsed + emb_dim + cond_dim
    dims = [input_dim] + list(down_dims)
    T = [horizon >> i for i in range(len(down_dims))]
    m = dsed * 4 * dsed * 2  # time MLP

    def block(ci, co, t):
        x = t * co * ci * 3 + t * co * co * 3 + gdim * 2 * co
        if ci != co:
            x += t * co * ci
        return x
    for i in range(3):
        m += block(dims[i], dims[i + 1], T[i]) + block(dims[i + 1], dims[i + 1], T[i])
        if i < 2:
            m += T[i + 1] * dims[i + 1] * dims[i + 1] * 3
    m += 2 * block(dims[3], dims[3], T[2])
    for u, (ci, co, t) in enumerate(((2 * dims[3], dims[2], T[2]), (2 * dims[2], dims[1], T[1]))):
        m += block(ci, co, t) + block(co, co, t)
        m += t * co * co * 4  # ConvTranspose1d(4,2,1): every input step feeds 4 taps
    m += T[0] * dims[1] * dims[1] * 3 + T[0] * input_dim * dims[1]
    return float(enc), float(2 * m)
