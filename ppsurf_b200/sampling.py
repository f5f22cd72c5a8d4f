"""Quantised support sampling on the device (``sampling_quantized``,
source/poco_data_loader.py:59-134): one representative per voxel of a randomly rotated grid, voxel edge
``||bbox||_2 / sqrt(n_support)``, halved until enough points are picked, random trim of the last round.

The reference is non-deterministic by construction (random rotations, randperm, scatter winner), so parity is
distributional.  This version keeps the points on the device and uses torch's sort/unique for the voxel bucketing
(index plumbing; the kNN that consumes the supports is the hand-written kernel).  Rotation angles and the final
trim come from a host ``numpy`` generator so a seed reproduces a sampling.
"""
import math

import numpy as np
import torch


def _rotation(gen: np.random.Generator) -> np.ndarray:
    mats = []
    for axis in range(3):
        deg = gen.uniform(-180.0, 180.0)
        c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
        m = np.eye(3)
        a, b = [(1, 2), (0, 2), (0, 1)][axis]
        m[a, a], m[a, b], m[b, a], m[b, b] = c, s, -s, c
        mats.append(m)
    return mats[2] @ mats[1] @ mats[0]  # rot_z(rot_y(rot_x(.)))  (poco_data_loader.py:103)


def sampling_quantized(pts: torch.Tensor, n_support: int, gen: np.random.Generator) -> torch.Tensor:
    """``pts [N,3]`` (device) -> indices ``[n_support]`` int64 (device)"""
    n = pts.shape[0]
    dev = pts.device
    if n_support >= n:
        return torch.arange(n, device=dev)
    ext = pts.max(dim=0).values - pts.min(dim=0).values
    vox = float(ext.norm(2)) / math.sqrt(n_support)
    cur = pts
    ids = torch.arange(n, device=dev)
    picked, count = [], 0
    while True:
        rot = torch.from_numpy(_rotation(gen)).to(dev, pts.dtype)
        pr = cur @ rot.T
        cell = torch.floor((pr - pr.min(dim=0).values) / vox).to(torch.int64)
        dims = cell.max(dim=0).values + 1
        key = (cell[:, 2] * dims[1] + cell[:, 1]) * dims[0] + cell[:, 0]
        order = torch.argsort(key, stable=True)
        sk = key[order]
        first = torch.ones_like(sk, dtype=torch.bool)
        first[1:] = sk[1:] != sk[:-1]
        rep = order[first]  # one representative per occupied voxel
        if count + rep.shape[0] < n_support:
            picked.append(ids[rep])
            count += rep.shape[0]
            keep = torch.ones(cur.shape[0], dtype=torch.bool, device=dev)
            keep[rep] = False
            cur, ids = cur[keep], ids[keep]
            vox = vox / 2
        else:
            sel = torch.from_numpy(gen.permutation(rep.shape[0])[:n_support - count]).to(dev)
            picked.append(ids[rep[sel]])
            break
    return torch.cat(picked)
