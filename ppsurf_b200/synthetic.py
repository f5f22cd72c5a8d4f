"""Synthetic workload of the benchmark: no checkpoint and no dataset can be downloaded, so the bench
and the smoke test use a seeded noisy-sphere cloud and seeded random weights of the PPSurf 50NN architecture.

Product-side twin of the generators in ``oracle/ppsurf_oracle.py`` (the product must not import the oracle);
``tests/test_host_logic.py`` asserts that both produce bit-identical tensors."""
import collections
import math

import numpy as np
import torch

DEFAULT_GAINS = {'encoder': 0.68, 'projection': 1.5, 'point_net': 1.21, 'mlp': 1.2}


def synthetic_cloud(n: int, seed: int = 42, radius: float = 0.4, noise: float = 0.005) -> np.ndarray:
    """noisy sphere ``[n,3]`` float32 inside [-0.5,0.5]^3 (reference-normalised extent)"""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return (radius * d + noise * rng.standard_normal((n, 3))).astype(np.float32)


def make_state_dict(network: torch.nn.Module, seed: int = 42, gains=None) -> 'collections.OrderedDict[str, torch.Tensor]':
    """seeded weights for every entry of ``network.state_dict()`` (names/shapes come from the module itself):
    U(+-gain*sqrt(3/fan_in)) matrices, non-trivial BatchNorm statistics and FKAConv scalars, so that latents are O(1)
    and logits O(1-10) on the synthetic cloud."""
    rng = np.random.default_rng(seed)
    gains = dict(DEFAULT_GAINS, **(gains or {}))
    spec = collections.OrderedDict((k, tuple(v.shape)) for k, v in network.state_dict().items())
    out = collections.OrderedDict()
    for name, shape in spec.items():
        owner, leaf = name.rsplit('.', 1)
        parent = owner.rsplit('.', 1)[0] if '.' in owner else ''
        is_norm = (owner + '.running_mean') in spec or (owner.endswith(('.bn1', '.bn2')) and (parent + '.alpha') in spec)
        if leaf == 'num_batches_tracked':
            v = np.array(7, dtype=np.int64)
        elif leaf in ('alpha', 'beta'):
            v = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif leaf == 'norm_radius':
            v = rng.uniform(0.05, 0.2, size=shape).astype(np.float32)
        elif leaf == 'running_mean':
            v = (0.1 * rng.standard_normal(size=shape)).astype(np.float32)
        elif leaf == 'running_var':
            v = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif is_norm and leaf == 'weight':
            v = rng.uniform(0.75, 1.25, size=shape).astype(np.float32)
        elif is_norm and leaf == 'bias':
            v = (0.1 * rng.standard_normal(size=shape)).astype(np.float32)
        elif leaf == 'weight':
            bound = gains[name.split('.', 1)[0]] * math.sqrt(3.0 / int(np.prod(shape[1:])))
            v = rng.uniform(-bound, bound, size=shape).astype(np.float32)
        elif leaf == 'bias':
            v = rng.uniform(-0.1, 0.1, size=shape).astype(np.float32)
        else:
            raise KeyError(name)
        out[name] = torch.from_numpy(np.asarray(v))
    return out
