"""Training / test side data preparation on the device.

The reference prepares every item in DataLoader worker processes on the CPU (``PPSurfDataset.__getitem__``,
source/ppsurf_data_loader.py:61-81): a kd-tree over the raw cloud, the k = ``num_pts_local`` nearest raw points of every query
(``get_local_subsamples`` 83-89), their patch normalisation (91-123), then ``get_data_poco`` (source/poco_data_loader.py:243-270):
the occupancy labels, four quantised support samplings, 13 kNN index tensors (``get_fkaconv_ids`` 137-209) and the k = 64
projection neighbours of the queries (``get_proj_ids`` 212-240).  With a fast forward / backward those workers starve the GPUs;
``prepare_batch`` does the same work for a whole collated batch with the kernels of the predict path and returns the dict
``PPSurfNetwork.forward`` (and the reference's ``network.forward``) consumes.  File I/O and the random augmentation rotation stay on
the host (they are a handful of flops per point).
"""
import typing

import torch

from . import ops


def occupancy_labels(imp_surf_dist_ms: torch.Tensor) -> torch.Tensor:
    """source/poco_data_loader.py:251-255: 1 where the signed distance is positive, else 0 (int64)"""
    occ = torch.zeros_like(imp_surf_dist_ms, dtype=torch.int64)
    occ[torch.sign(imp_surf_dist_ms) > 0.0] = 1
    return occ


def local_patches(pts_raw: torch.Tensor, pts_query: torch.Tensor, num_pts_local: int) -> torch.Tensor:
    """``pts_raw [N,3]``, ``pts_query [Q,3]`` (device, fp32) -> ``pts_local_ps [Q,P,3]``: the P nearest raw points of every query,
    centred on the query and divided by the distance of the farthest one (ppsurf_data_loader.py:83-123)"""
    idx, d2 = ops.knn(pts_raw, pts_query, num_pts_local, return_dist=True)
    return ops.patch_normalize(pts_raw, pts_query, idx, d2, idx.shape[1])


def prepare_batch(network, batch: typing.Dict[str, torch.Tensor], k: typing.Optional[int] = None) -> typing.Dict[str, torch.Tensor]:
    """``batch``: ``pts_ms [B,N,3]``, ``pts_query_ms [B,Q,3]``, optional ``imp_surf_dist_ms [B,Q]`` and ``pts_raw_ms [B,Nraw,3]``
    (defaults to ``pts_ms``), any device.  Adds, in the reference's layouts: ``pts [B,3,N]``, ``pts_query [B,3,Q]``, ``occ [B,Q]``,
    ``support1-4``, ``ids00 .. ids10`` (int64), ``proj_ids [B,Q,k]`` (int64), ``pts_local_ps [B,Q,P,3]``."""
    dev = next(network.parameters()).device
    k = network.k if k is None else k
    out = dict(batch)
    pts_ms = batch['pts_ms'].to(dev, torch.float32).contiguous()
    qry_ms = batch['pts_query_ms'].to(dev, torch.float32).contiguous()
    raw_ms = batch['pts_raw_ms'].to(dev, torch.float32).contiguous() if 'pts_raw_ms' in batch else pts_ms
    out['pts'] = pts_ms.transpose(1, 2).contiguous()
    out['pts_query'] = qry_ms.transpose(1, 2).contiguous()
    if 'imp_surf_dist_ms' in batch:
        out['occ'] = occupancy_labels(batch['imp_surf_dist_ms'].to(dev))
    else:
        out['occ'] = torch.zeros(qry_ms.shape[:2], dtype=torch.int64, device=dev)
    out.update(network.spatial_ids(out['pts']))
    proj, loc = [], []
    for b in range(pts_ms.shape[0]):
        proj.append(ops.knn(pts_ms[b], qry_ms[b], k).long())
        loc.append(local_patches(raw_ms[b], qry_ms[b], network.num_pts_local))
    out['proj_ids'] = torch.stack(proj)
    out['pts_local_ps'] = torch.stack(loc)
    return out
