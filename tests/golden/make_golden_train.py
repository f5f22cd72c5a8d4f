"""Golden vectors of ONE TRAINING STEP (BASELINE config 5) from the UNMODIFIED reference, run in the build container only:

    python tests/golden/make_golden_train.py

The reference ``PPSurfNetwork`` (imported read-only through ``ref_import``) is put in train mode, fed a seeded batch of two
clouds and back-propagated through ``cross_entropy`` exactly like ``PocoModel.training_step`` does (source/poco_model.py:75-125).
The only change to its configuration is ``Dropout.p = 0`` on the two MLP dropout modules: a random mask cannot be reproduced by
another implementation (dropout is tested separately).  Stored: the batch, loss, logits, the gradient of every parameter computed in FLOAT64 as
(L2 norm, sum, 48 sampled entries) next to the L2 error of the reference's own float32 gradient, and every buffer after the step (BatchNorm running statistics, FKAConv norm_radius).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_import  # noqa: E402
from oracle import ppsurf_oracle as O  # noqa: E402

SAMPLES = 48


def sample_positions(index: int, numel: int) -> np.ndarray:
    return np.random.default_rng(9000 + index).integers(0, numel, size=SAMPLES)


def make_batch(seed=200, b=2, n=1200, q=96, p=50, k=64):
    """two different noisy-sphere clouds with query points near and off the surface; ids from the oracle's get_fkaconv_ids"""
    rng = np.random.default_rng(seed)
    out = {}
    per = []
    for s in range(b):
        pts = O.synthetic_cloud(n, seed=seed + 1 + s)
        qry = np.concatenate([pts[rng.integers(0, n, q // 2)] + 0.03 * rng.standard_normal((q // 2, 3)),
                              rng.uniform(-0.5, 0.5, (q - q // 2, 3))]).astype(np.float32)
        d = {'pts': pts.T[None].copy()}
        d.update(O.get_fkaconv_ids(d['pts'], rng))
        d['pts_query'] = qry.T[None].copy()
        d['proj_ids'] = O.knn(pts, qry, k)[0][None].astype(np.int64)
        d['pts_local_ps'] = O.get_pts_local_ps(pts, qry, p)[None].astype(np.float32)
        d['occ'] = (np.linalg.norm(qry, axis=1) > 0.4).astype(np.int64)[None]  # outside the sphere = 1
        per.append(d)
    for key in per[0]:
        out[key] = np.concatenate([d[key] for d in per], axis=0)
    return out


def main():
    ref_import.import_reference()
    from source.ppsurf_model import PPSurfNetwork

    torch.set_float32_matmul_precision('highest')
    torch.manual_seed(0)
    p = O.make_state_dict(42)
    digest = O.state_dict_digest(p)
    with contextlib.redirect_stdout(io.StringIO()):
        net = PPSurfNetwork(3, 256, 2, 64, 50, 256)
    net.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in p.items()}, strict=True)
    net.train()
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    batch = make_batch()

    def step(dtype):
        data = {k: torch.from_numpy(v).to(dtype) if v.dtype == np.float32 else torch.from_numpy(v) for k, v in batch.items()}
        net.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in p.items()}, strict=True)  # fresh buffers
        net.to(dtype)
        net.zero_grad(set_to_none=True)
        pred = net.forward(dict(data))
        loss = torch.nn.functional.cross_entropy(input=pred, target=data['occ'], reduction='none').mean()
        loss.backward()
        grads = [(name, par.grad.detach().numpy().astype(np.float64).reshape(-1)) for name, par in net.named_parameters()]
        bufs = {name: buf.detach().numpy().astype(np.float32) for name, buf in net.named_buffers()
                if not name.endswith('num_batches_tracked')}
        return loss.item(), pred.detach().numpy(), grads, bufs

    loss64, logits64, grads64, _ = step(torch.float64)   # the truth the CUDA gradients are measured against
    loss, logits, grads32, bufs = step(torch.float32)    # what the reference's fp32 CPU path returns
    store = {'loss': np.float64(loss), 'loss64': np.float64(loss64), 'logits': logits, 'logits64': logits64.astype(np.float32),
             'digest': digest}
    for key, val in batch.items():
        store['in_' + key] = val.astype(np.int32) if val.dtype == np.int64 else val
    names, norms, sums, samples, ref32_err = [], [], [], [], []
    for i, ((name, g), (_, g32)) in enumerate(zip(grads64, grads32)):
        names.append(name)
        norms.append(np.sqrt((g * g).sum()))
        sums.append(g.sum())
        samples.append(g[sample_positions(i, g.size)])
        ref32_err.append(np.sqrt(((g32 - g) ** 2).sum()))  # absolute L2 error of the fp32 reference gradient
    store['grad_names'] = np.array(names)
    store['grad_norm'] = np.array(norms)
    store['grad_sum'] = np.array(sums)
    store['grad_samples'] = np.array(samples, dtype=np.float64)
    store['grad_ref32_l2err'] = np.array(ref32_err)
    for name, buf in bufs.items():
        store['buf_' + name] = buf
    np.savez_compressed(os.path.join(HERE, 'train_step.npz'), **store)
    rel = np.array(ref32_err) / np.maximum(np.array(norms), 1e-30)
    print('loss', loss, loss64, 'params', len(names), 'fp32 reference gradient error vs float64: median', np.median(rel), 'tensors above 1e-2:',
          [n for n, r, m in zip(names, rel, norms) if r > 1e-2 and m > 1e-9])


if __name__ == '__main__':
    main()
