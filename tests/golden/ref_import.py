"""Import the UNMODIFIED reference (cg-tuwien/ppsurf, read-only at /root/reference) in this container.

Test infrastructure only.  Used by ``make_golden.py`` (golden-vector generation) and by CPU tests that
can see ``/root/reference``; nothing on the GPU box may depend on it.

The reference needs packages that are not installed here.  They are replaced by the
smallest possible stubs *before* the reference is imported:

* ``pytorch_lightning``            -> ``LightningModule = torch.nn.Module``
* ``trimesh``, ``overrides``, ``pysdf`` -> empty classes / identity decorator
* ``pykdtree.kdtree.KDTree``       -> ``scipy.spatial.cKDTree(leafsize=10)`` with pykdtree's ``query`` signature
  (float64 distances instead of pykdtree's float32: ties at the k-th neighbour may resolve differently)
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('PPSURF_REFERENCE', '/root/reference')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'source', 'ppsurf_model.py'))


def _install_stubs():
    import numpy as np
    import torch

    if 'pytorch_lightning' not in sys.modules:
        pl = types.ModuleType('pytorch_lightning')

        class LightningModule(torch.nn.Module):
            pass

        class LightningDataModule:
            pass

        pl.LightningModule = LightningModule
        pl.LightningDataModule = LightningDataModule
        cb = types.ModuleType('pytorch_lightning.callbacks')
        prog = types.ModuleType('pytorch_lightning.callbacks.progress')
        tq = types.ModuleType('pytorch_lightning.callbacks.progress.tqdm_progress')

        class TQDMProgressBar:
            pass

        tq.TQDMProgressBar = TQDMProgressBar
        pl.callbacks = cb
        cb.progress = prog
        prog.tqdm_progress = tq
        sys.modules['pytorch_lightning'] = pl
        sys.modules['pytorch_lightning.callbacks'] = cb
        sys.modules['pytorch_lightning.callbacks.progress'] = prog
        sys.modules['pytorch_lightning.callbacks.progress.tqdm_progress'] = tq

    if 'trimesh' not in sys.modules:
        tm = types.ModuleType('trimesh')

        class Trimesh:
            pass

        tm.Trimesh = Trimesh
        sys.modules['trimesh'] = tm

    if 'overrides' not in sys.modules:
        ov = types.ModuleType('overrides')

        class EnforceOverrides:
            pass

        ov.EnforceOverrides = EnforceOverrides
        ov.overrides = lambda f: f
        sys.modules['overrides'] = ov

    if 'pysdf' not in sys.modules:
        ps = types.ModuleType('pysdf')

        class SDF:
            pass

        ps.SDF = SDF
        sys.modules['pysdf'] = ps

    if 'pykdtree' not in sys.modules:
        from scipy.spatial import cKDTree
        pk = types.ModuleType('pykdtree')
        pkk = types.ModuleType('pykdtree.kdtree')

        class KDTree:
            def __init__(self, pts, leafsize=10):
                self._dtype = pts.dtype
                self._tree = cKDTree(pts, leafsize=leafsize)

            def query(self, query_pts, k=1, sqr_dists=False, **_):
                d, i = self._tree.query(query_pts, k=k)
                if sqr_dists:
                    d = d * d
                return d.astype(self._dtype), i.astype(np.uint32)

        pkk.KDTree = KDTree
        pk.kdtree = pkk
        sys.modules['pykdtree'] = pk
        sys.modules['pykdtree.kdtree'] = pkk


def import_reference():
    """Returns the reference's ``source`` package (``source.ppsurf_model`` etc. importable afterwards)."""
    if not reference_available():
        raise RuntimeError('reference checkout not found at {}'.format(REFERENCE_ROOT))
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    src = importlib.import_module('source')
    importlib.import_module('source.ppsurf_model')
    importlib.import_module('source.poco_utils')
    importlib.import_module('source.ppsurf_data_loader')
    return src
