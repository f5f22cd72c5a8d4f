"""Tensor-level wrappers over the C ABI (``include/ppsurf_b200.h``).  torch supplies device memory and the current
stream; every wrapper validates dtype / device / contiguity and raises ``PpsError`` on a non-zero status.  No wrapper
falls back to torch math."""
import ctypes

import torch

from . import _lib
from ._lib import check, lib


# queries per decode launch: 148 SMs x 128-row tiles, i.e. whole waves for every persistent tensor-core kernel
# queries per launch of the per-chunk kernels: a multiple of 2 * 148 * 128 (whole waves for every tensor-core kernel: tiles of 2 or 128
# queries on 148 or 296 CTAs).  Four of those units: every launch has a ramp and a tail, and 16 chunks share one neighbour search and one
# projection launch -- 37 888 / 75 776 / 151 552 / 303 104 queries per chunk decode the 131^3 grid in 276 / 275 / 267 / 266 ms; the scratch
# buffers are ~80 KB per query of the chunk (12 GB here, one buffer per device shared by all decoders)
DEFAULT_CHUNK = 4 * 2 * 148 * 128


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError('expected a CUDA tensor')
    if not t.is_contiguous():
        raise ValueError('expected a contiguous tensor')
    if dtype is not None and t.dtype != dtype:
        raise ValueError('expected dtype {}, got {}'.format(dtype, t.dtype))
    return ctypes.c_void_p(t.data_ptr())


def require_device():
    """Raises unless the current CUDA device can run the sm_100a kernels (no CPU fallback exists)."""
    if not torch.cuda.is_available():
        raise _lib.PpsError('ppsurf_b200 needs a CUDA (sm_100a) device; there is no CPU fallback')
    check(lib.pps_check_device())


class KnnIndex:
    """Morton-sorted implicit octree over ``pts [N,3]`` (replaces ``make_kdtree``, source/base/proximity.py:40-64)."""

    def __init__(self, pts: torch.Tensor):
        self.pts = pts
        self.n = pts.shape[0]
        self.buf = torch.empty(lib.pps_knn_index_bytes(self.n), dtype=torch.uint8, device=pts.device)
        check(lib.pps_knn_build(_ptr(pts, torch.float32), self.n, _ptr(self.buf), self.buf.numel(), _stream()))

    def query(self, queries: torch.Tensor, k: int, return_dist: bool = False):
        """``queries [Q,3]`` -> ``idx [Q,k'] int32`` ascending by (dist2, index), ``k' = min(k, N)`` like the reference's
        ``knn`` (source/poco_utils.py:259-260)."""
        k = min(int(k), self.n)
        q = queries.shape[0]
        idx = torch.empty((q, k), dtype=torch.int32, device=queries.device)
        d2 = torch.empty((q, k), dtype=torch.float32, device=queries.device) if return_dist else None
        check(lib.pps_knn_query(_ptr(self.buf), self.n, _ptr(queries, torch.float32), q, k, _ptr(idx), _ptr(d2), _stream()))
        return (idx, d2) if return_dist else idx


def knn(points: torch.Tensor, queries: torch.Tensor, k: int, return_dist: bool = False):
    """one-shot build + query, point-major ``[N,3]`` / ``[Q,3]``"""
    return KnnIndex(points).query(queries, k, return_dist)


def patch_normalize(pts, queries, idx, d2, p):
    q = queries.shape[0]
    out = torch.empty((q, p, 3), dtype=torch.float32, device=pts.device)
    check(lib.pps_patch_normalize(_ptr(pts, torch.float32), _ptr(queries, torch.float32), _ptr(idx, torch.int32),
                                  _ptr(d2, torch.float32), q, p, idx.shape[1], _ptr(out), _stream()))
    return out


def linear(x, w, bias=None, residual=None, gather=None, relu=False, out=None, rows=None):
    """``act(x[gather] @ w.T + bias + residual)``; x ``[M,K]`` (or ``[Nsrc,K]`` with ``gather [M]`` int32), w ``[N,K]``."""
    m = (gather.shape[0] if gather is not None else x.shape[0]) if rows is None else rows
    n, k = w.shape
    if x.shape[-1] != k:
        raise ValueError('linear: x has {} columns, w expects {}'.format(x.shape[-1], k))
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=x.device)
    check(lib.pps_linear(_ptr(x, torch.float32), _ptr(w, torch.float32), _ptr(bias, torch.float32), _ptr(residual, torch.float32),
                         _ptr(gather, torch.int32) if gather is not None else None, _ptr(out), m, n, k, x.stride(-2)
                         if x.dim() > 1 else k, n, 1 if relu else 0, _stream()))
    return out


def grid_queries(r, step, bmin_pad, first=0, count=None, device='cuda'):
    count = r ** 3 - first if count is None else count
    out = torch.empty((count, 3), dtype=torch.float32, device=device)
    check(lib.pps_grid_queries(r, float(step), float(bmin_pad), first, count, _ptr(out), _stream()))
    return out


class RegionVolume:
    """Occupancy volume with the region-growing bookkeeping of ``_create_volume`` (source/poco_utils.py:178-254) on the
    device: ``volume [r,r,r]`` fp32 (NaN = not decoded), ``to_see`` mask, frontier lists as C-order linear indices."""

    def __init__(self, r: int, dilation: int, device):
        self.r, self.dilation = int(r), int(dilation)
        total = self.r ** 3
        self.volume = torch.empty((total,), dtype=torch.float32, device=device)
        self.to_see = torch.empty((total,), dtype=torch.uint8, device=device)
        self.ws = torch.empty(lib.pps_region_workspace_bytes(self.r), dtype=torch.uint8, device=device)
        self.lists = [torch.empty((total,), dtype=torch.int32, device=device) for _ in range(3)]
        self.count = torch.zeros((1,), dtype=torch.int64, device=device)
        check(lib.pps_region_init(self.r, _ptr(self.volume), _ptr(self.to_see), _stream()))

    def _count(self) -> int:
        return int(self.count.item())  # the one host synchronisation per list

    def pending(self, seeds: torch.Tensor) -> torch.Tensor:
        """voxels within ``dilation`` of a seed that have no value yet (ascending index order)"""
        out = self.lists[0]
        check(lib.pps_region_pending(_ptr(seeds, torch.int32), seeds.shape[0], self.r, self.dilation, _ptr(self.volume),
                                     _ptr(self.ws), self.ws.numel(), _ptr(out), _ptr(self.count), _stream()))
        return out[:self._count()]

    def queries(self, ids: torch.Tensor, step, bmin_pad) -> torch.Tensor:
        out = torch.empty((ids.shape[0], 3), dtype=torch.float32, device=ids.device)
        check(lib.pps_region_queries(_ptr(ids, torch.int32), ids.shape[0], self.r, float(step), float(bmin_pad), _ptr(out),
                                     _stream()))
        return out

    def scatter(self, ids: torch.Tensor, values: torch.Tensor):
        check(lib.pps_region_scatter(_ptr(ids, torch.int32), _ptr(values, torch.float32), ids.shape[0], _ptr(self.volume),
                                     _stream()))

    def frontier(self, seeds: torch.Tensor, slot: int) -> torch.Tensor:
        """takes ``seeds`` off ``to_see`` and returns the sign-change frontier around them (written to list ``slot``)"""
        out = self.lists[1 + slot]
        check(lib.pps_region_frontier(_ptr(seeds, torch.int32), seeds.shape[0], self.r, self.dilation, _ptr(self.volume),
                                      _ptr(self.to_see), _ptr(self.ws), self.ws.numel(), _ptr(out), _ptr(self.count), _stream()))
        return out[:self._count()]

    def finish(self, padding: int, out_value: float) -> torch.Tensor:
        check(lib.pps_region_finish(_ptr(self.volume), self.r, int(padding), float(out_value), _stream()))
        return self.volume.view(self.r, self.r, self.r)


_shared_workspace = {}


class Decoder:
    """Per-cloud decoder state: kNN index + hoisted fc1 table; decodes query batches through ``pps_decoder_decode``."""

    def __init__(self, packed, pts: torch.Tensor, latents: torch.Tensor, chunk: int = 18944, path: int = 0):
        """``pts [N,3]``, ``latents [N,C]`` point-major fp32 on the device."""
        self.packed = packed
        self.pts = pts.contiguous()
        self.latents = latents.contiguous()
        self.n = pts.shape[0]
        self.chunk = int(chunk)
        self.path = int(path)
        self.kmax = max(packed.struct.k, packed.struct.num_pts_local)
        if self.n < self.kmax:
            raise ValueError('cloud has {} points, the decoder needs at least {}'.format(self.n, self.kmax))
        self.index = KnnIndex(self.pts)
        self.table = torch.empty((self.n, packed.struct.latent), dtype=torch.float32, device=pts.device)
        check(lib.pps_decoder_point_table(packed.ref, _ptr(self.pts, torch.float32), _ptr(self.latents, torch.float32),
                                          self.n, _ptr(self.table), _stream()))
        self._staging = None
        self._copy_stream = None

    def workspace(self, chunk):
        """decode scratch (about 3 GB for the default chunk), ONE buffer per device shared by every Decoder: decodes are ordered on
        the caller's stream, and a per-decoder buffer made every new cloud pay a multi-GB cudaMalloc"""
        nbytes = lib.pps_decoder_workspace_bytes(self.packed.ref, chunk)
        key = str(self.pts.device)
        ws = _shared_workspace.get(key)
        if ws is None or ws.numel() < nbytes:
            _shared_workspace.pop(key, None)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.pts.device)
            _shared_workspace[key] = ws
        return ws

    def decode(self, queries: torch.Tensor, want_logits=True, want_occ=False, want_idx=False):
        """``queries [Q,3]`` device -> dict(logits [Q,2], occ [Q], idx [Q,kmax])"""
        q = queries.shape[0]
        chunk = max(1, min(self.chunk, q))
        ws = self.workspace(chunk)
        dev = queries.device
        logits = torch.empty((q, 2), dtype=torch.float32, device=dev) if want_logits else None
        occ = torch.empty((q,), dtype=torch.float32, device=dev) if want_occ else None
        idx = torch.empty((q, self.kmax), dtype=torch.int32, device=dev) if want_idx else None
        check(lib.pps_decoder_decode(self.packed.ref, _ptr(self.index.buf), _ptr(self.pts), _ptr(self.table), self.n,
                                     _ptr(queries, torch.float32), q, chunk, _ptr(ws), ws.numel(), _ptr(logits),
                                     _ptr(occ), _ptr(idx), self.path, _stream()))
        return {'logits': logits, 'occ': occ, 'idx': idx}

    def decode_host(self, queries_host: torch.Tensor, occ_host: torch.Tensor = None):
        """``queries_host [Q,3]`` pinned CPU tensor -> ``occ_host [Q]`` pinned CPU tensor (synchronous)."""
        q = queries_host.shape[0]
        if occ_host is None:
            occ_host = torch.empty((q,), dtype=torch.float32).pin_memory()
        chunk = max(1, min(self.chunk, q))
        ws = self.workspace(chunk)
        if self._staging is None or self._staging.numel() < q * 16:
            self._staging = torch.empty(q * 16, dtype=torch.uint8, device=self.pts.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.pts.device)
        assert queries_host.dtype == torch.float32 and queries_host.is_contiguous() and not queries_host.is_cuda
        check(lib.pps_decoder_decode_host(self.packed.ref, _ptr(self.index.buf), _ptr(self.pts), _ptr(self.table), self.n,
                                          ctypes.c_void_p(queries_host.data_ptr()), q, chunk, _ptr(ws), ws.numel(),
                                          _ptr(self._staging), self._staging.numel(),
                                          ctypes.c_void_p(occ_host.data_ptr()), self.path, _stream(),
                                          ctypes.c_void_p(self._copy_stream.cuda_stream)))
        return occ_host

    def projection(self, queries, idx):
        q = queries.shape[0]
        ws = self.workspace(max(q, 1))
        out = torch.empty((q, self.packed.struct.latent), dtype=torch.float32, device=queries.device)
        check(lib.pps_decoder_projection(self.packed.ref, _ptr(self.pts), _ptr(self.table), _ptr(queries, torch.float32),
                                         _ptr(idx, torch.int32), idx.shape[1], q, _ptr(ws), ws.numel(), _ptr(out),
                                         self.path, _stream()))
        return out


def pointnet(packed, patches: torch.Tensor, path: int = 0):
    """``patches [Q,P,3]`` -> ``[Q,C]`` local-branch features; ``P`` must be the ``num_pts_local`` the weights were packed for
    (``pps_decoder_pointnet`` takes the patch size from the weight struct)"""
    if patches.dim() != 3 or patches.shape[1] != packed.struct.num_pts_local or patches.shape[2] != 3:
        raise ValueError('pointnet: patches must be [Q,{},3], got {}'.format(packed.struct.num_pts_local, tuple(patches.shape)))
    q = patches.shape[0]
    nbytes = lib.pps_decoder_workspace_bytes(packed.ref, max(q, 1))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=patches.device)
    out = torch.empty((q, packed.struct.latent), dtype=torch.float32, device=patches.device)
    check(lib.pps_decoder_pointnet(packed.ref, _ptr(patches, torch.float32), q, _ptr(ws), ws.numel(), _ptr(out), int(path),
                                   _stream()))
    return out


def sample_quantized(pts: torch.Tensor, n_support: int, rotations: torch.Tensor, seed: int) -> torch.Tensor:
    """``pts [N,3]`` -> indices ``[n_support]`` int32 of the quantised support sampling; ``rotations [R,9]`` on the device"""
    n = pts.shape[0]
    ws = torch.empty(lib.pps_sample_workspace_bytes(n), dtype=torch.uint8, device=pts.device)
    sel = torch.empty((n_support,), dtype=torch.int32, device=pts.device)
    check(lib.pps_sample_quantized(_ptr(pts, torch.float32), n, n_support, _ptr(rotations, torch.float32), rotations.shape[0],
                                   int(seed) & 0xFFFFFFFF, _ptr(ws), ws.numel(), _ptr(sel), _stream()))
    return sel


def encoder_ids(pts: torch.Tensor, rotations: torch.Tensor, seed: int) -> dict:
    """get_fkaconv_ids for a batch: ``pts [B,N0,3]``, ``rotations [B,4,R,9]`` -> point-major supports ``support1..4 [B,Nl,3]``
    and int32 index tensors ``ids00 ... ids10``"""
    b, n0, _ = pts.shape
    dev = pts.device
    sizes = [n0]
    for _ in range(4):
        sizes.append(max(1, sizes[-1] // 4))
    out = {}
    st = _lib.EncoderIdsOut()
    for lv in range(1, 5):
        out['support%d' % lv] = torch.empty((b, sizes[lv], 3), dtype=torch.float32, device=dev)
        st.support[lv - 1] = out['support%d' % lv].data_ptr()
    for p, (a, c) in enumerate(((0, 0), (0, 1), (1, 1), (1, 2), (2, 2), (2, 3), (3, 3), (3, 4), (4, 4))):
        t = torch.empty((b, sizes[c], min(16, sizes[a])), dtype=torch.int32, device=dev)
        out['ids%d%d' % (a, c)] = t
        st.ids16[p] = t.data_ptr()
    for p, (a, c) in enumerate(((4, 3), (3, 2), (2, 1), (1, 0))):
        t = torch.empty((b, sizes[c], 1), dtype=torch.int32, device=dev)
        out['ids%d%d' % (a, c)] = t
        st.ids1[p] = t.data_ptr()
    ws = torch.empty(lib.pps_encoder_ids_workspace_bytes(n0), dtype=torch.uint8, device=dev)
    check(lib.pps_encoder_ids(_ptr(pts, torch.float32), b, n0, _ptr(rotations, torch.float32), rotations.shape[2],
                              int(seed) & 0xFFFFFFFF, _ptr(ws), ws.numel(), ctypes.byref(st), _stream()))
    return out


def fkaconv(packed, x, pts, support, ids):
    """``x [B,Nin,Cin]``, ``pts [B,Nin,3]``, ``support [B,Ns,3]``, ``ids [B,Ns,kn<=16] int32`` -> ``[B,Ns,Cout]``"""
    b, n_in, cin = x.shape
    if cin != packed.struct.cin:
        raise ValueError('fkaconv: x has {} channels, the packed layer expects {}'.format(cin, packed.struct.cin))
    n_s = support.shape[1]
    kn = ids.shape[-1]
    nbytes = lib.pps_fkaconv_workspace_bytes_for(packed.ref, kn, b, n_s)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    out = torch.empty((b, n_s, packed.struct.cout), dtype=torch.float32, device=x.device)
    check(lib.pps_fkaconv_forward(packed.ref, _ptr(x, torch.float32), _ptr(pts, torch.float32), _ptr(support, torch.float32),
                                  _ptr(ids, torch.int32), kn, b, n_in, n_s, _ptr(ws), ws.numel(), _ptr(out), _stream()))
    return out


def gather_rows(x, rows):
    """``x[rows]`` for ``x [N,C]``, ``rows [M]`` int32 (stream-ordered, no index conversion)"""
    return torch.index_select(x, 0, rows)


def gather_max(x, ids):
    b, n_in, c = x.shape
    n_s, kn = ids.shape[1], ids.shape[2]
    out = torch.empty((b, n_s, c), dtype=torch.float32, device=x.device)
    check(lib.pps_gather_max(_ptr(x, torch.float32), _ptr(ids, torch.int32), b, n_in, n_s, c, kn, _ptr(out), _stream()))
    return out


def global_max(x):
    b, n, c = x.shape
    out = torch.empty((b, c), dtype=torch.float32, device=x.device)
    check(lib.pps_global_max(_ptr(x, torch.float32), b, n, c, _ptr(out), _stream()))
    return out


def latent_accumulate(partial, ids, latent, counts):
    check(lib.pps_latent_accumulate(_ptr(partial, torch.float32), _ptr(ids, torch.int32), ids.shape[0], latent.shape[1],
                                    _ptr(latent, torch.float32), _ptr(counts, torch.float32), _stream()))


def latent_accumulate_rows(partial, rows, ids, latent, counts):
    """``latent[ids[i]] += partial[rows[i]]``, ``counts[ids[i]] += 1``; ``partial [R,C]`` point-major, ``rows`` / ``ids`` int32"""
    check(lib.pps_latent_accumulate_rows(_ptr(partial, torch.float32), _ptr(rows, torch.int32), _ptr(ids, torch.int32),
                                         ids.shape[0], latent.shape[1], _ptr(latent, torch.float32), _ptr(counts, torch.float32),
                                         _stream()))


def latent_finalize(latent, counts):
    check(lib.pps_latent_finalize(_ptr(latent, torch.float32), _ptr(counts, torch.float32), latent.shape[0],
                                  latent.shape[1], _stream()))


# ---- f3: marching cubes + bisection refinement on the device (csrc/mcubes.cu) ------------------------------------------------

_mc_state = {}


def _mc_tables(device):
    """case table on the device, cell-edge tables in constant memory (once per device)"""
    key = str(device)
    if key not in _mc_state:
        import numpy as np
        from . import mc_tables
        corner = np.ascontiguousarray(mc_tables.EDGE_CORNERS.astype(np.int8))
        axis = np.ascontiguousarray(mc_tables.EDGE_AXIS.astype(np.int8))
        origin = np.ascontiguousarray(mc_tables.EDGE_ORIGIN.astype(np.int8))
        check(lib.pps_mc_set_edges(corner.ctypes.data, axis.ctypes.data, origin.ctypes.data))
        _mc_state[key] = torch.from_numpy(mc_tables.TRI_TABLE.copy()).to(device)
    return _mc_state[key]


def marching_cubes(volume: torch.Tensor, level: float = 0.0):
    """``volume [r,r,r]`` fp32 on the device -> ``verts [nv,3]`` fp32 (volume-index coordinates), ``vert_edge [nv]`` int32,
    ``faces [nt,3]`` int32, all on the device.  One host synchronisation (the two counts)."""
    r = volume.shape[0]
    if volume.dim() != 3 or volume.shape[1] != r or volume.shape[2] != r:
        raise ValueError('marching_cubes: expected a cubic [r,r,r] volume')
    dev = volume.device
    table = _mc_tables(dev)
    ws = torch.empty(lib.pps_mc_workspace_bytes(r), dtype=torch.uint8, device=dev)
    counts = torch.zeros((2,), dtype=torch.int64, device=dev)
    vol = volume.contiguous()
    check(lib.pps_mc_count(_ptr(vol, torch.float32), r, float(level), _ptr(table), table.shape[1], _ptr(ws), ws.numel(), _ptr(counts),
                           _stream()))
    nv, nt = (int(c) for c in counts.cpu())
    verts = torch.empty((nv, 3), dtype=torch.float32, device=dev)
    vert_edge = torch.empty((nv,), dtype=torch.int32, device=dev)
    faces = torch.empty((nt, 3), dtype=torch.int32, device=dev)
    if nt > 0:
        check(lib.pps_mc_emit(_ptr(vol, torch.float32), r, float(level), _ptr(table), table.shape[1], _ptr(ws), _ptr(verts),
                              _ptr(vert_edge), _ptr(faces), _stream()))
    return verts, vert_edge, faces


class VertexRefiner:
    """bisection state of the vertex refinement (source/poco_utils.py:111-168) on the device"""

    def __init__(self, volume: torch.Tensor, verts: torch.Tensor, vert_edge: torch.Tensor, step: float, bmin_pad: float):
        nv, dev = verts.shape[0], verts.device
        r = volume.shape[0]
        self.verts_ms = torch.empty((nv, 3), dtype=torch.float32, device=dev)  # every vertex in model space
        va, vb = torch.empty_like(self.verts_ms), torch.empty_like(self.verts_ms)
        pa = torch.empty((nv,), dtype=torch.float32, device=dev)
        pb = torch.empty_like(pa)
        active = torch.empty((nv,), dtype=torch.uint8, device=dev)
        check(lib.pps_refine_init(_ptr(volume.contiguous(), torch.float32), r, _ptr(verts, torch.float32), _ptr(vert_edge, torch.int32), nv,
                                  float(step), float(bmin_pad), _ptr(va), _ptr(vb), _ptr(pa), _ptr(pb), _ptr(self.verts_ms), _ptr(active),
                                  _stream()))
        self.index = torch.nonzero(active).reshape(-1)  # the refined subset, compacted
        self.va, self.vb = va[self.index].contiguous(), vb[self.index].contiguous()
        self.pa, self.pb = pa[self.index].contiguous(), pb[self.index].contiguous()
        self.v = self.verts_ms[self.index].contiguous()

    def update(self, pred: torch.Tensor):
        """``pred [n_active]`` = occupancy at ``self.v``; moves the bracket and sets ``self.v`` to its midpoint"""
        check(lib.pps_refine_update(_ptr(pred, torch.float32), self.v.shape[0], _ptr(self.va), _ptr(self.vb), _ptr(self.pa),
                                    _ptr(self.pb), _ptr(self.v), _stream()))

    def result(self) -> torch.Tensor:
        out = self.verts_ms.clone()
        out[self.index] = self.v
        return out
