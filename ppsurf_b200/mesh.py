"""Host-side mesh post-processing of the reconstruction driver, restated without trimesh (not installed in the build image; the
reference calls it for exactly these steps): ``clean_simple_inplace`` and ``remove_small_connected_components``
(source/base/mesh.py:7-38, used at source/poco_utils.py:105-107,173-174), the PLY export of ``predict_step``
(source/poco_model.py:268) and the reload of the input cloud for the de-normalisation of single-file reconstructions
(source/poco_model.py:256-263, source/occupancy_data_module.py:174-224, source/base/math.py:111-135).

These run once per mesh on a few 10^5 faces; they are plumbing around the hot path, not part of it.
"""
import os
import struct

import numpy as np

MERGE_DIGITS = 8  # trimesh tol.merge = 1e-8: vertices equal after rounding to 8 decimals are one vertex


def remove_unreferenced_vertices(verts: np.ndarray, faces: np.ndarray):
    used = np.zeros(verts.shape[0], dtype=bool)
    used[faces.reshape(-1)] = True
    remap = np.cumsum(used) - 1
    return verts[used], remap[faces]


def clean_simple(verts: np.ndarray, faces: np.ndarray):
    """source/base/mesh.py:7-13 in trimesh's order: unreferenced vertices, faces with a non-finite vertex, vertex merge, degenerate
    faces (a repeated vertex), duplicate faces (same vertex set)"""
    verts = np.asarray(verts, dtype=np.float64)
    faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    verts, faces = remove_unreferenced_vertices(verts, faces)
    finite = np.isfinite(verts).all(axis=1)
    faces = faces[finite[faces].all(axis=1)]
    verts, faces = remove_unreferenced_vertices(verts, faces)
    if verts.shape[0]:
        _, first, inverse = np.unique(np.round(verts, MERGE_DIGITS), axis=0, return_index=True, return_inverse=True)
        order = np.argsort(first)  # keep the vertices in their original order
        rank = np.empty_like(order)
        rank[order] = np.arange(order.shape[0])
        verts, faces = verts[first[order]], rank[inverse.reshape(-1)][faces]
    faces = faces[(faces[:, 0] != faces[:, 1]) & (faces[:, 1] != faces[:, 2]) & (faces[:, 0] != faces[:, 2])]
    if faces.shape[0]:
        _, keep = np.unique(np.sort(faces, axis=1), axis=0, return_index=True)
        faces = faces[np.sort(keep)]
    return remove_unreferenced_vertices(verts, faces)


def face_components(faces: np.ndarray):
    """connected components over face adjacency = two faces sharing an edge that belongs to exactly two faces (trimesh's
    ``face_adjacency``); returns one label per face"""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    nf = faces.shape[0]
    edges = np.sort(faces[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2), axis=1)
    owner = np.repeat(np.arange(nf), 3)
    key = edges[:, 0] * (edges.max() + 1) + edges[:, 1]
    order = np.argsort(key, kind='stable')
    key, owner = key[order], owner[order]
    start = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])
    count = np.diff(np.r_[start, key.shape[0]])
    pair = start[count == 2]
    a, b = owner[pair], owner[pair + 1]
    graph = coo_matrix((np.ones(a.shape[0]), (a, b)), shape=(nf, nf))
    return connected_components(graph, directed=False)[1]


def remove_small_connected_components(verts: np.ndarray, faces: np.ndarray, num_faces: int = 6):
    """source/base/mesh.py:16-38: keep the face components with MORE than ``num_faces`` faces, then clean again"""
    if faces.shape[0] == 0:
        return verts, faces
    labels = face_components(faces)
    sizes = np.bincount(labels)
    keep = sizes[labels] > max(num_faces, 2)  # trimesh lists only components of at least 3 faces (min_len=3)
    return clean_simple(verts, faces[keep])


def write_ply(path: str, verts: np.ndarray, faces: np.ndarray):
    """binary little-endian PLY with float vertices and int32 triangle indices (what ``trimesh.Trimesh.export`` writes for a mesh
    without attributes)"""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    v = np.asarray(verts, dtype='<f4')
    f = np.asarray(faces, dtype='<i4').reshape(-1, 3)
    header = ('ply\nformat binary_little_endian 1.0\nelement vertex {}\nproperty float x\nproperty float y\nproperty float z\n'
              'element face {}\nproperty list uchar int vertex_indices\nend_header\n').format(v.shape[0], f.shape[0])
    rec = np.empty(f.shape[0], dtype=[('n', 'u1'), ('idx', '<i4', (3,))])
    rec['n'] = 3
    rec['idx'] = f
    with open(path, 'wb') as out:
        out.write(header.encode('ascii'))
        out.write(v.tobytes())
        out.write(rec.tobytes())


_PLY_TYPES = {'char': 'i1', 'uchar': 'u1', 'short': 'i2', 'ushort': 'u2', 'int': 'i4', 'uint': 'u4', 'float': 'f4', 'double': 'f8',
              'int8': 'i1', 'uint8': 'u1', 'int16': 'i2', 'uint16': 'u2', 'int32': 'i4', 'uint32': 'u4', 'float32': 'f4', 'float64': 'f8'}


def read_ply_vertices(path: str) -> np.ndarray:
    """vertex coordinates (and whatever scalar vertex properties follow) of an ASCII or binary PLY file, ``[n, props]``"""
    with open(path, 'rb') as f:
        fmt, count, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline().decode('ascii', 'replace').strip()
            if line.startswith('format'):
                fmt = line.split()[1]
            elif line.startswith('element'):
                in_vertex = line.split()[1] == 'vertex'
                if in_vertex:
                    count = int(line.split()[2])
            elif line.startswith('property') and in_vertex:
                parts = line.split()
                if parts[1] == 'list':
                    raise ValueError('list property on PLY vertices is not supported: {}'.format(path))
                props.append((parts[2], _PLY_TYPES[parts[1]]))
            elif line == 'end_header':
                break
            elif line == '' and f.tell() > 1 << 16:
                raise ValueError('no PLY header in {}'.format(path))
        if fmt == 'ascii':
            data = np.loadtxt(f, max_rows=count, ndmin=2)
            return data[:, :len(props)]
        endian = '<' if fmt == 'binary_little_endian' else '>'
        rec = np.frombuffer(f.read(count * sum(np.dtype(t).itemsize for _, t in props)),
                            dtype=[(n, endian + t) for n, t in props], count=count)
        return np.stack([rec[n].astype(np.float64) for n, _ in props], axis=1)


def load_pts(pts_file: str) -> np.ndarray:
    """``OccupancyDataModule.load_pts`` (source/occupancy_data_module.py:174-216) for the formats that need no mesh library:
    NPY / NPZ, whitespace-separated XYZ, PLY"""
    ext = os.path.splitext(pts_file)[1].lower()
    if ext == '.npy':
        return np.load(pts_file)
    if ext == '.npz':
        return np.load(pts_file)['arr_0']
    if ext == '.xyz':
        return np.loadtxt(pts_file, ndmin=2)
    if ext == '.ply':
        return read_ply_vertices(pts_file)
    raise ValueError('Unknown point cloud type: {} (ppsurf_b200 reads npy, npz, xyz and ply without trimesh / laspy)'.format(pts_file))


def get_points_normalization_info(pts: np.ndarray, padding_factor: float = 0.05):
    """source/base/math.py:111-117"""
    bb_min, bb_max = np.min(pts, axis=0), np.max(pts, axis=0)
    return (bb_min + bb_max) * 0.5, np.max(bb_max - bb_min) * (1.0 + padding_factor)


def denormalize_points_with_info(pts: np.ndarray, bb_center: np.ndarray, scale: float):
    """source/base/math.py:130-133"""
    return pts * scale + bb_center[None, :]
