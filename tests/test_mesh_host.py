"""CPU tests of the reconstruction back end's host side: the generated marching-cubes case table (table-independent properties), the
oracle restatement of the vertex refinement, and the trimesh-free mesh cleaning / PLY / point-file helpers."""
import os

import numpy as np
import pytest

from conftest import ROOT  # noqa: F401


def _field(r=25):
    ax = np.linspace(-0.6, 0.6, r, dtype=np.float32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing='ij')
    a = np.sqrt((x + 0.15) ** 2 + y ** 2 + z ** 2) - 0.3
    b = np.sqrt((x - 0.3) ** 2 + (y - 0.1) ** 2 + z ** 2) - 0.2
    c = np.sqrt((x + 0.3) ** 2 + (y + 0.35) ** 2 + (z - 0.3) ** 2) - 0.12  # a separate blob
    return np.minimum(np.minimum(a, b), c).astype(np.float32)


def test_case_table_is_watertight_and_oriented(oracle):
    from ppsurf_b200 import mc_tables as T
    assert T.TRI_TABLE.shape == (256, 3 * T.MAX_TRIS) and T.MAX_TRIS == 5
    assert T.TRI_COUNT[0] == 0 and T.TRI_COUNT[255] == 0
    # complementary cases cut the same edges
    for c in range(256):
        a, b = T.TRI_TABLE[c], T.TRI_TABLE[255 - c]
        assert set(a[a >= 0].tolist()) == set(b[b >= 0].tolist())
    rng = np.random.default_rng(0)
    for vol in (_field(), rng.standard_normal((9, 9, 9)).astype(np.float32)):  # the noise volume hits every ambiguous configuration
        v, ve, f = oracle.marching_cubes(vol, 0.0)
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).astype(np.int64)
        key, rev = e[:, 0] * len(v) + e[:, 1], e[:, 1] * len(v) + e[:, 0]
        assert len(np.unique(key)) == len(key)  # no directed edge twice: consistently oriented
        interior = np.ones(len(key), dtype=bool)  # edges on the volume border have no partner
        g = ve.astype(np.int64) // 3
        r = vol.shape[0]
        xyz = np.stack([g // (r * r), (g // r) % r, g % r], axis=1)
        on_border = ((xyz == 0) | (xyz == r - 1)).any(axis=1)
        interior &= ~(on_border[e[:, 0]] & on_border[e[:, 1]])
        assert np.isin(rev[interior], key).all()  # closed: every interior edge has its reverse
        # vertices sit where the linear interpolant of their grid edge crosses the level
        axis = ve % 3
        lo = vol.reshape(-1)[g]
        hi = vol.reshape(-1)[g + np.where(axis == 0, r * r, np.where(axis == 1, r, 1))]
        t = v[np.arange(len(v)), axis] - xyz[np.arange(len(v)), axis]
        assert np.all((t >= 0) & (t <= 1)) and np.abs(lo + t * (hi - lo)).max() < 1e-5
    v, ve, f = oracle.marching_cubes(_field(), 0.0)
    p = v[f].astype(np.float64)
    assert np.einsum('ij,ij->i', p[:, 0], np.cross(p[:, 1], p[:, 2])).sum() > 0  # normals towards the positive (outside) side
    from ppsurf_b200 import mesh
    labels = mesh.face_components(f.astype(np.int64))
    assert labels.max() + 1 == 2  # two merged spheres + one blob
    for lab in range(2):
        ff = f[labels == lab]
        nv, ne, nf = len(np.unique(ff)), 3 * len(ff) // 2, len(ff)
        assert nv - ne + nf == 2  # each component is a sphere
    nan_vol = _field().copy()
    nan_vol[:6] = np.nan
    v2, _, f2 = oracle.marching_cubes(nan_vol, 0.0)
    assert 0 < len(f2) < len(f) and np.isfinite(v2).all()


def test_refine_vertices_converges(oracle):
    vol = _field().astype(np.float64)
    r = vol.shape[0]
    step, bmin_pad = np.float32(1.2 / (r - 1)), np.float32(-0.6)
    v, ve, f = oracle.marching_cubes(vol, 0.0)

    def predict(q):
        q = q.astype(np.float64)
        a = np.sqrt((q[:, 0] + 0.15) ** 2 + q[:, 1] ** 2 + q[:, 2] ** 2) - 0.3
        b = np.sqrt((q[:, 0] - 0.3) ** 2 + (q[:, 1] - 0.1) ** 2 + q[:, 2] ** 2) - 0.2
        c = np.sqrt((q[:, 0] + 0.3) ** 2 + (q[:, 1] + 0.35) ** 2 + (q[:, 2] - 0.3) ** 2) - 0.12
        return np.minimum(np.minimum(a, b), c).astype(np.float32)

    before = np.abs(predict(v * step + bmin_pad)).max()
    out = oracle.refine_vertices(predict, vol, v, step, bmin_pad, 10)
    after = np.abs(predict(out)).max()
    assert after < before / 20 and after < 2e-4
    same = oracle.refine_vertices(predict, vol, v, step, bmin_pad, 0)
    np.testing.assert_array_equal(same, v * step + bmin_pad)


def test_mesh_cleaning_and_io(tmp_path):
    from ppsurf_b200 import mesh
    # a tetrahedron with a duplicated vertex, a degenerate face, a duplicate face, a NaN vertex with its face, an unreferenced vertex
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0], [np.nan, 0, 0], [5, 5, 5]], dtype=np.float64)
    f = np.array([[0, 1, 2], [0, 3, 1], [0, 2, 3], [4, 3, 2], [1, 1, 2], [2, 1, 0], [0, 1, 5]])
    cv, cf = mesh.clean_simple(v, f)
    assert cv.shape == (4, 3) and cf.shape == (4, 3) and np.isfinite(cv).all() and cf.max() == 3
    # a big component survives, a small one (4 faces <= 6) is dropped
    from oracle import ppsurf_oracle as oracle
    bv, _, bf = oracle.marching_cubes(_field(), 0.0)
    allv = np.concatenate([bv.astype(np.float64), cv + 100.0])
    allf = np.concatenate([bf.astype(np.int64), cf + len(bv)])
    kv, kf = mesh.remove_small_connected_components(*mesh.clean_simple(allv, allf), num_faces=6)
    ev, ef = mesh.clean_simple(bv, bf)  # vertices that fall on a grid vertex are merged, their collapsed faces dropped
    assert kf.shape[0] == ef.shape[0] and kv.shape[0] == ev.shape[0] and kv.max() < 50 and 0.9 * bf.shape[0] < ef.shape[0] <= bf.shape[0]
    # PLY round trip and the point-file readers used by the de-normalisation
    path = str(tmp_path / 'm.ply')
    mesh.write_ply(path, kv, kf)
    back = mesh.read_ply_vertices(path)
    assert back.shape == (kv.shape[0], 3) and np.abs(back - kv).max() < 1e-5
    pts = np.random.default_rng(1).uniform(-3, 7, (50, 3))
    np.save(str(tmp_path / 'p.npy'), pts)
    np.savetxt(str(tmp_path / 'p.xyz'), pts)
    np.testing.assert_allclose(mesh.load_pts(str(tmp_path / 'p.npy')), pts)
    np.testing.assert_allclose(mesh.load_pts(str(tmp_path / 'p.xyz')), pts)
    with pytest.raises(ValueError):
        mesh.load_pts(str(tmp_path / 'p.las'))
    center, scale = mesh.get_points_normalization_info(pts, 0.05)
    norm = (pts - center) / scale
    assert np.abs(norm).max() <= 0.5
    np.testing.assert_allclose(mesh.denormalize_points_with_info(norm, center, scale), pts, atol=1e-12)
    real = os.path.join('/root/reference/datasets/abc_minimal/04_pts_vis')
    if os.path.isdir(real):  # binary PLY written by the reference's tooling
        name = sorted(os.listdir(real))[0]
        got = mesh.load_pts(os.path.join(real, name))
        assert got.ndim == 2 and got.shape[1] >= 3 and np.isfinite(got).all() and got.shape[0] > 1000
