"""Marching-cubes case table, GENERATED (not transcribed): for each of the 256 sign configurations of a cell's corners the
triangles of the zero-level set as triples of cell-edge numbers.

Construction: on every cell face the crossed edges are joined by segments (two crossings: one segment; four crossings, i.e.
an ambiguous face with diagonal inside corners: each inside corner is cut off on its own  --  the rule depends on the face's
corner signs only, so the two cells sharing the face agree and the mesh is watertight); the segments close into loops, each
loop is triangulated as a fan and oriented with its normal towards the positive (outside) side, which is the orientation
``skimage.measure.marching_cubes(gradient_direction='descent')`` gives the reference (source/poco_utils.py:96).  Every vertex
lies on a grid edge, so the reference's bisection refinement (poco_utils.py:111-168) applies to all of them.

Corner i of a cell sits at offset (i & 1, (i >> 1) & 1, (i >> 2) & 1) in (x, y, z) = volume index order; edge e joins
``EDGE_CORNERS[e]``, runs along axis ``EDGE_AXIS[e]`` and starts at the cell-local grid vertex ``EDGE_ORIGIN[e]``.
"""
import numpy as np

CORNERS = np.array([[i & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)], dtype=np.int32)
EDGE_CORNERS = np.array([(a, a | (1 << d)) for d in range(3) for a in range(8) if not a & (1 << d)], dtype=np.int32)  # [12,2]
EDGE_AXIS = np.array([d for d in range(3) for a in range(8) if not a & (1 << d)], dtype=np.int32)
EDGE_ORIGIN = CORNERS[EDGE_CORNERS[:, 0]]  # [12,3]
_EDGE_OF = {(int(a), int(b)): e for e, (a, b) in enumerate(EDGE_CORNERS)}


def _edge(a, b):
    return _EDGE_OF[(min(a, b), max(a, b))]


def _faces():
    """the 6 faces as 4 corners in cyclic order"""
    out = []
    for d in range(3):
        u, v = [1 << k for k in range(3) if k != d]
        for side in (0, 1):
            base = side << d
            out.append([base, base | u, base | u | v, base | v])
    return out


def _edge_faces(e):
    """the two cell faces (axis, side) a cell edge lies on"""
    a, b = EDGE_CORNERS[e]
    return {(d, (int(a) >> d) & 1) for d in range(3) if ((int(a) >> d) & 1) == ((int(b) >> d) & 1)}


def _triangulate(loop, mid):
    """triangulation of a loop of edge vertices.  A loop that visits an ambiguous face twice must not get a DIAGONAL inside that
    face: the neighbouring cell could lay the same diagonal into the shared face, which doubles a directed edge or glues the two
    sheets along a non-manifold edge.  All triangulations of the polygon are enumerated (at most 12 vertices), the ones with an in-face
    diagonal are dropped, the shortest total diagonal length wins."""
    n = len(loop)
    faces_of = [_edge_faces(e) for e in loop]

    def diagonal_in_face(a, b):
        adjacent = b - a == 1 or (a == 0 and b == n - 1)
        return not adjacent and bool(faces_of[a] & faces_of[b])

    def in_face(i, k, j):
        return diagonal_in_face(i, k) or diagonal_in_face(k, j) or diagonal_in_face(i, j)

    best = {}

    def solve(i, j):  # best triangulation of the sub-polygon i..j (indices into loop): (cost, triangles) or None
        if j - i < 2:
            return 0.0, []
        if (i, j) in best:
            return best[(i, j)]
        res = None
        for k in range(i + 1, j):
            if in_face(i, k, j):
                continue
            left, right = solve(i, k), solve(k, j)
            if left is None or right is None:
                continue
            cost = left[0] + right[0]
            for a, b in ((i, k), (k, j)):
                if b - a > 1:
                    cost += float(np.sum((mid[loop[a]] - mid[loop[b]]) ** 2))
            if res is None or cost < res[0] - 1e-12:
                res = (cost, left[1] + right[1] + [(loop[i], loop[k], loop[j])])
        best[(i, j)] = res
        return res

    res = solve(0, n - 1)
    assert res is not None, 'no triangulation without an in-face diagonal for loop {}'.format(loop)
    return res[1]


def _case_triangles(case):
    inside = [(case >> i) & 1 for i in range(8)]
    nbr = {}
    for cyc in _faces():
        cross = [k for k in range(4) if inside[cyc[k]] != inside[cyc[(k + 1) % 4]]]
        segs = []
        if len(cross) == 2:
            segs.append((_edge(cyc[cross[0]], cyc[(cross[0] + 1) % 4]), _edge(cyc[cross[1]], cyc[(cross[1] + 1) % 4])))
        elif len(cross) == 4:
            for k in range(4):
                if inside[cyc[k]]:
                    segs.append((_edge(cyc[k - 1], cyc[k]), _edge(cyc[k], cyc[(k + 1) % 4])))
        for a, b in segs:
            nbr.setdefault(a, []).append(b)
            nbr.setdefault(b, []).append(a)
    assert all(len(v) == 2 for v in nbr.values())
    mid = (CORNERS[EDGE_CORNERS[:, 0]] + CORNERS[EDGE_CORNERS[:, 1]]) * 0.5
    tris, seen = [], set()
    for start in sorted(nbr):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            nxt = [n for n in nbr[cur] if n != prev]
            nxt = nxt[0] if nxt else nbr[cur][0]
            if nxt == start:
                break
            loop.append(nxt)
            seen.add(nxt)
            prev, cur = cur, nxt
        fan = _triangulate(loop, mid)
        area = sum(np.cross(mid[b] - mid[a], mid[c] - mid[a]) for a, b, c in fan)
        out_dir = np.zeros(3)
        for e in loop:  # inside corner -> outside corner along every crossed edge of the loop
            a, b = EDGE_CORNERS[e]
            out_dir += (CORNERS[b] - CORNERS[a]) * (1.0 if inside[a] else -1.0)
        if float(np.dot(area, out_dir)) < 0:
            fan = [(a, c, b) for a, b, c in fan]
        tris += fan
    return tris


def build_tri_table():
    """``[256, 3 * MAX_TRIS]`` int8, -1 padded; bit i of the case index = corner i is INSIDE (value < level)"""
    cases = [_case_triangles(c) for c in range(256)]
    width = 3 * max(len(t) for t in cases)
    table = np.full((256, width), -1, dtype=np.int8)
    for c, tris in enumerate(cases):
        flat = [e for t in tris for e in t]
        table[c, :len(flat)] = flat
    return table


TRI_TABLE = build_tri_table()
MAX_TRIS = TRI_TABLE.shape[1] // 3
TRI_COUNT = (TRI_TABLE >= 0).sum(axis=1).astype(np.int32) // 3
