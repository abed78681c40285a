"""Synthetic Kuhn/Freudenthal tetrahedral box meshes.

The reference ships one mesh (``data/sandwich.msh``); the large benchmark configurations are
structured boxes generated here with the same tag legend as ``data/sandwich.txt:1-16``:
volume tags 1/2/3 = bottom/middle/top z-layers, surface tags 4 (x = 0), 5 (x = Lx), 6 (z = 0),
7 (z = Lz), vertex tag 8 = node nearest the middle of the top face.

Conventions (all deterministic, no randomness unless ``jitter`` is set):

* node id = (i*(ny+1) + j)*(nz+1) + k  (x-major, z fastest)
* every cell is split into 6 tets along the (0,0,0)-(1,1,1) diagonal: for each permutation pi of
  the axes the tet (v, v+e_pi1, v+e_pi1+e_pi2, v+(1,1,1)); odd permutations have their last two
  vertices swapped so that det(Jt) > 0 under the reference's convention (``elements.py:199-210``)
* element order = tag 1 cells, then tag 2, then tag 3 (dict insertion order, ``mesh.py:51``), cells
  lexicographic in (i, j, k) inside a layer, 6 tets per cell in the fixed permutation order below
* boundary triangles are the tet faces lying in the plane, emitted with an ACUTE vertex first so
  that the reference's "area" 0.5*|X12.X13| (``model.py:153``) is non-zero.
"""
import numpy as np

# permutations of (0,1,2) and their parity (True = odd -> swap last two vertices)
_PERMS = [((0, 1, 2), False), ((0, 2, 1), True), ((1, 0, 2), True),
          ((1, 2, 0), False), ((2, 0, 1), False), ((2, 1, 0), True)]


def _cell_tets():
    """[6,4,3] integer corner offsets of the 6 tets of one cell."""
    out = np.zeros((6, 4, 3), dtype=np.int64)
    for t, (p, odd) in enumerate(_PERMS):
        v1 = np.zeros(3, dtype=np.int64)
        v1[p[0]] = 1
        v2 = v1.copy()
        v2[p[1]] = 1
        v3 = np.ones(3, dtype=np.int64)
        verts = [np.zeros(3, dtype=np.int64), v1, v2, v3]
        if odd:
            verts[2], verts[3] = verts[3], verts[2]
        out[t] = np.array(verts)
    return out


CELL_TETS = _cell_tets()


def layer_bounds(nz, n_layers=3):
    """k-ranges [k0, k1) of the z-layers (tags 1..n_layers), as even as possible, bottom first."""
    edges = [round(l * nz / n_layers) for l in range(n_layers + 1)]
    return [(edges[l], edges[l + 1]) for l in range(n_layers)]


def kuhn_box(nx, ny, nz, lx=2.0, ly=0.5, lz=0.3, n_layers=3, jitter=0.0, seed=0, index_dtype=np.int64):
    """Returns (points[N,3] float64, tet_groups {tag: [n,4]}, tri_groups {tag: [n,3]}, node_groups {8: [1]})."""
    sx, sy, sz = ny + 1, nz + 1, 1
    sx *= (nz + 1)
    I, J, K = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing='ij')
    pts = np.stack([I * (lx / nx), J * (ly / ny), K * (lz / nz)], axis=-1).reshape(-1, 3).astype(np.float64)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        interior = ((I > 0) & (I < nx) & (J > 0) & (J < ny) & (K > 0) & (K < nz)).reshape(-1)
        h = np.array([lx / nx, ly / ny, lz / nz])
        pts[interior] += (rng.uniform(-1.0, 1.0, size=(int(interior.sum()), 3)) * jitter) * h

    def nid(i, j, k):
        return (i * (ny + 1) + j) * (nz + 1) + k

    tet_groups = {}
    off = CELL_TETS
    for tag, (k0, k1) in enumerate(layer_bounds(nz, n_layers), start=1):
        if k1 <= k0:
            continue
        ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(k0, k1), indexing='ij')
        ci, cj, ck = ci.reshape(-1, 1, 1), cj.reshape(-1, 1, 1), ck.reshape(-1, 1, 1)
        conn = nid(ci + off[None, :, :, 0], cj + off[None, :, :, 1], ck + off[None, :, :, 2])
        tet_groups[tag] = conn.reshape(-1, 4).astype(index_dtype)

    def plane_tris(axis, at_max):
        """Triangles of the two-triangle split of each boundary quad induced by the Kuhn tets."""
        a, b = [d for d in range(3) if d != axis]
        n = [nx, ny, nz]
        fixed = n[axis] if at_max else 0
        ca, cb = np.meshgrid(np.arange(n[a]), np.arange(n[b]), indexing='ij')
        ca, cb = ca.reshape(-1), cb.reshape(-1)

        def node(da, db):
            ijk = [None, None, None]
            ijk[axis] = np.full_like(ca, fixed)
            ijk[a] = ca + da
            ijk[b] = cb + db
            return nid(*ijk)
        # the Kuhn split cuts every axis-aligned face along its (0,0)-(1,1) diagonal.
        # acute vertex first: start at (0,0) [acute: 45 deg], then the right-angle corner, then (1,1).
        t1 = np.stack([node(0, 0), node(1, 0), node(1, 1)], axis=1)
        t2 = np.stack([node(0, 0), node(0, 1), node(1, 1)], axis=1)
        return np.stack([t1, t2], axis=1).reshape(-1, 3).astype(index_dtype)

    tri_groups = {4: plane_tris(0, False), 6: plane_tris(2, False), 5: plane_tris(0, True), 7: plane_tris(2, True)}
    target = np.array([lx / 2, ly / 2, lz])
    top = nid(I[:, :, nz], J[:, :, nz], K[:, :, nz]).reshape(-1)
    d = np.linalg.norm(pts[top] - target, axis=1)
    node_groups = {8: np.array([top[np.argmin(d)]], dtype=index_dtype)}
    return pts, tet_groups, tri_groups, node_groups


# ---------------------------------------------------------------------------------------------------
# Pieces of the same box for the row-partitioned multi-GPU path: a rank generates only its slab
# ---------------------------------------------------------------------------------------------------
def kuhn_node_coords(ids, nx, ny, nz, lx=2.0, ly=0.5, lz=0.3):
    """Coordinates of the (global) node ids of an un-jittered box: the same values kuhn_box() produces."""
    ids = np.asarray(ids, dtype=np.int64)
    k = ids % (nz + 1)
    j = (ids // (nz + 1)) % (ny + 1)
    i = ids // ((nz + 1) * (ny + 1))
    return np.stack([i * (lx / nx), j * (ly / ny), k * (lz / nz)], axis=-1).astype(np.float64)


def kuhn_slab_tets(nx, ny, nz, ci0, ci1, index_dtype=np.int64):
    """Tets of the cells with x-index in [ci0, ci1) of a single-layer (n_layers = 1) box: GLOBAL node ids and
    GLOBAL element numbers, in the element order of kuhn_box(..., n_layers=1)."""
    ci0, ci1 = max(ci0, 0), min(ci1, nx)
    if ci1 <= ci0:
        return np.zeros((0, 4), dtype=index_dtype), np.zeros(0, dtype=np.int64)
    ci, cj, ck = np.meshgrid(np.arange(ci0, ci1), np.arange(ny), np.arange(nz), indexing='ij')
    ci, cj, ck = ci.reshape(-1, 1, 1), cj.reshape(-1, 1, 1), ck.reshape(-1, 1, 1)
    off = CELL_TETS
    conn = ((ci + off[None, :, :, 0]) * (ny + 1) + (cj + off[None, :, :, 1])) * (nz + 1) + (ck + off[None, :, :, 2])
    first = 6 * ci0 * ny * nz
    return conn.reshape(-1, 4).astype(index_dtype), np.arange(first, first + 6 * (ci1 - ci0) * ny * nz, dtype=np.int64)


def kuhn_face_tris(nx, ny, nz, axis, at_max, index_dtype=np.int64):
    """Boundary triangles (GLOBAL node ids) of one face, identical to the tri groups of kuhn_box()."""
    a, b = [d for d in range(3) if d != axis]
    n = [nx, ny, nz]
    fixed = n[axis] if at_max else 0
    ca, cb = np.meshgrid(np.arange(n[a]), np.arange(n[b]), indexing='ij')
    ca, cb = ca.reshape(-1), cb.reshape(-1)

    def node(da, db):
        ijk = [None, None, None]
        ijk[axis] = np.full_like(ca, fixed)
        ijk[a] = ca + da
        ijk[b] = cb + db
        return (ijk[0] * (ny + 1) + ijk[1]) * (nz + 1) + ijk[2]
    t1 = np.stack([node(0, 0), node(1, 0), node(1, 1)], axis=1)
    t2 = np.stack([node(0, 0), node(0, 1), node(1, 1)], axis=1)
    return np.stack([t1, t2], axis=1).reshape(-1, 3).astype(index_dtype)
