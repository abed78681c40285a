"""Minimal Gmsh MSH 4.1 ASCII reader / writer (meshio-free).

The reference reads meshes through ``meshio.read`` (``mesh.py:22``) and uses only: points in file
order, one cell block per ``$Elements`` entity block in file order, its type (vertex / triangle /
tetra), its connectivity with node tags mapped to 0-based positions, and the first physical tag of
the block's entity (``cell_data['gmsh:physical']``, ``mesh.py:27-44``).  This module provides exactly
that, plus a writer used to materialise fixtures as ``.msh`` files.
"""
import numpy as np

_TYPE_NAME = {15: "vertex", 1: "line", 2: "triangle", 4: "tetra"}
_TYPE_NODES = {15: 1, 1: 2, 2: 3, 4: 4}
_NAME_TYPE = {"vertex": (15, 0), "line": (1, 1), "triangle": (2, 2), "tetra": (4, 3)}


def read_msh41(path):
    """Returns (points[N,3] float64, blocks) with blocks = list of (type_name, physical_tag, conn[n,k] int64)."""
    with open(path, "r") as fh:
        lines = fh.read().splitlines()
    start = {ln.strip(): i for i, ln in enumerate(lines) if ln.startswith("$")}
    fmt = lines[start["$MeshFormat"] + 1].split()
    if not fmt[0].startswith("4.") or int(fmt[1]) != 0:
        raise ValueError(f"only Gmsh MSH 4.x ASCII files are supported (got '{' '.join(fmt)}')")
    # $Entities -> physical tags
    phys = [dict(), dict(), dict(), dict()]
    if "$Entities" in start:
        i = start["$Entities"] + 1
        counts = [int(v) for v in lines[i].split()]
        i += 1
        for dim in range(4):
            first = 4 if dim == 0 else 7
            for _ in range(counts[dim]):
                f = lines[i].split()
                i += 1
                n_phys = int(f[first])
                phys[dim][int(f[0])] = [int(v) for v in f[first + 1:first + 1 + n_phys]]
    # $Nodes
    i = start["$Nodes"] + 1
    n_blocks, n_nodes = (int(v) for v in lines[i].split()[:2])
    i += 1
    tags = np.empty(n_nodes, dtype=np.int64)
    pts = np.empty((n_nodes, 3), dtype=np.float64)
    k = 0
    for _ in range(n_blocks):
        n = int(lines[i].split()[3])
        i += 1
        tags[k:k + n] = np.array(lines[i:i + n], dtype=np.int64)
        i += n
        pts[k:k + n] = np.array([ln.split()[:3] for ln in lines[i:i + n]], dtype=np.float64).reshape(n, 3)
        i += n
        k += n
    lookup = np.full(int(tags.max()) + 1, -1, dtype=np.int64)
    lookup[tags] = np.arange(n_nodes)
    # $Elements
    i = start["$Elements"] + 1
    n_blocks = int(lines[i].split()[0])
    i += 1
    blocks = []
    for _ in range(n_blocks):
        dim, ent, gtype, n = (int(v) for v in lines[i].split())
        i += 1
        if gtype not in _TYPE_NAME:
            i += n
            continue
        nn = _TYPE_NODES[gtype]
        conn = np.array([ln.split()[1:1 + nn] for ln in lines[i:i + n]], dtype=np.int64).reshape(n, nn)
        i += n
        tag_list = phys[dim].get(ent, [])
        if not tag_list:
            continue                      # meshio keeps untagged blocks; the reference would crash on them
        blocks.append((_TYPE_NAME[gtype], tag_list[0], lookup[conn]))
    return pts, blocks


def write_msh41(path, points, blocks):
    """Writes points and (type_name, physical_tag, conn) blocks as MSH 4.1 ASCII, one entity per block,
    preserving block order (and therefore the reference's element order)."""
    points = np.asarray(points, dtype=np.float64)
    lo, hi = points.min(axis=0), points.max(axis=0)
    ents = [[], [], [], []]
    for b, (name, tag, conn) in enumerate(blocks):
        gtype, dim = _NAME_TYPE[name]
        ents[dim].append((len(ents[dim]) + 1, tag, b))
    with open(path, "w") as fh:
        fh.write("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$Entities\n")
        fh.write("%d %d %d %d\n" % tuple(len(e) for e in ents))
        for dim in range(4):
            for etag, ptag, b in ents[dim]:
                if dim == 0:
                    p = points[np.asarray(blocks[b][2]).reshape(-1)[0]]
                    fh.write("%d %.17g %.17g %.17g 1 %d\n" % (etag, p[0], p[1], p[2], ptag))
                else:
                    fh.write("%d %.17g %.17g %.17g %.17g %.17g %.17g 1 %d 0\n" % (etag, *lo, *hi, ptag))
        fh.write("$EndEntities\n$Nodes\n")
        n = points.shape[0]
        fh.write("1 %d 1 %d\n3 1 0 %d\n" % (n, n, n))
        fh.write("\n".join(str(t) for t in range(1, n + 1)) + "\n")
        fh.write("\n".join("%.17g %.17g %.17g" % tuple(p) for p in points) + "\n")
        fh.write("$EndNodes\n$Elements\n")
        total = sum(len(c) for _, _, c in blocks)
        fh.write("%d %d 1 %d\n" % (len(blocks), total, total))
        eid = 1
        counters = [0, 0, 0, 0]
        for name, tag, conn in blocks:
            gtype, dim = _NAME_TYPE[name]
            counters[dim] += 1
            conn = np.asarray(conn).reshape(len(conn), -1)
            fh.write("%d %d %d %d\n" % (dim, counters[dim], gtype, conn.shape[0]))
            for row in conn:
                fh.write(str(eid) + " " + " ".join(str(int(v) + 1) for v in row) + "\n")
                eid += 1
        fh.write("$EndElements\n")
