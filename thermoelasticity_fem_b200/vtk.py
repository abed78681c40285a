"""Binary export of a tetrahedral mesh with nodal results for plotting.

The reference plots through meshio + pyvista (``plots.py:6-24``: ``mesh.make_meshio_mesh()``, point data ``T``, vector
``U``, ``warp_by_vector``).  Neither package is needed here: this module writes VTK XML unstructured grids (``.vtu``,
raw appended binary, little endian, 64-bit headers) that ParaView / pyvista / VisIt open directly, and a ``.pvd`` index
for time series.  Pure numpy, host side, no device code: results are already on the host when they are plotted.
"""
import os
import struct

import numpy as np

_VTK_TETRA = 10
_VTK_TYPES = {np.dtype(np.float64): "Float64", np.dtype(np.float32): "Float32", np.dtype(np.int64): "Int64",
              np.dtype(np.int32): "Int32", np.dtype(np.uint8): "UInt8"}


def _as_point_array(name, a, n):
    a = np.ascontiguousarray(a)
    if a.dtype not in _VTK_TYPES:
        a = a.astype(np.float64)
    if a.ndim == 1 and a.shape[0] == n:
        return a, 1
    if a.ndim == 1 and a.shape[0] == 3 * n:          # the reference's interleaved displacement vector (3 per node)
        return a.reshape(n, 3), 3
    if a.ndim == 2 and a.shape[0] == n and a.shape[1] in (1, 3):
        return np.ascontiguousarray(a), int(a.shape[1])
    raise ValueError(f"{name}: expected {n} values, 3 * {n} values or an array [{n}, 3], got shape {a.shape}")


def write_vtu(path, points, tets, point_data=None, cell_data=None):
    """One ``.vtu`` file: points [N, 3], tets [E, 4] (node numbers), point_data / cell_data: name -> array
    (scalars [N] / [E], vectors [N, 3] or the interleaved [3 N] layout of the solvers' ``U``)."""
    points = np.ascontiguousarray(points, dtype=np.float64)
    tets = np.ascontiguousarray(tets, dtype=np.int64)
    if points.ndim != 2 or points.shape[1] != 3 or tets.ndim != 2 or tets.shape[1] != 4:
        raise ValueError("points must be [N, 3] and tets [E, 4]")
    n, e = points.shape[0], tets.shape[0]
    if e and (tets.min() < 0 or tets.max() >= n):
        raise ValueError("tets reference nodes outside the point table")
    blocks, offset = [], 0

    def array_xml(name, a, ncomp, indent):
        nonlocal offset
        raw = a.tobytes()
        xml = (f'{indent}<DataArray type="{_VTK_TYPES[a.dtype]}" Name="{name}" NumberOfComponents="{ncomp}" '
               f'format="appended" offset="{offset}"/>\n')
        blocks.append(struct.pack("<Q", len(raw)) + raw)
        offset += 8 + len(raw)
        return xml

    out = ['<?xml version="1.0"?>\n',
           '<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n',
           '  <UnstructuredGrid>\n', f'    <Piece NumberOfPoints="{n}" NumberOfCells="{e}">\n']
    pd = {}
    for name, a in (point_data or {}).items():
        pd[name] = _as_point_array(name, a, n)
    scalars = next((k for k, (_, c) in pd.items() if c == 1), None)
    vectors = next((k for k, (_, c) in pd.items() if c == 3), None)
    attrs = (f' Scalars="{scalars}"' if scalars else "") + (f' Vectors="{vectors}"' if vectors else "")
    out.append(f'      <PointData{attrs}>\n')
    for name, (a, c) in pd.items():
        out.append(array_xml(name, a, c, "        "))
    out.append('      </PointData>\n      <CellData>\n')
    for name, a in (cell_data or {}).items():
        a = np.ascontiguousarray(a)
        if a.shape != (e,):
            raise ValueError(f"{name}: cell data must have one value per tetrahedron")
        if a.dtype not in _VTK_TYPES:
            a = a.astype(np.float64)
        out.append(array_xml(name, a, 1, "        "))
    out.append('      </CellData>\n      <Points>\n')
    out.append(array_xml("Points", points, 3, "        "))
    out.append('      </Points>\n      <Cells>\n')
    out.append(array_xml("connectivity", tets.reshape(-1), 1, "        "))
    out.append(array_xml("offsets", np.arange(4, 4 * e + 1, 4, dtype=np.int64), 1, "        "))
    out.append(array_xml("types", np.full(e, _VTK_TETRA, dtype=np.uint8), 1, "        "))
    out.append('      </Cells>\n    </Piece>\n  </UnstructuredGrid>\n  <AppendedData encoding="raw">\n_')
    with open(path, "wb") as f:
        f.write("".join(out).encode("ascii"))
        for b in blocks:
            f.write(b)
        f.write(b"\n  </AppendedData>\n</VTKFile>\n")
    return path


def read_vtu(path):
    """Reader for the files ``write_vtu`` produces (round-trip checks and small post-processing scripts; not a general VTK
    reader).  Returns dict(points, tets, point_data, cell_data)."""
    import re
    raw = open(path, "rb").read()
    marker = b'<AppendedData encoding="raw">'
    k = raw.index(marker)
    start = raw.index(b"_", k + len(marker)) + 1
    header = raw[:k].decode("ascii")
    dtypes = {v: k_ for k_, v in _VTK_TYPES.items()}
    arrays, section, where = {}, None, {}
    for line in header.splitlines():
        s = line.strip()
        for sec in ("PointData", "CellData", "Points", "Cells"):
            if s.startswith("<" + sec):
                section = sec
        m = re.match(r'<DataArray type="(\w+)" Name="([^"]+)" NumberOfComponents="(\d+)" format="appended" offset="(\d+)"/>', s)
        if m:
            t, name, ncomp, off = m.group(1), m.group(2), int(m.group(3)), int(m.group(4))
            nbytes = struct.unpack("<Q", raw[start + off:start + off + 8])[0]
            a = np.frombuffer(raw, dtype=dtypes[t], count=nbytes // dtypes[t].itemsize, offset=start + off + 8).copy()
            arrays[(section, name)] = a.reshape(-1, ncomp) if ncomp > 1 else a
            where[(section, name)] = section
    tets = arrays[("Cells", "connectivity")].reshape(-1, 4)
    return dict(points=arrays[("Points", "Points")], tets=tets,
                point_data={n: a for (s, n), a in arrays.items() if s == "PointData"},
                cell_data={n: a for (s, n), a in arrays.items() if s == "CellData"})


def _material_tags(mesh):
    tags = []
    for tag, t in mesh.dict_tet_groups.items():
        tags.append(np.full(len(t), int(tag), dtype=np.int32))
    return np.concatenate(tags) if tags else np.zeros(0, dtype=np.int32)


def export_U_T(mesh, temperature, displacement, path):
    """The data of the reference's ``plot_U_T(mesh, temperature, displacement)`` (``plots.py:6-24``) as one ``.vtu``:
    point data ``T`` [n_nodes] and ``U`` [3 n_nodes] (warp by ``U`` in the viewer), cell data ``tag`` (material group)."""
    return write_vtu(path, mesh.table_nodes, mesh.all_tets(), {"T": temperature, "U": displacement}, {"tag": _material_tags(mesh)})


def export_animation(mesh, temperature, displacement, vec_t, path, step=1):
    """The data of the reference's ``animate_U_T`` (``plots.py:26-75``): histories ``temperature`` [n_nodes, n_t] and
    ``displacement`` [3 n_nodes, n_t] at the times ``vec_t``, every ``step``-th frame, as ``<path>_NNNN.vtu`` files plus
    the ParaView index ``<path>.pvd``.  Returns the list of files written."""
    temperature, displacement, vec_t = np.asarray(temperature), np.asarray(displacement), np.asarray(vec_t)
    if temperature.shape[1] != vec_t.size or displacement.shape[1] != vec_t.size:
        raise ValueError("histories and vec_t disagree on the number of time steps")
    base = path[:-4] if path.endswith(".pvd") else path
    frames = list(range(0, vec_t.size, max(1, int(step))))
    tets, tags = mesh.all_tets(), _material_tags(mesh)
    files = []
    for k, i in enumerate(frames):
        f = f"{base}_{k:04d}.vtu"
        write_vtu(f, mesh.table_nodes, tets, {"T": temperature[:, i], "U": displacement[:, i]}, {"tag": tags})
        files.append(f)
    with open(base + ".pvd", "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1" byte_order="LittleEndian">\n  <Collection>\n')
        for k, i in enumerate(frames):
            f.write(f'    <DataSet timestep="{float(vec_t[i]):.17g}" part="0" file="{os.path.basename(files[k])}"/>\n')
        f.write('  </Collection>\n</VTKFile>\n')
    return files + [base + ".pvd"]
