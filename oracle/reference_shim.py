"""Import the UNMODIFIED reference package from /root/reference/src (build container only).

TEST INFRASTRUCTURE ONLY - used by ``oracle/make_golden.py`` and by tests that are
skipped when ``/root/reference`` is absent (it does not exist on the GPU box).

The reference cannot be imported as is here because two third-party modules are
imported at module top level and are not installed (no network):

* ``meshio``            (reference ``mesh.py:1``; only ``meshio.read`` at ``mesh.py:22`` and
                         ``meshio.Mesh`` at ``mesh.py:65`` are used)
* ``random_matrices``   (reference ``solvers.py:5``; only the out-of-scope UQ solver uses it)

This file installs two stand-in modules into ``sys.modules`` before importing the
reference: a ``meshio`` whose ``read`` is backed by a small Gmsh 4.1 ASCII parser that
follows meshio 5.3's conventions (points in file order, node tags -> 0-based positions,
one cell block per ``$Elements`` entity block in file order, ``cell_data['gmsh:physical']``
= first physical tag of the block's entity), and an empty ``random_matrices.generators``.
Nothing is written anywhere; the reference sources are read where they lie.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("TEFEM_REFERENCE_ROOT", "/root/reference")
REFERENCE_SRC = os.path.join(REFERENCE_ROOT, "src")
SANDWICH_MSH = os.path.join(REFERENCE_ROOT, "data", "sandwich.msh")
# byte code of the unmodified reference built by oracle/build_ref.py (travels to the GPU box; git-ignored)
REFERENCE_PYC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
# the reference's data/sandwich.msh as arrays (committed fixture): what `meshio.read` returns for a path ending in .npz
SANDWICH_NPZ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sandwich_mesh.npz")

_GMSH_TYPE = {15: ("vertex", 1), 1: ("line", 2), 2: ("triangle", 3), 4: ("tetra", 4)}


def sources_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "thermoelasticity_fem"))


def bytecode_available():
    d = os.path.join(REFERENCE_PYC, "thermoelasticity_fem")
    ver = os.path.join(d, "PYTHON_VERSION")
    if not (os.path.exists(os.path.join(d, "solvers.bc")) and os.path.exists(ver)):
        return False
    with open(ver) as fh:                         # byte code is specific to the minor version of the interpreter
        return fh.read().split(".")[:2] == sys.version.split()[0].split(".")[:2]


def available():
    return sources_available() or bytecode_available()


def sandwich_path():
    """What to pass to the reference's Mesh.load_mesh: its own data/sandwich.msh, or the committed array copy."""
    return SANDWICH_MSH if os.path.exists(SANDWICH_MSH) else SANDWICH_NPZ


class _CellBlock:
    def __init__(self, type_, data):
        self.type = type_
        self.data = data


class _MeshioMesh:
    def __init__(self, points, cells, cell_data=None):
        self.points = points
        self.cells = cells
        self.cell_data = cell_data or {}


def parse_gmsh41(path):
    """Gmsh MSH 4.1 ASCII -> (points[N,3], [(type, phys_tag, conn[n,k]) ...] in file order)."""
    with open(path, "r") as fh:
        tok = fh.read().split("\n")
    pos = {name: i for i, name in enumerate(tok) if name.startswith("$")}
    # --- $Entities: physical tags per (dim, tag)
    phys = {0: {}, 1: {}, 2: {}, 3: {}}
    i = pos["$Entities"] + 1
    counts = [int(v) for v in tok[i].split()]
    i += 1
    for dim in range(4):
        for _ in range(counts[dim]):
            f = tok[i].split()
            i += 1
            tag = int(f[0])
            off = 4 if dim == 0 else 7
            nphys = int(f[off])
            phys[dim][tag] = [int(v) for v in f[off + 1: off + 1 + nphys]]
    # --- $Nodes
    i = pos["$Nodes"] + 1
    nblocks, nnodes = [int(v) for v in tok[i].split()[:2]]
    i += 1
    tags = np.empty(nnodes, dtype=np.int64)
    pts = np.empty((nnodes, 3), dtype=np.float64)
    k = 0
    for _ in range(nblocks):
        n = int(tok[i].split()[3])
        i += 1
        for j in range(n):
            tags[k + j] = int(tok[i + j])
        i += n
        for j in range(n):
            pts[k + j] = [float(v) for v in tok[i + j].split()]
        i += n
        k += n
    tag2idx = np.full(int(tags.max()) + 1, -1, dtype=np.int64)
    tag2idx[tags] = np.arange(nnodes)
    # --- $Elements
    i = pos["$Elements"] + 1
    nblocks = int(tok[i].split()[0])
    i += 1
    blocks = []
    for _ in range(nblocks):
        dim, etag, gtype, n = [int(v) for v in tok[i].split()]
        i += 1
        name, nn = _GMSH_TYPE[gtype]
        conn = np.array([[int(v) for v in tok[i + j].split()[1:1 + nn]] for j in range(n)], dtype=np.int64)
        i += n
        blocks.append((name, phys[dim][etag][0], tag2idx[conn]))
    return pts, blocks


def _read_npz_blocks(path):
    g = np.load(path)
    blocks = [(str(g[f"b{i}_type"]), int(g[f"b{i}_tag"]), g[f"b{i}_conn"].astype(np.int64)) for i in range(int(g["n_blocks"]))]
    return g["points"], blocks


def _meshio_read(path):
    pts, blocks = _read_npz_blocks(path) if str(path).endswith(".npz") else parse_gmsh41(path)
    cells = [_CellBlock(t, c) for t, _, c in blocks]
    cd = {"gmsh:physical": [np.full(c.shape[0], p, dtype=np.int64) for _, p, c in blocks]}
    return _MeshioMesh(pts, cells, cd)


def _install_stand_ins():
    if "meshio" not in sys.modules:
        m = types.ModuleType("meshio")
        m.read = _meshio_read
        m.Mesh = _MeshioMesh
        sys.modules["meshio"] = m
    if "random_matrices" not in sys.modules:
        rm = types.ModuleType("random_matrices")
        gen = types.ModuleType("random_matrices.generators")
        for name in ("SE_0_plus", "SE_plus0", "SE_rect"):
            setattr(gen, name, None)
        rm.generators = gen
        sys.modules["random_matrices"] = rm
        sys.modules["random_matrices.generators"] = gen


def load_reference():
    """Returns the reference's modules (elements, materials, mesh, model, solvers) as a namespace."""
    if not available():
        raise RuntimeError("reference not present: neither sources at %s nor byte code at %s" % (REFERENCE_SRC, REFERENCE_PYC))
    _install_stand_ins()
    import importlib
    ns = types.SimpleNamespace()
    if sources_available():
        if REFERENCE_SRC not in sys.path:
            sys.path.insert(0, REFERENCE_SRC)
    elif "thermoelasticity_fem" not in sys.modules:
        # byte code built by oracle/build_ref.py: <module>.bc files in the .pyc format, imported with sourceless loaders
        import importlib.machinery
        import importlib.util
        d = os.path.join(REFERENCE_PYC, "thermoelasticity_fem")

        def load(fullname, fname, is_pkg):
            path = os.path.join(d, fname)
            loader = importlib.machinery.SourcelessFileLoader(fullname, path)
            spec = importlib.util.spec_from_file_location(fullname, path, loader=loader,
                                                          submodule_search_locations=[d] if is_pkg else None)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[fullname] = mod
            loader.exec_module(mod)
            return mod

        pkg = load("thermoelasticity_fem", "__init__.bc", True)
        for name in ("elements", "materials", "mesh", "model", "solvers"):     # dependency order of the reference's imports
            setattr(pkg, name, load("thermoelasticity_fem." + name, name + ".bc", False))
    for name in ("elements", "materials", "mesh", "model", "solvers"):
        setattr(ns, name, importlib.import_module("thermoelasticity_fem." + name))
    return ns
