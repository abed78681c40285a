"""Mesh container of the drop-in API (reference ``mesh.py:8-67``).

Same public surface: ``load_mesh(path)``, ``set_materials(dict)``, ``make_elements()``,
``make_meshio_mesh()`` and the attributes ``table_nodes, n_nodes, n_dofs, n_elements,
dict_nodes_groups, dict_tri_groups, dict_tet_groups, dict_materials, elements``.  Differences: the
Gmsh file is read without meshio (``gmsh.py``), ``elements`` is a lazy sequence, and
``make_elements()`` builds the flat element tables (connectivity in the reference's element order =
tag first-appearance order, ``mesh.py:51``; per-element material index) that the device kernels use.
"""
import numpy as np

from .elements import ElementList
from . import gmsh


class Mesh:
    def __init__(self):
        self.path_to_mesh = None
        self.n_nodes = None
        self.n_dofs = None
        self.n_elements = None
        self.table_nodes = None
        self.dict_materials = {}
        self.dict_nodes_groups = {}
        self.dict_tri_groups = {}
        self.dict_tet_groups = {}
        self.elements = []
        self._conn = None
        self._mat_id = None
        self._materials_by_index = None
        self._device = None

    # ---- loading ----------------------------------------------------------------------------------
    def load_mesh(self, path_to_mesh):
        """reference mesh.py:20-44 (grouping of cell blocks by first physical tag, insertion-ordered)."""
        self.path_to_mesh = path_to_mesh
        points, blocks = gmsh.read_msh41(path_to_mesh)
        self.from_blocks(points, blocks)

    def from_blocks(self, points, blocks):
        self.table_nodes = np.ascontiguousarray(points, dtype=np.float64)
        self.n_nodes = self.table_nodes.shape[0]
        self.n_dofs = self.n_nodes * 4
        target = {"vertex": self.dict_nodes_groups, "triangle": self.dict_tri_groups, "tetra": self.dict_tet_groups}
        for name, tag, conn in blocks:
            if name not in target:
                continue
            d = target[name]
            data = np.asarray(conn).reshape(-1) if name == "vertex" else np.asarray(conn)
            d[tag] = data if tag not in d else np.concatenate((d[tag], data), axis=0)
        self._invalidate()

    def from_arrays(self, points, dict_tet_groups, dict_tri_groups=None, dict_nodes_groups=None):
        """Direct construction from arrays (synthetic meshes, ``synthetic.kuhn_box``)."""
        self.table_nodes = np.ascontiguousarray(points, dtype=np.float64)
        self.n_nodes = self.table_nodes.shape[0]
        self.n_dofs = self.n_nodes * 4
        self.dict_tet_groups = dict(dict_tet_groups)
        self.dict_tri_groups = dict(dict_tri_groups or {})
        self.dict_nodes_groups = dict(dict_nodes_groups or {})
        self._invalidate()
        return self

    def _invalidate(self):
        self._conn = None
        self._mat_id = None
        self._device = None
        self.elements = []

    def set_materials(self, dict_materials):
        self.dict_materials = dict_materials
        self._invalidate()

    # ---- element tables -----------------------------------------------------------------------------
    def make_elements(self):
        """reference mesh.py:49-58: element order = iteration over dict_tet_groups, rows in order."""
        conns, ids, mats = [], [], []
        for k, (tag, table_tets) in enumerate(self.dict_tet_groups.items()):
            material = self.dict_materials[tag]          # KeyError like the reference when a tag has no material
            conns.append(np.asarray(table_tets, dtype=np.int64))
            ids.append(np.full(len(table_tets), k, dtype=np.int32))
            mats.append(material)
        self._conn = np.concatenate(conns, axis=0)
        self._mat_id = np.concatenate(ids)
        self._materials_by_index = mats
        self.n_elements = self._conn.shape[0]
        self.elements = ElementList(self)
        self._device = None

    def element_tables(self):
        if self._conn is None:
            self.make_elements()
        table = np.array([m.table_row() for m in self._materials_by_index])
        return self._conn, self._mat_id, table

    def all_tets(self):
        return np.concatenate([np.asarray(v) for v in self.dict_tet_groups.values()], axis=0)

    def device_mesh(self):
        """Mesh tables + symbolic schedule on the GPU (built once, cached)."""
        if self._device is None:
            from .device import DeviceMesh
            conn, mat_id, table = self.element_tables()
            self._device = DeviceMesh(self.table_nodes, conn, mat_id, table)
        return self._device

    def element_matrices(self, kinds=("Muu", "Dtu", "Dtt", "Kuu", "Kut", "Ktt"), elem_idx=None):
        """Batched element matrices as numpy arrays (computed on the GPU)."""
        out = self.device_mesh().element_matrices(tuple(kinds), elem_idx)
        return {k: v.cpu().numpy() for k, v in out.items()}

    def reference_temperature(self):
        """Mean of T0 over the material tags whose tets contain the node (reference solvers.py:51-60),
        vectorised."""
        acc = np.zeros(self.n_nodes)
        cnt = np.zeros(self.n_nodes)
        for tag, material in self.dict_materials.items():
            nodes = np.unique(np.asarray(self.dict_tet_groups[tag]).reshape(-1))
            acc[nodes] += material.T0
            cnt[nodes] += 1
        with np.errstate(invalid="ignore", divide="ignore"):
            return acc / cnt

    def make_meshio_mesh(self):
        """reference mesh.py:60-67 (needs meshio; plotting only)."""
        import meshio
        return meshio.Mesh(self.table_nodes, [('tetra', self.all_tets().tolist())])
