"""Tet4 element view of the drop-in API (reference ``elements.py:181-286``).

The reference instantiates one Python object per tetrahedron (``mesh.py:51-58``) and computes the six
element matrices with small dense numpy products.  Here an ``Element`` is a light view into the mesh
tables; its ``compute_mat_*_e`` methods return the same 12x12 / 4x12 / 4x4 / 12x4 numpy arrays, but
the arithmetic runs in the CUDA library (``tefem_element_matrices``).  The per-object route exists for
API compatibility and parity tests; assembly never goes through it (kernel K1 works on the tables).
"""
import numpy as np


class Element:
    def __init__(self, number, material, nodes_nums, nodes_coords, _mesh=None):
        self.number = number
        self.material = material
        self.nodes_nums = nodes_nums
        self.nodes_coords = nodes_coords
        self.vec_nodes_coords = np.reshape(self.nodes_coords, 12)
        # DOF numbering, reference elements.py:192-197: dof = 4*node + {0,1,2: u, 3: theta}
        self.dofs_nums_u = [int(4 * n + c) for n in nodes_nums for c in range(3)]
        self.dofs_nums_t = [int(4 * n + 3) for n in nodes_nums]
        self._mesh = _mesh

    def _compute(self, kind):
        if self._mesh is None:
            raise RuntimeError("Element is not attached to a Mesh; use Mesh.make_elements()")
        return self._mesh.element_matrices((kind,), [self.number])[kind][0]

    def compute_mat_Muu_e(self):
        return self._compute("Muu")

    def compute_mat_Dtu_e(self):
        return self._compute("Dtu")

    def compute_mat_Dtt_e(self):
        return self._compute("Dtt")

    def compute_mat_Kuu_e(self):
        return self._compute("Kuu")

    def compute_mat_Kut_e(self):
        return self._compute("Kut")

    def compute_mat_Ktt_e(self):
        return self._compute("Ktt")


class ElementList:
    """Lazy sequence standing in for the reference's ``mesh.elements`` list: Element objects are built
    on access, so a 40 M-tet mesh does not allocate 40 M Python objects."""

    def __init__(self, mesh):
        self._mesh = mesh

    def __len__(self):
        return self._mesh.n_elements

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        m = self._mesh
        nodes = m._conn[i]
        return Element(i, m._materials_by_index[int(m._mat_id[i])], nodes, m.table_nodes[nodes, :], _mesh=m)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]
