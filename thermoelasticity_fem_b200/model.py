"""Model of the drop-in API: DOF maps, operator assembly, loads, Dirichlet elimination.

Mirrors the reference ``model.py:7-339`` (constructor signature, method names, attribute names and
the order in which methods must be called), but the operators live in HBM in plane storage and are
assembled by the CUDA kernel K1; the scipy ``csc_array`` attributes of the reference (``mat_M``,
``mat_D``, ``mat_K``, ``mat_*_f_f``, ``vec_F``, ``vec_F_f``) are exported lazily, on attribute access,
for API compatibility and parity checks - the solvers never touch them.

Deliberate deviations (documented in DESIGN.md):
* DOF lists are numpy int64 arrays instead of Python lists (a 28 M-entry list is unusable);
  ``dirichlet_dofs_U`` is sorted (the reference's ``list(set(...))`` order is CPython-specific).
* the exported matrices carry the reference's STORED pattern approximately (exact zeros dropped like
  ``csc_array(dense)``, ``model.py:92``); ``touched(name)`` gives the reproducible touched pattern.
* the reduced-order bases (``model.py:341-363``) come from a block shift-invert Lanczos iteration on the GPU operators
  (``rom.py``) instead of ARPACK: same subspaces, arbitrary signs / rotations inside multiple eigenvalues.
"""
import numpy as np
import torch

from . import _lib, layout
from .device import load_axpy

_MODIF = {'x': 0, 'y': 1, 'z': 2}


def _letter(str_dof):
    if str_dof not in _MODIF:
        raise ValueError('Unrecognized DOF.')             # reference model.py:63-64, 321-322
    return _MODIF[str_dof]


class Model:
    def __init__(self, mesh,
                 dict_dirichlet_U=None, dict_dirichlet_theta=None,
                 dict_nodal_forces=None, dict_surface_forces=None, dict_volume_forces=None,
                 dict_heat_flux=None, dict_heat_source=None,
                 alpha_M=0., alpha_K=0.):
        self.mesh = mesh

        self.dict_dirichlet_U = dict_dirichlet_U
        self.dict_nodal_forces = dict_nodal_forces
        self.dict_surface_forces = dict_surface_forces
        self.dict_volume_forces = dict_volume_forces

        self.dict_dirichlet_theta = dict_dirichlet_theta
        self.dict_heat_flux = dict_heat_flux
        self.dict_heat_source = dict_heat_source

        self.alpha_M = alpha_M
        self.alpha_K = alpha_K

        self.free_dofs = None
        self.free_dofs_U = None
        self.free_dofs_theta = None
        self.dirichlet_dofs_U = None
        self.dirichlet_dofs_theta = None

        # device state
        self._planes = {}            # name -> [n_planes, P] device tensor
        self._d_alpha = (0., 0.)     # (alpha_M, alpha_K) the D planes were assembled with
        self._export = {}            # name -> cached scipy export
        self._override = {}          # attributes explicitly assigned by the user
        self._reduced = set()
        self._cleared = set()        # operators whose full-matrix view was cleared while the reduced view stays alive
        self._F = None               # device [4N]
        self._rhs = None             # device [4N]: (F - lift) on free DOFs, 0 on constrained DOFs
        self._bc_dev = None          # (dir_mask uint8 [4N], g_fill [4N], g_lift [4N]) on the device, uploaded lazily
        self._mask_host = None       # bool [4N]: constrained DOFs
        self._g_fill_host = None     # [4N]: prescribed values (last tag wins), 0 elsewhere
        self._g_lift_host = None     # [4N]: prescribed values weighted by their multiplicity in the tag lists
        self._lift = None            # device [4N]: K @ g_lift (cached)
        self._has_nonzero_dirichlet = False
        self._load_cache = {}
        self._rom_force_cache = {}
        self._phi_u_dev = self._phi_t_dev = None
        self.n_q_u = self.n_q_t = None
        self.mat_Mrom = self.mat_Drom = self.mat_Krom = self.vec_From = None

    # ------------------------------------------------------------------------------------------
    # DOF maps (reference model.py:51-85)
    # ------------------------------------------------------------------------------------------
    def create_free_dofs_lists(self):
        n_nodes, n_dofs = self.mesh.n_nodes, self.mesh.n_dofs
        dofs_U = []
        g_fill = np.zeros(n_dofs)
        g_lift = np.zeros(n_dofs)
        if self.dict_dirichlet_U is not None:
            for tag, list_dir_u in self.dict_dirichlet_U.items():
                nodes = np.asarray(self.mesh.dict_tri_groups[tag]).reshape(-1)
                uniq = np.unique(nodes)
                last = {}
                for str_dof, val in list_dir_u:
                    modif = _letter(str_dof)
                    dofs_U.append(4 * nodes + modif)
                    last[modif] = val
                for str_dof, _ in list_dir_u:               # model.py:312-325: every list entry repeats the column,
                    modif = _MODIF[str_dof]                 # with the LAST value written for that DOF in this tag
                    g_lift[4 * uniq + modif] += last[modif]
                for modif, val in last.items():             # solvers.py:29-42: later tags overwrite
                    g_fill[4 * uniq + modif] = val
        self.dirichlet_dofs_U = np.unique(np.concatenate(dofs_U)) if dofs_U else np.zeros(0, dtype=np.int64)
        nodes_T = []
        if self.dict_dirichlet_theta is not None:
            for tag, theta in self.dict_dirichlet_theta.items():
                uniq = np.unique(np.asarray(self.mesh.dict_tri_groups[tag]).reshape(-1))
                nodes_T.append(uniq)
                g_lift[4 * uniq + 3] += theta
                g_fill[4 * uniq + 3] = theta
        nodes_T = np.unique(np.concatenate(nodes_T)) if nodes_T else np.zeros(0, dtype=np.int64)
        self.dirichlet_dofs_theta = 4 * nodes_T + 3
        mask = np.zeros(n_dofs, dtype=bool)
        mask[self.dirichlet_dofs_U] = True
        mask[self.dirichlet_dofs_theta] = True
        self.free_dofs = np.nonzero(~mask)[0]
        free4 = (~mask).reshape(n_nodes, 4)
        nodes = np.arange(n_nodes, dtype=np.int64)
        self.free_dofs_U = np.concatenate([4 * nodes[free4[:, c]] + c for c in range(3)])   # x of all nodes, then y, z
        self.free_dofs_theta = 4 * nodes[free4[:, 3]] + 3
        self._mask_host = mask
        self._g_fill_host = g_fill
        self._g_lift_host = g_lift
        self._has_nonzero_dirichlet = bool(np.any(g_lift != 0.0))
        self._bc_dev = None
        self._lift = None
        self._rhs = None

    def _bc(self):
        if self._mask_host is None:
            raise RuntimeError("create_free_dofs_lists() must be called first")
        if self._bc_dev is None:
            dev = _lib.require_cuda()
            self._bc_dev = (torch.from_numpy(self._mask_host.astype(np.uint8)).to(dev),
                            torch.from_numpy(self._g_fill_host).to(dev), torch.from_numpy(self._g_lift_host).to(dev))
        return self._bc_dev

    _dir_mask = property(lambda self: self._bc()[0])
    _g_fill = property(lambda self: self._bc()[1])
    _g_lift = property(lambda self: self._bc()[2])

    # ------------------------------------------------------------------------------------------
    # Operator assembly (reference model.py:87-118) - kernel K1
    # ------------------------------------------------------------------------------------------
    def _assemble(self, ops):
        dm = self.mesh.device_mesh()
        for nm in ops:
            self._planes.pop(nm, None)
            self._cleared.discard(nm)
            self._export.pop("mat_" + nm, None)
            self._override.pop("mat_" + nm, None)
        out = dm.assemble(ops, self.alpha_M, self.alpha_K)
        self._planes.update(out)
        if "D" in ops:
            self._d_alpha = (self.alpha_M, self.alpha_K)
        if "K" in ops:
            self._lift = None
        self._rhs = None

    def assemble_M(self):
        self._assemble("M")

    def assemble_K(self):
        self._assemble("K")

    def assemble_D(self):
        self._assemble("D")

    def assemble_MDK(self):
        """M, D and K in ONE launch (the transient solver's path; same results as the three calls)."""
        self._assemble("MKD")

    def plane_map(self, name):
        if name == "M":
            return layout.plane_map_M()
        if name == "K":
            return layout.plane_map_K()
        return layout.plane_map_D(*self._d_alpha)

    def touched(self, name):
        """scipy CSC of operator 'M' | 'D' | 'K' on the TOUCHED pattern (explicit zeros kept)."""
        return self.mesh.device_mesh().export_csc(self._planes[name], self.plane_map(name), drop_zeros=False)

    # ------------------------------------------------------------------------------------------
    # Loads (reference model.py:120-286)
    # ------------------------------------------------------------------------------------------
    def _tri_weights(self, tag):
        key = ("tri", tag)
        if key not in self._load_cache:
            tri = np.asarray(self.mesh.dict_tri_groups[tag])
            X = self.mesh.table_nodes
            X12 = X[tri[:, 1]] - X[tri[:, 0]]
            X13 = X[tri[:, 2]] - X[tri[:, 0]]
            area = 0.5 * np.abs(np.einsum('ij,ij->i', X12, X13))       # model.py:153: dot product, as is
            self._load_cache[key] = self._node_weights(tri, area / 3)
        return self._load_cache[key]

    def _tet_weights(self, tag):
        key = ("tet", tag)
        if key not in self._load_cache:
            tet = np.asarray(self.mesh.dict_tet_groups[tag])
            X = self.mesh.table_nodes
            X12 = X[tet[:, 1]] - X[tet[:, 0]]
            X13 = X[tet[:, 2]] - X[tet[:, 0]]
            X14 = X[tet[:, 3]] - X[tet[:, 0]]
            vol = np.abs(np.einsum('ij,ij->i', X14, np.cross(X13, X12))) / 6   # model.py:189
            self._load_cache[key] = self._node_weights(tet, vol / 4)
        return self._load_cache[key]

    def _point_weights(self, tag):
        key = ("pt", tag)
        if key not in self._load_cache:
            nodes = np.asarray(self.mesh.dict_nodes_groups[tag]).reshape(-1, 1)
            self._load_cache[key] = self._node_weights(nodes, np.ones(nodes.shape[0]))
        return self._load_cache[key]

    def _node_weights(self, cells, per_cell_weight):
        """Lumped nodal weights: every vertex of a cell receives the cell's weight; summed per node in
        cell order (sequential bincount: deterministic)."""
        nodes = cells.reshape(-1)
        w = np.repeat(per_cell_weight, cells.shape[1])
        uniq, inv = np.unique(nodes, return_inverse=True)
        acc = np.bincount(inv, weights=w, minlength=uniq.shape[0])
        dev = _lib.require_cuda()
        return (torch.from_numpy(uniq.astype(np.int32)).to(dev), torch.from_numpy(acc).to(dev))

    @staticmethod
    def _vector_at(arr, idx_t, message):
        if idx_t is None:
            return np.asarray(arr, dtype=np.float64)
        arr = np.asarray(arr)
        if arr.ndim == 2:
            return arr[:, idx_t]
        if arr.ndim == 1:
            return arr
        raise ValueError(message)

    @staticmethod
    def _scalar_at(val, idx_t, message):
        if idx_t is None or isinstance(val, float):
            return float(val)
        if getattr(val, "ndim", None) == 1:
            return float(val[idx_t])
        raise ValueError(message)

    def load_terms(self, idx_t=None):
        """[(node_idx, weights, coef[4])]: F[4 node + c] += w * coef[c]."""
        terms = []
        if self.dict_nodal_forces is not None:
            for tag, f in self.dict_nodal_forces.items():
                f = self._vector_at(f, idx_t, 'Prescribed nodal force array should have 1 or 2 dimensions.')
                terms.append(self._point_weights(tag) + ((f[0], f[1], f[2], 0.),))
        if self.dict_surface_forces is not None:
            for tag, f in self.dict_surface_forces.items():
                f = self._vector_at(f, idx_t, 'Prescribed surface force array should have 1 or 2 dimensions.')
                terms.append(self._tri_weights(tag) + ((f[0], f[1], f[2], 0.),))
        if self.dict_volume_forces is not None:
            for tag, f in self.dict_volume_forces.items():
                f = self._vector_at(f, idx_t, 'Prescribed volume force array should have 1 or 2 dimensions.')
                terms.append(self._tet_weights(tag) + ((f[0], f[1], f[2], 0.),))
        if self.dict_heat_flux is not None:
            for tag, q in self.dict_heat_flux.items():
                q = self._scalar_at(q, idx_t, 'Prescribed flux array should be a float or have 1 dimension.')
                terms.append(self._tri_weights(tag) + ((0., 0., 0., -q),))        # model.py:227: flux SUBTRACTS
        if self.dict_heat_source is not None:
            for tag, R in self.dict_heat_source.items():
                R = self._scalar_at(R, idx_t, 'Prescribed heat source array should be a float or have 1 dimension.')
                terms.append(self._tet_weights(tag) + ((0., 0., 0., R),))
        return terms

    def assemble_F(self, idx_t=None):
        dev = _lib.require_cuda()
        lib = _lib.load()
        if self._F is None:
            self._F = torch.zeros(self.mesh.n_dofs, dtype=torch.float64, device=dev)
        else:
            self._F.zero_()
        for node, w, coef in self.load_terms(idx_t):
            load_axpy(lib, node, w, coef, self._F)
        self._export.pop("vec_F", None)
        self._override.pop("vec_F", None)
        self._rhs = None

    # ------------------------------------------------------------------------------------------
    # clearing (reference model.py:288-294)
    # ------------------------------------------------------------------------------------------
    def clear_full_matrices(self):
        self.mat_M = None
        self.mat_D = None
        self.mat_K = None

    def clear_full_F(self):
        self.vec_F = None

    # ------------------------------------------------------------------------------------------
    # Dirichlet elimination (reference model.py:296-339) - kernel K2 + plane SpMV
    # ------------------------------------------------------------------------------------------
    def apply_dirichlet_matrices(self):
        self._reduced = {nm for nm in "MDK" if nm in self._planes}
        for nm in "MDK":
            self._export.pop(f"mat_{nm}_f_f", None)

    def lift_vector(self):
        """K @ g (all rows) with g the prescribed values; constant, cached (reference recomputes it every
        call, model.py:326-339).  None when K is not assembled (the reference then skips lifting)."""
        if "K" not in self._planes:
            return None
        if self._lift is None:
            dm = self.mesh.device_mesh()
            self._lift = torch.zeros(self.mesh.n_dofs, dtype=torch.float64, device=dm.device)
            if self._has_nonzero_dirichlet:
                dm.planes_spmv(self._planes["K"], layout.plane_map_K(), self._g_lift, self._lift)
        return self._lift

    def apply_dirichlet_F(self):
        if self._F is None:
            return
        rhs = self._F.clone()
        lift = self.lift_vector()
        if lift is not None and self._has_nonzero_dirichlet:
            rhs -= lift
        rhs.masked_fill_(self._dir_mask.bool(), 0.0)
        self._rhs = rhs
        self._export.pop("vec_F_f", None)

    def rhs_device(self):
        """(F - K g) on free DOFs, 0 on constrained DOFs: the right-hand side in the uniform block layout."""
        if self._rhs is None:
            self.apply_dirichlet_F()
        return self._rhs

    # ------------------------------------------------------------------------------------------
    # scipy / numpy views of the device state (API compatibility; lazy)
    # ------------------------------------------------------------------------------------------
    def _get_full(self, nm):
        key = "mat_" + nm
        if key in self._override:
            return self._override[key]
        if nm not in self._planes or nm in self._cleared:
            return None
        if key not in self._export:
            self._export[key] = self.mesh.device_mesh().export_csc(self._planes[nm], self.plane_map(nm), drop_zeros=True)
        return self._export[key]

    def _set_full(self, nm, value):
        key = "mat_" + nm
        self._export.pop(key, None)
        if value is None:
            # reference: clear_full_matrices() after apply_dirichlet_matrices() frees the full matrices but the reduced
            # ones stay usable (model.py:288-302).  Here the planes back BOTH views: once reduced, they are kept (only the
            # full-matrix view reads None); otherwise they are freed.
            self._override.pop(key, None)
            if nm in self._reduced:
                self._cleared.add(nm)
            else:
                self._planes.pop(nm, None)
                if nm == "K":
                    self._lift = None
        else:
            self._override[key] = value

    def _get_reduced(self, nm):
        key = f"mat_{nm}_f_f"
        if nm not in self._reduced:
            return None
        if key not in self._export:
            A = self.touched(nm).tocsr()[self.free_dofs, :].tocsc()[:, self.free_dofs]
            A.eliminate_zeros()
            self._export[key] = A
        return self._export[key]

    mat_M = property(lambda self: self._get_full("M"), lambda self, v: self._set_full("M", v))
    mat_D = property(lambda self: self._get_full("D"), lambda self, v: self._set_full("D", v))
    mat_K = property(lambda self: self._get_full("K"), lambda self, v: self._set_full("K", v))
    mat_M_f_f = property(lambda self: self._get_reduced("M"))
    mat_D_f_f = property(lambda self: self._get_reduced("D"))
    mat_K_f_f = property(lambda self: self._get_reduced("K"))

    @property
    def vec_F(self):
        if "vec_F" in self._override:
            return self._override["vec_F"]
        if self._F is None:
            return None
        if "vec_F" not in self._export:
            self._export["vec_F"] = self._F.cpu().numpy()
        return self._export["vec_F"]

    @vec_F.setter
    def vec_F(self, value):
        self._export.pop("vec_F", None)
        if value is None:
            self._F = None
            self._override.pop("vec_F", None)
        else:
            self._override["vec_F"] = value

    @property
    def vec_F_f(self):
        if self._rhs is None:
            return None
        if "vec_F_f" not in self._export:
            self._export["vec_F_f"] = self._rhs.cpu().numpy()[self.free_dofs]
        return self._export["vec_F_f"]

    # ------------------------------------------------------------------------------------------
    # reduced-order model (reference model.py:341-440) - rom.py: block shift-invert Lanczos on the GPU operators
    # ------------------------------------------------------------------------------------------
    def _compute_ROB(self, field, n_q, tol, **kw):
        from .rom import FieldEigenProblem
        if self._mask_host is None:
            raise RuntimeError("create_free_dofs_lists() must be called first")            # model.py:342,354
        prob = FieldEigenProblem(self, field, **kw)
        lam, Phi, info = prob.smallest(int(n_q), tol=tol)
        if not info.get("converged", False):
            # scipy's eigs raises ArpackNoConvergence in the reference (model.py:346-347): do not hand an unconverged
            # basis to the reduced-order model silently
            raise RuntimeError(f"reduced-order basis for '{field}' did not converge to tol={tol:g}: {info}")
        return lam, Phi, info

    def compute_ROB_u(self, n_q_u, tol=1e-9, **kw):
        """eigs(K_uu, k=n_q_u, M=M_uu, which='SM') on free_dofs_U (model.py:341-351).  mat_phi_u has unit 2-norm
        columns like ARPACK's; signs (and rotations inside a multiple eigenvalue) are arbitrary there and here."""
        self.n_q_u = int(n_q_u)
        self.eig_u, self._phi_u_dev, self.rob_info_u = self._compute_ROB("u", n_q_u, tol, **kw)
        self._export.pop("mat_phi_u", None)
        self._rom_force_cache = {}

    def compute_ROB_theta(self, n_q_t, tol=1e-9, **kw):
        """eigs(K_tt, k=n_q_t, M=D_tt, which='SM') on free_dofs_theta (model.py:353-363)."""
        self.n_q_t = int(n_q_t)
        self.eig_t, self._phi_t_dev, self.rob_info_t = self._compute_ROB("theta", n_q_t, tol, **kw)
        self._export.pop("mat_phi_t", None)
        self._rom_force_cache = {}

    @property
    def mat_phi_u(self):
        """(len(free_dofs_U), n_q_u) numpy array in the reference's row order (host copy, lazy)."""
        if getattr(self, "_phi_u_dev", None) is None:
            return None
        if "mat_phi_u" not in self._export:
            self._export["mat_phi_u"] = self._phi_u_dev.cpu().numpy()[:, self.free_dofs_U].T.copy()
        return self._export["mat_phi_u"]

    @property
    def mat_phi_t(self):
        if getattr(self, "_phi_t_dev", None) is None:
            return None
        if "mat_phi_t" not in self._export:
            self._export["mat_phi_t"] = self._phi_t_dev.cpu().numpy()[:, self.free_dofs_theta].T.copy()
        return self._export["mat_phi_t"]

    def compute_ROM_matrices(self):
        """Phi^T M Phi, Phi^T D Phi, Phi^T K Phi block by block (model.py:365-393): one plane SpMV per basis vector and
        operator; the bases vanish on the constrained DOFs, so the [free, free] restriction is implicit."""
        from .rom import project
        dm = self.mesh.device_mesh()
        Pu, Pt = self._phi_u_dev, self._phi_t_dev
        nu, nt = self.n_q_u, self.n_q_t
        (Muu,) = project(dm, self._planes["M"], self.plane_map("M"), Pu, [Pu])
        Duu, Dtu = project(dm, self._planes["D"], self.plane_map("D"), Pu, [Pu, Pt])
        (Dtt,) = project(dm, self._planes["D"], self.plane_map("D"), Pt, [Pt])
        (Kuu,) = project(dm, self._planes["K"], self.plane_map("K"), Pu, [Pu])
        Kut, Ktt = project(dm, self._planes["K"], self.plane_map("K"), Pt, [Pu, Pt])
        self.mat_Mrom = np.zeros((nu + nt, nu + nt))
        self.mat_Mrom[:nu, :nu] = Muu
        self.mat_Drom = np.zeros((nu + nt, nu + nt))
        self.mat_Drom[:nu, :nu] = Duu
        self.mat_Drom[nu:, :nu] = Dtu
        self.mat_Drom[nu:, nu:] = Dtt
        self.mat_Krom = np.zeros((nu + nt, nu + nt))
        self.mat_Krom[:nu, :nu] = Kuu
        self.mat_Krom[:nu, nu:] = Kut
        self.mat_Krom[nu:, nu:] = Ktt

    def _rom_lift(self):
        """Phi^T (K_uu g_u) and Phi^T (K_tt g_theta): compute_ROM_forces lifts every field through its OWN diagonal block
        of K only (model.py:399-418 use K[free_U, dir_U], :420-433 K[free_theta, dir_theta]), unlike apply_dirichlet_F."""
        if "lift" not in self._rom_force_cache:
            out = np.zeros(self.n_q_u + self.n_q_t)
            if self._has_nonzero_dirichlet and "K" in self._planes:
                dm = self.mesh.device_mesh()
                full = layout.plane_map_K()
                for sl, Phi, off in ((slice(0, 3), self._phi_u_dev, 0), (slice(3, 4), self._phi_t_dev, self.n_q_u)):
                    pm = np.full((4, 4), -1, dtype=np.int8)
                    pm[sl, sl] = full[sl, sl]
                    y = torch.zeros(self.mesh.n_dofs, dtype=torch.float64, device=dm.device)
                    dm.planes_spmv(self._planes["K"], pm, self._g_lift, y)
                    out[off:off + Phi.shape[0]] = (Phi @ y).cpu().numpy()
            self._rom_force_cache["lift"] = out
        return self._rom_force_cache["lift"]

    def rom_forces(self, idx_t=None):
        """Reduced load vector of time index idx_t without forming F: the loads are lumped nodal weights times an
        amplitude (load_terms), so Phi^T F = sum_terms (Phi[:, 4 node + c] @ w) coef[c]; the projections of the weight
        vectors are cached, which leaves an (n_q x 4) product per load and step."""
        Phi = None
        out = -self._rom_lift().copy()
        for node, w, coef in self.load_terms(idx_t):
            key = (node.data_ptr(), w.data_ptr())
            if key not in self._rom_force_cache:
                if Phi is None:
                    Phi = torch.cat([self._phi_u_dev, self._phi_t_dev])
                idx = 4 * node.long()
                self._rom_force_cache[key] = torch.stack([Phi[:, idx + c] @ w for c in range(4)], dim=1).cpu().numpy()
            out += self._rom_force_cache[key] @ np.asarray(coef, dtype=np.float64)
        return out

    def compute_ROM_forces(self):
        """vec_From from the assembled vec_F (model.py:395-440); assemble_F must have been called."""
        if self._F is None:
            raise RuntimeError("assemble_F() must be called first")
        Phi = torch.cat([self._phi_u_dev, self._phi_t_dev])
        self.vec_From = (Phi @ self._F).cpu().numpy() - self._rom_lift()
