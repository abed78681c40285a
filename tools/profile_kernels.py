"""Small driver for ncu captures (GPU box): a few launches of the assembly kernel (K1) and of the block SpMV (K3)
on a Kuhn box whose operands exceed L2.  usage: python tools/profile_kernels.py [cells]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thermoelasticity_fem_b200 import LinearThermoElastic, Mesh, Model  # noqa: E402
from thermoelasticity_fem_b200.synthetic import kuhn_box  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 100
pts, tets, tris, nodes = kuhn_box(cells, cells, cells, 1.0, 1.0, 1.0, n_layers=1, index_dtype=np.int32)
mesh = Mesh().from_arrays(pts, tets, tris, nodes)
mesh.set_materials({1: LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.)})
mesh.make_elements()
dm = mesh.device_mesh()
model = Model(mesh, dict_dirichlet_U={4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]},
              dict_dirichlet_theta={4: 0., 5: 0.}, alpha_M=0.2, alpha_K=0.2)
model.create_free_dofs_lists()
out = dm.assemble("MKD", 0.2, 0.2)
for _ in range(2):
    dm.assemble("MKD", 0.2, 0.2, out=out, check=False)
outK = {"K": out["K"]}
for _ in range(2):
    dm.assemble("K", 0.2, 0.2, out=outK, check=False)
sys_ = dm.form_system(out, 1.0, 0.9, 0.81, 13, model._dir_mask)
dinv = dm.block_jacobi_scale(sys_)
x = torch.randn(4 * mesh.n_nodes, dtype=torch.float64, device="cuda")
y = torch.zeros_like(x)
for _ in range(4):
    dm.spmv(sys_, x, y)
torch.cuda.synchronize()
print("done", mesh.n_elements, dm.schedule.n_blocks)
