"""BASELINE config 1: the reference's examples/example_steadystate.py with the import line changed.

Same mesh, material, Dirichlet sets, surface force and heat flux (reference example_steadystate.py:19-77)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from thermoelasticity_fem_b200.materials import LinearThermoElastic
from thermoelasticity_fem_b200.mesh import Mesh
from thermoelasticity_fem_b200.model import Model
from thermoelasticity_fem_b200.solvers import LinearSteadyState
from thermoelasticity_fem_b200.plots import export_U_T

from _mesh_fixture import sandwich_msh_path

if __name__ == '__main__':
    t0 = time.time()

    mesh = Mesh()
    mesh.load_mesh(sandwich_msh_path())

    material = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=750., alpha=4e-6, T0=20.)
    mesh.set_materials({1: material, 2: material, 3: material})
    mesh.make_elements()

    dict_dirichlet_U = {4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('x', 0.), ('y', 0.), ('z', 0.)]}
    dict_dirichlet_theta = {4: 0., 5: 0.}
    dict_surface_forces = {7: np.array([0., 0., -1e9])}
    dict_heat_flux = {6: -25.}

    model = Model(mesh,
                  dict_dirichlet_U=dict_dirichlet_U, dict_dirichlet_theta=dict_dirichlet_theta,
                  dict_surface_forces=dict_surface_forces, dict_volume_forces=None,
                  dict_heat_flux=dict_heat_flux, dict_heat_source=None)

    solver = LinearSteadyState(model)
    solver.solve()

    # The reference ends with plot_U_T(solver.model.mesh, solver.T, solver.U, save_path='./steadystate.png') (pyvista).
    # The same data as a binary VTK file that ParaView / pyvista open (no optional package needed):
    if len(sys.argv) > 1:
        print('wrote', export_U_T(solver.model.mesh, solver.T, solver.U, sys.argv[1]))

    print(f'max|U| = {np.abs(solver.U).max():.16g}   T in [{solver.T.min():.16g}, {solver.T.max():.16g}]')
    print(f'BiCGSTAB: {solver.solve_info}   timings: {solver.timings}')
    print(f'Total time: {time.time() - t0:.2f} seconds.')
