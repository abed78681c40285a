"""The reference's examples/example_transient_3.py with the import line changed: the model of example_transient_2
solved with the REDUCED-ORDER model, 30 + 30 modes (reference example_transient_3.py:19-103, solvers.py:219-354).

The bases come from a block shift-invert Lanczos iteration on the GPU operators (thermoelasticity_fem_b200/rom.py)
instead of scipy's eigs; measured on a B200: 2.9 s for the whole solve_ROM (the reference: 67 s), U / T histories
equal to the reference's to 1e-12 / 2e-11 (tests/test_rom.py)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from thermoelasticity_fem_b200.materials import LinearThermoElastic
from thermoelasticity_fem_b200.mesh import Mesh
from thermoelasticity_fem_b200.model import Model
from thermoelasticity_fem_b200.solvers import LinearTransient

from _mesh_fixture import sandwich_msh_path

if __name__ == '__main__':
    t0 = time.time()

    t_end = 1800e0
    n_t = int(1e3)
    vec_t = np.linspace(0., t_end, n_t)

    mesh = Mesh()
    mesh.load_mesh(sandwich_msh_path())

    material = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=1., alpha=4e-6, T0=20.)
    mesh.set_materials({1: material, 2: material, 3: material})
    mesh.make_elements()

    dict_dirichlet_U = {4: [('x', 0.), ('y', 0.), ('z', 0.)], 5: [('y', 0.), ('z', 0.)]}
    dict_dirichlet_theta = {4: 0., 5: 0.}

    amplitude_fx = -1e9
    arr_f_surf = np.zeros((3, n_t))
    arr_f_surf[0, :] = np.sin(2 * np.pi * vec_t / 900.) * amplitude_fx
    dict_surface_forces = {5: arr_f_surf}

    model = Model(mesh,
                  dict_dirichlet_U=dict_dirichlet_U, dict_dirichlet_theta=dict_dirichlet_theta,
                  dict_surface_forces=dict_surface_forces, dict_volume_forces=None,
                  dict_heat_flux=None, dict_heat_source=None,
                  alpha_M=2e-1, alpha_K=2e-1)

    initial_U = np.zeros((model.mesh.n_nodes * 3, ))
    initial_Udot = np.zeros((model.mesh.n_nodes * 3, ))
    initial_theta = np.zeros((model.mesh.n_nodes, ))
    initial_thetadot = np.zeros((model.mesh.n_nodes, ))

    solver = LinearTransient(model, initial_U, initial_Udot, initial_theta, initial_thetadot, t_end, n_t, 1 / 2, 1 / 4)
    solver.solve_ROM(n_modes_u=30, n_modes_theta=30)

    print(f'max|U| = {np.abs(solver.U).max():.16g}   max|T - T0| = {np.abs(solver.X[3::4]).max():.16g}')
    print(f'timings: {solver.timings}')
    print(f'bases: u {model.rob_info_u}  theta {model.rob_info_t}')
    print(f'smallest eigenvalues: u {model.eig_u[:3]}  theta {model.eig_t[:3]}')
    print(f'Total time: {time.time() - t0:.2f} seconds.')
