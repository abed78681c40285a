"""The reference's ``plots.py`` names (post-processing, outside the hot path), import-safe.

Every reference example imports a plotting function at module top level (``example_steadystate.py:13``), so a drop-in
package needs the names.  Rendering needs pyvista (3-D views) / matplotlib (curves), which are optional here: a function
raises a clear ImportError only when it is CALLED without its package.  What does not need them is provided without:
``export_U_T`` / ``export_animation`` write the same data as binary VTK files (``vtk.py``) for any viewer, and the
statistics of the ``*_random`` plots (``confidence_bounds``) are plain numpy.
"""
import numpy as np

from .vtk import export_U_T, export_animation  # noqa: F401  (re-exported: the dependency-free way to look at results)


def _pyvista():
    try:
        import pyvista as pv
        return pv
    except ImportError as exc:  # pragma: no cover
        raise ImportError("3-D rendering needs pyvista (not part of the B200 hot path); export_U_T / export_animation write "
                          "the same data as .vtu / .pvd files without it") from exc


def _pyplot():
    try:
        import matplotlib.pyplot as plt
        return plt
    except ImportError as exc:  # pragma: no cover
        raise ImportError("curve plots need matplotlib (not part of the B200 hot path)") from exc


def _grid(pv, mesh):
    tets = mesh.all_tets()
    cells = np.hstack([np.full((tets.shape[0], 1), 4), tets]).ravel()
    return pv.UnstructuredGrid(cells, np.full(tets.shape[0], 10), mesh.table_nodes)


def plot_U_T(mesh, temperature, displacement, amplification_factor_U=1., save_path=None):
    """reference plots.py:6-24: temperature on the mesh warped by the displacement."""
    pv = _pyvista()
    grid = _grid(pv, mesh)
    grid.point_data['T'] = temperature
    grid['U'] = np.asarray(displacement).reshape((mesh.n_nodes, 3))
    warped = grid.warp_by_vector('U', factor=amplification_factor_U)
    p = pv.Plotter(off_screen=save_path is not None)
    p.add_mesh(warped, show_edges=True, scalars='T')
    p.show_axes()
    p.show(screenshot=save_path)


def animate_U_T(mesh, temperature, displacement, vec_t, save_path, amplification_factor_U=1., fps=10, quality=5, step=1):
    """reference plots.py:26-75: a movie of the warped temperature field, colour range fixed over the whole history."""
    pv = _pyvista()
    temperature, displacement = np.asarray(temperature), np.asarray(displacement)
    clim = [float(temperature.min()), float(temperature.max())]
    p = pv.Plotter(off_screen=True)
    p.open_movie(save_path, framerate=fps, quality=quality)
    actor = None
    for i in [0] + list(range(1, np.asarray(vec_t).size, step)):
        if actor is not None:
            p.remove_actor(actor)
        grid = _grid(pv, mesh)
        grid.point_data['T'] = temperature[:, i]
        grid['U'] = displacement[:, i].reshape((mesh.n_nodes, 3))
        actor = p.add_mesh(grid.warp_by_vector('U', factor=amplification_factor_U), cmap="jet", clim=clim, show_edges=True,
                           scalars='T')
        p.show_axes()
        p.add_title(f't = {vec_t[i]:.3f} s')
        p.write_frame()
    p.close()


def _curve(vec_t, values, ylabel, title, save_path):
    plt = _pyplot()
    _, ax = plt.subplots()
    ax.plot(vec_t, np.asarray(values).flatten(), '-b')
    ax.set_xlabel('t [s]')
    ax.set_ylabel(ylabel)
    ax.set_title(title)
    ax.grid()
    if save_path is not None:
        plt.savefig(save_path)
    else:
        plt.show()


def plot_dofU_vs_t(Udof_num, vec_t, displacement, save_path=None):
    """reference plots.py:77-89."""
    _curve(vec_t, np.asarray(displacement)[Udof_num, :], 'displacement [m]', f'Displacement, U_DOF {Udof_num}', save_path)


def plot_nodeT_vs_t(node_num, vec_t, temperature, save_path=None):
    """reference plots.py:91-103."""
    _curve(vec_t, np.asarray(temperature)[node_num, :], 'Temperature [degrees]', f'Temperature, node {node_num}', save_path)


def confidence_bounds(samples, confidence_level):
    """Bounds of the ``*_random`` plots (reference plots.py:117-125, 150-158): per time step (row of ``samples``
    [n_t, n_samples]) the order statistics left after rejecting ``floor((1 - level) n_samples / 2)`` samples on each
    side.  Returns (lower [n_t], upper [n_t])."""
    samples = np.asarray(samples, dtype=np.float64)
    n_rejected = int(np.floor((1 - confidence_level) * samples.shape[1] / 2))
    ordered = np.sort(samples, axis=1)
    return ordered[:, n_rejected].copy(), ordered[:, -(n_rejected + 1)].copy()


def _random_curves(vec_t, random, deterministic, confidence_level, ylabel, title, save_path):
    plt = _pyplot()
    random = np.asarray(random)
    _, ax = plt.subplots()
    ax.plot(vec_t, np.mean(random, axis=-1).flatten(), '-b', label='statistical mean')
    ax.plot(vec_t, np.asarray(deterministic).flatten(), '-r', label='deterministic response')
    ax.set_xlabel('t [s]')
    ax.set_ylabel(ylabel)
    ax.set_title(title)
    if confidence_level is not None:
        lower, upper = confidence_bounds(random, confidence_level)
        ax.plot(vec_t, lower, '--g', label=f'{confidence_level * 100}% confidence interval')
        ax.plot(vec_t, upper, '--g')
        ax.fill_between(vec_t, lower, upper, color='cyan')
    ax.legend(loc='upper right')
    ax.grid()
    if save_path is not None:
        plt.savefig(save_path)
    else:
        plt.show()


def plot_dofU_vs_t_random(vec_t, displacement_random, displacement_deterministic, confidence_level=None, save_path=None):
    """reference plots.py:105-136: mean of the samples, deterministic response, optional confidence band."""
    _random_curves(vec_t, displacement_random, displacement_deterministic, confidence_level, 'displacement [m]',
                   'Displacement, observed dof', save_path)


def plot_nodeT_vs_t_random(vec_t, temperature_random, temperature_deterministic, confidence_level=None, save_path=None):
    """reference plots.py:138-169."""
    _random_curves(vec_t, temperature_random, temperature_deterministic, confidence_level, 'temperature [degrees]',
                   'Temperature, observed node', save_path)
