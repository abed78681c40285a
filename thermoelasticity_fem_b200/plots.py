"""Import-safe stand-in for the reference's ``plots.py`` (out of scope: pure post-processing).

Every reference example imports a plotting function at module top level
(``example_steadystate.py:13``), so a drop-in package needs the names.  The functions need pyvista /
matplotlib, which are optional; they raise a clear error only when CALLED without them.
"""


def _need():
    try:
        import pyvista  # noqa: F401
        import matplotlib  # noqa: F401
    except ImportError as exc:  # pragma: no cover
        raise ImportError("plotting needs pyvista and matplotlib (not part of the B200 hot path)") from exc


def plot_U_T(mesh, temperature, displacement, amplification_factor_U=1., save_path=None):
    _need()
    import pyvista as pv
    pts = mesh.table_nodes
    tets = mesh.all_tets()
    cells = __import__("numpy").hstack([__import__("numpy").full((tets.shape[0], 1), 4), tets]).ravel()
    grid = pv.UnstructuredGrid(cells, __import__("numpy").full(tets.shape[0], 10), pts)
    grid.point_data['T'] = temperature
    grid['U'] = displacement.reshape((mesh.n_nodes, 3))
    warped = grid.warp_by_vector('U', factor=amplification_factor_U)
    p = pv.Plotter(off_screen=save_path is not None)
    p.add_mesh(warped, show_edges=True, scalars='T')
    p.show_axes()
    p.show(screenshot=save_path)


def animate_U_T(*args, **kwargs):
    _need()
    raise NotImplementedError("animation is outside the B200 hot path (reference plots.py:26-75)")


def plot_dofU_vs_t(*args, **kwargs):
    _need()
    raise NotImplementedError("outside the B200 hot path (reference plots.py:77-89)")


def plot_nodeT_vs_t(*args, **kwargs):
    _need()
    raise NotImplementedError("outside the B200 hot path (reference plots.py:91-103)")


def plot_dofU_vs_t_random(*args, **kwargs):
    _need()
    raise NotImplementedError("outside the B200 hot path (reference plots.py:105-136)")


def plot_nodeT_vs_t_random(*args, **kwargs):
    _need()
    raise NotImplementedError("outside the B200 hot path (reference plots.py:138-169)")
