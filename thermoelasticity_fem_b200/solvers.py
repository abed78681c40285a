"""Steady-state and implicit transient solvers of the drop-in API (reference ``solvers.py:8-217``).

Same classes, constructor signatures and result attributes as the reference; the linear algebra runs
on the GPU: ``spsolve`` (``solvers.py:26,172``; SuperLU, re-factorised at every time step) is replaced by
a block-Jacobi-scaled BiCGSTAB (kernels K3-K6) on the Dirichlet-eliminated operator embedded in the
uniform 4x4 node-block layout (kernel K2), and the Newmark update (``solvers.py:167-178``) by fused
vector kernels (K7).  Extra keyword arguments (``rtol``, ``max_iter``, ``history`` ...) have defaults,
so reference scripts run unchanged.

``solve_ROM`` (``solvers.py:219-354``) runs on reduced-order bases computed on the GPU (``rom.py``); the
non-parametric UQ solver ``solve_ROM_nonparametric`` (``solvers.py:356-668``) is out of scope.
"""
import time

import numpy as np
import torch

from . import _lib, layout
from .device import GuessExtrapolator, KrylovSolver, newmark_correct, newmark_predict, load_axpy

try:                                    # the reference wraps its time loop in tqdm (solvers.py:163)
    from tqdm import tqdm as _tqdm
except ImportError:                     # pragma: no cover
    def _tqdm(it, **kw):
        return it


MULTILEVEL_MIN_NODES = 20000     # below this the iteration is launch-latency bound and block-Jacobi alone is faster


def _deinterleave(X, n_nodes):
    """reference solvers.py:62-66 / 188-202: X (4N[, n_t]) -> U (3N[, n_t]), theta (N[, n_t])."""
    U = np.zeros((n_nodes * 3,) + X.shape[1:])
    U[::3] = X[::4]
    U[1::3] = X[1::4]
    U[2::3] = X[2::4]
    return U, X[3::4]


def reduced_newmark(Mr, Dr, Kr, forces, q_init, v_init, dt, gamma, beta, n_t, steps=None):
    """Newmark integration (acceleration form, zero initial acceleration) of the dense reduced system
    Mr q'' + Dr q' + Kr q = forces(i), i = 1 .. n_t - 1 (reference solvers.py:255-297).  Host-side: the matrices are
    (n_q x n_q) with n_q ~ 10 - 100; Mr + gamma dt Dr + beta dt^2 Kr is factorised once (the reference calls
    np.linalg.solve every step).  Returns q, q', q'' of shape (n_q, n_t)."""
    import scipy.linalg
    nq = Mr.shape[0]
    Q = np.zeros((3, nq, n_t))
    Q[0, :, 0], Q[1, :, 0] = q_init, v_init
    lu = scipy.linalg.lu_factor(Mr + gamma * dt * Dr + beta * dt ** 2 * Kr)                  # solvers.py:264-266
    for i in (range(1, n_t) if steps is None else steps):
        q0, q1, q2 = Q[0, :, i - 1], Q[1, :, i - 1], Q[2, :, i - 1]
        v_pred = q1 + (1 - gamma) * dt * q2                                                  # predictors, solvers.py:271-275
        x_pred = q0 + dt * q1 + 0.5 * dt ** 2 * (1 - 2 * beta) * q2
        acc = scipy.linalg.lu_solve(lu, forces(i) - Dr @ v_pred - Kr @ x_pred)               # solvers.py:269-277
        Q[2, :, i] = acc
        Q[1, :, i] = v_pred + gamma * dt * acc                                               # solvers.py:278-283
        Q[0, :, i] = x_pred + beta * dt ** 2 * acc
    return Q[0], Q[1], Q[2]


class _SystemSolve:
    """Shared by both solvers: forms the eliminated system and solves with it - BiCGSTAB on the block-Jacobi-scaled
    system (method='bicgstab', any operator), or block-Jacobi-preconditioned CG on the unscaled system (method='cg',
    symmetric positive definite operators: the field systems of the staged steady-state solve)."""

    def __init__(self, model, cM, cD, cK, rtol, max_iter, check_every, precond="auto", dir_mask=None, method="bicgstab",
                 floor_factor=1.0, precond_options=None):
        """dir_mask (uint8 [4N], default: the model's Dirichlet DOFs): DOFs eliminated from the system (identity rows
        and columns); the reduced-order bases pass the complement of one field's free DOFs (rom.py).
        floor_factor: see KrylovSolver.bicgstab (1 = a solve that stalls above rtol raises)."""
        if method not in ("bicgstab", "cg"):
            raise ValueError("method must be 'bicgstab' or 'cg'")
        self.model = model
        self.method, self.floor_factor = method, floor_factor
        self.dm = model.mesh.device_mesh()
        self.dir_mask = model._dir_mask if dir_mask is None else dir_mask
        self.rtol, self.max_iter, self.check_every = rtol, max_iter, check_every
        d_layout = layout.n_planes(layout.plane_map_D(*model._d_alpha)) if "D" in model._planes else 4
        planes = {k: v for k, v in model._planes.items()
                  if (k == "M" and cM != 0.) or (k == "D" and cD != 0.) or (k == "K" and cK != 0.)}
        self.sys = self.dm.form_system(planes, cM, cD, cK, d_layout, self.dir_mask)
        self.dinv = self.dm.block_jacobi_scale(self.sys, scale=(method == "bicgstab"))
        self.krylov = KrylovSolver(self.dm)
        if method == "cg":
            if precond not in ("auto", "block_jacobi"):
                raise ValueError("method='cg' runs with the block-Jacobi preconditioner")
            precond = "block_jacobi"
        # right preconditioner on top of the block-Jacobi scaling: aggregation coarse correction (multilevel.py);
        # precond="block_jacobi" keeps the scaling alone
        self.precond = None
        if precond == "auto":      # the coarse correction pays off once the iteration is bandwidth- not launch-bound
            precond = "multilevel" if self.dm.n_nodes >= MULTILEVEL_MIN_NODES else "block_jacobi"
        if precond == "multilevel":
            from .multilevel import TwoLevelPreconditioner
            self.precond = TwoLevelPreconditioner(self.dm, self.sys, **(precond_options or {}))
            self.krylov.set_precond(self.precond)
        elif precond != "block_jacobi":
            raise ValueError("precond must be 'auto', 'multilevel' or 'block_jacobi'")
        self.b = torch.empty(model.mesh.n_dofs, dtype=torch.float64, device=self.dm.device)
        self.info = []

    def solve(self, rhs, x):
        """x <- solution of A_ff x_f = rhs_f (constrained entries of x stay 0)."""
        if self.method == "cg":
            self.b.copy_(rhs)
            self.b.masked_fill_(self.dir_mask.bool(), 0.0)
            info = self.krylov.cg(self.sys, self.dinv, self.b, x, self.rtol, self.max_iter, self.check_every, self.floor_factor)
        else:
            self.dm.block_jacobi_apply(self.dinv, rhs, self.b, self.dir_mask)
            info = self.krylov.bicgstab(self.sys, self.b, x, self.rtol, self.max_iter, self.check_every, self.floor_factor)
        self.info.append(info)
        return info


class LinearSteadyState:
    """Steady-state solver for linear thermoelasticity (reference solvers.py:8-66).

    solver: how spsolve(K_ff, F_f) (solvers.py:26) is replaced.
      'staged'  - K has no theta -> u block (model.py:94-103: uu <- Kuu, ut <- Kut, tt <- Ktt), so the solve is a block
                  back-substitution: K_tt theta = F_t, then K_uu u = F_u - K_ut theta, two SYMMETRIC POSITIVE DEFINITE
                  systems solved with block-Jacobi-preconditioned CG (one SpMV per iteration, no breakdown);
      'coupled' - preconditioned BiCGSTAB on the whole non-symmetric operator (the transient solver's path; with the
                  aggregation coarse correction on large meshes);
      'auto'    - 'staged' below MULTILEVEL_MIN_NODES nodes (where block-Jacobi is the preconditioner anyway), 'coupled'
                  above (the coarse correction needs 20 x fewer iterations there)."""

    def __init__(self, model, rtol=1e-10, max_iter=50000, check_every=32, precond="auto", solver="auto", floor_factor=1.0,
                 comm=None, precond_options=None):
        """comm: None (one GPU) | 'auto' | distributed.Comm - row-partitioned solve on the GPUs of the torch.distributed job
        (dist_api.py; the coupled BiCGSTAB path); every rank receives the full X, U, T."""
        if solver not in ("auto", "staged", "coupled"):
            raise ValueError("solver must be 'auto', 'staged' or 'coupled'")
        self.model = model
        self.comm = comm
        self.precond_options = precond_options      # keyword arguments of multilevel.TwoLevelPreconditioner (experiments)
        self.rtol, self.max_iter, self.check_every, self.precond = rtol, max_iter, check_every, precond
        self.solver, self.floor_factor = solver, floor_factor
        self.X = None
        self.U = None
        self.T = None
        self.X_device = None
        self.solve_info = None
        self.timings = {}

    def _solve_distributed(self):
        from . import dist_api
        m = self.model
        t0 = time.perf_counter()
        comm = dist_api.make_comm(self.comm)
        m.create_free_dofs_lists()
        zeros = np.zeros(m.mesh.n_dofs)
        pm, ds = dist_api.distributed_transient(m, zeros, zeros, 1.0, 2, 0.5, 0.25, comm, self.rtol, self.max_iter,
                                                self.check_every, self.precond, self.floor_factor, MULTILEVEL_MIN_NODES,
                                                self.precond_options)
        ds.prepare(steady=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        self.solve_info = ds.solve_steady()
        self.X = pm.gather(ds.owned_state()[0])
        t2 = time.perf_counter()
        self.U, theta = _deinterleave(self.X, m.mesh.n_nodes)
        self.T = theta + m.mesh.reference_temperature()
        self.timings = dict(assemble_s=t1 - t0, solve_s=t2 - t1, post_s=time.perf_counter() - t2)
        self._distributed = (pm, ds)

    def solve(self):
        if self.comm is not None:
            return self._solve_distributed()
        m = self.model
        t0 = time.perf_counter()
        m.create_free_dofs_lists()
        m.assemble_K()
        m.assemble_F()
        m.apply_dirichlet_matrices()
        m.apply_dirichlet_F()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        solver = self.solver
        if solver == "auto":
            solver = "staged" if (m.mesh.n_nodes < MULTILEVEL_MIN_NODES and self.precond in ("auto", "block_jacobi")) else "coupled"
        if solver == "staged":
            x = self._solve_staged(m)
        else:
            ss = _SystemSolve(m, 0., 0., 1., self.rtol, self.max_iter, self.check_every, self.precond,
                              floor_factor=self.floor_factor, precond_options=self.precond_options)
            x = torch.zeros(m.mesh.n_dofs, dtype=torch.float64, device=ss.dm.device)
            self.solve_info = ss.solve(m.rhs_device(), x)
        x.masked_fill_(m._dir_mask.bool(), 0.0)   # eliminated unknowns (a coarse correction may leave round-off there)
        x += m._g_fill                       # prescribed values on the constrained DOFs (solvers.py:27-49)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        self.X_device = x
        self.X = x.cpu().numpy()
        self.U, theta = _deinterleave(self.X, m.mesh.n_nodes)
        self.T = theta + m.mesh.reference_temperature()       # solvers.py:51-66
        self.timings = dict(assemble_s=t1 - t0, solve_s=t2 - t1, post_s=time.perf_counter() - t2)


    def _solve_staged(self, m):
        """theta from K_tt, then u from K_uu with the coupling moved to the right-hand side; each system keeps the uniform
        4-DOF-per-node layout with the other field's DOFs as identity rows (like the reduced-order bases, rom.py)."""
        dm = m.mesh.device_mesh()
        n = m.mesh.n_dofs
        is_t = (torch.arange(n, device=dm.device) % 4) == 3
        dirm = m._dir_mask.bool()
        rhs = m.rhs_device()
        kw = dict(precond="block_jacobi", method="cg", floor_factor=self.floor_factor)
        ss = _SystemSolve(m, 0., 0., 1., self.rtol, self.max_iter, self.check_every,
                          dir_mask=(dirm | ~is_t).to(torch.uint8), **kw)
        x = torch.zeros(n, dtype=torch.float64, device=dm.device)
        info_t = ss.solve(rhs, x)                                        # K_tt theta = F_t
        del ss
        rhs_u = rhs.clone()
        dm.planes_spmv(m._planes["K"], layout.plane_map_K(), x, rhs_u, alpha=-1.0, beta_y=1.0)    # F_u - K_ut theta
        ss = _SystemSolve(m, 0., 0., 1., self.rtol, self.max_iter, self.check_every,
                          dir_mask=(dirm | is_t).to(torch.uint8), **kw)
        xu = torch.zeros_like(x)
        info_u = ss.solve(rhs_u, xu)                                     # K_uu u = F_u - K_ut theta
        x += xu
        self.solve_info = dict(method="staged-cg", iterations=info_t["iterations"] + info_u["iterations"],
                               restarts=info_t["restarts"] + info_u["restarts"], relres=max(info_t["relres"], info_u["relres"]),
                               bnorm=info_u["bnorm"], floor=info_t["floor"] or info_u["floor"], stages=dict(theta=info_t, u=info_u))
        return x


class LinearTransient:
    """Transient solver (Newmark, acceleration form) for linear thermoelasticity (reference solvers.py:69-217).

    history: 'full' keeps X, Xdot, Xdotdot of shape (n_dofs, n_t) like the reference (solvers.py:125-127);
    'last' keeps only the final state (needed at 28 M DOF where one history array would be 45 GB);
    an integer k keeps every k-th step.  `on_step(i, x, v, a)` is called with device tensors after each step."""

    def __init__(self, model,
                 initial_U, initial_Udot,
                 initial_theta, initial_thetadot,
                 t_end, n_t,
                 gamma=0.5, beta=0.25,
                 rtol=1e-10, max_iter=50000, check_every=32, history='full', warm_start=True, guess_order=0,
                 on_step=None, progress=True, precond="auto", floor_factor=1.0, comm=None, precond_options=None):
        self.model = model
        self.gamma = gamma
        self.beta = beta
        self.t_end = t_end
        self.n_t = n_t
        self.initial_U = initial_U
        self.initial_Udot = initial_Udot
        self.initial_theta = initial_theta
        self.initial_thetadot = initial_thetadot

        self.vec_t = np.linspace(0., t_end, n_t)
        self.dt = t_end / (n_t - 1)                       # solvers.py:89

        self.rtol, self.max_iter, self.check_every = rtol, max_iter, check_every
        self.history, self.warm_start, self.on_step, self.progress = history, warm_start, on_step, progress
        self.precond = precond
        self.precond_options = precond_options      # keyword arguments of multilevel.TwoLevelPreconditioner (experiments)
        self.floor_factor = floor_factor
        # comm: None (one GPU) | 'auto' | distributed.Comm - row-partitioned solve on the GPUs of the torch.distributed
        # job (dist_api.py); every rank receives the kept histories
        self.comm = comm
        # initial guess of each solve: 0 (warm_start=False) or the previous accelerations extrapolated in time
        # with a polynomial of degree guess_order (0 = previous acceleration, 1 = linear, 2 = quadratic).
        # Measured on the B200 (sandwich mesh, 100^3 and 190^3 boxes): the iteration count of BiCGSTAB, with and without
        # the coarse correction,
        # is insensitive to the initial guess (420 +- 30 iterations for orders 0, 1, 2 alike - the Krylov space
        # has to be rebuilt to resolve the low modes whatever the starting residual), so the default stays 0.
        self.guess_order = guess_order

        self.X = None
        self.Xdot = None
        self.Xdotdot = None
        self.q = None
        self.qdot = None
        self.qdotdot = None
        self.U = None
        self.Udot = None
        self.Udotdot = None
        self.T = None
        self.Tdot = None
        self.Tdotdot = None
        self.solve_info = []
        self.timings = {}
        self.state_device = None

    # -- pieces reused by bench.py ----------------------------------------------------------------
    def prepare(self):
        """Everything before the time loop (solvers.py:105-161)."""
        m = self.model
        m.create_free_dofs_lists()
        m.assemble_MDK()                                  # assemble_M, assemble_K, assemble_D in one launch
        m.apply_dirichlet_matrices()
        dev = m._dir_mask.device
        n_nodes, n = m.mesh.n_nodes, m.mesh.n_dofs

        def interleave(u3, th):
            X = np.zeros(n)
            X[::4] = np.asarray(u3)[::3]
            X[1::4] = np.asarray(u3)[1::3]
            X[2::4] = np.asarray(u3)[2::3]
            X[3::4] = th
            return X

        x0 = interleave(self.initial_U, self.initial_theta)
        v0 = interleave(self.initial_Udot, self.initial_thetadot)
        free_mask = ~m._mask_host
        # the reference advances only the free DOFs (solvers.py:155-157); constrained DOFs stay at their
        # prescribed value in X and at 0 in Xdot / Xdotdot
        self._x = torch.from_numpy(x0 * free_mask).to(dev)
        self._v = torch.from_numpy(v0 * free_mask).to(dev)
        self._a = torch.zeros(n, dtype=torch.float64, device=dev)     # initial acceleration assumed zero (:131)
        self._x0_full, self._v0_full = x0, v0
        dt = self.dt
        self._ss = _SystemSolve(m, 1.0, self.gamma * dt, self.beta * dt ** 2, self.rtol, self.max_iter, self.check_every,
                                self.precond, floor_factor=self.floor_factor, precond_options=self.precond_options)
        self._w1 = torch.empty_like(self._x)
        self._w2 = torch.empty_like(self._x)
        self._rhs = torch.empty_like(self._x)
        self._a_new = torch.zeros_like(self._x)
        self._guess = GuessExtrapolator(_lib.load(), self._x, order=self.guess_order)
        self._lift = m.lift_vector() if m._has_nonzero_dirichlet else None
        self._mapD = layout.plane_map_D(*m._d_alpha)
        self._mapK = layout.plane_map_K()
        self._lib = _lib.load()

    def n_load_terms(self):
        return len(self.model.load_terms(idx_t=1))

    def load_coefficients(self, i):
        """[n_terms, 4] amplitudes of the load terms at step i (host): what an end-to-end caller copies to the device."""
        return np.array([coef for _, _, coef in self.model.load_terms(idx_t=i)], dtype=np.float64).reshape(-1, 4)

    def step(self, i, coef_dev=None):
        """One time step (solvers.py:164-186) entirely on the device.  coef_dev (optional, device [n_terms, 4]): the load
        amplitudes of this step already copied to the device (end-to-end path); default: taken from the model's host arrays
        and passed as kernel arguments."""
        m, dm = self.model, self._ss.dm
        rhs = self._rhs
        rhs.zero_()
        for k, (node, w, coef) in enumerate(m.load_terms(idx_t=i)):        # F(i), model.py:120-286
            load_axpy(self._lib, node, w, coef if coef_dev is None else coef_dev[k], rhs)
        if self._lift is not None:                                         # F_f -= K_fd x_d, model.py:326-339
            rhs -= self._lift
        newmark_predict(self._lib, self._x, self._v, self._a, self.dt, self.gamma, self.beta, m._dir_mask,
                        self._w1, self._w2)
        dm.planes_spmv(m._planes["D"], self._mapD, self._w1, rhs, alpha=-1.0, beta_y=1.0)   # - D_ff @ (...)
        dm.planes_spmv(m._planes["K"], self._mapK, self._w2, rhs, alpha=-1.0, beta_y=1.0)   # - K_ff @ (...)
        if not self.warm_start:
            self._a_new.zero_()
        else:
            self._guess.guess(self._a, self._a_new)                        # extrapolated initial guess
        info = self._ss.solve(rhs, self._a_new)                            # spsolve(K_dyn, rhs), solvers.py:172
        self._guess.push(self._a)
        newmark_correct(self._lib, self._x, self._v, self._a, self._a_new, self.dt, self.gamma, self.beta, m._dir_mask)
        return info

    def _solve_distributed(self):
        from . import dist_api
        m = self.model
        t0 = time.perf_counter()
        comm = dist_api.make_comm(self.comm)
        m.create_free_dofs_lists()
        n, n_t, n_nodes = m.mesh.n_dofs, self.n_t, m.mesh.n_nodes

        def interleave(u3, th):
            X = np.zeros(n)
            X[::4], X[1::4], X[2::4], X[3::4] = np.asarray(u3)[::3], np.asarray(u3)[1::3], np.asarray(u3)[2::3], th
            return X

        x0, v0 = interleave(self.initial_U, self.initial_theta), interleave(self.initial_Udot, self.initial_thetadot)
        pm, ds = dist_api.distributed_transient(m, x0, v0, self.t_end, n_t, self.gamma, self.beta, comm, self.rtol,
                                                self.max_iter, self.check_every, self.precond, self.floor_factor,
                                                MULTILEVEL_MIN_NODES, self.precond_options)
        ds.prepare()
        keep_every = 1 if self.history == 'full' else (None if self.history == 'last' else int(self.history))
        cols = list(range(0, n_t, keep_every)) if keep_every is not None else [n_t - 1]
        col_of = {s_: k for k, s_ in enumerate(cols)}
        self.kept_steps = np.array(cols)
        self.X, self.Xdot, self.Xdotdot = (np.zeros((n, len(cols))) for _ in range(3))
        if 0 in col_of:
            self.X[:, 0], self.Xdot[:, 0] = x0, v0
            self.X[m._mask_host, 0] = m._g_fill_host[m._mask_host]
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        it = range(1, n_t)
        for i in (_tqdm(it) if (self.progress and comm.rank == 0) else it):
            self.solve_info.append(ds.step(i))
            if i in col_of:
                k = col_of[i]
                xo, vo, ao = ds.owned_state()
                self.X[:, k], self.Xdot[:, k], self.Xdotdot[:, k] = pm.gather(xo), pm.gather(vo), pm.gather(ao)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        self.U, theta = _deinterleave(self.X, n_nodes)
        self.Udot, self.Tdot = _deinterleave(self.Xdot, n_nodes)
        self.Udotdot, self.Tdotdot = _deinterleave(self.Xdotdot, n_nodes)
        self.T = theta + m.mesh.reference_temperature()[:, None]
        self.timings = dict(prepare_s=t1 - t0, loop_s=t2 - t1, steps=n_t - 1, post_s=time.perf_counter() - t2)
        self._distributed = (pm, ds)

    def solve(self):
        if self.comm is not None:
            return self._solve_distributed()
        m = self.model
        t0 = time.perf_counter()
        self.prepare()
        n, n_t, n_nodes = m.mesh.n_dofs, self.n_t, m.mesh.n_nodes
        keep_every = 1 if self.history == 'full' else (None if self.history == 'last' else int(self.history))
        if keep_every is not None:
            cols = list(range(0, n_t, keep_every))
            col_of = {s: k for k, s in enumerate(cols)}
            self.kept_steps = np.array(cols)
            self.X = np.zeros((n, len(cols)))
            self.Xdot = np.zeros((n, len(cols)))
            self.Xdotdot = np.zeros((n, len(cols)))
            self.X[:, 0] = self._x0_full
            self.Xdot[:, 0] = self._v0_full
            fill = m._g_fill_host
            self.X[m._mask_host, :] = fill[m._mask_host][:, None]          # solvers.py:133-153
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        it = range(1, n_t)
        for i in (_tqdm(it) if self.progress else it):
            self.solve_info.append(self.step(i))
            if self.on_step is not None:
                self.on_step(i, self._x, self._v, self._a)
            if keep_every is not None and i in col_of:
                k = col_of[i]
                free = m.free_dofs
                self.X[free, k] = self._x.cpu().numpy()[free]
                self.Xdot[free, k] = self._v.cpu().numpy()[free]
                self.Xdotdot[free, k] = self._a.cpu().numpy()[free]
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        self.state_device = (self._x + m._g_fill, self._v, self._a)
        if keep_every is None:
            X = self.state_device[0].cpu().numpy()[:, None]
            self.X, self.Xdot, self.Xdotdot = X, self._v.cpu().numpy()[:, None], self._a.cpu().numpy()[:, None]
            self.kept_steps = np.array([n_t - 1])
        self.U, theta = _deinterleave(self.X, n_nodes)
        self.Udot, self.Tdot = _deinterleave(self.Xdot, n_nodes)
        self.Udotdot, self.Tdotdot = _deinterleave(self.Xdotdot, n_nodes)
        self.T = theta + m.mesh.reference_temperature()[:, None]          # solvers.py:204-215
        self.timings = dict(prepare_s=t1 - t0, loop_s=t2 - t1, steps=n_t - 1, post_s=time.perf_counter() - t2)

    def solve_ROM(self, n_modes_u, n_modes_theta, history='full', eig_tol=1e-9):
        """Newmark integration of the Galerkin-reduced model (reference solvers.py:219-354).

        The bases (model.compute_ROB_u / compute_ROB_theta), the reduced matrices and the reduced load vectors are
        computed on the GPU; the (n_q x n_q) time loop runs on the host exactly like the reference's
        (solvers.py:273-297) and the histories are expanded with one GEMM per quantity.  history: 'full' like the
        reference (arrays of shape (., n_t)), 'last', or an integer stride."""
        m = self.model
        t0 = time.perf_counter()
        m.create_free_dofs_lists()
        m.assemble_MDK()                                   # assemble_M, assemble_K, assemble_D (solvers.py:221-223)
        m.compute_ROB_u(n_modes_u, tol=eig_tol)
        m.compute_ROB_theta(n_modes_theta, tol=eig_tol)
        m.compute_ROM_matrices()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        nu, nt, n_t, n = m.n_q_u, m.n_q_t, self.n_t, m.mesh.n_dofs
        dev = m._phi_u_dev.device
        Phi = torch.cat([m._phi_u_dev, m._phi_t_dev])       # [nu + nt, 4N], rows vanish outside their field's free DOFs

        def interleave_rom(u3, th):                         # solvers.py:237-241 / 246-250: y and z are initialised
            X = np.zeros(n)                                 # from the X components, as in the reference
            X[::4] = np.asarray(u3)[::3]
            X[1::4] = np.asarray(u3)[::3]
            X[2::4] = np.asarray(u3)[::3]
            X[3::4] = th
            return torch.from_numpy(X).to(dev)

        q_init = (Phi @ interleave_rom(self.initial_U, self.initial_theta)).cpu().numpy()          # solvers.py:243-244
        v_init = (Phi @ interleave_rom(self.initial_Udot, self.initial_thetadot)).cpu().numpy()    # solvers.py:252-253
        it = range(1, n_t)
        self.q, self.qdot, self.qdotdot = reduced_newmark(m.mat_Mrom, m.mat_Drom, m.mat_Krom, m.rom_forces, q_init, v_init,
                                                          self.dt, self.gamma, self.beta, n_t,
                                                          steps=_tqdm(it) if self.progress else it)
        t2 = time.perf_counter()
        # expansion to the physical DOFs (solvers.py:299-326) and prescribed values (:328-341)
        cols = (np.arange(n_t) if history == 'full' else np.array([n_t - 1]) if history == 'last'
                else np.arange(0, n_t, int(history)))
        self.kept_steps = cols
        PhiT = Phi.T.contiguous()

        def expand(q, with_dirichlet):
            X = (PhiT @ torch.from_numpy(np.ascontiguousarray(q[:, cols])).to(dev)).cpu().numpy()
            if with_dirichlet:
                X[m._mask_host, :] = m._g_fill_host[m._mask_host][:, None]
            return X

        self.X, self.Xdot, self.Xdotdot = expand(self.q, True), expand(self.qdot, False), expand(self.qdotdot, False)
        n_nodes = m.mesh.n_nodes
        self.U, theta = _deinterleave(self.X, n_nodes)
        self.Udot, self.Tdot = _deinterleave(self.Xdot, n_nodes)
        self.Udotdot, self.Tdotdot = _deinterleave(self.Xdotdot, n_nodes)
        self.T = theta + m.mesh.reference_temperature()[:, None]             # solvers.py:343-354
        self.timings = dict(prepare_s=t1 - t0, loop_s=t2 - t1, steps=n_t - 1, post_s=time.perf_counter() - t2)

    def solve_ROM_nonparametric(self, *args, **kwargs):
        raise NotImplementedError("non-parametric UQ solve (reference solvers.py:356-668) is outside the B200 hot path")
