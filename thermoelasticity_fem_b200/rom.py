"""Reduced-order bases and reduced operators (reference ``model.py:341-440``).

The reference computes, per field, the ``n_q`` eigenvectors of smallest eigenvalue magnitude of
``K_ff x = lambda B_ff x`` with ``scipy.sparse.linalg.eigs(K, k, M=B, which='SM')`` (``model.py:349,361``:
field u: K_uu / M_uu on ``free_dofs_U``; field theta: K_tt / D_tt on ``free_dofs_theta``) and projects M, D, K
and the load vector on them.  Here the operators stay on the GPU in plane storage and are applied by the
hand-written plane SpMV (``tefem_planes_spmv``); the eigenpairs come from a BLOCK SHIFT-INVERT LANCZOS iteration
with full re-orthogonalisation whose linear solves ``(K + sigma B) y = B x`` run on the Krylov solver of the hot
path (K3-K6) - ARPACK's role in the reference.  Vectors live in the uniform 4-DOF-per-node layout and vanish
outside the field's free DOFs, so that no gather / scatter between "reduced" and "full" numberings is needed.

``block_lanczos_smallest`` is written against three callables on torch tensors and is exercised on the CPU by the
``-m "not gpu"`` tests with scipy-backed operators; the product path (``FieldEigenProblem``) requires CUDA.
"""
import numpy as np
import torch


def _b_orthonormalize(W, BW, drop_tol=1e-10):
    """Rows of W made B-orthonormal (BW = rows B w); returns (Q, BQ) with possibly fewer rows."""
    G = (W @ BW.T).double().cpu().numpy()
    G = 0.5 * (G + G.T)
    lam, U = np.linalg.eigh(G)
    keep = lam > drop_tol * max(lam.max(), 0.0)
    if not keep.any():
        return W[:0], BW[:0]
    T = torch.from_numpy((U[:, keep] / np.sqrt(lam[keep])).T.copy()).to(W.device, W.dtype)     # [kept, b]
    return T @ W, T @ BW


def block_lanczos_smallest(apply_B, solve_A, x0, k, tol=1e-9, max_vectors=None, shift=0.0, log=None, apply_A=None):
    """The k eigenpairs of smallest eigenvalue of A x = lambda B x (A, B symmetric, B positive definite on the
    subspace the vectors live in, A + shift B positive definite).

    apply_B(X) -> B X and solve_A(R, X_guess) -> (A + shift B)^-1 R act on the ROWS of [b, n] tensors.
    x0: [b, n] start block (b >= 1; b bounds the eigenvalue multiplicity that can be resolved).
    Convergence: max_j |S x_j - theta_j x_j|_B / theta_j < tol (S = (A + shift B)^-1 B, theta = 1 / (lambda + shift)).
    With apply_A given the converged Ritz vectors are POLISHED by one step of subspace iteration (X <- S X, then
    Rayleigh-Ritz with the pencil (A, B) itself): S damps the rough error components that the re-orthogonalisation
    accumulates and that the residual |A x - lambda B x| / |A x| (the measure ARPACK reports for the reference's eigs)
    amplifies by up to lambda_max / lambda; info["residual"] is that measure after the polish.
    Returns (lam[k] ascending, X[k, n] B-orthonormal rows, info)."""
    b, n = x0.shape
    max_vectors = max_vectors or 8 * (k + b)
    V = torch.zeros((max_vectors, n), dtype=x0.dtype, device=x0.device)
    BV = torch.zeros_like(V)
    SV = torch.zeros_like(V)
    Q, BQ = _b_orthonormalize(x0, apply_B(x0))
    nv = 0
    lam = X = None
    info = dict(solves=0, blocks=0, converged=False, residual_estimate=None)
    while Q.shape[0] > 0 and nv + Q.shape[0] <= max_vectors:
        nb = Q.shape[0]
        V[nv:nv + nb], BV[nv:nv + nb] = Q, BQ
        W = solve_A(BQ, Q)                                   # S Q with S = (A + shift B)^-1 B
        SV[nv:nv + nb] = W
        nv += nb
        info["solves"] += nb
        info["blocks"] += 1
        # Rayleigh-Ritz of S in the B inner product on span(V): H = V B S V^T (symmetric)
        H = (BV[:nv] @ SV[:nv].T).double().cpu().numpy()
        H = 0.5 * (H + H.T)
        theta, C = np.linalg.eigh(H)
        order = np.argsort(-theta)[:k]                       # largest theta = smallest lambda + shift
        if order.size == k and theta[order].min() > 0:
            Ck = torch.from_numpy(C[:, order].T.copy()).to(V.device, V.dtype)
            X = Ck @ V[:nv]
            SX = Ck @ SV[:nv]
            th = torch.from_numpy(theta[order].copy()).to(V.device, V.dtype)
            R = SX - th[:, None] * X                         # S x - theta x; relative to theta it bounds the error of lambda + shift
            BR = apply_B(R)
            est = torch.sqrt(torch.clamp((R * BR).sum(dim=1), min=0.0)) / th
            lam = 1.0 / theta[order] - shift
            info["residual_estimate"] = float(est.max())
            if log is not None:
                log(f"block {info['blocks']}: {nv} vectors, residual estimate {info['residual_estimate']:.2e}")
            if info["residual_estimate"] < tol:
                info["converged"] = True
                break
        # next block: S Q orthogonalised against everything so far and B-orthonormalised, twice (the whitening of a
        # nearly dependent block amplifies what the first projection left behind)
        Q, BQ = W, apply_B(W)
        for _ in range(2):
            coef = Q @ BV[:nv].T
            Q = Q - coef @ V[:nv]
            BQ = BQ - coef @ BV[:nv]
            Q, BQ = _b_orthonormalize(Q, BQ)
            if Q.shape[0] == 0:
                break
            BQ = apply_B(Q)            # not the recurrence: its rounding errors are amplified by the whitening
    if X is None:
        raise RuntimeError("block Lanczos: the Krylov space is too small for the requested number of eigenpairs")
    if apply_A is not None:
        def true_residual(X, lam):
            AX, BX = apply_A(X), apply_B(X)
            lam_t = torch.from_numpy(np.asarray(lam)).to(X.device, X.dtype)
            return float(((AX - lam_t[:, None] * BX).norm(dim=1) / AX.norm(dim=1)).max())
        info["residual_before_polish"] = true_residual(X, lam)
        Y = solve_A(apply_B(X), X)
        info["solves"] += k
        GA = (Y @ apply_A(Y).T).double().cpu().numpy()
        GB = (Y @ apply_B(Y).T).double().cpu().numpy()
        import scipy.linalg
        lam_p, C = scipy.linalg.eigh(0.5 * (GA + GA.T), 0.5 * (GB + GB.T))
        Xp = torch.from_numpy(C.T.copy()).to(Y.device, Y.dtype) @ Y
        res_p = true_residual(Xp, lam_p)
        if res_p < info["residual_before_polish"]:
            lam, X = lam_p, Xp
            info["residual"] = res_p
        else:
            info["residual"] = info["residual_before_polish"]
        if log is not None:
            log(f"polish: residual {info['residual_before_polish']:.2e} -> {res_p:.2e}")
    return lam, X, info


class FieldEigenProblem:
    """K x = lambda B x restricted to the free DOFs of one field ('u': K_uu / M_uu, model.py:346-349; 'theta': K_tt / D_tt,
    model.py:358-361), operators in plane storage on the GPU, vectors in the uniform 4-DOF-per-node layout."""

    def __init__(self, model, field, rtol=1e-11, max_iter=50000, check_every=32, precond="auto", shift=None):
        from . import layout
        from .solvers import _SystemSolve
        assert field in ("u", "theta")
        self.model, self.field = model, field
        self.dm = model.mesh.device_mesh()
        dev = self.dm.device
        n = model.mesh.n_dofs
        comp = np.arange(n) % 4
        in_field = (comp < 3) if field == "u" else (comp == 3)
        free = in_field & ~model._mask_host
        self.free_host = free
        self.free = torch.from_numpy(free).to(dev)
        self.freef = self.free.to(torch.float64)
        self.not_free_u8 = torch.from_numpy((~free).astype(np.uint8)).to(dev)
        full_K = layout.plane_map_K()
        mapA = np.full((4, 4), -1, dtype=np.int8)
        mapB = np.full((4, 4), -1, dtype=np.int8)
        if field == "u":
            need = ("K", "M")
            mapA[:3, :3] = full_K[:3, :3]
            mapB[:3, :3] = layout.plane_map_M()[:3, :3]
            self.planesB = model._planes.get("M")
        else:
            need = ("K", "D")
            mapA[3, 3] = full_K[3, 3]
            mapB[3, 3] = model.plane_map("D")[3, 3]
            self.planesB = model._planes.get("D")
        for nm in need:
            if nm not in model._planes:
                raise RuntimeError(f"assemble_{nm}() must be called before computing the reduced-order basis")   # model.py:342,354
        self.planesA = model._planes["K"]
        self.mapA, self.mapB = mapA, mapB
        # a field without Dirichlet DOFs has a singular K (rigid-body / constant modes): shift the solves
        has_dirichlet = bool((in_field & model._mask_host).any())
        if shift is None:
            shift = 0.0
            if not has_dirichlet:
                g = torch.Generator(device="cpu").manual_seed(1)
                x = (torch.rand(n, dtype=torch.float64, generator=g).to(dev) - 0.5) * self.freef
                shift = 1e-6 * float((x * self.apply_A(x[None])[0]).sum() / (x * self.apply_B(x[None])[0]).sum())
        self.shift = float(shift)
        coef = dict(cM=self.shift if field == "u" else 0.0, cD=self.shift if field == "theta" else 0.0, cK=1.0)
        self._ss = _SystemSolve(model, coef["cM"], coef["cD"], coef["cK"], rtol, max_iter, check_every, precond,
                                dir_mask=self.not_free_u8)

    def _apply(self, planes, pmap, X):
        Y = torch.empty_like(X)
        for j in range(X.shape[0]):
            self.dm.planes_spmv(planes, pmap, X[j], Y[j])
        return Y * self.freef

    def apply_A(self, X):
        return self._apply(self.planesA, self.mapA, X)

    def apply_B(self, X):
        return self._apply(self.planesB, self.mapB, X)

    def solve_A(self, R, X_guess=None):
        Y = torch.zeros_like(R)
        for j in range(R.shape[0]):
            self._ss.solve(R[j].contiguous(), Y[j])
        return Y * self.freef

    def smallest(self, k, block=None, tol=1e-9, seed=0, log=None):
        """(lam[k], Phi[k, n_dofs] device tensor with unit Euclidean norm rows, info)."""
        n = self.model.mesh.n_dofs
        n_free = int(self.free_host.sum())
        if not 0 < k < n_free:
            raise ValueError("number of modes must be between 1 and the number of free DOFs of the field - 1")
        block = block or max(4, min(8, k))
        g = torch.Generator(device="cpu").manual_seed(seed)
        x0 = (torch.rand((block, n), dtype=torch.float64, generator=g) - 0.5).to(self.dm.device) * self.freef
        lam, X, info = block_lanczos_smallest(self.apply_B, self.solve_A, x0, k, tol=tol, shift=self.shift, log=log,
                                              max_vectors=min(n_free, 8 * (k + block)), apply_A=self.apply_A)
        # info["residual"] = max |K x - lam B x| / |K x| (what the golden fixture records for the reference's eigs)
        X = X / X.norm(dim=1, keepdim=True)                  # eigs returns unit 2-norm vectors
        return np.asarray(lam), X, info


def project(dm, planes, pmap, Phi_right, Phi_lefts):
    """[Phi_left @ (Op @ Phi_right^T) for Phi_left in Phi_lefts] with Op in plane storage: one plane SpMV per column."""
    W = torch.empty_like(Phi_right)
    for j in range(Phi_right.shape[0]):
        dm.planes_spmv(planes, pmap, Phi_right[j], W[j])
    return [(P @ W.T).cpu().numpy() for P in Phi_lefts]
