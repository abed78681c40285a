"""Parity metrics shared by the tests (SURVEY.md section 8a row A9 and hard part H1)."""
import numpy as np
import scipy.sparse as sp

VALUE_TOL = 1e-12       # BASELINE.json north_star: assembled entries within 1e-12 relative
SOLUTION_TOL = 1e-8     # BASELINE.json north_star: solutions within 1e-8 relative (residual-checked oracle)


def keys(A):
    A = A.tocoo()
    return A.row.astype(np.int64) * A.shape[1] + A.col


def block_class(row, col):
    return (row % 4 == 3) * 2 + (col % 4 == 3)      # uu = 0, ut = 1, tu = 2, tt = 3


def row_class_scale(A):
    """scale[node_row, class] = max |entry| of that block class in that node row (always contains the
    diagonal block, so it is the natural magnitude of the row)."""
    A = A.tocoo()
    n_nodes = A.shape[0] // 4
    scale = np.zeros((n_nodes, 4))
    np.maximum.at(scale, (A.row // 4, block_class(A.row, A.col)), np.abs(A.data))
    return scale


def scaled_error(A_new, A_ref):
    """max over entries of |a - b| / max(|b|, s) with s the row/class scale of the reference-side matrix.
    Entries present in only one of the two count with the other side = 0."""
    D = (A_new - A_ref).tocoo()
    if D.nnz == 0:
        return 0.0
    scale = np.maximum(row_class_scale(A_ref), row_class_scale(A_new))
    R = sp.csr_array(A_ref)
    ref_vals = np.abs(np.asarray(R[D.row, D.col]).reshape(-1))
    s = np.maximum(ref_vals, scale[D.row // 4, block_class(D.row, D.col)])
    return float(np.max(np.abs(D.data) / s))


def check_pattern(A_touched, A_ref_stored, A_oracle_touched=None):
    """(i) touched pattern == oracle touched pattern, bit-exact, sorted; (ii) reference.nonzero() is a subset."""
    A_touched = A_touched.tocsc()
    A_touched.sort_indices()
    if A_oracle_touched is not None:
        B = A_oracle_touched.tocsc()
        B.sort_indices()
        assert np.array_equal(A_touched.indptr, B.indptr), "touched pattern: indptr differs"
        assert np.array_equal(A_touched.indices, B.indices), "touched pattern: indices differ"
    if A_ref_stored is not None:
        assert np.isin(keys(A_ref_stored), keys(A_touched)).all(), "reference stores an entry outside the touched pattern"


def field_errors(x, x_ref):
    idx = np.arange(x.shape[0])
    out = []
    for m in (idx % 4 != 3, idx % 4 == 3):
        d = np.linalg.norm(x_ref[m])
        out.append(np.linalg.norm(x[m] - x_ref[m]) / d if d > 0 else np.linalg.norm(x[m]))
    return out
