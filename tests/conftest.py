import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        skip = pytest.mark.skip(reason="no CUDA device")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


def load_sandwich_blocks():
    g = np.load(os.path.join(GOLDEN, "sandwich_mesh.npz"))
    blocks = [(str(g[f"b{i}_type"]), int(g[f"b{i}_tag"]), g[f"b{i}_conn"].astype(np.int64)) for i in range(int(g["n_blocks"]))]
    return g["points"], blocks


def groups_from_blocks(blocks):
    tets, tris, nodes = {}, {}, {}
    for typ, tag, conn in blocks:
        d = {"tetra": tets, "triangle": tris, "vertex": nodes}[typ]
        c = conn.reshape(-1) if typ == "vertex" else conn
        d[tag] = c if tag not in d else np.concatenate([d[tag], c])
    return dict(nodes=nodes, tri=tris, tet=tets)


@pytest.fixture(scope="session")
def sandwich():
    pts, blocks = load_sandwich_blocks()
    return pts, blocks, groups_from_blocks(blocks)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def csc_from(f, prefix):
    import scipy.sparse as sp
    return sp.csc_array((f[prefix + "_data"], f[prefix + "_indices"], f[prefix + "_indptr"]), shape=tuple(f[prefix + "_shape"]))
