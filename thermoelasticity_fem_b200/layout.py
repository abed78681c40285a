"""Plane layouts of the assembled operators (see include/tefem.h "Storage formats").

An operator's values live in structure-of-arrays planes ``vals[plane, p]`` over the shared
node-block pattern; ``plane_map[r, c]`` gives the plane of scalar entry (4*i + r, 4*j + c) of block
(i, j), or -1 when the reference never adds into that entry (``model.py:91,98-102,109-117``):

* M: only uu, and Muu_e is component-diagonal (``elements.py:226-227``) -> 3 planes
* K: uu, ut, tt; no tu block (``model.py:98-102``)                      -> 13 planes
* D: tu, tt (``model.py:109-111``), plus uu when Rayleigh damping is on (``model.py:112-117``):
  component-diagonal for alpha_M only, full 3x3 when alpha_K != 0        -> 4 / 7 / 13 planes
"""
import numpy as np


def plane_map_K():
    m = -np.ones((4, 4), dtype=np.int8)
    for r in range(3):
        for c in range(3):
            m[r, c] = 3 * r + c
        m[r, 3] = 9 + r
    m[3, 3] = 12
    return m


def plane_map_M():
    m = -np.ones((4, 4), dtype=np.int8)
    for r in range(3):
        m[r, r] = r
    return m


def d_uu_mode(alpha_M, alpha_K):
    """0: no uu part, 1: alpha_M only (component diagonal), 2: alpha_K != 0 (full 3x3)."""
    return 2 if alpha_K != 0. else (1 if alpha_M != 0. else 0)


def plane_map_D(alpha_M, alpha_K):
    m = -np.ones((4, 4), dtype=np.int8)
    for c in range(3):
        m[3, c] = c
    m[3, 3] = 3
    mode = d_uu_mode(alpha_M, alpha_K)
    if mode == 1:
        for r in range(3):
            m[r, r] = 4 + r
    elif mode == 2:
        for r in range(3):
            for c in range(3):
                m[r, c] = 4 + 3 * r + c
    return m


def n_planes(plane_map):
    return int(plane_map.max()) + 1
