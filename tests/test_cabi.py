"""The C-ABI library loads on a machine without GPU and exports every symbol that
include/pnmol_b200.h declares; argument errors are reported through return codes."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__

    __graft_entry__.ensure_built()


def _declared():
    text = open(os.path.join(ROOT, "include", "pnmol_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pnmol_b200_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from pnmol_b200 import _lib

    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_lib.SIGNATURES) == names  # the ctypes table mirrors the header one to one
    assert lib.pnmol_b200_version() >= 100
    assert lib.pnmol_b200_launch_count() >= 0


def test_argument_errors_use_return_codes():
    from pnmol_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.pnmol_b200_create(ctypes.byref(h), 7, 4, 2, 2, 1, 1, 0, 0)
    assert rc < 0 and b"kind" in lib.pnmol_b200_last_error()
    rc = lib.pnmol_b200_create(ctypes.byref(h), 1, 4, 2, 2, 1, 1, 0, 0)  # semi-linear without reaction id
    assert rc < 0
    rc = lib.pnmol_b200_sqrt_propagate(None, None, None, 2, 2, 0, 1, 0, None)
    assert rc < 0
    with pytest.raises(_lib.PnmolB200Error):
        _lib.check(rc)


def test_structure_entry_point_runs_without_gpu():
    from pnmol_b200 import _lib

    lib = _lib.load()
    d, nu, nb = 5, 2, 2
    lcol = np.array([[0, 1, 2], [0, 1, 2], [1, 2, 3], [2, 3, 4], [2, 3, 4]], np.int32)
    bcol = np.array([[0], [4]], np.int32)
    D, m = 3 * d, d + nb
    te_p, be_p = np.zeros(D, np.int32), np.zeros(D, np.int32)
    te_u, be_u = np.zeros(m + D, np.int32), np.zeros(m + D, np.int32)
    rc = lib.pnmol_b200_structure(0, d, nu, nb, _lib.ptr(lcol), 3, _lib.ptr(bcol), 1, 1, 0, _lib.ptr(te_p), _lib.ptr(be_p),
                                  _lib.ptr(te_u), _lib.ptr(be_u))
    assert rc == 0
    assert te_p.tolist() == [2, 2, 2, 5, 5, 5, 8, 8, 8, 11, 11, 11, 14, 14, 14]
    assert be_p.tolist() == list(range(D, 2 * D))
    assert te_u[0] == 6 and te_u[m - 1] == 13 and te_u[-1] == D - 1 and be_u[0] == D and be_u[-1] == D + m - 1
